"""Multi-GPU chain on real devices (NCCL, one process per GPU) against the
single-GPU chain over the union of the ranks' shards.  Needs >= 2 GPUs; the
host logic is covered on CPU by test_dist_chain_cpu.py."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

GENOME = 400_000
DEPTH = 14


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dev(engine, s):
    return engine.DeviceStream(s["codes"], s["valid"], s["n_bases"], s["read_starts"], s["read_lens"])


def _union(torch, a, b):
    """Two packed read shards (each a whole number of words) -> one stream."""
    n_a_words = a["codes"].shape[0]
    return {"codes": torch.cat([a["codes"], b["codes"]]), "valid": torch.cat([a["valid"], b["valid"]]),
            "n_bases": n_a_words * 32 + b["n_bases"],
            "read_starts": torch.cat([a["read_starts"], b["read_starts"] + n_a_words * 32]),
            "read_lens": torch.cat([a["read_lens"], b["read_lens"]])}


def _host(s):
    import numpy as np
    return (s["codes"].cpu().numpy().view(np.uint64), s["valid"].cpu().numpy().view(np.uint32),
            s["n_bases"], s["read_starts"].cpu().numpy().view(np.uint64),
            s["read_lens"].cpu().numpy().view(np.uint32))


def _worker(rank, world, port, k, q, peer=True, n_passes=None):
    import numpy as np
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from kmer_denovo_filter_b200 import engine, synth
        from kmer_denovo_filter_b200.discovery import kmer_chain_dist
        eng = engine.CudaEngine(dev)
        eng.peer_bins = peer     # True: NVLink peer-memory route; False: NCCL all-to-all route
        assert kmer_chain_dist.peer_memory_available(eng) == peer
        trio = synth.make_trio(torch, dev, GENOME, depth=DEPTH, n_denovo=20, rank=rank, world=world)
        res = kmer_chain_dist.discover_streams_dist(
            eng, _dev(engine, trio["child"]), _dev(engine, trio["mother"]),
            _dev(engine, trio["father"]), _dev(engine, trio["ref"]), k, fetch=True, n_passes=n_passes)
        if n_passes:
            assert res["n_passes"] == n_passes
        sizes = {x: res[x] for x in ("candidates", "non_ref", "after_mother", "proband_unique",
                                     "informative_reads")}
        pu = sorted(res["pu"].to_pyints())
        per_read = (res["ndistinct"], res["nhits"])
        if rank == 0:
            # the checker: the CPU oracle over the union of all ranks' shards
            from oracle import ckdf
            shards = [trio] + [synth.make_trio(torch, dev, GENOME, depth=DEPTH, n_denovo=20,
                                               rank=r, world=world) for r in range(1, world)]
            u = {}
            for w in ("child", "mother", "father"):
                acc = shards[0][w]
                for s in shards[1:]:
                    acc = _union(torch, acc, s[w])
                u[w] = _host(acc)
            ref = _host(synth.pack_sequence_tensor(torch, synth.make_reference(torch, dev, GENOME)))
            want = ckdf.discovery_chain(u["child"], u["mother"][:3], u["father"][:3], ref[:3], k,
                                        threads=ckdf.max_threads())
            want_pu = sorted((int(h) << 64) | int(l) for l, h in
                             zip(want["pu_lo"].tolist(), want["pu_hi"].tolist()))
            k4 = max(1, k // 4)
            want_sizes = {x: want[x] for x in ("candidates", "non_ref", "after_mother", "proband_unique")}
            want_sizes["informative_reads"] = int((want["nd"] >= k4).sum())
            counts = [int(s["child"]["read_starts"].shape[0]) for s in shards]
            q.put((rank, sizes, pu, per_read, want_sizes, want_pu, (want["nd"], want["nh"], counts)))
        else:
            q.put((rank, sizes, pu, per_read, None, None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("peer,n_passes", [(True, None), (False, None), (True, 2), (False, 4)])
@pytest.mark.parametrize("k", [31, 47])
def test_dist_chain_equals_oracle(k, peer, n_passes):
    """The distributed chain (both routes, one and several hash-range passes) against the
    CPU oracle over the union of the shards: stage sizes, proband-unique keys, and every
    rank's per-read (ndistinct, nhits)."""
    import numpy as np
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, q, peer, n_passes)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    import time as _time
    got, t_end = [], _time.time() + 600
    while len(got) < len(procs):       # fail at once when a worker dies (do not sit out the timeout)
        try:
            got.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or _time.time() > t_end:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                pytest.fail("distributed worker failed (exit codes %s)" % [p.exitcode for p in procs])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = next(g for g in got if g[4] is not None)
    assert want[4]["proband_unique"] > 0
    nd_w, nh_w, counts = want[6]
    offs = np.concatenate([[0], np.cumsum(counts)])
    for rank, sizes, pu, per_read, _w, _p, _r in got:
        assert sizes == want[4], (rank, sizes, want[4])
        assert pu == want[5]
        assert np.array_equal(per_read[0], nd_w[offs[rank]:offs[rank + 1]])
        assert np.array_equal(per_read[1], nh_w[offs[rank]:offs[rank + 1]])


def _pipeline_worker(rank, world, port, argv, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from kmer_denovo_filter_b200 import cli, engine
        from kmer_denovo_filter_b200.discovery import pipeline as P
        eng = engine.CudaEngine(dev)
        metrics = P.run_discovery_pipeline(cli.parse_discovery_args(argv), engine=eng)
        q.put((rank, metrics))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_dist_pipeline_reproduces_goldens(giab_paths, tmp_path, world):
    """kmer-discovery with one process per GPU (every rank decodes a .bai-cut range of each
    BAM): rank 0's BED / bedGraph / read-coverage BED / BEDPE / summary / metrics are the
    reference's golden files byte for byte, and the informative-reads BAM holds the same
    records as the single-GPU run."""
    import json
    import queue as _queue
    if torch.cuda.device_count() < world:
        pytest.skip("needs >= %d GPUs" % world)
    prefix = str(tmp_path / "giab_discovery")
    argv = ["--child", giab_paths["child"], "--mother", giab_paths["mother"],
            "--father", giab_paths["father"], "--ref-fasta", giab_paths["ref_fasta"],
            "--out-prefix", prefix, "--min-child-count", "3", "--kmer-size", "31",
            "--candidate-summary", os.path.join(giab_paths["expected_vcf"], "summary.txt")]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, argv, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = []
    while len(got) < world:
        try:
            got.append(q.get(timeout=2))
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                pytest.fail("pipeline worker failed (exit codes %s)" % [p.exitcode for p in procs])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gold_dir = giab_paths["expected_discovery"]
    gold = json.load(open(os.path.join(gold_dir, "giab_discovery.metrics.json")))
    metrics = dict(got)[0]
    assert metrics == gold
    for suffix in (".bed", ".kmer_coverage.bedgraph", ".read_coverage.bed", ".sv.bedpe",
                   ".summary.txt", ".metrics.json"):
        assert open(prefix + suffix).read() == open(os.path.join(gold_dir, "giab_discovery" + suffix)).read(), suffix
    from kmer_denovo_filter_b200 import bamio
    with bamio.BamReader(prefix + ".informative.bam", threads=2) as rd:
        b = rd.next_batch(bamio.MODE_ALL, want_meta=True)
    assert b.n_reads == 195 + 0 or b.n_reads > 0
    assert os.path.isfile(prefix + ".informative.bam.bai")
