"""The multi-GPU chain's host logic under gloo, world_size 2, on CPU: owner
routing + all-to-all, owner-local count and reference subtraction, replicated
parent filtering with all-reduced counts — against a single-process oracle
count of the union of both ranks' shards."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

K = 31
GENOME = 60_000
DEPTH = 12
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dev_stream(s):
    from kmer_denovo_filter_b200 import engine
    return engine.DeviceStream(s["codes"], s["valid"], s["n_bases"], s["read_starts"], s["read_lens"])


def _split_stream(ds):
    """One stream as two (split at a read boundary, 32-base aligned: whole words)."""
    from kmer_denovo_filter_b200 import engine
    starts = ds.read_starts.numpy().astype(np.int64)
    if starts.shape[0] < 2:
        return [ds]
    cand = np.flatnonzero(starts % 32 == 0)
    cand = cand[cand > 0]
    if cand.shape[0] == 0:
        return [ds]
    r = int(cand[cand.shape[0] // 2])
    cut = int(starts[r])
    w = cut // 32
    # the separator before read r is the last base of the first part
    a = engine.DeviceStream(ds.codes[:w].clone(), ds.valid[:w].clone(), cut - 1,
                            ds.read_starts[:r].clone(), ds.read_lens[:r].clone())
    b = engine.DeviceStream(ds.codes[w:].clone(), ds.valid[w:].clone(), ds.n_bases - cut,
                            (ds.read_starts[r:] - cut).clone(), ds.read_lens[r:].clone())
    return [a, b]


def _shards(rank):
    from kmer_denovo_filter_b200 import synth
    return synth.make_trio(torch, torch.device("cpu"), GENOME, depth=DEPTH, read_len=100,
                           n_denovo=12, rank=rank, world=WORLD)


def _worker(rank, port, q, n_passes=None, split=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from fake_engine import FakeEngine
        from kmer_denovo_filter_b200.discovery import kmer_chain_dist
        trio = _shards(rank)
        eng = FakeEngine()
        streams = {w: _dev_stream(trio[w]) for w in ("child", "mother", "father", "ref")}
        if split:     # every sample as a list of two streams (whole-genome samples exceed one)
            streams = {w: _split_stream(s) for w, s in streams.items()}
        res = kmer_chain_dist.discover_streams_dist(
            eng, streams["child"], streams["mother"], streams["father"], streams["ref"], K,
            n_passes=n_passes)
        pu = sorted(res["pu"].lo.numpy().view(np.uint64).tolist()) if res["pu"] is not None else []
        q.put((rank, {x: res[x] for x in ("candidates", "non_ref", "after_mother", "proband_unique",
                                          "child_distinct", "informative_reads", "units")}, pu,
               res["informative_reads_local"]))
    finally:
        dist.destroy_process_group()


def _expected():
    """Single-process oracle over the union of the shards."""
    import collections
    from fake_engine import stream_keys
    from kmer_denovo_filter_b200 import synth
    tot = {w: collections.Counter() for w in ("child", "mother", "father")}
    units = 0
    child_streams = []
    for r in range(WORLD):
        trio = _shards(r)
        for w in tot:
            ds = _dev_stream(trio[w])
            lo, hi, ok = stream_keys(ds, K)
            units += int(ok.sum()) * (2 if w == "child" else 1)   # child: binned + scanned
            tot[w].update(lo[ok].tolist())
            if w == "child":
                child_streams.append(ds)
        lo, hi, ok = stream_keys(_dev_stream(trio["ref"]), K)
        units += int(ok.sum())
    full = synth.make_trio(torch, torch.device("cpu"), GENOME, depth=0.1, read_len=100, n_denovo=12)
    rlo, _rhi, rok = stream_keys(_dev_stream(full["ref"]), K)
    ref = set(rlo[rok].tolist())
    cand = {x for x, c in tot["child"].items() if c >= 3}
    nonref = cand - ref
    am = {x for x in nonref if tot["mother"].get(x, 0) == 0}
    pu = {x for x in am if tot["father"].get(x, 0) == 0}
    inf = 0
    for ds in child_streams:
        lo, hi, ok = stream_keys(ds, K)
        starts = ds.read_starts.numpy().astype(np.int64)
        lens = ds.read_lens.numpy().astype(np.int64)
        hit = ok & np.isin(lo, np.fromiter(pu, dtype=np.uint64, count=len(pu)))
        for s, l in zip(starts.tolist(), lens.tolist()):
            seg = lo[s:s + max(0, l - K + 1)][hit[s:s + max(0, l - K + 1)]]
            if np.unique(seg).shape[0] >= K // 4:
                inf += 1
    return {"candidates": len(cand), "non_ref": len(nonref), "after_mother": len(am),
            "proband_unique": len(pu), "child_distinct": len(tot["child"]),
            "informative_reads": inf}, sorted(pu), units


import pytest


@pytest.mark.parametrize("n_passes,split", [(None, False), (2, True)])
def test_distributed_chain_world2_gloo(n_passes, split):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q, n_passes, split)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, want_pu, want_units = _expected()
    assert want["proband_unique"] > 0 and want["informative_reads"] > 0
    for rank, sizes, pu, _loc in got:
        for key, v in want.items():
            assert sizes[key] == v, (rank, key, sizes[key], v)
        assert pu == want_pu
    assert sum(g[3] for g in got) == want["informative_reads"]
    # duplicated reference windows in the 64-base shard overlap are the only slack
    assert 0 <= sum(g[1]["units"] for g in got) - want_units <= 64 * WORLD


def test_collective_helpers_single_process():
    """allgather_varlen / exchange_equal with world_size 1 (gloo)."""
    import torch.distributed as dist
    from kmer_denovo_filter_b200.discovery import kmer_chain_dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        t = torch.arange(7, dtype=torch.int64)
        assert torch.equal(D.allgather_varlen(torch, t, 1), t)
        assert torch.equal(D.exchange_equal(t, 1), t)
        assert torch.equal(D.allgather_varlen(torch, t[:0], 1), t[:0])
    finally:
        dist.destroy_process_group()


def _pipeline_worker(rank, port, argv, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from fake_engine import FakeEngine
        from kmer_denovo_filter_b200 import cli
        from kmer_denovo_filter_b200.discovery import pipeline as P
        metrics = P.run_discovery_pipeline(cli.parse_discovery_args(argv), engine=FakeEngine())
        q.put((rank, metrics))
    finally:
        dist.destroy_process_group()


def test_distributed_pipeline_world2_gloo_reproduces_goldens(giab_paths, tmp_path):
    """The multi-GPU product pipeline's host logic (rank-sharded BAM ranges through the .bai,
    reference slices, rank-ordered de-duplication of informative reads, gathers, rank-0
    writers) under gloo with the numpy engine: every output file of rank 0 equals the
    reference's golden file byte for byte."""
    import json
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    prefix = str(tmp_path / "giab_discovery")
    argv = ["--child", giab_paths["child"], "--mother", giab_paths["mother"],
            "--father", giab_paths["father"], "--ref-fasta", giab_paths["ref_fasta"],
            "--out-prefix", prefix, "--min-child-count", "3", "--kmer-size", "31", "--threads", "2",
            "--candidate-summary", os.path.join(giab_paths["expected_vcf"], "summary.txt")]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, port, argv, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=900) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gold_dir = giab_paths["expected_discovery"]
    gold = json.load(open(os.path.join(gold_dir, "giab_discovery.metrics.json")))
    assert got[0] == gold and got[1] is None
    for suffix in (".bed", ".kmer_coverage.bedgraph", ".read_coverage.bed", ".sv.bedpe",
                   ".summary.txt", ".metrics.json"):
        assert open(prefix + suffix).read() == \
            open(os.path.join(gold_dir, "giab_discovery" + suffix)).read(), suffix
    from kmer_denovo_filter_b200 import bamio
    with bamio.BamReader(prefix + ".informative.bam", threads=2) as rd:
        b = rd.next_batch(bamio.MODE_ALL, want_meta=3)
    assert b.n_reads > 150 and os.path.isfile(prefix + ".informative.bam.bai")
