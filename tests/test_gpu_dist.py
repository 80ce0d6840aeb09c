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


def _worker(rank, world, port, k, q, peer=True):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from kmer_denovo_filter_b200 import engine, synth
        from kmer_denovo_filter_b200.discovery import kmer_chain, kmer_chain_dist
        eng = engine.CudaEngine(dev)
        eng.peer_bins = peer     # True: NVLink peer-memory route; False: NCCL all-to-all route
        assert kmer_chain_dist.peer_memory_available(eng) == peer
        trio = synth.make_trio(torch, dev, GENOME, depth=DEPTH, n_denovo=20, rank=rank, world=world)
        res = kmer_chain_dist.discover_streams_dist(
            eng, _dev(engine, trio["child"]), _dev(engine, trio["mother"]),
            _dev(engine, trio["father"]), _dev(engine, trio["ref"]), k)
        sizes = {x: res[x] for x in ("candidates", "non_ref", "after_mother", "proband_unique",
                                     "child_distinct", "informative_reads")}
        pu = sorted(res["pu"].to_pyints())
        if rank == 0:
            shards = [trio] + [synth.make_trio(torch, dev, GENOME, depth=DEPTH, n_denovo=20,
                                               rank=r, world=world) for r in range(1, world)]
            u = {}
            for w in ("child", "mother", "father"):
                acc = shards[0][w]
                for s in shards[1:]:
                    acc = _union(torch, acc, s[w])
                u[w] = acc
            full = synth.make_trio(torch, dev, GENOME, depth=0.2, n_denovo=20)
            one = kmer_chain.discover_streams(eng, _dev(engine, u["child"]), _dev(engine, u["mother"]),
                                              _dev(engine, u["father"]), _dev(engine, full["ref"]), k)
            want = {x: one[x] for x in sizes}
            q.put((rank, sizes, pu, want, sorted(one["pu"].to_pyints())))
        else:
            q.put((rank, sizes, pu, None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("peer", [True, False])
@pytest.mark.parametrize("k", [31, 47])
def test_dist_chain_equals_single_gpu(k, peer):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, q, peer)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = next(g for g in got if g[3] is not None)
    assert want[3]["proband_unique"] > 0
    for rank, sizes, pu, _w, _p in got:
        assert sizes == want[3], (rank, sizes, want[3])
        assert pu == want[4]
