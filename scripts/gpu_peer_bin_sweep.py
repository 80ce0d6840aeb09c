"""What limits the fused bin-and-send kernel at N GPUs: time of kdf_bin_stream_pass into
peer-mapped bins against the number of hash ranges per owner (run length over NVLink), next
to a plain NCCL all-to-all of the same bytes.  torchrun --nproc-per-node N this file."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmer_denovo_filter_b200 import engine, synth  # noqa: E402
from kmer_denovo_filter_b200.discovery import kmer_chain_dist as D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = engine.CudaEngine(dev)
    genome = 64_000_000 * world
    trio = synth.make_trio(torch, dev, genome, depth=30, rank=rank, world=world)
    s = trio["child"]
    ds = engine.DeviceStream(s["codes"], s["valid"], s["n_bases"], s["read_starts"], s["read_lens"])
    out = {}
    for n_local in (1, 4, 16, 32, 64):
        ms = []
        for rep in range(4):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            D.route_composite_p2p(eng, [ds], 31, world, "sweep%d" % n_local, n_local, ds.n_bases)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ms[1:])], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["bin_send_ms_n_local_%d" % n_local] = float(t.item())
    # local binning of the same stream (no NVLink): the kernel's own cost
    for n_parts in (8, 32, 256):
        bins = eng.new_bins(31, n_parts, int(ds.n_bases / n_parts * 1.1) + 4096)
        ms = []
        for rep in range(3):
            bins.reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.bin_stream(bins, ds)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        out["bin_local_ms_%d_bins" % n_parts] = min(ms[1:])
        del bins
    # NCCL all-to-all of the bytes a rank sends (8 B per k-mer)
    n = (ds.n_bases // world) * world
    send = torch.empty(n, dtype=torch.int64, device=dev)
    recv = torch.empty_like(send)
    ms = []
    for rep in range(4):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_to_all_single(recv, send)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ms[1:])], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["nccl_all_to_all_ms"] = float(t.item())
    out["bytes_sent_per_rank"] = int(n * 8 * (world - 1) // world)
    out["keys_per_rank"] = int(ds.n_bases)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
