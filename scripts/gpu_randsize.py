"""Random 32-byte sector reads against buffer size: where L2 ends and where the TLB reach ends."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from kmer_denovo_filter_b200 import engine
dev = torch.device("cuda", 0)
eng = engine.CudaEngine(dev)
for mb in (32, 64, 96, 128, 192, 256, 384, 512, 1024, 2048, 4096, 8192):
    buf = torch.zeros(mb * (1 << 20) // 8, dtype=torch.int64, device=dev)
    n_ops = 1 << 28
    out = []
    for mode in (0, 10, 1):
        eng.bench_random_access(buf, n_ops, mode)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.bench_random_access(buf, n_ops, mode); e1.record()
        torch.cuda.synchronize()
        out.append(n_ops / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    print("%5d MB  read32 %6.1f  read256 %6.1f  read+RED.ADD.32 %6.1f  G ops/s" % (mb, *out), flush=True)
    del buf
