"""Random-access read-modify-write flavours on B200 (L2-resident and HBM-sized buffers)."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from kmer_denovo_filter_b200 import engine
dev = torch.device("cuda", 0)
eng = engine.CudaEngine(dev)
names = {0: "read32", 10: "read256", 1: "read+RED.ADD.32", 2: "read+ATOM.ADD.32", 3: "read+CAS.32",
         4: "read+CAS.64", 5: "read+RED.ADD.64", 6: "read+RED.AND.64", 7: "read+ATOM.EXCH.64",
         8: "RED.ADD.32", 11: "read256+CAS.64 on 1/6"}
for mb in (48, 96, 8192):
    buf = torch.zeros(mb * (1 << 20) // 8, dtype=torch.int64, device=dev)
    n_ops = 1 << 27
    for mode in (0, 10, 1, 2, 3, 4, 5, 6, 7, 8, 11):
        eng.bench_random_access(buf, n_ops, mode)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.bench_random_access(buf, n_ops, mode); e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        print("%5d MB  %-20s %7.1f G ops/s" % (mb, names[mode], n_ops / t / 1e9), flush=True)
    del buf
