"""Host-side tuning of the product pipeline on the GPU box: write the bench trio once
(--make), then time run_discovery_pipeline under the environment it is started with.
    python scripts/wall_sweep.py --dir /tmp/kdf_sweep --make
    KDF_BAM_CHUNK_KB=16384 python scripts/wall_sweep.py --dir /tmp/kdf_sweep --label chunk16m
Prints one JSON line per call (label, environment switches, wall seconds of every repeat,
stage times of the best one)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", required=True)
    ap.add_argument("--make", action="store_true")
    ap.add_argument("--genome-mbp", type=float, default=64.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--label", default="default")
    a = ap.parse_args()
    import torch
    import bench_wall
    from kmer_denovo_filter_b200 import engine
    from kmer_denovo_filter_b200.discovery import pipeline
    eng = engine.CudaEngine(torch.device("cuda", 0))
    meta = os.path.join(a.dir, "paths.json")
    if a.make:
        os.makedirs(a.dir, exist_ok=True)
        paths, _ev, gen = bench_wall.make_bam_trio(torch, eng.device, int(a.genome_mbp * 1e6), 30.0, 150, 100,
                                                    a.dir, os.cpu_count() or 4)
        json.dump(paths, open(meta, "w"))
        print(json.dumps({"made": paths, "seconds": gen["seconds"]}))
        return
    paths = json.load(open(meta))
    threads = os.cpu_count() or 4
    walls, best = [], None
    for _ in range(a.reps):
        pargs = bench_wall.discovery_args(paths, os.path.join(a.dir, "out_" + a.label), 31, threads)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipeline.run_discovery_pipeline(pargs, engine=eng)
        torch.cuda.synchronize()
        w = time.perf_counter() - t0
        walls.append(round(w, 4))
        if best is None or w < best[0]:
            best = (w, {k: round(v, 4) for k, v in pipeline.LAST_TIMINGS.items()})
    env = {k: v for k, v in os.environ.items() if k.startswith("KDF_") or k.startswith("OMP_") or k.startswith("GOMP_")}
    print(json.dumps({"label": a.label, "env": env, "wall_s": walls, "best_stages": best[1]}))


if __name__ == "__main__":
    main()
