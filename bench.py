#!/usr/bin/env python
"""bench.py — throughput of the trio discovery k-mer path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm

A step is one pass of the whole k-mer hot path over one synthetic trio
(BASELINE.json config "synthetic chr20-scale (64 Mbp) 30x trio ... 1 B200"):
child count -> threshold -> reference subtraction -> mother / father filtered
counts -> proband-unique set -> per-read distinct-hit scan.  The metric is
canonical k-mer instances counted + queried per second, summed over stages.
At N > 1 every rank holds 1/N of each sample's reads of an N x 64 Mbp genome
(weak scaling); the child table is partitioned by owner rank and k-mers are
routed to their owner by the binning kernel itself, over NVLink peer memory
(NCCL all-to-all where peer memory is unavailable).  `--total-genome-mbp 3000`
is BASELINE config 4 (whole genome, table hash-partitioned across the GPUs): the
genome is divided by N, samples come as lists of streams and the child count takes
as many hash-range passes as the memory plan asks for.
`e2e` is the same call on pinned HOST buffers in the decoder's batch format
(codes + the sparse validity list), H2D and D2H inside the timed region.
`parity_checked`: in the same run, the same chain (same N, same routes) on a bounded
sample is compared with the CPU port — stage sizes, the proband-unique keys and
every per-read (ndistinct, nhits) — and the run FAILS on any difference.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "canonical k-mers/s counted+queried (trio discovery k-mer path)"
UNIT = "k-mers/s"
SAMPLES = ("child", "mother", "father", "ref")
# algorithmic bytes per k-mer instance (SURVEY §8d / DESIGN.md "Roofline")
ALGO_BYTES = {
    # insert+count: key read + count read + count write
    ("count", 1): 16.0, ("count", 2): 24.0,
    # update-if-present / membership, miss: key read
    ("probe", 1): 8.0, ("probe", 2): 16.0,
    # hash-range binning: key written once
    ("bin", 1): 8.0, ("bin", 2): 16.0,
}


def kernel_class(name):
    """(class, reads the packed stream?) of a timed kernel name."""
    if name.startswith(("count_stream/mode0", "count_stream/mode1")):
        return "count", True
    if name.startswith("count_bins"):
        return "count", False
    if name.startswith("bin_stream"):
        return "bin", True
    if name.startswith("update_bins"):
        return "probe", False
    return "probe", True


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--genome-mbp", type=float, default=64.0, help="genome size per GPU (Mbp)")
    ap.add_argument("--total-genome-mbp", type=float, default=0.0,
                    help="whole-job genome size (Mbp), divided by the number of GPUs (BASELINE config 4: 3000)")
    ap.add_argument("--depth", type=float, default=30.0)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--denovo", type=int, default=100)
    ap.add_argument("--cpu-sample-mbp", type=float, default=8.0,
                    help="genome size of the bounded CPU / parity sample (same depth)")
    ap.add_argument("--n-passes", type=int, default=0, help="hash-range passes of the child count (0: memory plan)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-random-bench", action="store_true")
    ap.add_argument("--no-k-sweep", action="store_true")
    ap.add_argument("--no-wall", action="store_true", help="skip the BAM -> BED discovery wall time")
    ap.add_argument("--wall-mbp", type=float, default=0.0,
                    help="genome size of the BAM trio of the wall-time leg (0: the bench genome)")
    args = ap.parse_args()
    if args.total_genome_mbp:
        world = max(int(os.environ.get("WORLD_SIZE", "1")), 1)
        args.genome_mbp = args.total_genome_mbp / world
    return args


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region
    (B200_PROFILING.md "clocks DURING the timed region")."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.fh, stderr=subprocess.DEVNULL)
            # nvidia-smi takes a moment to print its first sample; the timed region
            # may last only a few hundred ms, so wait until the sampler is running
            t0 = time.time()
            while time.time() - t0 < 15.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        self.fh.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    pw.append(float(f[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def one_or_list(xs):
    return xs[0] if len(xs) == 1 else xs


def stream_to_host(torch, s, pin):
    """Device packed-stream dict -> HostStream on (pinned) host memory."""
    from kmer_denovo_filter_b200 import engine

    def h(t):
        c = torch.empty(t.shape, dtype=t.dtype, pin_memory=pin)
        c.copy_(t)
        return c
    keep = [h(s["codes"]), h(s["valid"]), h(s["read_starts"]), h(s["read_lens"])]
    hs = engine.HostStream(keep[0].numpy().view(np.uint64), keep[1].numpy().view(np.uint32),
                           s["n_bases"], keep[2].numpy().view(np.uint64),
                           keep[3].numpy().view(np.uint32))
    # the sparse form of the validity bitmap, as the BAM decoder emits it with every batch
    # (kdf_bam_batch.invalid_pos): computed once here, outside any timed region
    inv = engine.invalid_positions(hs.valid, hs.n_bases)
    if inv is not None:
        t = torch.empty(max(inv.shape[0], 1), dtype=torch.int32, pin_memory=pin)
        t[:inv.shape[0]].copy_(torch.from_numpy(inv.view(np.int32)))
        keep.append(t)
        hs.invalid = t.numpy().view(np.uint32)[:inv.shape[0]]
    return hs, keep


def host_bytes(hs, with_reads):
    """Bytes engine.upload copies for this stream (what crosses PCIe per step)."""
    sparse = hs.invalid is not None and os.environ.get("KDF_SPARSE_VALID", "1") != "0"
    n = hs.codes.nbytes + (hs.invalid.nbytes if sparse else hs.valid.nbytes)
    if with_reads:
        n += hs.read_starts.nbytes + hs.read_lens.nbytes
    return n


def to_device_stream(engine_mod, s):
    return engine_mod.DeviceStream(s["codes"], s["valid"], s["n_bases"], s["read_starts"],
                                   s["read_lens"])


def stream_host_tuple(s):
    return (s["codes"].cpu().numpy().view(np.uint64), s["valid"].cpu().numpy().view(np.uint32),
            int(s["n_bases"]), s["read_starts"].cpu().numpy().view(np.uint64),
            s["read_lens"].cpu().numpy().view(np.uint32))


def concat_host(parts):
    """Several host stream tuples as one: parts are laid word after word (the bits between
    a part's n_bases and its last word are invalid, i.e. separators)."""
    if len(parts) == 1:
        return parts[0]
    codes = np.concatenate([p[0] for p in parts])
    valid = np.concatenate([p[1] for p in parts])
    starts, lens, w = [], [], 0
    for p in parts:
        starts.append(p[3] + np.uint64(32 * w))
        lens.append(p[4])
        w += p[0].shape[0]
    n_bases = 32 * (w - parts[-1][0].shape[0]) + parts[-1][2]
    return codes, valid, int(n_bases), np.concatenate(starts), np.concatenate(lens)


def cpu_chain(sample, k, threads):
    """One pass of the CPU restatement (oracle C twin) over the sample."""
    from oracle import ckdf
    t0 = time.perf_counter()
    res = ckdf.discovery_chain(sample["child"], sample["mother"][:3], sample["father"][:3],
                               sample["ref"][:3], k, threads=threads,
                               child_capacity=max(int(sample["child"][2]) // 4, 1024))
    dt = time.perf_counter() - t0
    return res, dt


def make_host_sample(torch, dev, genome_bp, depth, read_len, denovo, world=1):
    """The bounded CPU / parity sample on the host: the union of the `world` rank shards the
    GPU arm runs (same generator, same seeds), the whole reference."""
    from kmer_denovo_filter_b200 import synth
    parts = {w: [] for w in ("child", "mother", "father")}
    for r in range(world):
        trio = synth.make_trio(torch, dev, int(genome_bp), depth=depth, read_len=read_len,
                               n_denovo=denovo, rank=r, world=world)
        for w in parts:
            parts[w].append(stream_host_tuple(trio[w]))
        del trio
    out = {w: concat_host(v) for w, v in parts.items()}
    out["child_read_counts"] = [int(p[3].shape[0]) for p in parts["child"]]
    ref = synth.pack_sequence_tensor(torch, synth.make_reference(torch, dev, int(genome_bp)))
    out["ref"] = stream_host_tuple(ref)
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return out


def sample_text(args):
    return ("1 pass of the same chain over a synthetic trio of a %.0f Mbp genome at %gx "
            "(same generator, read length, error/variant rates; %.3g of one GPU's k-mers per step)"
            % (args.cpu_sample_mbp, args.depth, args.cpu_sample_mbp / args.genome_mbp))


def digest(lo, hi):
    """sha256 of the sorted key list (hi:lo), the form both arms can produce."""
    lo = np.asarray(lo, dtype=np.uint64)
    hi = np.asarray(hi, dtype=np.uint64) if hi is not None else np.zeros_like(lo)
    order = np.lexsort((lo, hi))
    return hashlib.sha256(np.stack([hi[order], lo[order]], axis=1).tobytes()).hexdigest()[:16]


# --------------------------------------------------------------------------
# reference arm: the CPU restatement of the Jellyfish path, all host cores
# --------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import ckdf
    threads = ckdf.max_threads()     # every core of the box, whatever OMP_NUM_THREADS torchrun exported
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() \
        else torch.device("cpu")
    sample = make_host_sample(torch, dev, args.cpu_sample_mbp * 1e6, args.depth, args.read_len,
                              min(args.denovo, 100))
    for _ in range(args.warmup):
        cpu_chain(sample, args.k, threads)
    units, total = 0, 0.0
    for _ in range(args.steps):
        res, dt = cpu_chain(sample, args.k, threads)
        units += res["units"]
        total += dt
    value = units / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64" if args.k <= 32 else "u128", "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1)),
        # what a step of THIS arm actually ran: the bounded sample of that workload
        "ran": {"genome_bp_total": int(args.cpu_sample_mbp * 1e6), "depth": args.depth,
                "units_per_step": units // max(args.steps, 1), "host_threads": threads,
                "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample_text(args),
                         "note": "jellyfish/samtools/pysam are absent from this image; this is "
                                 "the C/OpenMP restatement of the Jellyfish path (oracle/kdf_oracle.c)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_sizes": {x: int(res[x]) for x in ("candidates", "non_ref", "after_mother",
                                                  "proband_unique")},
    }
    print(json.dumps(line))
    return 0


def workload_config(args, world):
    return {
        "workload": "synthetic trio, %g Mbp genome per GPU, %gx, 2x%d bp reads, %d de novo "
                    "events, discovery k-mer chain (child count, ref subtraction, 2 parent "
                    "count --if passes, per-read scan)" % (args.genome_mbp, args.depth,
                                                           args.read_len, args.denovo),
        "k": args.k, "min_child_count": 3, "parent_max_count": 0,
        "genome_bp_total": int(args.genome_mbp * 1e6) * world,
        "table": "hash-partitioned across %d GPU(s)" % world if world > 1 else "single GPU",
        "l2_policy": "inputs (0.7 GB per sample per 64 Mbp) and the child table (GBs) are larger than L2; no flush",
    }


# --------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------

def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from kmer_denovo_filter_b200 import engine, synth
    from kmer_denovo_filter_b200.discovery import kmer_chain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = engine.CudaEngine(dev)   # raises if the library or a GPU is missing: no fallback
    n_passes = args.n_passes or None

    # ---- synthetic trio, resident in HBM ---------------------------------
    genome_bp = int(args.genome_mbp * 1e6) * world
    t_gen = time.perf_counter()
    trio = synth.make_trio(torch, dev, genome_bp, depth=args.depth, read_len=args.read_len,
                           n_denovo=args.denovo, rank=rank, world=world)
    d = {w: one_or_list([to_device_stream(engine, s) for s in as_list(trio[w])]) for w in SAMPLES}
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    t_gen = time.perf_counter() - t_gen

    if world > 1:
        from kmer_denovo_filter_b200.discovery import kmer_chain_dist

        def step(streams, k=None):
            return kmer_chain_dist.discover_streams_dist(
                eng, streams["child"], streams["mother"], streams["father"], streams["ref"],
                k or args.k, fetch=False, n_passes=n_passes)
    else:
        def step(streams, k=None):
            return kmer_chain.discover_streams(
                eng, streams["child"], streams["mother"], streams["father"], streams["ref"],
                k or args.k, fetch=False, n_passes=n_passes)

    # valid k-mer instances of each resident stream (exact, measured once: a probe
    # pass against an empty table), for the per-kernel roofline arithmetic
    windows = {}
    tiny = eng.new_table(args.k, n_keys=16)
    for wname in SAMPLES:
        st0 = eng.new_stats()
        for s in as_list(d[wname]):
            eng.count_stream(tiny, s, engine.MODE_COUNT_IF_PRESENT, 0, 1, st0)
        windows[wname] = eng.read_stats(st0)["windows"]
    tiny.close()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(streams, n_steps, k=None):
        """-> (device ms max over ranks, units summed over ranks, last result)."""
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        units = 0
        res = None
        for _ in range(n_steps):
            res = step(streams, k)
            units += res["units"]
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            u = torch.tensor([units], dtype=torch.int64, device=dev)
            dist.all_reduce(u, op=dist.ReduceOp.SUM)
            units = int(u.item())
        return ms, units, res

    # ---- device-resident: warm-up, then K timed steps ---------------------
    for _ in range(args.warmup):
        step(d)
    torch.cuda.reset_peak_memory_stats(dev)
    eng.timers = {}
    launches0 = eng.launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, units, res = timed_region(d, args.steps)
    clk = clocks.stop() if rank == 0 else None
    launches = eng.launches - launches0
    ktimes = eng.kernel_times_ms()
    eng.timers = None
    value = units / (ms * 1e-3)
    peak_hbm = int(torch.cuda.max_memory_allocated(dev))

    # ---- roofline of the dominant call -----------------------------------
    kw = 1 if args.k <= 32 else 2
    per_kernel = {n: {"launches": len(v), "ms_total": float(sum(v)), "ms_avg": float(np.mean(v))}
                  for n, v in ktimes.items()}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    in_bytes = args.read_len / (4.0 * (args.read_len - args.k + 1))
    passes = int(res.get("n_passes", 1) or 1)

    def kmers_per_launch(name):
        """valid k-mer instances one launch of this (timed) call processes, averaged
        over its launches in the step (a multi-pass count re-reads the stream once per
        pass but bins every k-mer once: the algorithmic work is the k-mers)"""
        per_step = max(per_kernel[name]["launches"] / float(args.steps), 1.0)
        if name.startswith(("count_stream/mode2", "update_bins/mode2")):
            binned = name.startswith("update_bins")
            sel = [w for w, b in zip(("mother", "father"), res.get("parents_binned", [False, False]))
                   if bool(b) == binned] or ["mother", "father"]
            return sum(windows[w] for w in sel) / per_step
        if name.startswith("bin_stream"):
            # child + reference, and the parents that took the binned route (filter table > L2)
            tot = windows["child"] + windows["ref"]
            for who, b in zip(("mother", "father"), res.get("parents_binned", [])):
                tot += windows[who] if b else 0
            return tot / per_step
        if name.startswith("count_bins"):
            return float(windows["child"]) / per_step
        if name.startswith(("scan_stream_hits", "scan_reads", "count_stream/mode0")):
            return float(windows["child"]) / per_step
        if name.startswith("bin_keys"):
            return float(windows["child"] + windows["ref"]) / per_step
        return None

    def roofline_of(name):
        n = kmers_per_launch(name)
        if n is None:
            return None
        kclass, reads_stream = kernel_class(name)
        per_unit = ALGO_BYTES[(kclass, kw)] + (in_bytes if reads_stream else 0.0)
        ach = n * per_unit / (per_kernel[name]["ms_avg"] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                "algorithmic_bytes_per_kmer": per_unit, "kmers_per_launch": int(n),
                "gkmers_per_s": n / (per_kernel[name]["ms_avg"] * 1e-3) / 1e9,
                "kernel_ms_avg": per_kernel[name]["ms_avg"],
                "kernel_share_of_step": per_kernel[name]["ms_total"] / ms, "traffic": None}

    # the dominant call is the one with the largest total time in the step, whatever it
    # launches (count_bins is a C call that runs three kernels per table slice)
    timed = [n for n in per_kernel if kmers_per_launch(n) is not None]
    dom = max(timed, key=lambda n: per_kernel[n]["ms_total"]) if timed else None
    roofline = roofline_of(dom) if dom else None
    roofline_all = [r for r in (roofline_of(n) for n in per_kernel) if r is not None]
    prof = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.isfile(prof):
        try:
            traffic = json.load(open(prof))
            for r in ([roofline] if roofline else []) + roofline_all:
                t = traffic.get(r["kernel"])
                if t:
                    r["traffic"] = t.get("dram_bytes_per_launch")
                    r["traffic_source"] = t.get("source")
        except Exception:
            pass

    # ---- sector-granular random-access roofline (same box, same run) ------
    random_access = None
    if rank == 0 and not args.no_random_bench:
        buf = torch.zeros(1 << 30, dtype=torch.int64, device=dev)   # 8 GiB
        n_ops = 1 << 28
        random_access = {}
        for name, atomic in (("gather32", 0), ("gather32_atomic", 1)):
            eng.bench_random_access(buf, n_ops, atomic)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.bench_random_access(buf, n_ops, atomic)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3
            random_access[name] = {"gops": n_ops / t / 1e9,
                                   "sector_gbs": n_ops * (64 if atomic else 32) / t / 1e9}
        del buf
        # the ceiling that actually binds the L2-sliced child count: an L2-resident
        # 48 MB buffer, one 256-bit read per op, alone and with a returning compare-and-swap on
        # one op in six — the access mix of k_packed_keys
        small = torch.zeros((48 << 20) // 8, dtype=torch.int64, device=dev)
        for name, mode in (("l2_read256", 10), ("l2_read256_cas_1_in_6", 11)):
            eng.bench_random_access(small, n_ops, mode)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.bench_random_access(small, n_ops, mode)
            e1.record()
            torch.cuda.synchronize()
            random_access[name] = {"gops": n_ops / (e0.elapsed_time(e1) * 1e-3) / 1e9, "buffer_mb": 48}
        del small
        for r in ([roofline] if roofline else []) + roofline_all:
            if r["kernel"].startswith("count_bins"):
                r["l2_mixed_rate_fraction"] = r["gkmers_per_s"] / random_access["l2_read256_cas_1_in_6"]["gops"]
                r["l2_mixed_rate_note"] = (
                    "k-mers/s of the whole count_bins call (insert + mark + emit kernels) / ops/s of the "
                    "microbenchmark 'one 32 B read per op + returning CAS on one op in six' on an "
                    "L2-resident 48 MB buffer, same run: the rate the L2 sustains for this kernel's "
                    "access mix; occupancy, instruction count and atomic latency were each varied "
                    "without moving the kernel's time (profiles/r2c_count_variants.md)")
        if roofline is not None:
            for r in [roofline] + roofline_all:
                is_count = kernel_class(r["kernel"])[0] == "count"
                if kernel_class(r["kernel"])[0] == "bin":
                    continue
                ref_ops = random_access["gather32_atomic" if is_count else "gather32"]["gops"] * 1e9
                r["random_fraction"] = r["gkmers_per_s"] * 1e9 / ref_ops
                r["random_fraction_note"] = (
                    "k-mers/s of this kernel / ops/s of uniformly random 32 B sector %s over 8 GiB "
                    "(table >> L2) measured in this run; > 1 means the table traffic stays in L2 / "
                    "shared memory" % ("read + atomic add" if is_count else "reads"))
    for r in ([roofline] if roofline else []) + roofline_all:
        if r["kernel"].startswith("bin_stream_to_peers"):
            sent = r["kmers_per_launch"] * 8.0 * kw * (world - 1) / max(world, 1)
            r["interconnect"] = {
                "bound": "nvlink", "bytes_to_peers_per_launch": sent,
                "achieved_gbs_per_gpu": sent / (r["kernel_ms_avg"] * 1e-3) / 1e9,
                "reference": "profiles/r2g_peer_bin_sweep_n8.json: on 8 B200 the same kernel with ONE "
                             "bin per owner (longest runs) takes 22.8 ms per 1.53 G keys = 470 GB/s per "
                             "GPU; ncclAllToAll of the same bytes takes 48.8 ms (242 GB/s)"}
    if world > 1:
        dist.barrier()

    # ---- k sweep (BASELINE config 5) on the same resident trio ------------
    k_sweep = None
    if not args.no_k_sweep and world == 1 and not args.total_genome_mbp:
        k_sweep = {}
        for kk in (21, 31, 47, 63):
            if kk == args.k:
                k_sweep[str(kk)] = {"value": value, "ms_per_step": ms / args.steps}
                continue
            for _ in range(2):
                step(d, kk)
            sms, sunits, sres = timed_region(d, 3, kk)
            k_sweep[str(kk)] = {"value": sunits / (sms * 1e-3), "ms_per_step": sms / 3,
                                "key_bits": 64 if kk <= 32 else 128,
                                "proband_unique": int(sres["proband_unique"])}

    # ---- end to end: host buffers, H2D + D2H inside the timed region ------
    e2e = None
    if not args.no_e2e:
        hosts, keep, h2d = {}, [], 0
        for w in SAMPLES:
            hs_list = []
            for s in as_list(trio[w]):
                hs, kp = stream_to_host(torch, s, pin=True)
                hs_list.append(hs)
                keep.append(kp)
                h2d += host_bytes(hs, w == "child")
            hosts[w] = one_or_list(hs_list)
        del d
        del trio
        torch.cuda.empty_cache()
        step(hosts)
        ems, eunits, eres = timed_region(hosts, args.steps)
        d2h = 0
        if eres.get("reads") is not None:     # sparse per-read records + sorted hits
            d2h += sum(int(a.nbytes) for a in eres["reads"].values())
        elif eres["ndistinct"] is not None:
            d2h += 8 * int(eres["ndistinct"].shape[0])
        d2h += 8 * 16 + 4 * 32   # stage counters, bin cursors' flags and stats read back per step
        e2e = {"value": eunits / (ems * 1e-3), "unit": UNIT, "ms_per_step": ems / args.steps,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "kmer_denovo_filter_b200.discovery.kmer_chain.discover_streams(HostStream...) "
                      "-> libkdf_sm100 C ABI; pinned host buffers"}
        del hosts, keep
    else:
        del d
        del trio
    torch.cuda.empty_cache()

    # ---- parity (every N) + CPU baseline (rank 0, N = 1): the CPU port on a bounded sample,
    #      the GPU chain on the same sample with the routes of the run above ------------------
    cpu, parity = None, None
    if not args.no_parity or (world == 1 and not args.no_cpu_baseline):
        parity, cpu = parity_and_cpu(args, torch, dist, eng, engine, synth, kmer_chain, step, res,
                                     rank, world, dev, genome_bp)
        if args.no_cpu_baseline or world > 1:
            cpu = None

    # ---- discovery wall time: BAM trio -> candidate BED through the product pipeline ------
    wall = None
    if not args.no_wall and not args.total_genome_mbp:
        import bench_wall
        wall = bench_wall.discovery_wall(args, eng, rank, world)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "u64" if args.k <= 32 else "u128", "data": "synthetic",
            "config": workload_config(args, world),
            "units_per_step": units // args.steps,
            "stage_sizes": {x: int(res[x]) for x in ("candidates", "non_ref", "after_mother",
                                                      "proband_unique", "informative_reads")},
            "count_passes": passes, "peak_hbm_bytes_rank0": peak_hbm, "synth_seconds": t_gen,
            "roofline": roofline, "roofline_all": roofline_all, "kernels": per_kernel,
            "random_access": random_access, "parity_checked": parity,
            "cpu_baseline": cpu, "e2e": e2e, "k_sweep": k_sweep, "discovery_wall": wall,
            "gpu_launches": launches, "clocks": clk,
            "device": eng.props["name"],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def parity_and_cpu(args, torch, dist, eng, engine, synth, kmer_chain, step, main_res, rank, world, dev,
                   genome_bp):
    """Run the GPU chain (this N, the routes the timed run took) on the bounded sample and
    compare it with the CPU port on the union of the shards.  → (parity dict, cpu dict).
    Raises SystemExit on any difference: a bench line is only printed for equal results."""
    sample_bp = int(args.cpu_sample_mbp * 1e6)
    den = min(args.denovo, 100)
    trio = synth.make_trio(torch, dev, sample_bp, depth=args.depth, read_len=args.read_len,
                           n_denovo=den, rank=rank, world=world)
    ds = {w: to_device_stream(engine, trio[w]) for w in SAMPLES}
    # same routes as the timed run: table slices scaled with the genome so that the sample is
    # cut into as many hash ranges, and the parent route (filter / binned) forced to match
    saved = (kmer_chain.SLICE_BYTES, kmer_chain.PROBE_DIRECT_BYTES, kmer_chain.L2_TABLE_BYTES,
             os.environ.get("KDF_TABLE_FILTER"))
    scale = max(sample_bp / float(max(genome_bp, 1)), 1e-4)
    kmer_chain.SLICE_BYTES = max(1 << 20, int(saved[0] * min(scale, 1.0)))
    binned = bool(main_res.get("parents_binned")) and all(main_res["parents_binned"])
    if binned:
        kmer_chain.PROBE_DIRECT_BYTES = 0
        kmer_chain.L2_TABLE_BYTES = 0
        os.environ["KDF_TABLE_FILTER"] = "0"
    try:
        got = step(ds)
    finally:
        kmer_chain.SLICE_BYTES, kmer_chain.PROBE_DIRECT_BYTES, kmer_chain.L2_TABLE_BYTES = saved[:3]
        if saved[3] is None:
            os.environ.pop("KDF_TABLE_FILTER", None)
        else:
            os.environ["KDF_TABLE_FILTER"] = saved[3]
    sp = got["reads"] or {"read": np.zeros(0, np.uint64), "ndistinct": np.zeros(0, np.uint32),
                          "nhits": np.zeros(0, np.uint32)}
    mine = {"read": sp["read"], "nd": sp["ndistinct"], "nh": sp["nhits"]}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    else:
        gathered = [mine]
    out = (None, None)
    if rank == 0:
        from oracle import ckdf
        threads = ckdf.max_threads()
        sample = make_host_sample(torch, dev, sample_bp, args.depth, args.read_len, den, world)
        want, dt = cpu_chain(sample, args.k, threads)
        diffs = []
        for key in ("candidates", "non_ref", "after_mother", "proband_unique"):
            if int(got[key]) != int(want[key]):
                diffs.append("%s: gpu %d != cpu %d" % (key, got[key], want[key]))
        pu = got["pu"]
        g_lo = pu.lo.cpu().numpy().view(np.uint64) if pu is not None else np.zeros(0, np.uint64)
        g_hi = (pu.hi.cpu().numpy().view(np.uint64) if (pu is not None and pu.hi is not None)
                else np.zeros_like(g_lo))
        dg, dw = digest(g_lo, g_hi), digest(want["pu_lo"], want["pu_hi"])
        if dg != dw:
            diffs.append("proband-unique key digest: gpu %s != cpu %s" % (dg, dw))
        n_reads, off, n_with_hits = 0, 0, 0
        nd_w = want["nd"] if want["nd"] is not None else np.zeros(sum(sample["child_read_counts"]), np.uint32)
        nh_w = want["nh"] if want["nh"] is not None else np.zeros_like(nd_w)
        for r, cnt in enumerate(sample["child_read_counts"]):
            nd = np.zeros(cnt, np.uint32)
            nh = np.zeros(cnt, np.uint32)
            idx = gathered[r]["read"].astype(np.int64)
            nd[idx] = gathered[r]["nd"]
            nh[idx] = gathered[r]["nh"]
            if not (np.array_equal(nd, nd_w[off:off + cnt]) and np.array_equal(nh, nh_w[off:off + cnt])):
                diffs.append("per-read (ndistinct, nhits) of rank %d's shard differ" % r)
            n_with_hits += int((nh > 0).sum())
            off += cnt
            n_reads += cnt
        parity = {"ok": not diffs, "against": "oracle/kdf_oracle.c (CPU port), same run",
                  "sample": "%.0f Mbp genome x %gx trio, %d shard(s): the union of what the %d rank(s) ran"
                            % (args.cpu_sample_mbp, args.depth, world, world),
                  "routes": {"hash_ranges_per_gpu": int(got.get("n_passes", 1) or 1) * int(got.get("n_local", 1) or 1),
                             "parents": "binned + update_bins" if binned else "stream probe behind the L2 filter / shared-memory table",
                             "count_passes": int(got.get("n_passes", 1) or 1)},
                  "stage_sizes": {x: int(want[x]) for x in ("candidates", "non_ref", "after_mother",
                                                             "proband_unique")},
                  "pu_digest": dw, "reads_compared": n_reads, "reads_with_hits": n_with_hits,
                  "differences": diffs}
        cpu = {"value": want["units"] / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "seconds": dt, "sample": sample_text(args)}
        out = (parity, cpu)
        if diffs:
            sys.stderr.write("PARITY FAILURE: " + "; ".join(diffs) + "\n")
    flag = torch.tensor([0 if (out[0] is None or out[0]["ok"]) else 1], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()):
        raise SystemExit("bench.py: GPU results differ from the CPU port on the parity sample")
    del trio, ds
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    sys.exit(main())
