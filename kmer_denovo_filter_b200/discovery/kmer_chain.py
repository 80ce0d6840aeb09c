"""The k-mer part of the discovery path on packed streams (no BAM I/O).

This is the hot path BASELINE.json's metric is quoted on, in one call: what the
reference does with ``samtools fasta | jellyfish count``, ``dump -L``,
``query ref.jf``, two ``count --if`` + ``query`` parent passes and the per-read
``jellyfish query`` scan (reference ``discovery/pipeline.py:69-612`` and
``core/bam_scanner.py:340-507``), given the four read streams already decoded
and 2-bit packed (``include/kdf.h`` stream layout).

``run_discovery_pipeline`` reaches the same kernels batch by batch while it
decodes BAMs; ``bench.py`` and the parity tests call this function with whole
samples, either resident in HBM (``DeviceStream``) or in host memory
(``HostStream`` — uploaded here, inside the caller's timed region).
"""

import os

import numpy as np

from .. import engine as _engine
from ..kmer_utils import KmerSet

REF_PLANE = 1
L2_TABLE_BYTES = int(os.environ.get("KDF_L2_TABLE_MB", "100")) << 20   # keys of a read-only table that can stay L2-resident next to the stream
SMEM_TABLE_BYTES = 160 * 1024   # keys of a read-only table that the stream kernels copy to shared memory


class _Uploader:
    """Host → device copies on a side stream, in the order the chain consumes
    them, so that the parents' upload overlaps the child count (the copy engine
    and the SMs run concurrently; each consumer waits on its own event)."""

    def __init__(self, eng):
        self.eng = eng
        self.torch = eng.torch
        self.side = None

    def put(self, s, with_reads):
        """-> (DeviceStream, ready event | None)"""
        if isinstance(s, _engine.DeviceStream):
            return s, None
        torch = self.torch
        if self.eng.device.type != "cuda":      # the numpy engine of the CPU tests: no streams
            return self.eng.upload(s, with_reads=with_reads), None
        if self.side is None:
            self.side = getattr(self.eng, "_copy_stream", None)
            if self.side is None:
                self.side = self.eng._copy_stream = torch.cuda.Stream(device=self.eng.device)
        d = self.eng.upload(s, with_reads=with_reads, copy_stream=self.side)
        ev = torch.cuda.Event()
        ev.record(self.side)
        d.ready = ev
        return d, ev

    def put_chunked(self, s, with_reads):
        """Like :meth:`put`, but the stream is copied in chunks and can be consumed
        chunk by chunk (``DeviceStream.chunks``); the returned event covers all of it."""
        if isinstance(s, _engine.DeviceStream):
            return s, None
        torch = self.torch
        if self.side is None:
            self.side = getattr(self.eng, "_copy_stream", None)
            if self.side is None:
                self.side = self.eng._copy_stream = torch.cuda.Stream(device=self.eng.device)
        d, ev = self.eng.upload_chunked(s, self.side, with_reads=with_reads)
        d.ready = ev
        return d, ev

    def put_read_index(self, d, s):
        """read_starts / read_lens of host stream ``s`` onto ``d`` → event (or None)."""
        if isinstance(s, _engine.DeviceStream) or s.read_lens is None:
            return None
        return self.eng.upload_read_index(d, s, self.side)

    def wait(self, ev):
        if ev is not None:
            self.torch.cuda.current_stream(self.eng.device).wait_event(ev)


def default_child_capacity(eng, n_bases):
    return eng.capacity_for(max(n_bases // 8, 512))


# One table slice of the partitioned child count: 64 MB of keys + planes, half of
# B200's 126 MB L2, so a slice stays L2-resident while its bin is streamed through
# it with evict-first loads.  Measured on the 64 Mbp x 30x trio (k = 31): 32 MB
# slices (256 bins) 43.8 ms for bin + count, 64 MB (128 bins) 37.9 ms, 64 MB at
# load 0.56 (64 bins) 32.6 ms, 128 MB 44 ms (the slice no longer fits).
SLICE_BYTES = int(os.environ.get("KDF_SLICE_MB", "64")) << 20
MAX_PARTS = 256
# table slots per k-mer instance: 30x data holds ~12 instances per distinct k-mer
SLOTS_DIV = int(os.environ.get("KDF_COUNT_SLOTS_DIV", "8"))


def _pow2_at_least(x):
    p = 1
    while p < x:
        p *= 2
    return p


def _as_list(x):
    if x is None:
        return []
    return list(x) if isinstance(x, (list, tuple)) else [x]


def plan_partitions(n_windows_max, slots_needed=None, key_words=1, packed=False, max_parts=MAX_PARTS):
    """(n_parts, slice_capacity) for about ``slots_needed`` table slots in total.
    Default sizing: 8 bases per distinct k-mer (30x data holds ~14); a slice that
    turns out too small is reported by the kernel and the pass is redone.
    ``packed``: the slice holds keys only (``CudaEngine.count_bins_packed``), so the
    same L2 budget covers twice the slots and half as many bins are needed."""
    if slots_needed is None:
        slots_needed = max(n_windows_max // SLOTS_DIV, 1024)
    slice_slots = SLICE_BYTES // (8 * key_words + (0 if packed else 8))
    n_parts = _pow2_at_least((slots_needed + slice_slots - 1) // slice_slots)
    if max_parts:
        n_parts = min(max_parts, n_parts)
    slice_cap = max(1024, (slots_needed + n_parts - 1) // n_parts)
    return n_parts, (slice_cap + 3) & ~3


def _bin_capacity(n_keys_max, n_parts):
    """Slots per bin for ``n_keys_max`` k-mer INSTANCES spread over ``n_parts``
    hash ranges.  All copies of a k-mer share a bin, so the spread of a bin's
    size is that of ~mean/depth keys of weight ~depth, not of ``mean`` unit
    keys: 4 % + 64 sqrt(mean) covers depths into the hundreds.  A bin that is
    still too small is reported by the kernel and the caller re-bins with the
    exact sizes."""
    mean = n_keys_max / n_parts
    return int(mean * 1.04 + 64 * (mean ** 0.5) + 1024)


def bin_budget_bytes(eng):
    """Device memory the bins of one counting pass may take: ``KDF_BIN_BUDGET_GB``, else
    half of what is free on the device now (free + the allocator's cached blocks)."""
    env = os.environ.get("KDF_BIN_BUDGET_GB")
    if env:
        return int(float(env) * (1 << 30))
    fn = getattr(eng, "free_device_bytes", None)
    if fn is None:
        return 1 << 62
    # cudaMemGetInfo takes the driver lock (measured: several ms per call while nvidia-smi
    # polls the device, as bench.py's clock sampler does), so the answer is kept for a while
    import time
    now = time.monotonic()
    cached = getattr(eng, "_bin_budget", None)
    if cached is None or now - cached[0] > 30.0:
        cached = eng._bin_budget = (now, int(fn() * 0.5))
    return cached[1]


def plan_child_count(eng, n_child, n_ref, k, min_child_count, n_parts=None, slice_capacity=None,
                     n_passes=None, max_local=MAX_PARTS):
    """How the child count of ``n_child`` k-mer instances (+ ``n_ref`` reference ones) is
    laid out: ``(n_passes, n_local, slice_capacity)``.  The hash space is cut into
    ``n_passes * n_local`` ranges, each counted in one L2-sized table slice; a pass
    re-extracts the streams and bins only its own ``n_local`` ranges, so the bins of a
    pass hold ``1 / n_passes`` of the k-mers.  One pass when the bins of the whole sample
    fit :func:`bin_budget_bytes` and ``max_local`` ranges are enough — which is what
    Jellyfish's sized hash + spill/merge does for the reference
    (``core/jellyfish_wrappers.py:73-107, 335-366``)."""
    kw = eng.lib.kdf_key_words(k)
    packed = eng.count_bins_packed(k, min_child_count)
    n_total, s_auto = plan_partitions(n_child, key_words=kw, packed=packed, max_parts=0)
    if n_parts is not None:
        n_total = n_parts * (n_passes or 1)
    if n_passes is None:
        need = (n_child + n_ref) * 1.06 * 8 * kw
        n_passes = _pow2_at_least(int(-(-need // max(bin_budget_bytes(eng), 1))))
        n_passes = max(n_passes, n_total // max_local)
    n_total = max(n_total, n_passes)
    n_local = n_total // n_passes
    if slice_capacity is None:
        slots = max(n_child // SLOTS_DIV, 1024)
        slice_capacity = (max(1024, -(-slots // n_total)) + 3) & ~3
    return n_passes, n_local, slice_capacity


def _wait_ready(eng, s):
    if getattr(s, "ready", None) is not None:
        eng.torch.cuda.current_stream(eng.device).wait_event(s.ready)


def _bin_into(eng, bins, streams, stats, pass_):
    for s in streams:
        if getattr(s, "chunks", None):
            main = eng.torch.cuda.current_stream(eng.device)
            for first, n, ev in s.chunks:      # bin a chunk as soon as it has landed
                main.wait_event(ev)
                eng.bin_stream(bins, s, stats, word_range=(first, n), pass_=pass_)
        else:
            _wait_ready(eng, s)
            eng.bin_stream(bins, s, stats, pass_=pass_)


def count_child_partitioned(eng, child_streams, ref_streams, k, min_child_count,
                            n_parts=None, slice_capacity=None, n_passes=None):
    """Module 1 + reference subtraction without a table in HBM: bin the child's
    (and the reference's) canonical k-mers by hash range, then count each bin in
    an L2-resident table slice and emit the keys with count >= min_child_count
    that are not in the reference.  When the bins of the whole sample would not fit
    the device the hash ranges are taken in several passes (:func:`plan_child_count`).

    ``child_streams`` / ``ref_streams``: lists of DeviceStream.  Returns dict:
    child_windows, ref_windows, child_distinct, candidates, non_ref, lo, hi."""
    torch = eng.torch
    n_max = sum(s.n_bases for s in child_streams)
    r_max = sum(s.n_bases for s in ref_streams)
    n_passes, n_local, slice_capacity = plan_child_count(
        eng, n_max, r_max, k, min_child_count, n_parts, slice_capacity, n_passes)
    plog = n_passes.bit_length() - 1
    if (1 << plog) != n_passes:
        raise _engine.KdfError("the number of counting passes must be a power of two")
    bin_cap = _bin_capacity(n_max / n_passes, n_local)
    ref_cap = _bin_capacity(r_max / n_passes, n_local)
    out_cap = max(1 << 16, n_max // 64 // n_passes)
    st_c, st_r = eng.new_stats(), eng.new_stats()
    cb = rb = None
    tot = {"distinct": 0, "n_count": 0, "n_out": 0}
    los, his = [], []
    for p in range(n_passes):
        pass_ = (plog, p) if n_passes > 1 else None
        while True:   # bins: retry with exact sizes if the hash ranges are skewed
            if cb is None:
                cb = eng.new_bins(k, n_local, bin_cap)
            if rb is None and ref_streams:
                rb = eng.new_bins(k, n_local, ref_cap)
            st_c1, st_r1 = eng.new_stats(), eng.new_stats()
            _bin_into(eng, cb, child_streams, st_c1, pass_)
            if rb is not None:
                _bin_into(eng, rb, ref_streams, st_r1, pass_)
            over_c = cb.overflowed()
            over_r = rb.overflowed() if rb is not None else False
            if not over_c and not over_r:
                st_c += st_c1
                st_r += st_r1
                break
            if over_c:
                bin_cap = int(cb.counts().max()) + 4
                cb = None
            else:
                cb.reset()
            if over_r:
                ref_cap = int(rb.counts().max()) + 4
                rb = None
            elif rb is not None:
                rb.reset()
        while True:   # slices / output: retry when a slice was full or the output too small
            res = eng.count_bins(cb, rb, slice_capacity, min0=min_child_count, max1=0,
                                 count_min0=min_child_count, out_cap=out_cap, pass_=pass_)
            if res["full"]:
                if slice_capacity >= 2 * bin_cap:
                    raise _engine.KdfError("child k-mer table slice full at %d slots" % slice_capacity)
                slice_capacity = min(slice_capacity * 4, 2 * bin_cap + 4)
                continue
            if res["n_out"] > out_cap:
                out_cap = res["n_out"]
                continue
            break
        for key in tot:
            tot[key] += res[key]
        los.append(res["lo"])
        if res["hi"] is not None:
            his.append(res["hi"])
        if p + 1 < n_passes:
            cb.reset()
            if rb is not None:
                rb.reset()
    del cb, rb
    lo = los[0] if len(los) == 1 else torch.cat(los)
    hi = (his[0] if len(his) == 1 else torch.cat(his)) if his else None
    return {"child_windows": eng.read_stats(st_c)["windows"],
            "ref_windows": eng.read_stats(st_r)["windows"] if ref_streams else 0,
            "child_distinct": tot["distinct"], "candidates": tot["n_count"],
            "non_ref": tot["n_out"], "lo": lo, "hi": hi,
            "n_parts": n_local, "n_passes": n_passes, "slice_capacity": slice_capacity}


# A parent stream is probed against a read-only table (``count --if``).  While the
# table's keys fit L2 the stream kernel probes it directly; past that every probe
# would be a random DRAM sector (measured 37 G/s against ~150 G/s from L2), so the
# parent's k-mers are first binned by hash range — a streaming pass — and the bins are
# applied one after the other, each touching only its own L2-sized share of the table.
PROBE_DIRECT_BYTES = int(os.environ.get("KDF_PROBE_DIRECT_MB", "128")) << 20   # keys of a table probed straight from the stream
PROBE_SLICE_BYTES = 48 << 20
PROBE_CHUNK_BASES = 1 << 30      # bins of one chunk: <= 8.6 GB (64-bit keys)


def count_if_present(eng, table, d_stream, stats, plane=0, arg=1):
    """``jellyfish count --if`` of one parent stream, or of a list of streams that
    together are the parent (discovery/pipeline.py:377-386).
    → True when the stream was binned first (the large-table route)."""
    if isinstance(d_stream, (list, tuple)):
        binned = False
        for s in d_stream:
            _wait_ready(eng, s)
            binned = count_if_present(eng, table, s, stats, plane, arg) or binned
        return binned
    key_bytes = table.capacity * 8 * table.key_words
    if (key_bytes <= PROBE_DIRECT_BYTES or getattr(table, "filter_buf", None) is not None
            or os.environ.get("KDF_PROBE_DIRECT") == "1"):
        eng.count_stream(table, d_stream, _engine.MODE_COUNT_IF_PRESENT, plane, arg, stats)
        return False
    # at least 16 bins: with <= 8 the binning kernel takes its owner-routing form
    n_parts = min(MAX_PARTS, max(16, _pow2_at_least((key_bytes + PROBE_SLICE_BYTES - 1) // PROBE_SLICE_BYTES)))
    n_words = (d_stream.n_bases + 31) // 32
    chunk_words = PROBE_CHUNK_BASES // 32
    bin_cap = _bin_capacity(min(d_stream.n_bases, PROBE_CHUNK_BASES), n_parts)
    first = 0
    while first < n_words:
        n = min(chunk_words, n_words - first)
        while True:   # a skewed hash range: retry this chunk with the exact size
            bins = eng.new_bins(table.k, n_parts, bin_cap)
            eng.bin_stream(bins, d_stream, None, word_range=(first, n))
            if not bins.overflowed():
                break
            bin_cap = int(bins.counts().max()) + 4
            del bins
        eng.update_bins(table, bins, _engine.MODE_COUNT_IF_PRESENT, plane, arg, stats)
        del bins
        first += n
    return True


def _primed_table(eng, k, lo, hi, n):
    # small sets get load 0.25: they stay within the shared-memory budget of the
    # stream kernels and almost no probe has to look past its home bucket
    # larger read-only sets get load 0.33: a third as many probes meet a full home
    # bucket (4.6 % instead of 14 %), which is worth more than the extra L2 footprint
    n_keys = max(n, 1)
    kw = eng.lib.kdf_key_words(k)
    # a two-bit filter of 4 bytes per key (rounded up to a power of two) in front of the
    # table keeps the stream probe L2-resident whatever the table's size — as long as the
    # FILTER fits the ~64 MB of L2 that data shared by both dies can use
    filtered = eng.filter_applies(k, n_keys)
    if n_keys * 4 * 8 * kw <= SMEM_TABLE_BYTES:
        n_keys *= 2
    elif n_keys * 3 * 8 * kw <= L2_TABLE_BYTES:
        n_keys = n_keys * 3 // 2     # still L2-resident at load 0.33
    elif not filtered and n_keys * 2 * 8 * kw > PROBE_DIRECT_BYTES:
        n_keys = n_keys * 3 // 2     # probed bin by bin (count_if_present): its size is free
    t = eng.new_table(k, n_keys=n_keys)
    eng.update_keys(t, lo, hi, _engine.MODE_INSERT_ONLY, 0, 0)
    if filtered:
        eng.build_filter(t, max(n, 1), eng.FILTER_MAX_BYTES)
    return t


def merge_sparse_records(parts, streams):
    """Per-read records of several streams of one sample (``scan_reads_sparse`` each) as
    one record set: read indices, hit positions and ``first`` continue from one stream to
    the next (stream i starts at read sum(n_reads[:i]), base sum(n_bases[:i]))."""
    if len(parts) == 1:
        return parts[0]
    out = {key: [] for key in parts[0]}
    r0 = b0 = h0 = 0
    for sp, d in zip(parts, streams):
        out["read"].append(sp["read"] + np.uint64(r0))
        out["first"].append(sp["first"] + np.uint64(h0))
        out["hit_pos"].append(sp["hit_pos"] + np.uint64(b0))
        for key in ("ndistinct", "nhits", "hit_slot"):
            out[key].append(sp[key])
        r0 += d.n_reads
        b0 += d.n_bases
        h0 += int(sp["hit_pos"].shape[0])
    return {key: np.concatenate(v) for key, v in out.items()}


def discover_streams(eng, child, mother, father, ref, k, min_child_count=3,
                     parent_max_count=0, min_distinct_kmers_per_read=None,
                     child_capacity=None, want_hits=False, fetch=True, partitioned=None,
                     sparse_scan=True, n_passes=None):
    """Child count → threshold → reference subtraction → mother / father
    filtered counts → proband-unique set → per-read distinct-hit reduction.

    Returns a dict: stage sizes (``candidates``, ``non_ref``, ``after_mother``,
    ``proband_unique``), ``pu`` (:class:`KmerSet` or None), per-read
    ``ndistinct`` / ``nhits`` (numpy u32 when ``fetch`` else device tensors),
    ``informative_reads`` and ``units`` = valid k-mer instances processed over
    all stages (the unit of BASELINE.json's metric)."""
    if min_distinct_kmers_per_read is None:
        min_distinct_kmers_per_read = max(1, k // 4)   # discovery/pipeline.py:2119-2121
    if partitioned is None:
        partitioned = child_capacity is None
    stats = eng.new_stats()
    up = _Uploader(eng)
    # a sample may come as a list of streams (a whole-genome sample exceeds the 2^32
    # bases one stream's sparse validity list can address): lists take the partitioned,
    # sparse-scan route
    multi = any(isinstance(x, (list, tuple)) for x in (child, mother, father, ref))
    if multi and not (partitioned and sparse_scan):
        raise _engine.KdfError("lists of streams need partitioned=True and sparse_scan=True")
    # copy order = consumption order: child (in chunks, binned as they land), ref, mother,
    # father, and last the child's read index (12 bytes per read, needed by the scan only:
    # ahead of the parents it delayed the father's arrival, which the chain waits for)
    if partitioned:
        d_childs = [up.put_chunked(x, False)[0] for x in _as_list(child)]
        d_refs = [up.put(x, False)[0] for x in _as_list(ref)]
        d_child, d_ref = d_childs[0], (d_refs[0] if d_refs else None)
        ev_child = None     # every stream carries its own `ready` event
    else:
        d_child, ev_child = up.put(child, True)
        d_ref, ev_ref = up.put(ref, False)
        up.wait(ev_child)
        up.wait(ev_ref)
        d_childs, d_refs = [d_child], [d_ref]
    d_mothers = [up.put(x, False)[0] for x in _as_list(mother)]
    d_fathers = [up.put(x, False)[0] for x in _as_list(father)]
    d_mother, d_father = d_mothers, d_fathers
    ev_mother = ev_father = None
    ev_reads = ([up.put_read_index(d, h) for d, h in zip(d_childs, _as_list(child))]
                if partitioned else [])

    # Module 1 + reference subtraction: jellyfish count -C ; dump -L ; query ref.jf
    if partitioned:
        c = count_child_partitioned(eng, d_childs, d_refs, k, min_child_count, n_passes=n_passes)
        n_cand, n_nonref, lo, hi = c["candidates"], c["non_ref"], c["lo"], c["hi"]
        child_windows, child_distinct = c["child_windows"], c["child_distinct"]
        child_capacity = c["n_passes"] * c["n_parts"] * c["slice_capacity"]
        units_binned = c["child_windows"] + c["ref_windows"]
    else:
        # direct form: one table in HBM, sized for 4 bases per distinct k-mer; when
        # that is too small the count is redone in a larger table — never a silent
        # drop (Jellyfish would spill to .jf_N files and merge).
        units_binned = 0
        if child_capacity is None:
            child_capacity = default_child_capacity(eng, d_child.n_bases)
        while True:
            table = eng.new_table(k, capacity=child_capacity)
            eng.count_stream(table, d_child, _engine.MODE_INSERT_COUNT, 0, 1, stats)
            st = eng.read_stats(stats)
            if not st["full"]:
                child_windows, child_distinct = st["windows"], st["new"]
                break
            table.close()
            if child_capacity >= 2 * d_child.n_bases:
                raise _engine.KdfError("child k-mer table full at %d slots" % child_capacity)
            child_capacity = min(child_capacity * 4, eng.capacity_for(d_child.n_bases))
            stats.zero_()
        # reference subtraction: stream the reference against the child table
        eng.count_stream(table, d_ref, _engine.MODE_MARK_IF_PRESENT, REF_PLANE, 1, stats)
        n_cand = eng.threshold_count(table, min0=min_child_count)
        n_nonref, lo, hi, _a, _b = eng.threshold_compact(table, min0=min_child_count, max1=0)
        eng.check_not_full(stats)
        table.close()
    out = {"child_windows": child_windows, "child_distinct": child_distinct,
           "child_capacity": child_capacity, "candidates": n_cand,
           "n_passes": c["n_passes"] if partitioned else 1,
           "n_local": c["n_parts"] if partitioned else 1,
           "non_ref": n_nonref, "after_mother": 0, "proband_unique": 0,
           "pu": None, "ndistinct": None, "nhits": None, "informative_reads": 0, "hits": None,
           "reads": None, "parents_binned": []}

    # Module 2: count --if against mother, then father
    n_pu = 0
    if n_nonref:
        mt = _primed_table(eng, k, lo, hi, n_nonref)
        up.wait(ev_mother)
        out["parents_binned"].append(count_if_present(eng, mt, d_mother, stats))
        n_am, lo, hi, _a, _b = eng.threshold_compact(mt, max0=parent_max_count)
        mt.close()
        out["after_mother"] = n_am
        if n_am:
            ft = _primed_table(eng, k, lo, hi, n_am)
            up.wait(ev_father)
            out["parents_binned"].append(count_if_present(eng, ft, d_father, stats))
            n_pu, lo, hi, _a, _b = eng.threshold_compact(ft, max0=parent_max_count)
            ft.close()
    out["proband_unique"] = n_pu

    # Modules 2b + 3: membership table, per-read scan
    if n_pu:
        out["pu"] = KmerSet(eng, k, lo, hi)
        pt = _primed_table(eng, k, lo, hi, n_pu)
        up.wait(ev_child)
        for ev in ev_reads:
            up.wait(ev)
        if sparse_scan:
            # hits are rare: emit them from a streaming probe, reduce per read on the
            # device, return one record per read that has hits
            parts = []
            for d in d_childs:
                _wait_ready(eng, d)
                parts.append(eng.scan_reads_sparse(pt, d, stats=stats))
            sp = merge_sparse_records(parts, d_childs)
            out["reads"] = sp
            out["informative_reads"] = int((sp["ndistinct"] >= min_distinct_kmers_per_read).sum())
            if fetch:
                n_reads = sum(d.n_reads for d in d_childs)
                nd = np.zeros(n_reads, dtype=np.uint32)
                nh = np.zeros(n_reads, dtype=np.uint32)
                nd[sp["read"].astype(np.int64)] = sp["ndistinct"]
                nh[sp["read"].astype(np.int64)] = sp["nhits"]
                out["ndistinct"], out["nhits"] = nd, nh
            pt.close()
        else:
            res = eng.scan_reads(pt, d_child, min_distinct=min_distinct_kmers_per_read,
                                 stats=stats, want_hits=want_hits)
            nd, nh = res["ndistinct"], res["nhits"]
            if fetch:
                nd = nd.cpu().numpy().view(np.uint32)
                nh = nh.cpu().numpy().view(np.uint32)
                out["informative_reads"] = int((nd >= min_distinct_kmers_per_read).sum())
            out["ndistinct"], out["nhits"] = nd, nh
            if want_hits:
                out["hits"] = (res["hit_pos"], res["hit_slot"], pt)
            else:
                pt.close()
    st = eng.read_stats(stats)
    out["units"] = st["windows"] + units_binned
    return out
