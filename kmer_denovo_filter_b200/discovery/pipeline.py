"""VCF-free discovery pipeline on the GPU k-mer engine.

Function names, argument meaning, return values and every output file follow
the reference's ``discovery/pipeline.py``; the Jellyfish / samtools /
``jellyfish query`` work of Modules 0-3 runs as sm_100a kernels through
``libkdf_sm100.so``.  The clustering, SV annotation and the writers are the
reference's CPU steps, restated here (they consume the per-read reduction the
GPU produces).

Reference map (``discovery/pipeline.py``):
  _extract_child_kmers_discovery :69    _subtract_reference_kmers :271
  _count_parent_jellyfish :322          _filter_parents_discovery :462
  _anchor_and_cluster :615              _write_bed :1156
  _write_bedgraph :1197                 _write_read_coverage_bed :1281
  _annotate_and_link_from_metadata :1351  _write_bedpe :1492
  _classify_regions :1517               run_discovery_pipeline :2093
"""

import bisect
import collections
import json
import logging
import os
import sys
import time

import numpy as np

from .. import bamio
from .. import engine as _engine
from ..core import kmer_engine_wrappers as kw
from ..core.kmer_engine_wrappers import (  # noqa: F401  (re-exported like the reference)
    _build_proband_jf_index,
    _ensure_ref_jf,
    get_engine,
)
from ..kmer_utils import KmerSet, canonicalize
from . import kmer_chain

logger = logging.getLogger(__name__)

REF_PLANE = 1  # plane 1 of the child table holds the "in reference" flag

# wall-clock seconds of the stages of the last run_discovery_pipeline call (bench.py's
# discovery_wall leg reads it; the reference logs per-module times, discovery/pipeline.py:2185-2194)
LAST_TIMINGS = {}


class ChildScanCache:
    """The child BAM decoded ONCE.  The reference reads it three times (``samtools fasta``
    for the count, the pysam scan of Module 3, the informative-reads writer); here one
    scan-mode decode (with metadata) feeds all three: the counting stream is the same
    packed codes behind the ``fasta_keep`` mask (``bamio.counting_view``), the batches stay
    in host memory for the per-read scan (and for the later passes of a multi-pass
    count), and the writer fetches the few informative records back by their offsets
    (``kdf_bam_fetch_records``).  ``KDF_CHILD_CACHE_GB`` bounds the host memory (default
    64; 0 disables the cache); past it the batches are dropped and every later consumer
    decodes the file again, as the reference does."""

    def __init__(self, path, threads, batch_bases=kw.BATCH_BASES):
        self.path = path
        self.threads = threads
        self.batch_bases = batch_bases
        self.pf = kw.take_prefetch(path, bamio.MODE_SCAN, threads, True, batch_bases)
        self.reader = self.pf.reader
        self.batches = []
        self.nbytes = 0
        self.limit = int(float(os.environ.get("KDF_CHILD_CACHE_GB", "64")) * (1 << 30))
        self.complete = False      # every batch of the file is held in `batches`
        self.decoded = False       # the first pass over the file is over
        self.reads = self.bases = 0
        self.hit_reads = []        # (uncompressed offset, qname, is_supplementary) of reads with hits

    @staticmethod
    def _batch_bytes(b):
        n = b.codes.nbytes + b.valid.nbytes + b.read_starts.nbytes + b.read_lens.nbytes
        for name in ("qname_blob", "cigar_blob", "sa_blob", "invalid"):
            a = getattr(b, name, None)
            n += a.nbytes if a is not None else 0
        return n + 64 * b.n_reads

    def stream(self):
        """The scan-mode batches of the file: the first call decodes it (keeping the
        batches while they fit), later calls replay the kept batches or decode again.
        The caller must not close what it gets."""
        if self.complete:
            for b in self.batches:
                yield b
            return
        if self.decoded:       # the batches did not fit: another pass over the file
            pf = kw.BamPrefetcher(self.path, bamio.MODE_SCAN, self.threads, True, self.batch_bases)
            try:
                for b in pf:
                    yield b
                    b.close()
            finally:
                pf.close()
            return
        keep = self.limit > 0
        for b in self.pf:
            self.reads += b.n_reads
            self.bases += b.n_bases
            if keep:
                self.nbytes += self._batch_bytes(b)
                if self.nbytes > self.limit:
                    logger.info("  child batches exceed KDF_CHILD_CACHE_GB: later stages decode %s again",
                                self.path)
                    keep = False
                    for old in self.batches:
                        old.close()
                    self.batches = []
            if keep:
                self.batches.append(b)
            yield b
            if not keep:
                b.close()
        self.decoded = True
        self.complete = keep

    def release_batches(self):
        """The per-read scan is done: only the hit list and the reader are still needed."""
        for b in self.batches:
            b.close()
        self.batches = []
        self.complete = False

    def note_hit(self, batch, r):
        rec = batch.record(r)
        self.hit_reads.append((int(batch.rec_uoff[r]), rec.query_name, rec.is_supplementary))

    def close(self):
        self.release_batches()
        self.pf.close()


_CHILD_CACHE = {}


def child_cache(path):
    return _CHILD_CACHE.get(path)


def drop_child_caches():
    for key in list(_CHILD_CACHE):
        _CHILD_CACHE.pop(key).close()


class ChildCandidates:
    """Module-1 handle: the child's canonical k-mers in hash-range bins (HBM) plus
    the counts of the first pass; replaces ``child.jf`` / ``child_candidates.fa``.
    The count table itself never exists in HBM: every pass re-counts the bins in
    an L2-resident slice (``kdf_count_bins``).  When the bins of the whole sample would
    not fit the device (``kmer_chain.plan_child_count``: a whole-genome child is 8 bytes
    x 9e10 k-mer instances) the hash ranges are taken in ``n_passes`` groups: each pass
    re-bins the packed batches (from host memory, or by decoding the file again) and
    counts only its own ranges — the analogue of Jellyfish's sized hash with spill files
    (``core/jellyfish_wrappers.py:73-107, 335-366``)."""

    def __init__(self, eng, k, source, n_passes, n_local, slice_capacity, min_child_count, stats,
                 bins0=None, bin_cap=None):
        self.engine = eng
        self.k = k
        self.source = source            # ChildScanCache
        self.n_passes = n_passes
        self.n_local = n_local
        self.slice_capacity = slice_capacity
        self.min_child_count = min_child_count
        self.n_candidates = 0
        self.stats = stats              # windows (k-mer instances), new (distinct), reads, bases
        self.bins = bins0               # the bins of pass 0 (the only pass, usually): kept
        self.bin_cap = bin_cap

    def _pass(self, p):
        return (self.n_passes.bit_length() - 1, p) if self.n_passes > 1 else None

    def _bins_of_pass(self, p):
        """Hash-range bins of pass ``p``: the kept ones, or re-binned from the batches."""
        eng = self.engine
        if p == 0 and self.bins is not None:
            return self.bins
        while True:
            bins = eng.new_bins(self.k, self.n_local, self.bin_cap)
            for b in self.source.stream():
                eng.bin_stream(bins, eng.upload(bamio.counting_view(b), with_reads=False), None,
                               pass_=self._pass(p))
            if not bins.overflowed():
                return bins
            self.bin_cap = int(bins.counts().max()) + 4
            logger.info("  hash ranges are skewed: re-binning with %d slots per bin", self.bin_cap)

    def count(self, ref_index=None, out_cap=1, want_planes=False, **thresholds):
        """One counting pass over all hash ranges → the ``eng.count_bins`` result (summed
        over the passes); a slice that turns out too small is enlarged and redone."""
        eng = self.engine
        torch = eng.torch
        tot = {"n_out": 0, "keys": 0, "hits": 0, "distinct": 0, "n_count": 0, "occupied": 0}
        parts = {"lo": [], "hi": [], "p0": [], "p1": []}
        for p in range(self.n_passes):
            bins = self._bins_of_pass(p)
            ref_bins = ref_index.to_bins(eng, self.n_local, self._pass(p)) if ref_index is not None else None
            cap = out_cap
            while True:
                res = eng.count_bins(bins, ref_bins, self.slice_capacity, out_cap=cap,
                                     count_min0=self.min_child_count, want_planes=want_planes,
                                     pass_=self._pass(p), **thresholds)
                if res["full"]:
                    if self.slice_capacity >= 2 * bins.bin_cap:
                        raise _engine.KdfError("k-mer table slice full at %d slots" % self.slice_capacity)
                    self.slice_capacity = min(self.slice_capacity * 4, 2 * bins.bin_cap + 4)
                    logger.info("  growing the table slice to %d slots", self.slice_capacity)
                    continue
                if res["n_out"] > cap and cap > 1:
                    cap = res["n_out"]
                    continue
                break
            for key in tot:
                tot[key] += res[key]
            for key in parts:
                if res.get(key) is not None:
                    parts[key].append(res[key])
            if not (p == 0 and bins is self.bins):
                del bins
        out = dict(tot)
        out["full"] = 0
        for key, v in parts.items():
            out[key] = (v[0] if len(v) == 1 else torch.cat(v)) if v else None
        return out

    def dump(self, **thresholds):
        """All (key, count, in-reference flag) that pass the thresholds — what
        ``jellyfish dump -c`` would print; used by the tests."""
        return self.count(out_cap=max(self.stats["new"], 2), want_planes=True, **thresholds)

    def close(self):
        self.bins = None


# ── Module 1 ───────────────────────────────────────────────────────

def _extract_child_kmers_discovery(child_bam, ref_fasta, kmer_size, min_child_count,
                                   threads, tmpdir, jf_hash_size=None, engine=None):
    """Count every canonical k-mer of the child and threshold by count.

    Returns ``(ChildCandidates, n_candidates)``; the reference returns the
    path of a FASTA holding the same set (``dump -c -L min_child_count``).
    ``jf_hash_size`` (Jellyfish ``-s``) is accepted and used as the total slot
    budget of the table slices when given.
    """
    eng = get_engine(engine)
    if not eng.lib.kdf_key_words(kmer_size):
        raise _engine.KdfError(
            "k=%d is outside the GPU engine's range (k <= 63; 64-bit keys for k <= 32, "
            "128-bit above)" % kmer_size)
    t0 = time.monotonic()
    logger.info("Extracting child k-mers from BAM (k=%d)…", kmer_size)
    old = _CHILD_CACHE.pop(child_bam, None)
    if old is not None:
        old.close()
    cache = _CHILD_CACHE[child_bam] = ChildScanCache(child_bam, threads)
    # The layout of the count is planned before the file has been read, from its size: a BAM
    # holds at most ~4 bases per compressed byte.  Too few bases guessed → the bins overflow
    # and the pass is redone with exact sizes; too many → bins larger than needed.
    est = max(int(os.path.getsize(child_bam) * 4), 1 << 16)
    n_passes, n_local, slice_capacity = kmer_chain.plan_child_count(eng, est, 0, kmer_size, min_child_count)
    bin_cap = kmer_chain._bin_capacity(est / n_passes, n_local)
    pass0 = (n_passes.bit_length() - 1, 0) if n_passes > 1 else None
    bins = eng.new_bins(kmer_size, n_local, bin_cap)
    st = eng.new_stats()
    for b in cache.stream():        # decode + bin pass 0, batch by batch
        eng.bin_stream(bins, eng.upload(bamio.counting_view(b), with_reads=False), st, pass_=pass0)
    tot = {"reads": cache.reads, "bases": cache.bases}
    # now that the size is known: the real plan
    n_max = max(cache.bases, 1)
    real = kmer_chain.plan_child_count(eng, n_max, 0, kmer_size, min_child_count)
    keep0 = (not bins.overflowed()) and real[0] <= n_passes
    if keep0:
        # fewer hash ranges would have done, but these are already binned: keep the layout,
        # size the slices for the real number of k-mers
        slice_capacity = (max(1024, -(-max(n_max // kmer_chain.SLOTS_DIV, 1024) // (n_passes * n_local))) + 3) & ~3
    else:
        logger.info("  the size estimate was off (%d bases): re-binning with %d pass(es)", n_max, real[0])
        n_passes, n_local, slice_capacity = real
        bins = None
        bin_cap = kmer_chain._bin_capacity(n_max / n_passes, n_local)
    n_keys = kw._parse_hash_size(jf_hash_size)
    if n_keys:
        slice_capacity = max(slice_capacity, (n_keys // (n_passes * n_local) + 3) & ~3)
    cand = ChildCandidates(eng, kmer_size, cache, n_passes, n_local, slice_capacity, min_child_count, tot,
                           bins0=bins, bin_cap=bins.bin_cap if bins is not None else bin_cap)
    res = cand.count(min0=min_child_count)
    tot["windows"] = res["keys"]
    tot["new"] = res["distinct"]
    tot["hits"] = res["hits"]
    cand.n_candidates = n_candidates = res["n_count"]
    logger.info("Child k-mer counting complete (%.1fs): %d reads, %d k-mer instances, "
                "%d distinct (%d pass(es) x %d hash ranges x %d slots)", time.monotonic() - t0,
                tot["reads"], tot["windows"], tot["new"], n_passes, n_local, cand.slice_capacity)
    logger.info("Child candidate k-mers (count >= %d): %d", min_child_count, n_candidates)
    return cand, n_candidates


def _subtract_reference_kmers(ref_jf, child_candidates_fa, tmpdir):
    """Remove candidates that occur in the reference → ``(KmerSet, n_non_ref)``.

    The reference queries every candidate against ``ref.jf`` and keeps count
    == 0; here the reference k-mers are binned by the same hash ranges and
    marked inside the counting pass, and the survivors are emitted from the
    L2-resident slice.  The candidates handle is released afterwards, as the
    reference deletes its input FASTA."""
    cand = child_candidates_fa
    eng = cand.engine
    res = cand.count(ref_jf, out_cap=max(cand.n_candidates, 2), min0=cand.min_child_count, max1=0)
    k = cand.k
    cand.close()
    logger.info("Non-reference child k-mers after subtraction: %d", res["n_out"])
    return KmerSet(eng, k, res["lo"], res["hi"]), res["n_out"]


# ── Module 2 ───────────────────────────────────────────────────────

def _count_parent_jellyfish(parent_bam, ref_fasta, kmer_fasta, kmer_size, parent_dir, threads,
                            label="Parent", n_filter_kmers=None, engine=None):
    """``samtools fasta | jellyfish count --if`` → table whose plane 0 holds the
    parent's count of every filter k-mer (the reference returns a ``.jf`` path)."""
    eng = get_engine(engine)
    kset = kw._as_kmer_set(eng, kmer_fasta, kmer_size)
    table = kset.build_table(n_min=n_filter_kmers or 0)
    t0 = time.monotonic()
    table, tot = kw.count_bam_into_table(eng, parent_bam, table, _engine.MODE_COUNT_IF_PRESENT,
                                         0, threads)
    logger.info("  %s counting complete (%.1fs): %d reads, %d k-mer instances probed, %d hits",
                label, time.monotonic() - t0, tot["reads"], tot["windows"], tot["hits"])
    return table


def _filter_parents_discovery(mother_bam, father_bam, ref_fasta, child_non_ref_fa, kmer_size,
                              threads, tmpdir, parent_max_count=0, engine=None):
    """Keep k-mers seen at most ``parent_max_count`` times in the mother, then in
    the father → ``(n_proband_unique, KmerSet | None)``."""
    eng = get_engine(engine)
    kset = kw._as_kmer_set(eng, child_non_ref_fa, kmer_size)
    n_input = len(kset)
    if n_input == 0:
        return 0, None
    logger.info("Filtering %d non-reference k-mers against parents…", n_input)
    mt = _count_parent_jellyfish(mother_bam, ref_fasta, kset, kmer_size,
                                 os.path.join(tmpdir or "", "mother"), threads, "Mother",
                                 n_input, engine=eng)
    n_surv, lo, hi, _a, _b = eng.threshold_compact(mt, max0=parent_max_count)
    mt.close()
    logger.info("Mother: %d / %d non-ref k-mers found (count > %d), %d surviving",
                n_input - n_surv, n_input, parent_max_count, n_surv)
    if n_surv == 0:
        return 0, None
    after_mother = KmerSet(eng, kmer_size, lo, hi)
    ft = _count_parent_jellyfish(father_bam, ref_fasta, after_mother, kmer_size,
                                 os.path.join(tmpdir or "", "father"), threads, "Father",
                                 n_surv, engine=eng)
    n_pu, lo, hi, _a, _b = eng.threshold_compact(ft, max0=parent_max_count)
    ft.close()
    logger.info("Father: %d / %d surviving k-mers found (count > %d), %d proband-unique",
                n_surv - n_pu, n_surv, parent_max_count, n_pu)
    if n_pu == 0:
        return 0, None
    return n_pu, KmerSet(eng, kmer_size, lo, hi)


# ── Module 3 ───────────────────────────────────────────────────────

def _collect_kmer_ref_positions(read, kmer_hit_indices, kmer_size):
    """Counter{ref position: #hit k-mers covering it} (``core/bam_scanner.py:97-117``)."""
    cig = read.cigartuples or ()
    qlen = sum(ln for op, ln in cig if op in (0, 1, 4, 7, 8))
    q2r = np.full(max(qlen, 1), -1, dtype=np.int64)
    q = 0
    r = read.reference_start
    for op, ln in cig:
        if op in (0, 7, 8):
            q2r[q:q + ln] = np.arange(r, r + ln)
            q += ln
            r += ln
        elif op in (1, 4):
            q += ln
        elif op in (2, 3):
            r += ln
    cov = collections.Counter()
    if not len(kmer_hit_indices):
        return cov
    starts = np.fromiter(kmer_hit_indices, dtype=np.int64)
    qpos = (starts[:, None] + np.arange(kmer_size)[None, :]).ravel()
    qpos = qpos[qpos < q2r.shape[0]]
    rpos = q2r[qpos]
    rpos = rpos[rpos >= 0]
    u, c = np.unique(rpos, return_counts=True)
    for p, n in zip(u.tolist(), c.tolist()):
        cov[p] = n
    return cov


def _accumulate_coverage(eng, batch, reads, hit_slices, off, kmer_size, kmer_coverage, read_coverage):
    """Per-position k-mer and read coverage of the kept reads of one batch, on the
    device (``kdf_hit_coverage``, K7): what the reference computes read by read with
    ``_collect_kmer_ref_positions`` and merges with ``Counter.update``
    (``core/bam_scanner.py:97-117``, ``discovery/pipeline.py:851-855``)."""
    if not reads:
        return
    rr = np.asarray(reads, dtype=np.int64)
    c0 = batch.cigar_off[rr].astype(np.int64)
    c1 = batch.cigar_off[rr + 1].astype(np.int64)
    n_ops = c1 - c0
    cig_off = np.zeros(rr.shape[0] + 1, dtype=np.uint64)
    cig_off[1:] = np.cumsum(n_ops)
    if int(cig_off[-1]):
        idx = np.repeat(c0 - cig_off[:-1].astype(np.int64), n_ops) + np.arange(int(cig_off[-1]))
        cigar = batch.cigar_blob[idx]
    else:
        cigar = np.zeros(0, dtype=np.uint32)
    n_hits = np.asarray([b - a for a, b in hit_slices], dtype=np.int64)
    hit_read = np.repeat(np.arange(rr.shape[0], dtype=np.uint32), n_hits)
    hit_off = np.concatenate([off[a:b] for a, b in hit_slices]).astype(np.uint32) if n_hits.sum() else \
        np.zeros(0, dtype=np.uint32)
    contig, pos, kc, rc = eng.hit_coverage(hit_read, hit_off, kmer_size, batch.ref_id[rr],
                                           batch.pos[rr].astype(np.int64), cig_off, cigar)
    for c in np.unique(contig).tolist():
        sel = contig == c
        name = batch.ref_names[c]
        kcov, rcov = kmer_coverage[name], read_coverage[name]
        for p, a, b in zip(pos[sel].tolist(), kc[sel].tolist(), rc[sel].tolist()):
            kcov[p] += a
            rcov[p] += b


def scan_child_reads(eng, child_bam, table, kmer_size, min_distinct_kmers_per_read, threads,
                     batch_bases=kw.BATCH_BASES):
    """GPU part of Module 3: per-read distinct / hit counts for every scanned
    record plus, for reads meeting the threshold, their hit positions.

    Yields ``(batch, ndistinct u32[], nhits u32[], hit_read_idx, hit_offset)``
    per batch, hits sorted by (read, offset)."""
    cache = child_cache(child_bam)
    pf = None
    if cache is not None and cache.complete:
        source, owned = cache.batches, False          # decoded once, in Module 1
    else:
        pf = kw.take_prefetch(child_bam, bamio.MODE_SCAN, threads, True, batch_bases)
        source, owned = pf, True
    try:
        for batch in source:
            ds = eng.upload(batch)
            # sparse form: a streaming probe emits the (rare) hit windows, the device
            # reduces them per read; dense per-read arrays are rebuilt here for the
            # CPU cluster / anchor step
            sp = eng.scan_reads_sparse(table, ds)
            _b, nd, nh, ridx, off, slot = sparse_to_scan_item(batch, sp, min_distinct_kmers_per_read)
            yield batch, nd, nh, ridx, off, slot
            if owned:
                batch.close()
    finally:
        if pf is not None:
            pf.close()


def _cluster_hits(read_hits, merge_distance):
    """Sort by (chrom, start) and merge greedily (reference ``:1107-1144``)."""
    regions, region_reads, region_kmers = [], {}, {}
    if not read_hits:
        return regions, region_reads, region_kmers
    read_hits.sort(key=lambda h: (h[0], h[1]))
    cur = None
    for chrom, start, end, name, kmers, _supp in read_hits:
        if cur is not None and chrom == cur[0] and start <= cur[2] + merge_distance:
            cur[2] = max(cur[2], end)
            cur[3].add(name)
            cur[4].update(kmers)
            continue
        if cur is not None:
            key = (cur[0], cur[1], cur[2])
            regions.append(key)
            region_reads[key] = cur[3]
            region_kmers[key] = cur[4]
        cur = [chrom, start, end, {name}, set(kmers)]
    key = (cur[0], cur[1], cur[2])
    regions.append(key)
    region_reads[key] = cur[3]
    region_kmers[key] = cur[4]
    return regions, region_reads, region_kmers


class AnchorCollector:
    """The CPU half of Module 3 for the batches one process scans: informative reads
    de-duplicated on ``(query_name, is_supplementary)`` in file order, their region
    tuples, SV metadata and per-position coverage (K7 on the device).  ``preseen``: keys
    already claimed by reads that come EARLIER in the file (the shards of lower ranks in
    a multi-GPU run): such a read is a duplicate here exactly as it would be in a
    sequential pass."""

    def __init__(self, eng, kmer_size, min_distinct_kmers_per_read, note_hit=None, preseen=None):
        self.eng = eng
        self.k = kmer_size
        self.min_dk = max(1, min_distinct_kmers_per_read)
        self.note_hit = note_hit
        self.reads_seen = set(preseen) if preseen else set()
        self.read_hits = []
        self.read_sv_meta = {}
        self.kmer_coverage = collections.defaultdict(collections.Counter)
        self.read_coverage = collections.defaultdict(collections.Counter)
        self.unmapped_informative = 0
        self.total_scanned = 0
        self.per_read = []  # (record index, n_distinct, n_hits) of reads with >= 1 hit

    def add(self, batch, nd, nh, ridx, off):
        kmer_size = self.k
        self.total_scanned += batch.n_reads
        hit_reads = np.flatnonzero(nh > 0)
        for r in hit_reads.tolist():
            self.per_read.append((int(batch.rec_index[r]), int(nd[r]), int(nh[r])))
            if self.note_hit is not None:
                self.note_hit(batch, r)     # the informative-reads writer fetches these by offset
        informative = np.flatnonzero(nd >= self.min_dk)
        lo_i = np.searchsorted(ridx, informative, side="left")
        hi_i = np.searchsorted(ridx, informative, side="right")
        cov_reads, cov_hits = [], []   # kept reads of this batch and their hit slices (K7 input)
        for r, a, b in zip(informative.tolist(), lo_i.tolist(), hi_i.tolist()):
            read = batch.record(r)
            dedup_key = (read.query_name, read.is_supplementary)
            if dedup_key in self.reads_seen:
                continue
            self.reads_seen.add(dedup_key)
            if read.is_unmapped:
                self.unmapped_informative += 1
                continue
            hit_idx = off[a:b].tolist()
            seq = read.query_sequence
            unique_in_read = {canonicalize(seq[i:i + kmer_size]) for i in hit_idx}
            chrom = read.reference_name
            self.read_hits.append((chrom, read.reference_start, read.reference_end,
                                   dedup_key[0], unique_in_read, dedup_key[1]))
            cov_reads.append(r)
            cov_hits.append((a, b))
            max_clip = 0
            for op, ln in read.cigartuples or ():
                if op == 4 and ln > max_clip:
                    max_clip = ln
            has_sa = read.has_tag("SA")
            self.read_sv_meta[dedup_key] = {
                "has_sa": has_sa,
                "sa_str": read.get_tag("SA") if (has_sa and not dedup_key[1]) else None,
                "is_paired": read.is_paired,
                "is_proper_pair": read.is_proper_pair,
                "mate_is_unmapped": read.mate_is_unmapped if read.is_paired else False,
                "max_clip": max_clip,
            }
        _accumulate_coverage(self.eng, batch, cov_reads, cov_hits, off, kmer_size, self.kmer_coverage,
                             self.read_coverage)


def sparse_to_scan_item(batch, sp, min_distinct_kmers_per_read):
    """One batch's ``scan_reads_sparse`` result → the tuple ``scan_child_reads`` yields."""
    nd = np.zeros(batch.n_reads, dtype=np.uint32)
    nh = np.zeros(batch.n_reads, dtype=np.uint32)
    rr = sp["read"].astype(np.int64)
    nd[rr] = sp["ndistinct"]
    nh[rr] = sp["nhits"]
    thr = max(1, min_distinct_kmers_per_read)
    pos = sp["hit_pos"]
    slot = sp["hit_slot"]
    if pos.shape[0]:
        ridx = np.searchsorted(batch.read_starts, pos, side="right") - 1
        keep = nd[ridx] >= thr
        pos, slot, ridx = pos[keep], slot[keep], ridx[keep]
        off = (pos - batch.read_starts[ridx]).astype(np.int64)
    else:
        ridx = np.zeros(0, dtype=np.int64)
        off = np.zeros(0, dtype=np.int64)
        slot = np.zeros(0, dtype=np.uint32)
    return batch, nd, nh, ridx, off, slot


def _anchor_and_cluster(child_bam, ref_fasta, proband_unique_kmers, kmer_size,
                        merge_distance=500, threads=1, min_distinct_kmers_per_read=1,
                        proband_unique_fa=None, proband_jf=None, n_proband_unique=None,
                        tmpdir=None, memory_limit_gb=None, engine=None):
    """Find reads carrying proband-unique k-mers and cluster them into regions.

    ``proband_jf`` is the device membership table (or pass the set through
    ``proband_unique_kmers`` / ``proband_unique_fa``).  Returns the reference's
    8-tuple ``(regions, region_reads, total_informative, region_kmers,
    unmapped_informative, read_sv_meta, kmer_coverage, read_coverage)``.
    Informative reads are de-duplicated on ``(query_name, is_supplementary)``
    in BAM file order.
    """
    eng = get_engine(engine)
    table = proband_jf
    owns = False
    if table is None:
        if proband_unique_fa is not None:
            kset = kw._as_kmer_set(eng, proband_unique_fa, kmer_size)
        else:
            kset = KmerSet.from_strings(eng, kmer_size, proband_unique_kmers or ())
        table = kset.build_table()
        owns = True
    t0 = time.monotonic()
    cache = child_cache(child_bam)
    if cache is not None:
        cache.hit_reads = []
    col = AnchorCollector(eng, kmer_size, min_distinct_kmers_per_read,
                          note_hit=cache.note_hit if cache is not None else None)
    for item in scan_child_reads(eng, child_bam, table, kmer_size, min_distinct_kmers_per_read, threads):
        col.add(*item[:5])
        item[0].close()
    read_hits, read_sv_meta = col.read_hits, col.read_sv_meta
    kmer_coverage, read_coverage = col.kmer_coverage, col.read_coverage
    unmapped_informative, total_scanned, per_read = col.unmapped_informative, col.total_scanned, col.per_read
    if cache is not None:
        cache.scanned = True
        cache.release_batches()      # only the hit list and the reader are needed from here on
    if owns:
        table.close()
    total_informative = len(read_hits) + unmapped_informative
    logger.info("Anchoring complete: %d informative reads (%d mapped, %d unmapped) from %d "
                "scanned (%.1fs)", total_informative, len(read_hits), unmapped_informative,
                total_scanned, time.monotonic() - t0)
    _anchor_and_cluster.last_per_read = per_read
    if not read_hits:
        return ([], {}, total_informative, {}, unmapped_informative, read_sv_meta,
                kmer_coverage, read_coverage)
    regions, region_reads, region_kmers = _cluster_hits(read_hits, merge_distance)
    logger.info("Clustered %d mapped informative reads into %d regions", len(read_hits),
                len(regions))
    return (regions, region_reads, total_informative, region_kmers, unmapped_informative,
            read_sv_meta, kmer_coverage, read_coverage)


# ── Module 4: annotation + writers (CPU, output contract) ──────────

def _infer_sv_type(region_a, region_b):
    return "BND" if region_a[0] != region_b[0] else "INTRA"


def _annotate_and_link_from_metadata(regions, region_reads, read_sv_meta):
    """Per-region SV evidence and SA-tag links (reference ``:1351-1489``)."""
    membership = collections.defaultdict(set)
    for rk in regions:
        for qname in region_reads.get(rk, ()):
            membership[qname].add(rk)
    annotations = {rk: {"split_reads": 0, "discordant_pairs": 0, "max_clip_len": 0,
                        "unmapped_mates": 0} for rk in regions}
    if not membership:
        return annotations, []
    split_done = set()
    for (qname, _supp), meta in read_sv_meta.items():
        for rk in membership.get(qname, ()):
            ann = annotations[rk]
            if meta["has_sa"] and (qname, rk) not in split_done:
                split_done.add((qname, rk))
                ann["split_reads"] += 1
            if meta["is_paired"]:
                if meta["mate_is_unmapped"]:
                    ann["unmapped_mates"] += 1
                elif not meta["is_proper_pair"]:
                    ann["discordant_pairs"] += 1
            ann["max_clip_len"] = max(ann["max_clip_len"], meta["max_clip"])
    by_chrom = collections.defaultdict(list)
    for rk in regions:
        by_chrom[rk[0]].append(rk)
    starts = {}
    for chrom, lst in by_chrom.items():
        lst.sort(key=lambda x: x[1])
        starts[chrom] = [x[1] for x in lst]
    bridges = collections.defaultdict(set)
    for (qname, _supp), meta in read_sv_meta.items():
        sa = meta.get("sa_str")
        if not sa or qname not in membership:
            continue
        for entry in sa.rstrip(";").split(";"):
            fields = entry.split(",")
            if len(fields) < 3 or fields[0] not in starts:
                continue
            try:
                sa_pos = int(fields[1]) - 1
            except ValueError:
                continue
            j = bisect.bisect_right(starts[fields[0]], sa_pos) - 1
            if j < 0:
                continue
            target = by_chrom[fields[0]][j]
            if not (target[1] <= sa_pos < target[2]):
                continue
            for prim in membership[qname]:
                if prim != target:
                    bridges[tuple(sorted((prim, target)))].add(qname)
    for qname, rset in membership.items():
        if len(rset) > 1:
            ordered = sorted(rset)
            for i, a in enumerate(ordered):
                for b in ordered[i + 1:]:
                    bridges[(a, b)].add(qname)
    links = [{"region_a": a, "region_b": b, "supporting_reads": bridges[(a, b)],
              "sv_type_hint": _infer_sv_type(a, b)} for a, b in sorted(bridges)]
    return annotations, links


def _classify_regions(regions, region_annotations, sv_links):
    """SV / SMALL / AMBIGUOUS (reference ``:1517-1546``)."""
    linked = {l["region_a"] for l in sv_links} | {l["region_b"] for l in sv_links}
    for rk in regions:
        ann = region_annotations.get(rk, {})
        evid = [ann.get("split_reads", 0), ann.get("discordant_pairs", 0),
                ann.get("unmapped_mates", 0)]
        if max(evid) >= 2 or rk in linked:
            ann["class"] = "SV"
        elif not any(evid):
            ann["class"] = "SMALL"
        else:
            ann["class"] = "AMBIGUOUS"
        region_annotations[rk] = ann


def _write_bed(regions, region_reads, region_kmers, bed_path, region_annotations=None,
               filters=None):
    region_annotations = region_annotations or {}
    with open(bed_path, "w") as fh:
        if filters:
            fh.write("#filters: %s\n" % " ".join("%s=%s" % kv for kv in sorted(filters.items())))
        fh.write("#chrom\tstart\tend\treads\tunique_kmers\tsplit_reads\tdiscordant_pairs"
                 "\tmax_clip_len\tunmapped_mates\tclass\n")
        for rk in regions:
            ann = region_annotations.get(rk, {})
            cols = [rk[0], rk[1], rk[2], len(region_reads.get(rk, ())),
                    len(region_kmers.get(rk, ())), ann.get("split_reads", 0),
                    ann.get("discordant_pairs", 0), ann.get("max_clip_len", 0),
                    ann.get("unmapped_mates", 0), ann.get("class", "SMALL")]
            fh.write("\t".join(str(c) for c in cols) + "\n")
    logger.info("BED file written: %s (%d regions)", bed_path, len(regions))


def _runs(sorted_items):
    """Merge (pos, value) pairs, sorted by pos, into (start, end, value) runs."""
    run = None
    for pos, val in sorted_items:
        if run is not None and pos == run[1] and val == run[2]:
            run[1] = pos + 1
            continue
        if run is not None:
            yield tuple(run)
        run = [pos, pos + 1, val]
    if run is not None:
        yield tuple(run)


def _write_bedgraph(kmer_coverage, bedgraph_path, read_coverage=None, min_reads=3):
    """4-column bedGraph of k-mer coverage (reference ``:1197-1278``)."""
    with open(bedgraph_path, "w") as fh:
        fh.write("#track type=bedGraph description=\"De novo k-mer coverage (unique k-mer base "
                 "overlaps per position, min_reads>=%d)\"\n" % min_reads)
        for chrom in sorted(kmer_coverage):
            cov = kmer_coverage[chrom]
            if not cov:
                continue
            rc = read_coverage.get(chrom, {}) if read_coverage else None
            kept = [(p, cov[p]) for p in sorted(cov)
                    if rc is None or rc.get(p, 0) >= min_reads]
            for s, e, v in _runs(kept):
                fh.write("%s\t%d\t%d\t%s\n" % (chrom, s, e, v))


def _write_read_coverage_bed(kmer_coverage, read_coverage, bed_path, min_reads=3):
    """Per-position read support BED (reference ``:1281-1348``)."""
    with open(bed_path, "w") as fh:
        fh.write("#track description=\"De novo k-mer read support (min_reads>=%d)\"\n"
                 "#chrom\tstart\tend\tread_count\tavg_kmers_per_read\n" % min_reads)
        for chrom in sorted(read_coverage):
            rc = read_coverage[chrom]
            kc = kmer_coverage.get(chrom, {})
            kept = [(p, (n, round(kc.get(p, 0) / n, 1)))
                    for p, n in sorted(rc.items()) if n >= min_reads]
            for s, e, (n, avg) in _runs(kept):
                fh.write("%s\t%d\t%d\t%s\t%s\n" % (chrom, s, e, n, avg))


def _write_bedpe(links, bedpe_path):
    with open(bedpe_path, "w") as fh:
        fh.write("#chrom1\tstart1\tend1\tchrom2\tstart2\tend2\tsv_id\tsupporting_reads\tsv_type\n")
        for i, link in enumerate(links, 1):
            a, b = link["region_a"], link["region_b"]
            fh.write("%s\t%d\t%d\t%s\t%d\t%d\tSV_%d\t%d\t%s\n" % (
                a[0], a[1], a[2], b[0], b[1], b[2], i, len(link["supporting_reads"]),
                link["sv_type_hint"]))


def _parse_candidate_summary(summary_path, dka_dkt_min=0.25, dka_min=10):
    """High-quality candidates of a VCF-mode summary.txt (reference ``:1549-1606``)."""
    out = []
    in_table = False
    with open(summary_path) as fh:
        for raw in fh:
            text = raw.strip()
            if not in_table:
                in_table = text.startswith("Variant") and "DKU" in text
                continue
            if text.startswith("-------"):
                continue
            if not text or text.startswith("="):
                break
            parts = text.split()
            if len(parts) < 12:
                continue
            chrom, pos = parts[0].rsplit(":", 1)
            ref, alt = parts[1].split(">")
            dka, dka_dkt = int(parts[4]), float(parts[6])
            if dka_dkt > dka_dkt_min and dka > dka_min:
                out.append({"chrom": chrom, "pos": int(pos), "ref": ref, "alt": alt,
                            "dka": dka, "dka_dkt": dka_dkt, "call": parts[-1]})
    return out


def _compare_candidates_to_regions(candidates, regions):
    res = []
    for cand in candidates:
        hit = next(((c, s, e) for c, s, e in regions
                    if c == cand["chrom"] and s < cand["pos"] <= e), None)
        res.append(dict(cand, captured=hit is not None,
                        region=("%s:%d-%d" % (hit[0], hit[1] + 1, hit[2])) if hit else None))
    return res


#: curated DNM loci the reference reports on (Sulovari et al. 2023; reference ``:1641-1649``)
SULOVARI_DNM_REGIONS = [
    ("chr17", 53340465, 107, "deletion"),
    ("chr14", 23280711, None, "microsatellite_expansion"),
    ("chr3", 85552367, 64, "sv_like"),
    ("chr5", 97089276, 43, "sv_like"),
    ("chr8", 125785998, 43, "sv_like"),
    ("chr18", 62805217, 34, "sv_like"),
    ("chr7", 142786222, 10607, "deletion"),
]


def _evaluate_dnm_regions(discovery_regions, region_detail, dnm_regions=None):
    """Per-locus detection summary (reference ``:1653-1783``)."""
    dnm_regions = SULOVARI_DNM_REGIONS if dnm_regions is None else dnm_regions
    detail = {(d["chrom"], d["start"], d["end"]): d for d in region_detail}
    rank = {"SV": 3, "AMBIGUOUS": 2, "SMALL": 1}
    out = []
    for chrom, pos, size, event_type in dnm_regions:
        lo, hi = pos, pos + (size if size else 1)
        matches = [r for r in discovery_regions if r[0] == chrom and r[1] < hi and lo < r[2]]
        ds = [detail.get(m, {}) for m in matches]
        span_lo = min([lo] + [m[1] for m in matches])
        span_hi = max([hi] + [m[2] for m in matches])
        kmers = sum(d.get("unique_kmers", 0) for d in ds)
        classes = [d.get("class", "SMALL") for d in ds]
        out.append({
            "locus": "%s:%d" % (chrom, pos),
            "event_type": event_type,
            "event_size": size,
            "detected": bool(matches),
            "discovery_regions": ["%s:%d-%d" % (m[0], m[1] + 1, m[2]) for m in matches],
            "total_reads": sum(d.get("reads", 0) for d in ds),
            "total_unique_kmers": kmers,
            "max_clip_len": max([0] + [d.get("max_clip_len", 0) for d in ds]),
            "unmapped_mates": sum(d.get("unmapped_mates", 0) for d in ds),
            "discordant_pairs": sum(d.get("discordant_pairs", 0) for d in ds),
            "split_reads": sum(d.get("split_reads", 0) for d in ds),
            "sv_class": max(classes, key=lambda c: rank.get(c, 0)) if classes else "NONE",
            "kmer_signal": round(kmers / max(span_hi - span_lo, 1), 4) if matches else 0.0,
            "assessment": "DETECTED" if matches else "NOT_DETECTED",
        })
    return out


def _write_informative_reads_discovery(child_bam, ref_fasta, proband_unique_kmers_or_path, kmer_size,
                                       output_bam, engine=None, threads=4):
    """Child reads carrying >= 1 proband-unique k-mer → coordinate-sorted, indexed
    BAM, each read tagged ``dk:i:1`` (reference ``:1979-2079``): primary and
    supplementary, non-duplicate, mapped or not, first record per
    ``(query_name, is_supplementary)`` in file order.  ``proband_unique_kmers_or_path``
    is the device membership table (or a KmerSet)."""
    eng = get_engine(engine)
    table = proband_unique_kmers_or_path
    owns = False
    if isinstance(table, KmerSet):
        table = table.build_table()
        owns = True
    records = []
    written = set()
    cache = child_cache(child_bam)
    if cache is not None and getattr(cache, "scanned", False):
        # the anchoring scan has already seen every read with a hit: fetch those records
        # back by their offsets (a few BGZF blocks) instead of decoding the file again
        rd = cache.reader
        uoffs = []
        for uoff, qname, supp in cache.hit_reads:      # file order
            if (qname, supp) in written:
                continue
            written.add((qname, supp))
            uoffs.append(uoff)
        records = [bamio.append_int_tag(raw, "dk", 1) for raw in rd.fetch_records(uoffs)]
        if owns:
            table.close()
        n = bamio.write_sorted_bam(output_bam, rd.header_text, rd.references, rd.lengths, records)
        logger.info("Informative reads BAM written: %s (%d reads)", output_bam, n)
        return n
    with bamio.BamReader(child_bam, threads=threads) as rd:
        header_text, names, lens = rd.header_text, rd.references, rd.lengths
        for batch in rd.batches(bamio.MODE_SCAN, max_bases=kw.BATCH_BASES, want_meta=3):
            sp = eng.scan_reads_sparse(table, eng.upload(batch))
            for i in sp["read"].astype(np.int64).tolist():
                rec = batch.record(i)
                key = (rec.query_name, rec.is_supplementary)
                if key in written:
                    continue
                written.add(key)
                raw = batch.raw_blob[int(batch.raw_off[i]):int(batch.raw_off[i + 1])].tobytes()
                records.append(bamio.append_int_tag(raw, "dk", 1))
            batch.close()
    if owns:
        table.close()
    n = bamio.write_sorted_bam(output_bam, header_text, names, lens, records)
    logger.info("Informative reads BAM written: %s (%d reads)", output_bam, n)
    return n


def _write_empty_discovery_outputs(bed_path, metrics_path, summary_path, metrics,
                                   bedpe_path=None):
    _write_bed([], {}, {}, bed_path)
    if bedpe_path:
        _write_bedpe([], bedpe_path)
    with open(metrics_path, "w") as fh:
        json.dump(metrics, fh, indent=2)
    from .summary import _write_discovery_summary
    _write_discovery_summary(summary_path, [], {}, {}, metrics)


def _validate(args):
    """Input rules of the reference (``utils.py:230-350``) that concern this path."""
    errs = []
    k = args.kmer_size
    if k < 3 or k % 2 == 0 or k > 201:
        errs.append("--kmer-size must be an odd number in [3, 201], got %d" % k)
    elif k > 63:
        errs.append("--kmer-size %d: the GPU engine supports k <= 63 (128-bit keys)" % k)
    for name in ("child", "mother", "father"):
        p = getattr(args, name)
        if not os.path.isfile(p):
            errs.append("--%s file not found: %s" % (name, p))
    if not args.ref_fasta and not getattr(args, "ref_jf", None):
        errs.append("discovery mode needs --ref-fasta or --ref-jf")
    if args.ref_fasta and not os.path.isfile(args.ref_fasta):
        errs.append("--ref-fasta file not found: %s" % args.ref_fasta)
    if errs:
        for e in errs:
            logger.error(e)
        sys.exit(1)


def run_discovery_pipeline(args, engine=None):
    """Run the VCF-free discovery pipeline (reference ``:2093``).  Returns the
    metrics dict that is also written to ``{out_prefix}.metrics.json``."""
    start = time.monotonic()
    logging.basicConfig(level=logging.DEBUG if getattr(args, "debug_kmers", False) else logging.INFO,
                        format="%(asctime)s %(levelname)s %(message)s")
    _validate(args)
    eng = get_engine(engine)
    out_prefix = args.out_prefix
    bed_path = out_prefix + ".bed"
    info_bam_path = out_prefix + ".informative.bam"
    metrics_path = out_prefix + ".metrics.json"
    summary_path = out_prefix + ".summary.txt"
    bedpe_path = getattr(args, "sv_bedpe", None) or out_prefix + ".sv.bedpe"
    bedgraph_path = out_prefix + ".kmer_coverage.bedgraph"
    read_cov_bed_path = out_prefix + ".read_coverage.bed"
    min_bedgraph_reads = getattr(args, "min_bedgraph_reads", 3)
    min_dk = getattr(args, "min_distinct_kmers_per_read", None)
    if min_dk is None:
        min_dk = max(1, args.kmer_size // 4)
    k = args.kmer_size
    threads = max(1, args.threads)

    def finish_empty(n_cand, n_non_ref):
        m = {"mode": "discovery", "child_candidate_kmers": n_cand, "non_ref_kmers": n_non_ref,
             "proband_unique_kmers": 0, "informative_reads": 0,
             "unmapped_informative_reads": 0, "candidate_regions": 0}
        _write_empty_discovery_outputs(bed_path, metrics_path, summary_path, m, bedpe_path)
        return m

    kw.reset_times()
    timings = LAST_TIMINGS
    timings.clear()
    t_stage = time.perf_counter()

    def lap(name):
        nonlocal t_stage
        now = time.perf_counter()
        timings[name] = timings.get(name, 0.0) + (now - t_stage)
        t_stage = now

    # one process per GPU under torch.distributed: every rank takes a range of each BAM
    world = 1
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size()
    except ImportError:
        pass
    if world > 1:
        from . import pipeline_dist
        paths = {"bed": bed_path, "metrics": metrics_path, "summary": summary_path, "bedpe": bedpe_path,
                 "bedgraph": bedgraph_path, "read_cov": read_cov_bed_path, "info_bam": info_bam_path}
        try:
            return pipeline_dist.run_discovery_pipeline_dist(args, eng, paths, min_dk, min_bedgraph_reads,
                                                             finish_empty, lap)
        finally:
            timings["total_s"] = time.monotonic() - start
    # the parents are decoded in the background while the child is read and counted (bounded
    # look-ahead: two batches each)
    # ... and the child's decode (what the first GPU stage waits for) starts before the
    # reference index is built
    if os.environ.get("KDF_PREFETCH_PARENTS", "1") != "0":
        kw.start_prefetch(args.child, bamio.MODE_SCAN, threads, True)
        kw.start_prefetch(args.mother, bamio.MODE_FASTA, threads)
        kw.start_prefetch(args.father, bamio.MODE_FASTA, threads)
    try:
        return _run_discovery(args, eng, k, threads, min_dk, min_bedgraph_reads, bed_path, info_bam_path,
                              metrics_path, summary_path, bedpe_path, bedgraph_path, read_cov_bed_path,
                              finish_empty, lap, start)
    finally:
        kw.drop_prefetch()
        drop_child_caches()
        timings["decode_threads_s"] = kw.TIMES["decode_s"]
        timings["decode_wait_s"] = kw.TIMES["decode_wait_s"]
        timings["total_s"] = time.monotonic() - start


def _run_discovery(args, eng, k, threads, min_dk, min_bedgraph_reads, bed_path, info_bam_path,
                   metrics_path, summary_path, bedpe_path, bedgraph_path, read_cov_bed_path,
                   finish_empty, lap, start):
    out_prefix = args.out_prefix
    ref_index = _ensure_ref_jf(args.ref_fasta, k, threads, getattr(args, "ref_jf", None), eng)
    lap("reference_index_s")
    cand, n_candidates = _extract_child_kmers_discovery(
        args.child, args.ref_fasta, k, args.min_child_count, threads, None,
        jf_hash_size=getattr(args, "jf_hash_size", None), engine=eng)
    lap("child_decode_and_count_s")
    if n_candidates == 0:
        cand.close()
        return finish_empty(0, 0)
    non_ref, n_non_ref = _subtract_reference_kmers(ref_index, cand, None)
    lap("reference_subtraction_s")
    if n_non_ref == 0:
        return finish_empty(n_candidates, 0)
    n_pu, pu = _filter_parents_discovery(args.mother, args.father, args.ref_fasta, non_ref, k,
                                         threads, None, args.parent_max_count, engine=eng)
    lap("parents_decode_and_filter_s")
    if n_pu == 0:
        return finish_empty(n_candidates, n_non_ref)
    pu_table = _build_proband_jf_index(pu, k, None, n_pu, engine=eng)
    (regions, region_reads, total_informative, region_kmers, unmapped_informative,
     read_sv_meta, kmer_coverage, read_coverage) = _anchor_and_cluster(
        args.child, args.ref_fasta, None, k, merge_distance=args.cluster_distance,
        threads=threads, min_distinct_kmers_per_read=min_dk, proband_jf=pu_table,
        n_proband_unique=n_pu, engine=eng)
    lap("anchor_scan_and_cluster_s")
    logger.info("[Module 4] Writing informative reads BAM: %s", info_bam_path)
    _write_informative_reads_discovery(args.child, getattr(args, "ref_fasta", None), pu_table, k,
                                       info_bam_path, engine=eng, threads=threads)
    pu_table.close()
    lap("informative_bam_s")

    paths = {"bed": bed_path, "metrics": metrics_path, "summary": summary_path, "bedpe": bedpe_path,
             "bedgraph": bedgraph_path, "read_cov": read_cov_bed_path, "info_bam": info_bam_path}
    return _finish_discovery(args, regions, region_reads, region_kmers, read_sv_meta, kmer_coverage,
                             read_coverage, total_informative, unmapped_informative, n_candidates,
                             n_non_ref, n_pu, min_dk, min_bedgraph_reads, paths, lap, start)


def _finish_discovery(args, regions, region_reads, region_kmers, read_sv_meta, kmer_coverage,
                      read_coverage, total_informative, unmapped_informative, n_candidates, n_non_ref,
                      n_pu, min_dk, min_bedgraph_reads, paths, lap, start=None):
    """Module 4: filter, annotate, classify and write every output file (reference
    ``discovery/pipeline.py:2330-2548``) from the anchored reads."""
    bed_path, metrics_path, summary_path = paths["bed"], paths["metrics"], paths["summary"]
    bedpe_path, bedgraph_path, read_cov_bed_path = paths["bedpe"], paths["bedgraph"], paths["read_cov"]
    if start is None:
        start = time.monotonic()
    min_reads, min_kmers = args.min_supporting_reads, args.min_distinct_kmers
    if min_reads > 1 or min_kmers > 1:
        regions = [r for r in regions if len(region_reads.get(r, ())) >= min_reads
                   and len(region_kmers.get(r, ())) >= min_kmers]
    annotations, links = _annotate_and_link_from_metadata(regions, region_reads, read_sv_meta)
    _classify_regions(regions, annotations, links)
    _write_bed(regions, region_reads, region_kmers, bed_path, region_annotations=annotations,
               filters={"min_distinct_kmers_per_read": min_dk, "min_supporting_reads": min_reads,
                        "min_distinct_kmers": min_kmers})
    _write_bedgraph(kmer_coverage, bedgraph_path, read_coverage, min_bedgraph_reads)
    _write_read_coverage_bed(kmer_coverage, read_coverage, read_cov_bed_path, min_bedgraph_reads)
    _write_bedpe(links, bedpe_path)

    metrics = {
        "mode": "discovery",
        "child_candidate_kmers": n_candidates,
        "non_ref_kmers": n_non_ref,
        "proband_unique_kmers": n_pu,
        "informative_reads": total_informative,
        "unmapped_informative_reads": unmapped_informative,
        "candidate_regions": len(regions),
        "filters": {"min_distinct_kmers_per_read": min_dk, "min_supporting_reads": min_reads,
                    "min_distinct_kmers": min_kmers, "min_bedgraph_reads": min_bedgraph_reads},
        "regions": [],
    }
    for rk in regions:
        ann = annotations.get(rk, {})
        metrics["regions"].append({
            "chrom": rk[0], "start": rk[1], "end": rk[2], "size": rk[2] - rk[1],
            "reads": len(region_reads.get(rk, ())),
            "unique_kmers": len(region_kmers.get(rk, ())),
            "split_reads": ann.get("split_reads", 0),
            "discordant_pairs": ann.get("discordant_pairs", 0),
            "max_clip_len": ann.get("max_clip_len", 0),
            "unmapped_mates": ann.get("unmapped_mates", 0),
            "class": ann.get("class", "SMALL"),
        })
    comparison = None
    cs = getattr(args, "candidate_summary", None)
    if cs and os.path.isfile(cs):
        comparison = _compare_candidates_to_regions(_parse_candidate_summary(cs), regions)
        n_cap = sum(1 for c in comparison if c["captured"])
        metrics["candidate_comparison"] = {
            "hq_candidates": len(comparison), "captured": n_cap,
            "capture_rate": (n_cap / len(comparison)) if comparison else 0.0,
            "candidates": [{"variant": "%s:%d %s>%s" % (c["chrom"], c["pos"], c["ref"], c["alt"]),
                            "dka": c["dka"], "dka_dkt": c["dka_dkt"], "captured": c["captured"],
                            "region": c["region"]} for c in comparison],
        }
    dnm = _evaluate_dnm_regions(regions, metrics["regions"])
    n_det = sum(1 for d in dnm if d["detected"])
    metrics["dnm_evaluation"] = {"total_loci": len(dnm), "detected": n_det,
                                 "detection_rate": (n_det / len(dnm)) if dnm else 0.0,
                                 "loci": dnm}
    with open(metrics_path, "w") as fh:
        json.dump(metrics, fh, indent=2)
    from .summary import _write_discovery_summary
    _write_discovery_summary(summary_path, regions, region_reads, region_kmers, metrics,
                             candidate_comparison=comparison, region_annotations=annotations,
                             dnm_evaluation=dnm)
    lap("annotate_and_write_s")
    logger.info("Discovery pipeline finished in %.1fs", time.monotonic() - start)
    return metrics
