"""GPU-engine wrappers — same names and contracts as the reference's
``core/jellyfish_wrappers.py``, with the Jellyfish / samtools subprocesses
replaced by calls into ``libkdf_sm100.so``.

Where the reference passes k-mer sets between stages as FASTA paths and
``.jf`` index paths, these wrappers pass in-memory handles
(:class:`~kmer_denovo_filter_b200.kmer_utils.KmerSet`, :class:`RefIndex`,
``KmerTable``); every wrapper also accepts the reference's path arguments
(FASTA of k-mers, Jellyfish ``binary/sorted`` file) and converts them.
"""

import json
import logging
import os

import numpy as np

from .. import bamio
from .. import engine as _engine
from ..kmer_utils import KmerSet, kmer_of

logger = logging.getLogger(__name__)

# bases per host→device batch while streaming a BAM (≈ 0.4 GB of packed input)
BATCH_BASES = 1 << 30

_default_engine = None


def get_engine(engine=None):
    """The process-wide :class:`CudaEngine` (one per GPU / process)."""
    global _default_engine
    if engine is not None:
        return engine
    if _default_engine is None:
        _default_engine = _engine.CudaEngine()
    return _default_engine


def set_engine(engine):
    global _default_engine
    _default_engine = engine


# ---------------------------------------------------------------------------
# sizing (reference: _estimate_jf_hash_size :73-107, max(2n, 10M) :155)
# ---------------------------------------------------------------------------

def _parse_hash_size(text):
    """'2G' / '500M' / '1000' → entries (Jellyfish ``-s`` syntax)."""
    if text is None:
        return None
    t = str(text).strip().upper()
    mult = 1
    if t and t[-1] in "KMG":
        mult = {"K": 1000, "M": 1000_000, "G": 1000_000_000}[t[-1]]
        t = t[:-1]
    return int(float(t) * mult)


# ---------------------------------------------------------------------------
# decode ahead of the GPU
# ---------------------------------------------------------------------------

# wall-clock accounting of the last pipeline run (seconds): time the decoder threads
# spent inside kdf_bam_next_batch, and time the consumer waited for a batch
TIMES = {"decode_s": 0.0, "decode_wait_s": 0.0}


def reset_times():
    for key in list(TIMES):
        TIMES[key] = 0.0


class BamPrefetcher:
    """Decode a BAM on a background thread, ``depth`` batches ahead of the consumer
    (the C++ decoder runs without the GIL, so batch i + 1 is inflated and packed while
    batch i is uploaded and processed on the GPU; started early, a parent is decoded
    while the child is still being counted).  Iterate it like ``BamReader.batches``;
    ``reader`` stays open until :meth:`close` (``fetch_records`` needs it)."""

    def __init__(self, path, mode, threads, want_meta=False, batch_bases=BATCH_BASES, depth=2):
        import queue
        import threading
        self.path = path
        self.reader = bamio.BamReader(path, threads=threads)
        self._q = queue.Queue(maxsize=max(1, depth))
        self._err = None
        self._stop = False

        def work():
            import time
            try:
                while not self._stop:
                    t0 = time.perf_counter()
                    b = self.reader.next_batch(mode, batch_bases, want_meta)
                    TIMES["decode_s"] += time.perf_counter() - t0
                    last = b.at_eof
                    if b.n_reads:
                        self._q.put(b)
                    else:
                        b.close()
                    if last:
                        break
            except BaseException as e:      # surfaced in the consumer
                self._err = e
            self._q.put(None)
        self._t = threading.Thread(target=work, name="kdf-decode", daemon=True)
        self._t.start()

    def __iter__(self):
        import time
        while True:
            t0 = time.perf_counter()
            b = self._q.get()
            TIMES["decode_wait_s"] += time.perf_counter() - t0
            if b is None:
                if self._err is not None:
                    raise self._err
                return
            yield b

    def close(self):
        self._stop = True
        try:
            while self._t.is_alive():
                try:
                    b = self._q.get(timeout=0.05)
                    if b is not None:
                        b.close()
                except Exception:
                    pass
        finally:
            self.reader.close()


# prefetchers started ahead of their consumer, by BAM path (run_discovery_pipeline starts the
# parents' decode before the child has been counted)
_PREFETCH = {}


def start_prefetch(path, mode, threads, want_meta=False, batch_bases=BATCH_BASES, depth=2):
    pf = _PREFETCH.get((path, mode))
    if pf is None:
        pf = _PREFETCH[(path, mode)] = BamPrefetcher(path, mode, threads, want_meta, batch_bases, depth)
    return pf


def take_prefetch(path, mode, threads, want_meta=False, batch_bases=BATCH_BASES):
    """The prefetcher started for (path, mode), or a new one."""
    pf = _PREFETCH.pop((path, mode), None)
    if pf is None:
        pf = BamPrefetcher(path, mode, threads, want_meta, batch_bases)
    return pf


def drop_prefetch():
    for key in list(_PREFETCH):
        _PREFETCH.pop(key).close()


def count_bam_into_table(eng, bam_path, table, mode, plane, threads, batch_bases=BATCH_BASES):
    """Stream ``samtools fasta -F 0xD00``-equivalent reads of a BAM through K1+K2
    against ``table`` (filtered parent counts: the table is primed with the filter
    set and never grows).  Returns ``(table, stats dict)``."""
    from ..discovery import kmer_chain as _kmer_chain   # (discovery imports this module)
    total = {"windows": 0, "hits": 0, "new": 0, "reads": 0, "bases": 0}
    pf = take_prefetch(bam_path, bamio.MODE_FASTA, threads, False, batch_bases)
    try:
        for batch in pf:
            ds = eng.upload(batch, with_reads=False)
            st = eng.new_stats()
            if mode == _engine.MODE_COUNT_IF_PRESENT:
                # direct probe (behind the table's filter when it has one), or — for a filter
                # set far larger than L2 — bin the batch by hash range and apply the bins
                _kmer_chain.count_if_present(eng, table, ds, st, plane)
            else:
                eng.count_stream(table, ds, mode, plane, 1, st)
            s = eng.read_stats(st)
            if s["full"]:
                raise _engine.KdfError(
                    "k-mer table full while counting %s (capacity %d slots); "
                    "pass a larger --jf-hash-size" % (bam_path, table.capacity))
            for key in ("windows", "hits", "new"):
                total[key] += s[key]
            total["reads"] += batch.n_reads
            total["bases"] += batch.n_bases
            batch.close()
    finally:
        pf.close()
    return table, total


# ---------------------------------------------------------------------------
# reference index (reference: _ensure_ref_jf :286-332)
# ---------------------------------------------------------------------------

class RefIndex:
    """The reference k-mer set, held as what it is cheapest to stream against
    the child table: either the packed reference sequence (built from the
    FASTA) or an explicit key list (parsed from a Jellyfish ``.jf`` file)."""

    def __init__(self, k, host_stream=None, keys=None, source=None):
        self.k = k
        self.host_stream = host_stream
        self.keys = keys          # (lo u64, hi u64) numpy or None
        self.source = source

    def to_bins(self, eng, n_parts, pass_=None):
        """The reference's canonical k-mers in ``n_parts`` hash-range bins (the input
        of ``kdf_count_bins``; with ``pass_`` only those of one group of a multi-pass
        count), with an exact-size retry if a range is skewed."""
        from ..discovery import kmer_chain
        if self.host_stream is not None:
            n_max = self.host_stream.n_bases
        else:
            n_max = int(self.keys[0].shape[0])
        groups = (1 << pass_[0]) if pass_ else 1
        cap = kmer_chain._bin_capacity(max(n_max, 1) / groups, n_parts)
        while True:
            bins = eng.new_bins(self.k, n_parts, cap)
            if self.host_stream is not None:
                eng.bin_stream(bins, eng.upload(self.host_stream, with_reads=False), pass_=pass_)
            else:
                lo, hi = eng.keys_to_device(self.keys, bins.key_words)
                eng.bin_keys(bins, lo, hi, pass_=pass_)
            if not bins.overflowed():
                return bins
            cap = int(bins.counts().max()) + 4


def read_jf_binary_sorted(path):
    """Parse a Jellyfish ``binary/sorted`` index → ``(k, lo u64, hi u64, counts)``.

    Format (as found in the reference fixture ``mini_ref.fa.k31.jf``): 9 ASCII
    digits = header length, JSON header with ``key_len`` (bits) and
    ``counter_len`` (bytes), then fixed-width little-endian records."""
    with open(path, "rb") as fh:
        data = fh.read()
    hlen = int(data[:9].decode())
    hdr = json.loads(data[9:9 + hlen].decode().rstrip("\0"))
    fmt = hdr.get("format", "binary/sorted")
    if fmt != "binary/sorted":
        raise _engine.KdfError("unsupported Jellyfish format %r in %s" % (fmt, path))
    kbits = int(hdr["key_len"])
    cbytes = int(hdr["counter_len"])
    kbytes = (kbits + 7) // 8
    rec = kbytes + cbytes
    body = np.frombuffer(data, dtype=np.uint8, offset=9 + hlen)
    n = body.shape[0] // rec
    body = body[:n * rec].reshape(n, rec)
    keyb = np.zeros((n, 16), dtype=np.uint8)
    keyb[:, :kbytes] = body[:, :kbytes]
    words = keyb.view("<u8")
    cb = np.zeros((n, 8), dtype=np.uint8)
    cb[:, :cbytes] = body[:, kbytes:]
    return kbits // 2, words[:, 0].copy(), words[:, 1].copy(), cb.view("<u8")[:, 0].copy()


def _ensure_ref_jf(ref_fasta, kmer_size, threads, ref_jf=None, engine=None):
    """Reference k-mer index.  An existing ``ref_jf`` / ``{ref_fasta}.k{k}.jf``
    Jellyfish file is reused (parsed); otherwise the FASTA is packed.  Nothing is
    written next to the FASTA (the reference caches a ``.jf`` there)."""
    if ref_jf is None and ref_fasta:
        cand = "%s.k%d.jf" % (ref_fasta, kmer_size)
        if os.path.isfile(cand):
            ref_jf = cand
    if ref_jf and os.path.isfile(ref_jf) and not ref_fasta:
        k, lo, hi, _c = read_jf_binary_sorted(ref_jf)
        if k != kmer_size:
            raise _engine.KdfError("reference index %s has k=%d, expected %d" % (ref_jf, k, kmer_size))
        logger.info("Reference Jellyfish index found: %s (%d k-mers)", ref_jf, lo.shape[0])
        return RefIndex(kmer_size, keys=(lo, hi), source=ref_jf)
    if not ref_fasta:
        raise _engine.KdfError("a reference FASTA or a Jellyfish reference index is required")
    if ref_jf and os.path.isfile(ref_jf):
        # same k-mer set either way; the FASTA is 12x smaller than the index's records and is
        # extracted on the device, so it wins when both are given (the reference prefers the .jf)
        logger.info("Reference index %s ignored: the k-mers are taken from %s", ref_jf, ref_fasta)
    logger.info("Packing reference FASTA: %s (k=%d)", ref_fasta, kmer_size)
    hs, _n_records = _engine.pack_fasta_file(ref_fasta, threads)   # (all host threads, in the library)
    return RefIndex(kmer_size, host_stream=hs, source=ref_fasta)


# ---------------------------------------------------------------------------
# parent scans
# ---------------------------------------------------------------------------

def _as_kmer_set(eng, kmer_fasta, kmer_size):
    if isinstance(kmer_fasta, KmerSet):
        return kmer_fasta
    return KmerSet.from_fasta(eng, kmer_size, kmer_fasta)


def _scan_parent_jellyfish(parent_bam, ref_fasta, kmer_fasta, kmer_size, parent_dir,
                           threads=4, n_filter_kmers=None, engine=None):
    """Count the filter k-mers in a parent BAM → ``dict{canonical k-mer: count}``
    holding only k-mers seen at least once (reference ``:115-283``:
    ``samtools fasta | jellyfish count --if`` then ``dump -c -L 1``)."""
    eng = get_engine(engine)
    kset = _as_kmer_set(eng, kmer_fasta, kmer_size)
    table = kset.build_table(n_min=n_filter_kmers or 0)
    table, tot = count_bam_into_table(eng, parent_bam, table, _engine.MODE_COUNT_IF_PRESENT,
                                      0, threads)
    logger.info("  parent scan: %d reads, %d k-mer instances, %d hits",
                tot["reads"], tot["windows"], tot["hits"])
    n, lo, hi, p0, _p1 = eng.threshold_compact(table, min0=1, want_planes=True)
    keys = eng.keys_to_pyints(lo, hi)
    counts = p0.cpu().numpy().view(np.uint32).tolist()
    table.close()
    return {kmer_of(key, kmer_size): int(c) for key, c in zip(keys, counts)}


def _build_proband_jf_index(proband_unique_fa, kmer_size, tmpdir, n_proband_unique=None,
                            engine=None):
    """Membership table of the proband-unique k-mers (reference ``:369-436``)."""
    eng = get_engine(engine)
    kset = _as_kmer_set(eng, proband_unique_fa, kmer_size)
    return kset.build_table(n_min=n_proband_unique or 0)
