"""ctypes binding of ``libkdf_sm100.so`` (C ABI in ``include/kdf.h``).

PyTorch is used for device buffers and streams only; every computation on the
k-mer path is a hand-written sm_100a kernel reached through the C ABI.  There
is **no CPU fallback**: constructing :class:`CudaEngine` without the compiled
library or without a CUDA device raises.

Reference seams replaced (reference ``src/kmer_denovo_filter/``):
``core/jellyfish_wrappers.py`` (count / count --if / dump), ``kmer_utils.py``
``JellyfishKmerQuery`` (query), ``discovery/pipeline.py`` Modules 1-2.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KDF_LIB") or os.path.join(_HERE, "libkdf_sm100.so")

KDF_OK = 0
MODE_INSERT_COUNT = 0
MODE_INSERT_ONLY = 1
MODE_COUNT_IF_PRESENT = 2
MODE_MARK_IF_PRESENT = 3
STAT_WINDOWS, STAT_FULL, STAT_HITS, STAT_NEW, N_STATS = 0, 1, 2, 3, 4
NDISTINCT_OVERFLOW = 0xFFFFFFFF
U32_MAX = 0xFFFFFFFF


class KdfError(RuntimeError):
    """Raised when a libkdf call fails (mirrors the reference's
    ``RuntimeError("jellyfish ... failed: <stderr>")``,
    ``core/jellyfish_wrappers.py:239-242``)."""


class _Stream(ctypes.Structure):
    _fields_ = [("codes", ctypes.c_void_p), ("valid", ctypes.c_void_p),
                ("n_bases", ctypes.c_uint64)]


class _DeviceProps(ctypes.Structure):
    _fields_ = [("sm_count", ctypes.c_int), ("cc_major", ctypes.c_int),
                ("cc_minor", ctypes.c_int), ("l2_bytes", ctypes.c_int),
                ("hbm_bytes", ctypes.c_uint64), ("name", ctypes.c_char * 128)]


# name -> (restype, argtypes); every symbol include/kdf.h declares
_vp, _u64, _u32, _i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
_SIGNATURES = {
    "kdf_version": (_i, []),
    "kdf_last_error": (ctypes.c_char_p, []),
    "kdf_device_info": (_i, [_i, ctypes.POINTER(_DeviceProps)]),
    "kdf_key_words": (_i, [_i]),
    "kdf_table_bytes": (ctypes.c_size_t, [_u64, _i]),
    "kdf_table_capacity_for": (_u64, [_u64]),
    "kdf_table_create": (_i, [ctypes.POINTER(_vp), _i, _u64, _vp, _vp]),
    "kdf_table_destroy": (_i, [_vp]),
    "kdf_table_clear": (_i, [_vp, _vp]),
    "kdf_table_clear_plane": (_i, [_vp, _i, _vp]),
    "kdf_table_build_filter": (_i, [_vp, _vp, _u64, _vp]),
    "kdf_table_info": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_u64)]),
    "kdf_extract_canonical": (_i, [ctypes.POINTER(_Stream), _i, _vp, _vp, _vp, _vp]),
    "kdf_count_stream": (_i, [_vp, ctypes.POINTER(_Stream), _i, _i, _u32, _vp, _vp]),
    "kdf_update_keys": (_i, [_vp, _vp, _vp, _u64, _i, _i, _u32, _vp, _vp]),
    "kdf_add_planes": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp]),
    "kdf_threshold_compact": (_i, [_vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _u64, _vp, _vp]),
    "kdf_lookup_keys": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp]),
    "kdf_scan_reads": (_i, [_vp, ctypes.POINTER(_Stream), _vp, _vp, _u64, _u32, _vp, _vp, _vp, _vp,
                            _u64, _vp, _vp, _vp]),
    "kdf_scan_stream_hits": (_i, [_vp, ctypes.POINTER(_Stream), _vp, _vp, _u64, _vp, _vp, _vp]),
    "kdf_reduce_hits_scratch_bytes": (ctypes.c_size_t, [_u64]),
    "kdf_reduce_hits": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp, _vp]),
    "kdf_bin_stream": (_i, [ctypes.POINTER(_Stream), _i, _i, _i, _vp, _u64, _vp, _vp, _vp, _vp]),
    "kdf_bin_stream_range": (_i, [ctypes.POINTER(_Stream), _u64, _u64, _i, _i, _i, _vp, _u64, _vp, _vp,
                                  _vp, _vp]),
    "kdf_bin_stream_to_range": (_i, [ctypes.POINTER(_Stream), _u64, _u64, _i, _i, _i, _vp, _u64, _vp,
                                     _vp, _vp, _vp]),
    "kdf_bin_stream_to": (_i, [ctypes.POINTER(_Stream), _i, _i, _i, _vp, _u64, _vp, _vp, _vp, _vp]),
    "kdf_bin_keys": (_i, [_vp, _vp, _u64, _i, _i, _i, _vp, _u64, _vp, _vp, _vp]),
    "kdf_bin_stream_pass": (_i, [ctypes.POINTER(_Stream), _u64, _u64, _i, _i, _i, _vp, _vp, _u64, _vp, _vp,
                                 _vp, _i, _u32, _vp]),
    "kdf_bin_keys_pass": (_i, [_vp, _vp, _u64, _i, _i, _i, _i, _u32, _vp, _u64, _vp, _vp, _vp]),
    "kdf_hit_coverage_scratch_bytes": (ctypes.c_size_t, [_u64, _i]),
    "kdf_hit_coverage": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp]),
    "kdf_debug_hit_coverage_host": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _vp, _vp, _vp]),
    "kdf_update_bins": (_i, [_vp, _i, _vp, _u64, _vp, _i, _i, _u32, _vp, _vp]),
    "kdf_count_bins_packed": (_i, [_i, _u32, _u32, _u32, _u32, _u32, _i]),
    "kdf_count_bins": (_i, [_i, _i, _vp, _u64, _vp, _vp, _u64, _vp, _vp, _u64, _u32, _u32, _u32, _u32,
                            _vp, _vp, _vp, _vp, _u64, _vp, _u32, _vp, _vp]),
    "kdf_count_bins_multi": (_i, [_i, _i, _i, _i, _vp, _u64, _vp, _vp, _u64, _vp, _vp, _u64, _u32, _u32, _u32,
                                  _u32, _vp, _vp, _vp, _vp, _u64, _vp, _u32, _vp, _vp]),
    "kdf_count_bins_pass": (_i, [_i, _i, _i, _i, _i, _u32, _vp, _u64, _vp, _vp, _u64, _vp, _vp, _u64, _u32,
                                 _u32, _u32, _u32, _vp, _vp, _vp, _vp, _u64, _vp, _u32, _vp, _vp]),
    "kdf_debug_hash_host": (_i, [_vp, _vp, _u64, _i, _i, _u32, _u32, _vp, _vp, _vp]),
    "kdf_pack_sequences": (_u64, [_vp, _vp, _u64, _vp, _vp, _vp]),
    "kdf_invalid_positions": (_u64, [_vp, _u64, _vp, _u64]),
    "kdf_valid_from_invalid": (_i, [_vp, _u64, _vp, _u64, _vp]),
    "kdf_debug_extract_host": (_i, [_vp, _vp, _u64, _i, _i, _vp, _vp, _vp]),
    "kdf_bench_random_access": (_i, [_vp, _u64, _u64, _i, _vp, _vp]),
    # host BGZF/BAM decoder (bound in bamio.py)
    "kdf_bam_open": (_i, [ctypes.c_char_p, _i, ctypes.POINTER(_vp)]),
    "kdf_bam_close": (None, [_vp]),
    "kdf_bam_header_text": (ctypes.c_void_p, [_vp, ctypes.POINTER(_u64)]),
    "kdf_bam_n_refs": (_i, [_vp]),
    "kdf_bam_ref_name": (ctypes.c_char_p, [_vp, _i]),
    "kdf_bam_ref_len": (ctypes.c_int64, [_vp, _i]),
    "kdf_bam_next_batch": (_i, [_vp, _i, _u64, _i, _vp]),
    "kdf_bam_batch_free": (None, [_vp]),
    "kdf_host_last_error": (ctypes.c_char_p, []),
    "kdf_bam_fetch_records": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, _vp]),
    "kdf_bam_seek": (_i, [_vp, _u64]),
    "kdf_bam_set_end": (_i, [_vp, _u64]),
    "kdf_bam_set_begin": (_i, [_vp, _u64]),
    "kdf_bam_set_chunk_bytes": (_i, [_vp, _u64]),
    "kdf_bgzf_write": (_i, [ctypes.c_char_p, _vp, _u64, _i, _i, _vp, _u64, _vp]),
    "kdf_bgzf_inflate_block": (_i, [_vp, ctypes.c_uint32, _vp, ctypes.c_uint32, _i, _i]),
    "kdf_crc32": (ctypes.c_uint32, [_vp, _u64]),
    "kdf_fasta_layout": (_i, [_vp, _u64, _i, _vp, _vp]),
    "kdf_fasta_pack": (_i, [_vp, _u64, _i, _vp, _vp, _vp, _vp]),
}

_lib = None


def load_library(path=None):
    """Load the shared library once; fail loudly when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise KdfError(
            "libkdf_sm100.so not found at %s — build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % p)
    lib = ctypes.CDLL(p)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
# host-side packing (CPU helper of the library; no device needed)
# ---------------------------------------------------------------------------

class HostStream:
    """Packed stream in host memory (numpy): ``codes`` u64, ``valid`` u32,
    ``n_bases``, ``read_starts`` u64, ``read_lens`` u32, and optionally ``invalid``
    (u32, ascending positions of the invalid bases: the sparse form of ``valid`` —
    when present, uploads send it instead of the bitmap, include/kdf.h)."""

    __slots__ = ("codes", "valid", "n_bases", "read_starts", "read_lens", "invalid")

    def __init__(self, codes, valid, n_bases, read_starts, read_lens, invalid=None):
        self.codes = codes
        self.valid = valid
        self.n_bases = int(n_bases)
        self.read_starts = read_starts
        self.read_lens = read_lens
        self.invalid = invalid

    def with_sparse_validity(self):
        """Compute ``invalid`` from ``valid`` (what the BAM decoder emits with a batch)."""
        self.invalid = invalid_positions(self.valid, self.n_bases)
        return self

    @property
    def n_reads(self):
        return int(self.read_lens.shape[0])

    @property
    def n_words(self):
        return (self.n_bases + 31) // 32

    def window_count_upper_bound(self, k):
        l = self.read_lens.astype(np.int64) - (k - 1)
        return int(np.clip(l, 0, None).sum())


def invalid_positions(valid, n_bases):
    """Ascending positions of the 0 bits of a validity bitmap (``kdf_invalid_positions``),
    or None when the stream is longer than 2^32 bases."""
    lib = load_library()
    if n_bases > 0xFFFFFFFF:
        return None
    v = np.ascontiguousarray(valid, dtype=np.uint32)
    if v.shape[0] == 0:
        return np.zeros(0, dtype=np.uint32)
    n = lib.kdf_invalid_positions(_np_ptr(v), int(n_bases), None, 0)
    out = np.zeros(max(n, 1), dtype=np.uint32)
    lib.kdf_invalid_positions(_np_ptr(v), int(n_bases), _np_ptr(out), n)
    return out[:n]


def pack_sequences(seqs):
    """Pack an iterable of str/bytes sequences into a :class:`HostStream`."""
    lib = load_library()
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    n = len(bs)
    lens = np.fromiter((len(b) for b in bs), dtype=np.uint64, count=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    blob = b"".join(bs)
    buf = np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(1, dtype=np.uint8)
    total = int(lens.sum()) + max(n - 1, 0)
    n_words = (total + 31) // 32
    codes = np.zeros(max(n_words, 1), dtype=np.uint64)
    valid = np.zeros(max(n_words, 1), dtype=np.uint32)
    read_offsets = np.zeros(n + 1, dtype=np.uint64)
    got = lib.kdf_pack_sequences(_np_ptr(buf), _np_ptr(offsets), n, _np_ptr(codes),
                                 _np_ptr(valid), _np_ptr(read_offsets))
    assert got == total
    return HostStream(codes[:n_words], valid[:n_words], total,
                      read_offsets[:n].copy(), lens.astype(np.uint32)).with_sparse_validity()


def pack_fasta_file(path, threads=None):
    """The sequences of a (possibly gzip-compressed) FASTA file as one :class:`HostStream`
    (records separated by an invalid base), packed by the library with all host threads
    (``kdf_fasta_pack``).  → ``(HostStream, n_records)``."""
    lib = load_library()
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rb") as fh:
            text = np.frombuffer(fh.read(), dtype=np.uint8)
    else:
        text = np.fromfile(path, dtype=np.uint8)
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    n = int(text.shape[0])
    tp = _np_ptr(text) if n else None
    n_seqs, n_bases = ctypes.c_uint64(0), ctypes.c_uint64(0)
    if lib.kdf_fasta_layout(tp, n, int(threads), ctypes.byref(n_seqs), ctypes.byref(n_bases)) != KDF_OK:
        raise KdfError(lib.kdf_host_last_error().decode())
    ns, total = int(n_seqs.value), int(n_bases.value)
    n_words = (total + 31) // 32
    codes = np.empty(max(n_words, 1), dtype=np.uint64)
    valid = np.empty(max(n_words, 1), dtype=np.uint32)
    starts = np.zeros(max(ns, 1), dtype=np.uint64)
    lens = np.zeros(max(ns, 1), dtype=np.uint64)
    if lib.kdf_fasta_pack(tp, n, int(threads), _np_ptr(codes), _np_ptr(valid), _np_ptr(starts),
                          _np_ptr(lens)) != KDF_OK:
        raise KdfError(lib.kdf_host_last_error().decode())
    hs = HostStream(codes[:n_words], valid[:n_words], total, starts[:ns].copy(),
                    lens[:ns].astype(np.uint32)).with_sparse_validity()
    return hs, ns


def debug_extract_host(hs, k, random_access=False):
    """Run the device iterator templates on the CPU (test hook).  ``random_access``:
    False = rolling WindowIter, True = WindowAt, 2 / 3 = the chunked WindowChunks of the
    stream kernels (chunks of 4 / 16 windows)."""
    lib = load_library()
    n = hs.n_bases
    lo = np.zeros(max(n, 1), dtype=np.uint64)
    hi = np.zeros(max(n, 1), dtype=np.uint64)
    ok = np.zeros(max(n, 1), dtype=np.uint8)
    codes = np.ascontiguousarray(hs.codes)
    valid = np.ascontiguousarray(hs.valid)
    if codes.size == 0:
        return lo[:0], hi[:0], ok[:0].astype(bool)
    rc = lib.kdf_debug_extract_host(_np_ptr(codes), _np_ptr(valid), n, k,
                                    int(random_access), _np_ptr(lo), _np_ptr(hi), _np_ptr(ok))
    if rc != KDF_OK:
        raise KdfError(lib.kdf_last_error().decode())
    return lo[:n], hi[:n], ok[:n].astype(bool)


# ---------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------

class DeviceStream:
    """A packed stream resident in HBM (torch tensors as raw buffers)."""

    __slots__ = ("codes", "valid", "n_bases", "read_starts", "read_lens", "_c", "chunks", "ready")

    def __init__(self, codes, valid, n_bases, read_starts=None, read_lens=None):
        self.codes = codes          # torch.int64 (bit pattern of u64)
        self.valid = valid          # torch.int32 (bit pattern of u32)
        self.n_bases = int(n_bases)
        self.read_starts = read_starts  # torch.int64 or None
        self.read_lens = read_lens      # torch.int32 or None
        self._c = _Stream(codes.data_ptr(), valid.data_ptr(), self.n_bases)
        # optional upload schedule [(first_word, n_words, ready_event)]: the words of a
        # chunk (and the few after it that its windows span) are valid once the event fires
        self.chunks = None
        self.ready = None   # event after which the whole stream is in HBM (None = already)

    @property
    def n_reads(self):
        return 0 if self.read_lens is None else int(self.read_lens.shape[0])

    def c(self):
        return ctypes.byref(self._c)


class KmerTable:
    """Device hash table handle; owns the torch buffer that backs the slots."""

    def __init__(self, engine, k, capacity):
        self.engine = engine
        self.k = int(k)
        self.capacity = max(4, (int(capacity) + 3) & ~3)   # whole 32-byte buckets
        lib = engine.lib
        self.key_words = lib.kdf_key_words(self.k)
        if not self.key_words:
            raise KdfError("k=%d unsupported by the GPU engine (1..64)" % k)
        nbytes = lib.kdf_table_bytes(self.capacity, self.key_words)
        torch = engine.torch
        # int64 storage guarantees >= 256-byte alignment from the caching allocator
        self.buf = torch.empty(nbytes // 8, dtype=torch.int64, device=engine.device)
        h = ctypes.c_void_p()
        engine._check(lib.kdf_table_create(ctypes.byref(h), self.k, self.capacity,
                                           self.buf.data_ptr(), engine.stream_ptr()))
        self.handle = h
        self.filter_buf = None

    @property
    def nbytes(self):
        return self.buf.numel() * 8

    def close(self):
        if self.handle is not None:
            self.engine.lib.kdf_table_destroy(self.handle)
            self.handle = None
            self.buf = None
            self.filter_buf = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KeyBins:
    """Fixed-capacity bins of canonical k-mers in HBM (``kdf_bin_stream``)."""

    def __init__(self, engine, k, n_parts, bin_cap, by_owner=False):
        torch = engine.torch
        self.engine = engine
        self.k = int(k)
        self.key_words = engine.lib.kdf_key_words(self.k)
        if not self.key_words:
            raise KdfError("k=%d unsupported by the GPU engine (1..64)" % k)
        self.n_parts = int(n_parts)
        self.bin_cap = max(4, (int(bin_cap) + 3) & ~3)
        self.by_owner = bool(by_owner)
        self.data = torch.empty(self.n_parts * self.bin_cap * self.key_words, dtype=torch.int64,
                                device=engine.device)
        self.cursors = torch.zeros(self.n_parts, dtype=torch.int64, device=engine.device)
        self.overflow = torch.zeros(1, dtype=torch.int64, device=engine.device)

    def counts(self):
        """Keys *offered* to each bin (may exceed bin_cap after an overflow)."""
        return self.cursors.cpu().numpy().astype(np.int64)

    def overflowed(self):
        return bool(int(self.overflow.item()))

    def reset(self):
        """Empty the bins (the next pass of a multi-pass count refills them)."""
        self.cursors.zero_()
        self.overflow.zero_()

    def bin_keys(self, b, count=None):
        """(lo, hi|None) device views of bin b."""
        n = int(self.cursors[b].item()) if count is None else int(count)
        n = min(n, self.bin_cap)
        kw = self.key_words
        seg = self.data[b * self.bin_cap * kw:(b * self.bin_cap + n) * kw]
        if kw == 1:
            return seg, None
        pair = seg.view(-1, 2)
        return pair[:, 0].contiguous(), pair[:, 1].contiguous()


def _fold_coverage(keys, counts):
    """Runs of ``kdf_hit_coverage`` (key = contig << 40 | pos << 1 | first) →
    (contig, pos, k-mer count, read count) per position."""
    keep = keys != np.uint64(0xFFFFFFFFFFFFFFFF)
    keys, counts = keys[keep], counts[keep].astype(np.int64)
    pos_key = keys >> np.uint64(1)
    first = (keys & np.uint64(1)).astype(bool)
    u, inv = np.unique(pos_key, return_inverse=True)
    kc = np.zeros(u.shape[0], dtype=np.int64)
    rc = np.zeros(u.shape[0], dtype=np.int64)
    np.add.at(kc, inv, counts)
    np.add.at(rc, inv[first], counts[first])
    return ((u >> np.uint64(39)).astype(np.int64), (u & np.uint64((1 << 39) - 1)).astype(np.int64),
            kc, rc)


def debug_hit_coverage_host(hit_read, hit_off, k, read_contig, read_ref_start, read_cig_off, cigar):
    """The K7 expansion run on the host by the same code the kernel uses (test hook);
    same return value as :meth:`CudaEngine.hit_coverage`."""
    lib = load_library()
    hr = np.ascontiguousarray(hit_read, dtype=np.uint32)
    ho = np.ascontiguousarray(hit_off, dtype=np.uint32)
    rc = np.ascontiguousarray(read_contig, dtype=np.int32)
    rs = np.ascontiguousarray(read_ref_start, dtype=np.int64)
    co = np.ascontiguousarray(read_cig_off, dtype=np.uint64)
    cg = np.ascontiguousarray(cigar, dtype=np.uint32)
    if cg.shape[0] == 0:
        cg = np.zeros(1, np.uint32)
    n = hr.shape[0]
    keys = np.zeros(max(n * k, 1), np.uint64)
    if n:
        r = lib.kdf_debug_hit_coverage_host(_np_ptr(hr), _np_ptr(ho), n, int(k), _np_ptr(rc), _np_ptr(rs),
                                            _np_ptr(co), _np_ptr(cg), _np_ptr(keys))
        if r != KDF_OK:
            raise KdfError(lib.kdf_last_error().decode())
    keys = np.sort(keys[:n * k])
    u, c = np.unique(keys, return_counts=True)
    return _fold_coverage(u, c.astype(np.uint32))


def debug_hash_host(lo, hi, key_words, log2_parts, n_buckets, n_ranks):
    """Partition / bucket / owner of each key, computed on the host by the same
    code the kernels use (test hook)."""
    lib = load_library()
    lo = np.ascontiguousarray(lo, dtype=np.uint64)
    hi = np.ascontiguousarray(hi, dtype=np.uint64) if hi is not None else None
    n = lo.shape[0]
    part = np.zeros(max(n, 1), np.uint32)
    bucket = np.zeros(max(n, 1), np.uint32)
    owner = np.zeros(max(n, 1), np.uint32)
    rc = lib.kdf_debug_hash_host(_np_ptr(lo), _np_ptr(hi) if hi is not None else None, n, key_words,
                                 log2_parts, n_buckets, n_ranks, _np_ptr(part), _np_ptr(bucket),
                                 _np_ptr(owner))
    if rc != KDF_OK:
        raise KdfError(lib.kdf_last_error().decode())
    return part[:n], bucket[:n], owner[:n]


class CudaEngine:
    """One engine per GPU / process.  All calls are ordered on torch's current
    stream of ``device``."""

    def __init__(self, device=None):
        import torch
        self.torch = torch
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise KdfError("CUDA device required: the k-mer engine has no CPU fallback")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        props = _DeviceProps()
        self._check(self.lib.kdf_device_info(self.device.index or 0, ctypes.byref(props)))
        self.props = {"sm_count": props.sm_count, "cc": (props.cc_major, props.cc_minor),
                      "l2_bytes": props.l2_bytes, "hbm_bytes": props.hbm_bytes,
                      "name": props.name.decode()}
        self.launches = 0  # kernels launched through this engine (bench.py gpu_launches)
        # name -> [(start event, end event)] when per-kernel timing is switched on
        # (bench.py roofline: CUDA events on the launching stream); None = off
        self.timers = None

    # -- plumbing ----------------------------------------------------------
    def _t0(self):
        if self.timers is None:
            return None
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.torch.cuda.current_stream(self.device))
        return ev

    def _t1(self, name, ev0):
        if ev0 is None:
            return
        ev1 = self.torch.cuda.Event(enable_timing=True)
        ev1.record(self.torch.cuda.current_stream(self.device))
        self.timers.setdefault(name, []).append((ev0, ev1))

    def kernel_times_ms(self):
        """name -> list of launch durations (ms); call after a synchronize."""
        return {n: [a.elapsed_time(b) for a, b in evs] for n, evs in (self.timers or {}).items()}

    def _check(self, rc):
        if rc != KDF_OK:
            raise KdfError("libkdf error %d: %s" % (rc, self.lib.kdf_last_error().decode()))

    def stream_ptr(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def zeros(self, n, dtype):
        return self.torch.zeros(n, dtype=dtype, device=self.device)

    def empty(self, n, dtype):
        return self.torch.empty(n, dtype=dtype, device=self.device)

    def free_device_bytes(self):
        """Bytes a new allocation could take now: free on the device plus the blocks
        torch's caching allocator holds without using them."""
        torch = self.torch
        free, _total = torch.cuda.mem_get_info(self.device)
        cached = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        return int(free + max(cached, 0))

    def new_stats(self):
        return self.zeros(N_STATS, self.torch.int64)

    def read_stats(self, stats):
        v = stats.cpu().numpy().view(np.uint64)
        return {"windows": int(v[STAT_WINDOWS]), "full": int(v[STAT_FULL]),
                "hits": int(v[STAT_HITS]), "new": int(v[STAT_NEW])}

    # -- data movement -----------------------------------------------------
    def upload(self, hs, non_blocking=False, with_reads=True, copy_stream=None):
        """HostStream → DeviceStream.  Device buffers are allocated on the current
        stream; the H2D copies run on it too, or on ``copy_stream`` (which first
        waits for the current stream, so a recycled buffer is never overwritten
        while earlier kernels still read it) — the caller then orders consumers
        after the copies with an event."""
        torch = self.torch
        main = torch.cuda.current_stream(self.device)

        def host(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a).view(dt))

        sparse = self._sparse(hs)
        src = [host(hs.codes, np.int64), host(hs.invalid if sparse else hs.valid, np.int32)]
        if with_reads and hs.read_lens is not None:
            src += [host(hs.read_starts, np.int64), host(hs.read_lens, np.int32)]
        dst = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in src]
        valid = torch.empty((hs.n_bases + 31) // 32, dtype=torch.int32, device=self.device) if sparse else None
        if copy_stream is not None:
            copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):
                for d, h in zip(dst, src):
                    d.copy_(h, non_blocking=True)
                if sparse:
                    self._valid_from_invalid(valid, hs.n_bases, dst[1], copy_stream)
        else:
            for d, h in zip(dst, src):
                d.copy_(h, non_blocking=non_blocking)
            if sparse:
                self._valid_from_invalid(valid, hs.n_bases, dst[1], main)
        rs, rl = (dst[2], dst[3]) if len(dst) == 4 else (None, None)
        return DeviceStream(dst[0], valid if sparse else dst[1], hs.n_bases, rs, rl)

    def _sparse(self, hs):
        """Send the invalid-position list instead of the validity bitmap?"""
        return (getattr(hs, "invalid", None) is not None and hs.n_bases > 0
                and os.environ.get("KDF_SPARSE_VALID", "1") != "0")

    def _valid_from_invalid(self, valid, n_bases, invalid_dev, stream):
        """Rebuild the validity bitmap on ``stream`` from the uploaded list; the list
        is kept alive by the stream-ordered allocator until the kernel has run."""
        self._check(self.lib.kdf_valid_from_invalid(
            valid.data_ptr(), int(n_bases), invalid_dev.data_ptr() if invalid_dev.numel() else None,
            int(invalid_dev.numel()), ctypes.c_void_p(stream.cuda_stream)))
        invalid_dev.record_stream(stream)
        self.launches += 2

    def upload_chunked(self, hs, copy_stream, n_chunks=8, with_reads=True):
        """HostStream → DeviceStream copied chunk by chunk on ``copy_stream``; the
        result carries ``chunks`` so that a consumer can start on the first words
        while the rest is still in flight (``bin_stream(word_range=...)``).  A
        chunk's event is recorded after the NEXT chunk's copy, because the windows
        starting in its last words read the first words of the next one."""
        torch = self.torch
        main = torch.cuda.current_stream(self.device)
        n_words = (hs.n_bases + 31) // 32
        sparse = self._sparse(hs)
        codes_h = torch.from_numpy(np.ascontiguousarray(hs.codes).view(np.int64))
        valid_h = torch.from_numpy(np.ascontiguousarray(hs.invalid if sparse else hs.valid).view(np.int32))
        codes = torch.empty(codes_h.shape, dtype=torch.int64, device=self.device)
        valid = torch.empty(n_words, dtype=torch.int32, device=self.device)
        rs = rl = None
        step = max((n_words + n_chunks - 1) // n_chunks, 1)
        bounds = [(a, min(a + step, n_words)) for a in range(0, n_words, step)]
        copy_stream.wait_stream(main)
        events = []
        with torch.cuda.stream(copy_stream):
            if sparse:   # the whole (small) list first, the bitmap rebuilt before the first chunk
                inv = torch.empty(valid_h.shape, dtype=torch.int32, device=self.device)
                inv.copy_(valid_h, non_blocking=True)
                self._valid_from_invalid(valid, hs.n_bases, inv, copy_stream)
            for a, b in bounds:
                codes[a:b].copy_(codes_h[a:b], non_blocking=True)
                if not sparse:
                    valid[a:b].copy_(valid_h[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append(ev)
            if with_reads and hs.read_lens is not None:
                rs_h = torch.from_numpy(np.ascontiguousarray(hs.read_starts).view(np.int64))
                rl_h = torch.from_numpy(np.ascontiguousarray(hs.read_lens).view(np.int32))
                rs = torch.empty(rs_h.shape, dtype=torch.int64, device=self.device)
                rl = torch.empty(rl_h.shape, dtype=torch.int32, device=self.device)
                rs.copy_(rs_h, non_blocking=True)
                rl.copy_(rl_h, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        ds = DeviceStream(codes, valid, hs.n_bases, rs, rl)
        ds.chunks = [(a, b - a, events[min(i + 1, len(events) - 1)]) for i, (a, b) in enumerate(bounds)]
        return ds, done

    def upload_read_index(self, ds, hs, copy_stream):
        """Copy ``read_starts`` / ``read_lens`` of ``hs`` onto ``ds`` (copy stream);
        → event."""
        torch = self.torch
        main = torch.cuda.current_stream(self.device)
        rs_h = torch.from_numpy(np.ascontiguousarray(hs.read_starts).view(np.int64))
        rl_h = torch.from_numpy(np.ascontiguousarray(hs.read_lens).view(np.int32))
        ds.read_starts = torch.empty(rs_h.shape, dtype=torch.int64, device=self.device)
        ds.read_lens = torch.empty(rl_h.shape, dtype=torch.int32, device=self.device)
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            ds.read_starts.copy_(rs_h, non_blocking=True)
            ds.read_lens.copy_(rl_h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def keys_to_device(self, keys, key_words):
        """Python ints / numpy → (lo, hi) int64 device tensors."""
        torch = self.torch
        if isinstance(keys, tuple):
            lo_np, hi_np = keys
        else:
            ks = list(keys)
            lo_np = np.fromiter((x & 0xFFFFFFFFFFFFFFFF for x in ks), dtype=np.uint64, count=len(ks))
            hi_np = np.fromiter((x >> 64 for x in ks), dtype=np.uint64, count=len(ks))
        lo = torch.from_numpy(lo_np.view(np.int64)).to(self.device)
        hi = torch.from_numpy(hi_np.view(np.int64)).to(self.device) if key_words == 2 else None
        return lo, hi

    @staticmethod
    def keys_to_pyints(lo, hi=None):
        lo_np = lo.cpu().numpy().view(np.uint64)
        if hi is None:
            return [int(x) for x in lo_np.tolist()]
        hi_np = hi.cpu().numpy().view(np.uint64)
        return [(int(h) << 64) | int(l) for h, l in zip(hi_np.tolist(), lo_np.tolist())]

    # -- tables --------------------------------------------------------------
    def capacity_for(self, n_keys):
        return int(self.lib.kdf_table_capacity_for(int(n_keys)))

    def new_table(self, k, n_keys=None, capacity=None):
        if capacity is None:
            capacity = self.capacity_for(n_keys or 0)
        self.launches += 1
        ev = self._t0()
        t = KmerTable(self, k, capacity)
        self._t1("table_clear", ev)
        return t

    @staticmethod
    def filter_words(n_keys, max_bytes):
        """32-bit words of the filter for ``n_keys`` keys within ``max_bytes``: 32 bits per
        key (false positives ~0.4 %) when that fits, else 16 (~1.5 %), else 8 (~5 %: the
        replicated filter set of an 8-GPU run, 31 M keys — one probe in twenty then goes to
        the table, still far cheaper than binning the whole parent), else 0 (no filter).
        ``KDF_FILTER_MIN_BITS=16`` drops the last tier."""
        tiers = (1, 2, 4) if os.environ.get("KDF_FILTER_MIN_BITS", "8") == "8" else (1, 2)
        for per_word in tiers:
            n_words = 1024
            while n_words * per_word < n_keys:
                n_words *= 2
            if n_words * 4 <= max_bytes:
                return n_words
        return 0

    FILTER_MAX_BYTES = int(os.environ.get("KDF_FILTER_MAX_MB", "32")) << 20
    SMEM_TABLE_BYTES = 160 * 1024

    def filter_applies(self, k, n_keys):
        """A table of ``n_keys`` final keys gets a filter: too large for the stream
        kernels' shared-memory copy, and its filter fits the L2 budget."""
        kw = self.lib.kdf_key_words(k)
        return (max(n_keys, 1) * 4 * 8 * kw > self.SMEM_TABLE_BYTES
                and self.filter_words(max(n_keys, 1), self.FILTER_MAX_BYTES) > 0
                and os.environ.get("KDF_TABLE_FILTER", "1") != "0")

    def build_filter(self, table, n_keys, max_bytes=1 << 30):
        """Attach a two-bit membership filter (4 bytes per key, ``kdf_table_build_filter``)
        to a table whose keys are final; the probing stream kernels then read one 32-bit
        word per window and probe the table only for the few windows it cannot rule out."""
        n_words = self.filter_words(n_keys, max_bytes)
        if not n_words:
            raise KdfError("build_filter: %d keys do not fit a filter of %d bytes" % (n_keys, max_bytes))
        ev = self._t0()
        table.filter_buf = self.zeros(n_words, self.torch.int32)   # kept alive with the table
        self._check(self.lib.kdf_table_build_filter(table.handle, table.filter_buf.data_ptr(), n_words,
                                                    self.stream_ptr()))
        self._t1("build_filter", ev)
        self.launches += 1

    def clear_plane(self, table, plane):
        self._check(self.lib.kdf_table_clear_plane(table.handle, plane, self.stream_ptr()))
        self.launches += 1

    # -- kernels -------------------------------------------------------------
    def extract_canonical(self, ds, k):
        """K1 → (lo, hi|None, ok_words) device tensors."""
        torch = self.torch
        kw = self.lib.kdf_key_words(k)
        if not kw:
            raise KdfError("k=%d unsupported (1..64)" % k)
        n = max(ds.n_bases, 1)
        lo = self.empty(n, torch.int64)
        hi = self.empty(n, torch.int64) if kw == 2 else None
        ok = self.zeros(max((ds.n_bases + 31) // 32, 1), torch.int32)
        self._check(self.lib.kdf_extract_canonical(
            ds.c(), k, lo.data_ptr(), hi.data_ptr() if hi is not None else None,
            ok.data_ptr(), self.stream_ptr()))
        self.launches += 1
        return lo[:ds.n_bases], (hi[:ds.n_bases] if hi is not None else None), ok

    def count_stream(self, table, ds, mode=MODE_INSERT_COUNT, plane=0, arg=1, stats=None):
        """K1+K2 fused."""
        ev = self._t0()
        self._check(self.lib.kdf_count_stream(
            table.handle, ds.c(), mode, plane, arg,
            stats.data_ptr() if stats is not None else None, self.stream_ptr()))
        self._t1("count_stream/mode%d/kw%d" % (mode, table.key_words), ev)
        self.launches += 1

    def update_keys(self, table, lo, hi=None, mode=MODE_INSERT_ONLY, plane=0, arg=1, stats=None):
        n = int(lo.shape[0])
        ev = self._t0()
        self._check(self.lib.kdf_update_keys(
            table.handle, lo.data_ptr() if n else None,
            hi.data_ptr() if (hi is not None and n) else None, n, mode, plane, arg,
            stats.data_ptr() if stats is not None else None, self.stream_ptr()))
        self._t1("update_keys/mode%d" % mode, ev)
        if n:
            self.launches += 1

    def add_planes(self, table, lo, hi=None, add0=None, add1=None):
        """Add per-key values to the planes of existing keys; returns #missing."""
        n = int(lo.shape[0])
        if not n:
            return 0
        miss = self.zeros(1, self.torch.int64)
        self._check(self.lib.kdf_add_planes(
            table.handle, lo.data_ptr(), hi.data_ptr() if hi is not None else None, n,
            add0.data_ptr() if add0 is not None else None,
            add1.data_ptr() if add1 is not None else None, miss.data_ptr(), self.stream_ptr()))
        self.launches += 1
        return int(miss.item())

    def check_not_full(self, stats):
        if self.read_stats(stats)["full"]:
            raise KdfError("k-mer table full (libkdf error %d): size it with more slots" % -4)

    def threshold_count(self, table, min0=0, max0=U32_MAX, min1=0, max1=U32_MAX):
        n_out = self.zeros(1, self.torch.int64)
        ev = self._t0()
        self._check(self.lib.kdf_threshold_compact(
            table.handle, min0, max0, min1, max1, None, None, None, None, 0,
            n_out.data_ptr(), self.stream_ptr()))
        self._t1("threshold_compact", ev)
        self.launches += 1
        return int(n_out.item())

    def threshold_compact(self, table, min0=0, max0=U32_MAX, min1=0, max1=U32_MAX,
                          want_planes=False, n_expected=None):
        """K3 → (n, lo, hi|None, p0|None, p1|None); exact two-pass sizing."""
        torch = self.torch
        n = self.threshold_count(table, min0, max0, min1, max1) if n_expected is None else int(n_expected)
        cap = max(n, 1)
        lo = self.empty(cap, torch.int64)
        hi = self.empty(cap, torch.int64) if table.key_words == 2 else None
        p0 = self.empty(cap, torch.int32) if want_planes else None
        p1 = self.empty(cap, torch.int32) if want_planes else None
        n_out = self.zeros(1, torch.int64)
        ev = self._t0()
        self._check(self.lib.kdf_threshold_compact(
            table.handle, min0, max0, min1, max1, lo.data_ptr(),
            hi.data_ptr() if hi is not None else None,
            p0.data_ptr() if p0 is not None else None,
            p1.data_ptr() if p1 is not None else None,
            cap, n_out.data_ptr(), self.stream_ptr()))
        self._t1("threshold_compact", ev)
        self.launches += 1
        got = int(n_out.item())
        if got > cap:
            raise KdfError("threshold_compact: %d matches exceed buffer %d" % (got, cap))
        return (got, lo[:got], hi[:got] if hi is not None else None,
                p0[:got] if p0 is not None else None, p1[:got] if p1 is not None else None)

    def lookup_keys(self, table, lo, hi=None, want_planes=True):
        """K4 → (found u8, p0, p1) device tensors."""
        torch = self.torch
        n = int(lo.shape[0])
        found = self.zeros(max(n, 1), torch.uint8)
        p0 = self.zeros(max(n, 1), torch.int32) if want_planes else None
        p1 = self.zeros(max(n, 1), torch.int32) if want_planes else None
        if n:
            ev = self._t0()
            self._check(self.lib.kdf_lookup_keys(
                table.handle, lo.data_ptr(), hi.data_ptr() if hi is not None else None, n,
                found.data_ptr(), p0.data_ptr() if p0 is not None else None,
                p1.data_ptr() if p1 is not None else None, self.stream_ptr()))
            self._t1("lookup_keys", ev)
            self.launches += 1
        return found[:n], (p0[:n] if p0 is not None else None), (p1[:n] if p1 is not None else None)

    def scan_reads(self, table, ds, min_distinct=1, hit_cap=None, stats=None, want_hits=True):
        """K4+K5 → dict(ndistinct, nhits, hit_pos, hit_slot, n_hits)."""
        torch = self.torch
        n_reads = ds.n_reads
        nd = self.zeros(max(n_reads, 1), torch.int32)
        nh = self.zeros(max(n_reads, 1), torch.int32)
        if hit_cap is None:
            hit_cap = 1 << 16
        while True:
            n_hits = self.zeros(1, torch.int64)
            hp = self.empty(max(hit_cap, 1), torch.int64) if want_hits else None
            hsl = self.empty(max(hit_cap, 1), torch.int32) if want_hits else None
            st = None
            if stats is not None:
                st = self.zeros(N_STATS, torch.int64)
            if n_reads:
                ev = self._t0()
                self._check(self.lib.kdf_scan_reads(
                    table.handle, ds.c(), ds.read_starts.data_ptr(), ds.read_lens.data_ptr(),
                    n_reads, min_distinct, nd.data_ptr(), nh.data_ptr(),
                    hp.data_ptr() if hp is not None else None,
                    hsl.data_ptr() if hsl is not None else None,
                    hit_cap if want_hits else 0, n_hits.data_ptr(),
                    st.data_ptr() if st is not None else None, self.stream_ptr()))
                self._t1("scan_reads/kw%d" % table.key_words, ev)
                self.launches += 1
            total = int(n_hits.item())
            if not want_hits or total <= hit_cap:
                break
            hit_cap = total  # exact retry (rare: only when hits are dense)
        if stats is not None and st is not None:
            stats += st
        return {"ndistinct": nd[:n_reads], "nhits": nh[:n_reads],
                "hit_pos": hp[:total] if hp is not None else None,
                "hit_slot": hsl[:total] if hsl is not None else None, "n_hits": total}

    def scan_reads_sparse(self, table, ds, stats=None, hit_cap=None):
        """K4+K5, sparse form: emit every hit window, then reduce per read on the
        device.  Returns dict(read u64[], ndistinct u32[], nhits u32[], first u64[],
        hit_pos u64[] sorted, hit_slot u32[] sorted) as numpy arrays holding one
        record per read with at least one hit, ordered by read index."""
        torch = self.torch
        if hit_cap is None:
            hit_cap = 1 << 20
        while True:
            n_hits = self.zeros(1, torch.int64)
            hp = self.empty(max(hit_cap, 1), torch.int64)
            hsl = self.empty(max(hit_cap, 1), torch.int32)
            st = self.zeros(N_STATS, torch.int64) if stats is not None else None
            ev = self._t0()
            self._check(self.lib.kdf_scan_stream_hits(
                table.handle, ds.c(), hp.data_ptr(), hsl.data_ptr(), hit_cap, n_hits.data_ptr(),
                st.data_ptr() if st is not None else None, self.stream_ptr()))
            self._t1("scan_stream_hits/kw%d" % table.key_words, ev)
            self.launches += 1
            total = int(n_hits.item())
            if total <= hit_cap:
                break
            hit_cap = total
        if stats is not None:
            stats += st
        empty = {"read": np.zeros(0, np.uint64), "ndistinct": np.zeros(0, np.uint32),
                 "nhits": np.zeros(0, np.uint32), "first": np.zeros(0, np.uint64),
                 "hit_pos": np.zeros(0, np.uint64), "hit_slot": np.zeros(0, np.uint32)}
        if total == 0:
            return empty
        n_reads = ds.n_reads
        nbytes = int(self.lib.kdf_reduce_hits_scratch_bytes(total))
        scratch = self.empty(nbytes, torch.uint8)
        spos = self.empty(total, torch.int64)
        sslot = self.empty(total, torch.int32)
        rr = self.empty(total, torch.int64)
        rnd = self.empty(total, torch.int32)
        rnh = self.empty(total, torch.int32)
        rf = self.empty(total, torch.int64)
        n_recs = self.zeros(1, torch.int64)
        ev = self._t0()
        self._check(self.lib.kdf_reduce_hits(
            hp.data_ptr(), hsl.data_ptr(), total, ds.read_starts.data_ptr(), n_reads,
            scratch.data_ptr(), nbytes, spos.data_ptr(), sslot.data_ptr(), rr.data_ptr(),
            rnd.data_ptr(), rnh.data_ptr(), rf.data_ptr(), n_recs.data_ptr(), self.stream_ptr()))
        self._t1("reduce_hits", ev)
        self.launches += 3   # radix sort passes + reduce
        n = int(n_recs.item())
        read = rr[:n].cpu().numpy().view(np.uint64)
        order = np.argsort(read, kind="stable")
        return {"read": read[order],
                "ndistinct": rnd[:n].cpu().numpy().view(np.uint32)[order],
                "nhits": rnh[:n].cpu().numpy().view(np.uint32)[order],
                "first": rf[:n].cpu().numpy().view(np.uint64)[order],
                "hit_pos": spos.cpu().numpy().view(np.uint64),
                "hit_slot": sslot.cpu().numpy().view(np.uint32)}

    # -- binning ---------------------------------------------------------------
    def new_bins(self, k, n_parts, bin_cap, by_owner=False):
        return KeyBins(self, k, n_parts, bin_cap, by_owner)

    def bin_stream(self, bins, ds, stats=None, word_range=None, pass_=None):
        """K2p / K6: append the canonical k-mers of a stream to hash-range (or owner)
        bins; ``word_range=(first, n)`` restricts the window starts to those words;
        ``pass_=(pass_log2, pass_val)`` bins only hash-range group ``pass_val`` of
        ``2**pass_log2`` (``kdf_bin_stream_pass``: the multi-pass count)."""
        ev = self._t0()
        first, n = word_range if word_range is not None else (0, (ds.n_bases + 31) // 32)
        plog, pval = pass_ if pass_ is not None else (0, 0)
        self._check(self.lib.kdf_bin_stream_pass(
            ds.c(), int(first), int(n), bins.k, 1 if bins.by_owner else 0, bins.n_parts,
            bins.data.data_ptr(), None, bins.bin_cap, bins.cursors.data_ptr(), bins.overflow.data_ptr(),
            stats.data_ptr() if stats is not None else None, int(plog), int(pval), self.stream_ptr()))
        self._t1("bin_stream/kw%d" % bins.key_words, ev)
        self.launches += 1

    def bin_stream_to(self, ds, k, bin_ptrs, bin_cap, cursors, overflow, by_owner=1, stats=None,
                      word_range=None, pass_=None):
        """K6 fused with the transfer: bin p goes to ``bin_ptrs[p]`` (device int64
        tensor of raw pointers, possibly peer memory over NVLink).  ``by_owner``: 1 =
        one bin per owner rank, R >= 2 = composite R owners x hash ranges."""
        ev = self._t0()
        first, n = word_range if word_range is not None else (0, (ds.n_bases + 31) // 32)
        plog, pval = pass_ if pass_ is not None else (0, 0)
        self._check(self.lib.kdf_bin_stream_pass(
            ds.c(), int(first), int(n), int(k), int(by_owner), int(bin_ptrs.shape[0]),
            None, bin_ptrs.data_ptr(),
            int(bin_cap), cursors.data_ptr(), overflow.data_ptr(),
            stats.data_ptr() if stats is not None else None, int(plog), int(pval), self.stream_ptr()))
        self._t1("bin_stream_to_peers/kw%d" % self.lib.kdf_key_words(int(k)), ev)
        self.launches += 1

    def bin_keys(self, bins, lo, hi=None, n=None, pass_=None):
        n = int(lo.shape[0]) if n is None else int(n)
        if not n:
            return
        ev = self._t0()
        plog, pval = pass_ if pass_ is not None else (0, 0)
        self._check(self.lib.kdf_bin_keys_pass(
            lo.data_ptr(), hi.data_ptr() if hi is not None else None, n, bins.k,
            1 if bins.by_owner else 0, bins.n_parts, int(plog), int(pval), bins.data.data_ptr(),
            bins.bin_cap, bins.cursors.data_ptr(), bins.overflow.data_ptr(), self.stream_ptr()))
        self._t1("bin_keys/kw%d" % bins.key_words, ev)
        self.launches += 1

    def hit_coverage(self, hit_read, hit_off, k, read_contig, read_ref_start, read_cig_off, cigar):
        """K7 (``kdf_hit_coverage``): reference positions covered by hit k-mers.
        Inputs are numpy arrays (hits sorted by (read, offset); per-read contig id,
        reference start, CIGAR offsets; BAM CIGAR words).  Returns numpy
        ``(contig i64[], pos i64[], kmer_count i64[], read_count i64[])``, one entry
        per covered position, sorted by (contig, pos)."""
        torch = self.torch
        n_hits = int(len(hit_read))
        if n_hits == 0:
            z = np.zeros(0, dtype=np.int64)
            return z, z, z, z

        def dev(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        d_hr, d_ho = dev(hit_read, np.uint32).view(torch.int32), dev(hit_off, np.uint32).view(torch.int32)
        d_rc, d_rs = dev(read_contig, np.int32), dev(read_ref_start, np.int64)
        d_co = dev(np.asarray(read_cig_off, dtype=np.uint64).view(np.int64), np.int64)
        d_cg = dev(np.asarray(cigar, dtype=np.uint32).view(np.int32), np.int32)
        if d_cg.numel() == 0:
            d_cg = self.zeros(1, torch.int32)
        n = n_hits * int(k)
        nbytes = self.lib.kdf_hit_coverage_scratch_bytes(n_hits, int(k))
        scratch = self.empty((nbytes + 7) // 8, torch.int64)
        keys = self.empty(n, torch.int64)
        counts = self.empty(n, torch.int32)
        n_out = self.zeros(1, torch.int64)
        ev = self._t0()
        self._check(self.lib.kdf_hit_coverage(
            d_hr.data_ptr(), d_ho.data_ptr(), n_hits, int(k), d_rc.data_ptr(), d_rs.data_ptr(),
            d_co.data_ptr(), d_cg.data_ptr(), scratch.data_ptr(), nbytes, keys.data_ptr(),
            counts.data_ptr(), n_out.data_ptr(), self.stream_ptr()))
        self._t1("hit_coverage", ev)
        self.launches += 3
        m = int(n_out.item())
        return _fold_coverage(keys[:m].cpu().numpy().view(np.uint64),
                              counts[:m].cpu().numpy().view(np.uint32))

    def update_bins(self, table, bins, mode=MODE_COUNT_IF_PRESENT, plane=0, arg=1, stats=None):
        """K2 over hash-range bins, bin after bin (``kdf_update_bins``): each bin only
        touches its own share of the table, which therefore stays in L2."""
        if bins.by_owner:
            raise KdfError("update_bins needs hash-range bins")
        ev = self._t0()
        self._check(self.lib.kdf_update_bins(
            table.handle, bins.n_parts, bins.data.data_ptr(), bins.bin_cap, bins.cursors.data_ptr(),
            mode, plane, arg, stats.data_ptr() if stats is not None else None, self.stream_ptr()))
        self._t1("update_bins/mode%d/kw%d" % (mode, bins.key_words), ev)
        self.launches += bins.n_parts

    def count_bins_packed(self, k, min_child_count):
        """True when the discovery chain's count_bins call (count >= min_child_count,
        not in the reference, no count planes wanted) takes the packed form of
        include/kdf.h: the table slice then holds keys only."""
        return bool(self.lib.kdf_count_bins_packed(k, min_child_count, U32_MAX, 0, 0,
                                                   min_child_count, 0))

    def count_bins(self, child_bins, ref_bins, slice_capacity, min0=0, max0=U32_MAX, min1=0,
                   max1=U32_MAX, count_min0=0, out_cap=1 << 20, want_planes=False, sub_split=1,
                   pass_=None):
        """Count every bin in an L2-resident slice and emit (see include/kdf.h).
        Returns dict(n_out, lo, hi, p0, p1, keys, full, hits, distinct, n_count, occupied);
        lo/hi/p0/p1 hold min(n_out, out_cap) entries.  ``pass_=(pass_log2, pass_val)``:
        the bins are those of one pass of a multi-pass count (``bin_stream(pass_=...)``)."""
        torch = self.torch
        k = child_bins.k
        kw = child_bins.key_words
        slice_capacity = max(4, (int(slice_capacity) + 3) & ~3)
        slice_buf = self.empty(self.lib.kdf_table_bytes(slice_capacity, kw) // 8, torch.int64)
        lo = self.empty(max(out_cap, 1), torch.int64)
        hi = self.empty(max(out_cap, 1), torch.int64) if kw == 2 else None
        p0 = self.empty(max(out_cap, 1), torch.int32) if want_planes else None
        p1 = self.empty(max(out_cap, 1), torch.int32) if want_planes else None
        n_out = self.zeros(1, torch.int64)
        ctr = self.zeros(6, torch.int64)
        ev = self._t0()
        n_src = getattr(child_bins, "n_src", 1)
        if ref_bins is not None and getattr(ref_bins, "n_src", 1) != n_src:
            raise KdfError("count_bins: child and reference bins must have the same sources")
        plog, pval = pass_ if pass_ is not None else (0, 0)
        self._check(self.lib.kdf_count_bins_pass(
            k, child_bins.n_parts, n_src, int(sub_split), int(plog), int(pval),
            child_bins.data.data_ptr(),
            child_bins.bin_cap,
            child_bins.cursors.data_ptr(),
            ref_bins.data.data_ptr() if ref_bins is not None else None,
            ref_bins.bin_cap if ref_bins is not None else 0,
            ref_bins.cursors.data_ptr() if ref_bins is not None else None,
            slice_buf.data_ptr(), slice_capacity, min0, max0, min1, max1, lo.data_ptr(),
            hi.data_ptr() if hi is not None else None,
            p0.data_ptr() if p0 is not None else None, p1.data_ptr() if p1 is not None else None,
            out_cap, n_out.data_ptr(), count_min0, ctr.data_ptr(), self.stream_ptr()))
        self._t1("count_bins/kw%d" % kw, ev)
        packed = bool(self.lib.kdf_count_bins_packed(k, min0, max0, min1, max1, count_min0,
                                                     1 if want_planes else 0))
        self.launches += (1 if packed else 2) + child_bins.n_parts * int(sub_split) * (
            1 + n_src * (1 + (1 if ref_bins is not None else 0)))
        c = ctr.cpu().numpy().view(np.uint64)
        n = int(n_out.item())
        m = min(n, out_cap)
        return {"n_out": n, "lo": lo[:m], "hi": hi[:m] if hi is not None else None,
                "p0": p0[:m] if p0 is not None else None, "p1": p1[:m] if p1 is not None else None,
                "keys": int(c[0]), "full": int(c[1]), "hits": int(c[2]), "distinct": int(c[3]),
                "n_count": int(c[4]), "occupied": int(c[5])}

    def debug_hash_host(self, lo, hi, key_words, log2_parts, n_buckets, n_ranks):
        return debug_hash_host(lo, hi, key_words, log2_parts, n_buckets, n_ranks)

    def bench_random_access(self, buf, n_ops, atomic):
        sink = self.zeros(1, self.torch.int64)
        self._check(self.lib.kdf_bench_random_access(
            buf.data_ptr(), buf.numel() * buf.element_size(), n_ops, int(atomic),
            sink.data_ptr(), self.stream_ptr()))
        self.launches += 1
