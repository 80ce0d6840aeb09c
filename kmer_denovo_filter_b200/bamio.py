"""Host BAM input: ctypes binding of the library's multi-threaded BGZF/BAM decoder.

Replaces ``samtools fasta -F 0xD00`` (reference ``core/jellyfish_wrappers.py:
159-165``, ``discovery/pipeline.py:106-112, 369-375``) and the pysam iteration
of the anchoring scan (``core/bam_scanner.py:405-414``) for BAM input.  Reads are
delivered as packed batches (:class:`HostBatch`) ready for upload; alignment
metadata comes as flat numpy arrays with pysam-like per-record accessors for
the (few) informative reads the CPU cluster step needs.
"""

import ctypes
import struct
import zlib
import os

import numpy as np

from . import engine as _engine

MODE_FASTA = 0   # samtools fasta -F 0xD00 semantics (counting input)
MODE_SCAN = 1    # skip secondary + duplicate (anchoring scan)
MODE_ALL = 2

_vp, _u64, _i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int


class _Batch(ctypes.Structure):
    _fields_ = [
        ("impl", _vp), ("n_reads", _u64), ("n_bases", _u64),
        ("codes", _vp), ("valid", _vp), ("read_starts", _vp), ("read_lens", _vp),
        ("rec_index", _vp), ("ref_id", _vp), ("pos", _vp), ("next_ref_id", _vp),
        ("next_pos", _vp), ("flag", _vp), ("mapq", _vp), ("qname_off", _vp),
        ("qname_blob", _vp), ("cigar_off", _vp), ("cigar_blob", _vp),
        ("sa_off", _vp), ("sa_blob", _vp), ("qual_off", _vp), ("qual_blob", _vp), ("raw_off", _vp), ("raw_blob", _vp),
        ("at_eof", _i), ("has_invalid", _i), ("invalid_pos", _vp), ("n_invalid", _u64),
        ("rec_uoff", _vp), ("fasta_keep", _vp),
    ]


def _arr(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    ct = np.ctypeslib.as_ctypes_type(dtype)
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(int(n),))


class Record:
    """pysam.AlignedSegment look-alike built lazily from a :class:`HostBatch`."""

    __slots__ = ("batch", "i")

    def __init__(self, batch, i):
        self.batch = batch
        self.i = i

    @property
    def flag(self):
        return int(self.batch.flag[self.i])

    @property
    def query_name(self):
        b = self.batch
        return bytes(b.qname_blob[int(b.qname_off[self.i]):int(b.qname_off[self.i + 1])]).decode()

    @property
    def reference_id(self):
        return int(self.batch.ref_id[self.i])

    @property
    def reference_name(self):
        r = self.reference_id
        return self.batch.ref_names[r] if r >= 0 else None

    @property
    def reference_start(self):
        return int(self.batch.pos[self.i])

    @property
    def cigartuples(self):
        b = self.batch
        c = b.cigar_blob[int(b.cigar_off[self.i]):int(b.cigar_off[self.i + 1])]
        if c.shape[0] == 0:
            return None
        return [(int(x) & 15, int(x) >> 4) for x in c.tolist()]

    @property
    def reference_end(self):
        cig = self.cigartuples
        if self.is_unmapped or not cig:
            return None
        end = self.reference_start
        for op, ln in cig:
            if op in (0, 2, 3, 7, 8):
                end += ln
        return end

    @property
    def mapping_quality(self):
        return int(self.batch.mapq[self.i])

    @property
    def query_length(self):
        return int(self.batch.read_lens[self.i])

    is_paired = property(lambda s: bool(s.flag & 0x1))
    is_proper_pair = property(lambda s: bool(s.flag & 0x2))
    is_unmapped = property(lambda s: bool(s.flag & 0x4))
    mate_is_unmapped = property(lambda s: bool(s.flag & 0x8))
    is_reverse = property(lambda s: bool(s.flag & 0x10))
    is_secondary = property(lambda s: bool(s.flag & 0x100))
    is_duplicate = property(lambda s: bool(s.flag & 0x400))
    is_supplementary = property(lambda s: bool(s.flag & 0x800))

    def has_tag(self, tag):
        if tag != "SA":
            raise KeyError("only the SA tag is decoded")
        b = self.batch
        return b.sa_off[self.i + 1] > b.sa_off[self.i]

    def get_tag(self, tag):
        if not self.has_tag(tag):
            raise KeyError(tag)
        b = self.batch
        return bytes(b.sa_blob[int(b.sa_off[self.i]):int(b.sa_off[self.i + 1])]).decode()

    def get_aligned_pairs(self, matches_only=True):
        """pysam ``get_aligned_pairs``: (query_pos, ref_pos) per CIGAR column;
        with ``matches_only=False`` insertions / soft clips give (q, None) and
        deletions / skips (None, r) (``core/bam_scanner.py:111``,
        ``vcf/pipeline.py:683``)."""
        pairs = []
        q = 0
        r = self.reference_start
        for op, ln in self.cigartuples or ():
            if op in (0, 7, 8):
                pairs.extend((q + j, r + j) for j in range(ln))
                q += ln
                r += ln
            elif op in (1, 4):
                if not matches_only:
                    pairs.extend((q + j, None) for j in range(ln))
                q += ln
            elif op in (2, 3):
                if not matches_only:
                    pairs.extend((None, r + j) for j in range(ln))
                r += ln
        return pairs

    def get_reference_positions(self, full_length=False):
        if full_length:
            out = [None] * self.query_length
            for q, r in self.get_aligned_pairs(matches_only=True):
                out[q] = r
            return out
        return [r for _q, r in self.get_aligned_pairs(matches_only=True)]

    @property
    def query_qualities(self):
        """Phred qualities as a list, or None when absent (needs want_meta=2)."""
        b = self.batch
        if not getattr(b, "has_quals", False):
            raise _engine.KdfError("batch was decoded without base qualities")
        q = b.qual_blob[int(b.qual_off[self.i]):int(b.qual_off[self.i + 1])]
        if q.shape[0] == 0 or int(q[0]) == 0xFF:
            return None
        return q.tolist()

    @property
    def query_sequence(self):
        """Decode the read from the packed stream (N for invalid bases)."""
        b = self.batch
        n = int(b.read_lens[self.i])
        if n == 0:
            return None
        p = int(b.read_starts[self.i]) + np.arange(n, dtype=np.uint64)
        c = (b.codes[p >> np.uint64(5)] >> (np.uint64(62) - np.uint64(2) * (p & np.uint64(31)))) & np.uint64(3)
        v = (b.valid[p >> np.uint64(5)] >> (np.uint32(31) - (p & np.uint64(31)).astype(np.uint32))) & np.uint32(1)
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[c.astype(np.int64)]
        s = np.where(v.astype(bool), s, ord("N")).astype(np.uint8)
        return s.tobytes().decode()


class HostBatch(_engine.HostStream):
    """A packed batch plus optional per-record metadata (numpy views into the
    library-owned buffers; call :meth:`close` or let it be collected)."""

    def __init__(self, lib, raw, ref_names, want_meta):
        self._lib = lib
        self._raw = raw
        n = int(raw.n_reads)
        nw = (int(raw.n_bases) + 31) // 32
        super().__init__(_arr(raw.codes, nw, np.uint64), _arr(raw.valid, nw, np.uint32),
                         raw.n_bases, _arr(raw.read_starts, n, np.uint64),
                         _arr(raw.read_lens, n, np.uint32))
        self.ref_names = ref_names
        if raw.has_invalid:   # sparse form of `valid`: uploads send this list instead of the bitmap
            self.invalid = _arr(raw.invalid_pos, int(raw.n_invalid), np.uint32)
        self.rec_index = _arr(raw.rec_index, n, np.uint64)
        self.rec_uoff = _arr(raw.rec_uoff, n, np.uint64)       # kdf_bam_fetch_records keys
        self.fasta_keep = _arr(raw.fasta_keep, n, np.uint8)    # member of the MODE_FASTA stream
        self.at_eof = bool(raw.at_eof)
        self.has_meta = bool(want_meta)
        if want_meta:
            self.ref_id = _arr(raw.ref_id, n, np.int32)
            self.pos = _arr(raw.pos, n, np.int32)
            self.next_ref_id = _arr(raw.next_ref_id, n, np.int32)
            self.next_pos = _arr(raw.next_pos, n, np.int32)
            self.flag = _arr(raw.flag, n, np.uint16)
            self.mapq = _arr(raw.mapq, n, np.uint8)
            self.qname_off = _arr(raw.qname_off, n + 1, np.uint64)
            self.qname_blob = _arr(raw.qname_blob, int(self.qname_off[-1]) if n else 0, np.uint8)
            self.cigar_off = _arr(raw.cigar_off, n + 1, np.uint64)
            self.cigar_blob = _arr(raw.cigar_blob, int(self.cigar_off[-1]) if n else 0, np.uint32)
            self.sa_off = _arr(raw.sa_off, n + 1, np.uint64)
            self.sa_blob = _arr(raw.sa_blob, int(self.sa_off[-1]) if n else 0, np.uint8)
            self.has_raw = int(want_meta) >= 3
            if self.has_raw:
                self.raw_off = _arr(raw.raw_off, n + 1, np.uint64)
                self.raw_blob = _arr(raw.raw_blob, int(self.raw_off[-1]) if n else 0, np.uint8)
            self.has_quals = int(want_meta) >= 2
            if self.has_quals:
                self.qual_off = _arr(raw.qual_off, n + 1, np.uint64)
                self.qual_blob = _arr(raw.qual_blob, int(self.qual_off[-1]) if n else 0, np.uint8)

    def record(self, i):
        if not self.has_meta:
            raise _engine.KdfError("batch was decoded without metadata")
        return Record(self, int(i))

    def close(self):
        if self._raw is not None:
            self._lib.kdf_bam_batch_free(ctypes.byref(self._raw))
            self._raw = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BamReader:
    """Sequential BAM decoder.  ``for batch in BamReader(p).batches(MODE_FASTA)``."""

    def __init__(self, path, threads=None):
        self.lib = _engine.load_library()
        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        lower = path.lower()
        if lower.endswith(".cram"):
            raise _engine.KdfError(
                "CRAM input is not supported by the built-in decoder (needs htslib codecs); "
                "convert with `samtools view -b` first")
        if threads is None:
            threads = os.cpu_count() or 1
        h = _vp()
        rc = self.lib.kdf_bam_open(path.encode(), int(threads), ctypes.byref(h))
        if rc != 0:
            raise _engine.KdfError("cannot read BAM %s: %s" % (
                path, self.lib.kdf_host_last_error().decode()))
        self.handle = h
        self.path = path
        n = self.lib.kdf_bam_n_refs(h)
        ln = ctypes.c_uint64()
        ptr = self.lib.kdf_bam_header_text(h, ctypes.byref(ln))
        self.header_text = ctypes.string_at(ptr, ln.value) if ptr and ln.value else b""
        self.references = [self.lib.kdf_bam_ref_name(h, i).decode() for i in range(n)]
        self.lengths = [int(self.lib.kdf_bam_ref_len(h, i)) for i in range(n)]

    def next_batch(self, mode, max_bases=0, want_meta=False):
        raw = _Batch()
        rc = self.lib.kdf_bam_next_batch(self.handle, mode, int(max_bases), int(want_meta),
                                         ctypes.byref(raw))
        if rc != 0:
            raise _engine.KdfError("BAM decode failed for %s: %s" % (
                self.path, self.lib.kdf_host_last_error().decode()))
        return HostBatch(self.lib, raw, self.references, want_meta)

    def batches(self, mode, max_bases=0, want_meta=False):
        while True:
            b = self.next_batch(mode, max_bases, want_meta)
            if b.n_reads:
                yield b
            if b.at_eof:
                break

    def seek(self, voffset, begin=None, end=None):
        """Continue the decode at BGZF virtual offset ``voffset`` (the start of a record:
        an entry of the .bai linear index).  ``begin``: records before this virtual offset
        are parsed for the QNAME-run state only; ``end``: end of file is reported at the
        first record at or after this virtual offset."""
        for fn, v in ((self.lib.kdf_bam_set_end, end if end is not None else 0xFFFFFFFFFFFFFFFF),
                      (self.lib.kdf_bam_seek, voffset)):
            if fn(self.handle, int(v)) != 0:
                raise _engine.KdfError("cannot seek in %s: %s" % (
                    self.path, self.lib.kdf_host_last_error().decode()))
        if begin is not None and self.lib.kdf_bam_set_begin(self.handle, int(begin)) != 0:
            raise _engine.KdfError(self.lib.kdf_host_last_error().decode())

    def fetch(self, tid, beg, end, want_meta=2, mode=MODE_ALL):
        """Batches holding every record of reference ``tid`` that overlaps ``[beg, end)``
        (and possibly a few more around it) — ``pysam.AlignmentFile.fetch`` through the
        .bai linear index: the decode starts at the first record that overlaps the 16 kbp
        window holding ``beg`` and stops at the first batch that lies wholly past ``end``.
        Needs ``<bam>.bai``."""
        if getattr(self, "_bai", None) is None:
            bai = find_bai(self.path)
            if bai is None:
                raise _engine.KdfError("region fetch needs an index: %s.bai not found" % self.path)
            self._bai = read_bai(bai)
            self.lib.kdf_bam_set_chunk_bytes(self.handle, 256 << 10)
        if tid < 0 or tid >= len(self._bai):
            return
        lin = self._bai[tid]["linear"]
        if lin.shape[0] == 0:
            return
        # SAM spec 5.1.3: the linear index holds, per 16 kbp window, the smallest offset of the
        # alignments that OVERLAP the window — so a record that starts before `beg` and
        # reaches it is at or after the entry of beg's window.  Windows nothing overlaps are 0
        # (samtools) or repeat a neighbour: take the first non-empty one the region touches.
        voff = 0
        for w in range(min(int(beg) >> 14, lin.shape[0] - 1), min(max(int(end) - 1, int(beg)) >> 14, lin.shape[0] - 1) + 1):
            voff = int(lin[w])
            if voff:
                break
        if voff == 0:
            return
        self.seek(voff)
        while True:
            b = self.next_batch(mode, 1 << 18, want_meta)
            if b.n_reads == 0:
                b.close()
                return
            past = (b.ref_id != tid) | (b.pos.astype(np.int64) >= int(end))
            yield b
            if bool(past.all()) or bool(past[-1]) or b.at_eof:
                return

    def set_end(self, voffset):
        if self.lib.kdf_bam_set_end(self.handle, int(voffset) if voffset is not None else 0xFFFFFFFFFFFFFFFF) != 0:
            raise _engine.KdfError(self.lib.kdf_host_last_error().decode())

    def fetch_records(self, uoffs):
        """Raw BAM records (bytes after block_size) at uncompressed offsets ``uoffs``
        (``HostBatch.rec_uoff`` of records this reader has decoded): only the BGZF blocks
        that hold them are inflated (``kdf_bam_fetch_records``).  → list of bytes."""
        u = np.ascontiguousarray(uoffs, dtype=np.uint64)
        n = int(u.shape[0])
        if n == 0:
            return []
        off = np.zeros(n + 1, dtype=np.uint64)
        need = ctypes.c_uint64(0)
        cap = max(1024, 512 * n)
        while True:
            buf = np.empty(cap, dtype=np.uint8)
            rc = self.lib.kdf_bam_fetch_records(self.handle, u.ctypes.data_as(_vp), n,
                                                buf.ctypes.data_as(_vp), cap, off.ctypes.data_as(_vp),
                                                ctypes.byref(need))
            if rc != 0:
                raise _engine.KdfError("cannot fetch records from %s: %s" % (
                    self.path, self.lib.kdf_host_last_error().decode()))
            if need.value <= cap:
                break
            cap = int(need.value)
        o = off.astype(np.int64)
        return [buf[o[i]:o[i + 1]].tobytes() for i in range(n)]

    def close(self):
        if self.handle is not None:
            self.lib.kdf_bam_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_bai(path):
    """Parse a ``.bai`` → list (one per reference) of dicts ``{"bins": {bin: [(beg, end)
    virtual offsets]}, "linear": uint64 array of 16 kbp-window virtual offsets}``."""
    data = open(path, "rb").read()
    if data[:4] != b"BAI\x01":
        raise _engine.KdfError("%s is not a BAI index" % path)
    n_ref = struct.unpack_from("<i", data, 4)[0]
    off = 8
    out = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", data, off)[0]
        off += 4
        bins = {}
        for _b in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", data, off)
            off += 8
            ch = np.frombuffer(data, dtype="<u8", count=2 * n_chunk, offset=off).reshape(-1, 2)
            off += 16 * n_chunk
            bins[b] = ch
        n_intv = struct.unpack_from("<i", data, off)[0]
        off += 4
        lin = np.frombuffer(data, dtype="<u8", count=n_intv, offset=off).copy()
        off += 8 * n_intv
        out.append({"bins": bins, "linear": lin})
    return out


def find_bai(bam_path):
    for cand in (bam_path + ".bai", os.path.splitext(bam_path)[0] + ".bai"):
        if os.path.isfile(cand):
            return cand
    return None


def shard_offsets(bam_path, world):
    """Virtual offsets that cut a coordinate-sorted, indexed BAM into ``world`` contiguous
    ranges of about equal compressed size, each starting at a record (an entry of the .bai
    linear index), plus for every cut a slightly earlier entry to warm the QNAME-run state
    up from.  → list of ``(warm, begin, end)`` per rank (None = file start / end)."""
    bai = find_bai(bam_path)
    if bai is None:
        raise _engine.KdfError("sharding %s over %d ranks needs its .bai index" % (bam_path, world))
    entries = np.unique(np.concatenate([r["linear"] for r in read_bai(bai)] + [np.zeros(0, np.uint64)]))
    entries = entries[entries > 0]
    size = os.path.getsize(bam_path)
    cuts, warms = [], []
    for r in range(1, world):
        target = np.uint64((size * r // world) << 16)
        j = int(np.searchsorted(entries, target))
        if j >= entries.shape[0] or (cuts and int(entries[j]) <= cuts[-1]):
            cuts.append(None)
            warms.append(None)
            continue
        cuts.append(int(entries[j]))
        warms.append(int(entries[j - 1]) if j > 0 else None)
    out = []
    for r in range(world):
        begin = cuts[r - 1] if r > 0 else None
        warm = warms[r - 1] if r > 0 else None
        end = cuts[r] if r < world - 1 else None
        if r > 0 and begin is None:      # an earlier cut could not be placed: this rank gets nothing
            out.append(("empty", None, None))
            continue
        if end is None and r < world - 1:
            nxt = [c for c in cuts[r:] if c is not None]
            end = nxt[0] if nxt else None
        out.append((warm, begin, end))
    return out


def open_shard(bam_path, rank, world, threads=None):
    """A reader over rank ``rank``'s range of the file (``shard_offsets``)."""
    rd = BamReader(bam_path, threads=threads)
    if world <= 1:
        return rd
    warm, begin, end = shard_offsets(bam_path, world)[rank]
    if warm == "empty":
        rd.seek(0, end=0)      # nothing
        rd.empty_shard = True
        return rd
    if begin is None:
        rd.set_end(end)
    else:
        rd.seek(warm if warm is not None else begin, begin=begin, end=end)
    return rd


def counting_view(batch):
    """The ``samtools fasta -F 0xD00`` stream of a batch decoded in scan mode: the same
    packed codes with the bases of the records outside that stream (supplementary reads,
    later records of a QNAME run: ``fasta_keep == 0``) marked invalid, so that they start
    no k-mer window.  → ``HostStream`` sharing the batch's code words (the batch must
    outlive it)."""
    drop = np.flatnonzero(batch.fasta_keep == 0)
    if drop.shape[0] == 0:
        hs = _engine.HostStream(batch.codes, batch.valid, batch.n_bases, batch.read_starts,
                                batch.read_lens, invalid=getattr(batch, "invalid", None))
        return hs
    valid = np.array(batch.valid, dtype=np.uint32, copy=True)
    extra = []
    for r in drop.tolist():
        s, ln = int(batch.read_starts[r]), int(batch.read_lens[r])
        if ln == 0:
            continue
        pos = np.arange(s, s + ln, dtype=np.uint64)
        np.bitwise_and.at(valid, (pos >> np.uint64(5)).astype(np.int64),
                          ~(np.uint32(0x80000000) >> (pos & np.uint64(31)).astype(np.uint32)))
        extra.append(pos.astype(np.uint32))
    invalid = getattr(batch, "invalid", None)
    if invalid is not None:
        invalid = np.union1d(invalid, np.concatenate(extra)) if extra else invalid
    return _engine.HostStream(batch.codes, valid, batch.n_bases, batch.read_starts, batch.read_lens,
                              invalid=invalid)


def bgzf_write(path, data, level=6, threads=None):
    """``data`` (bytes-like) as a BGZF file + EOF marker (``kdf_bgzf_write``: every thread
    deflates blocks).  → uint64 array with the file offset of every 0xff00-byte block and
    of the EOF block (BAI / TBI virtual offsets: ``block_offset << 16 | offset in block``)."""
    lib = _engine.load_library()
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    n = int(buf.shape[0])
    nb = (n + BGZF_WRITE_BLOCK - 1) // BGZF_WRITE_BLOCK
    coff = np.zeros(nb + 1, dtype=np.uint64)
    got = ctypes.c_uint64(0)
    rc = lib.kdf_bgzf_write(os.fsencode(path), buf.ctypes.data_as(_vp) if n else None, n, int(level),
                            int(threads or os.cpu_count() or 1), coff.ctypes.data_as(_vp), nb + 1,
                            ctypes.byref(got))
    if rc != 0:
        raise _engine.KdfError("cannot write %s: %s" % (path, lib.kdf_host_last_error().decode()))
    return coff


BGZF_WRITE_BLOCK = 0xff00


def read_fasta_sequences(path):
    """Sequences of a (possibly gzip-compressed) FASTA as a list of bytes."""
    import gzip
    opener = gzip.open if path.endswith(".gz") else open
    names, seqs, cur = [], [], []
    with opener(path, "rb") as fh:
        for line in fh:
            if line.startswith(b">"):
                if names:
                    seqs.append(b"".join(cur))
                names.append(line[1:].split()[0].decode() if len(line) > 1 else "")
                cur = []
            else:
                cur.append(line.strip())
    if names:
        seqs.append(b"".join(cur))
    return names, seqs


# ---------------------------------------------------------------------------
# BAM + BAI writer (replaces pysam's AlignmentFile("wb") / sort / index for the
# informative-reads BAM, reference discovery/pipeline.py:1979-2079)
# ---------------------------------------------------------------------------

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_BGZF_BLOCK = 0xFF00


def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def append_int_tag(raw, tag, value):
    """BAM record bytes (without block_size) + an ``i``-typed aux tag."""
    return bytes(raw) + tag.encode() + b"i" + struct.pack("<i", int(value))


def append_z_tag(raw, tag, text):
    """BAM record bytes (without block_size) + a ``Z``-typed (string) aux tag."""
    return bytes(raw) + tag.encode() + b"Z" + text.encode() + b"\0"


def _record_span(raw):
    """(ref_id, pos, end, flag) of a raw BAM record."""
    ref_id, pos, l_name, _mq, _bin, n_cig, flag, _l_seq = struct.unpack_from("<iiBBHHHi", raw, 0)
    end = pos
    off = 32 + l_name
    for c in range(n_cig):
        v = struct.unpack_from("<I", raw, off + 4 * c)[0]
        if (v & 15) in (0, 2, 3, 7, 8):
            end += v >> 4
    if end == pos:
        end = pos + 1
    return ref_id, pos, end, flag


def write_sorted_bam(path, header_text, ref_names, ref_lens, records, index=True):
    """Coordinate-sort ``records`` (raw BAM records without block_size), write them
    as BGZF-compressed BAM at ``path`` and, with ``index``, a ``.bai`` beside it.
    Order: (reference id, position, strand) with unplaced reads last, ties in input
    order — what ``samtools sort`` produces."""
    keyed = []
    for i, raw in enumerate(records):
        ref_id, pos, end, flag = _record_span(raw)
        tid = ref_id if ref_id >= 0 else 1 << 31
        keyed.append((tid, ((pos + 1) << 1) | (1 if flag & 0x10 else 0), i, raw, ref_id, pos, end, flag))
    keyed.sort(key=lambda t: (t[0], t[1], t[2]))
    text = header_text if isinstance(header_text, bytes) else header_text.encode()
    if b"SO:" in text.split(b"\n", 1)[0]:
        first, _, rest = text.partition(b"\n")
        import re as _re
        first = _re.sub(rb"SO:\S+", b"SO:coordinate", first)
        text = first + b"\n" + rest
    head = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(ref_names)))
    for name, ln in zip(ref_names, ref_lens):
        nb = name.encode() + b"\0"
        head += struct.pack("<i", len(nb)) + nb + struct.pack("<i", int(ln))
    # uncompressed stream with the offset of every record
    stream = bytearray(head)
    starts = []
    for t in keyed:
        starts.append(len(stream))
        stream += struct.pack("<i", len(t[3])) + t[3]
    starts.append(len(stream))
    # BGZF blocks; voffset(u) = file offset of u's block << 16 | offset inside it
    block_file_off = []
    with open(path, "wb") as fh:
        for off in range(0, len(stream), _BGZF_BLOCK):
            block_file_off.append(fh.tell())
            chunk = bytes(stream[off:off + _BGZF_BLOCK])
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) +
                     comp + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        eof_off = fh.tell()
        fh.write(_BGZF_EOF)
    block_file_off.append(eof_off)

    def voff(u):
        b = u // _BGZF_BLOCK
        return (block_file_off[b] << 16) | (u - b * _BGZF_BLOCK) if u < len(stream) else (eof_off << 16)

    if not index:
        return len(keyed)
    n_ref = len(ref_names)
    bins = [dict() for _ in range(n_ref)]
    linear = [dict() for _ in range(n_ref)]
    n_no_coor = 0
    for j, t in enumerate(keyed):
        ref_id, pos, end = t[4], t[5], t[6]
        if ref_id < 0 or pos < 0:
            n_no_coor += 1
            continue
        v0, v1 = voff(starts[j]), voff(starts[j + 1])
        chunks = bins[ref_id].setdefault(_reg2bin(pos, end), [])
        if chunks and chunks[-1][1] == v0:
            chunks[-1][1] = v1
        else:
            chunks.append([v0, v1])
        for w in range(pos >> 14, ((end - 1) >> 14) + 1):
            if w not in linear[ref_id]:
                linear[ref_id][w] = v0
    out = bytearray(b"BAI\1" + struct.pack("<i", n_ref))
    for r in range(n_ref):
        out += struct.pack("<i", len(bins[r]))
        for b in sorted(bins[r]):
            out += struct.pack("<Ii", b, len(bins[r][b]))
            for v0, v1 in bins[r][b]:
                out += struct.pack("<QQ", v0, v1)
        n_intv = (max(linear[r]) + 1) if linear[r] else 0
        out += struct.pack("<i", n_intv)
        last = 0
        for w in range(n_intv):
            last = linear[r].get(w, last)
            out += struct.pack("<Q", last)
    out += struct.pack("<Q", n_no_coor)
    with open(path + ".bai", "wb") as fh:
        fh.write(out)
    return len(keyed)
