"""Stdlib BGZF/BAM reader and writer (oracle; test infrastructure only).

The reference reads alignments through pysam/htslib and streams reads to
Jellyfish through ``samtools fasta -F 0xD00`` (reference
``core/jellyfish_wrappers.py:159-165``, ``discovery/pipeline.py:106-112,
369-375``).  Neither tool exists in this image, so the oracle decodes the BAM
container itself (SAM/BAM spec v1, section 4) and restates the two record
streams the reference consumes:

* :func:`fasta_stream`  — what ``samtools fasta -F 0xD00`` emits:
  drop records with ``flag & 0xD00``; within a maximal run of consecutive
  records that share a QNAME keep only the first record of each read-part
  (READ1 / READ2 / other).
* :func:`scan_stream`   — what the anchoring loop iterates
  (reference ``core/bam_scanner.py:405-414``): every record except secondary
  (0x100) and duplicate (0x400); supplementary and unmapped records are kept;
  records without a sequence are skipped by the caller.
"""

import gzip
import struct
import zlib

SEQ_NT16 = "=ACMGRSVTWYHKDBN"
CIGAR_OPS = "MIDNSHP=X"

FLAG_PAIRED = 0x1
FLAG_PROPER_PAIR = 0x2
FLAG_UNMAP = 0x4
FLAG_MUNMAP = 0x8
FLAG_REVERSE = 0x10
FLAG_READ1 = 0x40
FLAG_READ2 = 0x80
FLAG_SECONDARY = 0x100
FLAG_QCFAIL = 0x200
FLAG_DUP = 0x400
FLAG_SUPPLEMENTARY = 0x800

_NT16_PAIR = [SEQ_NT16[b >> 4] + SEQ_NT16[b & 15] for b in range(256)]


class BamRecord:
    """One alignment record with the pysam-like accessors the reference uses."""

    __slots__ = (
        "ref_id", "pos", "mapq", "flag", "next_ref_id", "next_pos", "tlen",
        "qname", "cigar", "seq", "qual", "tags_raw", "ref_names", "_tags",
    )

    # --- pysam.AlignedSegment look-alikes ---------------------------------
    @property
    def query_name(self):
        return self.qname

    @property
    def query_sequence(self):
        return self.seq if self.seq else None

    @property
    def query_qualities(self):
        return self.qual

    @property
    def reference_name(self):
        return self.ref_names[self.ref_id] if self.ref_id >= 0 else None

    @property
    def reference_start(self):
        return self.pos

    @property
    def reference_end(self):
        if self.is_unmapped or not self.cigar:
            return None
        end = self.pos
        for op, ln in self.cigar:
            if op in (0, 2, 3, 7, 8):
                end += ln
        return end

    @property
    def cigartuples(self):
        return self.cigar if self.cigar else None

    @property
    def mapping_quality(self):
        return self.mapq

    is_paired = property(lambda s: bool(s.flag & FLAG_PAIRED))
    is_proper_pair = property(lambda s: bool(s.flag & FLAG_PROPER_PAIR))
    is_unmapped = property(lambda s: bool(s.flag & FLAG_UNMAP))
    mate_is_unmapped = property(lambda s: bool(s.flag & FLAG_MUNMAP))
    is_reverse = property(lambda s: bool(s.flag & FLAG_REVERSE))
    is_read1 = property(lambda s: bool(s.flag & FLAG_READ1))
    is_read2 = property(lambda s: bool(s.flag & FLAG_READ2))
    is_secondary = property(lambda s: bool(s.flag & FLAG_SECONDARY))
    is_duplicate = property(lambda s: bool(s.flag & FLAG_DUP))
    is_supplementary = property(lambda s: bool(s.flag & FLAG_SUPPLEMENTARY))

    def _parse_tags(self):
        if self._tags is not None:
            return self._tags
        tags = {}
        b = self.tags_raw
        i = 0
        n = len(b)
        while i + 3 <= n:
            tag = b[i:i + 2].decode()
            typ = chr(b[i + 2])
            i += 3
            if typ == "A":
                val = chr(b[i]); i += 1
            elif typ in "cC":
                val = struct.unpack_from("<b" if typ == "c" else "<B", b, i)[0]; i += 1
            elif typ in "sS":
                val = struct.unpack_from("<h" if typ == "s" else "<H", b, i)[0]; i += 2
            elif typ in "iI":
                val = struct.unpack_from("<i" if typ == "i" else "<I", b, i)[0]; i += 4
            elif typ == "f":
                val = struct.unpack_from("<f", b, i)[0]; i += 4
            elif typ in "ZH":
                j = b.index(0, i)
                val = b[i:j].decode(); i = j + 1
            elif typ == "B":
                sub = chr(b[i]); cnt = struct.unpack_from("<I", b, i + 1)[0]; i += 5
                size = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
                fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
                val = list(struct.unpack_from("<%d%s" % (cnt, fmt), b, i)); i += size * cnt
            else:
                raise ValueError("bad BAM tag type %r" % typ)
            tags[tag] = val
        self._tags = tags
        return tags

    def has_tag(self, tag):
        return tag in self._parse_tags()

    def get_tag(self, tag):
        return self._parse_tags()[tag]

    def get_aligned_pairs(self, matches_only=False):
        """pysam ``get_aligned_pairs`` (reference ``core/bam_scanner.py:111``)."""
        pairs = []
        q = 0
        r = self.pos
        for op, ln in self.cigar or ():
            if op in (0, 7, 8):
                for i in range(ln):
                    pairs.append((q + i, r + i))
                q += ln; r += ln
            elif op in (1, 4):
                if not matches_only:
                    for i in range(ln):
                        pairs.append((q + i, None))
                q += ln
            elif op in (2, 3):
                if not matches_only:
                    for i in range(ln):
                        pairs.append((None, r + i))
                r += ln
            # H, P consume nothing
        return pairs

    def get_reference_positions(self, full_length=False):
        if full_length:
            out = [None] * len(self.seq)
            for q, r in self.get_aligned_pairs(matches_only=True):
                out[q] = r
            return out
        return [r for _q, r in self.get_aligned_pairs(matches_only=True)]


def _inflate_all(path):
    """Concatenated BGZF members are plain multi-member gzip."""
    with gzip.open(path, "rb") as fh:
        return fh.read()


def read_bam(path):
    """Return ``(ref_names, ref_lengths, records)`` for a BAM file."""
    data = _inflate_all(path)
    if data[:4] != b"BAM\x01":
        raise ValueError("not a BAM file: %s" % path)
    l_text = struct.unpack_from("<i", data, 4)[0]
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", data, off)[0]
    off += 4
    ref_names = []
    ref_lengths = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, off)[0]
        off += 4
        ref_names.append(data[off:off + l_name - 1].decode())
        off += l_name
        ref_lengths.append(struct.unpack_from("<i", data, off)[0])
        off += 4
    records = []
    n = len(data)
    unpack_core = struct.Struct("<iiBBHHHiiii").unpack_from
    while off + 4 <= n:
        block_size = struct.unpack_from("<i", data, off)[0]
        off += 4
        end = off + block_size
        (ref_id, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq,
         next_ref_id, next_pos, tlen) = unpack_core(data, off)
        p = off + 32
        rec = BamRecord()
        rec.ref_names = ref_names
        rec.ref_id = ref_id
        rec.pos = pos
        rec.mapq = mapq
        rec.flag = flag
        rec.next_ref_id = next_ref_id
        rec.next_pos = next_pos
        rec.tlen = tlen
        rec.qname = data[p:p + l_read_name - 1].decode()
        p += l_read_name
        cig = struct.unpack_from("<%dI" % n_cigar, data, p)
        rec.cigar = [(c & 15, c >> 4) for c in cig]
        p += 4 * n_cigar
        nb = (l_seq + 1) // 2
        rec.seq = "".join(_NT16_PAIR[b] for b in data[p:p + nb])[:l_seq]
        p += nb
        q = data[p:p + l_seq]
        rec.qual = None if (l_seq and q[0] == 0xFF) else list(q)
        p += l_seq
        rec.tags_raw = data[p:end]
        rec._tags = None
        records.append(rec)
        off = end
    return ref_names, ref_lengths, records


# ---------------------------------------------------------------------------
# The two record streams of the reference
# ---------------------------------------------------------------------------

def _read_part(flag):
    r1 = bool(flag & FLAG_READ1)
    r2 = bool(flag & FLAG_READ2)
    if r1 and not r2:
        return 1
    if r2 and not r1:
        return 2
    return 0


def fasta_stream(records):
    """Records that ``samtools fasta -F 0xD00`` would emit, in file order.

    Flag filter first, then the same-QNAME collapse over *consecutive kept*
    records (samtools bam2fq groups adjacent records of one template and emits
    at most one per read-part).  Established against the reference goldens
    (SURVEY "Five facts" 4): no collapse gives 51223/6777/728, this rule gives
    the committed 51125/6679/630.
    """
    out = []
    cur_name = None
    seen_parts = set()
    for rec in records:
        if rec.flag & 0xD00:
            continue
        if rec.qname != cur_name:
            cur_name = rec.qname
            seen_parts = set()
        part = _read_part(rec.flag)
        if part in seen_parts:
            continue
        seen_parts.add(part)
        out.append(rec)
    return out


def scan_stream(records):
    """Records the anchoring scan visits (``core/bam_scanner.py:405-414``)."""
    return [r for r in records
            if not (r.flag & (FLAG_SECONDARY | FLAG_DUP))]


# ---------------------------------------------------------------------------
# Writer (used to build synthetic trios for tests; the reference's own test
# helpers do this with pysam, reference ``tests/helpers.py:6-107``)
# ---------------------------------------------------------------------------

_NT16_CODE = {c: i for i, c in enumerate(SEQ_NT16)}


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


def encode_record(ref_id, pos, qname, flag, mapq, cigar, seq, qual=None,
                  next_ref_id=-1, next_pos=-1, tlen=0, tags=b""):
    """Serialise one BAM alignment record (without BGZF framing)."""
    name = qname.encode() + b"\0"
    l_seq = len(seq)
    ref_len = sum(ln for op, ln in cigar if op in (0, 2, 3, 7, 8))
    end = pos + (ref_len if ref_len else 1)
    bin_ = _reg2bin(max(pos, 0), max(end, 1)) if pos >= 0 else 4680
    core = struct.pack(
        "<iiBBHHHiiii", ref_id, pos, len(name), mapq, bin_, len(cigar), flag,
        l_seq, next_ref_id, next_pos, tlen)
    cig = b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cigar)
    codes = [_NT16_CODE.get(c.upper(), 15) for c in seq]
    if l_seq & 1:
        codes.append(0)
    packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
    if qual is None:
        q = b"\xff" * l_seq
    else:
        q = bytes(qual)
    body = core + name + cig + packed + q + tags
    return struct.pack("<i", len(body)) + body


def _bgzf_block(payload, level=6):
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = comp.compress(payload) + comp.flush()
    bsize = len(cdata) + 25
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return hdr + cdata + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def write_bam(path, ref_names, ref_lengths, encoded_records, header_text=None,
              block_payload=60000):
    """Write a BAM from already-encoded records (see :func:`encode_record`)."""
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(
            "@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in zip(ref_names, ref_lengths))
    text = header_text.encode()
    head = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(ref_names))
    for n, l in zip(ref_names, ref_lengths):
        nb = n.encode() + b"\0"
        head += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    with open(path, "wb") as fh:
        buf = bytearray(head)
        for rec in encoded_records:
            if len(buf) + len(rec) > block_payload and buf:
                fh.write(_bgzf_block(bytes(buf)))
                buf = bytearray()
            buf += rec
            while len(buf) > block_payload:
                fh.write(_bgzf_block(bytes(buf[:block_payload])))
                del buf[:block_payload]
        if buf:
            fh.write(_bgzf_block(bytes(buf)))
        fh.write(BGZF_EOF)


def read_fasta(path):
    """Return ``[(name, sequence)]`` from a FASTA file (plain or gzip)."""
    opener = gzip.open if path.endswith(".gz") else open
    out = []
    name = None
    chunks = []
    with opener(path, "rt") as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    out.append((name, "".join(chunks)))
                name = line[1:].split()[0] if len(line) > 1 else ""
                chunks = []
            elif line:
                chunks.append(line)
    if name is not None:
        out.append((name, "".join(chunks)))
    return out
