"""K-mer utilities — drop-in for the reference's ``kmer_utils.py`` on the k-mer path.

Same names, arguments and results as the reference for the string helpers
(``reverse_complement``, ``canonicalize``, ``_extract_read_kmers``;
reference ``kmer_utils.py:30-38, 91-121``), and :class:`GpuKmerQuery` replaces
``JellyfishKmerQuery`` (``kmer_utils.py:124-245``): the same duck-typed
interface (``query_batch``, ``scan_read``, ``close``; pinned by the fake in the
reference's ``tests/discovery/test_pipeline.py:1532-1542``) backed by the GPU
table instead of a ``jellyfish query`` subprocess per batch.
"""

import numpy as np

from . import engine as _engine

_COMP = str.maketrans("ACGTacgt", "TGCAtgca")
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def reverse_complement(seq):
    """Return the reverse complement of a DNA sequence."""
    return seq.translate(_COMP)[::-1]


def canonicalize(kmer):
    """Return the canonical (lexicographically smaller) form of a k-mer."""
    rc = reverse_complement(kmer)
    return rc if rc < kmer else kmer


def _extract_read_kmers(seq, kmer_size):
    """``(canon_at_pos, unique_candidates)`` of a read; windows with N skipped."""
    if len(seq) < kmer_size:
        return {}, []
    up = seq.upper()
    canon_at_pos = {}
    for i in range(len(up) - kmer_size + 1):
        w = up[i:i + kmer_size]
        if "N" not in w:
            canon_at_pos[i] = canonicalize(w)
    return canon_at_pos, list(dict.fromkeys(canon_at_pos.values()))


def _is_symbolic(allele):
    """True for VCF alleles without a literal sequence (``<DEL>``, ``*``, BND)."""
    if not allele:
        return True
    return allele[0] == "<" or allele == "*" or "[" in allele or "]" in allele


# ---- VCF-mode read helpers (CPU: a few dozen reads per variant) ------------

def read_supports_alt(read, variant_pos, ref, alt, min_baseq=0, *, aligned_pairs=None, seq=None,
                      quals=None):
    """True iff the read carries exactly ``alt`` over the reference span of the
    variant (reference ``kmer_utils.py:1037-1099``)."""
    if alt is None or _is_symbolic(alt):
        return False
    if seq is None:
        seq = read.query_sequence
    if seq is None:
        return False
    if min_baseq > 0 and quals is None:
        quals = read.query_qualities
    if aligned_pairs is None:
        aligned_pairs = read.get_aligned_pairs(matches_only=False)
    extracted = []
    in_region = False
    end = variant_pos + len(ref)
    for qpos, rpos in aligned_pairs:
        if rpos is not None and rpos >= end:
            break
        if rpos == variant_pos:
            in_region = True
        if in_region and qpos is not None:
            if min_baseq > 0 and quals is not None and quals[qpos] < min_baseq:
                return False
            extracted.append(seq[qpos])
    if not in_region:
        return False
    return "".join(extracted).upper() == alt.upper()


def extract_variant_spanning_kmers(read, variant_pos, k, min_baseq=0, ref=None, alt=None, *,
                                   aligned_pairs=None, seq=None, quals=None):
    """Canonical k-mers of the read that cover the variant, skipping windows with an
    N or a base below ``min_baseq`` (reference ``kmer_utils.py:1102-1172``)."""
    try:
        rp = read.get_reference_positions(full_length=True).index(variant_pos)
    except ValueError:
        return set()
    if seq is None:
        seq = read.query_sequence
    if seq is None:
        return set()
    if quals is None:
        quals = read.query_qualities
    alt_len = len(alt) if alt and not _is_symbolic(alt) else 1
    start_min = max(0, rp - k + 1)
    start_max = min(len(seq) - k, rp + alt_len - 1)
    window_end = start_max + k
    if window_end <= start_min:
        return set()
    up = seq[start_min:window_end].upper()
    bad = bytearray(len(up))
    for i, ch in enumerate(up):
        if ch == "N":
            bad[i] = 1
    if quals is not None and min_baseq > 0:
        for i in range(len(up)):
            if quals[start_min + i] < min_baseq:
                bad[i] = 1
    out = set()
    bad_count = sum(bad[:min(k, len(bad))])
    for s in range(start_min, start_max + 1):
        off = s - start_min
        if off > 0:
            bad_count += bad[off + k - 1] - bad[off - 1]
        if bad_count:
            continue
        out.add(canonicalize(seq[s:s + k]))
    return out


# ---- integer keys ---------------------------------------------------------

def key_of(kmer):
    """2-bit key (A0 C1 G2 T3, first base most significant) or None if non-ACGT."""
    v = 0
    for ch in kmer:
        c = _CODE.get(ch)
        if c is None:
            c = _CODE.get(ch.upper())
            if c is None:
                return None
        v = (v << 2) | c
    return v


def kmer_of(key, k):
    return "".join("ACGT"[(key >> (2 * (k - 1 - i))) & 3] for i in range(k))


def keys_to_arrays(keys):
    """Python-int keys → (lo u64, hi u64) numpy arrays."""
    ks = list(keys)
    lo = np.fromiter((x & 0xFFFFFFFFFFFFFFFF for x in ks), dtype=np.uint64, count=len(ks))
    hi = np.fromiter((x >> 64 for x in ks), dtype=np.uint64, count=len(ks))
    return lo, hi


class KmerSet:
    """A set of canonical k-mers held on the device (what the reference keeps
    as a FASTA file of k-mers between pipeline stages)."""

    def __init__(self, eng, k, lo, hi=None):
        self.engine = eng
        self.k = int(k)
        self.lo = lo
        self.hi = hi

    def __len__(self):
        return int(self.lo.shape[0])

    def to_pyints(self):
        return self.engine.keys_to_pyints(self.lo, self.hi)

    def to_strings(self):
        return [kmer_of(x, self.k) for x in self.to_pyints()]

    def write_fasta(self, path):
        """Same format as the reference's k-mer FASTA (``>i\\nKMER\\n``)."""
        with open(path, "w") as fh:
            for i, s in enumerate(self.to_strings()):
                fh.write(">%d\n%s\n" % (i, s))

    @classmethod
    def from_strings(cls, eng, k, kmers):
        keys = []
        for s in kmers:
            v = key_of(canonicalize(s.upper()))
            if v is not None:
                keys.append(v)
        kw = eng.lib.kdf_key_words(k)
        lo, hi = eng.keys_to_device(keys, kw)
        return cls(eng, k, lo, hi)

    @classmethod
    def from_fasta(cls, eng, k, path):
        seqs = []
        with open(path) as fh:
            for line in fh:
                line = line.strip()
                if line and not line.startswith(">"):
                    seqs.append(line)
        return cls.from_strings(eng, k, seqs)

    def build_table(self, n_min=0):
        """Membership table primed with this set (``jellyfish count --if``
        priming / ``_build_proband_jf_index``)."""
        eng = self.engine
        t = eng.new_table(self.k, n_keys=max(len(self), n_min, 1))
        eng.update_keys(t, self.lo, self.hi, _engine.MODE_INSERT_ONLY, 0, 0)
        if eng.filter_applies(self.k, len(self)):   # final keys: front the table with a filter
            eng.build_filter(t, max(len(self), 1), eng.FILTER_MAX_BYTES)
        return t


class GpuKmerQuery:
    """Membership queries against a device k-mer table.

    Drop-in for ``JellyfishKmerQuery``: ``query_batch(list[str]) -> set[str]``,
    ``scan_read(seq, k) -> (set[str], set[int])``, ``close()``.
    """

    def __init__(self, kmer_set_or_table, k=None, engine=None):
        if isinstance(kmer_set_or_table, KmerSet):
            self.engine = kmer_set_or_table.engine
            self.k = kmer_set_or_table.k
            self.table = kmer_set_or_table.build_table()
            self._owns = True
        else:
            self.engine = engine or kmer_set_or_table.engine
            self.table = kmer_set_or_table
            self.k = k or kmer_set_or_table.k
            self._owns = False

    def query_batch(self, canonical_kmers):
        """Set of the given canonical k-mer strings present in the table."""
        if not canonical_kmers:
            return set()
        kmers = list(canonical_kmers)
        keys, idx = [], []
        for i, s in enumerate(kmers):
            v = key_of(s) if len(s) == self.k else None
            if v is not None:
                keys.append(v)
                idx.append(i)
        if not keys:
            return set()
        lo, hi = self.engine.keys_to_device(keys, self.table.key_words)
        found, _p0, _p1 = self.engine.lookup_keys(self.table, lo, hi, want_planes=False)
        f = found.cpu().numpy().astype(bool)
        return {kmers[i] for i, ok in zip(idx, f.tolist()) if ok}

    def scan_read(self, seq, kmer_size):
        """``(unique_in_read, kmer_hit_indices)`` for one read sequence."""
        if kmer_size != self.k:
            raise ValueError("table was built for k=%d" % self.k)
        if seq is None or len(seq) < kmer_size:
            return set(), set()
        hs = _engine.pack_sequences([seq])
        ds = self.engine.upload(hs)
        res = self.engine.scan_reads(self.table, ds, min_distinct=1)
        pos = res["hit_pos"].cpu().numpy().view(np.uint64).tolist()
        up = seq.upper()
        uniq = {canonicalize(up[p:p + kmer_size]) for p in pos}
        return uniq, set(int(p) for p in pos)

    def close(self):
        """The reference clears its per-object result cache here and keeps using
        the object; the device table has no cache, so this is a no-op."""

    def release(self):
        """Free an owned device table."""
        if self._owns and self.table is not None:
            self.table.close()
            self.table = None
