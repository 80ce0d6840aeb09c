"""Canonical k-mer semantics (oracle; test infrastructure only).

Two restatements of the same rules:

* string level — literal restatement of reference ``kmer_utils.py:15-38``
  (``reverse_complement``, ``canonicalize``) and ``kmer_utils.py:91-121``
  (``_extract_read_kmers``); slow, used on small cases and to pin the numeric
  level.
* numeric level — the 2-bit encoding A=0 C=1 G=2 T=3, first base most
  significant.  Under this code lexicographic string order equals unsigned
  integer order, so ``canonical = min(fwd, rc)``; it is also Jellyfish's on-disk
  key encoding (checked against ``mini_ref.fa.k31.jf`` in
  tests/test_oracle_golden.py).  Any base outside ``ACGTacgt`` invalidates every
  window that covers it (Jellyfish ``count``/``query`` break k-mers at non-ACGT;
  SURVEY §8 A2).

A "stream" is the layout shared by oracle and GPU engine: all sequences of a
batch concatenated, with exactly one invalid separator base between
consecutive sequences, so that a window is valid iff all ``k`` of its bases are
valid.
"""

import numpy as np

_COMP = str.maketrans("ACGTacgt", "TGCAtgca")
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
_LUT = np.full(256, 4, dtype=np.uint8)
for _c, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    _LUT[ord(_c)] = _v
    _LUT[ord(_c.lower())] = _v


# --------------------------------------------------------------------------
# string level (reference kmer_utils.py)
# --------------------------------------------------------------------------

def reverse_complement(seq):
    """reference ``kmer_utils.py:30-32``"""
    return seq.translate(_COMP)[::-1]


def canonicalize(kmer):
    """reference ``kmer_utils.py:35-38``"""
    rc = kmer.translate(_COMP)[::-1]
    return kmer if kmer < rc else rc


def extract_read_kmers(seq, k):
    """reference ``kmer_utils.py:91-121`` (``_extract_read_kmers``).

    Returns ``(canon_at_pos, unique_candidates)``.  Faithful to the reference:
    only windows containing ``N`` are skipped here.
    """
    n = len(seq)
    if n < k:
        return {}, []
    up = seq.upper()
    canon_at_pos = {}
    cands = []
    for i in range(n - k + 1):
        kmer = up[i:i + k]
        if "N" in kmer:
            continue
        c = canonicalize(kmer)
        canon_at_pos[i] = c
        cands.append(c)
    return canon_at_pos, list(dict.fromkeys(cands))


def is_acgt(kmer):
    return all(ch in "ACGT" for ch in kmer)


def count_canonical_strings(seqs, k):
    """``jellyfish count -m k -C`` over sequences, pure Python (small inputs)."""
    counts = {}
    for s in seqs:
        up = s.upper()
        for i in range(len(up) - k + 1):
            w = up[i:i + k]
            if not is_acgt(w):
                continue
            c = canonicalize(w)
            counts[c] = counts.get(c, 0) + 1
    return counts


# --------------------------------------------------------------------------
# numeric level
# --------------------------------------------------------------------------

def key_of(kmer):
    """2-bit integer key of an ACGT string (first base most significant)."""
    v = 0
    for ch in kmer:
        v = (v << 2) | _CODE[ch]
    return v


def kmer_of(key, k):
    """Inverse of :func:`key_of`."""
    return "".join("ACGT"[(key >> (2 * (k - 1 - i))) & 3] for i in range(k))


def encode_stream(seqs):
    """Concatenate sequences into ``(codes u8, valid bool, starts i64, lens i64)``.

    One invalid separator base (code 0, valid False) sits between consecutive
    sequences; none before the first or after the last.
    """
    n = len(seqs)
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=n)
    starts = np.zeros(n, dtype=np.int64)
    if n:
        starts[1:] = np.cumsum(lens[:-1] + 1)
    total = int(lens.sum() + max(n - 1, 0))
    raw = np.zeros(total, dtype=np.uint8)
    sep = np.zeros(total, dtype=bool)
    if n:
        joined = "\0".join(seqs).encode("latin-1")
        raw = np.frombuffer(joined, dtype=np.uint8)
    c = _LUT[raw] if total else np.zeros(0, dtype=np.uint8)
    valid = c < 4
    codes = np.where(valid, c, 0).astype(np.uint8)
    del sep
    return codes, valid, starts, lens


def window_valid(valid, k):
    """valid[p] for window p covering bases p..p+k-1."""
    n = valid.shape[0] - k + 1
    if n <= 0:
        return np.zeros(0, dtype=bool)
    bad = np.concatenate(([0], np.cumsum(~valid, dtype=np.int64)))
    return (bad[k:k + n] - bad[:n]) == 0


def canonical_windows(codes, valid, k):
    """Canonical key of every window of a stream.

    Returns ``(hi u64, lo u64, ok bool)`` each of length ``len(codes)-k+1``;
    the key is ``hi<<64 | lo`` (``hi`` is all zero for k <= 32).
    """
    if not (1 <= k <= 64):
        raise ValueError("k must be in 1..64")
    n = codes.shape[0] - k + 1
    if n <= 0:
        z = np.zeros(0, dtype=np.uint64)
        return z, z.copy(), np.zeros(0, dtype=bool)
    c64 = codes.astype(np.uint64)
    two = np.uint64(2)
    s62 = np.uint64(62)
    three = np.uint64(3)
    fhi = np.zeros(n, dtype=np.uint64)
    flo = np.zeros(n, dtype=np.uint64)
    rhi = np.zeros(n, dtype=np.uint64)
    rlo = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        cj = c64[j:j + n]
        fhi = (fhi << two) | (flo >> s62)
        flo = (flo << two) | cj
        cr = three - c64[k - 1 - j:k - 1 - j + n]
        rhi = (rhi << two) | (rlo >> s62)
        rlo = (rlo << two) | cr
    if k <= 32:
        fhi[:] = 0
        rhi[:] = 0
        if k < 32:
            m = np.uint64((1 << (2 * k)) - 1)
            flo &= m
            rlo &= m
    elif k < 64:
        m = np.uint64((1 << (2 * k - 64)) - 1)
        fhi &= m
        rhi &= m
    fwd_smaller = (fhi < rhi) | ((fhi == rhi) & (flo <= rlo))
    hi = np.where(fwd_smaller, fhi, rhi)
    lo = np.where(fwd_smaller, flo, rlo)
    return hi, lo, window_valid(valid, k)


def to_pyints(hi, lo):
    """Combine hi/lo arrays into Python ints."""
    return [(int(h) << 64) | int(l) for h, l in zip(hi.tolist(), lo.tolist())]


def count_stream(codes, valid, k):
    """Exact canonical multiset count → ``dict{int key: int count}``.

    Restates ``jellyfish count -m k -C`` (reference
    ``discovery/pipeline.py:114-122``; ``core/jellyfish_wrappers.py:313-321``).
    """
    hi, lo, ok = canonical_windows(codes, valid, k)
    hi = hi[ok]
    lo = lo[ok]
    if hi.size == 0:
        return {}
    if k <= 32:
        u, c = np.unique(lo, return_counts=True)
        return dict(zip(u.tolist(), c.tolist()))
    order = np.lexsort((lo, hi))
    hi = hi[order]
    lo = lo[order]
    new = np.ones(hi.size, dtype=bool)
    new[1:] = (hi[1:] != hi[:-1]) | (lo[1:] != lo[:-1])
    idx = np.flatnonzero(new)
    cnt = np.diff(np.append(idx, hi.size))
    keys = to_pyints(hi[idx], lo[idx])
    return dict(zip(keys, cnt.tolist()))


def count_sequences(seqs, k):
    """``jellyfish count -C`` of an iterable of sequence strings."""
    codes, valid, _s, _l = encode_stream(list(seqs))
    return count_stream(codes, valid, k)


def count_stream_filtered(codes, valid, k, filter_keys):
    """``jellyfish count -C --if filter`` (reference
    ``core/jellyfish_wrappers.py:167-176``): only keys already in the filter
    are counted.  Returns ``dict{key: count}`` for *every* filter key
    (count 0 when never seen), mirroring what ``jellyfish query`` reports.
    """
    full = count_stream(codes, valid, k)
    return {key: full.get(key, 0) for key in filter_keys}


def read_jf_binary_sorted(path):
    """Decode a Jellyfish ``binary/sorted`` file → ``(k, dict{key: count})``.

    Layout (decoded from the reference fixture ``mini_ref.fa.k31.jf``): nine
    ASCII digits = JSON header length, JSON header (``key_len`` bits,
    ``counter_len`` bytes), NUL padding, then fixed-width little-endian
    records of ceil(key_len/8) key bytes + counter_len count bytes.
    """
    import json
    with open(path, "rb") as fh:
        data = fh.read()
    hlen = int(data[:9].decode())
    hdr = json.loads(data[9:9 + hlen].decode().rstrip("\0"))
    key_bits = hdr["key_len"]
    cbytes = hdr["counter_len"]
    kbytes = (key_bits + 7) // 8
    off = 9 + hlen
    rec = kbytes + cbytes
    body = data[off:]
    n = len(body) // rec
    out = {}
    for i in range(n):
        r = body[i * rec:(i + 1) * rec]
        out[int.from_bytes(r[:kbytes], "little")] = int.from_bytes(r[kbytes:], "little")
    return key_bits // 2, out
