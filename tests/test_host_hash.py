"""Quality of the device hash functions, run on the host through the test hook
(same code as the kernels): partition, bucket and owner must each be uniform
and pairwise independent on canonical k-mers of a random genome, otherwise
bins overflow and buckets cluster."""
import random

import numpy as np
import pytest

from kmer_denovo_filter_b200 import engine
from oracle import kmers


def _canonical_keys(k, n_bases=200_000, seed=5):
    rng = random.Random(seed + k)
    g = "".join(rng.choice("ACGT") for _ in range(n_bases))
    codes, valid, _s, _l = kmers.encode_stream([g])
    hi, lo, ok = kmers.canonical_windows(codes, valid, k)
    return lo[ok], hi[ok]


def _chi2_ok(counts):
    exp = counts.sum() / counts.shape[0]
    chi2 = ((counts - exp) ** 2 / exp).sum()
    dof = counts.shape[0] - 1
    return chi2 < dof + 6 * np.sqrt(2 * dof) + 10


@pytest.mark.parametrize("k", [21, 31, 32, 33, 47, 63])
def test_partition_bucket_owner_uniform(k):
    lo, hi = _canonical_keys(k)
    kw = 1 if k <= 32 else 2
    part, bucket, owner = engine.debug_hash_host(lo, hi if kw == 2 else None, kw, 7, 4096, 8)
    assert part.max() < 128 and bucket.max() < 4096 and owner.max() < 8
    assert _chi2_ok(np.bincount(part, minlength=128).astype(np.float64))
    assert _chi2_ok(np.bincount(bucket, minlength=4096).astype(np.float64))
    assert _chi2_ok(np.bincount(owner, minlength=8).astype(np.float64))
    # pairwise independence: joint histograms are uniform too
    assert _chi2_ok(np.bincount(part.astype(np.int64) * 8 + owner, minlength=1024).astype(np.float64))
    assert _chi2_ok(np.bincount((bucket.astype(np.int64) & 127) * 128 + part,
                                minlength=128 * 128).astype(np.float64))


def test_owner_counts_for_non_power_of_two_ranks():
    lo, _hi = _canonical_keys(31)
    for n in (1, 3, 5, 6):
        _p, _b, owner = engine.debug_hash_host(lo, None, 1, 0, 1024, n)
        assert owner.max() < n
        assert _chi2_ok(np.bincount(owner, minlength=n).astype(np.float64))


def test_low_entropy_keys_do_not_collapse():
    """Homopolymer / dinucleotide-repeat neighbourhoods (real genomes have them)."""
    seqs = []
    for unit in ("A", "AC", "AG", "AAT", "ACG", "AAAC"):
        base = (unit * 400)[:400]
        for i in range(0, 400, 7):
            s = list(base)
            s[i] = "G" if s[i] != "G" else "T"
            seqs.append("".join(s))
    codes, valid, _s, _l = kmers.encode_stream(seqs)
    hi, lo, ok = kmers.canonical_windows(codes, valid, 31)
    keys = np.unique(lo[ok])
    _p, bucket, _o = engine.debug_hash_host(keys, None, 1, 0, 1 << 16, 1)
    # distinct keys spread: no bucket holds more than a handful
    assert np.bincount(bucket).max() <= 6


def test_filter_words_and_partition_plans():
    """Host-side sizing: the two-bit filter takes 32 bits per key, 16 or 8 when only that
    fits the budget, none beyond; the packed count plans half as many hash ranges as the plane
    form for the same L2 slice."""
    from kmer_denovo_filter_b200.discovery import kmer_chain
    fw = engine.CudaEngine.filter_words
    mb = 1 << 20
    assert fw(1, 32 * mb) == 1024
    assert fw(3_893_677, 32 * mb) * 4 == 16 * mb          # 32 bits per key
    assert fw(7_790_648, 32 * mb) * 4 == 32 * mb
    assert fw(15_602_227, 32 * mb) * 4 == 32 * mb         # 16 bits per key
    assert fw(31_243_893, 32 * mb) * 4 == 32 * mb         # 8 bits per key (the 8-GPU filter set)
    assert fw(70_000_000, 32 * mb) == 0                   # no filter: the binned route
    for n in (1, 1000, 5_000_000):
        w = fw(n, 1 << 30)
        assert w & (w - 1) == 0 and w >= min(n, 1024)
    n_win = 1_920_000_000
    p_plane, c_plane = kmer_chain.plan_partitions(n_win, key_words=1, packed=False)
    p_pack, c_pack = kmer_chain.plan_partitions(n_win, key_words=1, packed=True)
    assert (p_plane, p_pack) == (64, 32)
    assert c_pack * 8 <= kmer_chain.SLICE_BYTES and c_plane * 16 <= kmer_chain.SLICE_BYTES
    assert p_pack * c_pack >= n_win // 8 and c_pack % 4 == 0
    assert kmer_chain.plan_partitions(100, key_words=2, packed=True)[0] == 1
