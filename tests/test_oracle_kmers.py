"""Known-answer tests for the canonical k-mer semantics.

Vectors follow the reference's own unit tests (reference
``tests/test_kmer_utils.py:14-44`` canonical / reverse-complement,
``:537-584`` ``_extract_read_kmers``), plus string-vs-numeric equivalence.
"""
import random

import pytest

from oracle import kmers


def test_reverse_complement_known_answers():
    assert kmers.reverse_complement("ACGT") == "ACGT"
    assert kmers.reverse_complement("AAAA") == "TTTT"
    assert kmers.reverse_complement("ATCG") == "CGAT"
    assert kmers.reverse_complement("GATTACA") == "TGTAATC"
    assert kmers.reverse_complement("acgt") == "acgt"


def test_canonicalize_known_answers():
    assert kmers.canonicalize("AAAA") == "AAAA"
    assert kmers.canonicalize("TTTT") == "AAAA"
    assert kmers.canonicalize("ACGT") == "ACGT"
    assert kmers.canonicalize("CGAT") == "ATCG"
    assert kmers.canonicalize("TGTAATC") == "GATTACA"
    for s in ("GATTACA", "CCCGGGA", "TTGCA"):
        assert kmers.canonicalize(s) == kmers.canonicalize(kmers.reverse_complement(s))


def test_extract_read_kmers_reference_semantics():
    cap, uniq = kmers.extract_read_kmers("ACGTACGT", 4)
    assert set(cap) == {0, 1, 2, 3, 4}
    assert uniq == list(dict.fromkeys(cap[i] for i in range(5)))
    cap, uniq = kmers.extract_read_kmers("ACGNACGT", 4)
    assert set(cap) == {4}
    assert kmers.extract_read_kmers("ACG", 4) == ({}, [])
    cap, _ = kmers.extract_read_kmers("acgtacgt", 4)
    assert cap[0] == "ACGT"


def test_key_roundtrip_and_order():
    rng = random.Random(7)
    for k in (5, 21, 31, 33, 47, 63):
        ks = ["".join(rng.choice("ACGT") for _ in range(k)) for _ in range(50)]
        keys = [kmers.key_of(s) for s in ks]
        assert [kmers.kmer_of(x, k) for x in keys] == ks
        assert sorted(ks) == [kmers.kmer_of(x, k) for x in sorted(keys)]


@pytest.mark.parametrize("k", [3, 5, 21, 31, 32, 33, 47, 63, 64])
def test_numeric_count_equals_string_count(k):
    rng = random.Random(k)
    seqs = []
    for _ in range(40):
        n = rng.randint(0, 150)
        seqs.append("".join(rng.choice("ACGTACGTACGTNacgtR") for _ in range(n)))
    want = kmers.count_canonical_strings(seqs, k)
    got = kmers.count_sequences(seqs, k)
    assert {kmers.kmer_of(key, k): c for key, c in got.items()} == want


def test_stream_separator_breaks_windows():
    codes, valid, starts, lens = kmers.encode_stream(["ACGT", "ACGT"])
    assert codes.shape[0] == 9 and not valid[4]
    assert starts.tolist() == [0, 5] and lens.tolist() == [4, 4]
    hi, lo, ok = kmers.canonical_windows(codes, valid, 3)
    assert ok.tolist() == [True, True, False, False, False, True, True]
    assert kmers.count_sequences(["ACGT", "ACGT"], 3) == kmers.count_sequences(["ACGT"] * 2, 3)
    assert kmers.count_sequences([], 3) == {}
    assert kmers.count_sequences(["AC"], 3) == {}
