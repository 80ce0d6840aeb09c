"""kdf_fasta_pack (the reference FASTA packed by all host threads) against the line-by-line
Python reader + kdf_pack_sequences it replaces on the reference-index path."""
import glob
import gzip
import os

import numpy as np
import pytest

from kmer_denovo_filter_b200 import bamio, engine


def _same(path, threads):
    _names, seqs = bamio.read_fasta_sequences(path)
    seqs = [s.replace(b" ", b"").replace(b"\t", b"") for s in seqs]
    got, n = engine.pack_fasta_file(path, threads)
    assert n == len(seqs)
    if not seqs:
        assert got.n_bases == 0
        return got
    want = engine.pack_sequences(seqs)
    assert got.n_bases == want.n_bases
    for name in ("codes", "valid", "read_starts", "read_lens", "invalid"):
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
    return got


@pytest.mark.parametrize("text", [
    b"", b">a\n", b">a", b">a\nACGT\n>b\n\n>c\nNNAC\nGT", b"junk before\n>a desc\nAC GT\r\nacgtn\n>b\nA",
    b">x\nA>C\nG\n", b">only\n" + b"ACGTACGTACGTACGTACGTACGTACGTACGTA\n" * 3, b">a\n\n\n>b\nT\n\n",
])
@pytest.mark.parametrize("threads", [1, 3])
def test_small_cases(tmp_path, text, threads):
    p = tmp_path / "x.fa"
    p.write_bytes(text)
    _same(str(p), threads)


def test_many_records_and_block_boundaries(tmp_path):
    """Records longer than the 1 MiB work blocks, lines of odd widths, lower case, Ns."""
    rng = np.random.default_rng(4)
    parts = []
    for i, (n, width) in enumerate(((2_500_000, 61), (33, 7), (1_048_576 + 5, 1000), (0, 60), (70_001, 50))):
        seq = rng.choice(np.frombuffer(b"ACGTacgtNn", dtype=np.uint8), size=n,
                         p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .02]).tobytes()
        lines = b"\n".join(seq[j:j + width] for j in range(0, n, width))
        parts.append(b">chr%d some description\n" % i + lines + (b"\n" if i % 2 == 0 else b"\r\n"))
    p = tmp_path / "big.fa"
    p.write_bytes(b"".join(parts))
    hs = _same(str(p), 4)
    assert hs.read_lens.tolist() == [2_500_000, 33, 1_048_576 + 5, 0, 70_001]
    gz = tmp_path / "big.fa.gz"
    gz.write_bytes(gzip.compress(p.read_bytes(), 1))
    hz, _n = engine.pack_fasta_file(str(gz), 2)
    assert np.array_equal(hz.codes, hs.codes) and np.array_equal(hz.valid, hs.valid)


def test_golden_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    fas = sorted(glob.glob(os.path.join(here, "golden", "**", "*.fa"), recursive=True))
    assert fas
    for fa in fas[:4]:
        _same(fa, 4)
