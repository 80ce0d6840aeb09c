"""CPU checks of the library's host helpers and of the *device* window-iterator
templates instantiated on the host (``kdf_debug_extract_host``): packing layout,
rolling canonical k-mers (64- and 128-bit), random-access extraction, validity.
No GPU needed; the oracle is the checker."""
import random
import re

import numpy as np
import pytest

from kmer_denovo_filter_b200 import engine
from oracle import kmers

KS = [1, 3, 5, 21, 31, 32, 33, 47, 63, 64]


def _rand_seqs(seed, n=30, maxlen=300, alphabet="ACGTACGTACGTACGTNacgtRY"):
    rng = random.Random(seed)
    return ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, maxlen))) for _ in range(n)]


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    hdr = open(engine.os.path.join(engine._HERE, "..", "include", "kdf.h")).read()
    declared = set(re.findall(r"\b(kdf_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(engine.exported_symbols())
    assert lib.kdf_version() == 2
    assert lib.kdf_key_words(31) == 1 and lib.kdf_key_words(33) == 2 and lib.kdf_key_words(65) == 0
    assert lib.kdf_table_bytes(12, 1) == 192 and lib.kdf_table_bytes(12, 2) == 288


def test_pack_layout_matches_oracle_stream():
    seqs = _rand_seqs(1)
    hs = engine.pack_sequences(seqs)
    codes, valid, starts, lens = kmers.encode_stream(seqs)
    assert hs.n_bases == codes.shape[0]
    assert hs.read_starts.tolist() == starts.tolist()
    assert hs.read_lens.tolist() == lens.tolist()
    n = hs.n_bases
    idx = np.arange(n)
    got_codes = (hs.codes[idx >> 5] >> (62 - 2 * (idx & 31)).astype(np.uint64)) & np.uint64(3)
    got_valid = (hs.valid[idx >> 5] >> (31 - (idx & 31)).astype(np.uint32)) & np.uint32(1)
    assert np.array_equal(got_codes.astype(np.uint8), codes)
    assert np.array_equal(got_valid.astype(bool), valid)


def test_pack_empty_inputs():
    hs = engine.pack_sequences([])
    assert hs.n_bases == 0 and hs.n_reads == 0
    hs = engine.pack_sequences(["", ""])
    assert hs.n_bases == 1 and hs.read_lens.tolist() == [0, 0]


@pytest.mark.parametrize("k", KS)
@pytest.mark.parametrize("random_access", [False, True, 2, 3])
def test_device_iterator_on_host_equals_oracle(k, random_access):
    seqs = _rand_seqs(100 + k)
    hs = engine.pack_sequences(seqs)
    lo, hi, ok = engine.debug_extract_host(hs, k, random_access=random_access)
    codes, valid, _s, _l = kmers.encode_stream(seqs)
    ohi, olo, ook = kmers.canonical_windows(codes, valid, k)
    n = ook.shape[0]
    assert not ok[n:].any()          # windows running past the end are invalid
    assert np.array_equal(ok[:n], ook)
    assert np.array_equal(lo[:n][ook], olo[ook])
    assert np.array_equal(hi[:n][ook], ohi[ook])


def test_engine_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.KdfError):
        engine.CudaEngine()


def test_invalid_positions_is_the_sparse_form_of_valid():
    """kdf_invalid_positions (and the list the BAM decoder attaches to a batch) names
    exactly the 0 bits of the validity bitmap inside n_bases, ascending: separators,
    N / IUPAC bases, nothing past the end of the stream."""
    for seed in range(6):
        seqs = _rand_seqs(seed, n=40) + ["", "N", "ACGT" * 8, "A" * 31, "C" * 33]
        hs = engine.pack_sequences(seqs)
        bits = np.unpackbits(np.ascontiguousarray(hs.valid).view(np.uint8).reshape(-1, 4)[:, ::-1].reshape(-1))
        want = np.flatnonzero(bits[:hs.n_bases] == 0).astype(np.uint32)
        assert np.array_equal(hs.invalid, want)
        assert np.array_equal(engine.invalid_positions(hs.valid, hs.n_bases), want)
        # rebuilding the bitmap from the list (what kdf_valid_from_invalid does on the device)
        rebuilt = np.ones(hs.n_words * 32, dtype=np.uint8)
        rebuilt[hs.n_bases:] = 0
        rebuilt[want] = 0
        assert np.array_equal(np.packbits(rebuilt).view(">u4").astype(np.uint32), hs.valid)
    empty = engine.pack_sequences([])
    assert empty.invalid is not None and empty.invalid.shape[0] == 0
