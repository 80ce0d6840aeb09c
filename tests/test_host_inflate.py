"""The library's own DEFLATE decoder and CRC-32 (csrc/kdf_inflate.cpp) against zlib, block by
block through the C ABI (kdf_bgzf_inflate_block / kdf_crc32): every block type, the code
shapes zlib's strategies produce, real BAM blocks, and corrupted streams (which must be
rejected — or, without the CRC check, decode to exactly what zlib decodes them to)."""
import ctypes
import random
import struct
import zlib

import numpy as np
import pytest

from kmer_denovo_filter_b200 import engine


def _bgzf_block(payload, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, raw=None):
    if raw is None:
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        raw = co.compress(payload) + co.flush()
    head = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(raw) + 25)
    return head + raw + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


def _inflate(block, usize, verify_crc=1, impl=0):
    lib = engine.load_library()
    src = np.frombuffer(block, dtype=np.uint8)
    # the output sits between two guard zones: nothing may be written outside [0, usize)
    out = np.full(usize + 128, 0xA5, dtype=np.uint8)
    rc = lib.kdf_bgzf_inflate_block(src.ctypes.data_as(ctypes.c_void_p), len(block),
                                    out[64:].ctypes.data_as(ctypes.c_void_p), usize, verify_crc, impl)
    assert (out[:64] == 0xA5).all() and (out[64 + usize:] == 0xA5).all(), "wrote outside the output"
    return rc, out[64:64 + usize].tobytes()


def _payloads():
    rng = random.Random(5)
    yield b""
    yield b"A"
    yield b"ACGT" * 16000                                  # long matches, short distance
    yield bytes(65280)                                     # one value: distance-1 runs
    yield bytes(rng.getrandbits(8) for _ in range(65280))  # incompressible -> stored blocks
    yield bytes(rng.choice(b"ACGTN") for _ in range(60000))
    yield bytes(min(255, max(0, int(rng.gauss(70, 12)))) for _ in range(65000))   # quality-like literals
    yield b"".join(b"read%07d/1\tchr1\t%d\n" % (i, rng.randrange(1 << 28)) for i in range(2000))
    # > 32 KiB apart repeats (window edge) and every byte value
    blob = bytes(rng.getrandbits(8) for _ in range(20000))
    yield blob + bytes(range(256)) * 40 + blob
    yield bytes(rng.getrandbits(8) for _ in range(300))    # shorter than the fast-path margins
    # geometric symbol frequencies: Huffman codes up to the 15-bit limit (sub-tables of both tables)
    yield bytes(min(int(rng.expovariate(0.55)), 255) for _ in range(65000))
    yield b"".join(bytes([min(int(rng.expovariate(0.4)), 255)]) * rng.choice((1, 1, 1, 3, 40, 258, 300))
                   for _ in range(6000))[:65000]


@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE,
                                      zlib.Z_FILTERED])
def test_blocks_of_every_kind_match_zlib(strategy):
    for payload in _payloads():
        for level in (0, 1, 6, 9):
            blk = _bgzf_block(payload, level, strategy)
            rc, got = _inflate(blk, len(payload))
            assert rc == 0, (len(payload), level, strategy, engine.load_library().kdf_host_last_error())
            assert got == payload
            rc1, got1 = _inflate(blk, len(payload), impl=1)
            assert rc1 == 0 and got1 == payload


def test_many_deflate_blocks_in_one_bgzf_block():
    rng = random.Random(11)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts, raw = [], b""
    for i in range(40):     # Z_FULL_FLUSH ends a deflate block (and adds an empty stored one)
        p = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 1500)))
        parts.append(p)
        raw += co.compress(p) + co.flush(zlib.Z_FULL_FLUSH if i % 2 else zlib.Z_SYNC_FLUSH)
    raw += co.flush()
    payload = b"".join(parts)
    rc, got = _inflate(_bgzf_block(payload, raw=raw), len(payload))
    assert rc == 0 and got == payload


def test_real_bam_blocks_match_zlib(giab_paths):
    data = open(giab_paths["child"], "rb").read()
    p = n = 0
    while p < len(data):
        bsize = struct.unpack_from("<H", data, p + 16)[0] + 1
        isize = struct.unpack_from("<I", data, p + bsize - 4)[0]
        blk = data[p:p + bsize]
        want = zlib.decompress(blk[18:-8], -15)
        rc, got = _inflate(blk, isize)
        assert rc == 0 and got == want
        p += bsize
        n += 1
    assert n > 20


def test_wrong_size_crc_and_truncation_are_rejected():
    payload = b"".join(b"line %d of the payload\n" % i for i in range(2000))
    blk = _bgzf_block(payload)
    assert _inflate(blk, len(payload))[0] == 0
    assert _inflate(blk, len(payload) - 1)[0] != 0          # the stream holds more than ISIZE says
    assert _inflate(blk, len(payload) + 1)[0] != 0          # ... and less
    bad_crc = blk[:-8] + struct.pack("<I", (zlib.crc32(payload) ^ 1) & 0xFFFFFFFF) + blk[-4:]
    assert _inflate(bad_crc, len(payload))[0] != 0
    assert _inflate(bad_crc, len(payload), verify_crc=0) == (0, payload)
    for cut in (1, 2, 7, 100, len(blk) - 40):
        raw = blk[18:-8][:-cut]
        assert _inflate(_bgzf_block(payload, raw=raw), len(payload))[0] != 0
    assert _inflate(blk[:20], 0)[0] != 0                    # shorter than any BGZF block


def test_corrupted_streams_agree_with_zlib():
    """Bit flips in the compressed bytes, CRC check off: either both decoders reject the
    stream or both produce the same bytes (and nothing is written outside the output)."""
    rng = random.Random(23)
    payloads = [p for p in _payloads() if len(p) > 1000]
    accepted = rejected = 0
    for trial in range(1500):
        payload = payloads[trial % len(payloads)]
        co = zlib.compressobj(rng.choice((1, 6)), zlib.DEFLATED, -15)
        raw = bytearray(co.compress(payload) + co.flush())
        for _ in range(rng.randrange(1, 4)):
            at = rng.randrange(min(len(raw), 60)) if rng.random() < 0.5 else rng.randrange(len(raw))
            raw[at] ^= 1 << rng.randrange(8)
        blk = _bgzf_block(payload, raw=bytes(raw))
        rc, got = _inflate(blk, len(payload), verify_crc=0)
        try:
            d = zlib.decompressobj(-15)
            want = d.decompress(bytes(raw))
            zok = d.eof and len(want) == len(payload)
        except zlib.error:
            zok = False
        assert (rc == 0) == zok, trial
        if zok:
            assert got == want
            accepted += 1
        else:
            rejected += 1
    assert accepted > 20 and rejected > 500


def test_crc32_matches_zlib():
    lib = engine.load_library()
    rng = np.random.default_rng(3)
    buf = rng.integers(0, 256, size=70016, dtype=np.uint8)
    for n in list(range(0, 200)) + [255, 256, 1023, 4096, 65280, 65536, 69999]:
        for shift in (0, 1, 7):
            a = buf[shift:shift + n]
            got = lib.kdf_crc32(a.ctypes.data_as(ctypes.c_void_p), n)
            assert got == (zlib.crc32(a.tobytes()) & 0xFFFFFFFF), (n, shift)
