"""Randomised corruption of a BAM: flip bytes of the UNCOMPRESSED stream (record headers, field
sizes, sequence, tags) or of the compressed file (BGZF headers, payload, CRC, ISIZE) and decode.
The decoder must either succeed or raise KdfError — never crash, hang or run away with memory.
    python scripts/fuzz_bam_corrupt.py tests/golden/giab/HG004_mother.bam [seed] [seconds]
Each trial runs in a child process with an address-space limit, so a segfault or an
allocation bomb shows up as a failed trial instead of taking the fuzzer down."""
import gzip
import os
import random
import resource
import struct
import subprocess
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import resource, sys
resource.setrlimit(resource.RLIMIT_AS, (8 << 30, 8 << 30))
sys.path.insert(0, %r)
from kmer_denovo_filter_b200 import bamio, engine
try:
    with bamio.BamReader(sys.argv[1], threads=2) as rd:
        for mode, meta in ((bamio.MODE_FASTA, 0), (bamio.MODE_ALL, 3)):
            pass
        n = 0
        for b in rd.batches(int(sys.argv[2]), max_bases=200000, want_meta=int(sys.argv[3])):
            n += b.n_reads
            b.close()
    print("ok", n)
except engine.KdfError as e:
    print("kdferror", str(e)[:80])
except MemoryError:
    print("memoryerror")
''' % ROOT


def bgzf(payload_blocks):
    out = bytearray()
    for p in payload_blocks:
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        c = co.compress(p) + co.flush()
        out += struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(c) + 25)
        out += c + struct.pack("<II", zlib.crc32(p) & 0xFFFFFFFF, len(p))
    out += bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    return bytes(out)


def main():
    path = sys.argv[1]
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = gzip.open(path, "rb").read()
    comp = open(path, "rb").read()
    # where the records start (after the header), to aim half of the mutations at record headers
    l_text = struct.unpack_from("<i", raw, 4)[0]
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, off)[0]
    off += 4
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", raw, off)[0]
        off += 8 + ln
    rec_starts = []
    o = off
    while o + 4 <= len(raw) and len(rec_starts) < 5000:
        rec_starts.append(o)
        o += 4 + struct.unpack_from("<i", raw, o)[0]
    t0, trials, outcomes = time.time(), 0, {}
    tmp = tempfile.mkdtemp(prefix="kdf_fuzz_")
    p = os.path.join(tmp, "x.bam")
    while time.time() - t0 < seconds:
        kind = rng.choice(["header_field", "record_byte", "compressed_byte", "truncate"])
        if kind in ("header_field", "record_byte"):
            data = bytearray(raw)
            for _ in range(rng.randint(1, 3)):
                r = rng.choice(rec_starts)
                if kind == "header_field":
                    field = rng.choice([(0, "<i"), (12, "<B"), (16, "<H"), (20, "<i"), (4, "<i")])
                    val = rng.choice([0, -1, 1, 255, 65535, 0x7fffffff, 0x7fffff00, -0x80000000, rng.randint(-10**9, 10**9)])
                    size = struct.calcsize(field[1])
                    mask = (1 << (8 * size)) - 1
                    data[r + field[0]:r + field[0] + size] = (val & mask).to_bytes(size, "little")
                else:
                    q = min(len(data) - 1, r + rng.randint(0, 400))
                    data[q] = rng.randint(0, 255)
            blocks = [bytes(data[i:i + 60000]) for i in range(0, len(data), 60000)]
            open(p, "wb").write(bgzf(blocks))
        elif kind == "compressed_byte":
            data = bytearray(comp)
            for _ in range(rng.randint(1, 4)):
                data[rng.randrange(len(data))] = rng.randint(0, 255)
            open(p, "wb").write(data)
        else:
            open(p, "wb").write(comp[:rng.randrange(1, len(comp))])
        mode, meta = rng.choice([(0, 0), (1, 1), (2, 2), (2, 3)])
        # KDF_FUZZ_CMD: an AddressSanitizer build of scripts/asan_bam_decode.cpp decodes the file
        # instead of the Python child (same "ok" / "kdferror" protocol), with a random chunk
        # size, headroom and thread count
        cmd = [sys.executable, "-c", CHILD, p, str(mode), str(meta)]
        env = dict(os.environ)
        if os.environ.get("KDF_FUZZ_CMD"):
            cmd = [os.environ["KDF_FUZZ_CMD"], p, str(mode), str(meta), str(rng.choice([1, 2, 3, 8])),
                   str(rng.choice([0, 50_000, 1_000_000]))]
            env["ASAN_OPTIONS"] = "detect_leaks=0"
            if rng.random() < 0.7:
                env["KDF_BAM_CHUNK_KB"] = str(rng.choice([64, 65, 100, 300, 1000]))
                env["KDF_BAM_GAP"] = str(rng.choice([0, 1, 17, 300, 5000]))
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=env)
        except subprocess.TimeoutExpired:
            print("HANG", kind, mode, meta)
            sys.exit(1)
        word = (r.stdout.split() or ["crash"])[0]
        if r.returncode != 0 or word not in ("ok", "kdferror", "memoryerror"):
            keep = os.path.join(tempfile.gettempdir(), "kdf_fuzz_failure.bam")
            os.replace(p, keep)
            print("FAILURE rc=%d kind=%s mode=%d meta=%d file=%s\n%s" % (r.returncode, kind, mode, meta, keep, r.stderr[-500:]))
            sys.exit(1)
        outcomes[(kind, word)] = outcomes.get((kind, word), 0) + 1
        trials += 1
    print("%d trials, no crash:" % trials, dict(sorted(outcomes.items())))


if __name__ == "__main__":
    main()
