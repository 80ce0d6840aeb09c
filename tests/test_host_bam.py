"""The library's C++ BGZF/BAM decoder against the oracle's stdlib reader on the
GIAB fixtures: record streams (samtools-fasta semantics, scan semantics),
packed bases, metadata, batching."""
import numpy as np
import pytest

from kmer_denovo_filter_b200 import bamio, engine
from oracle import bam as obam


def _decode(batch, i):
    return batch.record(i).query_sequence if batch.has_meta else bamio.Record(batch, i).query_sequence


@pytest.mark.parametrize("who", ["child", "mother", "father"])
def test_fasta_stream_matches_oracle(giab_paths, giab_records, who):
    want = obam.fasta_stream(giab_records[who])
    with bamio.BamReader(giab_paths[who], threads=4) as rd:
        batches = list(rd.batches(bamio.MODE_FASTA, want_meta=True))
    assert len(batches) == 1
    b = batches[0]
    assert b.n_reads == len(want)
    assert [b.record(i).query_name for i in range(0, b.n_reads, 97)] == \
        [r.qname for r in want[::97]]
    hs = engine.pack_sequences([r.seq for r in want])
    assert b.n_bases == hs.n_bases
    assert np.array_equal(b.codes, hs.codes) and np.array_equal(b.valid, hs.valid)
    assert np.array_equal(b.read_starts, hs.read_starts)
    assert np.array_equal(b.read_lens, hs.read_lens)
    assert np.array_equal(b.invalid, hs.invalid)     # the decoder's sparse form of `valid`


def test_scan_stream_and_metadata(giab_paths, giab_records):
    recs = giab_records["child"]
    want = [(i, r) for i, r in enumerate(recs) if not (r.flag & 0x500)]
    with bamio.BamReader(giab_paths["child"], threads=2) as rd:
        assert rd.references == giab_records["child_refs"][0]
        assert rd.lengths == giab_records["child_refs"][1]
        b = rd.next_batch(bamio.MODE_SCAN, want_meta=True)
    assert b.at_eof and b.n_reads == len(want)
    assert b.rec_index.tolist() == [i for i, _r in want]
    for j in list(range(0, len(want), 53)) + [len(want) - 1]:
        o = want[j][1]
        m = b.record(j)
        assert m.query_name == o.qname and m.flag == o.flag
        assert m.reference_name == o.reference_name
        assert m.reference_start == o.pos and m.reference_end == o.reference_end
        assert m.mapping_quality == o.mapq
        assert (m.cigartuples or []) == (o.cigar or [])
        assert m.has_tag("SA") == o.has_tag("SA")
        if o.has_tag("SA"):
            assert m.get_tag("SA") == o.get_tag("SA")
        seq = o.seq.upper()
        assert m.query_sequence == "".join(c if c in "ACGT" else "N" for c in seq)
        assert m.get_aligned_pairs() == o.get_aligned_pairs(matches_only=True)
    n_sa = sum(1 for _i, r in want if r.has_tag("SA"))
    assert int((b.sa_off[1:] > b.sa_off[:-1]).sum()) == n_sa


@pytest.mark.parametrize("mode", [bamio.MODE_FASTA, bamio.MODE_SCAN, bamio.MODE_ALL])
def test_batching_is_equivalent_to_one_shot(giab_paths, mode):
    with bamio.BamReader(giab_paths["mother"], threads=3) as rd:
        whole = rd.next_batch(mode, want_meta=True)
    names, lens, idx = [], [], []
    with bamio.BamReader(giab_paths["mother"], threads=3) as rd:
        nb = 0
        for b in rd.batches(mode, max_bases=200_000, want_meta=True):
            nb += 1
            assert b.n_bases <= 200_000
            lens += b.read_lens.tolist()
            idx += b.rec_index.tolist()
            names += [b.record(i).query_name for i in (0, b.n_reads - 1)]
    assert nb > 5
    assert lens == whole.read_lens.tolist()
    assert idx == whole.rec_index.tolist()


def test_synthetic_bam_roundtrip(tmp_path):
    """Writer (oracle) → C++ reader: flags, collapse, odd lengths, IUPAC, empty seq."""
    recs = [
        obam.encode_record(0, 10, "q1", 0x41, 60, [(0, 7)], "ACGTNAC"),
        obam.encode_record(0, 10, "q1", 0x41, 60, [(0, 7)], "TTTTTTT"),      # collapsed (same part)
        obam.encode_record(0, 12, "q1", 0x81, 60, [(0, 5)], "GGRCC"),        # other part: kept
        obam.encode_record(0, 20, "q2", 0x100, 0, [(0, 4)], "AAAA"),         # secondary
        obam.encode_record(0, 21, "q3", 0x400, 0, [(0, 4)], "CCCC"),         # duplicate
        obam.encode_record(0, 22, "q4", 0x800, 9, [(4, 2), (0, 2)], "GGTT",
                           tags=b"SAZchr1,5,+,4M,60,0;\0"),                  # supplementary
        obam.encode_record(-1, -1, "q5", 0x4, 0, [], "ACGTACGTA"),           # unmapped
        obam.encode_record(0, 30, "q6", 0, 60, [], ""),                      # no sequence
    ]
    p = str(tmp_path / "t.bam")
    obam.write_bam(p, ["chr1"], [1000], recs)
    with bamio.BamReader(p) as rd:
        f = rd.next_batch(bamio.MODE_FASTA, want_meta=True)
    assert [f.record(i).query_name for i in range(f.n_reads)] == ["q1", "q1", "q5", "q6"]
    assert [f.record(i).query_sequence for i in range(f.n_reads)] == ["ACGTNAC", "GGNCC", "ACGTACGTA", None]
    with bamio.BamReader(p) as rd:
        s = rd.next_batch(bamio.MODE_SCAN, want_meta=True)
    assert [s.record(i).query_name for i in range(s.n_reads)] == ["q1", "q1", "q1", "q4", "q5", "q6"]
    assert s.record(3).get_tag("SA") == "chr1,5,+,4M,60,0;"
    assert s.record(3).is_supplementary and s.record(4).is_unmapped
    assert s.record(3).get_aligned_pairs() == [(2, 22), (3, 23)]
    assert [s.record(i).has_tag("SA") for i in range(s.n_reads)] == [False] * 3 + [True, False, False]
    # qualities (want_meta=2) and the raw records (want_meta=3), several decoder threads
    with bamio.BamReader(p, threads=3) as rd:
        a = rd.next_batch(bamio.MODE_ALL, want_meta=3)
    assert a.n_reads == len(recs)
    for i, raw in enumerate(recs):
        assert bytes(a.raw_blob[int(a.raw_off[i]):int(a.raw_off[i + 1])]) == raw[4:]   # after block_size
        l_seq = int(a.read_lens[i])
        assert int(a.qual_off[i + 1] - a.qual_off[i]) == l_seq
    assert [a.record(i).cigartuples for i in (0, 5, 6)] == [[(0, 7)], [(4, 2), (0, 2)], None]


def test_not_a_bam(tmp_path):
    p = tmp_path / "x.bam"
    p.write_bytes(b"hello world, definitely not bgzf")
    with pytest.raises(engine.KdfError):
        bamio.BamReader(str(p))
    with pytest.raises(engine.KdfError):
        c = tmp_path / "x.cram"
        c.write_bytes(b"CRAM")
        bamio.BamReader(str(c))


@pytest.mark.parametrize("chunk_kb,gap", [(64, 1 << 20), (64, 16), (200, 0)])
def test_decoder_pipeline_chunking_is_invisible(giab_paths, monkeypatch, chunk_kb, gap):
    """The decoder's read / inflate / parse pipeline works chunk by chunk; records that
    straddle chunks are completed in the next chunk's headroom, or — when the unparsed
    tail is larger than the headroom — by rebuilding that chunk.  Small chunks and tiny
    headrooms exercise every transition on the fixture: same batches as the defaults."""
    def decode(mode, max_bases):
        out = []
        with bamio.BamReader(giab_paths["child"], threads=3) as rd:
            for b in rd.batches(mode, max_bases=max_bases, want_meta=True):
                out.append((b.n_reads, b.n_bases, b.codes.copy(), b.valid.copy(), b.invalid.copy(),
                            b.rec_index.copy(), b.pos.copy(), b.cigar_blob.copy(), bytes(b.qname_blob)))
                b.close()
        return out
    want = {(m, mb): decode(m, mb) for m in (bamio.MODE_FASTA, bamio.MODE_SCAN) for mb in (0, 300_000)}
    monkeypatch.setenv("KDF_BAM_CHUNK_KB", str(chunk_kb))
    monkeypatch.setenv("KDF_BAM_GAP", str(gap))
    for key, w in want.items():
        got = decode(*key)
        assert len(got) == len(w)
        for a, b in zip(got, w):
            assert a[0] == b[0] and a[1] == b[1] and a[8] == b[8]
            for x, y in zip(a[2:8], b[2:8]):
                assert np.array_equal(x, y)


def test_truncated_bam_is_an_error(giab_paths, tmp_path):
    """A file cut in the middle of a record (or of a BGZF block) is reported, not read
    as a shorter file and not looped on."""
    data = open(giab_paths["father"], "rb").read()
    p = str(tmp_path / "cut.bam")
    open(p, "wb").write(data[:len(data) // 2])
    with pytest.raises(Exception):
        with bamio.BamReader(p, threads=2) as rd:
            for b in rd.batches(bamio.MODE_FASTA):
                b.close()


def test_count_bam_into_table_routes_probes_through_the_chain_helper(giab_paths):
    """The pipeline's parent count (`samtools fasta | jellyfish count --if`,
    core/jellyfish_wrappers.py:159-176) decodes batch by batch and applies every batch
    with kmer_chain.count_if_present (direct / filtered / binned route); other modes
    go straight to the stream kernel.  Host logic only: the engine is a recorder."""
    from kmer_denovo_filter_b200.core import kmer_engine_wrappers as kw

    class Table:
        capacity, key_words, k, filter_buf = 1024, 1, 31, None

    class Recorder:
        def __init__(self):
            self.calls = []

        def upload(self, batch, with_reads=True):
            return batch

        def new_stats(self):
            return {}

        def count_stream(self, table, ds, mode, plane, arg, stats):
            self.calls.append((mode, plane, arg, ds.n_bases))
            stats["windows"] = ds.n_bases

        def read_stats(self, st):
            return {"windows": st["windows"], "full": 0, "hits": 0, "new": 0}

    for mode in (engine.MODE_COUNT_IF_PRESENT, engine.MODE_INSERT_COUNT):
        eng = Recorder()
        _t, tot = kw.count_bam_into_table(eng, giab_paths["mother"], Table(), mode, 1, 2,
                                          batch_bases=500_000)
        assert len(eng.calls) > 3 and all(c[:3] == (mode, 1, 1) for c in eng.calls)
        assert tot["windows"] == sum(c[3] for c in eng.calls) == tot["bases"]


# ---------------------------------------------------------------------------
# round 2: one decode for the counting stream and the scan, random access to
# records, the BGZF writer, corrupt input
# ---------------------------------------------------------------------------

def _kmer_multiset(hs, k=21):
    lo, hi, ok = engine.debug_extract_host(hs, k)
    return np.sort(lo[ok])


@pytest.mark.parametrize("who", ["child", "mother"])
def test_fasta_keep_mask_gives_the_counting_stream(giab_paths, who):
    """A scan-mode decode + the fasta_keep mask = the `samtools fasta -F 0xD00` stream:
    same records, and (through bamio.counting_view) the same canonical k-mers."""
    with bamio.BamReader(giab_paths[who], threads=3) as rd:
        f = rd.next_batch(bamio.MODE_FASTA)
    with bamio.BamReader(giab_paths[who], threads=3) as rd:
        s = rd.next_batch(bamio.MODE_SCAN, want_meta=True)
    with bamio.BamReader(giab_paths[who], threads=3) as rd:
        a = rd.next_batch(bamio.MODE_ALL)
    assert f.fasta_keep.all()
    assert s.rec_index[s.fasta_keep == 1].tolist() == f.rec_index.tolist()
    assert a.rec_index[a.fasta_keep == 1].tolist() == f.rec_index.tolist()
    assert 0 < int((s.fasta_keep == 0).sum()) < s.n_reads // 2     # the fixture has supplementary reads
    view = bamio.counting_view(s)
    assert np.array_equal(_kmer_multiset(view), _kmer_multiset(f))
    assert np.array_equal(view.invalid, engine.invalid_positions(view.valid, view.n_bases))
    assert np.array_equal(s.valid, engine.HostStream(s.codes, s.valid, s.n_bases, s.read_starts,
                                                     s.read_lens).valid)      # the batch itself is untouched


def test_fasta_keep_survives_batch_limits(giab_paths):
    """The QNAME-run state is carried across batches (and rolled back for a record that a
    batch limit postponed) in every mode."""
    with bamio.BamReader(giab_paths["child"], threads=2) as rd:
        whole = rd.next_batch(bamio.MODE_SCAN)
    keep = []
    with bamio.BamReader(giab_paths["child"], threads=2) as rd:
        for b in rd.batches(bamio.MODE_SCAN, max_bases=50_000):
            keep += b.fasta_keep.tolist()
    assert keep == whole.fasta_keep.tolist()


@pytest.mark.parametrize("chunk_kb,gap", [(None, None), (64, 16)])
def test_fetch_records_by_offset(giab_paths, monkeypatch, chunk_kb, gap):
    """rec_uoff + kdf_bam_fetch_records return exactly the raw records of a full decode,
    whatever the chunking of the sequential pass that produced the offsets."""
    if chunk_kb:
        monkeypatch.setenv("KDF_BAM_CHUNK_KB", str(chunk_kb))
        monkeypatch.setenv("KDF_BAM_GAP", str(gap))
    with bamio.BamReader(giab_paths["child"], threads=3) as rd:
        raws, uoffs = [], []
        for b in rd.batches(bamio.MODE_ALL, max_bases=400_000, want_meta=3):
            ro = b.raw_off.astype(np.int64)
            raws += [bytes(b.raw_blob[ro[i]:ro[i + 1]]) for i in range(b.n_reads)]
            uoffs += b.rec_uoff.tolist()
        assert len(set(uoffs)) == len(uoffs) and uoffs == sorted(uoffs)
        pick = list(range(0, len(uoffs), 37)) + [len(uoffs) - 1, 0, 5, 5]
        got = rd.fetch_records([uoffs[i] for i in pick])
        assert got == [raws[i] for i in pick]
        assert rd.fetch_records([]) == []
        with pytest.raises(engine.KdfError):
            rd.fetch_records([uoffs[3] + 1])        # not a record boundary


def test_bgzf_write_roundtrip(tmp_path):
    import gzip
    rng = np.random.default_rng(5)
    for n in (0, 1, 0xff00, 0xff00 + 1, 300_001):
        data = rng.integers(0, 7, size=n, dtype=np.uint8).tobytes()
        p = str(tmp_path / ("x%d.gz" % n))
        coff = bamio.bgzf_write(p, data, level=4, threads=3)
        assert gzip.open(p, "rb").read() == data
        raw = open(p, "rb").read()
        assert raw.endswith(obam.BGZF_EOF)
        assert coff.shape[0] == (n + 0xff00 - 1) // 0xff00 + 1
        for o in coff.tolist():
            assert raw[o:o + 4] == b"\x1f\x8b\x08\x04"
        assert int(coff[-1]) == len(raw) - len(obam.BGZF_EOF)


def _one_record_bam(tmp_path, mutate):
    rec = bytearray(obam.encode_record(0, 10, "read1", 0x41, 60, [(0, 12)], "ACGTACGTACGT"))
    mutate(rec)
    good = obam.encode_record(0, 20, "read2", 0x81, 60, [(0, 8)], "ACGTACGT")
    p = str(tmp_path / "m.bam")
    obam.write_bam(p, ["chr1"], [1000], [bytes(rec), good])
    return p


@pytest.mark.parametrize("field,value", [("l_seq", 0x7fffff00), ("l_seq", -5), ("n_cigar", 60000),
                                         ("l_name", 255), ("block_size", 0x7fffffff), ("block_size", 8)])
def test_corrupt_record_fields_are_reported(tmp_path, field, value):
    """A record whose l_seq / n_cigar_op / l_read_name do not fit its block_size (or a
    block_size out of range) is an error — not a segfault in the packer, not garbage."""
    import struct

    def mutate(rec):
        if field == "l_seq":
            struct.pack_into("<i", rec, 4 + 16, value)
        elif field == "n_cigar":
            struct.pack_into("<H", rec, 4 + 12, value)
        elif field == "l_name":
            rec[4 + 8] = value
        else:
            struct.pack_into("<i", rec, 0, value)
    p = _one_record_bam(tmp_path, mutate)
    for mode, meta in ((bamio.MODE_FASTA, 0), (bamio.MODE_ALL, 3)):
        with pytest.raises(engine.KdfError):
            with bamio.BamReader(p, threads=2) as rd:
                for b in rd.batches(mode, want_meta=meta):
                    b.close()


def test_corrupt_bgzf_block_is_reported(giab_paths, tmp_path, monkeypatch):
    """A flipped payload byte fails the block's CRC32 (or the inflate); ISIZE above 64 KiB
    is rejected before any allocation; KDF_BAM_CRC=0 skips only the CRC comparison."""
    import struct
    data = bytearray(open(giab_paths["father"], "rb").read())
    bsize = struct.unpack_from("<H", data, 16)[0] + 1
    second = bsize                                     # start of the second block
    bsize2 = struct.unpack_from("<H", data, second + 16)[0] + 1
    bad = bytearray(data)
    struct.pack_into("<I", bad, second + bsize2 - 8, 0xdeadbeef)        # stored CRC32
    p = str(tmp_path / "crc.bam")
    open(p, "wb").write(bad)
    with pytest.raises(engine.KdfError):
        with bamio.BamReader(p, threads=2) as rd:
            for b in rd.batches(bamio.MODE_FASTA):
                b.close()
    monkeypatch.setenv("KDF_BAM_CRC", "0")
    with bamio.BamReader(p, threads=2) as rd:
        n = sum(b.n_reads for b in rd.batches(bamio.MODE_FASTA))
    assert n > 0
    monkeypatch.delenv("KDF_BAM_CRC")
    big = bytearray(data)
    struct.pack_into("<I", big, second + bsize2 - 4, 1 << 24)           # ISIZE
    open(p, "wb").write(big)
    with pytest.raises(engine.KdfError):
        with bamio.BamReader(p, threads=2) as rd:
            for b in rd.batches(bamio.MODE_FASTA):
                b.close()


def _records(b):
    return [(b.record(i).query_name, int(b.flag[i]), int(b.ref_id[i]), int(b.pos[i])) for i in range(b.n_reads)]


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("mode", [bamio.MODE_FASTA, bamio.MODE_SCAN])
def test_shards_partition_the_file_exactly(giab_paths, world, mode):
    """Rank ranges cut at .bai linear-index entries: the ranks' streams, concatenated, are
    the sequential stream — record for record, including the QNAME-run collapse of the
    FASTA stream across a boundary (each rank warms the state up from an earlier entry)."""
    with bamio.BamReader(giab_paths["child"], threads=2) as rd:
        whole = _records(rd.next_batch(mode, want_meta=True))
    got = []
    sizes = []
    for r in range(world):
        rd = bamio.open_shard(giab_paths["child"], r, world, threads=2)
        n0 = len(got)
        for b in rd.batches(mode, max_bases=300_000, want_meta=True):
            got += _records(b)
            b.close()
        rd.close()
        sizes.append(len(got) - n0)
    assert got == whole
    assert sum(1 for x in sizes if x > 0) >= 2        # the file really was split


def test_region_fetch_equals_a_linear_filter(giab_paths, giab_records):
    """BamReader.fetch through the .bai returns every record overlapping the region (the
    reference's bam.fetch(chrom, pos, pos + 1))."""
    recs = giab_records["child"]
    names = giab_records["child_refs"][0]
    rng = np.random.default_rng(3)
    mapped = [r for r in recs if r.ref_id >= 0 and r.pos >= 0 and r.reference_end is not None]
    with bamio.BamReader(giab_paths["child"], threads=2) as rd:
        for r in [mapped[int(i)] for i in rng.integers(0, len(mapped), 12)]:
            tid, pos = r.ref_id, r.pos + 3
            want = sorted((x.qname, x.flag, x.pos) for x in recs
                          if x.ref_id == tid and x.reference_end is not None and x.pos <= pos < x.reference_end)
            got = []
            for b in rd.fetch(tid, pos, pos + 1, want_meta=True):
                start = b.pos.astype(np.int64)
                for i in range(b.n_reads):
                    rec = b.record(i)
                    if int(b.ref_id[i]) == tid and rec.reference_end is not None and start[i] <= pos < rec.reference_end:
                        got.append((rec.query_name, int(b.flag[i]), int(start[i])))
                b.close()
            assert sorted(got) == want and want, names[tid]


def test_concurrent_readers_equal_sequential(giab_paths):
    """Three readers decoding at once (what the discovery pipeline does with child, mother and
    father: they share the cores and the batch-buffer pool) deliver what each delivers alone."""
    import hashlib
    import threading

    def digest(path, mode, meta, out, key):
        h = hashlib.sha256()
        n = 0
        with bamio.BamReader(path, threads=4) as rd:
            for b in rd.batches(mode, max_bases=400_000, want_meta=meta):
                for name in ("codes", "valid", "invalid", "read_starts", "read_lens", "rec_index", "rec_uoff",
                             "fasta_keep") + (("pos", "flag", "qname_blob", "cigar_blob") if meta else ()):
                    h.update(np.ascontiguousarray(getattr(b, name)).tobytes())
                n += b.n_reads
                b.close()
        out[key] = (n, h.hexdigest())

    jobs = [("child", bamio.MODE_SCAN, True), ("mother", bamio.MODE_FASTA, False), ("father", bamio.MODE_FASTA, False)]
    alone = {}
    for who, mode, meta in jobs:
        digest(giab_paths[who], mode, meta, alone, who)
    for _round in range(3):
        together = {}
        ths = [threading.Thread(target=digest, args=(giab_paths[who], mode, meta, together, who))
               for who, mode, meta in jobs]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        assert together == alone


@pytest.mark.parametrize("chunk_kb,gap,threads,block_payload", [
    (None, None, 4, 60000), (64, 0, 3, 60000), (100, 17, 8, 777), (300, 5000, 1, 65280), (64, 1, 2, 10_000)])
def test_long_and_tiny_records_across_blocks_and_chunks(tmp_path, monkeypatch, chunk_kb, gap, threads, block_payload):
    """Records from 40 bytes to 400 KB (long reads: one record spans many BGZF blocks and more
    than a pipeline chunk), block sizes that cut block_size fields in two: the chained record
    walk must hand every record on intact."""
    import random
    rng = random.Random(7)
    recs, want = [], []
    for i in range(260):
        n = rng.choice([0, 1, 2, 31, 150, 151, 1000, 20_000, 70_000, 200_000 if i % 40 == 0 else 300])
        seq = "".join(rng.choice("ACGTN" if i % 7 == 0 else "ACGT") for _ in range(n))
        flag = rng.choice([0, 0x10, 0x41, 0x81, 0x100, 0x400, 0x800, 0x4])
        name = "read%d" % (i // 2 if i % 5 else i)      # some runs of equal QNAMEs
        cigar = [(0, n)] if n and not (flag & 0x4) else []
        recs.append(obam.encode_record(0, 100 + i, name, flag, 60, cigar, seq,
                                       qual=bytes(rng.randrange(0, 42) for _ in range(n))))
        want.append((name, flag, seq))
    p = str(tmp_path / "long.bam")
    obam.write_bam(p, ["chr1"], [10_000_000], recs, block_payload=block_payload)
    if chunk_kb is not None:
        monkeypatch.setenv("KDF_BAM_CHUNK_KB", str(chunk_kb))
        monkeypatch.setenv("KDF_BAM_GAP", str(gap))
    for max_bases in (0, 50_000):
        got = []
        with bamio.BamReader(p, threads=threads) as rd:
            for b in rd.batches(bamio.MODE_ALL, max_bases=max_bases, want_meta=3):
                for i in range(b.n_reads):
                    r = b.record(i)
                    got.append((r.query_name, r.flag, r.query_sequence or ""))
                    assert bytes(b.raw_blob[int(b.raw_off[i]):int(b.raw_off[i + 1])]) == recs[len(got) - 1][4:]
                b.close()
        assert got == want
    # the `samtools fasta` view of the same file against the oracle's collapse rule
    orecs = obam.read_bam(p)[2] if hasattr(obam, "read_bam") else None
    if orecs is not None:
        keep = obam.fasta_stream(orecs)
        with bamio.BamReader(p, threads=threads) as rd:
            names = []
            for b in rd.batches(bamio.MODE_FASTA, max_bases=80_000, want_meta=True):
                names += [b.record(i).query_name for i in range(b.n_reads)]
                b.close()
        assert names == [r.qname for r in keep]


def test_allocation_failure_is_an_error_not_a_crash(giab_paths):
    """A batch buffer that cannot grow (KDF_BAM_FAIL_ALLOC: the n-th growth throws) must come
    back as KdfError — also when it happens inside the decoder's parallel region, which no C++
    exception may leave.  Each case runs in its own process: the hook counts process-wide."""
    import subprocess
    import sys
    code = (
        "import sys\n"
        "from kmer_denovo_filter_b200 import bamio, engine\n"
        "try:\n"
        "    n = 0\n"
        "    with bamio.BamReader(sys.argv[1], threads=3) as rd:\n"
        "        for b in rd.batches(bamio.MODE_SCAN, max_bases=300_000, want_meta=2):\n"
        "            n += b.n_reads; b.close()\n"
        "    print('ok', n)\n"
        "except engine.KdfError as e:\n"
        "    print('kdferror', e)\n")
    import os
    outcomes = set()
    for fail_at in (1, 2, 5, 9, 14, 23, 10_000):
        env = dict(os.environ, KDF_BAM_FAIL_ALLOC=str(fail_at), KDF_BAM_POOL_MB="0",
                   PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        r = subprocess.run([sys.executable, "-c", code, giab_paths["child"]], capture_output=True, text=True,
                           timeout=120, env=env)
        assert r.returncode == 0, (fail_at, r.returncode, r.stderr[-400:])
        word = r.stdout.split()[0]
        assert word in ("ok", "kdferror"), r.stdout
        if word == "kdferror":
            assert "memory" in r.stdout
        outcomes.add((word, fail_at == 10_000))
    assert ("kdferror", False) in outcomes and ("ok", True) in outcomes


@pytest.mark.parametrize("block_payload", [1, 5, 37, 65280])
def test_stored_blocks_tiny_blocks_and_no_eof_marker(tmp_path, block_payload):
    """Uncompressed (stored) BGZF blocks down to ONE byte each — every block_size field and
    every header is then split over several blocks — in a file without the EOF marker block;
    and a BAM with no record at all."""
    import struct
    import zlib
    recs = [obam.encode_record(0, i, "r%d" % i, 0x10 if i % 3 else 0, 60, [(0, 4)], "ACGT") for i in range(50)]
    p0 = str(tmp_path / "z.bam")
    obam.write_bam(p0, ["chr1"], [1000], recs)
    raw = obam._inflate_all(p0)

    def stored(payload):
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        c = co.compress(payload) + co.flush()
        return (struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(c) + 25) + c +
                struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload)))
    p = str(tmp_path / "stored.bam")
    with open(p, "wb") as fh:
        for i in range(0, len(raw), block_payload):
            fh.write(stored(raw[i:i + block_payload]))
    for threads in (1, 3):
        n = 0
        with bamio.BamReader(p, threads=threads) as rd:
            for b in rd.batches(bamio.MODE_ALL, max_bases=60, want_meta=3):
                for i in range(b.n_reads):
                    assert bytes(b.raw_blob[int(b.raw_off[i]):int(b.raw_off[i + 1])]) == recs[n][4:]
                    n += 1
                b.close()
        assert n == len(recs)
    pe = str(tmp_path / "empty.bam")
    obam.write_bam(pe, ["chr1"], [1000], [])
    with bamio.BamReader(pe, threads=2) as rd:
        b = rd.next_batch(bamio.MODE_ALL, want_meta=3)
        assert b.n_reads == 0 and b.at_eof


_DIGEST_CODE = r'''
import hashlib, os, sys
import numpy as np
from kmer_denovo_filter_b200 import bamio
path = sys.argv[1]
for mode, mb, meta, ck, gap, thr in ((0, 0, 0, None, None, 4), (1, 100000, 1, 64, 0, 3), (2, 333333, 3, 100, 17, 8),
                                     (1, 0, 2, 300, 5000, 1), (0, 1000000, 1, 65, 1, 2)):
    if ck is None:
        os.environ.pop("KDF_BAM_CHUNK_KB", None); os.environ.pop("KDF_BAM_GAP", None)
    else:
        os.environ["KDF_BAM_CHUNK_KB"] = str(ck); os.environ["KDF_BAM_GAP"] = str(gap)
    h = hashlib.sha256(); n = 0
    with bamio.BamReader(path, threads=thr) as rd:
        for b in rd.batches(mode, max_bases=mb, want_meta=meta):
            names = ["codes", "valid", "invalid", "read_starts", "read_lens", "rec_index", "rec_uoff", "fasta_keep"]
            if meta: names += ["ref_id", "pos", "flag", "mapq", "qname_off", "qname_blob", "cigar_off", "cigar_blob", "sa_blob"]
            if meta >= 2: names += ["qual_blob"]
            if meta >= 3: names += ["raw_off", "raw_blob"]
            for name in names:
                h.update(np.ascontiguousarray(getattr(b, name)).tobytes())
            h.update(b"|%d|%d|" % (b.n_reads, b.n_bases)); n += b.n_reads
            b.close()
    print(mode, mb, meta, n, h.hexdigest())
'''


@pytest.mark.parametrize("who", ["child", "mother"])
def test_decoder_switches_do_not_change_the_batches(giab_paths, who):
    """Headers classified by the walker while a block is hot / all of them after the barrier;
    own inflate / zlib; vector / scalar base packing: the same batches, bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for label, extra in (("default", {}), ("cold_classify", {"KDF_BAM_WALK_CLASSIFY": "0"}),
                         ("zlib", {"KDF_BAM_ZLIB": "1", "KDF_CRC_ZLIB": "1"}), ("scalar_pack", {"KDF_PACK_SCALAR": "1"})):
        env = dict(os.environ, PYTHONPATH=root, **extra)
        r = subprocess.run([sys.executable, "-c", _DIGEST_CODE, giab_paths[who]], capture_output=True, text=True,
                           timeout=300, env=env)
        assert r.returncode == 0, r.stderr[-500:]
        outs[label] = r.stdout
    assert len(outs["default"].splitlines()) == 5
    for label in ("cold_classify", "zlib", "scalar_pack"):
        assert outs[label] == outs["default"], label
