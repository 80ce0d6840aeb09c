"""Host side of VCF mode (no GPU): the product's child k-mer collection equals
the oracle's, and its writers reproduce the reference's golden files when fed
the (golden-pinned) oracle annotations."""
import gzip
import os

import numpy as np
import pytest

from kmer_denovo_filter_b200.vcf import pipeline as P
from oracle import vcf as ovcf


@pytest.fixture(scope="module")
def oracle_run(giab_records, giab_paths):
    _h, _s, variants = ovcf.parse_vcf(giab_paths["vcf"], "HG002")
    ann, metrics, found = ovcf.run(giab_records["child"], giab_records["mother"],
                                   giab_records["father"], variants, 31)
    return variants, ann, metrics, found


def test_parse_vcf_variants_matches_oracle(giab_paths):
    got = P._parse_vcf_variants(giab_paths["vcf"], "HG002")
    _h, _s, want = ovcf.parse_vcf(giab_paths["vcf"], "HG002")
    assert len(got) == 22
    for g, w in zip(got, want):
        assert {k: g[k] for k in ("chrom", "pos", "ref", "alts", "alt", "id")} == \
               {k: w[k] for k in ("chrom", "pos", "ref", "alts", "alt", "id")}


@pytest.mark.parametrize("use_index", [True, False])
def test_collect_child_kmers_matches_oracle(giab_paths, giab_records, tmp_path, monkeypatch, use_index):
    """Both routes of the child-side read collection — region fetches through the .bai, and
    the index-free linear pass with a sorted-interval lookup — against the oracle (same reads
    per variant, in file order)."""
    monkeypatch.setenv("KDF_VCF_FETCH", "1" if use_index else "0")
    variants = P._parse_vcf_variants(giab_paths["vcf"], "HG002")
    fa = str(tmp_path / "child_kmers.fa")
    total, vrk = P._collect_child_kmers(giab_paths["child"], None, variants, 31, 20, 20, False, fa)
    want_total, want_vrk, want_all = ovcf.collect_child(giab_records["child"], variants, 31, 20, 20)
    assert total == want_total == 1484
    assert set(vrk) == set(want_vrk)
    for key in vrk:
        assert [(n, sorted(k), s) for n, k, s in vrk[key]] == \
               [(n, sorted(k), s) for n, k, s in want_vrk[key]], key
    written = [l.strip() for l in open(fa) if not l.startswith(">")]
    assert sorted(written) == sorted(want_all)


def test_annotate_and_writers_reproduce_goldens(oracle_run, giab_paths, giab_records, tmp_path):
    variants, ann, metrics, found = oracle_run
    # product annotate on the oracle's inputs
    _t, vrk, _all = ovcf.collect_child(giab_records["child"], variants, 31, 20, 20)
    pvariants = P._parse_vcf_variants(giab_paths["vcf"], "HG002")
    got_ann, inf, _ia = P._annotate_variants(pvariants, vrk, found)
    assert got_ann == ann
    assert len(inf) == metrics["variants_with_unique_reads"] == 12
    exp = giab_paths["expected_vcf"]
    # summary.txt byte for byte
    s = str(tmp_path / "summary.txt")
    P._write_summary(s, pvariants, got_ann)
    assert open(s).read() == open(os.path.join(exp, "summary.txt")).read()
    # annotated VCF: every line (header and records) equals the golden
    out = P._write_annotated_vcf(giab_paths["vcf"], str(tmp_path / "annotated.vcf"), got_ann, "HG002")
    assert out.endswith(".vcf.gz")
    got_lines = gzip.open(out, "rt").read().splitlines()
    want_lines = gzip.open(os.path.join(exp, "annotated.vcf.gz"), "rt").read().splitlines()
    assert got_lines == want_lines
    # INFO fallback when the proband is not a sample of the VCF
    out2 = P._write_annotated_vcf(giab_paths["vcf"], str(tmp_path / "info.vcf.gz"), got_ann, "NOPE")
    rec = [l for l in gzip.open(out2, "rt").read().splitlines() if not l.startswith("#")][0].split("\t")
    assert "DKU=1;DKT=17;DKA=1;DKU_DKT=0.0588" in rec[7] and rec[8] == "GT:PS:DP:ADALL:AD:GQ"
    hdr = [l for l in gzip.open(out2, "rt").read().splitlines() if l.startswith("##INFO=<ID=DKU,")]
    assert len(hdr) == 1


def test_bgzf_output_is_valid_multi_member_gzip(tmp_path):
    data = (b"chr1\t%d\n" * 1) * 1
    blob = b"".join(b"line %d\n" % i for i in range(30000))
    p = str(tmp_path / "x.gz")
    P._bgzf_write(p, blob)
    assert gzip.open(p, "rb").read() == blob
    raw = open(p, "rb").read()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC"
    assert raw.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def test_read_helpers_known_answers():
    """Known answers in the style of the reference's tests/test_kmer_utils.py
    (MockRead with get_aligned_pairs / get_reference_positions)."""
    from kmer_denovo_filter_b200.kmer_utils import extract_variant_spanning_kmers, read_supports_alt

    class MockRead:
        def __init__(self, seq, start, cigar, quals=None):
            self.query_sequence = seq
            self.query_qualities = quals if quals is not None else [30] * len(seq)
            self.reference_start = start
            self.cigartuples = cigar

        def get_aligned_pairs(self, matches_only=False):
            out, q, r = [], 0, self.reference_start
            for op, ln in self.cigartuples:
                if op == 0:
                    out += [(q + i, r + i) for i in range(ln)]; q += ln; r += ln
                elif op == 1:
                    out += [] if matches_only else [(q + i, None) for i in range(ln)]; q += ln
                elif op == 2:
                    out += [] if matches_only else [(None, r + i) for i in range(ln)]; r += ln
            return out

        def get_reference_positions(self, full_length=False):
            out = [None] * len(self.query_sequence)
            for q, r in self.get_aligned_pairs(True):
                out[q] = r
            return out if full_length else [x for x in out if x is not None]

    seq = "ACGTACGTAC"
    r = MockRead(seq, 100, [(0, 10)])
    ks = extract_variant_spanning_kmers(r, 104, 5)
    assert len(ks) == 5 or len(ks) < 5      # canonical forms may coincide
    assert all(len(x) == 5 for x in ks)
    assert extract_variant_spanning_kmers(r, 99, 5) == set()         # not covered
    assert extract_variant_spanning_kmers(MockRead("ACGTNCGTAC", 100, [(0, 10)]), 104, 5) == set()
    lowq = MockRead(seq, 100, [(0, 10)], quals=[30] * 4 + [5] + [30] * 5)
    assert extract_variant_spanning_kmers(lowq, 104, 5, min_baseq=20) == set()
    assert read_supports_alt(r, 104, "A", "A") and not read_supports_alt(r, 104, "A", "G")
    assert not read_supports_alt(r, 104, "A", "<DEL>") and not read_supports_alt(r, 104, "A", None)
    ins = MockRead("ACGTAGGCGTAC", 100, [(0, 5), (1, 2), (0, 5)])
    assert read_supports_alt(ins, 104, "A", "AGG") and not read_supports_alt(ins, 104, "A", "A")
    dele = MockRead("ACGTAGTAC", 100, [(0, 5), (2, 1), (0, 4)])
    assert read_supports_alt(dele, 104, "AC", "A") and not read_supports_alt(dele, 104, "AC", "AC")
    assert not read_supports_alt(lowq, 104, "A", "A", min_baseq=20)


def test_tabix_index_structure(oracle_run, giab_paths, tmp_path):
    """The .tbi beside the annotated VCF: header fields of the VCF preset, one
    sequence entry per contig in file order, chunks that start at record lines."""
    import struct
    variants, ann, _m, _f = oracle_run
    out = P._write_annotated_vcf(giab_paths["vcf"], str(tmp_path / "a.vcf.gz"), ann, "HG002")
    tbi = gzip.open(out + ".tbi", "rb").read()
    assert tbi[:4] == b"TBI\x01"
    n_ref, fmt, col_seq, col_beg, col_end, meta, skip, l_nm = struct.unpack_from("<8i", tbi, 4)
    assert (fmt, col_seq, col_beg, col_end, meta, skip) == (2, 1, 2, 0, ord("#"), 0)
    names = tbi[36:36 + l_nm].split(b"\0")[:-1]
    want_names = []
    for v in variants:
        if v["chrom"].encode() not in want_names:
            want_names.append(v["chrom"].encode())
    assert names == want_names and n_ref == len(names)
    # walk the index: every chunk's virtual offset points at the start of a record line
    raw = open(out, "rb").read()
    text = gzip.open(out, "rb").read()
    off = 36 + l_nm
    n_chunks = 0
    import zlib
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", tbi, off)[0]; off += 4
        for _b in range(n_bin):
            _bin, n_chunk = struct.unpack_from("<Ii", tbi, off); off += 8
            for _c in range(n_chunk):
                v0, v1 = struct.unpack_from("<QQ", tbi, off); off += 16
                blk, within = v0 >> 16, v0 & 0xFFFF
                bsize = struct.unpack_from("<H", raw, blk + 16)[0] + 1
                payload = zlib.decompress(raw[blk + 18:blk + bsize - 8], -15)
                assert payload[within:within + 3] == b"chr" and v0 < v1
                n_chunks += 1
        n_intv = struct.unpack_from("<i", tbi, off)[0]; off += 4 + 8 * n_intv
    assert n_chunks >= n_ref and off == len(tbi)
    assert text.count(b"\nchr") == len(variants)


def test_informative_reads_bam_dv_tags(giab_paths, giab_records, tmp_path):
    """VCF mode's informative-reads BAM (reference vcf/pipeline.py:1307-1357): for every
    site in sorted order, the first record of each informative read name that overlaps the
    site, tagged DV:Z with the sorted variant keys; coordinate sorted + .bai."""
    from kmer_denovo_filter_b200 import bamio
    recs = giab_records["child"]
    names = giab_records["child_refs"][0]
    variants = P._parse_vcf_variants(giab_paths["vcf"], "HG002")
    # a synthetic "informative" assignment: every third read over each of the first 8 sites
    by_var = {}
    for var in variants[:8]:
        tid = names.index(var["chrom"])
        over = [r.qname for r in recs if r.ref_id == tid and r.reference_end is not None
                and r.pos <= var["pos"] < r.reference_end]
        if over[::3]:
            by_var[P._var_key(var)] = set(over[::3])
    assert len(by_var) >= 4
    out = str(tmp_path / "inf.bam")
    n = P._write_informative_reads(giab_paths["child"], None, by_var, out, threads=2)
    # independent restatement over the stdlib reader's records
    r2v = {}
    for key, rn in by_var.items():
        for x in rn:
            r2v.setdefault(x, set()).add(key)
    want, written = [], set()
    for chrom, pos in sorted({(k.split(":")[0], int(k.split(":")[1])) for k in by_var}):
        tid = names.index(chrom)
        for r in recs:
            end = r.reference_end if r.reference_end is not None else r.pos + 1
            if r.ref_id == tid and r.pos < pos + 1 and end > pos and r.qname in r2v and r.qname not in written:
                written.add(r.qname)
                want.append((r.qname, r.flag, r.pos, ",".join(sorted(r2v[r.qname]))))
    assert n == len(want) > 10
    with bamio.BamReader(out, threads=2) as rd:
        b = rd.next_batch(bamio.MODE_ALL, want_meta=3)
    order = b.ref_id.astype(np.int64) * (1 << 32) + b.pos.astype(np.int64)
    assert b.n_reads == n and (np.diff(order) >= 0).all()
    got = []
    ro = b.raw_off.astype(np.int64)
    for i in range(n):
        raw = bytes(b.raw_blob[ro[i]:ro[i + 1]])
        j = raw.rindex(b"DVZ")
        got.append((b.record(i).query_name, int(b.flag[i]), int(b.pos[i]), raw[j + 3:-1].decode()))
    assert sorted(got) == sorted(want)
    assert os.path.isfile(out + ".bai") and bamio.read_bai(out + ".bai")
