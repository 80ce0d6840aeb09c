"""The reference's own pipeline scenarios (tests/discovery/test_pipeline.py,
tests/vcf/test_pipeline.py: tiny synthetic trios, k = 5) through this package's
CLI parsers and pipelines on the GPU, each also checked against the oracle."""
import gzip
import json
import os

import pytest

from helpers import create_ref_fasta, create_vcf, simple_bam

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from kmer_denovo_filter_b200 import engine
    return engine.CudaEngine()


def _disc(eng, tmp, child, mother, father, ref_fa, extra=()):
    from kmer_denovo_filter_b200 import cli
    from kmer_denovo_filter_b200.discovery import pipeline as P
    prefix = os.path.join(tmp, "disc_out")
    args = cli.parse_discovery_args(["--child", child, "--mother", mother, "--father", father,
                                     "--ref-fasta", ref_fa, "--out-prefix", prefix,
                                     "--min-child-count", "3", "--kmer-size", "5", *extra])
    P.run_discovery_pipeline(args, engine=eng)
    metrics = json.load(open(prefix + ".metrics.json"))
    bed = [l.rstrip("\n").split("\t") for l in open(prefix + ".bed") if l.strip() and not l.startswith("#")]
    return prefix, metrics, bed


def _oracle_disc(child, mother, father, ref_fa, **kw):
    from oracle import bam as obam, discovery as odisc
    recs = [obam.read_bam(p)[2] for p in (child, mother, father)]
    return odisc.run(recs[0], recs[1], recs[2], [s for _n, s in obam.read_fasta(ref_fa)], 5, **kw)


def _trio(tmp, ref_seq, child_reads, parent_seq_reads=None):
    chrom = "chr1"
    paths = {w: os.path.join(tmp, w + ".bam") for w in ("child", "mother", "father")}
    simple_bam(paths["child"], chrom, child_reads)
    pr = parent_seq_reads or [(30, ref_seq[30:90])] * 3
    simple_bam(paths["mother"], chrom, [("m%d" % i, p, s) for i, (p, s) in enumerate(pr)])
    simple_bam(paths["father"], chrom, [("f%d" % i, p, s) for i, (p, s) in enumerate(pr)])
    return paths


def _mutate(seq, i, avoid):
    s = list(seq)
    s[i] = "G" if avoid != "G" else "T"
    return "".join(s)


def test_discovery_denovo_detected(eng, tmp_path):
    """reference tests/discovery/test_pipeline.py:39 — plus exact stage sizes."""
    tmp = str(tmp_path)
    ref_fa = os.path.join(tmp, "ref.fa")
    ref = create_ref_fasta(ref_fa)
    child_seq = _mutate(ref[30:90], 20, ref[50])
    p = _trio(tmp, ref, [("read%d" % i, 30, child_seq) for i in range(1, 5)])
    prefix, metrics, bed = _disc(eng, tmp, p["child"], p["mother"], p["father"], ref_fa)
    assert len(bed) >= 1 and len(bed[0]) == 10
    assert int(bed[0][3]) >= 1 and int(bed[0][4]) >= 1
    want = _oracle_disc(p["child"], p["mother"], p["father"], ref_fa)
    assert metrics["child_candidate_kmers"] == len(want["candidates"])
    assert metrics["non_ref_kmers"] == len(want["non_ref"])
    assert metrics["proband_unique_kmers"] == len(want["proband_unique"]) > 0
    assert metrics["informative_reads"] == want["informative"] == 4
    assert metrics["candidate_regions"] == len(want["regions"]) == len(bed)
    from oracle import bam as obam
    _n, _l, info = obam.read_bam(prefix + ".informative.bam")
    assert sorted(r.qname for r in info) == ["read1", "read2", "read3", "read4"]
    assert all(r.get_tag("dk") == 1 for r in info) and os.path.isfile(prefix + ".informative.bam.bai")
    for ext in (".summary.txt", ".kmer_coverage.bedgraph", ".read_coverage.bed", ".sv.bedpe"):
        assert os.path.isfile(prefix + ext)


def test_discovery_inherited_gives_no_unique_kmers(eng, tmp_path):
    """:228 — the variant is also in the mother: zero proband-unique k-mers, empty BED."""
    tmp = str(tmp_path)
    ref_fa = os.path.join(tmp, "ref.fa")
    ref = create_ref_fasta(ref_fa)
    var = _mutate(ref[30:90], 20, ref[50])
    p = _trio(tmp, ref, [("read%d" % i, 30, var) for i in range(4)], [(30, var)] * 3)
    _prefix, metrics, bed = _disc(eng, tmp, p["child"], p["mother"], p["father"], ref_fa)
    assert metrics["proband_unique_kmers"] == 0 and metrics["candidate_regions"] == 0 and bed == []
    assert metrics["non_ref_kmers"] > 0


def test_discovery_empty_child(eng, tmp_path):
    """:373 — a child BAM without reads: outputs exist, all counts zero."""
    tmp = str(tmp_path)
    ref_fa = os.path.join(tmp, "ref.fa")
    ref = create_ref_fasta(ref_fa)
    p = _trio(tmp, ref, [])
    prefix, metrics, bed = _disc(eng, tmp, p["child"], p["mother"], p["father"], ref_fa)
    assert metrics["child_candidate_kmers"] == 0 and metrics["informative_reads"] == 0 and bed == []
    assert os.path.isfile(prefix + ".summary.txt")


def test_parent_max_count_zero_vs_one(eng, tmp_path):
    """:717-782 — one parental read carrying the variant: removed at 0, kept at 1."""
    tmp = str(tmp_path)
    ref_fa = os.path.join(tmp, "ref.fa")
    ref = create_ref_fasta(ref_fa)
    var = _mutate(ref[30:90], 20, ref[50])
    chrom = "chr1"
    p = {w: os.path.join(tmp, w + ".bam") for w in ("child", "mother", "father")}
    simple_bam(p["child"], chrom, [("read%d" % i, 30, var) for i in range(4)])
    simple_bam(p["mother"], chrom, [("m0", 30, var), ("m1", 30, ref[30:90]), ("m2", 30, ref[30:90])])
    simple_bam(p["father"], chrom, [("f%d" % i, 30, ref[30:90]) for i in range(3)])
    _x, m0, bed0 = _disc(eng, tmp, p["child"], p["mother"], p["father"], ref_fa)
    _x, m1, bed1 = _disc(eng, tmp, p["child"], p["mother"], p["father"], ref_fa,
                         ("--parent-max-count", "1"))
    assert m0["proband_unique_kmers"] == 0 and bed0 == []
    w1 = _oracle_disc(p["child"], p["mother"], p["father"], ref_fa, parent_max_count=1)
    assert m1["proband_unique_kmers"] == len(w1["proband_unique"]) > 0 and len(bed1) == len(w1["regions"]) >= 1


def test_two_regions_with_cluster_distance_zero(eng, tmp_path):
    """:784-856 — two separated de novo sites stay two regions at --cluster-distance 0."""
    tmp = str(tmp_path)
    ref_fa = os.path.join(tmp, "ref.fa")
    ref = create_ref_fasta(ref_fa)
    a = _mutate(ref[10:60], 25, ref[35])
    b = _mutate(ref[120:170], 25, ref[145])
    child = [("a%d" % i, 10, a) for i in range(4)] + [("b%d" % i, 120, b) for i in range(4)]
    parents = [(10, ref[10:60])] * 3 + [(120, ref[120:170])] * 3
    p = _trio(tmp, ref, child, parents)
    _x, m, bed = _disc(eng, tmp, p["child"], p["mother"], p["father"], ref_fa, ("--cluster-distance", "0"))
    want = _oracle_disc(p["child"], p["mother"], p["father"], ref_fa, merge_distance=0)
    assert len(bed) == len(want["regions"]) == 2
    assert [int(r[3]) for r in bed] == [4, 4]
    _x, m2, bed2 = _disc(eng, tmp, p["child"], p["mother"], p["father"], ref_fa,
                         ("--cluster-distance", "500"))
    assert len(bed2) == 1 and int(bed2[0][3]) == 8


# ---- VCF mode -------------------------------------------------------------

def _vcf(eng, tmp, child, mother, father, vcf, extra=()):
    from kmer_denovo_filter_b200 import cli
    from kmer_denovo_filter_b200.vcf import pipeline as P
    out = os.path.join(tmp, "out.vcf.gz")
    args = cli.parse_vcf_args(["--child", child, "--mother", mother, "--father", father, "--vcf", vcf,
                               "--output", out, "--metrics", os.path.join(tmp, "m.json"),
                               "--summary", os.path.join(tmp, "s.txt"), "--kmer-size", "5",
                               "--proband-id", "HG002", *extra])
    res = P.run_pipeline(args, engine=eng)
    recs = [l.rstrip("\n").split("\t") for l in gzip.open(res["paths"]["vcf"], "rt") if not l.startswith("#")]
    vals = [dict(zip(r[8].split(":"), r[9].split(":"))) for r in recs]
    return res, vals


def _oracle_vcf(child, mother, father, vcf):
    from oracle import bam as obam, vcf as ovcf
    _h, _s, variants = ovcf.parse_vcf(vcf, "HG002")
    recs = [obam.read_bam(p)[2] for p in (child, mother, father)]
    ann, metrics, _f = ovcf.run(recs[0], recs[1], recs[2], variants, 5)
    return [ann[ovcf.var_key(v)] for v in variants], metrics


def test_vcf_denovo_and_inherited(eng, tmp_path):
    """tests/vcf/test_pipeline.py:41, :427 — DKU > 0 for a de novo SNV, 0 when a parent has it."""
    tmp = str(tmp_path)
    ref = create_ref_fasta(os.path.join(tmp, "ref.fa"))
    alt = "G" if ref[50] != "G" else "T"
    var = _mutate(ref[30:90], 20, ref[50])
    vcf = os.path.join(tmp, "in.vcf")
    create_vcf(vcf, "chr1", [(51, ref[50], alt)])
    p = _trio(tmp, ref, [("read%d" % i, 30, var) for i in range(4)])
    res, vals = _vcf(eng, tmp, p["child"], p["mother"], p["father"], vcf)
    want, wm = _oracle_vcf(p["child"], p["mother"], p["father"], vcf)
    assert res["metrics"] == wm
    assert int(vals[0]["DKU"]) == want[0]["dku"] == 4 and int(vals[0]["DKT"]) == 4 and int(vals[0]["DKA"]) == 4
    p2 = _trio(tmp, ref, [("read%d" % i, 30, var) for i in range(4)], [(30, var)] * 3)
    res2, vals2 = _vcf(eng, tmp, p2["child"], p2["mother"], p2["father"], vcf)
    assert int(vals2[0]["DKU"]) == 0 and int(vals2[0]["DKT"]) == 4
    assert res2["metrics"]["child_unique_kmers"] == 0


def test_vcf_indel_allele_specificity(eng, tmp_path):
    """:525-1043 — DKA counts only reads that carry exactly the candidate allele."""
    tmp = str(tmp_path)
    ref = create_ref_fasta(os.path.join(tmp, "ref.fa"))
    # child: 3 reads with a 2-bp deletion after position 50, 2 reads with a SNV at 50
    del_seq = ref[30:51] + ref[53:90]
    del_cigar = [(0, 21), (2, 2), (0, 37)]
    snv = _mutate(ref[30:90], 20, ref[50])
    child = [("d%d" % i, 30, del_seq, None, del_cigar) for i in range(3)] + \
            [("s%d" % i, 30, snv) for i in range(2)]
    p = _trio(tmp, ref, child)
    vcf = os.path.join(tmp, "in.vcf")
    create_vcf(vcf, "chr1", [(51, ref[50:53], ref[50])])
    res, vals = _vcf(eng, tmp, p["child"], p["mother"], p["father"], vcf, ("--min-baseq", "0"))
    want, wm = _oracle_vcf(p["child"], p["mother"], p["father"], vcf)
    assert res["metrics"] == wm
    got = {k: vals[0][k] for k in ("DKU", "DKT", "DKA")}
    assert got == {"DKU": str(want[0]["dku"]), "DKT": str(want[0]["dkt"]), "DKA": str(want[0]["dka"])}
    assert want[0]["dka"] == 3 and want[0]["dkt"] == 5
    # insertion of 3 bases after position 60
    ins_seq = ref[30:61] + "TTT" + ref[61:90]
    ins_cigar = [(0, 31), (1, 3), (0, 29)]
    p2 = _trio(tmp, ref, [("i%d" % i, 30, ins_seq, None, ins_cigar) for i in range(4)])
    vcf2 = os.path.join(tmp, "in2.vcf")
    create_vcf(vcf2, "chr1", [(61, ref[60], ref[60] + "TTT"), (61, ref[60], ref[60] + "TTA")])
    res2, vals2 = _vcf(eng, tmp, p2["child"], p2["mother"], p2["father"], vcf2)
    want2, _wm2 = _oracle_vcf(p2["child"], p2["mother"], p2["father"], vcf2)
    assert [int(v["DKA"]) for v in vals2] == [w["dka"] for w in want2] == [4, 0]
    assert [int(v["DKU"]) for v in vals2] == [w["dku"] for w in want2]


def test_vcf_multiallelic_uses_proband_genotype(eng, tmp_path):
    """:1317-1572 — GT 0/2 evaluates the second ALT."""
    tmp = str(tmp_path)
    ref = create_ref_fasta(os.path.join(tmp, "ref.fa"))
    alts = [b for b in "ACGT" if b != ref[50]]
    s = list(ref[30:90]); s[20] = alts[1]
    var = "".join(s)
    p = _trio(tmp, ref, [("read%d" % i, 30, var) for i in range(4)])
    vcf = os.path.join(tmp, "in.vcf")
    create_vcf(vcf, "chr1", [(51, ref[50], alts[0] + "," + alts[1])], genotypes=["0/2"])
    res, vals = _vcf(eng, tmp, p["child"], p["mother"], p["father"], vcf)
    assert int(vals[0]["DKA"]) == 4 and int(vals[0]["DKU"]) == 4
    create_vcf(vcf, "chr1", [(51, ref[50], alts[0] + "," + alts[1])], genotypes=["0/1"])
    res1, vals1 = _vcf(eng, tmp, p["child"], p["mother"], p["father"], vcf)
    assert int(vals1[0]["DKA"]) == 0 and int(vals1[0]["DKU"]) == 4
