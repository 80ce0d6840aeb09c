"""VCF mode end to end on the GPU (BASELINE config 1: GIAB mini trio, k = 31)
against the reference's committed goldens, plus the parent scan against the
oracle for 64- and 128-bit keys."""
import argparse
import gzip
import json
import os

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from kmer_denovo_filter_b200 import engine
    return engine.CudaEngine()


def test_vcf_mode_reproduces_goldens(eng, giab_paths, tmp_path):
    from kmer_denovo_filter_b200.vcf import pipeline as P
    args = argparse.Namespace(
        child=giab_paths["child"], mother=giab_paths["mother"], father=giab_paths["father"],
        vcf=giab_paths["vcf"], output=str(tmp_path / "annotated.vcf.gz"),
        metrics=str(tmp_path / "metrics.json"), summary=str(tmp_path / "summary.txt"),
        ref_fasta=None, kmer_size=31, min_baseq=20, min_mapq=20, threads=4, debug_kmers=False,
        proband_id="HG002", tmp_dir=None)
    res = P.run_pipeline(args, engine=eng)
    exp = giab_paths["expected_vcf"]
    want = json.load(open(os.path.join(exp, "metrics.json")))
    assert res["metrics"] == want == json.load(open(args.metrics))
    assert open(args.summary).read() == open(os.path.join(exp, "summary.txt")).read()
    got = gzip.open(res["paths"]["vcf"], "rt").read().splitlines()
    assert got == gzip.open(os.path.join(exp, "annotated.vcf.gz"), "rt").read().splitlines()
    # the k-mers found in parents and their summed counts (MAX/AVG/MIN_PKC come from these)
    assert len(res["parent_found_kmers"]) == 1294
    assert max(res["parent_found_kmers"].values()) == 2177


@pytest.mark.parametrize("k", [31, 47])
def test_scan_parent_equals_oracle(eng, giab_paths, giab_records, tmp_path, k):
    from kmer_denovo_filter_b200.core.kmer_engine_wrappers import _scan_parent_jellyfish
    from kmer_denovo_filter_b200.vcf import pipeline as P
    from oracle import vcf as ovcf
    variants = P._parse_vcf_variants(giab_paths["vcf"], "HG002")
    fa = str(tmp_path / "k.fa")
    total, _vrk = P._collect_child_kmers(giab_paths["child"], None, variants, k, 20, 20, False, fa)
    filt = [l.strip() for l in open(fa) if not l.startswith(">")]
    assert total == len(filt) > 500
    got = _scan_parent_jellyfish(giab_paths["father"], None, fa, k, None, 4, n_filter_kmers=total,
                                 engine=eng)
    want = ovcf.parent_counts(giab_records["father"], k, filt)
    assert got == want
    assert all(v >= 1 for v in got.values())


def test_vcf_and_discovery_on_a_synthetic_bam_trio_equal_the_oracle(eng, tmp_path):
    """BASELINE config 3 in small (200 kbp x 30x trio written as sorted, indexed BAMs, the
    injected de novo events as the candidate VCF): VCF mode — every variant's DKU / DKT /
    DKA / PKC fields, the metrics and the DV-tagged informative reads — and discovery mode
    — stage sizes and BED rows — against the oracle run on the same files read with the
    oracle's own stdlib BAM reader."""
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench_wall
    from kmer_denovo_filter_b200.discovery import pipeline as DP
    from kmer_denovo_filter_b200.vcf import pipeline as P
    from oracle import bam as obam, discovery as odisc, vcf as ovcf
    genome, depth, L, k = 200_000, 30, 150, 31        # (the oracle is pure Python: ~50 s at this size)
    paths, events, _st = bench_wall.make_bam_trio(torch, eng.device, genome, depth, L, 8, str(tmp_path), threads=4)
    vcf_path = str(tmp_path / "truth.vcf")
    bench_wall.write_truth_vcf(vcf_path, bench_wall.contigs_for(genome), events)
    recs = {w: obam.read_bam(paths[w])[2] for w in ("child", "mother", "father")}
    # ---- VCF mode
    _h, _s, variants = ovcf.parse_vcf(vcf_path, "child")
    want_ann, want_metrics, _found = ovcf.run(recs["child"], recs["mother"], recs["father"], variants, k)
    res = P.run_pipeline(bench_wall.vcf_args(paths, vcf_path, str(tmp_path), k, threads=4), engine=eng)
    assert res["metrics"] == want_metrics
    assert res["annotations"] == want_ann
    assert all(a["dku"] > 0 for a in want_ann.values())                  # every injected event is found
    n_snv = sum(1 for e in events if e[2] == "snv")
    assert sum(1 for a in want_ann.values() if a["dka"] > 0) >= n_snv - 1
    # ---- discovery mode on the same files
    ref_seqs = [s for _n, s in obam.read_fasta(paths["ref"])]
    want = odisc.run(recs["child"], recs["mother"], recs["father"], ref_seqs, k)
    m = DP.run_discovery_pipeline(bench_wall.discovery_args(paths, str(tmp_path / "disc"), k, threads=4), engine=eng)
    assert m["child_candidate_kmers"] == len(want["candidates"])
    assert m["non_ref_kmers"] == len(want["non_ref"])
    assert m["proband_unique_kmers"] == len(want["proband_unique"])
    assert m["informative_reads"] == want["informative"]
    rows = [l.rstrip("\n").split("\t") for l in open(str(tmp_path / "disc.bed")) if not l.startswith("#")]
    assert rows == [[str(x) for x in r] for r in want["bed"]]
