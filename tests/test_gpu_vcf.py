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
