"""The oracle's VCF mode against the reference's committed goldens
(BASELINE config 1: GIAB mini trio, k = 31): metrics.json and every
per-variant FORMAT value of annotated.vcf.gz."""
import gzip
import json
import os

import pytest

from oracle import vcf as ovcf

FIELDS = ["DKU", "DKT", "DKA", "DKU_DKT", "DKA_DKT", "MAX_PKC", "AVG_PKC", "MIN_PKC",
          "MAX_PKC_ALT", "AVG_PKC_ALT", "MIN_PKC_ALT"]


def golden_format_values(path):
    out = {}
    with gzip.open(path, "rt") as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            f = line.rstrip("\n").split("\t")
            fmt = f[8].split(":")
            val = f[9].split(":")
            out["%s:%d:%s:%s" % (f[0], int(f[1]) - 1, f[3], f[4])] = {k: val[fmt.index(k)] for k in FIELDS}
    return out


def ann_as_vcf_strings(a):
    g = lambda x: "%g" % x
    return {"DKU": str(a["dku"]), "DKT": str(a["dkt"]), "DKA": str(a["dka"]),
            "DKU_DKT": g(a["dku_dkt"]), "DKA_DKT": g(a["dka_dkt"]),
            "MAX_PKC": str(a["max_pkc"]), "AVG_PKC": g(a["avg_pkc"]), "MIN_PKC": str(a["min_pkc"]),
            "MAX_PKC_ALT": str(a["max_pkc_alt"]), "AVG_PKC_ALT": g(a["avg_pkc_alt"]),
            "MIN_PKC_ALT": str(a["min_pkc_alt"])}


@pytest.fixture(scope="module")
def oracle_vcf(giab_records, giab_paths):
    _h, _s, variants = ovcf.parse_vcf(giab_paths["vcf"], "HG002")
    ann, metrics, found = ovcf.run(giab_records["child"], giab_records["mother"],
                                   giab_records["father"], variants, 31)
    return variants, ann, metrics, found


def test_metrics_equal_golden(oracle_vcf, giab_paths):
    _v, _a, metrics, _f = oracle_vcf
    want = json.load(open(os.path.join(giab_paths["expected_vcf"], "metrics.json")))
    assert metrics == want


def test_every_format_value_equals_golden(oracle_vcf, giab_paths):
    variants, ann, _m, _f = oracle_vcf
    want = golden_format_values(os.path.join(giab_paths["expected_vcf"], "annotated.vcf.gz"))
    assert len(want) == len(variants) == 22
    for var in variants:
        key = ovcf.var_key(var)
        assert ann_as_vcf_strings(ann[key]) == want[key], key


def test_select_alt_from_gt_known_answers():
    """tests/vcf/test_pipeline.py multiallelic cases."""
    assert ovcf.select_alt_from_gt(("A", "C"), (0, 2)) == ("C", [2])
    assert ovcf.select_alt_from_gt(("A", "C"), (1, 2)) == ("A", [1, 2])
    assert ovcf.select_alt_from_gt(("A", "C"), (0, 0)) == ("A", [])
    assert ovcf.select_alt_from_gt(("A", "C"), None) == ("A", [])
    assert ovcf.is_symbolic("<DEL>") and ovcf.is_symbolic("*") and ovcf.is_symbolic("G]17:198982]")
    assert not ovcf.is_symbolic("ACGT") and not ovcf.is_symbolic(None)
