"""Shared pytest configuration: markers and golden-fixture paths."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GIAB = os.path.join(GOLDEN, "giab")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def giab_paths():
    return {
        "child": os.path.join(GIAB, "HG002_child.bam"),
        "mother": os.path.join(GIAB, "HG004_mother.bam"),
        "father": os.path.join(GIAB, "HG003_father.bam"),
        "ref_fasta": os.path.join(GIAB, "mini_ref.fa"),
        "ref_jf": os.path.join(GIAB, "mini_ref.fa.k31.jf"),
        "vcf": os.path.join(GIAB, "candidates.vcf.gz"),
        "expected_discovery": os.path.join(GOLDEN, "expected_discovery"),
        "expected_vcf": os.path.join(GOLDEN, "expected_vcf"),
        "expected_json": os.path.join(GOLDEN, "giab_expected.json"),
    }


@pytest.fixture(scope="session")
def giab_records(giab_paths):
    """Decoded GIAB trio (oracle BAM reader) shared across the session."""
    from oracle import bam
    out = {}
    for who in ("child", "mother", "father"):
        names, lens, recs = bam.read_bam(giab_paths[who])
        out[who] = recs
        out[who + "_refs"] = (names, lens)
    out["ref"] = bam.read_fasta(giab_paths["ref_fasta"])
    return out


@pytest.fixture(scope="session")
def oracle_discovery(giab_records):
    """Oracle discovery run on the GIAB mini trio (k=31, defaults)."""
    from oracle import discovery
    return discovery.run(giab_records["child"], giab_records["mother"],
                         giab_records["father"],
                         [s for _n, s in giab_records["ref"]], 31)
