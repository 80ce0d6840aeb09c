#!/usr/bin/env python
"""Regenerate tests/golden/ from the upstream reference checkout.

Run in the authoring container only (``/root/reference`` does not exist on the
GPU box):  ``python tests/golden/make_golden.py``

What it copies (DATA fixtures, not source code): the reference's public GIAB
mini-trio inputs (``tests/data/giab``) and the outputs its CI produced with the
real samtools + Jellyfish + pysam stack (``tests/example_output*``).  These are
the golden vectors that pin the oracle.

What it derives: ``giab_expected.json`` — stage-by-stage k-mer set sizes and
SHA-256 digests of the sorted k-mer sets computed by the oracle *after* the
oracle has been checked equal to the reference goldens (stage counts, BED,
bedGraph, read-coverage BED, and the Jellyfish ``mini_ref.fa.k31.jf`` table).
The digests let GPU-box tests check whole sets without shipping them.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("KDF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

COPY = {
    "giab": [
        "tests/data/giab/HG002_child.bam",
        "tests/data/giab/HG003_father.bam",
        "tests/data/giab/HG004_mother.bam",
        "tests/data/giab/HG002_child.bam.bai",
        "tests/data/giab/HG003_father.bam.bai",
        "tests/data/giab/HG004_mother.bam.bai",
        "tests/data/giab/mini_ref.fa",
        "tests/data/giab/mini_ref.fa.k31.jf",
        "tests/data/giab/candidates.vcf.gz",
    ],
    "expected_vcf": [
        "tests/example_output/metrics.json",
        "tests/example_output/summary.txt",
        "tests/example_output/annotated.vcf.gz",
    ],
    "expected_discovery": [
        "tests/example_output_discovery/giab_discovery.bed",
        "tests/example_output_discovery/giab_discovery.metrics.json",
        "tests/example_output_discovery/giab_discovery.summary.txt",
        "tests/example_output_discovery/giab_discovery.sv.bedpe",
        "tests/example_output_discovery/giab_discovery.kmer_coverage.bedgraph",
        "tests/example_output_discovery/giab_discovery.read_coverage.bed",
    ],
}


def digest(keys):
    h = hashlib.sha256()
    for key in sorted(keys):
        h.update(key.to_bytes(16, "little"))
    return h.hexdigest()


def main():
    for sub, files in COPY.items():
        os.makedirs(os.path.join(HERE, sub), exist_ok=True)
        for f in files:
            shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, sub, os.path.basename(f)))

    from oracle import bam, discovery, kmers
    g = os.path.join(HERE, "giab")
    k = 31
    ref = [s for _n, s in bam.read_fasta(os.path.join(g, "mini_ref.fa"))]
    _, _, child = bam.read_bam(os.path.join(g, "HG002_child.bam"))
    _, _, mother = bam.read_bam(os.path.join(g, "HG004_mother.bam"))
    _, _, father = bam.read_bam(os.path.join(g, "HG003_father.bam"))
    r = discovery.run(child, mother, father, ref, k)
    gold = json.load(open(os.path.join(HERE, "expected_discovery", "giab_discovery.metrics.json")))
    assert len(r["candidates"]) == gold["child_candidate_kmers"]
    assert len(r["non_ref"]) == gold["non_ref_kmers"]
    assert len(r["proband_unique"]) == gold["proband_unique_kmers"]
    assert r["informative"] == gold["informative_reads"]
    assert r["unmapped_informative"] == gold["unmapped_informative_reads"]
    kk, jf = kmers.read_jf_binary_sorted(os.path.join(g, "mini_ref.fa.k31.jf"))
    assert kk == k and jf == kmers.count_sequences(ref, k)
    exp = {
        "k": k,
        "child_distinct": len(r["child_counts"]),
        "child_total": sum(r["child_counts"].values()),
        "child_counts_digest": digest([(key << 32) | c for key, c in r["child_counts"].items()]),
        "candidates": len(r["candidates"]),
        "candidates_digest": digest(r["candidates"]),
        "non_ref": len(r["non_ref"]),
        "non_ref_digest": digest(r["non_ref"]),
        "after_mother": len(r["after_mother"]),
        "after_mother_digest": digest(r["after_mother"]),
        "proband_unique": len(r["proband_unique"]),
        "proband_unique_digest": digest(r["proband_unique"]),
        "proband_unique_kmers": sorted(kmers.kmer_of(x, k) for x in r["proband_unique"]),
        "ref_distinct": len(jf),
        "ref_total": sum(jf.values()),
        "scanned": r["scanned"],
        "per_read_informative": [[ridx, nd, len(idx)] for ridx, nd, idx in r["per_read"] if nd > 0],
    }
    with open(os.path.join(HERE, "giab_expected.json"), "w") as fh:
        json.dump(exp, fh, indent=1)
    print("golden fixtures regenerated under", HERE)


if __name__ == "__main__":
    main()
