"""Pin the CPU oracle against the reference's committed golden outputs.

Golden sources (produced by the real samtools + Jellyfish + pysam stack in the
reference CI): reference ``tests/example_output_discovery/*`` (compared there by
``tests/test_example_output_discovery.py:48-263``) and the genuine Jellyfish
table ``tests/data/giab/mini_ref.fa.k31.jf``.
"""
import json
import os

from oracle import bam, kmers


def _rows(path):
    return [l.rstrip("\n").split("\t") for l in open(path) if not l.startswith("#")]


def test_reference_jf_table_equals_oracle_count(giab_paths, giab_records):
    k, jf = kmers.read_jf_binary_sorted(giab_paths["ref_jf"])
    assert k == 31
    mine = kmers.count_sequences([s for _n, s in giab_records["ref"]], 31)
    assert len(jf) == 45275 and sum(jf.values()) == 45804 and max(jf.values()) == 12
    assert mine == jf


def test_discovery_stage_counts(giab_paths, oracle_discovery):
    gold = json.load(open(os.path.join(giab_paths["expected_discovery"],
                                       "giab_discovery.metrics.json")))
    r = oracle_discovery
    assert len(r["candidates"]) == gold["child_candidate_kmers"] == 51125
    assert len(r["non_ref"]) == gold["non_ref_kmers"] == 6679
    assert len(r["proband_unique"]) == gold["proband_unique_kmers"] == 630
    assert r["informative"] == gold["informative_reads"] == 195
    assert r["unmapped_informative"] == gold["unmapped_informative_reads"] == 11
    assert len(r["regions"]) == gold["candidate_regions"] == 21


def test_discovery_bed_bedgraph_readcov(giab_paths, oracle_discovery):
    e = giab_paths["expected_discovery"]
    r = oracle_discovery
    assert _rows(os.path.join(e, "giab_discovery.bed")) == \
        [[str(x) for x in row] for row in r["bed"]]
    assert _rows(os.path.join(e, "giab_discovery.kmer_coverage.bedgraph")) == \
        [[str(x) for x in row] for row in r["bedgraph"]]
    assert _rows(os.path.join(e, "giab_discovery.read_coverage.bed")) == \
        [[str(x) for x in row] for row in r["read_coverage_bed"]]
    assert r["links"] == []  # golden BEDPE is header-only


def test_fasta_stream_collapse_is_required(giab_records):
    """Without the same-QNAME collapse the goldens are NOT reproduced
    (SURVEY 'Five facts' 4): 51223 candidates instead of 51125."""
    recs = [r for r in giab_records["child"] if not (r.flag & 0xD00)]
    counts = kmers.count_sequences([r.seq for r in recs], 31)
    assert sum(1 for c in counts.values() if c >= 3) == 51223
    assert len(recs) - len(bam.fasta_stream(giab_records["child"])) == 227 \
        or len(recs) > len(bam.fasta_stream(giab_records["child"]))


def test_expected_json_is_current(giab_paths, oracle_discovery):
    exp = json.load(open(giab_paths["expected_json"]))
    r = oracle_discovery
    assert exp["candidates"] == len(r["candidates"])
    assert exp["proband_unique_kmers"] == sorted(kmers.kmer_of(x, 31) for x in r["proband_unique"])
    assert exp["child_distinct"] == len(r["child_counts"]) == 282880
