"""End-to-end discovery pipeline on the GPU against the reference's committed
golden outputs (BASELINE config 2: GIAB mini trio, k=31, min-child-count 3) and
against the oracle on other k (config 5 parity: 64- vs 128-bit keys)."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _digest(keys):
    h = hashlib.sha256()
    for key in sorted(keys):
        h.update(int(key).to_bytes(16, "little"))
    return h.hexdigest()


@pytest.fixture(scope="module")
def eng():
    from kmer_denovo_filter_b200 import engine
    return engine.CudaEngine()


def test_discovery_chain_stage_sets(eng, giab_paths):
    """Every intermediate k-mer set equals the oracle's (digest in giab_expected.json)."""
    from kmer_denovo_filter_b200.discovery import pipeline as P
    exp = json.load(open(giab_paths["expected_json"]))
    k = 31
    ref = P._ensure_ref_jf(giab_paths["ref_fasta"], k, 4, None, eng)
    cand, n_cand = P._extract_child_kmers_discovery(giab_paths["child"], None, k, 3, 4, None,
                                                    engine=eng)
    assert n_cand == exp["candidates"] == 51125
    assert cand.stats["new"] == exp["child_distinct"]
    assert cand.stats["windows"] == exp["child_total"]
    d = cand.dump()
    keys = eng.keys_to_pyints(d["lo"], d["hi"])
    counts = d["p0"].cpu().numpy().view(np.uint32).tolist()
    assert d["n_out"] == exp["child_distinct"]
    assert _digest([(key << 32) | c for key, c in zip(keys, counts)]) == exp["child_counts_digest"]
    d = cand.dump(min0=3)
    assert _digest(eng.keys_to_pyints(d["lo"], d["hi"])) == exp["candidates_digest"]
    non_ref, n_non_ref = P._subtract_reference_kmers(ref, cand, None)
    assert n_non_ref == exp["non_ref"] == 6679
    assert _digest(non_ref.to_pyints()) == exp["non_ref_digest"]
    n_pu, pu = P._filter_parents_discovery(giab_paths["mother"], giab_paths["father"], None,
                                           non_ref, k, 4, None, 0, engine=eng)
    assert n_pu == exp["proband_unique"] == 630
    assert sorted(pu.to_strings()) == exp["proband_unique_kmers"]


def test_reference_index_from_jf_file(eng, giab_paths):
    """--ref-jf without a FASTA: the Jellyfish binary/sorted file is parsed."""
    from kmer_denovo_filter_b200.discovery import pipeline as P
    ref = P._ensure_ref_jf(None, 31, 4, giab_paths["ref_jf"], eng)
    assert ref.keys is not None and ref.keys[0].shape[0] == 45275
    cand, n_cand = P._extract_child_kmers_discovery(giab_paths["child"], None, 31, 3, 4, None,
                                                    engine=eng)
    _non_ref, n_non_ref = P._subtract_reference_kmers(ref, cand, None)
    assert (n_cand, n_non_ref) == (51125, 6679)


def _args(giab_paths, out_prefix, **kw):
    from kmer_denovo_filter_b200 import cli
    argv = ["--child", giab_paths["child"], "--mother", giab_paths["mother"],
            "--father", giab_paths["father"], "--ref-fasta", giab_paths["ref_fasta"],
            "--out-prefix", out_prefix, "--min-child-count", "3", "--kmer-size", "31",
            "--candidate-summary", os.path.join(giab_paths["expected_vcf"], "summary.txt")]
    for key, v in kw.items():
        argv += ["--" + key.replace("_", "-"), str(v)]
    return cli.parse_discovery_args(argv)


def test_discovery_pipeline_golden_files(eng, giab_paths, tmp_path):
    from kmer_denovo_filter_b200.discovery import pipeline as P
    prefix = str(tmp_path / "giab_discovery")
    metrics = P.run_discovery_pipeline(_args(giab_paths, prefix), engine=eng)
    gold_dir = giab_paths["expected_discovery"]
    gold = json.load(open(os.path.join(gold_dir, "giab_discovery.metrics.json")))
    assert metrics == gold
    assert json.load(open(prefix + ".metrics.json")) == gold
    for suffix in (".bed", ".kmer_coverage.bedgraph", ".read_coverage.bed", ".sv.bedpe",
                   ".summary.txt", ".metrics.json"):
        got = open(prefix + suffix).read()
        want = open(os.path.join(gold_dir, "giab_discovery" + suffix)).read()
        assert got == want, suffix
    # per-read unique counts (north-star parity item) against the oracle's
    exp = json.load(open(giab_paths["expected_json"]))
    got = sorted(P._anchor_and_cluster.last_per_read)
    assert [list(x) for x in got] == sorted(exp["per_read_informative"])


@pytest.mark.parametrize("k", [21, 47, 63])
def test_discovery_other_k_against_oracle(eng, giab_paths, giab_records, tmp_path, k):
    from kmer_denovo_filter_b200.discovery import pipeline as P
    from oracle import discovery
    want = discovery.run(giab_records["child"], giab_records["mother"], giab_records["father"],
                         [s for _n, s in giab_records["ref"]], k)
    prefix = str(tmp_path / ("k%d" % k))
    m = P.run_discovery_pipeline(_args(giab_paths, prefix, kmer_size=k), engine=eng)
    assert m["child_candidate_kmers"] == len(want["candidates"])
    assert m["non_ref_kmers"] == len(want["non_ref"])
    assert m["proband_unique_kmers"] == len(want["proband_unique"])
    assert m["informative_reads"] == want["informative"]
    assert m["unmapped_informative_reads"] == want["unmapped_informative"]
    rows = [l.rstrip("\n").split("\t") for l in open(prefix + ".bed") if not l.startswith("#")]
    assert rows == [[str(x) for x in r] for r in want["bed"]]
    rows = [l.rstrip("\n").split("\t") for l in open(prefix + ".kmer_coverage.bedgraph")
            if not l.startswith("#")]
    assert rows == [[str(x) for x in r] for r in want["bedgraph"]]


def test_parent_max_count_and_filters(eng, giab_paths, giab_records, tmp_path):
    from kmer_denovo_filter_b200.discovery import pipeline as P
    from oracle import discovery
    want = discovery.run(giab_records["child"], giab_records["mother"], giab_records["father"],
                         [s for _n, s in giab_records["ref"]], 31, parent_max_count=1,
                         min_dk_per_read=3, merge_distance=0, min_supporting_reads=2)
    prefix = str(tmp_path / "pmc1")
    m = P.run_discovery_pipeline(
        _args(giab_paths, prefix, parent_max_count=1, min_distinct_kmers_per_read=3,
              cluster_distance=0, min_supporting_reads=2), engine=eng)
    assert m["proband_unique_kmers"] == len(want["proband_unique"])
    assert m["informative_reads"] == want["informative"]
    rows = [l.rstrip("\n").split("\t") for l in open(prefix + ".bed") if not l.startswith("#")]
    assert rows == [[str(x) for x in r] for r in want["bed"]]


def test_gpu_kmer_query_interface(eng, giab_paths):
    """Duck-typed query object (reference tests/discovery/test_pipeline.py:1532-1542)."""
    from kmer_denovo_filter_b200.kmer_utils import GpuKmerQuery, KmerSet, canonicalize
    exp = json.load(open(giab_paths["expected_json"]))
    pu = exp["proband_unique_kmers"]
    q = GpuKmerQuery(KmerSet.from_strings(eng, 31, pu))
    probe = pu[:50] + ["A" * 31, "ACGT" * 7 + "ACG", "N" * 31]
    assert q.query_batch(probe) == set(pu[:50])
    assert q.query_batch([]) == set()
    seq = "TTGACC" + pu[3] + "G" + pu[3][:10]
    uniq, idx = q.scan_read(seq, 31)
    want = {i: canonicalize(seq[i:i + 31]) for i in range(len(seq) - 30)
            if canonicalize(seq[i:i + 31]) in set(pu)}
    assert uniq == set(want.values()) and idx == set(want)
    q.close()
    assert q.query_batch(pu[:3]) == set(pu[:3])  # close() only clears caches


def test_informative_reads_bam(eng, giab_paths, giab_records, oracle_discovery, tmp_path):
    """{prefix}.informative.bam: the reads with >= 1 proband-unique k-mer (primary +
    supplementary, non-duplicate, first per (qname, is_supplementary)), tagged
    dk:i:1, coordinate-sorted, with a .bai whose chunks land on record boundaries.
    Read back with the oracle's BAM reader."""
    import struct
    from kmer_denovo_filter_b200.discovery import pipeline as P
    from kmer_denovo_filter_b200.kmer_utils import KmerSet
    from oracle import bam as obam, discovery as odisc
    pu = oracle_discovery["proband_unique"]
    k = 31
    kset = KmerSet(eng, k, *eng.keys_to_device(sorted(pu), 1))
    out = str(tmp_path / "giab.informative.bam")
    n = P._write_informative_reads_discovery(giab_paths["child"], None, kset, k, out, engine=eng)
    # expected selection straight from the oracle's records
    pu_keys = set(kset.to_pyints())
    want, seen = [], set()
    for r in giab_records["child"]:
        if r.is_secondary or r.is_duplicate or not r.seq:
            continue
        uniq, _idx = odisc.scan_read_numeric(r.seq, k, pu_keys)
        key = (r.qname, r.is_supplementary)
        if uniq and key not in seen:
            seen.add(key)
            want.append((r.ref_id, r.pos, r.qname, r.flag))
    names, lens, got = obam.read_bam(out)
    assert n == len(got) == len(want) > 150
    assert sorted((r.ref_id, r.pos, r.qname, r.flag) for r in got) == sorted(want)
    keys = [((r.ref_id if r.ref_id >= 0 else 1 << 31), r.pos) for r in got]
    assert keys == sorted(keys)
    assert all(r.get_tag("dk") == 1 for r in got)
    src = {(r.qname, r.flag): r for r in giab_records["child"]}
    for r in got[:50]:
        o = src[(r.qname, r.flag)]
        assert (r.seq, r.cigar, r.mapq, r.next_pos, r.tlen) == (o.seq, o.cigar, o.mapq, o.next_pos, o.tlen)
    # the index parses and every chunk starts inside the file
    bai = open(out + ".bai", "rb").read()
    assert bai[:4] == b"BAI\x01"
    n_ref = struct.unpack_from("<i", bai, 4)[0]
    assert n_ref == len(names)
    size = os.path.getsize(out)
    off, n_chunks = 8, 0
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", bai, off)[0]; off += 4
        for _b in range(n_bin):
            _bin, n_chunk = struct.unpack_from("<Ii", bai, off); off += 8
            for _c in range(n_chunk):
                v0, v1 = struct.unpack_from("<QQ", bai, off); off += 16
                assert (v0 >> 16) < size and v0 < v1
                n_chunks += 1
        n_intv = struct.unpack_from("<i", bai, off)[0]; off += 4 + 8 * n_intv
    assert n_chunks > 0 and off + 8 == len(bai)
