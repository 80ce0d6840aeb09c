"""CPU tests of the product's host-side post-processing (cluster, annotate,
classify, writers) fed with the oracle's read hits, against the goldens."""
import json
import os

import numpy as np

from kmer_denovo_filter_b200.discovery import pipeline as P
from kmer_denovo_filter_b200.discovery import summary as S
from oracle import discovery, kmers


def test_cluster_annotate_write_match_golden(giab_paths, giab_records, oracle_discovery, tmp_path):
    r = oracle_discovery
    a = discovery.anchor(giab_records["child"], 31, r["proband_unique"], 7)
    hits = [(c, s, e, q, {kmers.kmer_of(x, 31) for x in ks}, supp)
            for c, s, e, q, ks, supp in a["read_hits"]]
    regions, rreads, rkmers = P._cluster_hits(list(hits), 500)
    ann, links = P._annotate_and_link_from_metadata(regions, rreads, a["read_sv_meta"])
    P._classify_regions(regions, ann, links)
    gold = giab_paths["expected_discovery"]
    bed = str(tmp_path / "o.bed")
    P._write_bed(regions, rreads, rkmers, bed, ann,
                 {"min_distinct_kmers_per_read": 7, "min_supporting_reads": 1,
                  "min_distinct_kmers": 1})
    assert open(bed).read() == open(os.path.join(gold, "giab_discovery.bed")).read()
    bg = str(tmp_path / "o.bedgraph")
    P._write_bedgraph(a["kmer_coverage"], bg, a["read_coverage"], 3)
    assert open(bg).read() == open(os.path.join(gold, "giab_discovery.kmer_coverage.bedgraph")).read()
    rc = str(tmp_path / "o.rc.bed")
    P._write_read_coverage_bed(a["kmer_coverage"], a["read_coverage"], rc, 3)
    assert open(rc).read() == open(os.path.join(gold, "giab_discovery.read_coverage.bed")).read()
    pe = str(tmp_path / "o.bedpe")
    P._write_bedpe(links, pe)
    assert open(pe).read() == open(os.path.join(gold, "giab_discovery.sv.bedpe")).read()
    metrics = json.load(open(os.path.join(gold, "giab_discovery.metrics.json")))
    comp = P._compare_candidates_to_regions(
        P._parse_candidate_summary(os.path.join(giab_paths["expected_vcf"], "summary.txt")), regions)
    dnm = P._evaluate_dnm_regions(regions, metrics["regions"])
    assert dnm == metrics["dnm_evaluation"]["loci"]
    assert [c["captured"] for c in comp] == [True, True, True]
    text = S._write_discovery_summary(str(tmp_path / "s.txt"), regions, rreads, rkmers, metrics,
                                      comp, ann, dnm)
    assert text == open(os.path.join(gold, "giab_discovery.summary.txt")).read()


def test_kmer_ref_positions_with_insertion():
    """reference tests/discovery/test_pipeline.py:1383-1439 (insertion splits coverage)."""
    class R:
        reference_start = 100
        cigartuples = [(0, 10), (1, 3), (0, 10)]
    cov = P._collect_kmer_ref_positions(R, [8], 5)      # q 8..12: 8,9 aligned; 10..12 inserted
    assert dict(cov) == {108: 1, 109: 1}
    cov = P._collect_kmer_ref_positions(R, [0, 1], 5)
    assert dict(cov) == {100: 1, 101: 2, 102: 2, 103: 2, 104: 2, 105: 1}
    assert dict(P._collect_kmer_ref_positions(R, [], 5)) == {}


def _random_alignments(seed, n_reads=60, k=7):
    """Reads with random CIGARs (all op kinds) and sorted random hit offsets."""
    import random
    rng = random.Random(seed)
    contig, start, cig_off, cigar, hr, ho, recs = [], [], [0], [], [], [], []
    for r in range(n_reads):
        ops = []
        if rng.random() < 0.3:
            ops.append((5, rng.randint(1, 9)))
        if rng.random() < 0.5:
            ops.append((4, rng.randint(1, 12)))
        for _ in range(rng.randint(1, 6)):
            ops.append((rng.choice([0, 0, 0, 7, 8]), rng.randint(1, 30)))
            if rng.random() < 0.6:
                ops.append((rng.choice([1, 2, 3, 6]), rng.randint(1, 8)))
        if rng.random() < 0.4:
            ops.append((4, rng.randint(1, 12)))
        qlen = sum(ln for op, ln in ops if op in (0, 1, 4, 7, 8))
        c = rng.randint(0, 3)
        st = rng.randint(0, 5000)
        offs = sorted(rng.sample(range(0, max(qlen - k + 1, 1)), min(rng.randint(0, 6), max(qlen - k + 1, 1))))
        contig.append(c)
        start.append(st)
        cigar += [(ln << 4) | op for op, ln in ops]
        cig_off.append(len(cigar))
        hr += [r] * len(offs)
        ho += offs
        recs.append((c, st, ops, offs))
    return contig, start, cig_off, cigar, hr, ho, recs


def test_hit_coverage_host_matches_reference_helper():
    """K7 (kdf_hit_coverage): the device function, instantiated on the host, against the
    reference-named per-read helper and its Counter merges (core/bam_scanner.py:97-117,
    discovery/pipeline.py:851-855), on random CIGARs with clips, insertions, deletions,
    skips and padding."""
    import collections
    from kmer_denovo_filter_b200 import engine
    for seed, k in ((1, 7), (2, 31), (3, 1)):
        contig, start, cig_off, cigar, hr, ho, recs = _random_alignments(seed, k=k)
        kc = collections.defaultdict(collections.Counter)
        rc = collections.defaultdict(collections.Counter)
        for c, st, ops, offs in recs:
            class R:
                reference_start = st
                cigartuples = ops
            cov = P._collect_kmer_ref_positions(R, offs, k)
            kc[c] += cov
            for p in cov:
                rc[c][p] += 1
        gc, gp, gk, gr = engine.debug_hit_coverage_host(hr, ho, k, contig, start, cig_off, cigar)
        got_k = collections.defaultdict(dict)
        got_r = collections.defaultdict(dict)
        for c, p, a, b in zip(gc.tolist(), gp.tolist(), gk.tolist(), gr.tolist()):
            got_k[c][p] = a
            got_r[c][p] = b
        assert {c: dict(v) for c, v in kc.items() if v} == dict(got_k)
        assert {c: dict(v) for c, v in rc.items() if v} == dict(got_r)
    z = engine.debug_hit_coverage_host([], [], 31, [0], [0], [0, 0], [])
    assert all(a.shape[0] == 0 for a in z)


def test_accumulate_coverage_glue_on_fixture_reads(giab_paths):
    """The pipeline's K7 glue (CIGAR gather, hit lists, Counter merges) on real records of
    the fixture, with the device call replaced by the host instantiation of the same code:
    equal to the reference-named helper applied read by read."""
    import collections
    from kmer_denovo_filter_b200 import bamio, engine

    class HostK7:
        def hit_coverage(self, *a):
            return engine.debug_hit_coverage_host(*a)

    k = 31
    with bamio.BamReader(giab_paths["child"], threads=2) as rd:
        batch = rd.next_batch(bamio.MODE_SCAN, want_meta=True)
    reads, slices, off = [], [], []
    for r in range(0, batch.n_reads, 23):
        rec = batch.record(r)
        if rec.is_unmapped or not rec.cigartuples or int(batch.read_lens[r]) < k + 40:
            continue
        offs = [3, 4, 20, int(batch.read_lens[r]) - k]
        reads.append(r)
        slices.append((len(off), len(off) + len(offs)))
        off += offs
    assert len(reads) > 100
    kc = collections.defaultdict(collections.Counter)
    rc = collections.defaultdict(collections.Counter)
    P._accumulate_coverage(HostK7(), batch, reads, slices, np.asarray(off, dtype=np.int64), k, kc, rc)
    want_k = collections.defaultdict(collections.Counter)
    want_r = collections.defaultdict(collections.Counter)
    for r, (a, b) in zip(reads, slices):
        rec = batch.record(r)
        cov = P._collect_kmer_ref_positions(rec, off[a:b], k)
        want_k[rec.reference_name] += cov
        for p in cov:
            want_r[rec.reference_name][p] += 1
    assert {c: dict(v) for c, v in kc.items()} == {c: dict(v) for c, v in want_k.items() if v}
    assert {c: dict(v) for c, v in rc.items()} == {c: dict(v) for c, v in want_r.items() if v}
    batch.close()


def test_sv_linking_and_classes():
    regions = [("chr1", 100, 300), ("chr1", 5000, 5200), ("chr2", 10, 90)]
    rreads = {regions[0]: {"a", "b"}, regions[1]: {"a"}, regions[2]: {"c"}}
    meta = {
        ("a", False): {"has_sa": True, "sa_str": "chr2,50,+,50M,60,0;", "is_paired": True,
                       "is_proper_pair": False, "mate_is_unmapped": False, "max_clip": 40},
        ("a", True): {"has_sa": True, "sa_str": None, "is_paired": True, "is_proper_pair": False,
                      "mate_is_unmapped": False, "max_clip": 10},
        ("b", False): {"has_sa": False, "sa_str": None, "is_paired": True, "is_proper_pair": True,
                       "mate_is_unmapped": True, "max_clip": 0},
        ("c", False): {"has_sa": False, "sa_str": None, "is_paired": False, "is_proper_pair": False,
                       "mate_is_unmapped": False, "max_clip": 3},
    }
    ann, links = P._annotate_and_link_from_metadata(regions, rreads, meta)
    assert ann[regions[0]] == {"split_reads": 1, "discordant_pairs": 2, "max_clip_len": 40,
                               "unmapped_mates": 1}
    pairs = [(l["region_a"], l["region_b"], l["sv_type_hint"]) for l in links]
    assert (regions[0], regions[1], "INTRA") in pairs and (regions[0], regions[2], "BND") in pairs
    P._classify_regions(regions, ann, links)
    assert [ann[r]["class"] for r in regions] == ["SV", "SV", "SV"]
    ann2, links2 = P._annotate_and_link_from_metadata([regions[2]], {regions[2]: {"c"}}, meta)
    P._classify_regions([regions[2]], ann2, links2)
    assert ann2[regions[2]]["class"] == "SMALL"


def test_cli_defaults_match_reference():
    from kmer_denovo_filter_b200 import cli
    a = cli.parse_discovery_args(["--child", "c", "--mother", "m", "--father", "f",
                                  "--out-prefix", "o"])
    assert (a.kmer_size, a.min_child_count, a.cluster_distance, a.min_supporting_reads,
            a.min_distinct_kmers, a.min_bedgraph_reads, a.parent_max_count, a.threads,
            a.min_distinct_kmers_per_read, a.min_baseq) == (31, 3, 500, 1, 1, 3, 0, 4, None, 20)
    v = cli.parse_vcf_args(["--child", "c", "--mother", "m", "--father", "f", "--vcf", "v",
                            "--output", "o"])
    assert (v.kmer_size, v.min_mapq, v.min_baseq, v.proband_id) == (31, 20, 20, None)
