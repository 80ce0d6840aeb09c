"""The synthetic-BAM writer of the wall-time benchmark (bench_wall.py) on CPU: the BAMs it
writes are valid, coordinate sorted, and carry exactly the reads of the packed streams
bench.py times."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_bam_trio_matches_the_packed_streams(tmp_path):
    import bench_wall
    from kmer_denovo_filter_b200 import bamio, engine, synth
    from oracle import bam as obam
    genome, depth, L = 150_000, 8, 100
    dev = torch.device("cpu")
    paths, events, stats = bench_wall.make_bam_trio(torch, dev, genome, depth, L, 12, str(tmp_path), threads=2)
    trio = synth.make_trio(torch, dev, genome, depth=depth, read_len=L, n_denovo=12)
    assert len(events) == 12 and all(e[0] == "chr1" for e in events)
    for who in ("child", "mother", "father"):
        refs, lens, recs = None, None, None
        with bamio.BamReader(paths[who], threads=2) as rd:
            assert rd.references == ["chr1"] and rd.lengths == [genome]
            b = rd.next_batch(bamio.MODE_FASTA, want_meta=True)
        s = trio[who]
        assert b.n_reads == int(s["read_starts"].shape[0]) == stats[who]["reads"]
        assert (np.diff(b.pos.astype(np.int64)) >= 0).all()              # coordinate sorted
        assert set(b.flag.tolist()) == {99, 147}
        # same k-mer multiset as the packed stream bench.py times (reads are a permutation,
        # reverse-strand reads are stored reverse-complemented: canonical k-mers do not care)
        hs = engine.HostStream(s["codes"].numpy().view(np.uint64), s["valid"].numpy().view(np.uint32),
                               s["n_bases"], s["read_starts"].numpy().view(np.uint64),
                               s["read_lens"].numpy().view(np.uint32))
        lo_a, _h, ok_a = engine.debug_extract_host(hs, 21)
        lo_b, _h, ok_b = engine.debug_extract_host(b, 21)
        assert np.array_equal(np.sort(lo_a[ok_a]), np.sort(lo_b[ok_b]))
        # the stdlib reader of the oracle agrees record by record
        orecs = obam.read_bam(paths[who])[-1] if isinstance(obam.read_bam(paths[who]), tuple) else obam.read_bam(paths[who])
        assert len(orecs) == b.n_reads
        for i in (0, 1, b.n_reads // 2, b.n_reads - 1):
            assert orecs[i].qname == b.record(i).query_name and orecs[i].pos == int(b.pos[i])
            assert orecs[i].cigar == [(0, L)]
    # parents carry no indels: a read sits where it was sampled, so its bases equal the
    # reference up to SNPs / errors; check the mapping on the child across its indels instead
    ref_seq = open(paths["ref"]).read().split("\n", 1)[1].replace("\n", "")
    assert len(ref_seq) == genome
    with bamio.BamReader(paths["child"], threads=2) as rd:
        b = rd.next_batch(bamio.MODE_FASTA, want_meta=True)
    mism = []
    for i in range(0, b.n_reads, 7):
        r = b.record(i)
        seq = r.query_sequence
        p = int(b.pos[i])
        ref = ref_seq[p:p + L]
        mism.append(sum(1 for x, y in zip(seq, ref) if x != y))
    mism = np.array(mism)
    assert np.median(mism) <= 1 and (mism <= 3).mean() > 0.97      # only reads over an indel disagree


def test_synthetic_bai_supports_fetch_and_shards(tmp_path):
    """The .bai bench_wall writes for its BAMs: region fetches return exactly the reads over
    a site, and rank shards partition the file."""
    import bench_wall
    from kmer_denovo_filter_b200 import bamio
    genome, depth, L = 300_000, 10, 100
    paths, events, _stats = bench_wall.make_bam_trio(torch, torch.device("cpu"), genome, depth, L, 10,
                                                     str(tmp_path), threads=2)
    assert os.path.isfile(paths["child"] + ".bai")
    with bamio.BamReader(paths["child"], threads=2) as rd:
        whole = rd.next_batch(bamio.MODE_ALL, want_meta=True)
        pos_all = whole.pos.astype(np.int64)
        names = [whole.record(i).query_name for i in range(whole.n_reads)]
        for site in (500, 16384, 16385, 100_000, 163_840, genome - 1000):
            want = sorted((names[i], int(pos_all[i])) for i in np.flatnonzero((pos_all <= site) & (site < pos_all + L)))
            got = []
            for b in rd.fetch(0, site, site + 1, want_meta=True):
                p = b.pos.astype(np.int64)
                got += [(b.record(i).query_name, int(p[i])) for i in np.flatnonzero((p <= site) & (site < p + L))]
                b.close()
            assert sorted(got) == want and want
    total = 0
    for r in range(3):
        rd = bamio.open_shard(paths["child"], r, 3, threads=2)
        n = sum(b.n_reads for b in rd.batches(bamio.MODE_FASTA, max_bases=1 << 20))
        rd.close()
        assert n > 0
        total += n
    assert total == whole.n_reads
