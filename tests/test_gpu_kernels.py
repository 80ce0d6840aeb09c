"""GPU parity tests: every kernel of libkdf_sm100.so, called through the C ABI
(ctypes), against the CPU oracle on the same seeded inputs.  Bit-exact."""
import random

import numpy as np
import pytest

from oracle import kmers

pytestmark = pytest.mark.gpu

KS = [5, 21, 31, 32, 33, 47, 63]


@pytest.fixture(scope="module")
def eng():
    from kmer_denovo_filter_b200 import engine
    return engine.CudaEngine()


def _rand_seqs(seed, n=400, maxlen=260, alphabet="ACGTACGTACGTACGTACGTN"):
    rng = random.Random(seed)
    return ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, maxlen))) for _ in range(n)]


def _genome_reads(seed, glen=20000, n=1500, rl=150, err=0.01):
    """Reads sampled from a random genome so that k-mers repeat (counts > 1)."""
    rng = random.Random(seed)
    g = "".join(rng.choice("ACGT") for _ in range(glen))
    out = []
    for _ in range(n):
        s = rng.randrange(0, glen - rl)
        r = list(g[s:s + rl])
        for i in range(rl):
            x = rng.random()
            if x < err:
                r[i] = rng.choice("ACGT")
            elif x < err + 0.001:
                r[i] = "N"
        r = "".join(r)
        if rng.random() < 0.5:
            r = kmers.reverse_complement(r)
        out.append(r)
    return g, out


def _table_dict(eng, table):
    n, lo, hi, p0, p1 = eng.threshold_compact(table, want_planes=True)
    keys = eng.keys_to_pyints(lo, hi)
    a = p0.cpu().numpy().view(np.uint32).tolist()
    b = p1.cpu().numpy().view(np.uint32).tolist()
    assert len(set(keys)) == len(keys), "duplicate keys in table"
    return {k: (x, y) for k, x, y in zip(keys, a, b)}


@pytest.mark.parametrize("k", KS)
def test_extract_canonical(eng, k):
    from kmer_denovo_filter_b200 import engine
    seqs = _rand_seqs(k)
    hs = engine.pack_sequences(seqs)
    ds = eng.upload(hs)
    lo, hi, okw = eng.extract_canonical(ds, k)
    codes, valid, _s, _l = kmers.encode_stream(seqs)
    ohi, olo, ook = kmers.canonical_windows(codes, valid, k)
    n = ook.shape[0]
    okw = okw.cpu().numpy().view(np.uint32)
    idx = np.arange(hs.n_bases)
    ok = ((okw[idx >> 5] >> (31 - (idx & 31)).astype(np.uint32)) & 1).astype(bool)
    assert not ok[n:].any()
    assert np.array_equal(ok[:n], ook)
    glo = lo.cpu().numpy().view(np.uint64)
    assert np.array_equal(glo[:n][ook], olo[ook])
    if hi is not None:
        ghi = hi.cpu().numpy().view(np.uint64)
        assert np.array_equal(ghi[:n][ook], ohi[ook])


@pytest.mark.parametrize("k", KS)
def test_count_stream_equals_oracle(eng, k):
    from kmer_denovo_filter_b200 import engine
    _g, reads = _genome_reads(k)
    want = kmers.count_sequences(reads, k)
    hs = engine.pack_sequences(reads)
    ds = eng.upload(hs)
    t = eng.new_table(k, n_keys=len(want))
    st = eng.new_stats()
    eng.count_stream(t, ds, engine.MODE_INSERT_COUNT, 0, 1, st)
    s = eng.read_stats(st)
    assert s["full"] == 0
    assert s["windows"] == sum(want.values())
    assert s["new"] == len(want)
    assert s["hits"] == s["windows"] - s["new"]
    got = _table_dict(eng, t)
    assert {key: v[0] for key, v in got.items()} == want
    assert all(v[1] == 0 for v in got.values())


def test_count_is_deterministic_and_high_load(eng):
    """Two runs give the same multiset; load factor 0.95 still exact."""
    from kmer_denovo_filter_b200 import engine
    k = 31
    _g, reads = _genome_reads(7, glen=8000, n=800)
    want = kmers.count_sequences(reads, k)
    ds = eng.upload(engine.pack_sequences(reads))
    cap = int(len(want) / 0.95) + 1
    for _ in range(2):
        t = eng.new_table(k, capacity=cap)
        st = eng.new_stats()
        eng.count_stream(t, ds, engine.MODE_INSERT_COUNT, 0, 1, st)
        assert eng.read_stats(st)["full"] == 0
        assert {key: v[0] for key, v in _table_dict(eng, t).items()} == want


def test_table_full_is_reported(eng):
    from kmer_denovo_filter_b200 import engine
    k = 21
    _g, reads = _genome_reads(3, glen=5000, n=300)
    want = kmers.count_sequences(reads, k)
    ds = eng.upload(engine.pack_sequences(reads))
    t = eng.new_table(k, capacity=len(want) // 2)
    st = eng.new_stats()
    eng.count_stream(t, ds, engine.MODE_INSERT_COUNT, 0, 1, st)
    assert eng.read_stats(st)["full"] == 1
    with pytest.raises(engine.KdfError):
        eng.check_not_full(st)


@pytest.mark.parametrize("k", [31, 47])
def test_filtered_count_mark_lookup_threshold(eng, k):
    """count --if (plane 1), reference mark, lookup and threshold compaction."""
    from kmer_denovo_filter_b200 import engine
    g, child = _genome_reads(11 + k, glen=12000, n=900)
    _g2, parent = _genome_reads(12 + k, glen=12000, n=900)
    parent = parent[:300] + child[:200]
    cc = kmers.count_sequences(child, k)
    pc = kmers.count_sequences(parent, k)
    rc = kmers.count_sequences([g[:6000]], k)
    ds_c = eng.upload(engine.pack_sequences(child))
    ds_p = eng.upload(engine.pack_sequences(parent))
    ds_r = eng.upload(engine.pack_sequences([g[:6000]]))
    t = eng.new_table(k, n_keys=len(cc))
    eng.count_stream(t, ds_c, engine.MODE_INSERT_COUNT, 0, 1)
    st = eng.new_stats()
    eng.count_stream(t, ds_p, engine.MODE_COUNT_IF_PRESENT, 1, 1, st)
    s = eng.read_stats(st)
    assert s["windows"] == sum(pc.values()) and s["new"] == 0
    assert s["hits"] == sum(c for key, c in pc.items() if key in cc)
    got = _table_dict(eng, t)
    assert set(got) == set(cc)
    for key, (a, b) in got.items():
        assert a == cc[key] and b == pc.get(key, 0)
    # threshold: child count >= 3 and parent count <= 0
    n, lo, hi, _a, _b = eng.threshold_compact(t, min0=3, max1=0)
    want = {key for key in cc if cc[key] >= 3 and pc.get(key, 0) == 0}
    assert n == len(want) and set(eng.keys_to_pyints(lo, hi)) == want
    assert eng.threshold_count(t, min0=3) == sum(1 for c in cc.values() if c >= 3)
    # mark-if-present with the reference, on a cleared plane 1
    eng.clear_plane(t, 1)
    eng.count_stream(t, ds_r, engine.MODE_MARK_IF_PRESENT, 1, 1)
    got = _table_dict(eng, t)
    for key, (a, b) in got.items():
        assert a == cc[key] and b == (1 if key in rc else 0)
    # lookup: present and absent keys
    probe = list(cc)[:500] + [key for key in pc if key not in cc][:500]
    lo, hi = eng.keys_to_device(probe, t.key_words)
    found, p0, _p1 = eng.lookup_keys(t, lo, hi)
    found = found.cpu().numpy().astype(bool).tolist()
    p0 = p0.cpu().numpy().view(np.uint32).tolist()
    for key, f, c in zip(probe, found, p0):
        assert f == (key in cc) and c == cc.get(key, 0)


@pytest.mark.parametrize("k", [31, 63])
def test_update_keys_insert_only_then_count_if(eng, k):
    from kmer_denovo_filter_b200 import engine
    _g, reads = _genome_reads(21 + k, glen=9000, n=700)
    full = kmers.count_sequences(reads, k)
    filt = sorted(full)[::3]
    t = eng.new_table(k, n_keys=len(filt))
    lo, hi = eng.keys_to_device(filt + filt[:10], t.key_words)  # duplicates are harmless
    st = eng.new_stats()
    eng.update_keys(t, lo, hi, engine.MODE_INSERT_ONLY, 0, 0, st)
    assert eng.read_stats(st)["new"] == len(filt)
    ds = eng.upload(engine.pack_sequences(reads))
    eng.count_stream(t, ds, engine.MODE_COUNT_IF_PRESENT, 0, 1)
    got = _table_dict(eng, t)
    assert {key: v[0] for key, v in got.items()} == {key: full[key] for key in filt}


@pytest.mark.parametrize("k", [5, 31, 47])
@pytest.mark.parametrize("min_distinct", [1, 3])
def test_scan_reads(eng, k, min_distinct):
    from kmer_denovo_filter_b200 import engine
    from oracle import discovery
    _g, reads = _genome_reads(31 + k, glen=6000, n=600, rl=120)
    reads += ["", "ACG", "N" * 50]
    full = kmers.count_sequences(reads, k)
    pu = set(sorted(full)[::7])
    t = eng.new_table(k, n_keys=len(pu))
    lo, hi = eng.keys_to_device(sorted(pu), t.key_words)
    eng.update_keys(t, lo, hi, engine.MODE_INSERT_ONLY, 0, 0)
    hs = engine.pack_sequences(reads)
    ds = eng.upload(hs)
    res = eng.scan_reads(t, ds, min_distinct=min_distinct, hit_cap=16)  # forces the retry path
    nd = res["ndistinct"].cpu().numpy().view(np.uint32).tolist()
    nh = res["nhits"].cpu().numpy().view(np.uint32).tolist()
    # slot -> key map to decode hit slots
    slot_keys = {}
    hp = res["hit_pos"].cpu().numpy().view(np.uint64)
    hsl = res["hit_slot"].cpu().numpy().view(np.uint32)
    starts = hs.read_starts
    want_hits = {}
    for r, seq in enumerate(reads):
        uniq, idx = discovery.scan_read_numeric(seq, k, pu)
        assert nd[r] == len(uniq), r
        assert nh[r] == len(idx), r
        if len(uniq) >= max(1, min_distinct):
            for i in idx:
                want_hits[int(starts[r]) + i] = None
    assert sorted(hp.tolist()) == sorted(want_hits)
    # every emitted slot decodes to the canonical key at that position
    klo, khi, kok = engine.debug_extract_host(hs, k)
    found, _a, _b = eng.lookup_keys(t, lo, hi, want_planes=False)
    assert bool(found.all())
    n, tlo, thi, _p0, _p1 = eng.threshold_compact(t)
    assert n == len(pu)


def test_scan_reads_overflow_path(eng):
    """A read with more than 1024 hit windows reports the overflow sentinel and
    still emits every hit."""
    from kmer_denovo_filter_b200 import engine
    k = 5
    rng = random.Random(5)
    long_read = "".join(rng.choice("ACGT") for _ in range(3000))
    pu = set(kmers.count_sequences([long_read], k))
    t = eng.new_table(k, n_keys=len(pu))
    lo, hi = eng.keys_to_device(sorted(pu), t.key_words)
    eng.update_keys(t, lo, hi, engine.MODE_INSERT_ONLY, 0, 0)
    hs = engine.pack_sequences([long_read, "ACGTACGTAC"])
    res = eng.scan_reads(t, eng.upload(hs), min_distinct=1)
    nd = res["ndistinct"].cpu().numpy().view(np.uint32).tolist()
    nh = res["nhits"].cpu().numpy().view(np.uint32).tolist()
    assert nd[0] == engine.NDISTINCT_OVERFLOW and nh[0] == 3000 - k + 1
    assert res["n_hits"] == nh[0] + nh[1]
    slots = res["hit_slot"].cpu().numpy()
    pos = res["hit_pos"].cpu().numpy().view(np.uint64)
    first = slots[pos < 3000]
    assert len(set(first.tolist())) == len(pu)


@pytest.mark.parametrize("k", [31, 47])
@pytest.mark.parametrize("n_parts,by_owner", [(1, True), (2, True), (3, True), (8, True),
                                              (1, False), (16, False), (32, False), (128, False),
                                              (256, False), (512, False)])
def test_bin_stream(eng, k, n_parts, by_owner):
    """K2p / K6: every valid canonical k-mer lands in exactly one bin, the bin is
    the one the host-side hash predicts, equal keys share a bin."""
    from kmer_denovo_filter_b200 import engine
    _g, reads = _genome_reads(41 + k, glen=5000, n=500)
    hs = engine.pack_sequences(reads)
    ds = eng.upload(hs)
    codes, valid, _s, _l = kmers.encode_stream(reads)
    ohi, olo, ook = kmers.canonical_windows(codes, valid, k)
    want = sorted(kmers.to_pyints(ohi[ook], olo[ook]))
    bins = eng.new_bins(k, n_parts, bin_cap=len(want) // n_parts * 4 + 512, by_owner=by_owner)
    st = eng.new_stats()
    eng.bin_stream(bins, ds, st)
    assert not bins.overflowed()
    counts = bins.counts()
    assert int(counts.sum()) == len(want) == eng.read_stats(st)["windows"]
    got = []
    kw = 1 if k <= 32 else 2
    log2p = n_parts.bit_length() - 1
    for b in range(n_parts):
        lo, hi = bins.bin_keys(b)
        keys = eng.keys_to_pyints(lo, hi)
        got.extend(keys)
        if keys:
            lo_np = np.array([x & 0xFFFFFFFFFFFFFFFF for x in keys], dtype=np.uint64)
            hi_np = np.array([x >> 64 for x in keys], dtype=np.uint64)
            part, _bk, owner = engine.debug_hash_host(lo_np, hi_np if kw == 2 else None, kw,
                                                      0 if by_owner else log2p, 1024,
                                                      n_parts if by_owner else 1)
            assert ((owner if by_owner else part) == b).all()
    assert sorted(got) == want
    if n_parts in (2, 3, 8, 16):
        assert counts.min() > 0 and counts.max() < 2.0 * counts.mean()
    # appending the same stream again doubles every bin (cursors accumulate)
    eng.bin_stream(bins, ds)
    assert (bins.counts() == 2 * counts).all()
    assert not bins.overflowed()


def test_bin_overflow_is_reported(eng):
    from kmer_denovo_filter_b200 import engine
    _g, reads = _genome_reads(77, glen=5000, n=500)
    ds = eng.upload(engine.pack_sequences(reads))
    bins = eng.new_bins(31, 4, bin_cap=100)
    eng.bin_stream(bins, ds)
    assert bins.overflowed()
    assert bins.counts().sum() == kmers.canonical_windows(*kmers.encode_stream(reads)[:2], 31)[2].sum()


@pytest.mark.parametrize("k", [31, 47])
@pytest.mark.parametrize("threads,wpr", [(256, 4), (512, 16), (1024, 8), (1024, 32)])
def test_bin_stream_launch_shapes(eng, k, threads, wpr, monkeypatch):
    """Every launch shape of the binning kernel (round size, flat / per-bin copy-out,
    queues that overflow into the direct path) bins the same multiset of keys."""
    from kmer_denovo_filter_b200 import engine
    monkeypatch.setenv("KDF_BIN_THREADS", str(threads))
    monkeypatch.setenv("KDF_BIN_WPR", str(wpr))
    _g, reads = _genome_reads(43 + k, glen=30000, n=6000)
    reads = reads + ["A" * 150] * 300          # one key 36 000 times: its queue overflows
    ds = eng.upload(engine.pack_sequences(reads))
    codes, valid, _s, _l = kmers.encode_stream(reads)
    ohi, olo, ook = kmers.canonical_windows(codes, valid, k)
    want = sorted(kmers.to_pyints(ohi[ook], olo[ook]))
    for n_parts in (4, 64, 512):
        bins = eng.new_bins(k, n_parts, bin_cap=len(want))
        st = eng.new_stats()
        eng.bin_stream(bins, ds, st)
        assert not bins.overflowed()
        assert int(bins.counts().sum()) == len(want) == eng.read_stats(st)["windows"]
        got = []
        for b in range(n_parts):
            got.extend(eng.keys_to_pyints(*bins.bin_keys(b)))
        assert sorted(got) == want


@pytest.mark.parametrize("k", [31, 47])
@pytest.mark.parametrize("n_passes,n_parts", [(2, 8), (8, 64), (16, 1)])
def test_bin_stream_passes_partition_the_keys(eng, k, n_passes, n_parts):
    """Multi-pass binning (kdf_bin_stream_pass): pass p bins exactly the keys whose top
    hash bits are p, into the bin the next bits name; over all passes every key appears
    once; kdf_bin_keys_pass agrees; kdf_count_bins_pass over the passes equals the oracle."""
    from kmer_denovo_filter_b200 import engine
    g, reads = _genome_reads(61 + k, glen=6000, n=900)
    ds = eng.upload(engine.pack_sequences(reads))
    dr = eng.upload(engine.pack_sequences([g[:3000]]))
    want = kmers.count_sequences(reads, k)
    refk = set(kmers.count_sequences([g[:3000]], k))
    kw = 1 if k <= 32 else 2
    plog = n_passes.bit_length() - 1
    log2p = n_parts.bit_length() - 1
    n_win = sum(want.values())
    lo_all, hi_all, ok = eng.extract_canonical(ds, k)
    okb = np.unpackbits(ok.cpu().numpy().view(np.uint32).byteswap().view(np.uint8))[:ds.n_bases].astype(bool)
    import torch
    sel = torch.from_numpy(np.flatnonzero(okb)).to(eng.device)
    klo = lo_all[sel].contiguous()
    khi = hi_all[sel].contiguous() if hi_all is not None else None
    seen = []
    emitted, n_count, distinct = set(), 0, 0
    for p in range(n_passes):
        bins = eng.new_bins(k, n_parts, bin_cap=n_win + 64)
        rb = eng.new_bins(k, n_parts, bin_cap=4000)
        st = eng.new_stats()
        eng.bin_stream(bins, ds, st, pass_=(plog, p))
        eng.bin_stream(rb, dr, None, pass_=(plog, p))
        b2 = eng.new_bins(k, n_parts, bin_cap=n_win + 64)
        eng.bin_keys(b2, klo, khi, pass_=(plog, p))
        assert (bins.counts() == b2.counts()).all()
        assert int(bins.counts().sum()) == eng.read_stats(st)["windows"]
        for b in range(n_parts):
            keys = eng.keys_to_pyints(*bins.bin_keys(b))
            assert sorted(keys) == sorted(eng.keys_to_pyints(*b2.bin_keys(b)))
            seen.extend(keys)
            if keys:
                lo_np = np.array([x & 0xFFFFFFFFFFFFFFFF for x in keys], dtype=np.uint64)
                hi_np = np.array([x >> 64 for x in keys], dtype=np.uint64)
                part, _bk, _o = engine.debug_hash_host(lo_np, hi_np if kw == 2 else None, kw,
                                                       plog + log2p, 1024, 1)
                assert (part == p * n_parts + b).all()
        res = eng.count_bins(bins, rb, slice_capacity=2 * len(want) // (n_passes * n_parts) + 256,
                             min0=3, max1=0, count_min0=3, out_cap=len(want) + 10, pass_=(plog, p))
        assert res["full"] == 0
        got = eng.keys_to_pyints(res["lo"], res["hi"])
        assert not (set(got) & emitted)
        emitted |= set(got)
        n_count += res["n_count"]
        distinct += res["distinct"]
    assert sorted(seen) == sorted(x for x, c in want.items() for _ in range(c))
    assert emitted == {x for x, c in want.items() if c >= 3} - refk
    assert n_count == sum(1 for c in want.values() if c >= 3) and distinct == len(want)


@pytest.mark.parametrize("k", [31, 47])
@pytest.mark.parametrize("owners,n_local,n_passes", [(2, 4, 1), (4, 16, 2), (8, 64, 4)])
def test_bin_stream_to_composite_bins(eng, k, owners, n_local, n_passes):
    """The fused multi-GPU route's binning (one pointer per (owner, hash range) bin,
    kdf_bin_stream_pass with bin_ptrs) on local memory: bin = owner * n_local + range,
    restricted to one pass group."""
    import torch
    from kmer_denovo_filter_b200 import engine
    _g, reads = _genome_reads(71 + k, glen=6000, n=900)
    ds = eng.upload(engine.pack_sequences(reads))
    codes, valid, _s, _l = kmers.encode_stream(reads)
    ohi, olo, ook = kmers.canonical_windows(codes, valid, k)
    want = sorted(kmers.to_pyints(ohi[ook], olo[ook]))
    kw = 1 if k <= 32 else 2
    plog = n_passes.bit_length() - 1
    log2l = n_local.bit_length() - 1
    n_bins = owners * n_local
    cap = len(want) + 8
    got_all = []
    for p in range(n_passes):
        buf = torch.zeros(n_bins * cap * kw, dtype=torch.int64, device=eng.device)
        ptrs = torch.tensor([buf.data_ptr() + b * cap * kw * 8 for b in range(n_bins)],
                            dtype=torch.int64, device=eng.device)
        cursors = torch.zeros(n_bins, dtype=torch.int64, device=eng.device)
        overflow = torch.zeros(1, dtype=torch.int64, device=eng.device)
        st = eng.new_stats()
        eng.bin_stream_to(ds, k, ptrs, cap, cursors, overflow, by_owner=owners, stats=st,
                          pass_=(plog, p) if n_passes > 1 else None)
        assert int(overflow.item()) == 0
        cnt = cursors.cpu().numpy()
        assert int(cnt.sum()) == eng.read_stats(st)["windows"]
        host = buf.cpu().numpy().view(np.uint64)
        for b in range(n_bins):
            seg = host[b * cap * kw:(b * cap + int(cnt[b])) * kw]
            lo_np = seg if kw == 1 else seg[0::2]
            hi_np = np.zeros_like(lo_np) if kw == 1 else seg[1::2]
            got_all.extend(kmers.to_pyints(hi_np, lo_np))
            if lo_np.shape[0]:
                part, _bk, owner = engine.debug_hash_host(lo_np.copy(), hi_np.copy() if kw == 2 else None,
                                                          kw, plog + log2l, 1024, owners)
                assert (owner == b // n_local).all()
                assert (part == p * n_local + b % n_local).all()
    assert sorted(got_all) == want


@pytest.mark.parametrize("k", [31, 47])
def test_bin_keys_matches_bin_stream(eng, k):
    from kmer_denovo_filter_b200 import engine
    _g, reads = _genome_reads(51 + k, glen=4000, n=300)
    ds = eng.upload(engine.pack_sequences(reads))
    lo, hi, ok = eng.extract_canonical(ds, k)
    okb = np.unpackbits(ok.cpu().numpy().view(np.uint32).byteswap().view(np.uint8))[:ds.n_bases].astype(bool)
    import torch
    sel = torch.from_numpy(np.flatnonzero(okb)).to(eng.device)
    klo = lo[sel].contiguous()
    khi = hi[sel].contiguous() if hi is not None else None
    a = eng.new_bins(k, 8, bin_cap=int(sel.shape[0]))
    b = eng.new_bins(k, 8, bin_cap=int(sel.shape[0]))
    eng.bin_stream(a, ds)
    eng.bin_keys(b, klo, khi)
    assert (a.counts() == b.counts()).all()
    for p in range(8):
        assert sorted(eng.keys_to_pyints(*a.bin_keys(p))) == sorted(eng.keys_to_pyints(*b.bin_keys(p)))


@pytest.mark.parametrize("k", [21, 31, 47, 63])
@pytest.mark.parametrize("n_parts", [1, 8, 64])
def test_count_bins_equals_direct_table(eng, k, n_parts):
    """The L2-sliced partitioned count (bins -> slice -> emit) gives the same
    (key, count, in-reference) triples as the oracle."""
    from kmer_denovo_filter_b200 import engine
    g, child = _genome_reads(91 + k, glen=9000, n=1200)
    ref = [g[:6000]]
    want = kmers.count_sequences(child, k)
    refk = kmers.count_sequences(ref, k)
    dc = eng.upload(engine.pack_sequences(child))
    dr = eng.upload(engine.pack_sequences(ref))
    n_win = sum(want.values())
    cb = eng.new_bins(k, n_parts, bin_cap=n_win // n_parts * 2 + 512)
    rb = eng.new_bins(k, n_parts, bin_cap=6000 // n_parts * 2 + 512)
    eng.bin_stream(cb, dc)
    eng.bin_stream(rb, dr)
    assert not cb.overflowed() and not rb.overflowed()
    res = eng.count_bins(cb, rb, slice_capacity=2 * len(want) // n_parts + 64, want_planes=True,
                         count_min0=3, out_cap=len(want) + 10)
    assert res["full"] == 0
    assert res["keys"] == n_win and res["distinct"] == len(want) == res["occupied"] == res["n_out"]
    assert res["hits"] + res["distinct"] == n_win
    assert res["n_count"] == sum(1 for c in want.values() if c >= 3)
    keys = eng.keys_to_pyints(res["lo"], res["hi"])
    p0 = res["p0"].cpu().numpy().view(np.uint32).tolist()
    p1 = res["p1"].cpu().numpy().view(np.uint32).tolist()
    assert dict(zip(keys, p0)) == want
    assert {key for key, f in zip(keys, p1) if f} == set(want) & set(refk)
    # the same bins counted in sub-range passes (a bin spanning several table slices)
    if n_parts == 8:
        for sub in (2, 8):
            rs = eng.count_bins(cb, rb, slice_capacity=2 * len(want) // (n_parts * sub) + 64,
                                want_planes=True, count_min0=3, out_cap=len(want) + 10, sub_split=sub)
            assert rs["full"] == 0 and rs["keys"] == n_win and rs["distinct"] == len(want)
            ks = eng.keys_to_pyints(rs["lo"], rs["hi"])
            assert dict(zip(ks, rs["p0"].cpu().numpy().view(np.uint32).tolist())) == want
            assert {key for key, f in zip(ks, rs["p1"].cpu().numpy().view(np.uint32).tolist()) if f} == \
                set(want) & set(refk)
    # thresholded emit: count >= 3 and not in the reference; undersized output is reported
    res2 = eng.count_bins(cb, rb, slice_capacity=2 * len(want) // n_parts + 64, min0=3, max1=0,
                          out_cap=len(want) + 10)
    want2 = {key for key, c in want.items() if c >= 3 and key not in refk}
    assert set(eng.keys_to_pyints(res2["lo"], res2["hi"])) == want2 and res2["n_out"] == len(want2)
    res3 = eng.count_bins(cb, rb, slice_capacity=2 * len(want) // n_parts + 64, min0=3, max1=0,
                          out_cap=5)
    assert res3["n_out"] == len(want2) and res3["lo"].shape[0] == 5
    # a slice that cannot hold the bin's distinct keys says so
    res4 = eng.count_bins(cb, None, slice_capacity=max(4, len(want) // n_parts // 2), out_cap=4)
    assert res4["full"] == 1


@pytest.mark.parametrize("k", [21, 31, 47, 63])
@pytest.mark.parametrize("n_parts", [1, 8])
def test_count_bins_packed_form(eng, k, n_parts):
    """The packed form of kdf_count_bins (keys-only slice, saturating counter in the
    key's spare bits, include/kdf.h) answers threshold + reference subtraction exactly
    like the oracle, for every threshold that fits the spare bits, with and without a
    reference, in sub-range passes, and reports a full slice."""
    from kmer_denovo_filter_b200 import engine
    g, child = _genome_reads(191 + k, glen=9000, n=1500)
    ref = [g[:6000]]
    want = kmers.count_sequences(child, k)
    refk = set(kmers.count_sequences(ref, k))
    dc = eng.upload(engine.pack_sequences(child))
    dr = eng.upload(engine.pack_sequences(ref))
    n_win = sum(want.values())
    cb = eng.new_bins(k, n_parts, bin_cap=n_win // n_parts * 2 + 512)
    rb = eng.new_bins(k, n_parts, bin_cap=6000 // n_parts * 2 + 512)
    eng.bin_stream(cb, dc)
    eng.bin_stream(rb, dr)
    cap = 2 * len(want) // n_parts + 64
    free_bits = 64 * cb.key_words - 2 * k
    for m in (1, 2, 3, 5):
        packed = bool(eng.lib.kdf_count_bins_packed(k, m, engine.U32_MAX, 0, 0, m, 0))
        assert packed == (m <= (1 << min(free_bits, 31)) - 1)
        assert eng.count_bins_packed(k, m) == packed
        # count >= m and not in the reference (the discovery chain's call)
        res = eng.count_bins(cb, rb, slice_capacity=cap, min0=m, max1=0, count_min0=m,
                             out_cap=len(want) + 10)
        cand = {key for key, c in want.items() if c >= m}
        assert res["full"] == 0 and res["keys"] == n_win
        assert res["distinct"] == len(want) == res["occupied"]
        assert res["hits"] + res["distinct"] == n_win
        assert res["n_count"] == len(cand)
        got = eng.keys_to_pyints(res["lo"], res["hi"])
        assert len(got) == res["n_out"] == len(set(got))
        assert set(got) == cand - refk
        # no reference bins: every candidate comes out
        res = eng.count_bins(cb, None, slice_capacity=cap, min0=m, max1=0, count_min0=m,
                             out_cap=len(want) + 10)
        assert set(eng.keys_to_pyints(res["lo"], res["hi"])) == cand and res["n_count"] == len(cand)
        # reference marks ignored (max1 open), n_count over all occupied slots (count_min0 = 0)
        res = eng.count_bins(cb, rb, slice_capacity=cap, min0=m, out_cap=len(want) + 10)
        assert set(eng.keys_to_pyints(res["lo"], res["hi"])) == cand and res["n_count"] == len(want)
    # sub-range passes over the same bins, and an undersized output / slice
    for sub in (2, 8):
        rs = eng.count_bins(cb, rb, slice_capacity=2 * len(want) // (n_parts * sub) + 64, min0=3,
                            max1=0, count_min0=3, out_cap=len(want) + 10, sub_split=sub)
        assert rs["full"] == 0 and rs["keys"] == n_win and rs["distinct"] == len(want)
        assert set(eng.keys_to_pyints(rs["lo"], rs["hi"])) == \
            {key for key, c in want.items() if c >= 3} - refk
    small = eng.count_bins(cb, rb, slice_capacity=cap, min0=3, max1=0, count_min0=3, out_cap=5)
    assert small["n_out"] == len({key for key, c in want.items() if c >= 3} - refk)
    assert small["lo"].shape[0] == 5
    full = eng.count_bins(cb, None, slice_capacity=max(4, len(want) // n_parts // 2), min0=3, max1=0,
                          count_min0=3, out_cap=4)
    assert full["full"] == 1


@pytest.mark.parametrize("k", [31, 47])
def test_update_bins_equals_count_stream(eng, k, monkeypatch):
    """`count --if` applied bin after bin (kdf_update_bins, the route for filter tables
    larger than L2) gives the same per-key counts as the direct stream probe and the
    oracle; the chain helper takes that route in chunks once the table is 'large'."""
    from kmer_denovo_filter_b200 import engine
    from kmer_denovo_filter_b200.discovery import kmer_chain
    g, parent = _genome_reads(301 + k, glen=8000, n=1300)
    parent = parent + ["", "ACGTN", g[:k]]
    filt = sorted(kmers.count_sequences([g[1000:5000]], k))[::2]
    want = kmers.count_sequences(parent, k)
    ds = eng.upload(engine.pack_sequences(parent))
    n_win = sum(want.values())

    def fresh():
        t = eng.new_table(k, n_keys=len(filt))
        lo, hi = eng.keys_to_device(filt, t.key_words)
        eng.update_keys(t, lo, hi, engine.MODE_INSERT_ONLY, 0, 0)
        return t, lo, hi

    t0, lo, hi = fresh()
    eng.count_stream(t0, ds, engine.MODE_COUNT_IF_PRESENT, 0, 1, None)
    direct = eng.lookup_keys(t0, lo, hi)[1].cpu().numpy().view(np.uint32)
    assert direct.tolist() == [want.get(x, 0) for x in filt]
    for n_parts in (1, 16):
        t1, lo, hi = fresh()
        bins = eng.new_bins(k, n_parts, bin_cap=n_win + 64)
        eng.bin_stream(bins, ds)
        st = eng.new_stats()
        eng.update_bins(t1, bins, engine.MODE_COUNT_IF_PRESENT, 0, 1, st)
        got = eng.lookup_keys(t1, lo, hi)[1].cpu().numpy().view(np.uint32)
        assert np.array_equal(got, direct)
        assert eng.read_stats(st)["windows"] == n_win
    # the chain helper: force the binned route, in several chunks, with a bin retry
    monkeypatch.setattr(kmer_chain, "PROBE_DIRECT_BYTES", 0)
    monkeypatch.setattr(kmer_chain, "PROBE_SLICE_BYTES", 4096)
    monkeypatch.setattr(kmer_chain, "PROBE_CHUNK_BASES", 32 * 1024)
    monkeypatch.setattr(kmer_chain, "_bin_capacity", lambda n, p: 64)
    t2, lo, hi = fresh()
    st = eng.new_stats()
    kmer_chain.count_if_present(eng, t2, ds, st)
    got = eng.lookup_keys(t2, lo, hi)[1].cpu().numpy().view(np.uint32)
    assert np.array_equal(got, direct)
    assert eng.read_stats(st)["windows"] == n_win


def test_hit_coverage_device_equals_oracle(eng):
    """K7 on the device (expand + radix sort + run-length encode) against the ORACLE's
    per-read restatement of core/bam_scanner.py:97-117 merged with Counters
    (discovery/pipeline.py:851-855) — an implementation that shares no code with the
    kernel — on random CIGARs with clips, insertions, deletions, skips and padding."""
    import collections
    from oracle import discovery as odisc
    from test_postprocess_cpu import _random_alignments

    from oracle import bam as obam

    class Rec:          # the oracle's own get_aligned_pairs (oracle/bam.py) over (pos, cigar)
        get_aligned_pairs = obam.BamRecord.get_aligned_pairs

        def __init__(self, pos, cigar):
            self.pos, self.cigar = pos, cigar

    for seed, k in ((11, 9), (12, 31), (13, 63)):
        contig, start, cig_off, cigar, hr, ho, recs = _random_alignments(seed, n_reads=400, k=k)
        kc = collections.defaultdict(collections.Counter)
        rc = collections.defaultdict(collections.Counter)
        for c, st, ops, offs in recs:
            cov = odisc.kmer_ref_positions(Rec(st, ops), offs, k)
            kc[c].update(cov)
            for p in cov:
                rc[c][p] += 1
        gc, gp, gk, gr = eng.hit_coverage(np.asarray(hr, np.uint32), np.asarray(ho, np.uint32), k,
                                          np.asarray(contig, np.int32), np.asarray(start, np.int64),
                                          np.asarray(cig_off, np.uint64), np.asarray(cigar, np.uint32))
        assert gc.shape[0] > 100
        got_k = collections.defaultdict(dict)
        got_r = collections.defaultdict(dict)
        for c, p, a, b in zip(gc.tolist(), gp.tolist(), gk.tolist(), gr.tolist()):
            got_k[c][p] = a
            got_r[c][p] = b
        assert {c: dict(v) for c, v in kc.items() if v} == dict(got_k)
        assert {c: dict(v) for c, v in rc.items() if v} == dict(got_r)
        # sorted by (contig, position)
        order = gc.astype(np.int64) * (1 << 40) + gp.astype(np.int64)
        assert (np.diff(order) > 0).all()
    empty = eng.hit_coverage(np.zeros(0, np.uint32), np.zeros(0, np.uint32), 31, np.zeros(1, np.int32),
                             np.zeros(1, np.int64), np.zeros(2, np.uint64), np.zeros(0, np.uint32))
    assert all(a.shape[0] == 0 for a in empty)


@pytest.mark.parametrize("k", [31, 47])
def test_table_filter_keeps_results_exact(eng, k):
    """A table behind its two-bit filter (kdf_table_build_filter) gives the same counts,
    marks, hit list and stats as the unfiltered table and the oracle; an insert after
    the build detaches the filter instead of producing false negatives."""
    from kmer_denovo_filter_b200 import engine
    g, parent = _genome_reads(501 + k, glen=9000, n=1500)
    parent = parent + ["", "ACGTN", g[:k]]
    filt = sorted(kmers.count_sequences([g[1000:6000]], k))[::2]
    want = kmers.count_sequences(parent, k)
    ds = eng.upload(engine.pack_sequences(parent))
    n_win = sum(want.values())
    res = []
    for use_filter in (False, True, "tiny"):
        t = eng.new_table(k, capacity=1 << 16)       # not shared-memory resident
        lo, hi = eng.keys_to_device(filt, t.key_words)
        eng.update_keys(t, lo, hi, engine.MODE_INSERT_ONLY, 0, 0)
        if use_filter:
            eng.build_filter(t, 1 if use_filter == "tiny" else len(filt))   # tiny: dense filter, many candidates
        st = eng.new_stats()
        eng.count_stream(t, ds, engine.MODE_COUNT_IF_PRESENT, 0, 1, st)
        eng.count_stream(t, ds, engine.MODE_MARK_IF_PRESENT, 1, 1, st)
        _f, p0, p1 = eng.lookup_keys(t, lo, hi)
        sp = eng.scan_reads_sparse(t, ds)
        res.append((p0.cpu().numpy().view(np.uint32).tolist(), p1.cpu().numpy().view(np.uint32).tolist(),
                    sp["read"].tolist(), sp["ndistinct"].tolist(), sp["nhits"].tolist(),
                    eng.read_stats(st)))
        t.close()
    assert res[0][0] == [want.get(x, 0) for x in filt]
    assert res[0][1] == [1 if x in want else 0 for x in filt]
    assert res[0][5]["windows"] == 2 * n_win and res[0][5]["hits"] == 2 * sum(want.get(x, 0) for x in filt)
    assert res[1] == res[0] and res[2] == res[0]
    # inserting after the build: the new key must be found
    t = eng.new_table(k, capacity=1 << 16)
    lo, hi = eng.keys_to_device(filt[:10], t.key_words)
    eng.update_keys(t, lo, hi, engine.MODE_INSERT_ONLY, 0, 0)
    eng.build_filter(t, 10)
    lo2, hi2 = eng.keys_to_device(filt[10:], t.key_words)
    eng.update_keys(t, lo2, hi2, engine.MODE_INSERT_ONLY, 0, 0)
    eng.count_stream(t, ds, engine.MODE_COUNT_IF_PRESENT, 0, 1, None)
    lo, hi = eng.keys_to_device(filt, t.key_words)
    assert eng.lookup_keys(t, lo, hi)[1].cpu().numpy().view(np.uint32).tolist() == res[0][0]


def test_sparse_validity_upload_rebuilds_the_bitmap(eng, monkeypatch):
    """Uploads send codes + the invalid-position list and rebuild the validity bitmap on
    the device (kdf_valid_from_invalid): same bitmap as the dense upload, for plain,
    copy-stream and chunked uploads, ragged ends and an empty stream."""
    from kmer_denovo_filter_b200 import engine
    import torch
    _g, reads = _genome_reads(701, glen=5000, n=400)
    reads = reads + ["", "N" * 40, "ACGTNNACGT", "A" * 33, "C"]
    hs = engine.pack_sequences(reads)
    assert hs.invalid is not None and hs.invalid.shape[0] >= len(reads) - 1
    monkeypatch.setenv("KDF_SPARSE_VALID", "0")
    dense = eng.upload(hs)
    monkeypatch.setenv("KDF_SPARSE_VALID", "1")
    side = torch.cuda.Stream(device=eng.device)
    a = eng.upload(hs)
    b = eng.upload(hs, copy_stream=side)
    c, ev = eng.upload_chunked(hs, side, n_chunks=5)
    torch.cuda.synchronize()
    want = dense.valid.cpu().numpy()
    assert np.array_equal(want.view(np.uint32), hs.valid)
    for d in (a, b, c):
        assert np.array_equal(d.valid.cpu().numpy(), want)
        assert np.array_equal(d.codes.cpu().numpy(), dense.codes.cpu().numpy())
    st = eng.new_stats()
    t = eng.new_table(31, n_keys=16)
    eng.count_stream(t, a, engine.MODE_COUNT_IF_PRESENT, 0, 1, st)
    assert eng.read_stats(st)["windows"] == sum(kmers.count_sequences(reads, 31).values())
    e = eng.upload(engine.pack_sequences([]))
    assert e.n_bases == 0


def test_count_bins_packed_heavy_duplicates(eng):
    """Many concurrent copies of few keys (the saturating CAS under contention) and
    keys whose top bases are all T (state bits next to an all-ones key prefix)."""
    from kmer_denovo_filter_b200 import engine
    k = 31
    seqs = ["T" * 15 + "C" + "A" * 15] * 3000 + ["ACGTTGCAAGGCTTAACCGGATATCGCGATTAGC"] * 5000 + \
           ["T" * 20 + "G" + "C" * 12] * 2 + ["G" * 31]
    want = kmers.count_sequences(seqs, k)
    dc = eng.upload(engine.pack_sequences(seqs))
    n_win = sum(want.values())
    cb = eng.new_bins(k, 4, bin_cap=n_win + 64)
    eng.bin_stream(cb, dc)
    for m in (1, 2, 3):
        res = eng.count_bins(cb, None, slice_capacity=1024, min0=m, max1=0, count_min0=m, out_cap=64)
        assert res["full"] == 0 and res["keys"] == n_win and res["distinct"] == len(want)
        assert set(eng.keys_to_pyints(res["lo"], res["hi"])) == {x for x, c in want.items() if c >= m}


@pytest.mark.parametrize("k", [31, 47])
@pytest.mark.parametrize("big_table", [False, True])
def test_scan_sparse_equals_dense(eng, k, big_table):
    """Sparse hit list + device reduction == dense per-read scan == oracle,
    for a shared-memory-resident table and for an L2/HBM one."""
    from kmer_denovo_filter_b200 import engine
    from oracle import discovery
    g, reads = _genome_reads(61 + k, glen=6000, n=700)
    reads = reads + ["", "ACGT", g[:k], g[100:100 + k + 1]]
    allk = sorted(kmers.count_sequences([g[2000:2300]], k))
    pu = set(allk[::3])
    ds = eng.upload(engine.pack_sequences(reads))
    t = eng.new_table(k, capacity=(1 << 16) if big_table else None, n_keys=len(pu))
    lo, hi = eng.keys_to_device(sorted(pu), t.key_words)
    eng.update_keys(t, lo, hi, engine.MODE_INSERT_ONLY, 0, 0)
    dense = eng.scan_reads(t, ds, min_distinct=1)
    nd = dense["ndistinct"].cpu().numpy().view(np.uint32)
    nh = dense["nhits"].cpu().numpy().view(np.uint32)
    st = eng.new_stats()
    sp = eng.scan_reads_sparse(t, ds, stats=st, hit_cap=64)   # forces the exact-size retry
    want_reads = np.flatnonzero(nh > 0)
    assert np.array_equal(sp["read"], want_reads.astype(np.uint64))
    assert np.array_equal(sp["ndistinct"], nd[want_reads])
    assert np.array_equal(sp["nhits"], nh[want_reads])
    assert eng.read_stats(st)["windows"] == sum(kmers.count_sequences(reads, k).values())
    starts = ds.read_starts.cpu().numpy().view(np.uint64)
    for j, r in enumerate(sp["read"].tolist()):
        uniq, idx = discovery.scan_read_numeric(reads[r], k, pu)
        assert sp["ndistinct"][j] == len(uniq) and sp["nhits"][j] == len(idx)
        f = int(sp["first"][j])
        got_idx = sp["hit_pos"][f:f + len(idx)] - starts[r]
        assert sorted(got_idx.tolist()) == sorted(idx)


@pytest.mark.parametrize("k", [31, 47])
def test_scan_sparse_long_reads_full_of_hits(eng, k):
    """Long reads (HiFi / ONT lengths) in which every window is a hit — a novel insertion:
    > 10 000 hits per read, with repeats so that distinct < hits.  The per-read reduction
    is sort based (linear in the run), so this finishes at once and stays exact."""
    from kmer_denovo_filter_b200 import engine
    rng = random.Random(7 + k)
    unit = "".join(rng.choice("ACGT") for _ in range(6000))
    reads = [unit * 3, unit[:3000] + unit[:3000] + unit, "ACGT" * 50, unit[100:400]]
    pu = sorted(kmers.count_sequences([unit], k))
    ds = eng.upload(engine.pack_sequences(reads))
    t = eng.new_table(k, n_keys=len(pu))
    lo, hi = eng.keys_to_device(pu, t.key_words)
    eng.update_keys(t, lo, hi, engine.MODE_INSERT_ONLY, 0, 0)
    sp = eng.scan_reads_sparse(t, ds)
    want = {}
    pus = set(pu)
    for r, seq in enumerate(reads):
        codes, valid, _s, _l = kmers.encode_stream([seq])
        whi, wlo, wok = kmers.canonical_windows(codes, valid, k)
        hits = [x for x in kmers.to_pyints(whi[wok], wlo[wok]) if x in pus]
        if hits:
            want[r] = (len(set(hits)), len(hits))
    assert sp["read"].tolist() == sorted(want)
    assert max(v[1] for v in want.values()) > 10000
    for j, r in enumerate(sp["read"].tolist()):
        assert (int(sp["ndistinct"][j]), int(sp["nhits"][j])) == want[r]
    # `first` indexes the position-sorted hit list
    starts = ds.read_starts.cpu().numpy().view(np.uint64)
    for j, r in enumerate(sp["read"].tolist()):
        f, n = int(sp["first"][j]), int(sp["nhits"][j])
        seg = sp["hit_pos"][f:f + n]
        assert seg.min() >= starts[r] and (np.diff(seg.astype(np.int64)) > 0).all()
        assert seg.max() < starts[r] + len(reads[r])


def test_empty_stream_is_a_noop(eng):
    from kmer_denovo_filter_b200 import engine
    hs = engine.pack_sequences([])
    ds = eng.upload(hs)
    t = eng.new_table(31, n_keys=10)
    st = eng.new_stats()
    eng.count_stream(t, ds, engine.MODE_INSERT_COUNT, 0, 1, st)
    assert eng.read_stats(st) == {"windows": 0, "full": 0, "hits": 0, "new": 0}
    assert eng.threshold_count(t) == 0
    res = eng.scan_reads(t, ds)
    assert res["n_hits"] == 0
