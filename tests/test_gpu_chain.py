"""The whole-sample k-mer chain (what bench.py times) against the oracle's C twin
on the same seeded synthetic trios, device-resident and from host buffers, and
size-independent properties at a larger size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from kmer_denovo_filter_b200 import engine
    return engine.CudaEngine()


def _trio(eng, genome_bp, depth, read_len=150):
    import torch
    from kmer_denovo_filter_b200 import synth
    return synth.make_trio(torch, eng.device, genome_bp, depth=depth, read_len=read_len,
                           n_denovo=20)


def _host(s):
    return (s["codes"].cpu().numpy().view(np.uint64), s["valid"].cpu().numpy().view(np.uint32),
            s["n_bases"], s["read_starts"].cpu().numpy().view(np.uint64),
            s["read_lens"].cpu().numpy().view(np.uint32))


def _dev(s):
    from kmer_denovo_filter_b200 import engine
    return engine.DeviceStream(s["codes"], s["valid"], s["n_bases"], s["read_starts"], s["read_lens"])


def _keys(lo, hi):
    return sorted((int(h) << 64) | int(l) for l, h in zip(lo.tolist(), hi.tolist()))


@pytest.mark.parametrize("k,genome,depth", [(31, 300_000, 12), (21, 200_000, 10),
                                            (47, 200_000, 10), (63, 200_000, 10)])
def test_chain_equals_oracle(eng, k, genome, depth):
    from kmer_denovo_filter_b200 import engine
    from kmer_denovo_filter_b200.discovery import kmer_chain
    from oracle import ckdf
    trio = _trio(eng, genome, depth)
    h = {w: _host(trio[w]) for w in ("child", "mother", "father", "ref")}
    want = ckdf.discovery_chain(h["child"], h["mother"][:3], h["father"][:3], h["ref"][:3], k,
                                threads=4)
    got = kmer_chain.discover_streams(eng, _dev(trio["child"]), _dev(trio["mother"]),
                                      _dev(trio["father"]), _dev(trio["ref"]), k)
    for key in ("candidates", "non_ref", "after_mother", "proband_unique", "units"):
        assert got[key] == want[key], key
    assert want["proband_unique"] > 0
    assert sorted(got["pu"].to_pyints()) == _keys(want["pu_lo"], want["pu_hi"])
    assert np.array_equal(got["ndistinct"], want["nd"])
    assert np.array_equal(got["nhits"], want["nh"])
    # same call with HOST buffers (the e2e path of bench.py)
    hs = {w: engine.HostStream(*h[w]) for w in h}
    got2 = kmer_chain.discover_streams(eng, hs["child"], hs["mother"], hs["father"], hs["ref"], k)
    assert got2["units"] == want["units"]
    assert sorted(got2["pu"].to_pyints()) == sorted(got["pu"].to_pyints())
    assert np.array_equal(got2["ndistinct"], want["nd"])


def test_chain_grows_an_undersized_child_table(eng):
    from kmer_denovo_filter_b200.discovery import kmer_chain
    trio = _trio(eng, 100_000, 6)
    a = kmer_chain.discover_streams(eng, _dev(trio["child"]), _dev(trio["mother"]),
                                    _dev(trio["father"]), _dev(trio["ref"]), 31)
    b = kmer_chain.discover_streams(eng, _dev(trio["child"]), _dev(trio["mother"]),
                                    _dev(trio["father"]), _dev(trio["ref"]), 31,
                                    child_capacity=4096)
    for key in ("candidates", "non_ref", "after_mother", "proband_unique", "units"):
        assert a[key] == b[key]


def test_chain_properties_at_scale(eng):
    """8 Mbp x 30x (bench-like): properties that do not need the oracle.
    - total of the child count plane == number of valid child windows
    - every proband-unique k-mer is absent from both parents and the reference
    - every de novo event yields informative reads."""
    import torch
    from kmer_denovo_filter_b200 import engine
    from kmer_denovo_filter_b200.discovery import kmer_chain
    k = 31
    trio = _trio(eng, 8_000_000, 30)
    child = _dev(trio["child"])
    t = eng.new_table(k, capacity=kmer_chain.default_child_capacity(eng, child.n_bases))
    st = eng.new_stats()
    eng.count_stream(t, child, engine.MODE_INSERT_COUNT, 0, 1, st)
    s = eng.read_stats(st)
    assert not s["full"]
    assert s["hits"] + s["new"] == s["windows"]
    n, lo, hi, p0, _p1 = eng.threshold_compact(t, want_planes=True)
    assert n == s["new"]
    assert int(p0.to(torch.int64).sum().item()) == s["windows"]
    t.close()
    res = kmer_chain.discover_streams(eng, child, _dev(trio["mother"]), _dev(trio["father"]),
                                      _dev(trio["ref"]), k)
    assert res["candidates"] >= res["non_ref"] >= res["after_mother"] >= res["proband_unique"] > 0
    pu = res["pu"]
    for who in ("mother", "father", "ref"):
        tp = pu.build_table()
        st = eng.new_stats()
        eng.count_stream(tp, _dev(trio[who]), engine.MODE_COUNT_IF_PRESENT, 0, 1, st)
        assert eng.read_stats(st)["hits"] == 0
        tp.close()
    assert res["informative_reads"] >= 20


def _pack_dev(eng, seqs):
    from kmer_denovo_filter_b200 import engine
    return eng.upload(engine.pack_sequences(seqs))


def test_partitioned_count_survives_skew_and_small_slices(eng):
    """One k-mer repeated 200 000 times lands in one hash range: the bins overflow
    and are re-made with exact sizes; a slice far too small is grown.  Counts stay
    exact (the reference's analogue is Jellyfish spilling to .jf_N files)."""
    import random
    from kmer_denovo_filter_b200.discovery import kmer_chain
    from oracle import kmers
    rng = random.Random(5)
    k = 31
    g = "".join(rng.choice("ACGT") for _ in range(3000))
    poly = "A" * 100
    reads = [poly] * 3000 + [g[i:i + 100] for i in range(0, 2900, 7)] * 3
    want = kmers.count_sequences(reads, k)
    ref = kmers.count_sequences([g[:1500]], k)
    res = kmer_chain.count_child_partitioned(eng, [_pack_dev(eng, reads)], [_pack_dev(eng, [g[:1500]])],
                                             k, 3, n_parts=16, slice_capacity=64)
    assert res["child_windows"] == sum(want.values())
    assert res["child_distinct"] == len(want)
    assert res["candidates"] == sum(1 for c in want.values() if c >= 3)
    got = set(eng.keys_to_pyints(res["lo"], res["hi"]))
    assert got == {x for x, c in want.items() if c >= 3 and x not in ref}
    assert 0 in want and want[0] == 3000 * 70          # poly-A canonical key


@pytest.mark.parametrize("k", [31, 47])
@pytest.mark.parametrize("n_passes", [2, 8])
def test_multi_pass_chain_equals_single_pass_and_oracle(eng, k, n_passes):
    """The child count in several hash-range passes (what bounds the bins' memory for a
    whole-genome sample) and with every sample given as a LIST of streams: identical
    stage sizes, proband-unique set and per-read records."""
    from kmer_denovo_filter_b200.discovery import kmer_chain
    from oracle import ckdf
    trio = _trio(eng, 250_000, 12)
    h = {w: _host(trio[w]) for w in ("child", "mother", "father", "ref")}
    want = ckdf.discovery_chain(h["child"], h["mother"][:3], h["father"][:3], h["ref"][:3], k, threads=4)
    d = {w: _dev(trio[w]) for w in ("child", "mother", "father", "ref")}
    got = kmer_chain.discover_streams(eng, d["child"], d["mother"], d["father"], d["ref"], k,
                                      n_passes=n_passes)
    assert got["n_passes"] == n_passes
    for key in ("candidates", "non_ref", "after_mother", "proband_unique", "units"):
        assert got[key] == want[key], key
    assert sorted(got["pu"].to_pyints()) == _keys(want["pu_lo"], want["pu_hi"])
    assert np.array_equal(got["ndistinct"], want["nd"]) and np.array_equal(got["nhits"], want["nh"])
    # every sample as two streams
    parts = {w: _split(d[w]) for w in d}
    got2 = kmer_chain.discover_streams(eng, parts["child"], parts["mother"], parts["father"],
                                       parts["ref"], k, n_passes=n_passes)
    for key in ("candidates", "non_ref", "after_mother", "proband_unique"):
        assert got2[key] == want[key], key
    assert sorted(got2["pu"].to_pyints()) == _keys(want["pu_lo"], want["pu_hi"])
    assert np.array_equal(got2["ndistinct"], want["nd"]) and np.array_equal(got2["nhits"], want["nh"])


def _split(ds):
    """One DeviceStream as two, cut at a read boundary that is word aligned."""
    from kmer_denovo_filter_b200 import engine
    starts = ds.read_starts.cpu().numpy().astype(np.int64)
    cand = np.flatnonzero(starts % 32 == 0)
    cand = cand[cand > 0]
    if cand.shape[0] == 0:
        return [ds]
    r = int(cand[cand.shape[0] // 2])
    cut = int(starts[r])
    w = cut // 32
    a = engine.DeviceStream(ds.codes[:w].clone(), ds.valid[:w].clone(), cut - 1,
                            ds.read_starts[:r].clone(), ds.read_lens[:r].clone())
    b = engine.DeviceStream(ds.codes[w:].clone(), ds.valid[w:].clone(), ds.n_bases - cut,
                            (ds.read_starts[r:] - cut).clone(), ds.read_lens[r:].clone())
    return [a, b]


def test_bench_config_routes_equal_oracle(eng, monkeypatch):
    """8 Mbp x 30x with the routes of the 64 Mbp bench configuration forced (32 hash-range
    bins over L2-sized slices, the two-bit filter in front of the parent tables, and the
    binned parent route): every stage size, the proband-unique keys and all per-read
    (ndistinct, nhits) equal the CPU oracle's.  ~4 s of oracle time."""
    from kmer_denovo_filter_b200.discovery import kmer_chain
    from oracle import ckdf
    k = 31
    trio = _trio(eng, 8_000_000, 30)
    h = {w: _host(trio[w]) for w in ("child", "mother", "father", "ref")}
    want = ckdf.discovery_chain(h["child"], h["mother"][:3], h["father"][:3], h["ref"][:3], k,
                                threads=ckdf.max_threads(),
                                child_capacity=max(int(h["child"][2]) // 4, 1024))
    d = {w: _dev(trio[w]) for w in ("child", "mother", "father", "ref")}
    monkeypatch.setattr(kmer_chain, "SLICE_BYTES", kmer_chain.SLICE_BYTES // 8)
    for route in ("filter", "binned"):
        if route == "binned":
            monkeypatch.setenv("KDF_TABLE_FILTER", "0")
            monkeypatch.setattr(kmer_chain, "PROBE_DIRECT_BYTES", 0)
            monkeypatch.setattr(kmer_chain, "L2_TABLE_BYTES", 0)
        got = kmer_chain.discover_streams(eng, d["child"], d["mother"], d["father"], d["ref"], k)
        for key in ("candidates", "non_ref", "after_mother", "proband_unique", "units"):
            assert got[key] == want[key], (route, key)
        assert got["parents_binned"] == [route == "binned"] * 2
        assert sorted(got["pu"].to_pyints()) == _keys(want["pu_lo"], want["pu_hi"])
        assert np.array_equal(got["ndistinct"], want["nd"]) and np.array_equal(got["nhits"], want["nh"])


@pytest.mark.parametrize("reads", [[], [""], ["ACGT"], ["N" * 200], ["ACGTN" * 40]])
def test_chain_on_degenerate_children(eng, reads):
    """Empty input, reads shorter than k, all-N reads: no k-mers, no crash."""
    import random
    from kmer_denovo_filter_b200.discovery import kmer_chain
    rng = random.Random(1)
    g = "".join(rng.choice("ACGT") for _ in range(500))
    parents = [g[i:i + 100] for i in range(0, 400, 10)]
    res = kmer_chain.discover_streams(eng, _pack_dev(eng, reads), _pack_dev(eng, parents),
                                      _pack_dev(eng, parents), _pack_dev(eng, [g]), 31)
    assert res["child_windows"] == 0 and res["candidates"] == 0 and res["proband_unique"] == 0
    assert res["pu"] is None and res["informative_reads"] == 0
    assert res["units"] == 2 * (100 - 30) * 0 + (500 - 30)     # only the reference was streamed


def test_child_only_kmers_are_all_proband_unique(eng):
    """Parents that share nothing with the child: every child k-mer with count >= 3
    that is not in the reference survives both filters; ragged read lengths."""
    import random
    from kmer_denovo_filter_b200.discovery import kmer_chain
    from oracle import kmers
    rng = random.Random(9)
    k = 21
    gc = "".join(rng.choice("ACGT") for _ in range(2000))
    gp = "".join(rng.choice("ACGT") for _ in range(2000))
    child = [gc[s:s + rng.randint(15, 140)] for s in [rng.randrange(0, 1850) for _ in range(900)]]
    parents = [gp[i:i + 90] for i in range(0, 1900, 5)]
    want = {x for x, c in kmers.count_sequences(child, k).items() if c >= 3}
    want -= set(kmers.count_sequences(parents, k)) | set(kmers.count_sequences([gp], k))
    res = kmer_chain.discover_streams(eng, _pack_dev(eng, child), _pack_dev(eng, parents),
                                      _pack_dev(eng, parents), _pack_dev(eng, [gp]), k)
    assert set(res["pu"].to_pyints()) == want and len(want) > 500
    assert res["after_mother"] == res["proband_unique"] == len(want)
