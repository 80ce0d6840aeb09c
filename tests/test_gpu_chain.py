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
