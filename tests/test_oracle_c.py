"""The oracle's C/OpenMP twin against the numpy oracle (and through it the goldens)."""
import json
import random

import pytest

from kmer_denovo_filter_b200 import engine
from oracle import bam as obam
from oracle import ckdf, kmers


def _pack(seqs):
    hs = engine.pack_sequences(seqs)
    return hs.codes, hs.valid, hs.n_bases, hs.read_starts, hs.read_lens


@pytest.mark.parametrize("k", [5, 31, 32, 33, 63])
@pytest.mark.parametrize("threads", [1, 4])
def test_c_count_equals_numpy(k, threads):
    rng = random.Random(k)
    g = "".join(rng.choice("ACGT") for _ in range(5000))
    reads = []
    for _ in range(400):
        s = rng.randrange(0, 4800)
        r = g[s:s + rng.randint(0, 180)]
        if rng.random() < 0.2 and r:
            i = rng.randrange(len(r))
            r = r[:i] + "N" + r[i + 1:]
        reads.append(r)
    want = kmers.count_sequences(reads, k)
    codes, valid, n, _s, _l = _pack(reads)
    t = ckdf.Table(2 * len(want) + 16)
    nwin = t.count_stream(codes, valid, n, k, ckdf.MODE_INSERT_COUNT, 0, 1, threads)
    assert nwin == sum(want.values())
    cnt, lo, hi, p0, _p1 = t.threshold()
    got = {(int(h) << 64) | int(l): int(c) for l, h, c in zip(lo.tolist(), hi.tolist(), p0.tolist())}
    assert got == want


def test_c_chain_on_giab(giab_records, giab_paths):
    exp = json.load(open(giab_paths["expected_json"]))
    streams = {}
    for who in ("child", "mother", "father"):
        streams[who] = _pack([r.seq for r in obam.fasta_stream(giab_records[who])])
    ref = _pack([s for _n, s in giab_records["ref"]])
    out = ckdf.discovery_chain(streams["child"][:3], streams["mother"][:3], streams["father"][:3],
                               ref[:3], 31, threads=4)
    assert (out["candidates"], out["non_ref"], out["after_mother"], out["proband_unique"]) == \
        (exp["candidates"], exp["non_ref"], exp["after_mother"], exp["proband_unique"])
    pu = sorted(kmers.kmer_of((int(h) << 64) | int(l), 31)
                for l, h in zip(out["pu_lo"].tolist(), out["pu_hi"].tolist()))
    assert pu == exp["proband_unique_kmers"]
    # per-read scan over the anchoring record stream
    scan = [r for r in giab_records["child"] if not (r.flag & 0x500)]
    codes, valid, n, rs, rl = _pack([r.seq for r in scan])
    tp = ckdf.Table(4096)
    tp.update_keys(out["pu_lo"], out["pu_hi"], threads=2)
    nd, nh, _nwin = tp.scan_reads(codes, valid, n, rs, rl, 31, threads=4)
    idx = [i for i, r in enumerate(giab_records["child"]) if not (r.flag & 0x500)]
    got = [[idx[j], int(nd[j]), int(nh[j])] for j in range(len(scan)) if nd[j] > 0]
    assert got == exp["per_read_informative"]
