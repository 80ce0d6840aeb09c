"""A numpy stand-in for CudaEngine (TEST INFRASTRUCTURE ONLY).

It implements just the engine surface that the distributed chain
(`kmer_chain_dist`) uses, on CPU tensors, so that the multi-rank host logic
(owner routing, all-to-all, replicated filtering, all-reduce of parent counts)
can run under gloo with world_size 2 in a container without a GPU.  The k-mer
arithmetic comes from the oracle; owner / partition assignment comes from the
library's own host-side hash hook, i.e. the same code the kernels use.
"""
import collections

import numpy as np
import torch

from kmer_denovo_filter_b200 import engine
from oracle import kmers

U32_MAX = 0xFFFFFFFF


def unpack_stream(ds):
    """DeviceStream on CPU tensors -> (codes u8[n], valid bool[n])."""
    n = ds.n_bases
    cw = ds.codes.numpy().view(np.uint64)
    vw = ds.valid.numpy().view(np.uint32)
    sh = (62 - 2 * np.arange(32, dtype=np.uint64)).astype(np.uint64)
    codes = ((cw[:, None] >> sh[None, :]) & np.uint64(3)).astype(np.uint8).ravel()[:n]
    vs = (31 - np.arange(32, dtype=np.uint32)).astype(np.uint32)
    valid = (((vw[:, None] >> vs[None, :]) & np.uint32(1)) != 0).ravel()[:n]
    return codes, valid


def stream_keys(ds, k):
    codes, valid = unpack_stream(ds)
    hi, lo, ok = kmers.canonical_windows(codes, valid, k)
    return lo, hi, ok


class FakeBins:
    def __init__(self, k, n_parts, bin_cap, by_owner):
        self.k = k
        self.key_words = 1 if k <= 32 else 2
        self.n_parts = n_parts
        self.bin_cap = max(4, (int(bin_cap) + 3) & ~3)
        self.by_owner = by_owner
        self.data = torch.zeros(n_parts * self.bin_cap * self.key_words, dtype=torch.int64)
        self.cursors = torch.zeros(n_parts, dtype=torch.int64)
        self.overflow = torch.zeros(1, dtype=torch.int64)

    def counts(self):
        return self.cursors.numpy().astype(np.int64)

    def overflowed(self):
        return bool(int(self.overflow.item()))

    def keys_of(self, b):
        n = min(int(self.cursors[b]), self.bin_cap)
        kw = self.key_words
        seg = self.data[b * self.bin_cap * kw:(b * self.bin_cap + n) * kw].numpy().view(np.uint64)
        if kw == 1:
            return seg, np.zeros_like(seg)
        return seg[0::2], seg[1::2]


class FakeTable:
    def __init__(self, k, capacity):
        self.k = k
        self.key_words = 1 if k <= 32 else 2
        self.capacity = capacity
        self.d = {}

    def close(self):
        self.d = None


class _Lib:
    @staticmethod
    def kdf_key_words(k):
        return 1 if 1 <= k <= 32 else (2 if k <= 64 else 0)


class FakeEngine:
    def __init__(self):
        self.torch = torch
        self.device = torch.device("cpu")
        self.lib = _Lib()
        self.launches = 0

    # -- plumbing
    def new_stats(self):
        return torch.zeros(4, dtype=torch.int64)

    def read_stats(self, st):
        v = st.numpy()
        return {"windows": int(v[0]), "full": int(v[1]), "hits": int(v[2]), "new": int(v[3])}

    # -- data movement: host stream -> "device" stream of CPU tensors
    def upload(self, hs, non_blocking=False, with_reads=True, copy_stream=None):
        def t(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a).view(dt).copy())
        rs = t(hs.read_starts, np.int64) if (with_reads and hs.read_lens is not None) else None
        rl = t(hs.read_lens, np.int32) if (with_reads and hs.read_lens is not None) else None
        return engine.DeviceStream(t(hs.codes, np.int64), t(hs.valid, np.int32), hs.n_bases, rs, rl)

    def hit_coverage(self, *a):
        """K7 through the host instantiation of the device code (pinned to the reference-named
        helper by tests/test_postprocess_cpu.py)."""
        return engine.debug_hit_coverage_host(*a)

    # -- binning
    def new_bins(self, k, n_parts, bin_cap, by_owner=False):
        return FakeBins(k, n_parts, bin_cap, by_owner)

    def _append(self, bins, lo, hi, pass_=None):
        kw = bins.key_words
        log2p = bins.n_parts.bit_length() - 1
        plog, pval = pass_ if pass_ is not None else (0, 0)
        # hash ranges: the top plog bits pick the pass, the next log2p bits the local bin
        part, _b, owner = engine.debug_hash_host(lo, hi if kw == 2 else None, kw,
                                                 plog + (0 if bins.by_owner else log2p), 1024,
                                                 bins.n_parts if bins.by_owner else 1)
        if bins.by_owner:
            in_pass = (part == pval) if plog else np.ones(lo.shape[0], dtype=bool)
            which = owner
        else:
            in_pass = ((part >> log2p) == pval) if plog else np.ones(lo.shape[0], dtype=bool)
            which = part & (bins.n_parts - 1)
        data = bins.data.numpy().view(np.uint64)
        for b in range(bins.n_parts):
            sel = np.flatnonzero((which == b) & in_pass)
            cur = int(bins.cursors[b])
            room = max(0, bins.bin_cap - cur)
            take = sel[:room]
            if take.size < sel.size:
                bins.overflow[0] = 1
            base = (b * bins.bin_cap + cur) * kw
            if kw == 1:
                data[base:base + take.size] = lo[take]
            else:
                data[base:base + 2 * take.size:2] = lo[take]
                data[base + 1:base + 2 * take.size:2] = hi[take]
            bins.cursors[b] = cur + sel.size
        return int(in_pass.sum())

    def bin_stream(self, bins, ds, stats=None, word_range=None, pass_=None):
        lo, hi, ok = stream_keys(ds, bins.k)
        n = self._append(bins, lo[ok], hi[ok], pass_)
        if stats is not None:
            stats[0] += n

    def bin_keys(self, bins, lo, hi=None, n=None, pass_=None):
        n = int(lo.shape[0]) if n is None else int(n)
        a = lo.numpy().view(np.uint64)
        if bins.key_words == 1:
            self._append(bins, a[:n].copy(), np.zeros(n, np.uint64), pass_)
        elif hi is None:
            self._append(bins, a[0:2 * n:2].copy(), a[1:2 * n:2].copy(), pass_)
        else:
            self._append(bins, a[:n].copy(), hi.numpy().view(np.uint64)[:n].copy(), pass_)

    def count_bins_packed(self, k, min_child_count):
        return False

    FILTER_MAX_BYTES = 32 << 20

    def filter_applies(self, k, n_keys):
        return False

    def build_filter(self, table, n_keys, max_bytes=1 << 30):
        pass

    def count_bins(self, cb, rb, slice_capacity, min0=0, max0=U32_MAX, min1=0, max1=U32_MAX,
                   count_min0=0, out_cap=1 << 20, want_planes=False, sub_split=1, pass_=None):
        cnt = collections.Counter()
        n_keys = 0
        full = 0
        for b in range(cb.n_parts):
            lo, hi = cb.keys_of(b)
            n_keys += lo.shape[0]
            part = collections.Counter(kmers.to_pyints(hi, lo))
            if len(part) > slice_capacity:
                full = 1
            cnt.update(part)
        ref = set()
        if rb is not None:
            for b in range(rb.n_parts):
                lo, hi = rb.keys_of(b)
                ref.update(kmers.to_pyints(hi, lo))
        keep = sorted(x for x, c in cnt.items()
                      if min0 <= c <= max0 and min1 <= (1 if x in ref else 0) <= max1)
        m = min(len(keep), out_cap)
        lo = np.array([x & 0xFFFFFFFFFFFFFFFF for x in keep[:m]], dtype=np.uint64)
        hi = np.array([x >> 64 for x in keep[:m]], dtype=np.uint64)
        return {"n_out": len(keep), "lo": torch.from_numpy(lo.view(np.int64)),
                "hi": torch.from_numpy(hi.view(np.int64)) if cb.key_words == 2 else None,
                "p0": None, "p1": None, "keys": n_keys, "full": full,
                "hits": n_keys - len(cnt), "distinct": len(cnt),
                "n_count": sum(1 for c in cnt.values() if c >= count_min0), "occupied": len(cnt)}

    # -- tables
    def new_table(self, k, n_keys=None, capacity=None):
        return FakeTable(k, capacity or 2 * (n_keys or 0))

    @staticmethod
    def _keys(lo, hi):
        l = lo.numpy().view(np.uint64)
        h = hi.numpy().view(np.uint64) if hi is not None else np.zeros_like(l)
        return kmers.to_pyints(h, l)

    def update_keys(self, table, lo, hi=None, mode=engine.MODE_INSERT_ONLY, plane=0, arg=1, stats=None):
        for x in self._keys(lo, hi):
            if mode in (engine.MODE_INSERT_COUNT, engine.MODE_INSERT_ONLY):
                v = table.d.setdefault(x, [0, 0])
                if mode == engine.MODE_INSERT_COUNT:
                    v[plane] += arg
            elif x in table.d:
                if mode == engine.MODE_COUNT_IF_PRESENT:
                    table.d[x][plane] += arg
                else:
                    table.d[x][plane] |= arg

    def count_stream(self, table, ds, mode=engine.MODE_INSERT_COUNT, plane=0, arg=1, stats=None):
        lo, hi, ok = stream_keys(ds, table.k)
        if stats is not None:
            stats[0] += int(ok.sum())
        assert mode == engine.MODE_COUNT_IF_PRESENT
        for x, c in collections.Counter(kmers.to_pyints(hi[ok], lo[ok])).items():
            if x in table.d:
                table.d[x][plane] += c * arg

    def lookup_keys(self, table, lo, hi=None, want_planes=True):
        keys = self._keys(lo, hi)
        found = np.array([x in table.d for x in keys], dtype=np.uint8)
        p0 = np.array([table.d.get(x, [0, 0])[0] for x in keys], dtype=np.int32)
        p1 = np.array([table.d.get(x, [0, 0])[1] for x in keys], dtype=np.int32)
        return torch.from_numpy(found), torch.from_numpy(p0), torch.from_numpy(p1)

    def scan_reads_sparse(self, table, ds, stats=None, hit_cap=None):
        lo, hi, ok = stream_keys(ds, table.k)
        if stats is not None:
            stats[0] += int(ok.sum())
        keys = kmers.to_pyints(hi, lo)
        starts = ds.read_starts.numpy().view(np.uint64).astype(np.int64)
        lens = ds.read_lens.numpy().view(np.uint32).astype(np.int64)
        rec = {"read": [], "ndistinct": [], "nhits": [], "first": [], "hit_pos": [], "hit_slot": []}
        for r, (s, l) in enumerate(zip(starts.tolist(), lens.tolist())):
            hits = [p for p in range(s, s + max(0, l - table.k + 1)) if ok[p] and keys[p] in table.d]
            if hits:
                rec["read"].append(r)
                rec["ndistinct"].append(len({keys[p] for p in hits}))
                rec["nhits"].append(len(hits))
                rec["first"].append(len(rec["hit_pos"]))
                rec["hit_pos"].extend(hits)
        return {"read": np.array(rec["read"], np.uint64), "ndistinct": np.array(rec["ndistinct"], np.uint32),
                "nhits": np.array(rec["nhits"], np.uint32), "first": np.array(rec["first"], np.uint64),
                "hit_pos": np.array(rec["hit_pos"], np.uint64),
                "hit_slot": np.zeros(len(rec["hit_pos"]), np.uint32)}
