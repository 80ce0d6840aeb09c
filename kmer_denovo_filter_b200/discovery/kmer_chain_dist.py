"""The discovery k-mer chain across the GPUs of one node (one process per GPU,
``torch.distributed``; NCCL over NVLink on the box, gloo in the CPU tests).

The reference has no multi-device mode; this is how the path
shards:

* reads shard by rank — every rank holds 1/R of each sample's reads and one
  slice of the reference;
* the child k-mer table is partitioned by *owner rank* (``owner_of(hash)``,
  independent of the bucket hash): K6 bins each rank's canonical k-mers by
  owner, one all-to-all routes every k-mer to its owner, and the owner counts
  what it received with the same L2-sliced partitioned count as on one GPU;
  the reference slice takes the same route, so reference subtraction is local
  to the owner;
* what survives (count >= min_child_count, not in the reference) is small:
  it is all-gathered and replicated, each rank probes its parent shards
  against the replica, and the per-key parent counts are summed with one
  all-reduce — in the fixed order of the gathered key list, so every rank
  filters identically;
* the per-read scan is purely data-parallel.
"""

import numpy as np

from .. import engine as _engine
from ..kmer_utils import KmerSet
from . import kmer_chain as _kc


# ---------------------------------------------------------------------------
# collective plumbing (device-agnostic: tested under gloo with CPU tensors)
# ---------------------------------------------------------------------------

def _dist():
    import torch.distributed as dist
    return dist


def exchange_equal(send, world):
    """All-to-all of ``world`` equal segments of ``send`` → tensor of the same shape
    whose segment r came from rank r."""
    dist = _dist()
    recv = send.new_empty(send.shape)
    dist.all_to_all_single(recv, send)
    return recv


def allreduce(t, op="sum"):
    dist = _dist()
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return t


def allgather_varlen(torch, t, world):
    """Concatenation over ranks (rank order) of 1-D tensors of different lengths."""
    dist = _dist()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes + [1])
    pad = t.new_zeros(m)
    pad[:t.shape[0]] = t
    parts = [t.new_empty(m) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


# ---------------------------------------------------------------------------
# stages
# ---------------------------------------------------------------------------

def route_to_owners(eng, streams, k, world):
    """K6 + all-to-all: the canonical k-mers of this rank's ``streams`` go to their
    owner ranks.  Returns ``(recv, recv_counts, bin_cap, windows)``: ``recv`` holds
    ``world`` segments of ``bin_cap`` keys (u64, or {lo, hi} pairs for k > 32),
    segment r = keys sent by rank r of which the first ``recv_counts[r]`` are valid."""
    torch = eng.torch
    n_max = sum(s.n_bases for s in streams)
    cap = torch.tensor([_kc._bin_capacity(max(n_max, 1), world)], dtype=torch.int64, device=eng.device)
    allreduce(cap, "max")          # every rank must use the same segment size
    bin_cap = (int(cap.item()) + 3) & ~3
    while True:
        bins = eng.new_bins(k, world, bin_cap, by_owner=True)
        st = eng.new_stats()
        for s in streams:
            eng.bin_stream(bins, s, st)
        need = torch.stack([bins.cursors.max() if world else bins.cursors.new_zeros(()),
                            bins.overflow[0]]).to(torch.int64)
        allreduce(need, "max")
        if int(need[1].item()) == 0:
            break
        bin_cap = (int(need[0].item()) + 4 + 3) & ~3     # exact size, agreed by all ranks
        del bins
    send_counts = torch.minimum(bins.cursors, torch.full_like(bins.cursors, bin_cap))
    recv_counts = exchange_equal(send_counts, world)
    recv = exchange_equal(bins.data, world)
    return recv, recv_counts.cpu().numpy().astype(np.int64), bins.bin_cap, eng.read_stats(st)["windows"]


# symmetric receive buffers are expensive to set up (handle exchange between the
# ranks): cached per (role, bytes) for the life of the process
_symm_cache = {}


def peer_memory_available(eng):
    """True when the ranks can map each other's memory (NCCL backend on CUDA with
    torch symmetric memory); the gloo / CPU tests take the all-to-all route."""
    if getattr(eng, "peer_bins", True) is False:
        return False
    dist = _dist()
    try:
        if dist.get_backend() != "nccl" or eng.device.type != "cuda":
            return False
        import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
        return True
    except Exception:
        return False


def _symm_buffer(eng, role, n_words):
    import torch.distributed._symmetric_memory as symm_mem
    dist = _dist()
    key = (role, int(n_words))
    hit = _symm_cache.get(key)
    if hit is None:
        # drop smaller buffers of the same role (a retry with larger segments)
        for old in [k2 for k2 in _symm_cache if k2[0] == role]:
            del _symm_cache[old]
        t = symm_mem.empty(int(n_words), dtype=eng.torch.int64, device=eng.device)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD)
        hit = _symm_cache[key] = (t, hdl)
    return hit


class RecvBins:
    """What a rank received through the fused route, seen as the input of
    ``kdf_count_bins_multi``: ``n_src`` x ``n_parts`` bins of ``bin_cap`` keys in the
    (symmetric) receive buffer, with the senders' counts as cursors."""

    def __init__(self, k, key_words, n_parts, n_src, bin_cap, data, cursors):
        self.k, self.key_words = k, key_words
        self.n_parts, self.n_src, self.bin_cap = n_parts, n_src, bin_cap
        self.data, self.cursors = data, cursors


def route_composite_p2p(eng, streams, k, world, role, n_local, n_expected):
    """K6 fused with the exchange AND with the owner's hash-range binning: one kernel
    per rank extracts its k-mers, picks (owner, hash range) and writes each key into
    bin [this rank][range] of the owner's receive buffer through NVLink peer pointers
    (``kdf_bin_stream_to``).  No send buffer, no NCCL data movement, no second
    binning pass on the receiver.  → (RecvBins, windows of this rank's streams)."""
    torch = eng.torch
    dist = _dist()
    rank = dist.get_rank()
    kw = eng.lib.kdf_key_words(k)
    n_bins = world * n_local
    seg_cap = (_kc._bin_capacity(max(n_expected, 1), n_bins) + 3) & ~3
    while True:
        recv, hdl = _symm_buffer(eng, role, n_bins * seg_cap * kw)
        base = [int(hdl.buffer_ptrs[o]) for o in range(world)]
        ptrs = torch.tensor([base[o] + (rank * n_local + p) * seg_cap * kw * 8
                             for o in range(world) for p in range(n_local)],
                            dtype=torch.int64, device=eng.device)
        cursors = torch.zeros(n_bins, dtype=torch.int64, device=eng.device)
        overflow = torch.zeros(1, dtype=torch.int64, device=eng.device)
        st = eng.new_stats()
        hdl.barrier(channel=0)          # every peer is done reading its buffer from the last use
        main = torch.cuda.current_stream(eng.device)
        for s in streams:
            if getattr(s, "chunks", None):           # bin each chunk as soon as it has landed
                for first, n, ev in s.chunks:
                    main.wait_event(ev)
                    eng.bin_stream_to(s, k, ptrs, seg_cap, cursors, overflow, by_owner=world,
                                      stats=st, word_range=(first, n))
            else:
                if getattr(s, "ready", None) is not None:
                    main.wait_event(s.ready)
                eng.bin_stream_to(s, k, ptrs, seg_cap, cursors, overflow, by_owner=world, stats=st)
        hdl.barrier(channel=1)          # every peer's writes into this rank's buffer have landed
        need = torch.stack([cursors.max(), overflow[0]]).to(torch.int64)
        allreduce(need, "max")
        if int(need[1].item()) == 0:
            break
        seg_cap = (int(need[0].item()) + 4 + 3) & ~3      # exact size, agreed by all ranks
    send_counts = torch.minimum(cursors, torch.full_like(cursors, seg_cap))
    recv_counts = exchange_equal(send_counts, world)      # [source][range], 8 bytes each
    return (RecvBins(k, kw, n_local, world, seg_cap, recv, recv_counts),
            eng.read_stats(st)["windows"])


def _segments(recv, counts, bin_cap, kw):
    for r, n in enumerate(counts.tolist()):
        if n:
            yield recv[r * bin_cap * kw:(r * bin_cap + n) * kw], int(n)


def count_child_dist(eng, child_streams, ref_streams, k, min_child_count, world):
    """Module 1 + reference subtraction with the table partitioned by owner rank.
    Returns dict(child_windows, ref_windows, child_distinct, candidates, non_ref — all
    LOCAL to this rank — and lo, hi: this owner's non-reference candidates)."""
    kw = eng.lib.kdf_key_words(k)
    if peer_memory_available(eng):
        return _count_child_fused(eng, child_streams, ref_streams, k, min_child_count, world, kw)
    c_recv, c_counts, c_cap, c_win = route_to_owners(eng, child_streams, k, world)
    r_recv, r_counts, r_cap, r_win = route_to_owners(eng, ref_streams, k, world)
    n_child = int(c_counts.sum())
    n_ref = int(r_counts.sum())
    n_parts, slice_capacity = _kc.plan_partitions(max(n_child, 1), key_words=kw,
                                                  packed=eng.count_bins_packed(k, min_child_count))
    bin_cap = _kc._bin_capacity(max(n_child, 1), n_parts)
    ref_cap = _kc._bin_capacity(max(n_ref, 1), n_parts)
    while True:
        cb = eng.new_bins(k, n_parts, bin_cap)
        rb = eng.new_bins(k, n_parts, ref_cap)
        for seg, n in _segments(c_recv, c_counts, c_cap, kw):
            eng.bin_keys(cb, seg, None, n)
        for seg, n in _segments(r_recv, r_counts, r_cap, kw):
            eng.bin_keys(rb, seg, None, n)
        oc, orf = cb.overflowed(), rb.overflowed()
        if not oc and not orf:
            break
        if oc:
            bin_cap = int(cb.counts().max()) + 4
        if orf:
            ref_cap = int(rb.counts().max()) + 4
        del cb, rb
    del c_recv, r_recv
    out_cap = max(1 << 16, n_child // 64)
    while True:
        res = eng.count_bins(cb, rb, slice_capacity, min0=min_child_count, max1=0,
                             count_min0=min_child_count, out_cap=out_cap)
        if res["full"]:
            if slice_capacity >= 2 * bin_cap:
                raise _engine.KdfError("child k-mer table slice full at %d slots" % slice_capacity)
            slice_capacity = min(slice_capacity * 4, 2 * bin_cap + 4)
            continue
        if res["n_out"] > out_cap:
            out_cap = res["n_out"]
            continue
        break
    return {"child_windows": c_win, "ref_windows": r_win, "child_distinct": res["distinct"],
            "candidates": res["n_count"], "non_ref": res["n_out"], "lo": res["lo"], "hi": res["hi"]}


def _count_child_fused(eng, child_streams, ref_streams, k, min_child_count, world, kw):
    """count_child_dist over NVLink peer memory (see route_composite_p2p)."""
    torch = eng.torch
    n_exp = torch.tensor([sum(s.n_bases for s in child_streams), sum(s.n_bases for s in ref_streams)],
                         dtype=torch.int64, device=eng.device)
    allreduce(n_exp, "max")              # weak scaling: a rank receives about what it sends
    n_child_exp, n_ref_exp = int(n_exp[0].item()), int(n_exp[1].item())
    # L2-sized slices want n_plan hash ranges; the binning kernel takes at most 512
    # (owner x range) bins, so beyond 8 ranks a bin spans `sub` slices and is counted
    # in `sub` passes (kdf_count_bins_multi sub_split).  Measured at 8 ranks: 512 bins
    # (bin + send 43.5 ms, count 33.8 ms) and 256 bins x 2 passes (31.2 + 48.1 ms) tie.
    n_plan, slice_capacity = _kc.plan_partitions(max(n_child_exp, 1), key_words=kw,
                                                 packed=eng.count_bins_packed(k, min_child_count))
    n_local = max(1, min(n_plan, 512 // _kc._pow2_at_least(world)))
    sub = max(1, n_plan // n_local)
    cb, c_win = route_composite_p2p(eng, child_streams, k, world, "child", n_local, n_child_exp)
    rb, r_win = route_composite_p2p(eng, ref_streams, k, world, "ref", n_local, n_ref_exp)
    n_child = int(cb.cursors.sum().item())
    out_cap = max(1 << 16, n_child // 64)
    while True:
        res = eng.count_bins(cb, rb, slice_capacity, min0=min_child_count, max1=0,
                             count_min0=min_child_count, out_cap=out_cap, sub_split=sub)
        if res["full"]:
            if slice_capacity >= 2 * cb.bin_cap * world:
                raise _engine.KdfError("child k-mer table slice full at %d slots" % slice_capacity)
            slice_capacity = min(slice_capacity * 4, 2 * cb.bin_cap * world + 4)
            continue
        if res["n_out"] > out_cap:
            out_cap = res["n_out"]
            continue
        break
    return {"child_windows": c_win, "ref_windows": r_win, "child_distinct": res["distinct"],
            "candidates": res["n_count"], "non_ref": res["n_out"], "lo": res["lo"], "hi": res["hi"]}


def filter_parent_dist(eng, parent_stream, k, lo, hi, parent_max_count, world, stats):
    """``count --if`` of this rank's parent shard against the replicated key list,
    all-reduce of the per-key counts (list order), threshold.  → (lo, hi) survivors,
    identical on every rank."""
    n = int(lo.shape[0])
    table = _kc._primed_table(eng, k, lo, hi, n)
    binned = _kc.count_if_present(eng, table, parent_stream, stats)
    _found, p0, _p1 = eng.lookup_keys(table, lo, hi)
    table.close()
    total = p0.to(eng.torch.int64)
    allreduce(total, "sum")
    keep = total <= parent_max_count
    return lo[keep].contiguous(), (hi[keep].contiguous() if hi is not None else None), binned


def discover_streams_dist(eng, child, mother, father, ref, k, min_child_count=3,
                          parent_max_count=0, min_distinct_kmers_per_read=None, fetch=False):
    """Multi-GPU form of :func:`kmer_chain.discover_streams`; every argument is this
    rank's shard.  Stage sizes in the result are GLOBAL (identical on all ranks);
    ``units`` and the per-read records are this rank's own."""
    dist = _dist()
    torch = eng.torch
    world = dist.get_world_size()
    if min_distinct_kmers_per_read is None:
        min_distinct_kmers_per_read = max(1, k // 4)
    stats = eng.new_stats()
    up = _kc._Uploader(eng)
    if peer_memory_available(eng):
        # fused route: the child is copied in chunks and binned (= sent) as it lands
        d_child, ev_child = up.put_chunked(child, False)
        d_ref, ev_ref = up.put(ref, False)
        fused = True
    else:
        d_child, ev_child = up.put(child, True)
        d_ref, ev_ref = up.put(ref, False)
        fused = False
        up.wait(ev_child)
        up.wait(ev_ref)
    d_mother, ev_mother = up.put(mother, False)
    d_father, ev_father = up.put(father, False)
    ev_reads = up.put_read_index(d_child, child) if fused else None   # needed last: copied last

    c = count_child_dist(eng, [d_child], [d_ref], k, min_child_count, world)
    tot = torch.tensor([c["candidates"], c["child_distinct"]], dtype=torch.int64, device=eng.device)
    allreduce(tot, "sum")
    lo = allgather_varlen(torch, c["lo"], world)
    hi = allgather_varlen(torch, c["hi"], world) if c["hi"] is not None else None
    out = {"child_windows": c["child_windows"], "child_distinct": int(tot[1].item()),
           "candidates": int(tot[0].item()), "non_ref": int(lo.shape[0]), "after_mother": 0,
           "proband_unique": 0, "pu": None, "ndistinct": None, "nhits": None,
           "informative_reads": 0, "reads": None, "hits": None, "parents_binned": []}
    units = c["child_windows"] + c["ref_windows"]

    n_pu = 0
    if out["non_ref"]:
        up.wait(ev_mother)
        lo, hi, b = filter_parent_dist(eng, d_mother, k, lo, hi, parent_max_count, world, stats)
        out["parents_binned"].append(b)
        out["after_mother"] = int(lo.shape[0])
        if out["after_mother"]:
            up.wait(ev_father)
            lo, hi, b = filter_parent_dist(eng, d_father, k, lo, hi, parent_max_count, world, stats)
            out["parents_binned"].append(b)
            n_pu = int(lo.shape[0])
    out["proband_unique"] = n_pu

    local_inf = 0
    if n_pu:
        out["pu"] = KmerSet(eng, k, lo, hi)
        pt = _kc._primed_table(eng, k, lo, hi, n_pu)
        up.wait(ev_child)
        up.wait(ev_reads)
        sp = eng.scan_reads_sparse(pt, d_child, stats=stats)
        pt.close()
        out["reads"] = sp
        local_inf = int((sp["ndistinct"] >= min_distinct_kmers_per_read).sum())
        if fetch:
            nd = np.zeros(d_child.n_reads, dtype=np.uint32)
            nh = np.zeros(d_child.n_reads, dtype=np.uint32)
            nd[sp["read"].astype(np.int64)] = sp["ndistinct"]
            nh[sp["read"].astype(np.int64)] = sp["nhits"]
            out["ndistinct"], out["nhits"] = nd, nh
    inf = torch.tensor([local_inf], dtype=torch.int64, device=eng.device)
    allreduce(inf, "sum")
    out["informative_reads"] = int(inf.item())
    out["informative_reads_local"] = local_inf
    out["units"] = eng.read_stats(stats)["windows"] + units
    return out
