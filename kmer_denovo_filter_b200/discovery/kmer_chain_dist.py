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

def route_to_owners(eng, streams, k, world, pass_=None, n_passes=1):
    """K6 + all-to-all: the canonical k-mers of this rank's ``streams`` go to their
    owner ranks (with ``pass_`` only those of one hash-range group, see
    ``kmer_chain.plan_child_count``).  Returns ``(recv, recv_counts, bin_cap, windows)``:
    ``recv`` holds ``world`` segments of ``bin_cap`` keys (u64, or {lo, hi} pairs for
    k > 32), segment r = keys sent by rank r of which the first ``recv_counts[r]`` are valid."""
    torch = eng.torch
    n_max = sum(s.n_bases for s in streams)
    cap = torch.tensor([_kc._bin_capacity(max(n_max, 1) / n_passes, world)], dtype=torch.int64,
                       device=eng.device)
    allreduce(cap, "max")          # every rank must use the same segment size
    bin_cap = (int(cap.item()) + 3) & ~3
    while True:
        bins = eng.new_bins(k, world, bin_cap, by_owner=True)
        st = eng.new_stats()
        for s in streams:
            _kc._wait_ready(eng, s)
            if pass_ is None:
                eng.bin_stream(bins, s, st)
            else:
                eng.bin_stream(bins, s, st, pass_=pass_)
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


def route_composite_p2p(eng, streams, k, world, role, n_local, n_expected, pass_=None, seg_cap=None):
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
    if seg_cap is None:
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
                                      stats=st, word_range=(first, n), pass_=pass_)
            else:
                if getattr(s, "ready", None) is not None:
                    main.wait_event(s.ready)
                eng.bin_stream_to(s, k, ptrs, seg_cap, cursors, overflow, by_owner=world, stats=st,
                                  pass_=pass_)
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


def _agree_max(eng, *vals):
    t = eng.torch.tensor(list(vals), dtype=eng.torch.int64, device=eng.device)
    allreduce(t, "max")
    return [int(x) for x in t.tolist()]


def count_child_dist(eng, child_streams, ref_streams, k, min_child_count, world, n_passes=None):
    """Module 1 + reference subtraction with the table partitioned by owner rank.
    Returns dict(child_windows, ref_windows, child_distinct, candidates, non_ref — all
    LOCAL to this rank — and lo, hi: this owner's non-reference candidates).  The hash
    ranges of an owner are taken in ``n_passes`` groups when what a rank would receive
    in one go does not fit its memory (``kmer_chain.plan_child_count``; every rank
    uses the same plan)."""
    kw = eng.lib.kdf_key_words(k)
    fused = peer_memory_available(eng)
    n_child_exp, n_ref_exp = _agree_max(eng, sum(s.n_bases for s in child_streams),
                                        sum(s.n_bases for s in ref_streams))
    # weak scaling: a rank receives about what it sends.  The fused route takes at most
    # 512 (owner x range) bins per pass
    max_local = max(1, 512 // _kc._pow2_at_least(world)) if fused else _kc.MAX_PARTS
    plan = _kc.plan_child_count(eng, max(n_child_exp, 1), n_ref_exp, k, min_child_count,
                                n_passes=n_passes, max_local=max_local)
    n_passes, n_local, slice_capacity = _agree_max(eng, *plan)
    n_total = _kc._pow2_at_least(n_passes * n_local)
    n_passes = _kc._pow2_at_least(n_passes)
    n_local = max(1, n_total // n_passes)
    plog = n_passes.bit_length() - 1
    tot = {"distinct": 0, "n_count": 0, "n_out": 0}
    c_win = r_win = 0
    los, his = [], []
    seg_c = seg_r = None
    for p in range(n_passes):
        pass_ = (plog, p) if n_passes > 1 else None
        if fused:
            cb, w1 = route_composite_p2p(eng, child_streams, k, world, "child", n_local,
                                         n_child_exp // n_passes, pass_, seg_c)
            rb, w2 = route_composite_p2p(eng, ref_streams, k, world, "ref", n_local,
                                         n_ref_exp // n_passes, pass_, seg_r)
            seg_c, seg_r = cb.bin_cap, rb.bin_cap       # sizes that worked: keep them
            n_child = int(cb.cursors.sum().item())
            cap_limit = 2 * cb.bin_cap * world
        else:
            c_recv, c_counts, c_cap, w1 = route_to_owners(eng, child_streams, k, world, pass_, n_passes)
            r_recv, r_counts, r_cap, w2 = route_to_owners(eng, ref_streams, k, world, pass_, n_passes)
            n_child = int(c_counts.sum())
            n_ref = int(r_counts.sum())
            bin_cap = _kc._bin_capacity(max(n_child, 1), n_local)
            ref_cap = _kc._bin_capacity(max(n_ref, 1), n_local)
            while True:
                cb = eng.new_bins(k, n_local, bin_cap)
                rb = eng.new_bins(k, n_local, ref_cap)
                for seg, n in _segments(c_recv, c_counts, c_cap, kw):
                    eng.bin_keys(cb, seg, None, n, pass_=pass_)
                for seg, n in _segments(r_recv, r_counts, r_cap, kw):
                    eng.bin_keys(rb, seg, None, n, pass_=pass_)
                oc, orf = cb.overflowed(), rb.overflowed()
                if not oc and not orf:
                    break
                if oc:
                    bin_cap = int(cb.counts().max()) + 4
                if orf:
                    ref_cap = int(rb.counts().max()) + 4
                del cb, rb
            del c_recv, r_recv
            cap_limit = 2 * bin_cap
        c_win += w1
        r_win += w2
        out_cap = max(1 << 16, n_child // 64)
        while True:
            res = eng.count_bins(cb, rb, slice_capacity, min0=min_child_count, max1=0,
                                 count_min0=min_child_count, out_cap=out_cap, pass_=pass_)
            if res["full"]:
                if slice_capacity >= cap_limit:
                    raise _engine.KdfError("child k-mer table slice full at %d slots" % slice_capacity)
                slice_capacity = min(slice_capacity * 4, cap_limit + 4)
                continue
            if res["n_out"] > out_cap:
                out_cap = res["n_out"]
                continue
            break
        for key in tot:
            tot[key] += res[key]
        los.append(res["lo"])
        if res["hi"] is not None:
            his.append(res["hi"])
        del cb, rb
    torch = eng.torch
    lo = los[0] if len(los) == 1 else torch.cat(los)
    hi = (his[0] if len(his) == 1 else torch.cat(his)) if his else None
    return {"child_windows": c_win, "ref_windows": r_win, "child_distinct": tot["distinct"],
            "candidates": tot["n_count"], "non_ref": tot["n_out"], "lo": lo, "hi": hi,
            "n_passes": n_passes, "n_local": n_local}


def filter_parent_dist(eng, parent_stream, k, lo, hi, parent_max_count, world, stats):
    """``count --if`` of this rank's parent shard against the replicated key list,
    all-reduce of the per-key counts (list order), threshold.  → (lo, hi) survivors,
    identical on every rank."""
    n = int(lo.shape[0])
    table = _kc._primed_table(eng, k, lo, hi, n)
    binned = _kc.count_if_present(eng, table, parent_stream, stats)
    _found, p0, _p1 = eng.lookup_keys(table, lo, hi)
    table.close()
    # (a count saturates nowhere near 2^31 per rank x 8 ranks on this path; int32 halves the bytes)
    total = p0.contiguous()
    allreduce(total, "sum")
    keep = total <= parent_max_count
    return lo[keep].contiguous(), (hi[keep].contiguous() if hi is not None else None), binned


def discover_streams_dist(eng, child, mother, father, ref, k, min_child_count=3,
                          parent_max_count=0, min_distinct_kmers_per_read=None, fetch=False,
                          n_passes=None, child_scan=None):
    """Multi-GPU form of :func:`kmer_chain.discover_streams`; every argument is this
    rank's shard (one stream, or a list of streams).  Stage sizes in the result are
    GLOBAL (identical on all ranks); ``units`` and the per-read records are this rank's own.
    ``child_scan``: the streams the per-read scan runs over when they differ from the
    counting streams (the BAM pipeline counts the ``samtools fasta`` view of a batch and
    scans the batch itself); ``reads_parts`` then holds one record set per scan stream."""
    dist = _dist()
    torch = eng.torch
    world = dist.get_world_size()
    if min_distinct_kmers_per_read is None:
        min_distinct_kmers_per_read = max(1, k // 4)
    stats = eng.new_stats()
    up = _kc._Uploader(eng)
    childs, refs = _kc._as_list(child), _kc._as_list(ref)
    if peer_memory_available(eng):
        # fused route: the child is copied in chunks and binned (= sent) as it lands
        d_childs = [up.put_chunked(x, False)[0] for x in childs]
        d_refs = [up.put(x, False)[0] for x in refs]
        fused = True
    else:
        d_childs = [up.put(x, True)[0] for x in childs]
        d_refs = [up.put(x, False)[0] for x in refs]
        fused = False
    d_mothers = [up.put(x, False)[0] for x in _kc._as_list(mother)]
    d_fathers = [up.put(x, False)[0] for x in _kc._as_list(father)]
    # needed last: copied last
    ev_reads = [up.put_read_index(d, h) for d, h in zip(d_childs, childs)] if (fused and child_scan is None) else []

    c = count_child_dist(eng, d_childs, d_refs, k, min_child_count, world, n_passes=n_passes)
    tot = torch.tensor([c["candidates"], c["child_distinct"]], dtype=torch.int64, device=eng.device)
    allreduce(tot, "sum")
    lo = allgather_varlen(torch, c["lo"], world)
    hi = allgather_varlen(torch, c["hi"], world) if c["hi"] is not None else None
    out = {"child_windows": c["child_windows"], "child_distinct": int(tot[1].item()),
           "candidates": int(tot[0].item()), "non_ref": int(lo.shape[0]), "after_mother": 0,
           "proband_unique": 0, "pu": None, "ndistinct": None, "nhits": None,
           "informative_reads": 0, "reads": None, "hits": None, "parents_binned": [],
           "n_passes": c["n_passes"], "n_local": c["n_local"]}
    units = c["child_windows"] + c["ref_windows"]

    n_pu = 0
    if out["non_ref"]:
        lo, hi, b = filter_parent_dist(eng, d_mothers, k, lo, hi, parent_max_count, world, stats)
        out["parents_binned"].append(b)
        out["after_mother"] = int(lo.shape[0])
        if out["after_mother"]:
            lo, hi, b = filter_parent_dist(eng, d_fathers, k, lo, hi, parent_max_count, world, stats)
            out["parents_binned"].append(b)
            n_pu = int(lo.shape[0])
    out["proband_unique"] = n_pu

    local_inf = 0
    if n_pu:
        out["pu"] = KmerSet(eng, k, lo, hi)
        pt = _kc._primed_table(eng, k, lo, hi, n_pu)
        for ev in ev_reads:
            up.wait(ev)
        parts = []
        if child_scan is not None:
            d_scan = []
            for h in _kc._as_list(child_scan):      # uploaded one at a time: they are only scanned
                d = eng.upload(h) if not isinstance(h, _engine.DeviceStream) else h
                parts.append(eng.scan_reads_sparse(pt, d, stats=stats))
                d_scan.append(d)
            d_childs_scan = d_scan
        else:
            d_childs_scan = d_childs
            for d in d_childs:
                _kc._wait_ready(eng, d)
                parts.append(eng.scan_reads_sparse(pt, d, stats=stats))
        pt.close()
        out["reads_parts"] = parts
        sp = _kc.merge_sparse_records(parts, d_childs_scan)
        out["reads"] = sp
        local_inf = int((sp["ndistinct"] >= min_distinct_kmers_per_read).sum())
        if fetch:
            n_reads = sum(d.n_reads for d in d_childs_scan)
            nd = np.zeros(n_reads, dtype=np.uint32)
            nh = np.zeros(n_reads, dtype=np.uint32)
            nd[sp["read"].astype(np.int64)] = sp["ndistinct"]
            nh[sp["read"].astype(np.int64)] = sp["nhits"]
            out["ndistinct"], out["nhits"] = nd, nh
    inf = torch.tensor([local_inf], dtype=torch.int64, device=eng.device)
    allreduce(inf, "sum")
    out["informative_reads"] = int(inf.item())
    out["informative_reads_local"] = local_inf
    out["units"] = eng.read_stats(stats)["windows"] + units
    return out
