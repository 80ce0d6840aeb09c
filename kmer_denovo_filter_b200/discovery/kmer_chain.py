"""The k-mer part of the discovery path on packed streams (no BAM I/O).

This is the hot path BASELINE.json's metric is quoted on, in one call: what the
reference does with ``samtools fasta | jellyfish count``, ``dump -L``,
``query ref.jf``, two ``count --if`` + ``query`` parent passes and the per-read
``jellyfish query`` scan (reference ``discovery/pipeline.py:69-612`` and
``core/bam_scanner.py:340-507``), given the four read streams already decoded
and 2-bit packed (``include/kdf.h`` stream layout).

``run_discovery_pipeline`` reaches the same kernels batch by batch while it
decodes BAMs; ``bench.py`` and the parity tests call this function with whole
samples, either resident in HBM (``DeviceStream``) or in host memory
(``HostStream`` — uploaded here, inside the caller's timed region).
"""

import numpy as np

from .. import engine as _engine
from ..kmer_utils import KmerSet

REF_PLANE = 1


def _on_device(eng, s, with_reads):
    if isinstance(s, _engine.DeviceStream):
        return s
    return eng.upload(s, non_blocking=True, with_reads=with_reads)


def default_child_capacity(eng, n_bases):
    return eng.capacity_for(max(n_bases // 8, 512))


def _primed_table(eng, k, lo, hi, n):
    t = eng.new_table(k, n_keys=max(n, 1))
    eng.update_keys(t, lo, hi, _engine.MODE_INSERT_ONLY, 0, 0)
    return t


def discover_streams(eng, child, mother, father, ref, k, min_child_count=3,
                     parent_max_count=0, min_distinct_kmers_per_read=None,
                     child_capacity=None, want_hits=False, fetch=True):
    """Child count → threshold → reference subtraction → mother / father
    filtered counts → proband-unique set → per-read distinct-hit reduction.

    Returns a dict: stage sizes (``candidates``, ``non_ref``, ``after_mother``,
    ``proband_unique``), ``pu`` (:class:`KmerSet` or None), per-read
    ``ndistinct`` / ``nhits`` (numpy u32 when ``fetch`` else device tensors),
    ``informative_reads`` and ``units`` = valid k-mer instances processed over
    all stages (the unit of BASELINE.json's metric)."""
    if min_distinct_kmers_per_read is None:
        min_distinct_kmers_per_read = max(1, k // 4)   # discovery/pipeline.py:2119-2121
    stats = eng.new_stats()
    d_child = _on_device(eng, child, True)
    d_ref = _on_device(eng, ref, False)
    d_mother = _on_device(eng, mother, False)
    d_father = _on_device(eng, father, False)

    # Module 1: jellyfish count -C ; dump -L min_child_count
    # Sized for 4 bases per distinct k-mer (30x data holds ~14 bases per distinct
    # k-mer); when that is too small the count is redone in a larger table — never
    # a silent drop (Jellyfish would spill to .jf_N files and merge).
    if child_capacity is None:
        child_capacity = default_child_capacity(eng, d_child.n_bases)
    while True:
        table = eng.new_table(k, capacity=child_capacity)
        eng.count_stream(table, d_child, _engine.MODE_INSERT_COUNT, 0, 1, stats)
        st = eng.read_stats(stats)
        if not st["full"]:
            child_windows, child_distinct = st["windows"], st["new"]
            break
        table.close()
        if child_capacity >= 2 * d_child.n_bases:
            raise _engine.KdfError("child k-mer table full at %d slots" % child_capacity)
        child_capacity = min(child_capacity * 4, eng.capacity_for(d_child.n_bases))
        stats.zero_()
    # reference subtraction: stream the reference against the child table
    eng.count_stream(table, d_ref, _engine.MODE_MARK_IF_PRESENT, REF_PLANE, 1, stats)
    n_cand = eng.threshold_count(table, min0=min_child_count)
    n_nonref, lo, hi, _a, _b = eng.threshold_compact(table, min0=min_child_count, max1=0)
    eng.check_not_full(stats)
    table.close()
    out = {"child_windows": child_windows, "child_distinct": child_distinct,
           "child_capacity": child_capacity, "candidates": n_cand, "non_ref": n_nonref, "after_mother": 0, "proband_unique": 0,
           "pu": None, "ndistinct": None, "nhits": None, "informative_reads": 0, "hits": None}

    # Module 2: count --if against mother, then father
    n_pu = 0
    if n_nonref:
        mt = _primed_table(eng, k, lo, hi, n_nonref)
        eng.count_stream(mt, d_mother, _engine.MODE_COUNT_IF_PRESENT, 0, 1, stats)
        n_am, lo, hi, _a, _b = eng.threshold_compact(mt, max0=parent_max_count)
        mt.close()
        out["after_mother"] = n_am
        if n_am:
            ft = _primed_table(eng, k, lo, hi, n_am)
            eng.count_stream(ft, d_father, _engine.MODE_COUNT_IF_PRESENT, 0, 1, stats)
            n_pu, lo, hi, _a, _b = eng.threshold_compact(ft, max0=parent_max_count)
            ft.close()
    out["proband_unique"] = n_pu

    # Modules 2b + 3: membership table, per-read scan
    if n_pu:
        out["pu"] = KmerSet(eng, k, lo, hi)
        pt = _primed_table(eng, k, lo, hi, n_pu)
        res = eng.scan_reads(pt, d_child, min_distinct=min_distinct_kmers_per_read, stats=stats,
                             want_hits=want_hits)
        nd, nh = res["ndistinct"], res["nhits"]
        if fetch:
            nd = nd.cpu().numpy().view(np.uint32)
            nh = nh.cpu().numpy().view(np.uint32)
            out["informative_reads"] = int((nd >= min_distinct_kmers_per_read).sum())
        out["ndistinct"], out["nhits"] = nd, nh
        if want_hits:
            out["hits"] = (res["hit_pos"], res["hit_slot"], pt)
        else:
            pt.close()
    st = eng.read_stats(stats)
    out["units"] = st["windows"]
    return out
