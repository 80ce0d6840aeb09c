"""``kmer-discovery`` across the GPUs of one node: one process per GPU
(``torchrun --nproc-per-node N -m kmer_denovo_filter_b200.cli --out-prefix ...``).

The reference has no multi-device mode; this is how its pipeline
(``discovery/pipeline.py:2093-2548``) shards:

* rank r decodes range r of every BAM (contiguous BGZF ranges cut at ``.bai`` linear-index
  entries, ``bamio.open_shard``; the child in scan mode with metadata, the parents as the
  ``samtools fasta`` stream) and one slice of the reference;
* the k-mer chain is :func:`kmer_chain_dist.discover_streams_dist` — child table
  partitioned by owner rank, candidates replicated, parent counts all-reduced — so the
  proband-unique set is identical on every rank;
* every rank anchors its own reads; the ``(query_name, is_supplementary)`` keys of the
  informative reads are exchanged first, so that a read is dropped as a duplicate exactly
  when an earlier record of the file (possibly another rank's) carries the same key;
* rank 0 gathers the per-read tuples, SV metadata, coverage and the informative records,
  clusters and writes the same files as the single-GPU run, byte for byte.
"""

import collections
import logging
import os
import time

import numpy as np

from .. import bamio
from .. import engine as _engine
from ..core import kmer_engine_wrappers as kw
from . import kmer_chain_dist
from . import pipeline as P

logger = logging.getLogger(__name__)


def _dist():
    import torch.distributed as dist
    return dist


def reference_slice(ref_fasta, k, rank, world):
    """Piece ``rank`` of every reference sequence (cut into ``world`` pieces that overlap by
    k - 1 bases, so that every k-mer of the reference lies in exactly one or two pieces —
    marking is idempotent) as a packed HostStream."""
    _names, seqs = bamio.read_fasta_sequences(ref_fasta)
    pieces = []
    for s in seqs:
        n = len(s)
        lo = n * rank // world
        hi = min(n, n * (rank + 1) // world + k - 1)
        if hi > lo:
            pieces.append(s[lo:hi])
    return _engine.pack_sequences(pieces)


def _gather(obj, world):
    out = [None] * world
    _dist().all_gather_object(out, obj)
    return out


def run_discovery_pipeline_dist(args, eng, paths, min_dk, min_bedgraph_reads, finish_empty, lap):
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    k = args.kmer_size
    threads = max(1, args.threads)
    if not args.ref_fasta:
        raise _engine.KdfError("the multi-GPU run needs --ref-fasta (every rank packs one slice of it)")

    def decode(path, mode, want_meta):
        rd = bamio.open_shard(path, rank, world, threads=threads)
        if getattr(rd, "empty_shard", False):
            return rd, []
        return rd, list(rd.batches(mode, max_bases=kw.BATCH_BASES, want_meta=want_meta))

    t0 = time.monotonic()
    child_rd, child_batches = decode(args.child, bamio.MODE_SCAN, True)
    mother_rd, mother_batches = decode(args.mother, bamio.MODE_FASTA, False)
    father_rd, father_batches = decode(args.father, bamio.MODE_FASTA, False)
    ref_hs = reference_slice(args.ref_fasta, k, rank, world)
    lap("decode_shards_s")
    logger.info("[rank %d/%d] decoded %d child / %d mother / %d father reads (%.1fs)", rank, world,
                sum(b.n_reads for b in child_batches), sum(b.n_reads for b in mother_batches),
                sum(b.n_reads for b in father_batches), time.monotonic() - t0)
    empty = _engine.pack_sequences([])
    res = kmer_chain_dist.discover_streams_dist(
        eng,
        [bamio.counting_view(b) for b in child_batches] or [empty],
        mother_batches or [empty], father_batches or [empty], ref_hs, k,
        min_child_count=args.min_child_count, parent_max_count=args.parent_max_count,
        min_distinct_kmers_per_read=min_dk, child_scan=child_batches or [empty])
    for b in mother_batches + father_batches:
        b.close()
    mother_rd.close()
    father_rd.close()
    lap("kmer_chain_s")
    n_candidates, n_non_ref, n_pu = res["candidates"], res["non_ref"], res["proband_unique"]
    if n_pu == 0:
        child_rd.close()
        return finish_empty(n_candidates, n_non_ref) if rank == 0 else None

    # ---- anchoring: every rank its own reads, duplicates resolved in file (= rank) order
    items = [P.sparse_to_scan_item(b, sp, min_dk) for b, sp in zip(child_batches, res.get("reads_parts") or [])]
    inf_keys, hit_keys = [], []
    for batch, nd, nh, _ridx, _off, _slot in items:
        for r in np.flatnonzero(nd >= max(1, min_dk)).tolist():
            rec = batch.record(r)
            inf_keys.append((rec.query_name, rec.is_supplementary))
        for r in np.flatnonzero(nh > 0).tolist():
            rec = batch.record(r)
            hit_keys.append((rec.query_name, rec.is_supplementary, int(batch.rec_uoff[r])))
    all_inf = _gather(inf_keys, world)
    all_hit = _gather([(a, b) for a, b, _u in hit_keys], world)
    preseen = set()
    for r in range(rank):
        preseen.update(all_inf[r])
    col = P.AnchorCollector(eng, k, min_dk, preseen=preseen)
    for item in items:
        col.add(*item[:5])
    # informative-reads BAM: reads with >= 1 hit, first record per key in file order
    seen = set()
    for r in range(rank):
        seen.update(all_hit[r])
    uoffs = []
    for qn, supp, uoff in hit_keys:
        if (qn, supp) in seen:
            continue
        seen.add((qn, supp))
        uoffs.append(uoff)
    raws = child_rd.fetch_records(uoffs) if uoffs else []
    header = (child_rd.header_text, child_rd.references, child_rd.lengths)
    for b in child_batches:
        b.close()
    child_rd.close()
    lap("anchor_s")
    part = {"read_hits": col.read_hits, "sv": col.read_sv_meta,
            "kcov": {c: dict(v) for c, v in col.kmer_coverage.items()},
            "rcov": {c: dict(v) for c, v in col.read_coverage.items()},
            "unmapped": col.unmapped_informative, "scanned": col.total_scanned, "raws": raws}
    parts = _gather(part, world)
    lap("gather_s")
    if rank != 0:
        return None
    read_hits, read_sv_meta, raws = [], {}, []
    kmer_coverage = collections.defaultdict(collections.Counter)
    read_coverage = collections.defaultdict(collections.Counter)
    unmapped = scanned = 0
    for pt in parts:
        read_hits += pt["read_hits"]
        read_sv_meta.update(pt["sv"])
        for c, v in pt["kcov"].items():
            kmer_coverage[c].update(v)
        for c, v in pt["rcov"].items():
            read_coverage[c].update(v)
        unmapped += pt["unmapped"]
        scanned += pt["scanned"]
        raws += pt["raws"]
    total_informative = len(read_hits) + unmapped
    logger.info("Anchoring complete: %d informative reads (%d mapped, %d unmapped) from %d scanned",
                total_informative, len(read_hits), unmapped, scanned)
    regions, region_reads, region_kmers = P._cluster_hits(read_hits, args.cluster_distance) \
        if read_hits else ([], {}, {})
    n = bamio.write_sorted_bam(paths["info_bam"], header[0], header[1], header[2],
                               [bamio.append_int_tag(raw, "dk", 1) for raw in raws])
    logger.info("Informative reads BAM written: %s (%d reads)", paths["info_bam"], n)
    lap("informative_bam_s")
    return P._finish_discovery(args, regions, region_reads, region_kmers, read_sv_meta, kmer_coverage,
                               read_coverage, total_informative, unmapped, n_candidates, n_non_ref, n_pu,
                               min_dk, min_bedgraph_reads, paths, lap)
