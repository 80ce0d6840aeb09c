"""bench_wall.py — the wall-time half of BASELINE.json's metric ("trio discovery wall time").

Writes the synthetic trio of bench.py as coordinate-sorted BAMs + a reference FASTA (what
a user of the reference would hand to ``kmer-discovery``), runs the product pipeline
(``run_discovery_pipeline``: BAM decode -> GPU k-mer chain -> clustering -> BED / bedGraph /
informative-reads BAM / metrics) on them and reports the wall time with its stages.
Bench / test infrastructure: nothing in the product imports this module.

Read alignment is by construction: a read's position is where it was sampled on its
haplotype, mapped back to reference coordinates across the de novo indels; every read gets
CIGAR ``<L>M`` (reads over an indel are therefore placed as a local aligner would place
their longer flank — the pipeline never re-aligns, it only projects k-mer hits through the
CIGAR).
"""
import argparse
import os
import shutil
import tempfile
import time

import numpy as np

CONTIG_BP = 64_000_000          # one contig per 64 Mbp
_NIB = np.array([1, 2, 4, 8, 15], dtype=np.uint8)      # A C G T N -> BAM nibble


def reg2bin_vec(beg, end):
    """BAM bin of [beg, end) (SAM spec §5.3), vectorised."""
    end = end - 1
    out = np.zeros(beg.shape[0], dtype=np.int64)
    done = np.zeros(beg.shape[0], dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        m = ~done & ((beg >> shift) == (end >> shift))
        out[m] = base + (beg[m] >> shift)
        done |= m
    return out.astype(np.uint16)


def hap_to_ref_map(events):
    """Piecewise shift from the coordinates of a haplotype carrying ``events`` (synth.apply_denovo)
    back to the coordinates of the haplotype it was derived from."""
    new_break, delta = [0], [0]
    cum = 0
    for pos, kind, ref, alt in sorted(events):
        if kind == "ins":
            ln = len(alt) - 1
            new_break.append(pos + 1 + cum + ln)
            cum += ln
            delta.append(cum)
        elif kind == "del":
            ln = len(ref) - 1
            new_break.append(pos + 1 + cum)
            cum -= ln
            delta.append(cum)
    nb, dl = np.asarray(new_break, dtype=np.int64), np.asarray(delta, dtype=np.int64)

    def to_ref(x):
        j = np.searchsorted(nb, x, side="right") - 1
        return np.maximum(x - dl[j], 0)
    return to_ref


def write_reads_bam(path, contigs, table, read_len, to_ref=None, threads=None, level=1, sample="S"):
    """``table`` (synth.make_reads(table=...)) as a coordinate-sorted BAM.  ``contigs``:
    [(name, length)].  ``to_ref``: {haplotype index: map to reference coordinates}."""
    from kmer_denovo_filter_b200 import bamio
    L = int(read_len)
    bases = table["bases"]                       # (2m, L) as sequenced; N = 4
    m = bases.shape[0] // 2
    start = table["start"].astype(np.int64)
    ins = table["insert"].astype(np.int64)
    hap = table["hap"]
    p1 = start.copy()                            # leftmost haplotype coordinate of read 1
    p2 = start + ins - L                         # ... of read 2 (sequenced from the reverse strand)
    if to_ref:
        for h, fn in to_ref.items():
            sel = hap == h
            p1[sel] = fn(p1[sel])
            p2[sel] = fn(p2[sel])
    # forward-strand sequence of read 2: reverse complement of what was sequenced (N stays N)
    fwd = bases.copy()
    r2 = bases[1::2, ::-1]
    fwd[1::2] = np.where(r2 < 4, 3 - r2, 4)
    pos = np.empty(2 * m, dtype=np.int64)
    pos[0::2], pos[1::2] = p1, p2
    mate = np.empty(2 * m, dtype=np.int64)
    mate[0::2], mate[1::2] = p2, p1
    tlen = np.empty(2 * m, dtype=np.int64)
    span = p2 + L - p1
    tlen[0::2], tlen[1::2] = span, -span
    flag = np.empty(2 * m, dtype=np.uint16)
    flag[0::2], flag[1::2] = 99, 147            # paired, proper, mate reverse / reverse, first / second
    pair = np.repeat(np.arange(m, dtype=np.int64), 2)
    off = np.cumsum([0] + [ln for _n, ln in contigs])
    tid = np.clip(np.searchsorted(off, pos, side="right") - 1, 0, len(contigs) - 1)
    mtid = np.clip(np.searchsorted(off, mate, side="right") - 1, 0, len(contigs) - 1)
    cpos = pos - off[tid]
    order = np.lexsort((flag & 16, cpos, tid))
    name_len = 12                                # "r%010d" + NUL
    dt = np.dtype([("block_size", "<i4"), ("refID", "<i4"), ("pos", "<i4"), ("l_read_name", "u1"),
                   ("mapq", "u1"), ("bin", "<u2"), ("n_cigar_op", "<u2"), ("flag", "<u2"),
                   ("l_seq", "<i4"), ("next_refID", "<i4"), ("next_pos", "<i4"), ("tlen", "<i4"),
                   ("read_name", "S%d" % name_len), ("cigar", "<u4"),
                   ("seq", "u1", ((L + 1) // 2,)), ("qual", "u1", (L,))])
    rec = np.zeros(2 * m, dtype=dt)
    rec["block_size"] = dt.itemsize - 4
    rec["refID"] = tid[order]
    rec["pos"] = cpos[order]
    rec["l_read_name"] = name_len
    rec["mapq"] = 60
    rec["bin"] = reg2bin_vec(cpos[order], cpos[order] + L)
    rec["n_cigar_op"] = 1
    rec["flag"] = flag[order]
    rec["l_seq"] = L
    rec["next_refID"] = mtid[order]
    rec["next_pos"] = (mate - off[mtid])[order]
    rec["tlen"] = tlen[order]
    nm = np.zeros((2 * m, name_len), dtype=np.uint8)      # "r%010d\0", built digit by digit
    nm[:, 0] = ord("r")
    x = pair[order].copy()
    for d in range(10, 0, -1):
        nm[:, d] = 48 + (x % 10)
        x //= 10
    rec["read_name"] = nm.view("S%d" % name_len).reshape(-1)
    rec["cigar"] = (L << 4) | 0
    nib = _NIB[fwd[order]]
    if L & 1:
        nib = np.concatenate([nib, np.zeros((2 * m, 1), dtype=np.uint8)], axis=1)
    rec["seq"] = (nib[:, 0::2] << 4) | nib[:, 1::2]
    rec["qual"] = 30
    text = ("@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in contigs) +
            "@RG\tID:%s\tSM:%s\n" % (sample, sample)).encode()
    head = bytearray(b"BAM\x01" + np.int32(len(text)).tobytes() + text + np.int32(len(contigs)).tobytes())
    for n, ln in contigs:
        nb = n.encode() + b"\0"
        head += np.int32(len(nb)).tobytes() + nb + np.int32(ln).tobytes()
    data = np.concatenate([np.frombuffer(bytes(head), dtype=np.uint8), rec.view(np.uint8).reshape(-1)])
    coff = bamio.bgzf_write(path, data, level=level, threads=threads)
    write_bai_fixed(path + ".bai", len(contigs), len(head), dt.itemsize, rec["refID"], rec["pos"].astype(np.int64),
                    L, rec["bin"].astype(np.int64), coff, int(data.shape[0]))
    return {"reads": int(2 * m), "bam_bytes": os.path.getsize(path), "uncompressed_bytes": int(data.shape[0])}


def write_bai_fixed(path, n_ref, head_len, rec_size, tid, pos, read_len, bins, coff, total_len):
    """The .bai of a BAM whose records all have ``rec_size`` bytes (and ``read_len`` aligned
    bases), sorted by (tid, pos): bins with their chunks and the 16 kbp linear index, from
    arithmetic on the record number (record i starts at head_len + i * rec_size of the
    uncompressed stream; BGZF blocks hold 0xff00 bytes)."""
    import struct
    BLK = 0xff00

    def voff(u):
        u = np.asarray(u, dtype=np.int64)
        b = u // BLK
        return (coff[np.minimum(b, coff.shape[0] - 1)].astype(np.uint64) << np.uint64(16)) | \
            (u - b * BLK).astype(np.uint64)
    n = tid.shape[0]
    ustart = head_len + np.arange(n + 1, dtype=np.int64) * rec_size
    ustart[-1] = total_len
    v = voff(ustart)
    if total_len % BLK == 0:          # the end of the data is the start of the EOF block
        v[-1] = np.uint64(int(coff[-1]) << 16)
    out = bytearray(b"BAI\x01" + struct.pack("<i", n_ref))
    bounds = np.searchsorted(tid, np.arange(n_ref + 1))
    for r in range(n_ref):
        a, b = int(bounds[r]), int(bounds[r + 1])
        if a == b:
            out += struct.pack("<ii", 0, 0)
            continue
        bb = bins[a:b]
        run_start = np.flatnonzero(np.concatenate(([True], bb[1:] != bb[:-1])))
        run_end = np.concatenate((run_start[1:], [b - a]))
        run_bin = bb[run_start]
        order = np.argsort(run_bin, kind="stable")
        ub, first = np.unique(run_bin[order], return_index=True)
        cnt = np.diff(np.concatenate((first, [order.shape[0]])))
        out += struct.pack("<i", ub.shape[0])
        for j in range(ub.shape[0]):
            sel = order[first[j]:first[j] + cnt[j]]
            out += struct.pack("<Ii", int(ub[j]), int(sel.shape[0]))
            ch = np.stack([v[a + run_start[sel]], v[a + run_end[sel]]], axis=1).astype("<u8")
            out += ch.tobytes()
        p = pos[a:b]
        n_intv = int((p[-1] + read_len - 1) >> 14) + 1
        first_read = np.searchsorted(p + read_len, np.arange(n_intv, dtype=np.int64) << 14, side="right")
        lin = v[a + np.minimum(first_read, b - a - 1)].astype("<u8")
        out += struct.pack("<i", n_intv) + lin.tobytes()
    out += struct.pack("<Q", 0)
    with open(path, "wb") as fh:
        fh.write(out)


def write_fasta(path, contigs, ref_codes):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as fh:
        o = 0
        for name, ln in contigs:
            fh.write((">%s\n" % name).encode())
            seq = lut[ref_codes[o:o + ln]]
            width = 100
            full = (ln // width) * width
            if full:
                lines = np.concatenate([seq[:full].reshape(-1, width),
                                        np.full((full // width, 1), 10, dtype=np.uint8)], axis=1)
                fh.write(lines.tobytes())
            if ln > full:
                fh.write(seq[full:].tobytes() + b"\n")
            o += ln


def contigs_for(genome_bp):
    out, o, i = [], 0, 1
    while o < genome_bp:
        ln = min(CONTIG_BP, genome_bp - o)
        out.append(("chr%d" % i, int(ln)))
        o += ln
        i += 1
    return out


def make_bam_trio(torch, dev, genome_bp, depth, read_len, n_denovo, outdir, threads=None):
    """The synthetic trio of bench.py (same generator and seeds) as BAMs + FASTA.
    → (paths dict, events in reference coordinates, stats)."""
    from kmer_denovo_filter_b200 import synth
    t0 = time.perf_counter()
    ref = synth.make_reference(torch, dev, genome_bp)
    m0, m1 = synth.make_haplotype(torch, ref, 2000), synth.make_haplotype(torch, ref, 2001)
    f0, f1 = synth.make_haplotype(torch, ref, 3000), synth.make_haplotype(torch, ref, 3001)
    child_a, events = synth.apply_denovo(torch, m0, n_denovo, 4000)
    pairs = int(depth * genome_bp / (2 * read_len))
    contigs = contigs_for(genome_bp)
    paths = {"ref": os.path.join(outdir, "ref.fa")}
    write_fasta(paths["ref"], contigs, ref.cpu().numpy())
    stats = {}
    for who, haps, seed, maps in (("child", [child_a, f1], 5000, {0: hap_to_ref_map(events)}),
                                  ("mother", [m0, m1], 5001, None), ("father", [f0, f1], 5002, None)):
        table = {}
        synth.make_reads(torch, haps, pairs, read_len, seed, table=table)
        paths[who] = os.path.join(outdir, who + ".bam")
        stats[who] = write_reads_bam(paths[who], contigs, table, read_len, maps, threads, sample=who)
        del table
    del ref, m0, m1, f0, f1, child_a
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    stats["seconds"] = time.perf_counter() - t0
    ev = []
    off = np.cumsum([0] + [ln for _n, ln in contigs])
    for pos, kind, r, a in events:
        c = int(np.searchsorted(off, pos, side="right") - 1)
        ev.append((contigs[c][0], int(pos - off[c]), kind, r, a))
    return paths, ev, stats


def write_truth_vcf(path, contigs, events, sample="child"):
    """The injected de novo events as the candidate VCF of VCF mode (config 3)."""
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n")
        for n, ln in contigs:
            fh.write("##contig=<ID=%s,length=%d>\n" % (n, ln))
        fh.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n" % sample)
        order = {n: i for i, (n, _l) in enumerate(contigs)}
        for chrom, pos, _kind, ref, alt in sorted(events, key=lambda e: (order[e[0]], e[1])):
            fh.write("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT\t0/1\n" % (chrom, pos + 1, ref, alt))


def vcf_args(paths, vcf_path, out_dir, k=31, threads=None):
    return argparse.Namespace(
        child=paths["child"], mother=paths["mother"], father=paths["father"], ref_fasta=paths["ref"],
        vcf=vcf_path, output=os.path.join(out_dir, "annotated.vcf.gz"),
        metrics=os.path.join(out_dir, "vcf_metrics.json"), summary=os.path.join(out_dir, "vcf_summary.txt"),
        informative_reads=os.path.join(out_dir, "vcf_informative.bam"), kmer_size=k, min_baseq=20,
        min_mapq=20, proband_id="child", threads=threads or (os.cpu_count() or 4), memory=None,
        debug_kmers=False, jf_hash_size=None, tmp_dir=None, kraken2_db=None, report=None)


def discovery_args(paths, out_prefix, k=31, threads=None):
    return argparse.Namespace(
        child=paths["child"], mother=paths["mother"], father=paths["father"], ref_fasta=paths["ref"],
        ref_jf=None, out_prefix=out_prefix, kmer_size=k, min_child_count=3, parent_max_count=0,
        threads=threads or (os.cpu_count() or 4), memory=None, debug_kmers=False, jf_hash_size=None,
        tmp_dir=None, min_baseq=20, candidate_summary=None, cluster_distance=500,
        min_supporting_reads=1, min_distinct_kmers=1, min_bedgraph_reads=3,
        min_distinct_kmers_per_read=None, sv_bedpe=None, report=None)


def events_detected(bed_path, events, slack=200):
    regions = []
    for line in open(bed_path):
        if line.startswith("#"):
            continue
        f = line.split("\t")
        regions.append((f[0], int(f[1]), int(f[2])))
    hit = 0
    for chrom, pos, _k, _r, _a in events:
        if any(c == chrom and s - slack <= pos <= e + slack for c, s, e in regions):
            hit += 1
    return hit, len(regions)


def discovery_wall(args, eng, rank=0, world=1):
    """bench.py's `discovery_wall` object: BAM trio -> candidate BED through the product CLI path.
    With several ranks the SAME trio (not a larger one) is run by the multi-GPU pipeline: every
    rank decodes 1 / world of each BAM."""
    import torch
    from kmer_denovo_filter_b200.discovery import pipeline
    dist = None
    if world > 1:
        import torch.distributed as dist
    genome_bp = int((args.wall_mbp or args.genome_mbp) * 1e6)
    threads = max(2, (os.cpu_count() or 4) // max(world, 1))
    box = [None]
    if rank == 0:
        tmp = tempfile.mkdtemp(prefix="kdf_wall_", dir=os.environ.get("KDF_WALL_TMP"))
        paths, events, gen = make_bam_trio(torch, eng.device, genome_bp, args.depth, args.read_len,
                                           args.denovo, tmp, os.cpu_count() or 4)
        box[0] = (tmp, paths)
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    tmp, paths = box[0]
    try:
        pargs = discovery_args(paths, os.path.join(tmp, "out"), args.k, threads)
        runs = []
        for _ in range(2):           # the first run pays page-cache and allocator warm-up: report the second
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            metrics = pipeline.run_discovery_pipeline(pargs, engine=eng)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            runs.append((time.perf_counter() - t0, dict(pipeline.LAST_TIMINGS)))
        if rank != 0:
            return None
        wall, stages = runs[-1]
        hit, n_regions = events_detected(os.path.join(tmp, "out.bed"), events)
        vcf_mode = None
        if world == 1:
            # VCF mode on the same BAMs (BASELINE config 3): the 100 injected events as candidates
            from kmer_denovo_filter_b200.vcf import pipeline as vpipe
            vcf_path = os.path.join(tmp, "truth.vcf")
            write_truth_vcf(vcf_path, contigs_for(genome_bp), events)
            vruns = []
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                vres = vpipe.run_pipeline(vcf_args(paths, vcf_path, tmp, args.k, threads), engine=eng)
                torch.cuda.synchronize()
                vruns.append(time.perf_counter() - t0)
            ann = vres["annotations"]
            vcf_mode = {"wall_s": vruns[-1], "first_run_wall_s": vruns[0], "variants": len(ann),
                        "variants_with_DKU": sum(1 for a in ann.values() if a["dku"] > 0),
                        "variants_with_DKA": sum(1 for a in ann.values() if a["dka"] > 0),
                        "total_child_kmers": int(vres["metrics"]["total_child_kmers"]),
                        "child_unique_kmers": int(vres["metrics"]["child_unique_kmers"]),
                        "api": "kmer_denovo_filter_b200.vcf.pipeline.run_pipeline (the kmer-denovo CLI entry)"}
        reads = sum(gen[w]["reads"] for w in ("child", "mother", "father"))
        return {
            "genome_bp": genome_bp, "depth": args.depth, "k": args.k, "n_gpus": world,
            "note": "the same 64 Mbp trio at every N (strong scaling of the product pipeline)" if world > 1 else None,
            "input": {"reads": reads, "bam_bytes": sum(gen[w]["bam_bytes"] for w in ("child", "mother", "father")),
                      "bam_level": 1, "bam_write_seconds": gen["seconds"]},
            "wall_s": wall, "first_run_wall_s": runs[0][0], "host_threads_per_rank": threads,
            "stages_s": {k: round(v, 4) for k, v in stages.items()},
            "reads_per_s": reads / wall,
            "outputs": {"candidate_regions": n_regions, "de_novo_events": len(events),
                        "events_inside_a_region": hit,
                        "proband_unique_kmers": int(metrics.get("proband_unique_kmers", 0)),
                        "informative_reads": int(metrics.get("informative_reads", 0))},
            "api": "kmer_denovo_filter_b200.discovery.pipeline.run_discovery_pipeline (the kmer-discovery CLI entry)",
            "vcf_mode": vcf_mode,
        }
    finally:
        if world > 1:
            dist.barrier()
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
