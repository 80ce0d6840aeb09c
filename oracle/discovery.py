"""Discovery-mode restatement (oracle; test infrastructure only).

Follows reference ``discovery/pipeline.py``:
  * Module 1  ``_extract_child_kmers_discovery`` :69-268 (count, ``dump -L``)
  * Module 1  ``_subtract_reference_kmers`` :271-319 (``query`` count == 0)
  * Module 2  ``_filter_parents_discovery`` :462-612 (``count --if`` + ``<= pmc``)
  * Module 3  ``_anchor_and_cluster`` :615-1153 with the per-contig worker of
    ``core/bam_scanner.py:340-507`` and ``_process_informative_read`` :284-337
  * Module 4  ``_annotate_and_link_from_metadata`` :1351-1489,
    ``_classify_regions`` :1517-1546, ``_write_bed`` :1156-1194,
    ``_write_bedgraph`` :1197-1278, ``_write_read_coverage_bed`` :1281-1348

All k-mers are Python ints in the 2-bit encoding of :mod:`oracle.kmers`.
Dedup order is BAM file order (SURVEY §8 parity traps; reproduces every
golden BED column).
"""

import bisect
import collections

from . import bam as obam
from . import kmers as okm


def child_candidates(child_records, k, min_child_count):
    """Module 1 step 1-2 → ``(set of candidate keys, dict all counts)``."""
    stream = obam.fasta_stream(child_records)
    counts = okm.count_sequences([r.seq for r in stream], k)
    cand = {key for key, c in counts.items() if c >= min_child_count}
    return cand, counts


def reference_kmers(ref_seqs, k):
    """Module 0: canonical k-mer set (with counts) of the reference FASTA."""
    return okm.count_sequences(ref_seqs, k)


def subtract_reference(candidates, ref_counts):
    """keep candidates whose reference count is 0 (``pipeline.py:302``)."""
    return {key for key in candidates if ref_counts.get(key, 0) == 0}


def parent_counts(parent_records, k, filter_keys):
    """``samtools fasta | jellyfish count --if`` → counts for filter keys."""
    stream = obam.fasta_stream(parent_records)
    codes, valid, _s, _l = okm.encode_stream([r.seq for r in stream])
    return okm.count_stream_filtered(codes, valid, k, filter_keys)


def filter_parents(non_ref, mother_records, father_records, k, parent_max_count=0):
    """Module 2 → ``(proband_unique set, after_mother set)``."""
    if not non_ref:
        return set(), set()
    mc = parent_counts(mother_records, k, non_ref)
    after_mother = {key for key in non_ref if mc[key] <= parent_max_count}
    if not after_mother:
        return set(), set()
    fc = parent_counts(father_records, k, after_mother)
    pu = {key for key in after_mother if fc[key] <= parent_max_count}
    return pu, after_mother


def scan_read_numeric(seq, k, pu):
    """Per-read reduction (``core/bam_scanner.py:434-443``): returns
    ``(set of distinct PU keys, sorted list of hit window starts)``."""
    codes, valid, _s, _l = okm.encode_stream([seq])
    hi, lo, ok = okm.canonical_windows(codes, valid, k)
    uniq = set()
    idx = []
    his = hi.tolist()
    los = lo.tolist()
    oks = ok.tolist()
    for i in range(len(oks)):
        if not oks[i]:
            continue
        key = (his[i] << 64) | los[i]
        if key in pu:
            uniq.add(key)
            idx.append(i)
    return uniq, idx


def kmer_ref_positions(rec, hit_indices, k):
    """``_collect_kmer_ref_positions`` (``core/bam_scanner.py:97-117``)."""
    cov = collections.Counter()
    q2r = dict(rec.get_aligned_pairs(matches_only=True))
    for s in hit_indices:
        for q in range(s, s + k):
            r = q2r.get(q)
            if r is not None:
                cov[r] += 1
    return cov


def anchor(child_records, k, pu, min_dk_per_read):
    """Module 3 scan + first-seen dedup in BAM file order.

    Returns dict with ``read_hits`` (chrom, start, end, qname, kmers, is_supp),
    ``per_read`` (list of (record index, n_distinct, hit indices) for every
    scanned record with a sequence), ``unmapped_informative``,
    ``read_sv_meta``, ``kmer_coverage``, ``read_coverage``, ``scanned``.
    """
    read_hits = []
    reads_seen = set()
    read_sv_meta = {}
    kmer_cov = collections.defaultdict(collections.Counter)
    read_cov = collections.defaultdict(collections.Counter)
    unmapped_inf = 0
    scanned = 0
    per_read = []
    for ridx, rec in enumerate(child_records):
        if rec.is_secondary or rec.is_duplicate:
            continue
        scanned += 1
        seq = rec.query_sequence
        if seq is None:
            continue
        uniq, idx = scan_read_numeric(seq, k, pu)
        per_read.append((ridx, len(uniq), idx))
        if len(uniq) < min_dk_per_read:
            continue
        key = (rec.qname, rec.is_supplementary)
        if key in reads_seen:
            continue
        reads_seen.add(key)
        if rec.is_unmapped:
            unmapped_inf += 1
            continue
        chrom = rec.reference_name
        read_hits.append((chrom, rec.reference_start, rec.reference_end,
                          rec.qname, uniq, rec.is_supplementary))
        cov = kmer_ref_positions(rec, idx, k)
        kmer_cov[chrom] += cov
        for pos in cov:
            read_cov[chrom][pos] += 1
        max_clip = 0
        for op, ln in rec.cigar or ():
            if op == 4 and ln > max_clip:
                max_clip = ln
        has_sa = rec.has_tag("SA")
        read_sv_meta[key] = {
            "has_sa": has_sa,
            "sa_str": rec.get_tag("SA") if (has_sa and not rec.is_supplementary) else None,
            "is_paired": rec.is_paired,
            "is_proper_pair": rec.is_proper_pair,
            "mate_is_unmapped": rec.mate_is_unmapped if rec.is_paired else False,
            "max_clip": max_clip,
        }
    return {
        "read_hits": read_hits, "per_read": per_read,
        "unmapped_informative": unmapped_inf, "read_sv_meta": read_sv_meta,
        "kmer_coverage": kmer_cov, "read_coverage": read_cov,
        "scanned": scanned,
    }


def cluster(read_hits, merge_distance):
    """Greedy interval clustering (``pipeline.py:1107-1153``)."""
    if not read_hits:
        return [], {}, {}
    hits = sorted(read_hits, key=lambda x: (x[0], x[1]))
    regions = []
    region_reads = {}
    region_kmers = {}
    cc, cs, ce = hits[0][0], hits[0][1], hits[0][2]
    names = {hits[0][3]}
    kms = set(hits[0][4])
    for chrom, start, end, name, uniq, _supp in hits[1:]:
        if chrom == cc and start <= ce + merge_distance:
            ce = max(ce, end)
            names.add(name)
            kms.update(uniq)
        else:
            key = (cc, cs, ce)
            regions.append(key)
            region_reads[key] = names
            region_kmers[key] = kms
            cc, cs, ce = chrom, start, end
            names = {name}
            kms = set(uniq)
    key = (cc, cs, ce)
    regions.append(key)
    region_reads[key] = names
    region_kmers[key] = kms
    return regions, region_reads, region_kmers


def annotate_and_link(regions, region_reads, read_sv_meta):
    """``_annotate_and_link_from_metadata`` (``pipeline.py:1351-1489``)."""
    read_to_regions = {}
    for rk in regions:
        for q in region_reads.get(rk, ()):
            read_to_regions.setdefault(q, set()).add(rk)
    ann = {r: {"split_reads": 0, "discordant_pairs": 0, "max_clip_len": 0,
               "unmapped_mates": 0} for r in regions}
    if not read_to_regions:
        return ann, []
    counted = set()
    for dkey, meta in read_sv_meta.items():
        q = dkey[0]
        if q not in read_to_regions:
            continue
        for rk in read_to_regions[q]:
            a = ann[rk]
            if meta["has_sa"] and (q, rk) not in counted:
                a["split_reads"] += 1
                counted.add((q, rk))
            if meta["is_paired"]:
                if meta["mate_is_unmapped"]:
                    a["unmapped_mates"] += 1
                elif not meta["is_proper_pair"]:
                    a["discordant_pairs"] += 1
            if meta["max_clip"] > a["max_clip_len"]:
                a["max_clip_len"] = meta["max_clip"]
    by_chrom = {}
    for r in regions:
        by_chrom.setdefault(r[0], []).append(r)
    starts = {}
    for c, lst in by_chrom.items():
        lst.sort(key=lambda x: x[1])
        starts[c] = [r[1] for r in lst]
    bridges = {}
    for dkey, meta in read_sv_meta.items():
        q = dkey[0]
        sa = meta.get("sa_str")
        if not sa or q not in read_to_regions:
            continue
        for entry in sa.rstrip(";").split(";"):
            parts = entry.split(",")
            if len(parts) < 3:
                continue
            try:
                pos = int(parts[1]) - 1
            except ValueError:
                continue
            if parts[0] not in starts:
                continue
            i = bisect.bisect_right(starts[parts[0]], pos) - 1
            if i >= 0:
                t = by_chrom[parts[0]][i]
                if t[1] <= pos < t[2]:
                    for p in read_to_regions[q]:
                        if p != t:
                            bridges.setdefault(tuple(sorted([p, t])), set()).add(q)
    for q, rset in read_to_regions.items():
        if len(rset) >= 2:
            rl = sorted(rset)
            for i in range(len(rl)):
                for j in range(i + 1, len(rl)):
                    bridges.setdefault((rl[i], rl[j]), set()).add(q)
    links = []
    for a, b in sorted(bridges):
        links.append({"region_a": a, "region_b": b,
                      "supporting_reads": bridges[(a, b)],
                      "sv_type_hint": "BND" if a[0] != b[0] else "INTRA"})
    return ann, links


def classify(regions, ann, links):
    """``_classify_regions`` (``pipeline.py:1517-1546``)."""
    linked = set()
    for l in links:
        linked.add(l["region_a"])
        linked.add(l["region_b"])
    for rk in regions:
        a = ann.get(rk, {})
        s, d, u = a.get("split_reads", 0), a.get("discordant_pairs", 0), a.get("unmapped_mates", 0)
        if s >= 2 or d >= 2 or u >= 2 or rk in linked:
            a["class"] = "SV"
        elif s == 0 and d == 0 and u == 0:
            a["class"] = "SMALL"
        else:
            a["class"] = "AMBIGUOUS"
        ann[rk] = a


def bed_rows(regions, region_reads, region_kmers, ann):
    """Data rows of ``_write_bed`` as tuples."""
    rows = []
    for c, s, e in regions:
        a = ann.get((c, s, e), {})
        rows.append((c, s, e, len(region_reads[(c, s, e)]), len(region_kmers[(c, s, e)]),
                     a.get("split_reads", 0), a.get("discordant_pairs", 0),
                     a.get("max_clip_len", 0), a.get("unmapped_mates", 0),
                     a.get("class", "SMALL")))
    return rows


def bedgraph_rows(kmer_cov, read_cov, min_reads):
    """Data rows of ``_write_bedgraph`` (``pipeline.py:1197-1278``)."""
    rows = []
    for chrom in sorted(kmer_cov):
        positions = kmer_cov[chrom]
        if not positions:
            continue
        rc = read_cov.get(chrom, {}) if read_cov else None
        rs = rv = re_ = None
        for pos in sorted(positions):
            if rc is not None and rc.get(pos, 0) < min_reads:
                if rs is not None:
                    rows.append((chrom, rs, re_, rv))
                    rs = None
                continue
            val = positions[pos]
            if rs is None:
                rs, rv, re_ = pos, val, pos + 1
            elif pos == re_ and val == rv:
                re_ = pos + 1
            else:
                rows.append((chrom, rs, re_, rv))
                rs, rv, re_ = pos, val, pos + 1
        if rs is not None:
            rows.append((chrom, rs, re_, rv))
    return rows


def read_coverage_rows(kmer_cov, read_cov, min_reads):
    """Data rows of ``_write_read_coverage_bed`` (``pipeline.py:1281-1348``)."""
    rows = []
    for chrom in sorted(read_cov):
        rc = read_cov[chrom]
        kc = kmer_cov.get(chrom, {})
        filt = {}
        for pos, n in rc.items():
            if n >= min_reads:
                filt[pos] = (n, round(kc.get(pos, 0) / n, 1))
        if not filt:
            continue
        sp = sorted(filt)
        rs = sp[0]
        rv = filt[rs]
        re_ = rs + 1
        for pos in sp[1:]:
            v = filt[pos]
            if pos == re_ and v == rv:
                re_ = pos + 1
            else:
                rows.append((chrom, rs, re_, rv[0], rv[1]))
                rs, rv, re_ = pos, v, pos + 1
        rows.append((chrom, rs, re_, rv[0], rv[1]))
    return rows


def run(child_records, mother_records, father_records, ref_seqs, k,
        min_child_count=3, parent_max_count=0, min_dk_per_read=None,
        merge_distance=500, min_supporting_reads=1, min_distinct_kmers=1,
        min_bedgraph_reads=3, ref_counts=None):
    """Whole discovery path on in-memory records → dict of results."""
    if min_dk_per_read is None:
        min_dk_per_read = max(1, k // 4)
    cand, child_counts = child_candidates(child_records, k, min_child_count)
    if ref_counts is None:
        ref_counts = reference_kmers(ref_seqs, k)
    non_ref = subtract_reference(cand, ref_counts)
    pu, after_mother = filter_parents(non_ref, mother_records, father_records, k,
                                      parent_max_count)
    res = {"candidates": cand, "non_ref": non_ref, "after_mother": after_mother,
           "proband_unique": pu, "child_counts": child_counts}
    if not pu:
        res.update({"regions": [], "bed": [], "informative": 0,
                    "unmapped_informative": 0, "per_read": []})
        return res
    a = anchor(child_records, k, pu, min_dk_per_read)
    regions, rreads, rkmers = cluster(a["read_hits"], merge_distance)
    if min_supporting_reads > 1 or min_distinct_kmers > 1:
        regions = [r for r in regions
                   if len(rreads[r]) >= min_supporting_reads
                   and len(rkmers[r]) >= min_distinct_kmers]
    ann, links = annotate_and_link(regions, rreads, a["read_sv_meta"])
    classify(regions, ann, links)
    res.update({
        "regions": regions, "region_reads": rreads, "region_kmers": rkmers,
        "annotations": ann, "links": links,
        "bed": bed_rows(regions, rreads, rkmers, ann),
        "bedgraph": bedgraph_rows(a["kmer_coverage"], a["read_coverage"], min_bedgraph_reads),
        "read_coverage_bed": read_coverage_rows(a["kmer_coverage"], a["read_coverage"], min_bedgraph_reads),
        "informative": len(a["read_hits"]) + a["unmapped_informative"],
        "unmapped_informative": a["unmapped_informative"],
        "per_read": a["per_read"], "scanned": a["scanned"],
    })
    return res
