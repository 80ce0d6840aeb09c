"""VCF-mode pipeline (``kmer-denovo``) on the GPU k-mer engine.

Function names, arguments, return values and output files follow the
reference's ``vcf/pipeline.py``; the two whole-genome parent scans — the hot
loop of this mode (``samtools fasta | jellyfish count --if`` + ``jellyfish
dump``, reference ``core/jellyfish_wrappers.py:115-283``) — run as sm_100a
kernels through ``libkdf_sm100.so``.  The child side touches only the reads
over candidate sites and stays on the CPU, as does the annotation arithmetic.

Reference map (``vcf/pipeline.py``):
  _collect_child_kmers :619      _select_alt_from_gt :730
  _parse_vcf_variants :747       _write_annotated_vcf :813
  _write_summary :1360           run_pipeline :1454 (parent scans :1587-1609,
                                 annotate :1662-1728, metrics :1925-1951)
  _write_informative_reads :1307 (DV:Z-tagged BAM of the informative reads)
Not restated: Kraken2 contamination fractions and the HTML report (SURVEY §2 rows 12, 16,
17 — out of scope; their flags are refused or warned about, see cli.py).
"""

import collections
import gzip
import json
import logging
import os
import statistics
import struct
import sys
import tempfile
import time
import zlib

import numpy as np

from .. import bamio
from ..core.kmer_engine_wrappers import _scan_parent_jellyfish, get_engine
from ..kmer_utils import _is_symbolic, extract_variant_spanning_kmers, read_supports_alt

logger = logging.getLogger(__name__)

_FIELD_DEFS = [
    ("DKU", "Integer", "Number of child fragments (unique read names) with at least one "
                       "variant-spanning k-mer unique to child (absent from both parents)"),
    ("DKT", "Integer", "Total child fragments (unique read names) with variant-spanning k-mers"),
    ("DKA", "Integer", "Number of child fragments (unique read names) with at least one "
                       "unique k-mer that also exactly supports the candidate allele"),
    ("DKU_DKT", "Float", "Proportion of child fragments with unique k-mers (DKU/DKT)"),
    ("DKA_DKT", "Float", "Proportion of child fragments with unique allele-supporting k-mers (DKA/DKT)"),
    ("MAX_PKC", "Integer", "Maximum k-mer count in parents for variant-spanning k-mers"),
    ("AVG_PKC", "Float", "Average k-mer count in parents for variant-spanning k-mers found in parents"),
    ("MIN_PKC", "Integer", "Minimum k-mer count in parents for variant-spanning k-mers"),
    ("MAX_PKC_ALT", "Integer", "Maximum k-mer count in parents for alt-allele-supporting k-mers"),
    ("AVG_PKC_ALT", "Float", "Average k-mer count in parents for alt-allele-supporting k-mers found in parents"),
    ("MIN_PKC_ALT", "Integer", "Minimum k-mer count in parents for alt-allele-supporting k-mers"),
]
_ANN_KEYS = ["dku", "dkt", "dka", "dku_dkt", "dka_dkt", "max_pkc", "avg_pkc", "min_pkc",
             "max_pkc_alt", "avg_pkc_alt", "min_pkc_alt"]


# ── VCF reading ────────────────────────────────────────────────────

def _open_text(path):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path)


def _select_alt_from_gt(alts, gt):
    """``(selected_alt, alt_indices)`` for a genotype tuple (reference ``:730-744``)."""
    if gt is None:
        return (alts[0] if alts else None), []
    alt_indices = sorted(set(i for i in gt if i is not None and i > 0))
    if not alt_indices:
        return (alts[0] if alts else None), []
    return alts[alt_indices[0] - 1], alt_indices


def _parse_gt(text):
    if text in (".", ""):
        return None
    return tuple(None if t == "." else int(t) for t in text.replace("|", "/").split("/"))


def _vcf_samples(vcf_path):
    with _open_text(vcf_path) as fh:
        for line in fh:
            if line.startswith("#CHROM"):
                return line.rstrip("\n").split("\t")[9:]
            if not line.startswith("#"):
                break
    return []


def _record_gt(fields, samples, proband_id):
    fmt = fields[8].split(":") if len(fields) > 8 else []
    if "GT" not in fmt:
        return None
    sv = fields[9 + samples.index(proband_id)].split(":")
    i = fmt.index("GT")
    return _parse_gt(sv[i]) if i < len(sv) else None


def _parse_vcf_variants(vcf_path, proband_id=None):
    """List of variant dicts: chrom, pos (0-based), ref, alts, alt, id (reference
    ``:747-810``; the proband's genotype picks the ALT of a multiallelic record)."""
    samples = _vcf_samples(vcf_path)
    proband_in_vcf = proband_id is not None and proband_id in samples
    variants = []
    with _open_text(vcf_path) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            f = line.rstrip("\n").split("\t")
            alts = None if f[4] == "." else tuple(f[4].split(","))
            alt = alts[0] if alts else None
            if alts and len(alts) > 1:
                if proband_in_vcf:
                    gt = _record_gt(f, samples, proband_id)
                    alt, idx = _select_alt_from_gt(alts, gt)
                    if len(idx) > 1:
                        logger.warning("Multiallelic variant %s:%s — proband is het non-ref; only "
                                       "the first non-ref ALT (%s) will be evaluated", f[0], f[1], alt)
                    elif not idx:
                        logger.warning("Multiallelic variant %s:%s has %d ALT alleles; only the "
                                       "first ALT (%s) will be evaluated", f[0], f[1], len(alts), alt)
                else:
                    logger.warning("Multiallelic variant %s:%s has %d ALT alleles; only the first "
                                   "ALT (%s) will be evaluated", f[0], f[1], len(alts), alt)
            variants.append({"chrom": f[0], "pos": int(f[1]) - 1, "ref": f[3], "alts": alts,
                             "alt": alt, "id": None if f[2] == "." else f[2]})
    return variants


def _var_key(var):
    alt = var["alt"] if var["alt"] is not None else "."
    return "%s:%d:%s:%s" % (var["chrom"], var["pos"], var["ref"], alt)


# ── Step 2: child k-mers over the candidate sites ──────────────────

def _reference_lengths(batch):
    """Reference bases consumed by every record of a batch (vectorised CIGAR walk)."""
    n = batch.n_reads
    out = np.zeros(n, dtype=np.int64)
    c = batch.cigar_blob
    if c.shape[0] == 0:
        return out
    op = c & np.uint32(15)
    ln = (c >> np.uint32(4)).astype(np.int64)
    consume = np.where(np.isin(op, (0, 2, 3, 7, 8)), ln, 0)
    cs = np.concatenate(([0], np.cumsum(consume)))
    off = batch.cigar_off.astype(np.int64)
    return cs[off[1:]] - cs[off[:-1]]


def _collect_child_kmers(child_bam, ref_fasta, variants, kmer_size, min_baseq, min_mapq,
                         debug_kmers, kmer_fasta, flush_threshold=500_000, threads=4):
    """Child k-mers spanning each variant → ``(total_child_kmers, variant_read_kmers)``
    and a k-mer FASTA at ``kmer_fasta`` (reference ``:619-726``).  ``bam.fetch(chrom,
    pos, pos + 1)`` goes through the .bai when the BAM has one (only the blocks over the
    sites are inflated); without an index it is one pass over the file with a sorted-interval
    lookup per batch."""
    by_chrom = collections.defaultdict(list)
    variant_read_kmers = {}
    for var in variants:
        key = _var_key(var)
        variant_read_kmers[key] = []
        if var["alt"] is not None and _is_symbolic(var["alt"]):
            logger.debug("Skipping variant %s with symbolic allele %s", key, var["alt"])
            continue
        by_chrom[var["chrom"]].append(var)
    batch_set = set()
    total_written = 0
    total_reads_scanned = 0
    fasta_fh = open(kmer_fasta, "w")

    def _flush():
        nonlocal total_written
        for kmer in batch_set:
            fasta_fh.write(">%d\n%s\n" % (total_written, kmer))
            total_written += 1
        batch_set.clear()

    def _take(batch, tid, vlist, vpos):
        """Reads of ``batch`` on reference ``tid`` against the (position-sorted) variants
        ``vlist``: one vectorised interval lookup per batch instead of a scan per variant;
        (variant, read) pairs are visited variant by variant in file order, as
        ``bam.fetch`` per variant would deliver them."""
        nonlocal total_reads_scanned
        flag = batch.flag
        ok = ((flag & np.uint16(0x4 | 0x100 | 0x800 | 0x400)) == 0) & (batch.mapq >= min_mapq) & \
            (batch.ref_id == tid)
        sel = np.flatnonzero(ok)
        if sel.size == 0:
            return
        start = batch.pos.astype(np.int64)[sel]
        end = start + _reference_lengths(batch)[sel]
        lo = np.searchsorted(vpos, start, side="left")
        hi = np.searchsorted(vpos, end, side="left")
        cover = np.flatnonzero(hi > lo)
        if cover.size == 0:
            return
        pairs = [(v, int(sel[j])) for j in cover.tolist() for v in range(int(lo[j]), int(hi[j]))]
        pairs.sort()
        for v, i in pairs:
            var = vlist[v]
            pos = var["pos"]
            key = _var_key(var)
            read = batch.record(i)
            total_reads_scanned += 1
            seq = read.query_sequence
            quals = read.query_qualities
            kmers = extract_variant_spanning_kmers(
                read, pos, kmer_size, min_baseq, ref=var["ref"], alt=var["alt"],
                seq=seq, quals=quals)
            if kmers:
                supports = read_supports_alt(read, pos, var["ref"], var["alt"],
                                             min_baseq=min_baseq, seq=seq, quals=quals)
                variant_read_kmers[key].append((read.query_name, kmers, supports))
                batch_set.update(kmers)
                if len(batch_set) >= flush_threshold:
                    _flush()

    for vlist in by_chrom.values():
        vlist.sort(key=lambda v: v["pos"])      # stable: equal positions keep their VCF order
    with bamio.BamReader(child_bam, threads=threads) as rd:
        tid_of = {name: i for i, name in enumerate(rd.references)}
        if bamio.find_bai(child_bam) is not None and os.environ.get("KDF_VCF_FETCH", "1") != "0":
            # indexed BAM: seek to the blocks over each cluster of sites (the reference's
            # bam.fetch(chrom, pos, pos + 1)) instead of decoding the whole file
            for chrom, vlist in by_chrom.items():
                tid = tid_of.get(chrom)
                if tid is None:
                    continue
                i = 0
                while i < len(vlist):
                    j = i
                    while j + 1 < len(vlist) and vlist[j + 1]["pos"] - vlist[j]["pos"] < 16384:
                        j += 1
                    group = vlist[i:j + 1]
                    gpos = np.asarray([v["pos"] for v in group], dtype=np.int64)
                    for batch in rd.fetch(tid, int(gpos[0]), int(gpos[-1]) + 1, want_meta=2):
                        _take(batch, tid, group, gpos)
                        batch.close()
                    i = j + 1
        else:
            vpos_of = {chrom: np.asarray([v["pos"] for v in vlist], dtype=np.int64)
                       for chrom, vlist in by_chrom.items()}
            for batch in rd.batches(bamio.MODE_ALL, max_bases=1 << 28, want_meta=2):
                for tid in np.unique(batch.ref_id).tolist():
                    chrom = rd.references[tid] if 0 <= tid < len(rd.references) else None
                    if chrom in by_chrom:
                        _take(batch, tid, by_chrom[chrom], vpos_of[chrom])
                batch.close()
    if batch_set:
        _flush()
    fasta_fh.close()
    if debug_kmers:
        for key, lst in variant_read_kmers.items():
            uniq = set().union(*(k for _, k, _ in lst)) if lst else set()
            logger.info("Variant %s: %d reads, %d unique k-mers", key, len(lst), len(uniq))
    logger.info("[Step 2/5] %d reads scanned, %d k-mers collected", total_reads_scanned, total_written)
    return total_written, variant_read_kmers


# ── informative reads (DV-tagged BAM) ──────────────────────────────

def _write_informative_reads(child_bam, ref_fasta, informative_reads_by_variant, output_bam,
                             threads=4):
    """Child reads carrying informative k-mers → coordinate-sorted, indexed BAM; each read
    is tagged ``DV:Z`` with the (sorted, comma-joined) variant keys it supports (reference
    ``vcf/pipeline.py:1307-1357``).  As there: the regions are visited in sorted (chrom, pos)
    order, every record that overlaps a site is considered (``bam.fetch``), and a read name
    is written once — the first record met.  Needs the BAM's ``.bai``."""
    read_to_variants = {}
    for var_key, read_names in informative_reads_by_variant.items():
        for rname in read_names:
            read_to_variants.setdefault(rname, set()).add(var_key)
    regions = set()
    for var_key in informative_reads_by_variant:
        parts = var_key.split(":")
        regions.add((parts[0], int(parts[1])))
    records, written = [], set()
    with bamio.BamReader(child_bam, threads=threads) as rd:
        if bamio.find_bai(child_bam) is None:
            raise bamio._engine.KdfError(
                "--informative-reads needs an indexed child BAM (%s.bai not found)" % child_bam)
        tid_of = {name: i for i, name in enumerate(rd.references)}
        for chrom, pos in sorted(regions):
            tid = tid_of.get(chrom)
            if tid is None:
                continue
            for batch in rd.fetch(tid, pos, pos + 1, want_meta=3):
                start = batch.pos.astype(np.int64)
                end = start + np.maximum(_reference_lengths(batch), 1)
                hit = np.flatnonzero((batch.ref_id == tid) & (start < pos + 1) & (end > pos))
                ro = batch.raw_off.astype(np.int64)
                for i in hit.tolist():
                    name = batch.record(i).query_name
                    if name in read_to_variants and name not in written:
                        written.add(name)
                        raw = batch.raw_blob[ro[i]:ro[i + 1]].tobytes()
                        records.append(bamio.append_z_tag(raw, "DV", ",".join(sorted(read_to_variants[name]))))
                batch.close()
        n = bamio.write_sorted_bam(output_bam, rd.header_text, rd.references, rd.lengths, records)
    logger.info("[Step 5/5] Informative reads BAM written: %s (%d reads)", output_bam, n)
    return n


# ── Step 4: annotate ───────────────────────────────────────────────

def _annotate_variants(variants, variant_read_kmers, parent_found_kmers):
    """Per-variant DKU / DKT / DKA and parent k-mer count statistics (reference
    ``:1662-1728``).  → (annotations, informative names, informative ALT names)."""
    parent_kmer_set = set(parent_found_kmers)
    annotations, inf_by_var, inf_alt_by_var = {}, {}, {}
    for var in variants:
        var_key = _var_key(var)
        spanning, informative, informative_alt = set(), set(), set()
        all_kmers, alt_kmers = set(), set()
        for read_name, kmers, supports_alt in variant_read_kmers.get(var_key, []):
            spanning.add(read_name)
            all_kmers.update(kmers)
            if supports_alt:
                alt_kmers.update(kmers)
            if not kmers.issubset(parent_kmer_set):
                informative.add(read_name)
                if supports_alt:
                    informative_alt.add(read_name)
        dkt, dku, dka = len(spanning), len(informative), len(informative_alt)
        pc = [parent_found_kmers[x] for x in all_kmers if x in parent_kmer_set]
        pca = [parent_found_kmers[x] for x in alt_kmers if x in parent_kmer_set]
        annotations[var_key] = {
            "dku": dku, "dkt": dkt, "dka": dka,
            "dku_dkt": round(dku / dkt, 4) if dkt > 0 else 0.0,
            "dka_dkt": round(dka / dkt, 4) if dkt > 0 else 0.0,
            "max_pkc": max(pc) if pc else 0,
            "avg_pkc": round(statistics.mean(pc), 2) if pc else 0.0,
            "min_pkc": min(pc) if pc else 0,
            "max_pkc_alt": max(pca) if pca else 0,
            "avg_pkc_alt": round(statistics.mean(pca), 2) if pca else 0.0,
            "min_pkc_alt": min(pca) if pca else 0,
        }
        if informative:
            inf_by_var[var_key] = informative
        if informative_alt:
            inf_alt_by_var[var_key] = informative_alt
    return annotations, inf_by_var, inf_alt_by_var


# ── Step 5: writers ────────────────────────────────────────────────

def _bgzf_write(path, data):
    """bgzip-compatible output (64 KiB BGZF blocks + EOF block).  → file offset of
    every block (for virtual offsets)."""
    offs = []
    with open(path, "wb") as fh:
        for off in range(0, len(data), 0xFF00):
            offs.append(fh.tell())
            chunk = data[off:off + 0xFF00]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" +
                     struct.pack("<H", len(comp) + 25) + comp +
                     struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        offs.append(fh.tell())
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return offs


def _write_tabix_index(vcf_gz, data, block_offs):
    """``.tbi`` for a bgzipped VCF whose uncompressed bytes are ``data`` (what
    ``pysam.tabix_index(preset="vcf")`` writes; reference ``:1302``): the BAI
    binning scheme over (CHROM, POS-1, POS-1+len(REF)), BGZF-compressed."""
    def voff(u):
        if u >= len(data):
            return block_offs[-1] << 16          # the EOF block
        b = u // 0xFF00
        return (block_offs[b] << 16) | (u - b * 0xFF00)

    names, bins, linear = [], [], []
    pos = 0
    for line in data.split(b"\n"):
        start, pos = pos, pos + len(line) + 1
        if not line or line[:1] == b"#":
            continue
        f = line.split(b"\t", 5)
        chrom = f[0].decode()
        if chrom not in names:
            names.append(chrom)
            bins.append({})
            linear.append({})
        t = names.index(chrom)
        beg = int(f[1]) - 1
        end = beg + max(len(f[3]), 1)
        v0, v1 = voff(start), voff(min(pos, len(data)))
        chunks = bins[t].setdefault(bamio._reg2bin(beg, end), [])
        if chunks and chunks[-1][1] == v0:
            chunks[-1][1] = v1
        else:
            chunks.append([v0, v1])
        for w in range(beg >> 14, ((end - 1) >> 14) + 1):
            linear[t].setdefault(w, v0)
    nm = b"".join(n.encode() + b"\0" for n in names)
    out = bytearray(b"TBI\1" + struct.pack("<iiiiiii", len(names), 2, 1, 2, 0, ord("#"), 0) +
                    struct.pack("<i", len(nm)) + nm)
    for t in range(len(names)):
        out += struct.pack("<i", len(bins[t]))
        for b in sorted(bins[t]):
            out += struct.pack("<Ii", b, len(bins[t][b]))
            for v0, v1 in bins[t][b]:
                out += struct.pack("<QQ", v0, v1)
        n_intv = (max(linear[t]) + 1) if linear[t] else 0
        out += struct.pack("<i", n_intv)
        last = 0
        for w in range(n_intv):
            last = linear[t].get(w, last)
            out += struct.pack("<Q", last)
    _bgzf_write(vcf_gz + ".tbi", bytes(out))


def _fmt_value(key, ann):
    v = ann[key]
    return "%g" % v if isinstance(v, float) else str(v)


def _write_annotated_vcf(input_vcf, output_vcf, annotations, proband_id=None):
    """Annotated VCF, bgzip-compressed (reference ``:813-1304``).  FORMAT fields on
    the proband's sample when it is in the VCF, INFO fields otherwise.  Records
    and header are passed through textually; the new header lines go last, as
    htslib appends them; a ``.tbi`` tabix index is written beside the output."""
    samples = _vcf_samples(input_vcf)
    use_format = proband_id is not None and proband_id in samples
    if proband_id is not None and not use_format:
        logger.warning("Proband '%s' not found in VCF samples (%s); falling back to INFO annotation",
                       proband_id, samples)
    category = "FORMAT" if use_format else "INFO"
    names = [d[0] for d in _FIELD_DEFS]
    out = []
    with _open_text(input_vcf) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith("#CHROM"):
                for name, typ, desc in _FIELD_DEFS:
                    out.append('##%s=<ID=%s,Number=1,Type=%s,Description="%s">' % (category, name, typ, desc))
                out.append(line)
                continue
            if line.startswith("#") or not line:
                out.append(line)
                continue
            f = line.split("\t")
            alts = None if f[4] == "." else tuple(f[4].split(","))
            alt_str = alts[0] if alts else "."
            if use_format and alts and len(alts) > 1:
                sel, _ = _select_alt_from_gt(alts, _record_gt(f, samples, proband_id))
                alt_str = sel if sel is not None else "."
            ann = annotations.get("%s:%d:%s:%s" % (f[0], int(f[1]) - 1, f[3], alt_str))
            if ann is not None:
                vals = [_fmt_value(k, ann) for k in _ANN_KEYS]
                if use_format:
                    f[8] = f[8] + ":" + ":".join(names)
                    pi = 9 + samples.index(proband_id)
                    for i in range(9, len(f)):
                        f[i] = f[i] + ":" + (":".join(vals) if i == pi else ":".join("." for _ in vals))
                else:
                    extra = ";".join("%s=%s" % kv for kv in zip(names, vals))
                    f[7] = extra if f[7] in (".", "") else f[7] + ";" + extra
            out.append("\t".join(f))
    if not output_vcf.endswith(".gz"):
        output_vcf += ".gz"
    data = ("\n".join(out) + "\n").encode()
    _write_tabix_index(output_vcf, data, _bgzf_write(output_vcf, data))
    return output_vcf


def _write_summary(summary_path, variants, annotations):
    """Human-readable summary, byte-for-byte the reference's format (``:1360-1451``)."""
    total = len(variants)
    vals = list(annotations.values())
    likely_dnm = sum(1 for a in vals if a["dku"] > 0)
    lines = ["=" * 60, "  kmer-denovo  —  De Novo Variant Summary", "=" * 60, "",
             "Variant Counts", "-" * 40,
             f"  Total candidates analyzed:   {total:>6}",
             f"  Likely de novo (DKU > 0):    {likely_dnm:>6}",
             f"  Inherited / unclear (DKU=0): {total - likely_dnm:>6}", ""]
    if vals:
        mean = lambda k: sum(a[k] for a in vals) / len(vals)
        median_dku = statistics.median([a["dku"] for a in vals])
        lines += ["Read Support Statistics", "-" * 40,
                  f"  DKU  mean:   {mean('dku'):>6.1f}   median: {median_dku:>4}",
                  f"  DKT  mean:   {mean('dkt'):>6.1f}",
                  f"  DKA  mean:   {mean('dka'):>6.1f}",
                  f"  DKU_DKT  mean: {mean('dku_dkt'):>6.4f}",
                  f"  DKA_DKT  mean: {mean('dka_dkt'):>6.4f}",
                  f"  MAX_PKC  mean: {mean('max_pkc'):>6.1f}",
                  f"  AVG_PKC  mean: {mean('avg_pkc'):>6.1f}",
                  f"  MIN_PKC  mean: {mean('min_pkc'):>6.1f}",
                  f"  MAX_PKC_ALT  mean: {mean('max_pkc_alt'):>6.1f}",
                  f"  AVG_PKC_ALT  mean: {mean('avg_pkc_alt'):>6.1f}",
                  f"  MIN_PKC_ALT  mean: {mean('min_pkc_alt'):>6.1f}", ""]
    dnm = [a["dku"] for a in vals if a["dku"] > 0]
    if dnm:
        lines += [f"  Avg DKU among likely DNMs:   {sum(dnm) / len(dnm):>6.1f}", ""]
    lines += ["Per-Variant Results", "-" * 120,
              f"  {'Variant':<30s} {'DKU':>5s} {'DKT':>5s} {'DKA':>5s} {'DKU_DKT':>8s} {'DKA_DKT':>8s} "
              f"{'MAX_PKC':>8s} {'AVG_PKC':>8s} {'MIN_PKC':>8s} {'MAX_PKC_ALT':>12s} {'AVG_PKC_ALT':>12s} "
              f"{'MIN_PKC_ALT':>12s}  Call",
              f"  {'-------':<30s} {'---':>5s} {'---':>5s} {'---':>5s} {'-------':>8s} {'-------':>8s} "
              f"{'-------':>8s} {'-------':>8s} {'-------':>8s} {'-----------':>12s} {'-----------':>12s} "
              f"{'-----------':>12s}  ----"]
    zero = {k: (0.0 if k in ("dku_dkt", "dka_dkt", "avg_pkc", "avg_pkc_alt") else 0) for k in _ANN_KEYS}
    for var in variants:
        alts = var["alts"]
        alt = var.get("alt") if var.get("alt") is not None else (alts[0] if alts else ".")
        a = annotations.get("%s:%d:%s:%s" % (var["chrom"], var["pos"], var["ref"], alt), zero)
        label = f"{var['chrom']}:{var['pos'] + 1} {var['ref']}>{alt}"
        call = "DE_NOVO" if a["dku"] > 0 else "inherited"
        lines.append(f"  {label:<30s} {a['dku']:>5d} {a['dkt']:>5d} {a['dka']:>5d} {a['dku_dkt']:>8.4f} "
                     f"{a['dka_dkt']:>8.4f} {a['max_pkc']:>8d} {a['avg_pkc']:>8.2f} {a['min_pkc']:>8d} "
                     f"{a['max_pkc_alt']:>12d} {a['avg_pkc_alt']:>12.2f} {a['min_pkc_alt']:>12d}  {call}")
    lines += ["", "=" * 60, ""]
    text = "\n".join(lines)
    with open(summary_path, "w") as fh:
        fh.write(text)
    return text


# ── driver ─────────────────────────────────────────────────────────

def _validate(args):
    k = args.kmer_size
    if k < 3 or k > 201 or k % 2 == 0:
        logger.error("--kmer-size must be an odd integer between 3 and 201 (got %d)", k)
        sys.exit(1)
    if k > 63:
        logger.error("k=%d: the GPU engine supports k <= 63 (64-bit keys for k <= 32, 128-bit above)", k)
        sys.exit(1)
    for what in ("child", "mother", "father", "vcf"):
        path = getattr(args, what)
        if not path or not os.path.isfile(path):
            logger.error("%s file not found: %s", what, path)
            sys.exit(1)


def run_pipeline(args, engine=None):
    """``kmer-denovo``: annotate candidate variants with child-unique k-mer evidence.
    → dict(metrics, annotations, output paths)."""
    logging.basicConfig(level=logging.DEBUG if getattr(args, "debug_kmers", False) else logging.INFO,
                        format="%(asctime)s %(levelname)s %(message)s")
    _validate(args)
    eng = get_engine(engine)
    threads = getattr(args, "threads", 4) or 4
    t0 = time.monotonic()
    variants = _parse_vcf_variants(args.vcf, getattr(args, "proband_id", None))
    logger.info("[Step 1/5] Parsed %d candidate variants", len(variants))
    tmp_root = getattr(args, "tmp_dir", None) or tempfile.gettempdir()
    # the two whole-file parent scans dominate this mode: start decoding both parents now, in the
    # background (bounded look-ahead), while the child's reads over the sites are collected
    from ..core import kmer_engine_wrappers as _kw
    if os.environ.get("KDF_PREFETCH_PARENTS", "1") != "0":
        _kw.start_prefetch(args.mother, bamio.MODE_FASTA, threads)
        _kw.start_prefetch(args.father, bamio.MODE_FASTA, threads)
    with tempfile.TemporaryDirectory(dir=tmp_root) as tmpdir:
        kmer_fasta = os.path.join(tmpdir, "child_kmers.fa")
        total_child_kmers, variant_read_kmers = _collect_child_kmers(
            args.child, getattr(args, "ref_fasta", None), variants, args.kmer_size,
            getattr(args, "min_baseq", 20), getattr(args, "min_mapq", 20),
            getattr(args, "debug_kmers", False), kmer_fasta, threads=threads)
        parent_found_kmers = collections.Counter()
        if total_child_kmers == 0:
            logger.info("[Step 3/5] No child k-mers found; skipping parent scans")
        else:
            logger.info("[Step 3/5] Scanning parent BAMs for %d child k-mers", total_child_kmers)
            for label, bam in (("Mother", args.mother), ("Father", args.father)):
                found = _scan_parent_jellyfish(bam, getattr(args, "ref_fasta", None), kmer_fasta,
                                               args.kmer_size, os.path.join(tmpdir, label.lower()),
                                               threads, n_filter_kmers=total_child_kmers, engine=eng)
                parent_found_kmers.update(found)
                logger.info("[Step 3/5] %s done — %d / %d child k-mers found", label, len(found),
                            total_child_kmers)
    _kw.drop_prefetch()      # (started above; consumed by the scans unless there was nothing to scan)
    child_unique_kmers = max(0, total_child_kmers - len(parent_found_kmers))
    annotations, inf_by_var, _inf_alt = _annotate_variants(variants, variant_read_kmers,
                                                           parent_found_kmers)
    out_vcf = _write_annotated_vcf(args.vcf, args.output, annotations,
                                   getattr(args, "proband_id", None))
    metrics = {
        "total_variants": len(variants),
        "total_child_kmers": total_child_kmers,
        "parent_found_kmers": len(parent_found_kmers),
        "child_unique_kmers": child_unique_kmers,
        "variants_with_unique_reads": sum(1 for a in annotations.values() if a["dku"] > 0),
    }
    paths = {"vcf": out_vcf}
    if getattr(args, "informative_reads", None):
        logger.info("[Step 5/5] Writing informative reads BAM: %s", args.informative_reads)
        _write_informative_reads(args.child, getattr(args, "ref_fasta", None), inf_by_var,
                                 args.informative_reads, threads=threads)
        paths["informative_reads"] = args.informative_reads
    if getattr(args, "metrics", None):
        with open(args.metrics, "w") as fh:
            json.dump(metrics, fh, indent=2)
        paths["metrics"] = args.metrics
    if getattr(args, "summary", None):
        _write_summary(args.summary, variants, annotations)
        paths["summary"] = args.summary
    logger.info("kmer-denovo finished in %.1fs: %d / %d variants with child-unique reads",
                time.monotonic() - t0, metrics["variants_with_unique_reads"], len(variants))
    return {"metrics": metrics, "annotations": annotations, "paths": paths,
            "parent_found_kmers": parent_found_kmers}
