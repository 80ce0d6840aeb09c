"""CPU restatement of the reference's VCF mode (TEST INFRASTRUCTURE ONLY).

Follows ``vcf/pipeline.py`` (``_parse_vcf_variants`` :747-810,
``_select_alt_from_gt`` :730-744, ``_collect_child_kmers`` :619-726, the parent
scans :1587-1609, the annotate loop :1662-1728, ``_write_summary`` :1360-1451)
and ``kmer_utils.py`` (``extract_variant_spanning_kmers`` :1102-1172,
``read_supports_alt`` :1037-1099, ``_is_symbolic`` :18-27).  Pinned against the
reference's committed goldens ``tests/example_output/{metrics.json,
summary.txt, annotated.vcf.gz}`` in ``tests/test_oracle_vcf.py``.
"""
import collections
import gzip
import statistics

from . import bam as obam
from . import kmers


def is_symbolic(allele):
    """kmer_utils.py:18-27."""
    if allele is None:
        return False
    return allele.startswith("<") or allele == "*" or "[" in allele or "]" in allele


def select_alt_from_gt(alts, gt):
    """vcf/pipeline.py:730-744."""
    if gt is None:
        return (alts[0] if alts else None), []
    idx = sorted(set(i for i in gt if i is not None and i > 0))
    if not idx:
        return (alts[0] if alts else None), []
    return alts[idx[0] - 1], idx


def parse_gt(text):
    if text in (".", ""):
        return None
    out = []
    for tok in text.replace("|", "/").split("/"):
        out.append(None if tok == "." else int(tok))
    return tuple(out)


def parse_vcf(path, proband_id=None):
    """→ (header lines, sample names, variants).  Each variant: chrom, pos
    (0-based), ref, alts (tuple or None), alt (the allele evaluated), id, fields."""
    opener = gzip.open if path.endswith(".gz") else open
    header, samples, variants = [], [], []
    with opener(path, "rt") as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith("##"):
                header.append(line)
                continue
            if line.startswith("#"):
                header.append(line)
                cols = line.split("\t")
                samples = cols[9:]
                continue
            if not line:
                continue
            f = line.split("\t")
            alts = None if f[4] == "." else tuple(f[4].split(","))
            alt = alts[0] if alts else None
            if alts and len(alts) > 1 and proband_id is not None and proband_id in samples:
                fmt = f[8].split(":")
                sv = f[9 + samples.index(proband_id)].split(":")
                gt = parse_gt(sv[fmt.index("GT")]) if "GT" in fmt else None
                alt, _idx = select_alt_from_gt(alts, gt)
            variants.append({"chrom": f[0], "pos": int(f[1]) - 1, "ref": f[3], "alts": alts,
                             "alt": alt, "id": None if f[2] == "." else f[2], "fields": f})
    return header, samples, variants


def spanning_kmers(rec, variant_pos, k, min_baseq=0, ref=None, alt=None):
    """kmer_utils.py:1102-1172 (canonical k-mer strings spanning the variant)."""
    try:
        rp = rec.get_reference_positions(full_length=True).index(variant_pos)
    except ValueError:
        return set()
    seq = rec.query_sequence
    if seq is None:
        return set()
    quals = rec.query_qualities
    alt_len = len(alt) if alt and not is_symbolic(alt) else 1
    v_end = rp + alt_len - 1
    start_min = max(0, rp - k + 1)
    start_max = min(len(seq) - k, v_end)
    out = set()
    up = seq.upper()
    for s in range(start_min, start_max + 1):
        win = up[s:s + k]
        if "N" in win:
            continue
        if quals is not None and min_baseq > 0 and min(quals[s:s + k]) < min_baseq:
            continue
        out.add(kmers.canonicalize(seq[s:s + k]))
    return out


def supports_alt(rec, variant_pos, ref, alt, min_baseq=0):
    """kmer_utils.py:1037-1099."""
    if alt is None or is_symbolic(alt):
        return False
    seq = rec.query_sequence
    if seq is None:
        return False
    quals = rec.query_qualities if min_baseq > 0 else None
    got = []
    inside = False
    for qpos, rpos in rec.get_aligned_pairs(matches_only=False):
        if rpos is not None and rpos >= variant_pos + len(ref):
            break
        if rpos == variant_pos:
            inside = True
        if inside and qpos is not None:
            if min_baseq > 0 and quals is not None and quals[qpos] < min_baseq:
                return False
            got.append(seq[qpos])
    if not inside:
        return False
    return "".join(got).upper() == alt.upper()


def var_key(var):
    alt = var["alt"] if var["alt"] is not None else "."
    return "%s:%d:%s:%s" % (var["chrom"], var["pos"], var["ref"], alt)


def collect_child(records, variants, k, min_baseq, min_mapq):
    """vcf/pipeline.py:619-726 with ``bam.fetch(chrom, pos, pos + 1)`` restated as
    a file-order scan.  → (total_child_kmers, {var_key: [(qname, kmers, supports)]})."""
    by_chrom = collections.defaultdict(list)
    for r in records:
        if r.ref_id >= 0:
            by_chrom[r.reference_name].append(r)
    all_kmers = set()
    vrk = {}
    for var in variants:
        key = var_key(var)
        if var["alt"] is not None and is_symbolic(var["alt"]):
            vrk[key] = []
            continue
        pos = var["pos"]
        lst = []
        for r in by_chrom.get(var["chrom"], ()):
            if r.is_unmapped or r.is_secondary or r.is_supplementary:
                continue
            if r.mapping_quality < min_mapq or r.is_duplicate:
                continue
            end = r.reference_end
            if end is None or not (r.reference_start <= pos < end):
                continue
            km = spanning_kmers(r, pos, k, min_baseq, var["ref"], var["alt"])
            if km:
                lst.append((r.query_name, km, supports_alt(r, pos, var["ref"], var["alt"], min_baseq)))
                all_kmers.update(km)
        vrk[key] = lst
    return len(all_kmers), vrk, all_kmers


def parent_counts(parent_records, k, filter_strings):
    """core/jellyfish_wrappers.py:115-283: ``count --if`` over the samtools-fasta
    stream, then ``dump -c -L 1`` → {kmer string: count >= 1}."""
    keys = {}
    for s in filter_strings:
        v = kmers.key_of(s)
        if v is not None:
            keys[v] = s
    codes, valid, _s, _l = kmers.encode_stream([r.seq for r in obam.fasta_stream(parent_records)])
    cnt = kmers.count_stream_filtered(codes, valid, k, keys.keys())
    return {keys[x]: c for x, c in cnt.items() if c >= 1}


def annotate(variants, vrk, parent_found):
    """vcf/pipeline.py:1662-1728."""
    pset = set(parent_found)
    ann = {}
    for var in variants:
        lst = vrk.get(var_key(var), [])
        spanning, inf, inf_alt = set(), set(), set()
        allk, altk = set(), set()
        for name, km, sup in lst:
            spanning.add(name)
            allk.update(km)
            if sup:
                altk.update(km)
            if not km.issubset(pset):
                inf.add(name)
                if sup:
                    inf_alt.add(name)
        dkt, dku, dka = len(spanning), len(inf), len(inf_alt)
        pc = [parent_found[x] for x in allk if x in pset]
        pca = [parent_found[x] for x in altk if x in pset]
        ann[var_key(var)] = {
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
    return ann


def run(child_records, mother_records, father_records, variants, k, min_baseq=20, min_mapq=20):
    total, vrk, allk = collect_child(child_records, variants, k, min_baseq, min_mapq)
    found = collections.Counter()
    if total:
        found.update(parent_counts(mother_records, k, allk))
        found.update(parent_counts(father_records, k, allk))
    ann = annotate(variants, vrk, found)
    metrics = {
        "total_variants": len(variants),
        "total_child_kmers": total,
        "parent_found_kmers": len(found),
        "child_unique_kmers": max(0, total - len(found)),
        "variants_with_unique_reads": sum(1 for a in ann.values() if a["dku"] > 0),
    }
    return ann, metrics, found
