"""Human-readable discovery summary (output contract of the reference's
``_write_discovery_summary``, ``discovery/pipeline.py:1786-1976``; the golden
file is ``tests/golden/expected_discovery/giab_discovery.summary.txt``)."""

import statistics

RULE = "=" * 60


def _row(spec, values):
    """Format one table row from ``[(width, align)]`` and the cell strings."""
    cells = []
    for (width, align), v in zip(spec, values):
        cells.append(("%-*s" if align == "<" else "%*s") % (width, v))
    return "  " + cells[0] + "".join(sep + c for sep, c in zip(spec.seps, cells[1:]))


class _Spec(list):
    def __init__(self, cols, seps):
        super().__init__(cols)
        self.seps = seps


_REGION = _Spec([(35, "<"), (8, ">"), (6, ">"), (14, ">"), (6, ">"), (5, ">"), (8, ">"),
                 (10, ">"), (10, ">")], [" "] * 8)
_CAND = _Spec([(30, "<"), (4, ">"), (8, ">"), (35, ">")], ["  "] * 3)
_DNM = _Spec([(20, "<"), (25, ">"), (8, ">"), (6, ">"), (6, ">"), (7, ">"), (8, ">"), (10, ">"),
              (14, ">")], [" "] * 8)


def _stat_line(label, values, mean_fmt, unit=""):
    med = statistics.median(values)
    return "  %s mean: %s%s   median: %4s%s   max: %4s%s" % (
        label, mean_fmt % (sum(values) / len(values)), unit, med, unit, max(values), unit)


def _write_discovery_summary(summary_path, regions, region_reads, region_kmers, metrics,
                             candidate_comparison=None, region_annotations=None,
                             dnm_evaluation=None):
    out = [RULE, "  kmer-denovo  —  Discovery Mode Summary", RULE, "",
           "K-mer Filtering", "-" * 40,
           "  Child candidate k-mers:      %8s" % metrics["child_candidate_kmers"],
           "  Non-reference k-mers:        %8s" % metrics["non_ref_kmers"],
           "  Proband-unique k-mers:       %8s" % metrics["proband_unique_kmers"], "",
           "Region Counts", "-" * 40,
           "  Candidate regions:           %8s" % metrics["candidate_regions"],
           "  Total informative reads:     %8s" % metrics["informative_reads"]]
    n_unmapped = metrics.get("unmapped_informative_reads", 0)
    if n_unmapped > 0:
        out.append("    (unmapped informative):     %8s" % n_unmapped)
    out.append("")
    region_annotations = region_annotations or {}
    if regions:
        n_reads = [len(region_reads.get(r, ())) for r in regions]
        n_kmers = [len(region_kmers.get(r, ())) for r in regions]
        sizes = [e - s for _c, s, e in regions]
        out += ["Region Statistics", "-" * 40,
                _stat_line("Reads/region  ", n_reads, "%6.1f"),
                _stat_line("K-mers/region ", n_kmers, "%6.1f"),
                _stat_line("Region size   ", sizes, "%6.0f", " bp"), "",
                "Per-Region Results", "-" * 120,
                _row(_REGION, ["Region", "Size", "Reads", "Unique K-mers", "Split", "Disc",
                               "MaxClip", "UnmapMate", "Class"]),
                _row(_REGION, ["------", "----", "-----", "-------------", "-----", "----",
                               "-------", "---------", "-----"])]
        for rk, nr, nk in zip(regions, n_reads, n_kmers):
            ann = region_annotations.get(rk, {})
            out.append(_row(_REGION, [
                "%s:%d-%d" % (rk[0], rk[1] + 1, rk[2]), "%7dbp" % (rk[2] - rk[1]), nr, nk,
                ann.get("split_reads", 0), ann.get("discordant_pairs", 0),
                ann.get("max_clip_len", 0), ann.get("unmapped_mates", 0),
                ann.get("class", "SMALL")]))
    if candidate_comparison:
        total = len(candidate_comparison)
        cap = sum(1 for c in candidate_comparison if c["captured"])
        out += ["Candidate Comparison (DKA_DKT > 0.25, DKA > 10)", "-" * 80,
                "  High-quality candidates:     %8s" % total,
                "  Captured by discovery:       %8s / %d (%.1f%%)" % (
                    cap, total, cap / total * 100 if total else 0.0), "",
                _row(_CAND, ["Candidate", "DKA", "DKA_DKT", "Region"]),
                _row(_CAND, ["---------", "---", "-------", "------"])]
        for c in candidate_comparison:
            out.append(_row(_CAND, [
                "%s:%d %s>%s" % (c["chrom"], c["pos"], c["ref"], c["alt"]), "%d" % c["dka"],
                "%.4f" % c["dka_dkt"], c["region"] if c["captured"] else "NOT CAPTURED"]))
        out.append("")
    if dnm_evaluation:
        total = len(dnm_evaluation)
        det = sum(1 for e in dnm_evaluation if e["detected"])
        out += ["Curated DNM Region Evaluation (Sulovari et al. 2023)", "-" * 80,
                "  Curated DNM loci:            %8s" % total,
                "  Detected by discovery:       %8s / %d (%.1f%%)" % (
                    det, total, det / total * 100 if total else 0.0), "",
                _row(_DNM, ["Locus", "Event", "Size", "Reads", "Kmers", "Signal", "MaxClip",
                            "Class", "Status"]),
                _row(_DNM, ["-----", "-----", "----", "-----", "-----", "------", "-------",
                            "-----", "------"])]
        for e in dnm_evaluation:
            out.append(_row(_DNM, [
                e["locus"], e["event_type"],
                ("%dbp" % e["event_size"]) if e["event_size"] else "–",
                e["total_reads"], e["total_unique_kmers"], "%.4f" % e["kmer_signal"],
                e["max_clip_len"], e["sv_class"], e["assessment"]]))
        out.append("")
    out += [RULE, ""]
    text = "\n".join(out)
    with open(summary_path, "w") as fh:
        fh.write(text)
    return text
