"""Command lines ``kmer-denovo`` (VCF mode) and ``kmer-discovery`` — the same
flags, defaults and destinations as the reference's ``cli.py`` (``:10-230``).
Flags that only tuned Jellyfish / the process pool (``--threads``,
``--memory``, ``--jf-hash-size``, ``--tmp-dir``) still parse; ``--threads``
drives the host BAM decoder, ``--jf-hash-size`` the initial table size, the
others are accepted and ignored with a warning.  Kraken2 and HTML-report outputs
are not part of this package: asking for one is an error (exit 2), never a silent
no-op."""

import argparse
import sys


def _add_shared_args(p):
    p.add_argument("--child", required=True, help="Child BAM file")
    p.add_argument("--mother", required=True, help="Mother BAM file")
    p.add_argument("--father", required=True, help="Father BAM file")
    p.add_argument("--ref-fasta", "-r", default=None, help="Reference FASTA")
    p.add_argument("--kmer-size", "-k", type=int, default=31, help="K-mer size (default: 31)")
    p.add_argument("--min-baseq", type=int, default=20,
                   help="Minimum base quality for read k-mers (default: 20)")
    p.add_argument("--threads", "-t", type=int, default=4,
                   help="Host threads for BAM decoding (default: 4)")
    p.add_argument("--memory", type=float, default=None, help="Accepted for compatibility")
    p.add_argument("--debug-kmers", action="store_true", default=False,
                   help="Enable per-variant debug output")
    p.add_argument("--jf-hash-size", default=None,
                   help="Initial k-mer table size in entries (e.g. '2G', '500M')")
    p.add_argument("--tmp-dir", default=None, help="Accepted for compatibility (no temp files)")


def parse_vcf_args(argv=None):
    p = argparse.ArgumentParser(prog="kmer-denovo",
                                description="De novo variant curation using k-mer analysis "
                                            "(VCF mode, B200 engine)")
    _add_shared_args(p)
    p.add_argument("--vcf", required=True, help="Input VCF with candidate variants")
    p.add_argument("--output", "-o", required=True, help="Output annotated VCF")
    p.add_argument("--metrics", default=None, help="Output summary metrics JSON file")
    p.add_argument("--summary", default=None, help="Output human-readable summary")
    p.add_argument("--informative-reads", default=None,
                   help="Output BAM of the informative child reads, tagged DV:Z (needs the child's .bai)")
    p.add_argument("--min-mapq", type=int, default=20,
                   help="Minimum mapping quality for child reads (default: 20)")
    p.add_argument("--proband-id", default=None, help="Sample ID of the proband in the VCF")
    p.add_argument("--kraken2-db", default=None, help="Not supported by this package")
    p.add_argument("--kraken2-confidence", type=float, default=0.0)
    p.add_argument("--kraken2-memory-mapping", action="store_true", default=False)
    p.add_argument("--kraken2-read-detail", default=None)
    p.add_argument("--kraken2-span-bed", default=None)
    p.add_argument("--no-expanded-bed", action="store_true", default=False)
    p.add_argument("--report", default=None, help="Not supported by this package")
    return p.parse_args(argv)


def parse_discovery_args(argv=None):
    p = argparse.ArgumentParser(prog="kmer-discovery",
                                description="VCF-free de novo k-mer discovery pipeline "
                                            "(B200 engine)")
    _add_shared_args(p)
    p.add_argument("--out-prefix", required=True, help="Output prefix for discovery mode files")
    p.add_argument("--ref-jf", default=None,
                   help="Precomputed Jellyfish reference index (binary/sorted)")
    p.add_argument("--min-child-count", type=int, default=3,
                   help="Minimum child k-mer occurrences (default: 3)")
    p.add_argument("--candidate-summary", default=None,
                   help="VCF-mode summary.txt for candidate comparison")
    p.add_argument("--cluster-distance", type=int, default=500,
                   help="Maximum gap (bp) for merging adjacent regions (default: 500)")
    p.add_argument("--min-supporting-reads", type=int, default=1)
    p.add_argument("--min-distinct-kmers", type=int, default=1)
    p.add_argument("--min-bedgraph-reads", type=int, default=3)
    p.add_argument("--min-distinct-kmers-per-read", type=int, default=None,
                   help="Minimum distinct proband-unique k-mers per read (default: k/4)")
    p.add_argument("--parent-max-count", type=int, default=0)
    p.add_argument("--sv-bedpe", default=None)
    p.add_argument("--report", default=None, help="Not supported by this package")
    return p.parse_args(argv)


def _check_unsupported(args):
    """Flags of the reference CLI whose subsystems are not part of this package: outputs that
    cannot be produced are refused (exit 2) rather than silently skipped; tuning flags that
    have no meaning here are reported and ignored."""
    import logging
    log = logging.getLogger(__name__)
    fatal = []
    if getattr(args, "report", None):
        fatal.append("--report: the HTML report is not part of this package")
    for flag in ("kraken2_db", "kraken2_read_detail", "kraken2_span_bed"):
        if getattr(args, flag, None):
            fatal.append("--%s: Kraken2 contamination screening is not part of this package"
                         % flag.replace("_", "-"))
    if fatal:
        for msg in fatal:
            sys.stderr.write("error: %s\n" % msg)
        sys.exit(2)
    for flag, why in (("memory", "no Jellyfish hash to size"), ("tmp_dir", "no temporary k-mer files")):
        if getattr(args, flag, None):
            log.warning("--%s is accepted for compatibility and ignored (%s)", flag.replace("_", "-"), why)
    if getattr(args, "kraken2_memory_mapping", False) or getattr(args, "kraken2_confidence", 0.0):
        log.warning("--kraken2-* tuning flags are ignored (no Kraken2 screening in this package)")


def vcf_main(argv=None):
    from .vcf.pipeline import run_pipeline
    args = parse_vcf_args(argv)
    _check_unsupported(args)
    run_pipeline(args)


def _init_distributed():
    """Under ``torchrun`` (one process per GPU): pick this rank's device and join the NCCL
    group; → engine for that device, or None for a plain single-process run."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    from . import engine
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return engine.CudaEngine(dev)


def discovery_main(argv=None):
    from .discovery.pipeline import run_discovery_pipeline
    args = parse_discovery_args(argv)
    _check_unsupported(args)
    eng = _init_distributed()       # torchrun --nproc-per-node N -m kmer_denovo_filter_b200.cli ...
    run_discovery_pipeline(args, engine=eng)
    if eng is not None:
        import torch.distributed as dist
        dist.destroy_process_group()


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if "--out-prefix" in argv:
        discovery_main(argv)
    else:
        vcf_main(argv)


if __name__ == "__main__":
    main()
