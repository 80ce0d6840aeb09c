"""kmer_denovo_filter_b200 — B200-native k-mer engine for trio de novo filtering.

Drop-in replacement for the k-mer hot path of jlanej/kmer_denovo_filter (the
Jellyfish / samtools subprocesses and the Python per-read scan).  See DESIGN.md.
"""
__version__ = "0.1.0"
