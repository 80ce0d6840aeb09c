"""CPU oracle for the k-mer hot path of jlanej/kmer_denovo_filter.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kmer_denovo_filter_b200/`` may
import, call, link or execute anything in this package.  The only allowed
consumers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` — and there only as the checker /
the timed CPU arm, never as the product.

Parity status: **pinned**.  Every function here is checked (tests/test_oracle_*)
against the reference's committed golden outputs, which were produced by the
real Jellyfish + samtools + pysam stack in the reference's CI
(reference ``tests/example_output*/``, ``tests/data/giab/mini_ref.fa.k31.jf``;
copies of the derived vectors live in ``tests/golden/``).

Modules
-------
bam        stdlib BGZF/BAM reader + writer (replaces pysam/samtools for tests)
kmers      canonical k-mer semantics (string level and numpy stream level)
discovery  discovery-mode filter chain + per-read scan + clustering
vcfmode    VCF-mode child k-mer collection, ALT support, DKU/DKT/DKA/PKC
"""
