"""VCF mode on the GPU k-mer engine."""
