"""Discovery mode (VCF-free) on the GPU k-mer engine."""
from .pipeline import run_discovery_pipeline  # noqa: F401
