"""Shared core: GPU-engine wrappers with the reference's core/ function names."""
