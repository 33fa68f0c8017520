"""Static version (the reference vendors versioneer, mpi4jax/_version.py; a git-tag
derived version is not needed for an in-tree, built-in-place package)."""

__version__ = "0.1.0"
