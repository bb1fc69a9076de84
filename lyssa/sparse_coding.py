"""lyssa.sparse_coding -> lyssandra_b200.sparse_coding (reference: lyssa/sparse_coding.py:587)."""
from lyssandra_b200.sparse_coding import sparse_encoder  # noqa: F401
