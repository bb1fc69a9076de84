"""lyssandra_b200 — B200-native Batch-OMP / approximate K-SVD / online dictionary learning
behind the API of ektormak/Lyssandra's hot path.

    from lyssandra_b200.sparse_coding import sparse_encoder
    from lyssandra_b200.dict_learning import ksvd_coder, online_dictionary_coder

The same names are importable from the ``lyssa`` alias package for drop-in use.
The CUDA library (liblyssa_b200.so, sm_100a) is loaded on first use; there is no CPU path.
"""
__version__ = "0.1.0"

from .sparse_coding import sparse_encoder  # noqa: F401
