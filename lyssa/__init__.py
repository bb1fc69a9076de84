"""Drop-in alias: the reference's import paths (README.md:17-21 of ektormak/Lyssandra),
served by the B200 engine in ``lyssandra_b200``.  Only the Batch-OMP / approximate K-SVD /
online-learning hot path exists here (see DESIGN.md for what is out of scope)."""
from lyssandra_b200 import __version__  # noqa: F401
