"""lyssa.utils -> lyssandra_b200.utils (reference: lyssa/utils/__init__.py:166-201, utils/math.py)."""
from lyssandra_b200.utils import (gen_batches, gen_even_batches, fast_dot, norm, norm_cols,  # noqa: F401
                                  normalize, frobenius_squared)
