"""Helpers mirrored from lyssa/utils/__init__.py that the hot path's callers use:
contiguous column partitions (gen_even_batches :166-180, gen_batches :183-201).  The process
pool of run_parallel (:40-163) has no counterpart — columns are sharded over GPUs instead."""
from __future__ import annotations

import numpy as np

from .math import fast_dot, norm, norm_cols, normalize, frobenius_squared  # noqa: F401


def gen_even_batches(N, n_batches):
    """n_batches contiguous ranges; the last takes the remainder (utils/__init__.py:166-180)."""
    size = int(np.floor(N / float(n_batches)))
    out = [range(j * size, (j + 1) * size) for j in range(n_batches - 1)]
    out.append(range((n_batches - 1) * size, N))
    return out


def gen_batches(N, batch_size=None):
    """Contiguous ranges of batch_size plus one short tail (utils/__init__.py:183-201)."""
    if batch_size is None:
        return [range(0, N)]
    n_full = int(np.floor(N / float(batch_size)))
    out = [range(j * batch_size, (j + 1) * batch_size) for j in range(n_full)]
    if N > n_full * batch_size:
        out.append(range(n_full * batch_size, N))
    return out


def shard_bounds(N, world_size, rank):
    """Contiguous column block of `rank` out of `world_size` (same scheme as gen_even_batches)."""
    size = N // world_size
    lo = rank * size
    hi = N if rank == world_size - 1 else lo + size
    return lo, hi
