"""lyssa/utils/math.py conventions for host-side glue (NumPy arrays or torch tensors):
``normalize`` x/(||x||+eps) (:61-62), ``norm_cols`` in place with +eps (:65-71), ``norm``
(:52-54), ``fast_dot`` (:11-24), ``frobenius_squared`` (:57-58).  eps = 2^-52 as in the
reference.  CUDA dictionaries are normalised by the library kernel (engine.norm_cols_)."""
from __future__ import annotations

import numpy as np
import torch

EPS = float(np.finfo(float).eps)


def fast_dot(a, b):
    if torch.is_tensor(a) or torch.is_tensor(b):
        return torch.matmul(a, b)
    return np.dot(a, b)


def norm(x):
    if torch.is_tensor(x):
        return torch.linalg.vector_norm(x)
    return float(np.sqrt(np.dot(x, x)))


def frobenius_squared(A):
    return (A * A).sum()


def normalize(x, eps=EPS):
    return x / (norm(x) + eps)


def norm_cols(X, eps=EPS):
    if torch.is_tensor(X):
        if X.is_cuda and X.dtype == torch.float32 and X.stride(1) == 1:
            from .. import engine
            return engine.norm_cols_(X)
        X /= (torch.sqrt((X * X).sum(dim=0)) + eps).unsqueeze(0)
        return X
    norms = np.sqrt(np.einsum("ij,ij->j", X, X)) + eps
    X /= norms[np.newaxis, :]
    return X
