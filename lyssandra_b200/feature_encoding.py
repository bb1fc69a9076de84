"""Mirror of ``lyssa.feature_encoding`` (/root/reference/lyssa/feature_encoding.py) for the encoder that sits on the hot
path's correlation front end: ``soft_thresholding`` (:26-37) and ``feature_encoder('soft_thresholding')`` (:40-89) —
Coates & Ng's "soft threshold" features as the reference implements them, i.e. the n_nonzero_coefs largest SIGNED
correlations of every signal (the same computation as sparse_coding.thresholding, :416-425).  ``sign_splitting``
(:14-23) splits a code matrix into its positive and negative parts.  Everything runs on the GPU through
liblyssa_b200.so (lys_thresh_encode / lys_topk_select); there is no CPU path."""
from __future__ import annotations

import numpy as np
import torch

from . import engine


def _k(n_atoms, nonzero_percentage, n_nonzero_coefs):
    if nonzero_percentage is not None:
        n_nonzero_coefs = int(np.floor(nonzero_percentage * n_atoms))              # :31-32
    if n_nonzero_coefs is None:
        raise ValueError("soft_thresholding needs n_nonzero_coefs or nonzero_percentage")
    return int(n_nonzero_coefs)


def soft_thresholding(Alpha, nonzero_percentage=None, n_nonzero_coefs=None):
    """soft_thresholding(Alpha) -> Z (K, N): Alpha = D^T X (feature_encoding.py:26-37).  CUDA tensor in -> CUDA tensor out,
    NumPy in -> NumPy out."""
    numpy_in = not (torch.is_tensor(Alpha) and Alpha.is_cuda)
    A = engine.as_device_matrix(Alpha, None if numpy_in else Alpha.device, "Alpha")
    _, Z = engine.topk_select(A, _k(A.shape[0], nonzero_percentage, n_nonzero_coefs), dense=True)
    return Z.cpu().numpy() if numpy_in else Z


def sign_splitting(X, D, sparse_coder=None):
    """feature_encoding.py:14-23: Z = sparse_coder.encode(X, D) -> (2K, N): positive parts on top, negated negative parts below.
    (The reference indexes ``np.where(Z > 0)[0]`` twice and cannot run; this is what it sets out to compute.)"""
    Z = sparse_coder.encode(X, D)
    if torch.is_tensor(Z):
        return torch.cat([Z.clamp(min=0), (-Z).clamp(min=0)], dim=0)
    return np.concatenate([np.maximum(Z, 0), np.maximum(-Z, 0)], axis=0)


class feature_encoder(object):
    """Same constructor and methods as the reference class (feature_encoding.py:40-89); algorithm 'soft_thresholding'."""

    def __init__(self, algorithm=None, params=None, n_jobs=1, verbose=True, mmap=False):
        self.algorithm = algorithm
        self.params = params
        if self.params is None:
            self.params = {}
        self.n_jobs = n_jobs
        self.verbose = verbose
        self.mmap = mmap

    def encode(self, X, D):
        return self.__call__(X, D)

    def __call__(self, X, D):
        if self.algorithm != "soft_thresholding":
            # the reference falls through with `func` unbound (UnboundLocalError, :64-81); say what is wrong instead
            raise ValueError("feature_encoder: unknown algorithm %r (only 'soft_thresholding' exists)" % (self.algorithm,))
        numpy_in = not (torch.is_tensor(X) and X.is_cuda)
        Xd = engine.as_device_matrix(X, None if numpy_in else X.device)
        Dd = engine.as_dictionary(D, Xd.device)
        k = _k(Dd.shape[1], self.params.get("nonzero_percentage"), self.params.get("n_nonzero_coefs"))
        _, Z = engine.thresh_encode(Xd, Dd, k, dense=True)               # Alpha = D^T X (:66) + selection (:68-69), one kernel
        return Z.cpu().numpy() if numpy_in else Z

    def encode_sparse(self, X, D):
        Xd = engine.as_device_matrix(X, X.device if torch.is_tensor(X) and X.is_cuda else None)
        Dd = engine.as_dictionary(D, Xd.device)
        return engine.thresh_encode(Xd, Dd, _k(Dd.shape[1], self.params.get("nonzero_percentage"), self.params.get("n_nonzero_coefs")))
