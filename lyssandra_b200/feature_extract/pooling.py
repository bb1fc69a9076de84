"""Pooling operators of the reference (lyssa/feature_extract/pooling.py:4-26).

The classes keep the reference's names and call signature (``op(Z) -> pooled``, Z of shape
(n_atoms, n_patches)); on CUDA tensors they reduce with torch (plumbing for callers that use them
stand-alone).  Inside ``sc_spm_extractor`` they only select the mode of the fused pooling kernel
(``lys_spm_pool``): the dense Z they would reduce is never formed there."""
import numpy as np
import torch


class _pool_op(object):
    mode = None          # lys_spm_pool `pooling` argument


class sc_max_pooling(_pool_op):
    """max pooling on the absolute values of the sparse codes (pooling.py:4-7)"""
    mode = 0

    def __call__(self, Z):
        return Z.abs().max(dim=1).values if torch.is_tensor(Z) else np.max(np.abs(Z), axis=1)


class sum_pooling(_pool_op):
    """sum pooling (pooling.py:16-19)"""
    mode = 1

    def __call__(self, Z):
        return Z.sum(dim=1) if torch.is_tensor(Z) else np.sum(Z, axis=1)


class average_pooling(_pool_op):
    """average pooling over the descriptors of the cell (pooling.py:22-26)"""
    mode = 2

    def __call__(self, Z):
        return (Z.sum(dim=1) if torch.is_tensor(Z) else np.sum(Z, axis=1)) / float(Z.shape[1])


class max_pooling(_pool_op):
    """signed max pooling (pooling.py:10-13): not served by the fused kernel (it needs the zeros of
    the atoms a patch did not select); stand-alone use only."""
    mode = None

    def __call__(self, Z):
        return Z.max(dim=1).values if torch.is_tensor(Z) else np.max(Z, axis=1)
