"""l2_normalizer of the reference (lyssa/feature_extract/preproc.py:8-15): x / (||x||_2 + eps)."""
import numpy as np
import torch

from ..utils.math import normalize


class l2_normalizer(object):
    def __call__(self, Z):
        if torch.is_tensor(Z):
            flat = Z.reshape(-1)
            return (flat / (torch.linalg.vector_norm(flat) + np.finfo(float).eps)).reshape(Z.shape)
        Z = np.asarray(Z)
        return normalize(Z.flatten()).reshape(Z.shape)
