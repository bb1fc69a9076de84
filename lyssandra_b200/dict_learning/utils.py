"""Host-side glue mirrored from lyssa/dict_learning/utils.py: ``approx_error`` (:14-19),
``average_mutual_coherence`` (:7-11), ``init_dictionary(method='data')`` (:35-75).

Random choices stay on the host and consume the GLOBAL NumPy RNG with the same calls as the
reference (quirk Q9), so a seeded run picks the same columns as the
oracle; only the gather / normalise runs on the device."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine


def approx_error(D, Z, X, n_jobs=1):
    """||X - D Z||_F^2 (dict_learning/utils.py:14-19).  Z may be engine.SparseCodes (no dense
    product) or a dense (K, N) array/tensor (densified codes are re-sparsified on the host
    side only for NumPy input — the device path takes SparseCodes)."""
    Xd = engine.as_device_matrix(X, X.device if torch.is_tensor(X) and X.is_cuda else None)
    Dd = engine.as_dictionary(D, Xd.device)
    if not isinstance(Z, engine.SparseCodes):
        Z = _codes_from_dense(Z, Xd.device)
    _, err = engine.residual(Xd, Dd, Z, want_residual=False, want_error=True)
    return float(err.item())


def _codes_from_dense(Z, device):
    Zt = torch.as_tensor(np.asarray(Z) if not torch.is_tensor(Z) else Z).to(device=device, dtype=torch.float32)
    K, N = Zt.shape
    nz = (Zt != 0)
    k = max(int(nz.sum(dim=0).max().item()), 1)
    # top-k by "is nonzero" keeps every nonzero; indices beyond a column's count are padded
    order = torch.argsort(nz.to(torch.int8), dim=0, descending=True, stable=True)[:k]        # (k, N)
    vals = torch.gather(Zt, 0, order)
    idx = torch.where(vals != 0, order, torch.full_like(order, -1)).t().contiguous().to(torch.int32)
    val = vals.t().contiguous()
    nsel = (idx >= 0).sum(dim=1).to(torch.int32)
    return engine.SparseCodes(idx, val, nsel, K)


def average_mutual_coherence(D):
    """mean off-diagonal |D^T D| (dict_learning/utils.py:7-11), from the Gram kernel."""
    Dd = engine.as_dictionary(D, D.device if torch.is_tensor(D) and D.is_cuda else None)
    G = engine.gram(Dd).abs()
    K = G.shape[0]
    return float((G.sum() - torch.diagonal(G).sum()).item() / float(K * (K - 1)))


def init_dictionary(X, n_atoms, method="data", return_unused_data=False, normalize=True):
    """method='data' (dict_learning/utils.py:49-70): candidates are the columns with
    sum(x^2) > 1e-6 (:55); ``np.random.choice(len(cands), n_atoms, replace=False)`` (:61);
    D = X[:, chosen] (a copy, :64); norm_cols (:65-66); the unused candidate indices are
    returned as a NumPy int64 array (the reference returns a Python list, :67-70).
    X may be NumPy (result NumPy, like the reference) or a CUDA tensor (result CUDA tensor)."""
    if method != "data":
        raise NotImplementedError("only method='data' is on the hot path (dict_learning/utils.py:49)")
    on_device = torch.is_tensor(X) and X.is_cuda
    if on_device:
        energy = (X * X).sum(dim=0)
        cands = torch.nonzero(energy > 1e-6).flatten().cpu().numpy()
    else:
        Xh = np.asarray(X)
        cands = np.flatnonzero(np.einsum("ij,ij->j", Xh, Xh) > 1e-6)
    if len(cands) < n_atoms:
        raise ValueError("not enough datapoints to initialize the dictionary")
    subset = np.random.choice(len(cands), size=n_atoms, replace=False)
    chosen = cands.astype(np.int64)[subset]
    if on_device:
        D = torch.empty((X.shape[0], n_atoms), dtype=torch.float32, device=X.device)
        engine.gather_cols_(X, chosen, D)
        if normalize:
            engine.norm_cols_(D)
    else:
        D = np.array(Xh[:, chosen])
        if normalize:
            from ..utils.math import norm_cols
            D = norm_cols(D.astype(np.float64) if D.dtype.kind != "f" else D)
    if return_unused_data:
        taken = np.zeros(X.shape[1], dtype=bool)
        taken[chosen] = True
        unused = cands[~taken[cands]]
        return D, unused
    return D
