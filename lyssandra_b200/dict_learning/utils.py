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


# ------------------------------------------------------------------ signals sharded over ranks (one process per GPU)
# The reference draws the initial atoms and the replacements of unused atoms from ALL the data
# (dict_learning/utils.py:55-70, ksvd.py:199-207).  With the signals sharded over ranks the same draws are made over the
# GLOBAL column index space — rank 0 consumes the NumPy RNG exactly as a single process holding the concatenated shards
# would — and every chosen column is fetched from the rank that owns it, so a sharded run picks the columns the
# single-GPU run on the gathered data picks.

def shard_offsets(dist, n_local):
    """offsets[r] .. offsets[r+1] = global column range of rank r's shard (shards concatenated in rank order)"""
    sizes = dist.allgather_object(int(n_local))
    return np.concatenate([[0], np.cumsum(np.asarray(sizes, dtype=np.int64))])


def global_candidates(dist, X, offsets):
    """global indices of the columns with sum(x^2) > 1e-6 (:55), ascending — the `idxs` of the reference on the
    concatenated data.  Only the (few) non-candidates travel."""
    energy = (X * X).sum(dim=0)
    bad_local = torch.nonzero(~(energy > 1e-6)).flatten().cpu().numpy().astype(np.int64) + int(offsets[dist.rank])
    bad = np.concatenate(dist.allgather_object(bad_local)) if dist.world > 1 else bad_local
    keep = np.ones(int(offsets[-1]), dtype=bool)
    keep[bad] = False
    return np.flatnonzero(keep)


def pick_columns_sharded(dist, X, cols_global, offsets):
    """(n, len(cols)) matrix of the global columns `cols_global`, identical on every rank: each rank copies the columns
    it owns into a zero matrix and the matrices are summed (x + 0 + ... + 0 is exact)."""
    cols_global = np.asarray(cols_global, dtype=np.int64)
    lo, hi = int(offsets[dist.rank]), int(offsets[dist.rank + 1])
    out = torch.zeros((X.shape[0], len(cols_global)), dtype=X.dtype, device=X.device)
    mine = np.flatnonzero((cols_global >= lo) & (cols_global < hi))
    if len(mine):
        src = torch.as_tensor(cols_global[mine] - lo, device=X.device)
        out[:, torch.as_tensor(mine, device=X.device)] = X.index_select(1, src)
    dist.allreduce_sum_(out)
    return out


def init_dictionary_sharded(dist, X, n_atoms, normalize_=None):
    """init_dictionary(method='data', return_unused_data=True) over the concatenation of every rank's shard `X`.
    -> (D replicated on every rank, unused global candidate indices (rank 0's copy is the one that is consumed),
    offsets).  `normalize_` defaults to engine.norm_cols_ (the device kernel of norm_cols, :65-66)."""
    offsets = shard_offsets(dist, X.shape[1])
    cands = global_candidates(dist, X, offsets)
    if len(cands) < n_atoms:
        raise ValueError("not enough datapoints to initialize the dictionary")
    chosen = None
    if dist.rank == 0:
        subset = np.random.choice(len(cands), size=n_atoms, replace=False)         # :61, the same RNG call
        chosen = cands[subset]
    chosen = np.asarray(dist.broadcast_object(chosen, src=0), dtype=np.int64)
    D = pick_columns_sharded(dist, X, chosen, offsets)
    (normalize_ or engine.norm_cols_)(D)
    taken = np.zeros(int(offsets[-1]), dtype=bool)
    taken[chosen] = True
    return D, cands[~taken[cands]], offsets


def replace_unused_atoms_sharded(dist, X, D, unused_atoms, unused_data, offsets, normalize_=None):
    """ksvd.py:199-207 with sharded signals: for every unused atom rank 0 draws `np.random.choice(len(unused_data))`
    (the reference's call), the drawn GLOBAL column is fetched from its owner, normalised and written into D on every
    rank.  Returns the shrunken unused_data (kept in step on every rank)."""
    plan = None
    if dist.rank == 0:
        plan = []
        for slot in unused_atoms:
            if len(unused_data) == 0:
                break
            pos = int(np.random.choice(len(unused_data), size=1)[0])
            plan.append((int(slot), int(unused_data[pos]), pos))
            unused_data = np.delete(unused_data, pos)
    plan = dist.broadcast_object(plan, src=0)
    if dist.rank != 0:
        for _, _, pos in plan:
            unused_data = np.delete(unused_data, pos)
    if plan:
        cols = pick_columns_sharded(dist, X, [c for _, c, _ in plan], offsets)
        for j, (slot, _, _) in enumerate(plan):
            one = cols[:, j:j + 1].contiguous()
            (normalize_ or engine.norm_cols_)(one)                                   # norm_cols of the single column, as :206
            D[:, slot:slot + 1] = one
    return unused_data
