"""Mairal online dictionary learning on the GPU behind the reference's API
(/root/reference/lyssa/dict_learning/online_dict_learn.py): ``online_dict_learn`` (:18-124)
and ``online_dictionary_coder`` (:127-160).

Per minibatch (:79-98): Batch-OMP encode -> sparse sufficient statistics
A = beta*A + Z Z^T, B = beta*B + X Z^T -> one GEMM D A with a fused Jacobi column update,
optional clamp and column normalisation.  Mirrored quirks: beta=None is
linspace(0,1,n_iter) restarted every epoch, so the first minibatch of every epoch wipes A, B;
a CUDA-tensor D_init is used without copying and updated in place as in the reference (a NumPy D_init is uploaded, so
the caller's array is NOT mutated — take the returned D); the end-of-epoch patience rule is reproduced as written."""
from __future__ import annotations

from itertools import cycle

import numpy as np
import torch

from .. import engine
from ..utils import gen_batches
from .utils import init_dictionary, init_dictionary_sharded


def _gather_minibatch(dist, Xb, codes):
    """all-gather this rank's slice of the minibatch — signals and sparse codes, (n + 2k) words per signal instead of
    the K x K + n x K statistics an all-reduce would move — so that every rank accumulates the WHOLE minibatch in the
    same fixed order and ends with bit-identical A, B and D."""
    import torch.distributed as td
    b_loc, k = codes.idx.shape
    n = Xb.shape[0]
    w = dist.world
    Xs = Xb.t().contiguous()                                                        # (b_loc, n) signal-major
    Xall = torch.empty((w * b_loc, n), dtype=torch.float32, device=Xs.device)
    iall = torch.empty((w * b_loc, k), dtype=torch.int32, device=Xs.device)
    vall = torch.empty((w * b_loc, k), dtype=torch.float32, device=Xs.device)
    td.all_gather_into_tensor(Xall, Xs, group=dist.group)
    td.all_gather_into_tensor(iall, codes.idx.contiguous(), group=dist.group)
    td.all_gather_into_tensor(vall, codes.val.contiguous(), group=dist.group)
    return Xall.t(), engine.SparseCodes(iall, vall, (iall >= 0).sum(dim=1).to(torch.int32), codes.n_atoms)


def online_dict_learn(X, n_atoms, sparse_coder=None, batch_size=None, A=None, B=None, D_init=None,
                      beta=None, n_epochs=1, verbose=False, n_jobs=1, non_neg=False, mmap=False,
                      dist=None):
    """-> (D, A, B).  CUDA tensors in -> CUDA tensors out; NumPy in -> NumPy out.

    Multi-GPU (one process per GPU): pass ``dist`` (distributed.DistContext).  X is then THIS rank's slice of every
    minibatch (``batch_size`` = signals per rank and minibatch, the same on every rank): the slices are encoded
    independently, all-gathered as (signal, idx, val), and the statistics and the dictionary update run replicated and
    order-deterministic on every rank — exact, because the block update is Jacobi (:91-94), and bit-identical to the
    single-GPU run on the gathered minibatch."""
    sparse_coder.verbose = False                                                   # :41
    numpy_in = not (torch.is_tensor(X) and X.is_cuda)
    Xd = engine.as_device_matrix(X, None if numpy_in else X.device)
    dev = Xd.device
    n_features, n_samples = Xd.shape
    if D_init is None:
        if dist is not None and dist.world > 1:          # one dictionary for all ranks, drawn over every rank's slice
            D, _unused, _ = init_dictionary_sharded(dist, Xd, n_atoms)
        else:
            D, _unused = init_dictionary(Xd, n_atoms, method="data", return_unused_data=True)   # :44-45
    else:
        D = engine.as_dictionary(D_init, dev)                                      # :47 (no copy)
    batch_idx = gen_batches(n_samples, batch_size=batch_size)                      # :52
    n_iter = len(batch_idx)
    if A is None and B is None:                                                    # :61-63
        A = torch.zeros((n_atoms, n_atoms), dtype=torch.float32, device=dev)
        B = torch.zeros((n_features, n_atoms), dtype=torch.float32, device=dev)
    else:
        A = engine.as_device_matrix(A, dev, "A").contiguous()
        B = engine.as_device_matrix(B, dev, "B").contiguous()
    if beta is None:                                                               # :65-69
        beta_seq = np.linspace(0, 1, num=n_iter)
    else:
        beta_seq = np.zeros(n_iter) + beta

    multi = dist is not None and dist.world > 1
    max_patience = 10
    error_curr = 0
    error_prev = 0
    patience = 0
    for e in range(n_epochs):
        for i, batch in zip(range(n_iter), cycle(batch_idx)):                      # :79
            Xb = Xd[:, batch.start:batch.stop]
            codes = sparse_coder.encode_sparse(Xb, D)                              # :82
            if multi:
                Xb, codes = _gather_minibatch(dist, Xb, codes)
            engine.odl_accumulate_(Xb, codes, beta_seq[i], A, B)                   # :84-85
            engine.odl_update_dict_(D, A, B, non_neg=non_neg)                      # :91-98
        if e < n_epochs - 1:                                                       # :101-118
            if patience >= max_patience:
                break
            error_curr = 0
            for i, batch in zip(range(n_iter), cycle(batch_idx)):
                Xb = Xd[:, batch.start:batch.stop]
                codes = sparse_coder.encode_sparse(Xb, D)
                _, err = engine.residual(Xb, D, codes, want_residual=False, want_error=True)
                if multi:
                    dist.allreduce_sum_(err)
                error_curr += float(err.item())
            if verbose:
                print("end of epoch %d: error %.6g (difference %.6g)" % (e, error_curr, error_curr - error_prev))
                error_prev = error_curr
            if (e > 0) and (error_curr > 0.9 * error_prev or error_curr > error_prev):
                patience += 1
    if numpy_in:
        return D.cpu().numpy(), A.cpu().numpy(), B.cpu().numpy()
    return D, A, B


class online_dictionary_coder():
    """Same constructor and methods as the reference wrapper (online_dict_learn.py:127-160);
    A, B, D are plain tensors/arrays the caller can save for a warm start (:139-140,:153)."""

    def __init__(self, n_atoms=None, sparse_coder=None, batch_size=None, beta=None, D_init=None,
                 n_epochs=1, verbose=False, memory="low", mmap=False, non_neg=False, n_jobs=1):
        self.n_atoms = n_atoms
        self.sparse_coder = sparse_coder
        self.batch_size = batch_size
        self.beta = beta
        self.n_epochs = n_epochs
        self.A = None
        self.B = None
        self.D_init = D_init
        self.memory = memory
        self.verbose = verbose
        self.n_jobs = n_jobs
        self.mmap = mmap
        self.non_neg = non_neg

    def __call__(self, X):
        self.fit(X)
        return self.encode(X)

    def fit(self, X):
        self.D, self.A, self.B = online_dict_learn(X, self.n_atoms, sparse_coder=self.sparse_coder,
                                                   batch_size=self.batch_size, A=self.A, B=self.B,
                                                   D_init=self.D_init, beta=self.beta, n_epochs=self.n_epochs,
                                                   verbose=self.verbose, n_jobs=self.n_jobs,
                                                   non_neg=self.non_neg, mmap=self.mmap)

    def encode(self, X):
        return self.sparse_coder(X, self.D)
