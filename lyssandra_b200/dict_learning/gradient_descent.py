"""Host-side projected-gradient dictionary learner, mirrored from
lyssa/dict_learning/gradient_descent.py:18-160 only as far as the reference's single
Batch-OMP-touching test drives it (lyssa/dict_learning/tests/test_dictionary_learn.py:11-21):
the learner is NOT on the hot path — it is here so that the GPU coder is
exercised "from a host-side GD loop" the way the reference's own test does.  The dense
gradient products are plain torch matmuls (plumbing, outside the measured path)."""
from __future__ import annotations

from itertools import cycle

import torch

from .. import engine
from ..utils import gen_batches
from .utils import init_dictionary


def projected_grad_desc(X, n_atoms=None, sparse_coder=None, batch_size=None, D_init=None, eta=None, mu=None,
                        n_epochs=None, non_neg=False, verbose=False, n_jobs=1, mmap=False):
    if eta is None:
        raise ValueError("Must specify learning rate.")                            # :41-42
    sparse_coder.verbose = False
    Xd = engine.as_device_matrix(X, X.device if torch.is_tensor(X) and X.is_cuda else None)
    if D_init is None:
        D, _ = init_dictionary(Xd, n_atoms, method="data", return_unused_data=True)
    else:
        D = engine.as_dictionary(D_init, Xd.device)
    batch_idx = gen_batches(Xd.shape[1], batch_size=batch_size)
    n_iter = len(batch_idx)
    eye = torch.eye(n_atoms, device=Xd.device)
    for e in range(n_epochs):
        for i, batch in zip(range(n_iter), cycle(batch_idx)):
            Xb = Xd[:, batch.start:batch.stop]
            Zb = sparse_coder(Xb, D)                                               # :77
            grad = (D @ Zb - Xb) @ Zb.t()                                          # :86
            incoh = 2 * mu * (D @ (D.t() @ D - eye)) if (mu is not None and mu > 0) else 0   # :88-91
            D = D - eta * grad + incoh                                             # :94 (sign as in the reference)
            if non_neg:
                D = D.clamp_min(0)
            D = engine.norm_cols_(D.contiguous())                                  # :99
    return D


class dictionary_learner():
    """dictionary_learner(n_atoms, sparse_coder, batch_size, D_init, eta, mu, n_epochs, ...) with
    fit / encode / __call__ and .D, as gradient_descent.py:128-160."""

    def __init__(self, n_atoms=None, sparse_coder=None, batch_size=None, eta=None, mu=None, D_init=None,
                 n_epochs=1, verbose=False, memory="low", mmap=False, non_neg=False, n_jobs=1):
        self.n_atoms = n_atoms
        self.sparse_coder = sparse_coder
        self.batch_size = batch_size
        self.eta = eta
        self.mu = mu
        self.n_epochs = n_epochs
        self.D_init = D_init
        self.memory = memory
        self.verbose = verbose
        self.n_jobs = n_jobs
        self.mmap = mmap
        self.non_neg = non_neg
        self.D = None

    def fit(self, X):
        D = projected_grad_desc(X, n_atoms=self.n_atoms, sparse_coder=self.sparse_coder, batch_size=self.batch_size,
                                D_init=self.D_init, eta=self.eta, mu=self.mu, n_epochs=self.n_epochs,
                                non_neg=self.non_neg, verbose=self.verbose, n_jobs=self.n_jobs, mmap=self.mmap)
        self.D = D if (torch.is_tensor(X) and X.is_cuda) else D.cpu().numpy()

    def encode(self, X):
        return self.sparse_coder(X, self.D)

    def __call__(self, X):
        self.fit(X)
        return self.encode(X)
