"""Class-specific K-SVD behind the reference's API (/root/reference/lyssa/dict_learning/class_dict_learn.py):
``class_dict_learn`` (:98-139) and ``class_ksvd_coder`` (:16-95) — a host loop over ``ksvd_dict_learn``, one
dictionary per class written side by side into a joint (n, sum(n_class_atoms)) dictionary.

Mirrored as written, quirks included:
  * the reference's ``return D`` sits INSIDE the class loop (:139), so it returns after class 0 with the other
    classes' columns still zero.  ``all_classes=False`` (the default) reproduces that; ``all_classes=True`` trains
    every class, which is what the module's docstring describes.
  * the block of class c starts at ``c * n_class_atoms[c]`` (:126), not at the sum of the previous sizes.
  * ``alpha`` (structural incoherence) imports a module that does not exist in the reference
    (``lyssa.dict_learn.utils``, :130): it raises there and is rejected here.
LC-KSVD (lc_ksvd.py:105-216) runs the exact atom update on signals stacked with the label matrices
(n + n_atoms + n_classes features), beyond the n <= 64 the device kernel is built for; it is not provided."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine
from .ksvd import ksvd_dict_learn


def class_dict_learn(X, y, n_class_atoms=None, sparse_coders=None, init_dict="data", max_iter=5, approx=False,
                     non_neg=False, eta=None, alpha=None, n_cycles=1, n_jobs=1, mmap=False, verbose=True,
                     all_classes=False):
    """-> joint dictionary D (n, sum(n_class_atoms)); NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out."""
    if alpha is not None:
        raise NotImplementedError("alpha (replace_coherent_atoms) does not exist in the reference either (class_dict_learn.py:130)")
    y = np.asarray(y.cpu() if torch.is_tensor(y) else y)
    n_classes = len(set(y.tolist()))                                               # :101
    n_total_atoms = int(np.sum(n_class_atoms))                                     # :103
    numpy_in = not (torch.is_tensor(X) and X.is_cuda)
    Xd = engine.as_device_matrix(X, None if numpy_in else X.device)
    D = torch.zeros((Xd.shape[0], n_total_atoms), dtype=torch.float32, device=Xd.device)   # :109
    for c in range(n_classes):
        if verbose:
            print("-------------------------------------")
            print("optimizing the dictionary of class", c)
        cols = torch.as_tensor(np.flatnonzero(y == c), device=Xd.device)          # :115-117
        Xc = Xd.index_select(1, cols)
        Dc, _ = ksvd_dict_learn(Xc, n_class_atoms[c], init_dict="data", sparse_coder=sparse_coders[c],   # :120-122
                                max_iter=max_iter, non_neg=non_neg, approx=approx, eta=eta, n_cycles=n_cycles,
                                n_jobs=n_jobs, mmap=mmap, verbose=verbose, return_codes=True)
        base = c * n_class_atoms[c]                                                # :124
        D[:, base:base + n_class_atoms[c]] = Dc                                    # :126
        if not all_classes:                                                        # :139, `return D` inside the loop
            break
    return D.cpu().numpy() if numpy_in else D


class class_ksvd_coder():
    """Same constructor and methods as the reference wrapper (class_dict_learn.py:16-95)."""

    def __init__(self, n_class_atoms=None, n_nonzero_coefs=None, atom_ratio=None, coef_ratio=None, sparse_coder=None,
                 non_neg=False, max_iter=None, approx=False, eta=None, alpha=None, n_cycles=1, n_jobs=1, mmap=False,
                 verbose=True, all_classes=False):
        self.sparse_coder = sparse_coder
        self.n_class_atoms = n_class_atoms
        self.n_nonzero_coefs = n_nonzero_coefs
        self.eta = eta
        self.alpha = alpha
        self.non_neg = non_neg
        self.max_iter = max_iter
        self.approx = approx
        self.atom_ratio = atom_ratio
        self.coef_ratio = coef_ratio
        self.n_cycles = n_cycles
        self.n_jobs = n_jobs
        self.mmap = mmap
        self.verbose = verbose
        self.all_classes = all_classes
        self.D = None

    def _fit(self, X, y):
        yh = np.asarray(y.cpu() if torch.is_tensor(y) else y)
        n_classes = len(set(yh.tolist()))
        if self.n_class_atoms is None:                                             # :58-63
            self.n_class_atoms = [int(np.sum(yh == c) * self.atom_ratio) for c in range(n_classes)]
        if self.n_nonzero_coefs is None and self.coef_ratio is not None:           # :65-69 (recorded, never used: the coder carries k)
            self.n_nonzero_coefs = [int(self.n_class_atoms[c] * self.coef_ratio) for c in range(n_classes)]
        if not isinstance(self.n_class_atoms, list):                               # :71-73
            self.n_class_atoms = [self.n_class_atoms for _ in range(n_classes)]
        sparse_coders = [self.sparse_coder for _ in range(n_classes)]              # :75
        self.D = class_dict_learn(X, yh, n_class_atoms=self.n_class_atoms, sparse_coders=sparse_coders,   # :77-81
                                  init_dict="data", max_iter=self.max_iter, non_neg=self.non_neg, approx=self.approx,
                                  eta=self.eta, alpha=self.alpha, n_cycles=self.n_cycles, n_jobs=self.n_jobs,
                                  mmap=self.mmap, verbose=self.verbose, all_classes=self.all_classes)

    def __call__(self, X, y):
        self._fit(X, y)
        return self.D

    def fit(self, X, y):
        self._fit(X, y)

    def encode(self, X):
        return self.sparse_coder.encode(X, self.D)

    def print_params(self):
        return
