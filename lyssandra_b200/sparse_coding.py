"""Mirror of ``lyssa.sparse_coding.sparse_encoder`` for the Batch-OMP hot path
(/root/reference/lyssa/sparse_coding.py:587-603 constructor + encode/__call__,
:629-635 the 'bomp' branch, :706 the unknown-algorithm error, :708-726 dispatch) and the two
thresholding coders that share its correlation front end (:636-641 'thresh', :671-690 'iht').

``algorithm`` 'bomp', 'omp' (:618-625 -> :19-66, the reference's default), 'thresh' and 'iht' run — on the GPU,
through liblyssa_b200.so.  The reference's other coders are outside this engine's scope and raise
NotImplementedError rather than silently running on the CPU; an unknown name raises ValueError
exactly like the reference.  The public attributes (algorithm, params, n_jobs, verbose, mmap, name) are plain
and mutable because reference callers mutate them (ksvd.py:159, online_dict_learn.py:41).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import engine

_REFERENCE_ALGORITHMS = ("omp", "bomp", "thresh", "nnomp", "group_omp", "sparse_group_omp",
                         "somp", "iht", "lasso", "llc")


class sparse_encoder(object):
    """sparse_encoder(algorithm, params, n_jobs, verbose, mmap, name) — same signature and
    defaults as the reference (sparse_coding.py:587).

    encode(X, D) / __call__(X, D): X (n_features, n_samples), D (n_features, n_atoms) ->
    Z (n_atoms, n_samples), float32.
      * torch CUDA tensors in  -> torch CUDA tensor out (a transposed view of signal-major
        storage; no host round trip);
      * NumPy arrays / CPU tensors in -> NumPy array out, through the library's overlapped
        host pipeline.  ``n_jobs`` > 1 (or -1 = all, sparse_coding.py:594-595) spreads
        contiguous column blocks over that many visible GPUs, the analogue of run_parallel's
        contiguous batches (lyssa/utils/__init__.py:93-97,166-180).
    encode_sparse(X, D) returns engine.SparseCodes (idx, val, nsel) without densifying — what
    the learners in lyssandra_b200.dict_learning consume.
    """

    def __init__(self, algorithm="omp", params=None, n_jobs=1, verbose=True, mmap=False, name="sparse_coder"):
        self.name = name
        self.algorithm = algorithm
        self.params = params
        if self.params is None:
            self.params = {}
        if n_jobs == -1:
            n_jobs = max(torch.cuda.device_count(), 1)
        self.n_jobs = n_jobs
        self.verbose = verbose
        self.mmap = mmap

    # ------------------------------------------------------------------ reference surface
    def encode(self, X, D):
        return self.__call__(X, D)

    def __call__(self, X, D):
        k = self._check(D.shape[1])
        if torch.is_tensor(X) and X.is_cuda:
            Dd = engine.as_dictionary(D, X.device)
            Xd = engine.as_device_matrix(X, X.device)
            return self._encode_device(Xd, Dd, k, dense=True)[1]
        Xh, Dh = self._host_arrays(X, D)
        if self.algorithm != "bomp":
            Xd = engine.as_device_matrix(Xh, None)
            Dd = engine.as_dictionary(Dh, Xd.device)
            return self._encode_device(Xd, Dd, k, dense=True)[1].cpu().numpy()
        return self._encode_host(Xh, Dh, k, dense=True)[3]

    def _encode_device(self, Xd, Dd, k, dense, G=None):
        if self.algorithm == "omp":
            out = engine.omp_encode(Xd, Dd, k, tol=self.params.get("tol"), G=G, dense=dense)
        elif self.algorithm == "thresh":
            out = engine.thresh_encode(Xd, Dd, k, dense=dense)
        elif self.algorithm == "iht":
            out = engine.iht_encode(Xd, Dd, k, self.params.get("eta"), self.params.get("n_iter"), dense=dense)
        else:
            out = engine.bomp_encode(Xd, Dd, k, G=G, dense=dense)
        return out if dense else (out, None)

    # ------------------------------------------------------------------------ extensions
    def encode_sparse(self, X, D, G=None):
        """Device-resident sparse codes; X/D may be NumPy (uploaded) or CUDA tensors."""
        k = self._check(D.shape[1])
        Xd = engine.as_device_matrix(X, X.device if torch.is_tensor(X) and X.is_cuda else None)
        Dd = engine.as_dictionary(D, Xd.device)
        return self._encode_device(Xd, Dd, k, dense=False, G=G)[0]

    def encode_sparse_host(self, X, D):
        """NumPy in, NumPy (idx, val, nsel) out — no dense Z crosses PCIe."""
        k = self._check(np.shape(D)[1])
        Xh, Dh = self._host_arrays(X, D)
        if self.algorithm != "bomp":
            codes = self.encode_sparse(Xh, Dh)
            return codes.idx.cpu().numpy(), codes.val.cpu().numpy(), codes.nsel.cpu().numpy()
        idx, val, nsel, _ = self._encode_host(Xh, Dh, k, dense=False)
        return idx, val, nsel

    # --------------------------------------------------------------------------- helpers
    def _check(self, n_atoms=None):
        alg = self.algorithm
        if alg not in ("bomp", "omp", "thresh", "iht"):
            if alg in _REFERENCE_ALGORITHMS:
                raise NotImplementedError(
                    "algorithm %r is outside the B200 engine's scope: 'bomp', 'omp', 'thresh' and 'iht' are "
                    "implemented (no CPU fallback by design)" % (alg,))
            raise ValueError("Sparse optimizer not found.")          # sparse_coding.py:706
        k = self.params.get("n_nonzero_coefs")
        if alg == "omp":                                             # :27-34: n_nonzero_coefs, or tol alone
            if k is None and self.params.get("tol") is None:
                raise ValueError("algorithm 'omp' needs params['n_nonzero_coefs'] or params['tol']")
            return None if k is None else int(k)
        pct = self.params.get("nonzero_percentage")
        if alg != "bomp" and pct is not None:
            if n_atoms is None:
                raise ValueError("nonzero_percentage needs the dictionary size")
            kp = int(np.floor(pct * n_atoms))                        # sparse_coding.py:419-420
            if alg == "iht" and k is not None and int(k) != kp:
                # the reference thresholds the start with floor(p*K) and the iterations with
                # n_nonzero_coefs (:676-690); the engine runs one sparsity throughout
                raise ValueError("'iht': nonzero_percentage and n_nonzero_coefs disagree (%d vs %d)" % (kp, int(k)))
            k = kp
        if k is None:
            # the reference crashes inside np.zeros((None, None)) (sparse_coding.py:317, quirk Q2)
            raise ValueError("params['n_nonzero_coefs'] must be set for algorithm %r" % (alg,))
        if alg == "iht" and (self.params.get("eta") is None or self.params.get("n_iter") is None):
            raise ValueError("params['eta'] and params['n_iter'] must be set for algorithm 'iht'")
        return int(k)

    @staticmethod
    def _host_arrays(X, D):
        Xh = X.detach().cpu().numpy() if torch.is_tensor(X) else np.asarray(X)
        Dh = D.detach().cpu().numpy() if torch.is_tensor(D) else np.asarray(D)
        if Xh.ndim != 2 or Dh.ndim != 2:
            raise ValueError("X and D must be 2-D (features x columns)")
        return Xh, Dh

    def _encode_host(self, Xh, Dh, k, dense):
        n_gpus = min(int(self.n_jobs), max(torch.cuda.device_count(), 1)) if self.n_jobs and self.n_jobs > 1 else 1
        N = Xh.shape[1]
        if n_gpus <= 1 or N < 2 * n_gpus:
            return engine.bomp_encode_host(Xh, Dh, k, dense=dense)
        K = Dh.shape[1]
        bounds = np.linspace(0, N, n_gpus + 1).astype(np.int64)
        Xh = np.ascontiguousarray(Xh, dtype=np.float32) if not (Xh.strides[0] == 4 or Xh.strides[1] == 4) else Xh

        def work(g):
            lo, hi = int(bounds[g]), int(bounds[g + 1])
            return engine.bomp_encode_host(Xh[:, lo:hi], Dh, k, dense=dense, device=g)

        with ThreadPoolExecutor(max_workers=n_gpus) as ex:
            parts = list(ex.map(work, range(n_gpus)))
        idx = np.concatenate([p[0] for p in parts], axis=0)
        val = np.concatenate([p[1] for p in parts], axis=0)
        nsel = np.concatenate([p[2] for p in parts], axis=0)
        Z = np.concatenate([p[3] for p in parts], axis=1) if dense else None
        return idx, val, nsel, Z
