"""Torch-facing wrappers over the C-ABI (include/lyssa_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every numerical step of the hot
path is a kernel of liblyssa_b200.so.  Nothing in this module computes on the CPU and nothing
falls back to torch ops when the library is missing (``_native.load()`` raises).

Shapes follow the reference ("datapoints in columns", lyssa/sparse_coding.py:604):
X is (n_features, n_signals), D is (n_features, n_atoms), dense codes Z are
(n_atoms, n_signals).  Any strides are accepted for X (the kernels take element strides);
D must have contiguous rows (atom index fastest), which is the reference's C-order layout.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _native as nat


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("lyssandra_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_workspaces = {}


def workspace(device, nbytes, tag="ws"):
    """Grow-only per-device scratch tensors (the library never allocates for the caller)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def as_device_matrix(A, device=None, name="X"):
    """numpy / CPU tensor / CUDA tensor -> float32 CUDA tensor (no copy when already there)."""
    _require_cuda()
    if isinstance(A, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(A, dtype=np.float32) if A.dtype != np.float32 else A)
    elif torch.is_tensor(A):
        t = A
    else:
        raise TypeError("%s must be a numpy array or a torch tensor, got %r" % (name, type(A)))
    if t.dim() != 2:
        raise ValueError("%s must be 2-D (features x columns), got shape %s" % (name, tuple(t.shape)))
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if t.dtype != torch.float32 or not t.is_cuda or t.device != device:
        t = t.to(device=device, dtype=torch.float32)
    return t


def as_dictionary(D, device):
    D = as_device_matrix(D, device, "D")
    if D.stride(1) != 1 or D.stride(0) < D.shape[1]:
        D = D.contiguous()
    return D


@dataclass
class SparseCodes:
    """(idx,val)[N][k] codes in selection order; idx = -1 / val = 0 pads signals that stopped
    early (lyssa/sparse_coding.py:323-325,:335,:345)."""
    idx: torch.Tensor      # (N, k) int32
    val: torch.Tensor      # (N, k) float32
    nsel: torch.Tensor     # (N,)  int32
    n_atoms: int

    @property
    def n_signals(self):
        return self.idx.shape[0]

    @property
    def k(self):
        return self.idx.shape[1]

    def to_dense(self):
        """Dense Z of logical shape (n_atoms, n_signals) — a transposed view of a
        signal-major (N, K) buffer (zero-fill + scatter on the device)."""
        return codes_to_dense(self)


def gram(D):
    """G = D^T D  (lyssa/sparse_coding.py:630)."""
    lib = nat.load()
    n, K = D.shape
    G = torch.empty((K, K), dtype=torch.float32, device=D.device)
    with torch.cuda.device(D.device):
        ws = workspace(D.device, lib.lys_gram_workspace_bytes(n, K), tag="gram")
        nat.check(lib.lys_gram_ws(_ptr(D), D.stride(0), n, K, _ptr(G), _ptr(ws), ws.numel(), _stream_ptr(D.device)))
    return G


def bomp_encode(X, D, k, G=None, dense=False, screen=False):
    """Batch-OMP of the columns of X over D (device tensors).  Returns SparseCodes and, if
    ``dense``, also Z as a (K, N) transposed view (lyssa/sparse_coding.py:629-635,:302-367).
    ``screen`` selects the one-product screened correlations of the fused kernel (LYS_BOMP_SCREEN, same codes; A/B checks)."""
    lib = nat.load()
    n, N = X.shape
    n2, K = D.shape
    if n != n2:
        raise ValueError("X has %d features but D has %d" % (n, n2))
    if k is None:
        raise ValueError("params['n_nonzero_coefs'] must be set for algorithm 'bomp'")
    k = int(k)
    dev = X.device
    with torch.cuda.device(dev):
        if G is None:
            G = gram(D)
        idx = torch.empty((N, k), dtype=torch.int32, device=dev)
        val = torch.empty((N, k), dtype=torch.float32, device=dev)
        nsel = torch.empty((N,), dtype=torch.int32, device=dev)
        Zt = torch.empty((N, K), dtype=torch.float32, device=dev) if dense else None
        wsb = lib.lys_bomp_workspace_bytes(n, K, N, k)
        ws = workspace(dev, wsb)
        nat.check(lib.lys_bomp_encode_ex(
            _ptr(X), X.stride(0), X.stride(1), _ptr(D), D.stride(0), _ptr(G), n, K, N, k,
            _ptr(idx), _ptr(val), _ptr(nsel), _ptr(Zt), 1, K, _ptr(ws), ws.numel(),
            nat.BOMP_SCREEN if screen else 0, _stream_ptr(dev)))
    codes = SparseCodes(idx, val, nsel, K)
    if dense:
        return codes, Zt.t()
    return codes


def omp_encode(X, D, k=None, tol=None, G=None, dense=False):
    """The reference's plain `omp` coder (lyssa/sparse_coding.py:618-625 -> :19-66) on device tensors.
    ``k`` = n_nonzero_coefs (then the reference forces tol = 1e-10 and stops at k atoms) or, with ``k`` None,
    ``tol`` alone (continue while ||r|| >= tol; at most nat.OMP_MAX_NONZERO atoms per signal are supported and a signal
    that needs more raises).  Returns like bomp_encode."""
    lib = nat.load()
    n, N = X.shape
    n2, K = D.shape
    if n != n2:
        raise ValueError("X has %d features but D has %d" % (n, n2))
    if k is None and tol is None:
        raise ValueError("algorithm 'omp' needs params['n_nonzero_coefs'] or params['tol']")
    strict = k is not None                                             # sparse_coding.py:27-34
    kmax = int(k) if strict else min(K, n, nat.OMP_MAX_NONZERO)
    tol = 1e-10 if strict else float(tol)
    dev = X.device
    with torch.cuda.device(dev):
        if G is None:
            G = gram(D)
        idx = torch.empty((N, kmax), dtype=torch.int32, device=dev)
        val = torch.empty((N, kmax), dtype=torch.float32, device=dev)
        nsel = torch.empty((N,), dtype=torch.int32, device=dev)
        Zt = torch.empty((N, K), dtype=torch.float32, device=dev) if dense else None
        trunc = torch.zeros((1,), dtype=torch.int32, device=dev)
        ws = workspace(dev, lib.lys_omp_workspace_bytes(n, K, N, kmax))
        nat.check(lib.lys_omp_encode(
            _ptr(X), X.stride(0), X.stride(1), _ptr(D), D.stride(0), _ptr(G), n, K, N, kmax, tol, 1 if strict else 0,
            _ptr(idx), _ptr(val), _ptr(nsel), _ptr(Zt), 1, K, _ptr(trunc), _ptr(ws), ws.numel(), _stream_ptr(dev)))
        if not strict and int(trunc.item()) > 0:
            raise NotImplementedError("'omp' with tol=%g: %d signal(s) need more than %d atoms (the engine's limit per signal)"
                                      % (tol, int(trunc.item()), kmax))
    codes = SparseCodes(idx, val, nsel, K)
    if dense:
        return codes, Zt.t()
    return codes


def topk_select(Alpha, k, dense=True):
    """thresholding / soft_thresholding on a given correlation matrix Alpha (K, N) (lyssa/sparse_coding.py:416-425,
    lyssa/feature_encoding.py:26-37): the k largest SIGNED entries of every column."""
    lib = nat.load()
    K, N = Alpha.shape
    At = Alpha.t().contiguous()                                        # (N, K) signal-major, what the selection kernels read
    dev = Alpha.device
    k = int(k)
    with torch.cuda.device(dev):
        idx = torch.empty((N, k), dtype=torch.int32, device=dev)
        val = torch.empty((N, k), dtype=torch.float32, device=dev)
        nsel = torch.empty((N,), dtype=torch.int32, device=dev)
        Zt = torch.empty((N, K), dtype=torch.float32, device=dev) if dense else None
        nat.check(lib.lys_topk_select(_ptr(At), K, N, k, _ptr(idx), _ptr(val), _ptr(nsel), _ptr(Zt), 1, K, _stream_ptr(dev)))
    codes = SparseCodes(idx, val, nsel, K)
    if dense:
        return codes, Zt.t()
    return codes


def _threshold_encode(kind, X, D, k, eta, n_iter, dense):
    lib = nat.load()
    n, N = X.shape
    n2, K = D.shape
    if n != n2:
        raise ValueError("X has %d features but D has %d" % (n, n2))
    if k is None:
        raise ValueError("params['n_nonzero_coefs'] (or 'nonzero_percentage') must be set for algorithm %r" % kind)
    k = int(k)
    dev = X.device
    with torch.cuda.device(dev):
        idx = torch.empty((N, k), dtype=torch.int32, device=dev)
        val = torch.empty((N, k), dtype=torch.float32, device=dev)
        nsel = torch.empty((N,), dtype=torch.int32, device=dev)
        Zt = torch.empty((N, K), dtype=torch.float32, device=dev) if dense else None
        if kind == "thresh":
            ws = workspace(dev, lib.lys_thresh_workspace_bytes(n, K, N))
            nat.check(lib.lys_thresh_encode(
                _ptr(X), X.stride(0), X.stride(1), _ptr(D), D.stride(0), n, K, N, k,
                _ptr(idx), _ptr(val), _ptr(nsel), _ptr(Zt), 1, K, _ptr(ws), ws.numel(), _stream_ptr(dev)))
        else:
            ws = workspace(dev, lib.lys_iht_workspace_bytes(n, K, N))
            nat.check(lib.lys_iht_encode(
                _ptr(X), X.stride(0), X.stride(1), _ptr(D), D.stride(0), n, K, N, k, float(eta), int(n_iter),
                _ptr(idx), _ptr(val), _ptr(nsel), _ptr(Zt), 1, K, _ptr(ws), ws.numel(), _stream_ptr(dev)))
    codes = SparseCodes(idx, val, nsel, K)
    if dense:
        return codes, Zt.t()
    return codes


def thresh_encode(X, D, k, dense=False):
    """Alpha = D^T X, keep the k largest signed correlations per signal
    (lyssa/sparse_coding.py:636-641,:416-425).  Device tensors; returns like bomp_encode."""
    return _threshold_encode("thresh", X, D, k, 0.0, 0, dense)


def iht_encode(X, D, k, eta, n_iter, dense=False):
    """Iterative hard thresholding started from thresh_encode (lyssa/sparse_coding.py:671-690,
    :433-446)."""
    return _threshold_encode("iht", X, D, k, eta, n_iter, dense)


def bomp_encode_host(X, D, k, dense=True, device=None, want_codes=True):
    """Same for HOST arrays (numpy float32): chunked, copy/compute-overlapped pipeline inside the
    library.  Returns (idx, val, nsel, Z) as numpy arrays (Z a transposed view, or None)."""
    _require_cuda()
    lib = nat.load()
    n, N = X.shape
    n2, K = D.shape
    if n != n2:
        raise ValueError("X has %d features but D has %d" % (n, n2))
    if k is None:
        raise ValueError("params['n_nonzero_coefs'] must be set for algorithm 'bomp'")
    k = int(k)
    if X.dtype != np.float32:
        X = X.astype(np.float32)
    D = np.ascontiguousarray(D, dtype=np.float32)
    es = X.itemsize
    xfs, xss = X.strides[0] // es, X.strides[1] // es
    if not ((xfs == 1 and xss >= n) or (xss == 1 and xfs >= N)) or N <= 1:
        X = np.ascontiguousarray(X)
        xfs, xss = N, 1
    idx = np.empty((N, k), dtype=np.int32) if want_codes else None
    val = np.empty((N, k), dtype=np.float32) if want_codes else None
    nsel = np.empty((N,), dtype=np.int32) if want_codes else None
    Zt = np.empty((N, K), dtype=np.float32) if dense else None

    def p(a):
        return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)

    dev = -1 if device is None else int(device)
    nat.check(lib.lys_bomp_encode_host(p(X), xfs, xss, p(D), K, n, K, N, k,
                                       p(idx), p(val), p(nsel), p(Zt), 1, K, dev))
    return idx, val, nsel, (Zt.T if dense else None)


def codes_to_dense(codes: SparseCodes):
    lib = nat.load()
    N, k = codes.idx.shape
    K = codes.n_atoms
    dev = codes.idx.device
    Zt = torch.empty((N, K), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.lys_codes_to_dense(_ptr(codes.idx), _ptr(codes.val), N, k, K, _ptr(Zt), 1, K, _stream_ptr(dev)))
    return Zt.t()


def residual(X, D, codes: SparseCodes, want_residual=True, want_error=True):
    """R = X - D Z as a signal-major (N, n) tensor and/or ||R||_F^2 as a 1-element float64
    device tensor (lyssa/dict_learning/ksvd.py:103; lyssa/dict_learning/utils.py:14-19)."""
    lib = nat.load()
    n, N = X.shape
    K = D.shape[1]
    dev = X.device
    R = torch.empty((N, n), dtype=torch.float32, device=dev) if want_residual else None
    err = torch.zeros((1,), dtype=torch.float64, device=dev) if want_error else None
    with torch.cuda.device(dev):
        ws = workspace(dev, lib.lys_residual_workspace_bytes(n, K, N))
        nat.check(lib.lys_residual(_ptr(X), X.stride(0), X.stride(1), _ptr(D), D.stride(0),
                                   _ptr(codes.idx), _ptr(codes.val), n, K, N, codes.k,
                                   _ptr(R), _ptr(err), _ptr(ws), ws.numel(), _stream_ptr(dev)))
    return R, err


def frobenius2(A):
    """||A||_F^2 as a 1-element float64 device tensor (one pass, deterministic)."""
    lib = nat.load()
    A = A if A.is_contiguous() else A.contiguous()
    dev = A.device
    out = torch.zeros((1,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        ws = workspace(dev, 32768, tag="frob")
        nat.check(lib.lys_frobenius2(_ptr(A), A.numel(), _ptr(out), _ptr(ws), ws.numel(), _stream_ptr(dev)))
    return out


def build_atom_csr(codes: SparseCodes):
    """users-of-atom index: rowptr (K+1), entries (N*k) = i*k + slot (ksvd.py:111)."""
    lib = nat.load()
    N, k = codes.idx.shape
    K = codes.n_atoms
    dev = codes.idx.device
    rowptr = torch.empty((K + 1,), dtype=torch.int32, device=dev)
    entries = torch.empty((max(N * k, 1),), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = workspace(dev, lib.lys_atom_csr_workspace_bytes(K, N, k), tag="csr")
        nat.check(lib.lys_build_atom_csr(_ptr(codes.idx), _ptr(codes.val), N, k, K, _ptr(rowptr), _ptr(entries),
                                         _ptr(ws), ws.numel(), _stream_ptr(dev)))
    return rowptr, entries


def approx_ksvd_sweep(R, D, codes: SparseCodes, rowptr, entries, n_cycles=1, comm=None):
    """In-place approximate K-SVD sweep over all atoms (ksvd.py:105-124).  Mutates D,
    codes.val and R; returns the int32 (K,) unused-atom flags."""
    lib = nat.load()
    n, K = D.shape
    N, k = codes.idx.shape
    dev = D.device
    if D.stride(1) != 1:
        raise ValueError("D must have contiguous rows")
    unused = torch.empty((K,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = workspace(dev, lib.lys_ksvd_sweep_workspace_bytes(n, K, N, k), tag="sweep")
        nat.check(lib.lys_approx_ksvd_sweep(_ptr(R), _ptr(D), D.stride(0), _ptr(codes.idx), _ptr(codes.val),
                                            _ptr(rowptr), _ptr(entries), n, K, N, k, int(n_cycles),
                                            _ptr(unused), ctypes.c_void_p(comm or 0), _ptr(ws), ws.numel(),
                                            _stream_ptr(dev)))
    return unused


def ksvd_exact_sweep(R, D, codes: SparseCodes, rowptr, entries, n_cycles=1):
    """In-place EXACT K-SVD sweep over all atoms (ksvd.py:19-43): (d, x) <- top singular triplet of R_k.  Mutates D,
    codes.val and R; returns the int32 (K,) unused-atom flags.  n <= 64, single GPU."""
    lib = nat.load()
    n, K = D.shape
    N, k = codes.idx.shape
    dev = D.device
    if D.stride(1) != 1:
        raise ValueError("D must have contiguous rows")
    unused = torch.empty((K,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = workspace(dev, lib.lys_ksvd_exact_workspace_bytes(n, K), tag="sweep")
        nat.check(lib.lys_ksvd_exact_sweep(_ptr(R), _ptr(D), D.stride(0), _ptr(codes.val), _ptr(rowptr), _ptr(entries),
                                           n, K, N, k, int(n_cycles), _ptr(unused), _ptr(ws), ws.numel(), _stream_ptr(dev)))
    return unused


def norm_cols_(D):
    """In-place D[:,c] /= (||D[:,c]|| + eps)  (lyssa/utils/math.py:65-71)."""
    lib = nat.load()
    n, K = D.shape
    if D.stride(1) != 1:
        raise ValueError("D must have contiguous rows")
    with torch.cuda.device(D.device):
        nat.check(lib.lys_norm_cols(_ptr(D), D.stride(0), n, K, _stream_ptr(D.device)))
    return D


def gather_cols_(X, cols, D, dst_cols=None):
    """D[:, dst_cols[j]] = X[:, cols[j]]  (lyssa/dict_learning/utils.py:64; ksvd.py:205)."""
    lib = nat.load()
    n = X.shape[0]
    dev = X.device
    cols_t = torch.as_tensor(np.asarray(cols, dtype=np.int64), device=dev)
    dst_t = None if dst_cols is None else torch.as_tensor(np.asarray(dst_cols, dtype=np.int32), device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.lys_gather_cols(_ptr(X), X.stride(0), X.stride(1), n, _ptr(cols_t), int(cols_t.numel()),
                                      _ptr(D), D.stride(0), _ptr(dst_t), _stream_ptr(dev)))
    return D


def odl_accumulate_(Xb, codes: SparseCodes, beta, A, B):
    """A = beta*A + Z Z^T ; B = beta*B + X Z^T  (online_dict_learn.py:84-85), in place, fixed summation order."""
    lib = nat.load()
    n, b = Xb.shape
    K = A.shape[0]
    dev = A.device
    with torch.cuda.device(dev):
        ws = workspace(dev, lib.lys_odl_accumulate_workspace_bytes(K, b, codes.k), tag="odl_acc")
        nat.check(lib.lys_odl_accumulate(_ptr(Xb), Xb.stride(0), Xb.stride(1), _ptr(codes.idx), _ptr(codes.val),
                                         n, K, b, codes.k, float(beta), _ptr(A), _ptr(B), _ptr(ws), ws.numel(),
                                         _stream_ptr(dev)))


def odl_update_dict_(D, A, B, non_neg=False):
    """D <- norm_cols(clamp(D + (B - D A)/(diag A + eps)))  (online_dict_learn.py:91-98), in place."""
    lib = nat.load()
    n, K = D.shape
    dev = D.device
    with torch.cuda.device(dev):
        ws = workspace(dev, lib.lys_odl_update_workspace_bytes(n, K), tag="odl")
        nat.check(lib.lys_odl_update_dict(_ptr(D), D.stride(0), _ptr(A), _ptr(B), n, K, int(bool(non_neg)),
                                          _ptr(ws), ws.numel(), _stream_ptr(dev)))
    return D
