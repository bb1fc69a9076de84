"""TEST INFRASTRUCTURE ONLY — ctypes wrapper over oracle/bomp_oracle.c (float64 C restatement of
the reference's batch_omp inner loop, lyssa/sparse_coding.py:302-367).  Used by tests for
full-size parity and by bench.py for the CPU baseline.  Never imported by the product."""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblyssa_oracle.so")
_lib = None


def _stale():
    return not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "bomp_oracle.c"))


def build(force=False):
    if force or _stale():
        import fcntl
        os.makedirs(os.path.join(_HERE, "_build"), exist_ok=True)
        with open(os.path.join(_HERE, "_build", ".lock"), "w") as lock:      # the ranks of one launch may all arrive here
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or _stale():
                    subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.lys_oracle_batch_omp.restype = ctypes.c_int
        _lib.lys_oracle_batch_omp.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.lys_oracle_approx_ksvd_sweep.restype = ctypes.c_int
        _lib.lys_oracle_approx_ksvd_sweep.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return _lib


def batch_omp_sparse(X, D, k, threads=None, trace=False, chunk=4096):
    """float64 Batch-OMP of the columns of X (n, N) over D (n, K), the way the reference's
    'bomp' encoder computes it: Gram = D^T D, Alpha = D^T X by NumPy dgemm
    (sparse_coding.py:630-631), then the per-signal loop in C.  Returns idx (N,k) int32
    (-1 padded), val (N,k) float64, nsel (N,) and, if trace, gap/vs (N,k)."""
    lib = _load()
    X = np.asarray(X, dtype=np.float64)
    D = np.asarray(D, dtype=np.float64)
    n_atoms, n_signals = D.shape[1], X.shape[1]
    gram = np.ascontiguousarray(D.T @ D)
    idx = np.empty((n_signals, k), dtype=np.int32)
    val = np.empty((n_signals, k), dtype=np.float64)
    nsel = np.empty(n_signals, dtype=np.int32)
    gap = np.empty((n_signals, k)) if trace else None
    vs = np.empty((n_signals, k)) if trace else None
    threads = threads or os.cpu_count() or 1
    Dt = np.ascontiguousarray(D.T)

    def work(lo):
        hi = min(lo + chunk, n_signals)
        alpha = np.ascontiguousarray((Dt @ X[:, lo:hi]).T)      # (chunk, K) signal-major
        rc = lib.lys_oracle_batch_omp(
            alpha.ctypes.data, 1, gram.ctypes.data, n_atoms, hi - lo, k,
            idx[lo:hi].ctypes.data, val[lo:hi].ctypes.data, nsel[lo:hi].ctypes.data,
            gap[lo:hi].ctypes.data if trace else None, vs[lo:hi].ctypes.data if trace else None)
        if rc != 0:
            raise RuntimeError("lys_oracle_batch_omp failed: %d" % rc)

    starts = list(range(0, n_signals, chunk))
    if threads == 1 or len(starts) == 1:
        for lo in starts:
            work(lo)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, starts))
    if trace:
        return idx, val, nsel, gap, vs
    return idx, val, nsel


def densify(idx, val, n_atoms):
    """(idx,val)[N,k] -> dense Z (K, N) float64, the reference's return layout."""
    n_signals, k = idx.shape
    Z = np.zeros((n_atoms, n_signals))
    rows = np.repeat(np.arange(n_signals), k)
    flat_i = idx.reshape(-1)
    ok = flat_i >= 0
    Z[flat_i[ok], rows[ok]] = val.reshape(-1)[ok]
    return Z


def approx_ksvd_sparse(Y, D, idx, val, n_cycles=1):
    """float64 approx_ksvd sweep (lyssa/dict_learning/ksvd.py:98-126) on sparse codes, in C: for sweeps too large
    for the dense NumPy restatement.  Y (n, N), D (n, K), idx/val (N, k).  Returns (D', val', unused list, R')
    with R' = Y - D' X' the (N, n) residual the sweep maintains (:123); inputs are not modified."""
    lib = _load()
    Y = np.asarray(Y, dtype=np.float64)
    D = np.array(D, dtype=np.float64, order="C")
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    val = np.array(val, dtype=np.float64, order="C")
    n, n_signals = Y.shape
    n_atoms, k = D.shape[1], idx.shape[1]
    R = np.ascontiguousarray(Y.T).copy()                       # :103  R = Y - D X, from the sparse codes
    Dt = np.ascontiguousarray(D.T)
    for s in range(k):
        a = idx[:, s]
        ok = a >= 0
        R[ok] -= Dt[a[ok]] * val[ok, s][:, None]
    unused = np.zeros(n_atoms, dtype=np.int32)
    rc = lib.lys_oracle_approx_ksvd_sweep(R.ctypes.data, D.ctypes.data, idx.ctypes.data, val.ctypes.data,
                                          n_signals, n, n_atoms, k, n_cycles, unused.ctypes.data)
    if rc != 0:
        raise RuntimeError("lys_oracle_approx_ksvd_sweep failed: %d" % rc)
    return D, val, np.nonzero(unused)[0].tolist(), R
