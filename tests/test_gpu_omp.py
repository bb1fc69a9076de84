"""GPU parity tests of the reference's plain `omp` coder (lyssa/sparse_coding.py:618-625 -> :19-66, its DEFAULT
algorithm) and of feature_encoder('soft_thresholding') (lyssa/feature_encoding.py:26-89): golden vectors written by the
live reference, seeded shapes against the float64 oracle."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parity import GAP_TOL, check_codes, dense_to_codes, omp_trace, thresh_trace  # noqa: E402
from oracle import lyssa_oracle as lo  # noqa: E402
from lyssandra_b200 import engine  # noqa: E402
from lyssandra_b200.sparse_coding import sparse_encoder  # noqa: E402
from lyssa.feature_encoding import feature_encoder, soft_thresholding  # noqa: E402  (the drop-in alias package)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# ||r|| is formed as sqrt(||x||^2 - y.y) in float32 (absolute error ~1e-6 ||x||^2 in the square): a stopping decision whose
# ||r|| is closer than this (relative to ||x||) to the threshold may flip
STOP_TOL = 5e-5


def _codes_of(Z):
    Z = np.asarray(Z)
    return dense_to_codes(Z, int((Z != 0).sum(axis=0).max()))


def _compare(Zg, Zref, ok, label, coef_tol=2e-5):
    k = int(max((Zg != 0).sum(axis=0).max(), (Zref != 0).sum(axis=0).max(), 1))
    ig, vg = dense_to_codes(Zg, k)
    ir, vr = dense_to_codes(Zref, k)
    rep = check_codes(ig, vg, ir, vr, ok=ok, coef_tol=coef_tol, label=label)
    assert rep["compared"] >= 0.9 * Zg.shape[1], rep
    return rep


@pytest.mark.parametrize("tag,params,dkey", [("k5", {"n_nonzero_coefs": 5}, "D"), ("tol1p2", {"tol": 1.2}, "D"),
                                              ("tol0p8", {"tol": 0.8}, "D"), ("scaled_k4", {"n_nonzero_coefs": 4}, "D_scaled")])
def test_omp_matches_golden(golden, tag, params, dkey):
    g = golden("omp")
    X, D = g["X"], g[dkey]
    _, gap, margin = omp_trace(X, D, params.get("n_nonzero_coefs"), params.get("tol"))
    ok = (gap >= GAP_TOL) & ((margin >= STOP_TOL) | ("tol" not in params))
    enc = sparse_encoder("omp", dict(params), verbose=False)
    Zd = enc.encode(torch.from_numpy(X).to(DEV), torch.from_numpy(D).to(DEV))
    assert tuple(Zd.shape) == (D.shape[1], X.shape[1]) and Zd.is_cuda
    _compare(Zd.cpu().numpy(), g["Z_" + tag], ok, "omp " + tag)
    Zh = enc.encode(X, D)                                        # NumPy in -> NumPy out
    assert isinstance(Zh, np.ndarray) and np.array_equal(Zh, Zd.cpu().numpy())


@pytest.mark.parametrize("n,K,N,params", [(64, 1024, 3000, {"n_nonzero_coefs": 5}), (64, 512, 2000, {"tol": 1.0}),
                                           (128, 2048, 500, {"n_nonzero_coefs": 10}), (33, 300, 700, {"tol": 0.9}),
                                           (16, 40, 300, {"n_nonzero_coefs": 8})])
def test_omp_seeded_vs_oracle(n, K, N, params):
    X = lo.synthetic_patches(N, n, seed=N + n)
    D = lo.synthetic_dictionary(K, n, seed=K + n)
    Zref, gap, margin = omp_trace(X, D, params.get("n_nonzero_coefs"), params.get("tol"))
    ok = (gap >= GAP_TOL) & ((margin >= STOP_TOL) | ("tol" not in params))
    Z = sparse_encoder("omp", dict(params), verbose=False).encode(torch.from_numpy(np.ascontiguousarray(X)).to(DEV), torch.from_numpy(D).to(DEV))
    _compare(Z.cpu().numpy(), Zref, ok, "omp n%d K%d" % (n, K), coef_tol=2e-5)


def test_omp_equals_bomp_on_unit_norm_dictionaries_and_errors():
    X = lo.synthetic_patches(5000, 64, seed=3); D = lo.synthetic_dictionary(1024, 64, seed=4)
    Xd = torch.from_numpy(np.ascontiguousarray(X)).to(DEV); Dd = torch.from_numpy(D).to(DEV)
    a = engine.omp_encode(Xd, Dd, 5)
    b = engine.bomp_encode(Xd, Dd, 5)
    same = (a.idx == b.idx).all(dim=1)
    assert int((~same).sum()) <= 5                               # only float32 near-ties may differ
    assert float((a.val[same] - b.val[same]).abs().max()) <= 2e-5 * float(b.val.abs().max())
    with pytest.raises(ValueError):
        sparse_encoder("omp", {}, verbose=False).encode(Xd, Dd)   # neither n_nonzero_coefs nor tol
    X128 = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(50, 128, seed=8))).to(DEV)
    D128 = torch.from_numpy(lo.synthetic_dictionary(512, 128, seed=9)).to(DEV)
    with pytest.raises(NotImplementedError, match="more than"):
        sparse_encoder("omp", {"tol": 0.05}, verbose=False).encode(X128, D128)        # ~100 atoms per signal: over the engine's 64
    assert sparse_encoder("omp", {"n_nonzero_coefs": 3}, verbose=False).encode(torch.empty((64, 0), device=DEV), Dd).shape == (1024, 0)


def test_feature_encoder_soft_thresholding(golden):
    g = golden("omp")
    X, D = g["X"], g["D"]
    _, gap = thresh_trace(X, D, 7)
    ok = gap >= GAP_TOL
    fe = feature_encoder(algorithm="soft_thresholding", params={"n_nonzero_coefs": 7}, verbose=False)
    Z = fe.encode(X, D)
    assert isinstance(Z, np.ndarray) and Z.shape == (256, 300)
    _compare(Z, g["Z_soft_k7"], ok, "feature_encoder")
    Zd = fe.encode(torch.from_numpy(X).to(DEV), torch.from_numpy(D).to(DEV))
    assert Zd.is_cuda and np.array_equal(Zd.cpu().numpy(), Z)
    # the bare function takes Alpha = D^T X (feature_encoding.py:26)
    A = (D.astype(np.float64).T @ X.astype(np.float64)).astype(np.float32)
    Zs = soft_thresholding(A, n_nonzero_coefs=7)
    _compare(Zs, g["Z_soft_k7"], ok, "soft_thresholding(Alpha)")
    Zp = soft_thresholding(torch.from_numpy(A).to(DEV), nonzero_percentage=7.5 / 256)     # floor(p*K) = 7
    assert np.array_equal(Zp.cpu().numpy(), Zs)
    with pytest.raises(ValueError):
        feature_encoder(algorithm="nope", params={"n_nonzero_coefs": 7}).encode(X, D)
