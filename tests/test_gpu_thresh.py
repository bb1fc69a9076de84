"""GPU parity of the 'thresh' and 'iht' coders against the committed golden
codes of the live reference and against the oracle on seeded inputs.  Reference semantics:
lyssa/sparse_coding.py:416-425 (thresholding), :433-446 (iterative_hard_thresh), :636-641 and
:671-690 (dispatch).  Supports are bit-identical on columns that are not near-ties
(tests/parity.py: GAP_TOL on the k-th/(k+1)-th selection keys); coefficients within COEF_TOL."""
import numpy as np
import pytest
import torch

from oracle import lyssa_oracle as lo
from parity import GAP_TOL, check_codes, dense_to_codes, thresh_trace

from lyssa.sparse_coding import sparse_encoder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

_CASES = (("thresh_k5", "thresh", {"n_nonzero_coefs": 5}, 5),
          ("thresh_p10", "thresh", {"nonzero_percentage": 0.1}, None),
          ("iht_k5", "iht", {"n_nonzero_coefs": 5, "eta": 0.2, "n_iter": 4}, 5),
          ("iht_k3_it0", "iht", {"n_nonzero_coefs": 3, "eta": 0.2, "n_iter": 0}, 3))


def _gpu_codes(alg, params, X, D):
    enc = sparse_encoder(alg, dict(params), verbose=False)
    codes = enc.encode_sparse(torch.as_tensor(np.ascontiguousarray(X), device=DEV), D)
    torch.cuda.synchronize()
    return codes


def test_thresh_iht_match_golden(golden):
    g = golden("thresh")
    for tag in ("a", "b"):
        X, D = g["X_" + tag], g["D_" + tag]
        for name, alg, params, k in _CASES:
            k = int(np.floor(0.1 * D.shape[1])) if k is None else k
            _, gap = thresh_trace(X, D, k, params.get("eta"), params.get("n_iter", 0) if alg == "iht" else 0)
            codes = _gpu_codes(alg, params, X, D)
            assert tuple(codes.idx.shape) == (X.shape[1], k) and bool((codes.nsel == k).all())
            rep = check_codes(codes.idx.cpu().numpy(), codes.val.cpu().numpy(),
                              g["idx_%s_%s" % (name, tag)], g["val_%s_%s" % (name, tag)],
                              ok=gap >= GAP_TOL, label="%s/%s" % (name, tag))
            assert rep["compared"] >= 0.98 * X.shape[1]


@pytest.mark.parametrize("n,K,N,alg,params", [
    (64, 1024, 3000, "thresh", {"n_nonzero_coefs": 5}),                 # fused tcgen05 kernel, CTA pairs
    (64, 512, 2500, "thresh", {"n_nonzero_coefs": 10}),                 # fused, single CTA, 10-entry lists
    (48, 768, 1111, "thresh", {"n_nonzero_coefs": 3}),                  # fused, zero-padded features
    (64, 256, 129, "thresh", {"n_nonzero_coefs": 1}),
    (64, 1024, 700, "thresh", {"nonzero_percentage": 0.1}),             # k = 102 > 32
    (64, 512, 2049, "iht", {"n_nonzero_coefs": 8, "eta": 0.1, "n_iter": 3}),
    (37, 130, 999, "iht", {"n_nonzero_coefs": 4, "eta": 0.3, "n_iter": 5}),   # ragged shapes, fp32 SIMT GEMM
    (128, 2048, 500, "thresh", {"n_nonzero_coefs": 32}),
    (16, 8, 65, "thresh", {"n_nonzero_coefs": 8}),                      # k == K: everything kept
    (64, 1000, 5000, "thresh", {"n_nonzero_coefs": 5}),                 # K not a multiple of 256: tcgen05 GEMM off, two-kernel path
])
def test_thresh_iht_seeded_vs_oracle(n, K, N, alg, params):
    X = lo.synthetic_patches(N, n, seed=N + K)
    D = lo.synthetic_dictionary(K, n, seed=N + K + 1)
    k = params.get("n_nonzero_coefs") or int(np.floor(params["nonzero_percentage"] * K))
    Zo, gap = thresh_trace(X, D, k, params.get("eta"), params.get("n_iter", 0))
    Zref = lo.sparse_encoder(alg, dict(params), verbose=False).encode(X.astype(np.float64), D.astype(np.float64))
    assert np.array_equal(Zo != 0, Zref != 0) or (gap < GAP_TOL).any()
    io, vo = dense_to_codes(Zref, k)
    enc = sparse_encoder(alg, dict(params), verbose=False)
    Xd = torch.as_tensor(np.ascontiguousarray(X), device=DEV)
    codes = enc.encode_sparse(Xd, D)
    rep = check_codes(codes.idx.cpu().numpy(), codes.val.cpu().numpy(), io, vo, ok=gap >= GAP_TOL,
                      label="%s n=%d K=%d" % (alg, n, K))
    assert rep["compared"] >= 0.97 * N
    # dense output: device tensor in -> (K, N) device tensor; NumPy in -> NumPy out; both equal the codes
    Zd = enc.encode(Xd, torch.as_tensor(D, device=DEV))
    assert tuple(Zd.shape) == (K, N) and Zd.is_cuda
    assert torch.equal(Zd, codes.to_dense())
    Zh = enc.encode(X, D)
    assert isinstance(Zh, np.ndarray) and np.array_equal(Zh, Zd.cpu().numpy())


@pytest.mark.parametrize("K,k", [(1024, 5), (512, 10), (256, 3)])
def test_thresh_fused_candidate_lists_stay_exact_on_sorted_and_constant_correlations(K, k):
    # the fused kernel collects, per signal, the columns that reach a running lower bound of the k-th largest
    # correlation and prunes the list when it fills up (bomp_fused.cu, MODE 1).  Correlations that ASCEND with the
    # column make every column a candidate (the bound always lags), all-equal correlations make every column a tie:
    # both must still return the k largest with ties to the lower column (sparse_coding.py:416-425).
    n, N = 64, 300
    rng = np.random.RandomState(K + k)
    theta = np.linspace(1.2, 0.1, K)                         # cos(theta) ascends with the column
    D = np.zeros((n, K), dtype=np.float32)
    D[0], D[1] = np.cos(theta), np.sin(theta)
    X = np.zeros((n, N), dtype=np.float32)
    X[0] = rng.uniform(0.5, 2.0, N)                          # ascending correlations: the last k columns win
    X[:, 100:200] *= -1.0                                    # descending: the first k columns win
    X[:, 200:] = 0.0                                         # all correlations equal (0): the first k columns
    codes = sparse_encoder("thresh", {"n_nonzero_coefs": k}, verbose=False).encode_sparse(
        torch.as_tensor(np.ascontiguousarray(X), device=DEV), D)
    idx = codes.idx.cpu().numpy(); val = codes.val.cpu().numpy()
    A = D.astype(np.float64).T @ X.astype(np.float64)
    want = np.argsort(-A, axis=0, kind="stable")[:k].T      # descending, ties to the lower column
    assert np.array_equal(idx, want)
    assert np.allclose(val, np.take_along_axis(A.T, want, axis=1), rtol=1e-5, atol=1e-6)


def test_thresh_keeps_largest_signed_not_absolute():
    # the reference sorts SIGNED correlations (sparse_coding.py:423): a large negative one is never kept
    D = np.eye(4, dtype=np.float32)
    X = np.array([[-9.0, 1.0, 2.0, 3.0]], dtype=np.float32).T
    Z = sparse_encoder("thresh", {"n_nonzero_coefs": 2}, verbose=False).encode(X, D)
    assert np.array_equal(Z[:, 0], np.array([0.0, 0.0, 2.0, 3.0], dtype=np.float32))
    # while 'iht' keeps by magnitude after its first gradient step (:442)
    Z = sparse_encoder("iht", {"n_nonzero_coefs": 2, "eta": 1.0, "n_iter": 1}, verbose=False).encode(X, D)
    assert np.array_equal(Z[:, 0], np.array([-9.0, 0.0, 0.0, 3.0], dtype=np.float32))


def test_thresh_errors():
    X = np.zeros((8, 5), dtype=np.float32); D = np.eye(8, dtype=np.float32)
    with pytest.raises(Exception, match="n_nonzero_coefs"):
        sparse_encoder("thresh", {"n_nonzero_coefs": 9}, verbose=False).encode(X, D)
    with pytest.raises(Exception, match="n_nonzero_coefs"):
        sparse_encoder("iht", {"n_nonzero_coefs": 40, "eta": 0.1, "n_iter": 1}, verbose=False).encode(
            np.zeros((64, 5), dtype=np.float32), np.zeros((64, 128), dtype=np.float32))
    assert sparse_encoder("thresh", {"n_nonzero_coefs": 2}, verbose=False).encode(np.zeros((8, 0), dtype=np.float32), D).shape == (8, 0)
