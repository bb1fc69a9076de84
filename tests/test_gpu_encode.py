"""GPU parity tests of the Batch-OMP encode path, through the reference-facing class and the
raw C-ABI: golden vectors from the live reference, KATs, layouts, and full-size properties."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import lyssa_oracle as lo  # noqa: E402
from lyssandra_b200 import _native, engine  # noqa: E402
from lyssandra_b200.sparse_coding import sparse_encoder  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _enc(k):
    return sparse_encoder(algorithm="bomp", params={"n_nonzero_coefs": k}, verbose=False)


def _sparse_from_dense(Z, k):
    Z = np.asarray(Z)
    K, N = Z.shape
    idx = -np.ones((N, k), dtype=np.int64); val = np.zeros((N, k))
    for i in range(N):
        nz = np.flatnonzero(Z[:, i])
        assert len(nz) <= k
        idx[i, :len(nz)] = nz; val[i, :len(nz)] = Z[nz, i]
    return idx, val


def _oracle_ok(X, D, k):
    idx, val, nsel, gap, vs = co.batch_omp_sparse(X.astype(np.float64), D.astype(np.float64), k, trace=True)
    return idx, val, parity.comparable_columns(gap, vs, nsel, k)


@pytest.mark.parametrize("name,ik,vk,k", [("bomp_cfg1", "idx", "val", 5), ("bomp_cfg4", "idx", "val", 5),
                                           ("bomp_K1024", "idx_k5", "val_k5", 5), ("bomp_K1024", "idx_k10", "val_k10", 10)])
def test_golden_vectors_device_tensors(golden, name, ik, vk, k):
    g = golden(name)
    X = torch.from_numpy(g["X"]).to(DEV); D = torch.from_numpy(g["D"]).to(DEV)
    codes = _enc(k).encode_sparse(X, D)
    _, _, ok = _oracle_ok(g["X"], g["D"], k)
    rep = parity.check_codes(codes.idx.cpu().numpy(), codes.val.cpu().numpy(), g[ik], g[vk], ok, label=name)
    assert rep["excluded"] <= max(2, rep["columns"] // 100), rep
    # dense contract: Z (K, N), zeros everywhere else
    Z = _enc(k).encode(X, D)
    assert tuple(Z.shape) == (g["D"].shape[1], g["X"].shape[1]) and Z.dtype == torch.float32
    assert torch.equal(Z, codes.to_dense())
    assert int((Z != 0).sum()) == int((codes.val != 0).sum())


def test_golden_vectors_numpy_in_numpy_out(golden):
    g = golden("bomp_cfg1")
    Z = _enc(5).encode(g["X"], g["D"])                      # reference-layout (n, N) C-order host array
    assert isinstance(Z, np.ndarray) and Z.shape == (256, 512)
    _, _, ok = _oracle_ok(g["X"], g["D"], 5)
    idx, val = _sparse_from_dense(Z, 5)
    parity.check_codes(idx, val, g["idx"], g["val"], ok, label="numpy")
    Z64 = _enc(5).encode(g["X"].astype(np.float64), g["D"].astype(np.float64))       # float64 callers
    assert np.array_equal(Z64, Z)
    Zt = _enc(5).encode(np.ascontiguousarray(g["X"].T).T, g["D"])                     # signal-major storage
    assert np.array_equal(Zt, Z)
    i2, v2, n2 = _enc(5).encode_sparse_host(g["X"], g["D"])
    assert np.array_equal(parity.sorted_codes(i2, v2)[0], parity.sorted_codes(idx, val)[0])


def test_kats(golden):
    g = golden("bomp_kat")

    def run(Xk, Dk, k):
        return _enc(k).encode(torch.from_numpy(g[Xk]).to(DEV), torch.from_numpy(g[Dk]).to(DEV)).cpu().numpy()

    Z = run("ortho_X", "ortho_D", 3)
    assert np.array_equal(Z != 0, g["ortho_Z"] != 0) and np.max(np.abs(Z - g["ortho_Z"])) < 1e-5
    Z = run("ties_X", "ties_D", 2)                          # exact ties -> lowest index
    assert np.array_equal(Z, g["ties_Z"])
    Z = run("kK_X", "kK_D", 4)                              # k = K (reference test regime)
    assert Z.shape == (4, 100)
    assert np.max(np.abs(Z - g["kK_Z"])) <= 2e-4 * np.max(np.abs(g["kK_Z"]))       # ill-conditioned 4 of 4 atoms
    for k, key in ((1, "k1_Z"), (2, "k2_Z")):
        Z = run("k12_X", "k12_D", k)
        assert np.array_equal(Z != 0, g[key] != 0) and np.max(np.abs(Z - g[key])) <= 1e-5 * np.max(np.abs(g[key]))
    # degenerate columns: oracle policy — coefficients agree; supports only where well-posed
    Z = run("degen_X", "k12_D", 4)
    ref = g["degen_Z"]
    assert np.max(np.abs(Z - ref)) <= 1e-5 * np.max(np.abs(ref))
    assert np.count_nonzero(np.abs(Z[:, 0]) > 1e-5) == 1 and abs(Z[7, 0] - 2.5) < 1e-5
    assert np.count_nonzero(Z[:, 1]) == 0                    # zero signal -> zeros, never NaN
    assert np.all(np.isfinite(Z))
    # duplicate atoms: never selected twice, coefficients finite and close
    Z = run("k12_X", "dup_D", 3)
    assert np.all(np.isfinite(Z)) and np.max(np.abs(Z - g["dup_Z"])) <= 1e-4 * np.max(np.abs(g["dup_Z"]))
    assert not np.any((Z[0] != 0) & (Z[1] != 0))


def test_layouts_and_raw_cabi(golden):
    """Same answer for feature-major (reference) and signal-major X, padded ldd, via ctypes."""
    g = golden("bomp_K1024")
    lib = _native.load()
    n, N = g["X"].shape; K = 1024; k = 5
    Xf = torch.from_numpy(g["X"]).to(DEV).contiguous()                 # (n, N): feat stride N
    Xs = Xf.t().contiguous()                                           # (N, n): feat stride 1
    Dp = torch.zeros((n, K + 32), device=DEV); Dp[:, :K] = torch.from_numpy(g["D"]).to(DEV)
    G = torch.empty((K, K), device=DEV)
    _native.check(lib.lys_gram(ctypes.c_void_p(Dp.data_ptr()), K + 32, n, K, ctypes.c_void_p(G.data_ptr()), None))
    Gref = g["D"].astype(np.float64).T @ g["D"].astype(np.float64)
    assert np.max(np.abs(G.cpu().numpy() - Gref)) < 5e-7
    outs = []
    for X, xfs, xss in ((Xf, N, 1), (Xs, 1, n)):
        idx = torch.empty((N, k), dtype=torch.int32, device=DEV); val = torch.empty((N, k), device=DEV)
        nsel = torch.empty((N,), dtype=torch.int32, device=DEV); Z = torch.full((N, K), 7.0, device=DEV)
        wsb = lib.lys_bomp_workspace_bytes(n, K, N, k)
        ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
        rc = lib.lys_bomp_encode(X.data_ptr(), xfs, xss, Dp.data_ptr(), K + 32, G.data_ptr(), n, K, N, k,
                                 idx.data_ptr(), val.data_ptr(), nsel.data_ptr(), Z.data_ptr(), 1, K,
                                 ws.data_ptr(), wsb, None)
        _native.check(rc)
        torch.cuda.synchronize()
        outs.append((idx.cpu().numpy(), val.cpu().numpy(), nsel.cpu().numpy(), Z.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.all(outs[0][2] == k)
    _, _, ok = _oracle_ok(g["X"], g["D"], k)
    parity.check_codes(outs[0][0], outs[0][1], g["idx_k5"], g["val_k5"], ok, label="raw C-ABI")
    Zd = outs[0][3]
    assert np.count_nonzero(Zd) == N * k and not np.any(Zd == 7.0)      # fully overwritten
    # too-small workspace and bad k are refused with a message, not a crash
    rc = lib.lys_bomp_encode(Xf.data_ptr(), N, 1, Dp.data_ptr(), K + 32, G.data_ptr(), n, K, N, k,
                             idx.data_ptr(), val.data_ptr(), None, None, 1, K, ws.data_ptr(), 16, None)
    assert rc == _native.LYS_EWORKSPACE
    rc = lib.lys_bomp_encode(Xf.data_ptr(), N, 1, Dp.data_ptr(), K + 32, G.data_ptr(), n, K, N, 40,
                             idx.data_ptr(), val.data_ptr(), None, None, 1, K, ws.data_ptr(), wsb, None)
    assert rc == _native.LYS_EINVAL


@pytest.mark.parametrize("n,K,N,k", [(64, 256, 1000, 5),        # BASELINE cfg1
                                     (64, 1024, 20000, 5), (64, 1024, 4099, 10), (128, 2048, 3000, 5),
                                     (10, 4, 100, 4), (64, 100, 777, 3), (33, 300, 500, 7), (64, 1024, 1, 5)])
def test_seeded_random_vs_oracle(n, K, N, k):
    X = lo.synthetic_patches(N, n, seed=100 + N % 97)
    D = lo.synthetic_dictionary(K, n, seed=7 + K)
    if K <= n:
        D = np.ascontiguousarray(lo.norm_cols(np.random.default_rng(3).random((n, K))).astype(np.float32))
    codes = _enc(k).encode_sparse(torch.from_numpy(np.ascontiguousarray(X)).to(DEV), torch.from_numpy(D).to(DEV))
    idx_r, val_r, ok = _oracle_ok(X, D, k)
    tol = parity.COEF_TOL if K > n else 1e-3          # K <= n with k = K: conditioning of random non-neg atoms
    rep = parity.check_codes(codes.idx.cpu().numpy(), codes.val.cpu().numpy(), idx_r, val_r, ok,
                             coef_tol=tol, label="n%d K%d N%d k%d" % (n, K, N, k))
    assert rep["excluded"] <= max(1, N // 200), rep
    assert np.array_equal(codes.nsel.cpu().numpy()[ok], np.full(int(ok.sum()), k))


def test_full_size_properties_cfg2():
    """BASELINE cfg2 at full size (1M patches, K=1024, k=5): oracle parity on a 131072-column
    prefix (the CPU-baseline subset) + size-independent properties on all 1M:
    k distinct atoms per signal, residual orthogonal to the selected atoms, determinism,
    idempotence of re-encoding the reconstruction's support, shard-invariance."""
    N, n, K, k = 1 << 20, 64, 1024, 5
    Xh = lo.synthetic_patches(N, n, seed=0); Dh = lo.synthetic_dictionary(K, n, seed=1)
    X = torch.from_numpy(np.ascontiguousarray(Xh.T)).to(DEV).t()        # signal-major storage, logical (n, N)
    D = torch.from_numpy(Dh).to(DEV)
    enc = _enc(k)
    codes = enc.encode_sparse(X, D)
    idx, val = codes.idx, codes.val
    sub = 131072
    idx_r, val_r, ok = _oracle_ok(Xh[:, :sub], Dh, k)
    rep = parity.check_codes(idx[:sub].cpu().numpy(), val[:sub].cpu().numpy(), idx_r, val_r, ok, label="cfg2 prefix")
    print("cfg2 parity report:", rep)
    assert rep["excluded"] <= sub // 500
    # properties on all columns
    assert int((idx < 0).sum()) == 0 and int(codes.nsel.min()) == k
    s, _ = torch.sort(idx, dim=1)
    assert bool((s[:, 1:] != s[:, :-1]).all())
    R, err = engine.residual(X, D, codes)
    Dsel = D.t()[idx.long()]                                             # (N, k, n)
    corr = torch.einsum("nkf,nf->nk", Dsel, R).abs().max(dim=1).values
    xnorm = torch.linalg.vector_norm(X, dim=0)
    assert float((corr / xnorm).max()) < 2e-5                            # normal equations hold
    assert abs(float(err.item()) - float((R.double() ** 2).sum())) <= 1e-6 * float(err.item())
    codes2 = enc.encode_sparse(X, D)
    assert torch.equal(codes2.idx, idx) and torch.equal(codes2.val, val)  # deterministic
    lo_, hi_ = 300000, 700001
    part = enc.encode_sparse(X[:, lo_:hi_], D)
    assert torch.equal(part.idx, idx[lo_:hi_]) and torch.equal(part.val, val[lo_:hi_])   # shard-invariant
    # dense output: checksum of checksums
    Z = enc.encode(X[:, :200000], D)
    assert abs(float(Z.double().sum()) - float(val[:200000].double().sum())) < 1e-3
    assert int((Z != 0).sum()) == int((val[:200000] != 0).sum())


def test_empty_and_errors():
    D = torch.from_numpy(lo.synthetic_dictionary(64, 64, seed=1)).to(DEV)
    Z = _enc(3).encode(torch.empty((64, 0), device=DEV), D)
    assert tuple(Z.shape) == (64, 0)
    with pytest.raises(ValueError):
        _enc(3).encode(torch.zeros((32, 5), device=DEV), D)              # feature mismatch
    with pytest.raises(_native.LyssaError):
        _enc(65).encode(torch.zeros((64, 5), device=DEV), D)             # k > K


# ---- shapes served by the fused tcgen05 kernel (csrc/bomp_fused.cu): single-CTA (K <= 512) and CTA-pair
# (K > 512) variants, zero-padded features (n < 64), ragged last tile, k = 1 .. 10
@pytest.mark.parametrize("n,K,N,k", [(64, 512, 3001, 5), (64, 768, 2000, 3), (48, 512, 1000, 4), (64, 256, 130, 1),
                                     (64, 1024, 257, 2), (64, 1024, 3000, 10), (20, 384, 500, 6)])
def test_fused_kernel_shapes_vs_oracle(n, K, N, k):
    X = lo.synthetic_patches(N, n, seed=300 + N % 89)
    D = lo.synthetic_dictionary(K, n, seed=11 + K)
    codes = _enc(k).encode_sparse(torch.from_numpy(np.ascontiguousarray(X)).to(DEV), torch.from_numpy(D).to(DEV))
    idx_r, val_r, ok = _oracle_ok(X, D, k)
    rep = parity.check_codes(codes.idx.cpu().numpy(), codes.val.cpu().numpy(), idx_r, val_r, ok,
                             label="fused n%d K%d N%d k%d" % (n, K, N, k))
    assert rep["excluded"] <= max(2, N // 100), rep


def test_fused_kernel_dense_layouts():
    """Dense Z from the fused kernel: contiguous rows, padded rows (z_sig_stride > K) and the sparse-only
    call give the same codes; every element of Z is written exactly once (zeros included)."""
    lib = _native.load()
    n, K, N, k = 64, 1024, 1000, 5
    Xh = lo.synthetic_patches(N, n, seed=5); Dh = lo.synthetic_dictionary(K, n, seed=6)
    X = torch.from_numpy(np.ascontiguousarray(Xh.T)).to(DEV)           # (N, n) signal-major
    D = torch.from_numpy(Dh).to(DEV)
    G = torch.empty((K, K), device=DEV)
    _native.check(lib.lys_gram(D.data_ptr(), K, n, K, G.data_ptr(), None))
    wsb = lib.lys_bomp_workspace_bytes(n, K, N, k)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    outs = []
    for zss in (K, K + 64, 0):
        idx = torch.full((N, k), -9, dtype=torch.int32, device=DEV); val = torch.full((N, k), 9.0, device=DEV)
        Z = torch.full((N, max(zss, 1)), 7.0, device=DEV) if zss else None
        _native.check(lib.lys_bomp_encode(X.data_ptr(), 1, n, D.data_ptr(), K, G.data_ptr(), n, K, N, k,
                                          idx.data_ptr(), val.data_ptr(), None, Z.data_ptr() if zss else None, 1, max(zss, K),
                                          ws.data_ptr(), wsb, None))
        torch.cuda.synchronize()
        outs.append((idx.cpu().numpy(), val.cpu().numpy(), None if Z is None else Z.cpu().numpy()))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])
    for o in outs[:2]:
        Zd = o[2][:, :K]
        ref = np.zeros((N, K), dtype=np.float32)
        ref[np.arange(N)[:, None], o[0]] = o[1]
        assert np.array_equal(Zd, ref)
        if o[2].shape[1] > K:
            assert np.all(o[2][:, K:] == 7.0)                           # padding columns untouched


def test_screen_and_split3_correlations_give_identical_codes():
    """The default fused path computes fp32-faithful correlations with three fp16 products; the LYS_BOMP_SCREEN path
    ranks with ONE product and certifies or exactly resolves every argmax.  Both must select the same atoms on
    every column that is not a near-tie in float64 — and, because everything after the argmax is the same fp32 code,
    then return bit-identical coefficients.  cfg2 data, 262144 columns, k = 5 and k = 10; K = 512 (single-CTA variant)."""
    for n, K, N, k, seed in ((64, 1024, 262144, 5, 0), (64, 1024, 65536, 10, 3), (64, 512, 65536, 5, 4), (40, 768, 20000, 7, 5)):
        Xh = lo.synthetic_patches(N, n, seed=seed); Dh = lo.synthetic_dictionary(K, n, seed=seed + 1)
        X = torch.from_numpy(np.ascontiguousarray(Xh.T)).to(DEV).t()
        D = torch.from_numpy(Dh).to(DEV)
        a = engine.bomp_encode(X, D, k)
        b = engine.bomp_encode(X, D, k, screen=True)
        same = (a.idx == b.idx).all(dim=1)
        differ = torch.nonzero(~same).flatten().cpu().numpy()
        if differ.size:
            # a split-product value carries a 2^-22 relative error, so the two paths may part on float64 near-ties only
            _, _, ok = _oracle_ok(Xh[:, differ], Dh, k)
            assert not ok.any(), "paths differ on %d well-separated columns" % int(ok.sum())
        assert differ.size <= N // 20000 + 1
        assert torch.equal(a.val[same], b.val[same]) and torch.equal(a.nsel[same], b.nsel[same])


@pytest.mark.parametrize("n,K,k", [(64, 1024, 5), (64, 512, 10), (128, 1024, 5), (33, 300, 4)])
def test_nan_and_inf_signals_are_memory_safe(n, K, k):
    """np.argmax on a NaN column returns an index in range (the first NaN) and the reference just produces NaN codes
    (sparse_coding.py:322); here every kernel family (fused, two-kernel, generic) must stay in bounds: indices in
    [-1, K), the clean columns unaffected, no sticky CUDA error."""
    N = 700
    Xh = lo.synthetic_patches(N, n, seed=9).copy(); Dh = lo.synthetic_dictionary(K, n, seed=10)
    bad = [3, 130, 131, 699]
    Xh[5, 3] = np.nan; Xh[:, 130] = np.nan; Xh[0, 131] = np.inf; Xh[7, 699] = -np.inf
    X = torch.from_numpy(np.ascontiguousarray(Xh)).to(DEV); D = torch.from_numpy(Dh).to(DEV)
    codes, Z = engine.bomp_encode(X, D, k, dense=True)
    torch.cuda.synchronize()
    idx = codes.idx.cpu().numpy()
    assert idx.min() >= -1 and idx.max() < K
    good = np.setdiff1d(np.arange(N), bad)
    clean = engine.bomp_encode(torch.from_numpy(np.ascontiguousarray(Xh[:, good])).to(DEV), D, k)
    assert np.array_equal(idx[good], clean.idx.cpu().numpy())
    assert torch.equal(codes.val[torch.as_tensor(good, device=DEV)], clean.val)
    assert bool(torch.isfinite(Z.t()[torch.as_tensor(good, device=DEV)]).all())


def test_dictionary_scale_does_not_matter_to_the_fused_path():
    """The fp16 planes of the dictionary use a power-of-two scale derived from max|D| (a fixed scale overflowed for
    |d| > 2047): a dictionary scaled by 4096 or 1/4096 must select like the unit-norm one.  'thresh' does not need unit
    norm in the reference (sparse_coding.py:416-425), so this matters for the fused thresholding coder as well."""
    n, K, N, k = 64, 1024, 4000, 5
    Xh = lo.synthetic_patches(N, n, seed=13); Dh = lo.synthetic_dictionary(K, n, seed=14)
    X = torch.from_numpy(np.ascontiguousarray(Xh)).to(DEV)
    base = engine.thresh_encode(X, torch.from_numpy(Dh).to(DEV), k)
    for scale in (4096.0, 1.0 / 4096.0):
        c = engine.thresh_encode(X, torch.from_numpy(Dh * np.float32(scale)).to(DEV), k)
        same = (c.idx == base.idx).all(dim=1)
        assert int((~same).sum()) <= N // 500
        assert torch.allclose(c.val[same], base.val[same] * scale, rtol=1e-6, atol=0)
    # Batch-OMP itself assumes unit-norm atoms (quirk Q1), so only the first pick is scale-free: k = 1
    b1 = engine.bomp_encode(X, torch.from_numpy(Dh).to(DEV), 1)
    b2 = engine.bomp_encode(X, torch.from_numpy(Dh * np.float32(4096.0)).to(DEV), 1)
    assert int((b1.idx != b2.idx).sum()) <= N // 500


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_n_jobs_spreads_host_arrays_over_gpus_bit_identically():
    """sparse_encoder(n_jobs=G).encode(numpy): contiguous column blocks over G GPUs, one host thread per GPU (the
    analogue of run_parallel's process pool, lyssa/utils/__init__.py:92-146).  The per-device pipelines take their own
    locks, so the devices run concurrently; the result must equal the one-GPU result bit for bit."""
    import time
    N, n, K, k = 400000, 64, 1024, 5
    X = lo.synthetic_patches(N, n, seed=17); D = lo.synthetic_dictionary(K, n, seed=18)
    one = sparse_encoder("bomp", {"n_nonzero_coefs": k}, n_jobs=1, verbose=False)
    two = sparse_encoder("bomp", {"n_nonzero_coefs": k}, n_jobs=2, verbose=False)
    i1, v1, s1 = one.encode_sparse_host(X, D)
    i2, v2, s2 = two.encode_sparse_host(X, D)
    assert np.array_equal(i1, i2) and np.array_equal(v1, v2) and np.array_equal(s1, s2)
    Z1 = one.encode(X[:, :50000], D); Z2 = two.encode(X[:, :50000], D)
    assert np.array_equal(Z1, Z2)
    two.encode_sparse_host(X, D)                                # warm both pipelines
    t0 = time.perf_counter(); one.encode_sparse_host(X, D); t1 = time.perf_counter(); two.encode_sparse_host(X, D); t2 = time.perf_counter()
    print("n_jobs=1 %.1f ms, n_jobs=2 %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
