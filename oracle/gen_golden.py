"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz from the LIVE reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
Every output array below is produced by the reference's own functions, loaded by
oracle/ref_loader.py (in-memory py3 token edits only; see that file's header).  The fixtures
pin (1) the NumPy oracle (tests/test_oracle.py, CPU) and (2) the CUDA path (tests -m gpu).
The reference's own tests hold no numbers for this path, so these vectors
are the known-answer tests SURVEY.md §8c asks the new repo to author.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import ref_loader as rl
from . import lyssa_oracle as lo

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _sparse(Z, k):
    """dense (K,N) -> idx/val (N,k) in ascending atom order, -1 padded (order-free support)."""
    K, N = Z.shape
    idx = -np.ones((N, k), dtype=np.int32)
    val = np.zeros((N, k))
    for i in range(N):
        nz = np.flatnonzero(Z[:, i])
        idx[i, :len(nz)] = nz
        val[i, :len(nz)] = Z[nz, i]
    return idx, val


def _bomp(ref, X, D, k):
    with rl.quiet():
        return ref.sparse_encoder(algorithm="bomp", params={"n_nonzero_coefs": k}, verbose=False).encode(X, D)


def thresh_section(ref):
    """the 'thresh' and 'iht' coders of the live reference on seeded patches
    (n=64; K=256 exercises the fp32 GEMM front end, K=100 the generic shapes)."""
    out = {}
    for tag, K, N in (("a", 256, 200), ("b", 100, 120)):
        X = lo.synthetic_patches(N, 64, seed=40 + K)
        D = lo.synthetic_dictionary(K, 64, seed=41 + K)
        out["X_" + tag], out["D_" + tag] = np.ascontiguousarray(X), D
        cases = (("thresh_k5", "thresh", {"n_nonzero_coefs": 5}, 5),
                 ("thresh_p10", "thresh", {"nonzero_percentage": 0.1}, int(np.floor(0.1 * K))),
                 ("iht_k5", "iht", {"n_nonzero_coefs": 5, "eta": 0.2, "n_iter": 4}, 5),
                 ("iht_k3_it0", "iht", {"n_nonzero_coefs": 3, "eta": 0.2, "n_iter": 0}, 3))
        for name, alg, params, k in cases:
            with rl.quiet():
                Z = ref.sparse_encoder(algorithm=alg, params=dict(params), verbose=False).encode(X.astype(float), D.astype(float))
            out["idx_%s_%s" % (name, tag)], out["val_%s_%s" % (name, tag)] = _sparse(Z, k)
    np.savez_compressed(os.path.join(OUT, "thresh.npz"), **out)


def ksvd_exact_section(ref):
    """one sweep of the live reference's exact ksvd() (ksvd.py:19-43) from seeded (X, D, Z),
    one atom unused; the global NumPy RNG (which randomized_svd draws from) is seeded with 5."""
    X = lo.synthetic_patches(300, 64, seed=60).astype(np.float64)
    D0 = lo.synthetic_dictionary(32, 64, seed=61).astype(np.float64)
    Z0 = lo.sparse_encoder("bomp", {"n_nonzero_coefs": 3}, verbose=False).encode(X, D0)
    Z0[9, :] = 0
    np.random.seed(5)
    with rl.quiet():
        D1, Z1, unused = ref.ksvd(X, D0.copy(), Z0.copy(), n_cycles=1, verbose=False)
    np.savez_compressed(os.path.join(OUT, "ksvd_exact.npz"), X=X.astype(np.float32), D0=D0, Z0=Z0, D1=D1, Z1=Z1,
                        unused=np.array(unused, dtype=np.int32), seed=5)


def gen_omp(ref):
    """'omp' (the reference's default algorithm, sparse_coding.py:618-625 -> :19-66) and feature_encoder's
    soft_thresholding (feature_encoding.py:26-37, loaded from the reference file) through the LIVE reference."""
    X = lo.synthetic_patches(300, 64, seed=60)
    D = lo.synthetic_dictionary(256, 64, seed=61)
    Xd, Dd = X.astype(float), D.astype(float)
    out = {"X": np.ascontiguousarray(X), "D": D}
    for tag, params in (("k5", {"n_nonzero_coefs": 5}), ("tol1p2", {"tol": 1.2}), ("tol0p8", {"tol": 0.8})):
        Z = ref.sparse_encoder(algorithm="omp", params=params, verbose=False).encode(Xd, Dd)
        out["Z_" + tag] = Z.astype(np.float32)
    Ds = (D * np.linspace(0.5, 2.0, 256, dtype=np.float32)[None, :]).astype(np.float32)      # NOT unit norm: omp is still exact least squares
    out["D_scaled"] = Ds
    out["Z_scaled_k4"] = ref.sparse_encoder(algorithm="omp", params={"n_nonzero_coefs": 4}, verbose=False).encode(Xd, Ds.astype(float)).astype(np.float32)
    out["Z_soft_k7"] = ref.soft_thresholding(np.dot(Dd.T, Xd), n_nonzero_coefs=7).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "omp.npz"), **out)


def main():
    ref = rl.load()
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["thresh"]:
        thresh_section(ref)
        return
    if sys.argv[1:] == ["ksvd_exact"]:
        ksvd_exact_section(ref)
        return

    # ---- (7) seeded random, BASELINE cfg1 shape cut to 512 signals: n=64, K=256, k=5
    X = lo.synthetic_patches(512, 64, seed=0)
    D = lo.synthetic_dictionary(256, 64, seed=1)
    Z = _bomp(ref, X.astype(float), D.astype(float), 5)
    idx, val = _sparse(Z, 5)
    np.savez_compressed(os.path.join(OUT, "bomp_cfg1.npz"), X=np.ascontiguousarray(X), D=D, k=5, idx=idx, val=val)

    # ---- K=1024 (cfg2 dictionary size), k=5 and k=10 (cfg3 coder), 256 signals
    X = lo.synthetic_patches(256, 64, seed=2)
    D = lo.synthetic_dictionary(1024, 64, seed=3)
    out = {"X": np.ascontiguousarray(X), "D": D}
    for k in (5, 10):
        Z = _bomp(ref, X.astype(float), D.astype(float), k)
        out["idx_k%d" % k], out["val_k%d" % k] = _sparse(Z, k)
    np.savez_compressed(os.path.join(OUT, "bomp_K1024.npz"), **out)

    # ---- cfg4 shape: n=128, K=2048, k=5, SIFT-like descriptors, 128 signals
    X = lo.synthetic_descriptors(128, 128, seed=4)
    rng = np.random.default_rng(5)
    Dd = np.abs(rng.standard_normal((128, 2048)))
    Dd = np.ascontiguousarray(lo.norm_cols(Dd).astype(np.float32))
    Z = _bomp(ref, X.astype(float), Dd.astype(float), 5)
    idx, val = _sparse(Z, 5)
    np.savez_compressed(os.path.join(OUT, "bomp_cfg4.npz"), X=np.ascontiguousarray(X), D=Dd, k=5, idx=idx, val=val)

    # ---- KATs; small, dense Z stored
    kat = {}
    rng = np.random.default_rng(11)
    # (1) orthonormal D: Z = top-k of alpha
    Q, _ = np.linalg.qr(rng.standard_normal((16, 16)))
    Xo = rng.standard_normal((16, 40))
    kat["ortho_D"], kat["ortho_X"] = Q.astype(np.float32), Xo.astype(np.float32)
    kat["ortho_Z"] = _bomp(ref, kat["ortho_X"].astype(float), kat["ortho_D"].astype(float), 3)
    # (2) exact ties in |alpha| -> lowest index wins: identity dictionary, repeated magnitudes
    I8 = np.eye(8, dtype=np.float32)
    Xt = np.array([[3, -3, 3, 1, 0, 0, 0, 0],
                   [0, 2, 2, -2, 2, 0, 0, 1],
                   [1, 1, 1, 1, 1, 1, 1, 1],
                   [0, 0, 0, 0, 0, 0, 5, -5]], dtype=np.float32).T
    kat["ties_D"], kat["ties_X"] = I8, np.ascontiguousarray(Xt)
    kat["ties_Z"] = _bomp(ref, Xt.astype(float), I8.astype(float), 2)
    # (3) k = K, the reference's own test regime (dict_learning/tests/test_dictionary_learn.py:11-21)
    Xr = rng.random((10, 100)).astype(np.float32)
    Dr = lo.norm_cols(rng.random((10, 4))).astype(np.float32)
    kat["kK_X"], kat["kK_D"] = Xr, np.ascontiguousarray(Dr)
    kat["kK_Z"] = _bomp(ref, Xr.astype(float), kat["kK_D"].astype(float), 4)
    # (4) k = 1 and k = 2 (special-cased branches sparse_coding.py:330-338,360-363)
    Xs = lo.synthetic_patches(64, 64, seed=6)
    Ds = lo.synthetic_dictionary(128, 64, seed=7)
    kat["k12_X"], kat["k12_D"] = np.ascontiguousarray(Xs), Ds
    kat["k1_Z"] = _bomp(ref, Xs.astype(float), Ds.astype(float), 1)
    kat["k2_Z"] = _bomp(ref, Xs.astype(float), Ds.astype(float), 2)
    # (5) degenerate columns: signal == 2.5 * atom 7, zero signal, signal in span of 2 atoms
    Xd = np.stack([2.5 * Ds[:, 7], np.zeros(64, np.float32), Ds[:, 3] - 0.5 * Ds[:, 90]], axis=1).astype(np.float32)
    kat["degen_X"] = np.ascontiguousarray(Xd)
    kat["degen_Z"] = _bomp(ref, Xd.astype(float), Ds.astype(float), 4)
    # (6) duplicate atoms -> pivot < eps break (sparse_coding.py:335,345)
    Dd2 = Ds.copy(); Dd2[:, 1] = Dd2[:, 0]; Dd2[:, 5] = Dd2[:, 4]
    kat["dup_D"] = Dd2
    kat["dup_Z"] = _bomp(ref, Xs.astype(float), Dd2.astype(float), 3)
    # (10) normalise of a zero vector -> zeros (utils/math.py:61-62)
    kat["norm_zero"] = ref.normalize(np.zeros(5))
    kat["norm_vec_in"] = np.array([3.0, 4.0, 0.0])
    kat["norm_vec_out"] = ref.normalize(np.array([3.0, 4.0, 0.0]))
    M = rng.standard_normal((6, 5)); M[:, 2] = 0
    kat["norm_cols_in"] = M.copy()
    kat["norm_cols_out"] = ref.norm_cols(M.copy())
    np.savez_compressed(os.path.join(OUT, "bomp_kat.npz"), **kat)

    # ---- (8) approx-K-SVD: one sweep and two cycles on seeded (X, D, Z) incl. unused atoms
    Xk = lo.synthetic_patches(400, 64, seed=8)
    Dk = lo.synthetic_dictionary(96, 64, seed=9)
    Dk[:, 50] = Dk[:, 0]; Dk[:, 77] = Dk[:, 3]      # duplicates lose every argmax tie -> unused atoms (ksvd.py:112-115)
    Zk = _bomp(ref, Xk.astype(float), Dk.astype(float), 4)
    ks = {"X": np.ascontiguousarray(Xk), "D": Dk, "k": 4}
    ks["Z_idx"], ks["Z_val"] = _sparse(Zk, 4)
    for cyc in (1, 2):
        D1, Z1 = Dk.astype(float).copy(), Zk.copy()
        with rl.quiet():
            _, _, unused = ref.approx_ksvd(Xk.astype(float), D1, Z1, n_cycles=cyc, verbose=False)
        ks["D_c%d" % cyc] = D1
        ks["Zval_c%d" % cyc] = Z1[np.maximum(ks["Z_idx"], 0), np.arange(400)[:, None]] * (ks["Z_idx"] >= 0)
        ks["unused_c%d" % cyc] = np.array(unused, dtype=np.int32)
        ks["err_c%d" % cyc] = ref.approx_error(D1, Z1, Xk.astype(float))
    np.savez_compressed(os.path.join(OUT, "ksvd_sweep.npz"), **ks)

    # ---- K-SVD outer loop: array init (no RNG unless atoms go unused) and 'data' init (seeded)
    kl = {"X": np.ascontiguousarray(Xk), "D0": Dk}
    for tag, init, seed in (("arr", Dk.astype(float), 21), ("data", "data", 22)):
        np.random.seed(seed)
        with rl.quiet():
            Dl, Zl = ref.ksvd_dict_learn(Xk.astype(float), 96, init_dict=init,
                                         sparse_coder=ref.sparse_encoder("bomp", {"n_nonzero_coefs": 4}, verbose=False),
                                         max_iter=3, approx=True, n_cycles=1, verbose=False)
        kl["D_%s" % tag] = Dl
        kl["err_%s" % tag] = ref.approx_error(Dl, Zl, Xk.astype(float))
        kl["seed_%s" % tag] = seed
    np.savez_compressed(os.path.join(OUT, "ksvd_learn.npz"), **kl)

    # ---- (9) ODL: 3 minibatches x 2 epochs, beta=None schedule and beta=0.9 with non_neg
    od = {"X": np.ascontiguousarray(Xk[:, :384]), "D0": Dk}
    for tag, beta, nn in (("lin", None, False), ("b09nn", 0.9, True)):
        D0 = Dk.astype(float).copy()
        with rl.quiet():
            Do, Ao, Bo = ref.online_dict_learn(Xk[:, :384].astype(float), 96,
                                               sparse_coder=ref.sparse_encoder("bomp", {"n_nonzero_coefs": 4}, verbose=False),
                                               batch_size=128, D_init=D0, beta=beta, n_epochs=2, non_neg=nn)
        od["D_%s" % tag], od["A_%s" % tag], od["B_%s" % tag] = Do, Ao, Bo
    # single-minibatch update from given sufficient statistics (isolates K12/K13)
    np.savez_compressed(os.path.join(OUT, "odl.npz"), **od)

    # ---- init_dictionary('data') with a seeded global RNG (dict_learning/utils.py:49-70)
    Xi = np.ascontiguousarray(lo.synthetic_patches(200, 16, seed=10)).copy()
    Xi[:, 5] = 0; Xi[:, 17] = 0            # tiny-norm columns are not candidates (:55)
    np.random.seed(33)
    Di, unused = ref.init_dictionary(Xi.astype(float), 12, method="data", return_unused_data=True)
    np.savez_compressed(os.path.join(OUT, "init_dict.npz"), X=Xi, D=Di, unused=np.array(unused, dtype=np.int32), seed=33)



    # ---- ScSPM pooling: the reference's sc_spm_extractor on 5 small synthetic images,
    # grid descriptors (stand-in for dense SIFT), K=96, k=3, levels (1,2,4), three pooling operators
    imgs = lo.synthetic_images(5, seed=3)
    fe = lo.grid_descriptor_extractor(step_size=4, patch_size=8)
    D = lo.synthetic_dictionary(96, 64, seed=5)
    out = {"D": D, "k": 3, "levels": np.array([1, 2, 4]), "step_size": 4, "patch_size": 8, "n_imgs": len(imgs)}
    for i, im in enumerate(imgs):
        out["img%d" % i] = im
    for name, op, nrm in (("absmax_l2", ref.sc_max_pooling(), True), ("sum", ref.sum_pooling(), False),
                          ("avg_l2", ref.average_pooling(), True)):
        with rl.quiet():
            enc = ref.sparse_encoder(algorithm="bomp", params={"n_nonzero_coefs": 3}, verbose=False)
            out["Z_" + name] = ref.sc_spm_extractor(feature_extractor=fe, levels=(1, 2, 4), sparse_coder=enc, pooling_operator=op,
                                                    normalizer=ref.l2_normalizer() if nrm else None).encode(imgs, D.astype(float))
    np.savez_compressed(os.path.join(OUT, "spm.npz"), **out)

    # ---- dense SIFT: the reference's DsiftExtractor on one 60x75 synthetic image
    img = lo.synthetic_images(1, seed=7, sizes=((60, 75),))[0] * 255.0
    out = {"img": img}
    for gs, ps in ((6, 16), (4, 8)):
        f, p = ref.DsiftExtractor(grid_spacing=gs, patch_size=ps).process_image(img)
        out["feat_%d_%d" % (gs, ps)] = f
        out["pos_%d_%d" % (gs, ps)] = p
    np.savez_compressed(os.path.join(OUT, "dsift.npz"), **out)

    thresh_section(ref)
    ksvd_exact_section(ref)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
