"""TEST INFRASTRUCTURE ONLY — CPU (NumPy, float64) restatement of the reference's hot path.

This is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``lyssandra_b200``) never does, and fails loudly when its CUDA library is missing.

Parity status: **pinned against the reference itself.**  The reference (Python 2, no tests
with numbers — SURVEY.md §4) is executed in the build container by ``oracle/ref_loader.py``
(in-memory py3 token edits only) and this restatement is compared with it bit-for-bit /
to 1e-12 in ``tests/test_oracle_vs_reference.py`` and through the committed fixtures
``tests/golden/*.npz`` written by ``oracle/gen_golden.py``.

Every function cites the reference lines it follows (paths relative to /root/reference).
All arithmetic is float64 exactly as in the reference ("datapoints in columns").
"""
from __future__ import annotations

import itertools
import multiprocessing
import os

import numpy as np
from scipy.linalg import solve_triangular

F64_EPS = np.finfo(float).eps


# --------------------------------------------------------------------------- utils/math.py
def normalize(x, eps=F64_EPS):
    """x / (||x||_2 + eps) — lyssa/utils/math.py:61-62 (zero vector -> zeros, never NaN)."""
    x = np.asarray(x, dtype=float)
    return x / (np.sqrt(np.dot(x, x)) + eps)


def norm_cols(M, eps=F64_EPS):
    """In-place column normalisation with the +eps convention — lyssa/utils/math.py:65-71."""
    scale = np.sqrt(np.einsum("ij,ij->j", M, M)) + eps
    M /= scale[np.newaxis, :]
    return M


def frobenius_squared(M):
    """sum(M**2) — lyssa/utils/math.py:57-58."""
    return np.sum(np.power(M, 2))


# ----------------------------------------------------------------------- utils/__init__.py
def gen_even_batches(n_items, n_batches):
    """n_batches contiguous ranges, the last one takes the remainder —
    lyssa/utils/__init__.py:166-180 (n_items < n_batches -> leading empty ranges, quirk Q8)."""
    width = int(np.floor(n_items / float(n_batches)))
    bounds = [(b * width, (b + 1) * width) for b in range(n_batches - 1)]
    bounds.append(((n_batches - 1) * width, n_items))
    return [range(lo, hi) for lo, hi in bounds]


def gen_batches(n_items, batch_size=None):
    """Fixed-size contiguous ranges + one short tail — lyssa/utils/__init__.py:183-201."""
    if batch_size is None:
        return [range(0, n_items)]
    full = int(np.floor(n_items / float(batch_size)))
    out = [range(b * batch_size, (b + 1) * batch_size) for b in range(full)]
    if n_items > full * batch_size:
        out.append(range(full * batch_size, n_items))
    return out


# ------------------------------------------------------------------------ sparse_coding.py
def batch_omp(X, Alpha, D, Gram, n_nonzero_coefs=None, tol=None, trace=None):
    """Batch-OMP over the columns of Alpha — lyssa/sparse_coding.py:302-367.

    Per signal: greedy argmax of |a| (first maximum wins, :322); stop if the atom is
    already selected (:323-325); incremental Cholesky row by forward substitution with a
    *literal 1* for the atom self-product (:330-349, quirk Q1) and stop when
    1 - w.w < eps (:335,:345); coefficients by two triangular solves (:353-354);
    a = a0 - G[:, I] z (:359).  ``tol`` is accepted and ignored (:302,:536).  X and D are
    not read (as in the reference).

    ``trace`` (optional dict) receives per-signal decision margins used by the parity tests'
    near-tie policy: 'gap' (N,k) relative top1-top2 gap of |a| at each executed step,
    'vs' (N,k) the Cholesky pivot 1 - w.w, 'nsel' (N,) number of selected atoms.
    """
    n_atoms, n_signals = Alpha.shape
    k_max = n_nonzero_coefs
    Z = np.zeros((n_atoms, n_signals))
    if trace is not None:
        trace["gap"] = np.full((n_signals, k_max), np.inf)
        trace["vs"] = np.full((n_signals, k_max), 1.0)
        trace["nsel"] = np.zeros(n_signals, dtype=int)

    for i in range(n_signals):
        support = np.array([]).astype(int)
        a0 = Alpha[:, i]
        a = a0
        L = np.zeros((k_max, k_max))
        for j in range(k_max):
            mag = np.abs(a)
            pick = np.argmax(mag)                                    # :322
            if trace is not None:
                top = mag[pick]
                mag2 = mag.copy(); mag2[pick] = -1.0
                trace["gap"][i, j] = (top - mag2.max()) / top if top > 0 else 0.0
            if pick in support:                                      # :323-325
                break
            g = Gram[support, pick]                                  # :327
            if j == 0:                                               # :360-363
                support = np.append(support, pick)
                z = a0[support]
                a = a0 - np.dot(Gram[:, support], z)
                continue
            if j == 1:                                               # :330-338
                w = g[0]
                pivot = 1 - w * w
                if trace is not None:
                    trace["vs"][i, j] = pivot
                if pivot < F64_EPS:
                    break
                L[:2, :2] = [[1, 0], [w, np.sqrt(pivot)]]
            else:                                                    # :340-349
                w = solve_triangular(L[:j, :j], g, lower=True, check_finite=False)
                pivot = 1 - np.dot(w, w)
                if trace is not None:
                    trace["vs"][i, j] = pivot
                if pivot < F64_EPS:
                    break
                L[j, :j] = w
                L[j, j] = np.sqrt(pivot)
            support = np.append(support, pick)                       # :351
            y = solve_triangular(L[:j + 1, :j + 1], a0[support], lower=True)          # :353
            z = solve_triangular(L[:j + 1, :j + 1], y, trans=1, lower=True)           # :354
            a = a0 - np.dot(Gram[:, support], z)                     # :359
        Z[support, i] = z                                            # :365
        if trace is not None:
            trace["nsel"][i] = len(support)
    return Z


def _omp(x, D, Gram, alpha, n_nonzero_coefs=None, tol=None):
    """Plain OMP of one signal — lyssa/sparse_coding.py:19-57.  Stopping rule (:27-34): with n_nonzero_coefs the
    tolerance is forced to 1e-10 and the loop runs while i < n_nonzero_coefs and ||r|| > 1e-10; with only ``tol`` it
    runs while ||r|| >= tol.  Per step: first-maximum argmax of |alpha| (:40); stop if already selected (:41-42);
    z = inv(G[I,I]) (D^T x)[I] with the TRUE Gram (no unit-norm assumption, unlike batch_omp) (:45-53); r = x - D_I z,
    alpha = D^T r (:54-55).  A singular G[I,I] (LinAlgError) stops with the previous z (:48-51)."""
    n_atoms = D.shape[1]
    support = np.array([]).astype(int)
    z = np.zeros(n_atoms)
    r = np.copy(x)
    i = 0
    if n_nonzero_coefs is not None:
        tol = 1e-10

        def cont():
            return i < n_nonzero_coefs and np.linalg.norm(r) > tol
    else:
        def cont():
            return np.linalg.norm(r) >= tol
    while cont():
        pick = np.argmax(np.abs(alpha))                              # :40
        if pick in support:                                          # :41-42
            break
        support = np.append(support, pick)
        G = np.atleast_2d(Gram[support, :][:, support])              # :45-46
        try:
            G_inv = np.linalg.inv(G)                                 # :48
        except np.linalg.LinAlgError:
            break
        z[support] = np.dot(G_inv, np.dot(D.T, x)[support])          # :53
        r = x - np.dot(D[:, support], z[support])                    # :54
        alpha = np.dot(D.T, r)                                       # :55
        i += 1
    return z


def omp(X, Alpha, D, Gram, n_nonzero_coefs=None, tol=None):
    """lyssa/sparse_coding.py:60-66: _omp over the columns of X."""
    n_atoms, n_signals = D.shape[1], X.shape[1]
    Z = np.zeros((n_atoms, n_signals))
    for i in range(n_signals):
        Z[:, i] = _omp(X[:, i], D, Gram, Alpha[:, i], n_nonzero_coefs=n_nonzero_coefs, tol=tol)
    return Z


def soft_thresholding(Alpha, nonzero_percentage=None, n_nonzero_coefs=None):
    """lyssa/feature_encoding.py:26-37 — line for line the same computation as ``thresholding``
    (sparse_coding.py:416-425): the n_nonzero_coefs largest SIGNED correlations of every signal."""
    return thresholding(Alpha, nonzero_percentage=nonzero_percentage, n_nonzero_coefs=n_nonzero_coefs)


class feature_encoder(object):
    """lyssa/feature_encoding.py:40-89: algorithm 'soft_thresholding' = soft_thresholding(D^T X)."""

    def __init__(self, algorithm=None, params=None, n_jobs=1, verbose=True, mmap=False):
        self.algorithm, self.params = algorithm, ({} if params is None else params)
        self.n_jobs, self.verbose, self.mmap = n_jobs, verbose, mmap

    def encode(self, X, D):
        return self.__call__(X, D)

    def __call__(self, X, D):
        if self.algorithm != "soft_thresholding":                   # the reference leaves `func` unbound here (:64-70)
            raise NameError("feature_encoder: unknown algorithm %r" % (self.algorithm,))
        return soft_thresholding(np.dot(D.T, X), nonzero_percentage=self.params.get("nonzero_percentage"),
                                 n_nonzero_coefs=self.params.get("n_nonzero_coefs"))


def _bomp_job(args):
    alpha_cols, gram, k = args
    return batch_omp(None, alpha_cols, None, gram, n_nonzero_coefs=k)


def _single_thread_blas():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:  # pragma: no cover
        pass


def thresholding(Alpha, nonzero_percentage=None, n_nonzero_coefs=None):
    """Keep the n_nonzero_coefs largest SIGNED correlations of every signal —
    lyssa/sparse_coding.py:416-425 (argsort ascending, reversed, first k)."""
    n_atoms, n_samples = Alpha.shape
    Z = np.zeros((n_atoms, n_samples))
    if nonzero_percentage is not None:
        n_nonzero_coefs = int(np.floor(nonzero_percentage * n_atoms))      # :419-420
    for i in range(n_samples):
        keep = Alpha[:, i].argsort()[::-1][:n_nonzero_coefs]                # :423
        Z[keep, i] = Alpha[keep, i]
    return Z


def iterative_hard_thresh(X, Z0, R0, D, eta=None, n_nonzero_coefs=None, n_iter=None):
    """Z <- Z - eta D^T R; zero all but the k largest |Z| per signal; R = D Z - X —
    lyssa/sparse_coding.py:433-446."""
    Z, R = Z0, R0
    for _ in range(n_iter):
        Z -= eta * np.dot(D.T, R)                                           # :439
        for i in range(X.shape[1]):
            drop = np.abs(Z[:, i]).argsort()[::-1][n_nonzero_coefs:]        # :442
            Z[drop, i] = 0
        R = np.dot(D, Z) - X                                                # :444
    return Z


class sparse_encoder(object):
    """The 'bomp' branch of lyssa.sparse_coding.sparse_encoder — sparse_coding.py:587-603,
    :629-635 (Gram, Alpha, partial(batch_omp)), :708-726 (run_parallel with n_batches=100) —
    plus the two thresholding coders on the same correlations: 'thresh' (:636-641) and 'iht'
    (:671-690).  Unknown algorithms raise ValueError (:706).  n_jobs>1 ('bomp') reproduces
    run_parallel's regime — a process pool over 100 contiguous column batches with one BLAS
    thread per worker (lyssa/utils/__init__.py:92-146, sparse_coding.py:713-716)."""

    def __init__(self, algorithm="omp", params=None, n_jobs=1, verbose=True, mmap=False, name="sparse_coder"):
        self.name = name
        self.algorithm = algorithm
        self.params = {} if params is None else params
        if n_jobs == -1:
            n_jobs = multiprocessing.cpu_count()
        self.n_jobs = n_jobs
        self.verbose = verbose
        self.mmap = mmap

    def encode(self, X, D):
        return self.__call__(X, D)

    def __call__(self, X, D):
        if self.algorithm in ("thresh", "iht"):
            alpha = np.dot(D.T, X)                                   # :637 / :672
            Z = thresholding(alpha, n_nonzero_coefs=self.params.get("n_nonzero_coefs"),
                             nonzero_percentage=self.params.get("nonzero_percentage"))
            if self.algorithm == "iht":
                R0 = np.dot(D, Z) - X                                # :682
                Z = iterative_hard_thresh(X, Z, R0, D, eta=self.params.get("eta"),
                                          n_nonzero_coefs=self.params.get("n_nonzero_coefs"),
                                          n_iter=self.params.get("n_iter"))
            return Z
        if self.algorithm == "omp":                                  # :618-625
            gram = np.dot(D.T, D)
            alpha = np.dot(D.T, X)
            return omp(X, alpha, D, gram, n_nonzero_coefs=self.params.get("n_nonzero_coefs"), tol=self.params.get("tol"))
        if self.algorithm != "bomp":
            if self.algorithm in ("nnomp", "group_omp", "sparse_group_omp", "somp", "lasso", "llc"):
                raise NotImplementedError("oracle restates only 'omp', 'bomp', 'thresh' and 'iht'")
            raise ValueError("Sparse optimizer not found.")
        k = self.params.get("n_nonzero_coefs")
        n_atoms, n_signals = D.shape[1], X.shape[1]
        gram = np.dot(D.T, D)                                        # :630
        alpha = np.dot(D.T, X)                                       # :631
        if self.n_jobs == 1:                                         # utils/__init__.py:78-90
            Z = np.zeros((n_atoms, n_signals))
            Z[:] = batch_omp(X, alpha, D, gram, n_nonzero_coefs=k, tol=self.params.get("tol"))
            return Z
        Z = np.zeros((n_atoms, n_signals))
        parts = gen_even_batches(n_signals, 100)                     # :93-97
        ctx = multiprocessing.get_context("fork")
        with ctx.Pool(processes=self.n_jobs, initializer=_single_thread_blas) as pool:
            jobs = [(alpha[:, r.start:r.stop], gram, k) for r in parts]
            for r, out in zip(parts, pool.imap(_bomp_job, jobs)):
                Z[:, r.start:r.stop] = out                           # :138-146
        return Z


# ------------------------------------------------------------------ dict_learning/utils.py
def approx_error(D, Z, X, n_jobs=1):
    """||X - D Z||_F^2 — lyssa/dict_learning/utils.py:14-19."""
    return frobenius_squared(X - np.dot(D, Z))


def average_mutual_coherence(D):
    """mean off-diagonal |D^T D| — lyssa/dict_learning/utils.py:7-11."""
    K = D.shape[1]
    C = np.abs(np.dot(D.T, D))
    np.fill_diagonal(C, 0)
    return np.sum(C) / float(K * (K - 1))


def init_dictionary(X, n_atoms, method="data", return_unused_data=False, normalize=True):
    """method='data' — lyssa/dict_learning/utils.py:49-70: candidate columns have
    sum(x^2) > 1e-6 (:55); ``np.random.choice(len(cands), n_atoms, replace=False)`` on the
    GLOBAL NumPy RNG (:61, quirk Q9); fancy-index copy (:64); norm_cols (:65-66); the
    unused candidates come back as a Python list (:67-70)."""
    if method != "data":
        raise NotImplementedError("oracle restates only method='data'")
    n_signals = X.shape[1]
    cands = [i for i in range(n_signals) if np.sum(X[:, i] ** 2) > 1e-6]
    if len(cands) < n_atoms:
        raise ValueError("not enough datapoints to initialize the dictionary")
    chosen = np.random.choice(len(cands), size=n_atoms, replace=False)
    chosen_cols = np.array(cands).astype(int)[chosen]
    D = X[:, chosen_cols]
    if normalize:
        D = norm_cols(D)
    if return_unused_data:
        taken = set(chosen_cols)
        return D, [c for c in cands if c not in taken]
    return D


# ------------------------------------------------------------------- dict_learning/ksvd.py
def approx_ksvd(Y, D, X, n_cycles=1, verbose=False):
    """Approximate K-SVD sweep, sequential over atoms, D and X mutated IN PLACE —
    lyssa/dict_learning/ksvd.py:98-126.  R = Y - D X (:103); per atom: users = X[k]!=0
    (:111), unused atoms recorded and skipped (:112-115), Rk = R[:,users] + d x (:116),
    d <- normalize(Rk x) (:118-119), x <- Rk^T d (:121), R[:,users] = Rk - d x (:123)."""
    n_atoms = D.shape[1]
    unused = []
    R = Y - np.dot(D, X)
    for _ in range(n_cycles):
        for k in range(n_atoms):
            users = X[k, :] != 0
            if not np.any(users):
                unused.append(k)
                continue
            Rk = R[:, users] + np.outer(D[:, k], X[k, users])
            D[:, k] = np.dot(Rk, X[k, users])
            D[:, k] = normalize(D[:, k])
            X[k, users] = np.dot(Rk.T, D[:, k])
            R[:, users] = Rk - np.outer(D[:, k], X[k, users])
    return D, X, unused


def ksvd(Y, D, X, n_cycles=1, verbose=False, svd="randomized"):
    """Exact K-SVD sweep, D and X mutated IN PLACE — lyssa/dict_learning/ksvd.py:19-43 (SURVEY.md section 8f
    row 4; no device path yet, this is the checker it will be built against).  Per atom: users = X[k]!=0 (:31),
    unused atoms recorded and skipped (:32-34), Rk = R[:,users] + d x (:36), (d, x) <- top singular triplet of Rk
    (:37-39), R[:,users] = Rk - d x (:41).  The reference takes the triplet from scikit-learn's
    ``randomized_svd(Rk, n_components=1, n_iter=10, flip_sign=False)`` (a third-party dependency, not pinned by the
    reference; 1.9.0 in this image) with the global NumPy RNG: the common sign of (d, x) is arbitrary, d x^T is not.
    ``svd="lapack"`` takes it from np.linalg.svd instead (deterministic; same d x^T wherever sigma_1 > sigma_2)."""
    n_atoms = D.shape[1]
    unused = []
    R = Y - np.dot(D, X)
    for _ in range(n_cycles):
        for k in range(n_atoms):
            users = X[k, :] != 0
            if not np.any(users):
                unused.append(k)
                continue
            Rk = R[:, users] + np.outer(D[:, k], X[k, users])
            if svd == "randomized":
                from sklearn.utils.extmath import randomized_svd
                U, S, V = randomized_svd(Rk, n_components=1, n_iter=10, flip_sign=False)          # :37
            else:
                U, S, V = np.linalg.svd(Rk, full_matrices=False)
            D[:, k] = U[:, 0]
            X[k, users] = V[0, :] * S[0]
            R[:, users] = Rk - np.outer(D[:, k], X[k, users])
    return D, X, unused


def ksvd_dict_learn(X, n_atoms, init_dict="data", sparse_coder=None, max_iter=20, non_neg=False,
                    approx=False, eta=None, n_cycles=1, n_jobs=1, mmap=False, verbose=True,
                    history=None):
    """Outer K-SVD loop — lyssa/dict_learning/ksvd.py:129-231 (approx=True branch only).
    Mirrors: array init_dict is copied (:155); encode (:177); sweep (:186); unused-atom
    replacement by ``np.random.choice(unused_data, size=1)`` + normalize (:199-207);
    approx_error (:220); the patience rule exactly as written, including quirk Q3
    (error_prev only advances when verbose, :222-229).  ``history`` (list) collects the
    per-iteration error."""
    if not approx or non_neg or eta is not None:
        raise NotImplementedError("oracle restates only approx=True, non_neg=False, eta=None")
    unused_data = []
    if isinstance(init_dict, str) and init_dict == "data":
        D, unused_data = init_dictionary(X, n_atoms, method=init_dict, return_unused_data=True)
    else:
        D = np.copy(init_dict)
    Z = np.zeros((n_atoms, X.shape[1]))
    max_patience = 10
    error_curr = 0
    error_prev = 0
    it = 0
    patience = 0
    while it < max_iter and patience < max_patience:
        Z = sparse_coder(X, D)
        D, _, unused_atoms = approx_ksvd(X, D, Z, n_cycles=n_cycles)
        for slot in unused_atoms:
            if len(unused_data) == 0:
                break
            col = np.random.choice(unused_data, size=1)[0]
            D[:, slot] = X[:, col]
            D[:, slot] = normalize(D[:, slot])
            unused_data.remove(col)
        error_curr = approx_error(D, Z, X, n_jobs=2)
        if history is not None:
            history.append(error_curr)
        if verbose:
            error_prev = error_curr
        if (it > 0) and (error_curr > 0.9 * error_prev or error_curr > error_prev):
            patience += 1
        it += 1
    return D, Z


# ------------------------------------------------------- dict_learning/online_dict_learn.py
def online_dict_learn(X, n_atoms, sparse_coder=None, batch_size=None, A=None, B=None, D_init=None,
                      beta=None, n_epochs=1, verbose=False, n_jobs=1, non_neg=False, mmap=False):
    """Mairal online dictionary learning as the reference implements it —
    lyssa/dict_learning/online_dict_learn.py:18-124.  D_init is used WITHOUT copying (:47);
    beta=None -> linspace(0,1,n_iter) restarted every epoch (:65-69, :79, quirk Q6);
    A = b*A + Z Z^T, B = b*B + X Z^T (:84-85); Jacobi block update with the stale D A
    (:91-94); optional clamp (:96-97); norm_cols (:98); end-of-epoch error + patience
    (:101-118, same quirk as K-SVD).  Returns (D, A, B)."""
    sparse_coder.verbose = False
    n_features, n_signals = X.shape
    if D_init is None:
        D, _unused = init_dictionary(X, n_atoms, method="data", return_unused_data=True)
    else:
        D = D_init
    batches = gen_batches(n_signals, batch_size=batch_size)
    n_iter = len(batches)
    if A is None and B is None:
        A = np.zeros((n_atoms, n_atoms))
        B = np.zeros((n_features, n_atoms))
    if beta is None:
        beta = np.linspace(0, 1, num=n_iter)
    else:
        beta = np.zeros(n_iter) + beta
    max_patience = 10
    error_prev = 0
    patience = 0
    for e in range(n_epochs):
        for i, cols in zip(range(n_iter), itertools.cycle(batches)):
            Xb = X[:, cols]
            Zb = sparse_coder(Xb, D)
            A = beta[i] * A + np.dot(Zb, Zb.T)
            B = beta[i] * B + np.dot(Xb, Zb.T)
            DA = np.dot(D, A)
            for k in range(n_atoms):
                D[:, k] = (1 / (A[k, k] + F64_EPS)) * (B[:, k] - DA[:, k]) + D[:, k]
            if non_neg:
                D[D < 0] = 0
            D = norm_cols(D)
        if e < n_epochs - 1:
            if patience >= max_patience:
                return D, A, B
            error_curr = 0
            for i, cols in zip(range(n_iter), itertools.cycle(batches)):
                Xb = X[:, cols]
                Zb = sparse_coder(Xb, D)
                error_curr += approx_error(D, Zb, Xb, n_jobs=n_jobs)
            if verbose:
                error_prev = error_curr
            if (e > 0) and (error_curr > 0.9 * error_prev or error_curr > error_prev):
                patience += 1
    return D, A, B


# ------------------------------------------------------------------------- synthetic data
# ----------------------------------------------------------------------------- ScSPM pooling
class sc_max_pooling(object):
    """lyssa/feature_extract/pooling.py:4-7"""
    def __call__(self, Z):
        return np.max(np.abs(Z), axis=1)


class sum_pooling(object):
    """lyssa/feature_extract/pooling.py:16-19"""
    def __call__(self, Z):
        return np.sum(Z, axis=1)


class average_pooling(object):
    """lyssa/feature_extract/pooling.py:22-26"""
    def __call__(self, Z):
        return np.sum(Z, axis=1) / float(Z.shape[1])


class l2_normalizer(object):
    """lyssa/feature_extract/preproc.py:8-15 (1-D case)"""
    def __call__(self, Z):
        return normalize(Z)


class sc_spm_extractor(object):
    """Restatement of lyssa/feature_extract/spatial_pyramid.py:34-97 (per image: descriptors + positions
    from the feature extractor, sparse codes, cell membership per pyramid level, pool, normalise)."""

    def __init__(self, feature_extractor=None, levels=(1, 2, 4), sparse_coder=None, pooling_operator=None, normalizer=None):
        self.feature_extractor = feature_extractor
        self.levels = levels
        self.sparse_coder = sparse_coder
        self.pooling_operator = pooling_operator
        self.normalizer = normalizer

    def encode(self, imgs, dictionary):
        psize = self.feature_extractor.patch_size                       # :47
        n_imgs = len(imgs)
        n_atoms = dictionary.shape[1]
        cells = np.array(self.levels) ** 2                              # :50
        n_features = np.sum(cells) * n_atoms
        Z = np.zeros((n_features, n_imgs))
        for k in range(n_imgs):                                         # :54
            img = imgs[k]
            desc, pos = self.feature_extractor.extract(img)             # :57
            py = pos[:, 0]
            px = pos[:, 1]
            cy = py + float(psize) / 2 - 0.5                            # :62-63
            cx = px + float(psize) / 2 - 0.5
            coded_patches = self.sparse_coder.encode(desc, dictionary)  # :66
            n_atoms = coded_patches.shape[0]
            n_total_cells = np.sum(cells)
            imsize = img.shape
            poolpatches = np.zeros((n_total_cells, n_atoms))            # :74
            cnt = 0
            for (i, lev) in enumerate(self.levels):                     # :77
                wunit = float(imsize[1]) / lev
                hunit = float(imsize[0]) / lev
                binidx = np.floor(cy / hunit) * lev + np.floor(cx / wunit)      # :83
                for j in range(cells[i]):
                    pidx = np.nonzero(binidx == j)[0]
                    if len(pidx) > 0:
                        poolpatches[cnt, :] = self.pooling_operator(coded_patches[:, pidx])   # :91
                        if self.normalizer is not None:
                            poolpatches[cnt, :] = self.normalizer(poolpatches[cnt, :])
                    cnt += 1
            Z[:, k] = poolpatches.flatten()                             # :96
        return Z


class DsiftExtractor(object):
    """Restatement of lyssa/feature_extract/dsift.py:16-162 (dense SIFT after Lazebnik); the only edit is
    ``np.int`` -> ``int`` (:27, removed from NumPy 1.24)."""
    n_angles = 8
    n_bins = 4
    alpha = 9.0

    @staticmethod
    def gen_dgauss(sigma):                                              # :23-35
        fwid = int(2 * np.ceil(sigma))
        G = np.array(range(-fwid, fwid + 1)) ** 2
        G = G.reshape((G.size, 1)) + G
        G = np.exp(- G / 2.0 / sigma / sigma)
        G /= np.sum(G)
        GH, GW = np.gradient(G)
        GH *= 2.0 / np.sum(np.abs(GH))
        GW *= 2.0 / np.sum(np.abs(GW))
        return GH, GW

    def __init__(self, grid_spacing=None, patch_size=None, nrml_thres=1.0, sigma_edge=0.8, sift_thres=0.2):
        n_bins = self.n_bins
        self.gs = grid_spacing
        self.ps = patch_size
        self.nrml_thres = nrml_thres
        self.sigma = sigma_edge
        self.sift_thres = sift_thres
        sample_res = self.ps / np.double(n_bins)                        # :57-73
        sample_p = np.array(range(self.ps))
        sample_ph, sample_pw = np.meshgrid(sample_p, sample_p)
        sample_ph = sample_ph.reshape(-1)
        sample_pw = sample_pw.reshape(-1)
        bincenter = np.array(range(1, n_bins * 2, 2)) / 2.0 / n_bins * self.ps - 0.5
        bincenter_h, bincenter_w = np.meshgrid(bincenter, bincenter)
        bincenter_h = bincenter_h.reshape((-1, 1))
        bincenter_w = bincenter_w.reshape((-1, 1))
        weights_h = abs(sample_ph - bincenter_h) / sample_res
        weights_w = abs(sample_pw - bincenter_w) / sample_res
        weights_h = (1 - weights_h) * (weights_h <= 1)
        weights_w = (1 - weights_w) * (weights_w <= 1)
        self.weights = weights_h * weights_w

    def process_image(self, image, positionNormalize=False):            # :75-118
        from math import floor
        image = image.astype(np.double)
        if image.ndim == 3:
            image = np.mean(image, axis=2)
        h, w = image.shape
        gs, ps = self.gs, self.ps
        rem_h = np.mod(h - ps, gs)
        rem_w = np.mod(w - ps, gs)
        offset_h = int(floor(rem_h / 2.))
        offset_w = int(floor(rem_w / 2.))
        grid_h, grid_w = np.meshgrid(range(offset_h, h - ps + 1, gs), range(offset_w, w - ps + 1, gs))
        grid_h = grid_h.flatten()
        grid_w = grid_w.flatten()
        feat_arr = self.extract_sift_patches(image, grid_h, grid_w)
        if positionNormalize:
            positions = np.vstack((grid_h / np.double(h), grid_w / np.double(w)))
        else:
            positions = np.vstack((grid_h, grid_w))
        return feat_arr, positions

    def extract_sift_patches(self, image, grid_h, grid_w):              # :120-144
        from scipy import signal
        n_angles, n_samples = self.n_angles, self.n_bins ** 2
        angles = np.array(range(n_angles)) * 2.0 * np.pi / n_angles
        h, w = image.shape
        n_patches = grid_h.size
        feat_arr = np.zeros((n_patches, n_samples * n_angles))
        gh, gw = self.gen_dgauss(self.sigma)
        ih = signal.convolve2d(image, gh, mode='same')
        iw = signal.convolve2d(image, gw, mode='same')
        i_mag = np.sqrt(ih ** 2 + iw ** 2)
        i_theta = np.arctan2(ih, iw)
        i_orient = np.zeros((n_angles, h, w))
        for i in range(n_angles):
            i_orient[i] = i_mag * np.maximum(np.cos(i_theta - angles[i]) ** self.alpha, 0)
        for i in range(n_patches):
            curr_feature = np.zeros((n_angles, n_samples))
            for j in range(n_angles):
                curr_feature[j] = np.dot(self.weights, i_orient[j, grid_h[i]:grid_h[i] + self.ps,
                                                      grid_w[i]:grid_w[i] + self.ps].flatten())
            feat_arr[i] = curr_feature.flatten()
        return self.normalize_sift(feat_arr)

    def normalize_sift(self, feat_arr):                                 # :146-162
        siftlen = np.sqrt(np.sum(feat_arr ** 2, axis=1))
        hcontrast = (siftlen >= self.nrml_thres)
        siftlen[siftlen < self.nrml_thres] = self.nrml_thres
        feat_arr /= siftlen.reshape((siftlen.size, 1))
        feat_arr[feat_arr > self.sift_thres] = self.sift_thres
        feat_arr[hcontrast] /= np.sqrt(np.sum(feat_arr[hcontrast] ** 2, axis=1)).reshape((feat_arr[hcontrast].shape[0], 1))
        return feat_arr


class dsift_extractor(object):
    """lyssa/feature_extract/spatial_pyramid.py:9-20"""

    def __init__(self, step_size=None, patch_size=None):
        self.patch_size = patch_size
        self.extractor = DsiftExtractor(grid_spacing=step_size, patch_size=patch_size)

    def extract(self, img):
        dsift_patches, pos = self.extractor.process_image(img, positionNormalize=False)
        return dsift_patches.T, pos.T


class grid_descriptor_extractor(object):
    """Synthetic stand-in for the reference's dsift_extractor (spatial_pyramid.py:9-20): descriptors on a
    regular grid of patch_size x patch_size patches (top-left positions, step_size apart), here simply the
    mean-removed pixels of the patch.  Returns (desc (n, P), pos (P, 2)) like ``extract`` does."""

    def __init__(self, step_size=4, patch_size=8):
        self.step_size = step_size
        self.patch_size = patch_size

    def extract(self, img):
        img = np.asarray(img, dtype=np.float64)
        H, W = img.shape
        ps, st = self.patch_size, self.step_size
        rows = np.arange(0, H - ps + 1, st)
        cols = np.arange(0, W - ps + 1, st)
        desc = np.empty((ps * ps, len(rows) * len(cols)))
        pos = np.empty((len(rows) * len(cols), 2))
        c = 0
        for r in rows:
            for q in cols:
                patch = img[r:r + ps, q:q + ps].reshape(-1)
                desc[:, c] = patch - patch.mean()
                pos[c] = (r, q)
                c += 1
        return desc, pos


def synthetic_images(n_imgs, seed=0, sizes=((48, 64), (40, 40), (33, 57))):
    """seeded smooth-noise grayscale images of a few different sizes (SURVEY.md section 8d, cfg5 in miniature)"""
    rng = np.random.default_rng(seed)
    imgs = []
    for i in range(n_imgs):
        H, W = sizes[i % len(sizes)]
        a = rng.random((H + 4, W + 4))
        a = (a[:-4, :-4] + a[2:-2, 2:-2] + a[4:, 4:] + a[:-4, 4:] + a[4:, :-4]) / 5.0
        imgs.append(np.ascontiguousarray(a, dtype=np.float64))
    return imgs


def synthetic_patches(n_signals, n_features=64, seed=0):
    """uniform [0,1) pixels, per-patch mean removed; returns float32 (n, N)
    'datapoints in columns' as a transposed view of signal-major storage."""
    rng = np.random.default_rng(seed)
    P = rng.random((n_signals, n_features), dtype=np.float32)
    P -= P.mean(axis=1, keepdims=True)
    return P.T


def synthetic_dictionary(n_atoms, n_features=64, seed=1):
    """K independent draws of the same distribution, mean removed, unit-normalised with the
    +eps convention (utils/math.py:65-71); float32 (n, K) C-contiguous."""
    rng = np.random.default_rng(seed)
    P = rng.random((n_atoms, n_features), dtype=np.float32)
    P -= P.mean(axis=1, keepdims=True)
    D = np.ascontiguousarray(P.T.astype(np.float64))
    norm_cols(D)
    return np.ascontiguousarray(D.astype(np.float32))


def synthetic_descriptors(n_signals, n_features=128, seed=0):
    """SIFT-like descriptors for cfg4: |N(0,1)| then Lowe normalisation (unit norm, clip 0.2,
    renormalise — lyssa/feature_extract/dsift.py:146-162); float32 (n, N) view."""
    rng = np.random.default_rng(seed)
    P = np.abs(rng.standard_normal((n_signals, n_features), dtype=np.float32))
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    np.minimum(P, 0.2, out=P)
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    return P.T
