"""Shared parity policy (SURVEY.md §8c "Parity definitions").

Support sets must be bit-identical on every column whose decisions are not *near-ties* in the
oracle's own float64 arithmetic; near-tie / degenerate columns are counted and reported, not
hidden.  A column is excluded from the support comparison iff, at any executed greedy step,
  * the oracle's relative top-1/top-2 gap of |alpha| is below GAP_TOL (float32 rounding of the
    inputs alone can flip such an argmax), or
  * the oracle's Cholesky pivot 1 - w.w is below PIVOT_TOL (near-duplicate atoms), or
  * the oracle selected fewer than k atoms (it hit a break: degenerate column).
Coefficients: max |Z_gpu - Z_ref| / max |Z_ref| <= COEF_TOL over the compared columns.
"""
import numpy as np

GAP_TOL = 1e-5
PIVOT_TOL = 1e-6
COEF_TOL = 1e-5


def sorted_codes(idx, val):
    """order-free view of (idx,val)[N,k]: ascending atom index, -1 pads last."""
    idx = np.asarray(idx).astype(np.int64)
    val = np.asarray(val, dtype=np.float64)
    key = np.where(idx < 0, np.iinfo(np.int64).max, idx)
    order = np.argsort(key, axis=1, kind="stable")
    return np.take_along_axis(idx, order, axis=1), np.take_along_axis(val, order, axis=1)


def comparable_columns(gap, vs, nsel, k):
    gap = np.asarray(gap); vs = np.asarray(vs); nsel = np.asarray(nsel)
    ok = (np.min(gap, axis=1) >= GAP_TOL) & (np.min(vs, axis=1) >= PIVOT_TOL) & (nsel == k)
    return ok


def check_codes(idx_gpu, val_gpu, idx_ref, val_ref, ok=None, coef_tol=COEF_TOL, label=""):
    """Assert identical supports and close coefficients on the comparable columns.
    Returns a small report dict."""
    ig, vg = sorted_codes(idx_gpu, val_gpu)
    ir, vr = sorted_codes(idx_ref, val_ref)
    N = ig.shape[0]
    if ok is None:
        ok = np.ones(N, dtype=bool)
    same = np.all(ig == ir, axis=1)
    bad = np.flatnonzero(ok & ~same)
    assert bad.size == 0, "%s: %d of %d comparable columns have a different support (first: col %d gpu=%s ref=%s)" % (
        label, bad.size, int(ok.sum()), bad[0], ig[bad[0]], ir[bad[0]])
    both = ok & same
    scale = np.max(np.abs(vr[both])) if both.any() else 1.0
    err = np.max(np.abs(vg[both] - vr[both])) / scale if both.any() else 0.0
    assert err <= coef_tol, "%s: coefficient rel-inf error %.3g > %.3g" % (label, err, coef_tol)
    return {"columns": int(N), "compared": int(both.sum()), "excluded": int((~ok).sum()),
            "mismatch_in_excluded": int((~ok & ~same).sum()), "coef_rel_inf": float(err)}


def thresh_trace(X, D, k, eta=None, n_iter=0):
    """float64 replay of the oracle's 'thresh' (n_iter == 0) / 'iht' coders that also returns, per
    column, the smallest relative gap between the k-th and (k+1)-th selection keys met at any stage
    (signed alpha for the start, |Z| for the iterations).  Columns whose gap is below GAP_TOL are
    near-ties: float32 rounding alone can swap the two entries."""
    X = np.asarray(X, dtype=np.float64); D = np.asarray(D, dtype=np.float64)
    K, N = D.shape[1], X.shape[1]
    A = D.T @ X
    gap = np.full(N, np.inf)

    def keep_top(keys, vals):
        order = np.argsort(-keys, axis=0, kind="stable")
        top = order[:k]
        Z = np.zeros_like(vals)
        np.put_along_axis(Z, top, np.take_along_axis(vals, top, axis=0), axis=0)
        if k < K:
            ks = np.take_along_axis(keys, order[k - 1:k + 1], axis=0)
            g = (ks[0] - ks[1]) / np.maximum(np.max(np.abs(keys), axis=0), 1e-300)
            np.minimum(gap, g, out=gap)
        return Z

    Z = keep_top(A, A)
    for _ in range(int(n_iter)):
        V = Z - eta * (D.T @ (D @ Z - X))
        Z = keep_top(np.abs(V), V)
    return Z, gap


def dense_to_codes(Z, k):
    """dense (K,N) -> (idx, val)[N,k], ascending atom index, -1 padded."""
    K, N = Z.shape
    idx = -np.ones((N, k), dtype=np.int64)
    val = np.zeros((N, k))
    for i in range(N):
        nz = np.flatnonzero(Z[:, i])
        idx[i, :len(nz)] = nz
        val[i, :len(nz)] = Z[nz, i]
    return idx, val


def omp_trace(X, D, n_nonzero_coefs=None, tol=None):
    """float64 replay of the oracle's `omp` (lyssa/sparse_coding.py:19-66) that also returns, per column, the
    smallest relative top-1/top-2 gap of |alpha| met at any step and the smallest relative distance of ||r|| from
    the stopping threshold at any test of the continue criterion (:27-34) — the decisions float32 can flip."""
    X = np.asarray(X, dtype=np.float64); D = np.asarray(D, dtype=np.float64)
    K, N = D.shape[1], X.shape[1]
    G = D.T @ D
    Z = np.zeros((K, N)); gap = np.full(N, np.inf); margin = np.full(N, np.inf)
    thr = 1e-10 if n_nonzero_coefs is not None else tol
    for i in range(N):
        x = X[:, i]; r = x.copy(); sel = []; z = None
        while True:
            rn = np.linalg.norm(r)
            margin[i] = min(margin[i], abs(rn - thr) / max(np.linalg.norm(x), 1e-300))
            if n_nonzero_coefs is not None:
                if not (len(sel) < n_nonzero_coefs and rn > thr):
                    break
            elif not rn >= thr:
                break
            a = np.abs(D.T @ r)
            p = int(np.argmax(a))
            top = a[p]; a[p] = -1.0
            gap[i] = min(gap[i], (top - a.max()) / top if top > 0 else 0.0)
            if p in sel:
                break
            sel.append(p)
            z = np.linalg.solve(G[np.ix_(sel, sel)], D[:, sel].T @ x)
            r = x - D[:, sel] @ z
        if sel:
            Z[sel, i] = z
    return Z, gap, margin
