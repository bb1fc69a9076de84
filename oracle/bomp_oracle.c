/* TEST INFRASTRUCTURE ONLY — float64 C restatement of the per-signal loop of
 * lyssa/sparse_coding.py:302-367 (batch_omp), used by tests/ for full-size parity checks
 * and by bench.py's cpu_baseline leg.  Never linked into the product library.
 *
 * Validated against the NumPy restatement (oracle/lyssa_oracle.py), which is itself pinned
 * bit-for-bit against the live reference (oracle/ref_loader.py), in
 * tests/test_oracle.py::test_c_oracle_matches_numpy_oracle.
 *
 * Inputs are the float64 Alpha = D^T X and Gram = D^T D exactly as the reference forms them
 * (sparse_coding.py:630-631, NumPy dgemm); this file restates only what runs per signal:
 *   :322     first-maximum argmax of |a|
 *   :323-325 stop if already selected
 *   :330-349 Cholesky row by forward substitution, literal 1 on the diagonal, stop if 1-w.w < eps
 *   :353-354 z = L^-T (L^-1 a0[I])
 *   :359     a = a0 - G[:, I] z
 *   :365     Z[I, i] = z
 * Layout: alpha is (K, N) C-order (signals in columns) when alpha_signal_major == 0, else
 * (N, K); gram is (K, K).  Outputs: idx (N, k) int32 padded with -1, val (N, k) float64,
 * nsel (N).  Single-threaded; the Python wrapper (oracle/c_oracle.py) fans column ranges out
 * over threads (ctypes releases the GIL).  Optional trace: gap (N, k) relative top1-top2 gap at every executed argmax,
 * vs (N, k) pivots.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KMAX 64

int lys_oracle_batch_omp(const double* alpha, int alpha_signal_major, const double* gram,
                         int64_t n_atoms, int64_t n_signals, int k,
                         int32_t* idx, double* val, int32_t* nsel, double* gap, double* vs_out)
{
    if (k < 1 || k > KMAX) return -1;
    int status = 0;
    {
        double* a0 = (double*)malloc(sizeof(double) * (size_t)n_atoms);
        double* a = (double*)malloc(sizeof(double) * (size_t)n_atoms);
        double L[KMAX][KMAX];
        double w[KMAX], y[KMAX], z[KMAX], zprev[KMAX];
        int32_t sup[KMAX];
        for (int64_t i = 0; i < n_signals; ++i) {
            if (alpha_signal_major) {
                memcpy(a0, alpha + i * n_atoms, sizeof(double) * (size_t)n_atoms);
            } else {
                for (int64_t c = 0; c < n_atoms; ++c) a0[c] = alpha[c * n_signals + i];
            }
            memcpy(a, a0, sizeof(double) * (size_t)n_atoms);
            int cnt = 0;
            for (int j = 0; j < k; ++j) {
                if (gap) gap[i * k + j] = INFINITY;
                if (vs_out) vs_out[i * k + j] = 1.0;
            }
            for (int j = 0; j < k; ++j) {
                /* :322 first maximum */
                int64_t pick = 0; double top = fabs(a[0]), second = -1.0;
                for (int64_t c = 1; c < n_atoms; ++c) {
                    double m = fabs(a[c]);
                    if (m > top) { second = top; top = m; pick = c; }
                    else if (m > second) second = m;
                }
                if (gap) gap[i * k + j] = top > 0 ? (top - second) / top : 0.0;
                int dup = 0;
                for (int m = 0; m < cnt; ++m) if (sup[m] == pick) dup = 1;
                if (dup) break;                                       /* :323-325 */
                if (j == 0) {                                         /* :360-363 */
                    sup[0] = (int32_t)pick; cnt = 1;
                    z[0] = a0[pick]; y[0] = z[0]; L[0][0] = 1.0;
                } else {
                    /* w = L[:j,:j]^-1 G[I, pick]   :330-334 / :342 */
                    double ww = 0.0;
                    for (int r = 0; r < j; ++r) {
                        double s = gram[(int64_t)sup[r] * n_atoms + pick];
                        for (int c = 0; c < r; ++c) s -= L[r][c] * w[c];
                        w[r] = s / L[r][r];
                        ww += w[r] * w[r];
                    }
                    double pivot = 1.0 - ww;
                    if (vs_out) vs_out[i * k + j] = pivot;
                    if (pivot < DBL_EPSILON) break;                   /* :335 / :345 */
                    for (int c = 0; c < j; ++c) L[j][c] = w[c];
                    L[j][j] = sqrt(pivot);
                    sup[j] = (int32_t)pick; cnt = j + 1;
                    /* y = L^-1 a0[I]  (:353) */
                    for (int r = 0; r < cnt; ++r) {
                        double s = a0[sup[r]];
                        for (int c = 0; c < r; ++c) s -= L[r][c] * y[c];
                        y[r] = s / L[r][r];
                    }
                    /* z = L^-T y  (:354) */
                    for (int r = cnt - 1; r >= 0; --r) {
                        double s = y[r];
                        for (int c = r + 1; c < cnt; ++c) s -= L[c][r] * z[c];
                        z[r] = s / L[r][r];
                    }
                }
                /* a = a0 - G[:, I] z  (:359); skipped after the last selection (unused) */
                if (j + 1 < k) {
                    for (int64_t c = 0; c < n_atoms; ++c) {
                        double s = 0.0;
                        for (int m = 0; m < cnt; ++m) s += gram[(int64_t)sup[m] * n_atoms + c] * z[m];
                        a[c] = a0[c] - s;
                    }
                }
                (void)zprev;
            }
            nsel[i] = cnt;
            for (int m = 0; m < k; ++m) {
                idx[i * k + m] = m < cnt ? sup[m] : -1;
                val[i * k + m] = m < cnt ? z[m] : 0.0;
            }
        }
        free(a0); free(a);
    }
    return status;
}

/* float64 restatement of the atom loop of approx_ksvd (lyssa/dict_learning/ksvd.py:105-124) on sparse codes, for
 * sweeps too large for the dense NumPy restatement (oracle/lyssa_oracle.py::approx_ksvd, against which this function
 * is validated in tests/test_oracle.py::test_c_sweep_matches_numpy_oracle).
 *   R   (N, n) signal-major: the residual Y - D X of :103, updated in place (:123)
 *   D   (n, K) C-order as in the reference, updated in place (:118-119)
 *   idx/val (N, k): the non-zeros of the reference's dense X; users of atom c = entries with idx == c and
 *       val != 0 (:111), visited in ascending signal order; val updated in place (:121)
 *   unused (K): set to 1 for atoms without users (:112-115)
 * Per atom, exactly the reference's sequence: Rk = R[:,users] + d x (:116); d = Rk x (:118);
 * d /= ||d|| + eps (:119, utils/math.py:61-62); x = Rk^T d (:121); R[:,users] = Rk - d x (:123). */
int lys_oracle_approx_ksvd_sweep(double* R, double* D, const int32_t* idx, double* val,
                                 int64_t n_signals, int n, int64_t n_atoms, int k, int n_cycles, int32_t* unused)
{
    const int64_t E = n_signals * (int64_t)k;
    int64_t* rowptr = (int64_t*)calloc((size_t)n_atoms + 1, sizeof(int64_t));
    int64_t* ent = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E > 0 ? E : 1));
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_atoms);
    double* Rk = NULL;
    double* dnew = (double*)malloc(sizeof(double) * (size_t)n);
    if (!rowptr || !ent || !fill || !dnew) return -1;
    for (int64_t e = 0; e < E; ++e)
        if (idx[e] >= 0 && idx[e] < n_atoms && val[e] != 0.0) rowptr[idx[e] + 1]++;
    int64_t max_users = 0;
    for (int64_t c = 0; c < n_atoms; ++c) {
        if (rowptr[c + 1] > max_users) max_users = rowptr[c + 1];
        rowptr[c + 1] += rowptr[c];
        fill[c] = rowptr[c];
    }
    for (int64_t e = 0; e < E; ++e)
        if (idx[e] >= 0 && idx[e] < n_atoms && val[e] != 0.0) ent[fill[idx[e]]++] = e;
    Rk = (double*)malloc(sizeof(double) * (size_t)(max_users > 0 ? max_users : 1) * (size_t)n);
    if (!Rk) return -1;
    memset(unused, 0, sizeof(int32_t) * (size_t)n_atoms);
    for (int cyc = 0; cyc < n_cycles; ++cyc) {
        for (int64_t c = 0; c < n_atoms; ++c) {
            const int64_t lo = rowptr[c], hi = rowptr[c + 1], m = hi - lo;
            if (m == 0) { unused[c] = 1; continue; }                               /* :112-115 */
            for (int f = 0; f < n; ++f) dnew[f] = 0.0;
            for (int64_t u = 0; u < m; ++u) {                                       /* :116, :118 */
                const int64_t e = ent[lo + u], i = e / k;
                const double x = val[e];
                for (int f = 0; f < n; ++f) {
                    const double v = R[i * n + f] + D[(int64_t)f * n_atoms + c] * x;
                    Rk[u * n + f] = v;
                    dnew[f] += v * x;
                }
            }
            double nrm = 0.0;
            for (int f = 0; f < n; ++f) nrm += dnew[f] * dnew[f];
            nrm = sqrt(nrm) + DBL_EPSILON;                                          /* :119 */
            for (int f = 0; f < n; ++f) { dnew[f] /= nrm; D[(int64_t)f * n_atoms + c] = dnew[f]; }
            for (int64_t u = 0; u < m; ++u) {                                       /* :121, :123 */
                const int64_t e = ent[lo + u], i = e / k;
                double x = 0.0;
                for (int f = 0; f < n; ++f) x += Rk[u * n + f] * dnew[f];
                val[e] = x;
                for (int f = 0; f < n; ++f) R[i * n + f] = Rk[u * n + f] - dnew[f] * x;
            }
        }
    }
    free(rowptr); free(ent); free(fill); free(Rk); free(dnew);
    return 0;
}
