// bomp_generic.cu — generic-shape Batch-OMP: one warp per signal, any K <= 4096, k <= 32 — and, in OMP mode,
// the reference's plain `omp` coder (lyssa/sparse_coding.py:19-66) on the same Gram/Alpha front end, k <= 64:
//   :27-34   continue while ||r|| > 1e-10 and i < n_nonzero_coefs, or, with only `tol`, while ||r|| >= tol
//            (||r||^2 = ||x||^2 - y.y with y = L^-1 alpha0[I]: the residual is orthogonal to the selected atoms)
//   :45-53   z = inv(G[I,I]) alpha0[I] with the TRUE Gram diagonal (no unit-norm assumption); here by the same
//            incremental Cholesky, pivot = G[p,p] - w.w; a non-positive pivot is the reference's LinAlgError (:48-51)
//   no pivot-epsilon stop (that is batch_omp's, :335/:345)
//
// Mirrors the per-signal loop of batch_omp, lyssa/sparse_coding.py:310-365, in float32:
//   :322     argmax |a|, FIRST maximum (lowest atom index) wins
//   :323-325 stop if the picked atom is already selected
//   :327     g = G[I, pick]
//   :330-349 new Cholesky row w = L^-1 g (forward substitution), literal 1 as the atom
//            self-product, stop if 1 - w.w < eps
//   :353-354 z = L^-T (L^-1 a0[I])   (y = L^-1 a0[I] grows by one entry per step because L is
//            lower-triangular and I is append-only — same numbers as re-solving from scratch)
//   :359     a = a0 - G[:, I] z       ("recompute form": s = sum_m G[I_m, :] z_m, a = a0 - s)
//   :365     Z[I, i] = z
// Alpha = X^T D comes from the fp32 GEMM in gemm.cu, chunk by chunk (never the full K x N).
//
// Layout in HBM: alpha chunk (C,K) signal-major in the workspace; G (K,K) row-major and
// symmetric, so "column" I_m is read as the contiguous row I_m (coalesced 128-byte lines);
// lane l of the warp owns atoms {e*32 + l}.  Codes are written as (idx,val)[N][k] and,
// optionally, as the dense row Z[i, 0..K) (signal-major fast path: 128-bit coalesced zero
// fill by the same warp, then the k scattered values).
#include "common.cuh"

namespace lys {
namespace {

constexpr int WARPS = 8;
template <int KMAXNZ> struct WarpScratchT {
    float L[KMAXNZ * KMAXNZ];
    float w[KMAXNZ];
    float y[KMAXNZ];
    float z[KMAXNZ];
    int   I[KMAXNZ];
};

// OMP: xnorm2 (C) = ||x_i||^2, tol/strict = the continue criterion, truncated += signals that reached k atoms with
// the criterion still true in tolerance-only mode (strict == 0)
template <int EPL, int KMAXNZ, bool OMP>
__global__ void __launch_bounds__(WARPS * 32)
bomp_warp_kernel(const float* __restrict__ alpha, const float* __restrict__ G,
                 int K, int64_t C, int k,
                 int32_t* __restrict__ idx, float* __restrict__ val, int32_t* __restrict__ nsel,
                 float* __restrict__ Z, int64_t z_atom_stride, int64_t z_sig_stride,
                 const float* __restrict__ xnorm2, float tol, int strict, int32_t* __restrict__ truncated)
{
    using WarpScratch = WarpScratchT<KMAXNZ>;
    extern __shared__ unsigned char smem_raw[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(smem_raw) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * WARPS;

    for (int64_t i = warp_global; i < C; i += n_warps) {
        float a0[EPL], a[EPL];
        const float* arow = alpha + i * (int64_t)K;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            int c = e * 32 + lane;
            a0[e] = (c < K) ? arow[c] : 0.f;
            a[e] = a0[e];
        }
        int cnt = 0;
        float rn2 = OMP ? xnorm2[i] : 0.f;
        bool more = false;                                // OMP: the continue criterion still holds
        for (int j = 0; j < k; ++j) {
            if (OMP) {                                    // sparse_coding.py:27-34
                const float rn = sqrtf(fmaxf(rn2, 0.f));
                more = strict ? (rn > tol) : (rn >= tol);
                if (!more) break;
            }
            // ---- :322 argmax |a|, lowest index on ties
            float best = -1.f; int bidx = 0x7fffffff;
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                int c = e * 32 + lane;
                float m = fabsf(a[e]);
                if (c < K && m > best) { best = m; bidx = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
            }
            // no maximum only if alpha holds NaNs (np.argmax would return the first NaN): stay in range, and make the
            // pick warp-uniform (NaN comparisons can leave the lanes of the butterfly with different winners, and the
            // breaks below must be taken by the whole warp)
            const int pick = __shfl_sync(0xffffffffu, (bidx < K) ? bidx : 0, 0);
            // ---- :323-325 already selected -> stop
            bool dup = false;
            for (int m = 0; m < cnt; ++m) dup |= (ws->I[m] == pick);
            if (dup) break;
            float a0p;   // a0[pick], broadcast from the owning lane
            {
                float mine = 0.f;
#pragma unroll
                for (int e = 0; e < EPL; ++e) if (e * 32 + lane == pick) mine = a0[e];
                a0p = __shfl_sync(0xffffffffu, mine, pick & 31);
            }
            const float gpp = OMP ? __ldg(G + (int64_t)pick * K + pick) : 1.f;      // batch_omp: the literal 1 of :334,:337,:344
            if (j == 0) {
                // :360-363  z = a0[pick]   (omp: z = a0[pick] / G[pick,pick])
                if (OMP && !(gpp > 0.f)) break;
                const float l0 = OMP ? sqrtf(gpp) : 1.f;
                if (lane == 0) { ws->I[0] = pick; ws->L[0] = l0; ws->y[0] = a0p / l0; ws->z[0] = a0p / (l0 * l0); }
                cnt = 1;
                if (OMP) rn2 = xnorm2[i] - (a0p / l0) * (a0p / l0);
            } else {
                // ---- :327 g = G[I, pick];  :330-334 / :342  w = L^-1 g
                float ww = 0.f;
                for (int r = 0; r < j; ++r) {
                    float s = G[(int64_t)ws->I[r] * K + pick];
                    for (int c = 0; c < r; ++c) s -= ws->L[r * KMAXNZ + c] * ws->w[c];
                    float wr = s / ws->L[r * KMAXNZ + r];
                    __syncwarp();
                    if (lane == 0) ws->w[r] = wr;
                    __syncwarp();
                    ww = fmaf(wr, wr, ww);
                }
                const float pivot = gpp - ww;
                if (OMP ? !(pivot > 0.f) : (pivot < kPivotEps)) break;          // :335 / :345 (omp: singular G[I,I], :48-51)
                const float diag = sqrtf(pivot);
                // y_j = (a0[pick] - L[j,:j] . y[:j]) / L[j,j]
                float s = a0p;
                for (int c = 0; c < j; ++c) s -= ws->w[c] * ws->y[c];
                const float yj = s / diag;
                __syncwarp();
                if (lane == 0) {
                    for (int c = 0; c < j; ++c) ws->L[j * KMAXNZ + c] = ws->w[c];
                    ws->L[j * KMAXNZ + j] = diag;
                    ws->y[j] = yj;
                    ws->I[j] = pick;
                }
                __syncwarp();
                cnt = j + 1;
                if (OMP) rn2 -= yj * yj;
                // ---- :354 z = L^-T y (back substitution)
                for (int r = cnt - 1; r >= 0; --r) {
                    float t = ws->y[r];
                    for (int c = r + 1; c < cnt; ++c) t -= ws->L[c * KMAXNZ + r] * ws->z[c];
                    float zr = t / ws->L[r * KMAXNZ + r];
                    __syncwarp();
                    if (lane == 0) ws->z[r] = zr;
                    __syncwarp();
                }
            }
            __syncwarp();
            // ---- :359 a = a0 - G[:, I] z   (not needed after the last selection)
            if (j + 1 < k) {
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    const int c = e * 32 + lane;
                    float s = 0.f;
                    if (c < K) {
                        for (int m = 0; m < cnt; ++m)
                            s = fmaf(__ldg(G + (int64_t)ws->I[m] * K + c), ws->z[m], s);
                    }
                    a[e] = a0[e] - s;
                }
            }
        }
        __syncwarp();
        // ---- outputs
        for (int m = lane; m < k; m += 32) {
            idx[i * k + m] = m < cnt ? ws->I[m] : -1;
            val[i * k + m] = m < cnt ? ws->z[m] : 0.f;
        }
        if (nsel && lane == 0) nsel[i] = cnt;
        if (OMP && truncated && !strict && cnt == k && lane == 0) {
            const float rn = sqrtf(fmaxf(rn2, 0.f));
            if (rn >= tol) atomicAdd(truncated, 1);
        }
        if (Z) {
            float* zrow = Z + i * z_sig_stride;
            if (z_atom_stride == 1) {
                for (int c = lane; c < K; c += 32) zrow[c] = 0.f;
            } else {
                for (int c = lane; c < K; c += 32) zrow[(int64_t)c * z_atom_stride] = 0.f;
            }
            __syncwarp();
            for (int m = lane; m < cnt; m += 32) zrow[(int64_t)ws->I[m] * z_atom_stride] = ws->z[m];
        }
        __syncwarp();
    }
}

template <int EPL, int KMAXNZ, bool OMP>
int launch_warp_kernel(const float* alpha, const float* G, int K, int64_t C, int k,
                       int32_t* idx, float* val, int32_t* nsel,
                       float* Z, int64_t zas, int64_t zss, const float* xnorm2, float tol, int strict, int32_t* truncated,
                       cudaStream_t stream)
{
    size_t smem = sizeof(WarpScratchT<KMAXNZ>) * WARPS;
    auto kern = bomp_warp_kernel<EPL, KMAXNZ, OMP>;
    // function attributes are per device and one process may drive several GPUs: set it on every launch
    LYS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (C + WARPS - 1) / WARPS;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, WARPS * 32, smem, stream>>>(alpha, G, K, C, k, idx, val, nsel, Z, zas, zss, xnorm2, tol, strict, truncated);
    LYS_LAUNCH_CHECK("bomp_warp_kernel");
    return LYS_OK;
}

template <int KMAXNZ, bool OMP>
int greedy_by_K(const float* alpha, const float* G, int K, int64_t C, int k, int32_t* idx, float* val, int32_t* nsel,
                float* Z, int64_t zas, int64_t zss, const float* xnorm2, float tol, int strict, int32_t* truncated, cudaStream_t stream)
{
#define LYS_GK(E) return launch_warp_kernel<E, KMAXNZ, OMP>(alpha, G, K, C, k, idx, val, nsel, Z, zas, zss, xnorm2, tol, strict, truncated, stream)
    if (K <= 32 * 4) LYS_GK(4);
    if (K <= 32 * 8) LYS_GK(8);
    if (K <= 32 * 16) LYS_GK(16);
    if (K <= 32 * 32) LYS_GK(32);
    if (K <= 32 * 64) LYS_GK(64);
    if (K <= 32 * 128) LYS_GK(128);
#undef LYS_GK
    set_error("bomp: K=%d exceeds LYS_MAX_ATOMS", K);
    return LYS_EUNSUPPORTED;
}

// ||x_i||^2 of a chunk of signals (one warp per signal)
__global__ void col_norm2_kernel(const float* __restrict__ X, int64_t xfs, int64_t xss, int n, int64_t C, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= C) return;
    float s = 0.f;
    for (int f = lane; f < n; f += 32) { const float v = X[(int64_t)f * xfs + w * xss]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) out[w] = s;
}

}  // namespace

// greedy phase on a chunk of C signals whose correlations are already in `alpha` (C,K)
int bomp_greedy_generic(const float* alpha, const float* G, int K, int64_t C, int k,
                        int32_t* idx, float* val, int32_t* nsel,
                        float* Z, int64_t zas, int64_t zss, cudaStream_t stream)
{
    return greedy_by_K<LYS_MAX_NONZERO, false>(alpha, G, K, C, k, idx, val, nsel, Z, zas, zss, nullptr, 0.f, 1, nullptr, stream);
}

// the same for the reference's `omp` coder (sparse_coding.py:19-66); k <= LYS_OMP_MAX_NONZERO
int omp_greedy_generic(const float* alpha, const float* G, int K, int64_t C, int k, const float* xnorm2, float tol, int strict,
                       int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zas, int64_t zss, int32_t* truncated,
                       cudaStream_t stream)
{
    return greedy_by_K<LYS_OMP_MAX_NONZERO, true>(alpha, G, K, C, k, idx, val, nsel, Z, zas, zss, xnorm2, tol, strict, truncated, stream);
}

int col_norm2(const float* X, int64_t xfs, int64_t xss, int n, int64_t C, float* out, cudaStream_t stream)
{
    if (C == 0) return LYS_OK;
    col_norm2_kernel<<<(unsigned)((C * 32 + 255) / 256), 256, 0, stream>>>(X, xfs, xss, n, C, out);
    LYS_LAUNCH_CHECK("col_norm2_kernel");
    return LYS_OK;
}

}  // namespace lys
