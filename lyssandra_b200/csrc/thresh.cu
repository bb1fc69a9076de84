// thresh.cu — the reference's two thresholding coders on the correlation front end:
//   'thresh' (lyssa/sparse_coding.py:416-425, dispatch :636-641): Alpha = D^T X, keep the
//            n_nonzero_coefs LARGEST SIGNED correlations of every signal, Z = Alpha there;
//   'iht'    (:433-446, dispatch :671-690): Z0 = thresh(Alpha), then n_iter times
//            Z <- Z - eta * D^T (D Z - X), keep the k largest |Z| of every signal.
// 'thresh' at the fused kernel's shapes (n <= 64, K in {256,512,768,1024}, k <= 10) runs inside it
// (bomp_fused.cu, MODE 1: correlations never leave TMEM).  Everything else here is a correlation
// GEMM (corr_gemm_tc.cu / gemm.cu) followed by a per-signal selection kernel, one warp per signal.
#include "common.cuh"
#include "../../include/lyssa_b200.h"

#include <algorithm>
#include <math.h>
#include <stdlib.h>

namespace lys {

bool corr_gemm_tc_supported(int n, int K);
size_t corr_gemm_tc_planes_bytes(int n, int K);
int corr_gemm_tc_prepare(const float* D, int64_t ldd, int n, int K, void* planes, cudaStream_t stream);
int corr_gemm_tc(const float* X, int64_t xfs, int64_t xss, const void* planes,
                 int n, int K, int64_t C, float* alpha, cudaStream_t stream);

// 'thresh' inside the fused tcgen05 kernel (bomp_fused.cu, MODE 1); LYS_EUNSUPPORTED for shapes it is not built for
int thresh_encode_fused(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd,
                        int n, int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel,
                        float* Z, int64_t zas, int64_t zss, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t thresh_fused_workspace_bytes(int n, int K, int64_t N, int k);

namespace {

constexpr int SEL_WARPS = 8;
constexpr int64_t kSelChunkBytes = 256ll << 20;

int64_t sel_chunk(int K, int64_t N)
{
    int64_t c = kSelChunkBytes / ((int64_t)K * 4);
    c = std::max<int64_t>(1024, c / 1024 * 1024);
    return std::min<int64_t>(c, std::max<int64_t>(N, 1));
}

// v[c] = scale * alpha[i][c] (+ the signal's previous sparse code, when given); the k entries
// with the largest key (v, or |v| when ABS) are selected in descending key order, ties to the
// lower atom index.  Selection r+1 is the largest (key, -index) strictly below selection r in
// that lexicographic order, so no per-entry mask is kept.
template <bool ABS>
__global__ void __launch_bounds__(SEL_WARPS * 32)
select_topk_kernel(const float* __restrict__ alpha, float scale,
                   const int32_t* prev_idx, const float* prev_val, int prev_k,
                   int K, int64_t C, int k,
                   int32_t* idx, float* val, int32_t* nsel,
                   float* __restrict__ Z, int64_t zas, int64_t zss)
{
    extern __shared__ float srow_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    float* srow = srow_all + (size_t)warp * K;
    for (int64_t i = (int64_t)blockIdx.x * wpc + warp; i < C; i += (int64_t)gridDim.x * wpc) {
        const float* a = alpha + i * K;
        for (int c = lane; c < K; c += 32) srow[c] = scale * __ldcs(a + c);
        __syncwarp();
        if (prev_idx) {
            for (int j = lane; j < prev_k; j += 32) {
                const int p = prev_idx[i * prev_k + j];
                if (p >= 0) srow[p] += prev_val[i * prev_k + j];     // indices of one signal are distinct
            }
            __syncwarp();
        }
        if (Z) {
            float* z = Z + i * zss;
            if (zas == 1 && (K % 4) == 0 && ((reinterpret_cast<uintptr_t>(z) & 15) == 0)) {
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = lane * 4; c < K; c += 128) *reinterpret_cast<float4*>(z + c) = zero;
            } else {
                for (int c = lane; c < K; c += 32) z[(int64_t)c * zas] = 0.f;
            }
            __syncwarp();
        }
        float last_key = INFINITY;
        int last_idx = -1;
        for (int r = 0; r < k; ++r) {
            float best = -INFINITY;
            int bidx = 0x7fffffff;
            for (int c = lane; c < K; c += 32) {
                const float v = srow[c];
                const float key = ABS ? fabsf(v) : v;
                const bool below = (key < last_key) || (key == last_key && c > last_idx);
                if (below && key > best) { best = key; bidx = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
            }
            last_key = best;
            last_idx = bidx;
            if (lane == 0) {
                const bool ok = bidx < K;                             // false only for non-finite rows
                const float v = ok ? srow[bidx] : 0.f;
                idx[i * k + r] = ok ? bidx : -1;
                val[i * k + r] = v;
                if (Z && ok) Z[i * zss + (int64_t)bidx * zas] = v;
            }
        }
        if (nsel && lane == 0) nsel[i] = k;
        __syncwarp();
    }
}

// Register variant for K <= 32*IPL <= 1024 and k <= 32 (the shapes the coders are used at).  A lane holds the
// keys of columns j*32+lane.  Nothing in it is a k-deep chain of warp reductions:
//  (1) T = the k-th largest of the 32 lane maxima, found by ranking them against each other with 31 butterfly
//      shuffles — at least k entries of the row are >= T, and every one of its k largest is;
//  (2) entries >= T (a handful per row) are appended to a per-warp candidate list in shared memory;
//  (3) with at most 32 candidates, lane l ranks candidate l among all of them in the order (key descending,
//      column ascending) with another 31 butterfly shuffles and, if its rank is below k, writes output slot
//      `rank` itself.  Rows with more candidates (heavy ties) take k sequential rounds over the masks instead.
// Same result as select_topk_kernel: descending key, ties to the lower column.
template <int IPL, bool ABS, bool PREV>
__global__ void __launch_bounds__(SEL_WARPS * 32, IPL == 32 ? 3 : 6)
select_topk_reg_kernel(const float* __restrict__ alpha, float scale,
                       const int32_t* prev_idx, const float* prev_val, int prev_k,
                       int K, int64_t C, int k,
                       int32_t* idx, float* val, int32_t* nsel,
                       float* __restrict__ Z, int64_t zas, int64_t zss)
{
    extern __shared__ float srow_all[];
    __shared__ int cand_cnt[SEL_WARPS];
    __shared__ int cand_col[SEL_WARPS][32];
    constexpr int NONE = 0x7fffffff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    float* srow = srow_all + (size_t)warp * (IPL * 32);
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < IPL; ++j) valid |= (j * 32 + lane < K) ? (1u << j) : 0u;

    for (int64_t i = (int64_t)blockIdx.x * wpc + warp; i < C; i += (int64_t)gridDim.x * wpc) {
        const float* a = alpha + i * K;
        float key[IPL];
#pragma unroll
        for (int j = 0; j < IPL; ++j) key[j] = ((valid >> j) & 1u) ? scale * __ldcs(a + j * 32 + lane) : 0.f;
        if (lane == 0) cand_cnt[warp] = 0;
#pragma unroll
        for (int j = 0; j < IPL; ++j) srow[j * 32 + lane] = key[j];
        __syncwarp();
        if (PREV) {
            // the previous code is added by column through shared memory
            for (int t = lane; t < prev_k; t += 32) {
                const int p = prev_idx[i * prev_k + t];
                if (p >= 0) srow[p] += prev_val[i * prev_k + t];
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < IPL; ++j) key[j] = srow[j * 32 + lane];
        }
#pragma unroll
        for (int j = 0; j < IPL; ++j) key[j] = ((valid >> j) & 1u) ? (ABS ? fabsf(key[j]) : key[j]) : -INFINITY;
        if (Z) {
            float* z = Z + i * zss;
            if (zas == 1 && (K % 4) == 0 && ((reinterpret_cast<uintptr_t>(z) & 15) == 0)) {
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = lane * 4; c < K; c += 128) *reinterpret_cast<float4*>(z + c) = zero;
            } else {
                for (int c = lane; c < K; c += 32) z[(int64_t)c * zas] = 0.f;
            }
            __syncwarp();
        }
        // lane maximum (pairwise tree) and its rank among the 32 lane maxima
        float t2[IPL];
#pragma unroll
        for (int j = 0; j < IPL; ++j) t2[j] = key[j];
#pragma unroll
        for (int w = IPL / 2; w > 0; w >>= 1)
#pragma unroll
            for (int j = 0; j < w; ++j) t2[j] = fmaxf(t2[j], t2[j + w]);
        const float m = t2[0];
        int rank = 0;
#pragma unroll
        for (int o = 1; o < 32; ++o) {
            // partner lane^o is below this lane iff the top set bit of o is set in lane
            const int hb = o >= 16 ? 16 : o >= 8 ? 8 : o >= 4 ? 4 : o >= 2 ? 2 : 1;
            const float om = __shfl_xor_sync(0xffffffffu, m, o);
            rank += (om > m || (om == m && (lane & hb))) ? 1 : 0;
        }
        const unsigned kth = __ballot_sync(0xffffffffu, rank == k - 1);
        const float T = __shfl_sync(0xffffffffu, m, kth ? __ffs(kth) - 1 : 0);
        uint32_t cm = 0;
#pragma unroll
        for (int j = 0; j < IPL; ++j) cm |= (key[j] >= T) ? (1u << j) : 0u;
        cm &= valid;
        for (uint32_t mk = cm; mk; mk &= mk - 1) {
            const int slot = atomicAdd(&cand_cnt[warp], 1);
            if (slot < 32) cand_col[warp][slot] = (__ffs(mk) - 1) * 32 + lane;
        }
        __syncwarp();
        const int n = cand_cnt[warp];
        if (n <= 32) {
            const int c = lane < n ? cand_col[warp][lane] : NONE;
            const float v = lane < n ? srow[c] : 0.f;
            const float kk = lane < n ? (ABS ? fabsf(v) : v) : -INFINITY;
            int rk = 0;
#pragma unroll
            for (int o = 1; o < 32; ++o) {
                const float ok = __shfl_xor_sync(0xffffffffu, kk, o);
                const int oc = __shfl_xor_sync(0xffffffffu, c, o);
                rk += (oc != NONE && (ok > kk || (ok == kk && oc < c))) ? 1 : 0;
            }
            if (lane < n && rk < k) {
                idx[i * k + rk] = c;
                val[i * k + rk] = v;
                if (Z) Z[i * zss + (int64_t)c * zas] = v;
            }
            for (int r = n + lane; r < k; r += 32) {                 // only for rows without k finite entries
                idx[i * k + r] = -1;
                val[i * k + r] = 0.f;
            }
        } else {
            float lb;
            int lj;
            auto rescan = [&]() {
                lb = -INFINITY;
                lj = -1;
                for (uint32_t mk = cm; mk; mk &= mk - 1) {
                    const int j = __ffs(mk) - 1;
                    const float v = srow[j * 32 + lane];
                    const float kk = ABS ? fabsf(v) : v;
                    if (kk > lb || lj < 0) { lb = kk; lj = j; }
                }
            };
            rescan();
            for (int r = 0; r < k; ++r) {
                float best = lb;
                int bc = lj >= 0 ? lj * 32 + lane : NONE;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                    if (oc != NONE && (bc == NONE || ob > best || (ob == best && oc < bc))) { best = ob; bc = oc; }
                }
                if (lane == 0) {
                    const bool ok = bc < K;
                    const float v = ok ? srow[bc] : 0.f;
                    idx[i * k + r] = ok ? bc : -1;
                    val[i * k + r] = v;
                    if (Z && ok) Z[i * zss + (int64_t)bc * zas] = v;
                }
                if (lj >= 0 && bc == lj * 32 + lane) {
                    cm &= ~(1u << lj);
                    rescan();
                }
            }
        }
        if (nsel && lane == 0) nsel[i] = k;
        __syncwarp();
    }
}

template <int IPL>
int launch_select_reg(bool use_abs, const float* alpha, float scale, const int32_t* pidx, const float* pval, int pk,
                      int K, int64_t C, int k, int32_t* idx, float* val, int32_t* nsel,
                      float* Z, int64_t zas, int64_t zss, cudaStream_t stream)
{
    const size_t smem = (size_t)SEL_WARPS * IPL * 32 * sizeof(float);
    const int64_t blocks = std::min<int64_t>((C + SEL_WARPS - 1) / SEL_WARPS, (int64_t)sm_count() * (IPL == 32 ? 3 : 6));
#define LYS_SEL_LAUNCH(A, P)                                                                                           \
    select_topk_reg_kernel<IPL, A, P><<<(unsigned)blocks, SEL_WARPS * 32, smem, stream>>>(alpha, scale, pidx, pval, pk, K, C, k, \
                                                                                          idx, val, nsel, Z, zas, zss)
    if (use_abs && pidx) LYS_SEL_LAUNCH(true, true);
    else if (use_abs) LYS_SEL_LAUNCH(true, false);
    else if (pidx) LYS_SEL_LAUNCH(false, true);
    else LYS_SEL_LAUNCH(false, false);
#undef LYS_SEL_LAUNCH
    LYS_LAUNCH_CHECK("select_topk_reg_kernel");
    return LYS_OK;
}

int launch_select(bool use_abs, const float* alpha, float scale, const int32_t* pidx, const float* pval, int pk,
                  int K, int64_t C, int k, int32_t* idx, float* val, int32_t* nsel,
                  float* Z, int64_t zas, int64_t zss, cudaStream_t stream)
{
    if (k <= 32 && K <= 1024) {
        if (K <= 256) return launch_select_reg<8>(use_abs, alpha, scale, pidx, pval, pk, K, C, k, idx, val, nsel, Z, zas, zss, stream);
        if (K <= 512) return launch_select_reg<16>(use_abs, alpha, scale, pidx, pval, pk, K, C, k, idx, val, nsel, Z, zas, zss, stream);
        return launch_select_reg<32>(use_abs, alpha, scale, pidx, pval, pk, K, C, k, idx, val, nsel, Z, zas, zss, stream);
    }
    const size_t smem = (size_t)SEL_WARPS * K * sizeof(float);
    // function attributes are per device and one process may drive several GPUs: set it on every launch
    if (use_abs)
        LYS_CUDA(cudaFuncSetAttribute(select_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        LYS_CUDA(cudaFuncSetAttribute(select_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200u << 10) / std::max<size_t>(smem, 1)));
    const int64_t blocks = std::min<int64_t>((C + SEL_WARPS - 1) / SEL_WARPS, (int64_t)sm_count() * per_sm);
    if (use_abs)
        select_topk_kernel<true><<<(unsigned)blocks, SEL_WARPS * 32, smem, stream>>>(alpha, scale, pidx, pval, pk, K, C, k,
                                                                                      idx, val, nsel, Z, zas, zss);
    else
        select_topk_kernel<false><<<(unsigned)blocks, SEL_WARPS * 32, smem, stream>>>(alpha, scale, pidx, pval, pk, K, C, k,
                                                                                       idx, val, nsel, Z, zas, zss);
    LYS_LAUNCH_CHECK("select_topk_kernel");
    return LYS_OK;
}

struct ThreshLayout {
    int64_t chunk;
    size_t alpha_off, planes_off, r_off, resid_ws_off, resid_ws_bytes, total;
};

ThreshLayout thresh_layout(int n, int K, int64_t N, bool iht)
{
    ThreshLayout L;
    L.chunk = sel_chunk(K, N);
    size_t off = 0;
    L.alpha_off = off;  off += align_up((size_t)L.chunk * K * sizeof(float), 256);
    L.planes_off = off; off += align_up(corr_gemm_tc_planes_bytes(n, K), 256);
    L.r_off = off;
    L.resid_ws_off = off;
    L.resid_ws_bytes = 0;
    if (iht) {
        off += align_up((size_t)L.chunk * n * sizeof(float), 256);
        L.resid_ws_off = off;
        L.resid_ws_bytes = lys_residual_workspace_bytes(n, K, L.chunk);
        off += align_up(L.resid_ws_bytes, 256);
    }
    L.total = std::max(off + 256, thresh_fused_workspace_bytes(n, K, N, 1));
    return L;
}

int thresh_common(bool iht, const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd,
                  int n, int K, int64_t N, int k, float eta, int n_iter,
                  int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zas, int64_t zss,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream)
{
    const char* who = iht ? "lys_iht_encode" : "lys_thresh_encode";
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES, "%s: n=%d out of range [1,%d]", who, n, LYS_MAX_FEATURES);
    LYS_CHECK_ARG(K >= 1 && K <= LYS_MAX_ATOMS, "%s: K=%d out of range [1,%d]", who, K, LYS_MAX_ATOMS);
    LYS_CHECK_ARG(k >= 1 && k <= K, "%s: n_nonzero_coefs=%d must be in [1, K=%d]", who, k, K);
    LYS_CHECK_ARG(!iht || k <= LYS_MAX_NONZERO, "%s: n_nonzero_coefs=%d > %d", who, k, LYS_MAX_NONZERO);
    LYS_CHECK_ARG(!iht || n_iter >= 0, "%s: n_iter < 0", who);
    LYS_CHECK_ARG(N >= 0, "%s: N < 0", who);
    if (N == 0) return LYS_OK;
    LYS_CHECK_ARG(X && D && idx && val, "%s: null pointer", who);
    LYS_CHECK_ARG(ldd >= K, "%s: ldd < K", who);
    LYS_CHECK_ARG(!Z || (zas >= 1 && zss >= 1), "%s: bad Z strides", who);
    const ThreshLayout L = thresh_layout(n, K, N, iht);
    if (!workspace || workspace_bytes < L.total) {
        set_error("%s: workspace %zu B < required %zu B", who, workspace_bytes, L.total);
        return LYS_EWORKSPACE;
    }
    // correlation, selection (and dense rows) in one kernel when the shape allows it; for 'iht' this is the start Z0
    bool have_start = false;
    {
        const bool final_codes = !iht || n_iter == 0;
        const int frc = thresh_encode_fused(X, xfs, xss, D, ldd, n, K, N, k, idx, val, nsel, final_codes ? Z : nullptr, zas, zss,
                                            workspace, workspace_bytes, stream);
        if (frc != LYS_EUNSUPPORTED) {
            if (frc != LYS_OK || final_codes) return frc;
            have_start = true;
        }
    }
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    float* alpha = reinterpret_cast<float*>(ws + L.alpha_off);
    void* planes = ws + L.planes_off;
    float* R = reinterpret_cast<float*>(ws + L.r_off);
    const bool use_tc = corr_gemm_tc_supported(n, K);
    int rc = LYS_OK;
    if (use_tc && (rc = corr_gemm_tc_prepare(D, ldd, n, K, planes, stream))) return rc;

    auto corr = [&](const float* S, int64_t sfs, int64_t sss, int64_t C) -> int {
        if (use_tc) return corr_gemm_tc(S, sfs, sss, planes, n, K, C, alpha, stream);
        return sgemm_strided(S, sss, sfs, D, ldd, 1, alpha, K, 1, C, K, n, stream);
    };

    for (int64_t s0 = 0; s0 < N; s0 += L.chunk) {
        const int64_t C = std::min(L.chunk, N - s0);
        const float* Xc = X + s0 * xss;
        int32_t* ci = idx + s0 * k;
        float* cv = val + s0 * k;
        int32_t* cn = nsel ? nsel + s0 : nullptr;
        float* Zc = Z ? Z + s0 * zss : nullptr;
        const bool last0 = !iht || n_iter == 0;
        if (!have_start) {
            if ((rc = corr(Xc, xfs, xss, C))) return rc;
            if ((rc = launch_select(false, alpha, 1.f, nullptr, nullptr, 0, K, C, k, ci, cv, cn,
                                    last0 ? Zc : nullptr, zas, zss, stream))) return rc;
        }
        for (int it = 0; iht && it < n_iter; ++it) {
            // R = X - D Z (the reference keeps D Z - X and subtracts eta * D^T R; same update)
            if ((rc = lys_residual(Xc, xfs, xss, D, ldd, ci, cv, n, K, C, k, R, nullptr,
                                   ws + L.resid_ws_off, L.resid_ws_bytes, stream))) return rc;
            if ((rc = corr(R, 1, n, C))) return rc;
            if ((rc = launch_select(true, alpha, eta, ci, cv, k, K, C, k, ci, cv, cn,
                                    it + 1 == n_iter ? Zc : nullptr, zas, zss, stream))) return rc;
        }
    }
    return LYS_OK;
}

}  // namespace
}  // namespace lys

using namespace lys;

extern "C" size_t lys_thresh_workspace_bytes(int n, int K, int64_t N)
{
    if (n < 1 || K < 1 || N < 0) return 0;
    return thresh_layout(n, K, N, false).total;
}

extern "C" int lys_thresh_encode(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd,
                                 int n, int K, int64_t N, int k,
                                 int32_t* idx, float* val, int32_t* nsel,
                                 float* Z, int64_t zas, int64_t zss,
                                 void* workspace, size_t workspace_bytes, void* stream)
{
    return thresh_common(false, X, xfs, xss, D, ldd, n, K, N, k, 0.f, 0, idx, val, nsel, Z, zas, zss,
                         workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int lys_topk_select(const float* alpha, int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel,
                               float* Z, int64_t zas, int64_t zss, void* stream)
{
    LYS_CHECK_ARG(K >= 1 && K <= LYS_MAX_ATOMS && k >= 1 && k <= K && N >= 0, "lys_topk_select: bad shape");
    if (N == 0) return LYS_OK;
    LYS_CHECK_ARG(alpha && idx && val, "lys_topk_select: null pointer");
    LYS_CHECK_ARG(!Z || (zas >= 1 && zss >= 1), "lys_topk_select: bad Z strides");
    return launch_select(false, alpha, 1.f, nullptr, nullptr, 0, K, N, k, idx, val, nsel, Z, zas, zss, (cudaStream_t)stream);
}

extern "C" size_t lys_iht_workspace_bytes(int n, int K, int64_t N)
{
    if (n < 1 || K < 1 || N < 0) return 0;
    return thresh_layout(n, K, N, true).total;
}

extern "C" int lys_iht_encode(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd,
                              int n, int K, int64_t N, int k, float eta, int n_iter,
                              int32_t* idx, float* val, int32_t* nsel,
                              float* Z, int64_t zas, int64_t zss,
                              void* workspace, size_t workspace_bytes, void* stream)
{
    return thresh_common(true, X, xfs, xss, D, ldd, n, K, N, k, eta, n_iter, idx, val, nsel, Z, zas, zss,
                         workspace, workspace_bytes, (cudaStream_t)stream);
}
