// ksvd_exact.cu — the atom loop of EXACT K-SVD (lyssa/dict_learning/ksvd.py:19-43) on sparse codes, n <= 64.
//
// Per atom c (sequential, as in the reference):
//     users = columns with Z[c,:] != 0                                :31    (CSR built by lys_build_atom_csr)
//     R_k   = R[:,users] + d x^T                                      :36    (never materialised in HBM)
//     (d, x) <- top singular triplet of R_k: d = u_1, x = s_1 v_1     :37-39 (the reference: scikit-learn's
//              randomized_svd(R_k, 1, n_iter=10, flip_sign=False), a third-party dependency; sign arbitrary)
//     R[:,users] = R_k - d x^T                                        :41
// The triplet is taken from the n x n matrix M = R_k R_k^T = sum over users of (r_i + d x_i)(r_i + d x_i)^T:
// ONE device-wide reduction per atom (n(n+1)/2 + 1 fixed-point words, the same integer all-reduce through L2 atomics
// as the approximate sweep, ksvd_sweep.cu — bitwise reproducible), after which every CTA solves the same small
// symmetric eigenproblem redundantly: 12 normalised squarings of M (M^4096: a relative gap of 0.4 % between the two
// largest singular values is resolved to 1e-6) and two power steps with M itself; then d = u_1 (sign chosen so that
// d.d_old >= 0) and, locally, x_i = (r_i + d_old x_i).d, r_i <- (r_i + d_old x_i) - d x_i.
// A power iteration over the users' rows themselves would need one reduction per iteration; the Gram form needs one.
//
// Ownership as in the approximate sweep: CTA b owns signals [bS, (b+1)S); the users of an atom are sorted by signal, so
// its users are one sub-range of the CSR segment (`bounds`).  Up to EX_CAP users per CTA and atom keep their rows
// R_k[:, i] in shared memory between the two phases; more are streamed again.
#include "sweep_common.cuh"
#include <algorithm>
#include <cmath>

namespace lys {
int transpose(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, cudaStream_t st);

namespace {

constexpr int EX_THREADS = 512;
constexpr int EX_WARPS = EX_THREADS / 32;
constexpr int EX_NF = 64;                     // feature extent (n <= 64, zero padded)
constexpr int EX_CAP = 256;                   // users per CTA and atom whose R_k rows stay in shared memory
constexpr int EX_LDR = EX_NF + 1;             // row stride of the staged rows (bank-conflict padding)
constexpr int EX_NE = (EX_NF * (EX_NF + 1) / 2 + EX_THREADS - 1) / EX_THREADS;      // symmetric entries per thread (5)
constexpr int EX_SQUARINGS = 12;

struct ExactArgs {
    float* R; const float* Dt; float* Dt_new; float* val;
    const int32_t* entries; const int32_t* bounds;
    int n, K, k, lda;
    int32_t* unused;
    u64* acc;                 // [K][lda] words, zeroed before the launch; word lda-1 of an atom = its user count
    const double* fx;
    FastDiv kdiv;
};

__device__ __forceinline__ void block_sync() { __syncthreads(); }

__global__ void __launch_bounds__(EX_THREADS, 1)
ksvd_exact_kernel(const ExactArgs a)
{
    extern __shared__ float sm[];
    float* rk = sm;                                   // [EX_CAP][EX_LDR]   rows of R_k owned by this CTA
    float* M0 = rk + EX_CAP * EX_LDR;                 // [64][64]           M = R_k R_k^T (all CTAs hold the same)
    float* S0 = M0 + EX_NF * EX_NF;                   // [64][64]           squaring ping
    float* S1 = S0 + EX_NF * EX_NF;                   // [64][64]           squaring pong
    float* dold = S1 + EX_NF * EX_NF;                 // [64]
    float* dnew = dold + EX_NF;                       // [64]
    float* vtmp = dnew + EX_NF;                       // [64]
    __shared__ float s_scal[4];
    __shared__ int s_used;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, K = a.K, G = gridDim.x, b = blockIdx.x;
    const int n_sym = EX_NF * (EX_NF + 1) / 2;
    const double fx = a.fx[0];

    // the symmetric entries (row >= col) this thread accumulates, publishes and reads back
    int ea[EX_NE], eb[EX_NE];
#pragma unroll
    for (int j = 0; j < EX_NE; ++j) {
        const int e = tid + j * EX_THREADS;
        int r = (int)floorf((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        while (r * (r + 1) / 2 > e) --r;
        while ((r + 1) * (r + 2) / 2 <= e) ++r;
        ea[j] = (e < n_sym) ? r : -1;
        eb[j] = e - r * (r + 1) / 2;
    }

    for (int c = 0; c < K; ++c) {
        const int lo = __ldg(a.bounds + (size_t)c * (G + 1) + b), hi = __ldg(a.bounds + (size_t)c * (G + 1) + b + 1);
        if (tid < EX_NF) dold[tid] = (tid < n) ? __ldg(a.Dt + (size_t)c * n + tid) : 0.f;
        block_sync();
        // ---- phase 1: M_local = sum over this CTA's users of (r_i + d x_i)(r_i + d x_i)^T            (ksvd.py:36)
        float m[EX_NE];
#pragma unroll
        for (int j = 0; j < EX_NE; ++j) m[j] = 0.f;
        for (int p0 = lo; p0 < hi; p0 += EX_CAP) {
            const int cnt = min(EX_CAP, hi - p0);
            for (int u = warp; u < cnt; u += EX_WARPS) {
                const int e = __ldg(a.entries + p0 + u);
                const float x = __ldcg(a.val + e);
                const float* r = a.R + (int64_t)fdiv(e, a.kdiv) * n;
#pragma unroll
                for (int q = 0; q < EX_NF / 32; ++q) {
                    const int f = lane + 32 * q;
                    rk[u * EX_LDR + f] = (f < n) ? fmaf(dold[f], x, __ldcg(r + f)) : 0.f;
                }
            }
            block_sync();
#pragma unroll
            for (int j = 0; j < EX_NE; ++j) {
                if (ea[j] >= 0) {
                    float s = m[j];
                    for (int u = 0; u < cnt; ++u) s = fmaf(rk[u * EX_LDR + ea[j]], rk[u * EX_LDR + eb[j]], s);
                    m[j] = s;
                }
            }
            if (p0 + EX_CAP < hi) block_sync();          // the next chunk overwrites the staged rows
        }
        // ---- device-wide sum of M and of the user count: fixed point, one 64-bit atomic per word (ksvd_sweep.cu)
        u64* acc = a.acc + (size_t)c * a.lda;
#pragma unroll
        for (int j = 0; j < EX_NE; ++j)
            if (ea[j] >= 0) red_add(acc + tid + j * EX_THREADS, ((u64)__double2ll_rn((double)m[j] * fx) << 8) + 1ull);
        if (tid == 0) red_add(acc + a.lda - 1, ((u64)(hi - lo) << 8) + 1ull);
#pragma unroll
        for (int j = 0; j < EX_NE; ++j) {
            if (ea[j] >= 0) {
                const u64* w = acc + tid + j * EX_THREADS;
                u64 v;
                do { v = ld_relaxed_gpu(w); } while ((v & 255ull) != (u64)G);
                const float val = (float)((double)((long long)(v - (u64)G) >> 8) / fx);
                M0[ea[j] * EX_NF + eb[j]] = val;
                M0[eb[j] * EX_NF + ea[j]] = val;
            }
        }
        if (tid == 0) {
            const u64* w = acc + a.lda - 1;
            u64 v;
            do { v = ld_relaxed_gpu(w); } while ((v & 255ull) != (u64)G);
            s_used = ((long long)(v - (u64)G) >> 8) != 0;
        }
        block_sync();
        const bool used = s_used != 0;                     // uniform over CTAs
        if (!used) {                                      // ksvd.py:32-34
            if (b == 0 && tid < n) a.Dt_new[(size_t)c * n + tid] = dold[tid];
            if (b == 0 && tid == 0) a.unused[c] = 1;
            block_sync();
            continue;
        }
        // ---- top eigenvector of M: normalised repeated squaring, then two power steps with M itself
        if (warp == 0) {
            float t = M0[lane * EX_NF + lane] + M0[(lane + 32) * EX_NF + lane + 32];
            t = warp_sum(t);
            if (lane == 0) s_scal[0] = (t > 0.f) ? 1.f / t : 0.f;
        }
        block_sync();
        {
            const float inv = s_scal[0];
            for (int e = tid; e < EX_NF * EX_NF; e += EX_THREADS) S0[e] = M0[e] * inv;
        }
        block_sync();
        float* src = S0;
        float* dst = S1;
        for (int it = 0; it < EX_SQUARINGS; ++it) {
            const int i = tid >> 3, j0 = (tid & 7) * 8;                 // 8 entries of row i per thread
            float o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = 0.f;
            for (int kk = 0; kk < EX_NF; ++kk) {
                const float sik = src[i * EX_NF + kk];
                const float4 r0 = *reinterpret_cast<const float4*>(src + kk * EX_NF + j0);
                const float4 r1 = *reinterpret_cast<const float4*>(src + kk * EX_NF + j0 + 4);
                o[0] = fmaf(sik, r0.x, o[0]); o[1] = fmaf(sik, r0.y, o[1]); o[2] = fmaf(sik, r0.z, o[2]); o[3] = fmaf(sik, r0.w, o[3]);
                o[4] = fmaf(sik, r1.x, o[4]); o[5] = fmaf(sik, r1.y, o[5]); o[6] = fmaf(sik, r1.z, o[6]); o[7] = fmaf(sik, r1.w, o[7]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[i * EX_NF + j0 + q] = o[q];
            block_sync();
            if (warp == 0) {                                            // trace, fixed order
                float t = dst[lane * EX_NF + lane] + dst[(lane + 32) * EX_NF + lane + 32];
                t = warp_sum(t);
                if (lane == 0) s_scal[0] = (t > 0.f) ? 1.f / t : 0.f;
            }
            block_sync();
            const float inv = s_scal[0];
            for (int e = tid; e < EX_NF * EX_NF; e += EX_THREADS) dst[e] *= inv;
            block_sync();
            float* tmp = src; src = dst; dst = tmp;
        }
        // column of the largest diagonal entry of M^(2^12) (first maximum), as the start vector
        if (warp == 0) {
            float best = -1.f; int bj = 0;
            for (int j = 0; j < EX_NF; ++j) { const float v = src[j * EX_NF + j]; if (v > best) { best = v; bj = j; } }
            vtmp[lane] = src[lane * EX_NF + bj];
            vtmp[lane + 32] = src[(lane + 32) * EX_NF + bj];
        }
        block_sync();
        for (int ps = 0; ps < 3; ++ps) {                                // ps = 0 only normalises; 1, 2: v <- M v / ||M v||
            if (warp == 0) {
                float v0, v1;
                if (ps == 0) { v0 = vtmp[lane]; v1 = vtmp[lane + 32]; }
                else {
                    v0 = 0.f; v1 = 0.f;
                    for (int kk = 0; kk < EX_NF; ++kk) {
                        const float vk = vtmp[kk];
                        v0 = fmaf(M0[lane * EX_NF + kk], vk, v0);
                        v1 = fmaf(M0[(lane + 32) * EX_NF + kk], vk, v1);
                    }
                }
                const float nrm = sqrtf(warp_sum(fmaf(v0, v0, v1 * v1)));
                const float inv = (nrm > 0.f) ? 1.f / nrm : 0.f;
                __syncwarp();
                vtmp[lane] = v0 * inv; vtmp[lane + 32] = v1 * inv;
                __syncwarp();
            }
        }
        if (warp == 0) {
            const float dot = warp_sum(fmaf(vtmp[lane], dold[lane], vtmp[lane + 32] * dold[lane + 32]));
            const float sg = (dot < 0.f) ? -1.f : 1.f;                  // the reference's sign is arbitrary (flip_sign=False)
            dnew[lane] = sg * vtmp[lane]; dnew[lane + 32] = sg * vtmp[lane + 32];
            if (b == 0) {
                if (lane < n) a.Dt_new[(size_t)c * n + lane] = sg * vtmp[lane];
                if (lane + 32 < n) a.Dt_new[(size_t)c * n + lane + 32] = sg * vtmp[lane + 32];
            }
        }
        block_sync();
        // ---- phase 2: x_i = R_k[:,i].d ; R[:,i] = R_k[:,i] - d x_i                                  (ksvd.py:38-41)
        const bool staged = (hi - lo) <= EX_CAP;
        for (int p0 = lo; p0 < hi; p0 += EX_CAP) {
            const int cnt = min(EX_CAP, hi - p0);
            for (int u = warp; u < cnt; u += EX_WARPS) {
                const int e = __ldg(a.entries + p0 + u);
                float* r = a.R + (int64_t)fdiv(e, a.kdiv) * n;
                float rv[EX_NF / 32];
                float dot = 0.f;
                const float xo = staged ? 0.f : __ldcg(a.val + e);
#pragma unroll
                for (int q = 0; q < EX_NF / 32; ++q) {
                    const int f = lane + 32 * q;
                    rv[q] = staged ? rk[u * EX_LDR + f] : ((f < n) ? fmaf(dold[f], xo, __ldcg(r + f)) : 0.f);
                    dot = fmaf(rv[q], dnew[f], dot);
                }
                dot = warp_sum(dot);
#pragma unroll
                for (int q = 0; q < EX_NF / 32; ++q) {
                    const int f = lane + 32 * q;
                    if (f < n) __stcg(r + f, fmaf(-dnew[f], dot, rv[q]));
                }
                if (lane == 0) __stcg(a.val + e, dot);
            }
        }
        block_sync();
    }
}

int exact_lda(int) { return (EX_NF * (EX_NF + 1) / 2 + 1 + 3) / 4 * 4; }

// scale of the fixed-point sums: every entry of M is bounded by ||R_k||_F^2 <= 2 (||R||_F^2 + max||d||^2 ||val||^2)
__global__ void exact_scale_kernel(const double* __restrict__ fro /* ||R||^2, ||val||^2 */, double* __restrict__ fx_out)
{
    const double bound = 4.0 * (fro[0] + fro[1]) * 1.000001;
    double fx = 1.0;
    if (bound > 0.0 && bound < 1e300) fx = exp2(45.0 - ceil(log2(bound)));
    fx_out[0] = fx;
}

}  // namespace
}  // namespace lys

using namespace lys;

namespace {
struct ExactWs { float* Dt; float* Dt_new; u64* acc; size_t acc_bytes; int32_t* bounds; double* fro; double* fx; void* frob_ws; size_t total; };
ExactWs carve_exact(void* base, int n, int K)
{
    ExactWs w;
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    auto take = [&](size_t bytes) { unsigned char* r = p; p += align_up(bytes, 256); return r; };
    w.Dt = reinterpret_cast<float*>(take((size_t)n * K * 4));
    w.Dt_new = reinterpret_cast<float*>(take((size_t)n * K * 4));
    w.acc_bytes = (size_t)K * exact_lda(n) * sizeof(u64);
    w.acc = reinterpret_cast<u64*>(take(w.acc_bytes));
    w.bounds = reinterpret_cast<int32_t*>(take((size_t)K * 256 * 4));
    w.fro = reinterpret_cast<double*>(take(4 * sizeof(double)));
    w.fx = w.fro + 2;
    w.frob_ws = take(sizeof(double) * 4096);
    w.total = (size_t)(p - reinterpret_cast<unsigned char*>(base)) + 256;
    return w;
}
}  // namespace

extern "C" size_t lys_ksvd_exact_workspace_bytes(int n, int K)
{
    if (n < 1 || K < 1) return 0;
    return carve_exact(nullptr, n, K).total;
}

extern "C" int lys_ksvd_exact_sweep(float* R, float* D, int64_t ldd, float* val,
                                    const int32_t* rowptr, const int32_t* entries,
                                    int n, int K, int64_t N, int k, int n_cycles,
                                    int32_t* unused, void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(R && D && val && rowptr && entries && unused && workspace, "lys_ksvd_exact_sweep: null pointer");
    LYS_CHECK_ARG(K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K && k >= 1 && k <= LYS_MAX_NONZERO && n_cycles >= 1 && N >= 0,
                  "lys_ksvd_exact_sweep: bad shape");
    if (n < 1 || n > EX_NF) { set_error("lys_ksvd_exact_sweep: n=%d, the exact atom update is built for n <= %d", n, EX_NF); return LYS_EUNSUPPORTED; }
    LYS_CHECK_ARG(N * (int64_t)k < (1ll << 31), "lys_ksvd_exact_sweep: N*k must fit int32");
    LYS_CHECK_ARG((reinterpret_cast<uintptr_t>(R) & 15) == 0 && (reinterpret_cast<uintptr_t>(val) & 15) == 0,
                  "lys_ksvd_exact_sweep: R and val must be 16-byte aligned");
    if (workspace_bytes < lys_ksvd_exact_workspace_bytes(n, K)) { set_error("lys_ksvd_exact_sweep: workspace too small"); return LYS_EWORKSPACE; }
    const int grid = sweep_grid_size();
    const ExactWs w = carve_exact(workspace, n, K);
    LYS_CUDA(cudaMemsetAsync(unused, 0, sizeof(int32_t) * (size_t)K, stream));
    int rc = transpose(D, ldd, w.Dt, n, n, K, stream);
    if (rc) return rc;
    if ((rc = sweep_bounds(rowptr, entries, K, k, grid, N, w.bounds, stream))) return rc;

    ExactArgs a{};
    a.R = R; a.Dt = w.Dt; a.Dt_new = w.Dt_new; a.val = val; a.entries = entries; a.bounds = w.bounds;
    a.n = n; a.K = K; a.k = k; a.lda = exact_lda(n); a.unused = unused; a.acc = w.acc; a.fx = w.fx;
    a.kdiv = make_fastdiv(k);
    const size_t smem = sizeof(float) * (size_t)(EX_CAP * EX_LDR + 3 * EX_NF * EX_NF + 3 * EX_NF);
    LYS_CUDA(cudaFuncSetAttribute(ksvd_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int cyc = 0; cyc < n_cycles; ++cyc) {
        rc = lys_frobenius2(R, N * (int64_t)n, w.fro, w.frob_ws, sizeof(double) * 4096, stream);
        if (rc) return rc;
        rc = lys_frobenius2(val, N * (int64_t)k, w.fro + 1, w.frob_ws, sizeof(double) * 4096, stream);
        if (rc) return rc;
        exact_scale_kernel<<<1, 1, 0, stream>>>(w.fro, w.fx);
        LYS_LAUNCH_CHECK("exact_scale_kernel");
        LYS_CUDA(cudaMemsetAsync(w.acc, 0, w.acc_bytes, stream));
        void* params[] = {&a};
        LYS_CUDA(cudaLaunchCooperativeKernel((const void*)ksvd_exact_kernel, dim3(grid), dim3(EX_THREADS), params, smem, stream));
        LYS_CUDA(cudaMemcpyAsync(w.Dt, w.Dt_new, (size_t)n * K * 4, cudaMemcpyDeviceToDevice, stream));
    }
    return transpose(w.Dt, n, D, ldd, K, n, stream);
}
