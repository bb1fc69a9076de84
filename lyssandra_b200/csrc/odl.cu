// odl.cu — Mairal online dictionary learning as the reference implements it
// (lyssa/dict_learning/online_dict_learn.py:79-98), on sparse codes.
//
//   K12 odl_accumulate   A = beta*A + Z_b Z_b^T ; B = beta*B + X_b Z_b^T        :84-85
//       The reference multiplies DENSE K x b matrices (2 K^2 b flop, >99 % zeros); with
//       (idx,val)[b][k] it is b*k^2 + b*k*n scattered adds.
//   K13 odl_update_dict  D <- norm_cols(clamp(D + (B - D A) diag(1/(A_kk + eps))))  :91-98
//       Jacobi: D A is formed ONCE from the old D (one GEMM), not refreshed per atom.
#include "common.cuh"
#include <algorithm>

namespace lys {

int da_gemm_tc_splits(int K);
bool da_gemm_tc_supported(int n, int K);
int da_gemm_tc(const float* D, int64_t ldd, const float* A, int n, int K, float* partial, cudaStream_t stream);

namespace {

// ---- users-of-atom lists of one minibatch, ONE CTA (b*k is a few 10^4 entries): histogram, scan and fill in shared
// memory.  The order inside a list is whatever the atomics produce; the consumer turns every list into a bitmap over
// the signals of the minibatch and walks that in ascending order, so the statistics do not depend on it.
__global__ void __launch_bounds__(1024)
odl_csr_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, int64_t E, int K,
               int32_t* __restrict__ rowptr /* K+1 */, int32_t* __restrict__ entries /* E */)
{
    extern __shared__ int32_t sh[];            // counts / cursors [K]
    __shared__ int32_t wtot[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int c = t; c < K; c += 1024) sh[c] = 0;
    __syncthreads();
    for (int64_t e0 = t; e0 < E; e0 += 4 * 1024) {          // four loads in flight per thread: the kernel is one CTA of latency
        int a[4];
        float v4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t e = e0 + (int64_t)q * 1024;
            a[q] = (e < E) ? __ldg(idx + e) : -1;
            v4[q] = (e < E) ? __ldg(val + e) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (a[q] >= 0 && v4[q] != 0.f) atomicAdd(&sh[a[q]], 1);
    }
    __syncthreads();
    // exclusive scan: every thread owns 4 consecutive atoms (K <= 4096)
    int32_t v[4], sum = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int c = t * 4 + q; v[q] = (c < K) ? sh[c] : 0; sum += v[q]; }
    int32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int32_t w = wtot[lane];
        int32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += u; }
        wtot[lane] = winc - w;
    }
    __syncthreads();
    int32_t base = wtot[warp] + inc - sum;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = t * 4 + q;
        if (c < K) { rowptr[c] = base; sh[c] = base; }
        base += v[q];
        if (c == K - 1) rowptr[K] = base;
    }
    __syncthreads();
    for (int64_t e0 = t; e0 < E; e0 += 4 * 1024) {
        int a[4];
        float v4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t e = e0 + (int64_t)q * 1024;
            a[q] = (e < E) ? __ldg(idx + e) : -1;
            v4[q] = (e < E) ? __ldg(val + e) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (a[q] >= 0 && v4[q] != 0.f) entries[atomicAdd(&sh[a[q]], 1)] = (int32_t)(e0 + (int64_t)q * 1024);
    }
}

// A = beta A + Z Z^T, B = beta B + X Z^T, one CTA per atom a:   row a of A and column a of B            (:84-85)
//   A[a][c] = beta A[a][c] + sum over the users i of a of z_ia z_ic
//   B[f][a] = beta B[f][a] + sum over the users i of a of x_if z_ia
// Fixed summation order, hence bitwise reproducible and identical on every rank that sees the same minibatch: a signal
// uses an atom at most once, so the atom's (unordered) list becomes a BITMAP over the b signals of the minibatch plus
// the code slot of every user (order-independent writes); the signal range is cut into `active` contiguous parts, every
// warp walks the set bits of its part in ascending order into its own partial row (lanes = the k codes of a signal /
// the features), and the partial rows are added in warp order.  An atom that most signals of the minibatch use (the
// "mean" atom of non-negative descriptors has ~b users) is thus neither a sequential tail nor a source of run-to-run
// differences.  (Round 2's first version sorted every list with a bitonic network in shared memory: 78 stages for the
// 4096-user atom, a third of the kernel's 102 us.)
// beta == 0 writes the sums alone (0 * x of the reference without its NaN propagation: beta_0 = 0 exists to wipe the
// statistics).
constexpr int ODL_WARPS = 8;
constexpr int ODL_THREADS = ODL_WARPS * 32;
__global__ void __launch_bounds__(ODL_THREADS)
odl_accumulate_kernel(const float* __restrict__ X, int64_t xfs, int64_t xss,
                      const int32_t* __restrict__ idx, const float* __restrict__ val,
                      const int32_t* __restrict__ rowptr, const int32_t* __restrict__ entries,
                      int n, int K, int k, int b /* signals in the minibatch */, float beta,
                      float* __restrict__ A, float* __restrict__ B)
{
    extern __shared__ float smf[];
    const int ldp = K + n;                                  // a warp's partial: [K] new terms of row a of A, [n] of column a of B
    const int words = (b + 31) >> 5;
    float* part = smf;                                      // [ODL_WARPS][ldp]
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(smf + (size_t)ODL_WARPS * ldp);   // [words] users of the atom
    unsigned char* slot = reinterpret_cast<unsigned char*>(bitmap + words);          // [b] code slot of every user
    const int a = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int lo = rowptr[a], cnt = rowptr[a + 1] - lo;
    if (cnt == 0) {                                         // nobody uses the atom in this minibatch: only the decay
        float* Arow = A + (int64_t)a * K;
        for (int c = t; c < K; c += ODL_THREADS) Arow[c] = (beta == 0.f) ? 0.f : beta * Arow[c];
        for (int f = t; f < n; f += ODL_THREADS) B[(int64_t)f * K + a] = (beta == 0.f) ? 0.f : beta * B[(int64_t)f * K + a];
        return;
    }
    const int active = min(ODL_WARPS, (cnt + 3) / 4);        // warps that get a part (about 4 users per warp at least)
    for (int w = t; w < words; w += ODL_THREADS) bitmap[w] = 0u;
    for (int c = t; c < active * ldp; c += ODL_THREADS) part[c] = 0.f;
    __syncthreads();
    for (int p = t; p < cnt; p += ODL_THREADS) {
        const int32_t e = entries[lo + p];
        const int i = e / k;
        atomicOr(&bitmap[i >> 5], 1u << (i & 31));
        slot[i] = (unsigned char)(e - i * k);
    }
    __syncthreads();
    if (warp < active) {
        const int per = (words + active - 1) / active;       // bitmap words per warp
        const int w0 = warp * per, w1 = min(words, w0 + per);
        float* dA = part + (size_t)warp * ldp;
        float* dB = dA + K;
        constexpr int UB = 4;                                // users whose loads are in flight together
        constexpr int FQ = LYS_MAX_FEATURES / 32;
        for (int w = w0; w < w1; ++w) {                      // this warp's users in ascending signal order
            uint32_t bits = bitmap[w];                       // the same word in every lane: the loops are warp-uniform
            while (bits) {
                {   // k <= LYS_MAX_NONZERO = 32: one code per lane
                    // the walk of a popular atom is a chain of dependent loads per user (slot -> coefficient -> code ->
                    // signal): the loads of UB users are issued together, the additions keep the ascending order
                    int ui[UB], uc[UB];
                    float uz[UB], uv[UB], ux[UB][FQ];
#pragma unroll
                    for (int q = 0; q < UB; ++q) {
                        ui[q] = bits ? (w << 5) + __ffs(bits) - 1 : -1;
                        bits &= bits - 1;                    // 0 & anything stays 0
                    }
#pragma unroll
                    for (int q = 0; q < UB; ++q) {
                        uc[q] = -1; uz[q] = 0.f; uv[q] = 0.f;
                        if (ui[q] >= 0) {
                            const int64_t base = (int64_t)ui[q] * k;
                            uz[q] = val[base + slot[ui[q]]];
                            if (lane < k) { uc[q] = idx[base + lane]; uv[q] = val[base + lane]; }
#pragma unroll
                            for (int m = 0; m < FQ; ++m) {
                                const int f = lane + 32 * m;
                                ux[q][m] = (f < n) ? X[(int64_t)f * xfs + ui[q] * xss] : 0.f;
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < UB; ++q) {
                        if (ui[q] >= 0) {                    // warp-uniform
                            if (uc[q] >= 0) dA[uc[q]] = fmaf(uz[q], uv[q], dA[uc[q]]);   // distinct atoms inside one code: no conflict
#pragma unroll
                            for (int m = 0; m < FQ; ++m) {
                                const int f = lane + 32 * m;
                                if (f < n) dB[f] = fmaf(ux[q][m], uz[q], dB[f]);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    float* Arow = A + (int64_t)a * K;
    for (int c = t; c < ldp; c += ODL_THREADS) {
        float s = 0.f;
        for (int w = 0; w < active; ++w) s += part[(size_t)w * ldp + c];        // fixed order
        if (c < K) Arow[c] = (beta == 0.f) ? s : fmaf(beta, Arow[c], s);
        else { float* bp = B + (int64_t)(c - K) * K + a; *bp = (beta == 0.f) ? s : fmaf(beta, *bp, s); }
    }
}

// one warp per atom: u = D[:,c] + (B[:,c] - DA[:,c]) / (A[c,c] + eps); clamp; normalise.  DA arrives as `splits` partial
// products over slices of the contraction (tensor-core GEMM, odl_gemm_tc.cu; splits == 1 for the SIMT GEMM): they are
// added here in split order — fixed, hence deterministic.
__global__ void __launch_bounds__(256)
odl_update_kernel(float* __restrict__ D, int64_t ldd, const float* __restrict__ A,
                  const float* __restrict__ B, const float* __restrict__ DA, int splits, int n, int K, int non_neg)
{
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= K) return;
    const float inv = 1.f / (A[(int64_t)c * K + c] + kRefEps);                 // :94
    float u[LYS_MAX_FEATURES / 32];
    float sq = 0.f;
#pragma unroll
    for (int q = 0; q < LYS_MAX_FEATURES / 32; ++q) {
        int f = lane + 32 * q;
        float v = 0.f;
        if (f < n) {
            float da = 0.f;
            for (int s = 0; s < splits; ++s) da += DA[((int64_t)s * n + f) * K + c];
            v = inv * (B[(int64_t)f * K + c] - da) + D[(int64_t)f * ldd + c];
            if (non_neg && v < 0.f) v = 0.f;                                  // :96-97
        }
        u[q] = v;
        sq = fmaf(v, v, sq);
    }
    sq = warp_sum(sq);
    const float s = 1.f / (sqrtf(sq) + kRefEps);                               // :98 norm_cols
#pragma unroll
    for (int q = 0; q < LYS_MAX_FEATURES / 32; ++q) {
        int f = lane + 32 * q;
        if (f < n) D[(int64_t)f * ldd + c] = u[q] * s;
    }
}

}  // namespace
}  // namespace lys

using namespace lys;

extern "C" size_t lys_odl_accumulate_workspace_bytes(int K, int64_t b, int k)
{
    if (K < 1 || b < 0 || k < 1) return 0;
    return align_up((size_t)(K + 1) * 4, 256) + align_up((size_t)std::max<int64_t>(b * k, 1) * 4, 256) + 256;
}

extern "C" int lys_odl_accumulate(const float* Xb, int64_t xfs, int64_t xss, const int32_t* idx, const float* val,
                                  int n, int K, int64_t b, int k, float beta, float* A, float* B,
                                  void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(A && B, "lys_odl_accumulate: null statistics");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && b >= 0 && k >= 1 && k <= LYS_MAX_NONZERO,
                  "lys_odl_accumulate: bad shape");
    LYS_CHECK_ARG(b * (int64_t)k < (1ll << 31), "lys_odl_accumulate: b*k must fit int32");
    LYS_CHECK_ARG(b == 0 || (Xb && idx && val), "lys_odl_accumulate: null pointer");
    if (!workspace || workspace_bytes < lys_odl_accumulate_workspace_bytes(K, b, k)) {
        set_error("lys_odl_accumulate: workspace too small");
        return LYS_EWORKSPACE;
    }
    int32_t* rowptr = reinterpret_cast<int32_t*>(workspace);
    int32_t* entries = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(workspace) + align_up((size_t)(K + 1) * 4, 256));
    odl_csr_kernel<<<1, 1024, sizeof(int32_t) * (size_t)K, stream>>>(idx, val, b * (int64_t)k, K, rowptr, entries);
    LYS_LAUNCH_CHECK("odl_csr_kernel");
    // partial rows + the bitmap of the atom's users over the minibatch + one slot byte per signal
    const size_t smem = sizeof(float) * (size_t)ODL_WARPS * (K + n) + sizeof(uint32_t) * (size_t)((b + 31) / 32) + align_up((size_t)b, 16);
    if (smem > 200 * 1024) { set_error("lys_odl_accumulate: minibatch of %lld signals with K=%d does not fit the statistics kernel's shared memory", (long long)b, K); return LYS_EUNSUPPORTED; }
    LYS_CUDA(cudaFuncSetAttribute(odl_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    odl_accumulate_kernel<<<K, ODL_THREADS, smem, stream>>>(Xb, xfs, xss, idx, val, rowptr, entries, n, K, k, (int)b, beta, A, B);
    LYS_LAUNCH_CHECK("odl_accumulate_kernel");
    return LYS_OK;
}

extern "C" size_t lys_odl_update_workspace_bytes(int n, int K)
{
    if (n < 1 || K < 1) return 0;
    return align_up((size_t)n * K * 4 * (size_t)std::max(1, da_gemm_tc_splits(K)), 256);
}

extern "C" int lys_odl_update_dict(float* D, int64_t ldd, const float* A, const float* B, int n, int K,
                                   int non_neg, void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(D && A && B && workspace, "lys_odl_update_dict: null pointer");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K, "lys_odl_update_dict: bad shape");
    if (workspace_bytes < lys_odl_update_workspace_bytes(n, K)) { set_error("lys_odl_update_dict: workspace too small"); return LYS_EWORKSPACE; }
    float* DA = reinterpret_cast<float*>(workspace);
    int splits = 1;
    int rc;
    if (da_gemm_tc_supported(n, K)) {                                          // DA = D A on tcgen05 (:91), partial products
        splits = da_gemm_tc_splits(K);
        rc = da_gemm_tc(D, ldd, A, n, K, DA, stream);
    } else {
        rc = sgemm_strided(D, ldd, 1, A, K, 1, DA, K, 1, n, K, K, stream);      // n > 128: fp32 SIMT
    }
    if (rc) return rc;
    odl_update_kernel<<<(K * 32 + 255) / 256, 256, 0, stream>>>(D, ldd, A, B, DA, splits, n, K, non_neg);
    LYS_LAUNCH_CHECK("odl_update_kernel");
    return LYS_OK;
}
