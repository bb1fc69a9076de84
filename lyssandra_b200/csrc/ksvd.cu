// ksvd.cu — the dictionary-update half of approximate K-SVD on sparse codes
// (lyssa/dict_learning/ksvd.py:98-126) plus the residual / error kernels that feed it.
//
//   K5/K10 residual_kernel      R = X - D Z, ||R||_F^2 from (idx,val)   ksvd.py:103, dict_learning/utils.py:14-19
//   K6     atom CSR (3 kernels)  users of every atom                    ksvd.py:111
//   K7-K9  ksvd_sweep_kernel     sequential atom refresh                ksvd.py:105-124
//   K8/K14 norm_cols, gather     utils/math.py:65-71, dict_learning/utils.py:64
//
// Data layout in HBM: R (N,n) signal-major fp32 (one residual = one contiguous 4n-byte row);
// Dt (K,n) atom-major working copy of the dictionary (an atom = one contiguous row); codes
// (idx,val)[N][k]; CSR entries int32 = i*k + slot, ascending per atom.
//
// The sweep is HBM/latency bound, not flop bound: per atom it streams the |users| residual
// rows twice (8nk bytes per signal per iteration) and needs two device-wide
// barriers, so it runs as ONE persistent cooperative kernel (one CTA per SM) with an
// in-kernel grid barrier and fixed-order (deterministic) reductions.
#include "common.cuh"
#include <algorithm>

namespace lys {
namespace {

// ------------------------------------------------------------------ transpose D <-> Dt
__global__ void transpose_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst,
                                 int64_t ldd, int rows, int cols)
{
    // dst[c][r] = src[r][c]
    __shared__ float tile[32][33];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int r = r0 + y, c = c0 + threadIdx.x;
        tile[y][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int c = c0 + y, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(int64_t)c * ldd + r] = tile[threadIdx.x][y];
    }
}

}  // namespace

int transpose(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, cudaStream_t st)
{
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, st>>>(src, lds, dst, ldd, rows, cols);
    LYS_LAUNCH_CHECK("transpose_kernel");
    return LYS_OK;
}

namespace {

// ------------------------------------------------------------------------- residual
constexpr int RES_WARPS = 8;
constexpr int RES_TILE = 32;       // signals per tile
constexpr int MAX_NPL = LYS_MAX_FEATURES / 32;

template <int NPL>
__global__ void __launch_bounds__(RES_WARPS * 32)
residual_kernel(const float* __restrict__ X, int64_t xfs, int64_t xss,
                const float* __restrict__ Dt, const int32_t* __restrict__ idx,
                const float* __restrict__ val, int n, int64_t N, int k,
                float* __restrict__ R, double* __restrict__ partial)
{
    extern __shared__ float xs[];          // [RES_TILE][n + 1]
    __shared__ double wsum[RES_WARPS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ldx = n + 1;
    const int64_t n_tiles = (N + RES_TILE - 1) / RES_TILE;
    double esum = 0.0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * RES_TILE;
        __syncthreads();
        if (xfs == 1) {              // signal-major: coalesce along features
            for (int s = warp; s < RES_TILE; s += RES_WARPS) {
                int64_t i = i0 + s;
                for (int f = lane; f < n; f += 32)
                    xs[s * ldx + f] = (i < N) ? X[i * xss + f] : 0.f;
            }
        } else {                     // feature-major (reference layout): coalesce along signals
            for (int f = warp; f < n; f += RES_WARPS) {
                int64_t i = i0 + lane;
                xs[lane * ldx + f] = (i < N) ? X[(int64_t)f * xfs + i * xss] : 0.f;
            }
        }
        __syncthreads();
        for (int s = warp; s < RES_TILE; s += RES_WARPS) {
            const int64_t i = i0 + s;
            if (i >= N) break;
            float r[NPL];
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                int f = lane + 32 * q;
                r[q] = (f < n) ? xs[s * ldx + f] : 0.f;
            }
            // the k codes of the signal in one coalesced load (lane j holds code j), broadcast by shuffle; atoms are
            // fetched four at a time so that their L2 latencies overlap (an absent code reads atom 0 with z = 0)
            int my_a = 0; float my_z = 0.f;
            if (lane < k) {
                const int a = idx[i * k + lane];
                if (a >= 0) { my_a = a; my_z = val[i * k + lane]; }
            }
            for (int j0 = 0; j0 < k; j0 += 4) {
                float dv[4][NPL], z[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int a = __shfl_sync(0xffffffffu, my_a, (j0 + u) & 31);
                    z[u] = (j0 + u < k) ? __shfl_sync(0xffffffffu, my_z, (j0 + u) & 31) : 0.f;
                    const float* d = Dt + (int64_t)a * n;
#pragma unroll
                    for (int q = 0; q < NPL; ++q) { const int f = lane + 32 * q; dv[u][q] = (f < n) ? __ldg(d + f) : 0.f; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int q = 0; q < NPL; ++q) r[q] = fmaf(-z[u], dv[u][q], r[q]);
            }
            float sq = 0.f;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                int f = lane + 32 * q;
                if (f < n) {
                    if (R) R[i * n + f] = r[q];
                    sq = fmaf(r[q], r[q], sq);
                }
            }
            esum += (double)sq;
        }
    }
    esum = warp_sum(esum);
    if (lane == 0) wsum[warp] = esum;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < RES_WARPS; ++w) s += wsum[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void sum_partials_kernel(const double* __restrict__ partial, int count, double* __restrict__ out)
{
    // fixed-order pairwise sum by one warp -> deterministic
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += 32) s += partial[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) *out = s;
}

// ||A||_F^2 of a contiguous array: per-block partial in double, fixed-order final sum (deterministic)
__global__ void __launch_bounds__(256)
frob2_kernel(const float* __restrict__ A, int64_t count, double* __restrict__ partial)
{
    __shared__ double wsum[8];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double s = 0.0;
    const int64_t n4 = count / 4;
    const float4* A4 = reinterpret_cast<const float4*>(A);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + t; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(A4 + i);
        s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && t < (int)(count - n4 * 4)) { const float v = A[n4 * 4 + t]; s += (double)v * v; }
    s = warp_sum(s);
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (t == 0) { double r = 0.0; for (int w = 0; w < 8; ++w) r += wsum[w]; partial[blockIdx.x] = r; }
}

// --------------------------------------------------------------------------- atom CSR
constexpr int CSR_WARPS = 16;    // warps per CTA (upper bound; fewer when K * 4 B per warp would not fit in shared memory)
// Each warp counts / places its own contiguous slice of the codes, so the output order is deterministic; the
// slices are walked 32 entries at a time with dependent shared-memory updates, i.e. latency-bound: 16 warps per
// SM instead of 4 hide it (2.0 -> 0.84 ms at cfg3).

__device__ __forceinline__ int csr_atom(const int32_t* idx, const float* val, int64_t e, int64_t E)
{
    if (e >= E) return -1;
    int a = idx[e];
    if (a < 0) return -1;
    return (val[e] != 0.f) ? a : -1;        // ksvd.py:111  omega_k = X[k,:] != 0
}

__global__ void __launch_bounds__(CSR_WARPS * 32)
csr_hist_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, int64_t E, int K,
                int64_t per_warp, int32_t* __restrict__ hist /* [n_warps][K] */)
{
    extern __shared__ int32_t sh[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t* h = sh + warp * K;
    for (int c = lane; c < K; c += 32) h[c] = 0;
    __syncwarp();
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t lo = gw * per_warp, hi = min(lo + per_warp, E);
    for (int64_t e = lo + lane; e < hi; e += 32) {
        int a = csr_atom(idx, val, e, E);
        if (a >= 0) atomicAdd(&h[a], 1);
    }
    __syncwarp();
    for (int c = lane; c < K; c += 32) hist[gw * K + c] = h[c];
}

__global__ void __launch_bounds__(1024)
csr_scan_kernel(int32_t* __restrict__ hist, int n_warps, int K, int32_t* __restrict__ rowptr)
{
    // 1) per atom: exclusive prefix over the warp-chunks (in chunk order), total per atom
    // 2) exclusive scan of the totals over atoms -> rowptr
    __shared__ int32_t tot[LYS_MAX_ATOMS];
    __shared__ int32_t wtot[32];
    const int t = threadIdx.x;
    for (int c = t; c < K; c += 1024) {
        int32_t run = 0;
        for (int w = 0; w < n_warps; ++w) {
            int32_t v = hist[(int64_t)w * K + c];
            hist[(int64_t)w * K + c] = run;
            run += v;
        }
        tot[c] = run;
    }
    __syncthreads();
    // each thread owns 4 consecutive atoms (K <= 4096)
    int32_t v[4], s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { int c = t * 4 + q; v[q] = (c < K) ? tot[c] : 0; s += v[q]; }
    int32_t inc = s;
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int32_t w = wtot[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t u = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += u;
        }
        wtot[lane] = winc - w;      // exclusive
    }
    __syncthreads();
    int32_t base = wtot[warp] + inc - s;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int c = t * 4 + q;
        if (c < K) rowptr[c] = base;
        base += v[q];
        if (c == K - 1) rowptr[K] = base;
    }
}

__global__ void __launch_bounds__(CSR_WARPS * 32)
csr_fill_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, int64_t E, int K,
                int64_t per_warp, const int32_t* __restrict__ hist, const int32_t* __restrict__ rowptr,
                int32_t* __restrict__ entries)
{
    extern __shared__ int32_t sh[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t* base = sh + warp * K;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    for (int c = lane; c < K; c += 32) base[c] = rowptr[c] + hist[gw * K + c];
    __syncwarp();
    const int64_t lo = gw * per_warp, hi = min(lo + per_warp, E);
    for (int64_t e0 = lo; e0 < hi; e0 += 32) {
        const int64_t e = e0 + lane;
        const int a = (e < hi) ? csr_atom(idx, val, e, E) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, a);
        if (a >= 0) {
            const int rank = __popc(peers & ((1u << lane) - 1u));
            entries[base[a] + rank] = (int32_t)e;
        }
        __syncwarp();
        if (a >= 0 && lane == (__ffs(peers) - 1)) base[a] += __popc(peers);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ norm_cols / gather
__global__ void norm_cols_kernel(float* __restrict__ D, int64_t ldd, int n, int K)
{
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= K) return;
    float sq = 0.f;
    for (int f = lane; f < n; f += 32) { float v = D[(int64_t)f * ldd + c]; sq = fmaf(v, v, sq); }
    sq = warp_sum(sq);
    const float inv = 1.f / (sqrtf(sq) + kRefEps);
    for (int f = lane; f < n; f += 32) D[(int64_t)f * ldd + c] *= inv;
}

__global__ void gather_cols_kernel(const float* __restrict__ X, int64_t xfs, int64_t xss, int n,
                                   const int64_t* __restrict__ cols, int n_cols,
                                   float* __restrict__ D, int64_t ldd, const int32_t* __restrict__ dst)
{
    const int j = blockIdx.x;
    if (j >= n_cols) return;
    const int64_t src = cols[j];
    const int d = dst ? dst[j] : j;
    for (int f = threadIdx.x; f < n; f += blockDim.x) D[(int64_t)f * ldd + d] = X[(int64_t)f * xfs + src * xss];
}

}  // namespace
}  // namespace lys

using namespace lys;

// ============================================================================ C-ABI
extern "C" size_t lys_residual_workspace_bytes(int n, int K, int64_t N)
{
    (void)N;
    return align_up((size_t)n * K * 4, 256) + align_up(sizeof(double) * 4096, 256);
}

extern "C" int lys_residual(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd,
                            const int32_t* idx, const float* val, int n, int K, int64_t N, int k,
                            float* R, double* err, void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(X && D && idx && val && workspace, "lys_residual: null pointer");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K && k >= 1 && k <= LYS_MAX_NONZERO && N >= 0,
                  "lys_residual: bad shape");
    if (workspace_bytes < lys_residual_workspace_bytes(n, K, N)) { set_error("lys_residual: workspace too small"); return LYS_EWORKSPACE; }
    float* Dt = reinterpret_cast<float*>(workspace);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + align_up((size_t)n * K * 4, 256));
    int rc = transpose(D, ldd, Dt, n, n, K, stream);
    if (rc) return rc;
    int64_t tiles = (N + RES_TILE - 1) / RES_TILE;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, std::min<int64_t>(4096, (int64_t)sm_count() * 8)));
    size_t smem = sizeof(float) * RES_TILE * (n + 1);
    if (n <= 32) residual_kernel<1><<<grid, RES_WARPS * 32, smem, stream>>>(X, xfs, xss, Dt, idx, val, n, N, k, R, partial);
    else if (n <= 64) residual_kernel<2><<<grid, RES_WARPS * 32, smem, stream>>>(X, xfs, xss, Dt, idx, val, n, N, k, R, partial);
    else if (n <= 128) residual_kernel<4><<<grid, RES_WARPS * 32, smem, stream>>>(X, xfs, xss, Dt, idx, val, n, N, k, R, partial);
    else residual_kernel<MAX_NPL><<<grid, RES_WARPS * 32, smem, stream>>>(X, xfs, xss, Dt, idx, val, n, N, k, R, partial);
    LYS_LAUNCH_CHECK("residual_kernel");
    if (err) {
        sum_partials_kernel<<<1, 32, 0, stream>>>(partial, grid, err);
        LYS_LAUNCH_CHECK("sum_partials_kernel");
    }
    return LYS_OK;
}

extern "C" int lys_frobenius2(const float* A, int64_t count, double* out, void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(A && out && workspace && count >= 0, "lys_frobenius2: bad argument");
    LYS_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0, "lys_frobenius2: A must be 16-byte aligned");
    const int grid = std::min(4096, sm_count() * 8);
    if (workspace_bytes < sizeof(double) * 4096) { set_error("lys_frobenius2: workspace too small (32 KB)"); return LYS_EWORKSPACE; }
    double* partial = reinterpret_cast<double*>(workspace);
    frob2_kernel<<<grid, 256, 0, stream>>>(A, count, partial);
    LYS_LAUNCH_CHECK("frob2_kernel");
    sum_partials_kernel<<<1, 32, 0, stream>>>(partial, grid, out);
    LYS_LAUNCH_CHECK("sum_partials_kernel");
    return LYS_OK;
}

static int csr_warps_per_cta(int K) { return std::max(1, std::min(CSR_WARPS, (int)((200 * 1024) / ((size_t)K * 4)))); }
static int csr_n_warps(int K) { return sm_count() * csr_warps_per_cta(K); }

extern "C" size_t lys_atom_csr_workspace_bytes(int K, int64_t N, int k)
{
    (void)N; (void)k;
    return align_up((size_t)csr_n_warps(K) * (size_t)K * 4, 256);
}

extern "C" int lys_build_atom_csr(const int32_t* idx, const float* val, int64_t N, int k, int K,
                                  int32_t* rowptr, int32_t* entries, void* workspace, size_t workspace_bytes,
                                  void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(idx && val && rowptr && entries && workspace, "lys_build_atom_csr: null pointer");
    LYS_CHECK_ARG(K >= 1 && K <= LYS_MAX_ATOMS && k >= 1 && N >= 0, "lys_build_atom_csr: bad shape");
    LYS_CHECK_ARG(N * (int64_t)k < (1ll << 31), "lys_build_atom_csr: N*k must fit int32");
    if (workspace_bytes < lys_atom_csr_workspace_bytes(K, N, k)) { set_error("lys_build_atom_csr: workspace too small"); return LYS_EWORKSPACE; }
    int32_t* hist = reinterpret_cast<int32_t*>(workspace);
    const int nw = csr_n_warps(K), wpc = csr_warps_per_cta(K);
    const int64_t E = N * (int64_t)k;
    int64_t per_warp = (E + nw - 1) / nw;
    per_warp = std::max<int64_t>(32, (per_warp + 31) / 32 * 32);
    size_t smem = sizeof(int32_t) * (size_t)wpc * K;
    if (smem > 48 * 1024) {
        LYS_CUDA(cudaFuncSetAttribute(csr_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LYS_CUDA(cudaFuncSetAttribute(csr_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    csr_hist_kernel<<<nw / wpc, wpc * 32, smem, stream>>>(idx, val, E, K, per_warp, hist);
    LYS_LAUNCH_CHECK("csr_hist_kernel");
    csr_scan_kernel<<<1, 1024, 0, stream>>>(hist, nw, K, rowptr);
    LYS_LAUNCH_CHECK("csr_scan_kernel");
    csr_fill_kernel<<<nw / wpc, wpc * 32, smem, stream>>>(idx, val, E, K, per_warp, hist, rowptr, entries);
    LYS_LAUNCH_CHECK("csr_fill_kernel");
    return LYS_OK;
}

extern "C" int lys_norm_cols(float* D, int64_t ldd, int n, int K, void* stream)
{
    LYS_CHECK_ARG(D && n >= 1 && K >= 1 && ldd >= K, "lys_norm_cols: bad argument");
    norm_cols_kernel<<<(K * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(D, ldd, n, K);
    LYS_LAUNCH_CHECK("norm_cols_kernel");
    return LYS_OK;
}

extern "C" int lys_gather_cols(const float* X, int64_t xfs, int64_t xss, int n, const int64_t* cols, int n_cols,
                               float* D, int64_t ldd, const int32_t* dst_cols, void* stream)
{
    LYS_CHECK_ARG(X && cols && D && n >= 1 && n_cols >= 0, "lys_gather_cols: bad argument");
    if (n_cols == 0) return LYS_OK;
    gather_cols_kernel<<<n_cols, 64, 0, (cudaStream_t)stream>>>(X, xfs, xss, n, cols, n_cols, D, ldd, dst_cols);
    LYS_LAUNCH_CHECK("gather_cols_kernel");
    return LYS_OK;
}
