// ksvd_sweep.cu — the atom loop of approximate K-SVD (lyssa/dict_learning/ksvd.py:105-124) as ONE
// persistent cooperative kernel, fused with its per-atom reduction across CTAs and, when the
// signals are sharded over GPUs, across ranks (peer-mapped NVLink buffers, no host round trip).
//
// Per atom c (sequential, Gauss-Seidel order is part of the reference's result):
//     users = columns with Z[c,:] != 0                      :111   (CSR built by lys_build_atom_csr)
//     s  = R[:,users] x + d (x.x)          == R_k x         :116-118   (R_k never materialised)
//     d' = s / (||s|| + eps)                                :119
//     x' = R[:,users]^T d' + x (d.d')      == R_k^T d'      :121
//     R[:,users] += d x^T - d' x'^T        == R_k - d' x'   :123
// Ownership: CTA b owns the contiguous signal range [b*S, (b+1)*S).  The users of an atom are
// sorted by signal, so each CTA's users form one sub-range of the CSR segment (`bounds`,
// precomputed by binary search).  Because a residual row is only ever touched by its owner CTA,
// the only device-wide dependency per atom is the (n+1)-float sum s, sxx:
//     phase 1 (users' rows -> registers, partial sums) -> per-CTA partial + release flag
//     -> every CTA polls the flags of all CTAs (barrier and data fetch in one L2 round trip)
//     -> fixed-order two-level sum (deterministic, replicated) [-> rank exchange] -> d'
//     -> phase 2 from the register-resident rows -> next atom.   ONE grid-wide sync per atom.
// Multi-rank: CTA 0 writes the rank's (n+2) floats (s, sxx, user count) into every peer's slot
// and raises its flag there (st.release.sys over NVLink); all CTAs poll their own rank's flags
// and sum the slots in rank order, so every rank computes bit-identical d'.
#include "comm.cuh"
#include <algorithm>

namespace lys {
namespace {

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_volatile_f32(const float* p)
{
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// bounds[c][b] = first CSR position of atom c whose signal is >= b*S   (b = 0..G), bounds[c][G] = rowptr[c+1]
__global__ void sweep_bounds_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ entries,
                                    int K, int k, int G, int64_t S, int32_t* __restrict__ bounds)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= K * (G + 1)) return;
    const int c = id / (G + 1), b = id % (G + 1);
    int lo = rowptr[c], hi = rowptr[c + 1];
    const int64_t first_signal = (int64_t)b * S;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)(entries[mid] / k) < first_signal) lo = mid + 1; else hi = mid;
    }
    bounds[id] = lo;
}

constexpr int SW_WARPS = 32;
constexpr int SW_U = 8;          // users per warp whose residual rows stay in registers between the phases

template <int NPL>
__global__ void __launch_bounds__(SW_WARPS * 32, 1)
ksvd_sweep_kernel(float* __restrict__ R, float* __restrict__ Dt, float* __restrict__ val,
                  const int32_t* __restrict__ rowptr, const int32_t* __restrict__ entries,
                  const int32_t* __restrict__ bounds,
                  int n, int K, int k, int n_cycles,
                  int32_t* __restrict__ unused,
                  float* __restrict__ partial /* [2][grid][n+1] */, unsigned* __restrict__ flags /* [2][grid] */,
                  PeerComm pc, unsigned comm_seq0)
{
    extern __shared__ float sm[];
    float* d_old = sm;                       // [n]
    float* d_new = d_old + n;                // [n]
    float* svec = d_new + n;                 // [n + 2]
    float* red = svec + (n + 2);             // [SW_WARPS][n + 1]
    __shared__ float s_g;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ldr = n + 1;
    const int G = gridDim.x, b = blockIdx.x;
    const int groups = (SW_WARPS * 32) / ldr;
    unsigned seq = 0;

    for (int cyc = 0; cyc < n_cycles; ++cyc) {
        for (int c = 0; c < K; ++c) {
            const int local_count = rowptr[c + 1] - rowptr[c];
            if (pc.world == 1 && local_count == 0) {         // ksvd.py:112-115
                if (b == 0 && t == 0) unused[c] = 1;
                // the one-thread store above diverges warp 0; without an explicit reconvergence the
                // `continue` let its lanes reach the next bar.sync separately (found by
                // compute-sanitizer racecheck: barrier phases slipped after every unused atom)
                __syncwarp();
                continue;
            }
            ++seq;
            const int par = seq & 1;
            const int lo = bounds[c * (G + 1) + b], hi = bounds[c * (G + 1) + b + 1];
            for (int f = t; f < n; f += blockDim.x) d_old[f] = Dt[(int64_t)c * n + f];

            // ---- phase 1: this CTA's users; ids, coefficients and rows as one batch of independent loads
            int ent[SW_U]; float x[SW_U]; float rv[SW_U][NPL];
            float acc[NPL];
#pragma unroll
            for (int q = 0; q < NPL; ++q) acc[q] = 0.f;
            float sxx = 0.f;
#pragma unroll
            for (int u = 0; u < SW_U; ++u) {
                const int p = lo + warp + u * SW_WARPS;
                ent[u] = (p < hi) ? entries[p] : -1;
            }
#pragma unroll
            for (int u = 0; u < SW_U; ++u) {
                x[u] = 0.f;
#pragma unroll
                for (int q = 0; q < NPL; ++q) rv[u][q] = 0.f;
                if (ent[u] >= 0) {
                    x[u] = val[ent[u]];
                    const float* r = R + (int64_t)(ent[u] / k) * n;
#pragma unroll
                    for (int q = 0; q < NPL; ++q) { const int f = lane + 32 * q; if (f < n) rv[u][q] = r[f]; }
                }
            }
#pragma unroll
            for (int u = 0; u < SW_U; ++u) {
#pragma unroll
                for (int q = 0; q < NPL; ++q) acc[q] = fmaf(rv[u][q], x[u], acc[q]);
                sxx = fmaf(x[u], x[u], sxx);
            }
            for (int p = lo + warp + SW_U * SW_WARPS; p < hi; p += SW_WARPS) {       // overflow users (rare)
                const int e2 = entries[p];
                const float x2 = val[e2];
                const float* r = R + (int64_t)(e2 / k) * n;
#pragma unroll
                for (int q = 0; q < NPL; ++q) { const int f = lane + 32 * q; if (f < n) acc[q] = fmaf(r[f], x2, acc[q]); }
                sxx = fmaf(x2, x2, sxx);
            }
#pragma unroll
            for (int q = 0; q < NPL; ++q) { const int f = lane + 32 * q; if (f < n) red[warp * ldr + f] = acc[q]; }
            if (lane == 0) red[warp * ldr + n] = sxx;
            __syncthreads();
            float* my_partial = partial + ((size_t)par * G + b) * ldr;
            if (t <= n) {
                double s = 0.0;
                for (int w = 0; w < SW_WARPS; ++w) s += (double)red[w * ldr + t];
                my_partial[t] = (float)s;
            }
            __syncthreads();
            if (t == 0) { __threadfence(); st_release_gpu(flags + (size_t)par * G + b, seq); }
            // ---- grid-wide: wait for every CTA's partial (flag poll = barrier + data-ready in one)
            if (t < G) { while (ld_acquire_gpu(flags + (size_t)par * G + t) != seq) { } }
            __syncthreads();
            if (t < groups * ldr) {
                const int f = t % ldr, g = t / ldr;
                double s = 0.0;
                for (int bb = g; bb < G; bb += groups) s += (double)__ldcg(partial + ((size_t)par * G + bb) * ldr + f);
                red[g * ldr + f] = (float)s;
            }
            __syncthreads();
            if (t <= n) {
                double s = 0.0;
                for (int g = 0; g < groups; ++g) s += (double)red[g * ldr + t];
                svec[t] = (float)s;
            }
            if (t == n + 1) svec[n + 1] = (float)local_count;
            __syncthreads();
            // ---- rank exchange over peer-mapped buffers (only when the signals are sharded)
            if (pc.world > 1) {
                const unsigned cseq = comm_seq0 + seq;
                if (b == 0) {
                    if (t <= n + 1) {
                        const float v = svec[t];
                        for (int r = 0; r < pc.world; ++r) pc.slots[r][((size_t)par * COMM_MAX_RANKS + pc.rank) * COMM_LD + t] = v;
                    }
                    __syncthreads();
                    if (t < pc.world) {
                        __threadfence_system();
                        st_release_sys(pc.flags[t] + par * COMM_MAX_RANKS + pc.rank, cseq);
                    }
                }
                if (t < pc.world) { while (ld_acquire_sys(pc.flags[pc.rank] + par * COMM_MAX_RANKS + t) != cseq) { } }
                __syncthreads();
                if (t <= n + 1) {
                    double s = 0.0;
                    for (int r = 0; r < pc.world; ++r)
                        s += (double)ld_volatile_f32(pc.slots[pc.rank] + ((size_t)par * COMM_MAX_RANKS + r) * COMM_LD + t);
                    svec[t] = (float)s;
                }
                __syncthreads();
                if (svec[n + 1] == 0.f) {                     // no user on any rank (uniform decision)
                    if (b == 0 && t == 0) unused[c] = 1;
                    __syncthreads();
                    continue;
                }
            }
            // ---- new atom                                                        (ksvd.py:118-119)
            if (warp == 0) {
                const float sxx_all = svec[n];
                float sv[NPL], dsq = 0.f;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int f = lane + 32 * q;
                    sv[q] = (f < n) ? fmaf(d_old[f], sxx_all, svec[f]) : 0.f;     // R_k x = R x + d (x.x)
                    dsq = fmaf(sv[q], sv[q], dsq);
                }
                dsq = warp_sum(dsq);
                const float inv = 1.f / (sqrtf(dsq) + kRefEps);                    // utils/math.py:61-62
                float g = 0.f;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int f = lane + 32 * q;
                    if (f < n) {
                        const float dnv = sv[q] * inv;
                        d_new[f] = dnv;
                        g = fmaf(d_old[f], dnv, g);
                        if (b == 0) Dt[(int64_t)c * n + f] = dnv;
                    }
                }
                g = warp_sum(g);
                if (lane == 0) s_g = g;
            }
            __syncthreads();
            const float g = s_g;
            // ---- phase 2: x' = R_k^T d' ; R <- R_k - d' x'                      (ksvd.py:121-123)
            float dn[NPL], dold[NPL];
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const int f = lane + 32 * q;
                dn[q] = (f < n) ? d_new[f] : 0.f;
                dold[q] = (f < n) ? d_old[f] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < SW_U; ++u) {
                if (ent[u] >= 0) {                               // warp-uniform
                    float dot = 0.f;
#pragma unroll
                    for (int q = 0; q < NPL; ++q) dot = fmaf(rv[u][q], dn[q], dot);
                    dot = warp_sum(dot);
                    const float xn = fmaf(x[u], g, dot);
                    float* r = R + (int64_t)(ent[u] / k) * n;
#pragma unroll
                    for (int q = 0; q < NPL; ++q) {
                        const int f = lane + 32 * q;
                        if (f < n) r[f] = fmaf(-dn[q], xn, fmaf(dold[q], x[u], rv[u][q]));
                    }
                    if (lane == 0) val[ent[u]] = xn;
                }
            }
            for (int p = lo + warp + SW_U * SW_WARPS; p < hi; p += SW_WARPS) {
                const int e2 = entries[p];
                const float x2 = val[e2];
                float* r = R + (int64_t)(e2 / k) * n;
                float r2[NPL], dot = 0.f;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int f = lane + 32 * q;
                    r2[q] = (f < n) ? r[f] : 0.f;
                    dot = fmaf(r2[q], dn[q], dot);
                }
                dot = warp_sum(dot);
                const float xn = fmaf(x2, g, dot);
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int f = lane + 32 * q;
                    if (f < n) r[f] = fmaf(-dn[q], xn, fmaf(dold[q], x2, r2[q]));
                }
                __syncwarp();
                if (lane == 0) val[e2] = xn;
            }
            // rows/coefficients of this CTA's signals are re-read by other warps of THIS CTA only
            __syncthreads();
        }
    }
}

template <int NPL>
int launch_sweep(float* R, float* Dt, float* val, const int32_t* rowptr, const int32_t* entries, const int32_t* bounds,
                 int n, int K, int k, int n_cycles, int32_t* unused, float* partial, unsigned* flags,
                 PeerComm pc, unsigned comm_seq0, int grid, cudaStream_t stream)
{
    auto kern = ksvd_sweep_kernel<NPL>;
    size_t smem = sizeof(float) * (size_t)(2 * n + (n + 2) + SW_WARPS * (n + 1));
    LYS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    LYS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SW_WARPS * 32, smem));
    if (per_sm < 1) { set_error("ksvd sweep kernel does not fit on an SM"); return LYS_ECUDA; }
    void* args[] = {&R, &Dt, &val, &rowptr, &entries, &bounds, &n, &K, &k, &n_cycles, &unused, &partial, &flags, &pc, &comm_seq0};
    LYS_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(SW_WARPS * 32), args, smem, stream));
    return LYS_OK;
}

int sweep_grid() { return std::min(sm_count(), 1024); }

}  // namespace
}  // namespace lys

using namespace lys;

extern "C" size_t lys_ksvd_sweep_workspace_bytes(int n, int K)
{
    const size_t G = 1024;     // upper bound on the grid
    return align_up((size_t)n * K * 4, 256) + align_up(2 * G * (size_t)(n + 1) * 4, 256) + align_up(2 * G * 4, 256) +
           align_up((size_t)K * (G + 1) * 4, 256) + 256;
}

extern "C" int lys_approx_ksvd_sweep(float* R, float* D, int64_t ldd, const int32_t* idx, float* val,
                                     const int32_t* rowptr, const int32_t* entries,
                                     int n, int K, int64_t N, int k, int n_cycles,
                                     int32_t* unused, void* comm, void* workspace, size_t workspace_bytes,
                                     void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)idx;
    LYS_CHECK_ARG(R && D && val && rowptr && entries && unused && workspace, "lys_approx_ksvd_sweep: null pointer");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K && k >= 1 && n_cycles >= 1 && N >= 0,
                  "lys_approx_ksvd_sweep: bad shape");
    if (workspace_bytes < lys_ksvd_sweep_workspace_bytes(n, K)) { set_error("lys_approx_ksvd_sweep: workspace too small"); return LYS_EWORKSPACE; }
    PeerComm pc{};
    pc.rank = 0; pc.world = 1;
    unsigned seq0 = 0;
    CommHost* host = reinterpret_cast<CommHost*>(comm);
    if (host) {
        LYS_CHECK_ARG(host->connected, "lys_approx_ksvd_sweep: comm not connected (call lys_comm_connect)");
        LYS_CHECK_ARG(n + 2 <= COMM_LD, "lys_approx_ksvd_sweep: n too large for the exchange slots");
        pc = host->dev;
    }
    const int grid = sweep_grid();
    unsigned char* p = reinterpret_cast<unsigned char*>(workspace);
    float* Dt = reinterpret_cast<float*>(p); p += align_up((size_t)n * K * 4, 256);
    float* partial = reinterpret_cast<float*>(p); p += align_up(2 * (size_t)1024 * (n + 1) * 4, 256);
    unsigned* flags = reinterpret_cast<unsigned*>(p); p += align_up(2 * (size_t)1024 * 4, 256);
    int32_t* bounds = reinterpret_cast<int32_t*>(p);
    LYS_CUDA(cudaMemsetAsync(flags, 0, 2 * (size_t)1024 * 4, stream));
    LYS_CUDA(cudaMemsetAsync(unused, 0, sizeof(int32_t) * (size_t)K, stream));
    int rc = transpose(D, ldd, Dt, n, n, K, stream);
    if (rc) return rc;
    const int64_t S = std::max<int64_t>(1, (N + grid - 1) / grid);
    const int items = K * (grid + 1);
    sweep_bounds_kernel<<<(items + 255) / 256, 256, 0, stream>>>(rowptr, entries, K, k, grid, S, bounds);
    LYS_LAUNCH_CHECK("sweep_bounds_kernel");
    if (host) {
        seq0 = host->epoch;
        host->epoch += (unsigned)(n_cycles * K + 1);
    }
    if (n <= 32) rc = launch_sweep<1>(R, Dt, val, rowptr, entries, bounds, n, K, k, n_cycles, unused, partial, flags, pc, seq0, grid, stream);
    else if (n <= 64) rc = launch_sweep<2>(R, Dt, val, rowptr, entries, bounds, n, K, k, n_cycles, unused, partial, flags, pc, seq0, grid, stream);
    else if (n <= 128) rc = launch_sweep<4>(R, Dt, val, rowptr, entries, bounds, n, K, k, n_cycles, unused, partial, flags, pc, seq0, grid, stream);
    else rc = launch_sweep<8>(R, Dt, val, rowptr, entries, bounds, n, K, k, n_cycles, unused, partial, flags, pc, seq0, grid, stream);
    if (rc) return rc;
    return transpose(Dt, n, D, ldd, K, n, stream);
}
