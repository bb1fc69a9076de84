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

// entry id -> signal index: ent / k by multiply-shift (k <= 32, ent < 2^31; M = ceil(2^(32+s)/k) is exact
// for every 32-bit numerator because M*k - 2^(32+s) < k <= 2^s)
struct FastDiv { unsigned long long M; int s; };
__device__ __forceinline__ int fdiv(int ent, FastDiv d)
{
    return (int)(((unsigned long long)(unsigned)ent * d.M) >> (32 + d.s));
}

constexpr int SW_WARPS = 16;
constexpr int SW_U = 12;         // users per warp whose ids/coefficients/rows stay in registers (192 per CTA)

// ids of this warp's users of one atom (issued a whole atom ahead) ...
__device__ __forceinline__ void load_user_ids(const int32_t* __restrict__ entries, int lo, int hi, int warp,
                                              int (&ent)[SW_U])
{
#pragma unroll
    for (int u = 0; u < SW_U; ++u) {
        const int p = lo + warp + u * SW_WARPS;
        ent[u] = (p < hi) ? __ldg(entries + p) : -1;
    }
}
// ... and their coefficients + an L2 prefetch of their residual rows.  Safe while the previous atom
// is still in flight: an atom's coefficients are only written by that atom's own phase 2, and the
// rows are only PREFETCHED here (they are re-read after the previous atom's phase 2).
__device__ __forceinline__ void load_user_coefs(const float* val, const float* R, int lane, int n, FastDiv k,
                                                const int (&ent)[SW_U], float (&x)[SW_U])
{
#pragma unroll
    for (int u = 0; u < SW_U; ++u) {
        x[u] = 0.f;
        if (ent[u] >= 0) {
            x[u] = val[ent[u]];
            const char* row = reinterpret_cast<const char*>(R + (int64_t)fdiv(ent[u], k) * n);
            if (lane * 128 < n * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + lane * 128));
        }
    }
}

template <int NPL>
__global__ void __launch_bounds__(SW_WARPS * 32, 1)
ksvd_sweep_kernel(float* __restrict__ R, float* __restrict__ Dt, float* __restrict__ val,
                  const int32_t* __restrict__ rowptr, const int32_t* __restrict__ entries,
                  const int32_t* __restrict__ bounds,
                  int n, int K, int k, int n_cycles,
                  int32_t* __restrict__ unused,
                  float* __restrict__ partial /* [2][grid][n+1] */, unsigned* __restrict__ flags /* [2][grid] */,
                  PeerComm pc, unsigned comm_seq0, FastDiv kdiv)
{
    extern __shared__ float sm[];
    float* d_old = sm;                       // [n]
    float* d_new = d_old + n;                // [n]
    float* svec = d_new + n;                 // [n + 2]
    float* red = svec + (n + 2);             // [max(SW_WARPS, groups)][n + 1]
    __shared__ float s_g;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ldr = n + 1;
    const int G = gridDim.x, b = blockIdx.x;
    const int groups = (SW_WARPS * 32) / ldr;
    unsigned seq = 0;

    for (int cyc = 0; cyc < n_cycles; ++cyc) {
        // Software pipeline over atoms: the ids/coefficients of atom c+1's users and this CTA's CSR
        // bounds of atom c+2 are loaded while atom c is in flight, so no dependent address chain
        // (bounds -> ids -> coefficients -> rows) sits on the per-atom critical path.  Every atom runs
        // the same protocol; an atom nobody uses (ksvd.py:112-115) is detected from the summed user
        // count and only skips the refresh (rare, so the wasted sync does not matter).
        int ent[SW_U]; float x[SW_U];
        int lo = bounds[b], hi = bounds[b + 1];
        int lo2 = 0, hi2 = 0;
        if (K > 1) { lo2 = bounds[(G + 1) + b]; hi2 = bounds[(G + 1) + b + 1]; }
        load_user_ids(entries, lo, hi, warp, ent);
        load_user_coefs(val, R, lane, n, kdiv, ent, x);

        for (int c = 0; c < K; ++c) {
            const int local_count = rowptr[c + 1] - rowptr[c];
            ++seq;
            const int par = seq & 1;
            // ---- rows of this CTA's users: issued first (L2 hits after the prefetch)
            float rv[SW_U][NPL];
#pragma unroll
            for (int u = 0; u < SW_U; ++u) {
#pragma unroll
                for (int q = 0; q < NPL; ++q) rv[u][q] = 0.f;
                if (ent[u] >= 0) {
                    const float* r = R + (int64_t)fdiv(ent[u], kdiv) * n;
#pragma unroll
                    for (int q = 0; q < NPL; ++q) { const int f = lane + 32 * q; if (f < n) rv[u][q] = r[f]; }
                }
            }
            for (int f = t; f < n; f += blockDim.x) d_old[f] = Dt[(int64_t)c * n + f];
            // ---- next atom's user ids (its bounds are already in registers), bounds of atom c+2
            int ent2[SW_U]; float x2n[SW_U];
            if (c + 1 < K) {
                load_user_ids(entries, lo2, hi2, warp, ent2);
            } else {
#pragma unroll
                for (int u = 0; u < SW_U; ++u) ent2[u] = -1;
            }
            int lo3 = 0, hi3 = 0;
            if (c + 2 < K) { lo3 = bounds[(c + 2) * (G + 1) + b]; hi3 = bounds[(c + 2) * (G + 1) + b + 1]; }

            // ---- phase 1: partial s = sum_i R[i,:] x_i, sxx = sum x_i^2         (ksvd.py:116-118)
            float acc[NPL];
#pragma unroll
            for (int q = 0; q < NPL; ++q) acc[q] = 0.f;
            float sxx = 0.f;
#pragma unroll
            for (int u = 0; u < SW_U; ++u) {
#pragma unroll
                for (int q = 0; q < NPL; ++q) acc[q] = fmaf(rv[u][q], x[u], acc[q]);
                sxx = fmaf(x[u], x[u], sxx);
            }
            for (int p = lo + warp + SW_U * SW_WARPS; p < hi; p += SW_WARPS) {       // overflow users (rare)
                const int e2 = entries[p];
                const float xo = val[e2];
                const float* r = R + (int64_t)fdiv(e2, kdiv) * n;
#pragma unroll
                for (int q = 0; q < NPL; ++q) { const int f = lane + 32 * q; if (f < n) acc[q] = fmaf(r[f], xo, acc[q]); }
                sxx = fmaf(xo, xo, sxx);
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

            // ---- overlap with the grid converging: coefficients + row prefetch of the next atom
            load_user_coefs(val, R, lane, n, kdiv, ent2, x2n);

            // ---- grid-wide: wait for every CTA's partial (flag poll = barrier + data-ready in one)
            for (int tt = t; tt < G; tt += blockDim.x) { while (ld_acquire_gpu(flags + (size_t)par * G + tt) != seq) { } }
            __syncthreads();
            if (t < groups * ldr) {
                const int f = t % ldr, g = t / ldr;
                const float* src = partial + (size_t)par * G * ldr + f;
                double s = 0.0;
                for (int b0 = g; b0 < G; b0 += 24 * groups) {         // up to 24 independent L2 loads in flight
                    float v[24];
#pragma unroll
                    for (int q = 0; q < 24; ++q) { const int bb = b0 + q * groups; v[q] = (bb < G) ? __ldcg(src + (size_t)bb * ldr) : 0.f; }
#pragma unroll
                    for (int q = 0; q < 24; ++q) s += (double)v[q];     // fixed order: deterministic
                }
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
            }
            const bool atom_used = svec[n + 1] != 0.f;            // uniform over CTAs and ranks
            if (!atom_used && b == 0 && t == 0) unused[c] = 1;
            __syncwarp();      // reconverge warp 0 before the next bar.sync (racecheck finding)
            if (atom_used) {
                // ---- new atom                                                    (ksvd.py:118-119)
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
                // ---- phase 2: x' = R_k^T d' ; R <- R_k - d' x'                  (ksvd.py:121-123)
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
                        float* r = R + (int64_t)fdiv(ent[u], kdiv) * n;
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
                    const float xo = val[e2];
                    float* r = R + (int64_t)fdiv(e2, kdiv) * n;
                    float r2[NPL], dot = 0.f;
#pragma unroll
                    for (int q = 0; q < NPL; ++q) {
                        const int f = lane + 32 * q;
                        r2[q] = (f < n) ? r[f] : 0.f;
                        dot = fmaf(r2[q], dn[q], dot);
                    }
                    dot = warp_sum(dot);
                    const float xn = fmaf(xo, g, dot);
#pragma unroll
                    for (int q = 0; q < NPL; ++q) {
                        const int f = lane + 32 * q;
                        if (f < n) r[f] = fmaf(-dn[q], xn, fmaf(dold[q], xo, r2[q]));
                    }
                    __syncwarp();
                    if (lane == 0) val[e2] = xn;
                }
            }
            // rows/coefficients of this CTA's signals are re-read by other warps of THIS CTA only
            __syncthreads();
            lo = lo2; hi = hi2; lo2 = lo3; hi2 = hi3;
#pragma unroll
            for (int u = 0; u < SW_U; ++u) { ent[u] = ent2[u]; x[u] = x2n[u]; }
        }
    }
}

template <int NPL>
int launch_sweep(float* R, float* Dt, float* val, const int32_t* rowptr, const int32_t* entries, const int32_t* bounds,
                 int n, int K, int k, int n_cycles, int32_t* unused, float* partial, unsigned* flags,
                 PeerComm pc, unsigned comm_seq0, int grid, cudaStream_t stream)
{
    FastDiv kdiv;
    kdiv.s = 0;
    while ((1 << kdiv.s) < k) ++kdiv.s;
    kdiv.M = ((1ull << (32 + kdiv.s)) + (unsigned long long)k - 1) / (unsigned long long)k;
    auto kern = ksvd_sweep_kernel<NPL>;
    const int red_rows = std::max(SW_WARPS, (SW_WARPS * 32) / (n + 1));
    size_t smem = sizeof(float) * (size_t)(2 * n + (n + 2) + red_rows * (n + 1));
    LYS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    LYS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SW_WARPS * 32, smem));
    if (per_sm < 1) { set_error("ksvd sweep kernel does not fit on an SM"); return LYS_ECUDA; }
    void* args[] = {&R, &Dt, &val, &rowptr, &entries, &bounds, &n, &K, &k, &n_cycles, &unused, &partial, &flags, &pc, &comm_seq0, &kdiv};
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
