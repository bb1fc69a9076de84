// bomp_fast.cu — the tuned Batch-OMP greedy kernel: one warp per signal, alpha in registers,
// orthogonalised Gram columns in shared memory, no per-step triangular solves.
//
// Same selections and coefficients as batch_omp (lyssa/sparse_coding.py:310-365) through the
// algebra of SURVEY.md Appendix C, restated here.  With support I_j = (k_0..k_j), G[I,I] = L L^T
// (unit diagonal assumed, quirk Q1), y = L^-1 alpha0[I] and q_m = column m of G[:,I] L^-T:
//     alpha_j = alpha0 - G[:,I] z = alpha0 - Q y            =>  alpha_j = alpha_{j-1} - y_j q_j
//     new Cholesky row  w = L^-1 G[I,pick]                  =>  w_m = q_m[pick]
//     y_j = (alpha0[pick] - L[j,:j].y[:j]) / L[j,j]         =>  y_j = alpha_{j-1}[pick] / d_j
//     q_j = (G[:,k_j] - sum_{m<j} w_m q_m) / d_j ,  d_j = sqrt(1 - w.w)     (:342-349)
// The kernel stores t_m = d_m q_m (saves K multiplies per step).  Only t_0..t_{k-3} are kept
// (they are needed to form later t's and for the w look-ups); the one missing look-up at the
// last step, w_{k-2}, comes from a single scalar read G[I_{k-2}, pick] and the stored L row —
// exactly the reference's forward substitution (:342).  Coefficients z = L^-T y once per
// signal (:354; the reference re-solves every step but stores only the last z, :365).
// Break rules mirrored: picked atom already selected (:323-325); pivot 1 - w.w < eps (:335,:345).
//
// Roofline that binds it: L2 bandwidth for the k-1 gathered Gram rows (4K bytes each) and
// latency of that dependent chain, not flops: per signal 4K(k-1) B of L2 reads, ~2K(k-1)k/2
// FMAs, 3K(k) compare/select ops.  HBM traffic per signal: 4K (alpha0 read, L2-resident when
// the chunk fits) + 4K (dense Z row write) + 8k (codes).
//
// Register layout: lane l owns atoms {128 v + 4 l + c : v < EPL/4, c < 4} so that alpha rows,
// Gram rows and t-vectors move as coalesced 128-bit accesses (512 B per warp instruction).
#include "common.cuh"
#include <algorithm>

namespace lys {
namespace {

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int EPL, int KNZ> struct FastCfg { static constexpr int kMaxWarps = (EPL >= 64 || KNZ > 5) ? 8 : 16; };

template <int EPL, int KNZ>
__global__ void __launch_bounds__(FastCfg<EPL, KNZ>::kMaxWarps * 32, 1)
bomp_fast_kernel(const float* __restrict__ alpha, const float* __restrict__ G,
                 int64_t C, int k, int warps_per_cta, int tvecs,
                 int32_t* __restrict__ idx, float* __restrict__ val, int32_t* __restrict__ nsel,
                 float* __restrict__ Z, int64_t z_sig_stride)
{
    constexpr int K = EPL * 32;
    constexpr int NV = EPL / 4;
    extern __shared__ __align__(16) float smem_t[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* tv = smem_t + (size_t)warp * tvecs * K;            // tvecs = max(k-2, 0) vectors of K floats
    // Cholesky rows: registers for k <= 5; per-warp shared memory for the k <= 10 instantiation, where
    // 45 extra live scalars per lane cost more (occupancy, spills) than a few broadcast LDS
    constexpr bool kLInSmem = (KNZ > 5);
    float* Ls = smem_t + (size_t)warps_per_cta * tvecs * K + (size_t)warp * KNZ * KNZ;
    const int lane_off = 4 * lane;
    const int64_t warp_global = (int64_t)blockIdx.x * warps_per_cta + warp;
    const int64_t n_warps = (int64_t)gridDim.x * warps_per_cta;

    for (int64_t i = warp_global; i < C; i += n_warps) {
        float a[EPL];
        if (i + n_warps < C) {
            const char* nxt = reinterpret_cast<const char*>(alpha + (i + n_warps) * (int64_t)K);
            for (int b = lane * 128; b < K * 4; b += 32 * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + b));
        }
        {
            const float* arow = alpha + i * (int64_t)K + lane_off;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                float4 x = ldg_f4(arow + 128 * v);
                a[4 * v] = x.x; a[4 * v + 1] = x.y; a[4 * v + 2] = x.z; a[4 * v + 3] = x.w;
            }
        }
        float Lr[kLInSmem ? 1 : KNZ][kLInSmem ? 1 : KNZ];      // L[j][m] = w_m of step j (m < j)
        auto Lget = [&](int r, int c2) -> float { if constexpr (kLInSmem) return Ls[r * KNZ + c2]; else return Lr[r][c2]; };
        auto Lset = [&](int r, int c2, float v2) { if constexpr (kLInSmem) { if (lane == 0) Ls[r * KNZ + c2] = v2; } else Lr[r][c2] = v2; };
        float wrow[KNZ];        // the Cholesky row of the current step (always in registers)
        float dinv[KNZ], y[KNZ];
        int sel[KNZ];
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < KNZ; ++j) {
            if (j >= k) break;
            // ---- :322 argmax |alpha|, FIRST maximum (np.argmax).  Two levels instead of a compare/select
            // scan (3 ALU ops per element made the ALU pipe the busiest unit, 58 % in ncu):
            //   1. max of every group of 4 consecutive atoms and of the lane (3-input FMNMX3 tree),
            //      warp maximum by redux.sync on the bit pattern (|x| >= 0 orders like an integer);
            //   2. lowest group v holding the maximum in each lane, lowest (v, lane) over the warp by
            //      redux.sync.min  ==  lowest atom index block, since index = 128 v + 4 lane + c;
            //   3. warp-uniform switch on v: the owner lane finds the first of its 4 atoms that equals
            //      the maximum.  Exact ties therefore resolve to the lowest atom index, as in the reference.
            float gm[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v)
                gm[v] = fmaxf(fmaxf(fabsf(a[4 * v]), fabsf(a[4 * v + 1])), fmaxf(fabsf(a[4 * v + 2]), fabsf(a[4 * v + 3])));
            float lmax = gm[0];
#pragma unroll
            for (int v = 1; v < NV; ++v) lmax = fmaxf(lmax, gm[v]);
            const unsigned gbits = __reduce_max_sync(0xffffffffu, __float_as_uint(lmax));
            const float gmax = __uint_as_float(gbits);
            int vsel = NV;
#pragma unroll
            for (int v = NV - 1; v >= 0; --v) vsel = (gm[v] == gmax) ? v : vsel;
            const unsigned kmin = __reduce_min_sync(0xffffffffu, vsel < NV ? (unsigned)(vsel * 32 + lane) : 0x7fffffffu);
            const int vstar = (kmin == 0x7fffffffu) ? 0 : (int)(kmin >> 5);      // no hit only if alpha holds NaNs
            const int lstar = (kmin == 0x7fffffffu) ? 0 : (int)(kmin & 31u);
            float e0 = a[0], e1 = a[1], e2 = a[2], e3 = a[3];
            switch (vstar) {
#define LYS_CASE(V) case V: if (V < NV) { e0 = a[4 * (V < NV ? V : 0)]; e1 = a[4 * (V < NV ? V : 0) + 1]; e2 = a[4 * (V < NV ? V : 0) + 2]; e3 = a[4 * (V < NV ? V : 0) + 3]; } break;
                LYS_CASE(1) LYS_CASE(2) LYS_CASE(3) LYS_CASE(4) LYS_CASE(5) LYS_CASE(6) LYS_CASE(7)
                LYS_CASE(8) LYS_CASE(9) LYS_CASE(10) LYS_CASE(11) LYS_CASE(12) LYS_CASE(13) LYS_CASE(14) LYS_CASE(15)
#undef LYS_CASE
                default: break;
            }
            const int cmine = (fabsf(e0) == gmax) ? 0 : (fabsf(e1) == gmax) ? 1 : (fabsf(e2) == gmax) ? 2 : 3;
            const float vmine = (cmine == 0) ? e0 : (cmine == 1) ? e1 : (cmine == 2) ? e2 : e3;
            const int pick = 128 * vstar + 4 * lstar + __shfl_sync(0xffffffffu, cmine, lstar);
            const float apick = __shfl_sync(0xffffffffu, vmine, lstar);          // alpha_{j-1}[pick], signed
            // issue the Gram-row gather NOW (address is always valid): its L2 latency overlaps the
            // scalar Cholesky work below instead of following it
            float4 g[NV];
            if (j + 1 < k) {
                const float* grow = G + (int64_t)pick * K + lane_off;
#pragma unroll
                for (int v = 0; v < NV; ++v) g[v] = ldg_f4(grow + 128 * v);
            }
            // ---- :323-325 already selected -> stop
            bool dup = false;
#pragma unroll
            for (int m = 0; m < KNZ; ++m) if (m < j) dup |= (sel[m] == pick);
            if (dup) break;
            // ---- new Cholesky row: w_m = q_m[pick] = t_m[pick] / d_m; the last one by a scalar G read
            float ww = 0.f;
#pragma unroll
            for (int m = 0; m < KNZ; ++m) {
                if (m < j) {
                    float wm;
                    if (m < tvecs) {
                        wm = tv[m * K + pick] * dinv[m];
                    } else {
                        float s = __ldg(G + (int64_t)sel[m] * K + pick);                 // :327
#pragma unroll
                        for (int c = 0; c < KNZ; ++c) if (c < m) s = fmaf(-Lget(m, c), wrow[c], s);   // :342
                        wm = s * dinv[m];
                    }
                    wrow[m] = wm;
                    Lset(j, m, wm);
                    ww = fmaf(wm, wm, ww);
                }
            }
            const float pivot = 1.f - ww;                                   // :334 / :344
            if (j > 0 && pivot < kPivotEps) break;                          // :335 / :345
            const float d = (j == 0) ? 1.f : sqrtf(pivot);
            const float di = (j == 0) ? 1.f : 1.f / d;
            dinv[j] = di;
            y[j] = apick * di;
            sel[j] = pick;
            cnt = j + 1;
            if (j + 1 >= k) break;                                          // alpha not needed after the last pick
            // ---- t_j = G[:,pick] - sum_m (w_m/d_m) t_m ; alpha -= (y_j/d_j) t_j
            float cm[KNZ];
#pragma unroll
            for (int m = 0; m < KNZ; ++m) if (m < j) cm[m] = wrow[m] * dinv[m];
            const float coef = y[j] * di;
            // m-outer / v-inner in half-row batches: the LDS of one t_m chunk batch are independent and
            // the FMA chains of the NV/2 chunks interleave (the v-outer order serialised on LDS latency:
            // ncu showed short-scoreboard + wait as the top stalls of the k = 10 instantiation)
            constexpr int HB = (NV >= 2) ? NV / 2 : 1;
#pragma unroll
            for (int v0 = 0; v0 < NV; v0 += HB) {
#pragma unroll
                for (int m = 0; m < KNZ; ++m) {
                    if (m < j) {       // m < j <= k-2  =>  m <= k-3 < tvecs: always stored
                        float4 q[HB];
#pragma unroll
                        for (int h = 0; h < HB; ++h)
                            q[h] = *reinterpret_cast<const float4*>(tv + m * K + 128 * (v0 + h) + lane_off);
#pragma unroll
                        for (int h = 0; h < HB; ++h) {
                            float4& t = g[v0 + h];
                            t.x = fmaf(-cm[m], q[h].x, t.x); t.y = fmaf(-cm[m], q[h].y, t.y);
                            t.z = fmaf(-cm[m], q[h].z, t.z); t.w = fmaf(-cm[m], q[h].w, t.w);
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < HB; ++h) {
                    const int v = v0 + h;
                    const float4 t = g[v];
                    if (j < tvecs) *reinterpret_cast<float4*>(tv + j * K + 128 * v + lane_off) = t;
                    a[4 * v] = fmaf(-coef, t.x, a[4 * v]);         a[4 * v + 1] = fmaf(-coef, t.y, a[4 * v + 1]);
                    a[4 * v + 2] = fmaf(-coef, t.z, a[4 * v + 2]); a[4 * v + 3] = fmaf(-coef, t.w, a[4 * v + 3]);
                }
            }
            __syncwarp();      // t_j visible to every lane before the next step's look-ups
        }
        // ---- :354 z = L^-T y
        if constexpr (kLInSmem) __syncwarp();
        float z[KNZ];
#pragma unroll
        for (int r = KNZ - 1; r >= 0; --r) {
            if (r < cnt) {
                float s = y[r];
#pragma unroll
                for (int c = KNZ - 1; c > r; --c) if (c < cnt) s = fmaf(-Lget(c, r), z[c], s);
                z[r] = s * dinv[r];
            } else {
                z[r] = 0.f;
            }
        }
        // ---- outputs: codes, then the dense row (zero fill + scatter, :308,:365)
        {
            int my_i = -1; float my_z = 0.f;
#pragma unroll
            for (int m = 0; m < KNZ; ++m) if (lane == m && m < cnt) { my_i = sel[m]; my_z = z[m]; }
            if (lane < k) { idx[i * k + lane] = my_i; val[i * k + lane] = my_z; }
            if (nsel && lane == 0) nsel[i] = cnt;
            if (Z) {
                float* zrow = Z + i * z_sig_stride + lane_off;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int v = 0; v < NV; ++v) *reinterpret_cast<float4*>(zrow + 128 * v) = zero;
                __syncwarp();
                if (my_i >= 0) Z[i * z_sig_stride + my_i] = my_z;
            }
        }
        __syncwarp();          // the next signal reuses the t-vectors
    }
}

template <int EPL, int KNZ>
int launch_fast(const float* alpha, const float* G, int64_t C, int k,
                int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zss, cudaStream_t stream)
{
    constexpr int K = EPL * 32;
    const int tvecs = k > 2 ? k - 2 : 0;
    const size_t per_warp = (size_t)tvecs * K * sizeof(float);
    int warps = FastCfg<EPL, KNZ>::kMaxWarps;
    const size_t budget = 220 * 1024;
    if (per_warp > 0) warps = (int)std::min<size_t>((size_t)warps, budget / (per_warp + (KNZ > 5 ? KNZ * KNZ * 4 : 0)));
    if (warps < 1) { set_error("bomp_fast: k=%d K=%d needs more shared memory than one SM has", k, K); return LYS_EUNSUPPORTED; }
    // registers: 64K / (warps*32); the kernel is compiled for <= 128 regs at 512 threads
    const size_t smem = per_warp * warps + (KNZ > 5 ? (size_t)warps * KNZ * KNZ * sizeof(float) : 0);
    auto kern = bomp_fast_kernel<EPL, KNZ>;
    LYS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = std::min<int64_t>((C + warps - 1) / warps, (int64_t)sm_count());
    kern<<<(unsigned)blocks, warps * 32, smem, stream>>>(alpha, G, C, k, warps, tvecs, idx, val, nsel, Z, zss);
    LYS_LAUNCH_CHECK("bomp_fast_kernel");
    return LYS_OK;
}

template <int KNZ>
int dispatch_epl(int K, const float* alpha, const float* G, int64_t C, int k,
                 int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zss, cudaStream_t stream)
{
    switch (K) {
        case 128:  return launch_fast<4, KNZ>(alpha, G, C, k, idx, val, nsel, Z, zss, stream);
        case 256:  return launch_fast<8, KNZ>(alpha, G, C, k, idx, val, nsel, Z, zss, stream);
        case 512:  return launch_fast<16, KNZ>(alpha, G, C, k, idx, val, nsel, Z, zss, stream);
        case 1024: return launch_fast<32, KNZ>(alpha, G, C, k, idx, val, nsel, Z, zss, stream);
        case 2048: return launch_fast<64, KNZ>(alpha, G, C, k, idx, val, nsel, Z, zss, stream);
        default:   return LYS_EUNSUPPORTED;
    }
}

}  // namespace

bool bomp_fast_supported(int K, int k, int64_t zas, bool has_Z, const float* Z, int64_t zss)
{
    if (!(K == 128 || K == 256 || K == 512 || K == 1024 || K == 2048)) return false;
    if (k < 1 || k > 10) return false;
    if ((size_t)(k > 2 ? k - 2 : 0) * K * 4 > 220 * 1024) return false;
    if (has_Z && (zas != 1 || (zss % 4) != 0 || (reinterpret_cast<uintptr_t>(Z) & 15) != 0)) return false;
    return true;
}

// greedy phase on a chunk of C signals whose correlations are in `alpha` (C,K), 16-byte aligned
int bomp_greedy_fast(const float* alpha, const float* G, int K, int64_t C, int k,
                     int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zss, cudaStream_t stream)
{
    if (k <= 5) return dispatch_epl<5>(K, alpha, G, C, k, idx, val, nsel, Z, zss, stream);
    return dispatch_epl<10>(K, alpha, G, C, k, idx, val, nsel, Z, zss, stream);
}

}  // namespace lys
