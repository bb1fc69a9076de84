// bomp_tc3.cu — third generation of the fused Batch-OMP kernel (see bomp_fused.cu for the algebra,
// the fp16 hi/lo split and the CTA-pair decomposition, which are unchanged).
//
// What changed, and why (measured on B200, profiles/README.md): with two tiles in flight per CTA
// the tensor pipe idled half of the time — every greedy step of a tile is a dependent chain
// (MMA -> scan -> gather -> residual -> MMA) of ~20 k cycles of which the MMAs are 6 k, and a
// third tile did not fit: the residual planes (A operand, 32 KB per tile) live next to 128 KB of
// dictionary planes in shared memory.  Here the A operand lives in TENSOR MEMORY instead:
//     TMEM columns   0..191   fp16 hi/lo planes of r for 3 tiles (64 columns each; written with
//                             tcgen05.st by the thread that owns the signal, read by
//                             tcgen05.mma [d], [a_tmem], b_desc — the "TS" form)
//     TMEM columns 256..511   two 128-column fp32 accumulator stages (MMA N = 128 atoms)
// which frees 96 KB of shared memory for the fp32 residuals of the three tiles ([feature][signal],
// conflict-free), so the signal threads keep no 64-float vector in registers across a step.
// Roles (512 threads, registers re-balanced with setmaxnreg 160 / 32):
//     warps  0-11  three tiles ("slots") x 4 warps, thread = signal: scan TMEM, Cholesky,
//                  gather, residual, fp16 planes of r -> TMEM
//     warp  12     TMEM allocation; one thread of the pair's leader CTA issues every MMA
//     warps 13-15  one per slot: zero-fill of the tile's dense code rows with bulk (TMA) stores
// Barriers: a_ready[slot] (4 warps per CTA, at the leader), acc_full / acc_empty[slot][chunk]
// (one per (slot, chunk) so that every waiter sees every phase), tile_begin / zf_done[slot]
// (signal warps <-> zero-fill warp).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace lys {

bool profile_begin(cudaStream_t st, const char* name, cudaEvent_t* stop_out);

namespace {

using namespace tc;

constexpr int TM = 128;                  // signals per tile
constexpr int CH = 128;                  // atoms per MMA / accumulator stage
constexpr int NF = 64;
constexpr int NS = 3;                    // tiles in flight per CTA
constexpr int THREADS = 512;
constexpr int R_SLOT = NF * TM * 4;      // fp32 residuals of one tile, [feature][signal]
constexpr int SMEM_BAR = 576;
constexpr int ZB = 2048;                 // block of zeros for the bulk stores
constexpr int ACC_COL0 = 256;
constexpr float kDictScale = 32.f;

template <int PAIR> struct Geo {
    static constexpr int ROWS_B = CH / PAIR;
    static constexpr int B_PLANE = ROWS_B * NF * 2;
    static constexpr int B_CHUNK = 2 * B_PLANE;
};

template <int PAIR> __host__ __device__ constexpr uint32_t make_idesc()
{
    return (1u << 4) | ((uint32_t)(CH >> 3) << 17) | ((uint32_t)((TM * PAIR) >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]
template <int CG>
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
            "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
            "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

#define LYS_TMEM_ST_X16(taddr, r)                                                                             \
    asm volatile(                                                                                             \
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                       \
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"                            \
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }

template <int N> struct Tree3 {
    static __device__ __forceinline__ float vmax(const float (&a)[N])
    {
        constexpr int M = (N + 2) / 3;
        float t[M];
#pragma unroll
        for (int i = 0; i < N / 3; ++i) t[i] = max3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
        if (N % 3 == 1) t[M - 1] = a[N - 1];
        if (N % 3 == 2) t[M - 1] = fmaxf(a[N - 2], a[N - 1]);
        return Tree3<M>::vmax(t);
    }
    static __device__ __forceinline__ float vmin(const float (&a)[N])
    {
        constexpr int M = (N + 2) / 3;
        float t[M];
#pragma unroll
        for (int i = 0; i < N / 3; ++i) t[i] = min3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
        if (N % 3 == 1) t[M - 1] = a[N - 1];
        if (N % 3 == 2) t[M - 1] = fminf(a[N - 2], a[N - 1]);
        return Tree3<M>::vmin(t);
    }
};
template <> struct Tree3<1> {
    static __device__ __forceinline__ float vmax(const float (&a)[1]) { return a[0]; }
    static __device__ __forceinline__ float vmin(const float (&a)[1]) { return a[0]; }
};

// two-level first-maximum argmax of |alpha| (lyssa/sparse_coding.py:322), see bomp_fused.cu
struct ArgmaxState {
    float run_max;
    int run_piece;
    uint32_t kept[32];
};

__device__ __forceinline__ void scan_piece(const uint32_t (&r)[32], int piece, ArgmaxState& am)
{
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fabsf(__uint_as_float(r[i]));
    const float m = Tree3<32>::vmax(v);
    if (m > am.run_max) {
        am.run_max = m;
        am.run_piece = piece;
#pragma unroll
        for (int i = 0; i < 32; ++i) am.kept[i] = __float_as_uint(2.0f * __uint_as_float(r[i]));   // FMA pipe, exact
    }
}

__device__ __forceinline__ int argmax_finish(const ArgmaxState& am)
{
    float key[32];
    const float m2 = 2.0f * am.run_max;
#pragma unroll
    for (int i = 0; i < 32; ++i) key[i] = fmaf(fabsf(__uint_as_float(am.kept[i])) - m2, -1.0e30f, (float)i);
    return am.run_piece * 32 + (int)Tree3<32>::vmin(key);
}

// r (64 floats of one signal) -> power-of-two scaling, fp16 hi/lo split, rows of the slot's A
// operand in tensor memory: column c of a plane holds K elements (2c, 2c+1).  Warp-collective.
__device__ __forceinline__ void write_planes(const float (&v)[NF], uint32_t taddr_hi)
{
    float t[22];
#pragma unroll
    for (int i = 0; i < 21; ++i) t[i] = max3(fabsf(v[3 * i]), fabsf(v[3 * i + 1]), fabsf(v[3 * i + 2]));
    t[21] = fabsf(v[63]);
    const float amax = Tree3<22>::vmax(t);
    int es = 258 - (int)(__float_as_uint(amax) >> 23);
    es = min(max(es, 1), 254);
    const float s = __uint_as_float((uint32_t)es << 23);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float a = v[32 * half + 2 * c] * s, b = v[32 * half + 2 * c + 1] * s;
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
            hi[c] = *reinterpret_cast<const uint32_t*>(&h);
            lo[c] = *reinterpret_cast<const uint32_t*>(&l);
        }
        LYS_TMEM_ST_X16(taddr_hi + 16 * half, hi);
        LYS_TMEM_ST_X16(taddr_hi + 32 + 16 * half, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <int KNZ> struct SigState {
    float L[KNZ][KNZ];
    float dinv[KNZ], y[KNZ];
    int sel[KNZ];
    int cnt;
    bool done;
};

// step J for one signal after its argmax (:323-359); same algebra as update_step in bomp_fused.cu.
// rs points at this signal's residual in shared memory (stride TM floats); on return (unless the
// signal stopped or this was the last step) d holds the new residual, which is also in rs.
template <int J, int KNZ>
__device__ __forceinline__ void update_step(SigState<KNZ>& st, int pick, bool last, int k,
                                            const float* __restrict__ Dt, const float* __restrict__ G, int K,
                                            float* U, float* rs, float (&d)[NF])
{
    bool dup = false;
#pragma unroll
    for (int m = 0; m < J; ++m) dup |= (st.sel[m] == pick);
    if (dup) { st.done = true; return; }                                 // :323-325
    float g[J > 0 ? J : 1];
#pragma unroll
    for (int m = 0; m < J; ++m) g[m] = __ldg(G + (int64_t)st.sel[m] * K + pick);      // :327
    {
        const float4* dp = reinterpret_cast<const float4*>(Dt + (int64_t)pick * NF);
#pragma unroll
        for (int q = 0; q < NF / 4; ++q) {
            const float4 v4 = __ldg(dp + q);
            d[4 * q] = v4.x; d[4 * q + 1] = v4.y; d[4 * q + 2] = v4.z; d[4 * q + 3] = v4.w;
        }
    }
    float w[J > 0 ? J : 1];
    float ww = 0.f;
#pragma unroll
    for (int m = 0; m < J; ++m) {                                         // :342
        float s = g[m];
#pragma unroll
        for (int c = 0; c < m; ++c) s = fmaf(-st.L[m][c], w[c], s);
        w[m] = s * st.dinv[m];
        ww = fmaf(w[m], w[m], ww);
    }
    const float pivot = 1.f - ww;                                         // :334 / :344
    if (J > 0 && pivot < kPivotEps) { st.done = true; return; }           // :335 / :345
    const float di = (J == 0) ? 1.f : 1.f / sqrtf(pivot);
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
    for (int q = 0; q < NF / 4; ++q) {
        p0 = fmaf(d[4 * q], rs[(4 * q) * TM], p0);         p1 = fmaf(d[4 * q + 1], rs[(4 * q + 1) * TM], p1);
        p2 = fmaf(d[4 * q + 2], rs[(4 * q + 2) * TM], p2); p3 = fmaf(d[4 * q + 3], rs[(4 * q + 3) * TM], p3);
    }
    const float yj = ((p0 + p1) + (p2 + p3)) * di;
#pragma unroll
    for (int m = 0; m < J; ++m) st.L[J][m] = w[m];
    st.dinv[J] = di;
    st.y[J] = yj;
    st.sel[J] = pick;
    st.cnt = J + 1;
    if (last) return;
#pragma unroll
    for (int m = 0; m < J; ++m) {
        const float wm = -w[m];
        const float* um = U + (size_t)m * NF * TM;
#pragma unroll
        for (int f = 0; f < NF; ++f) d[f] = fmaf(wm, __ldcg(um + f * TM), d[f]);
    }
    const bool keep = (J + 2 < k);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const float u = d[f] * di;
        if (keep) __stcg(U + ((size_t)J * NF + f) * TM, u);
        const float rv = fmaf(-yj, u, rs[f * TM]);
        rs[f * TM] = rv;
        d[f] = rv;
    }
}

template <int KNZ, int PAIR>
__global__ void __launch_bounds__(THREADS, 1)
bomp_tc3_kernel(const float* __restrict__ X, int64_t xfs, int64_t xss, int n,
                const uint4* __restrict__ planes, const float* __restrict__ Dt, const float* __restrict__ G,
                int K, int nch, int64_t N, int k, int n_units /* clusters */, int rounds,
                int32_t* __restrict__ idx, float* __restrict__ val, int32_t* __restrict__ nsel,
                float* __restrict__ Z, int64_t zss, float* __restrict__ scratch)
{
    using GE = Geo<PAIR>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;                                                   // nch * B_CHUNK (<= 128 KB)
    float* sR = reinterpret_cast<float*>(smem + (size_t)nch * GE::B_CHUNK);     // NS * 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sR) + NS * R_SLOT);
    // bars[0..2] a_ready, [3..5] tile_begin, [6..8] zf_done, [16 + 8 slot + chunk] acc_full, [40 + 8 slot + chunk] acc_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 64);
    unsigned char* zbuf = reinterpret_cast<unsigned char*>(bars) + SMEM_BAR;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = (PAIR == 2) ? cluster_ctarank() : 0u;
    const int unit = (PAIR == 2) ? (int)cluster_id_x() : (int)blockIdx.x;
    const int64_t n_tiles = (N + TM - 1) / TM;

    if (tid == 0) {
        for (int b = 0; b < NS; ++b) {
            mbar_init(smem_u32(&bars[b]), 4 * PAIR);
            mbar_init(smem_u32(&bars[3 + b]), 4);
            mbar_init(smem_u32(&bars[6 + b]), 1);
        }
        for (int b = 0; b < 8 * NS; ++b) { mbar_init(smem_u32(&bars[16 + b]), 1); mbar_init(smem_u32(&bars[40 + b]), 4 * PAIR); }
        mbar_init_fence();
    }
    if (warp == 12) tmem_alloc<PAIR>(smem_u32(tmem_slot), 512);
    {
        const int items = nch * GE::B_CHUNK / 16;
        const uint4* src = planes + (size_t)rank * items;
        uint4* dst = reinterpret_cast<uint4*>(sB);
        for (int it = tid; it < items; it += THREADS) dst[it] = __ldg(src + it);
    }
    for (int it = tid; it < ZB / 16; it += THREADS) reinterpret_cast<uint4*>(zbuf)[it] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (PAIR == 2) cluster_sync();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t bar_local = smem_u32(&bars[0]);
    const uint32_t bar_lead = mapa(bar_local, 0);

    if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp == 12) {
            // ------------------------------------------------------------------- MMA issuer
            if (rank == 0 && lane == 0) {
                const uint32_t b_base = smem_u32(sB);
                constexpr uint32_t LBO_B = GE::ROWS_B * 16, SBO = 128;
                constexpr uint32_t kIdesc = make_idesc<PAIR>();
                uint32_t u = 0;
                uint32_t prev_bar0 = 0u, prev_bar1 = 0u, prev_par0 = 0u, prev_par1 = 0u;
                for (int r = 0; r < rounds; ++r) {
                    for (int j = 0; j < k; ++j) {
                        const uint32_t q = (uint32_t)(r * k + j);
#pragma unroll 1
                        for (int s = 0; s < NS; ++s) {
                            mbar_wait(bar_local + 8 * s, q & 1);                 // planes of r_j are in TMEM (both CTAs)
                            fence_after();
                            const uint32_t a_hi = tmem_base + 64 * s, a_lo = a_hi + 32;
#pragma unroll 1
                            for (int c = 0; c < nch; ++c, ++u) {
                                const uint32_t stg = u & 1;
                                if (u >= 2) {
                                    mbar_wait(stg ? prev_bar1 : prev_bar0, stg ? prev_par1 : prev_par0);
                                    fence_after();
                                }
                                const uint32_t eb = bar_local + 8 * (40 + 8 * s + c);
                                if (stg) { prev_bar1 = eb; prev_par1 = q & 1; } else { prev_bar0 = eb; prev_par0 = q & 1; }
                                const uint32_t d_tmem = tmem_base + ACC_COL0 + stg * CH;
                                const uint32_t b_hi = b_base + (c * 2) * GE::B_PLANE, b_lo = b_hi + GE::B_PLANE;
                                // small products first: (lo,hi) (hi,lo) (hi,hi)
#pragma unroll
                                for (int ks = 0; ks < NF / 16; ++ks)
                                    mma_f16_ts<PAIR>(d_tmem, a_lo + 8 * ks, make_desc(b_hi + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, ks > 0);
#pragma unroll
                                for (int ks = 0; ks < NF / 16; ++ks)
                                    mma_f16_ts<PAIR>(d_tmem, a_hi + 8 * ks, make_desc(b_lo + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, 1);
#pragma unroll
                                for (int ks = 0; ks < NF / 16; ++ks)
                                    mma_f16_ts<PAIR>(d_tmem, a_hi + 8 * ks, make_desc(b_hi + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, 1);
                                commit<PAIR>(bar_local + 8 * (16 + 8 * s + c));
                            }
                        }
                    }
                }
            }
            __syncwarp();
        } else if (Z) {
            // ------------------------------------------------------------------- zero fill (:308)
            const int zs = warp - 13;
            const uint32_t zsrc = smem_u32(zbuf);
            for (int r = 0; r < rounds; ++r) {
                mbar_wait(bar_local + 8 * (3 + zs), (uint32_t)r & 1);             // the slot has started this tile
                const int64_t tile = (((int64_t)r * n_units + unit) * PAIR + rank) * NS + zs;
                const int64_t sig0 = tile * TM;
                if (tile < n_tiles) {
                    const int64_t rows = (N - sig0 < TM) ? (N - sig0) : TM;
                    if (zss == K) {
                        char* gp = reinterpret_cast<char*>(Z + sig0 * zss);
                        const int64_t total = rows * (int64_t)K * 4;
                        for (int64_t off = (int64_t)lane * ZB; off < total; off += 32 * ZB) {
                            const uint32_t bytes = (uint32_t)((total - off < ZB) ? (total - off) : ZB);
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                         ::"l"(gp + off), "r"(zsrc), "r"(bytes) : "memory");
                        }
                    } else {
                        const int row_bytes = K * 4;
                        for (int64_t rr = lane; rr < rows; rr += 32) {
                            char* gp = reinterpret_cast<char*>(Z + (sig0 + rr) * zss);
                            for (int off = 0; off < row_bytes; off += ZB) {
                                const uint32_t bytes = (uint32_t)((row_bytes - off < ZB) ? (row_bytes - off) : ZB);
                                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                             ::"l"(gp + off), "r"(zsrc), "r"(bytes) : "memory");
                            }
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa(bar_local + 8 * (6 + zs), rank));
            }
        }
    } else {
        // ------------------------------------------------------------- one thread = one signal
        asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
        const int s = warp >> 2;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        float* rs = sR + s * (NF * TM) + row;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t ta_hi = tlane + 64 * s;
        const uint32_t bar_self = mapa(bar_local, rank);
        SigState<KNZ> st;
        const int n_keep = k > 2 ? k - 2 : 0;
        float* U = scratch + ((size_t)(blockIdx.x * NS + s) * n_keep) * NF * TM + row;
        for (int r = 0; r < rounds; ++r) {
            const int64_t tile = (((int64_t)r * n_units + unit) * PAIR + rank) * NS + s;
            const int64_t sig = tile * TM + row;
            const bool live = (tile < n_tiles) && (sig < N);
            {
                float v[NF];
                if (live) {
                    const float* xp = X + sig * xss;
                    if (xfs == 1 && n == NF && ((reinterpret_cast<uintptr_t>(xp) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < NF / 4; ++q) {
                            const float4 v4 = __ldg(reinterpret_cast<const float4*>(xp) + q);
                            v[4 * q] = v4.x; v[4 * q + 1] = v4.y; v[4 * q + 2] = v4.z; v[4 * q + 3] = v4.w;
                        }
                    } else {
#pragma unroll
                        for (int f = 0; f < NF; ++f) v[f] = (f < n) ? __ldg(xp + (int64_t)f * xfs) : 0.f;
                    }
                } else {
#pragma unroll
                    for (int f = 0; f < NF; ++f) v[f] = 0.f;
                }
#pragma unroll
                for (int f = 0; f < NF; ++f) rs[f * TM] = v[f];
                write_planes(v, ta_hi);
            }
            fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(bar_lead + 8 * s);
                if (Z) mbar_arrive_cluster(bar_self + 8 * (3 + s));
            }
            st.cnt = 0;
            st.done = !live;
            for (int j = 0; j < k; ++j) {
                // ---- :322 argmax |alpha_j| over all atoms, first maximum
                ArgmaxState am;
                am.run_max = -1.f;
                am.run_piece = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) am.kept[i] = 0u;
                const uint32_t q = (uint32_t)(r * k + j);
                const uint32_t u0 = (q * NS + s) * (uint32_t)nch;
#pragma unroll 1
                for (int c = 0; c < nch; ++c) {
                    const uint32_t stg = (u0 + c) & 1;
                    mbar_wait(bar_local + 8 * (16 + 8 * s + c), q & 1);
                    fence_after();
                    const uint32_t ta = tlane + ACC_COL0 + stg * CH;
                    uint32_t b0[32], b1[32];
                    LYS_TMEM_LD_X32(ta, b0);
#pragma unroll
                    for (int sc = 0; sc < CH / 32; sc += 2) {
                        LYS_TMEM_WAIT_X32(b0);
                        LYS_TMEM_LD_X32(ta + (sc + 1) * 32, b1);
                        scan_piece(b0, c * (CH / 32) + sc, am);
                        LYS_TMEM_WAIT_X32(b1);
                        if (sc + 2 < CH / 32) LYS_TMEM_LD_X32(ta + (sc + 2) * 32, b0);
                        else {
                            fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(bar_lead + 8 * (40 + 8 * s + c));
                        }
                        scan_piece(b1, c * (CH / 32) + sc + 1, am);
                    }
                }
                const bool last = (j + 1 >= k);
                const int run_idx = argmax_finish(am);
                float d[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) d[f] = 0.f;
                if (!st.done) {
                    switch (j) {
#define LYS_STEP(JJ) case JJ: if constexpr (JJ < KNZ) update_step<JJ, KNZ>(st, run_idx, last, k, Dt, G, K, U, rs, d); break;
                        LYS_STEP(0) LYS_STEP(1) LYS_STEP(2) LYS_STEP(3) LYS_STEP(4)
                        LYS_STEP(5) LYS_STEP(6) LYS_STEP(7) LYS_STEP(8) LYS_STEP(9)
#undef LYS_STEP
                        default: break;
                    }
                }
                if (!last) {
                    // warp-collective: rows of signals that have stopped get whatever d holds (their
                    // correlations are never looked at again)
                    __syncwarp();
                    write_planes(d, ta_hi);
                    fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(bar_lead + 8 * s);
                }
            }
            // ---- :354 z = L^-T y, outputs
            if (Z) mbar_wait(bar_local + 8 * (6 + s), (uint32_t)r & 1);          // dense rows of this tile are zeroed
            if (live) {
                float z[KNZ];
#pragma unroll
                for (int rr = KNZ - 1; rr >= 0; --rr) {
                    if (rr < st.cnt) {
                        float sacc = st.y[rr];
#pragma unroll
                        for (int c = KNZ - 1; c > rr; --c) if (c < st.cnt) sacc = fmaf(-st.L[c][rr], z[c], sacc);
                        z[rr] = sacc * st.dinv[rr];
                    } else {
                        z[rr] = 0.f;
                    }
                }
#pragma unroll
                for (int m = 0; m < KNZ; ++m) {
                    if (m < k) {
                        const bool has = m < st.cnt;
                        idx[sig * k + m] = has ? st.sel[m] : -1;
                        val[sig * k + m] = has ? z[m] : 0.f;
                        if (Z && has) Z[sig * zss + st.sel[m]] = z[m];
                    }
                }
                if (nsel) nsel[sig] = st.cnt;
            }
        }
    }
    fence_before();
    __syncthreads();
    if (PAIR == 2) cluster_sync();
    if (warp == 12) {
        fence_after();
        tmem_dealloc<PAIR>(tmem_base, 512);
    }
}

// D (n <= 64, K) fp32 -> scaled fp16 hi/lo planes in the kernel's shared-memory layout
// [rank][chunk of 128 atoms][plane][k-chunk][row] (16-byte items) and Dt (K, 64) fp32 atom-major
template <int PAIR>
__global__ void prep_dict3_kernel(const float* __restrict__ D, int64_t ldd, int n, int K, int nch,
                                  unsigned char* __restrict__ planes, float* __restrict__ Dt)
{
    using GE = Geo<PAIR>;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= K * (NF / 8)) return;
    const int atom = item % K, kc = item / K;
    const int c = atom / CH, nn = atom % CH, h = nn / GE::ROWS_B, row = nn % GE::ROWS_B;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int f = kc * 8 + e;
        v[e] = (f < n) ? __ldg(D + (int64_t)f * ldd + atom) : 0.f;
        Dt[(int64_t)atom * NF + f] = v[e];
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float a = v[2 * e] * kDictScale, b = v[2 * e + 1] * kDictScale;
        const __half2 hh = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
        hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
        lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    unsigned char* base = planes + (size_t)h * nch * GE::B_CHUNK + (size_t)c * GE::B_CHUNK + kc * (GE::ROWS_B * 16) + row * 16;
    *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + GE::B_PLANE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <int KNZ, int PAIR>
int launch_tc3(const float* X, int64_t xfs, int64_t xss, int n, const void* planes, const float* Dt, const float* G,
               int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zss,
               float* scratch, cudaStream_t stream)
{
    using GE = Geo<PAIR>;
    const int nch = K / CH;
    const size_t smem = (size_t)nch * GE::B_CHUNK + (size_t)NS * R_SLOT + SMEM_BAR + ZB;
    auto kern = bomp_tc3_kernel<KNZ, PAIR>;
    LYS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_tiles = (N + TM - 1) / TM;
    const int64_t tiles_per_unit = (int64_t)PAIR * NS;
    int units = sm_count() / PAIR;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    if (PAIR == 2) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cfg.gridDim = dim3((unsigned)(2 * units), 1, 1);
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0)
            units = std::min(units, max_clusters);
        else
            (void)cudaGetLastError();
    }
    units = (int)std::max<int64_t>(1, std::min<int64_t>(units, (n_tiles + tiles_per_unit - 1) / tiles_per_unit));
    const int rounds = (int)((n_tiles + (int64_t)units * tiles_per_unit - 1) / ((int64_t)units * tiles_per_unit));
    cfg.gridDim = dim3((unsigned)(units * PAIR), 1, 1);
    LYS_CUDA(cudaLaunchKernelEx(&cfg, kern, X, xfs, xss, n, reinterpret_cast<const uint4*>(planes), Dt, G, K, nch, N, k,
                                units, rounds, idx, val, nsel, Z, zss, scratch));
    LYS_LAUNCH_CHECK("bomp_tc3_kernel");
    return LYS_OK;
}

}  // namespace

// same contract as bomp_encode_fused (bomp_fused.cu); the workspace layout is shared:
// planes (K * 256 B) | Dt (K * 256 B) | scratch for the orthonormalised directions
int bomp_encode_tc3(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd, const float* G,
                    int n, int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel,
                    float* Z, int64_t zss, void* planes_ws, float* Dt, float* scratch, cudaStream_t stream)
{
    unsigned char* planes = reinterpret_cast<unsigned char*>(planes_ws);
    const int pair = (K > 4 * CH) ? 2 : 1;
    const int nch = K / CH;
    const int items = K * (NF / 8);
    if (pair == 2) prep_dict3_kernel<2><<<(items + 255) / 256, 256, 0, stream>>>(D, ldd, n, K, nch, planes, Dt);
    else prep_dict3_kernel<1><<<(items + 255) / 256, 256, 0, stream>>>(D, ldd, n, K, nch, planes, Dt);
    LYS_LAUNCH_CHECK("prep_dict3_kernel");
    cudaEvent_t stop_ev;
    const bool prof = profile_begin(stream, "bomp_tc3_kernel", &stop_ev);
    int rc;
    if (k <= 5) {
        rc = (pair == 2) ? launch_tc3<5, 2>(X, xfs, xss, n, planes, Dt, G, K, N, k, idx, val, nsel, Z, zss, scratch, stream)
                         : launch_tc3<5, 1>(X, xfs, xss, n, planes, Dt, G, K, N, k, idx, val, nsel, Z, zss, scratch, stream);
    } else {
        rc = (pair == 2) ? launch_tc3<10, 2>(X, xfs, xss, n, planes, Dt, G, K, N, k, idx, val, nsel, Z, zss, scratch, stream)
                         : launch_tc3<10, 1>(X, xfs, xss, n, planes, Dt, G, K, N, k, idx, val, nsel, Z, zss, scratch, stream);
    }
    if (prof) cudaEventRecord(stop_ev, stream);
    return rc;
}

}  // namespace lys
