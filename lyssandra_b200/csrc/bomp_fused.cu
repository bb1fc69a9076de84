// bomp_fused.cu — fused Batch-OMP for n <= 64, K in {256,...,1024}: correlations on the 5th-generation
// tensor cores, greedy selection straight out of TMEM, one thread per signal.  Alpha never
// touches HBM and no Gram row is gathered.
//
// Replaces sparse_encoder('bomp') (lyssa/sparse_coding.py:629-635,:708-726) and batch_omp
// (:302-367) for these shapes.  Algebra (same selections and coefficients as the reference):
// batch_omp keeps alpha_j = alpha0 - G[:,I] z_j (:359); with G = D^T D this is D^T r_j for the
// residual r_j = x - D_I z_j, so every greedy step is ONE correlation GEMM of the residual tile
// against the whole dictionary:
//     step j:  alpha_j = D^T r_j                      tcgen05.mma, fp32-faithful (below)
//              pick    = first argmax |alpha_j|       (:322; scanned from TMEM, thread = signal)
//              stop if pick already selected          (:323-325)
//              w = L^-1 G[I,pick], pivot = 1 - w.w    (:327,:342-344; stop if pivot < eps :335,:345)
//              L <- [L; w, sqrt(pivot)]               (:337,:348-349)
//              z = L^-T L^-1 alpha0[I]                (:353-354; alpha0[I] = D_I^T x in fp32)
//              r_{j+1} = x - D_I z
// Precision (default): every fp32 operand is scaled by a power of two and split exactly into two fp16
// planes (hi = rn16(v), lo = rn16(v - hi), 22+ mantissa bits); hi*hi + lo*hi + hi*lo are
// accumulated in fp32 in TMEM (3 MMAs per k-step; the dropped lo*lo term is 2^-22 relative).
// "Screen" mode (LYS_BOMP_SCREEN, A/B option): the tensor cores only RANK.  ONE product hi*hi per k-step; its
// distance to the exact correlation is bounded, per signal and step, by
//     E = ||r - r~|| max_c||d~_c|| + ||r|| max_c||d_c - d~_c||  (+ the accumulation error),
// all four norms measured, none assumed.  The scan keeps the maximum M, the best value outside the winning
// 32-column piece and the piece maxima; if nothing else comes within 2E of M the winner is the exact argmax
// (certified: 96.6 % of the decisions at cfg2).  Otherwise the warp recomputes, cooperatively and in fp32 from the
// atom-major fp32 dictionary, every column of every piece whose maximum is within 2E of M (2.0 pieces on average)
// and takes the first maximum of those exact values.  Everything after the argmax is the same fp32 code in both
// modes, so they return identical codes (tests/test_gpu_encode.py).  Measured: a third of the tensor work but
// 2.21 ms instead of 1.41 ms per 1M patches — the phase timers of the bring-up build show the kernel bound by the
// per-signal chain (scan 5.5 k, update 6 k cycles per warp and step), not by the tensor pipe; the exact
// recomputations add 6.5 k cycles of L2 round trips to that chain.
//
// Decomposition: a tile is 128 signals (TMEM lanes = MMA M); atoms are processed in chunks of
// 256 (MMA N = one 256-column accumulator stage; two stages).  The fp16 planes of the whole
// dictionary must stay in shared memory: 256 B per atom, i.e. 128 KB for 512 atoms.  For
// K > 512 two CTAs of a cluster (one TPC) pair up: tcgen05.mma.cta_group::2 with M = 256
// (128 signals per CTA) reads half of every 256-atom chunk from each CTA's shared memory.
// Every CTA interleaves TWO tiles ("slots") so that the tensor pipe works on one tile while the
// other one is scanned / updated.  Roles (384 threads, registers re-balanced with setmaxnreg):
//     warps 0-3  slot 0: thread = signal; TMEM scan, Cholesky, residual, fp16 planes of r   (232 regs)
//     warps 4-7  slot 1: same
//     warp  8    TMEM allocation; one thread of the pair's leader CTA issues every MMA       (40 regs)
//     warps 9-11 idle (they complete the third warpgroup)
// Barriers: a_ready[slot] (planes of r written; 4 warps per CTA arrive at the leader CTA),
// acc_full[slot][chunk] (tcgen05.commit, multicast to both CTAs), acc_empty[slot][chunk] (4 warps
// per CTA arrive at the leader after their last tcgen05.ld of that accumulator).  One barrier
// per (slot, chunk) — not per TMEM stage — so that every waiter sees EVERY phase of the barrier
// it waits on (the two slots alternate on the two stages; a parity wait must never be two
// phases behind).
//
// Roofline: HBM, 4n + 4K bytes per signal (x in, dense Z row out); tensor work per signal k * 2*64*K*3 fp16 flop
// (1.97 MFLOP at K=1024, k=5; a third of that in screen mode).
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
// atoms per MMA / accumulator stage and number of stages in tensor memory (CH * NSTG = 512 columns).
// Measured at cfg2: 2 x 256 beats 4 x 128 (1.43 vs 1.69 ms sparse): per-unit hand-shakes cost more than
// the deeper pipeline hides.
#ifndef LYS_CH
#define LYS_CH 256
#endif
constexpr int CH = LYS_CH;
constexpr int NSTG = 512 / CH;
constexpr int NF = 64;                   // feature extent of the MMA (n <= 64, zero padded)
constexpr int A_PLANE = TM * NF * 2;     // 16 KB: one fp16 plane of a residual tile
constexpr int A_SLOT = 2 * A_PLANE;      // hi + lo
constexpr int SMEM_BAR = 384;
constexpr int MAX_SLOTS = 3;
constexpr int PM_SLOT = 32 * TM * 4;     // screen mode: maxima of the (<= 32) 32-column pieces of every signal of a tile
constexpr int RBUF = 8 * NF * 4;         // screen mode: one residual per signal warp, for the cooperative exact recompute
constexpr float kAccGamma = 1.52587890625e-05f;   // 2^-16: bound on the relative error of the 64-term fp32 accumulation in TMEM
// dictionary statistics written by dict_stats_kernel: [0] power-of-two scale s (max |s d| in [16,32)),
// [1] max_c ||s d_c - rn16(s d_c)||, [2] max_c ||rn16(s d_c)||

template <int PAIR> struct Geo {
    static constexpr int ROWS_B = CH / PAIR;            // atoms of one chunk held by one CTA
    static constexpr int B_PLANE = ROWS_B * NF * 2;     // bytes of one fp16 plane of one chunk in one CTA
    static constexpr int B_CHUNK = 2 * B_PLANE;
};

// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major
template <int PAIR> __host__ __device__ constexpr uint32_t make_idesc()
{
    return (1u << 4) | ((uint32_t)(CH >> 3) << 17) | ((uint32_t)((TM * PAIR) >> 4) << 24);
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }

// Argmax of |alpha| over all atoms for one signal, FIRST maximum (np.argmax, :322), in two levels.
// Level 1, per 32-column piece read from TMEM: exact maximum of |v| with a 3-input max tree; if it
// beats the running maximum (strictly: earlier pieces win ties) the piece is kept in registers.
// Level 2, once per step on the kept piece: e = |v| - m is 0 exactly where the maximum sits and
// <= -ulp elsewhere, so key = e * (-1e30) + column is the column itself or something huge and
// the minimum key is the first column attaining the maximum.
template <int N> struct Tree3 {           // 3-input reduction trees (FMNMX3)
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

__device__ __forceinline__ float piece_absmax(const uint32_t (&r)[32])
{
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fabsf(__uint_as_float(r[i]));
    return Tree3<32>::vmax(v);
}

// first-maximum argmax state of one signal: the winning 32-column piece is kept in registers as 2*v
// (exact; a multiply runs on the full-rate FMA pipe, a MOV would queue behind the max tree on the
// half-rate ALU pipe: FMNMX3 2.0, FMUL 0.5-1.0 cycles per warp instruction, scripts/ubench/pipes.cu)
struct ArgmaxStateR {
    float run_max;
    float p2;                 // screen mode: largest piece maximum that is not run_max (best value outside the kept piece)
    int run_piece;
    uint32_t kept[32];
};
// SCREEN: also records the piece maximum (pmcol[piece * TM]) and the runner-up among the piece maxima
template <bool SCREEN>
__device__ __forceinline__ void scan_piece_r(const uint32_t (&r)[32], int piece, ArgmaxStateR& am, float* pmcol)
{
    const float m = piece_absmax(r);
    if constexpr (SCREEN) {
        pmcol[piece * TM] = m;
        am.p2 = fmaxf(am.p2, fminf(m, am.run_max));
    }
    if (m > am.run_max) {
        am.run_max = m;
        am.run_piece = piece;
#pragma unroll
        for (int i = 0; i < 32; ++i) am.kept[i] = __float_as_uint(2.0f * __uint_as_float(r[i]));
    }
}
// first column of the kept piece that attains the maximum; SCREEN: *s2 = largest |value| of the piece's OTHER columns
template <bool SCREEN>
__device__ __forceinline__ int argmax_finish_r(const ArgmaxStateR& am, float* s2)
{
    float key[32];
    const float m2 = 2.0f * am.run_max;
#pragma unroll
    for (int i = 0; i < 32; ++i) key[i] = fmaf(fabsf(__uint_as_float(am.kept[i])) - m2, -1.0e30f, (float)i);
    const float first = Tree3<32>::vmin(key);
    if constexpr (SCREEN) {
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = (first == (float)i) ? 0.f : fabsf(__uint_as_float(am.kept[i]));
        *s2 = 0.5f * Tree3<32>::vmax(o);
    }
    return am.run_piece * 32 + (int)first;
}


// scale r by a power of two so that max|r| lands in [16,32), split into fp16 hi/lo planes and
// store row `row` of the slot's A operand (canonical K-major no-swizzle layout: 16-byte chunk
// kc of row r at kc*(TM*16) + r*16).  SCREEN: only the hi plane exists; returns the bound E on
// |tensor-core value - exact correlation| of this residual against any atom, in the accumulator's units
// (both operands scaled): ||rs - rn16(rs)|| max||d~|| + ||rs|| (max||d - d~|| + gamma max||d~||), slightly inflated.
template <bool SCREEN>
__device__ __forceinline__ float store_planes(unsigned char* slotA, int row, const float (&r)[NF], float d_err, float d_max,
                                              float* inv_scale = nullptr)
{
    float t[22];
#pragma unroll
    for (int i = 0; i < 21; ++i) t[i] = max3(fabsf(r[3 * i]), fabsf(r[3 * i + 1]), fabsf(r[3 * i + 2]));
    t[21] = fabsf(r[63]);
    float amax = t[21];
#pragma unroll
    for (int i = 0; i < 21; i += 3) amax = fmaxf(amax, max3(t[i], t[i + 1], t[i + 2]));
    int es = 258 - (int)(__float_as_uint(amax) >> 23);
    es = min(max(es, 1), 254);
    const float s = __uint_as_float((uint32_t)es << 23);
    if (inv_scale) *inv_scale = (es < 254) ? __uint_as_float((uint32_t)(254 - es) << 23) : __uint_as_float(0x00400000u);   // 1/s, exact
    float rho2 = 0.f, drho2 = 0.f;
#pragma unroll
    for (int kc = 0; kc < NF / 8; ++kc) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = r[kc * 8 + 2 * e] * s, b = r[kc * 8 + 2 * e + 1] * s;
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            const float da = a - hf.x, db = b - hf.y;
            hi[e] = *reinterpret_cast<const uint32_t*>(&h);
            if constexpr (SCREEN) {
                rho2 = fmaf(a, a, fmaf(b, b, rho2));
                drho2 = fmaf(da, da, fmaf(db, db, drho2));
            } else {
                const __half2 l = __floats2half2_rn(da, db);
                lo[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
        }
        *reinterpret_cast<uint4*>(slotA + kc * (TM * 16) + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if constexpr (!SCREEN)
            *reinterpret_cast<uint4*>(slotA + A_PLANE + kc * (TM * 16) + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if constexpr (SCREEN)
        return 1.001f * fmaf(sqrtf(drho2), d_max, sqrtf(rho2) * fmaf(kAccGamma, d_max, d_err));
    else
        return 0.f;
}

// 'thresh' mode (lyssa/sparse_coding.py:416-425): running top-KNZ of the SIGNED correlations of one signal,
// descending, ties to the lower column (columns arrive in ascending order and only a strictly larger value
// displaces an entry).
template <int KNZ> struct TopK {
    float t[KNZ];
    int c[KNZ];
};
template <int KNZ> __device__ __forceinline__ void topk_insert(TopK<KNZ>& tk, float v, int col)
{
    tk.t[KNZ - 1] = v;
    tk.c[KNZ - 1] = col;
#pragma unroll
    for (int m = KNZ - 1; m > 0; --m) {
        const bool sw = tk.t[m] > tk.t[m - 1];
        const float a = tk.t[m - 1], b = tk.t[m];
        const int ca = tk.c[m - 1], cb = tk.c[m];
        tk.t[m - 1] = sw ? b : a; tk.t[m] = sw ? a : b;
        tk.c[m - 1] = sw ? cb : ca; tk.c[m] = sw ? ca : cb;
    }
}
// 'thresh' scan in two passes per 256-column chunk while it sits in tensor memory (round 2; the one-pass running
// top-k insertion cost 55 cycles per column and warp: with 32 signals per warp some lane inserts at 40 % of the columns).
// Pass A: the maximum of every (sub-)piece goes into a sorted list of the KNZ largest (sub-)piece maxima; its k-th
// entry t is a lower bound on the k-th largest correlation of the signal (k distinct columns reach it).  Pass B reads
// the chunk again and appends every column with v >= t to the signal's candidate list (values are the fp32-faithful
// tensor-core correlations): ~12 candidates per signal at K = 1024, k = 5, one predicated store pair per column.  The
// list lives in an L2-resident scratch ([entry][signal], only its owner thread touches it); when a lane holds more
// than TCAP - 64 entries the warp prunes its lists to their top KNZ (same routine as the final selection), so any
// input — sorted, constant — stays exact.  Ties: candidates arrive in ascending column order and only a strictly
// larger value displaces an entry, so equal values keep the lower column.
constexpr int TCAP = 96;                 // entries per candidate list (pruned above 32; two pieces = 64 columns between checks)
template <int KNZ> __device__ __forceinline__ void pmx_insert(float (&pmx)[KNZ], float m)
{
    float v = (m == m) ? m : -INFINITY;              // a piece of NaNs must not duplicate an entry
#pragma unroll
    for (int q = 0; q < KNZ; ++q) {
        const float hi = fmaxf(pmx[q], v);
        v = fminf(pmx[q], v);
        pmx[q] = hi;
    }
}
// SUB = 32: one maximum per piece; SUB = 16 (KNZ > 8: the first chunk's 8 pieces would not give k maxima): two
template <int KNZ> __device__ __forceinline__ void scan_piece_max(const uint32_t (&r)[32], float (&pmx)[KNZ])
{
    if constexpr (KNZ > 8) {
        float a[16], b[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[16 + i]); }
        pmx_insert(pmx, Tree3<16>::vmax(a));
        pmx_insert(pmx, Tree3<16>::vmax(b));
    } else {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        pmx_insert(pmx, Tree3<32>::vmax(v));
    }
}
// append the columns of one piece that reach t; off = entries held * TM
__device__ __forceinline__ void scan_piece_collect(const uint32_t (&r)[32], int col0, float t, float* lv, int* lc, uint32_t& off)
{
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float v = __uint_as_float(r[i]);
        if (v >= t) {
            lv[off] = v;
            lc[off] = col0 + i;
            off += TM;
        }
    }
}
// top KNZ of a candidate list, descending, ties to the earlier entry (= lower column)
template <int KNZ> __device__ __forceinline__ void list_topk(TopK<KNZ>& tk, const float* lv, const int* lc, uint32_t off)
{
#pragma unroll
    for (int m = 0; m < KNZ; ++m) { tk.t[m] = -INFINITY; tk.c[m] = -1; }
    const uint32_t mx = __reduce_max_sync(0xffffffffu, off);
#pragma unroll 1
    for (uint32_t e = 0; e < mx; e += 8 * TM) {              // 16 loads in flight per L2 round trip
        float v[8];
        int c[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const bool in = e + q * TM < off;
            v[q] = in ? lv[e + q * TM] : -INFINITY;
            c[q] = in ? lc[e + q * TM] : -1;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (v[q] > tk.t[KNZ - 1]) topk_insert(tk, v[q], c[q]);
    }
}

template <int KNZ> struct SigState {
    float r[NF];                // residual r_j (r_0 = x)
    float L[KNZ][KNZ];          // L[j][m], m < j: Cholesky row of step j (unit diagonal of G assumed, quirk Q1)
    float dinv[KNZ], y[KNZ];
    int sel[KNZ];
    int cnt;
    bool done;
    float E;                    // screen mode: error bound of the tensor-core correlations of the CURRENT residual
};

// everything that follows the argmax of step J for one signal (:323-359).
// With u_m the orthonormalised directions of the selected atoms (u_j = (d_pick - sum_m w_m u_m) / L[j][j],
// w = L^-1 G[I,pick] is the new Cholesky row, :342) the reference's quantities are
//     y_j = (L^-1 alpha0[I])_j = (d_pick . r_j) / L[j][j]        r_{j+1} = r_j - y_j u_j
// so a step needs ONE scattered atom gather (d_pick); u_0..u_{k-3} live in a per-CTA scratch
// laid out [vector][feature][signal] so that a warp's accesses to them are coalesced.
template <int J, int KNZ, bool SCREEN>
__device__ __forceinline__ void update_step(SigState<KNZ>& st, int pick, bool last, int k,
                                            const float* __restrict__ Dt, const float* __restrict__ G, int K,
                                            float* U, unsigned char* slotA, int row, float d_err, float d_max)
{
    bool dup = false;
#pragma unroll
    for (int m = 0; m < J; ++m) dup |= (st.sel[m] == pick);
    if (dup) { st.done = true; return; }                                 // :323-325
    float g[J > 0 ? J : 1];
#pragma unroll
    for (int m = 0; m < J; ++m) g[m] = __ldg(G + (int64_t)st.sel[m] * K + pick);      // :327
    float d[NF];
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
    for (int m = 0; m < J; ++m) {                                         // :342 forward substitution
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
        p0 = fmaf(d[4 * q], st.r[4 * q], p0);         p1 = fmaf(d[4 * q + 1], st.r[4 * q + 1], p1);
        p2 = fmaf(d[4 * q + 2], st.r[4 * q + 2], p2); p3 = fmaf(d[4 * q + 3], st.r[4 * q + 3], p3);
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
    const bool keep = (J + 2 < k);                 // u_J is needed by steps J+1 .. k-2
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const float u = d[f] * di;
        if (keep) __stcg(U + ((size_t)J * NF + f) * TM, u);
        st.r[f] = fmaf(-yj, u, st.r[f]);
    }
    st.E = store_planes<SCREEN>(slotA, row, st.r, d_err, d_max);
}

// The exact first-maximum argmax for the signal of lane `src` when the tensor-core scan could not certify its
// winner: every lane recomputes, in fp32 from the atom-major fp32 dictionary, one column of every 32-column piece
// whose (screened) maximum lies within 2E of the best screened value — all other columns are provably smaller.
// `thr` = M - 2E and `prow` = the signal's row of piece maxima are lane src's; rbuf is this warp's 64-float scratch.
__device__ __forceinline__ int resolve_exact(const float (&r)[NF], int src, float thr, const float* prow, int n_pieces,
                                             const float* __restrict__ Dt, int K, float* rbuf, int lane, int* n_cand = nullptr)
{
    if (lane == src) {
#pragma unroll
        for (int q = 0; q < NF / 4; ++q) reinterpret_cast<float4*>(rbuf)[q] = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    }
    __syncwarp();                                                 // rbuf and the piece maxima are read by the other lanes
    const float thr_s = __shfl_sync(0xffffffffu, thr, src);
    const float* prow_s = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(prow), src));
    const float pmv = (lane < n_pieces) ? prow_s[lane * TM] : -1.f;
    unsigned cand = __ballot_sync(0xffffffffu, pmv >= thr_s);
    if (n_cand) *n_cand += __popc(cand);
    float best = -1.f;
    int bcol = 0x7fffffff;
    while (cand) {
        const int p = __ffs(cand) - 1;
        cand &= cand - 1;
        const int col = p * 32 + lane;
        const float4* dp = reinterpret_cast<const float4*>(Dt + (size_t)col * NF);
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
        for (int q = 0; q < NF / 4; ++q) {
            const float4 d4 = __ldg(dp + q);
            const float4 r4 = reinterpret_cast<const float4*>(rbuf)[q];
            p0 = fmaf(d4.x, r4.x, p0); p1 = fmaf(d4.y, r4.y, p1); p2 = fmaf(d4.z, r4.z, p2); p3 = fmaf(d4.w, r4.w, p3);
        }
        const float v = fabsf((p0 + p1) + (p2 + p3));
        if (v > best) { best = v; bcol = col; }                  // pieces ascend: the first maximum of this lane's columns
    }
    const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(best, 0.f)));
    const unsigned pick = __reduce_min_sync(0xffffffffu, (best >= 0.f && __float_as_uint(best) == mb) ? (unsigned)bcol : 0x7fffffffu);
    __syncwarp();                                                 // rbuf is rewritten by the next unresolved lane
    return (pick < (unsigned)K) ? (int)pick : 0;                  // no candidate at all only if the residual holds NaNs
}

// bring-up instrumentation (LYS_TC_TIMING=1): cycles per role/phase, summed over warps (lane 0)
__device__ unsigned long long g_tc_timing[16];
__device__ unsigned long long g_tc_trace[8192];      // CTA 0: (clock << 8 | warp << 4 | phase) of every lap
__device__ unsigned int g_tc_trace_n;
template <bool ON> struct PhaseTimer {
    long long t;
    __device__ __forceinline__ void start() { if (ON) t = clock64(); }
    __device__ __forceinline__ void lap(int phase, int lane)
    {
        if (ON) {
            const long long now = clock64();
            if (lane == 0) {
                atomicAdd(&g_tc_timing[phase], (unsigned long long)(now - t));
                if (blockIdx.x == 0 && ((threadIdx.x >> 5) & 3) == 0) {
                    const unsigned int i = atomicAdd(&g_tc_trace_n, 1u);
                    if (i < 8192) g_tc_trace[i] = ((unsigned long long)now << 8) | ((threadIdx.x >> 5) << 4) | (unsigned)phase;
                }
            }
            t = now;
        }
    }
};

constexpr int NS = 2;                    // tiles ("slots") interleaved per CTA
constexpr int THREADS = (NS + 1) * 128;
constexpr int ZB = 16384;                // block of zeros, source of the bulk stores that zero-fill dense rows

// MODE 0: Batch-OMP (k greedy steps per tile).  MODE 1: 'thresh' — one correlation pass per tile, the scan keeps
// the k largest signed correlations, coefficients are the exact fp32 dot products with the picked atoms.
// SCREEN (MODE 0 only): one fp16 product ranks, certified or resolved exactly (see the header); only the hi planes
// of the dictionary and of the residuals are staged.
template <int KNZ, int PAIR, bool TIMING, int MODE, bool SCREEN>
__global__ void __launch_bounds__(THREADS, 1)
bomp_tc_kernel(const float* __restrict__ X, int64_t xfs, int64_t xss, int n,
               const uint4* __restrict__ planes, const float* __restrict__ Dt, const float* __restrict__ G,
               const float* __restrict__ dstats,
               int K, int nch, int64_t N, int k, int n_units /* clusters */, int rounds,
               int32_t* __restrict__ idx, float* __restrict__ val, int32_t* __restrict__ nsel,
               float* __restrict__ Z, int64_t zss, float* __restrict__ scratch)
{
    static_assert(!SCREEN || MODE == 0, "screen mode is a Batch-OMP mode");
    using GE = Geo<PAIR>;
    constexpr int NP = CH / 32;
    constexpr int B_STRIDE = SCREEN ? GE::B_PLANE : GE::B_CHUNK;      // bytes of one chunk of the dictionary in this CTA
    constexpr int A_STRIDE = SCREEN ? A_PLANE : A_SLOT;               // bytes of one residual tile
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + (size_t)nch * B_STRIDE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + NS * A_STRIDE);
    // bars[slot] a_ready, [2 + slot] tile_begin, [4 + slot] zf_done, [8 + 8 slot + chunk] acc_full,
    // [24 + 8 slot + chunk] acc_empty; then the TMEM base
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
    uint32_t* tile_ready = tmem_slot + 2;              // [slot] signal warps that have published a tile's codes
    unsigned char* zbuf = reinterpret_cast<unsigned char*>(bars) + SMEM_BAR;
    float* pm_all = reinterpret_cast<float*>(zbuf + NS * ZB);          // [slot][piece][row]       (SCREEN)
    float* rbuf_all = pm_all + NS * (PM_SLOT / 4);                      // [signal warp][feature]   (SCREEN)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = (PAIR == 2) ? cluster_ctarank() : 0u;
    const int unit = (PAIR == 2) ? (int)cluster_id_x() : (int)blockIdx.x;
    const int64_t n_tiles = (N + TM - 1) / TM;
    const int steps = (MODE == 0) ? k : 1;             // correlation passes per tile

    if (tid == 0) {
        tile_ready[0] = 0u; tile_ready[1] = 0u;
        for (int b = 0; b < NS; ++b) {
            mbar_init(smem_u32(&bars[b]), 4 * PAIR);
            mbar_init(smem_u32(&bars[2 + b]), 4);
            mbar_init(smem_u32(&bars[4 + b]), 1);
        }
        for (int b = 0; b < 8 * NS; ++b) { mbar_init(smem_u32(&bars[8 + b]), 1); mbar_init(smem_u32(&bars[24 + b]), 4 * PAIR); }
        mbar_init_fence();
    }
    if (warp == 4 * NS) tmem_alloc<PAIR>(smem_u32(tmem_slot), 512);
    {   // this CTA's share of the dictionary planes: resident for the whole kernel (plain loads + st.shared, not TMA:
        // 64-128 KB once per persistent CTA is not worth a tensor map; the per-tile traffic is the A operand and Z)
        constexpr int per_chunk = B_STRIDE / 16, src_chunk = GE::B_CHUNK / 16;
        const uint4* src = planes + (size_t)rank * nch * src_chunk;
        uint4* dst = reinterpret_cast<uint4*>(sB);
        for (int it = tid; it < nch * per_chunk; it += THREADS) dst[it] = __ldg(src + (it / per_chunk) * src_chunk + (it % per_chunk));
    }
    for (int it = tid; it < NS * ZB / 16; it += THREADS) reinterpret_cast<uint4*>(zbuf)[it] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (PAIR == 2) cluster_sync();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // barrier addresses: waits are local, arrivals of a_ready / acc_empty go to the leader CTA
    const uint32_t bar_local = smem_u32(&bars[0]);
    const uint32_t bar_lead = mapa(bar_local, 0);
    
    if (warp >= 4 * NS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");       // 256 x 232 + 128 x 40 = 384 x 168: the CTA's allocation, no more
        if (warp == 4 * NS) {
            // ------------------------------------------------------------------- MMA issuer
            if (rank == 0 && lane == 0) {
                const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
                constexpr uint32_t LBO_A = TM * 16, LBO_B = GE::ROWS_B * 16, SBO = 128;
                constexpr uint32_t kIdesc = make_idesc<PAIR>();
                uint32_t pb0 = 0u, pb1 = 0u, pb2 = 0u, pb3 = 0u, pp0 = 0u, pp1 = 0u, pp2 = 0u, pp3 = 0u;   // who used each TMEM stage last
                PhaseTimer<TIMING> pt;
                pt.start();
                uint32_t u = 0;
                if constexpr (MODE == 1) {
                    // 'thresh': one pass per tile and the scan, not the tensor pipe, is what a tile waits for.  Each slot
                    // owns SPS of the TMEM stages and the chunks of the two slots are issued alternately, so both
                    // slots scan at the same time (issued slot after slot — as Batch-OMP needs it — the second slot
                    // cannot start before the first one has drained its third chunk: one scanning warp group at a time,
                    // measured 0.89 ms per 1M signals against 0.5x ms with this order)
                    constexpr int SPS = NSTG / NS;
                    for (int r = 0; r < rounds; ++r) {
#pragma unroll 1
                        for (int c = 0; c < nch; ++c) {
#pragma unroll 1
                            for (int s = 0; s < NS; ++s) {
                                if (c == 0) {
                                    mbar_wait(bar_local + 8 * s, (uint32_t)r & 1);            // planes of x landed (both CTAs)
                                    fence_after();
                                }
                                const uint32_t stg = (uint32_t)(s * SPS + (c % SPS));
                                if (r > 0 || c >= SPS) {                                       // stage drained by its last user
                                    const int pc = (c >= SPS) ? c - SPS : nch - SPS + c;
                                    const uint32_t ppar = (c >= SPS) ? ((uint32_t)r & 1) : ((uint32_t)(r - 1) & 1);
                                    mbar_wait(bar_local + 8 * (24 + 8 * s + pc), ppar);
                                    fence_after();
                                }
                                const uint32_t d_tmem = tmem_base + stg * CH;
                                const uint32_t a_hi = a_base + s * A_STRIDE, a_lo = a_hi + A_PLANE;
                                const uint32_t b_hi = b_base + c * B_STRIDE, b_lo = b_hi + GE::B_PLANE;
#pragma unroll
                                for (int ks = 0; ks < NF / 16; ++ks)
                                    mma_f16<PAIR>(d_tmem, make_desc(a_lo + ks * 2 * LBO_A, LBO_A, SBO), make_desc(b_hi + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, ks > 0);
#pragma unroll
                                for (int ks = 0; ks < NF / 16; ++ks)
                                    mma_f16<PAIR>(d_tmem, make_desc(a_hi + ks * 2 * LBO_A, LBO_A, SBO), make_desc(b_lo + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, 1);
#pragma unroll
                                for (int ks = 0; ks < NF / 16; ++ks)
                                    mma_f16<PAIR>(d_tmem, make_desc(a_hi + ks * 2 * LBO_A, LBO_A, SBO), make_desc(b_hi + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, 1);
                                commit<PAIR>(bar_local + 8 * (8 + 8 * s + c));            // accumulator ready
                            }
                        }
                    }
                } else
                for (int r = 0; r < rounds; ++r) {
                    for (int j = 0; j < steps; ++j) {
                        const uint32_t q = (uint32_t)(r * steps + j);
#pragma unroll 1
                        for (int s = 0; s < NS; ++s) {
                            mbar_wait(bar_local + 8 * s, q & 1);                  // planes of r_j landed (both CTAs)
                            fence_after();
                            pt.lap(8, 0);
#pragma unroll 1
                            for (int c = 0; c < nch; ++c, ++u) {
                                const uint32_t stg = u & (NSTG - 1);
                                if (u >= NSTG) {                                       // stage drained (both CTAs)
                                    mbar_wait(stg == 0 ? pb0 : stg == 1 ? pb1 : stg == 2 ? pb2 : pb3,
                                              stg == 0 ? pp0 : stg == 1 ? pp1 : stg == 2 ? pp2 : pp3);
                                    fence_after();
                                }
                                pt.lap(9, 0);
                                const uint32_t eb = bar_local + 8 * (24 + 8 * s + c);
                                if (stg == 0) { pb0 = eb; pp0 = q & 1; } else if (stg == 1) { pb1 = eb; pp1 = q & 1; }
                                else if (stg == 2) { pb2 = eb; pp2 = q & 1; } else { pb3 = eb; pp3 = q & 1; }
                                const uint32_t d_tmem = tmem_base + stg * CH;
                                const uint32_t a_hi = a_base + s * A_STRIDE, a_lo = a_hi + A_PLANE;
                                const uint32_t b_hi = b_base + c * B_STRIDE, b_lo = b_hi + GE::B_PLANE;
                                if constexpr (!SCREEN) {
#pragma unroll
                                    for (int ks = 0; ks < NF / 16; ++ks)
                                        mma_f16<PAIR>(d_tmem, make_desc(a_lo + ks * 2 * LBO_A, LBO_A, SBO), make_desc(b_hi + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, ks > 0);
#pragma unroll
                                    for (int ks = 0; ks < NF / 16; ++ks)
                                        mma_f16<PAIR>(d_tmem, make_desc(a_hi + ks * 2 * LBO_A, LBO_A, SBO), make_desc(b_lo + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, 1);
                                }
#pragma unroll
                                for (int ks = 0; ks < NF / 16; ++ks)
                                    mma_f16<PAIR>(d_tmem, make_desc(a_hi + ks * 2 * LBO_A, LBO_A, SBO), make_desc(b_hi + ks * 2 * LBO_B, LBO_B, SBO), kIdesc, SCREEN ? (ks > 0) : 1);
                                commit<PAIR>(bar_local + 8 * (8 + 8 * s + c));        // accumulator ready
                                pt.lap(10, 0);
                            }
                        }
                    }
                }
            }
            __syncwarp();
        } else if (Z && warp <= 4 * NS + NS) {
            // ------------------------------------------------------- dense code rows (:308, :365)
            // One otherwise idle warp per slot writes the dense rows of a tile AFTER its signal warps have
            // published the sparse codes: rows are composed in shared memory (a block of zeros + the k
            // coefficients of each row) and leave with one bulk (TMA) store per ZROWS rows, so every
            // byte of Z is written exactly once.  (Zero-filling first and scattering afterwards cost
            // 0.5 ms per million signals: the 4-byte scatters hit lines that had already left L2.)
            // The writer works one tile behind the signal warps; a counter in shared memory, not an
            // mbarrier, tracks how many tiles are ready, so it may lag by any number of tiles.
            const int zs = warp - 4 * NS - 1;
            float* zf = reinterpret_cast<float*>(zbuf + zs * ZB);                  // this slot's row block
            volatile uint32_t* ready = tile_ready + zs;
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            // the 16 KB block works as two halves: while the bulk store of one half reads shared memory the other half
            // is composed (one block with a wait after every store kept the writer at ~1.2 k cycles per 4 rows — as
            // long as a whole 'thresh' tile takes the signal warps)
            constexpr int ZH = ZB / 2;
            const int zrows = (zss == K) ? min(ZH / (K * 4), 32 / k) : 1;        // rows per bulk store (<= 32 codes: one per lane)
            // what each half holds from its last store (cleared before the half is composed again)
            int ha = -1, hb = -1;
            // lane c handles code c of a group: row c / k, selection slot c % k
            const int o0 = (lane / k) * K;
            for (int r = 0; r < rounds; ++r) {
                const int64_t tile = (((int64_t)r * n_units + unit) * PAIR + rank) * NS + zs;
                const int64_t sig0 = tile * TM;
                if (tile >= n_tiles) break;
                while (*ready < 4u * (uint32_t)(r + 1)) __nanosleep(200);        // 4 signal warps per tile
                __threadfence();
                const int rows = (int)((N - sig0 < TM) ? (N - sig0) : TM);
                auto fetch = [&](int g, int& fa, float& fv) {
                    const int nr = (rows - g < zrows) ? (rows - g) : zrows;
                    fa = -1;
                    if (g < rows && lane < nr * k) {
                        const int64_t e = (sig0 + g) * k + lane;
                        fa = __ldcg(idx + e); fv = __ldcg(val + e);
                    }
                };
                // one group of rows through one half: clear the half's previous coefficients, set the new ones, store
                auto put = [&](int g, float* zh, int& pa, int a, float v) {
                    const int nr = (rows - g < zrows) ? (rows - g) : zrows;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // this half's last store has read it
                    __syncwarp();
                    if (pa >= 0) zh[o0 + pa] = 0.f;
                    __syncwarp();
                    if (a >= 0) zh[o0 + a] = v;
                    pa = a;
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                                     ::"l"(Z + (sig0 + g) * zss), "r"(smem_u32(zh)), "r"((uint32_t)(nr * K * 4)), "l"(pol) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                };
                int a0, b0;
                float v0 = 0.f, w0 = 0.f;
                fetch(0, a0, v0);
                fetch(zrows, b0, w0);
                for (int g = 0; g < rows; g += 2 * zrows) {
                    int na, nb;
                    float nv = 0.f, nw = 0.f;
                    fetch(g + 2 * zrows, na, nv);                               // two groups ahead: an L2 round trip
                    fetch(g + 3 * zrows, nb, nw);
                    put(g, zf, ha, a0, v0);
                    if (g + zrows < rows) put(g + zrows, zf + ZH / 4, hb, b0, w0);
                    a0 = na; v0 = nv; b0 = nb; w0 = nw;
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        // ------------------------------------------------------------- one thread = one signal
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int s = warp >> 2;                      // slot
        const int quad = warp & 3;                    // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;             // row of the tile
        unsigned char* slotA = sA + s * A_STRIDE;
        const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16);
        float* pmcol = pm_all + s * (PM_SLOT / 4) + row;              // this signal's piece maxima: pmcol[piece * TM]
        float* rbuf = rbuf_all + warp * NF;
        const float d_err = SCREEN ? __ldg(dstats + 1) : 0.f, d_max = SCREEN ? __ldg(dstats + 2) : 0.f;
        SigState<KNZ> st;
        const int n_keep = k > 2 ? k - 2 : 0;
        float* U = scratch + ((size_t)(blockIdx.x * NS + s) * n_keep) * NF * TM + row;
        PhaseTimer<TIMING> pt;
        pt.start();
        auto load_x = [&](int rr, float (&xr)[NF]) {
            const int64_t tile_r = (((int64_t)rr * n_units + unit) * PAIR + rank) * NS + s;
            const int64_t sig_r = tile_r * TM + row;
            if ((tile_r < n_tiles) && (sig_r < N)) {
                const float* xp = X + sig_r * xss;
                if (xfs == 1 && n == NF && ((reinterpret_cast<uintptr_t>(xp) & 15) == 0)) {
#pragma unroll
                    for (int q = 0; q < NF / 4; ++q) {
                        const float4 v4 = __ldg(reinterpret_cast<const float4*>(xp) + q);
                        xr[4 * q] = v4.x; xr[4 * q + 1] = v4.y; xr[4 * q + 2] = v4.z; xr[4 * q + 3] = v4.w;
                    }
                } else {
#pragma unroll
                    for (int f = 0; f < NF; ++f) xr[f] = (f < n) ? __ldg(xp + (int64_t)f * xfs) : 0.f;
                }
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) xr[f] = 0.f;
            }
        };
        for (int r = 0; r < rounds; ++r) {
            const int64_t tile = (((int64_t)r * n_units + unit) * PAIR + rank) * NS + s;
            const int64_t sig = tile * TM + row;
            const bool live = (tile < n_tiles) && (sig < N);
            // ---- load x, publish its planes (:631 alpha0 = D^T x is step 0 of the loop)
            load_x(r, st.r);
            float inv_s = 1.f;
            st.E = store_planes<SCREEN>(slotA, row, st.r, d_err, d_max, MODE == 1 ? &inv_s : nullptr);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar_lead + 8 * s);
            // ('thresh' needs x only for its planes, but loading the next tile's x here, to have it in flight during the
            // scan, keeps 64 more registers live through the scan loops: measured 1.45 ms against 0.60 ms per 1M signals)
            st.cnt = 0;
            st.done = !live;
            pt.lap(0, lane);
            if constexpr (MODE == 1) {
                float* lv = scratch + ((size_t)(blockIdx.x * NS + s) * 2 * TCAP) * TM + row;     // candidate values [entry][row]
                int* lc = reinterpret_cast<int*>(lv + (size_t)TCAP * TM);                        // candidate columns
                uint32_t off = 0u;                                                                // entries held * TM
                float pmx[KNZ];
#pragma unroll
                for (int m = 0; m < KNZ; ++m) pmx[m] = -INFINITY;
                float t_floor = live ? -INFINITY : INFINITY;          // rows past N collect nothing
                TopK<KNZ> tk;
                const uint32_t qq = (uint32_t)r;
#pragma unroll 1
                for (int c = 0; c < nch; ++c) {
                    const uint32_t stg = (uint32_t)(s * (NSTG / NS) + (c % (NSTG / NS)));     // this slot's stages (see the issuer)
                    mbar_wait(bar_local + 8 * (8 + 8 * s + c), qq & 1);
                    fence_after();
                    const uint32_t ta = tq + stg * CH;
                    uint32_t b0[32], b1[32];
                    // ---- pass A: (sub-)piece maxima -> threshold
                    LYS_TMEM_LD_X32(ta, b0);
#pragma unroll 1
                    for (int sc = 0; sc < NP; sc += 2) {
                        LYS_TMEM_WAIT_X32(b0);
                        LYS_TMEM_LD_X32(ta + (sc + 1) * 32, b1);
                        scan_piece_max<KNZ>(b0, pmx);
                        LYS_TMEM_WAIT_X32(b1);
                        LYS_TMEM_LD_X32(ta + ((sc + 2 < NP) ? (sc + 2) * 32 : 0), b0);      // wraps to piece 0 for pass B
                        scan_piece_max<KNZ>(b1, pmx);
                    }
                    float t = pmx[KNZ - 1];
#pragma unroll
                    for (int m = 0; m < KNZ - 1; ++m) t = (m == k - 1) ? pmx[m] : t;
                    t = fmaxf(t, t_floor);
                    // ---- pass B: candidates of this chunk
#pragma unroll 1
                    for (int sc = 0; sc < NP; sc += 2) {
                        LYS_TMEM_WAIT_X32(b0);
                        LYS_TMEM_LD_X32(ta + (sc + 1) * 32, b1);
                        scan_piece_collect(b0, (c * NP + sc) * 32, t, lv, lc, off);
                        LYS_TMEM_WAIT_X32(b1);
                        if (sc + 2 < NP) LYS_TMEM_LD_X32(ta + (sc + 2) * 32, b0);
                        else {
                            fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(bar_lead + 8 * (24 + 8 * s + c));
                        }
                        scan_piece_collect(b1, (c * NP + sc + 1) * 32, t, lv, lc, off);
                        // a lane close to the capacity: the warp prunes its lists to their top KNZ entries
                        if (__any_sync(0xffffffffu, off > (uint32_t)((TCAP - 64) * TM))) {
                            list_topk<KNZ>(tk, lv, lc, off);
                            off = 0u;
#pragma unroll
                            for (int m = 0; m < KNZ; ++m)
                                if (tk.c[m] >= 0) { lv[off] = tk.t[m]; lc[off] = tk.c[m]; off += TM; }
                            float tkk = tk.t[KNZ - 1];
#pragma unroll
                            for (int m = 0; m < KNZ - 1; ++m) tkk = (m == k - 1) ? tk.t[m] : tkk;
                            t_floor = fmaxf(t_floor, tkk);
                            t = fmaxf(t, t_floor);
                        }
                    }
                }
                list_topk<KNZ>(tk, lv, lc, off);
                if (live) {
                    // Z = Alpha on the kept entries (:424).  The accumulator is the fp32-faithful correlation of the two
                    // scaled operands (three fp16 products, 22+ mantissa bits each side, fp32 accumulation: the same
                    // error as an fp32 FMA dot product, tests/test_gpu_gemm.py); both scales are powers of two, so
                    // undoing them is exact.  (Round 1 recomputed the k kept dots from the fp32 atoms: five 256-byte
                    // gathers per signal, 28 % of the kernel's stall samples, and x had to stay in registers.)
                    const float inv_d = 1.f / __ldg(dstats);
#pragma unroll
                    for (int m = 0; m < KNZ; ++m) {
                        if (m < k) {
                            idx[sig * k + m] = tk.c[m];
                            val[sig * k + m] = (tk.c[m] >= 0) ? (tk.t[m] * inv_s) * inv_d : 0.f;
                        }
                    }
                    if (nsel) nsel[sig] = k;
                }
            } else {
                for (int j = 0; j < k; ++j) {
                    // ---- :322 argmax |alpha_j| over all atoms, first maximum
                    const uint32_t qq = (uint32_t)(r * k + j);
                    const uint32_t u0 = (qq * NS + s) * (uint32_t)nch;
                    const bool last = (j + 1 >= k);
                    int run_idx;
                    if constexpr (!SCREEN) {
                        ArgmaxStateR am;
                        am.run_max = -1.f;
                        am.p2 = -1.f;
                        am.run_piece = 0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) am.kept[i] = 0u;
#pragma unroll 1
                        for (int c = 0; c < nch; ++c) {
                            const uint32_t stg = (u0 + c) & (NSTG - 1);
                            mbar_wait(bar_local + 8 * (8 + 8 * s + c), qq & 1);
                            fence_after();
                            pt.lap(2, lane);
                            const uint32_t ta = tq + stg * CH;
                            uint32_t b0[32], b1[32];
                            LYS_TMEM_LD_X32(ta, b0);
#pragma unroll
                            for (int sc = 0; sc < NP; sc += 2) {
                                LYS_TMEM_WAIT_X32(b0);
                                LYS_TMEM_LD_X32(ta + (sc + 1) * 32, b1);
                                scan_piece_r<false>(b0, c * NP + sc, am, pmcol);
                                LYS_TMEM_WAIT_X32(b1);
                                if (sc + 2 < NP) LYS_TMEM_LD_X32(ta + (sc + 2) * 32, b0);
                                else {
                                    fence_before();
                                    __syncwarp();
                                    if (lane == 0) mbar_arrive_cluster(bar_lead + 8 * (24 + 8 * s + c));
                                }
                                scan_piece_r<false>(b1, c * NP + sc + 1, am, pmcol);
                            }
                            pt.lap(3, lane);
                        }
                        float s2unused = 0.f;
                        run_idx = argmax_finish_r<false>(am, &s2unused);
                    } else {
                        ArgmaxStateR am;
                        am.run_max = -1.f;
                        am.p2 = -1.f;
                        am.run_piece = 0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) am.kept[i] = 0u;
#pragma unroll 1
                        for (int c = 0; c < nch; ++c) {
                            const uint32_t stg = (u0 + c) & (NSTG - 1);
                            mbar_wait(bar_local + 8 * (8 + 8 * s + c), qq & 1);
                            fence_after();
                            pt.lap(2, lane);
                            const uint32_t ta = tq + stg * CH;
                            uint32_t b0[32], b1[32];
                            LYS_TMEM_LD_X32(ta, b0);
#pragma unroll
                            for (int sc = 0; sc < NP; sc += 2) {
                                LYS_TMEM_WAIT_X32(b0);
                                LYS_TMEM_LD_X32(ta + (sc + 1) * 32, b1);
                                scan_piece_r<true>(b0, c * NP + sc, am, pmcol);
                                LYS_TMEM_WAIT_X32(b1);
                                if (sc + 2 < NP) LYS_TMEM_LD_X32(ta + (sc + 2) * 32, b0);
                                else {
                                    fence_before();
                                    __syncwarp();
                                    if (lane == 0) mbar_arrive_cluster(bar_lead + 8 * (24 + 8 * s + c));
                                }
                                scan_piece_r<true>(b1, c * NP + sc + 1, am, pmcol);
                            }
                            pt.lap(3, lane);
                        }
                        float s2 = 0.f;
                        run_idx = argmax_finish_r<true>(am, &s2);
                        // certified if nothing else comes within 2E of the winner; otherwise the warp resolves the
                        // signal exactly, one uncertified lane at a time
                        const bool certified = am.run_max - fmaxf(am.p2, s2) > 2.f * st.E;
                        unsigned todo = __ballot_sync(0xffffffffu, !st.done && !certified);
                        const float thr = am.run_max - 2.f * st.E;
                        pt.lap(4, lane);
                        int n_cand = 0;
                        if (TIMING) {
                            const unsigned alive = __ballot_sync(0xffffffffu, !st.done);
                            if (lane == 0) {
                                atomicAdd(&g_tc_timing[11], (unsigned long long)__popc(todo));
                                atomicAdd(&g_tc_timing[13], (unsigned long long)__popc(alive));
                            }
                        }
                        while (todo) {
                            const int src = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const int pick = resolve_exact(st.r, src, thr, pmcol, nch * NP, Dt, K, rbuf, lane, TIMING ? &n_cand : nullptr);
                            if (lane == src) run_idx = pick;
                        }
                        if (TIMING && lane == 0) atomicAdd(&g_tc_timing[12], (unsigned long long)n_cand);
                        pt.lap(6, lane);
                    }
                    if ((unsigned)run_idx >= (unsigned)K) run_idx = 0;       // only with NaNs in the signal (np.argmax would return the first NaN)
                    if (!st.done) {
                        switch (j) {
#define LYS_STEP(JJ) case JJ: if constexpr (JJ < KNZ) update_step<JJ, KNZ, SCREEN>(st, run_idx, last, k, Dt, G, K, U, slotA, row, d_err, d_max); break;
                            LYS_STEP(0) LYS_STEP(1) LYS_STEP(2) LYS_STEP(3) LYS_STEP(4)
                            LYS_STEP(5) LYS_STEP(6) LYS_STEP(7) LYS_STEP(8) LYS_STEP(9)
#undef LYS_STEP
                            default: break;
                        }
                    }
                    if (!last) {
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(bar_lead + 8 * s);
                    }
                    pt.lap(7, lane);
                }
                // ---- :354 z = L^-T y, outputs
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
                        }
                    }
                    if (nsel) nsel[sig] = st.cnt;
                }
            }
            if (Z && tile < n_tiles) {                 // hand the tile to the dense-row writer of this slot
                __syncwarp();
                if (lane == 0) { __threadfence(); atomicAdd(tile_ready + s, 1u); }
            }
            pt.lap(5, lane);
        }
    }
    fence_before();
    __syncthreads();
    if (PAIR == 2) cluster_sync();
    if (warp == 4 * NS) {
        fence_after();
        tmem_dealloc<PAIR>(tmem_base, 512);
    }
}

// Dictionary statistics, one CTA: [0] the power-of-two scale s that puts max |s d| in [16, 32) (so that neither the
// hi nor the lo fp16 plane overflows or goes subnormal whatever the norm of the atoms is), [1] max over atoms of
// ||s d - rn16(s d)||, [2] max over atoms of ||rn16(s d)|| — the two dictionary-side terms of the screen bound E.
__global__ void __launch_bounds__(1024)
dict_stats_kernel(const float* __restrict__ D, int64_t ldd, int n, int K, float* __restrict__ dstats)
{
    __shared__ float red[32];
    __shared__ float s_scale;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float amax = 0.f;
    for (int c = t; c < K; c += 1024)
        for (int f = 0; f < n; ++f) amax = fmaxf(amax, fabsf(__ldg(D + (int64_t)f * ldd + c)));
    auto block_max = [&](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        v = red[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    };
    amax = block_max(amax);
    if (t == 0) {
        int es = 258 - (int)(__float_as_uint(amax) >> 23);           // as store_planes does for a residual
        es = min(max(es, 1), 254);
        s_scale = __uint_as_float((uint32_t)es << 23);
    }
    __syncthreads();
    const float sc = s_scale;
    float emax = 0.f, hmax = 0.f;
    for (int c = t; c < K; c += 1024) {
        float e2 = 0.f, h2 = 0.f;
        for (int f = 0; f < n; ++f) {
            const float a = __ldg(D + (int64_t)f * ldd + c) * sc;
            const float h = __half2float(__float2half_rn(a));
            e2 = fmaf(a - h, a - h, e2);
            h2 = fmaf(h, h, h2);
        }
        emax = fmaxf(emax, e2);
        hmax = fmaxf(hmax, h2);
    }
    emax = block_max(emax);
    hmax = block_max(hmax);
    if (t == 0) {
        dstats[0] = sc;
        dstats[1] = sqrtf(emax) * 1.0001f;
        dstats[2] = sqrtf(hmax) * 1.0001f;
    }
}

// D (n <= 64, K) fp32 -> (a) the scaled fp16 hi/lo planes in the kernel's shared-memory layout,
// per CTA of the pair: [rank][chunk][plane][k-chunk][row] 16-byte items; (b) Dt (K, 64) fp32
// atom-major, zero padded, for the residual / alpha0 gathers.
template <int PAIR>
__global__ void prep_dict_kernel(const float* __restrict__ D, int64_t ldd, int n, int K, int nch,
                                 const float* __restrict__ dstats, unsigned char* __restrict__ planes, float* __restrict__ Dt)
{
    using GE = Geo<PAIR>;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;        // one 16-byte chunk of one atom
    if (item >= K * (NF / 8)) return;
    const float dscale = __ldg(dstats);
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
        const float a = v[2 * e] * dscale, b = v[2 * e + 1] * dscale;
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

bool fused_shape_ok(int n, int K, int k)
{
    return n >= 1 && n <= NF && K >= CH && (K % CH) == 0 && K <= 1024 && k >= 1 && k <= 10;
}

size_t planes_bytes(int K) { return (size_t)K * NF * 2 * 2; }      // hi + lo fp16 of every atom
size_t dt_bytes(int K) { return (size_t)K * NF * sizeof(float); }
// orthonormalised directions u_0..u_{k-3} of every signal in flight: [CTA][slot][vector][feature][signal]
size_t scratch_bytes(int k) { return (size_t)sm_count() * MAX_SLOTS * (k > 2 ? k - 2 : 0) * NF * TM * sizeof(float); }

// 'thresh' candidate lists of every signal in flight: [CTA][slot][values | columns][entry][signal]
size_t thresh_lists_bytes() { return (size_t)sm_count() * MAX_SLOTS * 2 * TCAP * TM * sizeof(float); }

struct FusedWs { unsigned char* planes; float* Dt; float* dstats; float* scratch; };
FusedWs carve_ws(void* workspace, int K)
{
    FusedWs w;
    unsigned char* p = reinterpret_cast<unsigned char*>(workspace);
    w.planes = p; p += align_up(planes_bytes(K), 256);
    w.Dt = reinterpret_cast<float*>(p); p += align_up(dt_bytes(K), 256);
    w.dstats = reinterpret_cast<float*>(p); p += 256;
    w.scratch = reinterpret_cast<float*>(p);
    return w;
}

int prepare_dictionary(const float* D, int64_t ldd, int n, int K, const FusedWs& w, cudaStream_t stream)
{
    const int pair = (K > 512) ? 2 : 1;
    const int nch = K / CH;
    const int items = K * (NF / 8);
    dict_stats_kernel<<<1, 1024, 0, stream>>>(D, ldd, n, K, w.dstats);
    LYS_LAUNCH_CHECK("dict_stats_kernel");
    if (pair == 2) prep_dict_kernel<2><<<(items + 255) / 256, 256, 0, stream>>>(D, ldd, n, K, nch, w.dstats, w.planes, w.Dt);
    else prep_dict_kernel<1><<<(items + 255) / 256, 256, 0, stream>>>(D, ldd, n, K, nch, w.dstats, w.planes, w.Dt);
    LYS_LAUNCH_CHECK("prep_dict_kernel");
    return LYS_OK;
}

template <int KNZ, int PAIR, int MODE, bool SCREEN>
int launch_tc(const float* X, int64_t xfs, int64_t xss, int n, const FusedWs& w, const float* G,
              int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zss,
              cudaStream_t stream)
{
    using GE = Geo<PAIR>;
    const int nch = K / CH;
    const size_t smem = (size_t)nch * (SCREEN ? GE::B_PLANE : GE::B_CHUNK) + (size_t)NS * (SCREEN ? A_PLANE : A_SLOT) + SMEM_BAR + NS * ZB +
                        (SCREEN ? (size_t)NS * PM_SLOT + RBUF : 0);
#ifdef LYS_BRINGUP      // phase timers: only in the bring-up build (python -m lyssandra_b200._build --bringup)
    auto kern = bomp_tc_kernel<KNZ, PAIR, (MODE == 0), MODE, SCREEN>;
#else
    auto kern = bomp_tc_kernel<KNZ, PAIR, false, MODE, SCREEN>;
#endif
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
    LYS_CUDA(cudaLaunchKernelEx(&cfg, kern, X, xfs, xss, n, reinterpret_cast<const uint4*>(w.planes), (const float*)w.Dt, G,
                                (const float*)w.dstats, K, nch, N, k, units, rounds, idx, val, nsel, Z, zss, w.scratch));
    LYS_LAUNCH_CHECK("bomp_tc_kernel");
    return LYS_OK;
}

template <int MODE, bool SCREEN>
int launch_by_shape(const float* X, int64_t xfs, int64_t xss, int n, const FusedWs& w, const float* G,
                    int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zss, cudaStream_t stream)
{
    const bool pair = K > 512;
    if (k <= 5)
        return pair ? launch_tc<5, 2, MODE, SCREEN>(X, xfs, xss, n, w, G, K, N, k, idx, val, nsel, Z, zss, stream)
                    : launch_tc<5, 1, MODE, SCREEN>(X, xfs, xss, n, w, G, K, N, k, idx, val, nsel, Z, zss, stream);
    return pair ? launch_tc<10, 2, MODE, SCREEN>(X, xfs, xss, n, w, G, K, N, k, idx, val, nsel, Z, zss, stream)
                : launch_tc<10, 1, MODE, SCREEN>(X, xfs, xss, n, w, G, K, N, k, idx, val, nsel, Z, zss, stream);
}

}  // namespace

size_t bomp_fused_workspace_bytes(int n, int K, int64_t, int k)
{
    if (!fused_shape_ok(n, K, k)) return 0;
    return align_up(planes_bytes(K), 256) + align_up(dt_bytes(K), 256) + 256 + align_up(scratch_bytes(k), 256) + 256;
}

int bomp_fused_launch_count(int n, int K, int64_t, int k) { return fused_shape_ok(n, K, k) ? 3 : 0; }

// returns LYS_EUNSUPPORTED for shapes this path is not built for (the caller then takes the
// two-kernel path); Z must be signal-major (atom stride 1) here, other layouts are filled by the caller.
// screen != 0 selects the one-product screened correlations (see the header); the codes are the same.
int bomp_encode_fused(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd, const float* G,
                      int n, int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel,
                      float* Z, int64_t zas, int64_t zss, void* workspace, size_t workspace_bytes, int screen, cudaStream_t stream)
{
    if (!fused_shape_ok(n, K, k)) return LYS_EUNSUPPORTED;
    if (Z && (zas != 1 || (zss % 4) != 0 || (reinterpret_cast<uintptr_t>(Z) & 15) != 0)) return LYS_EUNSUPPORTED;
    if (workspace_bytes < bomp_fused_workspace_bytes(n, K, N, k)) return LYS_EWORKSPACE;
    const FusedWs w = carve_ws(workspace, K);
    int rc = prepare_dictionary(D, ldd, n, K, w, stream);
    if (rc) return rc;
    cudaEvent_t stop_ev;
    const bool prof = profile_begin(stream, "bomp_tc_kernel", &stop_ev);
    rc = screen ? launch_by_shape<0, true>(X, xfs, xss, n, w, G, K, N, k, idx, val, nsel, Z, zss, stream)
                : launch_by_shape<0, false>(X, xfs, xss, n, w, G, K, N, k, idx, val, nsel, Z, zss, stream);
    if (prof) cudaEventRecord(stop_ev, stream);
    return rc;
}

size_t thresh_fused_workspace_bytes(int n, int K, int64_t, int k)
{
    if (!fused_shape_ok(n, K, k)) return 0;
    return align_up(planes_bytes(K), 256) + align_up(dt_bytes(K), 256) + 256 + align_up(thresh_lists_bytes(), 256) + 256;
}

// 'thresh' coder through the fused kernel (MODE 1): same shapes and Z layout rules as bomp_encode_fused
int thresh_encode_fused(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd,
                        int n, int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel,
                        float* Z, int64_t zas, int64_t zss, void* workspace, size_t workspace_bytes, cudaStream_t stream)
{
    if (!fused_shape_ok(n, K, k)) return LYS_EUNSUPPORTED;
    if (Z && (zas != 1 || (zss % 4) != 0 || (reinterpret_cast<uintptr_t>(Z) & 15) != 0)) return LYS_EUNSUPPORTED;
    if (workspace_bytes < thresh_fused_workspace_bytes(n, K, N, k)) return LYS_EWORKSPACE;
    const FusedWs w = carve_ws(workspace, K);          // w.scratch: the candidate lists
    const int rc = prepare_dictionary(D, ldd, n, K, w, stream);
    if (rc) return rc;
    return launch_by_shape<1, false>(X, xfs, xss, n, w, nullptr, K, N, k, idx, val, nsel, Z, zss, stream);
}

}  // namespace lys

#ifdef LYS_BRINGUP
// bring-up hooks (not part of the ABI, compiled only with -DLYS_BRINGUP): cycles per phase accumulated by the
// instrumented kernels; reading resets the counters
extern "C" __attribute__((visibility("default"))) int lys_debug_tc_timing(unsigned long long* out16)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, lys::g_tc_timing, sizeof(unsigned long long) * 16) != cudaSuccess) return -2;
    unsigned long long zero[16] = {0};
    if (cudaMemcpyToSymbol(lys::g_tc_timing, zero, sizeof(zero)) != cudaSuccess) return -2;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int lys_debug_tc_trace(unsigned long long* out8192, unsigned int* n)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out8192, lys::g_tc_trace, sizeof(unsigned long long) * 8192) != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(n, lys::g_tc_trace_n, sizeof(unsigned int)) != cudaSuccess) return -2;
    unsigned int z = 0;
    if (cudaMemcpyToSymbol(lys::g_tc_trace_n, &z, sizeof(z)) != cudaSuccess) return -2;
    return 0;
}
#endif  // LYS_BRINGUP
