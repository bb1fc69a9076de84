// ksvd_sweep.cu — the atom loop of approximate K-SVD (lyssa/dict_learning/ksvd.py:105-124) as ONE persistent
// cooperative kernel per cycle, fused with its per-atom reduction across CTAs and, when the signals are sharded over
// GPUs, across ranks (peer-mapped NVLink mailboxes, no host round trip).
//
// Per atom c (sequential; the Gauss-Seidel order is part of the reference's result):
//     users = columns with Z[c,:] != 0                      :111   (CSR built by lys_build_atom_csr)
//     s  = R[:,users] x + d (x.x)          == R_k x         :116-118   (R_k never materialised)
//     d' = s / (||s|| + eps)                                :119
//     x' = R[:,users]^T d' + x (d.d')      == R_k^T d'      :121
//     R[:,users] += d x^T - d' x'^T        == R_k - d' x'   :123
// The only device-wide dependency is the sum s (n floats), x.x and the user count of every atom: 1024 dependent
// all-reduces per sweep at cfg3.  Two things keep that chain short.
//
// (1) Look-ahead by one atom (exact).  The sums of atom c+1 are taken over R BEFORE atom c has been applied, together
//     with the terms that turn them into the sums over the updated R once d'_c is known.  With S = sum_i R_i x_{c+1,i}
//     over the users of c+1, T the same sum restricted to the signals that use BOTH c and c+1, and a = sum over
//     those signals of x_{c,i} x_{c+1,i} (old coefficients):
//         x'_{c,i} = R_i.d'_c + x_{c,i} g,  g = d_c.d'_c          (:121)
//         sum_i R^{new}_i x_{c+1,i} = S + d_c a - d'_c (d'_c.T + g a)
//     so the partial sums of atom c+1 are published before the reduction of atom c has even been read, and the
//     reduction latency (L2 round trips, NVLink hops) overlaps the row updates of the previous atom.  Which users
//     of c+1 also use c is precomputed per code entry (`link`, one byte: the slot of atom c in the same signal).
// (2) The reduction is integer arithmetic.  Every CTA converts its partial sums to fixed point (scale chosen from
//     ||R||_F and ||val||_2 so that no sum can overflow) and adds them into one 64-bit word per element with
//     red.global.add.u64; the low 8 bits of the word count the contributions, so a reader spins on the word itself
//     until all CTAs have added — no flags, no fences, no second round trip — and the result does not depend on the
//     order of the additions: the sweep is bitwise reproducible and every rank computes bit-identical atoms.
//     Residual rows and coefficients are private to the CTA that owns the signal (CTA b owns [b*S, (b+1)*S); the
//     users of an atom are sorted by signal, so its users are one sub-range of the CSR segment, `bounds`), hence
//     nothing else is ever communicated.
// Multi-rank: one warp of every CTA forwards the completed local sums of "its" atoms (atom c -> CTA c mod grid)
// into every peer's mailbox as tagged 8-byte words (comm.cuh); readers add the peers' words to their own sum.
#include "comm.cuh"
#include "sweep_common.cuh"
#include <algorithm>
#include <cmath>

namespace lys {
namespace {

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

// link[i*k + s] = slot of atom (idx[i][s] - 1) in the code of signal i, or 255 (both entries must be users, ksvd.py:111)
__global__ void sweep_link_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, int64_t N, int k,
                                  uint8_t* __restrict__ link)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int32_t* row = idx + i * k;
    const float* vrow = val + i * k;
    for (int s = 0; s < k; ++s) {
        const int a = row[s];
        int lk = 255;
        if (a >= 1 && vrow[s] != 0.f) {
            for (int s2 = k - 1; s2 >= 0; --s2)
                if (row[s2] == a - 1 && vrow[s2] != 0.f) lk = s2;
        }
        link[i * k + s] = (uint8_t)lk;
    }
}

// fixed-point scale of the sweep's sums: every partial sum is bounded by ||val|| max(||R||_F, ||val||) (Cauchy-Schwarz),
// summed over the ranks; fx = 2^e with |sum| fx <= 2^45 leaves 2^10 of headroom in the 55-bit field for the growth of
// the norms during the sweep.  All ranks must use the same scale: the bounds are exchanged through the mailboxes.
__global__ void sweep_scale_kernel(const double* __restrict__ fro /* ||R||^2, ||val||^2 */, double* __restrict__ fx_out,
                                   PeerComm pc, unsigned gen)
{
    const int lane = threadIdx.x;
    const double r = sqrt(fro[0]), v = sqrt(fro[1]);
    float bound = (float)(v * fmax(r, v) * 1.000001);
    if (!(bound >= 0.f)) bound = 3.0e38f;
    if (pc.world > 1) {
        const u64 tag = comm_tag(gen);
        if (lane < pc.world && lane != pc.rank)
            st_relaxed_sys(pc.box[lane] + comm_word(COMM_WINDOW, pc.rank, 0), ((u64)__float_as_uint(bound) << 16) | tag);
        float other = 0.f;
        if (lane < pc.world && lane != pc.rank) {
            const u64* w = pc.box[pc.rank] + comm_word(COMM_WINDOW, lane, 0);
            u64 x;
            do { x = ld_relaxed_sys(w); } while ((x & 0xFFFFull) != tag);
            other = __uint_as_float((unsigned)(x >> 16));
        }
        bound = fmaxf(bound, other);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, o));
        bound *= (float)pc.world;
    }
    if (lane == 0) {
        double fx = 1.0;
        if (bound > 0.f && bound < 3.0e38f) fx = exp2(45.0 - ceil(log2((double)bound)));
        fx_out[0] = fx;
    }
}

// 16 warps per CTA (one CTA per SM, 128 registers per thread).  Multi-rank: warp 15 forwards, 15 warps compute.
constexpr int SW_THREADS = 512;
template <int CW> __device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory"); }

struct SweepArgs {
    float* R; const float* Dt; float* Dt_new; float* val;
    const int32_t* entries; const uint8_t* link; const int32_t* bounds;
    int n, K, k, lda, acc_stride;
    int32_t* unused;
    u64* acc;                     // [K][lda] words, acc_stride words apart, zeroed before the launch
    const double* fx;
    PeerComm pc; unsigned gen0;   // mailbox generation of atoms [0, COMM_WINDOW)
    FastDiv kdiv;
};

// the users of one atom that one warp keeps in registers: lane u holds user u
struct UserSet { int ent; float x; int lk; float xp; };
struct Range { int p0, nu, ovf, hi, cnt; };

template <int CW>
__device__ __forceinline__ Range make_range(int lo, int hi, int warp, int U)
{
    Range r;
    r.cnt = hi - lo;
    const int pw = min(U, (r.cnt + CW - 1) / CW);
    r.p0 = lo + warp * pw;
    r.nu = max(0, min(pw, hi - r.p0));
    r.ovf = lo + CW * pw;              // users beyond CW * U per CTA and atom are streamed, not kept
    r.hi = hi;
    return r;
}

// ids, coefficients and links of a warp's users, fetched in three dependent stages that run in three consecutive
// iterations of the atom loop (each stage's loads are consumed an iteration later, so none of the three L2 round trips
// of the chain entries -> {val, link} -> linked val sits on the per-atom critical path):
//   stage 1 (4 atoms ahead): ent            stage 2 (3 ahead): x = val[ent], lk = link[ent], L2 prefetch of the row
//   stage 3 (2 ahead): xp = val of the previous atom in the same signal (read before that atom's own phase 2 runs:
//   the set of atom c+2 is completed during iteration c, phase 2 of atom c+1 runs in iteration c+1)
__device__ __forceinline__ int set_stage1(const SweepArgs& a, const Range& rg, int lane)
{
    return (lane < rg.nu) ? __ldg(a.entries + rg.p0 + lane) : -1;
}
__device__ __forceinline__ void set_stage2(const SweepArgs& a, int ent, float& x, int& lk)
{
    x = 0.f; lk = 255;
    if (ent >= 0) {
        x = __ldcg(a.val + ent);
        lk = __ldg(a.link + ent);
        const char* row = reinterpret_cast<const char*>(a.R + (int64_t)fdiv(ent, a.kdiv) * a.n);
        for (int off = 0; off < a.n * 4; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
    }
}
__device__ __forceinline__ float set_stage3(const SweepArgs& a, int ent, int lk)
{
    return (ent >= 0 && lk != 255) ? __ldcg(a.val + (int64_t)fdiv(ent, a.kdiv) * a.k + lk) : 0.f;
}
__device__ __forceinline__ UserSet load_set(const SweepArgs& a, const Range& rg, int lane)      // all stages at once (prologue)
{
    UserSet s;
    s.ent = set_stage1(a, rg, lane);
    set_stage2(a, s.ent, s.x, s.lk);
    s.xp = set_stage3(a, s.ent, s.lk);
    return s;
}

// Sum each of the CNT values v[0..CNT) over the 32 lanes with CNT-1 + log2(32/CNT) shuffles instead of 5 CNT: at every
// level the two halves of the value list are exchanged between lane l and lane l ^ OFF.  Returns the total of ONE
// value per lane: value index = multi_owner<U>(lane).
template <int CNT, int OFF>
__device__ __forceinline__ float multi_reduce(float (&v)[CNT], int lane)
{
    if constexpr (OFF == 0) {
        return v[0];
    } else if constexpr (CNT > 1) {
        constexpr int H = CNT / 2;
        float w[H];
        const bool upper = (lane & OFF) != 0;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float send = upper ? v[j] : v[j + H];
            const float keep = upper ? v[j + H] : v[j];
            w[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        return multi_reduce<H, OFF / 2>(w, lane);
    } else {
        float w[1] = {v[0] + __shfl_xor_sync(0xffffffffu, v[0], OFF)};
        return multi_reduce<1, OFF / 2>(w, lane);
    }
}
template <int U> __device__ __forceinline__ int multi_owner(int lane)          // which value a lane ends up with
{
    int u = 0, off = 16;
#pragma unroll
    for (int h = U / 2; h >= 1; h >>= 1, off >>= 1) u += (lane & off) ? h : 0;
    return u;
}
template <int U> __host__ __device__ constexpr int multi_lane(int u)           // lowest lane that holds value u
{
    int l = 0, off = 16;
    for (int h = U / 2; h >= 1; h >>= 1, off >>= 1) if (u & h) l |= off;
    return l;
}

// LYS_ROWMAP = 1: a lane owns NPL consecutive features and moves a row with one 4*NPL-byte access.  Measured SLOWER on the
// same box (6.8 vs 5.9 ms per cfg3 sweep: the stride-NPL feature index costs bank conflicts in the reduction arrays
// and the wider accesses buy nothing on 256-byte rows), so the default is features strided by 32 over the lanes.
#ifndef LYS_ROWMAP
#define LYS_ROWMAP 0
#endif
#if LYS_ROWMAP
#define LYS_FEAT(lane, q) ((lane) * NPL + (q))
#else
#define LYS_FEAT(lane, q) ((lane) + 32 * (q))
#endif
// A lane owns the NPL consecutive features [lane*NPL, lane*NPL + NPL) of a row, so a residual row moves with ONE
// 4*NPL-byte access per lane (a 256-byte row = one LDG.64 / STG.64 per warp at n = 64) whenever n is a multiple of NPL
// (rows are then 4*NPL-byte aligned: R is 16-byte aligned and the row stride is n floats).
template <int NPL>
__device__ __forceinline__ void load_row(const float* __restrict__ r, int n, int lane, float (&v)[NPL])
{
    const int f0 = lane * NPL;
    if (LYS_ROWMAP && (n % NPL) == 0) {
        if (f0 < n) {
            if constexpr (NPL == 1) v[0] = __ldcg(r + f0);
            else if constexpr (NPL == 2) { const float2 t = __ldcg(reinterpret_cast<const float2*>(r + f0)); v[0] = t.x; v[1] = t.y; }
            else {
#pragma unroll
                for (int q = 0; q < NPL; q += 4) {
                    const float4 t = __ldcg(reinterpret_cast<const float4*>(r + f0 + q));
                    v[q] = t.x; v[q + 1] = t.y; v[q + 2] = t.z; v[q + 3] = t.w;
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < NPL; ++q) v[q] = 0.f;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NPL; ++q) { const int f = LYS_FEAT(lane, q); v[q] = (f < n) ? __ldcg(r + f) : 0.f; }
    }
}
template <int NPL>
__device__ __forceinline__ void store_row(float* __restrict__ r, int n, int lane, const float (&v)[NPL])
{
    const int f0 = lane * NPL;
    if (LYS_ROWMAP && (n % NPL) == 0) {
        if (f0 < n) {
            if constexpr (NPL == 1) __stcg(r + f0, v[0]);
            else if constexpr (NPL == 2) __stcg(reinterpret_cast<float2*>(r + f0), make_float2(v[0], v[1]));
            else {
#pragma unroll
                for (int q = 0; q < NPL; q += 4) __stcg(reinterpret_cast<float4*>(r + f0 + q), make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]));
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < NPL; ++q) { const int f = LYS_FEAT(lane, q); if (f < n) __stcg(r + f, v[q]); }
    }
}

template <int NPL, int U>
__device__ __forceinline__ void load_rows(const SweepArgs& a, const UserSet& s, int nu, int lane, float (&rv)[U][NPL])
{
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int e = __shfl_sync(0xffffffffu, s.ent, u);
#pragma unroll
        for (int q = 0; q < NPL; ++q) rv[u][q] = 0.f;
        if (u < nu) load_row<NPL>(a.R + (int64_t)fdiv(e, a.kdiv) * a.n, a.n, lane, rv[u]);
    }
}

// partial sums of one atom over this CTA's users (rows as they are NOW), added into the atom's accumulators
template <int NPL, int U, int CW>
__device__ __forceinline__ void publish_atom(const SweepArgs& a, int c, const UserSet& s, const Range& rg,
                                             const float (&rv)[U][NPL], float* red, double fx, int tid, int lane, int warp)
{
    const int n = a.n, lda = a.lda;
    float S[NPL], T[NPL];
#pragma unroll
    for (int q = 0; q < NPL; ++q) { S[q] = 0.f; T[q] = 0.f; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float x = __shfl_sync(0xffffffffu, s.x, u);
        const bool linked = __shfl_sync(0xffffffffu, s.lk, u) != 255;
        if (u < rg.nu) {
#pragma unroll
            for (int q = 0; q < NPL; ++q) S[q] = fmaf(rv[u][q], x, S[q]);
            if (linked) {
#pragma unroll
                for (int q = 0; q < NPL; ++q) T[q] = fmaf(rv[u][q], x, T[q]);
            }
        }
    }
    float xx = (lane < rg.nu) ? s.x * s.x : 0.f;
    float aa = (lane < rg.nu && s.lk != 255) ? s.xp * s.x : 0.f;
    for (int p = rg.ovf + warp; p < rg.hi; p += CW) {                 // streamed users (only for very popular atoms)
        const int e = __ldg(a.entries + p);
        const float xo = __ldcg(a.val + e);
        const int lk = __ldg(a.link + e);
        const int i = fdiv(e, a.kdiv);
        const float* r = a.R + (int64_t)i * n;
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int f = LYS_FEAT(lane, q);
            const float rvv = (f < n) ? __ldcg(r + f) : 0.f;
            S[q] = fmaf(rvv, xo, S[q]);
            if (lk != 255) T[q] = fmaf(rvv, xo, T[q]);
        }
        if (lane == 0) {
            xx = fmaf(xo, xo, xx);
            if (lk != 255) aa = fmaf(__ldcg(a.val + (int64_t)i * a.k + lk), xo, aa);
        }
    }
    xx = warp_sum(xx);
    aa = warp_sum(aa);
    float* mine = red + warp * lda;
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
        const int f = LYS_FEAT(lane, q);
        if (f < n) { mine[f] = S[q]; mine[n + f] = T[q]; }
    }
    if (lane == 0) { mine[2 * n] = aa; mine[2 * n + 1] = xx; }
    compute_sync<CW>();
    for (int t = tid; t < lda; t += CW * 32) {
        long long q;
        if (t < 2 * n + 2) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < CW; ++w) sum += (double)red[w * lda + t];
            q = __double2ll_rn(sum * fx);
        } else {
            q = (long long)rg.cnt << 8;                                 // user count: exact, survives the 48-bit forward
        }
        red_add(a.acc + ((size_t)c * lda + t) * a.acc_stride, ((u64)q << 8) + 1ull);
    }
}

#ifdef LYS_BRINGUP
// bring-up: cycles per phase of the atom loop, CTA gridDim.x/2, thread 0 (scripts/sweep_phases.py)
__device__ unsigned long long g_sweep_timing[8];
#define LYS_SW_LAP(i) do { if (b == (int)gridDim.x / 2 && tid == 0) { const long long now_ = clock64(); g_sweep_timing[i] += (unsigned long long)(now_ - t_lap); t_lap = now_; } } while (0)
#else
#define LYS_SW_LAP(i) do { } while (0)
#endif

template <int NPL, bool MULTI>
__global__ void __launch_bounds__(SW_THREADS, 1)
ksvd_sweep_kernel(const SweepArgs a)
{
    constexpr int U = (NPL <= 2) ? 16 : 32 / NPL;
    constexpr int CW = MULTI ? 15 : 16;            // compute warps
    extern __shared__ float sm[];
    float* svec = sm;                         // [lda]   the reduced sums of the current atom
    float* red = svec + a.lda;                // [CW][lda]
    __shared__ int s_progress;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, K = a.K, lda = a.lda;
    const int G = gridDim.x, b = blockIdx.x;
    const double fx = a.fx[0];

    if (tid == 0) s_progress = -1;
    __syncthreads();

    if (MULTI && warp == CW) {
        // ---------------------------------------------------------------- forwarder (one warp, atoms c = b, b+G, ...)
        volatile int* progress = &s_progress;
        for (int c = b; c < K; c += G) {
            while (*progress < c - 2) __nanosleep(400);
            const u64 tag = comm_tag(a.gen0 + (unsigned)(c / COMM_WINDOW));
            for (int t = lane; t < lda; t += 32) {
                const u64* w = a.acc + ((size_t)c * lda + t) * a.acc_stride;
                u64 v;
                while (((v = ld_relaxed_gpu(w)) & 255ull) != (u64)G) __nanosleep(40);
                const u64 payload = ((v - (u64)G) & ~0xFFFFull) | tag;      // high 48 bits of the local sum
                for (int r = 0; r < a.pc.world; ++r)
                    if (r != a.pc.rank) st_relaxed_sys(a.pc.box[r] + comm_word(c % COMM_WINDOW, a.pc.rank, t), payload);
            }
        }
        return;
    }

    // -------------------------------------------------------------------------------- compute warps
    const int32_t* bnd = a.bounds + b;
    auto bound_of = [&](int c, int& lo, int& hi) {
        lo = 0; hi = 0;
        if (c < K) { lo = __ldg(bnd + (size_t)c * (G + 1)); hi = __ldg(bnd + (size_t)c * (G + 1) + 1); }
    };
    int lo, hi;
    bound_of(0, lo, hi);
    Range rgC = make_range<CW>(lo, hi, warp, U);
    bound_of(1, lo, hi);
    Range rgN = make_range<CW>(lo, hi, warp, U);
    // the pipeline of user sets: atom c+2 (ids, coefficients, links known), c+3 (ids known), c+4 (bounds known)
    int lo2, hi2, lo3, hi3, lo4, hi4;
    bound_of(2, lo2, hi2);
    bound_of(3, lo3, hi3);
    bound_of(4, lo4, hi4);
    UserSet setC = load_set(a, rgC, lane);
    UserSet setN = load_set(a, rgN, lane);
    UserSet set2;
    set2.ent = set_stage1(a, make_range<CW>(lo2, hi2, warp, U), lane);
    set_stage2(a, set2.ent, set2.x, set2.lk);
    set2.xp = 0.f;
    int ent3 = set_stage1(a, make_range<CW>(lo3, hi3, warp, U), lane);
    float rvA[U][NPL], rvB[U][NPL];
    load_rows<NPL, U>(a, setC, rgC.nu, lane, rvA);
    publish_atom<NPL, U, CW>(a, 0, setC, rgC, rvA, red, fx, tid, lane, warp);
    compute_sync<CW>();                           // `red` is rewritten by the first look-ahead below

    float dold[NPL], doldN[NPL], dnP[NPL], doldP[NPL];
    float gP = 0.f;
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
        const int f = LYS_FEAT(lane, q);
        dold[q] = (f < n) ? __ldg(a.Dt + f) : 0.f;
        dnP[q] = 0.f; doldP[q] = 0.f; doldN[q] = 0.f;
    }
    const int my_user = multi_owner<U>(lane);

#ifdef LYS_BRINGUP
    long long t_lap = clock64();
#endif
    // one atom; the rows of atom c are in `rvC`, those of the look-ahead atom c+1 are loaded into `rvN`.  The caller
    // alternates the two register arrays instead of copying one into the other (64 moves per atom and warp)
    auto atom_step = [&](const int c, float (&rvC)[U][NPL], float (&rvN)[U][NPL]) {
        if (tid == 0) *reinterpret_cast<volatile int*>(&s_progress) = c;
        // ---- rows of the look-ahead atom c+1 first: their latency covers everything up to the sums below
        if (c + 1 < K) load_rows<NPL, U>(a, setN, rgN.nu, lane, rvN);
        // ---- rows of atom c that atom c-1 has just rewritten (its users were loaded before that update)
        {
            const unsigned stale = __ballot_sync(0xffffffffu, lane < rgC.nu && setC.lk != 255);
            if (stale) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if ((stale >> u) & 1u) {
                        const int e = __shfl_sync(0xffffffffu, setC.ent, u);
                        load_row<NPL>(a.R + (int64_t)fdiv(e, a.kdiv) * n, n, lane, rvC[u]);
                    }
                }
            }
        }
        // ---- the three stages of the sets of atoms c+2, c+3, c+4 (consumed at the end of this iteration) and the
        //      CSR bounds of atom c+5
        const float xp2 = set_stage3(a, set2.ent, set2.lk);
        float x3; int lk3;
        set_stage2(a, ent3, x3, lk3);
        const int ent4 = set_stage1(a, make_range<CW>(lo4, hi4, warp, U), lane);
        int lo5, hi5;
        bound_of(c + 5, lo5, hi5);
        LYS_SW_LAP(0);
        if (c + 1 < K) {
#pragma unroll
            for (int q = 0; q < NPL; ++q) { const int f = LYS_FEAT(lane, q); doldN[q] = (f < n) ? __ldg(a.Dt + (size_t)(c + 1) * n + f) : 0.f; }
            // ---- look-ahead: sums of atom c+1 over the rows as they are before atom c is applied
            publish_atom<NPL, U, CW>(a, c + 1, setN, rgN, rvN, red, fx, tid, lane, warp);
        }
        LYS_SW_LAP(1);
        // ---- reduced sums of atom c (published one iteration ago)
        for (int t = tid; t < lda; t += CW * 32) {
            const u64* w = a.acc + ((size_t)c * lda + t) * a.acc_stride;
            u64 v;
            do { v = ld_relaxed_gpu(w); } while ((v & 255ull) != (u64)G);
            long long q;
            if (!MULTI) {
                q = (long long)(v - (u64)G) >> 8;
            } else {
                q = (long long)(v - (u64)G) >> 16;
                const u64 tag = comm_tag(a.gen0 + (unsigned)(c / COMM_WINDOW));
                for (int r = 0; r < a.pc.world; ++r) {
                    if (r == a.pc.rank) continue;
                    const u64* m = a.pc.box[a.pc.rank] + comm_word(c % COMM_WINDOW, r, t);
                    u64 x;
                    do { x = ld_relaxed_sys(m); } while ((x & 0xFFFFull) != tag);
                    q += (long long)x >> 16;
                }
            }
            const double scale = (t < 2 * n + 2) ? (MULTI ? 256.0 : 1.0) / fx : (MULTI ? 1.0 : 1.0 / 256.0);
            svec[t] = (float)((double)q * scale);
        }
        LYS_SW_LAP(2);
        compute_sync<CW>();
        LYS_SW_LAP(3);
        // ---- new atom (every warp computes it redundantly: no second barrier)          (ksvd.py:118-119)
        const float aa = svec[2 * n], sxx = svec[2 * n + 1];
        const bool used = svec[2 * n + 2] != 0.f;                     // uniform over CTAs and ranks
        float dn[NPL], sv[NPL], tv[NPL];
        float dT = 0.f;
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int f = LYS_FEAT(lane, q);
            sv[q] = (f < n) ? svec[f] : 0.f;
            tv[q] = (f < n) ? svec[n + f] : 0.f;
            dT = fmaf(dnP[q], tv[q], dT);
        }
        dT = warp_sum(dT);
        const float coef = fmaf(gP, aa, dT);                          // sum over shared users of x'_{c-1,i} x_{c,i}
        float dsq = 0.f, dsv = 0.f;
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            sv[q] = fmaf(dold[q], sxx, fmaf(-dnP[q], coef, fmaf(doldP[q], aa, sv[q])));      // R_k x = R x + d (x.x)
            dsq = fmaf(sv[q], sv[q], dsq);
            dsv = fmaf(dold[q], sv[q], dsv);
        }
        {   // ||s||^2 and d.s in one exchange: lanes < 16 carry one sum, lanes >= 16 the other
            float two[2] = {dsq, dsv};
            const float mine = multi_reduce<2, 16>(two, lane);
            dsq = __shfl_sync(0xffffffffu, mine, 0);
            dsv = __shfl_sync(0xffffffffu, mine, 16);
        }
        const float inv = 1.f / (sqrtf(dsq) + kRefEps);              // utils/math.py:61-62
        float g = 0.f;
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            dn[q] = used ? sv[q] * inv : dold[q];                    // an atom nobody uses is left alone (:112-115)
            g = fmaf(dold[q], dold[q], g);
        }
        g = used ? dsv * inv : warp_sum(g);                          // g = d.d'
        if (b == 0 && warp == 0) {
#pragma unroll
            for (int q = 0; q < NPL; ++q) { const int f = LYS_FEAT(lane, q); if (f < n) a.Dt_new[(size_t)c * n + f] = dn[q]; }
            if (!used && lane == 0) a.unused[c] = 1;
        }
        LYS_SW_LAP(4);
        // ---- phase 2: x' = R_k^T d' ; R <- R_k - d' x'                                  (ksvd.py:121-123)
        if (used) {
            // all dot products of the warp's users in one multi-value reduction; lane multi_lane(u) then owns user u
            float dots[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float d = 0.f;
#pragma unroll
                for (int q = 0; q < NPL; ++q) d = fmaf(rvC[u][q], dn[q], d);
                dots[u] = d;
            }
            const float dot_mine = multi_reduce<U, 16>(dots, lane);
            const float x_mine = __shfl_sync(0xffffffffu, setC.x, my_user);
            const int ent_mine = __shfl_sync(0xffffffffu, setC.ent, my_user);
            const float xn_mine = fmaf(x_mine, g, dot_mine);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float x = __shfl_sync(0xffffffffu, setC.x, u);
                const int e = __shfl_sync(0xffffffffu, setC.ent, u);
                const float xn = __shfl_sync(0xffffffffu, xn_mine, multi_lane<U>(u));
                if (u < rgC.nu) {
                    float o[NPL];
#pragma unroll
                    for (int q = 0; q < NPL; ++q) o[q] = fmaf(-dn[q], xn, fmaf(dold[q], x, rvC[u][q]));
                    store_row<NPL>(a.R + (int64_t)fdiv(e, a.kdiv) * n, n, lane, o);
                }
            }
            if (my_user < rgC.nu && (lane & (32 / U - 1)) == 0) __stcg(a.val + ent_mine, xn_mine);
            for (int p = rgC.ovf + warp; p < rgC.hi; p += CW) {
                const int e = __ldg(a.entries + p);
                const float xo = __ldcg(a.val + e);
                float* r = a.R + (int64_t)fdiv(e, a.kdiv) * n;
                float r2[NPL], dot = 0.f;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int f = LYS_FEAT(lane, q);
                    r2[q] = (f < n) ? __ldcg(r + f) : 0.f;
                    dot = fmaf(r2[q], dn[q], dot);
                }
                dot = warp_sum(dot);
                const float xn = fmaf(xo, g, dot);
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int f = LYS_FEAT(lane, q);
                    if (f < n) __stcg(r + f, fmaf(-dn[q], xn, fmaf(dold[q], xo, r2[q])));
                }
                if (lane == 0) __stcg(a.val + e, xn);
            }
        }
        LYS_SW_LAP(5);
        // rows / coefficients of this CTA's signals are re-read by other warps of THIS CTA only
        compute_sync<CW>();
        LYS_SW_LAP(6);
        // ---- rotate the pipeline (the row buffers swap roles in the caller)
        gP = g;
#pragma unroll
        for (int q = 0; q < NPL; ++q) { dnP[q] = dn[q]; doldP[q] = dold[q]; dold[q] = doldN[q]; }
        setC = setN; rgC = rgN;
        setN = set2; setN.xp = xp2; rgN = make_range<CW>(lo2, hi2, warp, U);
        set2.ent = ent3; set2.x = x3; set2.lk = lk3;
        ent3 = ent4;
        lo2 = lo3; hi2 = hi3; lo3 = lo4; hi3 = hi4; lo4 = lo5; hi4 = hi5;
    };
    for (int c = 0; c < K; c += 2) {
        atom_step(c, rvA, rvB);
        if (c + 1 < K) atom_step(c + 1, rvB, rvA);
    }
}

template <int NPL>
int launch_sweep(const SweepArgs& args, int grid, cudaStream_t stream)
{
    const bool multi = args.pc.world > 1;
    const void* kern = multi ? (const void*)ksvd_sweep_kernel<NPL, true> : (const void*)ksvd_sweep_kernel<NPL, false>;
    const int threads = SW_THREADS;
    const size_t smem = sizeof(float) * (size_t)(args.lda * (1 + 16));
    LYS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    LYS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) { set_error("ksvd sweep kernel does not fit on an SM"); return LYS_ECUDA; }
    SweepArgs a = args;
    void* params[] = {&a};
    LYS_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(threads), params, smem, stream));
    return LYS_OK;
}

// one CTA per SM; the contribution counter of an accumulator word has 8 bits
int sweep_grid() { return std::min(sm_count(), 255); }

int sweep_lda(int n) { return (2 * n + 3 + 3) / 4 * 4; }
// words between two accumulators: 32 bytes apart (their atomics then spread over more L2 slices) unless that makes
// the array larger than 64 MB
int sweep_acc_stride(int n, int K) { return ((size_t)K * sweep_lda(n) * 32 <= (64u << 20)) ? 4 : 1; }

}  // namespace

int sweep_grid_size() { return sweep_grid(); }

int sweep_bounds(const int32_t* rowptr, const int32_t* entries, int K, int k, int grid, int64_t N, int32_t* bounds, cudaStream_t stream)
{
    const int64_t S = std::max<int64_t>(1, (N + grid - 1) / grid);
    const int items = K * (grid + 1);
    sweep_bounds_kernel<<<(items + 255) / 256, 256, 0, stream>>>(rowptr, entries, K, k, grid, S, bounds);
    LYS_LAUNCH_CHECK("sweep_bounds_kernel");
    return LYS_OK;
}

}  // namespace lys

using namespace lys;

namespace {
struct SweepWs {
    float* Dt; float* Dt_new; u64* acc; size_t acc_bytes; int32_t* bounds; double* fro; double* fx; void* frob_ws; uint8_t* link;
    size_t total;
};
SweepWs carve(void* base, int n, int K, int64_t N, int k)
{
    SweepWs w;
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    auto take = [&](size_t bytes) { unsigned char* r = p; p += align_up(bytes, 256); return r; };
    w.Dt = reinterpret_cast<float*>(take((size_t)n * K * 4));
    w.Dt_new = reinterpret_cast<float*>(take((size_t)n * K * 4));
    w.acc_bytes = (size_t)K * sweep_lda(n) * sweep_acc_stride(n, K) * sizeof(u64);
    w.acc = reinterpret_cast<u64*>(take(w.acc_bytes));
    w.bounds = reinterpret_cast<int32_t*>(take((size_t)K * 256 * 4));
    w.fro = reinterpret_cast<double*>(take(4 * sizeof(double)));
    w.fx = w.fro + 2;
    w.frob_ws = take(sizeof(double) * 4096);
    w.link = reinterpret_cast<uint8_t*>(take((size_t)std::max<int64_t>(N * k, 1)));
    w.total = (size_t)(p - reinterpret_cast<unsigned char*>(base)) + 256;
    return w;
}
}  // namespace

extern "C" size_t lys_ksvd_sweep_workspace_bytes(int n, int K, int64_t N, int k)
{
    if (n < 1 || K < 1 || N < 0 || k < 1) return 0;
    return carve(nullptr, n, K, N, k).total;
}

extern "C" int lys_approx_ksvd_sweep(float* R, float* D, int64_t ldd, const int32_t* idx, float* val,
                                     const int32_t* rowptr, const int32_t* entries,
                                     int n, int K, int64_t N, int k, int n_cycles,
                                     int32_t* unused, void* comm, void* workspace, size_t workspace_bytes,
                                     void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(R && D && idx && val && rowptr && entries && unused && workspace, "lys_approx_ksvd_sweep: null pointer");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K && k >= 1 && k <= LYS_MAX_NONZERO &&
                  n_cycles >= 1 && N >= 0, "lys_approx_ksvd_sweep: bad shape");
    LYS_CHECK_ARG(N * (int64_t)k < (1ll << 31), "lys_approx_ksvd_sweep: N*k must fit int32");
    LYS_CHECK_ARG((reinterpret_cast<uintptr_t>(R) & 15) == 0 && (reinterpret_cast<uintptr_t>(val) & 15) == 0,
                  "lys_approx_ksvd_sweep: R and val must be 16-byte aligned");
    if (workspace_bytes < lys_ksvd_sweep_workspace_bytes(n, K, N, k)) { set_error("lys_approx_ksvd_sweep: workspace too small"); return LYS_EWORKSPACE; }
    PeerComm pc{};
    pc.rank = 0; pc.world = 1;
    CommHost* host = reinterpret_cast<CommHost*>(comm);
    if (host) {
        LYS_CHECK_ARG(host->connected, "lys_approx_ksvd_sweep: comm not connected (call lys_comm_connect)");
        pc = host->dev;
    }
    const int grid = sweep_grid();
    const SweepWs w = carve(workspace, n, K, N, k);
    LYS_CUDA(cudaMemsetAsync(unused, 0, sizeof(int32_t) * (size_t)K, stream));
    int rc = transpose(D, ldd, w.Dt, n, n, K, stream);
    if (rc) return rc;
    if ((rc = sweep_bounds(rowptr, entries, K, k, grid, N, w.bounds, stream))) return rc;
    if (N > 0) {
        sweep_link_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(idx, val, N, k, w.link);
        LYS_LAUNCH_CHECK("sweep_link_kernel");
    }

    SweepArgs a{};
    a.R = R; a.Dt = w.Dt; a.Dt_new = w.Dt_new; a.val = val;
    a.entries = entries; a.link = w.link; a.bounds = w.bounds;
    a.n = n; a.K = K; a.k = k; a.lda = sweep_lda(n); a.acc_stride = sweep_acc_stride(n, K);
    a.unused = unused; a.acc = w.acc; a.fx = w.fx; a.pc = pc;
    a.kdiv = make_fastdiv(k);

    for (int cyc = 0; cyc < n_cycles; ++cyc) {
        // fixed-point scale from the norms of the residual and of the coefficients as they are now
        rc = lys_frobenius2(R, N * (int64_t)n, w.fro, w.frob_ws, sizeof(double) * 4096, stream);
        if (rc) return rc;
        rc = lys_frobenius2(val, N * (int64_t)k, w.fro + 1, w.frob_ws, sizeof(double) * 4096, stream);
        if (rc) return rc;
        unsigned gen = 0;
        if (host) {
            gen = host->generation;
            host->generation += 1u + (unsigned)((K + COMM_WINDOW - 1) / COMM_WINDOW);
        }
        sweep_scale_kernel<<<1, 32, 0, stream>>>(w.fro, w.fx, pc, gen);
        LYS_LAUNCH_CHECK("sweep_scale_kernel");
        a.gen0 = gen + 1u;
        LYS_CUDA(cudaMemsetAsync(w.acc, 0, w.acc_bytes, stream));
        if (n <= 32) rc = launch_sweep<1>(a, grid, stream);
        else if (n <= 64) rc = launch_sweep<2>(a, grid, stream);
        else if (n <= 128) rc = launch_sweep<4>(a, grid, stream);
        else rc = launch_sweep<8>(a, grid, stream);
        if (rc) return rc;
        LYS_CUDA(cudaMemcpyAsync(w.Dt, w.Dt_new, (size_t)n * K * 4, cudaMemcpyDeviceToDevice, stream));
    }
    return transpose(w.Dt, n, D, ldd, K, n, stream);
}

#ifdef LYS_BRINGUP
extern "C" __attribute__((visibility("default"))) int lys_debug_sweep_timing(unsigned long long* out8)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out8, lys::g_sweep_timing, sizeof(unsigned long long) * 8) != cudaSuccess) return -2;
    unsigned long long zero[8] = {0};
    if (cudaMemcpyToSymbol(lys::g_sweep_timing, zero, sizeof(zero)) != cudaSuccess) return -2;
    return 0;
}
#endif
