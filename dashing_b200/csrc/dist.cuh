// dist.cuh — hot path (ii): all-pairs union cardinality / Jaccard / Mash over HLL register arrays.
//
// Replaces the per-pair work of
//   hllbase_t::union_size        bonsai/hll/include/sketch/hll.h:1125-1141  (byte max + 2^p histogram increments)
//   detail::ertl_ml_estimate     hll.h:567-627
//   hllbase_t::jaccard_index     hll.h:1174-1183, full_set_comparison :1165-1173
//   bns::result_cmp              src/dashing.h:568-592
// and the loop nests of perform_core_op (src/sketch_and_cmp.h:699-710), dist_loop (:785-880),
// partdist_loop (src/dashing.h:660-712).
//
// Exact reformulation of the pair histogram.  For threshold k let A_k be the 2^p-bit mask
// [reg_a >= k] ("threshold bit-plane").  Then #{ i : max(a_i,b_i) >= k } = popc(A_k | B_k), and the
// histogram the reference builds with 2^p scalar increments is c[k] = G[k] - G[k+1] with
// G[k] = popc(A_k | B_k).  It is integer-exact, needs no shared-memory atomics, and turns the pair
// loop into OR + POPC over 2^p/32 words per live threshold.  Thresholds outside the value range of
// the sketches involved are known without reading anything (G = 2^p below the minimum, 0 above the
// maximum).  Any fixed permutation of register positions may be used when building the planes —
// the histogram does not depend on it — so planes are built with whatever bit order coalesces best.
//
// HBM layout: planes[t][s][w] (uint32), t = k - gmin - 1 over the global live range (gmin, gmax],
// s = sketch, w < W = 2^p/32; viewed by TMA as a 3-D tensor {W, n, K} with 128-byte swizzled boxes
// of 32 words x 32 sketches.
#pragma once
#include "common.cuh"
#include "estimators.cuh"
#include <cuda.h>

namespace db200 {

constexpr int DT = 32;                 // sketches per panel (tile is DT x DT pairs)
constexpr int DIST_CONSUMERS = 256;    // 8 consumer warps
constexpr int DIST_THREADS = DIST_CONSUMERS + 32;  // + 1 TMA producer warp
constexpr int BOX_BYTES = DT * 128;    // 32 sketches x 32 words
constexpr int STAGE_BYTES = 2 * BOX_BYTES;
constexpr int SPARSE_C = 64;           // a sketch's "sparse tail" = its registers >= T_s, where T_s is the smallest
                                       // threshold with at most SPARSE_C registers at or above it

// ---------------------------------------------------------------------------------------------
// global register range (min / max over the whole matrix)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) range_kernel(const uint4 *__restrict__ regs16, uint64_t n16, uint32_t *minmax) {
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(regs16 + i);
        mn = __vminu4(__vminu4(mn, v.x), __vminu4(v.y, __vminu4(v.z, v.w)));
        mx = __vmaxu4(__vmaxu4(mx, v.x), __vmaxu4(v.y, __vmaxu4(v.z, v.w)));
    }
    uint32_t lo = min(min(mn & 0xFF, (mn >> 8) & 0xFF), min((mn >> 16) & 0xFF, mn >> 24));
    uint32_t hi = max(max(mx & 0xFF, (mx >> 8) & 0xFF), max((mx >> 16) & 0xFF, mx >> 24));
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(minmax, lo); atomicMax(minmax + 1, hi); }
}

// ---------------------------------------------------------------------------------------------
// threshold bit-planes + per-sketch threshold counts.  One CTA (4 warps) per sketch; a warp turns
// 1024 registers (8 coalesced 128-byte loads, 32 registers per lane) into one plane word per lane and
// threshold WITHOUT leaving the lane: registers are < 64, so byte-wise x + (0x80 - k) sets bit 7 of a
// byte exactly when the register is >= k and never carries into the next byte; shifting word gg's
// four flag bits right by 7 - gg interleaves the 8 words into 32 distinct bit positions.
// Register r = g*1024 + gg*128 + lane*4 + j lands in word g*32 + lane, bit 8*j + gg (any fixed
// permutation of register positions serves: the pair histogram does not depend on it).
// (Round 1 built the words with 32 ballots per threshold: 4x the instructions, 2.8 ms for 28,284
// sketches; profiles/r02_planes_kernel_ncu.txt.)
// counts[s][t] = #{ registers of sketch s >= gmin + 1 + t }.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) planes_kernel(const uint32_t *__restrict__ regs32, uint64_t n, uint64_t row0, int p, int gmin, int K,
                                                     uint32_t *__restrict__ planes, uint32_t *__restrict__ counts /*[n][64]*/,
                                                     uint32_t *__restrict__ lists /*[n][SPARSE_C]*/, uint8_t *__restrict__ sthr,
                                                     uint32_t *__restrict__ pthr, uint32_t *__restrict__ llists /*[n][SPARSE_C]*/,
                                                     uint32_t *__restrict__ ptl) {
    __shared__ uint32_t cnt[64];
    __shared__ int s_T, s_TL;
    const uint64_t s = row0 + blockIdx.x;
    // plane rows are at least 32 words (one TMA box) wide: for p < 10 the tail is zero = "below every threshold"
    const uint32_t m = 1u << p, W = max(m >> 5, 32u), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nwords = m >> 2, ngroups = max(m >> 10, 1u);
    if (threadIdx.x < 64) cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t *src = regs32 + s * (m >> 2);
    for (uint32_t g = warp; g < ngroups; g += 4) {
        uint32_t x[8];
        uint32_t valid = 0;                      // bit 8*j + gg set: that register exists (p < 10 only has holes)
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) {
            const uint32_t wi = g * 256 + gg * 32 + lane;
            x[gg] = wi < nwords ? __ldg(src + wi) : 0u;
            if (wi < nwords) valid |= 0x01010101u << gg;
        }
        // thresholds above the largest register of this group of 1024 have empty planes, thresholds up to the smallest
        // one full planes: neither needs the compare sweep
        uint32_t mx4 = x[0], mn4 = valid == 0xFFFFFFFFu ? x[0] : 0u;
#pragma unroll
        for (int gg = 1; gg < 8; ++gg) { mx4 = __vmaxu4(mx4, x[gg]); mn4 = __vminu4(mn4, valid == 0xFFFFFFFFu ? x[gg] : 0u); }
        mx4 = max(max(mx4 & 0xFFu, (mx4 >> 8) & 0xFFu), max((mx4 >> 16) & 0xFFu, mx4 >> 24));
        mn4 = min(min(mn4 & 0xFFu, (mn4 >> 8) & 0xFFu), min((mn4 >> 16) & 0xFFu, mn4 >> 24));
        const int th = min(K, max(0, (int)__reduce_max_sync(0xFFFFFFFFu, mx4) - gmin));    // thresholds [th, K) are empty
        const int tf = min(th, max(0, (int)__reduce_min_sync(0xFFFFFFFFu, mn4) - gmin));   // thresholds [0, tf) are full
        uint32_t *dst = planes + s * W + g * 32 + lane;
        const uint64_t tstride = n * W;
        for (int t = th; t < K; ++t) dst[(uint64_t)t * tstride] = 0u;
        for (int t = 0; t < tf; ++t) dst[(uint64_t)t * tstride] = 0xFFFFFFFFu;
        if (lane == 0)
            for (int t = 0; t < tf; ++t) atomicAdd(&cnt[t], 1024u);
        // byte-wise bias: bit 7 of a byte of x + c is set iff that register >= k = gmin + 1 + t
        uint32_t c4 = (uint32_t)(0x80 - (gmin + 1 + tf)) * 0x01010101u;
        for (int t = tf; t < th; ++t) {
            uint32_t word = 0;
#pragma unroll
            for (int gg = 0; gg < 8; ++gg) word |= ((x[gg] + c4) >> (7 - gg)) & (0x01010101u << gg);
            dst[(uint64_t)t * tstride] = word;
            const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(word));
            if (lane == 0) atomicAdd(&cnt[t], tot);
            c4 -= 0x01010101u;
        }
    }
    __syncthreads();
    if (threadIdx.x < 64) counts[s * 64 + threadIdx.x] = threadIdx.x < (uint32_t)K ? cnt[threadIdx.x] : 0u;
    // sparse tail: T_s = smallest threshold with <= SPARSE_C registers at or above it (gmax + 1 if there is none)
    if (threadIdx.x == 0) {
        int t = 0;
        while (t < K && cnt[t] > (uint32_t)SPARSE_C) ++t;
        s_T = gmin + 1 + t;
        sthr[s] = (uint8_t)s_T;
        atomicMax(pthr + s / DT, (uint32_t)s_T);
        // low tail: TL_s = largest threshold with <= SPARSE_C registers BELOW it (gmin if there is none)
        t = 0;
        while (t < K && m - cnt[t] <= (uint32_t)SPARSE_C) ++t;
        s_TL = gmin + t;
        atomicMin(ptl + s / DT, (uint32_t)s_TL);
    }
    __syncthreads();
    if (warp == 0) {
        // the registers >= T_s as (index << 8 | value), in index order, zero padded: <= SPARSE_C of them by construction
        const uint32_t T4 = (uint32_t)s_T * 0x01010101u;
        uint32_t *dst = lists + s * SPARSE_C;
        uint32_t total = 0;
        for (uint32_t w0 = 0; w0 < nwords; w0 += 32) {
            const uint32_t x = w0 + lane < nwords ? __ldg(src + w0 + lane) : 0u;
            const uint32_t ge = __vcmpgeu4(x, T4) & 0x01010101u;
            if (!__any_sync(0xFFFFFFFFu, ge != 0u)) continue;
            const uint32_t c = (uint32_t)__popc(ge);
            uint32_t incl = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
                if (lane >= (uint32_t)off) incl += v;
            }
            uint32_t pos = total + incl - c;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((ge >> (8 * j)) & 1u) {
                    if (pos < (uint32_t)SPARSE_C) dst[pos] = (((w0 + lane) * 4 + j) << 8) | ((x >> (8 * j)) & 0xFFu);
                    ++pos;
                }
            total += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        for (uint32_t i = total + lane; i < (uint32_t)SPARSE_C; i += 32) dst[i] = 0u;
    } else if (warp == 1) {
        // the registers < TL_s as (index << 8 | value + 1) — never zero —, index order, zero padded (<= SPARSE_C of them)
        const uint32_t T4 = (uint32_t)s_TL * 0x01010101u;
        uint32_t *dst = llists + s * SPARSE_C;
        uint32_t total = 0;
        for (uint32_t w0 = 0; w0 < nwords; w0 += 32) {
            const bool in = w0 + lane < nwords;
            const uint32_t x = in ? __ldg(src + w0 + lane) : 0xFFFFFFFFu;
            const uint32_t lt = __vcmpltu4(x, T4) & 0x01010101u;
            if (!__any_sync(0xFFFFFFFFu, lt != 0u)) continue;
            const uint32_t c = (uint32_t)__popc(lt);
            uint32_t incl = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
                if (lane >= (uint32_t)off) incl += v;
            }
            uint32_t pos = total + incl - c;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((lt >> (8 * j)) & 1u) {
                    if (pos < (uint32_t)SPARSE_C) dst[pos] = (((w0 + lane) * 4 + j) << 8) | (((x >> (8 * j)) & 0xFFu) + 1u);
                    ++pos;
                }
            total += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        for (uint32_t i = total + lane; i < (uint32_t)SPARSE_C; i += 32) dst[i] = 0u;
    }
}

// Histogram accessor over per-sketch threshold counts: G(k) = #{reg >= k}.
struct SketchCounts {
    const uint32_t *g;  // counts row of the sketch
    uint32_t m;
    int gmin, gmax;
    __device__ __forceinline__ uint32_t G(int k) const { return k <= gmin ? m : (k > gmax ? 0u : g[k - gmin - 1]); }
    __device__ __forceinline__ uint32_t operator()(int k) const { return G(k) - G(k + 1); }
};

// per-sketch cardinality + value range, per-panel value range
__global__ void __launch_bounds__(128) card_kernel(const uint32_t *__restrict__ counts, uint64_t row0, uint64_t nrows, int p, int gmin, int gmax, int estim,
                                                   double *__restrict__ card, uint8_t *__restrict__ smin, uint8_t *__restrict__ smax,
                                                   uint32_t *__restrict__ pmin, uint32_t *__restrict__ pmax) {
    const uint64_t s = row0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= row0 + nrows) return;
    SketchCounts c{counts + s * 64, 1u << p, gmin, gmax};
    int lo = gmin, hi = gmax;
    while (lo < gmax && c.G(lo + 1) == c.m) ++lo;   // largest k with every register >= k
    while (hi > gmin && c.G(hi) == 0) --hi;          // largest register value
    card[s] = calculate_estimate(c, estim, p, lo, hi);
    smin[s] = (uint8_t)lo;
    smax[s] = (uint8_t)hi;
    atomicMin(pmin + s / DT, (uint32_t)lo);
    atomicMax(pmax + s / DT, (uint32_t)hi);
}

// ---------------------------------------------------------------------------------------------
// S2 standalone: per-sketch 64-bin histogram (hll.h:515-532) + estimator, one CTA per sketch.
// ---------------------------------------------------------------------------------------------
struct ArrayCounts {
    const uint32_t *c;
    __device__ __forceinline__ uint32_t operator()(int k) const { return c[k]; }
};

__global__ void __launch_bounds__(128) cardinality_kernel(const uint32_t *__restrict__ regs32, int p, int estim, double *__restrict__ out) {
    __shared__ uint32_t wh[4][64];
    __shared__ uint32_t hist[64];
    const uint32_t m4 = 1u << (p - 2), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 256; i += 128) (&wh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t *src = regs32 + (uint64_t)blockIdx.x * m4;
    for (uint32_t i = threadIdx.x; i < m4; i += 128) {
        const uint32_t x = __ldg(src + i);
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&wh[warp][(x >> (8 * j)) & 63u], 1u);
    }
    (void)lane;
    __syncthreads();
    if (threadIdx.x < 64) hist[threadIdx.x] = wh[0][threadIdx.x] + wh[1][threadIdx.x] + wh[2][threadIdx.x] + wh[3][threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = calculate_estimate(ArrayCounts{hist}, estim, p);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// Producer-side wait: the TMA lane has nothing else to do, so it sleeps between polls instead of burning issue slots of
// its scheduler (first sparse-tail profile: 8 % of all warp instructions were this spin loop).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (done) break;
        __nanosleep(256);
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// the all-pairs kernel
// ---------------------------------------------------------------------------------------------
struct DistTile { uint32_t a, b; };  // panel indices: A panel = sketches [rowA(a), +32), B panel = [32*b, +32)

struct DistArgs {
    const DistTile *tiles;
    const uint8_t *smin, *smax;       // per sketch
    const uint32_t *pmin, *pmax;      // per 32-sketch panel
    const double *card;               // per sketch (estimator `estim`)
    const uint32_t *counts;           // per sketch threshold counts [n][64]
    const uint32_t *lists;            // per sketch sparse tail [n][SPARSE_C]
    const uint32_t *pthr;             // per panel: max over its sketches of the sparse-tail threshold T_s
    const uint32_t *llists;           // per sketch low tail [n][SPARSE_C]: registers below TL_s
    const uint32_t *ptl;              // per panel: min over its sketches of TL_s
    float *out;
    uint64_t n;                       // sketches in the plane tensor
    uint64_t row_begin, row_end;      // symmetric: rows computed; rect: unused
    uint64_t out_base;                // symmetric: distmat offset of row_begin
    uint64_t nr, nq, qbase;           // rect: references rows [0,nr), queries rows [qbase, qbase+nq); qbase % DT == 0
    double ksinv;
    int p, gmin, gmax, K;
    int estim, rtype;
    int rect;                         // 0 symmetric, 1 rectangular (A = queries, B = references)
    int stages;
    int low;                          // 1: low tails are staged too (16 + 16 + 3.6 KB: needs 5 stage buffers; sparse alone 4)
    int sparse;                       // 0: shared memory too tight to stage the sparse tails -> every live threshold is swept densely
    int one;                          // 1 — a runtime value so that ptxas keeps `popc * one + acc` as an IMAD (FMA pipe)
};

// Pair histogram accessor over the tile's threshold counts in shared memory (uint16, see wrap rule).
template <typename GT>
struct PairCounts {
    const GT *g;         // &G[0][pair]
    uint32_t m;
    int lo, hi;          // tile value range: thresholds lo+1..hi are stored
    int kmax_pair;       // upper bound on the histogram's largest value (for the 2^16 wrap rule)
    int stride;          // elements between consecutive thresholds
    int cap;             // G(k) = 0 for k > cap
    __device__ __forceinline__ uint32_t G(int k) const {
        if (k > cap) return 0u;
        if (k <= lo) return m;
        if (k > hi) return 0u;
        const uint32_t v = g[(k - lo - 1) * stride];
        // uint16 counts are stored mod 2^16; 0 inside the pair's live range can only mean 2^16 (p == 16)
        return (sizeof(GT) == 2 && v == 0u && k <= kmax_pair) ? m : v;
    }
    __device__ __forceinline__ uint32_t operator()(int k) const { return G(k) - G(k + 1); }
};

// Pair histogram with the bin counts themselves in shared memory (dist_kernel converts its threshold counts in place
// before the estimator runs: one LDS per bin instead of two plus range logic).  Slot s holds bin lo + s, s = 0..hi-lo.
template <typename GT>
struct PairHist {
    const GT *c;         // &C[0][pair]
    uint32_t m;
    int lo, hi;
    int fullk;           // bin holding all 2^16 registers (stored as 0), or -1
    __device__ __forceinline__ uint32_t operator()(int k) const {
        const unsigned s = (unsigned)(k - lo);
        if (s > (unsigned)(hi - lo)) return 0u;
        return k == fullk ? m : (uint32_t)c[s * (DT * DT)];
    }
};

// a / b for finite, normal, positive b: reciprocal seed + two Newton steps + one correction (<= 1 ulp), no special-case
// path.  The estimator's denominators (x' + 1 - h, with h in (0,1]) always qualify.
__device__ __forceinline__ double fast_div(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    double q = a * r;
    const double rem = fma(-b, q, a);
    return fma(rem, r, q);
}

// Carry-save adder over three bit-vectors: h = majority (carry), l = parity (sum); two LOP3.
#define DB200_CSA(h, l, a, b, c)                                        \
    do {                                                                \
        const uint32_t u__ = (a), v__ = (b), w__ = (c);                 \
        asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(h) : "r"(u__), "r"(v__), "r"(w__)); /* majority */ \
        asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(l) : "r"(u__), "r"(v__), "r"(w__)); /* parity   */ \
    } while (0)
// 8 OR-words (two 16-byte chunks of A row `x` and B row `y`) -> planes (o,t,f) + 8 * popc(eights)
#define DB200_HS8(acc, o, t, f, x0, x1, y0, y1)                                                   \
    do {                                                                                          \
        uint32_t c1, c2, c3, c4, d1, d2, e;                                                       \
        DB200_CSA(c1, o, o, x0.x | y0.x, x0.y | y0.y);                                            \
        DB200_CSA(c2, o, o, x0.z | y0.z, x0.w | y0.w);                                            \
        DB200_CSA(d1, t, t, c1, c2);                                                              \
        DB200_CSA(c3, o, o, x1.x | y1.x, x1.y | y1.y);                                            \
        DB200_CSA(c4, o, o, x1.z | y1.z, x1.w | y1.w);                                            \
        DB200_CSA(d2, t, t, c3, c4);                                                              \
        DB200_CSA(e, f, f, d1, d2);                                                               \
        acc = (uint32_t)__popc(e) * eight + acc;                                                  \
    } while (0)

struct NewtonDiv { __device__ __forceinline__ static double div(double a, double b) { return fast_div(a, b); } };

// GT: uint16_t for p <= 16 (two CTAs per SM), uint32_t above (counts no longer fit 16 bits)
template <typename GT>
__global__ void __launch_bounds__(DIST_THREADS, 2) dist_kernel(const __grid_constant__ CUtensorMap tmap, const DistArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = a.stages;
    uint8_t *stage_mem = smem;                                             // S x {A box, B box}
    GT *G = reinterpret_cast<GT *>(smem + (size_t)S * STAGE_BYTES);  // [K + 1][1024]
    const int Kcap = a.K + 1;   // slots for bins lo .. hi of the tile (slot 0 is filled when counts become bins)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * STAGE_BYTES + (size_t)Kcap * DT * DT * sizeof(GT));
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);

    const DistTile tile = a.tiles[blockIdx.x];
    const uint64_t rowA0 = a.rect ? a.qbase + (uint64_t)tile.a * DT : (uint64_t)tile.a * DT;
    const uint64_t rowB0 = (uint64_t)tile.b * DT;
    const uint32_t panA = (uint32_t)(rowA0 / DT), panB = tile.b;
    const int lo = (int)min(a.pmin[panA], a.pmin[panB]);
    const int hi = (int)max(a.pmax[panA], a.pmax[panB]);
    // Thresholds lo+1 .. Td-1 are counted densely from the bit-planes.  From Td on every sketch of the tile has at most
    // SPARSE_C registers at or above the threshold, and G(k) = #a(k) + #b(k) - #{i : a_i >= k and b_i >= k} is obtained by
    // merging the two sorted sparse tails — no plane traffic, no POPC.
    const int Tt = a.sparse ? (int)max(a.pthr[panA], a.pthr[panB]) : 255;
    const int Td = min(max(Tt, lo + 1), hi + 1);
    // Symmetrically, for the lowest thresholds almost every register is at or above k: while every sketch of the tile has at
    // most SPARSE_C registers below k (k <= TLe), G(k) = 2^p - #{i : a_i < k and b_i < k} comes from merging the low tails.
    const int TLe = (a.sparse && a.low) ? min((int)min(a.ptl[panA], a.ptl[panB]), Td - 1) : lo;
    const int kd0 = max(lo, TLe) + 1;                      // first densely swept threshold
    const int W = max(1 << (a.p - 5), 32), nbox = W >> 5;
    const int iters = max(Td - kd0, 0) * nbox;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, DIST_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == DIST_CONSUMERS / 32) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            int s = 0, t = kd0 - a.gmin - 1, wb = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait_relaxed(empty0 + 8 * s, ph ^ 1u);
                const uint32_t dst = smem_u32(stage_mem + (size_t)s * STAGE_BYTES);
                mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                tma_load_3d(dst, &tmap, wb * 32, (int)rowA0, t, full0 + 8 * s);
                tma_load_3d(dst + BOX_BYTES, &tmap, wb * 32, (int)rowB0, t, full0 + 8 * s);
                if (++s == S) { s = 0; ph ^= 1u; }
                if (++wb == nbox) { wb = 0; ++t; }
            }
        }
    } else {
        // ---------------- consumers: OR + POPC ----------------
        const uint32_t ti = threadIdx.x >> 4, tj = threadIdx.x & 15;  // A rows {ti, ti+16}, B rows {tj, tj+16}
        const uint32_t swA = (ti & 7) << 4, swB = (tj & 7) << 4;
        // POPC runs on the 16-lane/clk XU pipe, which binds this kernel (ncu: XU 91 %, ALU 45 %).  Half of every box is
        // therefore counted with a carry-save adder tree on the ALU pipe instead (Harley-Seal): 8 OR-words are reduced by
        // 7 CSAs (2 LOP3 each) into bit-planes of weight 1, 2, 4 that persist in registers, plus ONE weight-8 plane that is
        // popcounted.  The planes are flushed with three POPCs when the threshold ends.  Still integer-exact.
        uint32_t acc00 = 0, acc01 = 0, acc10 = 0, acc11 = 0;
        uint32_t o00 = 0, o01 = 0, o10 = 0, o11 = 0;   // weight-1 planes
        uint32_t t00 = 0, t01 = 0, t10 = 0, t11 = 0;   // weight-2 planes
        uint32_t f00 = 0, f01 = 0, f10 = 0, f11 = 0;   // weight-4 planes
        // accumulate with IMAD (x * one + acc, `one` a runtime 1): keeps the adds off the ALU pipe
#define DB200_POP4(acc, p, q)                                                    \
    do {                                                                         \
        acc = (uint32_t)__popc(p.x | q.x) * one + acc;                           \
        acc = (uint32_t)__popc(p.y | q.y) * one + acc;                           \
        acc = (uint32_t)__popc(p.z | q.z) * one + acc;                           \
        acc = (uint32_t)__popc(p.w | q.w) * one + acc;                           \
    } while (0)
        const uint32_t one = (uint32_t)a.one, eight = one << 3;
        int s = 0, wb = 0, tl = kd0 - lo;      // threshold lo + tl -> slot tl
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full0 + 8 * s, ph);
            const uint8_t *A = stage_mem + (size_t)s * STAGE_BYTES, *B = A + BOX_BYTES;
            const uint8_t *a0p = A + ti * 128, *a1p = A + (ti + 16) * 128;
            const uint8_t *b0p = B + tj * 128, *b1p = B + (tj + 16) * 128;
            // chunks 0-3: direct POPC
#pragma unroll
            for (uint32_t c = 0; c < 4; ++c) {
                const uint32_t oa = (c << 4) ^ swA, ob = (c << 4) ^ swB;
                const uint4 a0 = *reinterpret_cast<const uint4 *>(a0p + oa);
                const uint4 a1 = *reinterpret_cast<const uint4 *>(a1p + oa);
                const uint4 b0 = *reinterpret_cast<const uint4 *>(b0p + ob);
                const uint4 b1 = *reinterpret_cast<const uint4 *>(b1p + ob);
                DB200_POP4(acc00, a0, b0);
                DB200_POP4(acc01, a0, b1);
                DB200_POP4(acc10, a1, b0);
                DB200_POP4(acc11, a1, b1);
            }
            // chunks 4-7: carry-save adders
#pragma unroll
            for (uint32_t c = 4; c < 8; c += 2) {
                const uint32_t oa = (c << 4) ^ swA, ob = (c << 4) ^ swB, oa2 = ((c + 1) << 4) ^ swA, ob2 = ((c + 1) << 4) ^ swB;
                const uint4 a0 = *reinterpret_cast<const uint4 *>(a0p + oa), a0n = *reinterpret_cast<const uint4 *>(a0p + oa2);
                const uint4 a1 = *reinterpret_cast<const uint4 *>(a1p + oa), a1n = *reinterpret_cast<const uint4 *>(a1p + oa2);
                const uint4 b0 = *reinterpret_cast<const uint4 *>(b0p + ob), b0n = *reinterpret_cast<const uint4 *>(b0p + ob2);
                const uint4 b1 = *reinterpret_cast<const uint4 *>(b1p + ob), b1n = *reinterpret_cast<const uint4 *>(b1p + ob2);
                DB200_HS8(acc00, o00, t00, f00, a0, a0n, b0, b0n);
                DB200_HS8(acc01, o01, t01, f01, a0, a0n, b1, b1n);
                DB200_HS8(acc10, o10, t10, f10, a1, a1n, b0, b0n);
                DB200_HS8(acc11, o11, t11, f11, a1, a1n, b1, b1n);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
            if (++s == S) { s = 0; ph ^= 1u; }
            if (++wb == nbox) {
                wb = 0;
                GT *g = G + (size_t)(tl++) * (DT * DT);
                g[ti * DT + tj] = (GT)(acc00 + __popc(o00) + 2 * __popc(t00) + 4 * __popc(f00));
                g[ti * DT + tj + 16] = (GT)(acc01 + __popc(o01) + 2 * __popc(t01) + 4 * __popc(f01));
                g[(ti + 16) * DT + tj] = (GT)(acc10 + __popc(o10) + 2 * __popc(t10) + 4 * __popc(f10));
                g[(ti + 16) * DT + tj + 16] = (GT)(acc11 + __popc(o11) + 2 * __popc(t11) + 4 * __popc(f11));
                acc00 = acc01 = acc10 = acc11 = 0;
                o00 = o01 = o10 = o11 = t00 = t01 = t10 = t11 = f00 = f01 = f10 = f11 = 0;
            }
        }
#undef DB200_POP4
    }
    __syncthreads();

    // ---------------- sparse tails of the 64 sketches of the tile -> shared memory (re-using the stage buffers) --------
    const uint32_t m = 1u << a.p;
    const int ns = hi - Td + 1;                                       // sparse thresholds Td..hi
    uint32_t *L = reinterpret_cast<uint32_t *>(stage_mem);             // [2*DT][SPARSE_C]                        16 KB
    uint32_t *LL = L + 2 * DT * SPARSE_C;                              // [2*DT][SPARSE_C] low tails (a.low only)  16 KB
    // [2*DT][ns]: #{reg >= Td + kk} — at most SPARSE_C for every threshold >= Td >= T_s, so a byte each (<= 3.6 KB)
    uint8_t *CN = reinterpret_cast<uint8_t *>(L + (a.low ? 4 : 2) * DT * SPARSE_C);
    const int nl = kd0 - 1 - lo;                                       // low-sparse thresholds lo+1 .. kd0-1
    if (nl > 0) {
        for (uint32_t e = threadIdx.x; e < 2 * DT * SPARSE_C; e += DIST_THREADS) {
            const uint32_t row = e / SPARSE_C;
            const uint64_t sk = row < DT ? rowA0 + row : rowB0 + (row - DT);
            LL[e] = sk < a.n ? a.llists[sk * SPARSE_C + (e % SPARSE_C)] : 0u;
        }
    }
    if (ns > 0) {
        // one warp per sketch: keep only the entries >= Td (Td >= the sketch's own T_s), order preserved, zero terminated
        static_assert(SPARSE_C == 64, "compaction below assumes two entries per lane");
        for (uint32_t row = warp; row < 2 * DT; row += DIST_THREADS / 32) {
            const uint64_t sk = row < DT ? rowA0 + row : rowB0 + (row - DT);
            const uint32_t e0 = sk < a.n ? a.lists[sk * SPARSE_C + lane] : 0u, e1 = sk < a.n ? a.lists[sk * SPARSE_C + 32 + lane] : 0u;
            const bool k0 = (int)(e0 & 0xFFu) >= Td, k1 = (int)(e1 & 0xFFu) >= Td;
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, k0), b1 = __ballot_sync(0xFFFFFFFFu, k1);
            const uint32_t lt = (1u << lane) - 1u, n0 = (uint32_t)__popc(b0), tot = n0 + (uint32_t)__popc(b1);
            uint32_t *dst = L + row * SPARSE_C;
            if (k0) dst[__popc(b0 & lt)] = e0;
            if (k1) dst[n0 + __popc(b1 & lt)] = e1;
            for (uint32_t i2 = tot + lane; i2 < (uint32_t)SPARSE_C; i2 += 32) dst[i2] = 0u;
        }
        for (uint32_t e = threadIdx.x; e < (uint32_t)(2 * DT * ns); e += DIST_THREADS) {
            const uint32_t row = e / ns;
            const int k = Td + (int)(e % ns);
            const uint64_t sk = row < DT ? rowA0 + row : rowB0 + (row - DT);
            CN[e] = (uint8_t)((sk < a.n && k <= a.gmax) ? a.counts[sk * 64 + (k - a.gmin - 1)] : 0u);   // k >= Td > gmin
        }
    }
    __syncthreads();

    // ---------------- estimator + emission, one pair per thread at a time ----------------
    for (uint32_t pair = threadIdx.x; pair < DT * DT; pair += DIST_THREADS) {
        const uint32_t il = pair >> 5, jl = pair & 31;
        const uint64_t i = rowA0 + il, j = rowB0 + jl;
        uint64_t oidx;
        if (a.rect) {
            if (i >= a.qbase + a.nq || j >= a.nr) continue;
            oidx = (i - a.qbase) * a.nr + j;
        } else {
            if (i >= j || j >= a.n || i < a.row_begin || i >= a.row_end) continue;
            oidx = (i * (2 * a.n - i - 1)) / 2 - a.out_base + (j - i - 1);
        }
        if (nl > 0) {
            // thresholds lo+1 .. kd0-1: start from 2^p and take out the registers that are below k in both sketches
            GT *g = G + (size_t)(DT * DT) + pair;                  // slot of threshold lo + 1
            for (int kk = 0; kk < nl; ++kk) g[kk * (DT * DT)] = (GT)m;
            const uint32_t *la = LL + il * SPARSE_C, *lb = LL + (DT + jl) * SPARSE_C;
            int ia = 0, ib = 0;
            uint32_t ea = la[0], eb = lb[0];
            while (ea != 0u && eb != 0u) {
                const uint32_t xa = ea >> 8, xb = eb >> 8;
                if (xa == xb) {
                    const int mx = (int)max(ea & 0xFFu, eb & 0xFFu) - 1;   // values are stored + 1
                    for (int k = max(mx, lo) + 1; k < kd0; ++k) g[(k - lo - 1) * (DT * DT)] -= 1;
                }
                if (xa <= xb) ea = ++ia < SPARSE_C ? la[ia] : 0u;
                if (xb <= xa) eb = ++ib < SPARSE_C ? lb[ib] : 0u;
            }
        }
        if (ns > 0) {
            GT *g = G + (size_t)(Td - lo) * (DT * DT) + pair;
            const uint8_t *ca = CN + il * ns, *cb = CN + (DT + jl) * ns;
            for (int kk = 0; kk < ns; ++kk) g[kk * (DT * DT)] = (GT)((uint32_t)ca[kk] + cb[kk]);
            // merge the two index-sorted tails; a register present in both with min value mn was counted twice for k <= mn
            const uint32_t *la = L + il * SPARSE_C, *lb = L + (DT + jl) * SPARSE_C;
            int ia = 0, ib = 0;
            uint32_t ea = la[0], eb = lb[0];
            while (ea != 0u && eb != 0u) {
                const uint32_t xa = ea >> 8, xb = eb >> 8;
                if (xa == xb) {
                    const int mn = (int)min(ea & 0xFFu, eb & 0xFFu);
                    for (int k = Td; k <= mn; ++k) g[(k - Td) * (DT * DT)] -= 1;
                }
                if (xa <= xb) ea = ++ia < SPARSE_C ? la[ia] : 0u;
                if (xb <= xa) eb = ++ib < SPARSE_C ? lb[ib] : 0u;
            }
        }
        const int kmin_pair = max((int)a.smin[i], (int)a.smin[j]);
        const int kmax_pair = max((int)a.smax[i], (int)a.smax[j]);
        // threshold counts G(k) (slots 1..hi-lo) -> bin counts c(k) = G(k) - G(k+1) in place, slot 0 = bin lo
        int fullk = -1;
        {
            GT *col = G + pair;
            uint32_t gk = m;                                    // G(lo) = 2^p
            for (int k = lo; k <= hi; ++k) {
                uint32_t gn = 0u;                               // G(hi + 1) = 0
                if (k < hi) {
                    gn = col[(k + 1 - lo) * (DT * DT)];
                    if (sizeof(GT) == 2 && gn == 0u && k + 1 <= kmax_pair) gn = m;  // uint16 counts are stored mod 2^16 (only p == 16 can wrap)
                }
                const uint32_t ck = gk - gn;
                if (sizeof(GT) == 2 && ck > 0xFFFFu) fullk = k;
                col[(k - lo) * (DT * DT)] = (GT)ck;
                gk = gn;
            }
        }
        PairHist<GT> c{G + pair, m, lo, hi, fullk};
        const double us = calculate_estimate<PairHist<GT>, NewtonDiv>(c, a.estim, a.p, kmin_pair, kmax_pair);
        // non-joint path is symmetric in its operands (IEEE addition commutes): lhs = A, rhs = B
        const double cl = a.card[i], cr = a.card[j];
        // jaccard_index, hll.h:1179-1182
        const double r = (cl + cr - us) / us;
        const double ji = 0. < r ? r : 0.;
        // full_set_comparison, hll.h:1169-1172
        double is = cl + cr - us;
        is = is < 0. ? 0. : is;
        double t0 = cl - is, t1 = cr - is;
        t0 = t0 < 0. ? 0. : t0;
        t1 = t1 < 0. ? 0. : t1;
        // DB200_UNION_SIZE (extension): hll_t::union_size itself, hll.h:1125-1138
        a.out[oidx] = a.rtype == DB200_UNION_SIZE ? (float)us : emit_value(a.rtype, ji, t0, t1, is, a.ksinv);
    }
}


// ---------------------------------------------------------------------------------------------
// Joint-MLE (-J) variant: ertl_joint, hll.h:636-684.  Besides the union histogram the reference
// builds countsAXBhalf[v] = cg1[v] + ceq[v] + cg2[v+1] (hll.h:660-674) — which is exactly the histogram
// of max(a_i, b_i - 1) — and its mirror image.  In threshold form
//     #{max(a, b-1) >= k} = popc(A_k | B_{k+1}),      #{max(a-1, b) >= k} = popc(A_{k+1} | B_k),
// so one pass over the planes k and k+1 yields all three count families.  Tile = 16 x 32 pairs
// (three uint16 count families must fit in shared memory for up to 52 thresholds).
// ---------------------------------------------------------------------------------------------
constexpr int JT = 16;                                 // A-panel rows of the joint tile
constexpr int JBOX_A = JT * 128, JBOX_B = DT * 128;
constexpr int JSTAGE_BYTES = 2 * (JBOX_A + JBOX_B);    // A_k, A_{k+1}, B_k, B_{k+1}
constexpr int JPAIRS = JT * DT;
// One pair per consumer thread (16 consumer warps + the TMA producer warp): shared memory allows a single CTA per SM
// here, so the warps have to come from the CTA itself (8 consumer warps left the schedulers half idle: ALU 68 %,
// XU 64 %, issue 51 % — profiles/r01l_jmle_kernel_ncu.txt).
constexpr int JCONSUMERS = JPAIRS, JTHREADS = JCONSUMERS + 32;

template <typename GT>
__global__ void __launch_bounds__(JTHREADS, 1) dist_jmle_kernel(const __grid_constant__ CUtensorMap tmapA,
                                                                    const __grid_constant__ CUtensorMap tmapB, const DistArgs a,
                                                                    const int lhs_is_b) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = a.stages;
    uint8_t *stage_mem = smem;
    GT *G = reinterpret_cast<GT *>(smem + (size_t)S * JSTAGE_BYTES);  // [3][K][512]
    const int Kcap = a.K > 0 ? a.K : 1;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * JSTAGE_BYTES + (size_t)3 * Kcap * JPAIRS * sizeof(GT));
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);

    const DistTile tile = a.tiles[blockIdx.x];          // tile.a counts 16-row half panels
    const uint64_t rowA0 = (a.rect ? a.qbase : 0) + (uint64_t)tile.a * JT;
    const uint64_t rowB0 = (uint64_t)tile.b * DT;
    const uint32_t panA = (uint32_t)(rowA0 / DT), panB = tile.b;
    const int lo = (int)min(a.pmin[panA], a.pmin[panB]);
    const int hi = (int)max(a.pmax[panA], a.pmax[panB]);
    // dense thresholds lo+1 .. Td-1 from the planes, sparse thresholds Td .. hi from the merged tails (see dist_kernel)
    const int Tt = a.sparse ? (int)max(a.pthr[panA], a.pthr[panB]) : 255;
    const int Td = min(max(Tt, lo + 1), hi + 1);
    const int W = max(1 << (a.p - 5), 32), nbox = W >> 5;
    const int iters = (Td - 1 - lo) * nbox;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, JCONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == JCONSUMERS / 32) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
            int s = 0, t = lo - a.gmin, wb = 0;                          // t = plane of threshold k = lo + 1 + (stage / nbox)
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait_relaxed(empty0 + 8 * s, ph ^ 1u);
                const uint32_t dst = smem_u32(stage_mem + (size_t)s * JSTAGE_BYTES);
                mbar_expect_tx(full0 + 8 * s, JSTAGE_BYTES);
                // plane t+1 beyond the last stored threshold is out of bounds -> zero fill == "no register that large"
                tma_load_3d(dst, &tmapA, wb * 32, (int)rowA0, t, full0 + 8 * s);
                tma_load_3d(dst + JBOX_A, &tmapA, wb * 32, (int)rowA0, t + 1, full0 + 8 * s);
                tma_load_3d(dst + 2 * JBOX_A, &tmapB, wb * 32, (int)rowB0, t, full0 + 8 * s);
                tma_load_3d(dst + 2 * JBOX_A + JBOX_B, &tmapB, wb * 32, (int)rowB0, t + 1, full0 + 8 * s);
                if (++s == S) { s = 0; ph ^= 1u; }
                if (++wb == nbox) { wb = 0; ++t; }
            }
        }
    } else {
        const uint32_t ti = threadIdx.x >> 5, tj = lane;               // pair (A row ti, B row tj)
        const uint32_t swA = (ti & 7) << 4, swB = (tj & 7) << 4;
        uint32_t u0 = 0, x0 = 0, y0 = 0;
        // carry-save planes (weight 1, 2, 4) per family: half of every box is counted on the ALU pipe (see dist_kernel)
        uint32_t ou0 = 0, ox0 = 0, oy0 = 0, tu0 = 0, tx0 = 0, ty0 = 0, fu0 = 0, fx0 = 0, fy0 = 0;
        const uint32_t one = (uint32_t)a.one, eight = 8u * one;
        int s = 0, wb = 0;
        size_t tl = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full0 + 8 * s, ph);
            const uint8_t *A0 = stage_mem + (size_t)s * JSTAGE_BYTES, *A1 = A0 + JBOX_A, *B0 = A0 + 2 * JBOX_A, *B1 = B0 + JBOX_B;
#define DB200_POP4(acc, p, q)                                                    \
    do {                                                                         \
        acc = (uint32_t)__popc(p.x | q.x) * one + acc;                           \
        acc = (uint32_t)__popc(p.y | q.y) * one + acc;                           \
        acc = (uint32_t)__popc(p.z | q.z) * one + acc;                           \
        acc = (uint32_t)__popc(p.w | q.w) * one + acc;                           \
    } while (0)
#pragma unroll
            for (uint32_t c = 0; c < 4; ++c) {
                const uint32_t oa = ti * 128 + ((c << 4) ^ swA), ob = tj * 128 + ((c << 4) ^ swB);
                const uint4 ak = *reinterpret_cast<const uint4 *>(A0 + oa), an = *reinterpret_cast<const uint4 *>(A1 + oa);
                const uint4 bk = *reinterpret_cast<const uint4 *>(B0 + ob), bn = *reinterpret_cast<const uint4 *>(B1 + ob);
                DB200_POP4(u0, ak, bk); DB200_POP4(x0, ak, bn); DB200_POP4(y0, an, bk);
            }
#undef DB200_POP4
#pragma unroll
            for (uint32_t c = 4; c < 8; c += 2) {
                const uint32_t oa = ti * 128 + ((c << 4) ^ swA), oa2 = ti * 128 + (((c + 1) << 4) ^ swA);
                const uint32_t ob = tj * 128 + ((c << 4) ^ swB), ob2 = tj * 128 + (((c + 1) << 4) ^ swB);
                const uint4 ak = *reinterpret_cast<const uint4 *>(A0 + oa), ak2 = *reinterpret_cast<const uint4 *>(A0 + oa2);
                const uint4 an = *reinterpret_cast<const uint4 *>(A1 + oa), an2 = *reinterpret_cast<const uint4 *>(A1 + oa2);
                const uint4 bk = *reinterpret_cast<const uint4 *>(B0 + ob), bk2 = *reinterpret_cast<const uint4 *>(B0 + ob2);
                const uint4 bn = *reinterpret_cast<const uint4 *>(B1 + ob), bn2 = *reinterpret_cast<const uint4 *>(B1 + ob2);
                DB200_HS8(u0, ou0, tu0, fu0, ak, ak2, bk, bk2);
                DB200_HS8(x0, ox0, tx0, fx0, ak, ak2, bn, bn2);
                DB200_HS8(y0, oy0, ty0, fy0, an, an2, bk, bk2);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
            if (++s == S) { s = 0; ph ^= 1u; }
            if (++wb == nbox) {
                wb = 0;
                GT *gu = G + tl * JPAIRS, *gx = G + ((size_t)Kcap + tl) * JPAIRS, *gy = G + ((size_t)2 * Kcap + tl) * JPAIRS;
                const uint32_t p0 = ti * DT + tj;
#define DB200_FLUSH(acc, o, t, f) ((GT)((acc) + __popc(o) + 2 * __popc(t) + 4 * __popc(f)))
                gu[p0] = DB200_FLUSH(u0, ou0, tu0, fu0);
                gx[p0] = DB200_FLUSH(x0, ox0, tx0, fx0);
                gy[p0] = DB200_FLUSH(y0, oy0, ty0, fy0);
#undef DB200_FLUSH
                u0 = x0 = y0 = 0;
                ou0 = ox0 = oy0 = tu0 = tx0 = ty0 = fu0 = fx0 = fy0 = 0;
                ++tl;
            }
        }
    }
    __syncthreads();

    const uint32_t m = 1u << a.p;
    const int q = 64 - a.p;
    const int ns = hi - Td + 1;                                       // sparse thresholds Td..hi
    uint32_t *L = reinterpret_cast<uint32_t *>(stage_mem);             // [JT + DT][SPARSE_C]
    uint32_t *CN = L + (JT + DT) * SPARSE_C;                           // [JT + DT][ns + 1]: #{reg >= Td + kk}
    if (ns > 0) {
        for (uint32_t row = warp; row < (uint32_t)(JT + DT); row += JTHREADS / 32) {
            const uint64_t sk = row < JT ? rowA0 + row : rowB0 + (row - JT);
            const uint32_t e0 = sk < a.n ? a.lists[sk * SPARSE_C + lane] : 0u, e1 = sk < a.n ? a.lists[sk * SPARSE_C + 32 + lane] : 0u;
            const bool k0 = (int)(e0 & 0xFFu) >= Td, k1 = (int)(e1 & 0xFFu) >= Td;
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, k0), b1 = __ballot_sync(0xFFFFFFFFu, k1);
            const uint32_t lt = (1u << lane) - 1u, n0 = (uint32_t)__popc(b0), tot = n0 + (uint32_t)__popc(b1);
            uint32_t *dst = L + row * SPARSE_C;
            if (k0) dst[__popc(b0 & lt)] = e0;
            if (k1) dst[n0 + __popc(b1 & lt)] = e1;
            for (uint32_t i2 = tot + lane; i2 < (uint32_t)SPARSE_C; i2 += 32) dst[i2] = 0u;
        }
        for (uint32_t e = threadIdx.x; e < (uint32_t)((JT + DT) * (ns + 1)); e += JTHREADS) {
            const uint32_t row = e / (ns + 1);
            const int k = Td + (int)(e % (ns + 1));
            const uint64_t sk = row < JT ? rowA0 + row : rowB0 + (row - JT);
            CN[e] = (sk < a.n && k <= a.gmax) ? (k <= a.gmin ? m : a.counts[sk * 64 + (k - a.gmin - 1)]) : 0u;
        }
    }
    __syncthreads();

    for (uint32_t pair = threadIdx.x; pair < (uint32_t)JPAIRS; pair += JTHREADS) {
        const uint32_t il = pair >> 5, jl = pair & 31;
        const uint64_t i = rowA0 + il, j = rowB0 + jl;
        uint64_t oidx;
        if (a.rect) {
            if (i >= a.qbase + a.nq || j >= a.nr) continue;
            oidx = (i - a.qbase) * a.nr + j;
        } else {
            if (i >= j || j >= a.n || i < a.row_begin || i >= a.row_end) continue;
            oidx = (i * (2 * a.n - i - 1)) / 2 - a.out_base + (j - i - 1);
        }
        if (ns > 0) {
            // G_U(k) = #a(k) + #b(k)   - #{a_i >= k,   b_i >= k  }
            // G_X(k) = #a(k) + #b(k+1) - #{a_i >= k,   b_i >= k+1}      (histogram of max(a, b-1))
            // G_Y(k) = #a(k+1) + #b(k) - #{a_i >= k+1, b_i >= k  }      (histogram of max(a-1, b))
            GT *gu = G + (size_t)(Td - lo - 1) * JPAIRS + pair;
            GT *gx = gu + (size_t)Kcap * JPAIRS, *gy = gx + (size_t)Kcap * JPAIRS;
            const uint32_t *ca = CN + il * (ns + 1), *cb = CN + (JT + jl) * (ns + 1);
            for (int kk = 0; kk < ns; ++kk) {
                gu[kk * JPAIRS] = (GT)(ca[kk] + cb[kk]);
                gx[kk * JPAIRS] = (GT)(ca[kk] + cb[kk + 1]);
                gy[kk * JPAIRS] = (GT)(ca[kk + 1] + cb[kk]);
            }
            const uint32_t *la = L + il * SPARSE_C, *lb = L + (JT + jl) * SPARSE_C;
            int ia = 0, ib = 0;
            uint32_t ea = la[0], eb = lb[0];
            while (ea != 0u && eb != 0u) {
                const uint32_t xa = ea >> 8, xb = eb >> 8;
                if (xa == xb) {
                    const int va = (int)(ea & 0xFFu), vb = (int)(eb & 0xFFu);
                    for (int k = Td; k <= min(va, vb); ++k) gu[(k - Td) * JPAIRS] -= 1;
                    for (int k = Td; k <= min(va, vb - 1); ++k) gx[(k - Td) * JPAIRS] -= 1;
                    for (int k = Td; k <= min(va - 1, vb); ++k) gy[(k - Td) * JPAIRS] -= 1;
                }
                if (xa <= xb) ea = ++ia < SPARSE_C ? la[ia] : 0u;
                if (xb <= xa) eb = ++ib < SPARSE_C ? lb[ib] : 0u;
            }
        }
        const int amin = a.smin[i], amax = a.smax[i], bmin = a.smin[j], bmax = a.smax[j];
        // union histogram -> cABX (always the MLE, hll.h:658)
        PairCounts<GT> cu{G + pair, m, lo, hi, max(amax, bmax), JPAIRS, 64};
        const double cABX = ertl_mle(cu, a.p, q, max(amin, bmin), max(amax, bmax));
        // tile families: X = max(A, B-1), Y = max(A-1, B); bins above q-1 fold into bin q (hll.h:660-674)
        PairCounts<GT> cx{G + (size_t)Kcap * JPAIRS + pair, m, lo, hi, max(amax, bmax - 1), JPAIRS, q};
        PairCounts<GT> cy{G + (size_t)2 * Kcap * JPAIRS + pair, m, lo, hi, max(amax - 1, bmax), JPAIRS, q};
        const double eX = ertl_mle(cx, a.p, q - 1, max(amin, bmin - 1) < 0 ? 0 : max(amin, bmin - 1), max(amax, bmax - 1));
        const double eY = ertl_mle(cy, a.p, q - 1, max(amin - 1, bmin) < 0 ? 0 : max(amin - 1, bmin), max(amax - 1, bmax));
        // lhs / rhs of result_cmp: lhs = A (row sketch) unless lhs_is_b
        const double cAX = lhs_is_b ? a.card[j] : a.card[i], cBX = lhs_is_b ? a.card[i] : a.card[j];
        const double cAXBhalf = lhs_is_b ? eY : eX, cBXAhalf = lhs_is_b ? eX : eY;
        const double t0 = cABX - cBX, t1 = cABX - cAX;
        const double cX1 = 1.5 * cBX + 1.5 * cAX - cBXAhalf - cAXBhalf;
        const double cX2 = 2. * (cBXAhalf + cAXBhalf) - 3. * cABX;
        const double h = 0.5 * (cX1 + cX2);
        const double t2 = 0. < h ? h : 0.;
        const double ji = t2 / (t0 + t1 + t2);   // hll.h:1175-1178
        // DB200_UNION_SIZE (extension): union_size under the joint MLE is the sum of the triple, hll.h:1139-1140
        a.out[oidx] = a.rtype == DB200_UNION_SIZE ? (float)(t0 + t1 + t2) : emit_value(a.rtype, ji, t0, t1, t2, a.ksinv);
    }
}

// =============================================================================================
// k nearest neighbours (perform_nns / lock_update / lockfree_update, src/sketch_and_cmp.h:605-697)
// =============================================================================================
// The reference keeps, per sketch, a heap of `nneighbors` (value, index) pairs, visits the other sketches and
// replaces the heap top (the retained pair that sorts LAST under std::less / std::greater on the pair) whenever the
// new value is STRICTLY better than the top's value; the rows are sorted at the end.  With one thread (and always in
// the -Q/-F mode) the visiting order is ascending index, which fixes which of several equal values at the cut survive;
// with more threads the reference itself is schedule dependent there.  This kernel replays the ascending order.
//
// A pair is held as one 64-bit key whose unsigned order is the final output order (smaller key = sorts first):
//   distance measures   key =  (ord(value) << 32 | index)       ascending (value, index)      std::less
//   similarity measures key = ~(ord(value) << 32 | index)       descending (value, index)     std::greater
// ord() is the usual order-preserving float -> uint32 map (after -0 -> +0, so that equal floats tie on the index).
__device__ __forceinline__ uint32_t nn_ord(float v) {
    const uint32_t u = __float_as_uint(v + 0.0f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float nn_unord(uint32_t o) {
    return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
__device__ __forceinline__ uint64_t nn_key(float v, uint32_t idx, int sim) {
    const uint64_t k = ((uint64_t)nn_ord(v) << 32) | idx;
    return sim ? ~k : k;
}
// empty slot: (+-FLT_MAX, uint32(-1)), src/sketch_and_cmp.h:652-653
__device__ __forceinline__ uint64_t nn_empty(int sim) { return nn_key(sim ? -3.402823466e+38f : 3.402823466e+38f, 0xFFFFFFFFu, sim); }

__global__ void knn_init_kernel(uint64_t *__restrict__ keys, uint64_t total, int sim) {
    const uint64_t e = nn_empty(sim);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) keys[i] = e;
}

struct KnnArgs {
    const float *vals;        // symmetric: rows [rb, re) of the packed upper triangle, from vals[0]; rect: [nq_block][nr]
    uint64_t *keys;           // [rows][nn] retained pairs (state across row blocks)
    uint64_t n;               // symmetric: sketches; rect: unused
    uint64_t rb, re;          // symmetric: rows held in vals
    uint64_t nr;              // rect: references
    uint64_t q0, nq;          // rect: first query of this block, queries in this block
    uint32_t nn;
    int sim, rect;
};

constexpr int KNN_WARPS = 4;

// One warp per sketch (symmetric: every sketch r >= rb, which sees sources i in [rb, min(re, r)) — gathered down the
// column — and then, if its own row is in the block, j in (r, n) — contiguous —; rect: one query, j in [0, nr)).
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_update_kernel(KnnArgs a) {
    extern __shared__ uint64_t knn_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t w = (uint64_t)blockIdx.x * KNN_WARPS + warp;
    uint64_t *S = knn_smem + (size_t)warp * a.nn;
    uint64_t row;          // index into keys
    if (a.rect) { if (w >= a.nq) return; row = a.q0 + w; }
    else { row = a.rb + w; if (row >= a.n) return; }
    uint64_t *gk = a.keys + row * a.nn;
    for (uint32_t e = lane; e < a.nn; e += 32) S[e] = gk[e];
    __syncwarp();
    uint64_t wkey; uint32_t wpos;
    auto find_worst = [&]() {
        uint64_t bk = 0; uint32_t bp = 0;
        for (uint32_t e = lane; e < a.nn; e += 32) { const uint64_t k = S[e]; if (k >= bk) { bk = k; bp = e; } }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const uint64_t ok = __shfl_xor_sync(0xFFFFFFFFu, bk, off);
            const uint32_t op = __shfl_xor_sync(0xFFFFFFFFu, bp, off);
            if (ok > bk || (ok == bk && op > bp)) { bk = ok; bp = op; }
        }
        wkey = bk; wpos = bp;
    };
    find_worst();
    bool dirty = false;
    auto consider = [&](float v, uint32_t idx, bool valid) {
        // strictly better VALUE than the worst retained one (the index plays no part in admission)
        const uint32_t hi = a.sim ? ~nn_ord(v) : nn_ord(v);
        uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid && v == v && hi < (uint32_t)(wkey >> 32));
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t hl = __shfl_sync(0xFFFFFFFFu, hi, l);
            const uint32_t il = __shfl_sync(0xFFFFFFFFu, idx, l);
            if (hl < (uint32_t)(wkey >> 32)) {
                if (lane == 0) S[wpos] = ((uint64_t)hl << 32) | (a.sim ? ~il : il);
                __syncwarp();
                find_worst();
                dirty = true;
            }
        }
    };
    if (a.rect) {
        const float *src = a.vals + w * a.nr;
        for (uint64_t j0 = 0; j0 < a.nr; j0 += 32) {
            const uint64_t j = j0 + lane;
            const bool ok = j < a.nr;
            consider(ok ? __ldg(src + j) : 0.f, (uint32_t)j, ok);
        }
    } else {
        const uint64_t n = a.n, r = row;
        auto tri = [n](uint64_t i) { return (i * (2 * n - i - 1)) / 2; };
        const uint64_t base = tri(a.rb);
        const uint64_t iend = a.re < r ? a.re : r;
        for (uint64_t i0 = a.rb; i0 < iend; i0 += 32) {
            const uint64_t i = i0 + lane;
            const bool ok = i < iend;
            consider(ok ? __ldg(a.vals + (tri(i) - base + (r - i - 1))) : 0.f, (uint32_t)i, ok);
        }
        if (r < a.re) {
            const float *src = a.vals + (tri(r) - base);
            for (uint64_t j0 = r + 1; j0 < n; j0 += 32) {
                const uint64_t j = j0 + lane;
                const bool ok = j < n;
                consider(ok ? __ldg(src + (j - r - 1)) : 0.f, (uint32_t)j, ok);
            }
        }
    }
    if (dirty) {
        __syncwarp();
        for (uint32_t e = lane; e < a.nn; e += 32) gk[e] = S[e];
    }
}

struct Neighbor { float value; uint32_t index; };   // std::pair<float, uint32_t> (validx_t, src/sketch_and_cmp.h:605)

// Final per-row sort (std::sort with std::less / std::greater, :692-696) by ranking, and decode to (value, index).
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_sort_kernel(const uint64_t *__restrict__ keys, uint64_t rows, uint32_t nn, int sim,
                                                                 Neighbor *__restrict__ out) {
    extern __shared__ uint64_t knn_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t row = (uint64_t)blockIdx.x * KNN_WARPS + warp;
    if (row >= rows) return;
    uint64_t *S = knn_smem + (size_t)warp * nn;
    for (uint32_t e = lane; e < nn; e += 32) S[e] = keys[row * nn + e];
    __syncwarp();
    for (uint32_t e = lane; e < nn; e += 32) {
        const uint64_t k = S[e];
        uint32_t rank = 0;
        for (uint32_t f = 0; f < nn; ++f) { const uint64_t o = S[f]; rank += (o < k) || (o == k && f < e); }
        const uint64_t raw = sim ? ~k : k;
        out[row * nn + rank] = Neighbor{nn_unord((uint32_t)(raw >> 32)), (uint32_t)raw};
    }
}

} // namespace db200
