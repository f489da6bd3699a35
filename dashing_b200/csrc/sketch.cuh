// sketch.cuh — hot path (i): canonical k-mer wang_hash + HLL register update.
//
// Replaces, for the GPU-eligible configuration (DNA4, unspaced, unwindowed, k <= 32, WangHash):
//   Encoder::for_each_uncanon_unspaced_unwindowed   bonsai/include/bonsai/encoder.h:240-271
//   canonical_representation / reverse_complement   bonsai/include/bonsai/kmerutil.h:83-90, :137-140
//   WangHash::operator()                             bonsai/hll/include/sketch/hash.h:40-49
//   hllbase_t::add                                   bonsai/hll/include/sketch/hll.h:828-836
//
// HBM-resident input format ("packed genome store", one per batch of genomes):
//   bases2 : uint4[nblk]     64 bases per 16-byte word, base j of a block at bits [2j, 2j+1] of the
//                            128-bit little-endian word (A=0 C=1 G=2 T=3)
//   nb     : uint64[nblk]    bit j = base j is one of ACGTacgt
//   st     : uint64[nblk]    bit j = base j is the first base of a FASTA record (windows never span records)
// Work items: (genome, [pos_begin,pos_end)) — the CTA emits every k-mer whose LAST base lies in the
// range into a shared-memory copy of that genome's registers, then max-merges it into HBM.
#pragma once
#include "common.cuh"

namespace db200 {

struct SketchItem {
    uint64_t pos_begin;  // first k-mer end position owned by this item (absolute base index in the store)
    uint64_t pos_end;    // one past the last
    uint32_t genome;     // row of the register matrix
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------
// ASCII -> 2-bit + validity.  One thread per 16 bases: one coalesced 16-byte load, one 4-byte
// store of codes, one 2-byte store of validity bits.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack4(uint32_t x, uint32_t &code8, uint32_t &valid4) {
    const uint32_t up = x & 0xDFDFDFDFu;  // fold case
    const uint32_t v = (__vcmpeq4(up, 0x41414141u) | __vcmpeq4(up, 0x43434343u) |
                        __vcmpeq4(up, 0x47474747u) | __vcmpeq4(up, 0x54545454u)) & 0x01010101u;
    uint32_t c = ((x >> 1) ^ (x >> 2)) & 0x03030303u;  // A->0 C->1 G->2 T->3
    c |= c >> 6;
    c |= c >> 12;
    code8 = c & 0xFFu;
    uint32_t t = v | (v >> 7);
    t |= t >> 14;
    valid4 = t & 0xFu;
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ ascii, uint64_t nbases,
                                                   uint32_t *__restrict__ codes, uint16_t *__restrict__ valid,
                                                   uint64_t ngroups /* = nblk * 4 */) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const uint64_t base0 = g * 16;
    uint32_t w[4] = {0, 0, 0, 0};
    if (base0 + 16 <= nbases) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(ascii + base0));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else if (base0 < nbases) {
        for (uint64_t i = base0; i < nbases; ++i) w[(i - base0) >> 2] |= (uint32_t)ascii[i] << (8 * ((i - base0) & 3));
    }
    uint32_t code = 0, val = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t c8, v4;
        pack4(w[i], c8, v4);
        code |= c8 << (8 * i);
        val |= v4 << (4 * i);
    }
    codes[g] = code;
    valid[g] = (uint16_t)val;
}

// record-start bit-plane: one thread per record
__global__ void mark_starts_kernel(const uint64_t *__restrict__ starts, uint64_t nstarts, uint32_t *__restrict__ st32) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nstarts) return;
    const uint64_t pos = starts[i];
    atomicOr(&st32[pos >> 5], 1u << (pos & 31));
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
// Runtime "constants" handed to the kernel as arguments.  They exist only so that ptxas cannot strength-reduce
// the multiplications below into shift/LEA sequences: the ALU pipe (shifts, logic, compares) is the binding
// unit of this kernel while the FMA pipe (IMAD) is otherwise idle, so every op moved across is a win.
struct SketchConsts {
    uint32_t four;      // 4
    uint32_t neg1;      // 0xFFFFFFFF
    uint32_t rc_mul_lo; // 2^rcshift        if rcshift <  32 else 0
    uint32_t rc_mul_hi; // 2^(rcshift-32)   if rcshift >= 32 else 0
    uint32_t rho_mul;   // 2^p
    uint32_t rho_add;   // 2^(p-1)
    uint32_t neg_2p20;  // -(2^20)
    uint32_t c1087_2p20; // 1087 << 20
};

// r = a * c + add (64-bit a, 32-bit c): IMAD.WIDE.U32 + IMAD
__device__ __forceinline__ uint64_t mad64c(uint64_t a, uint32_t c, uint64_t add) {
    uint64_t r;
    asm("{\n\t.reg .u32 lo, ahi, rlo, rhi;\n\t"
        "mov.b64 {lo, ahi}, %1;\n\t"
        "mad.wide.u32 %0, lo, %2, %3;\n\t"
        "mov.b64 {rlo, rhi}, %0;\n\t"
        "mad.lo.u32 rhi, ahi, %2, rhi;\n\t"
        "mov.b64 %0, {rlo, rhi};\n\t}"
        : "=l"(r)
        : "l"(a), "r"(c), "l"(add));
    return r;
}

// Thomas Wang's 64-bit mix (bonsai/hll/include/sketch/hash.h:40-49) with the shift-add steps written as
// multiplications: ~k + (k<<21) = k*(2^21-1) - 1, k + (k<<3) + (k<<8) = 265k, k + (k<<2) + (k<<4) = 21k,
// k + (k<<31) = (2^31+1)k  (all mod 2^64).  `ones` must be 2^64-1 (see SketchConsts).
__device__ __forceinline__ uint64_t wang64(uint64_t key, uint64_t ones) {
    key = mad64c(key, 0x1FFFFFu, ones);
    key ^= key >> 24;
    key = mad64c(key, 265u, 0ull);
    key ^= key >> 14;
    key = mad64c(key, 21u, 0ull);
    key ^= key >> 28;
    key = mad64c(key, 0x80000001u, 0ull);
    return key;
}

// Emission mask of one 64-base block: bit j set iff the k bases ending at base j are all valid and
// none but the first is a record start.  prev_* are the planes of the preceding block.
__device__ __forceinline__ uint64_t emit_mask(uint64_t prev_nb, uint64_t prev_st, uint64_t nb, uint64_t st, int k) {
    // 128-bit window: low = previous block, high = this block
    const uint64_t clo = prev_nb & ~prev_st, chi = nb & ~st;  // "continuation" bits
    uint64_t alo = ~0ull, ahi = ~0ull;                        // run-AND of cont over the last (k-1) positions
    int L = k - 1;
    if (L > 0) {
        // binary lifting: r = AND of cont over a window of length `len`
        uint64_t rlo = clo, rhi = chi;
        int len = 1;
        int done = 0;  // positions already covered in (alo,ahi), counted back from the current one
        while (L) {
            if (L & 1) {
                uint64_t slo, shi;
                if (done == 0) { slo = rlo; shi = rhi; }
                else { shi = (rhi << done) | (rlo >> (64 - done)); slo = rlo << done; }
                alo &= slo; ahi &= shi;
                done += len;
            }
            L >>= 1;
            if (L) {
                const uint64_t shi = (rhi << len) | (rlo >> (64 - len));
                const uint64_t slo = rlo << len;
                rlo &= slo; rhi &= shi;
                len <<= 1;
            }
        }
    }
    // first base of the window must be valid (it may be a record start)
    const int s = k - 1;
    const uint64_t fhi = s ? ((nb << s) | (prev_nb >> (64 - s))) : nb;
    return ahi & fhi;
}

// CAS-based byte max on a packed register word (shared or global).
__device__ __forceinline__ void byte_max(uint32_t *word, uint32_t shift, uint32_t rho) {
    uint32_t old = *reinterpret_cast<volatile uint32_t *>(word);
    while (((old >> shift) & 0xFFu) < rho) {
        const uint32_t nw = (old & ~(0xFFu << shift)) | (rho << shift);
        const uint32_t prev = atomicCAS(word, old, nw);
        if (prev == old) break;
        old = prev;
    }
}

// element-wise byte max of a packed word into HBM/L2
__device__ __forceinline__ void merge_word(uint32_t *g, uint32_t v) {
    uint32_t old = *g;
    for (;;) {
        const uint32_t nw = __vmaxu4(old, v);
        if (nw == old) break;
        const uint32_t prev = atomicCAS(g, old, nw);
        if (prev == old) break;
        old = prev;
    }
}

// ---------------------------------------------------------------------------------------------
// The sketch kernel.  Persistent CTAs of 512 threads pull work items from an atomic counter.  Inside an
// item every warp sweeps warp-tiles of 32 x BPT consecutive 64-base blocks with WARP-UNIFORM loop bounds
// (lanes past the end run on an all-invalid guard block), so the only divergent code is the rare
// register update.
//   MODE 0: registers staged in shared memory as one 32-bit word each -> a single ATOMS.MAX, no retry loop
//           (2^p * 4 bytes: p <= 15)
//   MODE 1: registers staged in shared memory as bytes, CAS loop + explicit warp reconvergence (p <= 17)
//   MODE 2: no staging, CAS on HBM/L2 directly (larger p)
// The staged copy is seeded from the registers already merged into HBM by earlier items of the same
// genome, which keeps the update rate (and the ATOMS traffic) low.
// ---------------------------------------------------------------------------------------------
constexpr int SK_THREADS = 512;  // 16 warps share one staged sketch: 3 CTAs/SM at p=14 -> 48 resident warps
constexpr int SK_BPT = 4;

// KCLASS: 0 -> k <= 16 (k-mer lives in the low word), 1 -> 16 < k < 32, 2 -> k == 32 (no mask at all)
template <int MODE, int KCLASS, bool CANON>
__global__ void __launch_bounds__(SK_THREADS) sketch_kernel(const uint4 *__restrict__ bases2, const uint64_t *__restrict__ nbp,
                                                            const uint64_t *__restrict__ stp,
                                                            const SketchItem *__restrict__ items, uint32_t nitems, uint64_t guard_blk,
                                                            int k, int p, uint8_t *__restrict__ regs,
                                                            uint32_t *__restrict__ item_counter, const SketchConsts kc) {
    extern __shared__ __align__(16) uint8_t sregs[];
    __shared__ uint32_t s_item;
    const uint32_t m = 1u << p;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    const uint32_t mask_lo = (uint32_t)kmask, mask_hi = (uint32_t)(kmask >> 32);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t ones = ((uint64_t)kc.neg1 << 32) | kc.neg1;
    const int idx_shift = 32 - p;   // p <= 24
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sregs);
    constexpr uint32_t NWARPS = SK_THREADS / 32;

    // one base: roll the forward and reverse-complement k-mers
#define DB200_ROLL(c)                                                                                       \
    do {                                                                                                    \
        /* fwd = (fwd << 2) | c: three IMADs, no carry can leave the low word because c < 4 */              \
        const uint32_t flo = (uint32_t)fwd, fhi = (uint32_t)(fwd >> 32);                                    \
        const uint32_t nlo = flo * kc.four + (c);                                                           \
        uint32_t nhi = fhi * kc.four + __umulhi(flo, kc.four);                                              \
        if (KCLASS == 0) fwd = nlo & mask_lo;                                                               \
        else { if (KCLASS == 1) nhi &= mask_hi; fwd = ((uint64_t)nhi << 32) | nlo; }                        \
        /* rc = (rc >> 2) | ((3 - c) << 2(k-1)) */                                                          \
        const uint32_t cc = (c) * kc.neg1 + 3u;                                                             \
        const uint32_t rlo = __funnelshift_r((uint32_t)rc, (uint32_t)(rc >> 32), 2);                        \
        const uint32_t rhi = (uint32_t)(rc >> 32) >> 2;                                                     \
        rc = ((uint64_t)(cc * kc.rc_mul_hi + rhi) << 32) | (uint32_t)(cc * kc.rc_mul_lo + rlo);             \
    } while (0)

    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(item_counter, 1u);
        __syncthreads();
        const uint32_t it = s_item;
        if (it >= nitems) break;
        const SketchItem item = items[it];
        uint8_t *greg = regs + (uint64_t)item.genome * m;
        uint32_t *g32 = reinterpret_cast<uint32_t *>(greg);
        // seed the staged registers from HBM
        if (MODE == 0) {
            // staged in the EXPONENT domain: a slot holds (biased exponent << 20) of the smallest double(T) seen for its register
            // (rho = 1087 - biased exponent, so a smaller word is a larger rho); 0xFFFFFFFF = empty.
            uint4 *s4 = reinterpret_cast<uint4 *>(sregs);
            for (uint32_t w = threadIdx.x; w < m / 4; w += SK_THREADS) {
                const uint32_t v = g32[w];
                uint32_t e[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t r = (v >> (8 * i)) & 0xFFu;
                    e[i] = r ? ((1087u - r) << 20) : 0xFFFFFFFFu;
                }
                s4[w] = make_uint4(e[0], e[1], e[2], e[3]);
            }
        } else if (MODE == 1) {
            for (uint32_t w = threadIdx.x; w < m / 4; w += SK_THREADS) reinterpret_cast<uint32_t *>(sregs)[w] = g32[w];
        }
        __syncthreads();
        uint8_t *r8 = MODE == 1 ? sregs : greg;

        const uint64_t blk_lo = item.pos_begin >> 6, blk_hi = (item.pos_end + 63) >> 6;
        for (uint64_t wb = blk_lo + (uint64_t)warp * (32 * SK_BPT); wb < blk_hi; wb += (uint64_t)NWARPS * 32 * SK_BPT) {
            const uint64_t b0 = wb + (uint64_t)lane * SK_BPT;
            uint64_t fwd = 0, rc = 0;
            // warm the rolling k-mers with the last 32 bases of the preceding block
            const uint64_t pb = (b0 > 0 && b0 < blk_hi) ? b0 - 1 : guard_blk;
            const uint4 pw = __ldg(bases2 + pb);
            uint64_t prev_nb = __ldg(nbp + pb), prev_st = __ldg(stp + pb);
#pragma unroll 1
            for (int wi = 0; wi < 2; ++wi) {
                const uint32_t word = wi == 0 ? pw.z : pw.w;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t c = (word >> (2 * i)) & 3u;
                    DB200_ROLL(c);
                }
            }
#pragma unroll 1
            for (int j = 0; j < SK_BPT; ++j) {
                const uint64_t b = b0 + j;
                const bool valid = b < blk_hi;
                const uint64_t bl = valid ? b : guard_blk;
                const uint4 w4 = __ldg(bases2 + bl);
                const uint64_t nb = __ldg(nbp + bl), st = __ldg(stp + bl);
                uint64_t E = emit_mask(prev_nb, prev_st, nb, st, k);
                // clip to the item's position range
                const uint64_t base = b << 6;
                if (item.pos_begin > base) E &= (item.pos_begin - base >= 64) ? 0ull : (~0ull << (item.pos_begin - base));
                if (item.pos_end < base + 64) E &= (item.pos_end <= base) ? 0ull : (~0ull >> (64 - (item.pos_end - base)));
                if (!valid) E = 0;
                prev_nb = nb; prev_st = st;
#pragma unroll 1
                for (int wi = 0; wi < 4; ++wi) {
                    const uint32_t word = wi == 0 ? w4.x : wi == 1 ? w4.y : wi == 2 ? w4.z : w4.w;
                    const uint32_t e16 = (uint32_t)(E >> (16 * wi)) & 0xFFFFu;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t c = (word >> (2 * i)) & 3u;
                        DB200_ROLL(c);
                        const uint64_t x = (CANON && rc < fwd) ? rc : fwd;
                        const uint64_t h = wang64(x, ones);
                        const uint32_t hlo = (uint32_t)h, hhi = (uint32_t)(h >> 32);
                        const uint32_t idx = hhi >> idx_shift;
                        // rho = clz(((h << 1) | 1) << (p - 1)) + 1 (hll.h:830) = 64 - floor(log2 T), T = (h << p) | 2^(p-1) != 0.
                        // One u64 -> f64 conversion (round toward zero, so the exponent is exact) replaces clz of a 64-bit
                        // value with its rare "upper word is zero" branch.
                        const uint32_t tlo = hlo * kc.rho_mul + kc.rho_add;
                        const uint32_t thi = __funnelshift_l(hlo, hhi, p);
                        const double dT = __ull2double_rz(((uint64_t)thi << 32) | tlo);
                        const uint32_t ehi = (uint32_t)__double2hiint(dT);   // biased exponent in bits 20..30
                        const bool emit = (e16 >> i) & 1u;
                        if (MODE == 0) {
                            // staged registers are addressed in the shared window directly: no generic-address setup per k-mer
                            const uint32_t addr = sbase + idx * 4u;
                            uint32_t cur;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(addr));
                            // exponent domain: slots hold (biased exponent << 20) of the best T so far, so `ehi < cur` is exactly
                            // "strictly larger rho" (equal exponents: ehi >= cur because of its mantissa bits) — one compare, and
                            // on the rare path one mask + RED.MIN; no rho arithmetic per k-mer
                            if (emit && ehi < cur)
                                asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(addr), "r"(ehi & 0xFFF00000u) : "memory");
                        } else {
                            const uint32_t rho = 1087u - (ehi >> 20);
                            const bool upd = emit && r8[idx] < rho;
                            if (__any_sync(0xFFFFFFFFu, upd)) {
                                if (upd) byte_max(reinterpret_cast<uint32_t *>(r8) + (idx >> 2), (idx & 3u) * 8u, rho);
                                __syncwarp();  // the CAS loop is not a forced reconvergence point: re-join here
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (MODE == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(sregs);
            for (uint32_t w = threadIdx.x; w < m / 4; w += SK_THREADS) {
                const uint4 v = s4[w];
                const uint32_t e[4] = {v.x, v.y, v.z, v.w};
                uint32_t packed = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) packed |= (e[i] == 0xFFFFFFFFu ? 0u : 1087u - (e[i] >> 20)) << (8 * i);   // back to rho
                merge_word(g32 + w, packed);
            }
        } else if (MODE == 1) {
            const uint32_t *s32 = reinterpret_cast<const uint32_t *>(sregs);
            for (uint32_t w = threadIdx.x; w < m / 4; w += SK_THREADS) merge_word(g32 + w, s32[w]);
        }
        __syncthreads();
    }
#undef DB200_ROLL
}

} // namespace db200
