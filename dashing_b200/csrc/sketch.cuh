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
__device__ __forceinline__ uint64_t wang64(uint64_t key) {
    key = ~key + (key << 21);
    key ^= key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key ^= key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key ^= key >> 28;
    key += key << 31;
    return key;
}

// Emission mask of one 64-base block: bit j set iff the k bases ending at base j are all valid and
// none but the first is a record start.  prev_* are the planes of the preceding block.
__device__ __forceinline__ uint64_t emit_mask(uint64_t prev_nb, uint64_t prev_st, uint64_t nb, uint64_t st, int k) {
    // 128-bit window: low = previous block, high = this block
    uint64_t clo = prev_nb & ~prev_st, chi = nb & ~st;  // "continuation" bits
    uint64_t alo = ~0ull, ahi = ~0ull;                  // run-AND of cont over the last (k-1) positions
    int L = k - 1;
    if (L > 0) {
        // binary lifting: r = AND of cont over a window of length `len`
        uint64_t rlo = clo, rhi = chi;
        int len = 1;
        alo = ahi = ~0ull;
        int done = 0;  // positions already covered in (alo,ahi), counted back from the current one
        while (L) {
            if (L & 1) {
                // a &= r << done
                uint64_t slo, shi;
                if (done == 0) { slo = rlo; shi = rhi; }
                else { shi = (rhi << done) | (rlo >> (64 - done)); slo = rlo << done; }
                alo &= slo; ahi &= shi;
                done += len;
            }
            L >>= 1;
            if (L) {
                // r &= r << len
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

// ---------------------------------------------------------------------------------------------
// The sketch kernel.  256 threads; each thread owns runs of BPT consecutive 64-base blocks.
// SMEM_REGS: registers of the item's genome are staged in shared memory (2^p bytes) and merged into
// HBM once per item; otherwise (2^p too large) updates go straight to HBM/L2.
// ---------------------------------------------------------------------------------------------
constexpr int SK_THREADS = 256;
constexpr int SK_BPT = 4;

template <bool SMEM_REGS>
__global__ void __launch_bounds__(SK_THREADS) sketch_kernel(const uint4 *__restrict__ bases2, const uint64_t *__restrict__ nbp,
                                                            const uint64_t *__restrict__ stp,
                                                            const SketchItem *__restrict__ items, uint32_t nitems, int k,
                                                            int p, int canon, uint8_t *__restrict__ regs) {
    extern __shared__ __align__(16) uint8_t sregs[];
    const uint32_t m = 1u << p;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    const int rcshift = 2 * (k - 1);
    const int idx_shift = 64 - p;

    for (uint32_t it = blockIdx.x; it < nitems; it += gridDim.x) {
        const SketchItem item = items[it];
        uint8_t *greg = regs + (uint64_t)item.genome * m;
        if (SMEM_REGS) {
            for (uint32_t w = threadIdx.x; w < m / 16; w += SK_THREADS) reinterpret_cast<uint4 *>(sregs)[w] = make_uint4(0, 0, 0, 0);
            __syncthreads();
        }
        uint8_t *r8 = SMEM_REGS ? sregs : greg;

        const uint64_t blk_lo = item.pos_begin >> 6, blk_hi = (item.pos_end + 63) >> 6;
        for (uint64_t b0 = blk_lo + (uint64_t)threadIdx.x * SK_BPT; b0 < blk_hi; b0 += (uint64_t)SK_THREADS * SK_BPT) {
            uint64_t fwd = 0, rc = 0, prev_nb = 0, prev_st = 0;
            if (b0 > 0) {
                // warm the rolling k-mers with the last 32 bases of the preceding block
                const uint4 pw = __ldg(bases2 + b0 - 1);
                prev_nb = __ldg(nbp + b0 - 1);
                prev_st = __ldg(stp + b0 - 1);
                const uint32_t pws[2] = {pw.z, pw.w};
#pragma unroll
                for (int wi = 0; wi < 2; ++wi) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t c = (pws[wi] >> (2 * i)) & 3u;
                        fwd = ((fwd << 2) | c) & kmask;
                        rc = (rc >> 2) | ((uint64_t)(c ^ 3u) << rcshift);
                    }
                }
            }
            const uint64_t b1 = (b0 + SK_BPT < blk_hi) ? b0 + SK_BPT : blk_hi;
            for (uint64_t b = b0; b < b1; ++b) {
                const uint4 w4 = __ldg(bases2 + b);
                const uint64_t nb = __ldg(nbp + b), st = __ldg(stp + b);
                uint64_t E = emit_mask(prev_nb, prev_st, nb, st, k);
                // clip to the item's position range
                const uint64_t base = b << 6;
                if (item.pos_begin > base) E &= (item.pos_begin - base >= 64) ? 0ull : (~0ull << (item.pos_begin - base));
                if (item.pos_end < base + 64) E &= (item.pos_end <= base) ? 0ull : (~0ull >> (64 - (item.pos_end - base)));
                prev_nb = nb; prev_st = st;
#pragma unroll 1
                for (int wi = 0; wi < 4; ++wi) {
                    const uint32_t word = wi == 0 ? w4.x : wi == 1 ? w4.y : wi == 2 ? w4.z : w4.w;
                    const uint32_t e16 = (uint32_t)(E >> (16 * wi)) & 0xFFFFu;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t c = (word >> (2 * i)) & 3u;
                        fwd = ((fwd << 2) | c) & kmask;
                        rc = (rc >> 2) | ((uint64_t)(c ^ 3u) << rcshift);
                        if ((e16 >> i) & 1u) {
                            const uint64_t x = (canon && rc < fwd) ? rc : fwd;
                            const uint64_t h = wang64(x);
                            const uint32_t idx = (uint32_t)(h >> idx_shift);
                            const uint32_t rho = (uint32_t)__clzll((long long)(((h << 1) | 1ull) << (p - 1))) + 1u;
                            if (r8[idx] < rho) byte_max(reinterpret_cast<uint32_t *>(r8) + (idx >> 2), (idx & 3u) * 8u, rho);
                        }
                    }
                }
            }
        }
        if (SMEM_REGS) {
            __syncthreads();
            uint32_t *g32 = reinterpret_cast<uint32_t *>(greg);
            const uint32_t *s32 = reinterpret_cast<const uint32_t *>(sregs);
            for (uint32_t w = threadIdx.x; w < m / 4; w += SK_THREADS) {
                const uint32_t v = s32[w];
                if (v) {
                    uint32_t old = g32[w];
                    for (;;) {
                        const uint32_t nw = __vmaxu4(old, v);
                        if (nw == old) break;
                        const uint32_t prev = atomicCAS(g32 + w, old, nw);
                        if (prev == old) break;
                        old = prev;
                    }
                }
            }
            __syncthreads();
        }
    }
}

} // namespace db200
