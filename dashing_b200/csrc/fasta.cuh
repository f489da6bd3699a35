// fasta.cuh — device-side FASTA parsing: raw file text in (header lines, newlines, CR), the packed genome
// store of sketch.cuh out.  Replaces, for plain FASTA, the host's kseq_read record loop (bonsai/klib/kseq.h:177-218) that
// feeds Encoder::for_each (bonsai/include/bonsai/encoder.h:509-529) — the host only has to get the file's bytes (read() or
// gz inflate) into memory.
//
// kseq's rules, restated per byte (a "line start" is the first byte of a file or the byte after a '\n'):
//   * a line that starts with '>' is a header: it opens a new record and none of its bytes are sequence (kseq.h:183-193);
//   * every other byte of a line that follows a header is sequence, except '\n' and a '\r' directly before a '\n'
//     (ks_getuntil2 strips one trailing CR, kseq.h:140-141) — including blanks, digits, '>' in mid-line, ...: they are
//     invalid bases, exactly what the reference's k-mer loop sees (encoder.h:252-253);
//   * bytes before the first header of a file belong to no record (kseq.h:181-186);
//   * a sequence line that starts with '+' or '@' switches kseq to FASTQ parsing (kseq.h:196-216), and before a file's first
//     header kseq scans characters rather than lines for '>' / '@' (kseq.h:183): files showing either are not handled here —
//     the file is flagged and the host sketches it through the record interface instead.
// Windows never span records (encoder.h:444): the first sequence byte of every record is marked in the record-start plane.
//
// States: SKIP (before a file's first header) / HDR (inside a header line) / SEQN (after a header, no sequence byte yet)
// / SEQ.  A run of bytes maps an incoming state to (outgoing state, number of sequence bytes); these maps compose, so
// blocks are summarised independently (fa_summary_kernel), chained by a tiny sequential pass (fa_chain_kernel) and then
// emitted (fa_emit_kernel) with every byte knowing its state and its output position.
#pragma once
#include "common.cuh"
#include "sketch.cuh"
#include "fasta_logic.h"

namespace db200 {

constexpr int FA_THREADS = 512;                 // threads per CTA
constexpr int FA_BPT = 16;                      // bytes per thread: one 16-byte load
constexpr int FA_BLOCK = FA_THREADS * FA_BPT;   // 8 KiB of text per CTA; files start on multiples of this

// Transfer summary of a run of bytes: bits [2s, 2s+2) = state after the run when entered in state s;
// bits [8 + 14 s, 8 + 14 (s+1)) = sequence bytes the run yields when entered in state s (<= 8192 per block).
typedef uint64_t FaSum;
__device__ __forceinline__ uint32_t fa_tf(FaSum a, uint32_t s) { return (uint32_t)(a >> (2 * s)) & 3u; }
__device__ __forceinline__ uint32_t fa_cnt(FaSum a, uint32_t s) { return (uint32_t)(a >> (8 + 14 * s)) & 0x3FFFu; }
__device__ __forceinline__ FaSum fa_compose(FaSum a, FaSum b) {   // a first, then b
    FaSum r = 0;
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {
        const uint32_t t = fa_tf(a, s);
        r |= (FaSum)fa_tf(b, t) << (2 * s);
        r |= (FaSum)(fa_cnt(a, s) + fa_cnt(b, t)) << (8 + 14 * s);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// Byte classes of a lane's 16 bytes as 16-bit masks (SIMD within the four 32-bit words of one 16-byte load).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fa_eq_nibble(uint32_t w, uint32_t pat4) {
    const uint32_t x = w ^ pat4;
    const uint32_t t = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 in every byte of w equal to the pattern byte
    return ((t >> 7) * 0x01020408u) >> 24;                                      // -> one bit per byte, bits 0..3
}
__device__ __forceinline__ uint32_t fa_eq16(const uint4 &v, uint32_t c) {
    const uint32_t pat = c * 0x01010101u;
    return fa_eq_nibble(v.x, pat) | (fa_eq_nibble(v.y, pat) << 4) | (fa_eq_nibble(v.z, pat) << 8) | (fa_eq_nibble(v.w, pat) << 12);
}

struct FaCls {
    uint4 raw;
    uint32_t nl, cr, gt;       // byte-class masks; bytes past the file's end count as '\n'
    uint32_t prev_nl, next_nl;
};

// Loads the lane's 16 bytes and classifies them.  [file_start, file_end): the file this block belongs to — bytes of the block
// past the file's end (the gap up to the next file's block-aligned start) read as '\n', whatever the caller left there.
// `chunk_end` / `next_byte`: the first byte not yet resident on the device and its value (the host reads it from its own
// copy of the text).  The byte before / after a lane comes from the neighbouring lane; only the warp's edge lanes load it.
__device__ __forceinline__ void fa_classify(const uint8_t *__restrict__ text, uint64_t i0, uint64_t file_start, uint64_t file_end, uint64_t chunk_end,
                                            uint32_t next_byte, FaCls &c) {
    const uint32_t lane = threadIdx.x & 31;
    c.raw = __ldg(reinterpret_cast<const uint4 *>(text + i0));
    const uint32_t nvalid = file_end > i0 ? (uint32_t)min((uint64_t)FA_BPT, file_end - i0) : 0u;
    const uint32_t vm = fa_below(nvalid);
    c.nl = (fa_eq16(c.raw, '\n') & vm) | (~vm & 0xFFFFu);
    c.cr = fa_eq16(c.raw, '\r') & vm;
    c.gt = fa_eq16(c.raw, '>') & vm;
    uint32_t pn = __shfl_up_sync(0xFFFFFFFFu, c.nl >> 15, 1);
    uint32_t nn = __shfl_down_sync(0xFFFFFFFFu, c.nl & 1u, 1);
    if (lane == 0) pn = (i0 > file_start && i0 - 1 < file_end) ? (uint32_t)(__ldg(text + i0 - 1) == '\n') : 1u;
    if (lane == 31) {
        const uint64_t in = i0 + FA_BPT;
        nn = in >= file_end ? 1u : (uint32_t)((in < chunk_end ? (uint32_t)__ldg(text + in) : next_byte) == '\n');
    }
    c.prev_nl = pn; c.next_nl = nn;
}

struct FaWarp { uint32_t b_hs, b_ls, b_kab, b_kb; };
__device__ __forceinline__ FaWarp fa_ballots(const FaLane &L) {
    FaWarp w;
    w.b_hs = __ballot_sync(0xFFFFFFFFu, L.hs != 0);
    w.b_ls = __ballot_sync(0xFFFFFFFFu, L.ls != 0);
    w.b_kab = __ballot_sync(0xFFFFFFFFu, (L.kA | L.kB) != 0);
    w.b_kb = __ballot_sync(0xFFFFFFFFu, L.kB != 0);
    return w;
}
// Incoming state of this lane when the warp is entered in state S (fasta_logic.h), det_out fetched from the lane that holds
// the nearest header start below.
__device__ __forceinline__ uint32_t fa_my_state(uint32_t S, const FaLane &L, const FaWarp &w) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ph = w.b_hs & fa_below(lane);
    const uint32_t det_h = __shfl_sync(0xFFFFFFFFu, L.det_out, ph ? FA_MSB(ph) : 0);
    return fa_lane_state(S, lane, w.b_hs, w.b_ls, w.b_kab, w.b_kb, det_h);
}
__device__ __forceinline__ uint32_t fa_warp_out(uint32_t S, const FaLane &L, const FaWarp &w) {
    const uint32_t det_h = __shfl_sync(0xFFFFFFFFu, L.det_out, w.b_hs ? FA_MSB(w.b_hs) : 0);
    return fa_lane_state(S, 32u, w.b_hs, w.b_ls, w.b_kab, w.b_kb, det_h);
}

// A block's file: files start on block boundaries, `fblk` holds their first block (ascending).
__device__ __forceinline__ uint32_t fa_file_of_block(const uint64_t *__restrict__ fblk, uint32_t nfiles, uint64_t blk) {
    uint32_t lo = 0, hi = nfiles;          // last f with fblk[f] <= blk
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(fblk + mid) <= blk) lo = mid; else hi = mid; }
    return lo;
}

// Pass 1: one summary per block (sums[blk]) and one per warp (wsum[(blk - blk0) * 16 + warp], for pass 3).
// A warp resolves its lanes' states with four ballots (fasta_logic.h), once per possible incoming state of the warp
// (SEQN and SEQ differ in the outgoing state only), and adds the lanes' sequence-byte counts with redux.
__global__ void __launch_bounds__(FA_THREADS) fa_summary_kernel(const uint8_t *__restrict__ text, uint64_t blk0, const uint64_t *__restrict__ fblk,
                                                               const uint64_t *__restrict__ flen, uint32_t nfiles, uint64_t chunk_end, uint32_t next_byte,
                                                               FaSum *__restrict__ sums, FaSum *__restrict__ wsum) {
    constexpr uint32_t NW = FA_THREADS / 32;
    __shared__ FaSum s_warp[NW];
    __shared__ uint64_t s_fstart, s_fend;
    const uint64_t blk = blk0 + blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        const uint32_t f = fa_file_of_block(fblk, nfiles, blk);
        s_fstart = __ldg(fblk + f) * (uint64_t)FA_BLOCK;
        s_fend = s_fstart + __ldg(flen + f);
    }
    __syncthreads();
    FaCls c;
    fa_classify(text, blk * FA_BLOCK + (uint64_t)threadIdx.x * FA_BPT, s_fstart, s_fend, chunk_end, next_byte, c);
    const FaLane L = fa_lane(c.nl, c.cr, c.gt, c.prev_nl, c.next_nl);
    const FaWarp w = fa_ballots(L);
    uint32_t cnt[3];
    const uint32_t Sin[3] = {FS_SKIP, FS_HDR, FS_SEQ};
#pragma unroll
    for (int i = 0; i < 3; ++i) cnt[i] = __reduce_add_sync(0xFFFFFFFFu, FA_POPC(fa_keep(L, fa_my_state(Sin[i], L, w))));
    const uint32_t o_skip = fa_warp_out(FS_SKIP, L, w), o_hdr = fa_warp_out(FS_HDR, L, w), o_seqn = fa_warp_out(FS_SEQN, L, w),
                   o_seq = fa_warp_out(FS_SEQ, L, w);
    const FaSum mine = (FaSum)(o_skip | (o_hdr << 2) | (o_seqn << 4) | (o_seq << 6)) | ((FaSum)cnt[0] << 8) | ((FaSum)cnt[1] << 22) |
                       ((FaSum)cnt[2] << 36) | ((FaSum)cnt[2] << 50);
    if (lane == 0) { s_warp[warp] = mine; wsum[(uint64_t)blockIdx.x * NW + warp] = mine; }
    __syncthreads();
    if (warp == 0) {                   // the block's map: the warps' maps composed in order
        FaSum v = lane < NW ? s_warp[lane] : (FaSum)0xE4ull;
#pragma unroll
        for (int d = 1; d < (int)NW; d <<= 1) {
            const FaSum o = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= (uint32_t)d) v = fa_compose(o, v);
        }
        if (lane == NW - 1) sums[blk] = v;
    }
}

// Pass 2: chain the blocks of one chunk.  carry[0] = state, carry[1] = next output position after the previous chunk.  A file
// start resets the state to SKIP; the first file of a genome also moves the output position to that genome's window
// (gpos0[f] != ~0).  genome_end[g] = one past the last sequence byte written for genome g so far (final once a later genome has
// started or the text has ended).
// One CTA of 1024 threads per chunk.  Thread t owns the consecutive blocks [t * per, (t + 1) * per) (per <= 8 for a 64 MiB chunk):
// it first folds them into one map of the incoming (state, position) — speculating over the four possible incoming states —,
// the maps are scanned (shuffles inside a warp, the 32 warp totals by warp 0), and the thread then walks its blocks again
// with the now known incoming state and position.  File starts reset the state; a genome's first file makes the position
// absolute (its window), which the map records as (has_abs, abs).
constexpr int FA_CHAIN_THREADS = 1024;
struct FaChainMap {
    uint32_t tf;        // 2 bits per incoming state
    uint32_t c[4];      // sequence bytes per incoming state (since the last genome start if has_abs)
    uint32_t has_abs;
    uint64_t abs;       // window start of the last genome started inside the run
};
__device__ __forceinline__ FaChainMap fa_chain_compose(const FaChainMap &a, const FaChainMap &b) {   // a first, then b
    FaChainMap r;
    r.tf = 0;
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {
        const uint32_t t = (a.tf >> (2 * s)) & 3u;
        r.tf |= ((b.tf >> (2 * t)) & 3u) << (2 * s);
        r.c[s] = (b.has_abs ? 0u : a.c[s]) + b.c[t];
    }
    r.has_abs = a.has_abs | b.has_abs;
    r.abs = b.has_abs ? b.abs : a.abs;
    return r;
}
__device__ __forceinline__ FaChainMap fa_chain_shfl_up(const FaChainMap &m, int d) {
    FaChainMap r;
    r.tf = __shfl_up_sync(0xFFFFFFFFu, m.tf, d);
#pragma unroll
    for (int s = 0; s < 4; ++s) r.c[s] = __shfl_up_sync(0xFFFFFFFFu, m.c[s], d);
    r.has_abs = __shfl_up_sync(0xFFFFFFFFu, m.has_abs, d);
    r.abs = __shfl_up_sync(0xFFFFFFFFu, m.abs, d);
    return r;
}

__device__ __forceinline__ FaChainMap fa_chain_identity() {
    FaChainMap m;
    m.tf = 0xE4u; m.c[0] = m.c[1] = m.c[2] = m.c[3] = 0; m.has_abs = 0; m.abs = 0;
    return m;
}

__global__ void __launch_bounds__(FA_CHAIN_THREADS) fa_chain_kernel(const FaSum *__restrict__ sums, uint64_t blk0, uint64_t nblk, const uint64_t *__restrict__ fblk,
                                const uint64_t *__restrict__ gpos0, const uint32_t *__restrict__ fgenome, uint32_t nfiles,
                                uint64_t *__restrict__ carry, uint8_t *__restrict__ in_state, uint64_t *__restrict__ in_pos,
                                uint64_t *__restrict__ genome_end) {
    __shared__ FaChainMap s_wmap[FA_CHAIN_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t state0 = (uint32_t)carry[0];      // state / position after the previous chunk
    const uint64_t pos0 = carry[1];
    const uint64_t per = (nblk + FA_CHAIN_THREADS - 1) / FA_CHAIN_THREADS;
    const uint64_t b_lo = min(blk0 + (uint64_t)threadIdx.x * per, blk0 + nblk);
    const uint64_t b_hi = min(b_lo + per, blk0 + nblk);
    // first file starting at or after b_lo
    uint32_t nf0;
    {
        uint32_t lo = 0, hi = nfiles;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(fblk + mid) < b_lo) lo = mid + 1; else hi = mid; }
        nf0 = lo;
    }
    // ---- pass 1: the thread's map
    FaChainMap m = fa_chain_identity();
    {
        uint32_t st[4] = {0, 1, 2, 3};
        uint32_t nf = nf0;
        uint64_t next_fblk = nf < nfiles ? __ldg(fblk + nf) : ~0ull;
        for (uint64_t b = b_lo; b < b_hi; ++b) {
            while (b == next_fblk) {
                st[0] = st[1] = st[2] = st[3] = FS_SKIP;
                const uint64_t gp = __ldg(gpos0 + nf);
                if (gp != ~0ull) { m.has_abs = 1; m.abs = gp; m.c[0] = m.c[1] = m.c[2] = m.c[3] = 0; }
                ++nf;
                next_fblk = nf < nfiles ? __ldg(fblk + nf) : ~0ull;
            }
            const FaSum sm = __ldg(sums + b);
#pragma unroll
            for (int s = 0; s < 4; ++s) { m.c[s] += fa_cnt(sm, st[s]); st[s] = fa_tf(sm, st[s]); }
        }
        m.tf = st[0] | (st[1] << 2) | (st[2] << 4) | (st[3] << 6);
    }
    // ---- scan: inside the warp, then the warp totals
    FaChainMap inc = m;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const FaChainMap o = fa_chain_shfl_up(inc, d);
        if (lane >= (uint32_t)d) inc = fa_chain_compose(o, inc);
    }
    FaChainMap exc = fa_chain_shfl_up(inc, 1);
    if (lane == 0) exc = fa_chain_identity();
    if (lane == 31) s_wmap[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        FaChainMap w = s_wmap[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const FaChainMap o = fa_chain_shfl_up(w, d);
            if (lane >= (uint32_t)d) w = fa_chain_compose(o, w);
        }
        FaChainMap we = fa_chain_shfl_up(w, 1);
        if (lane == 0) we = fa_chain_identity();
        __syncwarp();
        s_wmap[lane] = we;                           // exclusive prefix of the warps
        if (lane == 31) {                            // chunk carry = everything composed, applied to the incoming carry
            carry[0] = (w.tf >> (2 * state0)) & 3u;
            carry[1] = (w.has_abs ? w.abs : pos0) + w.c[state0];
        }
    }
    __syncthreads();
    const FaChainMap pre = fa_chain_compose(s_wmap[warp], exc);
    uint32_t my_state = (pre.tf >> (2 * state0)) & 3u;
    uint64_t my_pos = (pre.has_abs ? pre.abs : pos0) + pre.c[state0];
    // ---- pass 2: walk the blocks with the real state
    if (b_hi > b_lo) {
        uint32_t nf = nf0;
        uint32_t g = nf0 ? __ldg(fgenome + nf0 - 1) : 0u;
        uint64_t next_fblk = nf < nfiles ? __ldg(fblk + nf) : ~0ull;
        for (uint64_t b = b_lo; b < b_hi; ++b) {
            while (b == next_fblk) {
                my_state = FS_SKIP;
                const uint64_t gp = __ldg(gpos0 + nf);
                if (gp != ~0ull) { atomicMax(reinterpret_cast<unsigned long long *>(genome_end + g), (unsigned long long)my_pos); my_pos = gp; g = __ldg(fgenome + nf); }
                ++nf;
                next_fblk = nf < nfiles ? __ldg(fblk + nf) : ~0ull;
            }
            in_state[b] = (uint8_t)my_state;
            in_pos[b] = my_pos;
            const FaSum sm = __ldg(sums + b);
            my_pos += fa_cnt(sm, my_state);
            my_state = fa_tf(sm, my_state);
        }
        // the genome current at the end of this thread's run has reached my_pos (positions only grow inside a genome: max)
        atomicMax(reinterpret_cast<unsigned long long *>(genome_end + g), (unsigned long long)my_pos);
    }
}

// Work items were laid out over each genome's WINDOW (an upper bound: the raw size of its files); cut them back to the
// sequence bytes the parse actually produced.
__global__ void fa_clip_items_kernel(SketchItem *__restrict__ items, uint32_t n, const uint64_t *__restrict__ genome_end) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t e = genome_end[items[i].genome];
    if (items[i].pos_end > e) items[i].pos_end = e > items[i].pos_begin ? e : items[i].pos_begin;
}

// Pass 3: emit 2-bit codes + validity + record-start planes.  The block's sequence bytes land on the contiguous positions
// [P, P + n): they are assembled in shared memory relative to P & ~63 and flushed with plain stores for words the block
// owns entirely and atomicOr for the (at most two) words per plane it shares with its neighbours; the planes start zeroed.
// A warp finds its incoming state and offset by walking the maps of the warps before it (pass 1 left them in wsum), its
// lanes' states with the ballots again, and the lanes' offsets with a shuffle scan of their sequence-byte counts.  Codes
// and validity come from the SIMD-in-word packer of sketch.cuh and are squeezed by the lane's keep mask (fa_compress: one
// step per run of dropped bytes — none for most lanes, one where a line ends).
// flags[f] |= 1 when file f shows FASTQ record syntax ('@' header, or a '+' / '@' line inside a record) or holds a '>' / '@' in the
// junk before its first header line (kseq would open a record there): the host re-sketches it through the record interface.
__global__ void __launch_bounds__(FA_THREADS) fa_emit_kernel(const uint8_t *__restrict__ text, uint64_t blk0, const uint64_t *__restrict__ fblk,
                                                            const uint64_t *__restrict__ flen, uint32_t nfiles, uint64_t chunk_end, uint32_t next_byte,
                                                            const uint8_t *__restrict__ in_state, const uint64_t *__restrict__ in_pos,
                                                            const FaSum *__restrict__ sums, const FaSum *__restrict__ wsum,
                                                            uint32_t *__restrict__ codes, uint32_t *__restrict__ valid32, uint32_t *__restrict__ start32,
                                                            uint32_t *__restrict__ flags) {
    constexpr int CW = (FA_BLOCK + 64) / 16 + 1, PW = (FA_BLOCK + 64) / 32 + 1;
    constexpr uint32_t NW = FA_THREADS / 32;
    __shared__ uint32_t s_codes[CW], s_valid[PW], s_start[PW];
    __shared__ uint64_t s_fstart, s_fend;
    __shared__ uint32_t s_file;
    const uint64_t blk = blk0 + blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_file = fa_file_of_block(fblk, nfiles, blk);
        s_fstart = __ldg(fblk + s_file) * (uint64_t)FA_BLOCK;
        s_fend = s_fstart + __ldg(flen + s_file);
    }
    for (int i = threadIdx.x; i < CW; i += FA_THREADS) s_codes[i] = 0;
    for (int i = threadIdx.x; i < PW; i += FA_THREADS) { s_valid[i] = 0; s_start[i] = 0; }
    __syncthreads();
    FaCls c;
    fa_classify(text, blk * FA_BLOCK + (uint64_t)threadIdx.x * FA_BPT, s_fstart, s_fend, chunk_end, next_byte, c);
    const FaLane L = fa_lane(c.nl, c.cr, c.gt, c.prev_nl, c.next_nl);
    const FaWarp w = fa_ballots(L);
    const uint32_t sin = in_state[blk];
    const uint64_t P = in_pos[blk], base0 = P & ~63ull;
    const uint32_t nseq = fa_cnt(__ldg(sums + blk), sin);
    // this warp's incoming state and offset: the maps of the warps before it, applied in order
    uint32_t wst = sin, woff = 0;
    for (uint32_t v = 0; v < warp; ++v) {
        const FaSum m = __ldg(wsum + (uint64_t)blockIdx.x * NW + v);
        woff += fa_cnt(m, wst);
        wst = fa_tf(m, wst);
    }
    const uint32_t st = fa_my_state(wst, L, w);
    const uint32_t K = fa_keep(L, st);
    const uint32_t n = FA_POPC(K);
    // what the parser does not cover (FASTQ syntax; '>' / '@' in the junk before a file's first header) can only show where a
    // line starts in the lane or while the file is still being skipped
    bool fq = false;
    if (L.ls || st == FS_SKIP) {
        const uint32_t vm = fa_below(s_fend > blk * FA_BLOCK + (uint64_t)threadIdx.x * FA_BPT
                                         ? (uint32_t)min((uint64_t)FA_BPT, s_fend - (blk * FA_BLOCK + (uint64_t)threadIdx.x * FA_BPT)) : 0u);
        fq = fa_fastq(L, st, fa_eq16(c.raw, '@') & vm, fa_eq16(c.raw, '+') & vm, c.gt);
    }
    if (__any_sync(0xFFFFFFFFu, fq) && lane == 0) atomicOr(&flags[s_file], 1u);
    // exclusive scan of the lanes' counts
    uint32_t inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    if (n) {
        const uint32_t li0 = (uint32_t)(P - base0) + woff + (inc - n);     // local index of this lane's first sequence byte
        uint32_t code = 0, vbits = 0;
        {
            const uint32_t ww[4] = {c.raw.x, c.raw.y, c.raw.z, c.raw.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t c8, v4;
                pack4(ww[i], c8, v4);
                code |= c8 << (8 * i);
                vbits |= v4 << (4 * i);
            }
        }
        uint32_t sbits = fa_starts(L, st, K);
        if (K != 0xFFFFu) {
            code = fa_compress(code, K, 2u);
            vbits = fa_compress(vbits & K, K, 1u);
            if (sbits) sbits = fa_compress(sbits, K, 1u);
        }
        if (n < 16u) { code &= fa_below(2u * n); vbits &= fa_below(n); }
        // codes: 2 bits per base, 16 bases per word; up to 32 bits of payload straddle at most 2 words
        const uint32_t cw = li0 >> 4, cs = 2 * (li0 & 15u);
        const uint64_t lo = (uint64_t)code << cs;
        atomicOr(&s_codes[cw], (uint32_t)lo);
        if ((uint32_t)(lo >> 32)) atomicOr(&s_codes[cw + 1], (uint32_t)(lo >> 32));
        const uint32_t pw = li0 >> 5, ps = li0 & 31u;
        const uint64_t v = (uint64_t)vbits << ps, sb = (uint64_t)sbits << ps;
        if ((uint32_t)v) atomicOr(&s_valid[pw], (uint32_t)v);
        if ((uint32_t)(v >> 32)) atomicOr(&s_valid[pw + 1], (uint32_t)(v >> 32));
        if ((uint32_t)sb) atomicOr(&s_start[pw], (uint32_t)sb);
        if ((uint32_t)(sb >> 32)) atomicOr(&s_start[pw + 1], (uint32_t)(sb >> 32));
    }
    __syncthreads();
    if (nseq == 0) return;
    const uint32_t lo_b = (uint32_t)(P - base0), hi_b = lo_b + nseq;            // local range of owned bases
    uint32_t *gc = codes + (base0 >> 4), *gv = valid32 + (base0 >> 5), *gs = start32 + (base0 >> 5);
    for (uint32_t x = threadIdx.x; x * 16 < hi_b; x += FA_THREADS) {
        const uint32_t y = s_codes[x];
        if (x * 16 >= lo_b && (x + 1) * 16 <= hi_b) gc[x] = y;
        else if (y) atomicOr(gc + x, y);
    }
    for (uint32_t x = threadIdx.x; x * 32 < hi_b; x += FA_THREADS) {
        const uint32_t y = s_valid[x], z = s_start[x];
        if (x * 32 >= lo_b && (x + 1) * 32 <= hi_b) { gv[x] = y; if (z) gs[x] = z; }
        else { if (y) atomicOr(gv + x, y); if (z) atomicOr(gs + x, z); }
    }
}

} // namespace db200
