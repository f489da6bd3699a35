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
//   * a sequence line that starts with '+' or '@' switches kseq to FASTQ parsing (kseq.h:196-216): such files are not
//     handled here — the file is flagged and the host sketches it through the record interface instead.
// Windows never span records (encoder.h:444): the first sequence byte of every record is marked in the record-start plane.
//
// States: SKIP (before a file's first header) / HDR (inside a header line) / SEQN (after a header, no sequence byte yet)
// / SEQ.  A run of bytes maps an incoming state to (outgoing state, number of sequence bytes); these maps compose, so
// blocks are summarised independently (fa_summary_kernel), chained by a tiny sequential pass (fa_chain_kernel) and then
// emitted (fa_emit_kernel) with every byte knowing its state and its output position.
#pragma once
#include "common.cuh"
#include "sketch.cuh"

namespace db200 {

constexpr int FA_THREADS = 512;                 // threads per CTA
constexpr int FA_BPT = 16;                      // bytes per thread: one 16-byte load
constexpr int FA_BLOCK = FA_THREADS * FA_BPT;   // 8 KiB of text per CTA; files start on multiples of this
enum : uint32_t { FS_SKIP = 0, FS_HDR = 1, FS_SEQN = 2, FS_SEQ = 3 };

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

// One byte.  `ls`: the byte is a line start; `drop`: it is '\n', or a '\r' directly before a '\n'.
// Returns whether the byte is a sequence byte; `first` = it is the first one of its record.
__device__ __forceinline__ bool fa_step(uint32_t &state, uint32_t c, bool ls, bool drop, bool &first) {
    if (ls) {
        if (c == '>') state = FS_HDR;
        else if (state == FS_HDR) state = FS_SEQN;
    }
    const bool kept = state >= FS_SEQN && !drop;
    first = kept && state == FS_SEQN;
    if (kept) state = FS_SEQ;
    return kept;
}

struct FaBytes {
    uint32_t c[FA_BPT];
    uint32_t ls, drop;     // bit i: byte i is a line start / is dropped
};

// Loads the 16 bytes of this thread, the byte before (line start of byte 0) and the byte after (CR before LF).
// [file_start, file_end): the file this block belongs to — bytes of the block past the file's end (the gap up to the next
// file's block-aligned start) read as '\n', whatever the caller left there.  `chunk_end` / `next_byte`: the first byte not
// yet resident on the device and its value (the host reads it from its own copy of the text).
__device__ __forceinline__ void fa_load(const uint8_t *__restrict__ text, uint64_t i0, uint64_t file_start, uint64_t file_end, uint64_t chunk_end,
                                        uint32_t next_byte, FaBytes &fb) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + i0));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) fb.c[i] = (i0 + i < file_end) ? ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu) : (uint32_t)'\n';
    const uint32_t prev = (i0 > file_start && i0 - 1 < file_end) ? (uint32_t)__ldg(text + i0 - 1) : (uint32_t)'\n';
    const uint64_t in = i0 + FA_BPT;
    const uint32_t next = in >= file_end ? (uint32_t)'\n' : (in < chunk_end ? (uint32_t)__ldg(text + in) : next_byte);
    uint32_t ls = 0, drop = 0;
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        const uint32_t p = i ? fb.c[i - 1] : prev, n = i + 1 < FA_BPT ? fb.c[i + 1] : next;
        ls |= (uint32_t)(p == '\n') << i;
        drop |= (uint32_t)(fb.c[i] == '\n' || (fb.c[i] == '\r' && n == '\n')) << i;
    }
    fb.ls = ls; fb.drop = drop;
}

__device__ __forceinline__ FaSum fa_thread_summary(const FaBytes &fb) {
    FaSum r = 0;
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {
        uint32_t st = s, cnt = 0;
#pragma unroll
        for (int i = 0; i < FA_BPT; ++i) {
            bool first;
            cnt += fa_step(st, fb.c[i], (fb.ls >> i) & 1u, (fb.drop >> i) & 1u, first);
        }
        r |= (FaSum)st << (2 * s);
        r |= (FaSum)cnt << (8 + 14 * s);
    }
    return r;
}

// Scan of the per-thread summaries over the CTA (thread order = byte order).  Returns this thread's EXCLUSIVE prefix (the map
// of all bytes before its own); s_warp[FA_THREADS / 32] afterwards holds the whole block's map.
__device__ __forceinline__ FaSum fa_block_scan(FaSum mine, FaSum *s_warp /* [FA_THREADS / 32 + 1] */) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr FaSum IDENT = 0xE4ull;   // tf: s -> s, counts 0
    constexpr uint32_t NW = FA_THREADS / 32;
    FaSum inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const FaSum o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc = fa_compose(o, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {                   // exclusive scan of the NW warp totals; the grand total goes to s_warp[NW]
        FaSum w = lane < NW ? s_warp[lane] : IDENT;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const FaSum o = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= (uint32_t)d) w = fa_compose(o, w);
        }
        const FaSum e = __shfl_up_sync(0xFFFFFFFFu, w, 1);
        __syncwarp();
        if (lane < NW) s_warp[lane] = lane ? e : IDENT;
        if (lane == NW - 1) s_warp[NW] = w;
    }
    __syncthreads();
    FaSum exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) exc = IDENT;
    return fa_compose(s_warp[warp], exc);
}

// A block's file: files start on block boundaries, `fblk` holds their first block (ascending).
__device__ __forceinline__ uint32_t fa_file_of_block(const uint64_t *__restrict__ fblk, uint32_t nfiles, uint64_t blk) {
    uint32_t lo = 0, hi = nfiles;          // last f with fblk[f] <= blk
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(fblk + mid) <= blk) lo = mid; else hi = mid; }
    return lo;
}

// Pass 1: one summary per block.
__global__ void __launch_bounds__(FA_THREADS) fa_summary_kernel(const uint8_t *__restrict__ text, uint64_t blk0, const uint64_t *__restrict__ fblk,
                                                               const uint64_t *__restrict__ flen, uint32_t nfiles, uint64_t chunk_end, uint32_t next_byte, FaSum *__restrict__ sums) {
    __shared__ FaSum s_warp[FA_THREADS / 32 + 1];
    __shared__ uint64_t s_fstart, s_fend;
    const uint64_t blk = blk0 + blockIdx.x;
    if (threadIdx.x == 0) {
        const uint32_t f = fa_file_of_block(fblk, nfiles, blk);
        s_fstart = __ldg(fblk + f) * (uint64_t)FA_BLOCK;
        s_fend = s_fstart + __ldg(flen + f);
    }
    __syncthreads();
    FaBytes fb;
    fa_load(text, blk * FA_BLOCK + (uint64_t)threadIdx.x * FA_BPT, s_fstart, s_fend, chunk_end, next_byte, fb);
    fa_block_scan(fa_thread_summary(fb), s_warp);
    if (threadIdx.x == 0) sums[blk] = s_warp[FA_THREADS / 32];
}

// Pass 2: chain the blocks of one chunk.  carry[0] = state, carry[1] = next output position after the previous chunk.  A file
// start resets the state to SKIP; the first file of a genome also moves the output position to that genome's window
// (gpos0[f] != ~0).  genome_end[g] = one past the last sequence byte written for genome g so far (final once a later genome has
// started or the text has ended).
// One warp.  Lane l owns 64 consecutive blocks of a 2048-block tile: it first folds them into one map of the incoming
// (state, position) — speculating over the four possible incoming states —, the 32 maps are scanned with shuffles, and the lane
// then walks its blocks again with the now known incoming state and position.  File starts reset the state; a genome's first
// file makes the position absolute (its window), which the map records as (has_abs, abs).
constexpr int FA_CHAIN_PER_LANE = 64, FA_CHAIN_TILE = 32 * FA_CHAIN_PER_LANE;
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

__global__ void __launch_bounds__(32) fa_chain_kernel(const FaSum *__restrict__ sums, uint64_t blk0, uint64_t nblk, const uint64_t *__restrict__ fblk,
                                const uint64_t *__restrict__ gpos0, const uint32_t *__restrict__ fgenome, uint32_t nfiles,
                                uint64_t *__restrict__ carry, uint8_t *__restrict__ in_state, uint64_t *__restrict__ in_pos,
                                uint64_t *__restrict__ genome_end) {
    const uint32_t lane = threadIdx.x;
    uint32_t state = (uint32_t)carry[0];      // carried from tile to tile (identical in every lane)
    uint64_t pos = carry[1];
    for (uint64_t t0 = 0; t0 < nblk; t0 += FA_CHAIN_TILE) {
        const uint64_t b_lo = min(blk0 + t0 + (uint64_t)lane * FA_CHAIN_PER_LANE, blk0 + nblk);
        const uint64_t b_hi = min(b_lo + FA_CHAIN_PER_LANE, blk0 + nblk);
        // first file starting at or after b_lo
        uint32_t nf0;
        {
            uint32_t lo = 0, hi = nfiles;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(fblk + mid) < b_lo) lo = mid + 1; else hi = mid; }
            nf0 = lo;
        }
        // ---- pass 1: the lane's map
        FaChainMap m;
        m.tf = 0xE4u; m.c[0] = m.c[1] = m.c[2] = m.c[3] = 0; m.has_abs = 0; m.abs = 0;
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
        // ---- inclusive scan over the lanes, then this lane's incoming (state, position)
        FaChainMap inc = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const FaChainMap o = fa_chain_shfl_up(inc, d);
            if (lane >= (uint32_t)d) inc = fa_chain_compose(o, inc);
        }
        FaChainMap exc = fa_chain_shfl_up(inc, 1);
        if (lane == 0) { exc.tf = 0xE4u; exc.c[0] = exc.c[1] = exc.c[2] = exc.c[3] = 0; exc.has_abs = 0; exc.abs = 0; }
        uint32_t my_state = (exc.tf >> (2 * state)) & 3u;
        uint64_t my_pos = (exc.has_abs ? exc.abs : pos) + exc.c[state];
        // ---- pass 2: walk the blocks with the real state
        {
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
            // the genome current at the end of this lane's run has reached my_pos (positions only grow inside a genome: max)
            if (b_hi > b_lo) atomicMax(reinterpret_cast<unsigned long long *>(genome_end + g), (unsigned long long)my_pos);
        }
        // tile carry = state / position after the last lane
        state = __shfl_sync(0xFFFFFFFFu, my_state, 31);
        pos = __shfl_sync(0xFFFFFFFFu, my_pos, 31);
    }
    if (lane == 0) { carry[0] = state; carry[1] = pos; }
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
// flags[f] |= 1 when file f shows FASTQ record syntax ('@' header, or a '+' / '@' line inside a record).
__global__ void __launch_bounds__(FA_THREADS) fa_emit_kernel(const uint8_t *__restrict__ text, uint64_t blk0, const uint64_t *__restrict__ fblk,
                                                            const uint64_t *__restrict__ flen, uint32_t nfiles, uint64_t chunk_end, uint32_t next_byte,
                                                            const uint8_t *__restrict__ in_state, const uint64_t *__restrict__ in_pos,
                                                            uint32_t *__restrict__ codes, uint32_t *__restrict__ valid32, uint32_t *__restrict__ start32,
                                                            uint32_t *__restrict__ flags) {
    constexpr int CW = (FA_BLOCK + 64) / 16 + 1, PW = (FA_BLOCK + 64) / 32 + 1;
    __shared__ FaSum s_warp[FA_THREADS / 32 + 1];
    __shared__ uint32_t s_codes[CW], s_valid[PW], s_start[PW];
    __shared__ uint64_t s_fstart, s_fend;
    __shared__ uint32_t s_file;
    const uint64_t blk = blk0 + blockIdx.x;
    if (threadIdx.x == 0) {
        s_file = fa_file_of_block(fblk, nfiles, blk);
        s_fstart = __ldg(fblk + s_file) * (uint64_t)FA_BLOCK;
        s_fend = s_fstart + __ldg(flen + s_file);
    }
    for (int i = threadIdx.x; i < CW; i += FA_THREADS) s_codes[i] = 0;
    for (int i = threadIdx.x; i < PW; i += FA_THREADS) { s_valid[i] = 0; s_start[i] = 0; }
    __syncthreads();
    FaBytes fb;
    fa_load(text, blk * FA_BLOCK + (uint64_t)threadIdx.x * FA_BPT, s_fstart, s_fend, chunk_end, next_byte, fb);
    const FaSum pre = fa_block_scan(fa_thread_summary(fb), s_warp);
    const FaSum total = s_warp[FA_THREADS / 32];
    const uint32_t sin = in_state[blk];
    const uint64_t P = in_pos[blk], base0 = P & ~63ull;
    const uint32_t nseq = fa_cnt(total, sin);
    uint32_t state = fa_tf(pre, sin);
    const uint32_t li0 = (uint32_t)(P - base0) + fa_cnt(pre, sin);     // local index of this thread's first sequence byte
    uint64_t code = 0;
    uint32_t vbits = 0, sbits = 0, n = 0;
    bool fastq = false;
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        const uint32_t c = fb.c[i];
        const bool ls = (fb.ls >> i) & 1u;
        // FASTQ syntax: '@' opening a record, or a '+' / '@' line inside one (kseq.h:183, :196)
        fastq |= ls && ((c == '@') || (c == '+' && state >= FS_SEQN));
        bool first;
        if (fa_step(state, c, ls, (fb.drop >> i) & 1u, first)) {
            const uint32_t up = c & 0xDFu;
            const uint32_t ok = (up == 'A') | (up == 'C') | (up == 'G') | (up == 'T');
            code |= (uint64_t)(((c >> 1) ^ (c >> 2)) & 3u) << (2 * n);     // A0 C1 G2 T3 (anything for invalid bytes)
            vbits |= ok << n;
            sbits |= (uint32_t)first << n;
            ++n;
        }
    }
    if (fastq) atomicOr(&flags[s_file], 1u);
    if (n) {
        // codes: 2 bits per base, 16 bases per word; up to 32 bits of payload straddle at most 2 words
        const uint32_t cw = li0 >> 4, cs = 2 * (li0 & 15u);
        const uint64_t lo = code << cs;
        atomicOr(&s_codes[cw], (uint32_t)lo);
        if ((uint32_t)(lo >> 32)) atomicOr(&s_codes[cw + 1], (uint32_t)(lo >> 32));
        const uint32_t pw = li0 >> 5, ps = li0 & 31u;
        const uint64_t v = (uint64_t)vbits << ps, s = (uint64_t)sbits << ps;
        if ((uint32_t)v) atomicOr(&s_valid[pw], (uint32_t)v);
        if ((uint32_t)(v >> 32)) atomicOr(&s_valid[pw + 1], (uint32_t)(v >> 32));
        if ((uint32_t)s) atomicOr(&s_start[pw], (uint32_t)s);
        if ((uint32_t)(s >> 32)) atomicOr(&s_start[pw + 1], (uint32_t)(s >> 32));
    }
    __syncthreads();
    if (nseq == 0) return;
    const uint32_t lo_b = (uint32_t)(P - base0), hi_b = lo_b + nseq;            // local range of owned bases
    uint32_t *gc = codes + (base0 >> 4), *gv = valid32 + (base0 >> 5), *gs = start32 + (base0 >> 5);
    for (uint32_t w = threadIdx.x; w * 16 < hi_b; w += FA_THREADS) {
        const uint32_t x = s_codes[w];
        if (w * 16 >= lo_b && (w + 1) * 16 <= hi_b) gc[w] = x;
        else if (x) atomicOr(gc + w, x);
    }
    for (uint32_t w = threadIdx.x; w * 32 < hi_b; w += FA_THREADS) {
        const uint32_t x = s_valid[w], y = s_start[w];
        if (w * 32 >= lo_b && (w + 1) * 32 <= hi_b) { gv[w] = x; if (y) gs[w] = y; }
        else { if (x) atomicOr(gv + w, x); if (y) atomicOr(gs + w, y); }
    }
}

} // namespace db200
