// fasta_logic.h — the bit-parallel form of kseq's record rules used by fasta.cuh, as plain host/device functions over
// 16-bit byte-class masks (one bit per byte of a 16-byte lane) and 32-bit lane ballots (one bit per lane of a 512-byte warp).
// No CUDA intrinsics: tests/fasta_logic_test.cpp compiles this header with g++ and checks it against a byte-by-byte state
// machine on random text, for every incoming state.
//
// States: SKIP (before a file's first header) / HDR (inside a header line) / SEQN (after a header, no sequence byte yet) /
// SEQ.  Per byte (kseq_read, bonsai/klib/kseq.h:177-218): at a line start, '>' -> HDR, otherwise HDR -> SEQN; a byte is a
// sequence byte iff the state is SEQN/SEQ and it is neither '\n' nor a '\r' directly before '\n'; a sequence byte makes the
// state SEQ and, if the state was SEQN, is the first byte of its record.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#else
#define FA_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define FA_POPC(x) ((uint32_t)__popc(x))
#define FA_FFS(x) ((uint32_t)__ffs((int)(x)))                 /* 1-based, 0 for x == 0 */
#define FA_MSB(x) (31u - (uint32_t)__clz((int)(x)))          /* x != 0 */
#else
#define FA_POPC(x) ((uint32_t)__builtin_popcount(x))
#define FA_FFS(x) ((uint32_t)__builtin_ffs((int)(x)))
#define FA_MSB(x) (31u - (uint32_t)__builtin_clz(x))
#endif

enum : uint32_t { FS_SKIP = 0, FS_HDR = 1, FS_SEQN = 2, FS_SEQ = 3 };

// bits [0, n) set; n in [0, 32]
FA_HD uint32_t fa_below(uint32_t n) { return n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u); }

// What one lane (16 bytes) knows about itself.
struct FaLane {
    uint32_t keepable;   // not '\n', not a CR before LF
    uint32_t ls, hs;     // line starts; line starts whose byte is '>'
    uint32_t lead;       // bytes before the lane's first line start (they continue a line begun in an earlier lane)
    uint32_t mid;        // bytes of non-header lines that start in the lane BEFORE its first header start
    uint32_t tail;       // bytes of non-header lines that start in the lane AFTER its first header start
    uint32_t det_out;    // lanes with a header start: the state after the lane (independent of the incoming state)
    uint32_t kA, kB;     // keepable bytes in lead / in mid | tail
};

// nl / cr / gt: byte-class masks; prev_nl: the byte before the lane is '\n' (or the lane starts its file);
// next_nl: the byte after the lane is '\n'
FA_HD FaLane fa_lane(uint32_t nl, uint32_t cr, uint32_t gt, uint32_t prev_nl, uint32_t next_nl) {
    FaLane L;
    L.ls = ((nl << 1) | (prev_nl & 1u)) & 0xFFFFu;
    L.hs = L.ls & gt;
    const uint32_t drop = nl | (cr & ((nl >> 1) | ((next_nl & 1u) << 15)));
    L.keepable = ~drop & 0xFFFFu;
    L.lead = L.ls ? fa_below(FA_FFS(L.ls) - 1u) : 0xFFFFu;
    // header lines that start in the lane: from each header start up to the next line start
    uint32_t hl = 0;
    for (uint32_t h = L.hs; h; h &= h - 1u) {
        const uint32_t i = FA_FFS(h) - 1u;
        const uint32_t above = L.ls & ~fa_below(i + 1u);
        const uint32_t nxt = above ? FA_FFS(above) - 1u : 16u;
        hl |= fa_below(nxt) & ~fa_below(i);
    }
    const uint32_t seql = ~L.lead & ~hl & 0xFFFFu;
    if (L.hs) {
        const uint32_t fh = FA_FFS(L.hs) - 1u;
        L.mid = seql & fa_below(fh);
        L.tail = seql & ~fa_below(fh);
        const uint32_t last_ls = FA_MSB(L.ls);
        if ((L.hs >> last_ls) & 1u) L.det_out = FS_HDR;
        else {
            const uint32_t lh = FA_MSB(L.hs);
            L.det_out = (L.keepable & seql & ~fa_below(lh + 1u)) ? FS_SEQ : FS_SEQN;
        }
    } else {
        L.mid = seql;
        L.tail = 0;
        L.det_out = FS_SEQ;   // unused
    }
    L.kA = (L.lead & L.keepable) != 0;
    L.kB = ((L.mid | L.tail) & L.keepable) != 0;
    return L;
}

// State T carried across the lanes of R (a set of consecutive lanes, none of which holds a header start).
// b_ls / b_kab / b_kb: ballots of "has a line start" / "kA or kB" / "kB".
FA_HD uint32_t fa_evolve(uint32_t T, uint32_t R, uint32_t b_ls, uint32_t b_kab, uint32_t b_kb) {
    if (T == FS_HDR) {
        const uint32_t m = b_ls & R;
        if (!m) return FS_HDR;
        const uint32_t j = FA_FFS(m) - 1u;                       // the header line ends in lane j
        if ((b_kb >> j) & 1u) return FS_SEQ;
        return (b_kab & R & ~fa_below(j + 1u)) ? FS_SEQ : FS_SEQN;
    }
    if (T == FS_SEQN) return (b_kab & R) ? FS_SEQ : FS_SEQN;
    return T;                                                     // SKIP stays SKIP (no header start in R), SEQ stays SEQ
}

// Incoming state of lane l (0..32; 32 = the state after the whole warp) when the warp is entered in state S.
// det_h: det_out of the highest lane below l that holds a header start (only read if there is one).
FA_HD uint32_t fa_lane_state(uint32_t S, uint32_t l, uint32_t b_hs, uint32_t b_ls, uint32_t b_kab, uint32_t b_kb, uint32_t det_h) {
    const uint32_t below = fa_below(l);
    const uint32_t ph = b_hs & below;
    if (ph) {
        const uint32_t h = FA_MSB(ph);
        return fa_evolve(det_h, below & ~fa_below(h + 1u), b_ls, b_kab, b_kb);
    }
    return fa_evolve(S, below, b_ls, b_kab, b_kb);
}

// Sequence bytes of a lane entered in state s.
FA_HD uint32_t fa_keep(const FaLane &L, uint32_t s) {
    return L.keepable & ((s >= FS_SEQN ? L.lead : 0u) | (s != FS_SKIP ? L.mid : 0u) | L.tail);
}

// Record-start bytes among K = fa_keep(L, s): the first sequence byte after each header line.
FA_HD uint32_t fa_starts(const FaLane &L, uint32_t s, uint32_t K) {
    uint32_t st = 0;
    if (s == FS_SEQN) { const uint32_t m = K & (L.lead | L.mid); st |= m & (0u - m); }
    else if (s == FS_HDR) { const uint32_t m = K & L.mid; st |= m & (0u - m); }
    for (uint32_t h = L.hs; h; h &= h - 1u) {
        const uint32_t i = FA_FFS(h) - 1u;
        const uint32_t nh = (h & (h - 1u)) ? FA_FFS(h & (h - 1u)) - 1u : 16u;   // next header start
        const uint32_t m = K & L.tail & ~fa_below(i) & fa_below(nh);
        st |= m & (0u - m);
    }
    return st;
}

// What this parser does not cover, seen by a lane entered in state s (the file is flagged and the host sketches it through the
// record interface):
//   * FASTQ record syntax: '@' at a line start, or '+' at a line start inside a record (kseq.h:183, :196);
//   * a '>' in mid-line, or an '@' anywhere, BEFORE the file's first header line: without a pending header kseq scans
//     characters, not lines, for the next '>' / '@' (kseq.h:183), so such a byte would open a record there.
// at / pl / gt: byte-class masks of '@', '+' and '>'.
FA_HD bool fa_fastq(const FaLane &L, uint32_t s, uint32_t at, uint32_t pl, uint32_t gt) {
    if ((L.ls & at) != 0 || (L.ls & pl & ((s != FS_SKIP ? L.mid : 0u) | L.tail)) != 0) return true;
    if (s == FS_SKIP) {
        const uint32_t skipped = L.hs ? fa_below(FA_FFS(L.hs) - 1u) : 0xFFFFu;     // bytes before the lane's first header line
        return (((gt & ~L.ls) | at) & skipped) != 0;
    }
    return false;
}

// Removes the bits of v at the positions NOT in keep (16-bit), closing the gaps: a software pext, one step per run of
// dropped positions.  `width` = bits per position (1 for flag masks, 2 for base codes: v then holds 2 bits per position).
FA_HD uint32_t fa_compress(uint32_t v, uint32_t keep, uint32_t width) {
    uint32_t drop = ~keep & 0xFFFFu;
    while (drop) {
        const uint32_t hi = FA_MSB(drop);                                     // top of the highest run of dropped positions
        const uint32_t kept_below = keep & fa_below(hi);
        const uint32_t lo = kept_below ? FA_MSB(kept_below) + 1u : 0u;         // its bottom
        const uint32_t len = hi - lo + 1u;
        const uint32_t lowmask = width == 1u ? fa_below(lo) : fa_below(2u * lo);
        const uint32_t sh = width * len;
        v = (v & lowmask) | (sh >= 32u ? 0u : ((v >> sh) & ~lowmask));
        drop &= fa_below(lo);
    }
    return v;
}
