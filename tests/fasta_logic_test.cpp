// CPU check of dashing_b200/csrc/fasta_logic.h (the bit-parallel FASTA record rules the device parser uses) against a
// byte-by-byte state machine, on random text, for every incoming state.  Built and run by tests/test_fasta_logic.py.
#include "../dashing_b200/csrc/fasta_logic.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

static bool step(uint32_t &state, uint32_t c, bool ls, bool drop, bool &first) {   // the rules, one byte at a time
    if (ls) { if (c == '>') state = FS_HDR; else if (state == FS_HDR) state = FS_SEQN; }
    const bool kept = state >= FS_SEQN && !drop;
    first = kept && state == FS_SEQN;
    if (kept) state = FS_SEQ;
    return kept;
}

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? std::atoi(argv[1]) : 20000;
    std::mt19937_64 rng(12345);
    const char alphabet[] = "ACGTacgtNn\n\n\n\r\r>>@+ \t-";
    long checked = 0;
    for (int r = 0; r < rounds; ++r) {
        // 512 bytes (one warp) + the byte before and after; several "textures": dense newlines, long lines, headers
        const int mode = r % 5;
        std::vector<uint8_t> t(514);
        for (auto &c : t) {
            const uint64_t x = rng();
            if (mode == 0) c = alphabet[x % (sizeof alphabet - 1)];
            else if (mode == 1) c = (x % 80 == 0) ? '\n' : "ACGT"[x & 3];
            else if (mode == 2) c = (x % 40 == 0) ? '\n' : (x % 97 == 0 ? '>' : (x % 89 == 0 ? '\r' : "ACGTN"[x % 5]));
            else if (mode == 3) c = (x % 6 == 0) ? '\n' : (x % 7 == 0 ? '>' : "AC\r+@"[x % 5]);
            else c = (x % 300 == 0) ? '\n' : "ACGT>"[x % 5];
        }
        const uint8_t *w = t.data() + 1;
        // lane infos and ballots
        FaLane L[32];
        uint32_t at[32], pl[32], gtm[32], b_hs = 0, b_ls = 0, b_kab = 0, b_kb = 0;
        for (int l = 0; l < 32; ++l) {
            uint32_t nl = 0, cr = 0, gt = 0; at[l] = pl[l] = 0;
            for (int i = 0; i < 16; ++i) {
                const uint8_t c = w[l * 16 + i];
                nl |= (uint32_t)(c == '\n') << i; cr |= (uint32_t)(c == '\r') << i; gt |= (uint32_t)(c == '>') << i;
                at[l] |= (uint32_t)(c == '@') << i; pl[l] |= (uint32_t)(c == '+') << i;
            }
            gtm[l] = gt;
            L[l] = fa_lane(nl, cr, gt, w[l * 16 - 1] == '\n', w[l * 16 + 16] == '\n');
            b_hs |= (uint32_t)(L[l].hs != 0) << l; b_ls |= (uint32_t)(L[l].ls != 0) << l;
            b_kab |= (uint32_t)(L[l].kA || L[l].kB) << l; b_kb |= (uint32_t)(L[l].kB != 0) << l;
        }
        for (uint32_t S = 0; S < 4; ++S) {
            uint32_t state = S;
            bool fq = false, fq_fast = false;
            for (int l = 0; l <= 32; ++l) {
                // the highest header-start lane below l
                const uint32_t ph = b_hs & fa_below(l);
                const uint32_t det_h = ph ? L[FA_MSB(ph)].det_out : 0;
                const uint32_t got = fa_lane_state(S, l, b_hs, b_ls, b_kab, b_kb, det_h);
                if (got != state) { std::printf("round %d S=%u lane %d: state %u, want %u\n", r, S, l, got, state); return 1; }
                if (l == 32) break;
                uint32_t K = 0, ST = 0;
                for (int i = 0; i < 16; ++i) {
                    const int p = l * 16 + i;
                    const uint32_t c = w[p];
                    const bool ls = w[p - 1] == '\n', drop = c == '\n' || (c == '\r' && w[p + 1] == '\n');
                    fq |= ls && (c == '@' || (c == '+' && state != FS_SKIP));
                    fq |= state == FS_SKIP && ((c == '>' && !ls) || c == '@');      // kseq's character scan for the first header
                    bool first;
                    if (step(state, c, ls, drop, first)) { K |= 1u << i; ST |= (uint32_t)first << i; }
                }
                const uint32_t gk = fa_keep(L[l], got), gs = fa_starts(L[l], got, gk);
                if (gk != K || gs != ST) { std::printf("round %d S=%u lane %d: keep %04x/%04x starts %04x/%04x\n", r, S, l, gk, K, gs, ST); return 1; }
                fq_fast |= fa_fastq(L[l], got, at[l], pl[l], gtm[l]);
                // compress: flags and 2-bit codes
                const uint32_t v = (uint32_t)rng() & 0xFFFFu, c2 = (uint32_t)rng();
                uint32_t wv = 0, wc = 0; int n = 0;
                for (int i = 0; i < 16; ++i) if ((K >> i) & 1u) { wv |= ((v >> i) & 1u) << n; wc |= ((c2 >> (2 * i)) & 3u) << (2 * n); ++n; }
                const uint32_t gv = fa_compress(v, K, 1) & fa_below(n), gc = fa_compress(c2, K, 2) & fa_below(2 * n);
                if (gv != wv || gc != wc) { std::printf("round %d lane %d: compress %x/%x %x/%x keep %04x\n", r, l, gv, wv, gc, wc, K); return 1; }
                ++checked;
            }
            if (fq != fq_fast) { std::printf("round %d S=%u: fastq %d, want %d\n", r, S, (int)fq_fast, (int)fq); return 1; }
        }
    }
    std::printf("ok %ld lane checks\n", checked);
    return 0;
}
