// CPU emulation of the device FASTA parser's per-warp logic (dashing_b200/csrc/fasta_logic.h, the functions fasta.cuh's kernels
// call) over a whole file: lanes of 16 bytes, warps of 32 lanes, the warp's incoming state carried along.  Writes the flag and
// the records it would emit, so that tests/test_fasta_edges_cpu.py can compare them with the reference's kseq on random files.
//   usage: fasta_device_emul <in> <out>      out: "flag\n" then per record: u64 length + bytes
#include "../dashing_b200/csrc/fasta_logic.h"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    std::FILE *fp = std::fopen(argv[1], "rb");
    if (!fp) return 3;
    std::vector<uint8_t> t;
    for (int c; (c = std::fgetc(fp)) != EOF;) t.push_back((uint8_t)c);
    std::fclose(fp);
    const size_t n = t.size();
    auto at_ = [&](size_t i) -> uint8_t { return i < n ? t[i] : (uint8_t)'\n'; };   // bytes past the file's end read as '\n'
    std::vector<std::string> recs;
    bool flag = false;
    uint32_t S = FS_SKIP;
    for (size_t w0 = 0; w0 < n; w0 += 512) {
        FaLane L[32];
        uint32_t atm[32], plm[32], gtm[32], b_hs = 0, b_ls = 0, b_kab = 0, b_kb = 0;
        for (int l = 0; l < 32; ++l) {
            const size_t i0 = w0 + (size_t)l * 16;
            uint32_t nl = 0, cr = 0, gt = 0;
            atm[l] = plm[l] = 0;
            for (int i = 0; i < 16; ++i) {
                const bool valid = i0 + i < n;
                const uint8_t c = at_(i0 + i);
                nl |= (uint32_t)(c == '\n') << i;
                if (valid) { cr |= (uint32_t)(c == '\r') << i; gt |= (uint32_t)(c == '>') << i; atm[l] |= (uint32_t)(c == '@') << i; plm[l] |= (uint32_t)(c == '+') << i; }
            }
            gtm[l] = gt;
            const uint32_t prev_nl = i0 == 0 ? 1u : (uint32_t)(at_(i0 - 1) == '\n');
            L[l] = fa_lane(nl, cr, gt, prev_nl, at_(i0 + 16) == '\n');
            b_hs |= (uint32_t)(L[l].hs != 0) << l; b_ls |= (uint32_t)(L[l].ls != 0) << l;
            b_kab |= (uint32_t)(L[l].kA || L[l].kB) << l; b_kb |= (uint32_t)(L[l].kB != 0) << l;
        }
        for (int l = 0; l < 32; ++l) {
            const uint32_t ph = b_hs & fa_below(l);
            const uint32_t s = fa_lane_state(S, l, b_hs, b_ls, b_kab, b_kb, ph ? L[FA_MSB(ph)].det_out : 0);
            const uint32_t K = fa_keep(L[l], s), ST = fa_starts(L[l], s, K);
            if (L[l].ls || s == FS_SKIP) flag |= fa_fastq(L[l], s, atm[l], plm[l], gtm[l]);
            for (int i = 0; i < 16; ++i) {
                if (!((K >> i) & 1u)) continue;
                if ((ST >> i) & 1u || recs.empty()) recs.emplace_back();
                recs.back().push_back((char)t[w0 + (size_t)l * 16 + i]);
            }
        }
        S = fa_lane_state(S, 32, b_hs, b_ls, b_kab, b_kb, b_hs ? L[FA_MSB(b_hs)].det_out : 0);
    }
    std::FILE *out = std::fopen(argv[2], "wb");
    if (!out) return 4;
    std::fprintf(out, "%d\n", (int)flag);
    for (auto &r : recs) { const uint64_t len = r.size(); std::fwrite(&len, 8, 1, out); std::fwrite(r.data(), 1, r.size(), out); }
    std::fclose(out);
    return 0;
}
