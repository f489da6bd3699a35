// dashing_host.cpp — see dashing_host.hpp.  Host-side mirror of the reference's sketch / dist drivers and formats.
#include "dashing_host.hpp"
#include "../../include/dashing_b200.h"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <cinttypes>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <memory>
#include <getopt.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace db200h {

static void check(int rc) {
    if (rc != DB200_OK) throw Error(db200_last_error());
}

// DB200_TIMING=1: phase timings of the drivers on stderr (wall clock since the first call)
static void phase(const char *what) {
    static const bool on = std::getenv("DB200_TIMING") != nullptr;
    if (!on) return;
    static const double t0 = omp_get_wtime();
    static double last = t0;
    const double now = omp_get_wtime();
    std::fprintf(stderr, "[db200 timing] %-28s +%8.3f ms  (t = %8.3f ms)\n", what, (now - last) * 1e3, (now - t0) * 1e3);
    last = now;
}

static bool isfile(const std::string &p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

// ---------------------------------------------------------------------------------------------------------------
// .hll container
// ---------------------------------------------------------------------------------------------------------------
std::vector<uint8_t> hll_payload(const uint8_t *regs, uint32_t p, int estim, int jestim, double value) {
    const size_t m = size_t(1) << p;
    std::vector<uint8_t> out(28 + m);
    const uint32_t hdr[5] = {value >= 0. ? 1u : 0u, (uint32_t)estim, (uint32_t)jestim, 1u, p};   // hll.h:1041-1044
    std::memcpy(out.data(), hdr, 20);
    std::memcpy(out.data() + 20, &value, 8);
    std::memcpy(out.data() + 28, regs, m);
    return out;
}

void write_hll(const std::string &path, const uint8_t *regs, uint32_t p, int estim, int jestim, double value) {
    gzFile fp = gzopen(path.c_str(), "wb");
    if (!fp) throw Error("Could not open file at '" + path + "' for writing");
    const auto buf = hll_payload(regs, p, estim, jestim, value);
    const bool ok = gzwrite(fp, buf.data(), (unsigned)buf.size()) == (int)buf.size();
    gzclose(fp);
    if (!ok) throw Error("Error writing to file.");
}

HllFile read_hll(const std::string &path) {
    gzFile fp = gzopen(path.c_str(), "rb");
    if (!fp) throw Error("Could not open file at '" + path + "' for reading");
    HllFile h;
    uint32_t bf[5];
    auto rd = [&](void *dst, size_t len) {
        if ((size_t)gzread(fp, dst, (unsigned)len) != len) { gzclose(fp); throw Error("Error reading from file " + path); }
    };
    rd(bf, 20);
    h.is_calculated = bf[0]; h.estim = bf[1]; h.jestim = bf[2]; h.marker = bf[3]; h.p = bf[4];
    rd(&h.value, 8);
    if (h.p > 40) { gzclose(fp); throw Error("implausible sketch size in " + path); }
    h.core.resize(size_t(1) << h.p);
    rd(h.core.data(), h.core.size());
    gzclose(fp);
    return h;
}

void write_hll_stream(gzFile fp, const uint8_t *regs, uint32_t p, int estim, int jestim, double value) {
    const auto buf = hll_payload(regs, p, estim, jestim, value);
    if (gzwrite(fp, buf.data(), (unsigned)buf.size()) != (int)buf.size()) throw Error("Error writing to file.");
}

// hll_t(gzFile) repeatedly (dist_by_seq, src/sketch_and_cmp.h:85-88): `count` sketches from one gzip stream
std::vector<HllFile> read_hll_container(const std::string &path, size_t count) {
    gzFile fp = gzopen(path.c_str(), "rb");
    if (!fp) throw Error("Failed to open file at " + path);
    std::vector<HllFile> out;
    auto rd = [&](void *dst, size_t len) {
        if ((size_t)gzread(fp, dst, (unsigned)len) != len) { gzclose(fp); throw Error("Error reading from file " + path); }
    };
    while (out.size() < count) {
        HllFile h;
        uint32_t bf[5];
        rd(bf, 20);
        h.is_calculated = bf[0]; h.estim = bf[1]; h.jestim = bf[2]; h.marker = bf[3]; h.p = bf[4];
        rd(&h.value, 8);
        if (h.p > 40) { gzclose(fp); throw Error("implausible sketch size in " + path); }
        h.core.resize(size_t(1) << h.p);
        rd(h.core.data(), h.core.size());
        out.push_back(std::move(h));
    }
    gzclose(fp);
    return out;
}

std::string make_fname(const char *path, size_t sketch_p, int /*wsz*/, int k, int /*csz*/, const std::string &spacing,
                       const std::string &suffix, const std::string &prefix) {
    std::string ret(prefix);
    if (!ret.empty()) ret += '/';
    const char *p = std::strchr(path, ' ');
    p = p ? p + 1 : path;                       // multi-file paths are named after what follows the first separator
    const char *p2;
    if (!ret.empty() && (p2 = std::strrchr(p, '/'))) ret += std::string(p2 + 1);
    else ret += p;
    ret += ".w";                                // the window size never makes it into the name (src/dashing.h:510 is a no-op)
    ret += ".";
    ret += std::to_string(k);
    ret += ".spacing";
    ret += spacing;
    ret += '.';
    if (!suffix.empty()) { ret += "suf"; ret += suffix; ret += '.'; }
    ret += std::to_string(sketch_p);
    ret += ".hll";
    return ret;
}

std::vector<std::string> split_paths(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        const size_t b = s.find(sep, a);
        std::string tok = s.substr(a, b == std::string::npos ? std::string::npos : b - a);
        const bool blank = std::all_of(tok.begin(), tok.end(), [](unsigned char c) { return std::isspace(c); });
        if (!out.empty() && blank) break;      // the reference stops at the first blank token (src/substrs.h:23)
        out.push_back(std::move(tok));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}

std::vector<std::string> get_paths(const std::string &file) {
    // get_lines (bonsai/include/bonsai/util.h:1176-1184): empty lines and lines starting with '#' are skipped; a file that
    // cannot be opened yields no paths (the callers then stop with "No paths")
    std::ifstream is(file);
    std::vector<std::string> out;
    for (std::string line; std::getline(is, line);)
        if (!line.empty() && line.front() != '#') out.push_back(line);
    return out;
}

void sort_paths_by_fsize(std::vector<std::string> &paths) {
    if (paths.size() < 2) return;
    std::vector<std::pair<uint32_t, std::string>> ps;
    for (auto &p : paths) {
        size_t tot = 0;
        for (auto &f : split_paths(p)) { struct stat st; if (::stat(f.c_str(), &st) == 0) tot += st.st_size; }
        ps.emplace_back((uint32_t)tot, p);
    }
    // the reference uses an unstable std::sort on size only; equal sizes are order-unspecified there — use --avoid-sorting for parity
    std::stable_sort(ps.begin(), ps.end(), [](const auto &x, const auto &y) { return x.first > y.first; });
    for (size_t i = 0; i < paths.size(); ++i) paths[i] = std::move(ps[i].second);
}

// ---------------------------------------------------------------------------------------------------------------
// FASTA / FASTQ
// ---------------------------------------------------------------------------------------------------------------
namespace {
// Destination of sequence bytes: a growing std::string, or a caller-owned window that must not be overrun.
struct SeqSink {
    std::string *str = nullptr;
    char *dst = nullptr;
    size_t len = 0, cap = 0;
    bool overflow = false;
    void append(const char *s, size_t l) {
        if (str) { str->append(s, l); len = str->size(); return; }
        if (len + l > cap) { overflow = true; return; }
        std::memcpy(dst + len, s, l);
        len += l;
    }
    char back() const { return str ? str->back() : dst[len - 1]; }
    void pop_back() { if (str) str->pop_back(); --len; }
};

struct LineReader {
    gzFile fp;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    explicit LineReader(gzFile f) : fp(f), buf(1 << 18) {}
    int peek() {
        if (pos == end) fill();
        return pos < end ? (unsigned char)buf[pos] : -1;
    }
    void fill() {
        if (eof) return;
        const int n = gzread(fp, buf.data(), (unsigned)buf.size());
        pos = 0; end = n > 0 ? (size_t)n : 0;
        if (n <= 0) eof = true;
    }
    // appends the rest of the current line (without the newline / trailing CR) to `out` (or discards it); `floor` is the
    // sink length at the start of the record: kseq strips a trailing CR only while the record holds more than it
    // (ks_getuntil2, bonsai/klib/kseq.h:140-141: `str->l > 1`)
    size_t line(SeqSink *out, size_t floor = 0) {
        size_t got = 0;
        for (;;) {
            if (pos == end) { fill(); if (pos == end) return got; }
            const char *s = buf.data() + pos;
            const char *nl = (const char *)std::memchr(s, '\n', end - pos);
            const size_t len = nl ? (size_t)(nl - s) : end - pos;
            if (out) out->append(s, len);
            got += len;
            pos += len + (nl ? 1 : 0);
            if (nl) {
                if (out && !out->overflow && out->len > floor + 1 && out->back() == '\r') { out->pop_back(); --got; }
                return got;
            }
        }
    }
};

// kseq_read's record loop (bonsai/klib/kseq.h:177-218) over one file; record boundaries (sink lengths) go to `ends`, and,
// when asked for, the record names (header up to the first whitespace, kseq.h:190) to `names`.
//   * Without a pending header (start of the file, and after every FASTQ record: last_char == 0) kseq scans CHARACTERS, not
//     lines, for the next '>' or '@' (:183) — a '>' in the middle of a junk line opens a record.
//   * Sequence lines are read until a line STARTS with '>', '@' (the next header: consumed and remembered, :199) or '+' (:193).
//   * A header character with nothing behind it (end of file) yields no record (:188).
void parse_records(const std::string &file, SeqSink &sink, std::vector<uint64_t> &ends, std::vector<std::string> *names = nullptr) {
    gzFile fp = gzopen(file.c_str(), "rb");
    if (!fp) throw Error("Could not open file at " + file + ". Abort!");
    gzbuffer(fp, 1 << 18);
    LineReader lr(fp);
    bool pending = false;                  // kseq's last_char != 0: the header character has been consumed already
    int c;
    while (!sink.overflow) {
        if (!pending) {
            while ((c = lr.peek()) != -1 && c != '>' && c != '@') ++lr.pos;
            if (c == -1) break;
            ++lr.pos;                      // the header character
        }
        if (lr.peek() == -1) break;        // ks_getuntil(name) < 0: "normal exit: EOF"
        if (names) {
            std::string hdr;
            SeqSink hs;
            hs.str = &hdr;
            lr.line(&hs, 0);
            size_t e = 0;
            while (e < hdr.size() && !std::isspace((unsigned char)hdr[e])) ++e;
            names->push_back(hdr.substr(0, e));
        } else lr.line(nullptr);           // name + comment
        const size_t rs = sink.len;
        while ((c = lr.peek()) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') { ++lr.pos; continue; }    // "skip empty lines" (:195): no ks_getuntil2, so no CR is stripped here
            lr.line(&sink, rs);
        }
        pending = c == '>' || c == '@';
        if (c != -1) ++lr.pos;             // the character that ended the record is consumed either way
        if (c == '+') {                    // FASTQ: skip the rest of the '+' line, then quality lines until they cover the sequence
            lr.line(nullptr);
            std::string qual;               // while (getuntil2(qual) >= 0 && qual.l < seq.l), kseq.h:210
            SeqSink qs;
            qs.str = &qual;
            do {
                if (lr.peek() == -1) break;
                lr.line(&qs, 0);
            } while (qs.len < sink.len - rs);
            pending = false;
            if (qs.len != sink.len - rs) {  // "truncated quality string": kseq_read returns -2 (:214) and the callers' loops
                sink.len = rs;              // (`while(kseq_read(ks) >= 0)`, encoder.h) stop — the record is never seen
                if (sink.str) sink.str->resize(rs);
                if (names) names->pop_back();
                break;
            }
        }
        ends.push_back(sink.len);
    }
    gzclose(fp);
}
} // namespace

void for_each_named_record(const std::string &file, const std::function<void(const std::string &, const char *, size_t)> &fn) {
    std::string seq;
    SeqSink sink;
    sink.str = &seq;
    std::vector<uint64_t> ends;
    std::vector<std::string> names;
    parse_records(file, sink, ends, &names);
    uint64_t b = 0;
    for (size_t i = 0; i < ends.size(); ++i) { fn(names[i], seq.data() + b, ends[i] - b); b = ends[i]; }
}

size_t parse_into_window(const std::string &file, char *dst, size_t cap, std::vector<uint64_t> &ends) {
    SeqSink sk;
    sk.dst = dst; sk.cap = cap;
    parse_records(file, sk, ends);
    return sk.overflow ? SIZE_MAX : sk.len;
}

void for_each_record(const std::string &file, const std::function<void(const char *, size_t)> &fn) {
    std::string seq;
    SeqSink sink;
    sink.str = &seq;
    std::vector<uint64_t> ends;
    parse_records(file, sink, ends);
    uint64_t b = 0;
    for (uint64_t e : ends) { fn(seq.data() + b, e - b); b = e; }
}

// ---------------------------------------------------------------------------------------------------------------
// emitters
// ---------------------------------------------------------------------------------------------------------------
// "%.6g" / "%0.6g" / "%g" of the reference's emitters: std::to_chars(general, precision 6) is specified to produce printf's
// digits (checked on 3e7 floats incl. inf / nan / denormals) at 2.5x the speed of snprintf — the text formats are bound by it.
static void append_g6(std::string &s, double v) {
    char b[64];
    const auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::general, 6);
    s.append(b, r.ptr - b);
}

std::string format_sizes(const std::vector<std::string> &paths, const double *card) {
    std::string s("#Path\tSize (est.)\n");                       // src/sketch_and_cmp.h:372
    char b[32];
    for (size_t i = 0; i < paths.size(); ++i) {
        s += paths[i];
        const int n = std::snprintf(b, sizeof b, "\t%zu\n", size_t(card[i]));   // :382
        s.append(b, n);
    }
    return s;
}

std::string format_ut_tsv_header(const std::vector<std::string> &paths) {
    std::string s("##Names\t");                                   // :388-393
    for (auto &p : paths) { s += p; s += '\t'; }
    s.back() = '\n';
    return s;
}

void append_ut_row(std::string &buf, const std::string &name, const float *row, size_t n, size_t index, EmissionFormat fmt) {
    buf += name;                                                  // submit_emit_dists, :16-35
    if (fmt == UT_TSV) {
        for (size_t k = 0; k < index + 1; ++k) buf += "\t-";
    } else if (name.size() < 9) {
        buf.append(9 - name.size(), ' ');
    }
    buf.reserve(buf.size() + (n - index - 1) * 10 + 2);
    for (size_t k = 0; k < n - index - 1; ++k) { buf += '\t'; append_g6(buf, row[k]); }
    buf += '\n';
}

std::string format_symmetric(const std::vector<std::string> &paths, const float *packed, EmissionFormat fmt, const float *packed_lower) {
    const size_t n = paths.size();
    std::string s;
    auto rowp = [n](const float *base, size_t i) { return base + (i * (2 * n - i - 1)) / 2; };
    if (fmt == UT_TSV || fmt == UPPER_TRIANGULAR) {
        if (fmt == UT_TSV) s = format_ut_tsv_header(paths);
        else s = std::to_string(n) + "\n";                       // :394-397
        for (size_t i = 0; i < n; ++i) append_ut_row(s, paths[i], rowp(packed, i), n, i, fmt);
        return s;
    }
    if (fmt != FULL_TSV) throw Error("Invalid emit_fmt");
    if (!packed_lower) packed_lower = packed;
    s = "#Names";                                                 // :853-857 — no separator after "#Names" in the reference
    for (size_t i = 0; i < n; ++i) { s += paths[i]; s += (i == n - 1 ? '\n' : '\t'); }
    for (size_t i = 0; i < n; ++i) {
        s += paths[i]; s += '\t';
        for (size_t j = 0; j < n; ++j) {
            const double v = j == i ? 0. : (i < j ? rowp(packed, i)[j - i - 1] : rowp(packed_lower, j)[i - j - 1]);
            append_g6(s, v);
            s += (j == n - 1 ? '\n' : '\t');
        }
    }
    return s;
}

std::string format_rect_row(const std::string &qname, const float *row, size_t nr) {
    std::string s(qname);                                         // src/dashing.h:695-699
    for (size_t i = 0; i < nr; ++i) { s += '\t'; append_g6(s, row[i]); }
    s += '\n';
    return s;
}

void write_binary_matrix(std::FILE *fp, const float *packed, uint64_t n) {
    std::fputc('\0', fp);                                         // distmat magic for float, distmat.h:188-208
    if (std::fwrite(&n, sizeof n, 1, fp) != 1) throw Error("Failure");
    const uint64_t cnt = n * (n - 1) / 2;
    if (cnt && std::fwrite(packed, sizeof(float), cnt, fp) != cnt) throw Error("Error writing to binary file");
}

// ---------------------------------------------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Genome {                 // the records of all files behind one input path
    std::string bases;
    std::vector<uint64_t> offs{0};
};

Genome load_genome(const std::string &path) {
    Genome g;
    for (auto &f : split_paths(path))
        for_each_record(f, [&](const char *s, size_t l) { g.bases.append(s, l); g.offs.push_back(g.bases.size()); });
    return g;
}

// One input file of a batch: parsed straight into its window of the batch arena.
struct FileSlot {
    size_t genome = 0;          // index into the batch
    std::string path;
    size_t cap = 0, off = 0, used = 0;
    std::vector<uint64_t> ends; // record ends, relative to `off`
    bool redo = false;          // window too small (multi-member gzip, ...): the genome goes through load_genome instead
};

// Upper bound on the sequence bytes of a file: its size, or for gzip the ISIZE trailer (RFC 1952; wrong for multi-member
// files such as bgzip output — those overflow their window and take the load_genome path).
size_t file_capacity(const std::string &f) {
    struct stat st;
    if (::stat(f.c_str(), &st) != 0) throw Error("Could not open file at " + f + ". Abort!");
    size_t cap = (size_t)st.st_size;
    std::FILE *fp = std::fopen(f.c_str(), "rb");
    if (!fp) throw Error("Could not open file at " + f + ". Abort!");
    unsigned char magic[2] = {0, 0};
    if (std::fread(magic, 1, 2, fp) == 2 && magic[0] == 0x1f && magic[1] == 0x8b && st.st_size >= 18) {
        uint32_t isize = 0;
        if (std::fseek(fp, -4, SEEK_END) == 0 && std::fread(&isize, 4, 1, fp) == 1) cap = isize;
    }
    std::fclose(fp);
    return cap;
}

struct Batch {
    std::vector<size_t> which;          // genome -> index into the caller's path list
    std::vector<FileSlot> slots;
    std::vector<uint64_t> file_off, file_len, gfb;
    std::vector<uint8_t> regs, status;
    std::vector<size_t> redo;           // batch genomes to be re-read through load_genome
    size_t bytes = 0;
};

// The raw bytes of a file (inflated if it is gzip) straight into a caller window; SIZE_MAX when the window is too small.
size_t slurp_into_window(const std::string &file, char *dst, size_t cap) {
    std::FILE *fp = std::fopen(file.c_str(), "rb");
    if (!fp) throw Error("Could not open file at " + file + ". Abort!");
    unsigned char magic[2] = {0, 0};
    const bool gz = std::fread(magic, 1, 2, fp) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    size_t used = 0;
    if (!gz) {
        std::rewind(fp);
        for (;;) {
            const size_t n = std::fread(dst + used, 1, cap - used, fp);
            used += n;
            if (n == 0 || used == cap) break;
        }
        const bool more = used == cap && std::fgetc(fp) != EOF;     // the file grew since it was measured
        std::fclose(fp);
        return more ? SIZE_MAX : used;
    }
    std::fclose(fp);
    gzFile g = gzopen(file.c_str(), "rb");
    if (!g) throw Error("Could not open file at " + file + ". Abort!");
    gzbuffer(g, 1 << 18);
    for (;;) {
        if (used == cap) {
            char extra;
            const int n = gzread(g, &extra, 1);                      // multi-member gzip: ISIZE covers the last member only
            gzclose(g);
            return n > 0 ? SIZE_MAX : used;
        }
        const int n = gzread(g, dst + used, (unsigned)std::min<size_t>(cap - used, 1u << 30));
        if (n < 0) { gzclose(g); throw Error("Error reading from file " + file); }
        if (n == 0) break;
        used += (size_t)n;
    }
    gzclose(g);
    return used;
}

// Sketch a list of paths.  The host only moves bytes: every file's raw text (inflated if gzip) is read in parallel straight
// into its window of one reusable arena per batch, and db200_sketch_fasta_batch parses, packs and sketches it on the GPU
// (kseq's record rules run there: dashing_b200/csrc/fasta.cuh).  The GPU call + register hand-off of batch b run on a helper
// thread while batch b+1 is being read.  Genomes holding a FASTQ file (by its first byte, or flagged by the device), or a
// file that outgrew its window, take the kseq-compatible host reader + db200_sketch_batch instead.
void sketch_paths(const SketchOptions &o, const std::vector<std::string> &paths, const std::vector<size_t> &which,
                  const std::function<void(size_t, const uint8_t *)> &sink) {
    if (which.empty()) return;
    const size_t m = size_t(1) << o.p;
    const int nt = std::max(1, o.nthreads);
    const size_t AL = DB200_FASTA_ALIGN;
    // CUDA initialisation (seconds on a multi-GPU node) overlaps the reading of the first batch
    std::future<std::string> warm = std::async(std::launch::async, [&o] {   // (the error text is thread-local: carry it over)
        return db200_warmup(o.device) == DB200_OK ? std::string() : std::string(db200_last_error());
    });
    // Two arenas (one being read into, one being sketched).  DB200_PINNED_ARENA=1 page-locks them (db200_host_alloc) so the
    // upload DMAs straight from them; that costs the page-locking time up front and needs the CUDA context first.
    static const bool pinned = std::getenv("DB200_PINNED_ARENA") != nullptr && std::atoi(std::getenv("DB200_PINNED_ARENA")) != 0;
    struct Arena {
        char *ptr = nullptr; size_t cap = 0; bool pinned = false;
        void release() { if (ptr) { if (pinned) db200_host_free(ptr); else delete[] ptr; } ptr = nullptr; cap = 0; }
        ~Arena() { release(); }
    } arena[2];
    std::future<void> inflight;
    int cur = 0;
    size_t at = 0;
    auto finish = [&](std::future<void> &f) { if (f.valid()) f.get(); };
    try {
        while (at < which.size()) {
            phase("batch begin");
            auto bt = std::make_shared<Batch>();
            // ---- batch layout from file sizes: every file on its own DB200_FASTA_ALIGN boundary
            bt->gfb.assign(1, 0);
            while (at < which.size() && (bt->which.empty() || bt->bytes < o.batch_bytes)) {
                for (auto &f : split_paths(paths[which[at]])) {
                    FileSlot sl;
                    sl.genome = bt->which.size(); sl.path = f; sl.cap = file_capacity(f); sl.off = bt->bytes;
                    bt->bytes += std::max(AL, (sl.cap + AL - 1) / AL * AL);
                    bt->slots.push_back(std::move(sl));
                }
                bt->which.push_back(which[at++]);
                bt->gfb.push_back(bt->slots.size());
            }
            if (arena[cur].cap < bt->bytes + 64) {
                arena[cur].release();
                // (later batches are never larger than batch_bytes + one genome: size the arena for that once)
                const size_t want = std::max(bt->bytes + 64, at < which.size() ? o.batch_bytes + (o.batch_bytes >> 2) : size_t(0));
                if (pinned) {
                    if (warm.valid()) { const std::string werr = warm.get(); if (!werr.empty()) throw Error(werr); }
                    void *pp = nullptr;
                    check(db200_host_alloc(&pp, want));
                    arena[cur].ptr = static_cast<char *>(pp); arena[cur].pinned = true;
                } else arena[cur].ptr = new char[want];
                arena[cur].cap = want;
            }
            char *base = arena[cur].ptr;
            // ---- read (file IO and gz inflate stay the host's job, as in the reference; nothing is parsed here)
            std::string err;
#pragma omp parallel for schedule(dynamic) num_threads(nt)
            for (size_t i = 0; i < bt->slots.size(); ++i) {
                FileSlot &sl = bt->slots[i];
                try {
                    const size_t used = slurp_into_window(sl.path, base + sl.off, sl.cap);
                    if (used == SIZE_MAX) { sl.redo = true; sl.used = 0; }
                    else {
                        sl.used = used;
                        size_t j = 0;
                        while (j < used && (base[sl.off + j] == '\n' || base[sl.off + j] == '\r')) ++j;
                        if (j < used && base[sl.off + j] == '@') { sl.redo = true; sl.used = 0; }   // FASTQ: the record interface
                    }
                } catch (const std::exception &e) {
#pragma omp critical
                    err = e.what();
                }
            }
            if (!err.empty()) throw Error(err);
            phase("files read");
            bt->file_off.resize(bt->slots.size());
            bt->file_len.resize(bt->slots.size());
            for (size_t i = 0; i < bt->slots.size(); ++i) { bt->file_off[i] = bt->slots[i].off; bt->file_len[i] = bt->slots[i].used; }
            bt->status.assign(bt->slots.size(), 0);
            bt->regs.resize(bt->which.size() * m);
            // ---- previous batch must be through before its arena's twin is reused two batches later; hand this one over
            finish(inflight);
            if (warm.valid()) { const std::string werr = warm.get(); if (!werr.empty()) throw Error(werr); }
            phase("previous batch done");
            inflight = std::async(std::launch::async, [bt, base, m, nt, &o, &paths, &sink] {
                check(db200_sketch_fasta_batch(o.device, o.p, o.k, o.canon, base, bt->file_off.data(), bt->file_len.data(), bt->slots.size(),
                                               bt->gfb.data(), bt->which.size(), bt->regs.data(), bt->status.data()));
                std::vector<char> redo_genome(bt->which.size(), 0);
                for (size_t i = 0; i < bt->slots.size(); ++i)
                    if (bt->slots[i].redo || bt->status[i]) redo_genome[bt->slots[i].genome] = 1;
                for (size_t g = 0; g < redo_genome.size(); ++g) if (redo_genome[g]) bt->redo.push_back(g);
                if (!bt->redo.empty()) {
                    // FASTQ, multi-member gzip, ...: the kseq-compatible host reader and the record interface
                    std::string all;
                    std::vector<uint64_t> ro{0}, gb{0};
                    for (size_t g : bt->redo) {
                        Genome gn = load_genome(paths[bt->which[g]]);
                        const uint64_t b0 = all.size();
                        all += gn.bases;
                        for (size_t r = 1; r < gn.offs.size(); ++r) ro.push_back(b0 + gn.offs[r]);
                        gb.push_back(ro.size() - 1);
                    }
                    if (all.empty()) all.push_back('N');
                    std::vector<uint8_t> rr(bt->redo.size() * m);
                    check(db200_sketch_batch(o.device, o.p, o.k, o.canon, all.data(), ro.data(), ro.size() - 1, gb.data(), bt->redo.size(), rr.data()));
                    for (size_t j = 0; j < bt->redo.size(); ++j) std::memcpy(&bt->regs[bt->redo[j] * m], &rr[j * m], m);
                }
                std::string serr;
#pragma omp parallel for schedule(dynamic) num_threads(nt)
                for (size_t g = 0; g < bt->which.size(); ++g) {
                    try { sink(bt->which[g], bt->regs.data() + g * m); }
                    catch (const std::exception &e) {
#pragma omp critical
                        serr = e.what();
                    }
                }
                if (!serr.empty()) throw Error(serr);
            });
            cur ^= 1;
        }
        finish(inflight);
        phase("last batch done");
    } catch (...) {
        if (inflight.valid()) { try { inflight.get(); } catch (...) {} }
        if (warm.valid()) warm.wait();
        throw;
    }
}

} // namespace

size_t file_window(const std::string &f) { return file_capacity(f); }
size_t slurp_file(const std::string &f, char *dst, size_t cap) { return slurp_into_window(f, dst, cap); }

void sketch_core(const SketchOptions &o, std::vector<std::string> paths) {
    if (!o.avoid_sorting) sort_paths_by_fsize(paths);
    std::vector<std::string> fnames(paths.size());
    std::vector<size_t> todo;
    const size_t m = size_t(1) << o.p;
    const bool container = !o.output_file.empty();
    // container mode: sketchdest (src/sketch_and_cmp.h:458-463) and the header each entry will be written with
    std::vector<uint8_t> all;
    struct Hdr { int estim; double value; };
    std::vector<Hdr> hdr(paths.size(), Hdr{o.estim, -1.});
    if (container) {
        // labels first (:470-478), then every sketch into one gzip stream (:527-535)
        gzFile lf = gzopen((o.output_file + ".labels.gz").c_str(), "w");
        if (!lf) throw Error("Failed to write sequence labels to file");
        for (auto &pth : paths) { gzwrite(lf, pth.data(), (unsigned)pth.size()); gzputc(lf, '\n'); }
        gzclose(lf);
        all.assign(paths.size() * m, 0);
    }
    for (size_t i = 0; i < paths.size(); ++i) {
        fnames[i] = make_fname(paths[i].c_str(), o.p, o.k, o.k, o.k, "", o.suffix, o.prefix);
        if (o.skip_cached && isfile(fnames[i])) {                   // src/sketch_and_cmp.h:499-504
            if (!container) continue;
            // h.read(fname) with no `continue` (:500-503): the path's k-mers are then added ON TOP of the cached registers,
            // and the value read() computed (hll.h:1078) is never invalidated (add() does not call not_ready()), so the entry
            // is written as "calculated" with the cached sketch's estimate under the cached file's estimator
            const HllFile h = read_hll(fnames[i]);
            if (h.p != (uint32_t)o.p) throw Error("cached sketch " + fnames[i] + " has the wrong size");
            std::memcpy(&all[i * m], h.core.data(), m);
            hdr[i].estim = (int)h.estim;
            if (h.value >= 0.) hdr[i].value = h.value;
            else {
                if (h.estim > 2) throw Error("cached sketch " + fnames[i] + " names an unknown estimation method");
                check(db200_cardinalities(o.device, h.core.data(), 1, o.p, (int)h.estim, &hdr[i].value));
            }
        }
        todo.push_back(i);
    }
    if (!container) {
        sketch_paths(o, paths, todo, [&](size_t i, const uint8_t *regs) { write_hll(fnames[i], regs, o.p, o.estim, o.jestim, -1.); });
        return;
    }
    sketch_paths(o, paths, todo, [&](size_t i, const uint8_t *regs) {
        uint8_t *dst = &all[i * m];
        for (size_t r = 0; r < m; ++r) dst[r] = std::max(dst[r], regs[r]);   // a plain copy unless a cached sketch was read first
    });
    gzFile fp = gzopen(o.output_file.c_str(), "w");
    if (!fp) throw Error("Failed to write sketches to file");
    // sketchdest[i] = h copies estim_ but not jestim_ (hll.h:947-956): the entry keeps the command line's joint estimator
    for (size_t i = 0; i < paths.size(); ++i) write_hll_stream(fp, &all[i * m], o.p, hdr[i].estim, o.jestim, hdr[i].value);
    gzclose(fp);
}

// TSV table of nndist_loop (src/sketch_and_cmp.h:741-781): "<path>(\t<index>:<%g value>)*"; the index goes through
// kstring's putw_(int), so the filler index uint32(-1) prints as -1.
std::string format_neighbors(const std::vector<std::string> &names, size_t qoffset, const Neighbor *nb, size_t rows, unsigned nn) {
    std::string s = "#File\tNeighbor ID:distance\t...\n";
    char buf[64];
    for (size_t i = 0; i < rows; ++i) {
        s += names[i + qoffset];
        for (unsigned j = 0; j < nn; ++j) {
            const Neighbor &e = nb[i * nn + j];
            std::snprintf(buf, sizeof buf, "\t%d:%g", (int)e.index, (double)e.value);
            s += buf;
        }
        s += '\n';
    }
    return s;
}

// Binary table (:733-740): uint32 number of paths, uint32 neighbours per row, then the (float, uint32) pairs.
void write_binary_neighbors(std::FILE *fp, uint32_t npaths, const Neighbor *nb, size_t rows, unsigned nn) {
    const uint32_t hdr[2] = {npaths, nn};
    static_assert(sizeof(Neighbor) == 8, "validx_t layout");
    if (std::fwrite(hdr, sizeof(uint32_t), 2, fp) != 2 || std::fwrite(nb, sizeof(Neighbor), rows * nn, fp) != rows * nn)
        throw Error("Failed to write neighbors to disk (binary)");
}

// dist_loop / partdist_loop / nndist_loop and their emitters (src/sketch_and_cmp.h:785-880, :712-783; src/dashing.h:660-712)
// over n = inpaths.size() sketches held as rows of `regs`; the last nq are queries.
void compare_and_emit(const DistOptions &o, const std::vector<std::string> &inpaths, const std::vector<uint8_t> &regs, size_t nq,
                      const double *cached_card) {
    const size_t n = inpaths.size(), m = size_t(1) << o.p;
    std::FILE *pfp = o.dist_path.empty() ? stdout : std::fopen(o.dist_path.c_str(), "wb");
    if (!pfp) throw Error("Could not open file at " + o.dist_path + " for writing.");
    // cached value_ of loaded sketches travel in the params (db200_dist_params.card); in the -Q/-F forms the references are the
    // first n - nq sketches and the queries the last nq
    db200_dist_params prm{o.p, o.k, o.estim, o.jestim, o.result_type, DB200_ORDER_ROW_FIRST, cached_card, (cached_card && nq) ? cached_card + (n - nq) : nullptr};
    const bool joint = o.jestim == DB200_ERTL_JOINT_MLE;
    if (o.nneighbors) {
        // nndist_loop (src/sketch_and_cmp.h:712-783): the all-pairs values never leave the GPU, only rows x nn pairs do
        if (nq >= n) throw Error("Wrong number of query/references.");
        const size_t rows = nq ? nq : n, possible = nq ? n - nq : n, qoffset = nq ? n - nq : 0;
        unsigned nn = o.nneighbors;
        if (nn > possible) {
            std::fprintf(stderr, "Only reporting %zu rather than %u neighbors due to their being only that many sets.\n", possible, nn);
            nn = (unsigned)possible;
        }
        std::vector<Neighbor> nb(rows * nn);
        static_assert(sizeof(Neighbor) == sizeof(db200_neighbor), "neighbour layout");
        prm.order = DB200_ORDER_COL_FIRST;                                     // func(sketches[j], h1), :670
        if (nq) check(db200_dist_knn_rect(o.device, regs.data(), n - nq, regs.data() + (n - nq) * m, nq, &prm, nn, reinterpret_cast<db200_neighbor *>(nb.data())));
        else check(db200_dist_knn_symmetric(o.device, regs.data(), n, &prm, nn, reinterpret_cast<db200_neighbor *>(nb.data())));
        // The reference's emitters loop over ALL paths even in the -Q/-F mode, where only nq rows exist (an out-of-bounds
        // read there); this writes the nq rows that do.
        if (o.emit_fmt == BINARY) write_binary_neighbors(pfp, (uint32_t)n, nb.data(), rows, nn);
        else { const std::string s = format_neighbors(inpaths, qoffset, nb.data(), rows, nn); std::fwrite(s.data(), 1, s.size(), pfp); }
    } else if (nq) {
        if (nq >= n) throw Error("Wrong number of query/references.");
        const size_t nr = n - nq;
        std::vector<float> out(nr * nq);
        check(db200_dist_rect(o.device, regs.data(), nr, regs.data() + nr * m, nq, &prm, out.data()));
        if (o.emit_fmt == UPPER_TRIANGULAR) std::fprintf(pfp, "%zu\n", n);     // :394-397 runs before dist_loop even in this mode
        for (size_t q = 0; q < nq; ++q) {
            if (o.emit_fmt == BINARY) { if (std::fwrite(&out[q * nr], sizeof(float), nr, pfp) != nr) throw Error("Error writing to binary file"); }
            else { const std::string s = format_rect_row(inpaths[nr + q], &out[q * nr], nr); std::fwrite(s.data(), 1, s.size(), pfp); }
        }
    } else {
        const int rt = o.result_type;
        if (rt == DB200_CONTAINMENT_INDEX || rt == DB200_CONTAINMENT_DIST || rt == DB200_FULL_CONTAINMENT_DIST)
            throw Error("Can't perform symmetric distance comparisons with a symmetric method. Provide the same list of filenames to both -Q and -F.");
        const size_t np = n * (n - 1) / 2;
        // operand order: TSV / PHYLIP rows come from perform_core_op (cmp(s[j], s[i])), binary and the upper half of FULL_TSV
        // from cmp(s[i], s[j]); only the joint MLE can tell the difference
        prm.order = (o.emit_fmt == UT_TSV || o.emit_fmt == UPPER_TRIANGULAR) ? DB200_ORDER_COL_FIRST : DB200_ORDER_ROW_FIRST;
        auto write_labels = [&] {
            if (o.dist_path.empty()) return;                                 // src/distmain.cpp:191-200
            std::FILE *lf = std::fopen((o.dist_path + ".labels").c_str(), "wb");
            if (!lf) throw Error("Could not open file at '" + o.dist_path + ".labels' for writing");
            for (auto &p : inpaths) { std::fwrite(p.data(), p.size(), 1, lf); std::fputc('\n', lf); }
            std::fclose(lf);
        };
        // Streaming forms (db200_dist_symmetric_stream): row blocks are written as they arrive, so the n(n-1)/2 floats are never
        // held in memory — the binary matrix into a regular file at its distmat offset (any number of devices), the
        // triangular text formats in row order (one device).
        const bool one_device = o.device != DB200_ALL_DEVICES || db200_device_count() < 2;
        struct StreamCtx {
            const std::vector<std::string> *names; size_t n; EmissionFormat fmt; std::FILE *fp; int fd; int nthreads; std::string err;
        } sc{&inpaths, n, o.emit_fmt, pfp, -1, std::max(1, o.nthreads), {}};
        if (o.emit_fmt == BINARY && !o.dist_path.empty() && n >= 2) {
            const uint64_t n64 = n;
            std::fputc('\0', pfp);                                           // distmat magic for float + u64 n (distmat.h:188-208)
            if (std::fwrite(&n64, sizeof n64, 1, pfp) != 1) throw Error("Failure");
            std::fflush(pfp);
            sc.fd = fileno(pfp);
            if (::ftruncate(sc.fd, (off_t)(9 + np * sizeof(float))) != 0) throw Error("Error writing to binary file");
            const int rc = db200_dist_symmetric_stream(o.device, regs.data(), n, &prm, 0, n, 0,
                [](void *ud, uint64_t rb, uint64_t, const float *vals, uint64_t nv) -> int {
                    StreamCtx *c = static_cast<StreamCtx *>(ud);
                    const uint64_t nn = c->n, off = 9 + 4 * ((rb * (2 * nn - rb - 1)) / 2);
                    const char *src = reinterpret_cast<const char *>(vals);
                    for (uint64_t done = 0; done < nv * 4;) {
                        const ssize_t w = ::pwrite(c->fd, src + done, nv * 4 - done, (off_t)(off + done));
                        if (w <= 0) return 1;
                        done += (uint64_t)w;
                    }
                    return 0;
                }, &sc);
            if (rc != DB200_OK) throw Error(std::string("Error writing to binary file: ") + db200_last_error());
            write_labels();
        } else if ((o.emit_fmt == UT_TSV || o.emit_fmt == UPPER_TRIANGULAR) && one_device && n >= 2) {
            {
                const std::string h = o.emit_fmt == UT_TSV ? format_ut_tsv_header(inpaths) : std::to_string(n) + "\n";   // :388-397
                std::fwrite(h.data(), 1, h.size(), pfp);
            }
            const int rc = db200_dist_symmetric_stream(o.device, regs.data(), n, &prm, 0, n, 0,
                [](void *ud, uint64_t rb, uint64_t re, const float *vals, uint64_t) -> int {
                    StreamCtx *c = static_cast<StreamCtx *>(ud);
                    const uint64_t nn = c->n;
                    auto tri = [nn](uint64_t r) { return (r * (2 * nn - r - 1)) / 2; };
                    std::vector<std::string> rows(re - rb);
#pragma omp parallel for schedule(dynamic, 4) num_threads(c->nthreads)
                    for (uint64_t i = rb; i < re; ++i) append_ut_row(rows[i - rb], (*c->names)[i], vals + (tri(i) - tri(rb)), nn, i, c->fmt);
                    for (auto &r : rows) if (std::fwrite(r.data(), 1, r.size(), c->fp) != r.size()) return 1;
                    return 0;
                }, &sc);
            if (rc != DB200_OK) throw Error(std::string("Error writing distances: ") + db200_last_error());
        } else {
            std::vector<float> out(std::max<size_t>(np, 1)), lower;
            if (n >= 2) { check(db200_dist_symmetric(o.device, regs.data(), n, &prm, out.data())); }
            if (o.emit_fmt == BINARY) {
                write_binary_matrix(pfp, out.data(), n);
                write_labels();
            } else {
                if (o.emit_fmt == FULL_TSV && joint && n >= 2) {
                    lower.resize(np);
                    prm.order = DB200_ORDER_COL_FIRST;
                    check(db200_dist_symmetric(o.device, regs.data(), n, &prm, lower.data()));
                }
                const std::string s = format_symmetric(inpaths, out.data(), o.emit_fmt, lower.empty() ? nullptr : lower.data());
                std::fwrite(s.data(), 1, s.size(), pfp);
            }
        }
    }
    if (pfp != stdout) std::fclose(pfp); else std::fflush(pfp);
    phase("distances written");
}

// hll_t::read() computes the cardinality at once under the FILE's estimator (csum(), hll.h:1078) unless the file already
// stores one, and the reference keeps that cached value_ for the sizes file and for the per-sketch terms of every pair,
// whatever -E/-I/-m says on the command line (set_estim_and_jestim only changes what LATER evaluations use).
struct Cached { bool loaded = false; uint32_t estim = 2; double value = -1.; };
static Cached cached_of(const HllFile &h) { return Cached{true, h.estim, h.value >= 0. ? h.value : -1.}; }   // is_calculated() is value_ >= 0 (hll.h:1038)

// card[] arrives evaluated under `estim` for every sketch; entries of loaded sketches are replaced by their cached value.
// Returns whether any sketch was loaded (the pair loop then needs db200_dist_use_cardinalities).
static bool apply_cached(int device, int p, int estim, const std::vector<uint8_t> &regs, const std::vector<Cached> &info, std::vector<double> &card) {
    bool any = false;
    const size_t m = size_t(1) << p;
    for (size_t i = 0; i < info.size(); ++i) {
        if (!info[i].loaded) continue;
        any = true;
        if (info[i].value >= 0.) card[i] = info[i].value;
        else if ((int)info[i].estim != estim) {
            if (info[i].estim > 2) throw Error("a loaded sketch names an unknown estimation method");
            check(db200_cardinalities(device == DB200_ALL_DEVICES ? 0 : device, &regs[i * m], 1, p, (int)info[i].estim, &card[i]));
        }
    }
    return any;
}

// Phase A of dist_sketch_and_cmp / size_sketch_and_emit (src/sketch_and_cmp.h:314-360, :155-183): sketches that exist as files
// (--presketched, or -W with the file present) are read — in parallel, as the reference's OpenMP loop does — and the rest are
// listed in `todo` for sketch_paths.
static void load_existing(const DistOptions &o, const std::vector<std::string> &inpaths, std::vector<uint8_t> &regs, std::vector<Cached> &info,
                          std::vector<std::string> &fnames, std::vector<size_t> &todo) {
    const size_t n = inpaths.size(), m = size_t(1) << o.p;
    std::vector<char> need(n, 0);
    std::string err;
#pragma omp parallel for schedule(dynamic, 16) num_threads(std::max(1, o.nthreads))
    for (size_t i = 0; i < n; ++i) {
        try {
            std::string from;
            if (o.presketched) from = inpaths[i];
            else {
                fnames[i] = make_fname(inpaths[i].c_str(), o.p, o.k, o.k, o.k, "", o.suffix, o.prefix);
                if (o.cache_sketches && isfile(fnames[i])) from = fnames[i];
            }
            if (from.empty()) { need[i] = 1; continue; }
            const HllFile h = read_hll(from);
            if (h.p != (uint32_t)o.p) throw Error("sketch " + from + " has p=" + std::to_string(h.p) + ", expected " + std::to_string(o.p));
            info[i] = cached_of(h);
            std::memcpy(&regs[i * m], h.core.data(), m);
        } catch (const std::exception &e) {
#pragma omp critical
            err = e.what();
        }
    }
    if (!err.empty()) throw Error(err);
    for (size_t i = 0; i < n; ++i) if (need[i]) todo.push_back(i);
}

void dist_sketch_and_cmp(const DistOptions &o_in, std::vector<std::string> inpaths, size_t nq) {
    DistOptions o = o_in;
    if (o.defer_hll) o.estim = o.jestim = DB200_ERTL_MLE;   // hllbase_t(p) inside make_hll(): ERTL_MLE for both (hll.h:765)
    const size_t n = inpaths.size(), m = size_t(1) << o.p;
    if (nq > n) throw Error("more queries than paths");
    // the CUDA context (0.6 s and more) comes up while the sketch files are being read
    std::future<int> warm = std::async(std::launch::async, [dev = o.device] { return db200_warmup(dev); });
    std::vector<uint8_t> regs(n * m);
    // ---- phase A: load or sketch (src/sketch_and_cmp.h:314-360)
    std::vector<size_t> todo;
    std::vector<std::string> fnames(n);
    std::vector<Cached> info(n);
    load_existing(o, inpaths, regs, info, fnames, todo);
    sketch_paths(o, inpaths, todo, [&](size_t i, const uint8_t *r) {
        std::memcpy(&regs[i * m], r, m);
        // sketches carry the command line's estimators (set_estim_and_jestim, src/sketch_and_cmp.h:285-288)
        if (o.cache_sketches && !o.defer_hll) write_hll(fnames[i], r, o.p, o.estim, o.jestim, -1.);
    });
    // ---- phase B: sizes (:372-385)
    phase("sketches ready");
    warm.wait();
    std::vector<double> card(n);
    check(db200_cardinalities(o.device, regs.data(), n, o.p, o.estim, card.data()));
    const bool any_cached = apply_cached(o.device, o.p, o.estim, regs, info, card);
    {
        const std::string s = format_sizes(inpaths, card.data());
        std::FILE *fp = o.sizes_path.empty() ? stdout : std::fopen(o.sizes_path.c_str(), "w");
        if (!fp) throw Error("Could not open file at " + o.sizes_path + " for writing.");
        std::fwrite(s.data(), 1, s.size(), fp);
        if (fp != stdout) std::fclose(fp); else std::fflush(fp);
    }
    if (o.defer_hll && o.cache_sketches)     // final_sketches[i].write(fpath) after make_hll()'s sum() (bbmh.h:1175-1184): value included
        for (size_t i : todo) write_hll(fnames[i], &regs[i * m], o.p, 2, 2, card[i]);
    // ---- phase C: all pairs
    phase("sizes written");
    compare_and_emit(o, inpaths, regs, nq, any_cached ? card.data() : nullptr);
}

// ---------------------------------------------------------------------------------------------------------------
// CLI (hot subset of src/distmain.cpp:47-100 and src/dashing.cpp:307-337)
// ---------------------------------------------------------------------------------------------------------------
[[noreturn]] static void unsupported(const char *flag) {
    throw Error(std::string("flag ") + flag + " selects a code path outside the B200 engine (spaced/windowed k-mers, count-min, non-HLL sketches): "
                "run the reference binary for it");
}

// Option tables: the reference declares most boolean long options as getopt FLAG entries (LO_FLAG, src/dashing.h:33) that
// set a variable and return 0, and gives them a short letter that its switch often has no case for — so `--presketched`
// works while `-H` is accepted and ignored, `-n` swallows an argument and does nothing, etc.  The tables below reproduce
// that: long flags get private codes (>= 256), short letters do exactly what the reference's switch does with them.
enum LongOnly {
    L_AVOID_SORTING = 256, L_CACHE, L_BINARY, L_FULL_MASH, L_FULL_TSV, L_NO_CANON, L_PHYLIP, L_PRESKETCHED, L_SIZES, L_SCI, L_MASH,
    L_CONT_INDEX, L_CONT_DIST, L_FULL_CONT_DIST, L_SYM_CONT_INDEX, L_SYM_CONT_DIST, L_SKIP_CACHED, L_DEFER_HLL, L_DEVICE,
    L_COUNTMIN, L_BY_FNAME, L_ENTROPY, L_OTHER_SKETCH, L_OTHER_HASH, L_WJ, L_NN, L_IGNORED_ARG
};

static int parse_device(const char *arg) {
    return (!std::strcmp(arg, "all") || !std::strcmp(arg, "-1")) ? DB200_ALL_DEVICES : std::atoi(arg);
}

int dist_main(int argc, char **argv) {
    DistOptions o;
    std::string paths_file;
    std::vector<std::string> querypaths;
    static option longopts[] = {   // DIST_LONG_OPTS, src/dashing.h:46-108
        {"avoid-sorting", no_argument, nullptr, L_AVOID_SORTING}, {"by-entropy", no_argument, nullptr, L_ENTROPY},
        {"cache-sketches", no_argument, nullptr, L_CACHE}, {"countmin", no_argument, nullptr, L_COUNTMIN},
        {"emit-binary", no_argument, nullptr, L_BINARY}, {"full-mash-dist", no_argument, nullptr, L_FULL_MASH},
        {"full-tsv", no_argument, nullptr, L_FULL_TSV}, {"no-canon", no_argument, nullptr, L_NO_CANON}, {"phylip", no_argument, nullptr, L_PHYLIP},
        {"presketched", no_argument, nullptr, L_PRESKETCHED}, {"sizes", no_argument, nullptr, L_SIZES},
        {"sketch-by-fname", no_argument, nullptr, L_BY_FNAME}, {"use-bb-minhash", no_argument, nullptr, L_OTHER_SKETCH},
        {"use-scientific", no_argument, nullptr, L_SCI}, {"bbits", required_argument, nullptr, 'B'}, {"cm-sketch-size", required_argument, nullptr, 't'},
        // LO_ARG: these four demand an argument in their LONG form (src/dashing.h:62-65, :71)
        {"ertl-joint-mle", required_argument, nullptr, 'J'}, {"ertl-mle", required_argument, nullptr, 'm'}, {"improved", required_argument, nullptr, 'I'},
        {"original", required_argument, nullptr, 'E'},
        {"kmer-length", required_argument, nullptr, 'k'}, {"min-count", required_argument, nullptr, 'c'}, {"nhashes", required_argument, nullptr, 'q'},
        {"nthreads", required_argument, nullptr, 'p'}, {"out-dists", required_argument, nullptr, 'O'}, {"out-sizes", required_argument, nullptr, 'o'},
        {"paths", required_argument, nullptr, 'F'}, {"prefix", required_argument, nullptr, 'P'}, {"query-paths", required_argument, nullptr, 'Q'},
        {"seed", required_argument, nullptr, 'R'}, {"sketch-size", required_argument, nullptr, 'S'}, {"spacing", required_argument, nullptr, 's'},
        {"suffix", required_argument, nullptr, 'x'}, {"window-size", required_argument, nullptr, 'w'},
        {"use-range-minhash", no_argument, nullptr, L_OTHER_SKETCH}, {"use-full-khash-sets", no_argument, nullptr, L_OTHER_SKETCH},
        {"use-full-hash-sets", no_argument, nullptr, L_OTHER_SKETCH}, {"use-hash-sets", no_argument, nullptr, L_OTHER_SKETCH},
        {"hash-sets", no_argument, nullptr, L_OTHER_SKETCH}, {"use-full-sets", no_argument, nullptr, L_OTHER_SKETCH},
        {"use-bloom-filter", no_argument, nullptr, L_OTHER_SKETCH}, {"use-wide-hll", no_argument, nullptr, L_OTHER_SKETCH},
        {"use-nthash", no_argument, nullptr, L_OTHER_HASH}, {"use-cyclic-hash", no_argument, nullptr, L_OTHER_HASH},
        {"full-containment-dist", no_argument, nullptr, L_FULL_CONT_DIST}, {"containment-index", no_argument, nullptr, L_CONT_INDEX},
        {"containment-dist", no_argument, nullptr, L_CONT_DIST}, {"mash-dist", no_argument, nullptr, L_MASH},
        {"symmetric-containment-index", no_argument, nullptr, L_SYM_CONT_INDEX}, {"symmetric-containment-dist", no_argument, nullptr, L_SYM_CONT_DIST},
        {"wj", no_argument, nullptr, L_WJ}, {"wj-exact", no_argument, nullptr, L_WJ}, {"wj-cm-sketch-size", required_argument, nullptr, L_WJ},
        {"wj-cm-nhashes", required_argument, nullptr, L_WJ}, {"nearest-neighbors", required_argument, nullptr, L_NN},
        {"defer-hll", no_argument, nullptr, L_DEFER_HLL}, {"nperbatch", required_argument, nullptr, L_IGNORED_ARG},
        {"device", required_argument, nullptr, L_DEVICE},   // this engine's only addition: a GPU index, or "all"
        {nullptr, 0, nullptr, 0}};
    optind = 1;
    int co;
    // the reference's own short-option string (src/distmain.cpp:47)
    while ((co = getopt_long(argc, argv, "n:Q:P:x:F:c:p:o:s:w:O:S:k:=:t:R:D:8TgazlICbMEeHJhZBNyUmqW?", longopts, nullptr)) >= 0) {
        switch (co) {
            case L_AVOID_SORTING: o.avoid_sorting = true; break;
            case 'W': case L_CACHE: o.cache_sketches = true; break;
            case 'b': case L_BINARY: o.emit_fmt = BINARY; break;
            case 'l': case L_FULL_MASH: o.result_type = DB200_FULL_MASH_DIST; break;
            case 'T': case L_FULL_TSV: o.emit_fmt = FULL_TSV; break;
            case 'C': case L_NO_CANON: o.canon = false; break;
            case 'U': case L_PHYLIP: o.emit_fmt = UPPER_TRIANGULAR; break;
            case L_PRESKETCHED: o.presketched = true; break;           // the short -H has no case in the reference's switch
            case L_SIZES: o.result_type = DB200_SIZES; break;          // nor has -Z
            case 'M': case L_MASH: o.result_type = DB200_MASH_DIST; break;
            case 'J': o.jestim = DB200_ERTL_JOINT_MLE; break;
            case 'm': o.jestim = o.estim = DB200_ERTL_MLE; break;
            case 'I': o.jestim = o.estim = DB200_ERTL_IMPROVED; break;
            case 'E': o.jestim = o.estim = DB200_ORIGINAL; break;
            case 'k': o.k = std::atoi(optarg); break;
            case 'p': o.nthreads = std::atoi(optarg); break;
            case 'S': o.p = std::atoi(optarg); break;       // bytesl2_to_arg(S, HLL) == S, src/sketch_and_cmp.h:42
            case 'O': o.dist_path = optarg; break;
            case 'o': o.sizes_path = optarg; break;
            case 'F': paths_file = optarg; break;
            case 'Q': querypaths = get_paths(optarg); break;
            case 'P': o.prefix = optarg; break;
            case 'x': o.suffix = optarg; break;
            case L_CONT_INDEX: o.result_type = DB200_CONTAINMENT_INDEX; break;
            case L_CONT_DIST: o.result_type = DB200_CONTAINMENT_DIST; break;
            case L_FULL_CONT_DIST: o.result_type = DB200_FULL_CONTAINMENT_DIST; break;
            case L_SYM_CONT_INDEX: o.result_type = DB200_SYMMETRIC_CONTAINMENT_INDEX; break;
            case L_SYM_CONT_DIST: o.result_type = DB200_SYMMETRIC_CONTAINMENT_DIST; break;
            case L_DEFER_HLL: o.defer_hll = true; break;
            case L_DEVICE: o.device = parse_device(optarg); break;
            case 's': if (*optarg) unsupported("-s/--spacing"); break;
            case 'w': o.wsz = std::atoi(optarg); break;   // judged once k is final: wsz <= k is "unwindowed" (src/distmain.cpp:168)
            case L_COUNTMIN: unsupported("--countmin");
            case L_BY_FNAME: unsupported("--sketch-by-fname");
            case L_ENTROPY: case 'g': unsupported("-g/--by-entropy");
            case L_OTHER_SKETCH: case '8': unsupported("a non-HLL sketch type");
            case L_OTHER_HASH: unsupported("--use-nthash/--use-cyclic-hash");
            case L_WJ: unsupported("--wj (weighted Jaccard)");
            case L_NN:                                                            // src/distmain.cpp:89-93
                if (std::atoi(optarg) <= 0) throw Error("--nearest-neighbors needs a positive count");
                o.nneighbors = (unsigned)std::atoi(optarg);
                break;
            case 'h': case '?': throw Error("see the reference's `dashing dist` usage: this binary takes the same flags");
            default: break;   // -n X, -c X, -q X, -t X, -R X, -D X, -B X, -e, -H, -Z, -N, -y, -a, -z, --nperbatch: accepted, no effect on the HLL path
        }
    }
    if (o.k > 32) throw Error("k must be <= 32 for non-rolling hashes.");
    if (o.wsz > o.k) unsupported("-w/--window-size (windowed minimizers: w > k)");   // option order must not matter: `-w 30 -k 21` is windowed     // src/distmain.cpp:101-102
    if (o.nthreads < 1) o.nthreads = 1;
    std::vector<std::string> inpaths = paths_file.empty() ? std::vector<std::string>(argv + optind, argv + argc) : get_paths(paths_file);
    if (inpaths.empty()) throw Error("No paths. See usage.");
    size_t nq = querypaths.size();
    const int rt = o.result_type;
    if (nq == 0 && (rt == DB200_CONTAINMENT_INDEX || rt == DB200_CONTAINMENT_DIST || rt == DB200_FULL_CONTAINMENT_DIST)) {
        querypaths = inpaths; nq = querypaths.size();                          // :119-124
    }
    if (!o.presketched && !o.avoid_sorting) { sort_paths_by_fsize(inpaths); sort_paths_by_fsize(querypaths); }
    for (auto &q : querypaths) inpaths.push_back(q);
    dist_sketch_and_cmp(o, inpaths, nq);
    return 0;
}

// SKETCH_LONG_OPTS (src/dashing.cpp:253-291) and the switch of sketch_main / sketch_by_seq_main (:307-337, :487-517).
// Returns false for options the caller has to look at itself.
static option sketch_longopts[] = {
    {"countmin", no_argument, nullptr, L_COUNTMIN}, {"sketch-by-fname", no_argument, nullptr, L_BY_FNAME}, {"no-canon", no_argument, nullptr, L_NO_CANON},
    {"skip-cached", no_argument, nullptr, L_SKIP_CACHED}, {"by-entropy", no_argument, nullptr, L_ENTROPY}, {"use-bb-minhash", no_argument, nullptr, L_OTHER_SKETCH},
    {"bbits", required_argument, nullptr, 'B'}, {"paths", required_argument, nullptr, 'F'}, {"prefix", required_argument, nullptr, 'P'},
    {"nhashes", required_argument, nullptr, 'H'}, {"original", required_argument, nullptr, 'E'}, {"improved", required_argument, nullptr, 'I'},
    {"ertl-joint-mle", required_argument, nullptr, 'J'}, {"seed", required_argument, nullptr, 'R'}, {"sketch-size", required_argument, nullptr, 'S'},
    {"kmer-length", required_argument, nullptr, 'k'}, {"min-count", required_argument, nullptr, 'n'}, {"nthreads", required_argument, nullptr, 'p'},
    {"cm-sketch-size", required_argument, nullptr, 'q'}, {"spacing", required_argument, nullptr, 's'}, {"window-size", required_argument, nullptr, 'w'},
    {"suffix", required_argument, nullptr, 'x'}, {"wj-cm-sketch-size", required_argument, nullptr, L_WJ}, {"wj-cm-nhashes", required_argument, nullptr, L_WJ},
    {"use-range-minhash", no_argument, nullptr, L_OTHER_SKETCH}, {"use-full-khash-sets", no_argument, nullptr, L_OTHER_SKETCH},
    {"use-bloom-filter", no_argument, nullptr, L_OTHER_SKETCH}, {"use-wide-hll", no_argument, nullptr, L_OTHER_SKETCH},
    {"use-nthash", no_argument, nullptr, L_OTHER_HASH}, {"use-cyclic-hash", no_argument, nullptr, L_OTHER_HASH},
    {"avoid-sorting", no_argument, nullptr, L_AVOID_SORTING}, {"wj", no_argument, nullptr, L_WJ}, {"wj-exact", no_argument, nullptr, L_WJ},
    {"defer-hll", no_argument, nullptr, L_DEFER_HLL}, {"device", required_argument, nullptr, L_DEVICE},
    {nullptr, 0, nullptr, 0}};

static bool sketch_option(int co, SketchOptions &o, bool &defer_hll) {
    switch (co) {
        case L_AVOID_SORTING: o.avoid_sorting = true; return true;
        case L_SKIP_CACHED: o.skip_cached = true; return true;         // the short -c / -C / -e / -f / -j have no case in the reference's switch
        case L_NO_CANON: o.canon = false; return true;
        case 'E': o.jestim = o.estim = DB200_ORIGINAL; return true;
        case 'I': o.jestim = o.estim = DB200_ERTL_IMPROVED; return true;
        case 'J': o.jestim = DB200_ERTL_JOINT_MLE; return true;
        case 'k': o.k = std::atoi(optarg); return true;
        case 'p': o.nthreads = std::atoi(optarg); return true;
        case 'S': o.p = std::atoi(optarg); return true;
        case 'P': o.prefix = optarg; return true;
        case 'x': o.suffix = optarg; return true;
        case L_DEVICE: o.device = parse_device(optarg); return true;
        case L_DEFER_HLL: defer_hll = true; return true;
        case 's': if (*optarg) unsupported("-s/--spacing"); return true;
        case 'w': o.wsz = std::atoi(optarg); return true;   // judged by the caller once k is final
        case 'b': case L_COUNTMIN: unsupported("-b/--countmin");
        case L_BY_FNAME: unsupported("--sketch-by-fname");
        case L_ENTROPY: unsupported("--by-entropy");
        case L_OTHER_SKETCH: case '8': unsupported("a non-HLL sketch type");
        case L_OTHER_HASH: unsupported("--use-nthash/--use-cyclic-hash");
        case L_WJ: unsupported("--wj (weighted Jaccard)");
        case 'h': case '?': throw Error("see the reference's `dashing sketch` usage: this binary takes the same flags");
        default: return false;
    }
}

int sketch_main(int argc, char **argv) {
    SketchOptions o;
    std::string paths_file;
    bool defer_hll = false;
    optind = 1;
    int co;
    while ((co = getopt_long(argc, argv, "n:P:F:o:p:x:R:s:S:k:w:H:q:B:8JbfjEIcCeh?", sketch_longopts, nullptr)) >= 0) {   // src/dashing.cpp:307
        if (sketch_option(co, o, defer_hll)) continue;
        if (co == 'F') paths_file = optarg;
        else if (co == 'o') o.output_file = optarg;
        // -n X (min count), -R X, -H X, -q X, -B X, -c, -C, -e, -f, -j: accepted, no effect on the HLL path
    }
    // sketch_core<HyperLogLogHasher<>>::write is BBitMinHasher's: it emits a b-bit minhash under the .hll name (bbmh.h:962-969)
    if (defer_hll) unsupported("--defer-hll (`sketch` then writes b-bit minhash files, not HLLs)");
    if (o.k > 32) throw Error("k must be <= 32 for non-rolling hashes.");
    if (o.wsz > o.k) unsupported("-w/--window-size (windowed minimizers: w > k)");   // option order must not matter: `-w 30 -k 21` is windowed
    o.nthreads = std::max(o.nthreads, 1);
    std::vector<std::string> inpaths = (!paths_file.empty() && isfile(paths_file)) ? get_paths(paths_file) : std::vector<std::string>(argv + optind, argv + argc);
    if (inpaths.empty()) throw Error("No paths. See usage.");
    sketch_core(o, inpaths);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// union / hll / fold / view / card / sketch_by_seq / dist_by_seq
// ---------------------------------------------------------------------------------------------------------------
static void write_hll_mode(const std::string &path, const char *mode, const uint8_t *regs, uint32_t p, int estim, int jestim, double value) {
    gzFile fp = gzopen(path.c_str(), mode);
    if (!fp) throw Error("Could not open file at " + path);
    write_hll_stream(fp, regs, p, estim, jestim, value);
    if (gzclose(fp) != Z_OK) throw Error("Failed to close ofp");
}

// union_main + union_core<hll_t>, src/union.cpp:33-108
int union_main(int argc, char **argv) {
    int compression_level = 6, device = 0, nthreads = 1;
    std::string opath = "/dev/stdout";
    std::vector<std::string> paths;
    static option longopts[] = {{"device", required_argument, nullptr, L_DEVICE}, {nullptr, 0, nullptr, 0}};
    optind = 1;
    for (int c; (c = getopt_long(argc, argv, "p:b:o:F:zZ:h?", longopts, nullptr)) >= 0;) {
        switch (c) {
            case 'h': case '?': throw Error("Usage: union [-o out.hll] [-Z level] [-F paths.txt] sketch1.hll <sketch2.hll>...");
            case 'Z': compression_level = std::atoi(optarg); [[fallthrough]];    // the reference falls through here, so -Z N also names
            case 'o': opath = optarg; break;                                     // the output file N unless a later -o overrides it
            case 'F': paths = get_paths(optarg); break;
            case 'b': unsupported("-b (the reference's getopt string gives it an argument and its switch then selects bloom filters)");
            case L_DEVICE: device = parse_device(optarg); break;
            case 'p': nthreads = std::max(1, std::atoi(optarg)); break;          // threads the reading of the inputs, as there
            default: break;                                                      // -z: nothing to do here
        }
    }
    for (int i = optind; i < argc; ++i) paths.push_back(argv[i]);
    if (paths.empty()) throw Error("require >= 1 paths. See usage.");
    // T(paths[i]) for every input, then += (element-wise max, hll.h:958-992) and sum() (perform_sum): the result keeps the FIRST
    // sketch's estimators and carries its cardinality
    const HllFile first = read_hll(paths[0]);
    const uint32_t p = first.p, estim = first.estim, jestim = first.jestim;
    const size_t m = size_t(1) << p;
    std::vector<uint8_t> regs(paths.size() * m);
    std::memcpy(regs.data(), first.core.data(), m);
    std::string err;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (size_t i = 1; i < paths.size(); ++i) {
        try {
            const HllFile h = read_hll(paths[i]);
            if (h.p != p) throw Error("mismatched sketch sizes.");            // PREC_REQ, hll.h:959
            std::memcpy(&regs[i * m], h.core.data(), m);
        } catch (const std::exception &e) {
#pragma omp critical
            err = e.what();
        }
    }
    if (!err.empty()) throw Error(err);
    if (estim > 2) throw Error("sketch " + paths[0] + " names an unknown estimation method");
    std::vector<uint8_t> out(size_t(1) << p);
    check(db200_union(device, regs.data(), paths.size(), (int)p, out.data()));
    double value = -1.;
    check(db200_cardinalities(device, out.data(), 1, (int)p, (int)estim, &value));
    char mode[8];
    if (compression_level) std::snprintf(mode, sizeof mode, "wb%d", compression_level % 23);
    else std::snprintf(mode, sizeof mode, "wT");
    write_hll_mode(opath, mode, out.data(), p, (int)estim, (int)jestim, value);
    return 0;
}

// hll_main -> estimate_cardinality -> make_hll -> fill_sketch, src/hllmain.cpp:4-40, src/dashing.h:618-657: every path folded into
// ONE sketch of 2^S registers (default S = 24), its ERTL_MLE estimate truncated to u64
int hll_main(int argc, char **argv) {
    SketchOptions o;
    o.p = 24;
    int wsz = 0;
    std::string spacing;
    if (argc < 2) throw Error("Usage: hll <opts> <paths>\nFlags:\n-k:\tkmer length (Default: 31. Max: 32)\n-S:\tsketch size (default: 24)\n-p:\tnumber of threads.\n-C:\tdo not canonicalize");
    static option longopts[] = {{"device", required_argument, nullptr, L_DEVICE}, {nullptr, 0, nullptr, 0}};
    optind = 1;
    for (int c; (c = getopt_long(argc, argv, "Cw:s:S:p:k:tfh?", longopts, nullptr)) >= 0;) {
        switch (c) {
            case 'C': o.canon = false; break;
            case 'h': case '?': throw Error("Usage: hll <opts> <paths>");
            case 'k': o.k = std::atoi(optarg); break;
            case 'p': o.nthreads = std::atoi(optarg); break;
            case 's': spacing = optarg; break;
            case 'S': o.p = std::atoi(optarg); break;
            case 'w': wsz = std::atoi(optarg); break;
            case L_DEVICE: o.device = parse_device(optarg); break;
            default: break;
        }
    }
    if (!spacing.empty()) unsupported("-s (spacing)");
    if (wsz > o.k) unsupported("-w (window size)");
    if (o.k > 32) throw Error("k must be <= 32 for non-rolling hashes.");
    if (o.wsz > o.k) unsupported("-w/--window-size (windowed minimizers: w > k)");   // option order must not matter: `-w 30 -k 21` is windowed
    if (o.nthreads < 1) o.nthreads = 16;                               // negative -> hardware_concurrency (src/dashing.h:622-625)
    std::vector<std::string> inpaths(argv + optind, argv + argc);      // (-F is parsed nowhere in the reference's getopt string)
    // one "genome" whose files are all the inputs: FNAME_SEP-joined paths are exactly that (src/substrs.h:7-26)
    std::string joined;
    for (auto &pth : inpaths) {
        if (pth.find(' ') != std::string::npos) throw Error("hll: paths with spaces are not supported");
        if (!joined.empty()) joined += ' ';
        joined += pth;
    }
    std::vector<uint8_t> regs(size_t(1) << o.p, 0);
    if (!inpaths.empty())
        sketch_paths(o, std::vector<std::string>{joined}, std::vector<size_t>{0}, [&](size_t, const uint8_t *r) { std::memcpy(regs.data(), r, regs.size()); });
    double est = 0.;
    check(db200_cardinalities(o.device == DB200_ALL_DEVICES ? 0 : o.device, regs.data(), 1, o.p, DB200_ERTL_MLE, &est));
    std::fprintf(stdout, "Estimated number of unique exact matches: %lf\n", (double)(uint64_t)est);
    return 0;
}

// fold_main, src/dashing.cpp:575-595: hll_t(in).compress(destp).write(out)
int fold_main(int argc, char **argv) {
    std::string out = "/dev/stdout", in = "/dev/stdin";
    int destp = -1, device = 0;
    static option longopts[] = {{"device", required_argument, nullptr, L_DEVICE}, {nullptr, 0, nullptr, 0}};
    optind = 1;
    for (int c; (c = getopt_long(argc, argv, "p:o:h?", longopts, nullptr)) >= 0;) {
        switch (c) {
            case 'o': out = optarg; break;
            case 'p': destp = std::atoi(optarg); break;
            case L_DEVICE: device = parse_device(optarg); break;
            default: throw Error("Usage: dashing fold <flags> [in1.hll]\n-o: Write to <path> instead of stdout\n-p: set destination p [must be smaller than the input sketch");
        }
    }
    if (argc - optind == 1) in = argv[optind];
    else if (argc - optind != 0) throw Error("Usage: dashing fold <flags> [in1.hll]");
    const HllFile h = read_hll(in);
    if (out == "-") out = "/dev/stdout";
    if (destp <= 0) destp = (int)h.p - 1;
    if ((uint32_t)destp == h.p) {
        // compress() returns a copy here (hll.h:907), cached cardinality included (read() computed it, hll.h:1078)
        double value = h.value;
        if (value < 0.) { if (h.estim > 2) throw Error("unknown estimation method in " + in); check(db200_cardinalities(device, h.core.data(), 1, (int)h.p, (int)h.estim, &value)); }
        write_hll(out, h.core.data(), h.p, (int)h.estim, (int)h.jestim, value);
        return 0;
    }
    if ((uint32_t)destp > h.p)
        throw Error("Can't compress to a larger size. Current: " + std::to_string(h.p) + ". Requested new size: " + std::to_string(destp));
    std::vector<uint8_t> folded(size_t(1) << destp);
    check(db200_compress(device, h.core.data(), 1, (int)h.p, destp, folded.data()));
    write_hll(out, folded.data(), (uint32_t)destp, (int)h.estim, (int)h.jestim, -1.);   // a fresh hllbase_t(new_np, estim, jestim): not calculated
    return 0;
}

// view_main, src/dashing.cpp:562-566: hll_t::printf (hll.h:888-893) of every file, back to back
int view_main(int argc, char **argv) {
    if (argc < 2) throw Error("Usage: dashing view f1.hll [f2.hll ...]. Only HLLs currently supported.");
    for (int i = 1; i < argc; ++i) {
        const HllFile h = read_hll(argv[i]);
        std::string s = "[";
        for (size_t r = 0; r + 1 < h.core.size(); ++r) { s += std::to_string((int)h.core[r]); s += ", "; }
        s += std::to_string((int)h.core.back());
        s += ']';
        std::fwrite(s.data(), 1, s.size(), stdout);
    }
    std::fflush(stdout);
    return 0;
}

// card_main + size_sketch_and_emit<hll_t>, src/cardmain.cpp, src/sketch_and_cmp.h:122-265: per-path cardinalities, as FLOATS
int card_main(int argc, char **argv) {
    DistOptions o;
    std::string paths_file, out_path;
    bool use_scientific = false;
    if (argc == 1) throw Error("see the reference's `dashing dist` usage: `card` takes the same flags");
    static option longopts[] = {
        {"avoid-sorting", no_argument, nullptr, L_AVOID_SORTING}, {"cache-sketches", no_argument, nullptr, L_CACHE}, {"emit-binary", no_argument, nullptr, L_BINARY},
        {"no-canon", no_argument, nullptr, L_NO_CANON}, {"use-scientific", no_argument, nullptr, L_SCI}, {"presketched", no_argument, nullptr, L_PRESKETCHED},
        {"countmin", no_argument, nullptr, L_COUNTMIN}, {"sketch-by-fname", no_argument, nullptr, L_BY_FNAME}, {"use-bb-minhash", no_argument, nullptr, L_OTHER_SKETCH},
        {"kmer-length", required_argument, nullptr, 'k'}, {"nthreads", required_argument, nullptr, 'p'}, {"out-sizes", required_argument, nullptr, 'o'},
        {"paths", required_argument, nullptr, 'F'}, {"prefix", required_argument, nullptr, 'P'}, {"sketch-size", required_argument, nullptr, 'S'},
        {"suffix", required_argument, nullptr, 'x'}, {"spacing", required_argument, nullptr, 's'}, {"window-size", required_argument, nullptr, 'w'},
        {"original", required_argument, nullptr, 'E'}, {"improved", required_argument, nullptr, 'I'}, {"ertl-joint-mle", required_argument, nullptr, 'J'},
        {"ertl-mle", required_argument, nullptr, 'm'}, {"defer-hll", no_argument, nullptr, L_DEFER_HLL}, {"device", required_argument, nullptr, L_DEVICE},
        {nullptr, 0, nullptr, 0}};
    optind = 1;
    int co;
    while ((co = getopt_long(argc, argv, "n:Q:P:x:F:c:p:o:s:w:O:S:k:=:t:R:D:8TgazlICbMEeHJhZBNyUmqW?", longopts, nullptr)) >= 0) {   // src/cardmain.cpp:27
        switch (co) {
            case L_AVOID_SORTING: o.avoid_sorting = true; break;
            case 'W': case L_CACHE: o.cache_sketches = true; break;
            case 'b': case L_BINARY: o.emit_fmt = BINARY; break;
            case 'C': case L_NO_CANON: o.canon = false; break;
            case 'e': case L_SCI: use_scientific = true; break;
            case L_PRESKETCHED: o.presketched = true; break;
            case 'J': o.jestim = DB200_ERTL_JOINT_MLE; break;
            case 'm': o.jestim = o.estim = DB200_ERTL_MLE; break;
            case 'I': o.jestim = o.estim = DB200_ERTL_IMPROVED; break;
            case 'E': o.jestim = o.estim = DB200_ORIGINAL; break;
            case 'k': o.k = std::atoi(optarg); break;
            case 'p': o.nthreads = std::atoi(optarg); break;
            case 'S': o.p = std::atoi(optarg); break;
            case 'o': out_path = optarg; break;
            case 'F': paths_file = optarg; break;
            case 'P': o.prefix = optarg; break;
            case 'x': o.suffix = optarg; break;
            case L_DEVICE: o.device = parse_device(optarg); break;
            case L_DEFER_HLL: o.defer_hll = true; break;
            case 's': if (*optarg) unsupported("-s/--spacing"); break;
            case 'w': o.wsz = std::atoi(optarg); break;
            case L_COUNTMIN: unsupported("--countmin");
            case L_BY_FNAME: unsupported("--sketch-by-fname");
            case L_OTHER_SKETCH: case '8': unsupported("a non-HLL sketch type");
            case 'g': unsupported("-g/--by-entropy");
            case 'h': case '?': throw Error("see the reference's `dashing dist` usage: `card` takes the same flags");
            default: break;
        }
    }
    if (o.k > 32) throw Error("k must be <= 32 for non-rolling hashes.");
    if (o.wsz > o.k) unsupported("-w/--window-size (windowed minimizers: w > k)");   // option order must not matter: `-w 30 -k 21` is windowed
    if (o.nthreads < 1) o.nthreads = 1;
    if (o.defer_hll) o.estim = o.jestim = DB200_ERTL_MLE;
    std::vector<std::string> inpaths = paths_file.empty() ? std::vector<std::string>(argv + optind, argv + argc) : get_paths(paths_file);
    if (inpaths.empty()) throw Error("No paths. See usage.");
    if (!o.presketched && !o.avoid_sorting) sort_paths_by_fsize(inpaths);
    // card_main hands (prefix, suffix) to a callee that takes (suffix, prefix) (src/cardmain.cpp:4-6 vs src/sketch_and_cmp.h:126)
    std::swap(o.prefix, o.suffix);
    const size_t n = inpaths.size(), m = size_t(1) << o.p;
    std::vector<uint8_t> regs(n * m);
    std::vector<size_t> todo;
    std::vector<std::string> fnames(n);
    std::vector<Cached> info(n);
    load_existing(o, inpaths, regs, info, fnames, todo);
    sketch_paths(o, inpaths, todo, [&](size_t i, const uint8_t *r) {
        std::memcpy(&regs[i * m], r, m);
        if (o.cache_sketches && !o.defer_hll) write_hll(fnames[i], r, o.p, o.estim, o.jestim, -1.);
    });
    std::vector<double> card(n);
    check(db200_cardinalities(o.device, regs.data(), n, o.p, o.estim, card.data()));
    apply_cached(o.device, o.p, o.estim, regs, info, card);
    if (o.defer_hll && o.cache_sketches) for (size_t i : todo) write_hll(fnames[i], &regs[i * m], o.p, 2, 2, card[i]);
    std::FILE *fp = out_path.empty() ? stdout : std::fopen(out_path.c_str(), "w");
    if (!fp) throw Error("Could not open file at " + out_path + " for writing.");
    std::vector<float> fbuf(n);
    for (size_t i = 0; i < n; ++i) fbuf[i] = (float)card[i];           // :231-236
    if (o.emit_fmt == BINARY) {
        if (std::fwrite(fbuf.data(), sizeof(float), n, fp) != n) throw Error("Failed to write cardinality estimates to file");
    } else {
        std::string s("#Path\tSize (est.)\n");
        char buf[64];
        for (size_t i = 0; i < n; ++i) {
            s += inpaths[i];
            const int l = std::snprintf(buf, sizeof buf, use_scientific ? "\t%0.12g\n" : "\t%0.8f\n", (double)fbuf[i]);   // :247-249
            s.append(buf, l);
        }
        std::fwrite(s.data(), 1, s.size(), fp);
    }
    if (fp != stdout) std::fclose(fp); else std::fflush(fp);
    return 0;
}

// sketch_by_seq_main + sketch_by_seq_core, src/dashing.cpp:470-556, src/sketch_and_cmp.h:540-602: one sketch per RECORD of one
// sequence file, all into one gzip stream, record names into <out>.names
int sketch_by_seq_main(int argc, char **argv) {
    SketchOptions o;
    std::string outpath = "/dev/stdout";
    bool defer_hll = false;
    optind = 1;
    int co;
    while ((co = getopt_long(argc, argv, "o:n:P:p:x:R:s:S:k:w:H:q:B:8JbfjEIcCeh?", sketch_longopts, nullptr)) >= 0) {   // src/dashing.cpp:487
        if (sketch_option(co, o, defer_hll)) continue;
        if (co == 'o') outpath = optarg;
    }
    if (o.k > 32) throw Error("k must be <= 32 for non-rolling hashes.");
    if (o.wsz > o.k) unsupported("-w/--window-size (windowed minimizers: w > k)");   // option order must not matter: `-w 30 -k 21` is windowed
    if (argc != optind + 1) throw Error("Usage: sketch_by_seq <opts> [same as sketch] -o out_path sequence_file");
    // src/dashing.cpp:541-544 tests the flag the wrong way round: WITHOUT --defer-hll it instantiates HyperLogLogHasher, whose
    // inherited write(gzFile) emits b-bit minhash records (bbmh.h:962-969); only --defer-hll yields hll_t records
    if (!defer_hll)
        unsupported("sketch_by_seq without --defer-hll (the reference then writes b-bit minhash records, not HLLs; pass --defer-hll for HLL records)");
    const std::string inpath = argv[optind];
    // kseq_read: names + sequences
    std::vector<std::string> names;
    std::string bases;
    std::vector<uint64_t> offs{0};
    for_each_named_record(inpath, [&](const std::string &name, const char *sq, size_t l) {
        names.push_back(name);
        bases.append(sq, l);
        offs.push_back(bases.size());
    });
    const std::string namepath = outpath == "/dev/stdout" ? std::string("stdout.names") : outpath + ".names";
    std::FILE *nfp = std::fopen(namepath.c_str(), "w");
    if (!nfp) throw Error("Failed to open file for writing at " + namepath);
    std::fprintf(nfp, "#k=%d:Names for sequences sketched\n", o.k);
    for (auto &nm : names) { std::fwrite(nm.data(), 1, nm.size(), nfp); std::fputc('\n', nfp); }
    std::fclose(nfp);
    const size_t n = names.size(), m = size_t(1) << o.p;
    std::vector<uint8_t> regs(n * m);
    if (n) {
        std::vector<uint64_t> grb(n + 1);
        for (size_t i = 0; i <= n; ++i) grb[i] = i;                   // every record is its own "genome"
        if (bases.empty()) bases.push_back('N');
        check(db200_sketch_batch(o.device, o.p, o.k, o.canon, bases.data(), offs.data(), n, grb.data(), n, regs.data()));
    }
    gzFile ofp = gzopen(outpath.c_str(), "wb");
    if (!ofp) throw Error("Failed to open file for writing at " + outpath);
    for (size_t i = 0; i < n; ++i) write_hll_stream(ofp, &regs[i * m], o.p, o.estim, o.jestim, -1.);
    gzclose(ofp);
    return 0;
}

// dist_by_seq_main + dist_by_seq<hll_t>, src/distbyseq.cpp:50-138, src/sketch_and_cmp.h:76-118: all pairs over the records of a
// sketch_by_seq container, labelled by its .names file
int dist_by_seq_main(int argc, char **argv) {
    DistOptions o;
    o.k = -1;
    std::string namefile, otherpath, outpath = "/dev/stdout";
    static option longopts[] = {
        {"containment-index", no_argument, nullptr, L_CONT_INDEX}, {"containment-dist", no_argument, nullptr, L_CONT_DIST}, {"mash-dist", no_argument, nullptr, L_MASH},
        {"symmetric-containment-index", no_argument, nullptr, L_SYM_CONT_INDEX}, {"symmetric-containment-dist", no_argument, nullptr, L_SYM_CONT_DIST},
        {"sizes", no_argument, nullptr, L_SIZES}, {"device", required_argument, nullptr, L_DEVICE}, {nullptr, 0, nullptr, 0}};
    optind = 1;
    int c;
    while ((c = getopt_long(argc, argv, "q:o:k:n:p:EIJMBbS8KCTrh?", longopts, nullptr)) >= 0) {   // src/distbyseq.cpp:71
        switch (c) {
            case 'B': case 'S': case '8': case 'K': case 'r': unsupported("a non-HLL sketch type");
            case 'p': o.nthreads = std::atoi(optarg); break;
            case 'o': outpath = optarg; break;
            case 'E': o.jestim = o.estim = DB200_ORIGINAL; break;
            case 'I': o.jestim = o.estim = DB200_ERTL_IMPROVED; break;
            case 'J': o.jestim = DB200_ERTL_JOINT_MLE; break;
            case 'k': o.k = std::atoi(optarg); break;
            case 'n': namefile = optarg; break;
            case 'q': otherpath = optarg; break;
            case 'b': o.emit_fmt = BINARY; break;
            case 'T': o.emit_fmt = FULL_TSV; break;
            case L_MASH: o.result_type = DB200_MASH_DIST; break;   // --mash-dist is a long FLAG entry; the short -M is accepted but has no case in the reference's switch
            case L_CONT_INDEX: o.result_type = DB200_CONTAINMENT_INDEX; break;
            case L_CONT_DIST: o.result_type = DB200_CONTAINMENT_DIST; break;
            case L_SYM_CONT_INDEX: o.result_type = DB200_SYMMETRIC_CONTAINMENT_INDEX; break;
            case L_SYM_CONT_DIST: o.result_type = DB200_SYMMETRIC_CONTAINMENT_DIST; break;
            case L_SIZES: o.result_type = DB200_SIZES; break;
            case L_DEVICE: o.device = parse_device(optarg); break;
            case 'h': case '?': throw Error("Usage: dist_by_seq <flags> -n [namefile] input_file");
            default: break;                                              // -C: listed, no case
        }
    }
    if (optind + 1 != argc || namefile.empty()) throw Error("Usage: dist_by_seq <flags> -n [namefile] input_file");
    // -q reads the query sketches from the SAME stream, past its end (src/sketch_and_cmp.h:89-98): nothing to reproduce
    if (!otherpath.empty()) unsupported("-q (the reference reads the extra sketches past the end of the input stream)");
    const std::vector<std::string> labels = get_paths(namefile);
    if (o.k <= 0) {                                                     // "#k=<k>:Names ..." (src/distbyseq.cpp:104-111)
        std::ifstream reader(namefile);
        std::string line;
        std::getline(reader, line);
        int tmp = line.size() > 3 ? std::atoi(line.c_str() + 3) : 0;
        o.k = tmp > 0 ? tmp : 31;
    }
    const int rt = o.result_type;
    if (rt == DB200_CONTAINMENT_INDEX || rt == DB200_CONTAINMENT_DIST || rt == DB200_FULL_CONTAINMENT_DIST)
        throw Error("Can't perform asymmetric comparison without query paths");   // src/sketch_and_cmp.h:100-102
    const std::vector<HllFile> hs = read_hll_container(argv[optind], labels.size());
    if (hs.empty()) throw Error("no sketches to compare");
    o.p = (int)hs[0].p;
    const size_t m = size_t(1) << o.p;
    std::vector<uint8_t> regs(hs.size() * m);
    std::vector<Cached> info(hs.size());
    for (size_t i = 0; i < hs.size(); ++i) {
        if (hs[i].p != hs[0].p) throw Error("mismatched sketch sizes.");
        info[i] = cached_of(hs[i]);
        std::memcpy(&regs[i * m], hs[i].core.data(), m);
    }
    // every sketch came from a file: its cardinality is what read() cached under the FILE's estimator (hll.h:1078)
    std::vector<double> card(hs.size());
    check(db200_cardinalities(o.device, regs.data(), hs.size(), o.p, o.estim, card.data()));
    apply_cached(o.device, o.p, o.estim, regs, info, card);
    o.dist_path = outpath;
    compare_and_emit(o, labels, regs, 0, card.data());
    return 0;
}

int cli_main(int argc, char **argv) {
    phase("start");
    try {
        if (argc < 2) throw Error("usage: dashing_b200 <sketch|dist|cmp|union|hll|fold|view|card|sketch_by_seq|dist_by_seq> [options] (dashing's own flags)");
        const std::string sub = argv[1];
        // src/main.cpp:24-40
        if (sub == "sketch") return sketch_main(argc - 1, argv + 1);
        if (sub == "dist" || sub == "cmp" || sub == "setdist") return dist_main(argc - 1, argv + 1);
        if (sub == "union") return union_main(argc - 1, argv + 1);
        if (sub == "hll") return hll_main(argc - 1, argv + 1);
        if (sub == "fold") return fold_main(argc - 1, argv + 1);
        if (sub == "view") return view_main(argc - 1, argv + 1);
        if (sub == "card") return card_main(argc - 1, argv + 1);
        if (sub == "sketch_by_seq" || sub == "sbs") return sketch_by_seq_main(argc - 1, argv + 1);
        if (sub == "dist_by_seq" || sub == "cmp_by_seq") return dist_by_seq_main(argc - 1, argv + 1);
        throw Error("subcommand '" + sub + "' is outside the B200 engine's scope (sketch, dist/cmp, union, hll, fold, view, card, "
                    "sketch_by_seq, dist_by_seq; not panel, printmat, mkdist, flatten)");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());                   // UNRECOVERABLE_ERROR: message + exit(1)
        return 1;
    }
}

} // namespace db200h

// ---- C hooks for ctypes-based tests (no compute: formats and naming only, plus the CLI) ---------------------------
extern "C" {
#define DB200H_API __attribute__((visibility("default")))
DB200H_API int db200h_cli(int argc, char **argv) { return db200h::cli_main(argc, argv); }
DB200H_API int db200h_make_fname(const char *path, int p, int wsz, int k, int csz, const char *spacing, const char *suffix, const char *prefix, char *out, uint64_t cap) {
    const std::string s = db200h::make_fname(path, p, wsz, k, csz, spacing, suffix, prefix);
    if (s.size() + 1 > cap) return 1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return 0;
}
DB200H_API int db200h_write_hll(const char *path, const uint8_t *regs, int p, int estim, int jestim, double value) {
    try { db200h::write_hll(path, regs, p, estim, jestim, value); } catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
DB200H_API int db200h_read_hll(const char *path, uint8_t *regs, uint64_t cap, uint32_t *hdr5, double *value) {
    try {
        const auto h = db200h::read_hll(path);
        if (h.core.size() > cap) return 2;
        std::memcpy(regs, h.core.data(), h.core.size());
        hdr5[0] = h.is_calculated; hdr5[1] = h.estim; hdr5[2] = h.jestim; hdr5[3] = h.marker; hdr5[4] = h.p;
        *value = h.value;
    } catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
// formats a symmetric result (packed upper triangle) into `out`; names are '\n'-separated.  Returns bytes needed.
DB200H_API uint64_t db200h_format_symmetric(const char *names_nl, uint64_t n, const float *packed, const float *packed_lower, int fmt, char *out, uint64_t cap) {
    std::vector<std::string> names;
    const char *p = names_nl;
    for (uint64_t i = 0; i < n; ++i) { const char *e = std::strchr(p, '\n'); names.emplace_back(p, e ? e - p : std::strlen(p)); p = e ? e + 1 : p + std::strlen(p); }
    const std::string s = db200h::format_symmetric(names, packed, (db200h::EmissionFormat)fmt, packed_lower);
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return s.size();
}
// nearest-neighbour table: TSV into `out` (fmt 0) or the binary form (fmt 1: 8-byte header + pairs).  Returns bytes needed.
DB200H_API uint64_t db200h_format_neighbors(const char *names_nl, uint64_t n, uint64_t qoffset, const void *nb, uint64_t rows, uint32_t nn, int fmt,
                                            char *out, uint64_t cap) {
    std::vector<std::string> names;
    const char *p = names_nl;
    for (uint64_t i = 0; i < n; ++i) { const char *e = std::strchr(p, '\n'); names.emplace_back(p, e ? e - p : std::strlen(p)); p = e ? e + 1 : p + std::strlen(p); }
    std::string s;
    if (fmt == 0) s = db200h::format_neighbors(names, qoffset, static_cast<const db200h::Neighbor *>(nb), rows, nn);
    else {
        const uint32_t hdr[2] = {(uint32_t)n, nn};
        s.assign(reinterpret_cast<const char *>(hdr), 8);
        s.append(static_cast<const char *>(nb), rows * nn * sizeof(db200h::Neighbor));
    }
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return s.size();
}
DB200H_API uint64_t db200h_format_sizes(const char *names_nl, uint64_t n, const double *card, char *out, uint64_t cap) {
    std::vector<std::string> names;
    const char *p = names_nl;
    for (uint64_t i = 0; i < n; ++i) { const char *e = std::strchr(p, '\n'); names.emplace_back(p, e ? e - p : std::strlen(p)); p = e ? e + 1 : p + std::strlen(p); }
    const std::string s = db200h::format_sizes(names, card);
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return s.size();
}
DB200H_API uint64_t db200h_format_rect_row(const char *qname, const float *row, uint64_t nr, char *out, uint64_t cap) {
    const std::string s = db200h::format_rect_row(qname, row, nr);
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return s.size();
}
// parses a FASTA/FASTQ(.gz) file; writes concatenated records and their offsets.  Returns the number of records.
DB200H_API int64_t db200h_read_records(const char *path, char *bases, uint64_t cap, uint64_t *offs, uint64_t maxrec) {
    try {
        // the window form the batch driver uses (records land in the caller's buffer; -2 = window too small)
        std::vector<uint64_t> ends;
        if (db200h::parse_into_window(path, bases, cap, ends) == SIZE_MAX || ends.size() > maxrec) return -2;
        offs[0] = 0;
        for (size_t i = 0; i < ends.size(); ++i) offs[i + 1] = ends[i];
        return (int64_t)ends.size();
    } catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return -1; }
}
// record names of a FASTA/FASTQ(.gz) file (what sketch_by_seq writes to <out>.names), newline-terminated.  Returns bytes needed.
DB200H_API int64_t db200h_record_names(const char *path, char *out, uint64_t cap) {
    try {
        std::string s;
        db200h::for_each_named_record(path, [&](const std::string &nm, const char *, size_t) { s += nm; s += '\n'; });
        if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
        return (int64_t)s.size();
    } catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return -1; }
}
// get_paths(): newline-terminated entries.  Returns bytes needed.
DB200H_API int64_t db200h_get_paths(const char *file, char *out, uint64_t cap) {
    std::string s;
    for (auto &p : db200h::get_paths(file)) { s += p; s += '\n'; }
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return (int64_t)s.size();
}
// the raw bytes of a file (inflated if gzip) into the caller's window, as the batch driver reads them; -2 = window too small
DB200H_API int64_t db200h_slurp(const char *path, char *dst, uint64_t cap) {
    try {
        const size_t n = db200h::slurp_file(path, dst, cap);
        return n == SIZE_MAX ? -2 : (int64_t)n;
    } catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return -1; }
}
// file_capacity(): the window the batch driver reserves for a file
DB200H_API uint64_t db200h_file_capacity(const char *path) {
    try { return db200h::file_window(path); } catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return UINT64_MAX; }
}
}
