// dashing_host.hpp — host side above the C ABI, mirroring the reference's interface for the two hot paths:
// file formats (.hll, sizes file, distance-matrix outputs), file naming, FASTA feeding and the `sketch` / `dist`
// drivers.  All arithmetic of the hot paths happens in libdashing_b200.so (CUDA); this layer only parses, moves
// bytes and formats.  Citations are relative to the reference tree.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace db200h {

// bns::EmissionFormat — src/enums.h:25-34
enum EmissionFormat : unsigned { UT_TSV = 0, BINARY = 1, UPPER_TRIANGULAR = 2, FULL_TSV = 3 };

// UNRECOVERABLE_ERROR (bonsai/include/bonsai/util.h:547-554) prints and exit(1)s; here it is an exception the
// CLI turns into exactly that.
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- .hll container — hll_t::write / read, bonsai/hll/include/sketch/hll.h:1039-1080 --------------------------
struct HllFile {
    uint32_t is_calculated = 0, estim = 2, jestim = 2, marker = 1, p = 0;
    double value = -1.;
    std::vector<uint8_t> core;
};
std::vector<uint8_t> hll_payload(const uint8_t *regs, uint32_t p, int estim, int jestim, double value);
void write_hll(const std::string &path, const uint8_t *regs, uint32_t p, int estim = 2, int jestim = 2, double value = -1.);
HllFile read_hll(const std::string &path);
// several sketches in one gzip stream: what `sketch -o` and `sketch_by_seq` write and `dist_by_seq` reads
std::vector<HllFile> read_hll_container(const std::string &path, size_t count);
// make_fname<hll_t>, src/dashing.h:497-526
std::string make_fname(const char *path, size_t sketch_p, int wsz, int k, int csz, const std::string &spacing,
                       const std::string &suffix = "", const std::string &prefix = "");
// for_each_substr over FNAME_SEP (src/substrs.h:7-26): one "path" may name several files folded into one sketch
std::vector<std::string> split_paths(const std::string &s, char sep = ' ');
// get_paths (bonsai/include/bonsai/util.h:1185): one path per line
std::vector<std::string> get_paths(const std::string &file);
// sort_paths_by_fsize (src/finalizers.cpp:6-21): descending total file size (uint32_t sizes as in the reference)
void sort_paths_by_fsize(std::vector<std::string> &paths);

// ---- FASTA / FASTQ records with kseq semantics (bonsai/klib/kseq.h:177-218), gz transparent --------------------
void for_each_record(const std::string &file, const std::function<void(const char *, size_t)> &fn);
// ... with the record name (ks->name.s: the header up to the first whitespace), for sketch_by_seq
void for_each_named_record(const std::string &file, const std::function<void(const std::string &, const char *, size_t)> &fn);

// ---- emitters — src/sketch_and_cmp.h:16-35, :372-397, :838-878; src/dashing.h:675-705 -------------------------
std::string format_sizes(const std::vector<std::string> &paths, const double *card);
std::string format_ut_tsv_header(const std::vector<std::string> &paths);
// one row of the upper-triangular TSV / PHYLIP output: `row` holds the n-1-index values (i, j>i)
void append_ut_row(std::string &buf, const std::string &name, const float *row, size_t n, size_t index, EmissionFormat fmt);
// rows of the packed upper triangle -> full text output
std::string format_symmetric(const std::vector<std::string> &paths, const float *packed, EmissionFormat fmt,
                             const float *packed_lower = nullptr /* FULL_TSV only: values for i > j */);
std::string format_rect_row(const std::string &qname, const float *row, size_t nr);
void write_binary_matrix(std::FILE *fp, const float *packed, uint64_t n);   // '\0' + u64 n + floats (distmat.h:188-208)

// ---- drivers ----------------------------------------------------------------------------------------------------
struct SketchOptions {
    int k = 31, p = 10, nthreads = 1, device = 0;
    int wsz = 0;                            // -w: checked against k AFTER all options are parsed (the reference builds Spacer(k, wsz) then)
    int estim = 2, jestim = 2;              // -E/-I/-J: stored in the header of every .hll written (set_estim_and_jestim)
    bool canon = true, skip_cached = false, avoid_sorting = false;
    std::string prefix, suffix;
    std::string output_file;                // sketch -o: all sketches into one gzip stream + <file>.labels.gz (src/sketch_and_cmp.h:466-536)
    size_t batch_bytes = size_t(1) << 30;   // ASCII handed to one db200_sketch_batch call
};
struct DistOptions : SketchOptions {
    int result_type = 1 /* JI */;
    EmissionFormat emit_fmt = UT_TSV;
    bool presketched = false, cache_sketches = false;
    // --defer-hll (dist_sketch_and_cmp<HyperLogLogHasher<>>, src/distmain.cpp:177): the per-bucket minima finalize to the
    // same registers (bbmh.h:1175-1184), but the final hll_t objects keep the constructor's ERTL_MLE / ERTL_MLE whatever
    // -E/-I/-J said, and cached sketches are written with their cardinality already computed
    bool defer_hll = false;
    unsigned nneighbors = 0;                 // --nearest-neighbors (gargs.number_neighbors, src/dashing.h:255); 0 = all pairs
    std::string sizes_path, dist_path;       // empty -> stdout
};
// sketch_core<hll_t>, src/sketch_and_cmp.h:445-538: one .hll per input path
// Parses a file's records straight into a caller window (what the batch driver does per file): returns the bytes used
// (record ends relative to the window in `ends`) or SIZE_MAX when the window is too small.  file_window(): the window
// the driver reserves — the file size, or the ISIZE trailer for gzip (RFC 1952; too small for multi-member files).
size_t parse_into_window(const std::string &file, char *dst, size_t cap, std::vector<uint64_t> &ends);
size_t file_window(const std::string &file);
// The raw bytes of a file (inflated if it is gzip) into a caller window — all the CLI does with sequence files before handing
// them to db200_sketch_fasta_batch; SIZE_MAX when the window is too small (a multi-member gzip: its trailer sizes the last member).
size_t slurp_file(const std::string &file, char *dst, size_t cap);
void sketch_core(const SketchOptions &o, std::vector<std::string> paths);
// dist_sketch_and_cmp<hll_t> + dist_loop / partdist_loop, src/sketch_and_cmp.h:268-417, :785-880; src/dashing.h:660-712.
// The last nq entries of inpaths are queries (rectangular mode).
void dist_sketch_and_cmp(const DistOptions &o, std::vector<std::string> inpaths, size_t nq);
// nndist_loop's two output forms (src/sketch_and_cmp.h:733-782): rows of `nn` (value, index) pairs, row i named names[i + qoffset]
struct Neighbor { float value; uint32_t index; };
std::string format_neighbors(const std::vector<std::string> &names, size_t qoffset, const Neighbor *nb, size_t rows, unsigned nn);
void write_binary_neighbors(std::FILE *fp, uint32_t npaths, const Neighbor *nb, size_t rows, unsigned nn);
// dist_loop / partdist_loop / nndist_loop over in-memory register rows (last nq rows are queries)
// cached_card: per-sketch cardinalities that sketches loaded from files carry (hll_t::value_), or nullptr
void compare_and_emit(const DistOptions &o, const std::vector<std::string> &names, const std::vector<uint8_t> &regs, size_t nq,
                      const double *cached_card = nullptr);
// sketch_main / dist_main (src/dashing.cpp:294-409, src/distmain.cpp:28-204): the hot subset of the flags
int sketch_main(int argc, char **argv);
int dist_main(int argc, char **argv);
// SURVEY.md §8(f)3 — the subcommands that reuse the same primitives:
int union_main(int argc, char **argv);          // src/union.cpp:60-108 (HLL sketches)
int hll_main(int argc, char **argv);            // src/hllmain.cpp:4-40
int fold_main(int argc, char **argv);           // src/dashing.cpp:575-595
int view_main(int argc, char **argv);           // src/dashing.cpp:562-566
int card_main(int argc, char **argv);           // src/cardmain.cpp (size_sketch_and_emit<hll_t>, src/sketch_and_cmp.h:122-265)
int sketch_by_seq_main(int argc, char **argv);  // src/dashing.cpp:470-556 (sketch_by_seq_core, src/sketch_and_cmp.h:540-602)
int dist_by_seq_main(int argc, char **argv);    // src/distbyseq.cpp:50-138 (dist_by_seq, src/sketch_and_cmp.h:76-118)
int cli_main(int argc, char **argv);

} // namespace db200h
