// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" driver around the UNMODIFIED reference headers under
// /root/reference (dnbaker/dashing @ 0635bea).  It contains no arithmetic of
// its own: every number it returns is produced by the reference's own
// templates, instantiated here exactly as the reference's drivers do:
//
//   k-mer stream      bns::Encoder<score::Lex>::for_each(func, str, len)
//                       bonsai/include/bonsai/encoder.h:415-441 -> :218-232 -> :240-271
//   register update   sketch::hll_t::addh           bonsai/hll/include/sketch/hll.h:843-846, :828-836
//   cardinality       sketch::hll_t::report         hll.h:773-803 (sum_counts :515-532, calculate_estimate :199-246)
//   MLE               hll::detail::ertl_ml_estimate hll.h:567-627
//   pair value        bns::result_cmp               src/dashing.h:568-592 (jaccard_index hll.h:1174-1183,
//                                                   full_set_comparison :1165-1173, ertl_joint :636-684)
//   all-pairs loops   mirrors perform_core_op (src/sketch_and_cmp.h:699-710), dist_loop BINARY branch
//                     (:838-850 operand order cmp(s[i], s[j])) and partdist_loop (src/dashing.h:675-681)
//   .hll files        sketch::hll_t::write/read(path) hll.h:1039-1087
//   whole drivers     bns::sketch_core<hll_t>          src/sketch_and_cmp.h:445-538   (FASTA in, .hll files out)
//                     bns::dist_sketch_and_cmp<hll_t>  src/sketch_and_cmp.h:268-417   (sizes file + every output format,
//                                                       through the reference's own dist_loop / partdist_loop / emitters)
//
// Built by oracle/Makefile with g++ directly on this one file (the reference's
// own build system is never run); the output goes to oracle/_ref/ only.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference
// legs may load the resulting library.
#include "sketch_and_cmp.h"   // pulls in dashing.h; gives access to the reference's own drivers
#include "union.cpp"          // the reference's union_core<T> / union_main, compiled where it lies (src/union.cpp)
#include "hllmain.cpp"        // the reference's hll_main (src/hllmain.cpp)
#include <omp.h>
#include <cstring>
#include <vector>
#include <memory>

// process-global option block of the reference (declared extern in src/dashing.h:264, defined in src/dashing.cpp:8,
// which is not linked here)
namespace bns { GlobalArgs gargs; }
// union_main's usage text lives in src/dashing.cpp (not linked here); it only runs on bad flags
namespace bns { void union_usage(char *) { std::exit(1); } }

using namespace bns;
using namespace sketch;

namespace {
using hll_t = hll::hll_t;

static hll_t make_sketch(const uint8_t *regs, int p, int estim, int jestim) {
    hll_t h(p, (hll::EstimationMethod)estim, (hll::JointEstimationMethod)jestim);
    std::memcpy(h.mutable_core().data(), regs, h.size());
    h.not_ready();
    return h;
}

static std::vector<hll_t> make_sketches(const uint8_t *regs, uint64_t n, int p, int estim, int jestim) {
    // Mirrors dist_sketch_and_cmp: construct, set_estim_and_jestim (src/sketch_and_cmp.h:285-288),
    // then the serial sizes loop calls cardinality_estimate -> report() on every sketch (:377-383),
    // so every sketch enters the pair loop with its cached value_ "ready".
    std::vector<hll_t> v;
    v.reserve(n);
    const size_t m = size_t(1) << p;
    for(uint64_t i = 0; i < n; ++i) {
        v.emplace_back(make_sketch(regs + i * m, p, estim, jestim));
        v.back().report();
    }
    return v;
}
} // namespace

extern "C" {

int dref_simd_tier() {
#if HAS_AVX_512 && __AVX512BW__
    return 512;
#elif __AVX2__
    return 256;
#else
    return 128;
#endif
}

int dref_max_threads() { return omp_get_max_threads(); }

uint64_t dref_wang(uint64_t x) { return hash::WangHash()(x); }

// Emits the k-mers of ONE record exactly as Encoder::for_each(func, str, l) does.
uint64_t dref_kmers(const char *s, uint64_t len, int k, int canon, uint64_t *out, uint64_t cap) {
    Encoder<score::Lex> enc(Spacer(k, k), canon != 0);
    uint64_t n = 0;
    enc.for_each([&](u64 km) { if(n < cap) out[n] = km; ++n; }, s, len);
    return n;
}

// One genome = records [0, nrec) delimited by offsets[0..nrec]; every record is fed separately
// (k-mers never span records: encoder.h:444, :201-205).
int dref_sketch(const char *bases, const uint64_t *offsets, uint64_t nrec, int k, int p, int canon, uint8_t *regs_out) {
    hll_t h(p);
    Encoder<score::Lex> enc(Spacer(k, k), canon != 0);
    for(uint64_t r = 0; r < nrec; ++r)
        enc.for_each([&](u64 km) { h.addh(km); }, bases + offsets[r], offsets[r + 1] - offsets[r]);
    std::memcpy(regs_out, h.data(), h.size());
    return 0;
}

// Many genomes, one OpenMP task per genome like sketch_core (src/sketch_and_cmp.h:484-523):
// genome g owns records [genome_rec_begin[g], genome_rec_begin[g+1]).
int dref_sketch_many(const char *bases, const uint64_t *offsets, const uint64_t *genome_rec_begin, uint64_t ngenomes,
                     int k, int p, int canon, int nthreads, uint8_t *regs_out) {
    if(nthreads <= 0) nthreads = omp_get_max_threads();
    const size_t m = size_t(1) << p;
    #pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for(uint64_t g = 0; g < ngenomes; ++g) {
        hll_t h(p);
        Encoder<score::Lex> enc(Spacer(k, k), canon != 0);
        for(uint64_t r = genome_rec_begin[g]; r < genome_rec_begin[g + 1]; ++r)
            enc.for_each([&](u64 km) { h.addh(km); }, bases + offsets[r], offsets[r + 1] - offsets[r]);
        std::memcpy(regs_out + g * m, h.data(), m);
    }
    return 0;
}

void dref_histogram(const uint8_t *regs, int p, uint32_t *out64) {
    hll_t h = make_sketch(regs, p, hll::ERTL_MLE, hll::ERTL_MLE);
    auto c = hll::detail::sum_counts(h.core());
    std::memcpy(out64, c.data(), sizeof(uint32_t) * 64);
}

double dref_mle(const uint32_t *c64, int p, int q) {
    std::array<uint32_t, 64> c;
    std::memcpy(c.data(), c64, sizeof(uint32_t) * 64);
    return hll::detail::ertl_ml_estimate(c, (unsigned)p, (unsigned)q);
}

double dref_estimate_from_counts(const uint32_t *c64, int p, int estim) {
    std::array<uint32_t, 64> c;
    std::memcpy(c.data(), c64, sizeof(uint32_t) * 64);
    const uint64_t m = uint64_t(1) << p;
    return hll::detail::calculate_estimate(c, (hll::EstimationMethod)estim, m, p, hll::make_alpha(m));
}

double dref_cardinality(const uint8_t *regs, int p, int estim) {
    hll_t h = make_sketch(regs, p, estim, hll::ERTL_MLE);
    return h.report();
}

void dref_cardinalities(const uint8_t *regs, uint64_t n, int p, int estim, double *out) {
    const size_t m = size_t(1) << p;
    for(uint64_t i = 0; i < n; ++i) out[i] = dref_cardinality(regs + i * m, p, estim);
}

// One pair through result_cmp, both sketches "ready" as in the dist flow.
float dref_pair(const uint8_t *lhs, const uint8_t *rhs, int p, int estim, int jestim, int rtype, int k) {
    hll_t A = make_sketch(lhs, p, estim, jestim), B = make_sketch(rhs, p, estim, jestim);
    A.report(); B.report();
    const float ksinv = 1. / k; // src/sketch_and_cmp.h:797
    return result_cmp(A, B, (EmissionType)rtype, ksinv);
}

double dref_jaccard(const uint8_t *lhs, const uint8_t *rhs, int p, int estim, int jestim) {
    hll_t A = make_sketch(lhs, p, estim, jestim), B = make_sketch(rhs, p, estim, jestim);
    A.report(); B.report();
    return static_cast<const hll_t &>(A).jaccard_index(static_cast<const hll_t &>(B));
}

double dref_union_size(const uint8_t *lhs, const uint8_t *rhs, int p, int estim, int jestim) {
    hll_t A = make_sketch(lhs, p, estim, jestim), B = make_sketch(rhs, p, estim, jestim);
    A.report(); B.report();
    return A.union_size(B);
}

void dref_triple(const uint8_t *lhs, const uint8_t *rhs, int p, int estim, int jestim, double *out3) {
    hll_t A = make_sketch(lhs, p, estim, jestim), B = make_sketch(rhs, p, estim, jestim);
    A.report(); B.report();
    auto t = A.full_set_comparison(B);
    out3[0] = t[0]; out3[1] = t[1]; out3[2] = t[2];
}

// Rows [row_begin, row_end) of the symmetric all-pairs matrix, written at their distmat offsets
// (distmat/distmat.h:260-276) into `out` (length n(n-1)/2).
//   order 0: value = cmp(sketches[i], sketches[j])   (BINARY paths, src/sketch_and_cmp.h:829, :849)
//   order 1: value = cmp(sketches[j], sketches[i])   (TSV / PHYLIP path via perform_core_op, :699-710)
int dref_dist_rows(const uint8_t *regs, uint64_t n, int p, int k, int estim, int jestim, int rtype, int order,
                   uint64_t row_begin, uint64_t row_end, int nthreads, float *out) {
    if(nthreads <= 0) nthreads = omp_get_max_threads();
    std::vector<hll_t> sk = make_sketches(regs, n, p, estim, jestim);
    const float ksinv = 1. / k;
    const EmissionType rt = (EmissionType)rtype;
    for(uint64_t i = row_begin; i < row_end && i + 1 < n; ++i) {
        float *dists = out + (i * (2 * n - i - 1)) / 2; // row_ptr(i)
        const hll_t &h1 = sk[i];
        #pragma omp parallel for schedule(dynamic) num_threads(nthreads)
        for(uint64_t j = i + 1; j < n; ++j)
            dists[j - i - 1] = order ? result_cmp(sk[j], h1, rt, ksinv) : result_cmp(h1, sk[j], rt, ksinv);
    }
    return 0;
}

// Prepared form of the same loop: the sketches are built (and report()ed) once, so a bounded sample of rows can
// be timed without the per-call construction of n hll_t objects.
struct dref_set { std::vector<hll_t> sk; uint64_t n; };

void *dref_set_create(const uint8_t *regs, uint64_t n, int p, int estim, int jestim) {
    dref_set *s = new dref_set;
    s->sk = make_sketches(regs, n, p, estim, jestim);
    s->n = n;
    return s;
}

void dref_set_free(void *h) { delete static_cast<dref_set *>(h); }

int dref_set_dist_rows(void *h, int k, int rtype, int order, uint64_t row_begin, uint64_t row_end, int nthreads, float *out) {
    dref_set *s = static_cast<dref_set *>(h);
    if(nthreads <= 0) nthreads = omp_get_max_threads();
    const uint64_t n = s->n;
    const float ksinv = 1. / k;
    const EmissionType rt = (EmissionType)rtype;
    const float *base = out;
    for(uint64_t i = row_begin; i < row_end && i + 1 < n; ++i) {
        // rows are written contiguously from out[0] (the caller sizes `out` for the sampled rows only)
        float *dists = const_cast<float *>(base) + ((i * (2 * n - i - 1)) / 2 - (row_begin * (2 * n - row_begin - 1)) / 2);
        const hll_t &h1 = s->sk[i];
        #pragma omp parallel for schedule(dynamic) num_threads(nthreads)
        for(uint64_t j = i + 1; j < n; ++j)
            dists[j - i - 1] = order ? result_cmp(s->sk[j], h1, rt, ksinv) : result_cmp(h1, s->sk[j], rt, ksinv);
    }
    return 0;
}

int dref_dist_symmetric(const uint8_t *regs, uint64_t n, int p, int k, int estim, int jestim, int rtype, int order,
                        int nthreads, float *out) {
    return dref_dist_rows(regs, n, p, k, estim, jestim, rtype, order, 0, n, nthreads, out);
}

// partdist_loop (src/dashing.h:675-681): out[q * nr + j] = result_cmp(refs[j], queries[q]).
int dref_dist_rect(const uint8_t *ref_regs, uint64_t nr, const uint8_t *qry_regs, uint64_t nq, int p, int k,
                   int estim, int jestim, int rtype, int nthreads, float *out) {
    if(nthreads <= 0) nthreads = omp_get_max_threads();
    std::vector<hll_t> refs = make_sketches(ref_regs, nr, p, estim, jestim);
    std::vector<hll_t> qrys = make_sketches(qry_regs, nq, p, estim, jestim);
    const float ksinv = 1. / k;
    const EmissionType rt = (EmissionType)rtype;
    for(uint64_t qi = 0; qi < nq; ++qi) {
        const hll_t &hq = qrys[qi];
        float *arr = out + qi * nr;
        #pragma omp parallel for schedule(dynamic) num_threads(nthreads)
        for(uint64_t j = 0; j < nr; ++j) arr[j] = result_cmp(refs[j], hq, rt, ksinv);
    }
    return 0;
}

// Nearest neighbours through the reference's own perform_nns (src/sketch_and_cmp.h:642-697), values as in nndist_loop
// (:729-730: result_cmp with ksinv = 1./k in DOUBLE).  regs = all n sketches; nq > 0: the last nq are queries.
// out: validx_t[rows][nn], rows = nq ? nq : n.  nthreads = 1 makes the symmetric mode deterministic.
int dref_knn(const uint8_t *regs, uint64_t n, int p, int k, int estim, int jestim, int rtype, uint64_t nq, uint32_t nn,
             int nthreads, void *out) {
    static_assert(sizeof(validx_t) == 8, "validx_t layout");
    if(nthreads <= 0) nthreads = omp_get_max_threads();
    const int prev = omp_get_max_threads();
    omp_set_num_threads(nthreads);
    std::vector<hll_t> sk = make_sketches(regs, n, p, estim, jestim);
    std::vector<std::string> inpaths(n);
    const EmissionType rt = (EmissionType)rtype;
    const double ksinv = 1. / k;
    auto call_cmp = [rt, ksinv](const auto &x, const auto &y) {return result_cmp(x, y, rt, ksinv);};
    perform_nns(static_cast<validx_t *>(out), sk.data(), inpaths, (unsigned)k, rt, (size_t)nq, (unsigned)nn,
                emt2nntype(rt) == SIMILARITY_MEASURE, call_cmp);
    omp_set_num_threads(prev);
    return 0;
}

// .hll container via the reference's own writer/reader (gz).  A freshly sketched hll_t is written
// with value_ = -1 ("not calculated"), as sketch_core does (src/sketch_and_cmp.h:522).
int dref_hll_write(const char *path, const uint8_t *regs, int p, int estim, int jestim, int calculated) {
    try {
        hll_t h = make_sketch(regs, p, estim, jestim);
        if(calculated) h.report();
        h.write(path);
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_hll_write: %s\n", e.what()); return 1; }
    return 0;
}

int dref_hll_read(const char *path, uint8_t *regs_out, uint64_t cap, int *p_out, int *estim_out, int *jestim_out, double *value_out) {
    try {
        hll_t h(path);
        if(h.size() > cap) return 2;
        std::memcpy(regs_out, h.data(), h.size());
        *p_out = h.p(); *estim_out = h.get_estim(); *jestim_out = h.get_jestim(); *value_out = h.creport();
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_hll_read: %s\n", e.what()); return 1; }
    return 0;
}

// make_fname (src/dashing.h:497-526) for the HLL sketch type.
int dref_make_fname(const char *path, int p, int wsz, int k, int csz, const char *spacing, const char *suffix,
                    const char *prefix, char *out, uint64_t cap) {
    std::string s = make_fname<hll_t>(path, p, wsz, k, csz, spacing, suffix, prefix);
    if(s.size() + 1 > cap) return 1;
    std::memcpy(out, s.data(), s.size() + 1);
    return 0;
}


// ---- the reference's own drivers, end to end -------------------------------------------------------------------
// `dashing dist` for the HLL sketch type: what dist_main does after option parsing (src/distmain.cpp:150-178), minus
// the file-size sort (callers pass paths in final order, i.e. --avoid-sorting).  The last nq paths are queries.
int dref_cli_dist(int npaths, const char **paths, int nq, int k, int p, int canon, int estim, int jestim, int rtype, int emit_fmt,
                  int presketched, int nthreads, const char *sizes_path, const char *dist_path, int cache, const char *prefix,
                  const char *suffix) {
    try {
        std::vector<std::string> inpaths(paths, paths + npaths);
        std::vector<CountingSketch> cms;
        KSeqBufferHolder kseqs(nthreads);
        std::FILE *ofp = std::fopen(sizes_path, "w"), *pairofp = std::fopen(dist_path, "wb");
        if(!ofp || !pairofp) return 2;
        omp_set_num_threads(nthreads);
        Spacer sp(k, 0);
        dist_sketch_and_cmp<hll::hll_t>(inpaths, cms, kseqs, ofp, pairofp, dist_path, sp, p, 5, (hll::EstimationMethod)estim,
                                        (hll::JointEstimationMethod)jestim, cache != 0, (EmissionType)rtype, (EmissionFormat)emit_fmt,
                                        presketched != 0, nthreads, false, suffix, prefix, canon != 0, false, "", nq, BONSAI);
        if(pairofp) std::fclose(pairofp);
        if(emit_fmt == BINARY) { // labels file, src/distmain.cpp:191-200
            std::FILE *fp = std::fopen((std::string(dist_path) + ".labels").data(), "wb");
            for(const auto &path: inpaths) std::fwrite(path.data(), path.size(), 1, fp), std::fputc('\n', fp);
            std::fclose(fp);
        }
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_cli_dist: %s\n", e.what()); return 1; }
    return 0;
}

// --nearest-neighbors N (src/distmain.cpp:89-93, :106-109): N > 0 and emit_fmt | NEAREST_NEIGHBOR_TABLE (8) make
// dist_sketch_and_cmp call nndist_loop instead of dist_loop (src/sketch_and_cmp.h:398-400).
void dref_set_nneighbors(uint32_t n) { gargs.number_neighbors = n; }

// `dashing sketch` for the HLL sketch type (src/dashing.cpp:374-394 -> sketch_core<hll_t>), paths in final order.
int dref_cli_sketch(int npaths, const char **paths, int k, int p, int canon, int nthreads, const char *prefix, const char *suffix,
                    int skip_cached) {
    try {
        std::vector<std::string> inpaths(paths, paths + npaths);
        std::vector<CountingSketch> cms;
        std::vector<bool> use_filter;
        KSeqBufferHolder kseqs(nthreads);
        omp_set_num_threads(nthreads);
        Spacer sp(k, 0);
        const int flags = skip_cached | (int(canon != 0) << 1);       // src/dashing.cpp:375
        sketch_core<hll::hll_t>(p, nthreads, sp.c_, k, sp, inpaths, suffix, prefix, cms, hll::ERTL_MLE,
                                (hll::JointEstimationMethod)hll::ERTL_MLE, kseqs, use_filter, "", flags, 1, BONSAI, "");
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_cli_sketch: %s\n", e.what()); return 1; }
    return 0;
}

// ---- union / fold / hll / sketch -o / sketch_by_seq / card ------------------------------------
// hll_t::compress (hll.h:903-924)
int dref_compress(const uint8_t *regs, int p, int new_p, uint8_t *out) {
    try {
        hll_t h = make_sketch(regs, p, 2, 2);
        hll_t c = h.compress(new_p);
        std::memcpy(out, c.data(), c.size());
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_compress: %s\n", e.what()); return 1; }
    return 0;
}

// hll_t::operator+= folded over n in-memory sketches (hll.h:958-992)
int dref_union(const uint8_t *regs, uint64_t n, int p, uint8_t *out) {
    hll_t acc(p);
    const size_t m = size_t(1) << p;
    for(uint64_t i = 0; i < n; ++i) acc += make_sketch(regs + i * m, p, 2, 2);
    std::memcpy(out, acc.data(), m);
    return 0;
}

// `dashing union` / `dashing hll` / `dashing fold` / `dashing view`: the reference's own mains (union.cpp, hllmain.cpp)
// or, for the two that live in src/dashing.cpp (not compiled here), their bodies (fold_main :575-595, view_main :562-566).
int dref_union_main(int argc, char **argv) {
    optind = 1;
    try { return union_main(argc, argv); } catch(const std::exception &e) { std::fprintf(stderr, "dref_union_main: %s\n", e.what()); return 1; }
}
int dref_hll_main(int argc, char **argv) {
    optind = 1;
    try { return hll_main(argc, argv); } catch(const std::exception &e) { std::fprintf(stderr, "dref_hll_main: %s\n", e.what()); return 1; }
}
int dref_fold(const char *in, const char *out, int destp) {
    try {
        hll_t h{std::string(in)};
        if(destp <= 0) destp = h.p() - 1;
        h.compress(destp).write(out);
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_fold: %s\n", e.what()); return 1; }
    return 0;
}
int dref_view(const char *in, const char *out) {
    try {
        std::FILE *fp = std::fopen(out, "w");
        if(!fp) return 2;
        hll_t(std::string(in)).printf(fp);
        std::fclose(fp);
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_view: %s\n", e.what()); return 1; }
    return 0;
}

// `dashing sketch -o FILE`: sketch_core<hll_t> with an output_file (src/sketch_and_cmp.h:466-536): one gzip stream of
// all sketches + FILE.labels.gz
int dref_cli_sketch_container(int npaths, const char **paths, int k, int p, int canon, int nthreads, const char *output_file) {
    try {
        std::vector<std::string> inpaths(paths, paths + npaths);
        std::vector<CountingSketch> cms;
        std::vector<bool> use_filter;
        KSeqBufferHolder kseqs(nthreads);
        omp_set_num_threads(nthreads);
        Spacer sp(k, 0);
        const int flags = (int(canon != 0) << 1);
        sketch_core<hll::hll_t>(p, nthreads, sp.c_, k, sp, inpaths, "", "", cms, hll::ERTL_MLE,
                                (hll::JointEstimationMethod)hll::ERTL_MLE, kseqs, use_filter, "", flags, 1, BONSAI, output_file);
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_cli_sketch_container: %s\n", e.what()); return 1; }
    return 0;
}

// `dashing sketch_by_seq --defer-hll` == sketch_by_seq_core<hll_t> (src/dashing.cpp:542: the flag test is inverted there; the
// default instantiates HyperLogLogHasher, whose inherited write() emits a b-bit minhash, not an HLL)
int dref_cli_sketch_by_seq(const char *inpath, const char *outpath, int k, int p, int canon, int estim, int jestim) {
    try {
        Spacer sp(k, 0);
        const int flags = (int(canon != 0) << 1);
        sketch_by_seq_core<hll::hll_t>(p, 1, sp, inpath, outpath, nullptr, (hll::EstimationMethod)estim, (hll::JointEstimationMethod)jestim,
                                       false, flags, 1, BONSAI);
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_cli_sketch_by_seq: %s\n", e.what()); return 1; }
    return 0;
}

// `dashing dist_by_seq`: dist_by_seq<hll_t> (src/sketch_and_cmp.h:76-118)
int dref_cli_dist_by_seq(const char *namefile, const char *datapath, const char *outpath, int k, int estim, int jestim, int rtype,
                         int emit_fmt, int nthreads, int skip_header) {
    try {
        auto labels = get_paths(namefile);
        if(skip_header && !labels.empty()) labels.erase(labels.begin());
        std::FILE *ofp = std::fopen(outpath, "wb");
        if(!ofp) return 2;
        omp_set_num_threads(nthreads);
        dist_by_seq<hll::hll_t>(labels, datapath, ofp, outpath, k, (hll::EstimationMethod)estim, (hll::JointEstimationMethod)jestim,
                                (EmissionType)rtype, (EmissionFormat)emit_fmt, nthreads, "");
        // dist_loop's BINARY branch closes the stream itself (src/sketch_and_cmp.h:845); dist_by_seq_main closes it again
        // (src/distbyseq.cpp:136) — a double fclose this driver does not repeat
        if(emit_fmt != BINARY) std::fclose(ofp);
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_cli_dist_by_seq: %s\n", e.what()); return 1; }
    return 0;
}

// `dashing card`: size_sketch_and_emit<hll_t> (src/sketch_and_cmp.h:122-265) as card_main calls it (src/cardmain.cpp:4-6 —
// note the call passes (prefix, suffix) where the callee expects (suffix, prefix))
int dref_cli_card(int npaths, const char **paths, int k, int p, int canon, int estim, int jestim, int presketched, int emit_binary,
                  int use_scientific, int nthreads, const char *outpath) {
    try {
        std::vector<std::string> inpaths(paths, paths + npaths);
        std::vector<CountingSketch> cms;
        KSeqBufferHolder kseqs(nthreads);
        omp_set_num_threads(nthreads);
        std::FILE *ofp = std::fopen(outpath, "w");
        if(!ofp) return 2;
        Spacer sp(k, 0);
        size_sketch_and_emit<hll::hll_t>(inpaths, cms, kseqs, ofp, sp, p, 5, BONSAI, (hll::EstimationMethod)estim,
                                         (hll::JointEstimationMethod)jestim, false, emit_binary != 0, use_scientific != 0, presketched != 0,
                                         nthreads, "", "", canon != 0, "");
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_cli_card: %s\n", e.what()); return 1; }
    return 0;
}

// dist --defer-hll (src/distmain.cpp:177): dist_sketch_and_cmp<HyperLogLogHasher<>>
int dref_cli_dist_defer(int npaths, const char **paths, int nq, int k, int p, int canon, int estim, int jestim, int rtype, int emit_fmt,
                        int nthreads, const char *sizes_path, const char *dist_path, int cache, const char *prefix, const char *suffix) {
    try {
        std::vector<std::string> inpaths(paths, paths + npaths);
        std::vector<CountingSketch> cms;
        KSeqBufferHolder kseqs(nthreads);
        std::FILE *ofp = std::fopen(sizes_path, "w"), *pairofp = std::fopen(dist_path, "wb");
        if(!ofp || !pairofp) return 2;
        omp_set_num_threads(nthreads);
        Spacer sp(k, 0);
        dist_sketch_and_cmp<HyperLogLogHasher<>>(inpaths, cms, kseqs, ofp, pairofp, dist_path, sp, p, 5, (hll::EstimationMethod)estim,
                                        (hll::JointEstimationMethod)jestim, cache != 0, (EmissionType)rtype, (EmissionFormat)emit_fmt,
                                        false, nthreads, false, suffix, prefix, canon != 0, false, "", nq, BONSAI);
        if(pairofp) std::fclose(pairofp);
    } catch(const std::exception &e) { std::fprintf(stderr, "dref_cli_dist_defer: %s\n", e.what()); return 1; }
    return 0;
}

} // extern "C"
