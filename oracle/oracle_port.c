/* oracle/oracle_port.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (no intrinsics, scalar) restatement of the two dashing hot paths, written from the
 * specification in SURVEY.md Appendix A and pinned against the real reference (oracle/_ref,
 * built from /root/reference by oracle/Makefile) plus the committed golden vectors under
 * tests/golden/.  Parity status: PINNED (tests/test_oracle_pinning.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (dashing_b200/) never links, imports or calls it.
 *
 * Compile with -ffp-contract=off so the floating-point operation order below is what runs.
 *
 * Each function cites the reference lines (relative to /root/reference) it restates.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------------
 * a4. Thomas Wang 64-bit mix — bonsai/hll/include/sketch/hash.h:40-49
 * ------------------------------------------------------------------------------------------- */
ORC_API uint64_t orc_wang(uint64_t x) {
    x = ~x + (x << 21);
    x ^= x >> 24;
    x += (x << 3) + (x << 8);
    x ^= x >> 14;
    x += (x << 2) + (x << 4);
    x ^= x >> 28;
    x += x << 31;
    return x;
}

/* ---------------------------------------------------------------------------------------------
 * a2. DNA4 alphabet — bonsai/include/bonsai/alphabet.h:128 (table built at :30-59).
 * ACGT/acgt -> 0..3; every other byte (including U/u, N, and all bytes >= 0x80) is invalid.
 * (The reference indexes its LUT with a plain `char`, so bytes >= 0x80 are undefined behaviour
 * there; this restatement — and the product — define them as invalid.)
 * ------------------------------------------------------------------------------------------- */
static int dna4(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

/* a3. reverse complement / canonical form — bonsai/include/bonsai/kmerutil.h:83-90, :137-140.
 * Written base-by-base (the reference uses a 5-stage bit swap; same function). */
static uint64_t revcomp(uint64_t kmer, int k) {
    uint64_t rc = 0;
    for (int i = 0; i < k; ++i) {
        rc = (rc << 2) | (3u - (kmer & 3u));
        kmer >>= 2;
    }
    return rc;
}

ORC_API uint64_t orc_canonical(uint64_t kmer, int k) {
    uint64_t rc = revcomp(kmer, k);
    return kmer < rc ? kmer : rc;
}

/* ---------------------------------------------------------------------------------------------
 * a5. register update — hll.h:828-836 (add) with hll.h:843-846 (addh).
 * idx = top p bits; rho = 1 + leading zeros of the low (64-p) bits, capped at 64-p+1.
 * ------------------------------------------------------------------------------------------- */
static void hll_add_hash(uint8_t *regs, int p, uint64_t h) {
    uint64_t idx = h >> (64 - p);
    uint64_t w = ((h << 1) | 1u) << (p - 1);
    uint8_t rho = (uint8_t)(__builtin_clzll(w) + 1);
    if (regs[idx] < rho) regs[idx] = rho;
}

ORC_API void orc_addh(uint8_t *regs, int p, uint64_t element) { hll_add_hash(regs, p, orc_wang(element)); }

/* ---------------------------------------------------------------------------------------------
 * a1. k-mer stream of one record — encoder.h:240-271 (window logic), :218-232 (canonicalisation).
 * Calls emit(kmer) for every window of k valid bases; an invalid byte restarts the window after it.
 * Returns the number of k-mers emitted; if out != NULL the first `cap` are stored.
 * ------------------------------------------------------------------------------------------- */
ORC_API uint64_t orc_kmers(const char *s, uint64_t len, int k, int canon, uint64_t *out, uint64_t cap) {
    const uint64_t mask = k == 32 ? ~UINT64_C(0) : ((UINT64_C(1) << (2 * k)) - 1);
    uint64_t kmer = 0, n = 0;
    int filled = 0;
    for (uint64_t pos = 0; pos < len; ++pos) {
        int v = dna4((unsigned char)s[pos]);
        if (v < 0) { kmer = 0; filled = 0; continue; }
        kmer = ((kmer << 2) | (uint64_t)v) & mask;
        if (++filled >= k) {
            filled = k;
            uint64_t e = canon ? orc_canonical(kmer, k) : kmer;
            if (out && n < cap) out[n] = e;
            ++n;
        }
    }
    return n;
}

/* a1+a4+a5 fused: sketch the records [0,nrec) of one genome into regs (2^p bytes, zeroed here).
 * Mirrors the lambda `h.addh(kmer)` fed to Encoder::for_each in sketch_core (src/sketch_and_cmp.h:512). */
ORC_API int orc_sketch(const char *bases, const uint64_t *offsets, uint64_t nrec, int k, int p, int canon, uint8_t *regs) {
    const uint64_t mask = k == 32 ? ~UINT64_C(0) : ((UINT64_C(1) << (2 * k)) - 1);
    memset(regs, 0, (size_t)1 << p);
    for (uint64_t r = 0; r < nrec; ++r) {
        const char *s = bases + offsets[r];
        const uint64_t len = offsets[r + 1] - offsets[r];
        uint64_t kmer = 0;
        int filled = 0;
        for (uint64_t pos = 0; pos < len; ++pos) {
            int v = dna4((unsigned char)s[pos]);
            if (v < 0) { kmer = 0; filled = 0; continue; }
            kmer = ((kmer << 2) | (uint64_t)v) & mask;
            if (++filled >= k) {
                filled = k;
                hll_add_hash(regs, p, orc_wang(canon ? orc_canonical(kmer, k) : kmer));
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * a9. histogram of one register array — hll.h:515-532; of a pair's element-wise max — hll.h:1131-1136
 * ------------------------------------------------------------------------------------------- */
ORC_API void orc_histogram(const uint8_t *regs, int p, uint32_t *c64) {
    memset(c64, 0, 64 * sizeof(uint32_t));
    for (size_t i = 0, m = (size_t)1 << p; i < m; ++i) ++c64[regs[i]];
}

static void union_histogram(const uint8_t *a, const uint8_t *b, int p, uint32_t *c64) {
    memset(c64, 0, 64 * sizeof(uint32_t));
    for (size_t i = 0, m = (size_t)1 << p; i < m; ++i) ++c64[a[i] > b[i] ? a[i] : b[i]];
}

/* ---------------------------------------------------------------------------------------------
 * a11. Ertl maximum-likelihood estimator — hll.h:567-627.  c has q+2 live bins (0..q+1).
 * ------------------------------------------------------------------------------------------- */
ORC_API double orc_mle(const uint32_t *c, int p, int q) {
    const uint64_t m = UINT64_C(1) << p;
    const double relerr = 1e-2 / sqrt((double)m);
    if (c[q + 1] == m) return INFINITY;

    int kmin = 0, kmax = q + 1;
    while (c[kmin] == 0) ++kmin;
    while (kmax && c[kmax] == 0) --kmax;
    const int lo = kmin > 1 ? kmin : 1;    /* kMin' */
    const int hi = kmax < q ? kmax : q;    /* kMax' */

    double z = 0.;
    for (int k = hi; k >= lo; --k) z = 0.5 * z + c[k];
    z = ldexp(z, -lo);

    unsigned cprime = c[q + 1];
    if (q) cprime += c[hi];

    const double a = z + c[0];
    const int mprime = (int)(m - c[0]);
    double g0 = z + ldexp((double)c[q + 1], -q);
    double x = g0 <= 1.5 * a ? mprime / (0.5 * g0 + a) : (mprime / g0) * log1p(g0 / a);
    double gprev = 0., dx = x;

    while (dx > x * relerr) {
        int kappa;
        frexp(x, &kappa);
        const int sh = (hi + 1) > (kappa + 2) ? (hi + 1) : (kappa + 2);
        double xp = ldexp(x, -sh);
        const double xp2 = xp * xp;
        double h = xp - xp2 / 3 + (xp2 * xp2) * (1. / 45. - xp2 / 472.5);
        for (int k = kappa; k >= hi; --k) {
            const double hc = 1. - h;
            h = (xp + h * hc) / (xp + hc);
            xp += xp;
        }
        double g = cprime * h;
        for (int k = hi - 1; k >= lo; --k) {
            const double hc = 1. - h;
            h = (xp + h * hc) / (xp + hc);
            xp += xp;
            g += c[k] * h;
        }
        g += x * a;
        if (gprev < g && g <= mprime) dx *= (g - mprime) / (gprev - g);
        else dx = 0;
        x += dx;
        gprev = g;
    }
    return x * (double)m;
}

/* gen_sigma / gen_tau — hll.h:23-51 */
static double ertl_sigma(double x) {
    if (x == 1.) return INFINITY;
    double z = x, zp = 0., y = 1.;
    while (z != zp) {
        x *= x; zp = z; z += x * y; y += y;
        if (isnan(z)) return zp;
    }
    return z;
}

static double ertl_tau(double x) {
    if (x == 0. || x == 1.) return 0.;
    double z = 1 - x, y = 1., zp = x;
    while (zp != z) {
        x = sqrt(x);
        zp = z;
        y *= 0.5;
        const double t = 1. - x;
        z -= t * t * y;
    }
    return z / 3.;
}

/* make_alpha — hll.h:694-701 */
static double hll_alpha(uint64_t m) {
    if (m == 16) return .673;
    if (m == 32) return .697;
    if (m == 64) return .709;
    return 0.7213 / (1 + 1.079 / (double)m);
}

/* a10. calculate_estimate — hll.h:199-246.  estim: 0 ORIGINAL, 1 ERTL_IMPROVED, 2 ERTL_MLE. */
ORC_API double orc_estimate(const uint32_t *c, int p, int estim) {
    const uint64_t m = UINT64_C(1) << p;
    const int q = 64 - p;
    if (estim == 0) {
        double sum = c[0];
        for (int i = 1; i < q + 1; ++i) if (c[i]) sum += ldexp((double)c[i], -i);
        double v = hll_alpha(m) * (double)m * (double)m / sum;
        if (v < 2.5 * (double)m) {
            if (c[0]) v = (double)m * log((double)m / c[0]);
        } else if (v > 4294967296. / 30.) {
            /* the reference evaluates -2^32 * log1p(-v/2^32) with a long double factor; the
             * scaling by 2^32 is exact, so double arithmetic gives the same double */
            const double corr = -4294967296. * log1p(-ldexp(v, -32));
            if (!isnan(corr)) v = corr;
        }
        return v;
    }
    if (estim == 1) {
        /* divinv = (double)(1 / (2 ln 2)) evaluated in long double by the reference (hll.h:233) */
        const double divinv = (double)(1.L / (2.L * 0.693147180559945309417232121458176568L));
        double z = (double)m * ertl_tau((double)(m - c[q + 1]) / (double)m);
        for (int i = q; i; --i) { z += c[i]; z *= 0.5; }
        z += (double)m * ertl_sigma((double)c[0] / (double)m);
        return (double)m * divinv * (double)m / z;
    }
    return orc_mle(c, p, q);
}

ORC_API double orc_cardinality(const uint8_t *regs, int p, int estim) {
    uint32_t c[64];
    orc_histogram(regs, p, c);
    return orc_estimate(c, p, estim);
}

/* ---------------------------------------------------------------------------------------------
 * a12-a14. pair quantities.  cA / cB are the cached per-sketch cardinalities (the dist driver has
 * called report() on every sketch before the pair loop, src/sketch_and_cmp.h:377-383).
 * jestim == 3 selects Ertl's joint MLE (hll.h:636-684); anything else the union path (hll.h:1125-1138).
 * out3 = {|A \ B|, |B \ A|, |A ∩ B|} as full_set_comparison (hll.h:1165-1173) returns it.
 * ------------------------------------------------------------------------------------------- */
static void joint_mle_triple(const uint8_t *a, const uint8_t *b, int p, double cA, double cB, double *out3) {
    const int q = 64 - p;
    const size_t m = (size_t)1 << p;
    uint32_t cu[64], cg1[64], cg2[64], ceq[64], ha[64], hb[64];
    memset(cu, 0, sizeof cu); memset(cg1, 0, sizeof cg1); memset(cg2, 0, sizeof cg2); memset(ceq, 0, sizeof ceq);
    for (size_t i = 0; i < m; ++i) {          /* joint_unroller, hll.h:441-503 */
        const uint8_t x = a[i], y = b[i];
        ++cu[x > y ? x : y];
        if (x > y) ++cg1[x];
        else if (y > x) ++cg2[y];
        else ++ceq[x];
    }
    const double cU = orc_mle(cu, p, q);
    memset(ha, 0, sizeof ha); memset(hb, 0, sizeof hb);
    ha[q] = hb[q] = (uint32_t)m;               /* hll.h:660-674 */
    for (int k = 0; k < q; ++k) {
        ha[k] = cg1[k] + ceq[k] + cg2[k + 1];
        ha[q] -= ha[k];
        hb[k] = cg2[k] + ceq[k] + cg1[k + 1];
        hb[q] -= hb[k];
    }
    const double hA = orc_mle(ha, p, q - 1);
    const double hB = orc_mle(hb, p, q - 1);
    out3[0] = cU - cB;
    out3[1] = cU - cA;
    const double x1 = 1.5 * cB + 1.5 * cA - hB - hA;
    const double x2 = 2. * (hB + hA) - 3. * cU;
    const double is = 0.5 * (x1 + x2);
    out3[2] = 0. < is ? is : 0.;                /* std::max(0., x) */
}

ORC_API double orc_union_size(const uint8_t *a, const uint8_t *b, int p, int estim, int jestim, double cA, double cB) {
    if (jestim == 3) {
        double t[3];
        joint_mle_triple(a, b, p, cA, cB, t);
        return t[0] + t[1] + t[2];
    }
    uint32_t c[64];
    union_histogram(a, b, p, c);
    return orc_estimate(c, p, estim);
}

ORC_API void orc_triple(const uint8_t *a, const uint8_t *b, int p, int estim, int jestim, double cA, double cB, double *out3) {
    if (jestim == 3) { joint_mle_triple(a, b, p, cA, cB, out3); return; }
    const double us = orc_union_size(a, b, p, estim, jestim, cA, cB);
    /* std::max(x, 0.) == (x < 0.) ? 0. : x — written out so NaN propagates as in the reference */
    double is = cA + cB - us;
    is = is < 0. ? 0. : is;
    const double ao = cA - is, bo = cB - is;
    out3[0] = ao < 0. ? 0. : ao;
    out3[1] = bo < 0. ? 0. : bo;
    out3[2] = is;
}

ORC_API double orc_jaccard(const uint8_t *a, const uint8_t *b, int p, int estim, int jestim, double cA, double cB) {
    if (jestim == 3) {
        double t[3];
        joint_mle_triple(a, b, p, cA, cB, t);
        return t[2] / (t[0] + t[1] + t[2]);
    }
    const double us = orc_union_size(a, b, p, estim, jestim, cA, cB);
    const double r = (cA + cB - us) / us;
    return 0. < r ? r : 0.;                    /* std::max(0., r): NaN r -> 0. */
}

/* a15. result_cmp — src/dashing.h:568-592 with dist_index :154-156, containment_dist :163-165,
 * full_dist_index :172-174, full_containment_dist :181-183.  EmissionType numbering: src/enums.h:13-23.
 * lhs / rhs keep the operand order of the call site. */
static float result_cmp_dks(const uint8_t *lhs, const uint8_t *rhs, int p, int estim, int jestim, int rtype, double ksinv,
                            double cL, double cR) {
    double ret;
    if (rtype == 0 || rtype == 1 || rtype == 3) {
        ret = orc_jaccard(lhs, rhs, p, estim, jestim, cL, cR);
        if (rtype == 0) ret = ret ? -log(2. * ret / (1. + ret)) * ksinv : 1.;
        else if (rtype == 3) ret = 1. - pow(2. * ret / (1. + ret), ksinv);
    } else {
        double t[3];
        orc_triple(lhs, rhs, p, estim, jestim, cL, cR, t);
        ret = t[2];
        if (rtype == 7 || rtype == 8) {
            ret /= ((t[1] < t[0] ? t[1] : t[0]) + t[2]);  /* std::min(t0, t1) */
            if (rtype == 8) ret = ret ? -log(ret) * ksinv : 1.;
        } else if (rtype == 4 || rtype == 5 || rtype == 6) {
            ret /= (t[0] + t[1] + t[2]);
            if (rtype == 6) ret = ret ? -log(ret) * ksinv : 1.;
            else if (rtype == 4) ret = 1. - pow(ret, ksinv);
        } /* rtype == 2 (SIZES): intersection size */
    }
    return (float)ret;
}

ORC_API float orc_result_cmp(const uint8_t *lhs, const uint8_t *rhs, int p, int estim, int jestim, int rtype, int k,
                             double cL, double cR) {
    /* float ksinv promoted, src/sketch_and_cmp.h:797 */
    return result_cmp_dks(lhs, rhs, p, estim, jestim, rtype, (double)(float)(1. / k), cL, cR);
}

/* a16. symmetric all-pairs, packed upper triangle in distmat order (distmat/distmat.h:260-276).
 * order 0: cmp(s[i], s[j]); order 1: cmp(s[j], s[i]) (src/sketch_and_cmp.h:699-710 vs :829,:849). */
ORC_API int orc_dist_rows(const uint8_t *regs, uint64_t n, int p, int k, int estim, int jestim, int rtype, int order,
                          uint64_t row_begin, uint64_t row_end, float *out) {
    const size_t m = (size_t)1 << p;
    double *card = (double *)malloc(sizeof(double) * (n ? n : 1));
    if (!card) return 1;
    for (uint64_t i = 0; i < n; ++i) card[i] = orc_cardinality(regs + i * m, p, estim);
    for (uint64_t i = row_begin; i < row_end && i + 1 < n; ++i) {
        float *row = out + (i * (2 * n - i - 1)) / 2;
        for (uint64_t j = i + 1; j < n; ++j)
            row[j - i - 1] = order ? orc_result_cmp(regs + j * m, regs + i * m, p, estim, jestim, rtype, k, card[j], card[i])
                                   : orc_result_cmp(regs + i * m, regs + j * m, p, estim, jestim, rtype, k, card[i], card[j]);
    }
    free(card);
    return 0;
}

/* partdist_loop — src/dashing.h:675-681: out[q*nr + j] = result_cmp(refs[j], queries[q]). */
ORC_API int orc_dist_rect(const uint8_t *ref_regs, uint64_t nr, const uint8_t *qry_regs, uint64_t nq, int p, int k,
                          int estim, int jestim, int rtype, float *out) {
    const size_t m = (size_t)1 << p;
    double *cr = (double *)malloc(sizeof(double) * (nr ? nr : 1));
    double *cq = (double *)malloc(sizeof(double) * (nq ? nq : 1));
    if (!cr || !cq) { free(cr); free(cq); return 1; }
    for (uint64_t i = 0; i < nr; ++i) cr[i] = orc_cardinality(ref_regs + i * m, p, estim);
    for (uint64_t i = 0; i < nq; ++i) cq[i] = orc_cardinality(qry_regs + i * m, p, estim);
    for (uint64_t qi = 0; qi < nq; ++qi)
        for (uint64_t j = 0; j < nr; ++j)
            out[qi * nr + j] = orc_result_cmp(ref_regs + j * m, qry_regs + qi * m, p, estim, jestim, rtype, k, cr[j], cq[qi]);
    free(cr); free(cq);
    return 0;
}

/* ---- k nearest neighbours: perform_nns + lock_update / lockfree_update (src/sketch_and_cmp.h:605-697) -------------
 * The reference keeps a binary heap per sketch and replaces its top whenever a new value is strictly better than the
 * top's value; the top of a std::less heap of pairs is the lexicographically largest (value, index), of a std::greater
 * heap the smallest.  Only WHICH pair is evicted matters, so a linear scan for the extreme stands in for the heap.
 * Visiting order: ascending index (the reference's order with one thread; its -Q/-F mode always).  Values: result_cmp
 * with ksinv = 1./k in double (:729), as cmp(sketches[j], sketches[i]) for j > i (:670) / cmp(ref, query) (:686). */
typedef struct { float v; uint32_t i; } orc_nb;

static int nb_less(orc_nb a, orc_nb b) { return a.v < b.v || (!(b.v < a.v) && a.i < b.i); }   /* std::pair operator< */

static void nb_update(orc_nb *heap, uint32_t nn, float val, uint32_t idx, int sim) {
    /* heap "top": sim -> min under nb_less (std::greater heap), else max */
    uint32_t top = 0;
    for (uint32_t e = 1; e < nn; ++e)
        if (sim ? nb_less(heap[e], heap[top]) : nb_less(heap[top], heap[e])) top = e;
    if (sim ? (val > heap[top].v) : (val < heap[top].v)) { heap[top].v = val; heap[top].i = idx; }
}

static int nb_cmp_less(const void *a, const void *b) {
    const orc_nb x = *(const orc_nb *)a, y = *(const orc_nb *)b;
    return nb_less(x, y) ? -1 : (nb_less(y, x) ? 1 : 0);
}
static int nb_cmp_greater(const void *a, const void *b) { return nb_cmp_less(b, a); }

static float result_cmp_dks(const uint8_t *lhs, const uint8_t *rhs, int p, int estim, int jestim, int rtype, double ksinv,
                            double cL, double cR);

/* regs: n sketches; nq > 0: the last nq are the queries and the first n - nq the references.  out: [rows][nn]. */
ORC_API int orc_knn(const uint8_t *regs, uint64_t n, int p, int k, int estim, int jestim, int rtype, uint64_t nq, uint32_t nn,
                    void *out_) {
    orc_nb *out = (orc_nb *)out_;
    const uint64_t m = 1ull << p, rows = nq ? nq : n;
    /* emt2nntype, src/dashing.h:268-280 */
    const int sim = !(rtype == 0 || rtype == 3 || rtype == 4 || rtype == 6 || rtype == 8);
    const double ksinv = 1. / k;
    double *card = (double *)malloc(sizeof(double) * (n ? n : 1));
    if (!card) return 1;
    for (uint64_t i = 0; i < n; ++i) card[i] = orc_cardinality(regs + i * m, p, estim);
    for (uint64_t i = 0; i < rows * nn; ++i) { out[i].v = sim ? -3.402823466e+38f : 3.402823466e+38f; out[i].i = 0xFFFFFFFFu; }
    if (nq == 0) {
        for (uint64_t i = 0; i < n; ++i)
            for (uint64_t j = i + 1; j < n; ++j) {
                const float val = result_cmp_dks(regs + j * m, regs + i * m, p, estim, jestim, rtype, ksinv, card[j], card[i]);
                nb_update(out + i * nn, nn, val, (uint32_t)j, sim);
                nb_update(out + j * nn, nn, val, (uint32_t)i, sim);
            }
    } else {
        const uint64_t nr = n - nq;
        for (uint64_t q = 0; q < nq; ++q)
            for (uint64_t j = 0; j < nr; ++j)
                nb_update(out + q * nn, nn,
                          result_cmp_dks(regs + j * m, regs + (nr + q) * m, p, estim, jestim, rtype, ksinv, card[j], card[nr + q]),
                          (uint32_t)j, sim);
    }
    for (uint64_t r = 0; r < rows; ++r) qsort(out + r * nn, nn, sizeof(orc_nb), sim ? nb_cmp_greater : nb_cmp_less);
    free(card);
    return 0;
}


/* a7. decompressed .hll payload — hll.h:1039-1047: u32[4]{is_calculated, estim, jestim, 1}, u32 p,
 * f64 value, u8[2^p]; 28 + 2^p bytes, little-endian.  Returns bytes written (0 if cap too small). */
ORC_API uint64_t orc_hll_payload(const uint8_t *regs, int p, int estim, int jestim, double value, uint8_t *out, uint64_t cap) {
    const uint64_t m = UINT64_C(1) << p, nb = 28 + m;
    if (cap < nb) return 0;
    uint32_t hdr[5] = { value >= 0. ? 1u : 0u, (uint32_t)estim, (uint32_t)jestim, 1u, (uint32_t)p };
    memcpy(out, hdr, 20);
    memcpy(out + 20, &value, 8);
    memcpy(out + 28, regs, m);
    return nb;
}

/* ---------------------------------------------------------------------------------------------
 * (f)3. union — hll_t::operator+= (bonsai/hll/include/sketch/hll.h:958-992) folded over n sketches as
 * union_core does (src/union.cpp:33-58): element-wise byte maximum.  out: 2^p bytes.
 * ------------------------------------------------------------------------------------------- */
ORC_API void orc_union(const uint8_t *regs, uint64_t n, int p, uint8_t *out) {
    const uint64_t m = 1ull << p;
    memset(out, 0, m);
    for (uint64_t s = 0; s < n; ++s)
        for (uint64_t i = 0; i < m; ++i)
            if (regs[s * m + i] > out[i]) out[i] = regs[s * m + i];
}

/* ---------------------------------------------------------------------------------------------
 * (f)3. fold — hll_t::compress(new_np) (hll.h:903-924).  ratio = 2^(p - new_p) old registers per new one; j = offset of
 * the first non-zero old register of the group; none -> 0, j == 0 -> min(q'+1, old + diff), else min(q'+1, clz64(j) + 1)
 * with q' = 64 - new_p.  (clz is the 64-bit overload, integral.h: j is a size_t.)  Returns 1 for new_p > p (the
 * reference throws "Can't compress to a larger size"), 0 otherwise; new_p == p copies.
 * ------------------------------------------------------------------------------------------- */
ORC_API int orc_compress(const uint8_t *regs, int p, int new_p, uint8_t *out) {
    if (new_p > p) return 1;
    if (new_p == p) { memcpy(out, regs, (size_t)1 << p); return 0; }
    const unsigned diff = (unsigned)(p - new_p);
    const uint64_t ratio = 1ull << diff, new_size = 1ull << new_p;
    const unsigned cap = (unsigned)(64 - new_p) + 1u;
    uint64_t b = 0;
    for (uint64_t i = 0; i < new_size; ++i, b += ratio) {
        uint64_t j = 0;
        while (j < ratio && regs[j + b] == 0) ++j;
        unsigned v = 0;
        if (j != ratio) {
            const unsigned cand = j ? (unsigned)__builtin_clzll(j) + 1u : (unsigned)regs[b] + diff;
            v = cand < cap ? cand : cap;
        }
        out[i] = (uint8_t)v;
    }
    return 0;
}
