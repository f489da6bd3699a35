// estimators.cuh — FP64 cardinality estimators on a 64-bin register histogram, device side.
//
// Every function takes a *counts accessor* `c(k)` -> uint32_t instead of an array so the same code
// serves a per-sketch histogram in global memory and a pair histogram that only exists as
// threshold counts in shared memory (dist.cu).  Operation order follows the reference so results
// agree to the last few ulps (contraction into FMA is the only difference):
//   calculate_estimate   bonsai/hll/include/sketch/hll.h:199-246
//   gen_sigma / gen_tau  hll.h:23-51
//   make_alpha           hll.h:694-701
//   ertl_ml_estimate     hll.h:567-627
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace db200 {

// Ertl's maximum-likelihood estimate (Algorithm 8 of arXiv:1702.01284 as coded at hll.h:567-627).
// kmin_hint <= first non-empty bin, kmax_hint >= last non-empty bin: callers that know the value
// range of the registers pass it so the two scans do not walk 50 empty bins.
// Division used inside the secant iteration.  Default: IEEE division (same rounding as the reference's).  dist.cuh
// specialises the hot pair path with a <= 1 ulp Newton division (tests bound the end-to-end effect at 1e-6 relative).
struct IeeeDiv { __device__ __forceinline__ static double div(double a, double b) { return a / b; } };

template <typename Counts, typename Div = IeeeDiv>
__device__ __forceinline__ double ertl_mle(const Counts &c, int p, int q, int kmin_hint = 0, int kmax_hint = -1) {
    const unsigned long long m = 1ull << p;
    if (c(q + 1) == m) return __longlong_as_double(0x7ff0000000000000ll); // +inf

    int kMin = kmin_hint, kMax = kmax_hint < 0 ? q + 1 : (kmax_hint > q + 1 ? q + 1 : kmax_hint);
    while (c(kMin) == 0) ++kMin;
    while (kMax && c(kMax) == 0) --kMax;
    const int lo = kMin > 1 ? kMin : 1;
    const int hi = kMax < q ? kMax : q;

    double z = 0.;
    for (int k = hi; k >= lo; --k) z = 0.5 * z + (double)c(k);
    z = ldexp(z, -lo);

    unsigned cPrime = c(q + 1);
    if (q) cPrime += c(hi);

    const double c0 = (double)c(0);
    const double a = z + c0;
    const int mPrime = (int)(m - c(0));
    const double g0 = z + ldexp((double)c(q + 1), -q);
    double x = g0 <= 1.5 * a ? mPrime / (0.5 * g0 + a) : (mPrime / g0) * log1p(g0 / a);
    double gprev = 0., dx = x;
    const double relerr = 1e-2 / sqrt((double)m);

    while (dx > x * relerr) {
        int kappa;
        (void)frexp(x, &kappa);
        const int sh = (hi + 1) > (kappa + 2) ? (hi + 1) : (kappa + 2);
        double xp = ldexp(x, -sh);
        const double xp2 = xp * xp;
        double h = xp - xp2 / 3 + (xp2 * xp2) * (1. / 45. - xp2 / 472.5);
        for (int k = kappa; k >= hi; --k) {
            const double hc = 1. - h;
            h = Div::div(xp + h * hc, xp + hc);
            xp += xp;
        }
        double g = cPrime * h;
        for (int k = hi - 1; k >= lo; --k) {
            const double hc = 1. - h;
            h = Div::div(xp + h * hc, xp + hc);
            xp += xp;
            g += (double)c(k) * h;
        }
        g += x * a;
        if (gprev < g && g <= mPrime) dx *= (g - mPrime) / (gprev - g);
        else dx = 0;
        x += dx;
        gprev = g;
    }
    return x * (double)m;
}

__device__ __forceinline__ double ertl_sigma(double x) {
    if (x == 1.) return __longlong_as_double(0x7ff0000000000000ll);
    double z = x, zp = 0., y = 1.;
    while (z != zp) {
        x *= x; zp = z; z += x * y; y += y;
        if (isnan(z)) return zp;
    }
    return z;
}

__device__ __forceinline__ double ertl_tau(double x) {
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

__device__ __forceinline__ double hll_alpha(unsigned long long m) {
    if (m == 16) return .673;
    if (m == 32) return .697;
    if (m == 64) return .709;
    return 0.7213 / (1 + 1.079 / (double)m);
}

// estim: 0 ORIGINAL, 1 ERTL_IMPROVED, 2 ERTL_MLE
template <typename Counts, typename Div = IeeeDiv>
__device__ __forceinline__ double calculate_estimate(const Counts &c, int estim, int p, int kmin_hint = 0, int kmax_hint = -1) {
    const unsigned long long m = 1ull << p;
    const int q = 64 - p;
    const double md = (double)m;
    if (estim == DB200_ORIGINAL) {
        double sum = (double)c(0);
        for (int i = 1; i < q + 1; ++i) {
            const uint32_t ci = c(i);
            if (ci) sum += ldexp((double)ci, -i);
        }
        double v = hll_alpha(m) * md * md / sum;
        if (v < 2.5 * md) {
            if (c(0)) v = md * log(md / (double)c(0));
        } else if (v > 4294967296. / 30.) {
            const double corr = -4294967296. * log1p(-ldexp(v, -32));
            if (!isnan(corr)) v = corr;
        }
        return v;
    }
    if (estim == DB200_ERTL_IMPROVED) {
        const double divinv = 0.72134752044448170367996234050095; // 1 / (2 ln 2), hll.h:233
        double z = md * ertl_tau((double)(m - c(q + 1)) / md);
        for (int i = q; i; --i) { z += (double)c(i); z *= 0.5; }
        z += md * ertl_sigma((double)c(0) / md);
        return md * divinv * md / z;
    }
    return ertl_mle<Counts, Div>(c, p, q, kmin_hint, kmax_hint);
}

// result_cmp epilogue (src/dashing.h:568-592) from the pair quantities.
//   ji    : jaccard_index(lhs, rhs)                      (hll.h:1174-1183)
//   t0..2 : full_set_comparison(lhs, rhs) = {lhs only, rhs only, intersection}  (hll.h:1165-1173)
__device__ __forceinline__ float emit_value(int rtype, double ji, double t0, double t1, double t2, double ksinv) {
    double ret;
    switch (rtype) {
        case DB200_JI: ret = ji; break;
        case DB200_MASH_DIST: ret = ji ? -log(2. * ji / (1. + ji)) * ksinv : 1.; break;       // dist_index, dashing.h:154-156
        case DB200_FULL_MASH_DIST: ret = 1. - pow(2. * ji / (1. + ji), ksinv); break;          // full_dist_index, :172-174
        case DB200_SIZES: ret = t2; break;
        case DB200_SYMMETRIC_CONTAINMENT_INDEX:
        case DB200_SYMMETRIC_CONTAINMENT_DIST: {
            ret = t2 / ((t1 < t0 ? t1 : t0) + t2);                                             // std::min(t0, t1)
            if (rtype == DB200_SYMMETRIC_CONTAINMENT_DIST) ret = ret ? -log(ret) * ksinv : 1.; // containment_dist, :163-165
        } break;
        default: { // CONTAINMENT_INDEX / CONTAINMENT_DIST / FULL_CONTAINMENT_DIST: intersection over UNION (dashing.h:584)
            ret = t2 / (t0 + t1 + t2);
            if (rtype == DB200_CONTAINMENT_DIST) ret = ret ? -log(ret) * ksinv : 1.;
            else if (rtype == DB200_FULL_CONTAINMENT_DIST) ret = 1. - pow(ret, ksinv);         // :181-183
        }
    }
    return (float)ret;
}

} // namespace db200
