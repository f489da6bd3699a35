"""Parity policy shared by the tests (DESIGN.md "Parity bar").

Integer / byte results (registers, histograms, k-mers, file payloads): bit-exact.

Floating point (estimator values, Jaccard / Mash / containment floats): BASELINE.json's tolerance,
1e-6 relative.  Two refinements, both forced by the reference's own arithmetic:

* Cancellation residues.  Quantities such as |A \\ B| = cA - I are differences of ~1e5-sized doubles;
  when the true value is 0 the reference returns whatever its compiler's FMA contraction leaves
  (0, 1.4e-11, 2.9e-11 ...).  The error is TRUE relative error, |got - want| / |want|, wherever the
  checker's value is not such a residue; only where |want| < 1e-9 * scale (scale = 1 for index / distance
  outputs, the largest finite cardinality for SIZES) is it measured against `scale` instead, because
  a value that small is decided by the last bits of ~scale-sized operands in the reference itself.
* Discontinuity at 0.  dist_index / containment_dist map ji == 0 to exactly 1 and ji = 1e-17 to ~1.17
  (src/dashing.h:154-165), and 0/0 vs 1e-11/1e-11 decides between NaN and 1.  Where the checker's own
  intersection size for a pair is such a residue (< 1e-9 of the largest cardinality), infinite or
  NaN, the values derived from it are not compared (`unstable_size`).
"""
import numpy as np

RTOL = 1e-6
RESIDUE = 1e-9      # |want| below RESIDUE * scale: a cancellation residue, compared on the absolute scale


def rel_err(got, want, scale=1.0):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    with np.errstate(invalid="ignore", divide="ignore"):
        den = np.where(np.abs(want) >= RESIDUE * scale, np.abs(want), scale)
        err = np.abs(got - want) / den
    err = np.where(same, 0.0, err)
    return np.where(np.isnan(err), np.inf, err)


def assert_close(got, want, scale=1.0, rtol=RTOL, what="", ignore=None):
    err = rel_err(got, want, scale)
    if ignore is not None:
        err = np.where(ignore, 0.0, err)
    bad = err > rtol
    if bad.any():
        idx = np.nonzero(bad.ravel())[0][:5]
        g = np.asarray(got, dtype=np.float64).ravel()[idx]
        w = np.asarray(want, dtype=np.float64).ravel()[idx]
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} values differ by > {rtol:g} rel; first idx {idx}, got {g}, want {w}")


def unstable_size(intersection_want, scale, eps=1e-9):
    """Pairs whose intersection size in the checker is a cancellation residue (|I| < eps * scale),
    infinite or NaN: every value derived from it (similarities, and distances through the
    discontinuity at 0) is decided by rounding noise in the reference itself and is not compared."""
    s = np.asarray(intersection_want, dtype=np.float64)
    return ~(np.isfinite(s) & (np.abs(s) >= eps * scale))


def assert_knn_close(got, want, rtol=RTOL, what=""):
    """Nearest-neighbour tables [rows][nn] of (value, index).

    Values: position by position within `rtol` (the filler +-FLT_MAX must match exactly).  Indices: equal wherever the
    values around them are distinct beyond `rtol`; where several candidates are within `rtol` of each other the order
    (and, at the cut, the membership) is decided by the last float bit, so an index mismatch at position j is accepted
    only if the checker's row holds that index at a value within `rtol` of position j's, or does not hold it at all and
    position j's value is within `rtol` of the row's last (worst retained) value."""
    gv, gi = np.asarray(got["value"], dtype=np.float64), np.asarray(got["index"])
    wv, wi = np.asarray(want["value"], dtype=np.float64), np.asarray(want["index"])
    assert gv.shape == wv.shape, f"{what}: shape {gv.shape} vs {wv.shape}"
    assert_close(gv, wv, rtol=rtol, what=what + " values")
    bad_rows, bad_cols = np.nonzero(gi != wi)
    for r, j in zip(bad_rows, bad_cols):
        v = gv[r, j]
        pos = np.nonzero(wi[r] == gi[r, j])[0]
        near = lambda x: abs(x - v) <= rtol * max(min(abs(x), abs(v)), RESIDUE)
        if pos.size:
            ok = near(wv[r, pos[0]])
        else:
            ok = near(wv[r, -1])
        if not ok:
            raise AssertionError(f"{what}: row {r} slot {j}: got ({v}, {gi[r, j]}), want ({wv[r, j]}, {wi[r, j]})\n"
                                 f" got  {got[r]}\n want {want[r]}")
