"""All-pairs helpers over HyperLogLog sketches — the second front-end onto the all-pairs path, mirroring the
functions of the reference's Python module `sketch_util` (bonsai/hll/python/util.cpp:148-165, functors in
bonsai/hll/python/pysketch.h:25-146) for its `hll` type (an hll_t built as `hll(p)`: ERTL_MLE estimates, union path).

Where the reference takes a list of `hll` objects these take the register arrays: a uint8 array [n][2^p] or a list of n
arrays of 2^p registers (what `hll_t::core()` holds).  Every comparison runs on the GPU through libdashing_b200's C ABI
(dashing_b200/capi.py); there is no CPU path here.

    jaccard_matrix(sketches)                -> float32[n(n-1)/2]   x.jaccard_index(y)                    CmpFunc / JIF
    intersection_matrix(sketches)           -> float32[n(n-1)/2]   intersection_size(x, y)               CmpFunc / ISF   (hll.h:1321-1323)
    union_size_matrix(sketches)             -> float32[n(n-1)/2]   x.union_size(y)                       CmpFunc / USF   (hll.h:1125-1141)
    symmetric_containment_matrix(sketches)  -> float32[n(n-1)/2]   intersection / min(|x|, |y|)          CmpFunc / SCF
    containment_matrix(sketches)            -> float32[n][n]       x_i.containment_index(x_j)            AsymmetricCmpFunc / CSF (hll.h:1161-1164)
    tri2full(flat), ij2ind(i, j, n)                                packed upper triangle <-> square matrix (util.cpp:133-148)

Flat results are in the packed upper-triangular order of util.cpp:85-88 / distmat.h:260-276: ij2ind(i, j, n) for i < j.
"""
from __future__ import annotations

import numpy as np

from . import capi


def _registers(sketches):
    regs = np.ascontiguousarray(np.stack([np.asarray(s, dtype=np.uint8).reshape(-1) for s in sketches]) if isinstance(sketches, (list, tuple))
                                else np.asarray(sketches, dtype=np.uint8))
    if regs.ndim != 2 or regs.shape[1] & (regs.shape[1] - 1) or regs.shape[1] < 128:
        raise ValueError("sketches: n register arrays of 2^p (p >= 7) uint8 values each")
    return regs, int(regs.shape[1]).bit_length() - 1


def ij2ind(i: int, j: int, n: int) -> int:
    """Index of the pair (i, j), i != j, in the packed upper triangle (util.cpp:148)."""
    if i > j:
        i, j = j, i
    return (i * (n * 2 - i - 1)) // 2 + j - (i + 1)


def flat2fullsz(nflat: int) -> int:
    """n such that n(n-1)/2 == nflat (pysketch.h:17-23)."""
    n = int((1 + (1 + 8 * nflat) ** 0.5) / 2)
    for cand in (n - 1, n, n + 1):
        if cand >= 0 and cand * (cand - 1) // 2 == nflat:
            return cand
    raise ValueError("Failed to extract correct size")


def tri2full(flat, diagonal: float = 1.0) -> np.ndarray:
    """Packed upper triangle -> symmetric n x n matrix; the reference fills the diagonal with 1 (util.cpp:133-147: "Jaccard index
    is 1 for this case") and, like it, this writes the UPPER triangle only — the lower one is mirrored here as well, which the
    reference leaves uninitialised."""
    flat = np.asarray(flat, dtype=np.float32)
    n = flat2fullsz(flat.size)
    out = np.empty((n, n), dtype=np.float32)
    iu = np.triu_indices(n, 1)
    out[iu] = flat
    out[(iu[1], iu[0])] = flat
    np.fill_diagonal(out, diagonal)
    return out


def jaccard_matrix(sketches, device: int = 0) -> np.ndarray:
    regs, p = _registers(sketches)
    return capi.dist_symmetric(regs, p, result_type=capi.JI, device=device)


def intersection_matrix(sketches, device: int = 0) -> np.ndarray:
    regs, p = _registers(sketches)
    return capi.dist_symmetric(regs, p, result_type=capi.SIZES, device=device)       # max(0, |x| + |y| - |x u y|)


def union_size_matrix(sketches, device: int = 0) -> np.ndarray:
    regs, p = _registers(sketches)
    return capi.dist_symmetric(regs, p, result_type=capi.UNION_SIZE, device=device)


def symmetric_containment_matrix(sketches, device: int = 0) -> np.ndarray:
    """intersection_size(x, y) / min(x.report(), y.report()) (pysketch.h:125-129)."""
    regs, p = _registers(sketches)
    n = regs.shape[0]
    inter = capi.dist_symmetric(regs, p, result_type=capi.SIZES, device=device).astype(np.float64)
    card = capi.cardinalities(regs, p, device=0 if device == capi.ALL_DEVICES else device)
    iu = np.triu_indices(n, 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / np.minimum(card[iu[0]], card[iu[1]])).astype(np.float32)


def containment_matrix(sketches, device: int = 0) -> np.ndarray:
    """out[i][j] = x_i.containment_index(x_j) = I / (I + max(|x_i| - I, 0)), I the intersection size (hll.h:1161-1173); n x n,
    asymmetric, 1 on the diagonal of non-empty sketches."""
    regs, p = _registers(sketches)
    n = regs.shape[0]
    inter = tri2full(capi.dist_symmetric(regs, p, result_type=capi.SIZES, device=device), diagonal=0.0).astype(np.float64)
    card = capi.cardinalities(regs, p, device=0 if device == capi.ALL_DEVICES else device)
    inter[np.diag_indices(n)] = card                    # |x u x| = |x|  ->  I = |x|
    only_i = np.maximum(card[:, None] - inter, 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / (inter + only_i)).astype(np.float32)
