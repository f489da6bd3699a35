"""Seeded synthetic inputs for tests and bench.  Host-side numpy only.

* ``genomes``: one ancestor of i.i.d. uniform ACGT per group, genome i = ancestor with per-base
  substitution rate from a fixed ladder so that Jaccard spans (0, 1].
* ``registers``: HLL register arrays synthesised directly (no sketching): each sketch is the
  element-wise max of a shared component and a private component, both drawn from the exact
  register distribution of an HLL holding ``n`` distinct items, so pairs are correlated.
"""
from __future__ import annotations

import numpy as np

RATE_LADDER = (0.0, .001, .002, .005, .01, .02, .03, .05, .07, .1, .15, .2, .3, .5, .75, 1.0)
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def genome(rng: np.random.Generator, length: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]


def mutate(rng: np.random.Generator, anc: np.ndarray, rate: float) -> np.ndarray:
    if rate <= 0.0:
        return anc.copy()
    out = anc.copy()
    hit = rng.random(anc.size) < rate
    nh = int(hit.sum())
    # substitute by a uniformly random *different* base
    code = (np.searchsorted(_ACGT, out[hit]) + rng.integers(1, 4, size=nh)) & 3
    out[hit] = _ACGT[code]
    return out


def genomes(seed: int, n: int, length: int, group: int = 16):
    """-> list of n uint8 ASCII arrays (single-record genomes)."""
    rng = np.random.default_rng(seed)
    out = []
    anc = None
    for i in range(n):
        if i % group == 0:
            anc = genome(rng, length)
        out.append(mutate(rng, anc, RATE_LADDER[i % len(RATE_LADDER)]))
    return out


def sprinkle(rng: np.random.Generator, seq: np.ndarray, n_runs: int = 4, lower_frac: float = 0.1) -> np.ndarray:
    """Targeted edge cases: runs of N (1..100 long), a lower-case stretch, a few IUPAC/other bytes."""
    s = seq.copy()
    L = s.size
    for _ in range(n_runs):
        st = int(rng.integers(0, max(L - 100, 1)))
        ln = int(rng.integers(1, 101))
        s[st:st + ln] = ord("N")
    if lower_frac > 0 and L > 10:
        st = int(rng.integers(0, L // 2))
        s[st:st + int(L * lower_frac)] |= 0x20
    for ch in b"RYKMUu-*\n":
        s[int(rng.integers(0, L))] = ch
    return s


def _rho_sample(rng: np.random.Generator, shape, lam: float, q: int) -> np.ndarray:
    """Register value of an HLL bucket that received Poisson(lam) items: max of geometric ranks.
    P(reg <= r) = exp(-lam * 2^-r) for 0 <= r <= q, P(reg <= q+1) = 1 (0 = empty bucket)."""
    u = rng.random(shape)
    # smallest r >= 0 with exp(-lam 2^-r) >= u  <=>  2^-r <= -ln(u)/lam
    t = -np.log(np.maximum(u, 1e-300)) / lam
    r = np.ceil(-np.log2(np.maximum(t, 2.0 ** -(q + 2))))
    return np.clip(r, 0, q + 1).astype(np.uint8)


def registers(seed: int, n: int, p: int, card: float = 5e6, group: int = 16) -> np.ndarray:
    """-> uint8 [n, 2^p] register matrix with correlated groups (shared ancestor component)."""
    rng = np.random.default_rng(seed)
    m, q = 1 << p, 64 - p
    out = np.empty((n, m), dtype=np.uint8)
    shared = None
    for i in range(n):
        if i % group == 0:
            shared = _rho_sample(rng, m, card / m, q)
            shared_frac = None
        frac = 1.0 - RATE_LADDER[i % len(RATE_LADDER)]
        # a `frac` share of the sketch's items comes from the shared set: thin the shared registers
        # by re-sampling with the reduced rate and taking the min with the full shared sample
        if frac >= 1.0:
            sh = shared
        elif frac <= 0.0:
            sh = np.zeros(m, dtype=np.uint8)
        else:
            sh = np.minimum(shared, _rho_sample(rng, m, frac * card / m, q))
        priv = _rho_sample(rng, m, max(1.0 - frac, 1e-9) * card / m, q) if frac < 1.0 else np.zeros(m, np.uint8)
        out[i] = np.maximum(sh, priv)
    return out


def _rho_table(lam: float, q: int) -> np.ndarray:
    """cdf[r] = P(reg <= r) of `_rho_sample`'s distribution, r = 0..q+1 (the last entry is 1)."""
    cdf = np.exp(-lam * np.exp2(-np.arange(q + 2, dtype=np.float64)))
    cdf[q + 1] = 1.0
    return cdf


def registers_block_mt(seed: int, start: int, count: int, p: int, card: float = 5e6, group: int = 16, threads: int = 1, out=None) -> np.ndarray:
    """`registers_block` over several host threads (numpy's generators and searchsorted release the GIL); `out`, if given,
    is a uint8 [count, 2^p] array (e.g. a view of page-locked memory) that receives the rows."""
    from concurrent.futures import ThreadPoolExecutor
    m = 1 << p
    if out is None:
        out = np.empty((count, m), dtype=np.uint8)
    step = max(group, (count + max(threads, 1) * 4 - 1) // (max(threads, 1) * 4) // group * group)
    cuts = list(range(start - start % group, start + count, step))
    def work(a):
        lo, hi = max(a, start), min(a + step, start + count)
        if hi > lo:
            out[lo - start:hi - start] = registers_block(seed, lo, hi - lo, p, card, group)
    if threads <= 1 or len(cuts) <= 1:
        for a in cuts:
            work(a)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, cuts))
    return out


def registers_block(seed: int, start: int, count: int, p: int, card: float = 5e6, group: int = 16) -> np.ndarray:
    """Rows [start, start + count) of an endless seeded register matrix with the construction of `registers`
    (correlated groups of `group` sketches sharing an ancestor component).  Row i depends on (seed, i // group) only, so
    any shard can be produced on its own — every rank of the multi-process bench, and the reference arm, draw the very
    same bytes for the rows they need.  Inverse-CDF sampling from the register distribution's table (3 uniforms per
    register, no transcendental per element)."""
    m, q = 1 << p, 64 - p
    out = np.empty((count, m), dtype=np.uint8)
    lam = card / m
    full = _rho_table(lam, q)
    g0, g1 = start // group, (start + count + group - 1) // group
    for g in range(g0, g1):
        rng = np.random.default_rng([seed, g])
        shared = np.searchsorted(full, rng.random(m)).astype(np.uint8)
        for j in range(group):
            i = g * group + j
            frac = 1.0 - RATE_LADDER[i % len(RATE_LADDER)]
            # every member consumes the same number of uniforms whatever its rate: rows stay independent of the shard cut
            u1, u2 = rng.random(m), rng.random(m)
            if not (start <= i < start + count):
                continue
            if frac >= 1.0:
                row = shared
            else:
                sh = np.minimum(shared, np.searchsorted(_rho_table(frac * lam, q), u1).astype(np.uint8)) if frac > 0.0 else np.zeros(m, np.uint8)
                row = np.maximum(sh, np.searchsorted(_rho_table((1.0 - frac) * lam, q), u2).astype(np.uint8))
            out[i - start] = row
    return out


def adversarial_registers(seed: int, p: int) -> np.ndarray:
    """Edge-case sketches: empty, saturated, constant, full value range, single hot register."""
    rng = np.random.default_rng(seed)
    m, q = 1 << p, 64 - p
    rows = [
        np.zeros(m, np.uint8),                                   # empty sketch
        np.full(m, q + 1, np.uint8),                             # every register saturated (MLE -> inf)
        np.full(m, 7, np.uint8),                                 # constant
        rng.integers(0, q + 2, size=m).astype(np.uint8),         # uniform over the full range 0..q+1
        rng.integers(0, q + 2, size=m).astype(np.uint8),
        np.where(rng.random(m) < 0.01, 3, 0).astype(np.uint8),   # nearly empty (linear-counting regime)
        np.where(rng.random(m) < 0.5, q + 1, q).astype(np.uint8),  # top two bins only
    ]
    hot = np.zeros(m, np.uint8)
    hot[int(rng.integers(0, m))] = q + 1
    rows.append(hot)
    a = np.full(m, 10, np.uint8); a[0] = 0
    b = np.full(m, 10, np.uint8); b[1] = 0                       # union is all-10 although both have a 0
    rows += [a, b]
    return np.stack(rows)
