"""GPU parity, hot path (ii): cardinalities and all-pairs values through the C ABI vs the oracle.
Floats: 1e-6 relative (BASELINE.json), policy in tests/parity.py."""
import os

import numpy as np
import pytest

from parity import assert_close, unstable_size
from dashing_b200 import synth

pytestmark = pytest.mark.gpu


def test_cardinalities_all_estimators(gpu, checker, golden_dir):
    d = np.load(os.path.join(golden_dir, "dist.npz"))
    for estim in (0, 1, 2):
        assert_close(gpu.cardinalities(d["regs"], 10, estim), d[f"card_e{estim}"], rtol=1e-9, what=f"golden card estim {estim}")
    for p in (7, 10, 14, 16, 18):
        regs = np.concatenate([synth.registers(p, 9, p, card=1e3 * (1 << (p // 2))), synth.adversarial_registers(2, p)])
        for estim in (0, 1, 2):
            assert_close(gpu.cardinalities(regs, p, estim), checker.cardinalities(regs, p, estim), rtol=1e-9, what=f"p={p} estim={estim}")


def test_pair_matrices_golden(gpu, golden_dir):
    d = np.load(os.path.join(golden_dir, "dist.npz"))
    p, k, regs = int(d["p"]), int(d["k"]), d["regs"]
    for estim in (0, 1, 2):
        card = d[f"card_e{estim}"]
        scale = float(np.max(card[np.isfinite(card)]))
        for jestim in (2, 3):
            for order in (0, 1):
                ign = unstable_size(d[f"pairs_e{estim}_j{jestim}_r2_o{order}"], scale)
                for rtype in range(9):
                    got = gpu.dist_symmetric(regs, p, k=k, estim=estim, jestim=jestim, result_type=rtype, order=order)
                    assert_close(got, d[f"pairs_e{estim}_j{jestim}_r{rtype}_o{order}"], scale=scale if rtype == 2 else 1.0, ignore=ign,
                                 what=f"estim={estim} jestim={jestim} rtype={rtype} order={order}")
    assert_close(gpu.dist_rect(regs[:20], regs[20:], p, k=k, result_type=1), d["rect_e2_j2_r1"], what="rect JI")
    assert_close(gpu.dist_rect(regs[:20], regs[20:], p, k=k, jestim=3, result_type=0), d["rect_e2_j3_r0"], what="rect JMLE mash",
                 ignore=unstable_size(gpu.dist_rect(regs[:20], regs[20:], p, k=k, jestim=3, result_type=2), 4e5))
    assert_close(gpu.dist_symmetric(d["regs14"], 14, k=k, result_type=1), d["pairs14_ji"], what="p14 JI")
    assert_close(gpu.dist_symmetric(d["regs14"], 14, k=k, result_type=0), d["pairs14_mash"], what="p14 Mash")
    assert_close(gpu.dist_symmetric(d["regs14"], 14, k=k, jestim=3, result_type=1), d["pairs14_jmle"], what="p14 JMLE")


def test_reference_bundled_genomes(gpu, golden_dir):
    kat = np.load(os.path.join(golden_dir, "kat.npz"))
    ji = gpu.dist_symmetric(kat["gcf_regs_p10"], 10, k=31, result_type=1)
    mash = gpu.dist_symmetric(kat["gcf_regs_p14"], 14, k=31, result_type=0)
    assert_close(ji, kat["gcf_ji_p10"]); assert_close(mash, kat["gcf_mash_p14"])
    assert "%.6g" % ji[5] == "0.550403"
    assert ["%.6g" % v for v in mash] == ["0.238361", "1", "1", "1", "0.152908", "0.0106825"]


@pytest.mark.parametrize("p", [7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20])
def test_all_pairs_vs_oracle(gpu, checker, p):
    n = 75 if p < 15 else (45 if p < 17 else 12)   # ragged: not a multiple of the 32-sketch panel
    adv = synth.adversarial_registers(4, p)
    if p >= 17:   # 32-bit count storage: the full-range rows (48 live thresholds) exceed shared memory for the joint kernel
        adv = adv[[0, 2, 5]]
    regs = np.concatenate([synth.registers(100 + p, n, p, card=40.0 * (1 << p)), adv])
    for estim, jestim, rtypes in ((2, 2, range(9)), (0, 2, (1, 2)), (1, 2, (0, 7)), (2, 3, (0, 1, 2, 5, 7)), (0, 3, (1,))):
        card = checker.cardinalities(regs, p, estim)
        scale = float(np.max(card[np.isfinite(card)]))
        ign = unstable_size(checker.dist_rows(regs, p, estim=estim, jestim=jestim, rtype=2), scale)
        for rtype in rtypes:
            for order in ((0, 1) if jestim == 3 else (0,)):
                got = gpu.dist_symmetric(regs, p, k=21, estim=estim, jestim=jestim, result_type=rtype, order=order)
                want = checker.dist_rows(regs, p, k=21, estim=estim, jestim=jestim, rtype=rtype, order=order)
                assert_close(got, want, scale=scale if rtype == 2 else 1.0, ignore=ign, what=f"p={p} estim={estim} jestim={jestim} rtype={rtype} order={order}")


def test_rect_and_row_sharding(gpu, checker):
    p = 12
    regs = synth.registers(31, 150, p, card=2e5)
    full = gpu.dist_symmetric(regs, p, result_type=0)
    # block-row shards (the multi-GPU decomposition) concatenate to the full packed triangle
    parts = [gpu.dist_symmetric(regs, p, result_type=0, row_begin=a, row_end=b) for a, b in ((0, 1), (1, 40), (40, 97), (97, 150))]
    np.testing.assert_array_equal(np.concatenate(parts), full)
    # rectangular mode == the corresponding entries of the symmetric matrix (union path is operand-symmetric)
    nr = 83
    rect = gpu.dist_rect(regs[:nr], regs[nr:], p, result_type=0)
    n = len(regs)
    idx = lambda i, j: i * (2 * n - i - 1) // 2 + j - i - 1
    want = np.array([[full[idx(j, nr + q)] for j in range(nr)] for q in range(n - nr)], dtype=np.float32)
    np.testing.assert_array_equal(rect, want)
    assert_close(rect, checker.dist_rect(regs[:nr], regs[nr:], p, rtype=0), what="rect vs oracle")


def test_properties_at_bench_scale(gpu, checker):
    """BASELINE dist config is 10,000 p=14 sketches; the oracle needs minutes for that, so at N=3000 check
    size-independent properties plus a random sample of pairs against the oracle."""
    p, n = 14, 3000
    regs = synth.registers(2026, n, p)
    regs[17] = regs[5]                                   # identical sketches -> JI == 1, Mash == 0
    regs[100] = np.maximum(regs[7], regs[9])             # union sketch
    ji = gpu.dist_symmetric(regs, p, result_type=1)
    mash = gpu.dist_symmetric(regs, p, result_type=0)
    idx = lambda i, j: i * (2 * n - i - 1) // 2 + j - i - 1
    assert ji.shape == (n * (n - 1) // 2,) and np.isfinite(ji).all() and (ji >= 0).all() and (ji <= 1.0 + 1e-6).all()
    assert ji[idx(5, 17)] == 1.0 and mash[idx(5, 17)] == 0.0
    # Mash is the documented function of JI (dist_index)
    ksinv = np.float64(np.float32(1.0 / 31))
    with np.errstate(divide="ignore"):
        want = np.where(ji > 0, -np.log(2.0 * ji.astype(np.float64) / (1.0 + ji)) * ksinv, 1.0)
    assert_close(mash, want, rtol=2e-6, what="mash = f(ji)")
    # containment of a sketch in the union that contains it: |A ∩ (A ∪ B)| = |A|  -> SIZES equals cardinality
    sizes = gpu.dist_symmetric(regs, p, result_type=2)
    card = gpu.cardinalities(regs, p)
    assert_close(sizes[idx(7, 100)], card[7], rtol=1e-6)
    # random sample against the oracle
    rng = np.random.default_rng(3)
    for _ in range(300):
        i, j = sorted(rng.choice(n, 2, replace=False))
        assert_close(ji[idx(i, j)], checker.pair(regs[i], regs[j], p, rtype=1), what=f"pair {i},{j}")
    # symmetric result == rect result on a slab
    rect = gpu.dist_rect(regs[:64], regs[2000:2040], p, result_type=1)
    np.testing.assert_array_equal(rect, np.array([[ji[idx(j, 2000 + q)] for j in range(64)] for q in range(40)], dtype=np.float32))


def test_unsupported_is_loud(gpu):
    regs = np.zeros((2, 1 << 21), np.uint8)
    with pytest.raises(gpu.Db200Error) as ei:
        gpu.dist_symmetric(regs, 21)
    assert ei.value.code == gpu.EUNSUPPORTED


def test_large_matrix_indexing(gpu, checker):
    """20,011 sketches (2.0e8 pairs, 800 MB of output): 64-bit distmat offsets, ragged last panel, row shards far into the
    triangle, all checked on samples against the oracle."""
    p, n = 10, 20011
    regs = synth.registers(424242, n, p, card=3e5, group=64)
    out = gpu.dist_symmetric(regs, p, k=31, result_type=0)
    assert out.size == n * (n - 1) // 2 and np.isfinite(out).all()
    idx = lambda i, j: i * (2 * n - i - 1) // 2 + j - i - 1
    rng = np.random.default_rng(9)
    samples = [(0, 1), (0, n - 1), (n - 2, n - 1), (31, 32), (19999, 20010), (12345, 20000)]
    samples += [tuple(sorted(map(int, rng.choice(n, 2, replace=False)))) for _ in range(150)]
    for i, j in samples:
        assert_close(out[idx(i, j)], checker.pair(regs[i], regs[j], p, rtype=0, k=31), what=f"pair {i},{j}")
    # a row shard deep in the triangle equals the corresponding slice of the full result
    rb, re_ = 17000, 17777
    shard = gpu.dist_symmetric(regs, p, k=31, result_type=0, row_begin=rb, row_end=re_)
    np.testing.assert_array_equal(shard, out[idx(rb, rb + 1): idx(re_, re_ + 1) if re_ < n - 1 else out.size])


def test_wide_counts_full_value_range(gpu, checker):
    """p = 17 with registers over the whole range 0..q+1: 48 live thresholds x 32-bit counts leave no room to stage the
    sparse tails, so the kernel sweeps every threshold densely — same results."""
    p = 17
    regs = np.concatenate([synth.registers(5, 6, p, card=40.0 * (1 << p)), synth.adversarial_registers(4, p)])
    got = gpu.dist_symmetric(regs, p, k=21, result_type=1)
    want = checker.dist_rows(regs, p, k=21, rtype=1)
    card = checker.cardinalities(regs, p, 2)
    ign = unstable_size(checker.dist_rows(regs, p, k=21, rtype=2), float(np.max(card[np.isfinite(card)])))
    assert_close(got, want, ignore=ign, what="p=17 full range")


@pytest.mark.parametrize("vmax,label", [(31, "6 stages"), (34, "5 stages, both tails"), (38, "4 stages, sparse tails only"), (45, "one CTA per SM")])
def test_live_threshold_count_selects_the_pipeline_shape(gpu, checker, vmax, label):
    """The number of live thresholds K = gmax - gmin decides how much shared memory is left for the TMA ring and the
    staged tails (plan_run): 2 CTAs/SM with 6, 5 or 4 stages, then 1 CTA/SM.  Same results in every shape — and at
    100,000 real sketches K is about 35, so these are not corner cases."""
    p = 14
    regs = synth.registers(23, 40, p, card=3e6, group=8)
    rng = np.random.default_rng(vmax)
    regs = np.minimum(regs, vmax).astype(np.uint8)
    # plant the extremes: a few empty registers (gmin = 0) and a few at vmax (gmax = vmax), plus a mid-range sprinkle
    for s in range(regs.shape[0]):
        idx = rng.choice(1 << p, size=40, replace=False)
        regs[s, idx[:5]] = 0
        regs[s, idx[5:8]] = vmax
        regs[s, idx[8:]] = rng.integers(1, vmax + 1, size=32)
    assert int(regs.min()) == 0 and int(regs.max()) == vmax
    for jestim, rtype in ((2, 1), (2, 0), (3, 1)):
        got = gpu.dist_symmetric(regs, p, k=21, jestim=jestim, result_type=rtype)
        want = checker.dist_rows(regs, p, k=21, jestim=jestim, rtype=rtype)
        card = checker.cardinalities(regs, p, 2)
        ign = unstable_size(checker.dist_rows(regs, p, k=21, jestim=jestim, rtype=2), float(np.max(card)))
        assert_close(got, want, ignore=ign, what=f"K={vmax} ({label}) j{jestim} r{rtype}")


def test_full_matrix_parity_at_bench_shape(gpu, checker):
    """Every one of the 4.5e6 pairs of 3,000 bench-shaped p=14 sketches against the reference, TRUE relative error (VERDICT r01
    weak #1/#2: a 1-in-1e6 flip of the secant iteration's trip count would show here; bench.py repeats the comparison on
    >= 1.8e7 pairs of the 10,000-sketch matrix it times)."""
    p, n = 14, 3000
    regs = synth.registers_block_mt(2026, 0, n, p, threads=8)
    for rtype in (1, 0):
        got = gpu.dist_symmetric(regs, p, k=31, result_type=rtype)
        want = checker.dist_rows(regs, p, k=31, rtype=rtype)
        ign = None
        if rtype == 0:
            sizes = checker.dist_rows(regs, p, k=31, rtype=2)
            ign = unstable_size(sizes, float(np.max(checker.cardinalities(regs[:64], p, 2))))
        assert_close(got, want, ignore=ign, what=f"full matrix n={n} p={p} rtype={rtype}")


def test_jmle_parity_at_c5_shape(gpu, checker):
    """Joint MLE at the shape of BASELINE configs[4] (p=16, k=21): 1,200 sketches = 7.2e5 pairs against the reference's
    ertl_joint, both operand orders (only n <= 49 was covered)."""
    p, n, k = 16, 1200, 21
    regs = synth.registers_block_mt(2027, 0, n, p, threads=8)
    for order in (0, 1):
        got = gpu.dist_symmetric(regs, p, k=k, jestim=3, result_type=1, order=order)
        want = checker.dist_rows(regs, p, k=k, jestim=3, rtype=1, order=order)
        assert_close(got, want, what=f"JMLE n={n} p={p} k={k} order={order}")


def test_joint_mle_threshold_limit_is_an_explicit_error(gpu):
    """the joint-MLE kernel keeps three count families of every live threshold in shared memory; with 32-bit counts
    (p > 16) that fits about 32 thresholds.  A matrix whose registers span the whole range at p=17 (48 live thresholds) must be
    declined loudly (DB200_EUNSUPPORTED, so that a host keeps the reference's code for it) — never answered wrongly.  The union
    path takes the same matrix (test_wide_counts_full_value_range)."""
    p = 17
    regs = synth.adversarial_registers(4, p)[[0, 3, 4]]          # empty sketch + two sketches uniform over 0..q+1
    assert int(regs.max()) - int(regs.min()) == 64 - p + 1
    with pytest.raises(gpu.Db200Error) as ei:
        gpu.dist_symmetric(regs, p, k=21, jestim=3, result_type=1)
    assert ei.value.code == gpu.EUNSUPPORTED and "live thresholds" in str(ei.value)
    assert np.isfinite(gpu.dist_symmetric(regs[1:], p, k=21, result_type=1)).all()
