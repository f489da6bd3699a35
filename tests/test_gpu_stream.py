"""GPU: row-block streaming all-pairs (db200_dist_symmetric_stream, SURVEY.md §8(b) S3 callback form) — every block handed to the
callback must be the same bits the one-shot call writes for those rows, rows arrive in order and exactly once, a failing
callback aborts the call; also through every (logical) device, where blocks arrive per device."""
import numpy as np
import pytest

from dashing_b200 import synth

pytestmark = pytest.mark.gpu


def collect(gpu, regs, p, **kw):
    n = regs.shape[0]
    tri = lambda r: r * (2 * n - r - 1) // 2
    out = np.full(tri(n), np.nan, dtype=np.float32)
    seen = []

    def on_rows(rb, re_, vals):
        assert vals.size == tri(re_) - tri(rb)
        out[tri(rb):tri(re_)] = vals
        seen.append((rb, re_))
    gpu.dist_symmetric_stream(regs, p, on_rows, **kw)
    return out, seen


@pytest.mark.parametrize("p,jestim,rtype", [(10, 2, 1), (12, 3, 0), (14, 2, 0)])
def test_stream_equals_one_shot(gpu, p, jestim, rtype):
    regs = np.concatenate([synth.registers(9 + p, 530, p, card=30.0 * (1 << p), group=8), synth.adversarial_registers(5, p)])
    n = regs.shape[0]
    want = gpu.dist_symmetric(regs, p, k=21, jestim=jestim, result_type=rtype)
    for block_pairs in (0, 5000, 40_000):
        got, seen = collect(gpu, regs, p, k=21, jestim=jestim, result_type=rtype, block_pairs=block_pairs)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (p, block_pairs)
        assert seen[0][0] == 0 and seen[-1][1] == n and all(a[1] == b[0] for a, b in zip(seen, seen[1:])), "rows in order, exactly once"
        if block_pairs:
            assert len(seen) > 3
    # a row range
    tri = lambda r: r * (2 * n - r - 1) // 2
    got, seen = collect(gpu, regs, p, k=21, jestim=jestim, result_type=rtype, block_pairs=7000, row_begin=100, row_end=400)
    assert seen[0][0] == 100 and seen[-1][1] == 400
    assert np.array_equal(got[tri(100):tri(400)].view(np.uint32), want[tri(100):tri(400)].view(np.uint32))
    assert np.isnan(got[:tri(100)]).all() and np.isnan(got[tri(400):]).all()


def test_stream_callback_failure_and_cached_cards(gpu):
    p = 10
    regs = synth.registers(4, 200, p, card=3e4, group=8)

    def boom(rb, re_, vals):
        raise ValueError("writer failed")
    with pytest.raises(ValueError):
        gpu.dist_symmetric_stream(regs, p, boom, block_pairs=3000)
    with pytest.raises(gpu.Db200Error):
        gpu.dist_symmetric_stream(regs, p, lambda rb, re_, v: 7, block_pairs=3000)
    # the library is usable afterwards, and the cached-cardinality override reaches the streaming form too
    card = gpu.cardinalities(regs, p) * 1.05
    want = gpu.dist_symmetric(regs, p, card=card)
    got, _ = collect(gpu, regs, p, block_pairs=3000, card=card)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_stream_all_devices(gpu, monkeypatch):
    if gpu.device_count() < 2:
        monkeypatch.setenv("DB200_VIRTUAL_DEVICES", "3")
    p = 10
    regs = synth.registers(8, 700, p, card=3e4, group=8)
    want = gpu.dist_symmetric(regs, p, k=21, result_type=gpu.MASH_DIST, device=0)
    got, seen = collect(gpu, regs, p, k=21, result_type=gpu.MASH_DIST, device=gpu.ALL_DEVICES, block_pairs=9000)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert sorted(seen)[0][0] == 0 and sorted(seen)[-1][1] == 700 and sum(b - a for a, b in seen) == 700


def test_rows_beyond_2_pow_32_pairs(gpu, checker):
    """BASELINE.json's C4 size: 100,000 sketches = 4.99995e9 pairs, so distmat offsets pass 2^32.  Rows deep in the triangle
    (first pair index 4.99e9) through the one-shot and the streaming row-range calls, against the rectangular call on the
    same sketches and the checker on samples."""
    p, n = 7, 100_000
    regs = synth.registers(31, n, p, card=4e3, group=32)
    rb, re_ = 95_000, 95_160
    tri = lambda r: r * (2 * n - r - 1) // 2
    assert tri(rb) > 2 ** 32
    rows = gpu.dist_symmetric(regs, p, k=21, result_type=gpu.MASH_DIST, row_begin=rb, row_end=re_)
    assert rows.size == tri(re_) - tri(rb)
    got = np.full(rows.size, np.nan, np.float32)

    def on_rows(b0, b1, vals):
        got[tri(b0) - tri(rb): tri(b1) - tri(rb)] = vals
    gpu.dist_symmetric_stream(regs, p, on_rows, k=21, result_type=gpu.MASH_DIST, row_begin=rb, row_end=re_, block_pairs=100_000)
    assert np.array_equal(got.view(np.uint32), rows.view(np.uint32))
    # row i of the triangle = (i, j > i): the rectangular call with queries = those rows, references = everything after rb
    rect = gpu.dist_rect(regs[rb:], regs[rb:re_], p, k=21, result_type=gpu.MASH_DIST)     # [q][j - rb]
    for i in (rb, rb + 77, re_ - 1):
        seg = rows[tri(i) - tri(rb): tri(i + 1) - tri(rb)]
        assert np.array_equal(seg.view(np.uint32), rect[i - rb, i - rb + 1:].view(np.uint32)), i
    from parity import assert_close
    for i, j in ((rb, rb + 1), (rb + 100, n - 1), (re_ - 1, 99_000)):
        assert_close(rows[tri(i) - tri(rb) + j - i - 1], checker.pair(regs[i], regs[j], p, rtype=0, k=21), what=f"pair {i},{j}")
