"""GPU nearest neighbours (`--nearest-neighbors`; perform_nns / nndist_loop, src/sketch_and_cmp.h:642-783) through the
C ABI, against the reference's own perform_nns outputs (tests/golden/knn.npz, generated with one thread = the
deterministic visiting order) and the live checker."""
import json
import os

import numpy as np
import pytest

from parity import assert_knn_close
from dashing_b200 import synth

pytestmark = pytest.mark.gpu


def _run(gpu, regs, p, nn, nq=0, **kw):
    if nq:
        return gpu.knn_rect(regs[:-nq], regs[-nq:], p, nn, **kw)
    return gpu.knn_symmetric(regs, p, nn, **kw)


def test_knn_golden(gpu, golden_dir):
    from oracle.make_golden import knn_inputs
    g = np.load(os.path.join(golden_dir, "knn.npz"))
    p, a, b = knn_inputs()
    for ci, (which, rtype, jestim, nn, nq) in enumerate(json.loads(str(g["cases"]))):
        regs = a if which == "a" else b
        got = _run(gpu, regs, p, nn, nq, k=int(g["k"]), jestim=jestim, result_type=rtype)
        assert_knn_close(got, g[f"case{ci}"], what=f"knn case {ci} {(which, rtype, jestim, nn, nq)}")


def test_knn_ties_follow_the_visiting_order(gpu, checker):
    """Duplicated sketches: exactly equal values, so the cut falls inside runs of ties (see the CPU twin in
    test_oracle_pinning.py).  Equal inputs give equal values on the GPU too, so the indices must match exactly."""
    p = 10
    base = synth.registers(5, 12, p, card=1e5, group=4)
    regs = base[np.array([0, 1, 0, 2, 1, 0, 3, 4, 4, 5, 0, 6, 7, 1, 8, 9, 10, 11, 4, 0])]
    for rtype in (0, 1, 2, 8):
        for nn in (1, 3, 7, 25):
            for nq in (0, 6):
                got = _run(gpu, regs, p, nn, nq, k=31, result_type=rtype)
                want = checker.knn(regs, p, nn, k=31, rtype=rtype, nq=nq, nthreads=1)
                assert_knn_close(got, want, what=f"knn ties r{rtype} nn{nn} nq{nq}")
                assert np.array_equal(got["index"], want["index"]), (rtype, nn, nq)


def _replay(vals_row, idx_row, nn, sim):
    """The reference's update rule on one row, visiting `idx_row` in the given (ascending) order."""
    fill = np.float32(-3.402823466e38 if sim else 3.402823466e38)
    S = [(fill, 0xFFFFFFFF)] * nn
    for v, j in zip(vals_row, idx_row):
        top = min(range(nn), key=lambda e: S[e]) if sim else max(range(nn), key=lambda e: S[e])
        if (v > S[top][0]) if sim else (v < S[top][0]):
            S[top] = (v, int(j))
    return sorted(S, reverse=bool(sim))


@pytest.mark.parametrize("p,jestim", [(10, 2), (12, 3)])
def test_knn_row_blocks_equal_full_matrix(gpu, p, jestim, monkeypatch):
    """Many row blocks (tiny DB200_KNN_BLOCK_PAIRS) must give exactly what the update rule gives on the full matrix of
    the same kernel's values (JI: ksinv plays no part)."""
    n, nn = 333, 9
    regs = synth.registers(77, n, p, card=4e4 if p == 10 else 2e5, group=16)
    monkeypatch.setenv("DB200_KNN_BLOCK_PAIRS", "20000")
    got = gpu.knn_symmetric(regs, p, nn, jestim=jestim, result_type=gpu.JI)
    monkeypatch.delenv("DB200_KNN_BLOCK_PAIRS")
    one = gpu.knn_symmetric(regs, p, nn, jestim=jestim, result_type=gpu.JI)
    assert np.array_equal(got, one)
    packed = gpu.dist_symmetric(regs, p, jestim=jestim, result_type=gpu.JI, order=gpu.ORDER_COL_FIRST)
    full = np.zeros((n, n), dtype=np.float32)
    iu = np.triu_indices(n, 1)
    full[iu] = packed
    full.T[iu] = packed
    for r in range(0, n, 7):
        others = np.array([j for j in range(n) if j != r])
        want = _replay(full[r, others], others, nn, sim=True)
        assert [(float(v), int(i)) for v, i in got[r]] == [(float(v), int(i)) for v, i in want], r
    # rect mode in query blocks: the last 70 sketches against the first 263
    nq = 70
    monkeypatch.setenv("DB200_KNN_BLOCK_PAIRS", "9000")
    gotq = gpu.knn_rect(regs[:-nq], regs[-nq:], p, nn, jestim=jestim, result_type=gpu.MASH_DIST, k=21)
    monkeypatch.delenv("DB200_KNN_BLOCK_PAIRS")
    oneq = gpu.knn_rect(regs[:-nq], regs[-nq:], p, nn, jestim=jestim, result_type=gpu.MASH_DIST, k=21)
    assert np.array_equal(gotq, oneq)


def test_knn_more_neighbours_than_sketches_and_limits(gpu):
    p = 10
    regs = synth.registers(3, 5, p, card=1e5, group=5)
    got = gpu.knn_symmetric(regs, p, 8, result_type=gpu.MASH_DIST)
    assert got.shape == (5, 8)
    # 4 real neighbours per row, then the reference's (FLT_MAX, uint32(-1)) filler (src/sketch_and_cmp.h:652-653)
    assert (got["index"][:, 4:] == 0xFFFFFFFF).all() and (got["value"][:, 4:] == np.float32(3.402823466e38)).all()
    assert (np.sort(got["index"][:, :4], axis=1) == np.array([[j for j in range(5) if j != i] for i in range(5)])).all()
    assert (np.diff(got["value"].astype(np.float64), axis=1) >= 0).all()
    sim = gpu.knn_symmetric(regs, p, 8, result_type=gpu.JI)
    assert (sim["value"][:, 4:] == np.float32(-3.402823466e38)).all() and (np.diff(sim["value"].astype(np.float64), axis=1) <= 0).all()
    one = gpu.knn_symmetric(regs[:1], p, 2)
    assert (one["index"] == 0xFFFFFFFF).all()
    with pytest.raises(gpu.Db200Error) as e:
        gpu.knn_symmetric(regs, p, 2000)
    assert e.value.code == gpu.EUNSUPPORTED


def test_knn_large(gpu):
    """n = 3000 at p = 12 (4.5e6 pairs): a sample of rows against the checker's values for those rows."""
    n, p, nn = 3000, 12, 20
    regs = synth.registers(99, n, p, card=3e5, group=32)
    got = gpu.knn_symmetric(regs, p, nn, result_type=gpu.MASH_DIST, k=31)
    assert (np.diff(got["value"].astype(np.float64), axis=1) >= 0).all()
    assert (got["index"] != np.arange(n)[:, None]).all()
    # rows of a full-matrix run of the same kernel (float ksinv there: compare neighbour SETS through JI instead)
    sim = gpu.knn_symmetric(regs, p, nn, result_type=gpu.JI)
    packed = gpu.dist_symmetric(regs, p, result_type=gpu.JI, order=gpu.ORDER_COL_FIRST)
    tri = lambda i: (i * (2 * n - i - 1)) // 2
    for r in (0, 1, 31, 32, 1500, 2998, 2999):
        row = np.empty(n, dtype=np.float32)
        for i in range(r):
            row[i] = packed[tri(i) + r - i - 1]
        row[r + 1:] = packed[tri(r):tri(r) + n - r - 1]
        others = np.array([j for j in range(n) if j != r])
        want = _replay(row[others], others, nn, sim=True)
        assert [(float(v), int(i)) for v, i in sim[r]] == [(float(v), int(i)) for v, i in want], r


def test_knn_partial_tables_merge_for_distance_measures(gpu):
    """db200_dist_plan_run_knn_rows_dev: partial tables over disjoint block rows (what each rank of the multi-process driver
    computes) merged by multigpu.merge_neighbor_tables equal the one-GPU table bit for bit — for a distance measure, with long
    runs of exactly-1 Mash distances (unrelated groups) at the cut.  The plan API takes device memory: torch is the plumbing."""
    import torch
    from dashing_b200 import multigpu
    p, nn = 10, 7
    regs = synth.registers(21, 300, p, card=2e4, group=8)
    n = regs.shape[0]
    dev = torch.device("cuda", 0)
    d_regs = torch.from_numpy(regs).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    plan = gpu.DistPlan(0)
    plan.prepare_dev(d_regs.data_ptr(), n, p, gpu.ERTL_MLE, stream)
    for rtype, jestim in ((gpu.MASH_DIST, gpu.ERTL_MLE), (gpu.FULL_MASH_DIST, gpu.ERTL_JOINT_MLE)):
        prm = gpu.dist_params(p, 21, gpu.ERTL_MLE, jestim, rtype, gpu.ORDER_COL_FIRST)
        want = gpu.knn_symmetric(regs, p, nn, k=21, jestim=jestim, result_type=rtype)
        for world in (2, 3, 7):
            tables = []
            for rb, re_ in multigpu.row_partition(n, world):
                d_out = torch.zeros(n * nn * 8, dtype=torch.uint8, device=dev)
                plan.run_knn_rows_dev(prm, rb, re_, nn, d_out.data_ptr(), stream)
                torch.cuda.synchronize()
                tables.append(d_out.cpu().numpy().view(gpu.NEIGHBOR_DTYPE).reshape(n, nn))
            got = multigpu.merge_neighbor_tables(np.stack(tables))
            assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["value"].view(np.uint32), want["value"].view(np.uint32)), (rtype, world)
    plan.close()
