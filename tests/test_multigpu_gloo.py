"""N>1 host-side logic on CPU (gloo, world_size 2): register all-gather (ragged shards), block-row partition of the
packed triangle, concatenation.  The per-rank compute is injected: here the oracle stands in for the GPU kernels
(tests may use it), in production it is DistPlan.run_symmetric_dev."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dashing_b200 import multigpu, synth


def test_row_partition_balanced_and_exact():
    for n in (2, 3, 17, 1000, 10_000, 28_284, 100_000):
        for world in (1, 2, 3, 4, 8):
            parts = multigpu.row_partition(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            pairs = [multigpu.tri_offset(n, e) - multigpu.tri_offset(n, b) for b, e in parts]
            assert sum(pairs) == n * (n - 1) // 2
            if n >= 1000:
                assert max(pairs) - min(pairs) <= 2 * n, (n, world, pairs)   # boundaries are whole rows: within ~2 rows of perfect balance


def test_genome_partition_and_shards():
    sizes = [5_000_000] * 10 + [100_000, 3_000_000_000, 7]
    parts = multigpu.genome_partition(sizes, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) == 3_000_000_000   # the giant genome sits alone
    assert multigpu.shard_counts(10, 4) == [3, 3, 2, 2]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, p, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        chk = O.port()
        regs = synth.registers(seed, n, p, card=1e5)
        counts = multigpu.shard_counts(n, world)
        start = sum(counts[:rank])
        local = torch.from_numpy(regs[start:start + counts[rank]].copy())

        def compute_rows(full, nn, rb, re_):
            tri = lambda r: r * (2 * nn - r - 1) // 2
            out = chk.dist_rows(full.numpy(), p, k=31, rtype=0, row_begin=rb, row_end=re_)
            return out[tri(rb):tri(re_)]

        (rb, re_), full_out = multigpu.dist_symmetric_sharded(local, counts, dist, compute_rows, gather_out=True)
        gathered = multigpu.allgather_registers(local, counts, dist)
        ok_regs = bool((gathered.numpy() == regs).all())
        if rank == 0:
            want = chk.dist_rows(regs, p, k=31, rtype=0)
            q.put((ok_regs, bool(np.array_equal(full_out.numpy(), want)), (rb, re_)))
        else:
            q.put((ok_regs, True, (rb, re_)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 64])
def test_sharded_all_pairs_gloo_world2(n):
    world, p = 2, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, p, 5, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(r[0] for r in res), "all-gathered register matrix differs from the source"
    assert all(r[1] for r in res), "concatenated block-rows differ from the single-process matrix"
    assert sorted(r[2] for r in res) == multigpu.row_partition(n, world)


def _partial_knn(full_dist, n, rb, re_, nn):
    """Reference semantics restricted to the pairs (i, j > i) with i in [rb, re): what one rank contributes."""
    out = np.zeros((n, nn), dtype=multigpu.NEIGHBOR_DTYPE)
    out["value"], out["index"] = np.float32(3.4028234663852886e38), np.uint32(0xFFFFFFFF)
    cand = [[] for _ in range(n)]
    for i in range(rb, re_):
        for j in range(i + 1, n):
            v = full_dist[i, j]
            cand[i].append((v, j)); cand[j].append((v, i))
    for r in range(n):
        best = sorted(cand[r])[:nn]
        for s, (v, j) in enumerate(best):
            out[r, s] = (v, j)
    return out


def test_merge_neighbor_tables_equals_the_sequential_heap(port):
    """Distance measures: per-rank partial tables over disjoint block rows, merged, equal the reference's one-thread heap on the
    whole set — including long runs of equal values (unrelated sketches: Mash distance exactly 1)."""
    p, nn = 10, 6
    regs = synth.registers(21, 48, p, card=2e4, group=8)
    n = regs.shape[0]
    want = port.knn(regs, p, nn, k=21, rtype=0)
    packed = port.dist_rows(regs, p, k=21, rtype=0, order=1)
    full = np.zeros((n, n), dtype=np.float32)
    iu = np.triu_indices(n, 1)
    full[iu] = packed
    for world in (2, 3, 5):
        parts = multigpu.row_partition(n, world)
        tables = np.stack([_partial_knn(full, n, rb, re_, nn) for rb, re_ in parts])
        got = multigpu.merge_neighbor_tables(tables)
        # (the matrix path rounds ksinv = 1/k to float, nndist_loop keeps it in double: values agree to the last bits only)
        assert np.array_equal(got["index"], want["index"]) and np.allclose(got["value"], want["value"], rtol=1e-6, atol=0), world
    with pytest.raises(ValueError):
        multigpu.knn_symmetric_sharded(None, [n], None, None, result_type=1, nneighbors=nn)


def _knn_worker(rank, world, port_no, n, p, nn, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        chk = O.port()
        regs = synth.registers(33, n, p, card=2e4, group=8)
        counts = multigpu.shard_counts(n, world)
        start = sum(counts[:rank])
        local = torch.from_numpy(regs[start:start + counts[rank]].copy())

        def compute_partial(full, nt, rb, re_, k_nn):
            packed = chk.dist_rows(full.numpy(), p, k=21, rtype=0, order=1)
            fd = np.zeros((nt, nt), dtype=np.float32)
            fd[np.triu_indices(nt, 1)] = packed
            return _partial_knn(fd, nt, rb, re_, k_nn)

        got = multigpu.knn_symmetric_sharded(local, counts, dist, compute_partial, result_type=0, nneighbors=nn)
        want = chk.knn(regs, p, nn, k=21, rtype=0)
        q.put(bool(np.array_equal(got["index"], want["index"]) and np.allclose(got["value"], want["value"], rtol=1e-6, atol=0)))
    finally:
        dist.destroy_process_group()


def test_sharded_knn_gloo_world2():
    world, n, p, nn = 2, 40, 10, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_knn_worker, args=(r, world, port_no, n, p, nn, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(res), "merged per-rank neighbour tables differ from the single-process table"


class _PlanRecorder:
    """Stands in for capi.DistPlan in allgather_prepare_overlapped: records what the driver hands to the C ABI."""

    def __init__(self):
        self.calls = []

    def begin_dev(self, n, p, estim, reg_min, reg_max, stream=0):
        self.calls.append(("begin", n, p, estim, reg_min, reg_max))

    def add_rows_dev(self, d_regs, row_begin, nrows, stream=0):
        self.calls.append(("add", row_begin, nrows))

    def finish_dev(self):
        self.calls.append(("finish",))


def _overlap_worker(rank, world, port, n, p, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        regs = synth.registers(9, n, p, card=1e5)
        regs[n - 1, 5] = 0                      # the global minimum lives on the last rank, the maximum on the first
        regs[0, 7] = 40
        counts = multigpu.shard_counts(n, world)
        start = sum(counts[:rank])
        local = torch.from_numpy(regs[start:start + counts[rank]].copy())
        plan = _PlanRecorder()
        full = multigpu.allgather_prepare_overlapped(plan, local, counts, dist, p, 2)
        q.put((rank, bool((full.numpy() == regs).all()), plan.calls, int(regs.min()), int(regs.max())))
    finally:
        dist.destroy_process_group()


def test_overlapped_exchange_and_plane_build_gloo_world2():
    """The chunked exchange (global register range by all-reduce, one broadcast per shard, planes built per shard as it lands):
    every rank ends with the full matrix, begins the plan with the GLOBAL range and adds every row exactly once."""
    world, n, p = 2, 37, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, n, p, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    counts = multigpu.shard_counts(n, world)
    for rank, ok, calls, gmin, gmax in res:
        assert ok
        assert calls[0] == ("begin", n, p, 2, gmin, gmax) and calls[-1] == ("finish",)
        adds = [c for c in calls if c[0] == "add"]
        assert adds == [("add", sum(counts[:r]), counts[r]) for r in range(world)]
