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
