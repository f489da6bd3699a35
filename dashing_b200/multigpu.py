"""Multi-GPU driver: one process per GPU, torch.distributed for the plumbing.

The path shards naturally:
  * sketching — genomes are independent: rank r sketches genomes r, r+W, r+2W, ... (or a contiguous,
    size-balanced slice); no communication.
  * all-pairs — ONE exchange step: an all-gather of each rank's (n_r x 2^p) uint8 register block into
    the full n x 2^p matrix on every GPU (NCCL over NVLink; 100,000 x 16 KiB = 1.6 GB), then every
    rank computes a block-row range of the packed upper triangle, balanced by pair count, with no
    further communication.  Rows are contiguous in distmat order, so the host just concatenates.

The compute callables are injected so the plumbing can be exercised on CPU with the gloo backend
(tests/test_multigpu_gloo.py); in production they are the C-ABI `_dev` entry points (DistPlan).
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import numpy as np


def tri_offset(n: int, row: int) -> int:
    """distmat offset of the first entry of `row` (distmat/distmat.h:273-276)."""
    return row * (2 * n - row - 1) // 2


def row_partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Split rows [0, n) into `world` contiguous ranges holding (nearly) equal numbers of pairs.
    Row i owns n-1-i pairs, so boundaries follow the inverse of the cumulative pair count."""
    total = n * (n - 1) // 2
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        # smallest row b with tri_offset(n, b) >= target
        b = int(n - 0.5 - math.sqrt(max((n - 0.5) ** 2 - 2.0 * target, 0.0)))
        b = max(min(b, n), bounds[-1])
        while b < n and tri_offset(n, b) < target:
            b += 1
        while b > bounds[-1] and tri_offset(n, b - 1) >= target:
            b -= 1
        bounds.append(b)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def genome_partition(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Greedy size-balanced assignment of genomes to ranks (largest first), deterministic."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(x) for x in out]


def shard_counts(n: int, world: int) -> List[int]:
    """Contiguous shard sizes of n sketches over `world` ranks (first n % world ranks get one more)."""
    return [n // world + (1 if r < n % world else 0) for r in range(world)]


def allgather_registers(local, counts: Sequence[int], dist, out=None):
    """All-gather ragged register blocks.  local: torch uint8 [counts[rank], m] on the rank's device.
    Returns the full [sum(counts), m] matrix (same on every rank).  One collective: shards are padded
    to the largest shard so that all_gather_into_tensor (NCCL: a single ring/NVLS all-gather) applies."""
    import torch

    world = dist.get_world_size()
    m = local.shape[1]
    cmax = max(counts)
    if local.shape[0] != cmax:
        pad = torch.zeros((cmax, m), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        local = pad
    gathered = torch.empty((world * cmax, m), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local.contiguous())
    if all(c == cmax for c in counts):
        return gathered
    n = sum(counts)
    if out is None:
        out = torch.empty((n, m), dtype=local.dtype, device=local.device)
    off = 0
    for r, c in enumerate(counts):
        out[off:off + c] = gathered[r * cmax: r * cmax + c]
        off += c
    return out


def allgather_prepare_overlapped(plan, local, counts: Sequence[int], dist, p: int, estim: int, stream: int = 0):
    """The exchange step and the plane build, overlapped: the global register range first (two scalars, one all-reduce), then
    one broadcast per shard; every shard's threshold planes / counts / cardinalities are built (plan.add_rows_dev) as soon as that
    shard has landed, while the later shards are still in flight on NCCL's stream.  Returns the full [n, 2^p] matrix.
    `plan` needs begin_dev / add_rows_dev / finish_dev (capi.DistPlan; the gloo test passes a recorder)."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    m = local.shape[1]
    n = int(sum(counts))
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + int(c))
    if local.shape[0]:
        mm = torch.stack([local.amin().to(torch.int32), -(local.amax().to(torch.int32))])
    else:
        mm = torch.tensor([255, 0], dtype=torch.int32, device=local.device)
    dist.all_reduce(mm, op=dist.ReduceOp.MIN)
    gmin, gmax = int(mm[0].item()), -int(mm[1].item())          # (synchronises: the plan's shapes depend on the range)
    full = torch.empty((n, m), dtype=local.dtype, device=local.device)
    full[offs[rank]:offs[rank + 1]] = local
    plan.begin_dev(n, p, estim, gmin, gmax, stream)
    works = [dist.broadcast(full[offs[r]:offs[r + 1]], src=r, async_op=True) if counts[r] else None for r in range(world)]
    for r in range(world):
        if works[r] is not None:
            works[r].wait()           # NCCL: the current stream waits for that broadcast; the host does not block
            plan.add_rows_dev(full.data_ptr(), offs[r], int(counts[r]), stream)
    plan.finish_dev()
    return full


def dist_symmetric_sharded(local_regs, counts: Sequence[int], dist, compute_rows: Callable, gather_out: bool = False):
    """All-gather + block-row computation.
    compute_rows(full_regs, n, row_begin, row_end) -> 1-D float32 tensor/array with that row range.
    Returns (row_range, local_rows) and, if gather_out, the full packed matrix on every rank (tests)."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    full = allgather_registers(local_regs, counts, dist)
    n = int(sum(counts))
    rb, re_ = row_partition(n, world)[rank]
    rows = compute_rows(full, n, rb, re_)
    if not gather_out:
        return (rb, re_), rows
    rows_t = torch.as_tensor(np.asarray(rows.cpu() if hasattr(rows, "cpu") else rows), dtype=torch.float32)
    sizes = [tri_offset(n, b) - tri_offset(n, a) for a, b in row_partition(n, world)]
    smax = max(sizes + [1])
    buf = torch.zeros(smax, dtype=torch.float32)
    buf[: rows_t.numel()] = rows_t
    outs = [torch.zeros(smax, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(outs, buf)  # host-side concatenation (gloo in tests)
    return (rb, re_), torch.cat([o[:s] for o, s in zip(outs, sizes)])


# ---- nearest neighbours across ranks (multi-process form) ---------------------------------------------
NEIGHBOR_DTYPE = np.dtype([("value", np.float32), ("index", np.uint32)])   # validx_t, src/sketch_and_cmp.h:605
DIST_MEASURES = (0, 3, 4, 6, 8)   # MASH_DIST, FULL_MASH_DIST, FULL_CONTAINMENT_DIST, CONTAINMENT_DIST, SYMMETRIC_CONTAINMENT_DIST


def merge_neighbor_tables(tables: np.ndarray) -> np.ndarray:
    """tables: [world][n][nn] partial nearest-neighbour tables over DISJOINT row ranges of the triangle (structured
    (value, index) entries, unused slots filled with (FLT_MAX, 0xFFFFFFFF)).  -> [n][nn], per row the nn smallest
    (value, index) keys of the union, ascending.

    Exact for the distance measures only: there the reference's heap (perform_nns, src/sketch_and_cmp.h:642-697: visit in
    ascending index, replace the worst when STRICTLY better) ends with precisely the nn smallest (value, index) pairs, a
    property of the multiset that survives any split of the sources.  For similarity measures the retained set depends on
    the visiting order and no merge of partial tables reproduces it."""
    tables = np.asarray(tables)
    world, n, nn = tables.shape
    cat = np.ascontiguousarray(np.transpose(tables, (1, 0, 2))).reshape(n, world * nn)
    # -0.0 and +0.0 are the same value for the reference's comparisons: the index decides between them
    order = np.lexsort((cat["index"], cat["value"] + np.float32(0.0)), axis=-1)[:, :nn]
    return np.take_along_axis(cat, order, axis=-1)


def knn_symmetric_sharded(local_regs, counts: Sequence[int], dist, compute_partial: Callable, result_type: int, nneighbors: int):
    """All-gather of the register shards, one partial table per rank over its block rows, all-gather of the tables, merge.
    compute_partial(full_regs, n, row_begin, row_end, nneighbors) -> structured array [n][nneighbors] (production:
    DistPlan.run_knn_rows_dev).  Distance measures only (see merge_neighbor_tables)."""
    import torch

    if result_type not in DIST_MEASURES:
        raise ValueError("nearest-neighbour tables of similarity measures do not merge across ranks: compute them on one GPU")
    rank, world = dist.get_rank(), dist.get_world_size()
    full = allgather_registers(local_regs, counts, dist)
    n = int(sum(counts))
    rb, re_ = row_partition(n, world)[rank]
    part = np.ascontiguousarray(compute_partial(full, n, rb, re_, nneighbors))
    assert part.dtype == NEIGHBOR_DTYPE and part.shape == (n, nneighbors)
    mine = torch.from_numpy(part.view(np.uint8).reshape(-1).copy())
    if getattr(local_regs, "is_cuda", False):      # NCCL moves device tensors; gloo (CPU tests) host tensors
        mine = mine.to(local_regs.device)
    outs = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)          # n * nn * 8 bytes per rank
    tables = np.stack([o.cpu().numpy().view(NEIGHBOR_DTYPE).reshape(n, nneighbors) for o in outs])
    return merge_neighbor_tables(tables)
