"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL): sharded sketching + all-gather + block-row
all-pairs must equal the single-GPU result bit for bit.  Used by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dashing_b200 import capi, multigpu, synth  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    k, p, ng = 21, 12, 37
    genomes = synth.genomes(99, ng, 60_000, group=8)
    # sketch: contiguous shards of the genome list (equal sizes -> shard_counts), no communication
    counts = multigpu.shard_counts(ng, world)
    start = sum(counts[:rank])
    mine = genomes[start:start + counts[rank]]
    local = torch.from_numpy(capi.sketch_genomes(mine, k, p, device=lr)).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    plan = capi.DistPlan(lr)
    prm = capi.dist_params(p, k, result_type=capi.MASH_DIST)

    def compute_rows(full, n, rb, re_):
        out = torch.empty(max(multigpu.tri_offset(n, re_) - multigpu.tri_offset(n, rb), 1), dtype=torch.float32, device=dev)
        # (the plan is rebuilt here through the overlapped exchange: it must hold exactly what prepare_dev builds)
        full2 = multigpu.allgather_prepare_overlapped(plan, local, counts, dist, p, capi.ERTL_MLE, stream)
        assert torch.equal(full2, full)
        plan.run_symmetric_dev(prm, rb, re_, out.data_ptr(), stream)
        torch.cuda.synchronize()
        return out[: multigpu.tri_offset(n, re_) - multigpu.tri_offset(n, rb)].cpu()

    (rb, re_), rows = multigpu.dist_symmetric_sharded(local, counts, dist, compute_rows)
    # gather row blocks on rank 0 over NCCL (padded)
    sizes = [multigpu.tri_offset(ng, b) - multigpu.tri_offset(ng, a) for a, b in multigpu.row_partition(ng, world)]
    buf = torch.zeros(max(sizes), dtype=torch.float32, device=dev)
    buf[: rows.numel()] = rows.to(dev)
    outs = [torch.zeros(max(sizes), dtype=torch.float32, device=dev) for _ in range(world)]
    dist.all_gather(outs, buf)
    ok = True
    if rank == 0:
        full = np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)])
        regs = capi.sketch_genomes(genomes, k, p, device=lr)
        want = capi.dist_symmetric(regs, p, k=k, result_type=capi.MASH_DIST, device=lr)
        ok = bool(np.array_equal(full, want))
        print(f"MGPU_CHECK world={world} pairs={want.size} equal={ok}", flush=True)
    # nearest neighbours, distance measure: per-rank partial tables over the rank's block rows, all-gathered and merged
    nn = 5
    regs_all = capi.sketch_genomes(genomes, k, p, device=lr)
    prm_nn = capi.dist_params(p, k, result_type=capi.MASH_DIST, order=capi.ORDER_COL_FIRST)

    def compute_partial(full, n, rb, re_, k_nn):
        d_out = torch.zeros(n * k_nn * 8, dtype=torch.uint8, device=dev)
        plan.prepare_dev(full.data_ptr(), n, p, capi.ERTL_MLE, stream)
        plan.run_knn_rows_dev(prm_nn, rb, re_, k_nn, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
        return d_out.cpu().numpy().view(capi.NEIGHBOR_DTYPE).reshape(n, k_nn)

    merged = multigpu.knn_symmetric_sharded(local, counts, dist, compute_partial, capi.MASH_DIST, nn)
    if rank == 0:
        want_nn = capi.knn_symmetric(regs_all, p, nn, k=k, result_type=capi.MASH_DIST, device=lr)
        ok_nn = bool(np.array_equal(merged["index"], want_nn["index"]) and np.array_equal(merged["value"], want_nn["value"]))
        print(f"MGPU_CHECK knn world={world} equal={ok_nn}", flush=True)
        ok = ok and ok_nn
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
