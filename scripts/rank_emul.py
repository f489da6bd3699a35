"""Emulates one rank of the N-GPU weak-scaling dist bench on a single GPU: n = round(10000 * sqrt(N)) sketches,
the rank's row range from multigpu.row_partition; prints prepare / kernel times and the plan's threshold count."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
from dashing_b200 import capi, multigpu

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
n, p = int(round(10000 * N ** 0.5)), 14
regs = B.synth_registers_torch(torch, n, p, 5, dev, card=5e6)
plan = capi.DistPlan(0)
prm = capi.dist_params(p, 31, 2, 2, capi.JI, 0)
def timeit(fn, steps=5, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
out = {"N": N, "n": n}
out["prepare_ms"] = timeit(lambda: plan.prepare_dev(regs.data_ptr(), n, p, 2, st))
for rank in sorted({0, N // 2, N - 1}):
    rb, re = multigpu.row_partition(n, N)[rank]
    tri = lambda r: (r * (2 * n - r - 1)) // 2
    d_out = torch.empty(tri(re) - tri(rb), dtype=torch.float32, device=dev)
    ms = timeit(lambda: plan.run_symmetric_dev(prm, rb, re, d_out.data_ptr(), st))
    out[f"rank{rank}"] = {"rows": [rb, re], "kernel_ms": ms, "pairs_per_s": (tri(re) - tri(rb)) / ms * 1e3, "info": plan.last_run_info()}
    del d_out
print(json.dumps(out))
