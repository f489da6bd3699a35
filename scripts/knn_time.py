"""Nearest-neighbour mode next to the all-pairs mode on the bench workload (10,000 p=14 sketches, Mash distance):
device-resident timing with CUDA events, and the host-pointer call (upload + prepare + kernels + n x nn pairs back)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
from dashing_b200 import capi

dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
n, p, nn = 10000, 14, 10
regs = B.synth_registers_torch(torch, n, p, 5, dev, card=5e6)
plan = capi.DistPlan(0); plan.prepare_dev(regs.data_ptr(), n, p, 2, st)
prm = capi.dist_params(p, 31, 2, 2, capi.MASH_DIST, capi.ORDER_COL_FIRST)
d_nb = torch.empty((n, nn, 2), dtype=torch.float32, device=dev)
d_out = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device=dev)
def timeit(fn, steps=5, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
out = {"n": n, "p": p, "nn": nn}
out["all_pairs_ms"] = timeit(lambda: plan.run_symmetric_dev(prm, 0, n, d_out.data_ptr(), st))
out["knn_ms"] = timeit(lambda: plan.run_knn_dev(prm, 0, 0, nn, d_nb.data_ptr(), st))
for nn2 in (100, 1000):
    d2 = torch.empty((n, nn2, 2), dtype=torch.float32, device=dev)
    out[f"knn{nn2}_ms"] = timeit(lambda: plan.run_knn_dev(prm, 0, 0, nn2, d2.data_ptr(), st), steps=3, warm=1)
h = regs.cpu().numpy()
t0 = time.perf_counter(); capi.knn_symmetric(h, p, nn, result_type=capi.MASH_DIST); capi.knn_symmetric(h, p, nn, result_type=capi.MASH_DIST); out["knn_host_call_ms"] = (time.perf_counter() - t0) * 500
t0 = time.perf_counter(); capi.dist_symmetric(h, p, result_type=capi.MASH_DIST); capi.dist_symmetric(h, p, result_type=capi.MASH_DIST); out["all_pairs_host_call_ms"] = (time.perf_counter() - t0) * 500
print(json.dumps(out))
