"""Side measurements for the other BASELINE configs (not the bench line): sketch at k=21/p=16, joint-MLE all-pairs,
p=10 and p=16 all-pairs.  Device-resident timing with CUDA events, 3 warm-ups."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
from dashing_b200 import capi

dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
def timeit(fn, steps=5, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
out = {}
# sketch variants
for (k, p, ng) in ((21, 16, 400), (31, 10, 400), (32, 14, 400), (15, 14, 400)):
    asc = B.synth_genomes_torch(torch, ng, 5_000_000, 7, dev)
    offs = np.arange(ng + 1, dtype=np.uint64) * np.uint64(5_000_000); grb = np.arange(ng + 1, dtype=np.uint64)
    pg = capi.PackedGenomes(int(asc.data_ptr()), offs, grb, k, device=0)
    regs = torch.empty((ng, 1 << p), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: pg.sketch_dev(p, True, regs.data_ptr(), st))
    out[f"sketch_k{k}_p{p}"] = {"kmers_per_s": pg.kmers / ms * 1e3, "ms": ms}
    pg.close(); del asc, regs; torch.cuda.empty_cache()
# dist variants
for (p, n, jestim, label) in ((14, 6000, 3, "jmle_p14"), (16, 4000, 3, "jmle_p16"), (16, 5000, 2, "mle_p16"), (10, 20000, 2, "mle_p10"), (12, 10000, 2, "mle_p12")):
    regs = B.synth_registers_torch(torch, n, p, 5, dev, card=5e6 if p >= 12 else 3e5)
    d_out = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device=dev)
    plan = capi.DistPlan(0); plan.prepare_dev(regs.data_ptr(), n, p, 2, st)
    prm = capi.dist_params(p, 21, 2, jestim, capi.JI, 0)
    ms = timeit(lambda: plan.run_symmetric_dev(prm, 0, n, d_out.data_ptr(), st))
    out["dist_" + label] = {"pairs_per_s": n * (n - 1) / 2 / ms * 1e3, "ms": ms, "n": n, "info": plan.last_run_info()}
    plan.close(); del regs, d_out; torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
