"""One joint-MLE all-pairs launch (for ncu): p and n from argv."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
from dashing_b200 import capi
p = int(sys.argv[1]); n = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
regs = B.synth_registers_torch(torch, n, p, 5, dev, card=5e6)
d_out = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device=dev)
plan = capi.DistPlan(0); plan.prepare_dev(regs.data_ptr(), n, p, 2, st)
prm = capi.dist_params(p, 21, 2, 3, capi.JI, 1)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
plan.run_symmetric_dev(prm, 0, n, d_out.data_ptr(), st)
e0.record()
for _ in range(reps): plan.run_symmetric_dev(prm, 0, n, d_out.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"p": p, "n": n, "ms": ms, "pairs_per_s": n * (n - 1) / 2 / ms * 1e3, "info": plan.last_run_info()}))
