"""db200_sketch_batch end to end from page-locked / pageable host ASCII: the two upload routes and their mix (DB200_DEBUG_UPLOAD=1
prints how the chunks were split)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
from dashing_b200 import capi
ng, L, k, p = int(os.environ.get("NG", 1000)), 5_000_000, 31, 14
dev = torch.device("cuda:0")
a = B.synth_genomes_torch(torch, ng, L, 4242, dev)
pinned = capi.pinned_empty(ng * L); torch.from_numpy(pinned).copy_(a.cpu()); del a
offs = np.arange(ng + 1, dtype=np.uint64) * np.uint64(L); grb = np.arange(ng + 1, dtype=np.uint64)
out = {}
for label, env in (("ascii", {"DB200_HOST_PACK": "0"}), ("hybrid", {"DB200_HOST_PACK": "1"})):
    os.environ.update(env)
    for src_name, src in (("page_locked", pinned),) + ((("pageable", np.array(pinned)),) if os.environ.get("PAGEABLE") else ()):
        ref = capi.sketch_batch(src, offs, grb, k, p, True)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); r = capi.sketch_batch(src, offs, grb, k, p, True); ts.append((time.perf_counter() - t0) * 1e3)
        out[f"{label}_{src_name}_ms"] = round(min(ts), 2)
        out.setdefault("reg_sum", int(r.sum())); assert int(r.sum()) == out["reg_sum"]
print(json.dumps(out))
