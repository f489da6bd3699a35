#!/bin/bash
# CLI sketch / dist on 1000 x 5 Mbp FASTA files in /dev/shm: phase timings (pageable and page-locked arena) next to the
# reference's own drivers on the box's CPU quota; .hll payloads compared byte for byte.
set -u
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from dashing_b200 import synth
root = "/dev/shm/db200_cli"; os.makedirs(root + "/gpu", exist_ok=True); os.makedirs(root + "/ref", exist_ok=True)
rng = np.random.default_rng(1); L = 5_000_000; ng = int(os.environ.get("NG", "1000"))
anc = synth.genome(rng, L)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
with open(root + "/paths.txt", "w") as pf:
    for i in range(ng):
        g = anc.copy(); idx = rng.integers(0, L, 20000); g[idx] = acgt[rng.integers(0, 4, idx.size)]
        lines = np.full(L // 80 * 81, 10, dtype=np.uint8); lines.reshape(-1, 81)[:, :80] = g.reshape(-1, 80)
        n = f"{root}/g{i}.fa"
        with open(n, "wb") as f: f.write(f">g{i}\n".encode()); f.write(lines.tobytes())
        pf.write(n + "\n")
PY
CLI=dashing_b200/host/dashing_b200
R=/dev/shm/db200_cli
echo "--- sketch (run 1: cold)"; ( time $CLI sketch -k31 -S14 -p16 -P $R/gpu -F $R/paths.txt ) 2>&1 | grep real
echo "--- sketch (run 2, phases)"; ( time DB200_TIMING=1 $CLI sketch -k31 -S14 -p16 -P $R/gpu -F $R/paths.txt ) 2>&1 | tail -30
echo "--- sketch, page-locked arena"; ( time DB200_TIMING=1 DB200_PINNED_ARENA=1 $CLI sketch -k31 -S14 -p16 -P $R/gpu -F $R/paths.txt ) 2>&1 | tail -30
echo "--- dist from FASTA, Mash, binary"; ( time $CLI dist -k31 -S14 -p16 -M -b -o $R/s.txt -O $R/d.bin -F $R/paths.txt ) 2>&1 | grep real
echo "--- reference drivers"
python - <<'PY' 2>/dev/null
import os, sys, time, json, gzip
sys.path.insert(0, os.getcwd())
from oracle import oracle as O
R = O.ref(); root = "/dev/shm/db200_cli"
names = open(root + "/paths.txt").read().split()
t = time.perf_counter(); R.cli_sketch(names, k=31, p=14, nthreads=O.usable_cores(), prefix=root + "/ref"); a = time.perf_counter() - t
same = all(gzip.open(f"{root}/gpu/g{i}.fa.w.31.spacing.14.hll").read() == gzip.open(f"{root}/ref/g{i}.fa.w.31.spacing.14.hll").read() for i in range(len(names)))
print(json.dumps({"ref_sketch_s": a, "hll_identical": same, "cores": O.usable_cores(), "genomes": len(names)}))
PY
rm -rf /dev/shm/db200_cli
