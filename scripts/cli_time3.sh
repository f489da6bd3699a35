#!/bin/bash
# CLI tests + phase timings of the CLI on 1000 x 5 Mbp FASTA files, next to the reference driver
set -u
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_sketch.py -x -q -m gpu 2>&1 | tail -4
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from dashing_b200 import synth
root = "/dev/shm/db200_cli"; os.makedirs(root + "/gpu", exist_ok=True); os.makedirs(root + "/ref", exist_ok=True)
rng = np.random.default_rng(1); L = 5_000_000; ng = 1000
anc = synth.genome(rng, L)
with open(root + "/paths.txt", "w") as pf:
    for i in range(ng):
        g = anc.copy(); idx = rng.integers(0, L, 20000); g[idx] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, idx.size)]
        lines = np.full(L // 80 * 81, 10, dtype=np.uint8); lines.reshape(-1, 81)[:, :80] = g.reshape(-1, 80)
        n = f"{root}/g{i}.fa"
        with open(n, "wb") as f: f.write(f">g{i}\n".encode()); f.write(lines.tobytes())
        pf.write(n + "\n")
PY
CLI=dashing_b200/host/dashing_b200
echo "--- sketch 1000 genomes (run 1, then timed run 2)"
$CLI sketch -k31 -S14 -p16 -P /dev/shm/db200_cli/gpu -F /dev/shm/db200_cli/paths.txt
( time DB200_TIMING=1 $CLI sketch -k31 -S14 -p16 -P /dev/shm/db200_cli/gpu -F /dev/shm/db200_cli/paths.txt ) 2>&1 | tail -40
echo "--- dist 1000 genomes from FASTA, Mash, binary"
( time $CLI dist -k31 -S14 -p16 -M -b -o /dev/shm/db200_cli/s.txt -O /dev/shm/db200_cli/d.bin -F /dev/shm/db200_cli/paths.txt ) 2>&1 | tail -4
echo "--- reference drivers (16 threads)"
python - <<'PY' 2>/dev/null
import os, sys, time, json
sys.path.insert(0, os.getcwd())
from oracle import oracle as O
R = O.ref(); root = "/dev/shm/db200_cli"
names = open(root + "/paths.txt").read().split()
t = time.perf_counter(); R.cli_sketch(names, k=31, p=14, nthreads=O.usable_cores(), prefix=root + "/ref"); a = time.perf_counter() - t
t = time.perf_counter(); R.cli_dist(names, root + "/rs.txt", root + "/rd.bin", k=31, p=14, rtype=0, emit_fmt=1, nthreads=O.usable_cores()); b = time.perf_counter() - t
import gzip
same = all(gzip.open(f"{root}/gpu/g{i}.fa.w.31.spacing.14.hll").read() == gzip.open(f"{root}/ref/g{i}.fa.w.31.spacing.14.hll").read() for i in range(len(names)))
print(json.dumps({"ref_sketch_s": a, "ref_dist_s": b, "hll_identical": same, "cores": O.usable_cores()}))
PY
rm -rf /dev/shm/db200_cli
