#!/bin/bash
# CLI all-pairs on 10,000 p=14 sketches (made from 10,000 x 60 kbp FASTA files in /dev/shm): sketch, then dist --presketched
# with binary / TSV output, one device and --device all; copy-thread sweep for the pageable upload path.
set -u
R=/dev/shm/db200_cli5; mkdir -p $R/fa $R/sk
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from dashing_b200 import synth
root = "/dev/shm/db200_cli5"; rng = np.random.default_rng(3); L = 60_000; ng = 10_000
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
anc = [synth.genome(rng, L) for _ in range(100)]
with open(root + "/paths.txt", "w") as pf:
    for i in range(ng):
        g = anc[i % 100].copy(); idx = rng.integers(0, L, 600 * (1 + i // 100 % 7)); g[idx] = acgt[rng.integers(0, 4, idx.size)]
        lines = np.full(L // 80 * 81, 10, dtype=np.uint8); lines.reshape(-1, 81)[:, :80] = g.reshape(-1, 80)
        n = f"{root}/fa/g{i}.fa"
        with open(n, "wb") as f: f.write(f">g{i}\n".encode()); f.write(lines.tobytes())
        pf.write(n + "\n")
PY
CLI=dashing_b200/host/dashing_b200
echo "--- sketch 10,000 x 60 kbp, p=14"; ( time $CLI sketch -k31 -S14 -p16 -P $R/sk -F $R/paths.txt ) 2>&1 | grep real
ls $R/sk | sed "s#^#$R/sk/#" > $R/hll.txt; wc -l $R/hll.txt
echo "--- dist --presketched, binary"; ( time DB200_TIMING=1 $CLI dist --presketched -k31 -S14 -p16 -M -b -o $R/s.txt -O $R/d.bin -F $R/hll.txt ) 2>&1 | tail -8
echo "--- dist --presketched, TSV"; ( time $CLI dist --presketched -k31 -S14 -p16 -M -o $R/s.txt -O $R/d.tsv -F $R/hll.txt ) 2>&1 | grep real; ls -la $R/d.bin $R/d.tsv
echo "--- dist --presketched, 10 nearest neighbours"; ( time $CLI dist --presketched -k31 -S14 -p16 -M --nearest-neighbors 10 -o $R/s.txt -O $R/nn.tsv -F $R/hll.txt ) 2>&1 | grep real
echo "--- copy threads (sketch of 400 x 5 Mbp, steady-state batch time)"
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from dashing_b200 import synth
root = "/dev/shm/db200_cli5"; rng = np.random.default_rng(1); L = 5_000_000; ng = 400
anc = synth.genome(rng, L); acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
with open(root + "/big.txt", "w") as pf:
    for i in range(ng):
        g = anc.copy(); idx = rng.integers(0, L, 20000); g[idx] = acgt[rng.integers(0, 4, idx.size)]
        lines = np.full(L // 80 * 81, 10, dtype=np.uint8); lines.reshape(-1, 81)[:, :80] = g.reshape(-1, 80)
        n = f"{root}/b{i}.fa"
        with open(n, "wb") as f: f.write(f">g{i}\n".encode()); f.write(lines.tobytes())
        pf.write(n + "\n")
PY
mkdir -p $R/sk2
for T in 4 8 16; do echo "copy threads $T"; DB200_COPY_THREADS=$T DB200_TIMING=1 $CLI sketch -k31 -S14 -p16 -P $R/sk2 -F $R/big.txt 2>&1 | grep "previous batch done" | tail -1; done
rm -rf $R
