"""CLI-level timing on the GPU box: `dashing_b200 sketch|dist` on synthetic FASTA files in /dev/shm, next to the
reference's own drivers (oracle/_ref: sketch_core / dist_sketch_and_cmp) with all usable host threads."""
import os, sys, time, json, shutil, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dashing_b200 import synth
from oracle import oracle as O

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
with_ref = (sys.argv[3] != "noref") if len(sys.argv) > 3 else True
root = "/dev/shm/db200_cli"
shutil.rmtree(root, ignore_errors=True); os.makedirs(root + "/gpu"); os.makedirs(root + "/ref")
rng = np.random.default_rng(1)
t0 = time.perf_counter()
anc = synth.genome(rng, L)
names = []
for i in range(ng):
    g = synth.mutate(rng, anc, 0.01 * (i % 16)) if i % 16 else anc
    lines = np.full((L + 79) // 80 * 81, ord("\n"), dtype=np.uint8)
    body = lines.reshape(-1, 81)
    pad = np.full(body.shape[0] * 80, ord("A"), dtype=np.uint8); pad[:L] = g
    body[:, :80] = pad.reshape(-1, 80)
    n = f"{root}/g{i}.fa"
    with open(n, "wb") as f:
        f.write(f">g{i}\n".encode()); f.write(lines.tobytes()[: L + L // 80])
        if (L % 80): f.write(b"\n")
    names.append(n)
gen_s = time.perf_counter() - t0
CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dashing_b200", "host", "dashing_b200")
cores = O.usable_cores()
out = {"genomes": ng, "bases_each": L, "gen_s": gen_s, "cores": cores}
def run(*a):
    t = time.perf_counter(); r = subprocess.run([CLI, *a], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return time.perf_counter() - t
open(root + "/paths.txt", "w").write("\n".join(names) + "\n")
run("sketch", "-k31", "-S14", f"-p{cores}", "-P", root + "/gpu", "-F", root + "/paths.txt")   # warm page cache / driver
out["gpu_sketch_s"] = run("sketch", "-k31", "-S14", f"-p{cores}", "-P", root + "/gpu", "-F", root + "/paths.txt")
out["gpu_sketch_kmers_per_s"] = ng * (L - 30) / out["gpu_sketch_s"]
out["gpu_dist_s"] = run("dist", "-k31", "-S14", f"-p{cores}", "-M", "-b", "-o", root + "/gs.txt", "-O", root + "/gd.bin", "-F", root + "/paths.txt")
out["gpu_dist_presketched_s"] = run("dist", "-k31", "-S14", f"-p{cores}", "-M", "-b", "--presketched", "-o", root + "/gs2.txt", "-O", root + "/gd2.bin",
                                    *[f"{root}/gpu/g{i}.fa.w.31.spacing.14.hll" for i in range(ng)])
if with_ref:
    R = O.ref()
    t = time.perf_counter(); R.cli_sketch(names, k=31, p=14, nthreads=cores, prefix=root + "/ref"); out["ref_sketch_s"] = time.perf_counter() - t
    out["ref_sketch_kmers_per_s"] = ng * (L - 30) / out["ref_sketch_s"]
    t = time.perf_counter(); R.cli_dist(names, root + "/rs.txt", root + "/rd.bin", k=31, p=14, rtype=0, emit_fmt=1, nthreads=cores); out["ref_dist_s"] = time.perf_counter() - t
    import gzip
    same = all(gzip.open(f"{root}/gpu/g{i}.fa.w.31.spacing.14.hll").read() == gzip.open(f"{root}/ref/g{i}.fa.w.31.spacing.14.hll").read() for i in range(ng))
    out["hll_identical"] = bool(same)
print(json.dumps(out)); open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "cli_time.json"), "w").write(json.dumps(out))
shutil.rmtree(root, ignore_errors=True)
