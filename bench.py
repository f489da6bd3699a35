#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload dist|sketch|both]

Primary line (default): pairwise HLL comparisons/s, all-pairs on p=14 sketches (configs[2]: 10,000 sketches
on one B200).  A "step" is one pass of the hot path over one batch: (N>1: NCCL all-gather of the register
shards) -> threshold-plane build + per-sketch cardinalities -> tiled all-pairs kernel.  `value` times it with
the register matrix resident in HBM; `e2e` times the reference-facing C-ABI call with HOST buffers (H2D of
the registers and D2H of the float matrix inside the timed region).
The same JSON line carries
  * `parity`  — the GPU values of THIS run against the reference's values for the rows its `cpu_baseline` leg
                computes (N=1) / a reference sample of rank 0's rows (N>1), true relative error (tests/parity.py policy);
  * `sketch`  — the second half of the metric (k-mers hashed/s, configs[1]: 1,000 x 5 Mbp genomes, k=31, p=14);
  * `jmle`    — (N=1) Ertl joint MLE all-pairs at p=16, k=21 (the estimator of configs[4]) with its own parity object;
  * `c4`,`c5` — (N=8, or one emulated rank with --emulate-world 8) BASELINE configs[3] (100,000 p=14 sketches over
                8 ranks) and configs[4] (50,000 x 5 Mbp genomes, k=21, p=16, joint MLE, sketch + all-gather + all pairs).
                At N=8 these two legs run in child processes (`--legs-child`, own NCCL group, at most --legs-timeout seconds)
                so that nothing they do can cost the primary line.

Multi-GPU (torchrun, one rank per GPU): weak scaling.  dist: n(N) = round(10000*sqrt(N)) sketches, each rank
holds n/N of them, one all-gather, block-rows balanced by pair count; sketch: 1,000 genomes per rank.

Both arms draw the register matrix from dashing_b200.synth.registers_block (row i depends on (seed, i // 16) only),
so the GPU ranks and the reference arm consume identical bytes and print identical `config` objects.

`--impl reference` times the reference's own CPU implementation (oracle/_ref: the unmodified dashing headers
compiled with g++; falls back to the pinned C port if that library did not travel) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_DIST, K_MER, P_SKETCH = 14, 31, 14
N_DIST_1GPU = 10_000
N_GENOMES, GENOME_LEN = 1000, 5_000_000
SEED_DIST, SEED_JMLE, SEED_C4, SEED_C5 = 2026, 2027, 2028, 2029
N_JMLE, P_JMLE, K_JMLE = 4000, 16, 21
C4_N, C5_GENOMES, C5_P, C5_K, C45_WORLD = 100_000, 50_000, 16, 21, 8
RTOL, RESIDUE = 1e-6, 1e-9          # parity policy of tests/parity.py


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        rows = [l for (t, l) in self.samples if t0 <= t <= t1] or [l for (_, l) in self.samples[-3:]]
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); smax = max(smax, float(f[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def host_registers(seed, start, count, p, out=None):
    """Rows [start, start+count) of the seeded register matrix both arms use (host numpy, all usable cores)."""
    from dashing_b200 import synth
    return synth.registers_block_mt(seed, start, count, p, threads=usable_cores(), out=out)


def synth_registers_torch(torch, n, p, seed, device, card=5e6, group=16):
    """Device-side register synthesis (same construction as dashing_b200.synth.registers, torch RNG).  Only used to stand in
    for the OTHER ranks' sketches when one rank of an 8-GPU configuration is emulated on a single GPU."""
    from dashing_b200.synth import RATE_LADDER
    g = torch.Generator(device=device); g.manual_seed(seed)
    m, q = 1 << p, 64 - p

    def rho(lam_col, rows):
        # register of a bucket that received Poisson(lam) items: P(reg <= r) = exp(-lam 2^-r)
        u = torch.rand((rows, m), generator=g, device=device, dtype=torch.float64).clamp_min(1e-300)
        t = (-torch.log(u) / lam_col.clamp_min(1e-30)).clamp_min(2.0 ** -(q + 2))
        return torch.ceil(-torch.log2(t)).clamp(0, q + 1).to(torch.uint8)

    out = torch.empty((n, m), dtype=torch.uint8, device=device)
    ladder = torch.tensor(RATE_LADDER, dtype=torch.float64, device=device)
    one = torch.ones((1, 1), dtype=torch.float64, device=device)
    for s in range(0, n, group):
        rows = min(group, n - s)
        shared = rho(one * (card / m), 1).expand(rows, m)
        frac = (1.0 - ladder[(torch.arange(s, s + rows, device=device)) % len(RATE_LADDER)]).view(-1, 1)
        thin = torch.where(frac >= 1.0, shared, torch.minimum(shared, rho(frac * (card / m), rows)))
        priv = rho((1.0 - frac) * (card / m), rows)
        out[s:s + rows] = torch.maximum(thin, priv)
    return out


def synth_genomes_torch(torch, n, length, seed, device, group=50, first=0, out=None):
    """n single-record genomes of `length` ASCII bases on the device: groups share an ancestor, members are
    mutated copies (substitution-rate ladder of SURVEY.md §8(d)).  Returns a uint8 [n*length] tensor.
    Genome `first + i` belongs to ancestor group (first + i) // group, seeded by (seed, group index)."""
    from dashing_b200.synth import RATE_LADDER
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    if out is None:
        out = torch.empty(n * length, dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    anc, anc_group = None, -1
    for i in range(n):
        gi = first + i
        if gi // group != anc_group:
            anc_group = gi // group
            ga = torch.Generator(device=device); ga.manual_seed(seed * 1_000_003 + anc_group)
            anc = torch.randint(0, 4, (length,), generator=ga, device=device, dtype=torch.uint8)
        rate = RATE_LADDER[gi % len(RATE_LADDER)]
        if rate <= 0:
            code = anc
        else:
            g.manual_seed(seed * 7_000_003 + gi)
            hit = torch.rand(length, generator=g, device=device) < rate
            shift = torch.randint(1, 4, (length,), generator=g, device=device, dtype=torch.uint8)
            code = torch.where(hit, (anc + shift) & 3, anc)
        out[i * length:(i + 1) * length] = lut[code.long()]
    return out


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def load_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(kind, n, world):
    """ncu dram bytes per launch of the dominant kernel — only when the capture summarised under profiles/traffic.json was
    taken on THIS problem (same sketch / genome count, one GPU); otherwise null (a constant from another size is not a
    measurement of this run)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kind)
        if isinstance(d, dict) and world == 1 and int(d.get("n", -1)) == int(n):
            return d.get("bytes")
    except Exception:
        pass
    return None


def usable_cores():
    from oracle import oracle as O   # shared with the tests: affinity capped by the cgroup CPU quota
    return O.usable_cores()


def dist_n_for(world):
    return int(round(N_DIST_1GPU * math.sqrt(world)))


def tri(n, r):
    return r * (2 * n - r - 1) // 2


def rows_for_pairs(n, row0, target_pairs):
    """Smallest R such that rows [row0, row0+R) hold at least target_pairs pairs (or all remaining rows)."""
    r = row0
    while r < n - 1 and tri(n, r) - tri(n, row0) < target_pairs:
        r += 1
    return max(r - row0, 1)


def dist_config(n, world, p=P_DIST, k=K_MER, seed=SEED_DIST, what="ERTL_MLE union JI"):
    """The `config` object of the dist workload — printed verbatim by BOTH arms."""
    m = 1 << p
    return {"workload": f"dist all-pairs {n} p={p} sketches ({n * (n - 1) // 2} pairs), {what}", "n_sketches": n, "p": p, "k": k,
            "estimator": "ERTL_MLE", "result": "JI",
            "generator": f"dashing_b200.synth.registers_block(seed={seed}, card=5e6, group=16), rows [0,{n}) — identical bytes in both arms",
            "parallelism": f"block-row x{world}" + (" + 1 NCCL all-gather" if world > 1 else ""),
            "l2": "inputs (register matrix %d MB + threshold planes) larger than the 126 MB L2" % (n * m >> 20)}


def sketch_config(world):
    return {"workload": f"sketch {N_GENOMES * world} x {GENOME_LEN} bp synthetic genomes, k={K_MER}, p={P_SKETCH}, canonical",
            "genomes_per_gpu": N_GENOMES, "k": K_MER, "p": P_SKETCH, "parallelism": f"genomes x{world} (no collective)",
            "l2": "packed input 1788 MB per GPU, larger than the 126 MB L2"}


def parity_stats(got, want, scale=1.0, ignore=None, what=""):
    """tests/parity.py policy: true relative error |got-want|/|want|; only where |want| < 1e-9*scale (a cancellation residue
    in the reference itself) is the error measured against `scale`."""
    got = np.asarray(got, dtype=np.float64).ravel()
    want = np.asarray(want, dtype=np.float64).ravel()
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    with np.errstate(invalid="ignore", divide="ignore"):
        den = np.where(np.abs(want) >= RESIDUE * scale, np.abs(want), scale)
        err = np.abs(got - want) / den
    err = np.where(same, 0.0, err)
    err = np.where(np.isnan(err), np.inf, err)
    n_ign = 0
    if ignore is not None:
        ignore = np.asarray(ignore).ravel()
        n_ign = int(ignore.sum())
        err = np.where(ignore, 0.0, err)
    bad = np.nonzero(err > RTOL)[0]
    out = {"pairs_compared": int(got.size - n_ign), "max_rel_err": float(err.max()) if err.size else 0.0, "n_over_1e-6": int(bad.size),
           "n_ignored": n_ign, "n_bit_identical": int(same.sum()), "policy": "true relative; floor only where |want| < 1e-9 (tests/parity.py)"}
    if what:
        out["against"] = what
    if bad.size:
        out["first_bad"] = [{"idx": int(i), "got": float(got[i]), "want": float(want[i])} for i in bad[:5]]
    return out


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------
def reference_checker():
    from oracle import oracle as O
    if O.ref_available():
        r = O.ref()
        return r, "reference", f"oracle/_ref/{r.libname} (unmodified dashing headers, g++ -O3, SIMD tier {r.simd_tier()})"
    return O.port(), "port", "oracle/oracle_port.c (scalar C restatement; oracle/_ref did not travel)"


def ref_threads(kind):
    # torchrun exports OMP_NUM_THREADS=1: the thread count is passed to the reference's loops explicitly (num_threads clause)
    return usable_cores() if kind == "reference" else 1


def cpu_dist_sample(chk, kind, regs_np, n, target_s, threads, p=P_DIST, k=K_MER, jestim=2, rtype=1, row0=0, min_pairs=0):
    """Time rows [row0, row0+R) of the SAME all-pairs workload on the host cores (the perform_core_op loop: OpenMP dynamic
    over one matrix row at a time).  The n sketches are constructed and report()ed once, outside the timed region,
    as the reference does before its pair loop.  Returns (pairs/s, pairs, rows, seconds, values of the last pass)."""
    if kind == "reference":
        hs = chk.set_create(regs_np, p, 2, jestim)
        run = lambda r: chk.set_dist_rows(hs, k, rtype, 0, row0, row0 + r, threads)
    else:
        hs = None
        run = lambda r: chk.dist_rows(regs_np, p, k=k, jestim=jestim, rtype=rtype, row_begin=row0, row_end=row0 + r)[tri(n, row0):tri(n, row0 + r)]
    run(2)  # warm caches / thread pool
    rows = max(8, rows_for_pairs(n, row0, min_pairs) if min_pairs else 8)
    while True:
        t0 = time.perf_counter(); vals = run(rows); dt = time.perf_counter() - t0
        if dt >= 0.5 * target_s or row0 + rows >= n - 1:
            break
        rows = min(n - 1 - row0, int(rows * min(8.0, max(2.0, target_s / max(dt, 1e-3)))))
    if hs is not None:
        chk.set_free(hs)
    pairs = tri(n, row0 + rows) - tri(n, row0)
    return pairs / dt, pairs, rows, dt, np.asarray(vals)[:pairs]


def cpu_sketch_sample(chk, kind, genomes_np, length, threads, target_s=6.0, k=K_MER, p=P_SKETCH):
    """Repeated passes over a block of in-memory genomes (Encoder::for_each + hll_t::addh, one genome per OpenMP task)
    until ~target_s of CPU work has been timed."""
    ng = genomes_np.size // length
    offs = (np.arange(ng + 1, dtype=np.uint64) * np.uint64(length))
    grb = np.arange(ng + 1, dtype=np.uint64)

    def one_pass():
        if kind == "reference":
            chk.sketch_many(genomes_np, offs, grb, k, p, True, threads)
        else:
            for gi in range(ng):
                chk.sketch([genomes_np[gi * length:(gi + 1) * length].tobytes()], k, p, True)
    one_pass()
    passes, t0 = 0, time.perf_counter()
    while True:
        one_pass(); passes += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or passes >= 200:
            break
    kmers = passes * ng * (length - k + 1)
    return kmers / dt, kmers, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    chk, kind, desc = reference_checker()
    threads = ref_threads(kind)
    from dashing_b200 import synth
    n = dist_n_for(args.gpus)
    if args.workload == "sketch":
        ng = max(16, min(2 * threads, 128))
        gen = np.concatenate(synth.genomes(1234, ng, GENOME_LEN, group=16))
        vals = []
        for it in range(args.warmup + args.steps):
            v, kmers, dt = cpu_sketch_sample(chk, kind, gen, GENOME_LEN, threads)
            if it >= args.warmup:
                vals.append((v, dt))
        value = float(np.mean([v for v, _ in vals])); ms = float(np.mean([d for _, d in vals])) * 1e3
        sample = f"{kmers} k-mers: repeated passes over {ng} of the {N_GENOMES * args.gpus} genomes ({GENOME_LEN} bp each), in-memory Encoder::for_each + hll_t::addh, {threads} threads"
        metric, unit, config = "k-mers hashed/s (sketch k=31 p=14)", "kmers/s", sketch_config(args.gpus)
    else:
        t0 = time.perf_counter()
        regs = host_registers(SEED_DIST, 0, n, P_DIST)
        log(f"[bench/reference] {n} x 2^{P_DIST} registers generated in {time.perf_counter() - t0:.1f}s; {threads} threads")
        vals = []
        for it in range(args.warmup + args.steps):
            v, pairs, rows, dt, _ = cpu_dist_sample(chk, kind, regs, n, 6.0, threads)
            if it >= args.warmup:
                vals.append((v, dt, rows, pairs))
        value = float(np.mean([v[0] for v in vals])); ms = float(np.mean([v[1] for v in vals])) * 1e3
        sample = f"rows [0,{vals[-1][2]}) = {vals[-1][3]} of the {n * (n - 1) // 2} pairs of the same matrix, perform_core_op loop (OpenMP dynamic), {threads} threads"
        metric, unit, config = "pairwise HLL cmp/s (dist p=14)", "pairs/s", dist_config(n, args.gpus)
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 registers / u32 counts / f64 estimator -> f32 out",
            "data": "synthetic", "config": config, "reference": desc,
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_result(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def attach_collectives(cx):
    """barrier / max_over_ranks / all_ok on cx, over cx.torch, cx.dist, cx.dev (NCCL on the GPU box; the CPU tier drives the same
    functions over gloo with a stubbed torch.cuda, tests/test_bench_legs_gloo.py)."""
    torch, dist, world, dev = cx.torch, cx.dist, cx.world, cx.dev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(ok):
        """True only if every rank says so — the gate in front of every collective of the optional legs, so that a rank
        that failed (caught exception) takes the others out of the leg with it instead of leaving them in a collective."""
        if world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    cx.barrier, cx.max_over_ranks, cx.all_ok = barrier, max_over_ranks, all_ok


def run_legs(cx, legs):
    """BASELINE configs[3] and [4] (world == C45_WORLD, or one emulated rank): every rank calls this; a leg counts only if it went
    through on EVERY rank (a rank that dropped out leaves the others' numbers meaningless)."""
    args, rank = cx.args, cx.rank
    for name, fn in (("c4", bench_c4), ("c5", bench_c5)):
        if args.only and name not in args.only.split(","):
            continue
        t0 = time.perf_counter()
        cx.torch.cuda.set_device(cx.local_rank)        # (library calls take a device index and leave that device current)
        try:
            res_leg = fn(cx)
        except Exception as e:
            import traceback
            log(f"[bench] rank {rank}: {name} leg failed: {e!r}\n{traceback.format_exc()}")
            res_leg = {"failed": repr(e)}
        ok_all = True
        try:
            ok_all = cx.all_ok("failed" not in res_leg)
        except Exception as e:
            log(f"[bench] rank {rank}: status exchange after {name} failed: {e!r}")
        if not ok_all and "failed" not in res_leg:
            res_leg = {"failed": "another rank failed in this leg; see stderr"}
        legs[name] = res_leg
        log(f"[bench] rank {rank}: {name} done in {time.perf_counter() - t0:.1f}s")
        cx.torch.cuda.empty_cache()


def legs_child_port():
    """Rendezvous port of the children's process group: away from the parent's."""
    p = int(os.environ.get("MASTER_PORT", "29500"))
    return p + 101 if p + 101 < 65536 else p - 101


def run_legs_isolated(cx, timeout, cmd=None):
    """Every rank starts `bench.py --legs-child` (same RANK / WORLD_SIZE / LOCAL_RANK, own rendezvous port), waits for it at most
    `timeout` seconds and kills it otherwise; rank 0's child leaves the legs' results in a file.  -> {"c4": ..., "c5": ...} on rank 0
    (failure notes when the file is missing), {} elsewhere.  `cmd` replaces the child command in the CPU-tier test."""
    import subprocess
    import tempfile
    args, rank = cx.args, cx.rank
    names = [n for n in ("c4", "c5") if not (args.only and n not in args.only.split(","))]
    out_path = os.path.join(tempfile.gettempdir(), f"db200_legs_{os.getpid()}_{rank}.json")
    if os.path.exists(out_path):
        os.remove(out_path)
    env = dict(os.environ)
    env["DB200_LEGS_OUT"] = out_path
    env["DB200_LEGS_PORT"] = str(legs_child_port())
    if cmd is None:
        cmd = [sys.executable, os.path.abspath(__file__), "--gpus", str(cx.world), "--legs-child"]
        if args.only:
            cmd += ["--only", args.only]
        if args.no_cpu_baseline:
            cmd += ["--no-cpu-baseline"]
    t0 = time.perf_counter()
    why = None
    try:
        # the child's stdout must never reach this process's stdout (the ONE line of the contract): send it to stderr
        r = subprocess.run(cmd, env=env, timeout=timeout, stdout=sys.stderr, stderr=sys.stderr)
        if r.returncode != 0:
            why = f"child exited with status {r.returncode}"
    except subprocess.TimeoutExpired:
        why = f"child did not finish within {timeout:.0f} s and was killed"
    except Exception as e:        # noqa: BLE001 - the legs are optional, the primary line is not
        why = f"could not run the child: {e!r}"
    log(f"[bench] rank {rank}: legs child finished in {time.perf_counter() - t0:.1f}s" + (f" ({why})" if why else ""))
    if rank != 0:
        return {}
    legs = {}
    try:
        if os.path.exists(out_path):
            legs = json.load(open(out_path))
            os.remove(out_path)
    except Exception as e:        # noqa: BLE001
        why = why or f"unreadable result file: {e!r}"
    for n in names:
        if n not in legs:
            legs[n] = {"failed": (why or "the child left no result for this leg") + "; see stderr"}
    return legs


def run_legs_child(args, backend="nccl", stubs=None):
    """`bench.py --legs-child` (started by run_legs_isolated, one per rank): its own process group, the two legs, rank 0 writes the
    results to $DB200_LEGS_OUT.  `backend` / `stubs` = (torch-like, capi-like, device) let the CPU tier run the very same function
    over gloo with the device stubbed (tests/test_bench_legs_gloo.py)."""
    import datetime
    import torch
    import torch.distributed as dist
    from dashing_b200 import multigpu
    cx = Ctx()
    cx.world = int(os.environ["WORLD_SIZE"]); cx.rank = int(os.environ["RANK"]); cx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if stubs is None:
        from dashing_b200 import capi
        torch.cuda.set_device(cx.local_rank)
        cx.torch, cx.capi, cx.dev = torch, capi, torch.device("cuda", cx.local_rank)
        pg_kw = {"device_id": cx.dev}
        cx.stream = torch.cuda.current_stream().cuda_stream
    else:
        cx.torch, cx.capi, cx.dev = stubs
        pg_kw = {}
        cx.stream = 0
    cx.dist, cx.multigpu, cx.args = dist, multigpu, args
    os.environ.setdefault("DB200_PACK_THREADS", str(max(2, usable_cores() // cx.world)))
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{os.environ['DB200_LEGS_PORT']}", rank=cx.rank, world_size=cx.world,
                            timeout=datetime.timedelta(seconds=240), **pg_kw)
    cx.peak, cx.peak_src = load_peaks()
    cx.sampler = None
    attach_collectives(cx)
    legs = {}
    run_legs(cx, legs)
    if cx.rank == 0:
        tmp = os.environ["DB200_LEGS_OUT"] + ".tmp"
        with open(tmp, "w") as f:
            json.dump(legs, f)
        os.replace(tmp, os.environ["DB200_LEGS_OUT"])
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:        # noqa: BLE001
        log(f"[bench/legs-child] rank {cx.rank}: shutdown: {e!r}")
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from dashing_b200 import capi, multigpu

    cx = Ctx()
    cx.torch, cx.dist, cx.capi, cx.multigpu, cx.args = torch, dist, capi, multigpu, args
    cx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = rank = int(os.environ.get("RANK", "0"))
    cx.local_rank = local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if capi.device_count() < 1:
        raise RuntimeError("bench.py: libdashing_b200 sees no CUDA device (there is no CPU fallback)")
    if world > 1:
        # the ranks of one node share its host cores: each rank's library gets its share for the host-side packer
        os.environ.setdefault("DB200_PACK_THREADS", str(max(2, usable_cores() // world)))
    torch.cuda.set_device(local_rank)
    cx.dev = dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cx.stream = torch.cuda.current_stream().cuda_stream
    cx.peak, cx.peak_src = load_peaks()

    attach_collectives(cx)
    cx.sampler = ClockSampler(local_rank) if rank == 0 else None
    results, extra = {}, {}

    want = ("dist", "sketch") if args.workload == "both" else (args.workload,)
    emu = args.emulate_world > 1
    if not emu:
        for w in want:
            t0 = time.perf_counter()
            results[w], extra[w] = bench_dist(cx) if w == "dist" else bench_sketch(cx)
            log(f"[bench] {w}: {results[w]['value']:.4g} {results[w]['unit']} ({time.perf_counter() - t0:.1f}s incl. setup)")
            torch.cuda.empty_cache()

    # ---- CPU baseline + parity (rank 0): the reference's own loop on the host cores, bounded sample, values kept for parity
    if rank == 0 and not args.no_cpu_baseline and not emu:
        try:
            chk, kind, desc = reference_checker()
            threads = ref_threads(kind)
            if "dist" in results:
                n = results["dist"]["config"]["n_sketches"]
                ex = extra["dist"]
                # N=1: ~12 s of reference work = the cpu_baseline sample; N>1: a short sample of rank 0's first rows, parity only
                target_s, min_pairs = (12.0, 0) if world == 1 else (2.0, 1_000_000)
                v, pairs, rows, dt, vals = cpu_dist_sample(chk, kind, ex["regs_np"], n, target_s, threads, min_pairs=min_pairs)
                if world == 1:
                    results["dist"]["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind,
                                                       "sample": f"rows [0,{rows}) = {pairs} of the {n * (n - 1) // 2} pairs of the same matrix in {dt:.1f}s; {desc}"}
                results["dist"]["parity"] = parity_stats(ex["gpu_rows_np"][:pairs], vals, what=f"{kind}: rows [0,{rows}) of the same matrix ({desc})")
            if "sketch" in results and world == 1:
                gen = extra["sketch"]
                ngs = min(gen.size // GENOME_LEN, max(16, min(2 * threads, 128)))
                v, kmers, dt = cpu_sketch_sample(chk, kind, np.ascontiguousarray(gen[: ngs * GENOME_LEN]), GENOME_LEN, threads)
                results["sketch"]["cpu_baseline"] = {"value": v, "unit": "kmers/s", "cores": threads, "kind": kind,
                                                     "sample": f"{kmers} k-mers in {dt:.1f}s: repeated passes over {ngs} of the {N_GENOMES} genomes, in-memory Encoder::for_each + addh; {desc}"}
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU numbers
            log(f"[bench] cpu_baseline / parity failed: {e!r}")
    extra.clear()

    # ---- optional legs: joint MLE at the C5 shape (N=1), BASELINE configs[3] and [4] (N=8 or one emulated rank)
    legs = {}

    def build_line():
        if results:
            primary = "dist" if "dist" in results else "sketch"
            r = results[primary]
            line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": r["dtype"],
                    "data": "synthetic", "config": r["config"], "details": r.get("details"), "clocks": r["clocks"], "e2e": r["e2e"],
                    "gpu_launches": r["gpu_launches"], "roofline": r["roofline"]}
            for key in ("cpu_baseline", "parity"):
                if key in r:
                    line[key] = r[key]
            for other in results:
                if other != primary:
                    o = results[other]
                    line[other] = {k: o[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "details", "roofline", "e2e", "gpu_launches", "clocks",
                                                     "cpu_baseline", "parity") if k in o}
        else:
            line = {"metric": "emulated rank of an 8-GPU configuration (see c4 / c5)", "value": None, "n_gpus": 1, "emulated": {"world": args.emulate_world, "rank": args.emulate_rank}}
        line.update(legs)
        return line

    if world == 1 and not emu and args.workload == "both" and not args.no_extra:
        try:
            legs["jmle"] = bench_jmle(cx)
        except Exception as e:
            log(f"[bench] jmle leg failed: {e!r}")
    if emu and not args.no_extra:
        run_legs(cx, legs)                       # one emulated rank, in this process
    elif world == C45_WORLD and not args.no_extra:
        # The c4 / c5 legs run collectives of their own.  In the first 8-GPU run of round 2 a rank failed inside one of them, the
        # others sat in a collective until NCCL's watchdog killed the job 10 minutes later, and the primary result died with it.
        # They therefore run in CHILD processes (one per rank, their own process group on another port): whatever happens in
        # there — exception, hang, NCCL abort, crash — this process survives, waits at most --legs-timeout and prints its line.
        torch.cuda.empty_cache()
        legs.update(run_legs_isolated(cx, args.legs_timeout))
    if cx.sampler:
        cx.sampler.stop()
    if rank == 0:
        emit_result(build_line())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------- dist (primary)
def bench_dist(cx):
    torch, dist, capi, multigpu, args = cx.torch, cx.dist, cx.capi, cx.multigpu, cx.args
    world, rank, dev, stream = cx.world, cx.rank, cx.dev, cx.stream
    p, n = P_DIST, dist_n_for(world)
    m = 1 << p
    counts = multigpu.shard_counts(n, world)
    start = sum(counts[:rank])
    # every rank draws ITS rows of the seeded matrix (stands in for "rank r sketched these genomes") into page-locked memory
    host_regs = capi.pinned_empty(counts[rank] * m)
    t0 = time.perf_counter()
    host_registers(SEED_DIST, start, counts[rank], p, out=host_regs.reshape(counts[rank], m))
    log(f"[bench] rank {rank}: rows [{start},{start + counts[rank]}) of the register matrix generated in {time.perf_counter() - t0:.1f}s")
    pin_t = torch.from_numpy(host_regs).view(counts[rank], m)
    local = pin_t.to(dev)
    rb, re_ = multigpu.row_partition(n, world)[rank]
    my_pairs = tri(n, re_) - tri(n, rb)
    total_pairs = n * (n - 1) // 2
    d_out = torch.empty(max(my_pairs, 1), dtype=torch.float32, device=dev)
    plan = capi.DistPlan(cx.local_rank)
    prm = capi.dist_params(p, K_MER, capi.ERTL_MLE, capi.ERTL_MLE, capi.JI, capi.ORDER_ROW_FIRST)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step():
        ev[0].record()
        # ONE collective.  (multigpu.allgather_prepare_overlapped — global range by all-reduce, one broadcast per shard, planes built per
        # shard as it lands — was measured instead: 1.47 ms at N=2 against 0.46 + 0.7 ms here, and 8 ms at N=8: once the plane build
        # takes 0.35 ms per 10,000 sketches there is nothing left to hide behind eight broadcasts.)
        full = multigpu.allgather_registers(local, counts, dist) if world > 1 else local
        ev[1].record()
        plan.prepare_dev(full.data_ptr(), n, p, capi.ERTL_MLE, stream)
        ev[2].record()
        plan.run_symmetric_dev(prm, rb, re_, d_out.data_ptr(), stream)
        ev[3].record()
        return full

    for _ in range(args.warmup):
        step()
    cx.barrier()
    l0 = capi.kernel_launches()
    t_wall0 = time.perf_counter()
    ker_ms, prep_ms, ag_ms = [], [], []
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(args.steps):
        full = step()
        ev[3].synchronize()
        ag_ms.append(ev[0].elapsed_time(ev[1])); prep_ms.append(ev[1].elapsed_time(ev[2])); ker_ms.append(ev[2].elapsed_time(ev[3]))
    e_stop.record()
    cx.barrier()
    t_wall1 = time.perf_counter()
    total_ms = cx.max_over_ranks(e_start.elapsed_time(e_stop))
    launches = capi.kernel_launches() - l0
    ms_per_step = total_ms / args.steps
    value = total_pairs / (ms_per_step * 1e-3)
    _, tiles, K = plan.last_run_info()
    ker = float(np.mean(ker_ms))
    alg_bytes = my_pairs * (2 * m + 4)
    achieved = alg_bytes / (ker * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "dist_kernel (TMA-tiled OR+POPC + fused Ertl MLE)", "achieved": achieved, "peak": cx.peak, "unit": "GB/s",
            "frac": achieved / cx.peak, "peak_source": cx.peak_src, "traffic": load_traffic("dist_kernel", n, world),
            "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_pair": 2 * m + 4, "kernel_ms": ker,
            "compulsory_bytes_per_launch": int(n * m + 4 * my_pairs),
            "note": "algorithmic bytes = what the reference streams per pair (SURVEY.md §8(d)); the tiled kernel re-uses planes from SMEM/L2 so "
                    "frac exceeds 1 and says nothing about kernel quality — the binding units are the integer pipes (int_bound)",
            "int_bound": {"word_ops_per_pair": K * (m // 32), "thresholds": K,
                          "word_ops_per_s": my_pairs * K * (m // 32) / (ker * 1e-3),
                          # static pipe model of the sweep (DESIGN.md §4 "Round 2"): on this data 11 of the K thresholds are swept densely
                          # (the others come from the merged sparse / low tails), 512 plane words each; at the optimum POPC : carry-save
                          # mix a word costs 2.0 cycles of the 64-lane ALU pipe (and as much of the 16-lane XU pipe), 148 SMs, 1.965 GHz
                          "pipe_model": {"dense_thresholds_assumed": 11, "alu_cycles_per_word_at_optimum": 2.0,
                                         "ceiling_pairs_per_s_per_gpu": 148 * 64 * 1.965e9 / (11 * (m // 32) * 2.0),
                                         "frac_of_ceiling": (my_pairs / (ker * 1e-3)) / (148 * 64 * 1.965e9 / (11 * (m // 32) * 2.0)),
                                         "measured_pipe_utilisation": "ALU 66 %, XU 63 %, issue 60 % (profiles/r01n_dist_kernel_ncu.txt; same with 24 resident warps: profiles/r02b)"}}}
    clocks = cx.sampler.window(t_wall0, t_wall1) if cx.sampler else None
    gpu_rows_np = None
    regs_np = None
    if rank == 0:
        # values of the device-resident path for the parity check (rank 0's first rows), and the matrix the reference needs
        keep = min(my_pairs, 64_000_000)
        gpu_rows_np = d_out[:keep].cpu().numpy()
        regs_np = full.cpu().numpy() if world > 1 else host_regs.reshape(n, m)

    # ---- e2e: host buffers through the reference-facing C ABI (N=1) / the multi-GPU driver (N>1)
    host_out = capi.pinned_empty(max(my_pairs, 1) * 4).view(np.float32)
    out_t = torch.from_numpy(host_out)
    # N>1: the rank's rows in four blocks of equal pair counts; the device->host copy of block b runs on a second stream
    # while block b+1 computes (what db200_dist_symmetric_rows does inside the library at N=1)
    cuts = [rb]
    for b in range(1, 4):
        target = tri(n, rb) + my_pairs * b // 4
        r = cuts[-1]
        while r < re_ and tri(n, r) < target:
            r += 1
        cuts.append(r)
    cuts.append(re_)
    blocks = [(cuts[i], cuts[i + 1], tri(n, cuts[i]) - tri(n, rb), tri(n, cuts[i + 1]) - tri(n, cuts[i])) for i in range(4) if cuts[i + 1] > cuts[i]]
    copy_stream = torch.cuda.Stream(device=dev)
    blk_ev = [torch.cuda.Event() for _ in blocks]

    def e2e_step():
        if world == 1:
            capi.dist_symmetric(host_regs, p, k=K_MER, result_type=capi.JI, device=cx.local_rank, out=host_out)
        else:
            loc = pin_t.to(dev, non_blocking=True)
            full = multigpu.allgather_registers(loc, counts, dist)
            plan.prepare_dev(full.data_ptr(), n, p, capi.ERTL_MLE, stream)
            for (b0, b1, off, cnt), e in zip(blocks, blk_ev):
                plan.run_symmetric_dev(prm, b0, b1, d_out.data_ptr() + off * 4, stream)
                e.record()
                if cnt:
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(e)
                        out_t[off:off + cnt].copy_(d_out[off:off + cnt], non_blocking=True)
            torch.cuda.synchronize()
    e2e_step()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    cx.barrier()
    e2e_ms = cx.max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    # the host-buffer path must deliver the very same floats as the device-resident step
    if world > 1:
        plan.run_symmetric_dev(prm, rb, re_, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
    if not np.array_equal(d_out[:my_pairs].cpu().numpy(), host_out[:my_pairs]):
        raise RuntimeError("bench: device-resident and host-buffer results differ")
    e2e = {"value": total_pairs / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(counts[rank] * m),
           "d2h_bytes_per_step": int(my_pairs * 4), "ms_per_step": e2e_ms,
           "api": "db200_dist_symmetric(host regs -> host packed float matrix)" if world == 1 else "multigpu driver: pinned shard -> all-gather -> rows in 4 blocks (device->host copy of block b under the kernel of block b+1) -> pinned out"}
    res = {"metric": "pairwise HLL cmp/s (dist p=14)", "value": value, "unit": "pairs/s", "ms_per_step": ms_per_step,
           "config": dist_config(n, world),
           "details": {"step_breakdown_ms": {"allgather": float(np.mean(ag_ms)), "planes+cardinalities": float(np.mean(prep_ms)), "all_pairs_kernel": ker},
                       "tiles": tiles, "live_thresholds": K, "rows_of_rank0": [rb, re_]},
           "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "dtype": "u8 registers / u32 popcounts / f64 estimator -> f32 out"}
    plan.close()
    return res, {"regs_np": regs_np, "gpu_rows_np": gpu_rows_np}


# ---------------------------------------------------------------- sketch (second half of the metric)
def bench_sketch(cx):
    torch, capi, args = cx.torch, cx.capi, cx.args
    world, rank, dev, stream, local_rank = cx.world, cx.rank, cx.dev, cx.stream, cx.local_rank
    k, p, ng, L = K_MER, P_SKETCH, N_GENOMES, GENOME_LEN
    ascii_dev = synth_genomes_torch(torch, ng, L, 4242, dev, first=rank * ng)
    offs = (np.arange(ng + 1, dtype=np.uint64) * np.uint64(L))
    grb = np.arange(ng + 1, dtype=np.uint64)
    pg = capi.PackedGenomes(int(ascii_dev.data_ptr()), offs, grb, k, device=local_rank)
    d_regs = torch.empty((ng, 1 << p), dtype=torch.uint8, device=dev)
    kmers_rank, total_kmers = pg.kmers, pg.kmers * world
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        pg.sketch_dev(p, True, d_regs.data_ptr(), stream)
    cx.barrier()
    l0 = capi.kernel_launches()
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        pg.sketch_dev(p, True, d_regs.data_ptr(), stream)
    e1.record()
    cx.barrier()
    t_wall1 = time.perf_counter()
    launches = capi.kernel_launches() - l0
    ms = cx.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = total_kmers / (ms * 1e-3)
    alg_bytes = pg.packed_bytes + ng * (1 << p)
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "sketch_kernel<smem registers>", "achieved": achieved, "peak": cx.peak, "unit": "GB/s", "frac": achieved / cx.peak,
            "peak_source": cx.peak_src, "traffic": load_traffic("sketch_kernel", ng, world), "algorithmic_bytes_per_launch": alg_bytes,
            "bytes_per_kmer": alg_bytes / max(kmers_rank, 1), "kernel_ms": ms,
            "note": "2-bit bases + validity + record-start planes + register write-back (SURVEY.md §8(d) + the start plane); "
                    "0.5 B per k-mer cannot be HBM bound: the binding unit is the ALU pipe",
            "int_bound": {"sass_instr_per_kmer": 48, "alu_pipe_instr_per_kmer": 25, "fma_pipe_instr_per_kmer": 18,
                          "alu_ceiling_kmers_per_s": 148 * 64 / 25 * 1.965e9 * 1.0,
                          "frac_of_alu_ceiling": value / world / (148 * 64 / 25 * 1.965e9),
                          "source": "cuobjdump -sass of sketch_kernel<0,1,true>, one unrolled base step (DESIGN.md §3)"}}
    clocks = cx.sampler.window(t_wall0, t_wall1) if cx.sampler else None
    # e2e: host ASCII (pinned) -> db200_sketch_batch -> host registers
    try:
        host_ascii = capi.pinned_empty(ng * L)
    except capi.Db200Error as e:   # a box that cannot pin 5 GB per rank still gets its device-resident numbers
        log(f"[bench] pinned host allocation failed ({e}); using pageable memory for the sketch e2e leg")
        host_ascii = np.empty(ng * L, dtype=np.uint8)
    torch.from_numpy(host_ascii).copy_(ascii_dev.cpu())
    del ascii_dev
    regs_ref = d_regs.cpu().numpy()
    pg.close()
    torch.cuda.empty_cache()
    e2e_variants = {}

    def time_e2e(label, env):
        old = {kk: os.environ.get(kk) for kk in env}
        os.environ.update(env)
        try:
            out = capi.sketch_batch(host_ascii, offs, grb, k, p, True, device=local_rank)
            if not np.array_equal(out, regs_ref):
                raise RuntimeError(f"bench: host-buffer sketch ({label}) differs from the device-resident sketch")
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                capi.sketch_batch(host_ascii, offs, grb, k, p, True, device=local_rank)
            cx.barrier()
            return cx.max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        finally:
            for kk, vv in old.items():
                if vv is None:
                    os.environ.pop(kk, None)
                else:
                    os.environ[kk] = vv

    for label, env in (("ascii_upload", {"DB200_HOST_PACK": "0"}), ("host_pack_hybrid", {"DB200_HOST_PACK": "1"})):
        try:
            e2e_variants[label] = time_e2e(label, env)
        except capi.Db200Error as ex:
            log(f"[bench] sketch e2e variant {label} skipped: {ex}")
    best_label = min(e2e_variants, key=e2e_variants.get)
    e2e_ms = e2e_variants[best_label]
    e2e = {"value": total_kmers / (e2e_ms * 1e-3), "unit": "kmers/s", "h2d_bytes_per_step": int(ng * L), "d2h_bytes_per_step": int(ng << p),
           "ms_per_step": e2e_ms, "api": "db200_sketch_batch(host ASCII records -> host registers)", "variant": best_label,
           "variants_ms": e2e_variants,
           "variants_note": "ascii_upload: 8 bits/base over PCIe, packed on the device; host_pack_hybrid: host threads pack to 2 bits + validity "
                            "(0.375 B/base) from one end of the batch while ASCII chunks go up from the other end, so the link never idles"}
    # what bounds it: the host->device link.  Probe: one 1 GiB copy from the same page-locked buffer, best of 3.
    try:
        probe_n = min(1 << 30, ng * L)
        d_probe = torch.empty(probe_n, dtype=torch.uint8, device=dev)
        h_probe = torch.from_numpy(host_ascii[:probe_n])
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            tp = time.perf_counter()
            d_probe.copy_(h_probe, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, probe_n / (time.perf_counter() - tp) / 1e9)
        del d_probe
        e2e["link"] = {"bound": "pcie h2d", "achieved_ascii_equivalent": ng * L / (e2e_ms * 1e-3) / 1e9, "peak": best, "unit": "GB/s",
                       "frac_ascii_equivalent": ng * L / (e2e_ms * 1e-3) / 1e9 / best,
                       "note": "ASCII bytes of input consumed per second through db200_sketch_batch against a plain 1 GiB cudaMemcpy from the same "
                               "page-locked buffer; above 1 means the host-packed share of the batch crossed the link at 0.375 B/base"}
    except Exception as ex:
        log(f"[bench] link probe skipped: {ex}")
    # second end-to-end form: RAW FASTA text (headers + 80-column lines) parsed on the device (db200_sketch_fasta_batch) —
    # what the CLI feeds the library with; informational, the e2e key above stays the record interface
    try:
        if world > 1:
            raise RuntimeError("single-GPU runs only (a second 5 GB page-locked buffer per rank)")
        W = 80
        assert L % W == 0
        hdr_len = 16
        per = hdr_len + (L // W) * (W + 1)
        stride = (per + capi.FASTA_ALIGN - 1) // capi.FASTA_ALIGN * capi.FASTA_ALIGN
        try:
            text = capi.pinned_empty(ng * stride + 64)
        except capi.Db200Error:
            text = np.empty(ng * stride + 64, dtype=np.uint8)
        tv = text[: ng * stride].reshape(ng, stride)
        tv[:, :hdr_len] = np.frombuffer(b">genome 0000000\n", dtype=np.uint8)
        body = tv[:, hdr_len:per].reshape(ng, L // W, W + 1)
        body[:, :, :W] = host_ascii.reshape(ng, L // W, W)
        body[:, :, W] = 10
        foff = (np.arange(ng, dtype=np.uint64) * np.uint64(stride))
        flen = np.full(ng, per, dtype=np.uint64)
        out2 = np.zeros((ng, 1 << p), dtype=np.uint8)
        status = np.zeros(ng, dtype=np.uint8)
        import ctypes as C

        def fasta_step():
            capi._check(capi.lib.db200_sketch_fasta_batch(local_rank, p, k, 1, text.ctypes.data_as(C.c_void_p), foff.ctypes.data_as(capi.u64p),
                                                          flen.ctypes.data_as(capi.u64p), ng, grb.ctypes.data_as(capi.u64p), ng,
                                                          out2.ctypes.data_as(capi.u8p), status.ctypes.data_as(capi.u8p)))
        fasta_step()
        if status.any() or not np.array_equal(out2, regs_ref):
            raise RuntimeError("device-parsed FASTA sketch differs from the device-resident sketch")
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fasta_step()
        cx.barrier()
        f_ms = cx.max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        e2e["fasta_text"] = {"value": total_kmers / (f_ms * 1e-3), "unit": "kmers/s", "ms_per_step": f_ms, "h2d_bytes_per_step": int(ng * per),
                             "api": "db200_sketch_fasta_batch(host FASTA text, 80-column lines -> host registers; kseq rules on the device)"}
        del text, tv, body
    except Exception as ex:   # informational leg: never lose the bench line over it
        log(f"[bench] FASTA-text e2e leg skipped: {ex}")
    res = {"metric": "k-mers hashed/s (sketch k=31 p=14)", "value": value, "unit": "kmers/s", "ms_per_step": ms,
           "config": sketch_config(world),
           "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "dtype": "2-bit bases / u64 k-mers / u8 registers"}
    sample_np = host_ascii[: min(ng, 128) * L] if (rank == 0 and world == 1) else None
    return res, sample_np


# ---------------------------------------------------------------- joint MLE at the C5 shape (N=1)
def bench_jmle(cx):
    """Ertl joint MLE all pairs, p=16, k=21 (estimator and sketch shape of BASELINE configs[4]) on N_JMLE sketches: kernel rate and
    parity of >= 1e6 pairs against the reference's ertl_joint path."""
    torch, capi, args = cx.torch, cx.capi, cx.args
    dev, stream = cx.dev, cx.stream
    n, p, k = N_JMLE, P_JMLE, K_JMLE
    m = 1 << p
    regs_np = host_registers(SEED_JMLE, 0, n, p)
    d_regs = torch.from_numpy(regs_np).to(dev)
    pairs = n * (n - 1) // 2
    d_out = torch.empty(pairs, dtype=torch.float32, device=dev)
    plan = capi.DistPlan(cx.local_rank)
    prm = capi.dist_params(p, k, capi.ERTL_MLE, capi.ERTL_JOINT_MLE, capi.JI, capi.ORDER_ROW_FIRST)
    plan.prepare_dev(d_regs.data_ptr(), n, p, capi.ERTL_MLE, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        plan.run_symmetric_dev(prm, 0, n, d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    steps = max(2, min(args.steps, 5))
    e0.record()
    for _ in range(steps):
        plan.run_symmetric_dev(prm, 0, n, d_out.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    _, tiles, K = plan.last_run_info()
    out = {"metric": "pairwise HLL cmp/s (dist p=16, Ertl joint MLE)", "value": pairs / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "steps": steps,
           "config": dist_config(n, 1, p=p, k=k, seed=SEED_JMLE, what="ERTL_JOINT_MLE JI"), "details": {"tiles": tiles, "live_thresholds": K},
           "roofline": {"bound": "hbm", "kernel": "dist_jmle (three count families per threshold + three MLE solves per pair)",
                        "achieved": pairs * (2 * m + 4) / (ms * 1e-3) / 1e9, "peak": cx.peak, "unit": "GB/s",
                        "frac": pairs * (2 * m + 4) / (ms * 1e-3) / 1e9 / cx.peak, "traffic": None, "bytes_per_pair": 2 * m + 4}}
    if not args.no_cpu_baseline:
        chk, kind, desc = reference_checker()
        threads = ref_threads(kind)
        got = d_out.cpu().numpy()
        v, npairs, rows, dt, vals = cpu_dist_sample(chk, kind, regs_np, n, 1.0, threads, p=p, k=k, jestim=3, rtype=1, min_pairs=1_000_000)
        out["parity"] = parity_stats(got[:npairs], vals, what=f"{kind}: ertl_joint JI, rows [0,{rows}) of the same matrix")
        out["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": f"rows [0,{rows}) = {npairs} pairs in {dt:.1f}s; {desc}"}
    plan.close()
    return out


# ---------------------------------------------------------------- helpers for the 8-GPU configurations
def emu_world(cx):
    """(world, rank) the configuration is sized for: the real ones, or the emulated rank of --emulate-world."""
    if cx.args.emulate_world > 1:
        return cx.args.emulate_world, cx.args.emulate_rank
    return cx.world, cx.rank


def stream_rows(cx, plan, prm, n, rb, re_, block_pairs, ring, on_block=None):
    """Rows [rb, re) of the symmetric matrix in blocks of whole rows (about block_pairs values): kernel -> device ring slot ->
    page-locked ring slot, copy of block b under the kernel of block b+1; the host touches every delivered block.  The multi-process
    counterpart of db200_dist_symmetric_stream (no buffer of the rank's 2.5 GB of floats exists on either side)."""
    torch = cx.torch
    d_ring, h_ring, copied, cstream = ring
    ns = len(d_ring)
    pending = []          # (slot, count, rb, re)
    acc = 0.0

    def deliver():
        nonlocal acc
        sl, cnt, b0, b1 = pending.pop(0)
        copied[sl].synchronize()
        if cnt:
            blk = h_ring[sl][:cnt]
            acc += float(blk[0]) + float(blk[cnt - 1])
            if on_block is not None:
                on_block(b0, b1, blk)
    r = rb
    issued = 0
    while r < re_:
        r1 = r + 1
        while r1 < re_ and tri(n, r1) - tri(n, r) < block_pairs:
            r1 += 1
        cnt = tri(n, r1) - tri(n, r)
        if len(pending) == ns:
            deliver()
        sl = issued % ns
        plan.run_symmetric_dev(prm, r, r1, d_ring[sl].data_ptr(), cx.stream)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(cstream):
            cstream.wait_event(ev)
            if cnt:
                h_ring[sl][:cnt].copy_(d_ring[sl][:cnt], non_blocking=True)
            copied[sl].record()
        pending.append((sl, cnt, r, r1))
        issued += 1
        r = r1
    while pending:
        deliver()
    torch.cuda.synchronize()
    return acc


def make_ring(cx, slot_vals, ns=3):
    torch = cx.torch
    d_ring = [torch.empty(slot_vals, dtype=torch.float32, device=cx.dev) for _ in range(ns)]
    h_ring = [torch.from_numpy(cx.capi.pinned_empty(slot_vals * 4).view(np.float32)) for _ in range(ns)]
    copied = [torch.cuda.Event() for _ in range(ns)]
    return d_ring, h_ring, copied, torch.cuda.Stream(device=cx.dev)


# ---------------------------------------------------------------- C4: 100,000 p=14 sketches over 8 ranks
def bench_c4(cx):
    torch, dist, capi, multigpu, args = cx.torch, cx.dist, cx.capi, cx.multigpu, cx.args
    dev, stream = cx.dev, cx.stream
    W, R = emu_world(cx)
    emulated = args.emulate_world > 1
    n, p, k = C4_N, P_DIST, K_MER
    m = 1 << p
    counts = multigpu.shard_counts(n, W)
    start = sum(counts[:R])
    rb, re_ = multigpu.row_partition(n, W)[R]
    my_pairs, total_pairs = tri(n, re_) - tri(n, rb), n * (n - 1) // 2
    steps, warm = 2, 1
    state = {}

    # phase 1 (local): this rank's shard of the seeded matrix in page-locked memory
    ok = True
    try:
        t0 = time.perf_counter()
        host_regs = capi.pinned_empty(counts[R] * m)
        host_registers(SEED_C4, start, counts[R], p, out=host_regs.reshape(counts[R], m))
        pin_t = torch.from_numpy(host_regs).view(counts[R], m)
        local = pin_t.to(dev)
        others = None
        if emulated:   # the other ranks' shards: device-side synthesis (any valid sketches do; only this rank's rows are checked bit for bit)
            others = synth_registers_torch(torch, n - counts[R], p, SEED_C4 + 1, dev)
        plan = capi.DistPlan(cx.local_rank)
        prm = capi.dist_params(p, k, capi.ERTL_MLE, capi.ERTL_MLE, capi.JI, capi.ORDER_ROW_FIRST)
        block_pairs = 32 << 20
        ring = make_ring(cx, block_pairs + n)
        d_out = torch.empty(my_pairs, dtype=torch.float32, device=dev)
        log(f"[bench/c4] rank {R}: shard rows [{start},{start + counts[R]}) ready in {time.perf_counter() - t0:.1f}s; block rows [{rb},{re_}) = {my_pairs} pairs")
    except Exception as e:
        log(f"[bench/c4] rank {R}: setup failed: {e!r}")
        ok = False
    if not cx.all_ok(ok):
        return {"failed": "setup failed on at least one rank"}

    def gather(loc):
        if emulated:
            full = torch.empty((n, m), dtype=torch.uint8, device=dev)
            full[:start] = others[:start]
            full[start:start + counts[R]] = loc
            full[start + counts[R]:] = others[start:]
            return full
        return multigpu.allgather_registers(loc, counts, dist)

    def gather_prepare(loc):
        full_ = gather(loc)
        ev[1].record()
        plan.prepare_dev(full_.data_ptr(), n, p, capi.ERTL_MLE, stream)
        return full_

    torch.cuda.set_device(cx.local_rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ag, prep, ker, tot = [], [], [], []
    for it in range(warm + steps):
        cx.barrier()
        ev[0].record()
        full = gather_prepare(local)
        ev[2].record()
        plan.run_symmetric_dev(prm, rb, re_, d_out.data_ptr(), stream)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= warm:
            ag.append(ev[0].elapsed_time(ev[1])); prep.append(ev[1].elapsed_time(ev[2])); ker.append(ev[2].elapsed_time(ev[3]))
            tot.append(cx.max_over_ranks(ev[0].elapsed_time(ev[3])))
    _, tiles, K = plan.last_run_info()
    ms = float(np.mean(tot))
    # e2e: pinned shard -> H2D -> all-gather -> planes -> row blocks streamed to page-locked host buffers
    e2e_ms = []
    for it in range(1 + steps):
        cx.barrier()
        t0 = time.perf_counter()
        loc = pin_t.to(dev, non_blocking=True)
        full = gather_prepare(loc)
        stream_rows(cx, plan, prm, n, rb, re_, block_pairs, ring)
        cx.barrier()
        if it >= 1:
            e2e_ms.append(cx.max_over_ranks((time.perf_counter() - t0) * 1e3))
    e2e_t = float(np.mean(e2e_ms))
    scale = (W if emulated else 1)     # an emulated rank reports the whole-job figure its time implies (every rank holds 1/W of the pairs)
    out = {"metric": "pairwise HLL cmp/s (dist p=14)", "value": total_pairs / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "steps": steps, "warmup": warm,
           "n_gpus": W, "config": dist_config(n, W, seed=SEED_C4),
           "details": {"step_breakdown_ms": {"allgather": float(np.mean(ag)), "planes+cardinalities": float(np.mean(prep)), "all_pairs_kernel": float(np.mean(ker))},
                       "tiles": tiles, "live_thresholds": K, "rows_of_this_rank": [rb, re_], "pairs_of_this_rank": my_pairs,
                       "hbm_bytes": {"registers": n * m, "planes": K * n * (m // 8), "out_per_rank": my_pairs * 4}},
           "e2e": {"value": total_pairs / (e2e_t * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_t, "h2d_bytes_per_step": int(counts[R] * m),
                   "d2h_bytes_per_step": int(my_pairs * 4),
                   "api": "multigpu driver: pinned shard -> all-gather -> planes -> row blocks of 32 Mi pairs through a 3-slot device/page-locked ring (the multi-process form of db200_dist_symmetric_stream)"},
           "roofline": {"bound": "hbm", "kernel": "dist_kernel", "achieved": my_pairs * (2 * m + 4) / (float(np.mean(ker)) * 1e-3) / 1e9, "peak": cx.peak, "unit": "GB/s",
                        "frac": my_pairs * (2 * m + 4) / (float(np.mean(ker)) * 1e-3) / 1e9 / cx.peak, "traffic": None}}
    if emulated:
        out["emulated"] = {"world": W, "rank": R, "note": "one rank's share on one GPU; the all-gather is replaced by a device copy, the other ranks' sketches are device-synthesised"}
    # parity (rank 0 of the job / the emulated rank): the first rows of this rank's block against the reference on the same matrix
    if cx.rank == 0 and not args.no_cpu_baseline:
        try:
            chk, kind, desc = reference_checker()
            threads = ref_threads(kind)
            regs_np = full.cpu().numpy()
            v, npairs, rows, dt, vals = cpu_dist_sample(chk, kind, regs_np, n, 1.0, threads, row0=rb, min_pairs=2_000_000)
            got = d_out[:npairs].cpu().numpy()
            out["parity"] = parity_stats(got, vals, what=f"{kind}: rows [{rb},{rb + rows}) of the gathered matrix")
            del regs_np
        except Exception as e:
            log(f"[bench/c4] parity failed: {e!r}")
    plan.close()
    return out


# ---------------------------------------------------------------- C5: 50,000 x 5 Mbp, k=21, p=16, joint MLE, sketch + all pairs
def bench_c5(cx):
    torch, dist, capi, multigpu, args = cx.torch, cx.dist, cx.capi, cx.multigpu, cx.args
    dev, stream = cx.dev, cx.stream
    W, R = emu_world(cx)
    emulated = args.emulate_world > 1
    n, p, k, L = C5_GENOMES, C5_P, C5_K, GENOME_LEN
    m = 1 << p
    counts = multigpu.shard_counts(n, W)
    start = sum(counts[:R])
    ng = counts[R]
    rb, re_ = multigpu.row_partition(n, W)[R]
    my_pairs, total_pairs = tri(n, re_) - tri(n, rb), n * (n - 1) // 2
    B = 625                                   # genomes per on-device batch (3.1 GB of ASCII)
    steps, warm = 1, 1
    ok = True
    try:
        local = torch.zeros((ng, m), dtype=torch.uint8, device=dev)
        ascii_buf = torch.empty(B * L, dtype=torch.uint8, device=dev)
        others = synth_registers_torch(torch, n - ng, p, SEED_C5 + 1, dev) if emulated else None
        plan = capi.DistPlan(cx.local_rank)
        prm = capi.dist_params(p, k, capi.ERTL_MLE, capi.ERTL_JOINT_MLE, capi.JI, capi.ORDER_ROW_FIRST)
        block_pairs = 32 << 20
        ring = make_ring(cx, block_pairs + n)
        keep_rows = rows_for_pairs(n, rb, 300_000)
        d_keep = torch.empty(tri(n, rb + keep_rows) - tri(n, rb), dtype=torch.float32, device=dev)
    except Exception as e:
        log(f"[bench/c5] rank {R}: setup failed: {e!r}")
        ok = False
    if not cx.all_ok(ok):
        return {"failed": "setup failed on at least one rank"}

    def gather(loc):
        if emulated:
            full = torch.empty((n, m), dtype=torch.uint8, device=dev)
            full[:start] = others[:start]
            full[start:start + ng] = loc
            full[start + ng:] = others[start:]
            return full
        return multigpu.allgather_registers(loc, counts, dist)

    torch.cuda.set_device(cx.local_rank)
    sample_genomes = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = []
    pg = None
    for it in range(warm + steps):
        # ---- sketch: genomes are generated on the device batch by batch (untimed), packed and sketched (timed)
        t_pack = t_sk = 0.0
        kmers = 0
        for b0 in range(0, ng, B):
            nb = min(B, ng - b0)
            synth_genomes_torch(torch, nb, L, SEED_C5, dev, first=start + b0, out=ascii_buf)
            if it == 0 and b0 == 0 and cx.rank == 0:
                for gi in (0, 1):
                    sample_genomes[start + gi] = ascii_buf[gi * L:(gi + 1) * L].cpu().numpy()
            offs = (np.arange(nb + 1, dtype=np.uint64) * np.uint64(L))
            grb = np.arange(nb + 1, dtype=np.uint64)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if pg is None:                                                                            # ASCII in HBM -> 2-bit store (synchronous)
                pg = capi.PackedGenomes(int(ascii_buf.data_ptr()), offs, grb, k, device=cx.local_rank)
            else:
                pg.repack(int(ascii_buf.data_ptr()), offs, grb, k)                                    # the store's device buffers are reused
            t_pack += (time.perf_counter() - t0) * 1e3
            e0.record()
            pg.sketch_dev(p, True, local[b0:b0 + nb].data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            t_sk += e0.elapsed_time(e1)
            kmers += pg.kmers
        # ---- all-gather + planes + joint-MLE all pairs, row blocks streamed to the host
        cx.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t0 = time.perf_counter()
        ev[0].record()
        full = gather(local)
        ev[1].record()
        plan.prepare_dev(full.data_ptr(), n, p, capi.ERTL_MLE, stream)
        ev[2].record()
        stream_rows(cx, plan, prm, n, rb, re_, block_pairs, ring)
        t_pairs_wall = (time.perf_counter() - t0) * 1e3
        cx.barrier()
        if it >= warm:
            res.append({"pack": t_pack, "sketch": t_sk, "allgather": ev[0].elapsed_time(ev[1]), "planes": ev[1].elapsed_time(ev[2]),
                        "gather+planes+pairs+d2h_wall": t_pairs_wall, "kmers": kmers})
    if pg is not None:
        pg.close()
    r0 = res[-1]
    _, tiles, K = plan.last_run_info()
    sketch_ms = cx.max_over_ranks(r0["pack"] + r0["sketch"])
    pairs_ms = cx.max_over_ranks(r0["gather+planes+pairs+d2h_wall"])
    total_ms = sketch_ms + pairs_ms
    out = {"metric": "end-to-end sketch + all-pairs time, 50,000 genomes (genomes/s)", "value": n / (total_ms * 1e-3), "unit": "genomes/s", "ms_per_step": total_ms,
           "steps": steps, "warmup": warm, "n_gpus": W,
           "config": {"workload": f"e2e sketch+dist: {n} x {L} bp synthetic genomes, k={k}, p={p}, Ertl joint MLE JI ({total_pairs} pairs)", "genomes_per_gpu": ng,
                      "data": "genomes generated on the device batch by batch (SURVEY.md §8(d): C5 is never written to disk); sketch time = ASCII in HBM -> 2-bit store -> registers",
                      "parallelism": f"genomes x{W}, 1 NCCL all-gather of {n * m >> 20} MB of registers, block-row x{W}"},
           "details": {"step_breakdown_ms": {"pack_ascii_to_2bit": r0["pack"], "sketch_kernel": r0["sketch"], "allgather": r0["allgather"], "planes+cardinalities": r0["planes"],
                                             "allgather+planes+all_pairs+d2h (wall)": r0["gather+planes+pairs+d2h_wall"]},
                       "sketch_kmers_per_s_whole_job": r0["kmers"] * W / (sketch_ms * 1e-3), "dist_pairs_per_s_whole_job": total_pairs / (pairs_ms * 1e-3),
                       "tiles": tiles, "live_thresholds": K, "rows_of_this_rank": [rb, re_], "pairs_of_this_rank": my_pairs,
                       "hbm_bytes": {"registers": n * m, "planes": K * n * (m // 8)}},
           "e2e": {"value": n / (total_ms * 1e-3), "unit": "genomes/s", "ms_per_step": total_ms, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(my_pairs * 4),
                   "note": "inputs are device-generated (31 GB of ASCII per rank has no host copy); the float matrix is streamed to page-locked host buffers inside the timed region"}}
    if emulated:
        out["emulated"] = {"world": W, "rank": R, "note": "one rank's share on one GPU; the other ranks' sketches are device-synthesised register arrays"}
    if cx.rank == 0 and not args.no_cpu_baseline:
        try:
            chk, kind, desc = reference_checker()
            # (a) registers of two device-generated genomes against the reference's Encoder + addh, bit for bit
            reg_ok = True
            for gidx, seq in sample_genomes.items():
                wantr = chk.sketch([seq.tobytes()], k, p, True)
                reg_ok = reg_ok and bool(np.array_equal(local[gidx - start].cpu().numpy(), wantr))
            # (b) a sample of this rank's first rows against the reference's joint MLE
            plan.run_symmetric_dev(prm, rb, rb + keep_rows, d_keep.data_ptr(), stream)
            torch.cuda.synchronize()
            got_rows = d_keep.cpu().numpy()
            rng = np.random.default_rng(5)
            ns = 2000
            ii = rb + rng.integers(0, max(1, min(keep_rows, n - 1 - rb)), size=ns)
            jj = np.array([rng.integers(i + 1, n) for i in ii])
            rows_needed = np.unique(np.concatenate([ii, jj]))
            host_rows = {int(r): full[int(r)].cpu().numpy() for r in rows_needed}
            got = np.array([got_rows[tri(n, int(i)) - tri(n, rb) + int(j) - int(i) - 1] for i, j in zip(ii, jj)])
            wantv = np.array([chk.pair(host_rows[int(i)], host_rows[int(j)], p, estim=2, jestim=3, rtype=1, k=k) for i, j in zip(ii, jj)])
            out["parity"] = parity_stats(got, wantv, what=f"{kind}: {ns} sampled pairs of rows [{rb},{rb + keep_rows}) through ertl_joint; registers of 2 device-generated genomes bit-identical: {reg_ok}")
            out["parity"]["registers_bit_identical"] = reg_ok
        except Exception as e:
            log(f"[bench/c5] parity failed: {e!r}")
    plan.close()
    return out


_RESULT_OUT = None


def emit_result(line: dict):
    """The ONE JSON line of the contract goes to the process's original stdout."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries write to file descriptor 1 behind Python's back (NCCL prints "NCCL version ..." there when the environment
    # sets NCCL_DEBUG=VERSION): keep the real stdout for the result line only and point fd 1 at stderr for everything else.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="both", choices=["dist", "sketch", "both"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the jmle / c4 / c5 legs")
    ap.add_argument("--only", default="", help="comma list of extra legs to run (c4,c5)")
    ap.add_argument("--emulate-world", type=int, default=0, help="single GPU: run ONE rank's share of the 8-GPU configurations c4 / c5")
    ap.add_argument("--emulate-rank", type=int, default=0)
    ap.add_argument("--legs-timeout", type=float, default=300.0, help="N=8: seconds the c4 + c5 legs (child processes) may take before the line is printed without them")
    ap.add_argument("--legs-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: fewer than 3 warm-up steps; timing rules ask for W >= 3")
    if args.impl == "reference":
        if args.workload == "both":
            args.workload = "dist"
        return run_reference(args)
    if args.legs_child:
        return run_legs_child(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
