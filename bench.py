#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload dist|sketch|both]

Primary line (default): pairwise HLL comparisons/s, all-pairs on p=14 sketches (configs[2]: 10,000 sketches
on one B200).  A "step" is one pass of the hot path over one batch: (N>1: NCCL all-gather of the register
shards) -> threshold-plane build + per-sketch cardinalities -> tiled all-pairs kernel.  `value` times it with
the register matrix resident in HBM; `e2e` times the reference-facing C-ABI call with HOST buffers (H2D of
the registers and D2H of the float matrix inside the timed region).
The same JSON line carries a `sketch` object with the second half of the metric (k-mers hashed/s,
configs[1]: 1,000 x 5 Mbp genomes, k=31, p=14), measured the same way; `--workload sketch` makes it primary.

Multi-GPU (torchrun, one rank per GPU): weak scaling.  dist: n(N) = round(10000*sqrt(N)) sketches, each rank
holds n/N of them, one all-gather, block-rows balanced by pair count; sketch: 1,000 genomes per rank.

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
def synth_registers_torch(torch, n, p, seed, device, card=5e6, group=16):
    """Device-side version of dashing_b200.synth.registers (same construction, torch RNG): correlated groups of
    HLL register arrays drawn from the exact register distribution of an HLL holding `card` items."""
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


def synth_genomes_torch(torch, n, length, seed, device, group=50):
    """n single-record genomes of `length` ASCII bases on the device: groups share an ancestor, members are
    mutated copies (substitution-rate ladder of SURVEY.md §8(d)).  Returns a uint8 [n*length] tensor."""
    from dashing_b200.synth import RATE_LADDER
    g = torch.Generator(device=device); g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    out = torch.empty(n * length, dtype=torch.uint8, device=device)
    anc = None
    for i in range(n):
        if i % group == 0:
            anc = torch.randint(0, 4, (length,), generator=g, device=device, dtype=torch.uint8)
        rate = RATE_LADDER[i % len(RATE_LADDER)]
        if rate <= 0:
            code = anc
        else:
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


def load_traffic(kind):
    """ncu dram bytes per launch of the dominant kernel, if a capture has been summarised under profiles/."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return d.get(kind)
    except Exception:
        return None


def usable_cores():
    from oracle import oracle as O   # shared with the tests: affinity capped by the cgroup CPU quota
    return O.usable_cores()


def dist_n_for(world):
    return int(round(N_DIST_1GPU * math.sqrt(world)))


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------
def reference_checker():
    from oracle import oracle as O
    if O.ref_available():
        r = O.ref()
        return r, "reference", f"oracle/_ref/{r.libname} (unmodified dashing headers, g++ -O3, SIMD tier {r.simd_tier()})"
    return O.port(), "port", "oracle/oracle_port.c (scalar C restatement; oracle/_ref did not travel)"


def cpu_dist_sample(chk, kind, regs_np, n, target_s, threads):
    """Time rows [0, R) of the SAME all-pairs workload on the host cores (the perform_core_op loop: OpenMP dynamic
    over one matrix row at a time).  The n sketches are constructed and report()ed once, outside the timed region,
    as the reference does before its pair loop."""
    p = P_DIST
    if kind == "reference":
        hs = chk.set_create(regs_np, p, 2, 2)
        run = lambda r: chk.set_dist_rows(hs, K_MER, 1, 0, 0, r, threads)
    else:
        hs = None
        run = lambda r: chk.dist_rows(regs_np, p, k=K_MER, rtype=1, row_begin=0, row_end=r)
    run(2)  # warm caches / thread pool
    rows = 8
    while True:
        t0 = time.perf_counter(); run(rows); dt = time.perf_counter() - t0
        if dt >= 0.5 * target_s or rows >= n - 1:
            break
        rows = min(n - 1, int(rows * min(8.0, max(2.0, target_s / max(dt, 1e-3)))))
    if hs is not None:
        chk.set_free(hs)
    pairs = rows * (2 * n - rows - 1) // 2
    return pairs / dt, pairs, rows, dt


def cpu_sketch_sample(chk, kind, genomes_np, length, threads, target_s=6.0):
    """Repeated passes over a block of in-memory genomes (Encoder::for_each + hll_t::addh, one genome per OpenMP task)
    until ~target_s of CPU work has been timed."""
    ng = genomes_np.size // length
    offs = (np.arange(ng + 1, dtype=np.uint64) * np.uint64(length))
    grb = np.arange(ng + 1, dtype=np.uint64)

    def one_pass():
        if kind == "reference":
            chk.sketch_many(genomes_np, offs, grb, K_MER, P_SKETCH, True, threads)
        else:
            for gi in range(ng):
                chk.sketch([genomes_np[gi * length:(gi + 1) * length].tobytes()], K_MER, P_SKETCH, True)
    one_pass()
    passes, t0 = 0, time.perf_counter()
    while True:
        one_pass(); passes += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or passes >= 200:
            break
    kmers = passes * ng * (length - K_MER + 1)
    return kmers / dt, kmers, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    chk, kind, desc = reference_checker()
    threads = (min(chk.max_threads(), usable_cores()) if kind == "reference" else 1)
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
        metric, unit, workload = "k-mers hashed/s (sketch k=31 p=14)", "kmers/s", f"sketch {N_GENOMES * args.gpus} x {GENOME_LEN} bp, k=31, p=14"
    else:
        regs = synth.registers(2026, n if n <= 12000 else 12000, P_DIST)  # matrix rows only matter through the sampled rows
        nn = regs.shape[0]
        vals = []
        for it in range(args.warmup + args.steps):
            v, pairs, rows, dt = cpu_dist_sample(chk, kind, regs, nn, 6.0, threads)
            if it >= args.warmup:
                vals.append((v, dt, rows, pairs))
        value = float(np.mean([v[0] for v in vals])); ms = float(np.mean([v[1] for v in vals])) * 1e3
        sample = f"rows [0,{vals[-1][2]}) = {vals[-1][3]} of the {n * (n - 1) // 2} pairs, perform_core_op loop (OpenMP dynamic), {threads} threads"
        metric, unit, workload = "pairwise HLL cmp/s (dist p=14)", "pairs/s", f"dist all-pairs {n} p=14 sketches"
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
            "config": {"workload": workload, "estimator": "ERTL_MLE", "result": "JI", "reference": desc},
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_result(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from dashing_b200 import capi, multigpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if capi.device_count() < 1:
        raise RuntimeError("bench.py: libdashing_b200 sees no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream
    peak, peak_src = load_peaks()

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

    sampler = ClockSampler(local_rank) if rank == 0 else None
    results = {}

    # ---------------------------------------------------------------- dist
    def bench_dist():
        p, n = P_DIST, dist_n_for(world)
        m = 1 << p
        counts = multigpu.shard_counts(n, world)
        start = sum(counts[:rank])
        # every rank generates the same matrix (seeded) and keeps its shard: stands in for "rank r sketched these genomes"
        full_src = synth_registers_torch(torch, n, p, 2026, dev)
        local = full_src[start:start + counts[rank]].contiguous()
        del full_src
        torch.cuda.empty_cache()
        rb, re_ = multigpu.row_partition(n, world)[rank]
        my_pairs = multigpu.tri_offset(n, re_) - multigpu.tri_offset(n, rb)
        total_pairs = n * (n - 1) // 2
        d_out = torch.empty(max(my_pairs, 1), dtype=torch.float32, device=dev)
        plan = capi.DistPlan(local_rank)
        prm = capi.dist_params(p, K_MER, capi.ERTL_MLE, capi.ERTL_MLE, capi.JI, capi.ORDER_ROW_FIRST)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

        def step(timed):
            ev[0].record()
            full = multigpu.allgather_registers(local, counts, dist) if world > 1 else local
            ev[1].record()
            plan.prepare_dev(full.data_ptr(), n, p, capi.ERTL_MLE, stream)
            ev[2].record()
            plan.run_symmetric_dev(prm, rb, re_, d_out.data_ptr(), stream)
            ev[3].record()
            return full

        for _ in range(args.warmup):
            step(False)
        barrier()
        l0 = capi.kernel_launches()
        t_wall0 = time.perf_counter()
        ker_ms, prep_ms, ag_ms = [], [], []
        e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record()
        for _ in range(args.steps):
            step(True)
            ev[3].synchronize()
            ag_ms.append(ev[0].elapsed_time(ev[1])); prep_ms.append(ev[1].elapsed_time(ev[2])); ker_ms.append(ev[2].elapsed_time(ev[3]))
        e_stop.record()
        barrier()
        t_wall1 = time.perf_counter()
        total_ms = max_over_ranks(e_start.elapsed_time(e_stop))
        launches = capi.kernel_launches() - l0
        ms_per_step = total_ms / args.steps
        value = total_pairs / (ms_per_step * 1e-3)
        _, tiles, K = plan.last_run_info()
        ker = float(np.mean(ker_ms))
        alg_bytes = my_pairs * (2 * m + 4)
        achieved = alg_bytes / (ker * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "dist_kernel (TMA-tiled OR+POPC + fused Ertl MLE)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": load_traffic("dist_kernel"),
                "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_pair": 2 * m + 4, "kernel_ms": ker,
                "note": "algorithmic bytes = what the reference streams per pair (SURVEY.md §8(d)); the tiled kernel re-uses planes from SMEM/L2 so "
                        "frac may exceed 1 — the binding unit is the integer POPC pipe",
                "int_bound": {"word_ops_per_pair": K * (m // 32), "thresholds": K,
                              "word_ops_per_s": my_pairs * K * (m // 32) / (ker * 1e-3)}}
        clocks = sampler.window(t_wall0, t_wall1) if sampler else None

        # ---- e2e: host buffers through the reference-facing C ABI (N=1) / the multi-GPU driver (N>1)
        host_regs = capi.pinned_empty(counts[rank] * m)
        host_regs[:] = local.cpu().numpy().reshape(-1)
        host_out = capi.pinned_empty(max(my_pairs, 1) * 4).view(np.float32)
        pin_t = torch.from_numpy(host_regs).view(counts[rank], m)
        out_t = torch.from_numpy(host_out)

        # N>1: the rank's rows in four blocks of equal pair counts; the device->host copy of block b runs on a second stream
        # while block b+1 computes (what db200_dist_symmetric_rows does inside the library at N=1)
        tri = lambda r: multigpu.tri_offset(n, r)
        cuts = [rb]
        for b in range(1, 4):
            target = tri(rb) + my_pairs * b // 4
            r = cuts[-1]
            while r < re_ and tri(r) < target:
                r += 1
            cuts.append(r)
        cuts.append(re_)
        blocks = [(cuts[i], cuts[i + 1], tri(cuts[i]) - tri(rb), tri(cuts[i + 1]) - tri(cuts[i])) for i in range(4) if cuts[i + 1] > cuts[i]]
        copy_stream = torch.cuda.Stream(device=dev)
        blk_ev = [torch.cuda.Event() for _ in blocks]

        def e2e_step():
            if world == 1:
                capi.dist_symmetric(host_regs, p, k=K_MER, result_type=capi.JI, device=local_rank, out=host_out)
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
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        # sanity: device-resident and host paths agree
        if world == 1 and not np.array_equal(d_out.cpu().numpy()[:1000], host_out[:1000]):
            raise RuntimeError("bench: device-resident and host-buffer results differ")
        if world > 1:    # the blocked e2e path leaves the same rows in d_out / host_out as the one-launch step
            plan.run_symmetric_dev(prm, rb, re_, d_out.data_ptr(), stream)
            torch.cuda.synchronize()
            if not np.array_equal(d_out[:my_pairs].cpu().numpy(), host_out[:my_pairs]):
                raise RuntimeError("bench: blocked multi-GPU e2e rows differ from the one-launch rows")
        e2e = {"value": total_pairs / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(counts[rank] * m),
               "d2h_bytes_per_step": int(my_pairs * 4), "ms_per_step": e2e_ms,
               "api": "db200_dist_symmetric(host regs -> host packed float matrix)" if world == 1 else "multigpu driver: pinned shard -> all-gather -> rows in 4 blocks (device->host copy of block b under the kernel of block b+1) -> pinned out"}
        res = {"metric": "pairwise HLL cmp/s (dist p=14)", "value": value, "unit": "pairs/s", "ms_per_step": ms_per_step,
               "config": {"workload": f"dist all-pairs {n} p=14 sketches ({total_pairs} pairs), ERTL_MLE union JI", "n_sketches": n, "p": p, "k": K_MER,
                          "estimator": "ERTL_MLE", "result": "JI", "parallelism": f"block-row x{world}" + (" + 1 NCCL all-gather" if world > 1 else ""),
                          "l2": "inputs (register matrix %d MB + threshold planes) larger than the 126 MB L2" % (n * m >> 20),
                          "step_breakdown_ms": {"allgather": float(np.mean(ag_ms)), "planes+cardinalities": float(np.mean(prep_ms)), "all_pairs_kernel": ker},
                          "tiles": tiles, "live_thresholds": K},
               "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "dtype": "u8 registers / u32 popcounts / f64 estimator -> f32 out"}
        host_regs_np = local.cpu().numpy() if (rank == 0 and world == 1) else None
        plan.close()
        return res, host_regs_np

    # ---------------------------------------------------------------- sketch
    def bench_sketch():
        k, p, ng, L = K_MER, P_SKETCH, N_GENOMES, GENOME_LEN
        ascii_dev = synth_genomes_torch(torch, ng, L, 4242 + rank, dev)
        offs = (np.arange(ng + 1, dtype=np.uint64) * np.uint64(L))
        grb = np.arange(ng + 1, dtype=np.uint64)
        pg = capi.PackedGenomes(int(ascii_dev.data_ptr()), offs, grb, k, device=local_rank)
        d_regs = torch.empty((ng, 1 << p), dtype=torch.uint8, device=dev)
        kmers_rank, total_kmers = pg.kmers, pg.kmers * world
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(args.warmup):
            pg.sketch_dev(p, True, d_regs.data_ptr(), stream)
        barrier()
        l0 = capi.kernel_launches()
        t_wall0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            pg.sketch_dev(p, True, d_regs.data_ptr(), stream)
        e1.record()
        barrier()
        t_wall1 = time.perf_counter()
        launches = capi.kernel_launches() - l0
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        value = total_kmers / (ms * 1e-3)
        alg_bytes = pg.packed_bytes + ng * (1 << p)
        achieved = alg_bytes / (ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "sketch_kernel<smem registers>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": load_traffic("sketch_kernel"), "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_kmer": alg_bytes / max(kmers_rank, 1), "kernel_ms": ms,
                "note": "2-bit bases + validity + record-start planes + register write-back (SURVEY.md §8(d) + the start plane); "
                        "0.5 B per k-mer cannot be HBM bound: the binding unit is the ALU pipe",
                "int_bound": {"sass_instr_per_kmer": 48, "alu_pipe_instr_per_kmer": 25, "fma_pipe_instr_per_kmer": 18,
                              "alu_ceiling_kmers_per_s": 148 * 64 / 25 * 1.965e9 * 1.0,
                              "frac_of_alu_ceiling": value / world / (148 * 64 / 25 * 1.965e9),
                              "source": "cuobjdump -sass of sketch_kernel<0,1,true>, one unrolled base step (DESIGN.md §3)"}}
        clocks = sampler.window(t_wall0, t_wall1) if sampler else None
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
        out = capi.sketch_batch(host_ascii, offs, grb, k, p, True, device=local_rank)
        if not np.array_equal(out, regs_ref):
            raise RuntimeError("bench: host-buffer sketch differs from the device-resident sketch")
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            capi.sketch_batch(host_ascii, offs, grb, k, p, True, device=local_rank)
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        e2e = {"value": total_kmers / (e2e_ms * 1e-3), "unit": "kmers/s", "h2d_bytes_per_step": int(ng * L), "d2h_bytes_per_step": int(ng << p),
               "ms_per_step": e2e_ms, "api": "db200_sketch_batch(host ASCII records -> host registers)"}
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
            e2e["link"] = {"bound": "pcie h2d", "achieved": ng * L / (e2e_ms * 1e-3) / 1e9, "peak": best, "unit": "GB/s",
                           "frac": ng * L / (e2e_ms * 1e-3) / 1e9 / best,
                           "note": "ASCII bytes per second through db200_sketch_batch against a plain 1 GiB cudaMemcpy from the same page-locked buffer"}
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
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fasta_step()
            barrier()
            f_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
            e2e["fasta_text"] = {"value": total_kmers / (f_ms * 1e-3), "unit": "kmers/s", "ms_per_step": f_ms, "h2d_bytes_per_step": int(ng * per),
                                 "api": "db200_sketch_fasta_batch(host FASTA text, 80-column lines -> host registers; kseq rules on the device)"}
            del text, tv, body
        except Exception as ex:   # informational leg: never lose the bench line over it
            log(f"[bench] FASTA-text e2e leg skipped: {ex}")
        res = {"metric": "k-mers hashed/s (sketch k=31 p=14)", "value": value, "unit": "kmers/s", "ms_per_step": ms,
               "config": {"workload": f"sketch {ng * world} x {L} bp synthetic genomes, k={k}, p={p}, canonical", "genomes_per_gpu": ng, "k": k, "p": p,
                          "parallelism": f"genomes x{world} (no collective)", "l2": "packed input %d MB per GPU, larger than the 126 MB L2" % (pg.packed_bytes >> 20)},
               "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "dtype": "2-bit bases / u64 k-mers / u8 registers"}
        sample_np = host_ascii[: min(ng, 128) * L] if (rank == 0 and world == 1) else None
        return res, sample_np

    want = ("dist", "sketch") if args.workload == "both" else (args.workload,)
    extra = {}
    for w in want:
        t0 = time.perf_counter()
        results[w], extra[w] = bench_dist() if w == "dist" else bench_sketch()
        log(f"[bench] {w}: {results[w]['value']:.4g} {results[w]['unit']} ({time.perf_counter() - t0:.1f}s incl. setup)")
        torch.cuda.empty_cache()
    if sampler:
        sampler.stop()

    # ---- CPU baseline (rank 0, N=1 only): the reference's own loop on the host cores, bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            chk, kind, desc = reference_checker()
            threads = min(chk.max_threads(), usable_cores()) if kind == "reference" else 1
            if "dist" in results:
                n = results["dist"]["config"]["n_sketches"]
                v, pairs, rows, dt = cpu_dist_sample(chk, kind, extra["dist"], n, 12.0, threads)
                results["dist"]["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind,
                                                   "sample": f"rows [0,{rows}) = {pairs} of the {n * (n - 1) // 2} pairs of the same matrix in {dt:.1f}s; {desc}"}
            if "sketch" in results:
                gen = extra["sketch"]
                ngs = min(gen.size // GENOME_LEN, max(16, min(2 * threads, 128)))
                v, kmers, dt = cpu_sketch_sample(chk, kind, np.ascontiguousarray(gen[: ngs * GENOME_LEN]), GENOME_LEN, threads)
                results["sketch"]["cpu_baseline"] = {"value": v, "unit": "kmers/s", "cores": threads, "kind": kind,
                                                     "sample": f"{kmers} k-mers in {dt:.1f}s: repeated passes over {ngs} of the {N_GENOMES} genomes, in-memory Encoder::for_each + addh; {desc}"}
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU numbers
            log(f"[bench] cpu_baseline failed: {e!r}")

    if rank == 0:
        primary = "dist" if "dist" in results else "sketch"
        r = results[primary]
        line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": r["dtype"],
                "data": "synthetic", "config": r["config"], "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
                "roofline": r["roofline"]}
        if "cpu_baseline" in r:
            line["cpu_baseline"] = r["cpu_baseline"]
        for other in results:
            if other != primary:
                o = results[other]
                line[other] = {k: o[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "roofline", "e2e", "gpu_launches", "clocks") if k in o}
                if "cpu_baseline" in o:
                    line[other]["cpu_baseline"] = o["cpu_baseline"]
        emit_result(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: fewer than 3 warm-up steps; timing rules ask for W >= 3")
    if args.impl == "reference":
        if args.workload == "both":
            args.workload = "dist"
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
