"""CPU tier, world_size 2 over gloo: the CONTROL FLOW of bench.py's multi-rank legs (c4: BASELINE configs[3], c5: configs[4]) with
the device stubbed out — every collective the legs issue (status gates, barriers, all-gathers, max-over-ranks), the row-block
streaming ring, the per-rank bookkeeping and the shape of the result objects.  The first 8-GPU run of round 2 died in these
legs on a bug a single-GPU emulation could not show (ranks other than 0); this test runs them with real ranks.
Nothing is computed here: the plan and the sketcher are recorders, so parity objects are present but meaningless."""
import contextlib
import importlib.util
import os
import socket
import sys
import time

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        assert self.t is not None, "synchronize() on an event that was never recorded"

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _FakeStream:
    def __init__(self, device=None):
        pass

    def wait_event(self, ev):
        assert ev.t is not None


class _FakeCuda:
    Event, Stream = _FakeEvent, _FakeStream
    current = [None]

    @staticmethod
    def synchronize():
        pass

    @staticmethod
    def set_device(d):
        _FakeCuda.current[0] = d

    @staticmethod
    def empty_cache():
        pass

    @staticmethod
    @contextlib.contextmanager
    def stream(s):
        yield


class _FakeTorch:
    """torch with .cuda replaced; tensors live on the CPU device."""
    cuda = _FakeCuda

    def __getattr__(self, name):
        return getattr(torch, name)


class _Plan:
    def __init__(self, device=0):
        self.calls, self.n = [], 0

    def prepare_dev(self, d_regs, n, p, estim, stream=0):
        self.calls.append(("prepare", n, p)); self.n = n

    def run_symmetric_dev(self, prm, rb, re_, d_out, stream=0):
        assert 0 <= rb <= re_ <= self.n
        self.calls.append(("run", rb, re_))
        # a recognisable, NaN-free value per pair (the output buffer is uninitialised memory otherwise): its distmat index
        import ctypes
        n = self.n
        tri = lambda r: r * (2 * n - r - 1) // 2
        cnt = tri(re_) - tri(rb)
        if cnt:
            np.ctypeslib.as_array((ctypes.c_float * cnt).from_address(d_out))[:] = np.arange(tri(rb), tri(re_), dtype=np.float32)

    def last_run_info(self):
        return 0, 7, 30

    def close(self):
        self.calls.append(("close",))


class _Packed:
    def __init__(self, bases, offs, grb, k, device=0):
        self.kmers = int(sum(max(0, int(offs[i + 1] - offs[i]) - k + 1) for i in range(len(offs) - 1)))
        self.n_repack = 0

    def repack(self, bases, offs, grb, k):
        self.n_repack += 1

    def sketch_dev(self, p, canon, d_regs, stream=0):
        pass

    def close(self):
        pass


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, fail_rank, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
        from dashing_b200 import capi as real_capi, multigpu
        # the configurations, shrunk: same code paths, toy sizes
        bench.C45_WORLD, bench.C4_N, bench.C5_GENOMES, bench.GENOME_LEN = world, 96, 12, 3000
        bench.usable_cores = lambda: 2

        class FakeCapi:
            ERTL_MLE, ERTL_JOINT_MLE, JI, ORDER_ROW_FIRST = real_capi.ERTL_MLE, real_capi.ERTL_JOINT_MLE, real_capi.JI, real_capi.ORDER_ROW_FIRST
            dist_params = staticmethod(real_capi.dist_params)
            DistPlan, PackedGenomes = _Plan, _Packed
            made_pinned = 0
            kernel_launches = staticmethod(lambda: 0)

            @staticmethod
            def pinned_empty(nbytes):
                FakeCapi.made_pinned += 1
                return np.zeros(nbytes, dtype=np.uint8)

        if fail_rank == rank:      # a rank that fails INSIDE a leg (after the setup gate): what sank the 8-GPU run
            class Boom(_Plan):
                def prepare_dev(self, *a, **k):
                    raise RuntimeError("injected failure on this rank")
            FakeCapi.DistPlan = Boom

        class Args:
            emulate_world, emulate_rank, no_cpu_baseline, only, steps, warmup = 0, 0, True, "", 2, 1

        cx = bench.Ctx()
        cx.torch, cx.dist, cx.capi, cx.multigpu, cx.args = _FakeTorch(), dist, FakeCapi, multigpu, Args
        cx.world, cx.rank, cx.local_rank, cx.dev, cx.stream = world, rank, rank, torch.device("cpu"), 0
        cx.peak, cx.peak_src, cx.sampler = 6467.7, "test", None
        bench.attach_collectives(cx)
        legs = {}
        if fail_rank is None:
            # the primary leg's multi-rank path too (all-gather -> planes -> rows in four blocks -> page-locked out)
            bench.N_DIST_1GPU = 64
            res, extra = bench.bench_dist(cx)
            legs["_primary"] = {"n": res["config"]["n_sketches"], "value": res["value"], "e2e": res["e2e"]["value"],
                                "breakdown": sorted(res["details"]["step_breakdown_ms"]), "has_rows": extra["gpu_rows_np"] is not None}
            bench.run_legs(cx, legs)
        else:
            # the other rank would wait in a collective for ever: bound the attempt as bench.py's timer does
            import threading
            done = threading.Event()
            t = threading.Thread(target=lambda: (bench.run_legs(cx, legs), done.set()), daemon=True)
            t.start()
            t.join(timeout=10)
            legs["_finished"] = done.is_set()
        q.put((rank, legs, _FakeCuda.current[0]))
    finally:
        if fail_rank is None:
            dist.destroy_process_group()


def _run(world, fail_rank):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fail_rank, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(world):
        rank, legs, dev = q.get(timeout=180)
        res[rank] = (legs, dev)
    for pr in procs:
        pr.join(timeout=30)
        if pr.is_alive():
            pr.terminate()
    return res


def test_c4_c5_legs_world2_control_flow():
    world = 2
    res = _run(world, None)
    for rank in range(world):
        legs, dev = res[rank]
        assert dev == rank, "a leg left another device current"
        prim = legs.pop("_primary")
        assert prim["n"] == 91 and prim["value"] > 0 and prim["e2e"] > 0 and prim["has_rows"] == (rank == 0)
        assert prim["breakdown"] == ["all_pairs_kernel", "allgather", "planes+cardinalities"]
        assert set(legs) == {"c4", "c5"}
        for name in ("c4", "c5"):
            leg = legs[name]
            assert "failed" not in leg, (rank, name, leg)
            assert leg["n_gpus"] == world and leg["value"] > 0 and leg["ms_per_step"] > 0
            assert {"config", "details", "e2e"} <= set(leg)
        c4, c5 = legs["c4"], legs["c5"]
        n = 96
        assert c4["config"]["n_sketches"] == n and c4["details"]["pairs_of_this_rank"] > 0
        assert {"allgather", "planes+cardinalities", "all_pairs_kernel"} <= set(c4["details"]["step_breakdown_ms"])
        assert c5["config"]["genomes_per_gpu"] == 6 and c5["details"]["sketch_kmers_per_s_whole_job"] > 0
    # the two ranks' block rows tile the triangle
    rows = sorted(res[r][0]["c4"]["details"]["rows_of_this_rank"] for r in range(world))
    assert rows[0][0] == 0 and rows[0][1] == rows[1][0] and rows[1][1] == 96
    assert sum(res[r][0]["c4"]["details"]["pairs_of_this_rank"] for r in range(world)) == 96 * 95 // 2


def test_a_rank_failing_inside_a_leg_is_reported_not_silently_mixed():
    """The failing rank catches its exception, records the failure and moves on to the status exchange; the healthy rank is left
    in the leg's collective (that is what bench.py's timer is for) — it must NOT come back with a result that looks fine."""
    res = _run(2, 1)
    legs1, _ = res[1]
    assert "failed" in legs1.get("c4", {"failed": "never returned"}) or not legs1.get("_finished", False)
    legs0, _ = res[0]
    if legs0.get("_finished"):
        assert "failed" in legs0["c4"]


# ---- the same legs through bench.py's process isolation: parent ranks (gloo) -> run_legs_isolated -> child ranks -> run_legs_child
def _stub_capi():
    from dashing_b200 import capi as real_capi

    class FakeCapi:
        ERTL_MLE, ERTL_JOINT_MLE, JI, ORDER_ROW_FIRST = real_capi.ERTL_MLE, real_capi.ERTL_JOINT_MLE, real_capi.JI, real_capi.ORDER_ROW_FIRST
        dist_params = staticmethod(real_capi.dist_params)
        DistPlan, PackedGenomes = _Plan, _Packed
        kernel_launches = staticmethod(lambda: 0)
        pinned_empty = staticmethod(lambda nbytes: np.zeros(nbytes, dtype=np.uint8))
    return FakeCapi


def child_main():
    """What `bench.py --legs-child` does, with the device stubbed and gloo instead of NCCL (run as a script by the test below)."""
    sys.path.insert(0, ROOT)
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    bench.C45_WORLD, bench.C4_N, bench.C5_GENOMES, bench.GENOME_LEN = int(os.environ["WORLD_SIZE"]), 96, 12, 3000
    bench.usable_cores = lambda: 2
    if os.environ.get("LEGS_TEST_MODE") == "hang" and os.environ["RANK"] == "1":
        time.sleep(120)                        # a rank that never shows up: the others wait in the rendezvous

    class Args:
        emulate_world, emulate_rank, no_cpu_baseline, only, steps, warmup = 0, 0, True, "", 2, 1
    return bench.run_legs_child(Args, backend="gloo", stubs=(_FakeTorch(), _stub_capi(), torch.device("cpu")))


def _parent(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      LEGS_TEST_MODE=mode)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Args:
        no_cpu_baseline, only = True, ""
    cx = bench.Ctx()
    cx.args, cx.rank, cx.world = Args, rank, world
    cmd = [sys.executable, "-c", "import sys; sys.path.insert(0, %r); import test_bench_legs_gloo as t; sys.exit(t.child_main())" % os.path.join(ROOT, "tests")]
    t0 = time.perf_counter()
    legs = bench.run_legs_isolated(cx, 15 if mode == "hang" else 150, cmd=cmd)
    dist.barrier()                              # the parents' own process group is alive and well afterwards
    q.put((rank, legs, time.perf_counter() - t0))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ok", "hang"])
def test_legs_in_child_processes_world2(mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_parent, args=(r, world, port, mode, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(world):
        rank, legs, dt = q.get(timeout=240)
        res[rank] = (legs, dt)
    for pr in procs:
        pr.join(timeout=30)
    assert res[1][0] == {}
    legs = res[0][0]
    assert set(legs) == {"c4", "c5"}
    if mode == "ok":
        assert all("failed" not in v and v["value"] > 0 and v["n_gpus"] == world for v in legs.values()), legs
    else:
        assert all("failed" in v for v in legs.values()) and res[0][1] < 60      # bounded, and the parents went on


def test_primary_and_jmle_legs_world1_stubbed():
    """The N=1 branches of bench_dist / bench_jmle (host-pointer e2e call, roofline object, traffic lookup) with the device stubbed."""
    sys.path.insert(0, ROOT)
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from dashing_b200 import multigpu
    bench.N_DIST_1GPU, bench.N_JMLE = 80, 40
    bench.usable_cores = lambda: 2
    FakeCapi = _stub_capi()

    def dist_symmetric(regs, p, k=31, result_type=1, device=0, out=None, **kw):
        n = regs.size >> p
        out[: n * (n - 1) // 2] = np.arange(n * (n - 1) // 2, dtype=np.float32)
        return out
    FakeCapi.dist_symmetric = staticmethod(dist_symmetric)

    class Args:
        emulate_world, emulate_rank, no_cpu_baseline, only, steps, warmup = 0, 0, True, "", 2, 1
    cx = bench.Ctx()
    cx.torch, cx.dist, cx.capi, cx.multigpu, cx.args = _FakeTorch(), dist, FakeCapi, multigpu, Args
    cx.world, cx.rank, cx.local_rank, cx.dev, cx.stream = 1, 0, 0, torch.device("cpu"), 0
    cx.peak, cx.peak_src, cx.sampler = 6467.7, "test", None
    bench.attach_collectives(cx)
    res, extra = bench.bench_dist(cx)
    assert res["config"] == bench.dist_config(80, 1) and res["value"] > 0 and res["e2e"]["value"] > 0
    assert res["roofline"]["traffic"] is None                                      # the ncu capture is for n = 10,000 only
    pm = res["roofline"]["int_bound"]["pipe_model"]
    assert 1.6e9 < pm["ceiling_pairs_per_s_per_gpu"] < 1.7e9 and pm["frac_of_ceiling"] > 0
    assert extra["regs_np"].shape == (80, 1 << 14) and extra["gpu_rows_np"].size == 80 * 79 // 2
    assert np.array_equal(extra["regs_np"], bench.host_registers(bench.SEED_DIST, 0, 80, 14))     # the bytes the reference arm draws
    j = bench.bench_jmle(cx)
    assert j["config"]["n_sketches"] == 40 and j["config"]["p"] == 16 and j["value"] > 0
    import json
    json.dumps(res); json.dumps(j)                                                  # everything in the line must serialise
