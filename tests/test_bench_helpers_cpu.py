"""CPU tier: the parts of bench.py that do not need a device — both arms print the same `config`, the parity object follows
tests/parity.py, the traffic table only answers for the problem size it was captured on, the shardable generator is shard-invariant."""
import importlib.util
import os

import numpy as np

import parity
from dashing_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_both_arms_share_one_config_object():
    for world in (1, 2, 4, 8):
        n = bench.dist_n_for(world)
        a, b = bench.dist_config(n, world), bench.dist_config(n, world)
        assert a == b and a["n_sketches"] == n and str(n) in a["workload"] and "identical bytes" in a["generator"]
    assert bench.dist_n_for(1) == 10_000 and bench.dist_n_for(8) == 28_284
    assert bench.sketch_config(8)["workload"].startswith("sketch 8000 x 5000000 bp")


def test_parity_stats_follow_the_test_policy():
    rng = np.random.default_rng(0)
    want = rng.random(10_000).astype(np.float32)
    got = want.copy()
    got[7] = np.nextafter(got[7], np.float32(2))           # one float ulp: 6e-8 relative, inside 1e-6
    got[11] = want[11] * np.float32(1 + 3e-6)              # outside
    want[13] = 2e-10; got[13] = 0.0                        # residue of the reference: absolute scale
    st = bench.parity_stats(got, want)
    assert st["pairs_compared"] == 10_000 and st["n_over_1e-6"] == 1 and st["first_bad"][0]["idx"] == 11
    err = parity.rel_err(got, want)
    assert abs(st["max_rel_err"] - float(err.max())) < 1e-12
    assert bench.parity_stats(want, want)["n_bit_identical"] == 10_000


def test_traffic_table_answers_only_for_its_problem_size():
    assert bench.load_traffic("dist_kernel", 10_000, 1) > 1e9
    assert bench.load_traffic("dist_kernel", 28_284, 8) is None and bench.load_traffic("dist_kernel", 10_000, 2) is None
    assert bench.load_traffic("no_such_kernel", 10_000, 1) is None


def test_rows_for_pairs_and_triangle_offsets():
    n = 1000
    assert bench.tri(n, 0) == 0 and bench.tri(n, n) == n * (n - 1) // 2
    r = bench.rows_for_pairs(n, 10, 5000)
    assert bench.tri(n, 10 + r) - bench.tri(n, 10) >= 5000 > bench.tri(n, 10 + r - 1) - bench.tri(n, 10)


def test_register_generator_is_shard_invariant():
    """Rank r of the multi-process bench and the reference arm must draw identical bytes for the rows they share."""
    full = synth.registers_block(2026, 0, 100, 10)
    for start, count in ((0, 100), (3, 40), (16, 16), (31, 2), (99, 1)):
        assert np.array_equal(synth.registers_block(2026, start, count, 10), full[start:start + count])
        assert np.array_equal(synth.registers_block_mt(2026, start, count, 10, threads=3), full[start:start + count])
    assert not np.array_equal(synth.registers_block(2027, 0, 4, 10), full[:4])
    # correlated groups: members of one group share far more registers than members of different groups
    same = (full[1] == full[2]).mean()
    diff = (full[1] == full[17]).mean()
    assert same > diff + 0.1


def _cx(rank=0, only=""):
    class Args:
        no_cpu_baseline = True
    Args.only = only
    cx = bench.Ctx()
    cx.args, cx.rank, cx.world = Args, rank, 8
    return cx


def test_isolated_legs_survive_whatever_the_child_does(tmp_path, capfd):
    """run_legs_isolated (bench.py at N=8): the c4 / c5 legs run in a child process per rank; the parent — which owns the primary
    result — must come back with results or failure notes whether the child succeeds, crashes, hangs or cannot be started."""
    import sys
    import time
    py = sys.executable
    ok = [py, "-c", "import json, os; json.dump({'c4': {'value': 1.0}, 'c5': {'value': 2.0}}, open(os.environ['DB200_LEGS_OUT'], 'w')); print('child chatter on stdout')"]
    legs = bench.run_legs_isolated(_cx(), 30, cmd=ok)
    assert legs == {"c4": {"value": 1.0}, "c5": {"value": 2.0}}
    out, err = capfd.readouterr()
    assert "child chatter" not in out, "the child's stdout leaked into the result stream"
    assert bench.run_legs_isolated(_cx(rank=3), 30, cmd=ok) == {}                      # only rank 0 reports
    crash = [py, "-c", "import os, signal; os.kill(os.getpid(), signal.SIGSEGV)"]
    legs = bench.run_legs_isolated(_cx(), 30, cmd=crash)
    assert set(legs) == {"c4", "c5"} and all("failed" in v and "status" in v["failed"] for v in legs.values())
    t0 = time.perf_counter()
    legs = bench.run_legs_isolated(_cx(), 2, cmd=[py, "-c", "import time; time.sleep(60)"])
    assert time.perf_counter() - t0 < 20 and all("killed" in v["failed"] for v in legs.values())
    half = [py, "-c", "import json, os; json.dump({'c4': {'value': 1.0}}, open(os.environ['DB200_LEGS_OUT'], 'w')); raise SystemExit(3)"]
    legs = bench.run_legs_isolated(_cx(), 30, cmd=half)
    assert legs["c4"] == {"value": 1.0} and "failed" in legs["c5"]                     # what finished is kept
    legs = bench.run_legs_isolated(_cx(only="c4"), 30, cmd=["/nonexistent/interpreter"])
    assert set(legs) == {"c4"} and "could not run" in legs["c4"]["failed"]


def test_legs_child_port_is_off_the_parents(monkeypatch):
    monkeypatch.setenv("MASTER_PORT", "29500")
    assert bench.legs_child_port() == 29601
    monkeypatch.setenv("MASTER_PORT", "65500")
    assert bench.legs_child_port() == 65399
