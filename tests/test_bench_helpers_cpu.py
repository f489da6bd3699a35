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
