"""CPU: what the CLI does before (or without) touching a GPU — the reference's option rules and error texts, the purely
host-side subcommands (`view`, `fold` to the same size of a sketch that carries its value), and the loud failure of every
compute path when no CUDA device exists (no CPU fallback)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import hostlib

pytestmark = pytest.mark.skipif(not os.path.exists(hostlib.CLI), reason="host layer not built (run __graft_entry__.build())")


def cli(cwd, *args):
    return subprocess.run([hostlib.CLI, *args], cwd=cwd, capture_output=True, timeout=120)


@pytest.fixture(scope="module")
def sub(golden_dir):
    return np.load(os.path.join(golden_dir, "subcmd.npz"))


def test_option_rules_and_declined_paths(tmp_path):
    (tmp_path / "x.fa").write_bytes(b">x\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n")
    r = cli(tmp_path, "dist", "-k33", "x.fa")
    assert r.returncode == 1 and b"k must be <= 32 for non-rolling hashes." in r.stderr          # src/distmain.cpp:101-102
    for flags in (["-s", "1,1,1"], ["-w", "50"], ["--countmin"], ["-8"], ["--use-nthash"], ["--wj"], ["--use-bloom-filter"]):
        r = cli(tmp_path, "dist", "-k31", *flags, "x.fa")
        assert r.returncode == 1 and b"outside the B200 engine" in r.stderr, flags
    # option order must not matter: `-w 30 -k 21` is windowed (30 > 21) although 30 <= the default k = 31
    for sub_cmd in ("dist", "sketch", "card"):
        r = cli(tmp_path, sub_cmd, "-w", "30", "-k", "21", "x.fa")
        assert r.returncode == 1 and b"outside the B200 engine" in r.stderr, sub_cmd
    r = cli(tmp_path, "dist", "-w", "20", "-k31", "--containment-index")                       # no paths at all
    assert r.returncode == 1 and b"No paths" in r.stderr
    r = cli(tmp_path, "sketch", "-k31", "--defer-hll", "x.fa")
    assert r.returncode == 1 and b"b-bit minhash" in r.stderr
    r = cli(tmp_path, "sketch_by_seq", "-k31", "-o", "o.bin", "x.fa")
    assert r.returncode == 1 and b"--defer-hll" in r.stderr
    r = cli(tmp_path, "union")
    assert r.returncode == 1 and b"require >= 1 paths" in r.stderr
    r = cli(tmp_path, "dist_by_seq", "x.bin")                                                   # -n is mandatory (src/distbyseq.cpp:100)
    assert r.returncode == 1 and b"Usage: dist_by_seq" in r.stderr
    r = cli(tmp_path, "dist", "--nearest-neighbors", "0", "x.fa")
    assert r.returncode == 1 and b"positive count" in r.stderr
    r = cli(tmp_path, "panel", "x.fa")
    assert r.returncode == 1 and b"outside the B200 engine" in r.stderr
    r = cli(tmp_path)
    assert r.returncode == 1 and b"usage" in r.stderr


def test_view_and_fold_without_a_device(sub, tmp_path):
    with gzip.open(tmp_path / "f8.hll", "wb") as f:
        f.write(sub["fold_a_8"].tobytes())
    r = cli(tmp_path, "view", "f8.hll")
    assert r.returncode == 0 and r.stdout == sub["view_f8"].tobytes()                           # hll_t::printf, hll.h:888-893
    r = cli(tmp_path, "view", "f8.hll", "f8.hll")
    assert r.stdout == sub["view_f8"].tobytes() * 2
    # fold to a LARGER size: the reference's error, raised before any device work (hll.h:907-909)
    r = cli(tmp_path, "fold", "-p", "9", "-o", "big.hll", "f8.hll")
    assert r.returncode == 1 and b"Can't compress to a larger size. Current: 8. Requested new size: 9" in r.stderr
    # fold to the SAME size of a sketch that carries its cardinality: a copy, value included
    with gzip.open(tmp_path / "u3.hll", "wb") as f:
        f.write(sub["union3"].tobytes())
    r = cli(tmp_path, "fold", "-p", "12", "-o", "same.hll", "u3.hll")
    assert r.returncode == 0, r.stderr
    assert gzip.open(tmp_path / "same.hll").read() == sub["union3"].tobytes()
    r = cli(tmp_path, "fold", "-p", "5", "missing.hll")
    assert r.returncode == 1 and b"Could not open file" in r.stderr


def test_compute_paths_fail_loudly_without_a_device(capi, sub, tmp_path):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is visible; this test documents the no-GPU behaviour")
    (tmp_path / "x.fa").write_bytes(b">x\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n")
    with gzip.open(tmp_path / "a.hll", "wb") as f:
        f.write(sub["hll_a.fa"].tobytes())
    for args in (["sketch", "-k21", "-S10", "x.fa"], ["dist", "-k21", "-S10", "x.fa", "x.fa"], ["hll", "-k21", "-S12", "x.fa"],
                 ["union", "-o", "u.hll", "a.hll"], ["fold", "-p", "8", "-o", "f.hll", "a.hll"], ["card", "-k21", "-S10", "x.fa"],
                 ["dist", "--presketched", "-S12", "a.hll", "a.hll"]):
        r = cli(tmp_path, *args)
        assert r.returncode == 1 and b"no usable CUDA device" in r.stderr and b"no CPU fallback" in r.stderr, (args, r.stderr)
