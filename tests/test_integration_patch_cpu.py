"""CPU tier for the reference-side binding (oracle/integration.patch): the patch applies cleanly to the reference
headers where they are available, and the patched build — loaded on a box without a usable GPU path — keeps the reference's own code
behind its eligibility guard and writes the same files as the unpatched build.  (The GPU tier, tests/test_gpu_integration.py,
compares the two builds with the hot paths actually running through libdashing_b200.)"""
import os
import shutil
import subprocess

import pytest

from dashing_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_patch_applies_to_the_reference_headers(tmp_path):
    if not os.path.isdir(REF) or not shutil.which("patch"):
        pytest.skip("needs the reference tree and patch(1) (build container only)")
    (tmp_path / "src").mkdir()
    for h in ("sketch_and_cmp.h", "dashing.h"):
        shutil.copy(os.path.join(REF, "src", h), tmp_path / "src" / h)
        os.chmod(tmp_path / "src" / h, 0o644)
    r = subprocess.run(["patch", "-p1", "--dry-run", "-d", str(tmp_path)], stdin=open(os.path.join(ROOT, "oracle", "integration.patch")),
                       capture_output=True, text=True)
    assert r.returncode == 0 and "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout + r.stderr
    txt = open(os.path.join(ROOT, "oracle", "integration.patch")).read()
    # every added line sits under the DASHING_B200 guard: an unpatched build (-UDASHING_B200) is the reference, token for token
    assert txt.count("#ifdef DASHING_B200") == txt.count("+#endif") >= 10
    for sym in ("db200_sketcher_add_record", "db200_sketcher_finish", "db200_cardinalities", "db200_dist_symmetric", "db200_dist_rect"):
        assert sym in txt


def test_patched_build_keeps_reference_code_when_gpu_path_is_off(tmp_path, monkeypatch):
    from oracle import oracle as O
    from oracle.make_golden import write_fasta
    if not (O.ref_available() and O.patched_available()):
        pytest.skip("oracle/_ref (reference + patched build) not built")
    monkeypatch.setenv("DASHING_GPU", "0")            # the guard
    ref, pat = O.ref(), O.ref_patched()
    names = []
    for i, g in enumerate(synth.genomes(3, 4, 40_000, group=4)):
        write_fasta(str(tmp_path / f"g{i}.fa"), [g.tobytes()], width=70, gz=(i == 2))
        names.append(f"g{i}.fa")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for fmt, rtype in ((0, 0), (1, 1)):
            ref.cli_dist(names, f"rs{fmt}.txt", f"rd{fmt}.out", k=21, p=10, rtype=rtype, emit_fmt=fmt, nthreads=2)
            pat.cli_dist(names, f"ps{fmt}.txt", f"pd{fmt}.out", k=21, p=10, rtype=rtype, emit_fmt=fmt, nthreads=2)
            assert open(f"rd{fmt}.out", "rb").read() == open(f"pd{fmt}.out", "rb").read()
            assert open(f"rs{fmt}.txt", "rb").read() == open(f"ps{fmt}.txt", "rb").read()
    finally:
        os.chdir(cwd)
