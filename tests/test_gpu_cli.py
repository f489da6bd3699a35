"""GPU end-to-end through the host layer's CLI (`dashing_b200 sketch|dist`): FASTA files in, .hll / sizes / distance
files out, compared with what the reference's own drivers wrote for the same inputs (tests/golden/cli.npz; and, where
oracle/_ref travelled, with the reference run live).  Registers bit-exact; text outputs same structure, numbers within
2e-5 (the reference prints 6 significant digits); binary outputs within 1e-6."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import hostlib
from parity import assert_close, assert_knn_close

pytestmark = pytest.mark.gpu

FMT_FLAG = {0: [], 1: ["-b"], 2: ["-U"], 3: ["-T"]}
RTYPE_FLAG = {0: ["-M"], 1: [], 2: ["--sizes"], 3: ["-l"], 4: ["--full-containment-dist"], 5: ["--containment-index"],
              6: ["--containment-dist"], 7: ["--symmetric-containment-index"], 8: ["--symmetric-containment-dist"]}


def run_cli(cwd, *args):
    r = subprocess.run([hostlib.CLI, *args], cwd=cwd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, f"dashing_b200 {' '.join(args)} failed:\n{r.stderr}"
    return r


@pytest.fixture(scope="module")
def cli(golden_dir):
    return np.load(os.path.join(golden_dir, "cli.npz"))


def test_cli_sketch_writes_reference_hll_files(gpu, cli, tmp_path):
    names = hostlib.materialise_inputs(cli, str(tmp_path))
    os.makedirs(tmp_path / "sk")
    run_cli(str(tmp_path), "sketch", "-k31", "-S10", "-p2", "-P", "sk", "--avoid-sorting", *names)
    for n in names:
        hp = tmp_path / str(cli["hllname_" + n])
        assert hp.exists(), hp
        assert gzip.open(hp, "rb").read() == cli["hll_" + n].tobytes(), f"decompressed .hll payload of {n} differs"
    # --skip-cached leaves existing files alone
    before = {n: os.path.getmtime(tmp_path / str(cli["hllname_" + n])) for n in names}
    run_cli(str(tmp_path), "sketch", "-k31", "-S10", "-P", "sk", "--avoid-sorting", "--skip-cached", *names)
    assert before == {n: os.path.getmtime(tmp_path / str(cli["hllname_" + n])) for n in names}


def test_cli_dist_all_formats(gpu, cli, tmp_path):
    names = hostlib.materialise_inputs(cli, str(tmp_path))
    for run in [str(r) for r in cli["runs"]]:
        kw = json.loads(str(cli[run + "_kw"]))
        nq = kw.get("nq", 0)
        args = ["dist", f"-k{kw.get('k', 31)}", f"-S{kw.get('p', 10)}", "-p2", "--avoid-sorting", "-o", "sizes.txt", "-O", "dist.out"]
        args += FMT_FLAG[kw.get("emit_fmt", 0)] + RTYPE_FLAG[kw.get("rtype", 1)]
        if kw.get("jestim") == 3:
            args.append("-J")
        if kw.get("estim") == 0:
            args.append("-E")
        if kw.get("nneighbors"):
            args += ["--nearest-neighbors", str(kw["nneighbors"])]
        if nq:
            (tmp_path / "refs.txt").write_text("\n".join(names[:-nq]) + "\n")
            (tmp_path / "qry.txt").write_text("\n".join(names[-nq:]) + "\n")
            args += ["-F", "refs.txt", "-Q", "qry.txt"]
        else:
            args += names
        run_cli(str(tmp_path), *args)
        assert (tmp_path / "sizes.txt").read_bytes() == cli[run + "_sizes"].tobytes(), f"{run}: sizes file"
        got, want = (tmp_path / "dist.out").read_bytes(), cli[run + "_dist"].tobytes()
        if kw.get("nneighbors") and kw.get("emit_fmt") == 1:
            dt = np.dtype([("value", np.float32), ("index", np.uint32)])
            assert got[:8] == want[:8] and len(got) == len(want), run
            nn = int(np.frombuffer(got[4:8], np.uint32)[0])
            assert_knn_close(np.frombuffer(got[8:], dt).reshape(-1, nn), np.frombuffer(want[8:], dt).reshape(-1, nn), what=run)
        elif kw.get("emit_fmt") == 1:
            hdr = 0 if nq else 9
            assert got[:hdr] == want[:hdr] and len(got) == len(want), run
            assert_close(np.frombuffer(got[hdr:], np.float32), np.frombuffer(want[hdr:], np.float32), what=run)
            if not nq:
                assert (tmp_path / "dist.out.labels").read_bytes() == cli[run + "_labels"].tobytes()
        else:
            hostlib.assert_text_matches(got, want, what=run)


def test_cli_presketched_and_cache(gpu, cli, tmp_path):
    names = hostlib.materialise_inputs(cli, str(tmp_path))
    os.makedirs(tmp_path / "sk")
    # -W (cache sketches) writes the same .hll files `sketch` would ...
    run_cli(str(tmp_path), "dist", "-k31", "-S10", "-M", "-W", "-P", "sk", "--avoid-sorting", "-o", "s1.txt", "-O", "d1.txt", *names)
    hpaths = [str(cli["hllname_" + n]) for n in names]
    for n, hp in zip(names, hpaths):
        assert gzip.open(tmp_path / hp, "rb").read() == cli["hll_" + n].tobytes()
    # ... and --presketched on them reproduces the reference's presketched output
    run_cli(str(tmp_path), "dist", "-k31", "-S10", "-M", "--presketched", "-o", "s2.txt", "-O", "d2.txt", *hpaths)
    assert (tmp_path / "s2.txt").read_bytes() == cli["presketched_tsv_mash_sizes"].tobytes()
    hostlib.assert_text_matches((tmp_path / "d2.txt").read_bytes(), cli["presketched_tsv_mash_dist"].tobytes(), what="presketched")


def test_cli_live_against_reference_drivers(gpu, ref, cli, tmp_path):
    """Where oracle/_ref travelled: the reference's dist_sketch_and_cmp<hll_t> run live on fresh inputs, p=14."""
    from dashing_b200 import synth
    from oracle.make_golden import write_fasta
    gs = synth.genomes(31337, 5, 120_000, group=5)
    names = []
    for i, g in enumerate(gs):
        n = f"x{i}.fa"
        write_fasta(str(tmp_path / n), [g[:50_000].tobytes(), g[50_000:].tobytes()], width=80, gz=False)
        names.append(n)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref.cli_dist(names, "ref_sizes.txt", "ref_dist.txt", k=21, p=14, rtype=0, emit_fmt=0, nthreads=2)
    finally:
        os.chdir(cwd)
    run_cli(str(tmp_path), "dist", "-k21", "-S14", "-M", "--avoid-sorting", "-o", "sizes.txt", "-O", "dist.txt", *names)
    assert (tmp_path / "sizes.txt").read_bytes() == (tmp_path / "ref_sizes.txt").read_bytes()
    hostlib.assert_text_matches((tmp_path / "dist.txt").read_bytes(), (tmp_path / "ref_dist.txt").read_bytes(), what="live dist")


def test_cli_declines_out_of_scope_flags(gpu, cli, tmp_path):
    names = hostlib.materialise_inputs(cli, str(tmp_path))
    for flags in (["-s", "1,1,1"], ["-w", "50"], ["--countmin"], ["-8"]):
        r = subprocess.run([hostlib.CLI, "dist", "-k31", *flags, *names], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 1 and "outside the B200 engine" in r.stderr
    r = subprocess.run([hostlib.CLI, "dist", "-k33", *names], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "k must be <= 32" in r.stderr


def test_cli_multi_member_gzip_takes_the_reread_path(gpu, cli, tmp_path):
    """A bgzip-style multi-member .gz has an ISIZE trailer for its last member only, so the window the batch driver
    reserves is too small; the genome is then re-read through the growing-string path.  Same registers as the plain file,
    and the neighbouring genomes of the batch are unaffected."""
    names = hostlib.materialise_inputs(cli, str(tmp_path))
    raw = (tmp_path / "a.fa").read_bytes()
    cut = raw.index(b"\n", len(raw) // 2) + 1
    (tmp_path / "amm.fa.gz").write_bytes(gzip.compress(raw[:cut]) + gzip.compress(raw[cut:]))
    os.makedirs(tmp_path / "sk")
    run_cli(str(tmp_path), "sketch", "-k31", "-S10", "-p3", "-P", "sk", "--avoid-sorting", "b.fa", "amm.fa.gz", "d.fa")
    got = gzip.open(tmp_path / "sk" / "amm.fa.gz.w.31.spacing.10.hll", "rb").read()
    assert got[28:] == cli["hll_a.fa"].tobytes()[28:]
    for n in ("b.fa", "d.fa"):
        assert gzip.open(tmp_path / str(cli["hllname_" + n]), "rb").read() == cli["hll_" + n].tobytes()


def test_cli_device_all(gpu, cli, tmp_path):
    """--device all: sketching, sizes and the streamed binary / text matrices sharded over every (logical) device give the files
    the single-device run gives (on a one-GPU box DB200_VIRTUAL_DEVICES makes three logical devices out of it)."""
    names = hostlib.materialise_inputs(cli, str(tmp_path))
    env = dict(os.environ)
    if gpu.device_count() < 2:
        env["DB200_VIRTUAL_DEVICES"] = "3"
    for run, flags in (("bin_mash", ["-b", "-M"]), ("tsv_ji", [])):
        r = subprocess.run([hostlib.CLI, "dist", "-k31", "-S10", "-p2", "--avoid-sorting", "--device", "all", "-o", "sizes.txt", "-O", "dist.out", *flags, *names],
                           cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr
        assert (tmp_path / "sizes.txt").read_bytes() == cli[run + "_sizes"].tobytes()
        got, want = (tmp_path / "dist.out").read_bytes(), cli[run + "_dist"].tobytes()
        if run == "bin_mash":
            assert got[:9] == want[:9] and len(got) == len(want)
            assert_close(np.frombuffer(got[9:], np.float32), np.frombuffer(want[9:], np.float32), what=run)
            assert (tmp_path / "dist.out.labels").read_bytes() == cli[run + "_labels"].tobytes()
        else:
            hostlib.assert_text_matches(got, want, what=run)
    os.makedirs(tmp_path / "sk")
    r = subprocess.run([hostlib.CLI, "sketch", "-k31", "-S10", "-p2", "-P", "sk", "--avoid-sorting", "--device", "all", *names],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    for n in names:
        assert gzip.open(tmp_path / str(cli["hllname_" + n]), "rb").read() == cli["hll_" + n].tobytes()
