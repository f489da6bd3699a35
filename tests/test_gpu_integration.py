"""INTEGRATION.md, compiled: the reference's own sketch_core<hll_t>, dist_sketch_and_cmp<hll_t>, dist_loop and partdist_loop
built from its unmodified headers + oracle/integration.patch (-DDASHING_B200), linked against libdashing_b200.so
(oracle/_ref/libdashing_ref_patched.so, `make -C oracle patched`), against the SAME driver built from the unpatched headers.
Output files must agree byte for byte: .hll payloads (S1), the sizes file (S2), the TSV / PHYLIP / binary matrices and the
-Q rectangle (S3).  VERDICT r01 "What's missing" #2."""
import gzip
import os

import numpy as np
import pytest

from dashing_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def builds(gpu):
    from oracle import oracle as O
    if not (O.ref_available() and O.patched_available()):
        pytest.skip("oracle/_ref (reference + patched build) did not travel: run `make -C oracle ref patched` where /root/reference exists")
    return O.ref(), O.ref_patched()


@pytest.fixture(scope="module")
def fasta_dir(tmp_path_factory):
    """Six related genomes as FASTA: plain and gzipped, multi-record, 60/70/80-column lines, N runs, lower case, IUPAC codes."""
    from oracle.make_golden import write_fasta
    d = tmp_path_factory.mktemp("integration")
    rng = np.random.default_rng(20261017)
    gs = synth.genomes(4711, 6, 300_000, group=6)
    gs[1] = synth.sprinkle(rng, gs[1], n_runs=6)
    names = []
    for i, g in enumerate(gs):
        recs = [g.tobytes()] if i % 2 == 0 else [g[:100_000].tobytes(), g[100_000:100_017].tobytes(), g[100_017:].tobytes()]
        name = f"g{i}.fa" + (".gz" if i == 3 else "")
        write_fasta(str(d / name), recs, width=(60, 70, 80)[i % 3], gz=(i == 3))
        names.append(name)
    for sub in ("ref_out", "gpu_out", "cache"):
        (d / sub).mkdir()
    return d, names


def in_dir(d, fn):
    cwd = os.getcwd()
    os.chdir(d)
    try:
        return fn()
    finally:
        os.chdir(cwd)


def payload(path):
    return gzip.open(path, "rb").read()


def test_patched_sketch_core_writes_identical_hll_files(builds, fasta_dir, gpu):
    ref, pat = builds
    d, names = fasta_dir
    for k, p, canon in ((31, 14, True), (21, 10, False)):
        l0 = gpu.kernel_launches()
        in_dir(d, lambda: ref.cli_sketch(names, k=k, p=p, canon=canon, nthreads=2, prefix="ref_out"))
        assert gpu.kernel_launches() == l0, "the unpatched reference must not touch the GPU"
        in_dir(d, lambda: pat.cli_sketch(names, k=k, p=p, canon=canon, nthreads=2, prefix="gpu_out"))
        assert gpu.kernel_launches() > l0, "the patched build did not go through libdashing_b200"
        for n in names:
            f = ref.make_fname(n, p, k, k, k, prefix="ref_out")
            g = pat.make_fname(n, p, k, k, k, prefix="gpu_out")
            assert payload(d / f) == payload(d / g), (n, k, p)


@pytest.mark.parametrize("jestim,rtype", [(2, 1), (2, 0), (3, 0), (2, 2)])
def test_patched_dist_outputs_identical(builds, fasta_dir, gpu, jestim, rtype):
    ref, pat = builds
    d, names = fasta_dir
    k, p = 31, 12
    # emit_fmt: 0 = upper-triangular TSV, 2 = PHYLIP, 1 = binary distance matrix (mmap'd DistanceMatrix path)
    for fmt in (0, 2, 1):
        tag = f"j{jestim}_r{rtype}_f{fmt}"
        in_dir(d, lambda: ref.cli_dist(names, f"ref_sizes_{tag}.txt", f"ref_dist_{tag}.out", k=k, p=p, jestim=jestim, rtype=rtype, emit_fmt=fmt, nthreads=2))
        l0 = gpu.kernel_launches()
        in_dir(d, lambda: pat.cli_dist(names, f"gpu_sizes_{tag}.txt", f"gpu_dist_{tag}.out", k=k, p=p, jestim=jestim, rtype=rtype, emit_fmt=fmt, nthreads=2))
        assert gpu.kernel_launches() > l0
        assert (d / f"gpu_sizes_{tag}.txt").read_bytes() == (d / f"ref_sizes_{tag}.txt").read_bytes(), tag
        assert (d / f"gpu_dist_{tag}.out").read_bytes() == (d / f"ref_dist_{tag}.out").read_bytes(), tag


def test_patched_rect_and_presketched(builds, fasta_dir, gpu):
    ref, pat = builds
    d, names = fasta_dir
    k, p = 31, 12
    # -Q form: the last two paths are queries (partdist_loop), containment index is legal there
    for rtype, fmt in ((5, 3), (0, 1)):
        in_dir(d, lambda: ref.cli_dist(names, "ref_qs.txt", "ref_q.out", nq=2, k=k, p=p, rtype=rtype, emit_fmt=fmt, nthreads=2))
        in_dir(d, lambda: pat.cli_dist(names, "gpu_qs.txt", "gpu_q.out", nq=2, k=k, p=p, rtype=rtype, emit_fmt=fmt, nthreads=2))
        assert (d / "gpu_q.out").read_bytes() == (d / "ref_q.out").read_bytes(), (rtype, fmt)
        assert (d / "gpu_qs.txt").read_bytes() == (d / "ref_qs.txt").read_bytes()
    # --cache-sketches, then --presketched from the files the PATCHED build wrote: loaded sketches carry their cached value_
    in_dir(d, lambda: pat.cli_dist(names, "gpu_cs.txt", "gpu_c.out", k=k, p=p, rtype=0, emit_fmt=0, nthreads=2, cache=True, prefix="cache"))
    hll = [pat.make_fname(n, p, k, k, k, prefix="cache") for n in names]
    in_dir(d, lambda: ref.cli_dist(hll, "ref_ps.txt", "ref_p.out", k=k, p=p, rtype=0, emit_fmt=0, nthreads=2, presketched=True))
    in_dir(d, lambda: pat.cli_dist(hll, "gpu_ps.txt", "gpu_p.out", k=k, p=p, rtype=0, emit_fmt=0, nthreads=2, presketched=True))
    assert (d / "gpu_p.out").read_bytes() == (d / "ref_p.out").read_bytes()
    assert (d / "gpu_ps.txt").read_bytes() == (d / "ref_ps.txt").read_bytes()


def test_patched_build_falls_back_when_disabled(builds, fasta_dir, gpu, monkeypatch):
    """DASHING_GPU=0: the eligibility guard of INTEGRATION.md §0 keeps the reference's own code."""
    ref, pat = builds
    d, names = fasta_dir
    monkeypatch.setenv("DASHING_GPU", "0")
    l0 = gpu.kernel_launches()
    in_dir(d, lambda: pat.cli_dist(names, "off_sizes.txt", "off_dist.out", k=21, p=10, rtype=1, emit_fmt=0, nthreads=2))
    assert gpu.kernel_launches() == l0
    monkeypatch.delenv("DASHING_GPU")
    in_dir(d, lambda: ref.cli_dist(names, "ref_off_sizes.txt", "ref_off_dist.out", k=21, p=10, rtype=1, emit_fmt=0, nthreads=2))
    assert (d / "off_dist.out").read_bytes() == (d / "ref_off_dist.out").read_bytes()
