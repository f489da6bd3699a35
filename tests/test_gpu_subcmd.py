"""GPU: SURVEY.md §8(f)3 — db200_union / db200_compress through the C ABI vs the checker, and the CLI subcommands that sit
on them (union, hll, fold, view, card, sketch -o, sketch_by_seq, dist_by_seq, dist --defer-hll) vs tests/golden/subcmd.npz,
which holds what the reference's own mains wrote for the same inputs (oracle/make_golden.py; cross-checked there against
the real `dashing` binary).  Registers / payload headers bit-exact; cardinalities inside payloads 1e-9; printed values
within the reference's printed precision."""
import gzip
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import hostlib
from parity import assert_close
from dashing_b200 import synth

pytestmark = pytest.mark.gpu


def run_cli(cwd, *args, ok=True):
    r = subprocess.run([hostlib.CLI, *args], cwd=cwd, capture_output=True, timeout=300)
    if ok:
        assert r.returncode == 0, f"dashing_b200 {' '.join(args)} failed:\n{r.stderr.decode()}"
    return r


@pytest.fixture(scope="module")
def sub(golden_dir):
    return np.load(os.path.join(golden_dir, "subcmd.npz"))


def assert_payload(got: bytes, want: bytes, what=""):
    """Decompressed .hll payload: header words + p + registers bit-exact, the cached cardinality within 1e-9."""
    assert len(got) == len(want), f"{what}: {len(got)} bytes vs {len(want)}"
    assert got[:20] == want[:20], f"{what}: header {struct.unpack('<5I', got[:20])} vs {struct.unpack('<5I', want[:20])}"
    gv, wv = struct.unpack("<d", got[20:28])[0], struct.unpack("<d", want[20:28])[0]
    assert gv == wv or abs(gv - wv) <= 1e-9 * abs(wv), f"{what}: value {gv!r} vs {wv!r}"
    assert got[28:] == want[28:], f"{what}: registers differ"


def assert_container(got: bytes, want: bytes, what=""):
    assert len(got) == len(want), f"{what}: {len(got)} bytes vs {len(want)}"
    off, i = 0, 0
    while off < len(want):
        p = struct.unpack("<I", want[off + 16:off + 20])[0]
        size = 28 + (1 << p)
        assert_payload(got[off:off + size], want[off:off + size], what=f"{what}[{i}]")
        off += size
        i += 1


def materialise(sub, d):
    names = [str(x) for x in sub["names"]]
    for n in names + ["multi.fa"]:
        with open(os.path.join(d, n), "wb") as f:
            f.write(sub["file_" + n].tobytes())
    os.makedirs(os.path.join(d, "sk"), exist_ok=True)
    hp = [str(x) for x in sub["hllnames"]]
    for n, h in zip(names, hp):
        with gzip.open(os.path.join(d, h), "wb") as f:     # the reference's own sketches, as inputs of union / fold / view
            f.write(sub["hll_" + n].tobytes())
    return names, hp


# ---- C ABI ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [7, 10, 12, 14, 16])
def test_union_vs_checker(gpu, checker, p):
    for n in (1, 2, 3, 37, 300):
        regs = np.concatenate([synth.registers(p * 31 + n, n, p, card=30.0 * (1 << p)), synth.adversarial_registers(1, p)[[0, 3, 5]]])
        assert np.array_equal(gpu.union(regs, p), checker.union(regs, p)), (p, n)
    assert not gpu.union(np.zeros((0, 1 << p), np.uint8), p).any()       # no inputs: the empty sketch


def test_union_golden(gpu, sub):
    regs = np.stack([sub["hll_" + n][28:] for n in ("a.fa", "b.fa", "d.fa")])
    assert np.array_equal(gpu.union(regs, 12), sub["union3"][28:])


@pytest.mark.parametrize("p,new_p", [(10, 8), (12, 7), (14, 13), (14, 10), (10, 10), (9, 4), (16, 12), (20, 14)])
def test_compress_vs_checker(gpu, checker, p, new_p):
    rng = np.random.default_rng(p * 100 + new_p)
    dense = synth.registers(p, 3, p, card=20.0 * (1 << p))
    rows = [dense[0], dense[1] * (rng.random(1 << p) < 0.3), dense[2] * (rng.random(1 << p) < 0.02), np.zeros(1 << p, np.uint8)]
    rows += list(synth.adversarial_registers(2, p)[[1, 3, 7]])
    regs = np.stack(rows).astype(np.uint8)
    got = gpu.compress(regs, p, new_p)
    for i, r in enumerate(regs):
        assert np.array_equal(got[i], checker.compress(r, p, new_p)), (p, new_p, i)


def test_compress_golden_and_errors(gpu, sub):
    assert np.array_equal(gpu.compress(sub["hll_a.fa"][28:], 12, 8)[0], sub["fold_a_8"][28:])
    assert np.array_equal(gpu.compress(sub["hll_b.fa"][28:], 12, 11)[0], sub["fold_b_default"][28:])
    assert np.array_equal(gpu.compress(sub["union3"][28:], 12, 10)[0], sub["fold_union3_10"][28:])
    with pytest.raises(gpu.Db200Error) as e:
        gpu.compress(sub["hll_a.fa"][28:], 12, 13)
    assert e.value.code == gpu.EINVAL and "Can't compress to a larger size" in str(e.value)


def test_cached_cardinality_override(gpu):
    """db200_dist_params.card / card_queries: the per-sketch terms of an all-pairs call come from the caller (the reference's
    cached hll_t::value_), the union term still from the registers; nothing lingers between calls."""
    p = 12
    regs = synth.registers(77, 40, p, card=1e5, group=8)
    card = gpu.cardinalities(regs, p)
    base = gpu.dist_symmetric(regs, p, result_type=gpu.JI)
    same = gpu.dist_symmetric(regs, p, result_type=gpu.JI, card=card)
    assert np.array_equal(same.view(np.uint32), base.view(np.uint32))
    scaled = card * np.linspace(0.9, 1.1, card.size)
    got = gpu.dist_symmetric(regs, p, result_type=gpu.JI, card=scaled).astype(np.float64)
    again = gpu.dist_symmetric(regs, p, result_type=gpu.JI)               # no override without the pointer
    assert np.array_equal(again.view(np.uint32), base.view(np.uint32))
    iu = np.triu_indices(card.size, 1)
    U = gpu.dist_symmetric(regs, p, result_type=gpu.UNION_SIZE).astype(np.float64)   # the union term itself (extension type)
    want = np.maximum(0.0, (scaled[iu[0]] + scaled[iu[1]] - U) / U)
    pos = want > 1e-3
    assert_close(got[pos], want[pos], rtol=2e-5, what="JI under overridden cardinalities")
    # rect: references and queries have their own pointers
    r = gpu.dist_rect(regs[:25], regs[25:], p, result_type=gpu.JI, card=scaled[:25], card_queries=scaled[25:]).astype(np.float64)
    full = np.zeros((card.size, card.size)); full[iu] = got; full = full + full.T
    assert_close(r, full[25:, :25], rtol=1e-6, what="rect under overridden cardinalities")
    with pytest.raises(ValueError):
        gpu.dist_symmetric(regs, p, card=scaled[:-1])


def test_union_size_extension(gpu, checker):
    """DB200_UNION_SIZE: hll_t::union_size (hll.h:1125-1141) for every pair — MLE of the register-wise maximum on the union
    path, the sum of the ertl_joint triple under the joint MLE."""
    p = 12
    regs = synth.registers(78, 30, p, card=2e5, group=6)
    got = gpu.dist_symmetric(regs, p, result_type=gpu.UNION_SIZE).astype(np.float64)
    iu = np.triu_indices(regs.shape[0], 1)
    want = checker.cardinalities(np.maximum(regs[iu[0]], regs[iu[1]]), p, 2)
    assert_close(got, want.astype(np.float32), what="union_size, union path")
    gotj = gpu.dist_symmetric(regs, p, jestim=3, result_type=gpu.UNION_SIZE)
    wantj = np.array([checker.triple(regs[i], regs[j], p, jestim=3).sum() for i, j in zip(*iu)])
    assert_close(gotj, wantj.astype(np.float32), what="union_size, joint MLE")


# ---- multi-device form of the host-pointer entry points ---------------------------------------------------------------
def test_all_devices_matches_single_device(gpu, monkeypatch):
    """device = DB200_ALL_DEVICES: genomes / sketches / block rows / queries sharded over every logical device.  On a
    single-GPU box DB200_VIRTUAL_DEVICES maps three logical devices onto the one GPU, so the partitioning, the per-device
    threads and the output slicing all run; results must be bit-identical to the one-device call."""
    if gpu.device_count() < 2:
        monkeypatch.setenv("DB200_VIRTUAL_DEVICES", "3")
    assert gpu.device_count() >= 2
    A = gpu.ALL_DEVICES
    gs = synth.genomes(5, 11, 70_000, group=4) + [np.frombuffer(b"ACGT" * 3, dtype=np.uint8)]
    genomes = [[g[:30_000], g[30_000:]] if i % 3 == 0 else g for i, g in enumerate(gs)]
    one = gpu.sketch_genomes(genomes, 21, 12, True, device=0)
    assert np.array_equal(gpu.sketch_genomes(genomes, 21, 12, True, device=A), one)
    p = 10
    regs = np.concatenate([synth.registers(3, 700, p, card=2e4, group=8), synth.adversarial_registers(3, p)])
    assert np.array_equal(gpu.cardinalities(regs, p, device=A), gpu.cardinalities(regs, p, device=0))
    for jestim, rt in ((2, gpu.JI), (3, gpu.MASH_DIST)):
        want = gpu.dist_symmetric(regs, p, k=21, jestim=jestim, result_type=rt, device=0)
        got = gpu.dist_symmetric(regs, p, k=21, jestim=jestim, result_type=rt, device=A)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (jestim, rt)
    n = regs.shape[0]
    tri = lambda r: r * (2 * n - r - 1) // 2
    want = gpu.dist_symmetric(regs, p, k=21, device=0)
    got = gpu.dist_symmetric(regs, p, k=21, device=A, row_begin=100, row_end=650)
    assert np.array_equal(got.view(np.uint32), want[tri(100):tri(650)].view(np.uint32))
    want = gpu.dist_rect(regs[:300], regs[300:], p, k=21, result_type=gpu.CONTAINMENT_INDEX, device=0)
    got = gpu.dist_rect(regs[:300], regs[300:], p, k=21, result_type=gpu.CONTAINMENT_INDEX, device=A)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    want = gpu.knn_rect(regs[:300], regs[300:], p, 7, k=21, result_type=gpu.MASH_DIST, device=0)
    got = gpu.knn_rect(regs[:300], regs[300:], p, 7, k=21, result_type=gpu.MASH_DIST, device=A)
    assert np.array_equal(got, want)
    assert np.array_equal(gpu.knn_symmetric(regs[:200], p, 5, k=21, device=A), gpu.knn_symmetric(regs[:200], p, 5, k=21, device=0))
    # streaming sketcher: slots spread over the devices
    sk = gpu.Sketcher(12, 21, True, device=A, nslots=4)
    for slot in range(4):
        g = genomes[slot]
        for rec in (g if isinstance(g, list) else [g]):
            sk.add_record(slot, np.asarray(rec).tobytes())
    for slot in range(4):
        assert np.array_equal(sk.finish(slot), one[slot])
    sk.close()


# ---- CLI ----------------------------------------------------------------------------------------------------------------
def test_cli_union(gpu, sub, tmp_path):
    names, hp = materialise(sub, str(tmp_path))
    run_cli(str(tmp_path), "union", "-o", "u3.hll", hp[0], hp[1], hp[2])
    assert_payload(gzip.open(tmp_path / "u3.hll").read(), sub["union3"].tobytes(), "union3")
    run_cli(str(tmp_path), "union", "-o", "u1.hll", hp[3])
    assert_payload(gzip.open(tmp_path / "u1.hll").read(), sub["union1"].tobytes(), "union1")
    (tmp_path / "plist.txt").write_text("# a comment line is skipped (get_lines)\n" + "\n".join(hp) + "\n")
    run_cli(str(tmp_path), "union", "-p", "3", "-o", "u4.hll", "-F", "plist.txt")
    assert_payload(gzip.open(tmp_path / "u4.hll").read(), sub["union4_F"].tobytes(), "union4")
    r = run_cli(str(tmp_path), "union", ok=False)
    assert r.returncode == 1 and b"require >= 1 paths" in r.stderr


def test_cli_hll(gpu, sub, tmp_path):
    names, _ = materialise(sub, str(tmp_path))
    r = run_cli(str(tmp_path), "hll", "-k", "21", "-S", "14", "-p", "2", *names)
    assert r.stdout == sub["hll_stdout"].tobytes()
    r = run_cli(str(tmp_path), "hll", "-k", "17", "-S", "12", "-p", "1", "-C", names[0], names[3])
    assert r.stdout == sub["hll_stdout_nocanon_p1"].tobytes()


def test_cli_fold_and_view(gpu, sub, tmp_path):
    names, hp = materialise(sub, str(tmp_path))
    run_cli(str(tmp_path), "fold", "-p", "8", "-o", "f8.hll", hp[0])
    assert_payload(gzip.open(tmp_path / "f8.hll").read(), sub["fold_a_8"].tobytes(), "fold -p 8")
    run_cli(str(tmp_path), "fold", "-o", "f11.hll", hp[1])
    assert_payload(gzip.open(tmp_path / "f11.hll").read(), sub["fold_b_default"].tobytes(), "fold default")
    run_cli(str(tmp_path), "fold", "-p", "12", "-o", "f12.hll", hp[0])
    assert_payload(gzip.open(tmp_path / "f12.hll").read(), sub["fold_a_same"].tobytes(), "fold to the same size")
    r = run_cli(str(tmp_path), "fold", "-p", "13", "-o", "f13.hll", hp[0], ok=False)
    assert r.returncode == 1 and b"Can't compress to a larger size" in r.stderr
    r = run_cli(str(tmp_path), "view", "f8.hll")
    assert r.stdout == sub["view_f8"].tobytes()


def test_cli_sketch_container(gpu, sub, tmp_path):
    names, _ = materialise(sub, str(tmp_path))
    run_cli(str(tmp_path), "sketch", "-k21", "-S12", "-p2", "--avoid-sorting", "-o", "cont.bin", *names)
    assert_container(gzip.open(tmp_path / "cont.bin").read(), sub["container"].tobytes(), "sketch -o")
    assert gzip.open(tmp_path / "cont.bin.labels.gz").read() == sub["container_labels"].tobytes()
    r = run_cli(str(tmp_path), "sketch", "-k21", "-S12", "--defer-hll", *names, ok=False)
    assert r.returncode == 1 and b"outside the B200 engine" in r.stderr


def test_cli_sketch_by_seq_and_dist_by_seq(gpu, sub, tmp_path):
    names, _ = materialise(sub, str(tmp_path))
    run_cli(str(tmp_path), "sketch_by_seq", "-k21", "-S10", "--defer-hll", "-o", "sbs.bin", "f.fa")
    assert_container(gzip.open(tmp_path / "sbs.bin").read(), sub["sbs"].tobytes(), "sketch_by_seq")
    assert (tmp_path / "sbs.bin.names").read_bytes() == sub["sbs_names"].tobytes()
    run_cli(str(tmp_path), "sbs", "-k15", "-S10", "-E", "--defer-hll", "-o", "m.bin", "multi.fa")
    assert_container(gzip.open(tmp_path / "m.bin").read(), sub["sbs_multi"].tobytes(), "sketch_by_seq -E")
    assert (tmp_path / "m.bin.names").read_bytes() == sub["sbs_multi_names"].tobytes()
    # without --defer-hll the reference writes b-bit minhash records: declined
    r = run_cli(str(tmp_path), "sketch_by_seq", "-k21", "-S10", "-o", "x.bin", "f.fa", ok=False)
    assert r.returncode == 1 and b"--defer-hll" in r.stderr
    # dist_by_seq on the reference-written container
    with gzip.open(tmp_path / "ref_m.bin", "wb") as f:
        f.write(sub["sbs_multi"].tobytes())
    (tmp_path / "ref_m.bin.names").write_bytes(sub["sbs_multi_names"].tobytes())
    flags = {"dbs_tsv_ji": [], "dbs_bin_mash": ["-b", "--mash-dist"], "dbs_full_jmle": ["-T", "-J"], "dbs_tsv_sizes_k": ["--sizes", "-k", "15"]}
    for rn in [str(x) for x in sub["dbs_runs"]]:
        kw = json.loads(str(sub[rn + "_kw"]))
        run_cli(str(tmp_path), "dist_by_seq", "-n", "ref_m.bin.names", "-o", "dbs.out", *flags[rn], "ref_m.bin")
        got, want = (tmp_path / "dbs.out").read_bytes(), sub[rn].tobytes()
        if kw.get("emit_fmt") == 1:
            assert got[:9] == want[:9] and len(got) == len(want), rn
            assert_close(np.frombuffer(got[9:], np.float32), np.frombuffer(want[9:], np.float32), what=rn)
        else:
            hostlib.assert_text_matches(got, want, what=rn)


def test_cli_card(gpu, sub, tmp_path):
    names, _ = materialise(sub, str(tmp_path))
    run_cli(str(tmp_path), "card", "-k21", "-S12", "--avoid-sorting", "-o", "card.txt", *names)
    hostlib.assert_text_matches((tmp_path / "card.txt").read_bytes(), sub["card_txt"].tobytes(), rtol=1e-6, what="card")
    run_cli(str(tmp_path), "card", "-k21", "-S12", "-I", "-e", "--avoid-sorting", "-o", "card_e.txt", *names)
    hostlib.assert_text_matches((tmp_path / "card_e.txt").read_bytes(), sub["card_sci_improved"].tobytes(), rtol=1e-6, what="card -e -I")
    run_cli(str(tmp_path), "card", "-k21", "-S12", "-b", "--avoid-sorting", "-o", "card.bin", *names)
    assert_close(np.frombuffer((tmp_path / "card.bin").read_bytes(), np.float32), np.frombuffer(sub["card_bin"].tobytes(), np.float32),
                 scale=1.0, what="card -b")


def test_cli_dist_defer_hll(gpu, sub, tmp_path):
    """--defer-hll: -E / -J have no effect (the final hll_t objects keep ERTL_MLE), cached sketches carry their value."""
    names, _ = materialise(sub, str(tmp_path))
    os.makedirs(tmp_path / "dk")
    run_cli(str(tmp_path), "dist", "-k21", "-S12", "-E", "-J", "-M", "--defer-hll", "-W", "-P", "dk", "--avoid-sorting",
            "-o", "ds.txt", "-O", "dd.txt", *names)
    assert (tmp_path / "ds.txt").read_bytes() == sub["defer_sizes"].tobytes()
    hostlib.assert_text_matches((tmp_path / "dd.txt").read_bytes(), sub["defer_dist"].tobytes(), what="defer dist")
    for n in names:
        assert_payload(gzip.open(tmp_path / "dk" / f"{n}.w.21.spacing.12.hll").read(), sub["defer_hll_" + n].tobytes(), "defer cache " + n)


def test_cli_live_subcommands_against_reference(gpu, ref, tmp_path):
    """Where oracle/_ref travelled: union + fold + card on fresh p=14 sketches against the reference's mains run live."""
    gs = synth.genomes(4242, 5, 150_000, group=5)
    from oracle.make_golden import write_fasta
    names = []
    for i, g in enumerate(gs):
        write_fasta(str(tmp_path / f"y{i}.fa"), [g.tobytes()], width=100)
        names.append(f"y{i}.fa")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        os.makedirs("sk")
        ref.cli_sketch(names, k=31, p=14, prefix="sk")
        hp = [ref.make_fname(n, 14, 31, 31, 31, "", "", "sk") for n in names]
        ref.union_main("-o", "ref_u.hll", *hp)
        ref.fold("ref_u.hll", "ref_f.hll", 9)
        ref.cli_card(names, "ref_card.txt", k=31, p=14)
    finally:
        os.chdir(cwd)
    run_cli(str(tmp_path), "union", "-o", "u.hll", *hp)
    assert_payload(gzip.open(tmp_path / "u.hll").read(), gzip.open(tmp_path / "ref_u.hll").read(), "live union")
    run_cli(str(tmp_path), "fold", "-p", "9", "-o", "f.hll", "u.hll")
    assert_payload(gzip.open(tmp_path / "f.hll").read(), gzip.open(tmp_path / "ref_f.hll").read(), "live fold")
    run_cli(str(tmp_path), "card", "-k31", "-S14", "--avoid-sorting", "-o", "card.txt", *names)
    hostlib.assert_text_matches((tmp_path / "card.txt").read_bytes(), (tmp_path / "ref_card.txt").read_bytes(), rtol=1e-6, what="live card")
