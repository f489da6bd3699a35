"""oracle/make_golden.py — regenerate tests/golden/*.npz from the REAL reference (oracle/_ref).

Run in the build container, where /root/reference exists:   python oracle/make_golden.py
The fixtures are small (inputs + outputs) so that the `-m gpu` tests and the CPU tests can run on
the GPU box, where neither /root/reference nor a rebuilt oracle/_ref is guaranteed.

Files written
  tests/golden/kat.npz        Wang KATs, k-mer streams of edge-case strings, phiX distinct 31-mer count,
                              the reference's bundled test genomes (bonsai/test/GCF_*.fna.gz) sketched at
                              p=10/14 with their sizes / JI / Mash values (SURVEY.md §8(c) smoke values)
  tests/golden/sketch.npz     small synthetic genomes (N runs, lower case, multi-record, short records)
                              and their register arrays for several (k, p, canon)
  tests/golden/dist.npz       a 34-sketch p=10 register matrix (correlated + adversarial rows) with
                              cardinalities for all estimators and the full pair matrix for every
                              estim x jestim x result_type x operand order; MLE on raw histograms
  tests/golden/hll_payload.npz  decompressed .hll bytes written by hll_t::write for known registers
"""
from __future__ import annotations

import glob
import gzip
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT  # the script dir (oracle/) would shadow the `oracle` package
from oracle import oracle as O  # noqa: E402
from dashing_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def read_fasta_records(path):
    op = gzip.open if path.endswith(".gz") else open
    recs, cur = [], []
    with op(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur:
                    recs.append(b"".join(cur))
                cur = []
            else:
                cur.append(line.strip())
    if cur:
        recs.append(b"".join(cur))
    return recs


def main():
    R = O.ref()
    os.makedirs(OUT, exist_ok=True)

    # ---------------- kat.npz ----------------
    wang_in = np.array([0, 1, 2, 0x123456789ABCDEF, 0xFFFFFFFFFFFFFFFF, 0x8000000000000000, 31, 1 << 40], dtype=np.uint64)
    wang_out = np.array([R.wang(int(v)) for v in wang_in], dtype=np.uint64)
    strings = [
        b"ACGTTGCATGCATGCCGATCGATTAGCTAGCTAGGCTAACGNACGTTGCATGCATGCCGATCGATTAGCTAGCTAGGCTAACG",
        b"ACGUACGUAC",
        b"acgtacgtacgtacgtacgtacgtacgtacgtacgtacgt",
        b"ACGT" * 3 + b"N" + b"TTTTGGGGCCCCAAAA" * 3,
        b"A" * 40,
        b"ACG",
        b"",
    ]
    kat = {"wang_in": wang_in, "wang_out": wang_out, "nstrings": np.array(len(strings))}
    for i, s in enumerate(strings):
        kat[f"str{i}"] = np.frombuffer(s, dtype=np.uint8)
        for k in (5, 31, 32):
            for canon in (0, 1):
                kat[f"kmers{i}_k{k}_c{canon}"] = R.kmers(s, k, bool(canon))
    phix = os.path.join(REF, "bonsai/test/phix.fa")
    recs = read_fasta_records(phix)
    allk = np.concatenate([R.kmers(r, 31, True) for r in recs])
    kat["phix_distinct_31mers"] = np.array(len(np.unique(allk)))  # bonsai/test/encoding.cpp:122 requires 5356
    assert int(kat["phix_distinct_31mers"]) == 5356, kat["phix_distinct_31mers"]
    gcf = sorted(glob.glob(os.path.join(REF, "bonsai/test/GCF_*.fna.gz")), key=lambda p: -os.path.getsize(p))
    kat["gcf_names"] = np.array([os.path.basename(p) for p in gcf])
    for p in (10, 14):
        regs = np.stack([R.sketch(read_fasta_records(path), 31, p, True) for path in gcf])
        kat[f"gcf_regs_p{p}"] = regs
        kat[f"gcf_sizes_p{p}"] = R.cardinalities(regs, p, 2)
        kat[f"gcf_ji_p{p}"] = R.dist_rows(regs, p, k=31, rtype=1)
        kat[f"gcf_mash_p{p}"] = R.dist_rows(regs, p, k=31, rtype=0)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **kat)
    print("gcf sizes p10:", kat["gcf_sizes_p10"].astype(np.uint64), "ji:", kat["gcf_ji_p10"])
    print("gcf mash p14:", kat["gcf_mash_p14"])

    # ---------------- sketch.npz ----------------
    rng = np.random.default_rng(20260925)
    base = synth.genomes(20260925, 4, 20000, group=2)
    g1 = synth.sprinkle(rng, base[1], n_runs=6)
    multi = [base[2][:700].tobytes(), base[2][700:720].tobytes(), b"", base[2][720:9000].tobytes(), b"ACGT", base[2][9000:].tobytes()]
    genomes = [[base[0].tobytes()], [g1.tobytes()], multi, [base[3][:30].tobytes()], [b""], [(b"ACGT" * 2000)]]
    sk = {"ngenomes": np.array(len(genomes))}
    for gi, g in enumerate(genomes):
        sk[f"g{gi}_nrec"] = np.array(len(g))
        for ri, r in enumerate(g):
            sk[f"g{gi}_r{ri}"] = np.frombuffer(r, dtype=np.uint8)
    combos = [(31, 10, 1), (31, 14, 1), (21, 16, 1), (32, 12, 1), (31, 10, 0), (4, 10, 1), (1, 10, 1), (16, 11, 0)]
    sk["combos"] = np.array(combos)
    for (k, p, canon) in combos:
        sk[f"regs_k{k}_p{p}_c{canon}"] = np.stack([R.sketch(g, k, p, bool(canon)) for g in genomes])
    np.savez_compressed(os.path.join(OUT, "sketch.npz"), **sk)

    # ---------------- dist.npz ----------------
    p = 10
    regs = np.concatenate([synth.registers(7, 24, p, card=3e5), synth.adversarial_registers(3, p)])
    d = {"p": np.array(p), "k": np.array(31), "regs": regs}
    for estim in (0, 1, 2):
        d[f"card_e{estim}"] = R.cardinalities(regs, p, estim)
        for jestim in (2, 3):
            for rtype in range(9):
                for order in (0, 1):
                    d[f"pairs_e{estim}_j{jestim}_r{rtype}_o{order}"] = R.dist_rows(regs, p, k=31, estim=estim, jestim=jestim,
                                                                                   rtype=rtype, order=order)
    d["rect_e2_j2_r1"] = R.dist_rect(regs[:20], regs[20:], p, k=31, rtype=1)
    d["rect_e2_j3_r0"] = R.dist_rect(regs[:20], regs[20:], p, k=31, jestim=3, rtype=0)
    hists = np.stack([R.histogram(r, p) for r in regs])
    d["hists"] = hists
    finite = [i for i in range(len(regs))]
    d["mle"] = np.array([R.mle(hists[i], p) for i in finite])
    for estim in (0, 1):
        d[f"est_e{estim}"] = np.array([R.estimate(hists[i], p, estim) for i in finite])
    # p = 14 spot check (8 sketches) so the committed fixtures also pin the headline sketch size
    r14 = synth.registers(11, 8, 14)
    d["regs14"] = r14
    d["card14"] = R.cardinalities(r14, 14, 2)
    d["pairs14_ji"] = R.dist_rows(r14, 14, k=31, rtype=1)
    d["pairs14_mash"] = R.dist_rows(r14, 14, k=31, rtype=0)
    d["pairs14_jmle"] = R.dist_rows(r14, 14, k=31, jestim=3, rtype=1)
    np.savez_compressed(os.path.join(OUT, "dist.npz"), **d)

    # ---------------- hll_payload.npz ----------------
    h = {}
    with tempfile.TemporaryDirectory() as td:
        for name, (rr, pp, est, jest, calc) in {"fresh_p10": (regs[3], 10, 2, 2, False), "calc_p10": (regs[5], 10, 2, 3, True),
                                                "fresh_p14": (r14[0], 14, 2, 2, False)}.items():
            path = os.path.join(td, name + ".hll")
            R.hll_write(path, rr, pp, est, jest, calc)
            with gzip.open(path, "rb") as f:
                h[name] = np.frombuffer(f.read(), dtype=np.uint8)
            h[name + "_regs"] = rr
    h["fname"] = np.array(R.make_fname("/data/genomes/g1.fna.gz", 14, 31, 31, 31, "", "", "/out"))
    h["fname_nopfx"] = np.array(R.make_fname("g1.fna.gz", 10, 0, 21, 21, "", "x", ""))
    np.savez_compressed(os.path.join(OUT, "hll_payload.npz"), **h)
    make_cli_golden(R)
    make_knn_golden(R)
    make_subcmd_golden(R)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def write_fasta(path, records, width=70, gz=False, crlf=False, fastq=False):
    # (mtime=0: the gzip header then holds no timestamp, so regenerating the fixtures is byte-reproducible)
    op = (lambda p_, m: gzip.GzipFile(p_, m, mtime=0)) if gz else open
    nl = b"\r\n" if crlf else b"\n"
    with op(path, "wb") as f:
        for i, r in enumerate(records):
            if fastq:
                f.write(b"@r%d some comment" % i + nl + r + nl + b"+" + nl + b"I" * len(r) + nl)
            else:
                f.write(b">rec%d description" % i + nl)
                for j in range(0, len(r), width):
                    f.write(r[j:j + width] + nl)


def cli_inputs():
    """Small FASTA/FASTQ inputs exercising the reader: multi-record, gz, CRLF, lower case, N runs, FASTQ, short records."""
    rng = np.random.default_rng(777)
    gs = synth.genomes(777, 6, 24000, group=3)
    g1 = synth.sprinkle(rng, gs[1], n_runs=5)
    files = {
        "a.fa": dict(records=[gs[0].tobytes()]),
        "b.fa": dict(records=[g1.tobytes()], width=60),
        "c.fa.gz": dict(records=[gs[2][:9000].tobytes(), gs[2][9000:9020].tobytes(), gs[2][9020:].tobytes()], gz=True),
        "d.fa": dict(records=[gs[3].tobytes()], crlf=True),
        "e.fq": dict(records=[gs[4][i:i + 150].tobytes() for i in range(0, 24000, 150)], fastq=True),
        "f.fa": dict(records=[gs[5][:18000].tobytes(), b"ACGTACGT", gs[5][18000:].tobytes()], width=80),
    }
    return files


def knn_inputs():
    """p=10: the dist.npz set (value ties through the adversarial rows) + 96 small sketches in groups of 8, where most
    pairs are unrelated (Jaccard exactly 0, Mash exactly 1: long runs of equal values at the cut)."""
    p = 10
    a = np.concatenate([synth.registers(7, 24, p, card=3e5), synth.adversarial_registers(3, p)])
    b = synth.registers(21, 96, p, card=2e4, group=8)
    return p, a, b


KNN_CASES = [  # (input set, rtype, jestim, nneighbors, nq)
    ("a", 0, 2, 5, 0), ("a", 1, 2, 5, 0), ("a", 1, 3, 7, 0), ("a", 2, 2, 3, 0), ("a", 7, 2, 4, 0), ("a", 8, 2, 64, 0),
    ("b", 0, 2, 10, 0), ("b", 1, 2, 10, 0), ("b", 3, 2, 1, 0), ("b", 5, 3, 12, 0), ("b", 1, 2, 96, 0),
    ("b", 0, 2, 6, 32), ("b", 1, 2, 6, 32), ("b", 6, 3, 9, 13), ("a", 1, 2, 40, 5),
]


def make_knn_golden(R):
    """perform_nns of the reference (one thread: the deterministic visiting order) on seeded register sets."""
    p, a, b = knn_inputs()
    out = {"p": np.array(p), "k": np.array(21), "cases": np.array(json.dumps(KNN_CASES))}
    for ci, (which, rtype, jestim, nn, nq) in enumerate(KNN_CASES):
        regs = a if which == "a" else b
        out[f"case{ci}"] = R.knn(regs, p, nn, k=21, jestim=jestim, rtype=rtype, nq=nq, nthreads=1)
    np.savez_compressed(os.path.join(OUT, "knn.npz"), **out)


def make_cli_golden(R):
    """Outputs of the reference's own drivers (sketch_core<hll_t>, dist_sketch_and_cmp<hll_t>) on the small inputs."""
    files = cli_inputs()
    names = list(files)
    out = {"names": np.array(names)}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            for n, kw in files.items():
                write_fasta(n, **kw)
                out["file_" + n] = np.frombuffer(open(n, "rb").read(), dtype=np.uint8)
            os.makedirs("sk", exist_ok=True)
            R.cli_sketch(names, k=31, p=10, prefix="sk")
            for n in names:
                hp = R.make_fname(n, 10, 31, 31, 31, "", "", "sk")
                out["hllname_" + n] = np.array(hp)
                out["hll_" + n] = np.frombuffer(gzip.open(hp, "rb").read(), dtype=np.uint8)
            runs = {
                "tsv_ji": dict(rtype=1, emit_fmt=0),
                "phylip_mash": dict(rtype=0, emit_fmt=2),
                "bin_mash": dict(rtype=0, emit_fmt=1),
                "full_jmle": dict(rtype=1, emit_fmt=3, jestim=3),
                "tsv_sizes_orig": dict(rtype=2, emit_fmt=0, estim=0, jestim=0),
                "tsv_symcont_k21_p12": dict(rtype=7, emit_fmt=0, k=21, p=12),
                "rect_tsv_cont": dict(rtype=5, emit_fmt=0, nq=2),
                "rect_bin_ji": dict(rtype=1, emit_fmt=1, nq=2),
                "rect_phylip_mash_jmle": dict(rtype=0, emit_fmt=2, nq=3, jestim=3),
                # --nearest-neighbors (nndist_loop); one thread = deterministic visiting and output order
                "nn_tsv_mash": dict(rtype=0, emit_fmt=0, nneighbors=3),
                "nn_bin_ji": dict(rtype=1, emit_fmt=1, nneighbors=2),
                "nn_tsv_ji_all": dict(rtype=1, emit_fmt=0, nneighbors=50),
            }
            out["runs"] = np.array(list(runs))
            for rn, kw in runs.items():
                R.cli_dist(names, "sizes.txt", "dist.out", **kw)
                out[f"{rn}_sizes"] = np.frombuffer(open("sizes.txt", "rb").read(), dtype=np.uint8)
                out[f"{rn}_dist"] = np.frombuffer(open("dist.out", "rb").read(), dtype=np.uint8)
                out[f"{rn}_kw"] = np.array(json.dumps(kw))
                if kw.get("emit_fmt") == 1 and not kw.get("nq"):
                    out[f"{rn}_labels"] = np.frombuffer(open("dist.out.labels", "rb").read(), dtype=np.uint8)
            # presketched: same distances from the .hll files written above
            hpaths = [str(out["hllname_" + n]) for n in names]
            R.cli_dist(hpaths, "sizes.txt", "dist.out", rtype=0, emit_fmt=0, presketched=True)
            out["presketched_tsv_mash_sizes"] = np.frombuffer(open("sizes.txt", "rb").read(), dtype=np.uint8)
            out["presketched_tsv_mash_dist"] = np.frombuffer(open("dist.out", "rb").read(), dtype=np.uint8)
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "cli.npz"), **out)


def _capture_stdout(fn):
    """Runs fn() with file descriptor 1 redirected to a temporary file (the reference prints with C stdio)."""
    sys.stdout.flush()
    with tempfile.TemporaryFile() as tf:
        saved = os.dup(1)
        os.dup2(tf.fileno(), 1)
        try:
            fn()
        finally:
            import ctypes
            ctypes.CDLL(None).fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
        tf.seek(0)
        return tf.read()


REAL_BINARY = "/tmp/oracle/dashing/dashing"   # `make dashing` of the unmodified reference, when someone built it


def _real(*args, cwd="."):
    """Runs the real reference binary, if present, for cross-checking the driver-made fixtures."""
    import subprocess
    if not os.path.exists(REAL_BINARY):
        return None
    return subprocess.run([REAL_BINARY, *args], cwd=cwd, capture_output=True, timeout=600)


def make_subcmd_golden(R):
    """SURVEY.md §8(f)3 subcommands — union, hll, fold, view, sketch -o, sketch_by_seq, dist_by_seq, card, dist --defer-hll —
    through the reference's own mains / templates (oracle/ref_driver.cpp), cross-checked against the real binary when built."""
    files = cli_inputs()
    names = ["a.fa", "b.fa", "d.fa", "f.fa"]
    out = {"names": np.array(names)}
    rd = lambda f: np.frombuffer(open(f, "rb").read(), dtype=np.uint8)
    gz = lambda f: np.frombuffer(gzip.open(f, "rb").read(), dtype=np.uint8)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            for n in names:
                write_fasta(n, **files[n])
                out["file_" + n] = rd(n)
            os.makedirs("sk", exist_ok=True)
            R.cli_sketch(names, k=21, p=12, prefix="sk")
            hp = [R.make_fname(n, 12, 21, 21, 21, "", "", "sk") for n in names]
            out["hllnames"] = np.array(hp)
            for n, h in zip(names, hp):
                out["hll_" + n] = gz(h)
            # union
            R.union_main("-o", "u3.hll", hp[0], hp[1], hp[2])
            out["union3"] = gz("u3.hll")
            R.union_main("-o", "u1.hll", hp[3])
            out["union1"] = gz("u1.hll")
            open("plist.txt", "w").write("\n".join(hp) + "\n")
            R.union_main("-p", "3", "-o", "u4.hll", "-F", "plist.txt")
            out["union4_F"] = gz("u4.hll")
            # hll
            out["hll_stdout"] = np.frombuffer(_capture_stdout(lambda: R.hll_main("-k", "21", "-S", "14", "-p", "2", *names)), dtype=np.uint8)
            out["hll_stdout_nocanon_p1"] = np.frombuffer(_capture_stdout(lambda: R.hll_main("-k", "17", "-S", "12", "-p", "1", "-C", names[0], names[3])), dtype=np.uint8)
            # fold / view
            R.fold(hp[0], "f8.hll", 8)
            out["fold_a_8"] = gz("f8.hll")
            R.fold(hp[1], "f11.hll", -1)
            out["fold_b_default"] = gz("f11.hll")
            R.fold(hp[0], "f12.hll", 12)
            out["fold_a_same"] = gz("f12.hll")
            R.fold("u3.hll", "fu.hll", 10)
            out["fold_union3_10"] = gz("fu.hll")
            R.view("f8.hll", "view.txt")
            out["view_f8"] = rd("view.txt")
            # sketch -o (one gzip stream of all sketches + labels)
            R.cli_sketch_container(names, "cont.bin", k=21, p=12)
            out["container"] = gz("cont.bin")
            out["container_labels"] = gz("cont.bin.labels.gz")
            # sketch_by_seq --defer-hll (hll_t records), dist_by_seq on its output
            R.cli_sketch_by_seq("f.fa", "sbs.bin", k=21, p=10)
            out["sbs"] = gz("sbs.bin")
            out["sbs_names"] = rd("sbs.bin.names")
            multi = [g[i:i + 4000].tobytes() for g in synth.genomes(99, 3, 16000, group=3) for i in range(0, 16000, 4000)]
            write_fasta("multi.fa", multi, width=90)
            out["file_multi.fa"] = rd("multi.fa")
            R.cli_sketch_by_seq("multi.fa", "m.bin", k=15, p=10, estim=0, jestim=0)
            out["sbs_multi"] = gz("m.bin")
            out["sbs_multi_names"] = rd("m.bin.names")
            dbs = {"dbs_tsv_ji": dict(), "dbs_bin_mash": dict(rtype=0, emit_fmt=1), "dbs_full_jmle": dict(emit_fmt=3, jestim=3),
                   "dbs_tsv_sizes_k": dict(rtype=2, k=15)}
            out["dbs_runs"] = np.array(list(dbs))
            for rn, kw in dbs.items():
                kk = kw.pop("k", 15)
                R.cli_dist_by_seq("m.bin.names", "m.bin", "dbs.out", k=kk, **kw)
                out[rn] = rd("dbs.out")
                out[rn + "_kw"] = np.array(json.dumps(dict(kw, k=kk)))
            # card
            R.cli_card(names, "card.txt", k=21, p=12)
            out["card_txt"] = rd("card.txt")
            R.cli_card(names, "card_e.txt", k=21, p=12, use_scientific=True, estim=1, jestim=1)
            out["card_sci_improved"] = rd("card_e.txt")
            R.cli_card(names, "card.bin", k=21, p=12, emit_binary=True)
            out["card_bin"] = rd("card.bin")
            # dist --defer-hll -W: -E and -J are ignored, cached sketches carry their value
            os.makedirs("dk", exist_ok=True)
            R.cli_dist_defer(names, "ds.txt", "dd.txt", k=21, p=12, estim=0, jestim=3, rtype=0, cache=True, prefix="dk")
            out["defer_sizes"] = rd("ds.txt")
            out["defer_dist"] = rd("dd.txt")
            for n in names:
                out["defer_hll_" + n] = gz(R.make_fname(n, 12, 21, 21, 21, "", "", "dk"))
            # ---- cross-check with the real binary where it was built
            if os.path.exists(REAL_BINARY):
                chk = lambda a, b, what: (_ for _ in ()).throw(AssertionError(what)) if bytes(a) != bytes(b) else None
                _real("union", "-o", "ru3.hll", hp[0], hp[1], hp[2]); chk(gz("ru3.hll"), out["union3"], "union3 vs real binary")
                r = _real("hll", "-k", "21", "-S", "14", "-p", "2", *names); chk(r.stdout, out["hll_stdout"], "hll vs real binary")
                _real("fold", "-p", "8", "-o", "rf8.hll", hp[0]); chk(gz("rf8.hll"), out["fold_a_8"], "fold vs real binary")
                r = _real("view", "f8.hll"); chk(r.stdout, out["view_f8"], "view vs real binary")
                _real("sketch", "-k21", "-S12", "-p8", "--avoid-sorting", "-o", "rcont.bin", *names); chk(gz("rcont.bin"), out["container"], "sketch -o vs real binary")
                _real("sketch_by_seq", "-k15", "-S10", "-E", "--defer-hll", "-o", "rm.bin", "multi.fa"); chk(gz("rm.bin"), out["sbs_multi"], "sketch_by_seq vs real binary")
                chk(rd("rm.bin.names"), out["sbs_multi_names"], "sketch_by_seq names vs real binary")
                _real("dist_by_seq", "-n", "m.bin.names", "-o", "rdbs.out", "m.bin"); chk(rd("rdbs.out"), out["dbs_tsv_ji"], "dist_by_seq vs real binary")
                _real("card", "-k21", "-S12", "--avoid-sorting", "-o", "rcard.txt", *names); chk(rd("rcard.txt"), out["card_txt"], "card vs real binary")
                _real("dist", "-k21", "-S12", "-E", "-J", "-M", "--defer-hll", "--avoid-sorting", "-o", "rds.txt", "-O", "rdd.txt", *names)
                chk(rd("rds.txt"), out["defer_sizes"], "defer sizes vs real binary"); chk(rd("rdd.txt"), out["defer_dist"], "defer dist vs real binary")
                print("subcmd fixtures agree with the real binary")
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "subcmd.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "subcmd":
        make_subcmd_golden(O.ref())
    elif len(sys.argv) > 1 and sys.argv[1] == "knn":       # only the nearest-neighbour fixtures (knn.npz + cli.npz)
        make_cli_golden(O.ref())
        make_knn_golden(O.ref())
    else:
        main()
