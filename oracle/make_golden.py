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
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def write_fasta(path, records, width=70, gz=False, crlf=False, fastq=False):
    op = gzip.open if gz else open
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


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "knn":       # only the nearest-neighbour fixtures (knn.npz + cli.npz)
        make_cli_golden(O.ref())
        make_knn_golden(O.ref())
    else:
        main()
