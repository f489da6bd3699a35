"""CPU: the FASTA edge cases the device parser is tested on (tests/test_gpu_fasta.py::edge_files — junk before the first header,
blank lines, '>' in mid-line, CR inside a line, CRLF, no trailing newline, header-only and empty files, records shorter than
k, 20 kB header lines ...) through the REAL reference — kseq_read + Encoder::for_each + hll_t::addh inside sketch_core<hll_t>
(oracle/_ref) — against the kseq-compatible host reader + the C restatement.  This pins the record rules both the host reader
and dashing_b200/csrc/fasta_logic.h implement to what the reference does with those bytes."""
import gzip
import importlib.util
import os

import numpy as np

import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _edge_files():
    spec = importlib.util.spec_from_file_location("_tgf", os.path.join(ROOT, "tests", "test_gpu_fasta.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.edge_files()


def test_reference_kseq_on_the_edge_case_files(ref, port, tmp_path):
    host = hostlib.load()
    k, p = 21, 10
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for name, raw in _edge_files().items():
            fn = name + ".fa"
            with open(fn, "wb") as f:
                f.write(raw)
            ref.cli_sketch([fn], k=k, p=p, nthreads=1)
            want = np.frombuffer(gzip.open(ref.make_fname(fn, p, k, k, k)).read()[28:], dtype=np.uint8)
            recs = hostlib.read_records(host, fn, cap=max(len(raw) + 64, 1 << 16))
            assert np.array_equal(port.sketch(recs, k, p, True), want), name
    finally:
        os.chdir(cwd)


def _random_text(rng, fastq_ok):
    """Adversarial line soup: headers, '@' / '+' lines, blank lines, CRLF, CR and '>' / '@' / '+' inside lines, no final newline."""
    parts = []
    for _ in range(int(rng.integers(0, 25))):
        kind, L = rng.random(), int(rng.integers(0, 40))
        body = bytes(rng.choice(list(b"ACGTACGTACGTacgtNn >@+-\r\t"), L).tolist()) if L else b""
        if kind < 0.25:
            line = b">" + body
        elif kind < 0.30 and fastq_ok:
            line = b"@" + body
        elif kind < 0.33 and fastq_ok:
            line = b"+" + body
        elif kind < 0.40:
            line = b""
        else:
            line = bytes(rng.choice(list(b"ACGT"), L).tolist()) if rng.random() < 0.7 else body
            if not fastq_ok and line[:1] in (b"@", b"+"):
                line = b"A" + line
        parts.append(line + (b"\r\n" if rng.random() < 0.2 else b"\n"))
    s = b"".join(parts)
    return s[:-1] if (rng.random() < 0.3 and s.endswith(b"\n")) else s


def test_host_reader_fuzz_against_reference_kseq(ref, port, tmp_path):
    """The kseq-compatible host reader on 400 random files against the reference's own kseq_read + sketch (k = 5 so that every
    misplaced record boundary or stray byte changes the registers).  This fuzz is what found the three rules a line-based
    reader gets wrong: the character scan for the next header when none is pending (kseq.h:183), truncated-quality records
    ending the file (:214), and empty lines stripping no CR (:195)."""
    host = hostlib.load()
    rng = np.random.default_rng(20261017)
    k, p = 5, 8
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for it in range(400):
            raw = _random_text(rng, fastq_ok=(it % 3 == 0))
            fn = f"f{it}.fa"
            with open(fn, "wb") as f:
                f.write(raw)
            ref.cli_sketch([fn], k=k, p=p, nthreads=1)
            want = np.frombuffer(gzip.open(ref.make_fname(fn, p, k, k, k)).read()[28:], dtype=np.uint8)
            recs = hostlib.read_records(host, fn, cap=1 << 16)
            assert np.array_equal(port.sketch(recs, k, p, True), want), (it, raw)
    finally:
        os.chdir(cwd)


def _emulated_records(exe, path, tmp):
    import subprocess
    out = os.path.join(tmp, "emul.out")
    subprocess.check_call([exe, path, out])
    data = open(out, "rb").read()
    nl = data.index(b"\n")
    flag, recs, at = int(data[:nl]), [], nl + 1
    while at < len(data):
        n = int.from_bytes(data[at:at + 8], "little")
        recs.append(data[at + 8:at + 8 + n])
        at += 8 + n
    return flag, recs


def test_device_parser_rules_against_reference_kseq(ref, port, tmp_path):
    """The per-warp logic the FASTA kernels run (fasta_logic.h), emulated lane by lane over whole files on the CPU
    (tests/fasta_device_emul.cpp), against the reference's kseq on the edge-case files and on random line soup: every file the
    parser does NOT flag must sketch to the reference's registers; files it flags go to the host reader (checked above)."""
    import subprocess
    exe = str(tmp_path / "fasta_device_emul")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "fasta_device_emul.cpp")])
    rng = np.random.default_rng(77)
    k, p = 5, 8
    cases = list(_edge_files().items()) + [(f"r{it}", _random_text(rng, fastq_ok=(it % 4 == 0))) for it in range(500)]
    # longer random files, so that lines, headers and CRs straddle the 16-byte lanes and 512-byte warps in every phase
    cases += [(f"long{it}", b"".join(_random_text(rng, fastq_ok=False) for _ in range(12))) for it in range(60)]
    cwd = os.getcwd()
    os.chdir(tmp_path)
    unflagged = 0
    try:
        for name, raw in cases:
            fn = name + ".fa"
            with open(fn, "wb") as f:
                f.write(raw)
            flag, recs = _emulated_records(exe, fn, str(tmp_path))
            if flag:
                continue
            unflagged += 1
            ref.cli_sketch([fn], k=k, p=p, nthreads=1)
            want = np.frombuffer(gzip.open(ref.make_fname(fn, p, k, k, k)).read()[28:], dtype=np.uint8)
            assert np.array_equal(port.sketch(recs, k, p, True), want), (name, raw)
    finally:
        os.chdir(cwd)
    assert unflagged > 300          # the comparison must not be vacuous
