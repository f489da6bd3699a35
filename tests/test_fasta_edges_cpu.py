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
