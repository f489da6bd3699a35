"""CPU tests of the host layer (no GPU compute): file naming, the .hll container, the FASTA/FASTQ reader and every
output format, against fixtures written by the reference's own drivers (tests/golden/cli.npz, hll_payload.npz)."""
import gzip
import os
import re

import numpy as np
import pytest

import hostlib

pytestmark = pytest.mark.skipif(not os.path.exists(hostlib.HOST_SO), reason="host layer not built (run __graft_entry__.build())")


@pytest.fixture(scope="module")
def host():
    return hostlib.load()


@pytest.fixture(scope="module")
def cli(golden_dir):
    return np.load(os.path.join(golden_dir, "cli.npz"))


def test_make_fname(host, golden_dir, cli):
    h = np.load(os.path.join(golden_dir, "hll_payload.npz"))
    assert hostlib.make_fname(host, "/data/genomes/g1.fna.gz", 14, 31, prefix="/out") == str(h["fname"]) == "/out/g1.fna.gz.w.31.spacing.14.hll"
    assert hostlib.make_fname(host, "g1.fna.gz", 10, 21, suffix="x") == str(h["fname_nopfx"]) == "g1.fna.gz.w.21.spacing.sufx.10.hll"
    for n in cli["names"]:
        assert hostlib.make_fname(host, str(n), 10, 31, prefix="sk") == str(cli["hllname_" + str(n)])


def test_hll_container_roundtrip(host, golden_dir, tmp_path):
    h = np.load(os.path.join(golden_dir, "hll_payload.npz"))
    for name, p, est, jest in (("fresh_p10", 10, 2, 2), ("fresh_p14", 14, 2, 2), ("calc_p10", 10, 2, 3)):
        want = h[name].tobytes()
        regs = np.ascontiguousarray(h[name + "_regs"])
        value = float(np.frombuffer(want[20:28], dtype=np.float64)[0])
        path = str(tmp_path / (name + ".hll"))
        assert host.db200h_write_hll(path.encode(), regs.ctypes.data, p, est, jest, value) == 0
        assert gzip.open(path, "rb").read() == want            # identical decompressed payload
        # and a file written by the reference (here: its payload re-gzipped) reads back
        ref_path = str(tmp_path / (name + ".ref.hll"))
        with gzip.open(ref_path, "wb") as f:
            f.write(want)
        back = np.zeros(1 << p, np.uint8); hdr = np.zeros(5, np.uint32)
        import ctypes as C
        v = C.c_double()
        assert host.db200h_read_hll(ref_path.encode(), back.ctypes.data, back.size, hdr.ctypes.data, C.byref(v)) == 0
        np.testing.assert_array_equal(back, regs)
        assert list(hdr) == [int(value >= 0), est, jest, 1, p] and (v.value == value)


def test_fasta_reader_matches_kseq(host, cli, port, tmp_path):
    """Records parsed by the host reader, sketched by the oracle, must give the registers the reference's kseq-fed
    sketch_core wrote (multi-line, multi-record, gz, CRLF, FASTQ, lower case, N runs)."""
    names = hostlib.materialise_inputs(cli, str(tmp_path))
    for n in names:
        recs = hostlib.read_records(host, str(tmp_path / n))
        assert all(b"\n" not in r and b"\r" not in r and b">" not in r for r in recs)
        want = cli["hll_" + n].tobytes()
        np.testing.assert_array_equal(port.sketch(recs, 31, 10, True), np.frombuffer(want[28:], dtype=np.uint8), err_msg=n)
    assert len(hostlib.read_records(host, str(tmp_path / "c.fa.gz"))) == 3
    assert len(hostlib.read_records(host, str(tmp_path / "e.fq"))) == 160


def _parse_ut(text: bytes, skip_header: int):
    rows = text.decode().strip("\n").split("\n")[skip_header:]
    vals = []
    for r in rows:
        vals += [float(x) for x in re.split(r"\t", r)[1:] if x not in ("-", "")]
    return np.array(vals, dtype=np.float32)


def test_symmetric_formats_byte_exact(host, cli):
    names = [str(x) for x in cli["names"]]
    n = len(names)
    raw = cli["bin_mash_dist"].tobytes()
    assert raw[0] == 0 and int(np.frombuffer(raw[1:9], dtype=np.uint64)[0]) == n          # distmat header
    packed = np.frombuffer(raw[9:], dtype=np.float32)
    assert packed.size == n * (n - 1) // 2
    # PHYLIP from the binary values must reproduce the reference's text byte for byte
    assert hostlib.format_symmetric(host, names, packed, 2) == cli["phylip_mash_dist"].tobytes()
    # upper-triangular TSV: the presketched run printed the same Mash values under the .hll names
    hnames = [str(cli["hllname_" + x]) for x in names]
    assert hostlib.format_symmetric(host, hnames, packed, 0) == cli["presketched_tsv_mash_dist"].tobytes()
    # the other TSV outputs: re-format the 6-digit values parsed from the reference's own text
    for run, fmt, skip in (("tsv_ji", 0, 1), ("tsv_sizes_orig", 0, 1), ("tsv_symcont_k21_p12", 0, 1)):
        want = cli[run + "_dist"].tobytes()
        assert hostlib.format_symmetric(host, names, _parse_ut(want, skip), fmt) == want, run
    # FULL_TSV (including the reference's "#Names<first name>" header quirk)
    want = cli["full_jmle_dist"].tobytes()
    rows = want.decode().strip("\n").split("\n")[1:]
    full = np.array([[float(x) for x in r.split("\t")[1:]] for r in rows], dtype=np.float32)
    iu = np.triu_indices(n, 1)
    assert hostlib.format_symmetric(host, names, full[iu], 3, lower=full.T[iu]) == want
    assert want.startswith(b"#Namesa.fa\tb.fa")
    assert cli["bin_mash_labels"].tobytes() == ("\n".join(names) + "\n").encode()


def test_sizes_and_rect_formats(host, cli):
    names = [str(x) for x in cli["names"]]
    want = cli["tsv_ji_sizes"].tobytes()
    card = [float(l.split(b"\t")[1]) + 0.75 for l in want.strip().split(b"\n")[1:]]      # size_t(card) truncates
    assert hostlib.format_sizes(host, names, card) == want
    rect = cli["rect_tsv_cont_dist"].tobytes().strip(b"\n").split(b"\n")
    for line in rect:
        f = line.split(b"\t")
        assert hostlib.format_rect_row(host, f[0].decode(), [float(x) for x in f[1:]]) == line + b"\n"
    raw = np.frombuffer(cli["rect_bin_ji_dist"].tobytes(), dtype=np.float32)
    assert raw.size == 2 * 4                                                              # nq x nr floats, no header


def test_neighbor_formats_byte_exact(host, cli):
    """nndist_loop's outputs (src/sketch_and_cmp.h:733-782) re-formatted from the values the reference printed / wrote."""
    names = [str(x) for x in cli["names"]]
    n = len(names)
    dt = np.dtype([("value", np.float32), ("index", np.uint32)])
    # binary table: uint32 n, uint32 nn, pairs
    raw = cli["nn_bin_ji_dist"].tobytes()
    hdr = np.frombuffer(raw[:8], dtype=np.uint32)
    assert hdr[0] == n and hdr[1] == 2
    nb = np.frombuffer(raw[8:], dtype=dt).reshape(n, 2)
    assert hostlib.format_neighbors(host, names, 0, nb, 1) == raw
    # TSV tables, including the (uint32(-1), -FLT_MAX) filler printed as "-1:-3.40282e+38"
    for run in ("nn_tsv_mash", "nn_tsv_ji_all"):
        want = cli[run + "_dist"].tobytes()
        lines = want.decode().strip("\n").split("\n")
        assert lines[0] == "#File\tNeighbor ID:distance\t..."
        rows = []
        for ln in lines[1:]:
            f = ln.split("\t")
            rows.append([(np.float32(x.split(":")[1]), np.uint32(int(x.split(":")[0]) & 0xFFFFFFFF)) for x in f[1:]])
        nb = np.array(rows, dtype=dt)
        assert hostlib.format_neighbors(host, names, 0, nb, 0) == want, run
    assert b"\t-1:-3.40282e+38\n" in cli["nn_tsv_ji_all_dist"].tobytes()


def test_file_windows(host, cli, tmp_path):
    """The batch driver parses every file into a window sized from the file (plain: its size; gzip: the ISIZE trailer).
    The window must hold the file's sequence, and a multi-member gzip file (ISIZE covers the last member only) must be
    reported as too small so the driver re-reads it through the growing-string path."""
    import ctypes as C
    import gzip
    names = hostlib.materialise_inputs(cli, str(tmp_path))
    for n in names:
        path = str(tmp_path / n)
        cap = host.db200h_file_capacity(path.encode())
        recs = hostlib.read_records(host, path)
        assert sum(len(r) for r in recs) <= cap, n
        raw = open(path, "rb").read()
        if raw[:2] == b"\x1f\x8b":
            assert cap == len(gzip.decompress(raw))
        else:
            assert cap == len(raw)
        # a window of the sequence length (+1: a CR is dropped only after it has been appended) works, a smaller one does not
        need = sum(len(r) for r in recs)
        bases = np.zeros(need + 1, dtype=np.uint8); offs = np.zeros(1000, dtype=np.uint64)
        assert host.db200h_read_records(path.encode(), bases.ctypes.data, need + 1, offs.ctypes.data, 999) == len(recs)
        if need:
            assert host.db200h_read_records(path.encode(), bases.ctypes.data, need - 1, offs.ctypes.data, 999) == -2
    # two gzip members back to back: zlib reads both, ISIZE describes the second only
    a = b">r1\n" + b"ACGT" * 500 + b"\n"
    b = b">r2\nGGCC\n"
    mm = tmp_path / "multi.fa.gz"
    mm.write_bytes(gzip.compress(a) + gzip.compress(b))
    assert host.db200h_file_capacity(str(mm).encode()) == len(b)
    recs = hostlib.read_records(host, str(mm))
    assert [len(r) for r in recs] == [2000, 4]
    small = np.zeros(len(b), dtype=np.uint8); offs = np.zeros(10, dtype=np.uint64)
    assert host.db200h_read_records(str(mm).encode(), small.ctypes.data, len(b), offs.ctypes.data, 9) == -2


def test_record_names_and_path_lists(host, golden_dir, tmp_path):
    """sketch_by_seq's .names content (kseq's name = header up to the first whitespace) and get_lines' rules (bonsai util.h:1176-1184:
    empty lines and '#' lines are skipped, a missing file gives no paths), against the reference-written names file."""
    sub = np.load(os.path.join(golden_dir, "subcmd.npz"))
    for fn, key in (("multi.fa", "sbs_multi_names"), ("f.fa", "sbs_names")):
        (tmp_path / fn).write_bytes(sub["file_" + fn].tobytes())
        want = sub[key].tobytes()
        assert want.startswith(b"#k=")
        assert hostlib.record_names(host, str(tmp_path / fn)) == want.split(b"\n", 1)[1]
        # and dist_by_seq reads that names file as its labels: the '#k=..' header line is not a label
        (tmp_path / (fn + ".names")).write_bytes(want)
        assert hostlib.get_paths(host, str(tmp_path / (fn + ".names"))) == want.decode().split("\n")[1:-1]
    (tmp_path / "fq.fq").write_bytes(b"@r1 comment here\nACGT\n+\nIIII\n@r2\tx\nGGCC\n+r2\nIIII\n")
    assert hostlib.record_names(host, str(tmp_path / "fq.fq")) == b"r1\nr2\n"
    (tmp_path / "list.txt").write_text("a.fa\n\n#skipped\nb c.fa\n")
    assert hostlib.get_paths(host, str(tmp_path / "list.txt")) == ["a.fa", "b c.fa"]
    assert hostlib.get_paths(host, str(tmp_path / "missing.txt")) == []


def test_text_numbers_are_printf_g6(host):
    """The emitters format with std::to_chars(general, 6); the reference with "%.6g" / "%g" / "%0.6g".  Same characters, including
    the exponent switch points, zeros, infinities and NaNs."""
    rng = np.random.default_rng(5)
    vals = np.concatenate([
        rng.random(4000).astype(np.float32), (rng.random(2000) * 10.0 ** rng.integers(-12, 12, 2000)).astype(np.float32),
        np.array([0.0, -0.0, 1.0, 0.5, 1e-5, 9.99999e-5, 1e-4, 999999.0, 999999.5, 1e6, 123456.7, np.inf, -np.inf, np.nan, 3.4028235e38, 1.4e-45],
                 dtype=np.float32)])
    got = hostlib.format_rect_row(host, "q", vals).decode().rstrip("\n").split("\t")[1:]
    want = ["%g" % float(v) for v in vals]
    assert got == want


def test_raw_file_windows(host, tmp_path):
    """What the CLI does with a sequence file: size a window (file size, or the gzip ISIZE trailer) and read the raw bytes into
    it — inflating gzip, overflowing on multi-member gzip (whose trailer covers the last member only), nothing parsed."""
    raw = b">h1 x\r\nACGT\r\n\n>h2\nGGCC" * 1000
    (tmp_path / "p.fa").write_bytes(raw)
    (tmp_path / "g.fa.gz").write_bytes(gzip.compress(raw))
    (tmp_path / "mm.fa.gz").write_bytes(gzip.compress(raw[:9000]) + gzip.compress(raw[9000:]))
    (tmp_path / "e.fa").write_bytes(b"")
    for name, want in (("p.fa", raw), ("g.fa.gz", raw), ("e.fa", b"")):
        cap = host.db200h_file_capacity(str(tmp_path / name).encode())
        assert cap == len(want), name
        buf = np.zeros(cap + 16, dtype=np.uint8)
        n = host.db200h_slurp(str(tmp_path / name).encode(), buf.ctypes.data, cap)
        assert n == len(want) and buf[:n].tobytes() == want, name
    cap = host.db200h_file_capacity(str(tmp_path / "mm.fa.gz").encode())
    assert cap == len(raw) - 9000                      # ISIZE of the LAST member
    buf = np.zeros(len(raw) + 16, dtype=np.uint8)
    assert host.db200h_slurp(str(tmp_path / "mm.fa.gz").encode(), buf.ctypes.data, cap) == -2     # -> the re-read path
    assert host.db200h_slurp(str(tmp_path / "mm.fa.gz").encode(), buf.ctypes.data, len(raw)) == len(raw) and buf[:len(raw)].tobytes() == raw
    assert host.db200h_slurp(str(tmp_path / "missing.fa").encode(), buf.ctypes.data, 10) == -1
