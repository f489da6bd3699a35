"""GPU: device-side FASTA parsing: raw file text in, registers out — against the
records the kseq-compatible host reader yields for the same bytes (tests/test_host_formats.py pins that reader to the
reference's kseq-fed sketches), sketched by the checker.  Registers bit-exact."""
import os

import numpy as np
import pytest

import hostlib
from dashing_b200 import synth

pytestmark = pytest.mark.gpu


def fasta(records, width=70, nl=b"\n", header=lambda i: b">r%d some description" % i, last_nl=True):
    out = []
    for i, r in enumerate(records):
        out.append(header(i) + nl)
        for j in range(0, len(r), width):
            out.append(r[j:j + width] + nl)
    s = b"".join(out)
    return s if last_nl else s[:-len(nl)]


def expected(host, checker, files, k, p, canon, tmp_path):
    recs = []
    for i, raw in enumerate(files):
        path = tmp_path / f"x{i}.fa"
        path.write_bytes(raw)
        recs += hostlib.read_records(host, str(path), cap=max(len(raw) + 64, 1 << 16))
    return checker.sketch(recs, k, p, canon)


@pytest.fixture(scope="module")
def host():
    return hostlib.load()


def edge_files():
    rng = np.random.default_rng(2024)
    g = [x.tobytes() for x in synth.genomes(11, 8, 30_000, group=4)]
    spr = synth.sprinkle(rng, np.frombuffer(g[1], np.uint8).copy(), n_runs=6).tobytes().replace(b"\n", b"N")
    files = {
        "plain70": fasta([g[0]]),
        "plain60_multi": fasta([g[2][:9000], g[2][9000:9010], g[2][9010:]], width=60),
        "crlf": fasta([g[3]], nl=b"\r\n"),
        "no_trailing_newline": fasta([g[4]], last_nl=False),
        "trailing_cr_only": fasta([g[4][:5000]], nl=b"\r\n", last_nl=False) + b"\r",
        "junk_before_first_header": b"this is not sequence\nACGTACGTACGTACGTACGTACGTACGTACGTAAAC\n\n" + fasta([g[5]]),
        "blank_lines_spaces_gt": b">h1\n" + g[6][:3000] + b"\n\n\n" + g[6][3000:6000] + b" \n" + g[6][6000:7000] + b">" + g[6][7000:9000] + b"\n \n" + g[6][9000:] + b"\n",
        "lower_n_iupac": fasta([spr], width=80),
        "empty": b"",
        "header_only": b">lonely header\n",
        "header_then_nothing_then_record": b">a\n>b\n\n>c\n" + g[7][:2000] + b"\n>d",
        "short_records": fasta([g[7][i:i + 25] for i in range(0, 3000, 25)], width=100),
        "one_base_per_line": b">x\n" + b"\n".join(bytes([c]) for c in g[7][:9000]) + b"\n",
        "one_long_line": b">x\n" + g[0] + b"\n>y\n" + g[3],
        "long_header": b">" + b"h" * 20000 + b" ACGTACGTACGTACGTACGTACGTACGTACGTACGT\n" + g[5][:4000] + b"\n",
        "cr_inside_line": b">x\n" + g[6][:100] + b"\r" + g[6][100:200] + b"\r\r\n" + g[6][200:4000] + b"\n",
    }
    return files


@pytest.mark.parametrize("k,p,canon", [(31, 12, True), (21, 10, False), (32, 14, True), (5, 8, True)])
def test_edge_case_files(gpu, checker, host, tmp_path, k, p, canon):
    files = edge_files()
    names = list(files)
    got, status = gpu.sketch_fasta([[files[n]] for n in names], k, p, canon)
    assert not status.any(), [n for n, s in zip(names, status) if s]
    for i, n in enumerate(names):
        want = expected(host, checker, [files[n]], k, p, canon, tmp_path)
        assert np.array_equal(got[i], want), f"{n}: registers differ (k={k}, p={p}, canon={canon})"


def test_multi_file_genomes_and_chunk_boundaries(gpu, checker, host, tmp_path, monkeypatch):
    """Genomes made of several files (FNAME_SEP paths), and upload chunks of a few blocks only so that chunk boundaries — where
    the parser's state, the output position and the byte after a CR cross from one launch to the next — fall everywhere."""
    files = edge_files()
    g = [x.tobytes() for x in synth.genomes(5, 6, 200_000, group=3)]
    genomes = [
        [files["plain70"], files["crlf"], files["empty"], files["junk_before_first_header"]],
        [fasta([g[0]], width=61), fasta([g[1]], width=80, nl=b"\r\n")],
        [files["header_only"]],
        [fasta([g[2][:100_000], g[2][100_000:]], width=8191), fasta([g[3]], width=8192), fasta([g[4]], width=16383, last_nl=False)],
        [files["empty"]],
        [fasta([g[5]], width=79)],
    ]
    want = [expected(host, checker, fs, 21, 12, True, tmp_path) for fs in genomes]
    for chunk in (None, "8192", "24576", "65536"):
        if chunk:
            monkeypatch.setenv("DB200_FASTA_CHUNK", chunk)
        got, status = gpu.sketch_fasta(genomes, 21, 12, True)
        assert not status.any()
        for i in range(len(genomes)):
            assert np.array_equal(got[i], want[i]), f"genome {i}, chunk {chunk}"
    monkeypatch.delenv("DB200_FASTA_CHUNK")
    # same through every (logical) device
    if gpu.device_count() < 2:
        monkeypatch.setenv("DB200_VIRTUAL_DEVICES", "3")
    got, status = gpu.sketch_fasta(genomes, 21, 12, True, device=gpu.ALL_DEVICES)
    for i in range(len(genomes)):
        assert np.array_equal(got[i], want[i]), f"genome {i}, all devices"


def test_fastq_syntax_is_flagged(gpu):
    g = synth.genomes(3, 2, 5000, group=2)
    fq = b"".join(b"@r%d\n" % i + g[0][i:i + 100].tobytes() + b"\n+\n" + b"I" * 100 + b"\n" for i in range(0, 5000, 100))
    plus_inside = b">x\n" + g[1][:1000].tobytes() + b"\n+\n" + g[1][1000:2000].tobytes() + b"\n"
    at_inside = b">x\n" + g[1][:1000].tobytes() + b"\n@y\n" + g[1][1000:2000].tobytes() + b"\n"
    ok = fasta([g[1].tobytes()])
    plus_in_header_or_midline = b">x + @\n" + g[1][:1000].tobytes() + b"+@\n"
    # before a file's first header kseq scans CHARACTERS for '>' / '@' (kseq.h:183): junk holding either would open a record
    gt_in_junk = b"some junk > here\n" + ok
    at_in_junk = b"mail me: a@b.c\n\n" + ok
    plain_junk = b"no special bytes in this junk + nor here\n+ or here\n" + ok
    _, status = gpu.sketch_fasta([[fq], [plus_inside], [ok], [at_inside], [plus_in_header_or_midline], [gt_in_junk], [at_in_junk], [plain_junk]],
                                 21, 10, True)
    assert list(status) == [1, 1, 0, 1, 0, 1, 1, 0]


def test_large_batch_default_chunks(gpu, checker, host, tmp_path):
    """~90 MB of text: crosses the default 64 MiB chunk, four genome groups."""
    gs = synth.genomes(77, 9, 10_000_000, group=3)
    files = [[fasta([x.tobytes()], width=80)] for x in gs]
    got, status = gpu.sketch_fasta(files, 31, 14, True)
    assert not status.any()
    want = gpu.sketch_genomes(gs, 31, 14, True)       # the record path (pinned against the checker elsewhere)
    assert np.array_equal(got, want)
    assert np.array_equal(got[4], checker.sketch([gs[4].tobytes()], 31, 14, True))


def test_argument_checks(gpu):
    text, offs, lens, gfb = gpu.fasta_layout([[b">a\nACGT\n"], [b">b\nACGT\n"]])
    import ctypes as C
    out = np.zeros((2, 1 << 10), np.uint8)
    bad = offs.copy(); bad[1] += 16
    rc = gpu.lib.db200_sketch_fasta_batch(0, 10, 21, 1, text.ctypes.data_as(C.c_void_p), bad.ctypes.data_as(gpu.u64p), lens.ctypes.data_as(gpu.u64p),
                                          2, gfb.ctypes.data_as(gpu.u64p), 2, out.ctypes.data_as(gpu.u8p), None)
    assert rc == gpu.EINVAL
    rc = gpu.lib.db200_sketch_fasta_batch(0, 10, 33, 1, text.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(gpu.u64p), lens.ctypes.data_as(gpu.u64p),
                                          2, gfb.ctypes.data_as(gpu.u64p), 2, out.ctypes.data_as(gpu.u8p), None)
    assert rc == gpu.EUNSUPPORTED


def test_device_parser_against_the_reference_drivers_directly(gpu, ref, tmp_path):
    """no link through this repo's host reader — the reference's own sketch_core<hll_t> (kseq_read +
    Encoder::for_each + hll_t::addh + hll_t::write, run from oracle/_ref) writes the .hll file of every edge-case input, and the
    device parser's registers for the same raw bytes must be that file's payload."""
    import gzip
    k, p = 21, 10
    files = edge_files()
    names = list(files)
    got, status = gpu.sketch_fasta([[files[n]] for n in names], k, p, True)
    assert not status.any()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for i, n in enumerate(names):
            fn = n + ".fa"
            with open(fn, "wb") as f:
                f.write(files[n])
            ref.cli_sketch([fn], k=k, p=p, nthreads=1)
            want = np.frombuffer(gzip.open(ref.make_fname(fn, p, k, k, k)).read()[28:], dtype=np.uint8)
            assert np.array_equal(got[i], want), n
    finally:
        os.chdir(cwd)
