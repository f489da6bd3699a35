"""GPU parity, hot path (i): k-mer hashing + register update through the C ABI vs the oracle.
Registers are integers: the bar is bit-exact."""
import os

import numpy as np
import pytest

from dashing_b200 import synth

pytestmark = pytest.mark.gpu


def _golden_genomes(golden_dir):
    sk = np.load(os.path.join(golden_dir, "sketch.npz"))
    genomes = [[sk[f"g{g}_r{r}"].tobytes() for r in range(int(sk[f"g{g}_nrec"]))] for g in range(int(sk["ngenomes"]))]
    return sk, genomes


def test_sketch_golden_fixture(gpu, golden_dir):
    sk, genomes = _golden_genomes(golden_dir)
    for k, p, canon in sk["combos"]:
        got = gpu.sketch_genomes(genomes, int(k), int(p), bool(canon))
        np.testing.assert_array_equal(got, sk[f"regs_k{k}_p{p}_c{canon}"], err_msg=f"k={k} p={p} canon={canon}")


def test_reference_bundled_genome_sizes(gpu, golden_dir):
    """Registers of the reference's own test genomes -> sizes file values 4829255 / 2718859 / 2433839 / 2368528."""
    kat = np.load(os.path.join(golden_dir, "kat.npz"))
    card = gpu.cardinalities(kat["gcf_regs_p10"], 10)
    assert [int(c) for c in card] == [4829255, 2718859, 2433839, 2368528]


@pytest.mark.parametrize("k,p,canon", [(31, 14, True), (21, 16, True), (32, 10, True), (31, 12, False), (7, 11, True), (1, 10, False),
                                       (31, 18, True), (25, 7, True)])
def test_sketch_vs_oracle_seeded(gpu, checker, k, p, canon):
    rng = np.random.default_rng(1000 + k * 64 + p)
    gs = synth.genomes(int(rng.integers(1 << 30)), 5, 180_000, group=5)
    gs[1] = synth.sprinkle(rng, gs[1], n_runs=8)
    ragged = [gs[2][:5000].tobytes(), b"", gs[2][5000:5003].tobytes(), gs[2][5003:5003 + k].tobytes(), gs[2][6000:].tobytes(), b"N" * 70]
    genomes = [gs[0], gs[1], ragged, gs[3][: k - 1] if k > 1 else gs[3][:1], gs[4][:k]]
    got = gpu.sketch_genomes(genomes, k, p, canon)
    for gi, g in enumerate(genomes):
        recs = g if isinstance(g, list) else [np.asarray(g).tobytes()]
        np.testing.assert_array_equal(got[gi], checker.sketch(recs, k, p, canon), err_msg=f"genome {gi}")


def test_streaming_sketcher_matches_batch(gpu, checker):
    k, p = 31, 13
    gs = synth.genomes(5, 3, 70_000, group=3)
    sk = gpu.Sketcher(p, k, True, nslots=2)
    # slot 0: one genome in three records; slot 1: another genome, interleaved calls
    sk.add_record(0, gs[0][:100].tobytes())
    sk.add_record(1, gs[1].tobytes())
    sk.add_record(0, gs[0][100:40_000].tobytes())
    sk.add_record(0, gs[0][40_000:].tobytes())
    r0, r1 = sk.finish(0), sk.finish(1)
    np.testing.assert_array_equal(r0, checker.sketch([gs[0][:100].tobytes(), gs[0][100:40_000].tobytes(), gs[0][40_000:].tobytes()], k, p, True))
    np.testing.assert_array_equal(r1, checker.sketch([gs[1].tobytes()], k, p, True))
    # finish() clears the slot
    sk.add_record(0, gs[2].tobytes())
    np.testing.assert_array_equal(sk.finish(0), checker.sketch([gs[2].tobytes()], k, p, True))
    np.testing.assert_array_equal(sk.finish(1), np.zeros(1 << p, np.uint8))
    sk.close()


def test_empty_and_degenerate_inputs(gpu):
    z = gpu.sketch_genomes([[b""], [b"NNNNNNNN"], [b"ACGTACGTAC"]], 31, 10)
    assert not z.any()
    with pytest.raises(gpu.Db200Error) as ei:
        gpu.sketch_genomes([b"ACGT" * 20], 33, 10)
    assert ei.value.code == gpu.EUNSUPPORTED


def test_properties_at_bench_scale(gpu, checker):
    """Size-independent properties at BASELINE genome size (5 Mbp, k=31, p=14), where the oracle is only spot-checked:
    union = element-wise max, invariance to record order and to strand, idempotence."""
    k, p = 31, 14
    gs = synth.genomes(77, 4, 5_000_000, group=2)
    regs = gpu.sketch_genomes(gs, k, p)
    # (a) sketch of both genomes as two records == max of the two sketches; record order does not matter
    ab = gpu.sketch_genomes([[gs[0], gs[2]], [gs[2], gs[0]]], k, p)
    np.testing.assert_array_equal(ab[0], np.maximum(regs[0], regs[2]))
    np.testing.assert_array_equal(ab[1], ab[0])
    # (b) canonical k-mers: the reverse-complement strand gives the same sketch
    comp = np.zeros(256, np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    rc = comp[gs[1]][::-1].copy()
    np.testing.assert_array_equal(gpu.sketch_genomes([rc], k, p)[0], regs[1])
    # (c) idempotence: a genome repeated as two records changes nothing
    np.testing.assert_array_equal(gpu.sketch_genomes([[gs[3], gs[3]]], k, p)[0], regs[3])
    # (d) one full-size genome against the oracle
    np.testing.assert_array_equal(regs[1], checker.sketch([gs[1].tobytes()], k, p, True))
    # (e) cardinality of a 5 Mbp random genome is within HLL error (1.04/sqrt(m)) x 3 of its distinct k-mer count
    card = gpu.cardinalities(regs, p)
    assert abs(card[0] - (5_000_000 - k + 1)) < 3 * 1.04 / np.sqrt(1 << p) * 5_000_000


@pytest.mark.parametrize("pinned", [False, True])
def test_host_packed_upload_matches_ascii_upload(gpu, checker, monkeypatch, pinned):
    """db200_sketch_batch uploads a batch by two routes at once (host-packed 2-bit chunks from the front, ASCII chunks from
    the back when the source is page-locked): registers must not depend on the route or on where the chunk cuts fall."""
    k, p = 21, 12
    rng = np.random.default_rng(11)
    genomes = synth.genomes(5, 9, 150_001, group=3)
    genomes[2] = synth.sprinkle(rng, genomes[2])
    genomes[7] = [genomes[7][:70_000], genomes[7][70_000:70_013], genomes[7][70_013:]]     # multi-record, one record shorter than k
    bases, offs, grb = gpu.records_layout(genomes)
    if pinned:
        buf = gpu.pinned_empty(bases.size)
        buf[:] = bases
        bases = buf
    monkeypatch.setenv("DB200_HOST_PACK", "0")
    want = gpu.sketch_batch(bases, offs, grb, k, p, True)
    for i in (0, 2, 7):
        g = genomes[i] if isinstance(genomes[i], list) else [genomes[i]]
        assert np.array_equal(want[i], checker.sketch([np.asarray(r).tobytes() for r in g], k, p, True))
    for chunk in ("4096", "65536", "1000000"):
        monkeypatch.setenv("DB200_UPLOAD_CHUNK", chunk)
        monkeypatch.setenv("DB200_HOST_PACK", "1")
        assert np.array_equal(gpu.sketch_batch(bases, offs, grb, k, p, True), want), (chunk, "hybrid")
        monkeypatch.setenv("DB200_HOST_PACK", "0")
        assert np.array_equal(gpu.sketch_batch(bases, offs, grb, k, p, True), want), (chunk, "ascii")
