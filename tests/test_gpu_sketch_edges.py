"""Randomised edge-case sweep for the sketch kernel's bit tricks: record lengths and invalid bases placed around the
64-base block boundaries, every k from 1 to 32, many tiny records — always bit-exact against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def test_block_boundary_sweep(gpu, checker):
    rng = np.random.default_rng(64)
    for k in list(range(1, 33, 3)) + [31, 32]:
        genomes = []
        for g in range(6):
            recs = []
            for _ in range(int(rng.integers(1, 12))):
                L = int(rng.choice([0, 1, k - 1, k, k + 1, 63, 64, 65, 127, 128, 129, 191, 500, 4097]))
                r = ACGT[rng.integers(0, 4, size=max(L, 0))].copy()
                for _ in range(int(rng.integers(0, 4))):          # invalid bases near block edges
                    if L:
                        pos = int(min(L - 1, max(0, int(rng.choice([0, 62, 63, 64, 65, 126, 127, 128])) + int(rng.integers(-1, 2)))))
                        r[pos] = ord(rng.choice(list("NnRUx-")))
                recs.append(r.tobytes())
            genomes.append(recs)
        got = gpu.sketch_genomes(genomes, k, 10, True)
        for gi, recs in enumerate(genomes):
            np.testing.assert_array_equal(got[gi], checker.sketch(recs, k, 10, True), err_msg=f"k={k} genome {gi}")


def test_many_tiny_genomes(gpu, checker):
    rng = np.random.default_rng(7)
    genomes = [ACGT[rng.integers(0, 4, size=int(rng.integers(0, 300)))].tobytes() for _ in range(400)]
    got = gpu.sketch_genomes(genomes, 21, 8, False)
    for gi in range(0, 400, 7):
        np.testing.assert_array_equal(got[gi], checker.sketch([genomes[gi]], 21, 8, False))
