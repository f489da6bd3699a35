"""The Python all-pairs front-end (dashing_b200/sketch_util.py): index helpers on the CPU tier, the matrices
against the reference's own hll_t methods on the GPU tier."""
import numpy as np
import pytest

from dashing_b200 import synth
from parity import assert_close


def test_index_helpers_cpu():
    from dashing_b200 import sketch_util as su
    n = 7
    flat = np.arange(n * (n - 1) // 2, dtype=np.float32)
    assert su.flat2fullsz(flat.size) == n and su.flat2fullsz(0) in (0, 1)
    full = su.tri2full(flat)
    for i in range(n):
        for j in range(n):
            if i == j:
                assert full[i, j] == 1.0
            else:
                assert full[i, j] == flat[su.ij2ind(i, j, n)] and su.ij2ind(i, j, n) == su.ij2ind(j, i, n)
    # the reference's formula (util.cpp:148) verbatim
    assert su.ij2ind(2, 5, n) == (2 * (n * 2 - 2 - 1)) // 2 + 5 - 3
    with pytest.raises(ValueError):
        su.flat2fullsz(5)
    with pytest.raises(ValueError):
        su.jaccard_matrix(np.zeros((3, 100), np.uint8))


@pytest.mark.gpu
def test_matrices_against_the_reference_methods(gpu, checker):
    from dashing_b200 import sketch_util as su
    p, n = 12, 14
    regs = synth.registers(4242, n, p, card=3e5, group=7)
    iu = np.triu_indices(n, 1)
    trip = np.array([checker.triple(regs[i], regs[j], p) for i, j in zip(*iu)])          # {x only, y only, intersection}
    card = checker.cardinalities(regs, p, 2)
    ji = np.array([checker.jaccard(regs[i], regs[j], p) for i, j in zip(*iu)])
    assert_close(su.jaccard_matrix(list(regs)), ji.astype(np.float32), what="jaccard_matrix")
    big = trip[:, 2] > 1e-3 * card.max()                                                   # (tiny intersections are cancellation residues)
    assert_close(su.intersection_matrix(regs)[big], trip[big, 2].astype(np.float32), what="intersection_matrix")
    union = np.array([checker.cardinalities(np.maximum(regs[i], regs[j])[None], p, 2)[0] for i, j in zip(*iu)])
    assert_close(su.union_size_matrix(regs), union.astype(np.float32), what="union_size_matrix")
    sc = trip[:, 2] / np.minimum(card[iu[0]], card[iu[1]])
    assert_close(su.symmetric_containment_matrix(regs)[big], sc[big].astype(np.float32), rtol=2e-6, what="symmetric_containment_matrix")
    cm = su.containment_matrix(regs)
    assert cm.shape == (n, n) and np.allclose(np.diag(cm), 1.0)
    want = np.zeros((n, n))
    for (i, j), t in zip(zip(*iu), trip):
        want[i, j] = t[2] / (t[2] + t[0])
        want[j, i] = t[2] / (t[2] + t[1])
    off = ~np.eye(n, dtype=bool) & (su.tri2full(trip[:, 2].astype(np.float32), 0.0) > 1e-3 * card.max())
    assert_close(cm[off], want[off].astype(np.float32), rtol=2e-6, what="containment_matrix")
