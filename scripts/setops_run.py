"""Small driver for profiling the set-operation and FASTA-parsing kernels (ncu / compute-sanitizer targets)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dashing_b200 import capi, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(os.environ.get("N", "4000"))
if what in ("union", "all"):
    p = 14
    regs = synth.registers(1, 64, p)
    regs = np.ascontiguousarray(np.tile(regs, (n // 64, 1)))
    out = capi.union(regs, p)
    t = time.perf_counter(); out = capi.union(regs, p); dt = time.perf_counter() - t
    assert np.array_equal(out, regs[:64].max(axis=0))
    print(f"union of {regs.shape[0]} p={p} sketches ({regs.nbytes >> 20} MB, host buffers): {dt * 1e3:.2f} ms")
    c = capi.compress(regs[:512], p, 10)
    t = time.perf_counter(); c = capi.compress(regs[:512], p, 10); dt = time.perf_counter() - t
    print(f"compress 512 sketches p=14 -> 10: {dt * 1e3:.2f} ms")
if what in ("fasta", "all"):
    L, ng = 2_000_000, int(os.environ.get("NG", "40"))
    gs = synth.genomes(3, ng, L, group=8)
    files = []
    for i, g in enumerate(gs):
        lines = np.full(L // 80 * 81, 10, dtype=np.uint8)
        lines.reshape(-1, 81)[:, :80] = g.reshape(-1, 80)
        files.append([f">g{i}\n".encode() + lines.tobytes()])
    got, st = capi.sketch_fasta(files, 31, 14, True)
    t = time.perf_counter(); got, st = capi.sketch_fasta(files, 31, 14, True); dt = time.perf_counter() - t
    want = capi.sketch_genomes(gs, 31, 14, True)
    assert np.array_equal(got, want) and not st.any()
    print(f"sketch_fasta {ng} x {L} bp (pageable text, layout built per call): {dt * 1e3:.1f} ms")
