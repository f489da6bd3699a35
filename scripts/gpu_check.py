"""Ad-hoc first GPU validation: sketch + dist vs the oracle, plus rough timings (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dashing_b200 import capi, synth
from oracle import oracle as O

chk = O.best()
print("checker:", chk.kind, "devices:", capi.device_count())

# ---- sketch parity
rng = np.random.default_rng(11)
gs = synth.genomes(3, 6, 300_000, group=3)
gs[1] = synth.sprinkle(rng, gs[1])
multi = [gs[2][:1000].tobytes(), gs[2][1000:1010].tobytes(), b"", gs[2][1010:150000].tobytes(), b"ACGT"]
genomes = [gs[0], gs[1], multi, gs[3], gs[4][:25], gs[5]]
for k in (31, 21, 32, 4, 1):
    for p in (10, 14, 16, 18):
        for canon in (True, False):
            got = capi.sketch_genomes(genomes, k, p, canon)
            for gi, g in enumerate(genomes):
                recs = g if isinstance(g, list) else [g.tobytes()]
                exp = chk.sketch(recs, k, p, canon)
                if not (got[gi] == exp).all():
                    print("SKETCH MISMATCH", k, p, canon, gi, int((got[gi] != exp).sum()))
print("sketch parity done")

# ---- dist parity
for p in (10, 12, 14, 16):
    regs = np.concatenate([synth.registers(7, 70, p, card=3e5 if p < 14 else 5e6), synth.adversarial_registers(3, p)])
    n = len(regs)
    cards = capi.cardinalities(regs, p)
    exp_c = chk.cardinalities(regs, p, 2)
    ok = (cards == exp_c) | (np.abs(cards - exp_c) <= 1e-12 * np.abs(exp_c))
    print("p", p, "cards ok", ok.all())
    for estim in (0, 1, 2):
        for rtype in range(9):
            got = capi.dist_symmetric(regs, p, k=31, estim=estim, result_type=rtype).astype(np.float64)
            exp = chk.dist_rows(regs, p, k=31, estim=estim, jestim=2, rtype=rtype).astype(np.float64)
            same = (got == exp) | (np.isnan(got) & np.isnan(exp))
            scale = 1.0 if rtype != 2 else float(np.nanmax(exp_c[np.isfinite(exp_c)]))
            err = np.abs(got - exp) / np.maximum(np.maximum(np.abs(got), np.abs(exp)), scale)
            err[same] = 0
            bad = ~(err <= 1e-6)
            if bad.any():
                idx = np.nonzero(bad)[0][:4]
                print("DIST MISMATCH p", p, "estim", estim, "rtype", rtype, int(bad.sum()), "of", len(got), got[idx], exp[idx])
    # rect
    got = capi.dist_rect(regs[:50], regs[50:], p, result_type=1)
    exp = chk.dist_rect(regs[:50], regs[50:], p, rtype=1)
    print("p", p, "rect max abs diff", np.nanmax(np.abs(got.astype(np.float64) - exp)))
print("dist parity done")

# ---- rough timing
import torch
dev = torch.device("cuda:0")
p = 14
regs = synth.registers(5, 2048, p)
t0 = time.time(); out = capi.dist_symmetric(regs, p); t1 = time.time()
print("dist e2e n=2048: %.3fs -> %.3g pairs/s" % (t1 - t0, 2048 * 2047 / 2 / (t1 - t0)))
d_regs = torch.from_numpy(regs).to(dev)
n = 2048
d_out = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device=dev)
plan = capi.DistPlan(0)
st = torch.cuda.current_stream().cuda_stream
plan.prepare_dev(d_regs.data_ptr(), n, p, 2, st)
prm = capi.dist_params(p)
for _ in range(2): plan.run_symmetric_dev(prm, 0, n, d_out.data_ptr(), st)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); plan.run_symmetric_dev(prm, 0, n, d_out.data_ptr(), st); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("dist kernel n=2048 p=14: %.3f ms -> %.3g pairs/s; info" % (ms, n * (n - 1) / 2 / ms * 1e3), plan.last_run_info())
assert np.array_equal(d_out.cpu().numpy(), out)
e0.record(); plan.prepare_dev(d_regs.data_ptr(), n, p, 2, st); e1.record(); torch.cuda.synchronize()
print("prepare: %.3f ms" % e0.elapsed_time(e1))

gs = synth.genomes(9, 64, 1_000_000)
bases, offs, grb = capi.records_layout(gs)
t0 = time.time(); r = capi.sketch_batch(bases, offs, grb, 31, 14); t1 = time.time()
print("sketch e2e 64x1Mbp: %.3fs" % (t1 - t0))
pg = capi.PackedGenomes(bases, offs, grb, 31)
d_r = torch.empty((64, 1 << 14), dtype=torch.uint8, device=dev)
for _ in range(2): pg.sketch_dev(14, True, d_r.data_ptr(), st)
e0.record(); pg.sketch_dev(14, True, d_r.data_ptr(), st); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("sketch kernel: %.3f ms -> %.3g kmers/s (%d kmers)" % (ms, pg.kmers / ms * 1e3, pg.kmers))
assert np.array_equal(d_r.cpu().numpy(), r)
exp = chk.sketch([gs[5].tobytes()], 31, 14, True)
print("sketch big parity:", (r[5] == exp).all())
