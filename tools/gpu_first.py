"""First GPU check: knn2/radius parity vs the oracle + measured POPC/FP64 peaks + a raw knn2 timing."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf-pl-slam_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from gfpl import api, synth, _structs as S
import orc
ctx = api.Context(0)
rng = np.random.default_rng(0)
ok = True
for (nq, nt) in [(300, 300), (100, 100), (2000, 2000), (500, 500), (1, 2), (37, 1025), (513, 255)]:
    for norm in (6, 7):
        for tie in (False, True):
            q = synth.tie_stress_descriptors(rng, nq) if tie else rng.integers(0, 256, (nq, 32), dtype=np.uint8)
            t = synth.tie_stress_descriptors(rng, nt) if tie else rng.integers(0, 256, (nt, 32), dtype=np.uint8)
            gi, gd = ctx.knn2(q, t, norm); oi, od = orc.knn2(q, t, norm)
            good = (gi == oi).all() and (gd == od).all()
            ok &= good
            if not good: print("KNN MISMATCH", nq, nt, norm, tie, (gi != oi).sum())
            r = 6.0 if tie else 110.0
            rc, off, ri, rd = ctx.radius(q, t, r, norm); rc2, off2, ri2, rd2 = orc.radius(q, t, r, norm)
            good = rc == rc2 == 0 and (off == off2).all() and (ri == ri2).all() and (rd == rd2).all()
            ok &= good
            if not good: print("RADIUS MISMATCH", nq, nt, norm, tie)
print("parity", ok)
p, f = ctx.measure_peaks()
print("peaks: popc %.3f T/s, fp64 %.2f TFLOP/s" % (p, f))
# batched knn timing (host-pointer call, includes copies) just for a first feel
B = 2048
q = rng.integers(0, 256, (B, 300, 32), dtype=np.uint8); t = rng.integers(0, 256, (B, 300, 32), dtype=np.uint8)
n = np.full(B, 300, np.int32)
ctx.knn2_batch(q, n, t, n)
t0 = time.time(); ctx.knn2_batch(q, n, t, n); dt = time.time() - t0
print("knn2_batch B=%d 300x300 incl copies: %.2f ms (%.1f Gpairs/s)" % (B, dt * 1e3, B * 9e4 / dt / 1e9))
json.dump({"popc_tops": p, "fp64_tflops": f, "parity": bool(ok)}, open(os.path.join(ROOT, "gpurun_out", "first.json"), "w"))
