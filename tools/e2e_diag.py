"""e2e decomposition: pure H2D time, pure run time, chunked pipeline with different chunk counts / contexts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf-pl-slam_b200", "python"))
import numpy as np
from gfpl import api, synth, _structs as S
B = 8192
d = synth.make_pairs(B, seed=1000, pinned=True)
ctx = api.Context(0, cam=d["cam"])
prev = S.StereoFramesNP(B, 300, 100, pinned=True)
for lo in range(0, B, 1024):
    part = ctx.stereo_match(d["prev_raw"].slice(lo, lo + 1024))
    for k in S.StereoFramesNP._keys: getattr(prev, k)[lo:lo + 1024] = getattr(part, k)
flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
rp = api.Replay(ctx, B, 300, 100)
pose = S._alloc(B, S.pose_result_dt, True)
def t(f, n=5):
    f(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(n): f()
    ctx.sync()
    return (time.perf_counter() - t0) / n * 1e3
up = t(lambda: rp.upload(prev, d["state"], d["curr_raw"]))
print("upload only: %.2f ms -> %.1f GB/s (%d MB)" % (up, rp.h2d_bytes() / up / 1e6, rp.h2d_bytes() >> 20))
print("run only: %.2f ms" % t(lambda: rp.run(flags)))
print("download pose only: %.3f ms" % t(lambda: rp.download(pose)))
import torch
x = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); y = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
y.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): y.copy_(x, non_blocking=True)
torch.cuda.synchronize()
print("torch pinned H2D 512MB: %.1f GB/s" % (5 * 512 / 1024 / (time.perf_counter() - t0) * 1.0737))
for nctx in (1, 2, 3):
    ctxs = [ctx] + [api.Context(0, cam=d["cam"]) for _ in range(nctx - 1)]
    for nchunk in (2, 4, 8, 16):
        bounds = np.linspace(0, B, nchunk + 1).astype(int)
        chunks = []
        for i in range(nchunk):
            lo, hi = int(bounds[i]), int(bounds[i + 1]); c = ctxs[i % nctx]
            chunks.append((api.Replay(c, hi - lo, 300, 100), prev.slice(lo, hi), d["state"][lo:hi], d["curr_raw"].slice(lo, hi), pose[lo:hi]))
        def step():
            for r, pv, st, raw, ph in chunks:
                r.upload(pv, st, raw); r.run(flags); r.download(ph)
            for c in ctxs: c.sync()
        step(); step()
        t0 = time.perf_counter()
        for _ in range(5): step()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print("ctx %d chunks %2d: %.2f ms/step -> %.0f pairs/s" % (nctx, nchunk, ms, B / ms * 1e3))
        del chunks
