"""cross-step pipelined e2e with 2/3/4 slots; checks where the residual ms go"""
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
prev = prev.repack(pinned=True)
flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
for nslot in (2, 3, 4):
    ctxs = [api.Context(0, cam=d["cam"]) for _ in range(nslot)]
    rps = [api.Replay(c, B, 300, 100, prev.cap_pt, prev.cap_ln) for c in ctxs]
    poses = [S._alloc(B, S.pose_result_dt, True) for _ in range(nslot)]
    def run(n, lag):
        for k in range(n):
            s = k % nslot
            rps[s].upload(prev, d["state"], d["curr_raw"]); rps[s].run(flags); rps[s].download(poses[s])
            if k >= lag: ctxs[(k - lag) % nslot].sync()
        for c in ctxs: c.sync()
    for lag in range(1, nslot):
        run(4, lag)
        t0 = time.perf_counter(); run(12, lag); ms = (time.perf_counter() - t0) / 12 * 1e3
        print("slots %d lag %d: %.2f ms/step -> %.0f pairs/s" % (nslot, lag, ms, B / ms * 1e3))
    # upload-only and run-only concurrently on two contexts
    t0 = time.perf_counter()
    for _ in range(8):
        rps[0].upload(prev, d["state"], d["curr_raw"]); rps[1].run(flags)
    ctxs[0].sync(); ctxs[1].sync()
    print("   concurrent upload(ctx0) + run(ctx1): %.2f ms per pair of ops" % ((time.perf_counter() - t0) / 8 * 1e3))
    del rps, ctxs
