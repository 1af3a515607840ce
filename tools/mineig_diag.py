"""Diagnostic (GPU box): where do the min-eigenvalue cut decisions of linecut_kernel leave the oracle's?
   python tools/mineig_diag.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf-pl-slam_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gfpl import api, synth, _structs as S
import orc

B = 64
d = synth.make_pairs(B, seed=501)
for max_vol in (1, 0):
    cfg = api.default_config()
    cfg.max_vol_line_cut = max_vol
    c = api.Context(0, cam=d["cam"], cfg=cfg)
    prev = orc.stereo_match(d["cam"], cfg, d["prev_raw"])
    curr = orc.stereo_match(d["cam"], cfg, d["curr_raw"])
    obs = orc.f2f_match(d["cam"], cfg, prev, curr, d["state"])
    cov = orc.stereo_uncertainty(d["cam"], cfg, prev)
    Tc = d["state"]["Tfw"].reshape(-1, 4, 4) @ d["state"]["DT"].reshape(-1, 4, 4)
    DT_inv = (np.linalg.inv(Tc) @ d["state"]["Tfw"].reshape(-1, 4, 4)).reshape(-1, 16)
    og, oo = obs.copy(), obs.copy()
    cg = c.linecut(DT_inv, og, cov)
    co = orc.linecut(d["cam"], cfg, DT_inv, oo, cov)
    nbit = ntot = nflip = 0
    first = None
    for p in range(B):
        n = obs.n_ln[p]
        for l in range(n):
            same = np.array_equal(cg["cutRatio"][p, l], co["cutRatio"][p, l])
            if not same:
                nflip += 1
                if first is None:
                    first = (p, l)
            elif first is None or first[0] != p:
                ntot += 1
                nbit += cg["invCovPose"][p, l].tobytes() == co["invCovPose"][p, l].tobytes()
    print("max_vol", max_vol, "lines with different decisions", nflip, "| invCovPose bit-identical on", nbit, "of", ntot, "lines")
    if first:
        p, l = first
        print(" first flip pair", p, "line", l, "gpu", cg["cutRatio"][p, l], "oracle", co["cutRatio"][p, l])
        for k in range(max(0, l - 2), l + 1):
            a, b = cg["invCovPose"][p, k], co["invCovPose"][p, k]
            print("  line", k, "cut", cg["cutRatio"][p, k], co["cutRatio"][p, k], "max rel diff", np.abs(a - b).max() / np.abs(b).max(),
                  "min eig", np.linalg.eigvalsh(b.reshape(6, 6))[0])
