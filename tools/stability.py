"""Repeat-run stability of the resident replay step (stands in for racecheck, which is closed on this GPU pool): the same
8192-pair batch is run N times; the CRC of the pose records of every step and of ALL outputs (observations, inlier masks, cut
records, stereo features and descriptors) of every 25th step must equal those of the first.  A race between the tensor-core
product pipeline, the TMEM read-back, the shared-memory minima or the warp-level selections would show as an intermittent
difference.   python tools/stability.py [steps] [pairs]"""
import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf-pl-slam_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gfpl import api, synth, _structs as S
import orc
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d = synth.make_pairs(B, seed=4242, ragged=True)
ctx = api.Context(0, cam=d["cam"])
prev = orc.stereo_match(d["cam"], ctx.cfg, d["prev_raw"])
rp = api.Replay(ctx, B, d["curr_raw"].cap_pt, d["curr_raw"].cap_ln)
rp.upload(prev, d["state"], d["curr_raw"])
flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
pose = np.zeros(B, S.pose_result_dt)
obs = S.TrackObsNP(B, ctx.cfg.max_point_match_num, ctx.cfg.max_line_match_num)
cut = np.zeros((B, obs.cap_ln), S.line_cut_dt)
curr = S.StereoFramesNP(B, d["curr_raw"].cap_pt, d["curr_raw"].cap_ln)


def crc_all():
    rp.download(pose, obs, cut, curr)
    ctx.sync()
    c = zlib.crc32(pose.tobytes())
    for p in range(0, B, max(1, B // 64)):      # every field of a spread of pairs, up to the counts
        npt, nln, cp, cl = obs.n_pt[p], obs.n_ln[p], curr.n_pt[p], curr.n_ln[p]
        for a in (obs.pt[p, :npt], obs.pt_inlier[p, :npt], obs.ln[p, :nln], obs.ln_inlier[p, :nln], cut[p, :nln],
                  curr.pt[p, :cp], curr.pdesc[p, :cp], curr.ln[p, :cl], curr.ldesc[p, :cl]):
            c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    c = zlib.crc32(obs.n_pt.tobytes() + obs.n_ln.tobytes() + curr.n_pt.tobytes() + curr.n_ln.tobytes(), c)
    return c


def crc_pose():
    rp.download(pose)
    ctx.sync()
    return zlib.crc32(pose.tobytes())


rp.run(flags)
ref_all, ref_pose = crc_all(), crc_pose()
bad_pose = bad_all = checked_all = 0
for k in range(1, steps + 1):
    rp.run(flags)
    bad_pose += crc_pose() != ref_pose
    if k % 25 == 0:
        checked_all += 1
        bad_all += crc_all() != ref_all
print("stability: %d steps of %d pairs (ragged counts): %d steps with different pose records, %d of %d full-output checks different; "
      "tracked %.3f" % (steps, B, bad_pose, bad_all, checked_all, float((pose["status"] == 0).mean())))
rp.close()
ctx.close()
sys.exit(1 if (bad_pose or bad_all) else 0)
