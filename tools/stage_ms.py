"""Per-kernel CUDA-event times of the resident replay step, without bench.py's result checks (kernel experiments):
   python tools/stage_ms.py [pairs] [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf-pl-slam_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gfpl import api, synth, _structs as S
import orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
NPT, NLN = int(os.environ.get("GFPL_NPT", "300")), int(os.environ.get("GFPL_NLN", "100"))   # GFPL_NPT=2000 GFPL_NLN=500: config C4
d = synth.make_pairs(B, n_pt=NPT, n_ln=NLN, seed=1000, pinned=True)
ctx = api.Context(0, cam=d["cam"])
prev = orc.stereo_match(d["cam"], ctx.cfg, d["prev_raw"]).repack()      # CPU: independent of the kernels under test
rp = api.Replay(ctx, B, NPT, NLN, prev.cap_pt, prev.cap_ln)
rp.upload(prev, d["state"], d["curr_raw"])
flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
acc = None
for k in range(steps + 2):
    rp.run(flags)
    ms = np.array(rp.stage_ms())
    if k >= 2:
        acc = ms if acc is None else acc + ms
names = ["knn_stereo_pt", "knn_stereo_ln", "stereo_accept", "f2f_pt_search", "knn_f2f_ln", "f2f_accept", "uncertainty+linecut", "pose_gn", "step"]
print({n: round(float(v) / steps, 4) for n, v in zip(names, acc)}, "pairs", rp.pair_counts())
# software pipeline over sub-batches (gfpl_replay_set_pipeline): ms per step of back-to-back runs, checksum of the poses
import zlib
ref = None
for n_sub in [int(x) for x in os.environ.get("GFPL_SWEEP", "1,2,4,6,8,12,16").split(",")]:
    rp.set_pipeline(n_sub)
    for _ in range(3):
        rp.run(flags)
    ctx.sync()
    ctx.timer_record(0)
    for _ in range(steps * 4):
        rp.run(flags)
    ctx.timer_record(1)
    ms = ctx.timer_elapsed_ms(0, 1) / (steps * 4)
    pose = np.zeros(B, S.pose_result_dt)
    rp.download(pose)
    ctx.sync()
    crc = zlib.crc32(pose.tobytes())
    ref = crc if ref is None else ref
    print("pipeline n_sub=%2d  %.3f ms/step  %.2f M pairs/s  poses %s" % (n_sub, ms, B / ms / 1e3, "identical" if crc == ref else "DIFFER"))
