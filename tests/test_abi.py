"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/gfpl_c.h
declares, struct layouts agree between C, ctypes and numpy, defaults agree with the oracle's restatement
of config.cpp, there is no CPU fallback, and the product never touches oracle/."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gfpl_c.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gfpl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gfpl import api
    L = api.lib()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.gfpl_abi_version() == 1


def test_struct_layouts_match_c(tmp_path):
    from gfpl import _structs as S
    structs = {"gfpl_camera": C.sizeof(S.Camera), "gfpl_config": C.sizeof(S.Config),
               "gfpl_keypoint": S.keypoint_dt.itemsize, "gfpl_keyline": S.keyline_dt.itemsize,
               "gfpl_point_feature": S.point_feature_dt.itemsize, "gfpl_line_feature": S.line_feature_dt.itemsize,
               "gfpl_line_cov": S.line_cov_dt.itemsize, "gfpl_pt_obs": S.pt_obs_dt.itemsize,
               "gfpl_ln_obs": S.ln_obs_dt.itemsize, "gfpl_line_cut": S.line_cut_dt.itemsize,
               "gfpl_line_cut_descent": S.line_cut_descent_dt.itemsize,
               "gfpl_frame_state": S.frame_state_dt.itemsize, "gfpl_pose_result": S.pose_result_dt.itemsize,
               "gfpl_raw_frames": C.sizeof(S.RawFrames), "gfpl_stereo_frames": C.sizeof(S.StereoFrames),
               "gfpl_track_obs": C.sizeof(S.TrackObs), "gfpl_lba_problem": C.sizeof(S.LbaProblem),
               "gfpl_lba_blocks": C.sizeof(S.LbaBlocks), "gfpl_local_map": C.sizeof(S.LocalMap),
               "gfpl_map_matches": C.sizeof(S.MapMatches), "gfpl_map_pair": S.map_pair_dt.itemsize, "gfpl_pyramid": C.sizeof(S.Pyramid), "gfpl_compact_input": C.sizeof(S.CompactInput),
               "gfpl_keypoint8": S.keypoint8_dt.itemsize, "gfpl_keyline20": S.keyline20_dt.itemsize,
               "gfpl_prev_point12": S.prev_point12_dt.itemsize, "gfpl_prev_line32": S.prev_line32_dt.itemsize}
    prog = '#include <stdio.h>\n#include "gfpl_c.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in structs) + \
        'printf("off_sigma2 %zu\\n", offsetof(gfpl_point_feature, sigma2));' \
        'printf("off_le %zu\\n", offsetof(gfpl_line_feature, le));' \
        'printf("off_status %zu\\n", offsetof(gfpl_pose_result, status));return 0;}'
    src = tmp_path / "sz.c"
    src.write_text(prog)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])   # header is plain C
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, sz in structs.items():
        assert int(got[n]) == sz, n
    assert int(got["off_sigma2"]) == S.point_feature_dt.fields["sigma2"][1]
    assert int(got["off_le"]) == S.line_feature_dt.fields["le"][1]
    assert int(got["off_status"]) == S.pose_result_dt.fields["status"][1]


def test_config_defaults_match_reference_values(orc):
    from gfpl import api
    a, b = api.default_config(), orc.default_config()
    assert bytes(a) == bytes(b)
    # reference src/config.cpp:26-154
    assert (a.max_point_match_num, a.max_line_match_num, a.min_features, a.max_iters, a.max_iters_ref) == (500, 300, 10, 5, 10)
    assert (a.max_ratio_12_p, a.point_match_radius, a.desc_th_l, a.line_cov_th, a.inlier_k) == (0.9, 50.0, 0.1, 10.0, 2.0)
    assert (a.ratio_disp_std, a.ratio_disp_std_hor, a.homog_th, a.orb_scale_factor) == (0.15, 0.9, 1e-7, 1.2)
    cam = api.euroc_camera()
    assert (cam.fx, cam.b, cam.width, cam.height) == (458.654, 0.110078, 752, 480)       # config/euroc_params.yaml
    cam = api.kitti_camera()
    assert (cam.fx, cam.b, cam.width, cam.height) == (718.856, 0.537166, 1241, 376)      # config/kitti/kitti00-02.yaml


def test_no_cpu_fallback():
    """without a CUDA device the context cannot be created and the error says why"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gfpl import api
    with pytest.raises(api.GfplError) as e:
        api.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "gf-pl-slam_b200")):
        if os.sep + "lib" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"import orc|from orc|oracle/|gfpl_oracle|orc_[a-z]", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_runtime_files_do_not_read_the_reference_tree():
    for f in ("bench.py", "__graft_entry__.py"):
        txt = open(os.path.join(ROOT, f)).read()
        assert "/root/reference/" not in txt.replace("/root/reference needs", "")
