"""ctypes / numpy mirrors of the structs in include/gfpl_c.h.

numpy structured dtypes (align=True) give the C layout of the record arrays; the `*NP` containers
own the arrays of one batched view and hand out the ctypes view struct the C ABI takes.
"""
import ctypes as C

import numpy as np

DESC_BYTES = 32
NORM_HAMMING = 6
NORM_HAMMING2 = 7

STEREO_INITIAL, STEREO_BRUTE = 0, 1
F2F_HYBRID, F2F_BRUTE = 0, 1
POSE_SEED_PREV_DT, POSE_SEED_GIVEN = 0, 1
TRACK_LINECUT, TRACK_F2F_BRUTE, TRACK_SEED_GIVEN, TRACK_STEREO_INITIAL = 1, 2, 4, 8
NUM_STAGES = 9

OK, E_INVALID, E_NO_DEVICE, E_CUDA, E_NOMEM, E_CAPACITY = 0, -1, -2, -3, -4, -5


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("b", C.c_double), ("width", C.c_int32), ("height", C.c_int32)]


class Config(C.Structure):
    _fields_ = [
        ("has_points", C.c_int32), ("has_lines", C.c_int32), ("lr_in_parallel", C.c_int32),
        ("best_lr_matches", C.c_int32), ("use_line_conf_cut", C.c_int32), ("max_vol_line_cut", C.c_int32),
        ("max_line_match_num", C.c_int32), ("max_point_match_num", C.c_int32), ("min_features", C.c_int32),
        ("max_iters", C.c_int32), ("max_iters_ref", C.c_int32), ("max_iters_lba", C.c_int32),
        ("ratio_disp_std", C.c_double), ("ratio_disp_std_hor", C.c_double), ("max_dist_epip", C.c_double),
        ("min_disp", C.c_double), ("max_ratio_12_p", C.c_double), ("point_match_radius", C.c_double),
        ("stereo_overlap_th", C.c_double), ("line_horiz_th", C.c_double), ("desc_th_l", C.c_double),
        ("line_cov_th", C.c_double), ("line_match_radius", C.c_double), ("homog_th", C.c_double),
        ("min_error", C.c_double), ("min_error_change", C.c_double), ("inlier_k", C.c_double),
        ("motion_step_th", C.c_double), ("orb_scale_factor", C.c_double), ("lsd_scale", C.c_double),
        ("lambda_lba_lm", C.c_double), ("lambda_lba_k", C.c_double), ("f2f_point_window", C.c_double),
        ("budget_slack", C.c_double), ("cut_step", C.c_double), ("cut_rng", C.c_double * 2),
        ("mad_uninit_value", C.c_double), ("max_kf_epip_p", C.c_double), ("max_kf_epip_l", C.c_double),
        ("orb_nlevels", C.c_int32), ("_pad0", C.c_int32),
    ]


keypoint_dt = np.dtype([("x", "f4"), ("y", "f4"), ("octave", "i4"), ("class_id", "i4")], align=True)
keyline_dt = np.dtype([("startPointX", "f4"), ("startPointY", "f4"), ("endPointX", "f4"), ("endPointY", "f4"),
                       ("angle", "f4"), ("octave", "i4"), ("lineLength", "f4"), ("response", "f4")], align=True)
point_feature_dt = np.dtype([("P", "f8", 3), ("pl", "f8", 2), ("disp", "f8"), ("sigma2", "f8"), ("level", "i4"),
                             ("idx", "i4"), ("src_l", "i4"), ("src_r", "i4")], align=True)
line_feature_dt = np.dtype([("sP", "f8", 3), ("eP", "f8", 3), ("spl", "f8", 2), ("epl", "f8", 2), ("sdisp", "f8"),
                            ("edisp", "f8"), ("le", "f8", 3), ("angle", "f8"), ("sigma2", "f8"), ("level", "i4"),
                            ("idx", "i4"), ("src_l", "i4"), ("src_r", "i4")], align=True)
line_cov_dt = np.dtype([("covS", "f8", 9), ("covE", "f8", 9)], align=True)
pt_obs_dt = np.dtype([("P", "f8", 3), ("pl_obs", "f8", 2), ("sigma2", "f8"), ("prev_idx", "i4"), ("curr_idx", "i4")],
                     align=True)
ln_obs_dt = np.dtype([("sP", "f8", 3), ("eP", "f8", 3), ("le_obs", "f8", 3), ("sigma2", "f8"), ("prev_idx", "i4"),
                      ("curr_idx", "i4")], align=True)
line_cut_dt = np.dtype([("cutRatio", "f8", 2), ("invCovPose", "f8", 36), ("spl", "f8", 2), ("epl", "f8", 2),
                        ("sdisp", "f8"), ("edisp", "f8")], align=True)
line_cut_descent_dt = np.dtype([("cutRatio", "f8", 2), ("volume", "f8"), ("n_volume", "i4"), ("fst_idx", "i4"), ("lst_idx", "i4"),
                                ("pad", "i4"), ("covlPt2D", "f8", (2, 9)), ("lPt2D", "f8", (4, 2)), ("spl", "f8", 2), ("epl", "f8", 2),
                                ("sdisp", "f8"), ("edisp", "f8")], align=True)
frame_state_dt = np.dtype([("Tfw", "f8", 16), ("DT", "f8", 16), ("DT_cov", "f8", 36), ("Tfw_cov", "f8", 36),
                           ("time_stamp", "f8"), ("time_stamp_curr", "f8")], align=True)
pose_result_dt = np.dtype([("DT", "f8", 16), ("Tfw", "f8", 16), ("DT_cov", "f8", 36), ("DT_cov_eig", "f8", 6),
                           ("Tfw_cov", "f8", 36), ("err_norm", "f8"), ("status", "i4"), ("n_inliers", "i4"),
                           ("n_inliers_pt", "i4"), ("n_inliers_ls", "i4"), ("iters_first", "i4"),
                           ("iters_ref", "i4")], align=True)

prev_point_dt = np.dtype([("P", "f8", 3), ("sigma2", "f8")], align=True)
prev_line_dt = np.dtype([("sP", "f8", 3), ("eP", "f8", 3), ("spl", "f8", 2), ("epl", "f8", 2), ("sdisp", "f8"),
                         ("edisp", "f8"), ("le0", "f8"), ("sigma2", "f8")], align=True)
assert prev_point_dt.itemsize == 32 and prev_line_dt.itemsize == 112
assert keypoint_dt.itemsize == 16 and keyline_dt.itemsize == 32
assert point_feature_dt.itemsize == 72 and line_feature_dt.itemsize == 152
assert line_cov_dt.itemsize == 144 and pt_obs_dt.itemsize == 56 and ln_obs_dt.itemsize == 88
assert line_cut_descent_dt.itemsize == 296
assert line_cut_dt.itemsize == 352 and frame_state_dt.itemsize == 848 and pose_result_dt.itemsize == 912

_vp = C.c_void_p


class RawFrames(C.Structure):
    _fields_ = [("B", C.c_int32), ("cap_pt", C.c_int32), ("cap_ln", C.c_int32),
                ("n_pt_l", _vp), ("n_pt_r", _vp), ("n_ln_l", _vp), ("n_ln_r", _vp),
                ("pt_l", _vp), ("pt_r", _vp), ("pdesc_l", _vp), ("pdesc_r", _vp),
                ("ln_l", _vp), ("ln_r", _vp), ("ldesc_l", _vp), ("ldesc_r", _vp)]


class StereoFrames(C.Structure):
    _fields_ = [("B", C.c_int32), ("cap_pt", C.c_int32), ("cap_ln", C.c_int32),
                ("n_pt", _vp), ("n_ln", _vp), ("pt", _vp), ("pdesc", _vp), ("ln", _vp), ("ldesc", _vp)]


MAX_PYR_LEVELS = 8


class Pyramid(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("width", C.c_int32 * MAX_PYR_LEVELS), ("height", C.c_int32 * MAX_PYR_LEVELS),
                ("row_stride", C.c_int32 * MAX_PYR_LEVELS), ("offset", C.c_int64 * MAX_PYR_LEVELS),
                ("frame_stride", C.c_int64), ("data", _vp)]


class PyramidNP:
    """right-image pyramids of B frames, levels packed back to back (row stride = width)"""

    def __init__(self, B, sizes):
        self.B, self.sizes = B, [(int(w), int(h)) for w, h in sizes]
        self.offsets = np.concatenate([[0], np.cumsum([w * h for w, h in self.sizes])]).astype(np.int64)
        self.data = np.zeros((B, int(self.offsets[-1])), np.uint8)

    def level(self, p, l):
        w, h = self.sizes[l]
        return self.data[p, self.offsets[l]:self.offsets[l + 1]].reshape(h, w)

    def view(self):
        v = Pyramid()
        v.n_levels = len(self.sizes)
        for l, (w, h) in enumerate(self.sizes):
            v.width[l], v.height[l], v.row_stride[l], v.offset[l] = w, h, w, int(self.offsets[l])
        v.frame_stride = int(self.offsets[-1])
        v.data = ptr(self.data)
        return v


class LocalMap(C.Structure):
    _fields_ = [("B", C.c_int32), ("cap_pt", C.c_int32), ("cap_ln", C.c_int32),
                ("n_pt", _vp), ("n_ln", _vp), ("pt3d", _vp), ("pdesc", _vp), ("ln3d", _vp), ("ldesc", _vp)]


class MapMatches(C.Structure):
    _fields_ = [("B", C.c_int32), ("cap_pt", C.c_int32), ("cap_ln", C.c_int32),
                ("n_pt", _vp), ("n_ln", _vp), ("pt", _vp), ("ln", _vp)]


map_pair_dt = np.dtype([("map_idx", "i4"), ("kf_idx", "i4")], align=True)


class PrevFrames(C.Structure):
    _fields_ = [("B", C.c_int32), ("cap_pt", C.c_int32), ("cap_ln", C.c_int32),
                ("n_pt", _vp), ("n_ln", _vp), ("pt", _vp), ("pdesc", _vp), ("ln", _vp), ("ldesc", _vp)]


class TrackObs(C.Structure):
    _fields_ = [("B", C.c_int32), ("cap_pt", C.c_int32), ("cap_ln", C.c_int32),
                ("n_pt", _vp), ("n_ln", _vp), ("pt", _vp), ("pt_inlier", _vp), ("ln", _vp), ("ln_inlier", _vp)]


class LbaProblem(C.Structure):
    _fields_ = [("Nkf", C.c_int32), ("Npt", C.c_int32), ("Nls", C.c_int32), ("n_pt_obs", C.c_int32),
                ("n_ls_obs", C.c_int32), ("kf_Tkw", _vp), ("X_pt", _vp), ("X_ls", _vp),
                ("pt_obs_lm", _vp), ("pt_obs_kf", _vp), ("pt_obs_fix", _vp), ("pt_obs_uv", _vp), ("pt_obs_s2", _vp),
                ("ls_obs_lm", _vp), ("ls_obs_kf", _vp), ("ls_obs_fix", _vp), ("ls_obs_le", _vp), ("ls_obs_s2", _vp),
                ("fix_Tkw", _vp), ("n_fix", C.c_int32)]


class LbaBlocks(C.Structure):
    _fields_ = [("Hkk", _vp), ("Hpp", _vp), ("Hll", _vp), ("Hpk", _vp), ("Hlk", _vp), ("g", _vp), ("err", _vp)]


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data


def _alloc(shape, dtype, pinned):
    """Host array; `pinned` allocates page-locked memory through torch (bench e2e path)."""
    if not pinned:
        return np.zeros(shape, dtype=dtype)
    import torch
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    buf = torch.zeros(max(n, 1), dtype=torch.uint8).pin_memory()
    # the numpy view keeps the pinned tensor's storage alive through its base chain
    return buf.numpy()[:n].view(dt).reshape(shape)


class RawFramesNP:
    """points_l/r, pdesc_l/r, lines_l/r, ldesc_l/r of B stereo frames."""

    def __init__(self, B, cap_pt, cap_ln, pinned=False):
        self.B, self.cap_pt, self.cap_ln = B, cap_pt, cap_ln
        a = lambda s, d: _alloc(s, d, pinned)
        self.n_pt_l, self.n_pt_r = a(B, "i4"), a(B, "i4")
        self.n_ln_l, self.n_ln_r = a(B, "i4"), a(B, "i4")
        self.pt_l, self.pt_r = a((B, cap_pt), keypoint_dt), a((B, cap_pt), keypoint_dt)
        self.pdesc_l, self.pdesc_r = a((B, cap_pt, 32), "u1"), a((B, cap_pt, 32), "u1")
        self.ln_l, self.ln_r = a((B, cap_ln), keyline_dt), a((B, cap_ln), keyline_dt)
        self.ldesc_l, self.ldesc_r = a((B, cap_ln, 32), "u1"), a((B, cap_ln, 32), "u1")

    def view(self):
        return RawFrames(self.B, self.cap_pt, self.cap_ln, ptr(self.n_pt_l), ptr(self.n_pt_r), ptr(self.n_ln_l),
                         ptr(self.n_ln_r), ptr(self.pt_l), ptr(self.pt_r), ptr(self.pdesc_l), ptr(self.pdesc_r),
                         ptr(self.ln_l), ptr(self.ln_r), ptr(self.ldesc_l), ptr(self.ldesc_r))

    def nbytes(self):
        return sum(getattr(self, k).nbytes for k in ("n_pt_l", "n_pt_r", "n_ln_l", "n_ln_r", "pt_l", "pt_r",
                                                      "pdesc_l", "pdesc_r", "ln_l", "ln_r", "ldesc_l", "ldesc_r"))

    def slice(self, lo, hi):
        out = RawFramesNP.__new__(RawFramesNP)
        out.B, out.cap_pt, out.cap_ln = hi - lo, self.cap_pt, self.cap_ln
        for k in ("n_pt_l", "n_pt_r", "n_ln_l", "n_ln_r", "pt_l", "pt_r", "pdesc_l", "pdesc_r", "ln_l", "ln_r",
                  "ldesc_l", "ldesc_r"):
            setattr(out, k, getattr(self, k)[lo:hi])
        return out


class LocalMapNP:
    """local landmarks handed to gfpl_map_match: world points / line endpoints + representative descriptors"""

    def __init__(self, B, cap_pt, cap_ln):
        self.B, self.cap_pt, self.cap_ln = B, cap_pt, cap_ln
        self.n_pt, self.n_ln = np.zeros(B, "i4"), np.zeros(B, "i4")
        self.pt3d, self.pdesc = np.zeros((B, cap_pt, 3)), np.zeros((B, cap_pt, 32), "u1")
        self.ln3d, self.ldesc = np.zeros((B, cap_ln, 6)), np.zeros((B, cap_ln, 32), "u1")

    def view(self):
        return LocalMap(self.B, self.cap_pt, self.cap_ln, ptr(self.n_pt), ptr(self.n_ln), ptr(self.pt3d),
                        ptr(self.pdesc), ptr(self.ln3d), ptr(self.ldesc))


class MapMatchesNP:
    def __init__(self, B, cap_pt, cap_ln):
        self.B, self.cap_pt, self.cap_ln = B, cap_pt, cap_ln
        self.n_pt, self.n_ln = np.zeros(B, "i4"), np.zeros(B, "i4")
        self.pt, self.ln = np.zeros((B, cap_pt), map_pair_dt), np.zeros((B, cap_ln), map_pair_dt)

    def view(self):
        return MapMatches(self.B, self.cap_pt, self.cap_ln, ptr(self.n_pt), ptr(self.n_ln), ptr(self.pt), ptr(self.ln))


class StereoFramesNP:
    """stereo_pt / stereo_ls and compacted descriptors of B stereo frames."""
    _keys = ("n_pt", "n_ln", "pt", "pdesc", "ln", "ldesc")

    def __init__(self, B, cap_pt, cap_ln, pinned=False):
        self.B, self.cap_pt, self.cap_ln = B, cap_pt, cap_ln
        a = lambda s, d: _alloc(s, d, pinned)
        self.n_pt, self.n_ln = a(B, "i4"), a(B, "i4")
        self.pt, self.pdesc = a((B, cap_pt), point_feature_dt), a((B, cap_pt, 32), "u1")
        self.ln, self.ldesc = a((B, cap_ln), line_feature_dt), a((B, cap_ln, 32), "u1")

    def view(self):
        return StereoFrames(self.B, self.cap_pt, self.cap_ln, ptr(self.n_pt), ptr(self.n_ln), ptr(self.pt),
                            ptr(self.pdesc), ptr(self.ln), ptr(self.ldesc))

    def nbytes(self):
        return sum(getattr(self, k).nbytes for k in self._keys)

    def slice(self, lo, hi):
        out = StereoFramesNP.__new__(StereoFramesNP)
        out.B, out.cap_pt, out.cap_ln = hi - lo, self.cap_pt, self.cap_ln
        for k in self._keys:
            setattr(out, k, getattr(self, k)[lo:hi])
        return out

    def copy(self):
        out = StereoFramesNP.__new__(StereoFramesNP)
        out.B, out.cap_pt, out.cap_ln = self.B, self.cap_pt, self.cap_ln
        for k in self._keys:
            setattr(out, k, getattr(self, k).copy())
        return out

    def repack(self, cap_pt=None, cap_ln=None, pinned=False):
        """same frames with tighter row strides (default: the largest count in the batch)"""
        cap_pt = cap_pt or max(int(self.n_pt.max()), 1)
        cap_ln = cap_ln or max(int(self.n_ln.max()), 1)
        assert cap_pt >= self.n_pt.max() and cap_ln >= self.n_ln.max()
        out = StereoFramesNP(self.B, cap_pt, cap_ln, pinned)
        out.n_pt[:], out.n_ln[:] = self.n_pt, self.n_ln
        out.pt[:], out.pdesc[:] = self.pt[:, :cap_pt], self.pdesc[:, :cap_pt]
        out.ln[:], out.ldesc[:] = self.ln[:, :cap_ln], self.ldesc[:, :cap_ln]
        return out


class PrevFramesNP:
    """slim previous-frame records (gfpl_prev_frames): what the tracking step reads of the previous stereo frame"""
    _keys = ("n_pt", "n_ln", "pt", "pdesc", "ln", "ldesc")

    def __init__(self, full, pinned=False):
        self.B, self.cap_pt, self.cap_ln = full.B, full.cap_pt, full.cap_ln
        a = lambda s, d: _alloc(s, d, pinned)
        self.n_pt, self.n_ln = a(self.B, "i4"), a(self.B, "i4")
        self.pt, self.pdesc = a((self.B, self.cap_pt), prev_point_dt), a((self.B, self.cap_pt, 32), "u1")
        self.ln, self.ldesc = a((self.B, self.cap_ln), prev_line_dt), a((self.B, self.cap_ln, 32), "u1")
        self.n_pt[:], self.n_ln[:] = full.n_pt, full.n_ln
        self.pt["P"], self.pt["sigma2"] = full.pt["P"], full.pt["sigma2"]
        self.pdesc[:], self.ldesc[:] = full.pdesc, full.ldesc
        for k in ("sP", "eP", "spl", "epl", "sdisp", "edisp", "sigma2"):
            self.ln[k] = full.ln[k]
        self.ln["le0"] = full.ln["le"][..., 0]

    def view(self):
        return PrevFrames(self.B, self.cap_pt, self.cap_ln, ptr(self.n_pt), ptr(self.n_ln), ptr(self.pt),
                          ptr(self.pdesc), ptr(self.ln), ptr(self.ldesc))


keypoint8_dt = np.dtype([("x", "f4"), ("y", "f4")], align=True)
keyline20_dt = np.dtype([("startPointX", "f4"), ("startPointY", "f4"), ("endPointX", "f4"), ("endPointY", "f4"),
                         ("angle", "f4")], align=True)
prev_point12_dt = np.dtype([("x", "f4"), ("y", "f4"), ("disp", "f4")], align=True)
prev_line32_dt = np.dtype([("spl", "f4", 2), ("epl", "f4", 2), ("sdisp", "f8"), ("edisp", "f8")], align=True)
assert keypoint8_dt.itemsize == 8 and keyline20_dt.itemsize == 20 and prev_point12_dt.itemsize == 12
assert prev_line32_dt.itemsize == 32

_COMPACT_KEYS = ("n_pt_l", "n_pt_r", "n_ln_l", "n_ln_r", "pt_l", "pt_r", "pt_oct_l", "pt_oct_r", "pdesc_l", "pdesc_r", "ln_l", "ln_r",
                 "ln_oct_l", "ln_oct_r", "ldesc_l", "ldesc_r", "prev_n_pt", "prev_n_ln", "prev_pt", "prev_pt_level", "prev_pdesc",
                 "prev_ln", "prev_ln_level", "prev_ldesc")


class CompactInput(C.Structure):
    _fields_ = [("B", C.c_int32), ("cap_pt", C.c_int32), ("cap_ln", C.c_int32), ("prev_cap_pt", C.c_int32),
                ("prev_cap_ln", C.c_int32)] + [(k, _vp) for k in _COMPACT_KEYS]


class CompactInputNP:
    """gfpl_compact_input of one step: raw current frames + previous stereo frames in their smallest exact form"""
    _keys = _COMPACT_KEYS

    def __init__(self, raw, prev_full=None, pinned=False):
        """prev_full=None: current frames only (sequence-chained replay, gfpl_replay_upload_curr_compact)"""
        B = raw.B
        self.B, self.cap_pt, self.cap_ln = B, raw.cap_pt, raw.cap_ln
        self.prev_cap_pt, self.prev_cap_ln = (prev_full.cap_pt, prev_full.cap_ln) if prev_full is not None else (raw.cap_pt, raw.cap_ln)
        a = lambda s, d: _alloc(s, d, pinned)

        def u8(x):
            assert x.min() >= 0 and x.max() <= 255
            out = a(x.shape, "u1"); out[:] = x
            return out
        for k in ("n_pt_l", "n_pt_r", "n_ln_l", "n_ln_r"):
            setattr(self, k, a(B, "i4")); getattr(self, k)[:] = getattr(raw, k)
        for side in ("l", "r"):
            kp = a((B, raw.cap_pt), keypoint8_dt); src = getattr(raw, "pt_" + side)
            kp["x"], kp["y"] = src["x"], src["y"]
            setattr(self, "pt_" + side, kp); setattr(self, "pt_oct_" + side, u8(src["octave"]))
            kl = a((B, raw.cap_ln), keyline20_dt); src = getattr(raw, "ln_" + side)
            for f in ("startPointX", "startPointY", "endPointX", "endPointY", "angle"):
                kl[f] = src[f]
            setattr(self, "ln_" + side, kl); setattr(self, "ln_oct_" + side, u8(src["octave"]))
            for k in ("pdesc_", "ldesc_"):
                d = a(getattr(raw, k + side).shape, "u1"); d[:] = getattr(raw, k + side); setattr(self, k + side, d)
        if prev_full is None:
            for k in self._keys:
                if k.startswith("prev_"):
                    setattr(self, k, None)
            return
        self.prev_n_pt, self.prev_n_ln = a(B, "i4"), a(B, "i4")
        self.prev_n_pt[:], self.prev_n_ln[:] = prev_full.n_pt, prev_full.n_ln
        self.prev_pdesc, self.prev_ldesc = a(prev_full.pdesc.shape, "u1"), a(prev_full.ldesc.shape, "u1")
        self.prev_pdesc[:], self.prev_ldesc[:] = prev_full.pdesc, prev_full.ldesc
        ln = a((B, prev_full.cap_ln), prev_line32_dt)
        for k in ("spl", "epl", "sdisp", "edisp"):
            ln[k] = prev_full.ln[k]
        # exact only for uncut stereo-stage lines: float end points, sP / eP = backProjection(end point, disparity)
        assert np.array_equal(ln["spl"].astype("f8"), prev_full.ln["spl"]) and np.array_equal(ln["epl"].astype("f8"), prev_full.ln["epl"])
        self.prev_ln, self.prev_ln_level = ln, u8(np.where(np.arange(prev_full.cap_ln)[None] < prev_full.n_ln[:, None], prev_full.ln["level"], 0))
        pp = a((B, prev_full.cap_pt), prev_point12_dt)
        pl, disp = prev_full.pt["pl"], prev_full.pt["disp"]
        pp["x"], pp["y"], pp["disp"] = pl[..., 0], pl[..., 1], disp
        # the compact form is exact only when pl / disp are float values (output of the stereo stage)
        assert np.array_equal(pp["x"].astype("f8"), pl[..., 0]) and np.array_equal(pp["disp"].astype("f8"), disp)
        self.prev_pt, self.prev_pt_level = pp, u8(np.where(np.arange(prev_full.cap_pt)[None] < prev_full.n_pt[:, None], prev_full.pt["level"], 0))

    def view(self):
        return CompactInput(self.B, self.cap_pt, self.cap_ln, self.prev_cap_pt, self.prev_cap_ln,
                            *[ptr(getattr(self, k)) if getattr(self, k) is not None else None for k in self._keys])

    def nbytes(self):
        return sum(getattr(self, k).nbytes for k in self._keys if getattr(self, k) is not None)


class TrackObsNP:
    """matched_pt / matched_ls of B frame pairs as observation lists + inlier masks."""
    _keys = ("n_pt", "n_ln", "pt", "pt_inlier", "ln", "ln_inlier")

    def __init__(self, B, cap_pt, cap_ln, pinned=False):
        self.B, self.cap_pt, self.cap_ln = B, cap_pt, cap_ln
        a = lambda s, d: _alloc(s, d, pinned)
        self.n_pt, self.n_ln = a(B, "i4"), a(B, "i4")
        self.pt, self.pt_inlier = a((B, cap_pt), pt_obs_dt), a((B, cap_pt), "u1")
        self.ln, self.ln_inlier = a((B, cap_ln), ln_obs_dt), a((B, cap_ln), "u1")

    def view(self):
        return TrackObs(self.B, self.cap_pt, self.cap_ln, ptr(self.n_pt), ptr(self.n_ln), ptr(self.pt),
                        ptr(self.pt_inlier), ptr(self.ln), ptr(self.ln_inlier))

    def copy(self):
        out = TrackObsNP.__new__(TrackObsNP)
        out.B, out.cap_pt, out.cap_ln = self.B, self.cap_pt, self.cap_ln
        for k in self._keys:
            setattr(out, k, getattr(self, k).copy())
        return out


_LBA_FIELDS = (("kf_Tkw", "f8"), ("X_pt", "f8"), ("X_ls", "f8"), ("pt_obs_lm", "i4"), ("pt_obs_kf", "i4"),
               ("pt_obs_fix", "i4"), ("pt_obs_uv", "f8"), ("pt_obs_s2", "f8"), ("ls_obs_lm", "i4"),
               ("ls_obs_kf", "i4"), ("ls_obs_fix", "i4"), ("ls_obs_le", "f8"), ("ls_obs_s2", "f8"), ("fix_Tkw", "f8"))


def lba_problem_from_arrays(a):
    """dict of numpy arrays (named like gfpl_lba_problem's pointer fields) -> (LbaProblem, keepalive dict)."""
    keep = {k: np.ascontiguousarray(a[k], dt) for k, dt in _LBA_FIELDS}
    p = LbaProblem()
    p.Nkf = len(keep["kf_Tkw"].reshape(-1, 16))
    p.Npt = len(keep["X_pt"].reshape(-1, 3))
    p.Nls = len(keep["X_ls"].reshape(-1, 6))
    p.n_pt_obs = len(keep["pt_obs_lm"])
    p.n_ls_obs = len(keep["ls_obs_lm"])
    p.n_fix = len(keep["fix_Tkw"].reshape(-1, 16))
    for k, _ in _LBA_FIELDS:
        setattr(p, k, ptr(keep[k]))
    return p, keep
