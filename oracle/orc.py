"""TEST INFRASTRUCTURE — ctypes wrapper of the CPU oracle (oracle/_build/libgfpl_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
The product package (gf-pl-slam_b200/python/gfpl) never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "gf-pl-slam_b200", "python"))
from gfpl import _structs as S  # noqa: E402  (struct mirrors only — no compute)

_LIB = None


def build(force=False):
    if os.environ.get("GFPL_ORACLE_LIB"):     # another build of the same sources (the ASan / UBSan build of `make asan`)
        return os.environ["GFPL_ORACLE_LIB"]
    so = os.path.join(_HERE, "_build", "libgfpl_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_line_descriptor_mad.restype = C.c_double
        _LIB.orc_budget_thres.restype = C.c_double
        _LIB.orc_vector_stdv_mad.restype = C.c_double
        _LIB.orc_logdet6.restype = C.c_double
        _LIB.orc_min_eig6.restype = C.c_double
        _LIB.orc_descent_volume.restype = C.c_double
    return _LIB


NATIVE_FLAGS = "-O3 -march=native (the reference's CMAKE_CXX_FLAGS, CMakeLists.txt:67-68)"
PARITY_FLAGS = "-O3 -march=x86-64-v2 -mpopcnt -ffp-contract=off"


def use_native_build():
    """bench.py's CPU arm: rebuild the oracle sources with -O3 -march=native on THIS machine and route every call of this
    module to that library (timing only; the parity tests keep the -ffp-contract=off build).  Returns the flags in use;
    falls back to the parity build if the compiler is missing."""
    global _LIB
    so = os.path.join(_HERE, "_build", "native", "libgfpl_oracle.so")
    try:
        # always rebuilt: a library built with -march=native on another machine (the snapshot travels) may not run here
        if os.path.exists(so):
            os.remove(so)
        subprocess.check_call(["make", "-C", _HERE, "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _LIB = None
        native = C.CDLL(so)
        lib_parity = lib()
        for name in ("orc_line_descriptor_mad", "orc_budget_thres", "orc_vector_stdv_mad", "orc_logdet6", "orc_min_eig6", "orc_descent_volume"):
            getattr(native, name).restype = C.c_double
        _LIB = native
        return NATIVE_FLAGS
    except Exception:
        _LIB = None
        lib()
        return PARITY_FLAGS


def default_config():
    cfg = S.Config()
    lib().orc_config_default(C.byref(cfg))
    return cfg


def euroc_camera():
    return S.Camera(458.654, 458.654, 367.215, 248.375, 0.110078, 752, 480)


def kitti_camera():
    return S.Camera(718.856, 718.856, 607.1928, 185.2157, 0.537166, 1241, 376)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2 and a.shape[1] == 32
    return a


def knn2(q, t, norm=S.NORM_HAMMING):
    q, t = _u8(q), _u8(t)
    idx = np.empty((len(q), 2), np.int32)
    dist = np.empty((len(q), 2), np.float32)
    rc = lib().orc_knn2_hamming256(C.c_void_p(S.ptr(q)), len(q), C.c_void_p(S.ptr(t)), len(t), norm,
                                   C.c_void_p(S.ptr(idx)), C.c_void_p(S.ptr(dist)))
    if rc != 0:
        raise ValueError(f"orc_knn2_hamming256 -> {rc}")
    return idx, dist


def radius(q, t, r, norm=S.NORM_HAMMING, cap=None):
    q, t = _u8(q), _u8(t)
    cap = len(q) * len(t) if cap is None else cap
    off = np.empty(len(q) + 1, np.int32)
    idx = np.empty(max(cap, 1), np.int32)
    dist = np.empty(max(cap, 1), np.float32)
    rc = lib().orc_radius_hamming256(C.c_void_p(S.ptr(q)), len(q), C.c_void_p(S.ptr(t)), len(t), norm, C.c_float(r),
                                     C.c_void_p(S.ptr(off)), C.c_void_p(S.ptr(idx)), C.c_void_p(S.ptr(dist)), cap)
    return rc, off, idx[:off[-1]] if rc == 0 else idx, dist[:off[-1]] if rc == 0 else dist


def line_descriptor_mad(d, U=0.0):
    d = np.ascontiguousarray(d, np.float32)
    return lib().orc_line_descriptor_mad(C.c_void_p(S.ptr(d)), len(d), C.c_double(U))


def budget_thres(d, max_num):
    d = np.ascontiguousarray(d, np.float32)
    return lib().orc_budget_thres(C.c_void_p(S.ptr(d)), len(d), max_num)


def vector_stdv_mad(r):
    r = np.ascontiguousarray(r, np.float64)
    return lib().orc_vector_stdv_mad(C.c_void_p(S.ptr(r)), len(r))


def stereo_match(cam, cfg, raw, variant=S.STEREO_BRUTE):
    out = S.StereoFramesNP(raw.B, raw.cap_pt, raw.cap_ln)
    rv, ov = raw.view(), out.view()
    rc = lib().orc_stereo_match(C.byref(cam), C.byref(cfg), C.byref(rv), variant, C.byref(ov))
    assert rc == 0, rc
    return out


def stereo_match_orbslam(cam, cfg, raw, pyr):
    out = S.StereoFramesNP(raw.B, raw.cap_pt, raw.cap_ln)
    rv, pv, ov = raw.view(), pyr.view(), out.view()
    rc = lib().orc_stereo_match_orbslam(C.byref(cam), C.byref(cfg), C.byref(rv), C.byref(pv), C.byref(ov))
    assert rc == 0, rc
    return out


def stereo_uncertainty(cam, cfg, frames):
    out = np.zeros((frames.B, frames.cap_ln), S.line_cov_dt)
    fv = frames.view()
    rc = lib().orc_stereo_uncertainty(C.byref(cam), C.byref(cfg), C.byref(fv), C.c_void_p(S.ptr(out)))
    assert rc == 0, rc
    return out


def f2f_match(cam, cfg, prev, curr, state, mode=S.F2F_HYBRID):
    out = S.TrackObsNP(prev.B, cfg.max_point_match_num, cfg.max_line_match_num)
    pv, cv, ov = prev.view(), curr.view(), out.view()
    rc = lib().orc_f2f_match(C.byref(cam), C.byref(cfg), C.byref(pv), C.byref(cv), C.c_void_p(S.ptr(state)), mode,
                             C.byref(ov))
    assert rc == 0, rc
    return out


def kf_match(cam, cfg, kf0, kf1, DT):
    out = S.TrackObsNP(kf0.B, kf0.cap_pt, kf0.cap_ln)
    DT = np.ascontiguousarray(DT, np.float64)
    v0, v1, ov = kf0.view(), kf1.view(), out.view()
    rc = lib().orc_kf_match(C.byref(cam), C.byref(cfg), C.byref(v0), C.byref(v1), C.c_void_p(S.ptr(DT)), C.byref(ov))
    assert rc == 0, rc
    return out


def map_match(cam, cfg, lmap, kf1, Twf):
    out = S.MapMatchesNP(lmap.B, lmap.cap_pt, lmap.cap_ln)
    Twf = np.ascontiguousarray(Twf, np.float64)
    mv, v1, ov = lmap.view(), kf1.view(), out.view()
    rc = lib().orc_map_match(C.byref(cam), C.byref(cfg), C.byref(mv), C.byref(v1), C.c_void_p(S.ptr(Twf)), C.byref(ov))
    assert rc == 0, rc
    return out


def medoid_desc(desc, offsets):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    out = np.zeros(len(offsets) - 1, np.int32)
    rc = lib().orc_medoid_desc(C.c_void_p(S.ptr(desc)), C.c_void_p(S.ptr(offsets)), len(out), C.c_void_p(S.ptr(out)))
    assert rc == 0, rc
    return out


def linecut(cam, cfg, DT_inv, obs, prev_cov):
    """obs is updated in place (sP/eP of cut lines); returns the gfpl_line_cut records."""
    cut = np.zeros((obs.B, obs.cap_ln), S.line_cut_dt)
    DT_inv = np.ascontiguousarray(DT_inv, np.float64)
    ov = obs.view()
    rc = lib().orc_linecut_submodular(C.byref(cam), C.byref(cfg), C.c_void_p(S.ptr(DT_inv)), C.byref(ov),
                                      C.c_void_p(S.ptr(prev_cov)), prev_cov.shape[1], C.c_void_p(S.ptr(cut)))
    assert rc == 0, rc
    return cut


def linecut_descent(cam, cfg, DT_inv, obs, prev_cov):
    """C6: obs is updated in place (sP/eP of every line); returns the gfpl_line_cut_descent records."""
    out = np.zeros((obs.B, obs.cap_ln), S.line_cut_descent_dt)
    DT_inv = np.ascontiguousarray(DT_inv, np.float64)
    ov = obs.view()
    rc = lib().orc_linecut_descent(C.byref(cam), C.byref(cfg), C.c_void_p(S.ptr(DT_inv)), C.byref(ov),
                                   C.c_void_p(S.ptr(prev_cov)), prev_cov.shape[1], C.c_void_p(S.ptr(out)))
    assert rc == 0, rc
    return out


def pose_cov_on_line(cam, cfg, DT_inv, ln, cov, cut):
    """getPoseCovOnLine for one line record (numpy scalars of ln_obs_dt / line_cov_dt) -> 6x6"""
    out = np.zeros(36)
    DT_inv = np.ascontiguousarray(DT_inv, np.float64)
    ln = np.ascontiguousarray(ln).reshape(1); cov = np.ascontiguousarray(cov).reshape(1)
    cut = np.ascontiguousarray(cut, np.float64)
    lib().orc_pose_cov_on_line(C.byref(cam), C.byref(cfg), C.c_void_p(S.ptr(DT_inv)), C.c_void_p(S.ptr(ln)),
                               C.c_void_p(S.ptr(cov)), C.c_void_p(S.ptr(cut)), C.c_void_p(S.ptr(out)))
    return out.reshape(6, 6)


def descent_volume(M):
    M = np.ascontiguousarray(M, np.float64)
    return float(lib().orc_descent_volume(C.c_void_p(S.ptr(M))))


def pose_hg(cam, cfg, DT, obs):
    DT = np.ascontiguousarray(DT, np.float64)
    H = np.zeros((obs.B, 36)); g = np.zeros((obs.B, 6)); e = np.zeros(obs.B)
    ov = obs.view()
    rc = lib().orc_pose_hg(C.byref(cam), C.byref(cfg), C.c_void_p(S.ptr(DT)), C.byref(ov), C.c_void_p(S.ptr(H)),
                           C.c_void_p(S.ptr(g)), C.c_void_p(S.ptr(e)))
    assert rc == 0, rc
    return H, g, e


def pose_optimize(cam, cfg, obs, state, DT_ini=None, overload=S.POSE_SEED_PREV_DT):
    """obs inlier masks are updated in place."""
    out = np.zeros(obs.B, S.pose_result_dt)
    ov = obs.view()
    if DT_ini is not None:
        DT_ini = np.ascontiguousarray(DT_ini, np.float64)
    rc = lib().orc_pose_optimize(C.byref(cam), C.byref(cfg), C.byref(ov), C.c_void_p(S.ptr(state)),
                                 C.c_void_p(S.ptr(DT_ini)), overload, C.c_void_p(S.ptr(out)))
    assert rc == 0, rc
    return out


def track(cam, cfg, prev, state, raw, flags, n_threads=1):
    curr = S.StereoFramesNP(raw.B, raw.cap_pt, raw.cap_ln)
    obs = S.TrackObsNP(raw.B, cfg.max_point_match_num, cfg.max_line_match_num)
    cut = np.zeros((raw.B, obs.cap_ln), S.line_cut_dt)
    pose = np.zeros(raw.B, S.pose_result_dt)
    pv, rv, cv, ov = prev.view(), raw.view(), curr.view(), obs.view()
    rc = lib().orc_track(C.byref(cam), C.byref(cfg), C.byref(pv), C.c_void_p(S.ptr(state)), C.byref(rv), flags,
                         C.byref(cv), C.byref(ov), C.c_void_p(S.ptr(cut)), C.c_void_p(S.ptr(pose)), n_threads)
    assert rc == 0, rc
    return curr, obs, cut, pose


def lba_assemble(cam, cfg, prob_arrays):
    """prob_arrays: dict of numpy arrays named like gfpl_lba_problem's fields."""
    p, keep = S.lba_problem_from_arrays(prob_arrays)
    Nkf, Npt, Nls = p.Nkf, p.Npt, p.Nls
    out = {"Hkk": np.zeros((Nkf, 36)), "Hpp": np.zeros((Npt, 9)), "Hll": np.zeros((Nls, 36)),
           "Hpk": np.zeros((p.n_pt_obs, 18)), "Hlk": np.zeros((p.n_ls_obs, 36)),
           "g": np.zeros(6 * Nkf + 3 * Npt + 6 * Nls), "err": np.zeros(1)}
    blk = S.LbaBlocks(*[S.ptr(out[k]) for k in ("Hkk", "Hpp", "Hll", "Hpk", "Hlk", "g", "err")])
    rc = lib().orc_lba_assemble(C.byref(cam), C.byref(cfg), C.byref(p), C.byref(blk))
    assert rc == 0, rc
    return out


def lba_step(cam, cfg, prob_arrays, lam, scale_by_hmax):
    p, keep = S.lba_problem_from_arrays(prob_arrays)
    N = 6 * p.Nkf + 3 * p.Npt + 6 * p.Nls
    DX = np.zeros(N); err = np.zeros(1); lam_eff = np.zeros(1)
    rc = lib().orc_lba_step(C.byref(cam), C.byref(cfg), C.byref(p), C.c_double(lam), int(scale_by_hmax),
                            C.c_void_p(S.ptr(DX)), C.c_void_p(S.ptr(err)), C.c_void_p(S.ptr(lam_eff)))
    assert rc == 0, rc
    return DX, float(err[0]), float(lam_eff[0])


def expmap_se3(x):
    x = np.ascontiguousarray(x, np.float64); T = np.zeros(16)
    lib().orc_expmap_se3(C.c_void_p(S.ptr(x)), C.c_void_p(S.ptr(T)))
    return T.reshape(4, 4)


def logmap_se3(T):
    T = np.ascontiguousarray(T, np.float64).reshape(16); x = np.zeros(6)
    lib().orc_logmap_se3(C.c_void_p(S.ptr(T)), C.c_void_p(S.ptr(x)))
    return x


def inverse_se3(T):
    T = np.ascontiguousarray(T, np.float64).reshape(16); Ti = np.zeros(16)
    lib().orc_inverse_se3(C.c_void_p(S.ptr(T)), C.c_void_p(S.ptr(Ti)))
    return Ti.reshape(4, 4)


def logdet6(M):
    M = np.ascontiguousarray(M, np.float64).reshape(36)
    return lib().orc_logdet6(C.c_void_p(S.ptr(M)))


def min_eig6(M):
    M = np.ascontiguousarray(M, np.float64).reshape(36)
    return lib().orc_min_eig6(C.c_void_p(S.ptr(M)))


def sym_eig(M):
    M = np.ascontiguousarray(M, np.float64)
    n = M.shape[0]
    ev = np.zeros(n)
    (lib().orc_sym_eig6 if n == 6 else lib().orc_sym_eig3)(C.c_void_p(S.ptr(M.reshape(-1))), C.c_void_p(S.ptr(ev)))
    return ev
