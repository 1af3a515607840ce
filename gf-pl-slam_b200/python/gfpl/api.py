"""ctypes binding of libgfpl.so (the C ABI of include/gfpl_c.h).

Every call goes through the C ABI with host buffers, exactly as the C++ facades do.  Loading fails
loudly when the CUDA library has not been built, and Context() raises when no CUDA device exists:
there is no CPU path in this package.
"""
import ctypes as C
import weakref
import os

import numpy as np

from . import _structs as S

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# GFPL_LIB: another build of the same library (kernel experiments, tools/stage_ms.py); there is no other implementation
LIB_PATH = os.environ.get("GFPL_LIB") or os.path.join(_ROOT, "lib", "libgfpl.so")
_LIB = None


class GfplError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gfpl error {code}: {msg}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.gfpl_last_error.restype = C.c_char_p
        L.gfpl_last_error.argtypes = [C.c_void_p]
        L.gfpl_ctx_launch_count.restype = C.c_int64
        L.gfpl_ctx_launch_count.argtypes = [C.c_void_p]
        L.gfpl_ctx_destroy.argtypes = [C.c_void_p]
        L.gfpl_ctx_stream.restype = C.c_void_p
        L.gfpl_ctx_stream.argtypes = [C.c_void_p]
        if hasattr(L, "gfpl_replay_create"):
            L.gfpl_replay_h2d_bytes.restype = C.c_int64
            L.gfpl_replay_d2h_bytes.restype = C.c_int64
            L.gfpl_replay_h2d_bytes.argtypes = [C.c_void_p]
            L.gfpl_replay_d2h_bytes.argtypes = [C.c_void_p]
            L.gfpl_replay_pose_device_ptr.restype = C.c_void_p
            L.gfpl_replay_pose_device_ptr.argtypes = [C.c_void_p]
            L.gfpl_replay_destroy.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def default_config():
    cfg = S.Config()
    lib().gfpl_config_default(C.byref(cfg))
    return cfg


def euroc_camera():
    cam = S.Camera()
    lib().gfpl_camera_euroc(C.byref(cam))
    return cam


def kitti_camera():
    cam = S.Camera()
    lib().gfpl_camera_kitti(C.byref(cam))
    return cam


def _vp(a):
    return C.c_void_p(S.ptr(a))


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim >= 2 and a.shape[-1] == 32
    return a


class Context:
    """gfpl_ctx: one CUDA device, one stream, camera + config."""

    def __init__(self, device=0, cam=None, cfg=None):
        self._h = C.c_void_p()
        self.cam = cam if cam is not None else euroc_camera()
        self.cfg = cfg if cfg is not None else default_config()
        rc = lib().gfpl_ctx_create(device, C.byref(self.cam), C.byref(self.cfg), C.byref(self._h))
        if rc != 0:
            raise GfplError(rc, lib().gfpl_last_error(None).decode())

    def close(self):
        # replays hold a pointer to their context: they go first (at interpreter shutdown the order of __del__ is arbitrary)
        for r in list(getattr(self, "_replays", ())):
            r.close()
        if self._h:
            lib().gfpl_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, ok=(0,)):
        if rc not in ok:
            raise GfplError(rc, lib().gfpl_last_error(self._h).decode())
        return rc

    def set_config(self, cfg):
        self.cfg = cfg
        self._check(lib().gfpl_ctx_set_config(self._h, C.byref(cfg)))

    def set_camera(self, cam):
        self.cam = cam
        self._check(lib().gfpl_ctx_set_camera(self._h, C.byref(cam)))

    def sync(self):
        self._check(lib().gfpl_ctx_sync(self._h))

    def stream(self):
        """cudaStream_t of the context (an integer handle, e.g. for torch.cuda.ExternalStream)"""
        return lib().gfpl_ctx_stream(self._h)

    def launch_count(self):
        return int(lib().gfpl_ctx_launch_count(self._h))

    def timer_record(self, slot):
        self._check(lib().gfpl_timer_record(self._h, slot))

    def timer_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._check(lib().gfpl_timer_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def measure_peaks(self):
        p, f = C.c_double(), C.c_double()
        self._check(lib().gfpl_measure_peaks(self._h, C.byref(p), C.byref(f)))
        return p.value, f.value

    def measure_int8_peak(self):
        p = C.c_double()
        self._check(lib().gfpl_measure_int8_peak(self._h, C.byref(p)))
        return p.value

    def measure_alu_peak(self):
        p = C.c_double()
        self._check(lib().gfpl_measure_alu_peak(self._h, C.byref(p)))
        return p.value

    # ---- M1-M3 ----
    def knn2(self, q, t, norm=S.NORM_HAMMING):
        q, t = _u8(q), _u8(t)
        idx = np.empty((len(q), 2), np.int32)
        dist = np.empty((len(q), 2), np.float32)
        self._check(lib().gfpl_knn2_hamming256(self._h, _vp(q), len(q), _vp(t), len(t), norm, _vp(idx), _vp(dist)))
        return idx, dist

    def knn2_batch(self, q, nq, t, nt, norm=S.NORM_HAMMING):
        q, t = _u8(q), _u8(t)
        B, cap_q, cap_t = q.shape[0], q.shape[1], t.shape[1]
        nq = np.ascontiguousarray(nq, np.int32)
        nt = np.ascontiguousarray(nt, np.int32)
        idx = np.empty((B, cap_q, 2), np.int32)
        dist = np.empty((B, cap_q, 2), np.float32)
        self._check(lib().gfpl_knn2_hamming256_batch(self._h, B, _vp(q), _vp(nq), cap_q, _vp(t), _vp(nt), cap_t, norm,
                                                     _vp(idx), _vp(dist)))
        return idx, dist

    def knn2_mutual(self, q, nq, t, nt, norm=S.NORM_HAMMING):
        """q: [B][cap_q][32], t: [B][cap_t][32] -> forward (idx, dist) [B][cap_q][2] and reverse best (idx, dist) [B][cap_t]"""
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        nq = np.ascontiguousarray(nq, np.int32); nt = np.ascontiguousarray(nt, np.int32)
        B, cap_q, cap_t = q.shape[0], q.shape[1], t.shape[1]
        idx = np.empty((B, cap_q, 2), np.int32); dist = np.empty((B, cap_q, 2), np.float32)
        bi = np.empty((B, cap_t), np.int32); bd = np.empty((B, cap_t), np.float32)
        self._check(lib().gfpl_knn2_mutual_batch(self._h, B, _vp(q), _vp(nq), cap_q, _vp(t), _vp(nt), cap_t, norm, _vp(idx),
                                                 _vp(dist), _vp(bi), _vp(bd)))
        return idx, dist, bi, bd

    def radius(self, q, t, r, norm=S.NORM_HAMMING, cap=None):
        q, t = _u8(q), _u8(t)
        cap = len(q) * len(t) if cap is None else cap
        off = np.empty(len(q) + 1, np.int32)
        idx = np.empty(max(cap, 1), np.int32)
        dist = np.empty(max(cap, 1), np.float32)
        rc = self._check(lib().gfpl_radius_hamming256(self._h, _vp(q), len(q), _vp(t), len(t), norm, C.c_float(r),
                                                      _vp(off), _vp(idx), _vp(dist), cap), ok=(0, S.E_CAPACITY))
        if rc == 0:
            return rc, off, idx[:off[-1]], dist[:off[-1]]
        return rc, off, idx, dist

    # ---- M4-M6 ----
    def stereo_match(self, raw, variant=S.STEREO_BRUTE):
        out = S.StereoFramesNP(raw.B, raw.cap_pt, raw.cap_ln)
        rv, ov = raw.view(), out.view()
        self._check(lib().gfpl_stereo_match(self._h, C.byref(rv), variant, C.byref(ov)))
        return out

    def stereo_match_orbslam(self, raw, pyr):
        """extractStereoFeatures_ORBSLAM after detection: row-band point stereo + SAD refinement; lines as _Brute"""
        out = S.StereoFramesNP(raw.B, raw.cap_pt, raw.cap_ln)
        rv, pv, ov = raw.view(), pyr.view(), out.view()
        self._check(lib().gfpl_stereo_match_orbslam(self._h, C.byref(rv), C.byref(pv), C.byref(ov)))
        return out

    def stereo_uncertainty(self, frames):
        out = np.zeros((frames.B, frames.cap_ln), S.line_cov_dt)
        fv = frames.view()
        self._check(lib().gfpl_stereo_uncertainty(self._h, C.byref(fv), _vp(out)))
        return out

    # ---- M7/M8 ----
    def f2f_match(self, prev, curr, state, mode=S.F2F_HYBRID):
        out = S.TrackObsNP(prev.B, self.cfg.max_point_match_num, self.cfg.max_line_match_num)
        pv, cv, ov = prev.view(), curr.view(), out.view()
        self._check(lib().gfpl_f2f_match(self._h, C.byref(pv), C.byref(cv), _vp(state), mode, C.byref(ov)))
        return out

    def kf_match(self, kf0, kf1, DT):
        """KF<->KF matching of MapHandler::lookForCommonMatches (matching + gates)"""
        out = S.TrackObsNP(kf0.B, kf0.cap_pt, kf0.cap_ln)
        DT = np.ascontiguousarray(DT, np.float64)
        v0, v1, ov = kf0.view(), kf1.view(), out.view()
        self._check(lib().gfpl_kf_match(self._h, C.byref(v0), C.byref(v1), _vp(DT), C.byref(ov)))
        return out

    def map_match(self, lmap, kf1, Twf):
        """local map <-> KF matching of MapHandler::lookForCommonMatches (selection + matching + gates)"""
        out = S.MapMatchesNP(lmap.B, lmap.cap_pt, lmap.cap_ln)
        Twf = np.ascontiguousarray(Twf, np.float64)
        mv, v1, ov = lmap.view(), kf1.view(), out.view()
        self._check(lib().gfpl_map_match(self._h, C.byref(mv), C.byref(v1), _vp(Twf), C.byref(ov)))
        return out

    def medoid_desc(self, desc, offsets):
        """descriptor part of MapPoint/MapLine::updateAverageDescDir for L landmarks (CSR lists of 32-byte rows)"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        offsets = np.ascontiguousarray(offsets, np.int32)
        out = np.zeros(len(offsets) - 1, np.int32)
        self._check(lib().gfpl_medoid_desc(self._h, _vp(desc), _vp(offsets), len(out), _vp(out)))
        return out

    # ---- C1-C5 ----
    def linecut(self, DT_inv, obs, prev_cov):
        """obs is updated in place (sP/eP of cut lines); returns gfpl_line_cut records."""
        cut = np.zeros((obs.B, obs.cap_ln), S.line_cut_dt)
        DT_inv = np.ascontiguousarray(DT_inv, np.float64)
        ov = obs.view()
        self._check(lib().gfpl_linecut_submodular(self._h, _vp(DT_inv), C.byref(ov), _vp(prev_cov),
                                                  prev_cov.shape[1], _vp(cut)))
        return cut

    def linecut_descent(self, DT_inv, obs, prev_cov):
        """C6, estimateProjUncertainty_descent: obs is updated in place (sP/eP of every line); returns gfpl_line_cut_descent records."""
        out = np.zeros((obs.B, obs.cap_ln), S.line_cut_descent_dt)
        DT_inv = np.ascontiguousarray(DT_inv, np.float64)
        ov = obs.view()
        self._check(lib().gfpl_linecut_descent(self._h, _vp(DT_inv), C.byref(ov), _vp(prev_cov), prev_cov.shape[1], _vp(out)))
        return out

    # ---- P1-P4 ----
    def vector_stdv_mad(self, res, n):
        """res: [B][cap] float64, n: [B] -> 1.4826 * MAD per row (vector_stdv_mad, auxiliar.cpp:521-537)"""
        res = np.ascontiguousarray(res, np.float64)
        n = np.ascontiguousarray(n, np.int32)
        out = np.zeros(len(n), np.float64)
        self._check(lib().gfpl_vector_stdv_mad(self._h, _vp(res), _vp(n), len(n), res.shape[1], _vp(out)))
        return out

    def pose_hg(self, DT, obs):
        DT = np.ascontiguousarray(DT, np.float64)
        H = np.zeros((obs.B, 36)); g = np.zeros((obs.B, 6)); e = np.zeros(obs.B)
        ov = obs.view()
        self._check(lib().gfpl_pose_hg(self._h, _vp(DT), C.byref(ov), _vp(H), _vp(g), _vp(e)))
        return H, g, e

    def pose_optimize(self, obs, state, DT_ini=None, overload=S.POSE_SEED_PREV_DT):
        """obs inlier masks are updated in place."""
        out = np.zeros(obs.B, S.pose_result_dt)
        ov = obs.view()
        if DT_ini is not None:
            DT_ini = np.ascontiguousarray(DT_ini, np.float64)
        self._check(lib().gfpl_pose_optimize(self._h, C.byref(ov), _vp(state), _vp(DT_ini), overload, _vp(out)))
        return out

    # ---- whole step ----
    def track(self, prev, state, raw, flags):
        curr = S.StereoFramesNP(raw.B, raw.cap_pt, raw.cap_ln)
        obs = S.TrackObsNP(raw.B, self.cfg.max_point_match_num, self.cfg.max_line_match_num)
        cut = np.zeros((raw.B, obs.cap_ln), S.line_cut_dt)
        pose = np.zeros(raw.B, S.pose_result_dt)
        pv, rv, cv, ov = prev.view(), raw.view(), curr.view(), obs.view()
        self._check(lib().gfpl_track(self._h, C.byref(pv), _vp(state), C.byref(rv), flags, C.byref(cv), C.byref(ov),
                                     _vp(cut), _vp(pose)))
        return curr, obs, cut, pose

    # ---- L2 ----
    def lba_assemble(self, prob_arrays):
        p, keep = S.lba_problem_from_arrays(prob_arrays)
        Nkf, Npt, Nls = p.Nkf, p.Npt, p.Nls
        out = {"Hkk": np.zeros((Nkf, 36)), "Hpp": np.zeros((Npt, 9)), "Hll": np.zeros((Nls, 36)),
               "Hpk": np.zeros((p.n_pt_obs, 18)), "Hlk": np.zeros((p.n_ls_obs, 36)),
               "g": np.zeros(6 * Nkf + 3 * Npt + 6 * Nls), "err": np.zeros(1)}
        blk = S.LbaBlocks(*[S.ptr(out[k]) for k in ("Hkk", "Hpp", "Hll", "Hpk", "Hlk", "g", "err")])
        self._check(lib().gfpl_lba_assemble(self._h, C.byref(p), C.byref(blk)))
        return out

    def lba_step(self, prob_arrays, lam, scale_by_hmax):
        """assembly + damped Schur-complement solve of one LM pass: (DX, err_sum, lambda_eff)"""
        p, keep = S.lba_problem_from_arrays(prob_arrays)
        N = 6 * p.Nkf + 3 * p.Npt + 6 * p.Nls
        DX = np.zeros(N); err = np.zeros(1); lam_eff = np.zeros(1)
        self._check(lib().gfpl_lba_step(self._h, C.byref(p), C.c_double(lam), int(scale_by_hmax), _vp(DX), _vp(err),
                                        _vp(lam_eff)))
        return DX, float(err[0]), float(lam_eff[0])


    def lba_batch_prepare(self, windows):
        """host-side views of W windows (the caller's own arrays, no copies): reusable across lba_step_batch_prepared calls"""
        W = len(windows)
        built = [S.lba_problem_from_arrays(w) for w in windows]
        probs = (S.LbaProblem * W)(*[b[0] for b in built])
        sizes = [6 * p.Nkf + 3 * p.Npt + 6 * p.Nls for p, _ in built]
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        return {"W": W, "probs": probs, "keep": built, "off": off, "DX": np.zeros(int(off[-1])), "err": np.zeros(W), "lam": np.zeros(W)}

    def lba_step_batch_prepared(self, h, lam, scale_by_hmax):
        self._check(lib().gfpl_lba_step_batch(self._h, h["W"], h["probs"], C.c_double(lam), int(scale_by_hmax), _vp(h["DX"]),
                                              _vp(h["err"]), _vp(h["lam"])))
        return h["DX"], h["err"], h["lam"]

    def lba_step_batch(self, windows, lam, scale_by_hmax):
        """one LM pass (assembly + damped Schur solve) of W independent windows in one call:
        ([DX_w], err_sum [W], lambda_eff [W])"""
        h = self.lba_batch_prepare(windows)
        DX, err, lam_eff = self.lba_step_batch_prepared(h, lam, scale_by_hmax)
        return [DX[h["off"][w]:h["off"][w + 1]] for w in range(h["W"])], err, lam_eff


class Replay:
    """gfpl_replay: device-resident batch of B frame pairs (upload once / run many / download)."""

    def __init__(self, ctx, B, cap_pt, cap_ln, prev_cap_pt=None, prev_cap_ln=None):
        self.ctx, self.B, self.cap_pt, self.cap_ln = ctx, B, cap_pt, cap_ln
        self._h = C.c_void_p()
        if not hasattr(ctx, "_replays"):
            ctx._replays = weakref.WeakSet()
        ctx._replays.add(self)
        ctx._check(lib().gfpl_replay_create_ex(ctx._h, B, cap_pt, cap_ln, prev_cap_pt or cap_pt, prev_cap_ln or cap_ln,
                                               C.byref(self._h)))

    def close(self):
        if self._h:
            lib().gfpl_replay_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, prev, state, raw):
        pv, rv = prev.view(), raw.view()
        self.ctx._check(lib().gfpl_replay_upload(self._h, C.byref(pv), _vp(state), C.byref(rv)))

    def upload_slim(self, prev_slim, state, raw):
        pv, rv = prev_slim.view(), raw.view()
        self.ctx._check(lib().gfpl_replay_upload_slim(self._h, C.byref(pv), _vp(state), C.byref(rv)))

    def upload_compact(self, compact, state):
        cv = compact.view()
        self.ctx._check(lib().gfpl_replay_upload_compact(self._h, C.byref(cv), _vp(state)))

    def upload_curr_compact(self, compact):
        """sequence-chained replay: the next current raw frames only"""
        cv = compact.view()
        self.ctx._check(lib().gfpl_replay_upload_curr_compact(self._h, C.byref(cv)))

    def advance(self, time_stamp_next=None):
        """updateFrame: prev <- curr on the device, pose records -> frame state"""
        ts = np.ascontiguousarray(time_stamp_next, np.float64) if time_stamp_next is not None else None
        self.ctx._check(lib().gfpl_replay_advance(self._h, _vp(ts)))

    def run(self, flags):
        self.ctx._check(lib().gfpl_replay_run(self._h, flags))

    def set_pipeline(self, n_sub):
        """software pipeline over n_sub sub-batches: matching of sub-batch k+1 overlaps the fp64 stages of sub-batch k"""
        self.ctx._check(lib().gfpl_replay_set_pipeline(self._h, int(n_sub)))

    def download(self, pose, obs=None, cut=None, curr=None):
        ov = obs.view() if obs is not None else None
        cv = curr.view() if curr is not None else None
        self.ctx._check(lib().gfpl_replay_download(self._h, _vp(pose), C.byref(ov) if ov is not None else None,
                                                   _vp(cut), C.byref(cv) if cv is not None else None))

    def h2d_bytes(self):
        return int(lib().gfpl_replay_h2d_bytes(self._h))

    def d2h_bytes(self):
        return int(lib().gfpl_replay_d2h_bytes(self._h))

    def pose_device_ptr(self):
        return lib().gfpl_replay_pose_device_ptr(self._h)

    def pair_counts(self):
        n = (C.c_int64 * 4)()
        self.ctx._check(lib().gfpl_replay_pair_counts(self._h, n))
        return list(n)

    def stage_ms(self):
        ms = (C.c_float * S.NUM_STAGES)()
        self.ctx._check(lib().gfpl_replay_stage_ms(self._h, ms))
        return list(ms)
