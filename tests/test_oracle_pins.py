"""Pins of the fp64 oracle (the reference ships no fixture for these stages and cannot be compiled
here).  Every fp64 row of §8(a) is compared, at 1e-12 relative (discrete decisions: equal), with tests/ref_numpy.py — a
restatement written independently from the reference text with LAPACK in place of Eigen.  CPU only.
"""
import ctypes as C

import numpy as np
import pytest

import ref_numpy as R
from gfpl import _structs as S
from gfpl import synth

RT = 1e-12


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if b.size else 0.0


@pytest.fixture(scope="module")
def world(orc):
    d = synth.make_pairs(4, seed=77, ragged=True)
    cfg = orc.default_config()
    prev = orc.stereo_match(d["cam"], cfg, d["prev_raw"])
    curr = orc.stereo_match(d["cam"], cfg, d["curr_raw"])
    obs = orc.f2f_match(d["cam"], cfg, prev, curr, d["state"])
    cov = orc.stereo_uncertainty(d["cam"], cfg, prev)
    Tc = d["state"]["Tfw"].reshape(-1, 4, 4) @ d["state"]["DT"].reshape(-1, 4, 4)
    DT_inv = np.linalg.inv(Tc) @ d["state"]["Tfw"].reshape(-1, 4, 4)
    return dict(d=d, cam=d["cam"], cfg=cfg, prev=prev, curr=curr, obs=obs, cov=cov, DT_inv=DT_inv)


# ---- M4 / M5: stereo accept loops and the features they emit (stereoFrame.cpp:795-1014) ---------------------------------
def test_stereo_points_pin(orc, world):
    cam, cfg, raw, fr = world["cam"], world["cfg"], world["d"]["curr_raw"], world["curr"]
    for p in range(raw.B):
        nl, nr = raw.n_pt_l[p], raw.n_pt_r[p]
        ref = R.stereo_points_brute(cam, cfg, raw.pt_l[p, :nl], raw.pt_r[p, :nr], raw.pdesc_l[p, :nl], raw.pdesc_r[p, :nr])
        assert fr.n_pt[p] == len(ref) > 100
        f = fr.pt[p, :len(ref)]
        assert np.array_equal(f["src_l"], [r["src_l"] for r in ref]) and np.array_equal(f["src_r"], [r["src_r"] for r in ref])
        assert np.array_equal(f["level"], [r["level"] for r in ref])
        for name in ("P", "pl", "disp", "sigma2"):
            assert rel(f[name], np.array([r[name] for r in ref])) <= RT, name


def test_stereo_lines_pin(orc, world):
    """mutual check, descriptor-gap MAD threshold, end-point disparities from the right line equation, overlap,
    horizontal-line gate, end-point covariance eigenvalue gate (SelfAdjointEigenSolver -> LAPACK eigvalsh)"""
    cam, cfg, raw, fr = world["cam"], world["cfg"], world["d"]["curr_raw"], world["curr"]
    total = 0
    for p in range(raw.B):
        nl, nr = raw.n_ln_l[p], raw.n_ln_r[p]
        ref = R.stereo_lines_brute(cam, cfg, raw.ln_l[p, :nl], raw.ln_r[p, :nr], raw.ldesc_l[p, :nl], raw.ldesc_r[p, :nr])
        assert fr.n_ln[p] == len(ref)
        f = fr.ln[p, :len(ref)]
        assert np.array_equal(f["src_l"], [r["src_l"] for r in ref]) and np.array_equal(f["src_r"], [r["src_r"] for r in ref])
        for name in ("sP", "eP", "spl", "epl", "sdisp", "edisp", "le", "angle"):
            assert rel(f[name], np.array([r[name] for r in ref])) <= RT, name
        assert (f["sigma2"] == 1.0).all() and (f["idx"] == -1).all()      # lsdScale = 1; _Brute passes idx = -1
        total += len(ref)
    assert total > 30


def test_stereo_line_gates_pin_with_thresholds_inside_the_data(orc, world):
    """the same comparison with every gate's threshold moved INTO the data range, so each gate rejects a share"""
    cam, raw = world["cam"], world["d"]["curr_raw"]
    for field, val in (("line_cov_th", 0.02), ("stereo_overlap_th", 0.97), ("line_horiz_th", 0.6), ("min_disp", 6.0),
                       ("desc_th_l", 8.0)):
        cfg = orc.default_config()
        setattr(cfg, field, val)
        fr = orc.stereo_match(cam, cfg, raw)
        base = world["curr"]
        changed = 0
        for p in range(raw.B):
            nl, nr = raw.n_ln_l[p], raw.n_ln_r[p]
            ref = R.stereo_lines_brute(cam, cfg, raw.ln_l[p, :nl], raw.ln_r[p, :nr], raw.ldesc_l[p, :nl], raw.ldesc_r[p, :nr])
            assert fr.n_ln[p] == len(ref), field
            assert np.array_equal(fr.ln[p, :len(ref)]["src_l"], [r["src_l"] for r in ref]), field
            changed += int(base.n_ln[p] != len(ref))
        assert changed > 0, f"{field}={val} rejects nothing: the gate is not exercised"


# ---- M7: crossFrameMatching_Hybrid (stereoFrameHandler.cpp:451-686) --------------------------------------------------------
def test_f2f_hybrid_pin(orc, world):
    """radius candidates + fp64 projection gate + per-current-point arg-min (the multimap resolver) for points; mutual
    check, budget and descriptor-gap MAD thresholds for lines; the observation records both emit"""
    cam, cfg, prev, curr, obs, d = world["cam"], world["cfg"], world["prev"], world["curr"], world["obs"], world["d"]
    for p in range(obs.B):
        npp, ncp, npl, ncl = prev.n_pt[p], curr.n_pt[p], prev.n_ln[p], curr.n_ln[p]
        pts, lns = R.f2f_hybrid(cam, cfg, prev.pt[p, :npp], prev.pdesc[p, :npp], curr.pt[p, :ncp], curr.pdesc[p, :ncp],
                                prev.ln[p, :npl], prev.ldesc[p, :npl], curr.ln[p, :ncl], curr.ldesc[p, :ncl],
                                d["state"]["Tfw"][p].reshape(4, 4), d["state"]["DT"][p].reshape(4, 4))
        assert obs.n_pt[p] == len(pts) > 50 and obs.n_ln[p] == len(lns)
        o = obs.pt[p, :len(pts)]
        assert np.array_equal(o["prev_idx"], [a for a, _ in pts]) and np.array_equal(o["curr_idx"], [b for _, b in pts])
        assert np.array_equal(o["P"], prev.pt[p]["P"][o["prev_idx"]]) and np.array_equal(o["pl_obs"], curr.pt[p]["pl"][o["curr_idx"]])
        assert np.array_equal(o["sigma2"], prev.pt[p]["sigma2"][o["prev_idx"]])
        ol = obs.ln[p, :len(lns)]
        assert np.array_equal(ol["prev_idx"], [a for a, _ in lns]) and np.array_equal(ol["curr_idx"], [b for _, b in lns])
        assert np.array_equal(ol["sP"], prev.ln[p]["sP"][ol["prev_idx"]]) and np.array_equal(ol["le_obs"], curr.ln[p]["le"][ol["curr_idx"]])


def test_f2f_hybrid_pin_tight_gates(orc, world):
    """a 2-pixel projection window, a radius inside the inlier distances and a 40-match budget: every gate binds"""
    cam, prev, curr, d = world["cam"], world["prev"], world["curr"], world["d"]
    cfg = orc.default_config()
    cfg.f2f_point_window = 2.0; cfg.point_match_radius = 24.0; cfg.max_point_match_num = 40; cfg.max_line_match_num = 5
    obs = orc.f2f_match(cam, cfg, prev, curr, d["state"])
    for p in range(obs.B):
        npp, ncp, npl, ncl = prev.n_pt[p], curr.n_pt[p], prev.n_ln[p], curr.n_ln[p]
        pts, lns = R.f2f_hybrid(cam, cfg, prev.pt[p, :npp], prev.pdesc[p, :npp], curr.pt[p, :ncp], curr.pdesc[p, :ncp],
                                prev.ln[p, :npl], prev.ldesc[p, :npl], curr.ln[p, :ncl], curr.ldesc[p, :ncl],
                                d["state"]["Tfw"][p].reshape(4, 4), d["state"]["DT"][p].reshape(4, 4))
        assert obs.n_pt[p] == len(pts) and obs.n_ln[p] == len(lns) and len(pts) <= 40 and len(lns) <= 5
        assert np.array_equal(obs.pt[p, :len(pts)]["prev_idx"], [a for a, _ in pts])
        assert np.array_equal(obs.pt[p, :len(pts)]["curr_idx"], [b for _, b in pts])
        assert np.array_equal(obs.ln[p, :len(lns)]["prev_idx"], [a for a, _ in lns])
    assert (obs.n_pt > 0).all() and (obs.n_pt < world["obs"].n_pt).all()


# ---- C1: estimateStereoUncertainty (stereoFrame.cpp:1434-1484) -------------------------------------------------------------
def test_stereo_uncertainty_pin(orc, world):
    cam, cfg, prev, cov = world["cam"], world["cfg"], world["prev"], world["cov"]
    n = 0
    for p in range(prev.B):
        for l in range(prev.n_ln[p]):
            cs, ce = R.stereo_uncertainty(cam, cfg, prev.ln[p, l])
            assert rel(cov[p, l]["covS"].reshape(3, 3), cs) <= RT and rel(cov[p, l]["covE"].reshape(3, 3), ce) <= RT
            n += 1
    assert n > 50
    le0 = np.abs(np.concatenate([prev.ln[p, :prev.n_ln[p]]["le"][:, 0] for p in range(prev.B)]))
    assert (le0 > 0.15).any()                                            # (both branches of the ratio: see the next test)


def test_stereo_uncertainty_horizontal_branch(orc, world):
    cam, cfg = world["cam"], world["cfg"]
    fr = world["prev"].copy()
    fr.ln[0, :fr.n_ln[0]]["le"][:, 0] *= 0.05                            # |le[0]| <= 0.15 -> ratioDispSTDHor
    cov = orc.stereo_uncertainty(cam, cfg, fr)
    for l in range(fr.n_ln[0]):
        assert abs(fr.ln[0, l]["le"][0]) <= 0.15
        cs, ce = R.stereo_uncertainty(cam, cfg, fr.ln[0, l])
        assert rel(cov[0, l]["covS"].reshape(3, 3), cs) <= RT and rel(cov[0, l]["covE"].reshape(3, 3), ce) <= RT


# ---- C2 / C3: getPoseInfoOnLine, getPoseInfoPoint (stereoFrameHandler.cpp:1342-1447) ------------------------------------------
def _orc_info_line(orc, cam, cfg, DT_inv, ln, cov, cut):
    info = np.zeros(36)
    cut = np.ascontiguousarray(cut, np.float64)
    orc.lib().orc_pose_info_on_line(C.byref(cam), C.byref(cfg), C.c_void_p(np.ascontiguousarray(DT_inv).ctypes.data),
                                    C.c_void_p(ln.ctypes.data), C.c_void_p(cov.ctypes.data), C.c_void_p(cut.ctypes.data),
                                    C.c_void_p(info.ctypes.data))
    return info.reshape(6, 6)


def test_pose_info_on_line_pin(orc, world):
    cam, cfg, obs, cov = world["cam"], world["cfg"], world["obs"], world["cov"]
    rng = np.random.default_rng(3)
    n = 0
    for p in range(obs.B):
        for l in range(obs.n_ln[p]):
            ln = obs.ln[p, l:l + 1].copy()
            cv = cov[p, ln[0]["prev_idx"]:ln[0]["prev_idx"] + 1].copy()
            for cut in ((0.0, 0.0), (0.05, 0.0), tuple(np.round(rng.uniform(0, 0.5, 2) / 0.05) * 0.05), (1.0, 0.0)):
                a = _orc_info_line(orc, cam, cfg, world["DT_inv"][p], ln, cv, cut)
                b = R.pose_info_on_line(cam, cfg, world["DT_inv"][p], ln[0]["sP"], ln[0]["eP"], ln[0]["le_obs"],
                                        cv[0]["covS"].reshape(3, 3), cv[0]["covE"].reshape(3, 3), cut)
                assert rel(a, b) <= RT, (p, l, cut)
                n += 1
    assert n > 100


def test_pose_info_point_pin(orc, world):
    cam, cfg, obs = world["cam"], world["cfg"], world["obs"]
    for p in range(obs.B):
        for k in range(0, obs.n_pt[p], 7):
            info = np.zeros(36)
            pt = obs.pt[p, k:k + 1].copy()
            orc.lib().orc_pose_info_point(C.byref(cam), C.byref(cfg), C.c_void_p(np.ascontiguousarray(world["DT_inv"][p]).ctypes.data),
                                          C.c_void_p(pt.ctypes.data), C.c_void_p(info.ctypes.data))
            assert rel(info.reshape(6, 6), R.pose_info_point(cam, cfg, world["DT_inv"][p], pt[0]["P"], pt[0]["pl_obs"])) <= RT


# ---- C4 / C5: the whole submodular search (stereoFrameHandler.cpp:1618-1764, logdet via LAPACK Cholesky) ---------------------
def test_linecut_submodular_search_pin(orc, world):
    cam, cfg, cov = world["cam"], world["cfg"], world["cov"]
    obs = world["obs"].copy()
    before = world["obs"].copy()
    cut = orc.linecut(cam, cfg, world["DT_inv"].reshape(-1, 16), obs, cov)
    moved = lines = 0
    for p in range(obs.B):
        n, npt = obs.n_ln[p], obs.n_pt[p]
        lns = before.ln[p, :n]
        covs = [(cov[p, i]["covS"].reshape(3, 3), cov[p, i]["covE"].reshape(3, 3)) for i in lns["prev_idx"]]
        rc, rinfo, rsP, reP = R.linecut_submodular(cam, cfg, world["DT_inv"][p], before.pt[p, :npt], lns, covs)
        assert np.array_equal(np.round(cut["cutRatio"][p, :n] / cfg.cut_step), np.round(rc / cfg.cut_step)), p   # decisions
        assert rel(cut["cutRatio"][p, :n], rc) <= RT
        assert rel(cut["invCovPose"][p, :n].reshape(n, 6, 6), rinfo) <= 1e-11
        assert rel(obs.ln[p, :n]["sP"], rsP) <= RT and rel(obs.ln[p, :n]["eP"], reP) <= RT
        m = np.abs(rc).max(1) > 1e-4
        for l in np.nonzero(m)[0]:                     # updateEndPointByRatio: re-projected end points and disparities
            if abs(rc[l, 0]) > 1e-4:
                assert rel(cut["spl"][p, l], R.projection(cam, rsP[l])) <= RT
                assert rel(cut["sdisp"][p, l], R.get_disparity(cam, rsP[l][2])) <= RT
            if abs(rc[l, 1]) > 1e-4:
                assert rel(cut["epl"][p, l], R.projection(cam, reP[l])) <= RT
                assert rel(cut["edisp"][p, l], R.get_disparity(cam, reP[l][2])) <= RT
        moved += int(m.sum()); lines += n
    assert lines > 20 and moved > 3


# ---- C6: estimateProjUncertainty_descent + getPoseCovOnLine (stereoFrameHandler.cpp:1476-1610, 1249-1338) -------------------
def test_pose_cov_on_line_pin(orc, world):
    """the rank-2 pose covariance and its volume: general eigen-solver (LAPACK dgeev, as Eigen::EigenSolver in getEigenValue)
    against the oracle's symmetric Jacobi solver; closed form of the volume: cov_sR * cov_eR / det(J^T J)"""
    cam, cfg, obs, cov = world["cam"], world["cfg"], world["obs"], world["cov"]
    n = 0
    for p in range(obs.B):
        for l in range(min(obs.n_ln[p], 6)):
            ln = obs.ln[p, l]
            c = cov[p, ln["prev_idx"]]
            for cut in ((0.0, 0.0), (0.05, 0.0), (0.25, 0.4), (0.0, 0.95)):
                M = orc.pose_cov_on_line(cam, cfg, world["DT_inv"][p], ln, c, cut)
                ref = R.pose_cov_on_line(cam, cfg, world["DT_inv"][p], ln["sP"], ln["eP"], ln["le_obs"], c["covS"].reshape(3, 3),
                                         c["covE"].reshape(3, 3), cut)
                # (J^T J)^-1 of two nearly parallel Jacobian columns: agreement is limited by cond(J^T J) * eps (adjugate
                # inverse in the oracle / Eigen, LAPACK LU here), not by the formulas
                assert rel(M, ref) <= 1e-9
                ev = R.eigen_values_desc(ref)
                assert abs(orc.descent_volume(M) - ev[0] * ev[1]) <= 1e-8 * abs(ev[0] * ev[1])
                n += 1
    assert n > 40


def test_linecut_descent_search_pin(orc, world):
    cam, cfg, cov = world["cam"], world["cfg"], world["cov"]
    obs = world["obs"].copy()
    before = world["obs"].copy()
    out = orc.linecut_descent(cam, cfg, world["DT_inv"].reshape(-1, 16), obs, cov)
    lines = moved = near_ties = 0
    for p in range(obs.B):
        n = obs.n_ln[p]
        lns = before.ln[p, :n]
        covs = [(cov[p, i]["covS"].reshape(3, 3), cov[p, i]["covE"].reshape(3, 3)) for i in lns["prev_idx"]]
        ref = R.linecut_descent(cam, cfg, world["DT_inv"][p], lns, covs)
        for l in range(n):
            r, o = ref[l], out[p, l]
            if r["min_margin"] < 1e-7:          # a comparison the two eigen-solvers may order differently: counted, not compared
                near_ties += 1
                continue
            assert np.array_equal(np.round(o["cutRatio"] / cfg.cut_step), np.round(r["cut"] / cfg.cut_step)), (p, l)   # decisions
            assert o["n_volume"] == len(r["vols"]) and abs(o["volume"] - r["vols"][-1]) <= 1e-8 * abs(r["vols"][-1])
            assert o["fst_idx"] == 0 and o["lst_idx"] == 3
            assert rel(o["covlPt2D"][0].reshape(3, 3), r["covl"][0]) <= 1e-11 and rel(o["covlPt2D"][1].reshape(3, 3), r["covl"][1]) <= 1e-11
            assert rel(o["lPt2D"], r["lPt2D"]) <= RT
            assert rel(obs.ln[p, l]["sP"], r["sP"]) <= RT and rel(obs.ln[p, l]["eP"], r["eP"]) <= RT
            assert rel(o["spl"], r["spl"]) <= RT and rel(o["epl"], r["epl"]) <= RT
            assert rel(o["sdisp"], r["sdisp"]) <= RT and rel(o["edisp"], r["edisp"]) <= RT
            lines += 1; moved += int(np.abs(r["cut"]).max() > 1e-4)
    assert lines > 20 and moved > 3 and near_ties <= lines // 10, (lines, moved, near_ties)


# ---- P1-P4: optimizePose = GN(5) + removeOutliers + GN(10) + gates (stereoFrameHandler.cpp:1939-2116) -------------------------
def _cmp_pose(pose, p, ref, tol=1e-9):
    # the iterates pass through a 6x6 solve per step (Eigen LDLT in the reference, LAPACK LU here, our own LDLT in the
    # oracle): agreement is limited by cond(H) * eps per iteration, not by the formulas
    for f in ("DT", "Tfw", "DT_cov", "Tfw_cov"):
        assert rel(pose[f][p], ref[f].reshape(-1)) <= tol, f
    assert rel(pose["err_norm"][p], ref["err_norm"]) <= tol
    assert rel(pose["DT_cov_eig"][p], ref["DT_cov_eig"]) <= 1e-7


@pytest.mark.parametrize("seed", [5, 6])
def test_optimize_pose_pin(orc, seed):
    d = synth.make_pairs(3, seed=seed, ragged=True)
    cam, cfg = d["cam"], orc.default_config()
    prev = orc.stereo_match(cam, cfg, d["prev_raw"])
    curr = orc.stereo_match(cam, cfg, d["curr_raw"])
    obs = orc.f2f_match(cam, cfg, prev, curr, d["state"])
    oo = obs.copy()
    pose = orc.pose_optimize(cam, cfg, oo, d["state"], d["state"]["DT"].copy(), S.POSE_SEED_GIVEN)
    for p in range(3):
        npt, nln = obs.n_pt[p], obs.n_ln[p]
        ref = R.optimize_pose(cam, cfg, d["state"]["DT"][p].reshape(4, 4), d["state"][p], obs.pt[p, :npt], obs.ln[p, :nln])
        assert pose["status"][p] == 0
        assert np.array_equal(oo.pt_inlier[p, :npt], ref["pin"]) and np.array_equal(oo.ln_inlier[p, :nln], ref["lin"])   # masks
        assert (ref["pin"] == 0).any()
        assert pose["n_inliers_pt"][p] == ref["pin"].sum() and pose["n_inliers_ls"][p] == ref["lin"].sum()
        assert pose["iters_first"][p] == ref["iters_first"] and pose["iters_ref"][p] == ref["iters_ref"]
        _cmp_pose(pose, p, ref)


def test_gauss_newton_single_stage_pin(orc, world):
    """one gaussNewtonOptimization from the seed (no outlier pass): H^-1, error and pose after max_iters = 1, 2, 5"""
    cam, obs, d = world["cam"], world["obs"], world["d"]
    for iters in (1, 2, 5):
        cfg = orc.default_config()
        cfg.max_iters = iters
        cfg.inlier_k = 1e9                                   # removeOutliers keeps everything
        cfg.max_iters_ref = iters
        oo = obs.copy()
        pose = orc.pose_optimize(cam, cfg, oo, d["state"], d["state"]["DT"].copy(), S.POSE_SEED_GIVEN)
        for p in range(obs.B):
            npt, nln = obs.n_pt[p], obs.n_ln[p]
            ones_p, ones_l = np.ones(npt, np.uint8), np.ones(nln, np.uint8)
            DT, cov, err, n = R.gauss_newton(cam, cfg, d["state"]["DT"][p].reshape(4, 4).copy(), obs.pt[p, :npt], ones_p,
                                             obs.ln[p, :nln], ones_l, iters)
            assert rel(pose["DT"][p], R.inverse_se3(DT).reshape(-1)) <= 1e-10
            assert rel(pose["DT_cov"][p], cov.reshape(-1)) <= 1e-9 and rel(pose["err_norm"][p], err) <= 1e-10
            assert pose["iters_ref"][p] == n


def test_optimize_pose_gates_pin(orc, world):
    """too few features -> identity; zero time step -> motion gate rejects the step (stereoFrameHandler.cpp:1953, 1992)"""
    cam, cfg, d = world["cam"], world["cfg"], world["d"]
    obs = world["obs"].copy()
    state = d["state"].copy()
    obs.n_pt[0] = 6; obs.n_ln[0] = 3
    state["time_stamp_curr"][1] = state["time_stamp"][1]
    pose = orc.pose_optimize(cam, cfg, obs, state, state["DT"].copy(), S.POSE_SEED_GIVEN)
    for p in (0, 1):
        ref = R.optimize_pose(cam, cfg, state["DT"][p].reshape(4, 4), state[p], obs.pt[p, :obs.n_pt[p]], obs.ln[p, :obs.n_ln[p]])
        assert pose["err_norm"][p] == -1.0 == ref["err_norm"] or p == 0
        assert rel(pose["DT"][p], ref["DT"].reshape(-1)) <= RT and rel(pose["Tfw"][p], ref["Tfw"].reshape(-1)) <= RT
        assert rel(pose["Tfw_cov"][p], ref["Tfw_cov"].reshape(-1)) <= 1e-9
        assert rel(pose["DT_cov"][p], ref["DT_cov"].reshape(-1)) <= 1e-9
    assert np.array_equal(pose["DT"][0].reshape(4, 4), np.eye(4)) and not pose["DT_cov"][0].any()


def test_vector_stdv_mad_pin(orc):
    rng = np.random.default_rng(8)
    for n in (1, 2, 5, 100, 257):
        r = np.abs(rng.normal(0, 1, n)) * rng.choice([1, 1, 1, 8], n)
        assert orc.vector_stdv_mad(r) == R.vector_stdv_mad(r)            # float-rounded deviations: exact


def test_line_descriptor_mad_pin(orc):
    rng = np.random.default_rng(9)
    for n in (2, 3, 50, 101):
        d = np.sort(rng.integers(0, 128, (n, 2)), axis=1).astype(np.float32)
        for U in (0.0, 3.5):
            assert orc.line_descriptor_mad(np.ascontiguousarray(d), U) == R.line_descriptor_mad_nn12(d, U)
