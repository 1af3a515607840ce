"""TEST INFRASTRUCTURE — independent numpy restatement of the reference's fp64 stages, written from the reference
TEXT (file:line cited per function), NOT from oracle/gfpl_oracle.cpp.  It pins the C++ oracle: two restatements made
independently from the same source agree to 1e-12 only if both follow it.  Eigen's solvers are replaced by LAPACK
(numpy.linalg.solve / inv / cholesky / eigvalsh), which is the point: nothing here shares code, loop order or
factorisation with the oracle.

Matrices are row-major 4x4 / 6x6 numpy arrays.  Twist order is (translation; rotation) (auxiliar.cpp:169-170).
"""
import numpy as np

F32 = np.float32


# ---- pinholeStereoCamera.cpp:133-170 ---------------------------------------------------------------------------------
def back_projection(cam, u, v, disp):
    bd = cam.b / disp
    return np.array([bd * (u - cam.cx), bd * (v - cam.cy), bd * cam.fx])


def projection(cam, P):
    return np.array([cam.cx + cam.fx * P[0] / P[2], cam.cy + cam.fy * P[1] / P[2]])


def get_disparity(cam, Z):
    return cam.fx * cam.b / Z


# ---- auxiliar.cpp:154-238 ----------------------------------------------------------------------------------------------
def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def inverse_se3(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def expmap_se3(x):
    t, w = x[:3].copy(), x[3:]
    theta = np.linalg.norm(w)
    if theta < 0.000001:
        R = np.eye(3)
    else:
        s = skew(w) / theta
        R = np.eye(3) + s * np.sin(theta) + s @ s * (1.0 - np.cos(theta))
        V = np.eye(3) + s * (1.0 - np.cos(theta)) / theta + s @ s * (theta - np.sin(theta)) / theta
        t = V @ t
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return T


def adjoint_se3(T):
    A = np.zeros((6, 6))
    R = T[:3, :3]
    A[:3, :3] = R
    A[:3, 3:] = skew(T[:3, 3]) @ R
    A[3:, 3:] = R
    return A


def unccomp_se3(T1, cov1, cov_inc):
    A = adjoint_se3(T1)
    return cov1 + A @ cov_inc @ A.T


def vector_stdv_mad(res):
    """auxiliar.cpp:521-537: element [n/2] of the sorted values; deviations through fabsf (float)"""
    if len(res) == 0:
        return 0.0
    r = np.sort(np.asarray(res, np.float64))
    n = len(r)
    med = r[n // 2]
    dev = np.sort(np.abs((r - med).astype(F32)).astype(np.float64))
    return 1.4826 * dev[n // 2]


# ---- stereoFrameHandler.cpp:2118-2245 (optimizeFunctions) ---------------------------------------------------------------
def _jac6(k, X, Y, Z, lx, ly):
    return k * np.array([lx * Z, ly * Z, -(X * lx + Y * ly), -(X * Y * lx + Y * Y * ly + Z * Z * ly),
                         X * X * lx + Z * Z * lx + X * Y * ly, X * Z * ly - Y * Z * lx])


def optimize_functions(cam, cfg, DT, pts, pin, lns, lin):
    R, t = DT[:3, :3], DT[:3, 3]
    H = np.zeros((6, 6)); g = np.zeros(6); e = 0.0; N = 0
    for k in range(len(pts)):
        if not pin[k]:
            continue
        P = R @ pts["P"][k] + t
        d = projection(cam, P) - pts["pl_obs"][k]
        r = np.sqrt(d @ d)
        J = _jac6(cam.fx / max(cfg.homog_th, P[2] * P[2]), P[0], P[1], P[2], d[0], d[1]) / max(cfg.homog_th, r)
        w = 1.0 / (1.0 + r * r * pts["sigma2"][k])
        H += np.outer(J, J) * w; g += J * r * w; e += r * r * w; N += 1
    for k in range(len(lns)):
        if not lin[k]:
            continue
        l = lns["le_obs"][k]
        Ps, Pe = R @ lns["sP"][k] + t, R @ lns["eP"][k] + t
        ds = l[:2] @ projection(cam, Ps) + l[2]
        de = l[:2] @ projection(cam, Pe) + l[2]
        r = np.sqrt(ds * ds + de * de)
        Js = _jac6(cam.fx / max(cfg.homog_th, Ps[2] * Ps[2]), Ps[0], Ps[1], Ps[2], l[0], l[1])
        Je = _jac6(cam.fx / max(cfg.homog_th, Pe[2] * Pe[2]), Pe[0], Pe[1], Pe[2], l[0], l[1])
        J = (Js * ds + Je * de) / max(cfg.homog_th, r)
        w = 1.0 / (1.0 + r * r * lns["sigma2"][k])
        H += np.outer(J, J) * w; g += J * r * w; e += r * r * w; N += 1
    return H, g, e / N


# ---- stereoFrameHandler.cpp:2032-2056 (gaussNewtonOptimization) ----------------------------------------------------------
def gauss_newton(cam, cfg, DT, pts, pin, lns, lin, max_iters):
    err_prev = 999999999.9
    iters_done = 0
    H = np.zeros((6, 6)); err = 0.0
    for it in range(max_iters):
        H, g, err = optimize_functions(cam, cfg, DT, pts, pin, lns, lin)
        iters_done = it + 1
        if abs(err - err_prev) < cfg.min_error_change or err < cfg.min_error:
            break
        inc = np.linalg.solve(H, g)                       # LDLT<Matrix6d>(H).solve(g)
        DT = DT @ inverse_se3(expmap_se3(inc))
        if np.linalg.norm(inc) < np.finfo(np.float64).eps:
            break
        err_prev = err
    return DT, np.linalg.inv(H), err, iters_done


# ---- stereoFrameHandler.cpp:2058-2116 (removeOutliers) --------------------------------------------------------------------
def remove_outliers(cam, cfg, DT, pts, pin, lns, lin):
    R, t = DT[:3, :3], DT[:3, 3]
    res_p = [np.linalg.norm(projection(cam, R @ pts["P"][k] + t) - pts["pl_obs"][k]) * np.sqrt(pts["sigma2"][k])
             for k in range(len(pts))]
    res_l = []
    for k in range(len(lns)):
        l = lns["le_obs"][k]
        e0 = l[:2] @ projection(cam, R @ lns["sP"][k] + t) + l[2]
        e1 = l[:2] @ projection(cam, R @ lns["eP"][k] + t) + l[2]
        res_l.append(np.sqrt(e0 * e0 + e1 * e1) * np.sqrt(lns["sigma2"][k]))
    th_p = cfg.inlier_k * vector_stdv_mad(res_p)
    th_l = cfg.inlier_k * vector_stdv_mad(res_l)
    pin, lin = pin.copy(), lin.copy()
    for k in range(len(pts)):
        if res_p[k] > th_p:
            pin[k] = 0
    for k in range(len(lns)):
        if res_l[k] > th_l:
            lin[k] = 0
    return pin, lin


# ---- stereoFrameHandler.cpp:1939-2030 (optimizePose(DT_ini)) ----------------------------------------------------------------
def optimize_pose(cam, cfg, DT_ini, state, pts, lns):
    """returns dict(DT, Tfw, DT_cov, Tfw_cov, err_norm, pin, lin, iters_first, iters_ref)"""
    pin = np.ones(len(pts), np.uint8); lin = np.ones(len(lns), np.uint8)      # the matchers set inlier = true
    prev_Tfw, prev_Tfw_cov = state["Tfw"].reshape(4, 4), state["Tfw_cov"].reshape(6, 6)
    DT = DT_ini.copy()
    out = dict(iters_first=0, iters_ref=0)
    n_inliers = len(pts) + len(lns)
    if n_inliers > cfg.min_features:
        DT_, DT_cov, err, out["iters_first"] = gauss_newton(cam, cfg, DT.copy(), pts, pin, lns, lin, cfg.max_iters)
        if np.isfinite(DT_).all():
            pin, lin = remove_outliers(cam, cfg, DT_, pts, pin, lns, lin)
            if int(pin.sum()) + int(lin.sum()) > cfg.min_features:
                # D8: the refinement restarts from the initial guess, not from DT_
                DT, DT_cov, err, out["iters_ref"] = gauss_newton(cam, cfg, DT, pts, pin, lns, lin, cfg.max_iters_ref)
            else:
                DT, DT_cov = np.eye(4), np.zeros((6, 6))
        else:
            DT, DT_cov = np.eye(4), np.zeros((6, 6))
    else:
        DT, DT_cov, err = np.eye(4), np.zeros((6, 6)), 0.0
    DTi = inverse_se3(DT)
    accept = np.isfinite(DT).all() and np.linalg.norm(DTi[:3, 3]) < cfg.motion_step_th * (state["time_stamp_curr"] - state["time_stamp"])
    if accept:
        out.update(DT=DTi, Tfw=prev_Tfw @ DTi, Tfw_cov=unccomp_se3(prev_Tfw, prev_Tfw_cov, DT_cov), err_norm=err)
    else:
        out.update(DT=np.eye(4), Tfw=prev_Tfw, Tfw_cov=prev_Tfw_cov, err_norm=-1.0)
    out.update(DT_cov=DT_cov, DT_cov_eig=np.linalg.eigvalsh(DT_cov), pin=pin, lin=lin)
    return out


# ---- stereoFrame.cpp:1375-1484 (getJacob2D_3D, getCovMat2D_3D, estimateStereoUncertainty) -----------------------------------
def cov_mat_2d_3d(cam, u, v, u_std, d, d_std):
    b = cam.b
    J = np.array([[b / d, 0, -(u - cam.cx) * b / (d * d)],
                  [0, b / d, -(v - cam.cy) * b / (d * d)],
                  [0, 0, -cam.fx * b / (d * d)]])
    cov2d = np.diag([u_std ** 2, u_std ** 2, d_std ** 2])      # D5: off-diagonals are zero
    return J @ cov2d @ J.T


def stereo_uncertainty(cam, cfg, line):
    """line: one line_feature record -> (covSpt3D, covEpt3D)"""
    ratio = cfg.ratio_disp_std if abs(line["le"][0]) > 0.15 else cfg.ratio_disp_std_hor
    return (cov_mat_2d_3d(cam, line["spl"][0], line["spl"][1], 1.0, line["sdisp"], ratio * line["sdisp"]),
            cov_mat_2d_3d(cam, line["epl"][0], line["epl"][1], 1.0, line["edisp"], ratio * line["edisp"]))


# ---- stereoFrameHandler.cpp:1342-1447 (getPoseInfoOnLine / getPoseInfoPoint), stereoFrame.cpp:1394-1413 ----------------------
def jacob_3d_2d(cam, px, py, pz):
    f, b = cam.fx, cam.b
    return np.array([[f / pz, 0, -f * px / (pz * pz)], [0, f / pz, -f * py / (pz * pz)], [0, 0, -f * b / (pz * pz)]])


def pose_info_on_line(cam, cfg, DT_inv, sP, eP, le_obs, covS, covE, cut):
    c0, c1 = cut
    sP_t = (1 - c0) * sP + c0 * eP
    eP_t = (1 - c1) * eP + c1 * sP
    cov_s = (1 - c0) ** 2 * covS + c0 ** 2 * covE
    cov_e = (1 - c1) ** 2 * covE + c1 ** 2 * covS
    Rm, t = DT_inv[:3, :3], DT_inv[:3, 3]
    jl = le_obs[:2]
    cols, var = [], []
    for P_t, cov in ((sP_t, cov_s), (eP_t, cov_e)):
        cur = Rm @ P_t + t
        Jp = jacob_3d_2d(cam, *cur)[:2]
        var.append(jl @ (Jp @ Rm @ cov @ Rm.T @ Jp.T) @ jl)
        cols.append(_jac6(cam.fx / max(cfg.homog_th, cur[2] * cur[2]), cur[0], cur[1], cur[2], jl[0], jl[1]))
    Jac = np.stack(cols, 1)                                          # 6 x 2
    return Jac @ np.linalg.inv(np.diag(var)) @ Jac.T


def pose_info_point(cam, cfg, DT_inv, P, pl_obs):
    cur = DT_inv[:3, :3] @ P + DT_inv[:3, 3]
    d = projection(cam, cur) - pl_obs
    J = _jac6(cam.fx / max(cfg.homog_th, cur[2] * cur[2]), cur[0], cur[1], cur[2], d[0], d[1])
    return np.outer(J, J)


def logdet(M):
    """linespec.h:43-56: 2 * sum(log(diag(LLT(M).matrixL())))"""
    return 2.0 * np.log(np.diag(np.linalg.cholesky(M))).sum()


# ---- stereoFrameHandler.cpp:1618-1764 (estimateProjUncertainty_submodular), :1451-1470 (updateEndPointByRatio) ----------------
def linecut_submodular(cam, cfg, DT_inv, pts, lns, covs):
    """lns: ln_obs records (sP, eP, le_obs), covs: [(covS, covE)] per line.  Returns cut ratios [n][2], info matrices, and
    the moved end points.  Metric = log-det (Config::cutWithMaxVol)."""
    step, rng = cfg.cut_step, (cfg.cut_rng[0], cfg.cut_rng[1])
    nb = [(step, 0), (-step, 0), (0, step), (0, -step), (step, step), (step, -step), (-step, step), (-step, -step)]
    n = len(lns)
    sP = lns["sP"].copy(); eP = lns["eP"].copy()
    cut = np.zeros((n, 2)); info = np.zeros((n, 6, 6))
    total = np.zeros((6, 6))
    for l in range(n):
        info[l] = pose_info_on_line(cam, cfg, DT_inv, sP[l], eP[l], lns["le_obs"][l], covs[l][0], covs[l][1], cut[l])
        total += info[l]
    for k in range(len(pts)):
        total += pose_info_point(cam, cfg, DT_inv, pts["P"][k], pts["pl_obs"][k])
    for l in range(n):
        metric_last = logdet(total)
        total -= info[l]
        while cut[l, 0] + cut[l, 1] <= 1.0:
            best, hit = metric_last, None
            for dx, dy in nb:
                ct = (cut[l, 0] + dx, cut[l, 1] + dy)
                if ct[0] + ct[1] > 1.0 or ct[0] < rng[0] or ct[0] > rng[1] or ct[1] < rng[0] or ct[1] > rng[1]:
                    continue
                tmp = pose_info_on_line(cam, cfg, DT_inv, sP[l], eP[l], lns["le_obs"][l], covs[l][0], covs[l][1], ct)
                m = logdet(tmp + total)
                if m > best:
                    best, hit = m, (ct, tmp)
            if hit is None:
                break
            cut[l] = hit[0]; info[l] = hit[1]; metric_last = best
        if not (abs(cut[l, 0]) < 0.0001 and abs(cut[l, 1]) < 0.0001):
            if abs(cut[l, 0]) > 0.0001:
                sP[l] = (1 - cut[l, 0]) * sP[l] + cut[l, 0] * eP[l]
            if abs(cut[l, 1]) > 0.0001:
                eP[l] = (1 - cut[l, 1]) * eP[l] + cut[l, 1] * sP[l]          # reads the MOVED start point
        total += info[l]
    return cut, info, sP, eP


# ---- stereoFrameHandler.cpp:1249-1338 (getPoseCovOnLine), :1476-1610 (estimateProjUncertainty_descent), linespec.cpp:37-51 ------
def pose_cov_on_line(cam, cfg, DT_inv, sP, eP, le_obs, covS, covE, cut):
    c0, c1 = cut
    sP_t = (1 - c0) * sP + c0 * eP
    eP_t = (1 - c1) * eP + c1 * sP
    cov_s = (1 - c0) ** 2 * covS + c0 ** 2 * covE
    cov_e = (1 - c1) ** 2 * covE + c1 ** 2 * covS
    Rm, t = DT_inv[:3, :3], DT_inv[:3, 3]
    jl = le_obs[:2]
    cols, var = [], []
    for P_t, cov in ((sP_t, cov_s), (eP_t, cov_e)):
        cur = Rm @ P_t + t
        Jp = jacob_3d_2d(cam, *cur)[:2]
        var.append(jl @ (Jp @ Rm @ cov @ Rm.T @ Jp.T) @ jl)
        cols.append(_jac6(cam.fx / max(cfg.homog_th, cur[2] * cur[2]), cur[0], cur[1], cur[2], jl[0], jl[1]))
    Jac = np.stack(cols, 1)                                          # Jac_r2x, 6 x 2
    Jinv = np.linalg.inv(Jac.T @ Jac) @ Jac.T                        # 2 x 6
    return Jinv.T @ np.diag(var) @ Jinv


def eigen_values_desc(M):
    """getEigenValue: real parts of the GENERAL eigen-solver's values, sorted descending (LAPACK dgeev for Eigen::EigenSolver)"""
    return np.sort(np.linalg.eigvals(M).real)[::-1]


def linecut_descent(cam, cfg, DT_inv, lns, covs):
    """Per line (independent): returns cut [n][2], the volume track per line, and the rewritten fields."""
    step, rng = cfg.cut_step, (cfg.cut_rng[0], cfg.cut_rng[1])
    nb = [(step, 0), (-step, 0), (0, step), (0, -step), (step, step), (step, -step), (-step, step), (-step, -step)]
    Rm, t = DT_inv[:3, :3], DT_inv[:3, 3]
    out = []
    for l in range(len(lns)):
        sP, eP, le = lns["sP"][l], lns["eP"][l], lns["le_obs"][l]
        covS, covE = covs[l]
        cut = [0.0, 0.0]
        ev = eigen_values_desc(pose_cov_on_line(cam, cfg, DT_inv, sP, eP, le, covS, covE, cut))
        vols = [ev[0] * ev[1]]
        margins = []                                               # relative gap of every comparison made (near ties)
        while cut[0] + cut[1] < 1.0:
            hit = False
            for dx, dy in nb:
                ct = (cut[0] + dx, cut[1] + dy)
                if ct[0] + ct[1] >= 1.0 or ct[0] < rng[0] or ct[0] > rng[1] or ct[1] < rng[0] or ct[1] > rng[1]:
                    continue
                ev = eigen_values_desc(pose_cov_on_line(cam, cfg, DT_inv, sP, eP, le, covS, covE, ct))
                v = ev[0] * ev[1]
                margins.append(abs(v - vols[-1]) / max(abs(vols[-1]), 1e-300))
                if v < vols[-1]:
                    vols.append(v); cut = [ct[0], ct[1]]; hit = True   # the cut moves inside the neighbour loop
            if not hit:
                break
        sP_cut = (1 - cut[0]) * sP + cut[0] * eP
        eP_cut = (1 - cut[1]) * eP + cut[1] * sP
        cov_s = (1 - cut[0]) ** 2 * covS + cut[0] ** 2 * covE
        cov_e = (1 - cut[1]) ** 2 * covE + cut[1] ** 2 * covS
        cur_s, cur_e = Rm @ sP_cut + t, Rm @ eP_cut + t
        Js, Je = jacob_3d_2d(cam, *cur_s), jacob_3d_2d(cam, *cur_e)
        out.append(dict(cut=np.array(cut), vols=vols, min_margin=min(margins) if margins else 1.0,
                        covl=(Js @ Rm @ cov_s @ Rm.T @ Js.T, Je @ Rm @ cov_e @ Rm.T @ Je.T),
                        lPt2D=np.stack([projection(cam, Rm @ sP + t), projection(cam, cur_s), projection(cam, cur_e),
                                        projection(cam, Rm @ eP + t)]),
                        sP=sP_cut, eP=eP_cut, spl=projection(cam, sP_cut), epl=projection(cam, eP_cut),
                        sdisp=get_disparity(cam, sP_cut[2]), edisp=get_disparity(cam, eP_cut[2])))
    return out


# ---- stereoFrame.cpp:881-1014 (line part of extractStereoFeatures_Brute) -------------------------------------------------------
def hamming2_matrix(a, b):
    """NORM_HAMMING2: number of differing 2-bit groups"""
    x = a[:, None, :] ^ b[None, :, :]
    g = (x | (x >> 1)) & 0x55
    return np.unpackbits(g, axis=2).sum(2).astype(np.int32)


def hamming_matrix(a, b):
    return np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2).astype(np.int32)


def knn2_from_matrix(D):
    """cv::BFMatcher::knnMatch(k=2): ascending distance, ties -> ascending trainIdx (SURVEY §8c probe)"""
    order = np.argsort(D, axis=1, kind="stable")[:, :2]
    return order, np.take_along_axis(D, order, 1).astype(F32)


def line_segment_overlap_stereo(spl_obs, epl_obs, spl_proj, epl_proj):
    sln, eln = min(spl_obs, epl_obs), max(spl_obs, epl_obs)
    spn, epn = min(spl_proj, epl_proj), max(spl_proj, epl_proj)
    length = eln - spn
    if epn < sln or spn > eln:
        overlap = 0.0
    elif epn > eln and spn < sln:
        overlap = eln - sln
    else:
        overlap = min(eln, epn) - max(sln, spn)
    return overlap / length if length > F32(0.01) else 0.0


def line_descriptor_mad_nn12(dist, U=0.0):
    """stereoFrame.cpp:1287-1313, second output.  `nn_dist_median` is read uninitialised there: U."""
    dev = np.abs((dist[:, 1] - dist[:, 0] - F32(U)).astype(F32))
    return 1.4826 * float(np.sort(dev)[len(dev) // 2])


def stereo_lines_brute(cam, cfg, kl_l, kl_r, desc_l, desc_r):
    """-> list of dicts (one per accepted line, in loop order) with the LineFeature fields + src_l/src_r"""
    D = hamming2_matrix(desc_l, desc_r)
    lr_i, lr_d = knn2_from_matrix(D)
    rl_i, _ = knn2_from_matrix(D.T)
    th = line_descriptor_mad_nn12(lr_d, cfg.mad_uninit_value) * cfg.desc_th_l
    out = []
    for i in range(min(len(desc_l), len(desc_r))):                 # D4: n_matches = min(size(lr), size(rl))
        q, tdx = i, lr_i[i, 0]
        if rl_i[tdx, 0] != q or not (float(lr_d[i, 1] - lr_d[i, 0]) > th):
            continue
        L, Rr = kl_l[q], kl_r[tdx]
        sp_l = np.array([L["startPointX"], L["startPointY"], 1.0], np.float64)
        ep_l = np.array([L["endPointX"], L["endPointY"], 1.0], np.float64)
        le_l = np.cross(sp_l, ep_l); le_l = le_l / np.sqrt(le_l[0] ** 2 + le_l[1] ** 2)
        sp_r = np.array([Rr["startPointX"], Rr["startPointY"], 1.0], np.float64)
        ep_r = np.array([Rr["endPointX"], Rr["endPointY"], 1.0], np.float64)
        le_r = np.cross(sp_r, ep_r)
        overlap = line_segment_overlap_stereo(sp_l[1], ep_l[1], sp_r[1], ep_r[1])
        with np.errstate(divide="ignore", invalid="ignore"):
            xs = -(le_r[2] + le_r[1] * np.float64(L["startPointY"])) / le_r[0]
            xe = -(le_r[2] + le_r[1] * np.float64(L["endPointY"])) / le_r[0]
        disp_s = np.float64(L["startPointX"]) - xs
        disp_e = np.float64(L["endPointX"]) - xe
        if not (disp_s >= cfg.min_disp and disp_e >= cfg.min_disp and float(abs(F32(le_l[0]))) > cfg.line_horiz_th
                and overlap > cfg.stereo_overlap_th):
            continue
        sP = back_projection(cam, sp_l[0], sp_l[1], disp_s)
        eP = back_projection(cam, ep_l[0], ep_l[1], disp_e)
        eig = []
        for p, disp in ((sp_l, disp_s), (ep_l, disp_e)):
            px, py, d2, f = p[0] - cam.cx, p[1] - cam.cy, disp * disp, cam.fx
            C = np.array([[d2 + 2 * px * px, 2 * px * py, 2 * f * px],
                          [2 * px * py, d2 + 2 * py * py, 2 * f * py],
                          [2 * f * px, 2 * f * py, 2 * f * f]]) * cam.b * cam.b / (d2 * d2)
            eig.append(np.linalg.eigvalsh(C)[2])
        if not max(eig) < cfg.line_cov_th:
            continue
        out.append(dict(spl=sp_l[:2], sdisp=disp_s, sP=sP, epl=ep_l[:2], edisp=disp_e, eP=eP, le=le_l,
                        angle=float(L["angle"]), level=int(L["octave"]), src_l=q, src_r=int(tdx)))
    return out


# ---- stereoFrame.cpp:795-879 (point part of extractStereoFeatures_Brute) ----------------------------------------------------------
def stereo_points_brute(cam, cfg, kp_l, kp_r, desc_l, desc_r):
    D = hamming_matrix(desc_l, desc_r)
    lr_i, lr_d = knn2_from_matrix(D)
    rl_i, _ = knn2_from_matrix(D.T)
    out = []
    for i in range(len(desc_l)):
        q, tdx = i, lr_i[i, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.float64(lr_d[i, 0] / lr_d[i, 1])              # float / float, widened
        if rl_i[tdx, 0] != q or not (ratio <= cfg.max_ratio_12_p):
            continue
        if not (float(abs(F32(kp_l["y"][q]) - F32(kp_r["y"][tdx]))) <= cfg.max_dist_epip):
            continue
        disp = np.float64(F32(kp_l["x"][q]) - F32(kp_r["x"][tdx]))   # float subtraction assigned to a double
        if not disp >= cfg.min_disp:
            continue
        disp_sub = F32(disp)
        if not disp_sub >= 0:
            continue
        u, v = np.float64(kp_l["x"][q]), np.float64(kp_l["y"][q])
        s = np.float64(1.0)
        for _ in range(int(kp_l["octave"][q]) + 1):                  # stereoFeatures.cpp:41-48
            s *= cfg.orb_scale_factor
        out.append(dict(pl=np.array([u, v]), disp=np.float64(disp_sub), P=back_projection(cam, u, v, np.float64(disp_sub)),
                        sigma2=1.0 / (s * s), level=int(kp_l["octave"][q]), src_l=q, src_r=int(tdx)))
    return out


# ---- stereoFrameHandler.cpp:451-686 (crossFrameMatching_Hybrid) -------------------------------------------------------------------
def f2f_hybrid(cam, cfg, prev_pt, prev_pdesc, curr_pt, curr_pdesc, prev_ln, prev_ldesc, curr_ln, curr_ldesc, prev_Tfw, prev_DT):
    """-> (list of (prev_idx, curr_idx) point matches in matched_pt order, list of line matches in matched_ls order)"""
    pts, lns = [], []
    if len(prev_pt) and len(curr_pt):
        D = hamming_matrix(prev_pdesc, curr_pdesc)
        curr_Tfw = prev_Tfw @ prev_DT                                     # :157 constant-velocity prediction
        Tcw = np.linalg.inv(curr_Tfw)
        cand = {}                                                          # multimap<curr idx, (prev idx, distance)>
        for i in range(len(prev_pt)):
            Pw = prev_Tfw @ np.append(prev_pt["P"][i], 1.0)
            proj = projection(cam, (Tcw @ Pw)[:3])
            for j in np.nonzero(D[i] <= cfg.point_match_radius)[0]:      # radiusMatch: inclusive
                if np.linalg.norm(proj - curr_pt["pl"][j]) > cfg.f2f_point_window:
                    continue
                cand.setdefault(int(j), []).append((i, F32(D[i, j])))
        for key in sorted(cand):                                          # D2: distinct keys ascending
            best_i, best_d = 0, F32(1e7)
            for i, d in cand[key]:                                        # insertion order = ascending prev index
                if d < best_d:
                    best_i, best_d = i, d
            pts.append((best_i, key))
            if len(pts) >= cfg.max_point_match_num:
                break
    if len(prev_ln) >= 2 and len(curr_ln) >= 2:                           # D3: [i][1] is read
        D = hamming_matrix(prev_ldesc, curr_ldesc)
        i12, d12 = knn2_from_matrix(D)
        i21, _ = knn2_from_matrix(D.T)
        th = line_descriptor_mad_nn12(d12, cfg.mad_uninit_value) * cfg.desc_th_l
        budget = np.float64(np.sort(d12[:, 0])[min(cfg.max_line_match_num, len(d12)) - 1])
        for i in range(len(prev_ln)):
            t = i12[i, 0]
            if np.float64(d12[i, 0]) > cfg.budget_slack * budget:
                continue
            if i21[t, 0] == i and np.float64(d12[i, 1] - d12[i, 0]) > th:
                lns.append((i, int(t)))
            if len(lns) >= cfg.max_line_match_num:
                break
    return pts, lns
