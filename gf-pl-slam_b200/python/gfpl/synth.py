"""Seeded synthetic EuRoC-/KITTI-shaped stereo frame pairs.

The reference consumes detector output (ORB key points, LSD key lines, 32-byte descriptors); detection
stays on the host and is out of scope, so the replay harness injects it.  This module synthesises that
detector output for B independent frame pairs: a random 3-D scene per pair, a small random SE(3) motion
(cf. reference src/simulate_line_cut.cpp:117-126), pixel noise, 256-bit descriptors with bit flips,
distractor features, per-image shuffling.  No image data is involved.
"""
import numpy as np

from . import _structs as S

EUROC = dict(fx=458.654, fy=458.654, cx=367.215, cy=248.375, b=0.110078, width=752, height=480,
             z_rng=(2.0, 20.0))   # reference config/euroc_params.yaml
KITTI = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, b=0.537166, width=1241, height=376,
             z_rng=(5.0, 80.0))   # reference config/kitti/kitti00-02.yaml


def camera(shape="euroc"):
    c = EUROC if shape == "euroc" else KITTI
    return S.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["b"], c["width"], c["height"])


def _expmap(x):
    """Batched SE(3) exponential, twist x=(t;w) [B,6] -> [B,4,4] (same formula as auxiliar.cpp:165-182)."""
    B = len(x)
    t, w = x[:, :3], x[:, 3:]
    th = np.linalg.norm(w, axis=1)
    T = np.tile(np.eye(4), (B, 1, 1))
    small = th < 1e-6
    ths = np.where(small, 1.0, th)
    s = np.zeros((B, 3, 3))
    s[:, 0, 1], s[:, 0, 2], s[:, 1, 2] = -w[:, 2], w[:, 1], -w[:, 0]
    s[:, 1, 0], s[:, 2, 0], s[:, 2, 1] = w[:, 2], -w[:, 1], w[:, 0]
    s = s / ths[:, None, None]
    ss = s @ s
    sn, cs = np.sin(ths)[:, None, None], np.cos(ths)[:, None, None]
    I = np.eye(3)[None]
    R = I + s * sn + ss * (1 - cs)
    V = I + s * (1 - cs) / ths[:, None, None] + ss * (ths[:, None, None] - sn) / ths[:, None, None]
    R = np.where(small[:, None, None], I, R)
    V = np.where(small[:, None, None], I, V)
    T[:, :3, :3] = R
    T[:, :3, 3] = np.einsum("bij,bj->bi", V, t)
    return T


def _inv(T):
    Ti = np.tile(np.eye(4), (len(T), 1, 1))
    Rt = np.transpose(T[:, :3, :3], (0, 2, 1))
    Ti[:, :3, :3] = Rt
    Ti[:, :3, 3] = -np.einsum("bij,bj->bi", Rt, T[:, :3, 3])
    return Ti


def _flip_mask(rng, shape):
    """random byte mask with each bit set with probability 1/16 (~0.05 flips)."""
    m = rng.integers(0, 256, size=shape, dtype=np.uint8)
    for _ in range(3):
        m &= rng.integers(0, 256, size=shape, dtype=np.uint8)
    return m


def _project(c, P):
    z = P[..., 2]
    u = c["cx"] + c["fx"] * P[..., 0] / z
    v = c["cy"] + c["fy"] * P[..., 1] / z
    d = c["fx"] * c["b"] / z
    return u, v, d


def _in_image(c, u, v, margin=2.0):
    return (u >= margin) & (u < c["width"] - margin) & (v >= margin) & (v < c["height"] - margin)


def _sample_points(rng, c, B, N):
    u = rng.uniform(20, c["width"] - 20, (B, N))
    v = rng.uniform(20, c["height"] - 20, (B, N))
    z = rng.uniform(c["z_rng"][0], c["z_rng"][1], (B, N))
    return np.stack([(u - c["cx"]) * z / c["fx"], (v - c["cy"]) * z / c["fy"], z], axis=-1)


def _fill_points(rng, c, raw, side_l, side_r, dsc_l, dsc_r, P, code, n, n_lm, noise):
    """observe landmarks P [B,n_lm,3] (camera frame) in a stereo pair; pad to n with distractors; shuffle."""
    B = P.shape[0]
    u, v, d = _project(c, P)
    ul = u + rng.normal(0, noise, u.shape)
    vl = v + rng.normal(0, noise, u.shape)
    ur = u - d + rng.normal(0, noise, u.shape)
    vr = v + rng.normal(0, noise, u.shape)
    ok_l = _in_image(c, ul, vl) & (P[..., 2] > 0.5)
    ok_r = _in_image(c, ur, vr) & (P[..., 2] > 0.5)

    def one(uu, vv, ok, side, dsc):
        x = np.empty((B, n), np.float32)
        y = np.empty((B, n), np.float32)
        desc = np.empty((B, n, 32), np.uint8)
        x[:, :n_lm] = uu
        y[:, :n_lm] = vv
        desc[:, :n_lm] = code ^ _flip_mask(rng, code.shape)
        # distractors: the padding rows and every landmark that left the image
        rx = rng.uniform(2, c["width"] - 2, (B, n)).astype(np.float32)
        ry = rng.uniform(2, c["height"] - 2, (B, n)).astype(np.float32)
        rd = rng.integers(0, 256, size=(B, n, 32), dtype=np.uint8)
        bad = np.ones((B, n), bool)
        bad[:, :n_lm] = ~ok
        x = np.where(bad, rx, x)
        y = np.where(bad, ry, y)
        desc = np.where(bad[..., None], rd, desc)
        perm = np.argsort(rng.random((B, n)), axis=1)
        bi = np.arange(B)[:, None]
        side["x"][:, :n] = x[bi, perm]
        side["y"][:, :n] = y[bi, perm]
        side["octave"][:, :n] = rng.integers(0, 4, (B, n))
        side["class_id"][:, :n] = -1
        dsc[:, :n] = desc[bi, perm]

    one(ul, vl, ok_l, side_l, dsc_l)
    one(ur, vr, ok_r, side_r, dsc_r)


def _fill_lines(rng, c, side_l, side_r, dsc_l, dsc_r, Ps, Pe, code, n, n_lm, noise):
    B = Ps.shape[0]
    us, vs, ds = _project(c, Ps)
    ue, ve, de = _project(c, Pe)
    nz = lambda: rng.normal(0, noise, us.shape)
    l = [us + nz(), vs + nz(), ue + nz(), ve + nz()]
    r = [us - ds + nz(), vs + nz(), ue - de + nz(), ve + nz()]
    zok = (Ps[..., 2] > 0.5) & (Pe[..., 2] > 0.5)
    ok_l = _in_image(c, l[0], l[1]) & _in_image(c, l[2], l[3]) & zok
    ok_r = _in_image(c, r[0], r[1]) & _in_image(c, r[2], r[3]) & zok

    def one(g, ok, side, dsc):
        geo = np.empty((B, n, 4), np.float32)
        desc = np.empty((B, n, 32), np.uint8)
        for k in range(4):
            geo[:, :n_lm, k] = g[k]
        desc[:, :n_lm] = code ^ _flip_mask(rng, code.shape)
        rg = np.empty((B, n, 4), np.float32)
        rg[..., 0] = rng.uniform(2, c["width"] - 2, (B, n))
        rg[..., 1] = rng.uniform(2, c["height"] - 2, (B, n))
        ang = rng.uniform(0, 2 * np.pi, (B, n))
        ln = rng.uniform(15, 80, (B, n))
        rg[..., 2] = np.clip(rg[..., 0] + ln * np.cos(ang), 2, c["width"] - 2)
        rg[..., 3] = np.clip(rg[..., 1] + ln * np.sin(ang), 2, c["height"] - 2)
        rd = rng.integers(0, 256, size=(B, n, 32), dtype=np.uint8)
        bad = np.ones((B, n), bool)
        bad[:, :n_lm] = ~ok
        geo = np.where(bad[..., None], rg, geo)
        desc = np.where(bad[..., None], rd, desc)
        perm = np.argsort(rng.random((B, n)), axis=1)
        bi = np.arange(B)[:, None]
        geo = geo[bi, perm]
        side["startPointX"][:, :n] = geo[..., 0]
        side["startPointY"][:, :n] = geo[..., 1]
        side["endPointX"][:, :n] = geo[..., 2]
        side["endPointY"][:, :n] = geo[..., 3]
        dx, dy = geo[..., 2] - geo[..., 0], geo[..., 3] - geo[..., 1]
        side["angle"][:, :n] = np.arctan2(dy, dx)
        side["octave"][:, :n] = 0
        side["lineLength"][:, :n] = np.sqrt(dx * dx + dy * dy)
        side["response"][:, :n] = side["lineLength"][:, :n]
        dsc[:, :n] = desc[bi, perm]

    one(l, ok_l, side_l, dsc_l)
    one(r, ok_r, side_r, dsc_r)


def make_pairs(B, n_pt=300, n_ln=100, shape="euroc", seed=0, cap_pt=None, cap_ln=None, distractor_frac=0.1,
               noise_px=0.5, pinned=False, ragged=False):
    """B independent frame pairs.  Returns dict(cam, prev_raw, curr_raw, state, DT_true).

    prev_raw / curr_raw: RawFramesNP (detector output of the previous / current stereo images);
    state: frame_state_dt[B] of the previous frame (Tfw, DT = stored inverse increment, covariances, stamps);
    DT_true: [B,4,4] ground-truth prev->curr transform.
    ragged=True draws per-pair feature counts in [60%,100%] of n_pt/n_ln (parity tests only).
    """
    c = EUROC if shape == "euroc" else KITTI
    rng = np.random.Generator(np.random.PCG64(seed))
    cap_pt = cap_pt or n_pt
    cap_ln = cap_ln or n_ln
    prev = S.RawFramesNP(B, cap_pt, cap_ln, pinned)
    curr = S.RawFramesNP(B, cap_pt, cap_ln, pinned)

    tw = np.zeros((B, 6))
    tdir = rng.normal(size=(B, 3))
    tw[:, :3] = tdir / np.linalg.norm(tdir, axis=1, keepdims=True) * rng.uniform(0, 0.1, (B, 1))
    wdir = rng.normal(size=(B, 3))
    tw[:, 3:] = wdir / np.linalg.norm(wdir, axis=1, keepdims=True) * rng.uniform(0, 0.05, (B, 1))
    DT_true = _expmap(tw)                          # prev -> curr coordinates
    # stored prev->DT is the INVERSE increment of the previous step; constant velocity
    # => ~ inverse(DT_true); "prediction = truth perturbed by 10 %"
    DT_prev = _inv(_expmap(tw * rng.uniform(0.9, 1.1, (B, 1))))

    n_lm_pt = max(int(round(n_pt * (1 - distractor_frac))), 0)
    n_lm_ln = max(int(round(n_ln * (1 - distractor_frac))), 0)

    P = _sample_points(rng, c, B, n_lm_pt)
    code = rng.integers(0, 256, size=(B, n_lm_pt, 32), dtype=np.uint8)
    _fill_points(rng, c, prev, prev.pt_l, prev.pt_r, prev.pdesc_l, prev.pdesc_r, P, code, n_pt, n_lm_pt, noise_px)
    Pc = np.einsum("bij,bnj->bni", DT_true[:, :3, :3], P) + DT_true[:, None, :3, 3]
    _fill_points(rng, c, curr, curr.pt_l, curr.pt_r, curr.pdesc_l, curr.pdesc_r, Pc, code, n_pt, n_lm_pt, noise_px)

    Ps = _sample_points(rng, c, B, n_lm_ln)
    dirs = rng.normal(size=(B, n_lm_ln, 3))
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    Pe = Ps + dirs * rng.uniform(0.3, 3.0, (B, n_lm_ln, 1))
    Pe[..., 2] = np.maximum(Pe[..., 2], 1.0)
    lcode = rng.integers(0, 256, size=(B, n_lm_ln, 32), dtype=np.uint8)
    _fill_lines(rng, c, prev.ln_l, prev.ln_r, prev.ldesc_l, prev.ldesc_r, Ps, Pe, lcode, n_ln, n_lm_ln, noise_px)
    R, t = DT_true[:, :3, :3], DT_true[:, None, :3, 3]
    _fill_lines(rng, c, curr.ln_l, curr.ln_r, curr.ldesc_l, curr.ldesc_r, np.einsum("bij,bnj->bni", R, Ps) + t,
                np.einsum("bij,bnj->bni", R, Pe) + t, lcode, n_ln, n_lm_ln, noise_px)

    for fr in (prev, curr):
        if ragged:
            fr.n_pt_l[:] = rng.integers(int(0.6 * n_pt), n_pt + 1, B)
            fr.n_pt_r[:] = rng.integers(int(0.6 * n_pt), n_pt + 1, B)
            fr.n_ln_l[:] = rng.integers(int(0.6 * n_ln), n_ln + 1, B)
            fr.n_ln_r[:] = rng.integers(int(0.6 * n_ln), n_ln + 1, B)
        else:
            fr.n_pt_l[:] = n_pt
            fr.n_pt_r[:] = n_pt
            fr.n_ln_l[:] = n_ln
            fr.n_ln_r[:] = n_ln

    state = S._alloc(B, S.frame_state_dt, pinned)
    Tfw = _expmap(np.concatenate([rng.uniform(-5, 5, (B, 3)), rng.uniform(-0.5, 0.5, (B, 3))], axis=1))
    state["Tfw"] = Tfw.reshape(B, 16)
    state["DT"] = DT_prev.reshape(B, 16)
    state["DT_cov"] = (np.eye(6) * 1e-4).reshape(36)
    state["Tfw_cov"] = (np.eye(6) * 1e-3).reshape(36)
    state["time_stamp"] = 0.0
    state["time_stamp_curr"] = 0.05
    return dict(cam=camera(shape), prev_raw=prev, curr_raw=curr, state=state, DT_true=DT_true)


def make_sequences(B, T, n_pt=300, n_ln=100, shape="euroc", seed=0, distractor_frac=0.1, noise_px=0.5, pinned=False):
    """B independent sequences of T + 1 stereo frames (sequence-chained replay: frame k is the previous frame of step k).
    One scene per sequence, near-constant velocity: the increment of step k is the sequence's base twist scaled by
    U(0.95, 1.05).  Returns dict(cam, frames [T + 1] RawFramesNP, state (of frame 0), DT_true [T][B,4,4], stamps [T + 1][B])."""
    c = EUROC if shape == "euroc" else KITTI
    rng = np.random.Generator(np.random.PCG64(seed))
    tw = np.zeros((B, 6))
    tdir = rng.normal(size=(B, 3))
    tw[:, :3] = tdir / np.linalg.norm(tdir, axis=1, keepdims=True) * rng.uniform(0, 0.1, (B, 1))
    wdir = rng.normal(size=(B, 3))
    tw[:, 3:] = wdir / np.linalg.norm(wdir, axis=1, keepdims=True) * rng.uniform(0, 0.05, (B, 1))
    n_lm_pt = max(int(round(n_pt * (1 - distractor_frac))), 0)
    n_lm_ln = max(int(round(n_ln * (1 - distractor_frac))), 0)
    P = _sample_points(rng, c, B, n_lm_pt)
    code = rng.integers(0, 256, size=(B, n_lm_pt, 32), dtype=np.uint8)
    Ps = _sample_points(rng, c, B, n_lm_ln)
    dirs = rng.normal(size=(B, n_lm_ln, 3))
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    Pe = Ps + dirs * rng.uniform(0.3, 3.0, (B, n_lm_ln, 1))
    Pe[..., 2] = np.maximum(Pe[..., 2], 1.0)
    lcode = rng.integers(0, 256, size=(B, n_lm_ln, 32), dtype=np.uint8)
    frames, DT_true = [], []
    Tk = np.tile(np.eye(4), (B, 1, 1))            # frame-0 coordinates -> frame-k coordinates
    for k in range(T + 1):
        fr = S.RawFramesNP(B, n_pt, n_ln, pinned)
        R, t = Tk[:, :3, :3], Tk[:, None, :3, 3]
        _fill_points(rng, c, fr, fr.pt_l, fr.pt_r, fr.pdesc_l, fr.pdesc_r, np.einsum("bij,bnj->bni", R, P) + t, code, n_pt, n_lm_pt, noise_px)
        _fill_lines(rng, c, fr.ln_l, fr.ln_r, fr.ldesc_l, fr.ldesc_r, np.einsum("bij,bnj->bni", R, Ps) + t,
                    np.einsum("bij,bnj->bni", R, Pe) + t, lcode, n_ln, n_lm_ln, noise_px)
        fr.n_pt_l[:] = n_pt; fr.n_pt_r[:] = n_pt; fr.n_ln_l[:] = n_ln; fr.n_ln_r[:] = n_ln
        frames.append(fr)
        if k < T:
            inc = _expmap(tw * rng.uniform(0.95, 1.05, (B, 1)))
            DT_true.append(inc)
            Tk = inc @ Tk
    state = S._alloc(B, S.frame_state_dt, pinned)
    state["Tfw"] = _expmap(np.concatenate([rng.uniform(-5, 5, (B, 3)), rng.uniform(-0.5, 0.5, (B, 3))], axis=1)).reshape(B, 16)
    state["DT"] = _inv(_expmap(tw * rng.uniform(0.9, 1.1, (B, 1)))).reshape(B, 16)
    state["DT_cov"] = (np.eye(6) * 1e-4).reshape(36)
    state["Tfw_cov"] = (np.eye(6) * 1e-3).reshape(36)
    state["time_stamp"] = 0.0
    state["time_stamp_curr"] = 0.05
    stamps = [np.full(B, 0.05 * k) for k in range(T + 2)]
    return dict(cam=camera(shape), frames=frames, state=state, DT_true=DT_true, stamps=stamps)


def tie_stress_descriptors(rng, n):
    """descriptors whose bytes 2..31 are zero (SURVEY §8d tie-stress set): most queries have d0 == d1."""
    d = np.zeros((n, 32), np.uint8)
    d[:, :2] = rng.integers(0, 256, size=(n, 2), dtype=np.uint8)
    return d


def make_lba_window(seed=0, shape="kitti", n_kf=10, n_pt=1000, n_ls=300, pt_obs_per=4, ls_obs_per=3, n_fix=2):
    """One local-BA window as flat arrays (SURVEY §8d LBA row): local KFs 0..n_kf-1, `n_fix` non-local
    observing KFs, landmarks observed by random subsets of the KFs, noisy observations."""
    c = EUROC if shape == "euroc" else KITTI
    rng = np.random.Generator(np.random.PCG64(seed))
    n_all = n_kf + n_fix
    tw = np.zeros((n_all, 6))
    tw[:, 2] = np.arange(n_all) * 0.8                      # forward motion
    tw[:, :3] += rng.normal(0, 0.05, (n_all, 3))
    tw[:, 3:] = rng.normal(0, 0.02, (n_all, 3))
    Tkw_all = _expmap(tw)                                  # T_kf_w: camera -> world
    Twk = _inv(Tkw_all)
    mid = Tkw_all[n_all // 2]

    def world_pts(n):
        Pc = _sample_points(rng, c, 1, n)[0]
        return Pc @ mid[:3, :3].T + mid[:3, 3]

    def observe(X, per):
        lm, kf, fix, obs = [], [], [], []
        for j in range(len(X)):
            ks = np.sort(rng.choice(n_all, size=min(per, n_all), replace=False))
            for k in ks:
                lm.append(j)
                if k < n_kf:
                    kf.append(k); fix.append(0)
                else:
                    kf.append(-1); fix.append(k - n_kf)
                obs.append(k)
        return np.array(lm, np.int32), np.array(kf, np.int32), np.array(fix, np.int32), np.array(obs, np.int64)

    X_pt = world_pts(n_pt)
    lm, kf, fix, k_all = observe(X_pt, pt_obs_per)
    Pc = np.einsum("nij,nj->ni", Twk[k_all][:, :3, :3], X_pt[lm]) + Twk[k_all][:, :3, 3]
    u, v, _ = _project(c, Pc)
    pt_uv = np.stack([u, v], 1) + rng.normal(0, 0.5, (len(lm), 2))

    Ls = world_pts(n_ls)
    dirs = rng.normal(size=(n_ls, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    X_ls = np.concatenate([Ls, Ls + dirs * rng.uniform(0.5, 3.0, (n_ls, 1))], axis=1)
    llm, lkf, lfix, lk_all = observe(X_ls, ls_obs_per)
    Pp = np.einsum("nij,nj->ni", Twk[lk_all][:, :3, :3], X_ls[llm, :3]) + Twk[lk_all][:, :3, 3]
    Pq = np.einsum("nij,nj->ni", Twk[lk_all][:, :3, :3], X_ls[llm, 3:]) + Twk[lk_all][:, :3, 3]
    up, vp, _ = _project(c, Pp)
    uq, vq, _ = _project(c, Pq)
    sp = np.stack([up, vp, np.ones_like(up)], 1) + np.pad(rng.normal(0, 0.5, (len(llm), 2)), ((0, 0), (0, 1)))
    ep = np.stack([uq, vq, np.ones_like(uq)], 1) + np.pad(rng.normal(0, 0.5, (len(llm), 2)), ((0, 0), (0, 1)))
    le = np.cross(sp, ep)
    le /= np.linalg.norm(le[:, :2], axis=1, keepdims=True)
    # perturb the state away from the truth so residuals / Jacobians are non-trivial
    X_pt_n = X_pt + rng.normal(0, 0.05, X_pt.shape)
    X_ls_n = X_ls + rng.normal(0, 0.05, X_ls.shape)
    return dict(kf_Tkw=Tkw_all[:n_kf].reshape(n_kf, 16), X_pt=X_pt_n, X_ls=X_ls_n, pt_obs_lm=lm, pt_obs_kf=kf,
                pt_obs_fix=fix, pt_obs_uv=pt_uv, pt_obs_s2=np.ones(len(lm)), ls_obs_lm=llm, ls_obs_kf=lkf,
                ls_obs_fix=lfix, ls_obs_le=le, ls_obs_s2=np.ones(len(llm)),
                fix_Tkw=Tkw_all[n_kf:].reshape(n_fix, 16), cam=camera(shape))


def make_right_pyramids(B, cam, n_levels=4, scale_factor=1.2, seed=0):
    """Synthetic right-image pyramids (CV_8UC1) for the ORB-SLAM style stereo refinement: smooth random texture per
    level, sizes as ORBextractor::ComputePyramid (ORBextractor.cc:1107-1112, float scale tables)."""
    rng = np.random.default_rng(seed)
    sf = [np.float32(1.0)]
    for _ in range(1, n_levels):
        sf.append(np.float32(sf[-1] * np.float32(scale_factor)))
    isf = [np.float32(1.0) / s for s in sf]
    sizes = [(int(np.rint(np.float32(cam.width) * s)), int(np.rint(np.float32(cam.height) * s))) for s in isf]
    pyr = S.PyramidNP(B, sizes)
    for p in range(B):
        for l, (w, h) in enumerate(sizes):
            coarse = rng.integers(0, 256, (h // 4 + 2, w // 4 + 2)).astype(np.float32)
            img = np.kron(coarse, np.ones((4, 4), np.float32))[:h, :w]
            img = 0.25 * (img + np.roll(img, 1, 0) + np.roll(img, 1, 1) + np.roll(img, (1, 1), (0, 1)))
            img += rng.integers(-6, 7, (h, w))
            pyr.level(p, l)[:] = np.clip(img, 0, 255).astype(np.uint8)
    return pyr
