"""GPU parity: every stage of the CUDA path, called through the C ABI (libgfpl.so via ctypes, host buffers),
against the CPU oracle on the same seeded inputs and against the committed cv2 golden vectors.

Bars (BASELINE.json north_star): bit-exact match indices, distances, accepted-feature lists, inlier masks
and cut ratios; fp64 outputs (H, g, e, poses, covariances, information matrices) within RTOL = 1e-9
relative to the magnitude of the quantity.
"""
import glob
import os

import numpy as np
import pytest

from gfpl import _structs as S
from gfpl import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-9
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bfmatcher_*.npz")))


def close(a, b, rtol=RTOL):
    """relative to the largest magnitude of the reference quantity (matrix-level tolerance)"""
    a, b = np.asarray(a), np.asarray(b)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-300)
    return bool(np.all(np.abs(a - b) <= rtol * scale) and np.array_equal(np.isnan(a), np.isnan(b)))


def assert_records_equal(a, b, n=None, exact_fields=None, tol_fields=()):
    names = a.dtype.names
    for f in names:
        x, y = (a[f], b[f]) if n is None else (a[f][:n], b[f][:n])
        if f in tol_fields:
            assert close(x, y), f
        else:
            assert np.array_equal(x, y), f


# ---------------------------------------------------------------------------------------------------------
# M1-M3
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[10:-4] for p in GOLDEN])
@pytest.mark.parametrize("norm", [6, 7])
def test_knn2_radius_vs_cv2_golden(ctx, path, norm):
    g = np.load(path)
    idx, dist = ctx.knn2(g["q"], g["t"], norm)
    assert np.array_equal(idx, g[f"knn_idx_{norm}"]) and np.array_equal(dist, g[f"knn_dist_{norm}"])
    rc, off, ri, rd = ctx.radius(g["q"], g["t"], float(g["radius"]), norm)
    assert rc == 0 and np.array_equal(off, g[f"rad_off_{norm}"])
    assert np.array_equal(ri, g[f"rad_idx_{norm}"]) and np.array_equal(rd, g[f"rad_dist_{norm}"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[10:-4] for p in GOLDEN])
@pytest.mark.parametrize("norm", [6, 7])
def test_knn2_mutual_vs_cv2_golden(ctx, path, norm):
    """the cv2.BFMatcher vectors through the both-directions entry point, i.e. through the kernels the tracking step
    runs: knn2_tc_kernel (tcgen05, >= 128 rows per side, NORM_HAMMING), knn2_tc_small_kernel (<= 128 rows, both norms),
    knn2_sym_kernel (NORM_HAMMING2 above 128 rows, tiny sides).  Forward = knnMatch(q, t, 2), reverse best =
    knnMatch(t, q, 2)[j][0] (stereoFrame.cpp:807-808)."""
    g = np.load(path)
    q, t = g["q"], g["t"]
    # the same problem twice, the second copy padded to a larger stride, plus an empty problem in between
    cap_q, cap_t = len(q) + 5, len(t) + 9
    Q = np.zeros((3, cap_q, 32), np.uint8); T = np.zeros((3, cap_t, 32), np.uint8)
    Q[0, :len(q)] = q; T[0, :len(t)] = t; Q[2, :len(q)] = q; T[2, :len(t)] = t
    nq = np.array([len(q), 0, len(q)], np.int32); nt = np.array([len(t), 0, len(t)], np.int32)
    idx, dist, bi, bd = ctx.knn2_mutual(Q, nq, T, nt, norm)
    for p in (0, 2):
        assert np.array_equal(idx[p, :len(q)], g[f"knn_idx_{norm}"]) and np.array_equal(dist[p, :len(q)], g[f"knn_dist_{norm}"])
        assert np.array_equal(bi[p, :len(t)], g[f"rknn_idx_{norm}"][:, 0])
        assert np.array_equal(bd[p, :len(t)], g[f"rknn_dist_{norm}"][:, 0])
    # exact strides as well (cap == n: the shapes the kernels' dispatch rules see in the step)
    idx, dist, bi, bd = ctx.knn2_mutual(q[None], [len(q)], t[None], [len(t)], norm)
    assert np.array_equal(idx[0], g[f"knn_idx_{norm}"]) and np.array_equal(dist[0], g[f"knn_dist_{norm}"])
    assert np.array_equal(bi[0], g[f"rknn_idx_{norm}"][:, 0]) and np.array_equal(bd[0], g[f"rknn_dist_{norm}"][:, 0])


@pytest.mark.parametrize("shape", [(300, 300), (100, 100), (2000, 2000), (500, 500), (1, 2), (37, 1025), (513, 255),
                                   (257, 256), (31, 33)])
@pytest.mark.parametrize("norm", [6, 7])
@pytest.mark.parametrize("tie", [False, True])
def test_knn2_vs_oracle(ctx, orc, shape, norm, tie):
    rng = np.random.default_rng(hash((shape, norm, tie)) % 2**32)
    nq, nt = shape
    mk = (lambda n: synth.tie_stress_descriptors(rng, n)) if tie else (lambda n: rng.integers(0, 256, (n, 32), dtype=np.uint8))
    q, t = mk(nq), mk(nt)
    gi, gd = ctx.knn2(q, t, norm)
    oi, od = orc.knn2(q, t, norm)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_knn2_batch_ragged(ctx, orc):
    rng = np.random.default_rng(7)
    B, cq, ct = 9, 130, 90
    q = rng.integers(0, 256, (B, cq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (B, ct, 32), dtype=np.uint8)
    nq = rng.integers(0, cq + 1, B).astype(np.int32)
    nt = rng.integers(2, ct + 1, B).astype(np.int32)
    nq[0], nq[1], nt[2] = 0, cq, 1          # empty query set, full, single train row
    idx, dist = ctx.knn2_batch(q, nq, t, nt)
    for p in range(B):
        if nt[p] >= 2 and nq[p] > 0:
            oi, od = orc.knn2(q[p, :nq[p]], t[p, :nt[p]])
            assert np.array_equal(idx[p, :nq[p]], oi) and np.array_equal(dist[p, :nq[p]], od)
    assert (idx[2, :nq[2], 1] == -1).all() and np.isinf(dist[2, :nq[2], 1]).all()   # nt == 1: no second neighbour
    assert (idx[2, :nq[2], 0] == 0).all()


def test_knn2_errors(ctx):
    from gfpl.api import GfplError
    q = np.zeros((4, 32), np.uint8)
    with pytest.raises(GfplError) as e:
        ctx.knn2(q, q[:1])                  # Nt >= 2
    assert e.value.code == S.E_INVALID
    with pytest.raises(GfplError):
        ctx.knn2(q, q, norm=4)
    idx, dist = ctx.knn2(q[:0], q)          # empty query set is fine
    assert idx.shape == (0, 2)


def test_radius_capacity_and_empty(ctx, orc):
    rng = np.random.default_rng(8)
    q = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (60, 32), dtype=np.uint8)
    rc, off, _, _ = ctx.radius(q, t, 300.0, cap=100)
    assert rc == S.E_CAPACITY and off[-1] == 3000
    rc, off, ri, rd = ctx.radius(q, t, 20.0)            # nothing within 20 bits of random rows
    assert rc == 0 and off[-1] == 0 and len(ri) == 0
    rc, off, ri, rd = ctx.radius(q, t, 128.0)
    rc2, off2, ri2, rd2 = orc.radius(q, t, 128.0)
    assert np.array_equal(off, off2) and np.array_equal(ri, ri2) and np.array_equal(rd, rd2)


def test_knn2_full_size_properties(ctx):
    """C4-sized problem: size-independent properties (self-match, symmetry of the distance, sortedness)."""
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (2000, 32), dtype=np.uint8)
    perm = rng.permutation(2000)
    idx, dist = ctx.knn2(a, a[perm])
    assert np.array_equal(perm[idx[:, 0]], np.arange(2000)) and (dist[:, 0] == 0).all()   # every row finds itself
    assert (dist[:, 0] <= dist[:, 1]).all()
    bits = np.unpackbits(a, axis=1).astype(np.int32)
    j = idx[:, 1]
    d1 = (bits != bits[perm][j]).sum(1)
    assert np.array_equal(d1.astype(np.float32), dist[:, 1])


# ---------------------------------------------------------------------------------------------------------
# M4-M8, C1
# ---------------------------------------------------------------------------------------------------------
def _data(B, seed, **kw):
    return synth.make_pairs(B, seed=seed, **kw)


def _cmp_stereo(g, o):
    assert np.array_equal(g.n_pt, o.n_pt) and np.array_equal(g.n_ln, o.n_ln)
    for p in range(g.B):
        assert_records_equal(g.pt[p], o.pt[p], n=g.n_pt[p])                      # bit-exact incl. fp64 back-projection
        assert np.array_equal(g.pdesc[p, :g.n_pt[p]], o.pdesc[p, :o.n_pt[p]])
        assert_records_equal(g.ln[p], o.ln[p], n=g.n_ln[p])
        assert np.array_equal(g.ldesc[p, :g.n_ln[p]], o.ldesc[p, :o.n_ln[p]])


@pytest.mark.parametrize("variant", [S.STEREO_BRUTE, S.STEREO_INITIAL])
@pytest.mark.parametrize("ragged", [False, True])
def test_stereo_match(ctx, orc, variant, ragged):
    d = _data(12, 100 + variant, ragged=ragged)
    g = ctx.stereo_match(d["curr_raw"], variant)
    o = orc.stereo_match(d["cam"], ctx.cfg, d["curr_raw"], variant)
    assert (o.n_pt > 100).all() and (o.n_ln > 3).all()
    _cmp_stereo(g, o)


def test_stereo_match_dense_and_degenerate(ctx, orc):
    d = _data(2, 7, n_pt=2000, n_ln=500)                # C4 sizes
    _cmp_stereo(ctx.stereo_match(d["curr_raw"]), orc.stereo_match(d["cam"], ctx.cfg, d["curr_raw"]))
    d = _data(4, 8, n_pt=64, n_ln=24)
    raw = d["curr_raw"]
    raw.n_pt_l[0] = 0; raw.n_pt_r[1] = 1; raw.n_ln_l[2] = 1; raw.n_ln_r[3] = 0    # empty / single-row sets (D3)
    raw.pdesc_r[2, :32] = raw.pdesc_l[2, :32]                                       # zero distances
    raw.pdesc_r[2, 32:64] = raw.pdesc_l[2, :32]                                     # duplicates: d0 == d1 == 0 -> NaN ratio
    g, o = ctx.stereo_match(raw), orc.stereo_match(d["cam"], ctx.cfg, raw)
    _cmp_stereo(g, o)
    assert g.n_pt[0] == 0 and g.n_pt[1] == 0 and g.n_ln[2] == 0 and g.n_ln[3] == 0


def test_stereo_line_covariance_gate_sweep(orc):
    """the largest-eigenvalue gate of the stereo lines (stereoFrame.cpp:965-1000) is decided by bounds and only falls
    back to the Jacobi iteration inside them: sweep the threshold through the eigenvalues of a batch so that both
    branches and their borders are exercised; accepted line lists must stay identical to the oracle's"""
    from gfpl import api
    d = _data(6, 131)
    seen = set()
    for th in np.geomspace(0.05, 400.0, 60):
        cfg = api.default_config()
        cfg.line_cov_th = float(th)
        c2 = api.Context(0, cam=d["cam"], cfg=cfg)
        g, o = c2.stereo_match(d["curr_raw"]), orc.stereo_match(d["cam"], cfg, d["curr_raw"])
        assert np.array_equal(g.n_ln, o.n_ln), th
        for p in range(6):
            assert np.array_equal(g.ln[p, :g.n_ln[p]]["src_l"], o.ln[p, :o.n_ln[p]]["src_l"]), (th, p)
        seen.add(int(o.n_ln.sum()))
    assert len(seen) > 10          # the threshold really cuts through the batch
    # thresholds INSIDE the bounds of chosen end points (largest eigenvalue in [2 k |v|^2, k (2 |v|^2 + d^2)]): the
    # Jacobi fallback decides those lines
    cam = d["cam"]
    base = orc.stereo_match(cam, api.default_config(), d["curr_raw"])
    for p in range(3):
        for i in range(0, int(base.n_ln[p]), 7):
            ln = base.ln[p, i]
            x, y, dd = ln["spl"][0] - cam.cx, ln["spl"][1] - cam.cy, ln["sdisp"]
            k = cam.b * cam.b / dd ** 4
            lo = 2.0 * k * (x * x + y * y + cam.fx * cam.fx)
            cfg = api.default_config()
            cfg.line_cov_th = float(lo + 0.5 * k * dd * dd)
            c2 = api.Context(0, cam=cam, cfg=cfg)
            g, o = c2.stereo_match(d["curr_raw"]), orc.stereo_match(cam, cfg, d["curr_raw"])
            assert np.array_equal(g.n_ln, o.n_ln)
            assert np.array_equal(g.ln[p, :g.n_ln[p]]["src_l"], o.ln[p, :o.n_ln[p]]["src_l"])


def test_stereo_uncertainty(ctx, orc):
    d = _data(6, 11)
    fr = orc.stereo_match(d["cam"], ctx.cfg, d["prev_raw"])
    g, o = ctx.stereo_uncertainty(fr), orc.stereo_uncertainty(d["cam"], ctx.cfg, fr)
    for p in range(6):
        n = fr.n_ln[p]
        assert np.array_equal(g["covS"][p, :n], o["covS"][p, :n]) and np.array_equal(g["covE"][p, :n], o["covE"][p, :n])


def _cmp_obs(g, o, exact=True):
    assert np.array_equal(g.n_pt, o.n_pt) and np.array_equal(g.n_ln, o.n_ln)
    for p in range(g.B):
        assert_records_equal(g.pt[p], o.pt[p], n=g.n_pt[p])
        assert_records_equal(g.ln[p], o.ln[p], n=g.n_ln[p])
        assert np.array_equal(g.pt_inlier[p, :g.n_pt[p]], o.pt_inlier[p, :o.n_pt[p]])
        assert np.array_equal(g.ln_inlier[p, :g.n_ln[p]], o.ln_inlier[p, :o.n_ln[p]])


def _frames(orc, cfg, d):
    return orc.stereo_match(d["cam"], cfg, d["prev_raw"]), orc.stereo_match(d["cam"], cfg, d["curr_raw"])


@pytest.mark.parametrize("mode", [S.F2F_HYBRID, S.F2F_BRUTE])
@pytest.mark.parametrize("ragged", [False, True])
def test_f2f_match(ctx, orc, mode, ragged):
    d = _data(10, 200 + mode, ragged=ragged)
    prev, curr = _frames(orc, ctx.cfg, d)
    g = ctx.f2f_match(prev, curr, d["state"], mode)
    o = orc.f2f_match(d["cam"], ctx.cfg, prev, curr, d["state"], mode)
    assert (o.n_pt > 50).all() and (o.n_ln > 2).all()
    _cmp_obs(g, o)


def test_f2f_hybrid_knn_candidates_and_caps(ctx, orc):
    from gfpl import api
    d = _data(6, 300)
    cfg = api.default_config()
    cfg.lr_in_parallel = 0                   # knn k=2 candidates instead of radius lists (:484-493)
    cfg.max_point_match_num = 40             # force the early exit (:595)
    cfg.max_line_match_num = 5
    c2 = api.Context(0, cam=d["cam"], cfg=cfg)
    prev, curr = _frames(orc, cfg, d)
    g = c2.f2f_match(prev, curr, d["state"])
    o = orc.f2f_match(d["cam"], cfg, prev, curr, d["state"])
    assert (o.n_pt == 40).all() and (o.n_ln == 5).all()
    _cmp_obs(g, o)


def test_f2f_hybrid_duplicates_tie_stress(ctx, orc):
    """low-entropy descriptors + zero motion: many candidates per previous point, previous points winning
    several current keys and ties resolved to the lowest previous index (D2)."""
    d = _data(3, 400, n_pt=200, n_ln=40)
    rng = np.random.default_rng(4)
    prev, curr = _frames(orc, ctx.cfg, d)
    for fr in (prev, curr):
        fr.pdesc[...] = 0
        fr.pdesc[..., :1] = rng.integers(0, 4, fr.pdesc[..., :1].shape, dtype=np.uint8)
    curr.pt["pl"][:, 1::2] = curr.pt["pl"][:, 0::2] + 3.0          # clusters of current points within the 10 px window
    g = ctx.f2f_match(prev, curr, d["state"])
    o = orc.f2f_match(d["cam"], ctx.cfg, prev, curr, d["state"])
    dup = max(np.bincount(o.pt[p, :o.n_pt[p]]["prev_idx"]).max() for p in range(3) if o.n_pt[p])
    assert dup >= 2
    _cmp_obs(g, o)


# ---------------------------------------------------------------------------------------------------------
# C2-C5, P1-P4
# ---------------------------------------------------------------------------------------------------------
def _obs(orc, cfg, d):
    prev, curr = _frames(orc, cfg, d)
    return prev, curr, orc.f2f_match(d["cam"], cfg, prev, curr, d["state"])


def _dt_inv(state):
    Tc = state["Tfw"].reshape(-1, 4, 4) @ state["DT"].reshape(-1, 4, 4)
    return (np.linalg.inv(Tc) @ state["Tfw"].reshape(-1, 4, 4)).reshape(-1, 16)


@pytest.mark.parametrize("max_vol", [1, 0])
def test_linecut(ctx, orc, max_vol):
    from gfpl import api
    d = _data(8, 500 + max_vol)
    cfg = api.default_config()
    cfg.max_vol_line_cut = max_vol
    c2 = api.Context(0, cam=d["cam"], cfg=cfg)
    prev, curr, obs = _obs(orc, cfg, d)
    cov = orc.stereo_uncertainty(d["cam"], cfg, prev)
    DT_inv = _dt_inv(d["state"])
    og, oo = obs.copy(), obs.copy()
    cg = c2.linecut(DT_inv, og, cov)
    co = orc.linecut(d["cam"], cfg, DT_inv, oo, cov)
    flips = 0
    for p in range(8):
        n = obs.n_ln[p]
        same = np.array_equal(cg["cutRatio"][p, :n], co["cutRatio"][p, :n])
        flips += not same
        if True:
            assert same, f"cut decisions differ in pair {p}"                     # exact cut ratios
            assert close(cg["invCovPose"][p, :n], co["invCovPose"][p, :n])
            for f in ("spl", "epl", "sdisp", "edisp"):
                assert close(cg[f][p, :n], co[f][p, :n]), f
            assert close(og.ln[p, :n]["sP"], oo.ln[p, :n]["sP"]) and close(og.ln[p, :n]["eP"], oo.ln[p, :n]["eP"])
    # the min-eigenvalue metric (non-default) runs the oracle's Jacobi iteration operation for operation (-fmad=false):
    # its decisions replay exactly as well (64 pairs / 1291 lines checked by tools/mineig_diag.py: no differing line,
    # every information matrix bit-identical)
    assert flips == 0, f"{flips} pairs with differing cut decisions"


@pytest.mark.parametrize("step,rng", [(0.1, (0.0, 1.0)), (0.02, (0.0, 1.0)), (0.3, (0.0, 1.0)), (0.05, (0.0, 0.35))])
def test_linecut_other_steps_and_ranges(orc, step, rng):
    """the half cache of the screening path is indexed by cut step: other step sizes (2..50 steps per axis) and a
    range that clips the search must leave every decision as the oracle makes it"""
    from gfpl import api
    d = _data(6, 520)
    cfg = api.default_config()
    cfg.cut_step = step
    cfg.cut_rng[0], cfg.cut_rng[1] = rng
    c2 = api.Context(0, cam=d["cam"], cfg=cfg)
    prev, curr, obs = _obs(orc, cfg, d)
    cov = orc.stereo_uncertainty(d["cam"], cfg, prev)
    DT_inv = _dt_inv(d["state"])
    og, oo = obs.copy(), obs.copy()
    cg = c2.linecut(DT_inv, og, cov)
    co = orc.linecut(d["cam"], cfg, DT_inv, oo, cov)
    moved = 0
    for p in range(6):
        n = obs.n_ln[p]
        assert np.array_equal(cg["cutRatio"][p, :n], co["cutRatio"][p, :n]), f"cut decisions differ in pair {p}"
        assert close(cg["invCovPose"][p, :n], co["invCovPose"][p, :n])
        assert close(og.ln[p, :n]["sP"], oo.ln[p, :n]["sP"]) and close(og.ln[p, :n]["eP"], oo.ln[p, :n]["eP"])
        assert (co["cutRatio"][p, :n] <= rng[1] + 1e-12).all()
        moved += int((co["cutRatio"][p, :n] > 0).any())
    assert moved > 0


@pytest.mark.parametrize("step,rng,B", [(0.05, (0.0, 1.0), 48), (0.1, (0.0, 1.0), 6), (0.02, (0.0, 0.35), 6), (0.05, (0.0, 1.0), 512)])
def test_linecut_descent(orc, step, rng, B):
    """C6, estimateProjUncertainty_descent: one thread per line, the oracle's operations in the oracle's order
    (-fmad=false, the same Jacobi iteration): cut decisions, volume tracks and every rewritten field bit-identical"""
    from gfpl import api
    d = _data(B, 530, ragged=True)
    cfg = api.default_config()
    cfg.cut_step = step
    cfg.cut_rng[0], cfg.cut_rng[1] = rng
    c2 = api.Context(0, cam=d["cam"], cfg=cfg)
    prev, curr, obs = _obs(orc, cfg, d)
    cov = orc.stereo_uncertainty(d["cam"], cfg, prev)
    DT_inv = _dt_inv(d["state"])
    og, oo = obs.copy(), obs.copy()
    cg = c2.linecut_descent(DT_inv, og, cov)
    co = orc.linecut_descent(d["cam"], cfg, DT_inv, oo, cov)
    moved = lines = 0
    for p in range(B):
        n = obs.n_ln[p]
        assert np.array_equal(cg["cutRatio"][p, :n], co["cutRatio"][p, :n]), f"cut decisions differ in pair {p}"
        assert np.array_equal(cg["n_volume"][p, :n], co["n_volume"][p, :n])
        assert_records_equal(cg[p], co[p], n=n)
        assert_records_equal(og.ln[p], oo.ln[p], n=n)
        assert (cg[p, n:]["n_volume"] == 0).all()                     # rows past the count are not touched
        moved += int((co["cutRatio"][p, :n] > 0).any(1).sum()); lines += n
    assert moved > lines // 20 and lines > 0


def test_vector_stdv_mad_ties_and_edges(ctx, orc):
    """the selection of removeOutliers alone, against the oracle's two sorts (exact): random rows, rows drawn from a handful
    of values (the order statistics fall inside long runs of equal keys: more candidates than a lane group holds), rows
    whose values share the high word and differ in the low word, constant rows, n = 0..5, ragged counts"""
    rng = np.random.default_rng(12)
    cap = 301
    rows, counts = [], []

    def add(v):
        r = np.zeros(cap); r[:len(v)] = v
        rows.append(r); counts.append(len(v))
    for n in (0, 1, 2, 3, 4, 5, 16, 17, 64, 255, 300, 301):
        add(rng.random(n) * 3.0)
    for n in (40, 230, 301):
        add(rng.choice(rng.random(3) * 2.0, n))                             # three distinct values
        add(rng.choice(rng.random(12) * 2.0, n))                            # a dozen
        add(np.full(n, 0.731))                                              # constant: MAD = 0
        add(1.0 + rng.integers(0, 1 << 20, n) * 2.0 ** -52)                # equal high words, different low words
        add(np.concatenate([rng.random(n // 3), np.full(n - n // 3, 0.5)]))  # one long run across the median
        add(np.abs(rng.standard_normal(n)) * 10.0 ** rng.integers(-8, 3, n))   # wide dynamic range
        add(np.round(rng.random(n) * 8) / 8)                                # nine distinct values, exact in float
    for _ in range(40):
        add(np.abs(rng.standard_normal(int(rng.integers(1, cap + 1)))))
    res, n = np.array(rows), np.array(counts, np.int32)
    got = ctx.vector_stdv_mad(res, n)
    want = np.array([orc.vector_stdv_mad(res[i, :n[i]]) for i in range(len(n))])
    assert np.array_equal(got, want), np.nonzero(got != want)[0]
    assert (want == 0).sum() >= 4 and (want > 0).sum() > 50


@pytest.mark.parametrize("B", [1, 2, 3, 5, 7])
def test_pose_optimize_batch_tails(ctx, orc, B):
    """two frame pairs share a warp and four a CTA: batches that leave lane groups, warps and CTAs partly empty"""
    d = _data(B, 740 + B, ragged=True)
    prev, curr, obs = _obs(orc, ctx.cfg, d)
    og, oo = obs.copy(), obs.copy()
    ini = d["state"]["DT"].copy()
    pg = ctx.pose_optimize(og, d["state"], ini, S.POSE_SEED_GIVEN)
    po = orc.pose_optimize(d["cam"], ctx.cfg, oo, d["state"], ini, S.POSE_SEED_GIVEN)
    assert np.array_equal(pg["status"], po["status"])
    assert np.array_equal(og.pt_inlier, oo.pt_inlier) and np.array_equal(og.ln_inlier, oo.ln_inlier)
    _cmp_pose(pg, po, B)


def test_pose_hg(ctx, orc):
    d = _data(16, 600, ragged=True)
    prev, curr, obs = _obs(orc, ctx.cfg, d)
    obs.pt_inlier[:, ::7] = 0                                  # some outliers flagged
    DT = np.linalg.inv(d["state"]["DT"].reshape(-1, 4, 4)).reshape(-1, 16)
    Hg, gg, eg = ctx.pose_hg(DT, obs)
    Ho, go, eo = orc.pose_hg(d["cam"], ctx.cfg, DT, obs)
    for p in range(16):
        assert close(Hg[p], Ho[p]) and close(gg[p], go[p]) and close(eg[p], eo[p])


def _cmp_pose(pg, po, B):
    assert np.array_equal(pg["status"], po["status"])
    for f in ("n_inliers", "n_inliers_pt", "n_inliers_ls", "iters_first", "iters_ref"):
        assert np.array_equal(pg[f], po[f]), f
    for p in range(B):
        for f in ("DT", "Tfw", "DT_cov", "Tfw_cov", "err_norm"):
            assert close(pg[f][p], po[f][p]), (f, p)
        assert close(pg["DT_cov_eig"][p], po["DT_cov_eig"][p], rtol=1e-7), p      # eigenvalues: relative to the largest


@pytest.mark.parametrize("overload", [S.POSE_SEED_PREV_DT, S.POSE_SEED_GIVEN])
def test_pose_optimize(ctx, orc, overload):
    d = _data(24, 700 + overload, ragged=True)
    prev, curr, obs = _obs(orc, ctx.cfg, d)
    og, oo = obs.copy(), obs.copy()
    ini = d["state"]["DT"].copy() if overload == S.POSE_SEED_GIVEN else None
    pg = ctx.pose_optimize(og, d["state"], ini, overload)
    po = orc.pose_optimize(d["cam"], ctx.cfg, oo, d["state"], ini, overload)
    assert (po["status"] == 0).all()
    assert np.array_equal(og.pt_inlier, oo.pt_inlier) and np.array_equal(og.ln_inlier, oo.ln_inlier)   # exact masks
    _cmp_pose(pg, po, 24)


def test_pose_optimize_gates(ctx, orc):
    d = _data(6, 800)
    prev, curr, obs = _obs(orc, ctx.cfg, d)
    obs.n_pt[0] = 5; obs.n_ln[0] = 2                                   # too few features (:1953)
    obs.pt[1]["P"][:, 2] = 0.0                                         # division by zero depth -> non-finite (:1959)
    d["state"]["time_stamp_curr"][2] = d["state"]["time_stamp"][2]     # zero time step -> motion gate rejects (:1992)
    obs.n_pt[3] = 0                                                     # lines only
    og, oo = obs.copy(), obs.copy()
    pg = ctx.pose_optimize(og, d["state"], d["state"]["DT"].copy(), S.POSE_SEED_GIVEN)
    po = orc.pose_optimize(d["cam"], ctx.cfg, oo, d["state"], d["state"]["DT"].copy(), S.POSE_SEED_GIVEN)
    assert po["status"][0] & 1 and po["status"][2] & 16 and po["status"][1] != 0
    assert np.array_equal(pg["status"], po["status"])
    assert np.array_equal(og.pt_inlier, oo.pt_inlier) and np.array_equal(og.ln_inlier, oo.ln_inlier)
    for p in (0, 2, 3, 4, 5):
        for f in ("DT", "Tfw", "Tfw_cov", "err_norm", "DT_cov"):
            assert close(pg[f][p], po[f][p]), (f, p)
    assert pg["err_norm"][1] == -1.0 and np.array_equal(pg["DT"][1], np.eye(4).reshape(16))


# ---------------------------------------------------------------------------------------------------------
# whole step + resident replay
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [S.TRACK_LINECUT | S.TRACK_SEED_GIVEN, 0, S.TRACK_F2F_BRUTE | S.TRACK_STEREO_INITIAL])
def test_track_whole_step(ctx, orc, flags):
    d = _data(16, 900 + flags, ragged=True)
    prev = orc.stereo_match(d["cam"], ctx.cfg, d["prev_raw"])
    cg, og, kg, pg = ctx.track(prev, d["state"], d["curr_raw"], flags)
    co, oo, ko, po = orc.track(d["cam"], ctx.cfg, prev, d["state"], d["curr_raw"], flags)
    _cmp_stereo(cg, co)
    assert np.array_equal(og.n_pt, oo.n_pt) and np.array_equal(og.n_ln, oo.n_ln)
    for p in range(16):
        assert_records_equal(og.pt[p], oo.pt[p], n=og.n_pt[p])
        assert_records_equal(og.ln[p], oo.ln[p], n=og.n_ln[p], tol_fields=("sP", "eP"))
        assert np.array_equal(og.pt_inlier[p, :og.n_pt[p]], oo.pt_inlier[p, :oo.n_pt[p]])
        assert np.array_equal(og.ln_inlier[p, :og.n_ln[p]], oo.ln_inlier[p, :oo.n_ln[p]])
        if flags & S.TRACK_LINECUT:
            assert np.array_equal(kg["cutRatio"][p, :og.n_ln[p]], ko["cutRatio"][p, :oo.n_ln[p]])
    _cmp_pose(pg, po, 16)


def test_replay_matches_track_and_is_deterministic(ctx, orc):
    from gfpl import api
    d = _data(64, 1000)
    flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
    prev = orc.stereo_match(d["cam"], ctx.cfg, d["prev_raw"])
    _, og, kg, pg = ctx.track(prev, d["state"], d["curr_raw"], flags)
    rp = api.Replay(ctx, 64, d["curr_raw"].cap_pt, d["curr_raw"].cap_ln)
    rp.upload(prev, d["state"], d["curr_raw"])
    poses = []
    for _ in range(2):
        rp.run(flags)
        pose = np.zeros(64, S.pose_result_dt)
        obs = S.TrackObsNP(64, ctx.cfg.max_point_match_num, ctx.cfg.max_line_match_num)
        rp.download(pose, obs)
        ctx.sync()
        poses.append(pose.copy())
    assert poses[0].tobytes() == poses[1].tobytes()                    # run-to-run bit-reproducible
    assert poses[0].tobytes() == pg.tobytes()                          # resident path == host-pointer path
    assert np.array_equal(obs.n_pt, og.n_pt) and np.array_equal(obs.n_ln, og.n_ln)
    for p in range(64):   # slots past n_pt / n_ln are never written on either path
        assert np.array_equal(obs.pt_inlier[p, :obs.n_pt[p]], og.pt_inlier[p, :og.n_pt[p]])
        assert np.array_equal(obs.ln_inlier[p, :obs.n_ln[p]], og.ln_inlier[p, :og.n_ln[p]])
    assert rp.h2d_bytes() > 0 and rp.d2h_bytes() > 0
    ms = rp.stage_ms()
    assert len(ms) == S.NUM_STAGES and ms[-1] > 0
    assert ctx.launch_count() > 0


@pytest.mark.parametrize("n_sub", [2, 4, 7])
def test_replay_pipelined_is_bit_identical(ctx, orc, n_sub):
    """gfpl_replay_set_pipeline: matching of sub-batch k+1 on one stream, line cut + pose of sub-batch k on their own
    high-priority streams, consecutive runs overlapping — every record must equal the unpipelined step bit for bit
    (uneven split, ragged counts, three back-to-back runs, an upload in between)"""
    from gfpl import api
    B = 90
    d = _data(B, 1300 + n_sub, ragged=True)
    flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
    prev = orc.stereo_match(d["cam"], ctx.cfg, d["prev_raw"])
    rp = api.Replay(ctx, B, d["curr_raw"].cap_pt, d["curr_raw"].cap_ln)

    def fetch():
        pose = np.zeros(B, S.pose_result_dt)
        obs = S.TrackObsNP(B, ctx.cfg.max_point_match_num, ctx.cfg.max_line_match_num)
        cut = np.zeros((B, obs.cap_ln), S.line_cut_dt)
        curr = S.StereoFramesNP(B, d["curr_raw"].cap_pt, d["curr_raw"].cap_ln)
        rp.download(pose, obs, cut, curr)
        ctx.sync()
        return pose, obs, cut, curr

    rp.upload(prev, d["state"], d["curr_raw"])
    rp.run(flags)
    ref = fetch()
    rp.set_pipeline(n_sub)
    for rounds in (3, 1):
        rp.upload(prev, d["state"], d["curr_raw"])
        for _ in range(rounds):
            rp.run(flags)                       # back to back: step n+1's matching overlaps step n's optimiser
        got = fetch()
        assert got[0].tobytes() == ref[0].tobytes()
        assert np.array_equal(got[1].n_pt, ref[1].n_pt) and np.array_equal(got[1].n_ln, ref[1].n_ln)
        for p in range(B):
            npt, nln = ref[1].n_pt[p], ref[1].n_ln[p]
            assert got[1].pt[p, :npt].tobytes() == ref[1].pt[p, :npt].tobytes()
            assert got[1].ln[p, :nln].tobytes() == ref[1].ln[p, :nln].tobytes()
            assert np.array_equal(got[1].pt_inlier[p, :npt], ref[1].pt_inlier[p, :npt])
            assert np.array_equal(got[1].ln_inlier[p, :nln], ref[1].ln_inlier[p, :nln])
            assert got[2][p, :nln].tobytes() == ref[2][p, :nln].tobytes()
        assert np.array_equal(got[3].n_pt, ref[3].n_pt) and np.array_equal(got[3].n_ln, ref[3].n_ln)
    with pytest.raises(api.GfplError):
        rp.stage_ms()                           # per-kernel events exist only for the unpipelined step
    with pytest.raises(api.GfplError):
        rp.set_pipeline(0)
    rp.set_pipeline(1)
    rp.run(flags)
    assert fetch()[0].tobytes() == ref[0].tobytes() and rp.stage_ms()[-1] > 0


def test_sequence_chained_replay_updateFrame(ctx, orc):
    """gfpl_replay_advance + gfpl_replay_upload_curr_compact (StereoFrameHandler::updateFrame: prev_frame <- curr_frame):
    4 steps over 24 sequences; every step's records equal the oracle driven the same way on the host (the oracle's
    current frame and pose become its next previous frame and state)"""
    from gfpl import api
    B, T = 24, 4
    sq = synth.make_sequences(B, T, seed=31)
    flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
    prev_o = orc.stereo_match(sq["cam"], ctx.cfg, sq["frames"][0])
    state_o = sq["state"].copy()
    rp = api.Replay(ctx, B, 300, 100)
    rp.upload(prev_o, sq["state"], sq["frames"][1])
    for k in range(1, T + 1):
        if k > 1:
            rp.advance(sq["stamps"][k])
            rp.upload_curr_compact(S.CompactInputNP(sq["frames"][k]))
        rp.run(flags)
        pose = np.zeros(B, S.pose_result_dt)
        obs = S.TrackObsNP(B, ctx.cfg.max_point_match_num, ctx.cfg.max_line_match_num)
        curr = S.StereoFramesNP(B, 300, 100)
        rp.download(pose, obs, None, curr)
        ctx.sync()
        co, oo, ko, po = orc.track(sq["cam"], ctx.cfg, prev_o, state_o, sq["frames"][k], flags)
        _cmp_stereo(curr, co)
        assert np.array_equal(obs.n_pt, oo.n_pt) and np.array_equal(obs.n_ln, oo.n_ln), k
        for p in range(B):
            assert np.array_equal(obs.pt[p, :oo.n_pt[p]]["prev_idx"], oo.pt[p, :oo.n_pt[p]]["prev_idx"]), (k, p)
            assert np.array_equal(obs.pt_inlier[p, :oo.n_pt[p]], oo.pt_inlier[p, :oo.n_pt[p]]), (k, p)
            assert np.array_equal(obs.ln_inlier[p, :oo.n_ln[p]], oo.ln_inlier[p, :oo.n_ln[p]]), (k, p)
        _cmp_pose(pose, po, B)
        assert (po["status"] == 0).mean() > 0.9 and (oo.n_pt > 80).all()
        # the tracked increment is the simulated one
        err = np.abs(np.linalg.inv(pose["DT"].reshape(B, 4, 4)) - sq["DT_true"][k - 1]).max(axis=(1, 2))
        assert np.median(err) < 1e-2
        # oracle side of updateFrame
        prev_o = co
        for f in ("Tfw", "DT", "DT_cov", "Tfw_cov"):
            state_o[f] = po[f]
        state_o["time_stamp"] = state_o["time_stamp_curr"]
        state_o["time_stamp_curr"] = sq["stamps"][k + 1]
    assert rp.h2d_bytes() < 0.75 * (S.CompactInputNP(sq["frames"][1], prev_o).nbytes())


def test_track_at_scale_no_flips(ctx, orc):
    """4096 frame pairs through the whole step against the oracle (all host threads): the count of pairs whose inlier
    masks, iteration counts, status words or cut decisions differ must be ZERO — k_pose.cu is built with FMA
    contraction and hardware reciprocal seeds, so a residual within rounding of a MAD threshold would show up here."""
    B = 4096
    d = _data(B, 2024)
    flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
    prev = orc.stereo_match(d["cam"], ctx.cfg, d["prev_raw"])
    co, oo, ko, po = orc.track(d["cam"], ctx.cfg, prev, d["state"], d["curr_raw"], flags, n_threads=os.cpu_count() or 1)
    flips = dict(mask_pt=0, mask_ln=0, iters=0, status=0, cut=0, counts=0)
    tol_bad = 0
    tol_cov = dict(Tfw=0, DT_cov=0, Tfw_cov=0, err_norm=0, DT_cov_eig=0)
    for lo in range(0, B, 1024):
        hi = lo + 1024
        sl = lambda a: a[lo:hi]
        cg, og, kg, pg = ctx.track(prev.slice(lo, hi), d["state"][lo:hi], d["curr_raw"].slice(lo, hi), flags)
        for p in range(1024):
            q = lo + p
            npt, nln = oo.n_pt[q], oo.n_ln[q]
            if og.n_pt[p] != npt or og.n_ln[p] != nln:
                flips["counts"] += 1
                continue
            flips["mask_pt"] += int(not np.array_equal(og.pt_inlier[p, :npt], oo.pt_inlier[q, :npt]))
            flips["mask_ln"] += int(not np.array_equal(og.ln_inlier[p, :nln], oo.ln_inlier[q, :nln]))
            flips["cut"] += int(not np.array_equal(kg["cutRatio"][p, :nln], ko["cutRatio"][q, :nln]))
            flips["iters"] += int(pg["iters_first"][p] != po["iters_first"][q] or pg["iters_ref"][p] != po["iters_ref"][q])
            flips["status"] += int(pg["status"][p] != po["status"][q])
            if pg["status"][p] == po["status"][q]:
                tol_bad += int(not close(pg["DT"][p], po["DT"][q]))
                for f in ("Tfw", "DT_cov", "Tfw_cov", "err_norm"):
                    tol_cov[f] += int(not close(pg[f][p], po[f][q]))
                tol_cov["DT_cov_eig"] += int(not close(pg["DT_cov_eig"][p], po["DT_cov_eig"][q], rtol=1e-7))
    assert (po["status"] == 0).mean() > 0.99
    assert sum(flips.values()) == 0, f"pairs that differ from the oracle, of {B}: {flips}"
    assert tol_bad == 0, f"{tol_bad} poses outside {RTOL}"
    # the finalize kernel (LU inverse, covariance propagation, Jacobi eigenvalues on reciprocal refinements) at scale
    assert sum(tol_cov.values()) == 0, f"pairs outside the tolerance, of {B}: {tol_cov}"


def test_count_validation_rejects_out_of_range(ctx):
    """a per-problem count outside [0, cap] or a side of 2^20 rows is GFPL_E_INVALID before anything is launched"""
    from gfpl import api
    d = _data(2, 5)
    raw = d["curr_raw"]
    good = raw.n_pt_l.copy()
    for bad in (-1, raw.cap_pt + 1):
        raw.n_pt_l[1] = bad
        with pytest.raises(api.GfplError) as e:
            ctx.stereo_match(raw)
        assert e.value.code == S.E_INVALID
    raw.n_pt_l[:] = good
    prev = ctx.stereo_match(raw)
    prev.n_ln[0] = prev.cap_ln + 3
    with pytest.raises(api.GfplError):
        ctx.track(prev, d["state"], raw, 0)
    q = np.zeros((1, 4, 32), np.uint8)
    with pytest.raises(api.GfplError):
        ctx.knn2_mutual(q, [5], q, [4])
    with pytest.raises(api.GfplError):
        ctx.knn2_batch(q, [4], q, [-2])
    ctx.stereo_match(raw)                        # the context is still usable


def test_linecut_rejects_bad_covariance_index(ctx, orc):
    """a matched line whose prev_idx lies outside the covariance array is GFPL_E_INVALID (both cut searches), not a device fault"""
    from gfpl import api
    d = _data(2, 531)
    prev, curr, obs = _obs(orc, ctx.cfg, d)
    cov = orc.stereo_uncertainty(d["cam"], ctx.cfg, prev)
    DT_inv = _dt_inv(d["state"])
    assert obs.n_ln[1] > 0
    for bad in (-1, cov.shape[1]):
        o = obs.copy()
        o.ln[1, 0]["prev_idx"] = bad
        for call in (ctx.linecut, ctx.linecut_descent):
            with pytest.raises(api.GfplError) as e:
                call(DT_inv, o, cov)
            assert e.value.code == S.E_INVALID
    ctx.linecut_descent(DT_inv, obs.copy(), cov)   # the context is still usable


def test_lba_assemble(ctx, orc):
    for seed, kw in ((1, dict(n_kf=10, n_pt=1000, n_ls=300)), (2, dict(n_kf=3, n_pt=20, n_ls=5, n_fix=1))):
        w = synth.make_lba_window(seed=seed, **kw)
        c2 = ctx
        c2.set_camera(w["cam"])
        g = c2.lba_assemble(w)
        o = orc.lba_assemble(w["cam"], ctx.cfg, w)
        for k in ("Hpk", "Hlk"):
            assert np.array_equal(g[k], o[k]), k                      # per-observation blocks: no reduction, bit-exact
        for k in ("Hkk", "Hpp", "Hll", "g", "err"):
            assert close(g[k], o[k]), k
    from gfpl import api
    ctx.set_camera(api.euroc_camera())


def test_measure_peaks(ctx):
    popc, fp64 = ctx.measure_peaks()
    assert 1.0 < popc < 10.0 and 5.0 < fp64 < 100.0


def test_replay_with_packed_previous_frames(ctx, orc):
    """tighter row strides for the previous frames (gfpl_replay_create_ex) do not change any result"""
    from gfpl import api
    d = _data(32, 1100, ragged=True)
    flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
    prev = orc.stereo_match(d["cam"], ctx.cfg, d["prev_raw"])
    _, _, _, pg = ctx.track(prev, d["state"], d["curr_raw"], flags)
    packed = prev.repack()
    assert packed.cap_pt < prev.cap_pt and packed.cap_ln < prev.cap_ln
    rp = api.Replay(ctx, 32, d["curr_raw"].cap_pt, d["curr_raw"].cap_ln, packed.cap_pt, packed.cap_ln)
    rp.upload(packed, d["state"], d["curr_raw"])
    rp.run(flags)
    pose = np.zeros(32, S.pose_result_dt)
    rp.download(pose)
    ctx.sync()
    assert pose.tobytes() == pg.tobytes()
    # slim previous-frame records (gfpl_replay_upload_slim) give the same poses as well
    rp.upload_slim(S.PrevFramesNP(packed), d["state"], d["curr_raw"])
    rp.run(flags)
    pose2 = np.zeros(32, S.pose_result_dt)
    rp.download(pose2)
    ctx.sync()
    assert pose2.tobytes() == pg.tobytes()
    # compact step input (gfpl_replay_upload_compact: 12 B key points, 24 B key lines, 16 B previous points whose P and
    # sigma2 are recomputed on the device): same poses, in fewer host-to-device bytes
    slim_bytes = rp.h2d_bytes()
    comp = S.CompactInputNP(d["curr_raw"], packed)
    rp.upload_compact(comp, d["state"])
    assert rp.h2d_bytes() < 0.9 * slim_bytes
    rp.run(flags)
    pose3 = np.zeros(32, S.pose_result_dt)
    curr3 = S.StereoFramesNP(32, d["curr_raw"].cap_pt, d["curr_raw"].cap_ln)
    rp.download(pose3, curr=curr3)
    ctx.sync()
    assert pose3.tobytes() == pg.tobytes()
    _cmp_stereo(curr3, orc.stereo_match(d["cam"], ctx.cfg, d["curr_raw"]))


def test_linecut_decisions_exact_at_scale(ctx, orc):
    """the determinant-lemma screening of the cut search must never change a decision: 1024 frame pairs
    (~20k lines, ~2M candidate comparisons) against the oracle, cut ratios and information matrices"""
    d = _data(1024, 1200)
    prev, curr, obs = _obs(orc, ctx.cfg, d)
    cov = orc.stereo_uncertainty(d["cam"], ctx.cfg, prev)
    DT_inv = _dt_inv(d["state"])
    og, oo = obs.copy(), obs.copy()
    cg = ctx.linecut(DT_inv, og, cov)
    co = orc.linecut(d["cam"], ctx.cfg, DT_inv, oo, cov)
    mask = np.arange(obs.cap_ln)[None, :] < obs.n_ln[:, None]
    assert mask.sum() > 10000
    diff = (cg["cutRatio"] != co["cutRatio"]).any(-1) & mask
    assert diff.sum() == 0, f"{diff.sum()} of {mask.sum()} lines with a different cut decision"
    assert close(cg["invCovPose"][mask], co["invCovPose"][mask])
    assert close(og.ln["sP"][mask], oo.ln["sP"][mask]) and close(og.ln["eP"][mask], oo.ln["eP"][mask])


def test_track_dense_c4_and_kitti_shapes(orc):
    """BASELINE configs[3] (2000 ORB + 500 lines per image) and the KITTI-shaped camera of configs[2], whole step"""
    from gfpl import api
    flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
    for shape, n_pt, n_ln, B in (("euroc", 2000, 500, 3), ("kitti", 300, 100, 6)):
        d = synth.make_pairs(B, n_pt=n_pt, n_ln=n_ln, shape=shape, seed=77)
        c = api.Context(0, cam=d["cam"])
        prev = orc.stereo_match(d["cam"], c.cfg, d["prev_raw"])
        cg, og, kg, pg = c.track(prev, d["state"], d["curr_raw"], flags)
        co, oo, ko, po = orc.track(d["cam"], c.cfg, prev, d["state"], d["curr_raw"], flags)
        _cmp_stereo(cg, co)
        assert np.array_equal(og.n_pt, oo.n_pt) and np.array_equal(og.n_ln, oo.n_ln)
        assert (og.n_pt <= c.cfg.max_point_match_num).all() and (og.n_ln <= c.cfg.max_line_match_num).all()
        if n_pt == 2000:
            assert (og.n_pt == 500).all()                   # the 500-match budget of Config::maxPointMatchNum binds
        for p in range(B):
            assert np.array_equal(og.pt_inlier[p, :og.n_pt[p]], oo.pt_inlier[p, :oo.n_pt[p]])
            assert np.array_equal(og.ln_inlier[p, :og.n_ln[p]], oo.ln_inlier[p, :oo.n_ln[p]])
            assert np.array_equal(kg["cutRatio"][p, :og.n_ln[p]], ko["cutRatio"][p, :oo.n_ln[p]])
        _cmp_pose(pg, po, B)


def test_full_batch_properties(ctx):
    """8192-pair replay (BASELINE configs[4]): size-independent properties — every pose is a rigid transform, Tfw
    composes, the recovered motion is close to the simulated one, covariances are symmetric PSD, and a second
    replay object fed the same inputs reproduces the records bit for bit (checksum of checksums)."""
    from gfpl import api
    import zlib
    B = 8192
    d = _data(B, 4242)
    flags = S.TRACK_LINECUT | S.TRACK_SEED_GIVEN
    prev = S.StereoFramesNP(B, 300, 100)
    for lo in range(0, B, 2048):
        part = ctx.stereo_match(d["prev_raw"].slice(lo, lo + 2048))
        for k in S.StereoFramesNP._keys:
            getattr(prev, k)[lo:lo + 2048] = getattr(part, k)
    sums = []
    for _ in range(2):
        rp = api.Replay(ctx, B, 300, 100)
        rp.upload(prev, d["state"], d["curr_raw"])
        rp.run(flags)
        pose = np.zeros(B, S.pose_result_dt)
        rp.download(pose)
        ctx.sync()
        sums.append(zlib.crc32(np.array([zlib.crc32(pose[i].tobytes()) for i in range(B)], np.uint32).tobytes()))
        rp.close()
    assert sums[0] == sums[1]
    ok = pose["status"] == 0
    assert ok.mean() > 0.995
    DT = pose["DT"].reshape(B, 4, 4)[ok]
    R = DT[:, :3, :3]
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-9 and np.allclose(np.linalg.det(R), 1.0, atol=1e-9)
    Tfw = pose["Tfw"].reshape(B, 4, 4)[ok]
    assert np.abs(Tfw - d["state"]["Tfw"].reshape(B, 4, 4)[ok] @ DT).max() < 1e-9
    err = np.abs(np.linalg.inv(DT) - d["DT_true"][ok]).max(axis=(1, 2))
    assert np.median(err) < 5e-3 and (err < 5e-2).mean() > 0.99
    cov = pose["DT_cov"].reshape(B, 6, 6)[ok]
    assert np.abs(cov - cov.transpose(0, 2, 1)).max() <= 1e-9 * np.abs(cov).max()
    assert (pose["DT_cov_eig"][ok][:, 0] > -1e-12).all()


def test_kf_match_lookForCommonMatches(ctx, orc):
    """KF<->KF matching + chi-square gates of MapHandler::lookForCommonMatches (mapHandler.cpp:211-490)"""
    d = _data(12, 1300, ragged=True)
    kf0, kf1 = _frames(orc, ctx.cfg, d)
    rng = np.random.default_rng(5)
    DT = d["DT_true"].copy()
    DT[:, :3, 3] += rng.normal(0, 0.01, (12, 3))          # imperfect relative pose: the gate rejects some pairs
    g = ctx.kf_match(kf0, kf1, DT.reshape(-1, 16))
    o = orc.kf_match(d["cam"], ctx.cfg, kf0, kf1, DT.reshape(-1, 16))
    # (the line gate compares the re-projected kf0 line with kf0's OWN 2-D line, mapHandler.cpp:398: few lines survive motion)
    assert (o.n_pt > 50).all() and o.n_ln.sum() > 10 and (o.n_pt < kf0.n_pt).all()
    _cmp_obs(g, o)


def _local_map(d, kf0, kf1, seed):
    """local map = kf0's stereo landmarks (world = kf0 frame) + landmarks outside kf1's view; some kf1 features
    already carry a landmark id (idx != -1) as after the KF<->KF half"""
    rng = np.random.default_rng(seed)
    B = kf0.B
    extra_p, extra_l = 40, 12
    m = S.LocalMapNP(B, kf0.cap_pt + extra_p, kf0.cap_ln + extra_l)
    kf1 = kf1.copy()
    for p in range(B):
        n, nl = int(kf0.n_pt[p]), int(kf0.n_ln[p])
        X = np.concatenate([kf0.pt["P"][p, :n], rng.uniform(-30, 30, (extra_p, 3))])   # extras: behind / outside
        D = np.concatenate([kf0.pdesc[p, :n], rng.integers(0, 256, (extra_p, 32), dtype=np.uint8)])
        perm = rng.permutation(len(X))
        m.n_pt[p] = len(X); m.pt3d[p, :len(X)] = X[perm]; m.pdesc[p, :len(X)] = D[perm]
        L = np.concatenate([np.concatenate([kf0.ln["sP"][p, :nl], kf0.ln["eP"][p, :nl]], 1),
                            rng.uniform(-30, 30, (extra_l, 6))])
        DL = np.concatenate([kf0.ldesc[p, :nl], rng.integers(0, 256, (extra_l, 32), dtype=np.uint8)])
        perm = rng.permutation(len(L))
        m.n_ln[p] = len(L); m.ln3d[p, :len(L)] = L[perm]; m.ldesc[p, :len(L)] = DL[perm]
        kf1.pt["idx"][p] = np.where(rng.random(kf1.cap_pt) < 0.6, -1, 7)
        kf1.ln["idx"][p] = np.where(rng.random(kf1.cap_ln) < 0.7, -1, 3)
    return m, kf1


def _cmp_map_matches(g, o):
    assert np.array_equal(g.n_pt, o.n_pt) and np.array_equal(g.n_ln, o.n_ln)
    for p in range(g.B):
        assert np.array_equal(g.pt[p, :g.n_pt[p]], o.pt[p, :o.n_pt[p]])
        assert np.array_equal(g.ln[p, :g.n_ln[p]], o.ln[p, :o.n_ln[p]])


@pytest.mark.parametrize("ragged", [False, True])
def test_map_match_lookForCommonMatches(ctx, orc, ragged):
    """local map <-> KF half of MapHandler::lookForCommonMatches (mapHandler.cpp:509-770): index pairs bit-exact"""
    d = _data(10, 1400, ragged=ragged)
    kf0, kf1 = _frames(orc, ctx.cfg, d)
    lmap, kf1 = _local_map(d, kf0, kf1, 9)
    Twf = d["DT_true"].copy()
    Twf[:, :3, 3] += np.random.default_rng(6).normal(0, 0.002, (10, 3))
    g = ctx.map_match(lmap, kf1, Twf.reshape(-1, 16))
    o = orc.map_match(d["cam"], ctx.cfg, lmap, kf1, Twf.reshape(-1, 16))
    assert (o.n_pt > 10).all() and o.n_ln.sum() > 10
    assert (kf1.pt["idx"][np.arange(10)[:, None], o.pt["kf_idx"]][np.arange(o.cap_pt)[None] < o.n_pt[:, None]] == -1).all()
    _cmp_map_matches(g, o)


def test_map_match_degenerate(ctx, orc):
    d = _data(4, 1401, n_pt=64, n_ln=24)
    kf0, kf1 = _frames(orc, ctx.cfg, d)
    lmap, kf1 = _local_map(d, kf0, kf1, 10)
    lmap.n_pt[0] = 0; lmap.n_ln[1] = 1                      # empty / single-row map sides (D3)
    kf1.pt["idx"][2] = 5; kf1.ln["idx"][3] = 5              # nothing unmatched
    Twf = d["DT_true"].reshape(-1, 16)
    g = ctx.map_match(lmap, kf1, Twf)
    o = orc.map_match(d["cam"], ctx.cfg, lmap, kf1, Twf)
    assert o.n_pt[0] == 0 and o.n_ln[1] == 0 and o.n_pt[2] == 0 and o.n_ln[3] == 0
    _cmp_map_matches(g, o)


def test_medoid_desc_updateAverageDescDir(ctx, orc):
    """representative descriptor of a landmark (mapFeatures.cpp:51-82): index bit-exact, incl. ties and long lists"""
    rng = np.random.default_rng(77)
    ns = np.concatenate([rng.integers(2, 12, 3000), [2, 33, 64, 150, 1, 0, 2]])
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    base = rng.integers(0, 256, (len(ns), 32), dtype=np.uint8)
    desc = np.repeat(base, ns, axis=0)
    flip = (rng.random((len(desc), 256)) < 0.05)
    desc ^= np.packbits(flip, axis=1)
    desc[off[5]:off[6]] = desc[off[5]]                       # all-equal list: every statistic ties -> index 0
    g = ctx.medoid_desc(desc, off)
    o = orc.medoid_desc(desc, off)
    assert np.array_equal(g, o) and g[5] == 0 and len(np.unique(g)) > 5
    # tie-stress: low-entropy descriptors
    t = np.zeros((4000, 32), np.uint8); t[:, 0] = rng.integers(0, 4, 4000)
    off2 = np.arange(0, 4001, 8, dtype=np.int32)
    assert np.array_equal(ctx.medoid_desc(t, off2), orc.medoid_desc(t, off2))


@pytest.mark.parametrize("ragged", [False, True])
def test_stereo_match_orbslam(ctx, orc, ragged):
    """extractStereoFeatures_ORBSLAM (stereoFrame.cpp:453-767): row-band search, SAD refinement, median cut — bit-exact"""
    d = _data(8, 1500, ragged=ragged)
    pyr = synth.make_right_pyramids(8, d["cam"], seed=3)
    g = ctx.stereo_match_orbslam(d["curr_raw"], pyr)
    o = orc.stereo_match_orbslam(d["cam"], ctx.cfg, d["curr_raw"], pyr)
    assert (o.n_pt > 15).all() and (o.n_ln > 3).all()
    _cmp_stereo(g, o)
    # output order = ascending (descriptor distance, left index): left indices are not monotone
    assert any((np.diff(o.pt["src_l"][p, :o.n_pt[p]]) < 0).any() for p in range(8))


def test_stereo_match_orbslam_dense_and_degenerate(ctx, orc):
    d = _data(1, 1501, n_pt=2500, n_ln=40)               # > 8 key points per thread
    pyr = synth.make_right_pyramids(1, d["cam"], seed=4)
    o = orc.stereo_match_orbslam(d["cam"], ctx.cfg, d["curr_raw"], pyr)
    assert o.n_pt[0] > 200
    _cmp_stereo(ctx.stereo_match_orbslam(d["curr_raw"], pyr), o)
    d = _data(4, 1502, n_pt=64, n_ln=24)
    pyr = synth.make_right_pyramids(4, d["cam"], seed=5)
    raw = d["curr_raw"]
    raw.n_pt_l[0] = 0; raw.n_pt_r[1] = 0                                   # :453
    raw.pdesc_r[2] = 255 - raw.pdesc_r[2]                                  # no candidate below thOrbDist -> empty vDistIdx
    raw.pt_l["x"][3, :32] = 3.0                                            # windows leaving the level (defined: no match)
    o = orc.stereo_match_orbslam(d["cam"], ctx.cfg, raw, pyr)
    assert o.n_pt[0] == 0 and o.n_pt[1] == 0 and o.n_pt[2] == 0
    _cmp_stereo(ctx.stereo_match_orbslam(raw, pyr), o)
    with pytest.raises(Exception):
        bad = synth.make_right_pyramids(4, d["cam"], n_levels=3, seed=5)   # n_levels != cfg.orb_nlevels
        ctx.stereo_match_orbslam(raw, bad)


def _lba_residual(w, blk, DX, lam):
    """|| (H + lam diag H) DX - g || / || g ||  from the block-sparse assembly (numpy, no dense matrix)"""
    Nkf, Npt, Nls = len(w["kf_Tkw"]), len(w["X_pt"]), len(w["X_ls"])
    xk = DX[:6 * Nkf].reshape(Nkf, 6); xp = DX[6 * Nkf:6 * Nkf + 3 * Npt].reshape(Npt, 3); xl = DX[6 * Nkf + 3 * Npt:].reshape(Nls, 6)
    damp = lambda Hb, n: Hb.reshape(-1, n, n) + lam * np.einsum("bii->bi", Hb.reshape(-1, n, n))[:, :, None] * np.eye(n)
    rk = np.einsum("bij,bj->bi", damp(blk["Hkk"], 6), xk)
    rp = np.einsum("bij,bj->bi", damp(blk["Hpp"], 3), xp)
    rl = np.einsum("bij,bj->bi", damp(blk["Hll"], 6), xl)
    for W, lm, kf, r_lm, x_lm, dim in ((blk["Hpk"].reshape(-1, 3, 6), w["pt_obs_lm"], w["pt_obs_kf"], rp, xp, 3),
                                       (blk["Hlk"].reshape(-1, 6, 6), w["ls_obs_lm"], w["ls_obs_kf"], rl, xl, 6)):
        ok = kf >= 0
        np.add.at(r_lm, lm[ok], np.einsum("oij,oj->oi", W[ok], xk[kf[ok]]))
        np.add.at(rk, kf[ok], np.einsum("oij,oi->oj", W[ok], x_lm[lm[ok]]))
    r = np.concatenate([rk.ravel(), rp.ravel(), rl.ravel()]) - blk["g"]
    return np.linalg.norm(r) / np.linalg.norm(blk["g"])


def test_lba_step_schur_solve(ctx, orc):
    """assembly + damped solve of one LM pass (mapHandler.cpp:1428-1441): DX vs the dense CPU solve, 1e-9 relative"""
    from gfpl import api
    try:
        for seed, kw, lam, scale in ((3, dict(n_kf=6, n_pt=120, n_ls=30), 0.001, 1),
                                     (4, dict(n_kf=3, n_pt=20, n_ls=5, n_fix=1), 0.001, 1),
                                     (5, dict(n_kf=8, n_pt=150, n_ls=40), 37.5, 0),
                                     (6, dict(n_kf=4, n_pt=60, n_ls=0), 0.5, 0),
                                     (7, dict(n_kf=4, n_pt=0, n_ls=25), 0.5, 0)):
            w = synth.make_lba_window(seed=seed, **kw)
            ctx.set_camera(w["cam"])
            DXg, eg, lg = ctx.lba_step(w, lam, scale)
            DXo, eo, lo = orc.lba_step(w["cam"], ctx.cfg, w, lam, scale)
            assert lg == lo and abs(eg - eo) <= 1e-9 * abs(eo)
            assert np.isfinite(DXo).all() and np.abs(DXo).max() > 0
            assert np.linalg.norm(DXg - DXo) <= 1e-9 * np.linalg.norm(DXo), (seed, np.linalg.norm(DXg - DXo) / np.linalg.norm(DXo))
        # full C3 window: too large for the dense CPU solve -> residual of the damped system instead
        w = synth.make_lba_window(seed=8, n_kf=10, n_pt=1000, n_ls=300)
        ctx.set_camera(w["cam"])
        DX, err, lam_eff = ctx.lba_step(w, 0.001, 1)
        blk = ctx.lba_assemble(w)
        assert lam_eff > 0 and _lba_residual(w, blk, DX, lam_eff) < 1e-9
        # unsorted observation list is rejected, not mis-solved
        w2 = dict(w); w2["pt_obs_lm"] = w["pt_obs_lm"][::-1].copy()
        with pytest.raises(api.GfplError):
            ctx.lba_step(w2, 0.001, 1)
    finally:
        ctx.set_camera(api.euroc_camera())


def test_lba_step_batch_over_windows(ctx, orc):
    """gfpl_lba_step_batch: ragged windows (different key-frame, landmark and observation counts, a points-only, a
    lines-only and an EMPTY window) in one call == the windows solved one at a time (1e-9 vs the dense CPU solve, and the
    same figures as the single-window entry point up to the order of the fp64 atomics)"""
    from gfpl import api
    specs = [dict(seed=11, n_kf=6, n_pt=120, n_ls=30), dict(seed=12, n_kf=3, n_pt=20, n_ls=5, n_fix=1),
             dict(seed=13, n_kf=8, n_pt=150, n_ls=40), dict(seed=14, n_kf=4, n_pt=60, n_ls=0),
             dict(seed=15, n_kf=4, n_pt=0, n_ls=25), dict(seed=16, n_kf=10, n_pt=300, n_ls=80)]
    wins = [synth.make_lba_window(**kw) for kw in specs]
    empty = {k: v[:0].copy() if isinstance(v, np.ndarray) else v for k, v in wins[1].items()}
    wins.insert(2, empty)
    try:
        ctx.set_camera(wins[0]["cam"])
        for lam, scale in ((0.001, 1), (0.5, 0)):
            DXs, err, lam_eff = ctx.lba_step_batch(wins, lam, scale)
            assert len(DXs) == len(wins)
            for k, w in enumerate(wins):
                if len(w["kf_Tkw"]) == 0:
                    assert len(DXs[k]) == 0 and err[k] == 0.0 and lam_eff[k] == lam
                    continue
                DXo, eo, lo = orc.lba_step(w["cam"], ctx.cfg, w, lam, scale)
                assert lam_eff[k] == lo and abs(err[k] - eo) <= 1e-9 * abs(eo), k
                assert np.linalg.norm(DXs[k] - DXo) <= 1e-9 * np.linalg.norm(DXo), (k, np.linalg.norm(DXs[k] - DXo) / np.linalg.norm(DXo))
                DX1, e1, l1 = ctx.lba_step(w, lam, scale)
                assert l1 == lam_eff[k] and np.linalg.norm(DXs[k] - DX1) <= 1e-9 * np.linalg.norm(DX1)
        # 64 full-size C3 windows: residual of every damped system
        big = [synth.make_lba_window(seed=100 + i, n_kf=10, n_pt=1000, n_ls=300) for i in range(4)] * 16
        DXs, err, lam_eff = ctx.lba_step_batch(big, 0.001, 1)
        for k in (0, 1, 2, 3, 17, 63):
            blk = ctx.lba_assemble(big[k])
            assert lam_eff[k] > 0 and _lba_residual(big[k], blk, DXs[k], lam_eff[k]) < 1e-9
        assert np.array_equal(lam_eff[:4], lam_eff[4:8])
        assert ctx.timer_elapsed_ms(6, 7) > 0
        bad = list(wins)
        bad[3] = dict(wins[3]); bad[3]["pt_obs_lm"] = wins[3]["pt_obs_lm"][::-1].copy()
        with pytest.raises(api.GfplError):
            ctx.lba_step_batch(bad, 0.001, 1)
    finally:
        ctx.set_camera(api.euroc_camera())


@pytest.mark.parametrize("norm", [S.NORM_HAMMING, S.NORM_HAMMING2])
@pytest.mark.parametrize("tie", [False, True])
def test_knn2_mutual_both_directions(ctx, orc, norm, tie):
    """forward k=2 + reverse best match from one pass (tensor-core int8 +-64 products) == two knnMatch calls"""
    rng = np.random.default_rng(41 + norm)
    mk = (lambda n: synth.tie_stress_descriptors(rng, n)) if tie else (lambda n: rng.integers(0, 256, (n, 32), dtype=np.uint8))
    for cap_q, cap_t, sizes in ((300, 310, [(300, 310), (257, 129), (128, 160), (2, 3), (0, 5), (7, 0), (130, 161)]),
                                (1000, 900, [(1000, 900), (513, 321)]),                       # column-blocked tensor-core form
                                (2000, 2000, [(2000, 2000), (1999, 577), (300, 2000), (129, 289), (2000, 0), (5, 1153)]),
                                (100, 100, [(100, 100), (64, 17), (2, 2)])):
        B = len(sizes)
        q = np.zeros((B, cap_q, 32), np.uint8); t = np.zeros((B, cap_t, 32), np.uint8)
        nq = np.array([s[0] for s in sizes], np.int32); nt = np.array([s[1] for s in sizes], np.int32)
        for p in range(B):
            q[p, :nq[p]] = mk(nq[p]); t[p, :nt[p]] = mk(nt[p])
            m = min(nq[p], nt[p]) // 2
            t[p, :m] = q[p, :m] ^ np.packbits(rng.random((m, 256)) < 0.04, axis=1)     # true matches
        idx, dist, bi, bd = ctx.knn2_mutual(q, nq, t, nt, norm)
        for p in range(B):
            if nq[p] and nt[p]:
                oi, od = orc.knn2(q[p, :nq[p]], t[p, :nt[p]], norm)
                ri, rd = orc.knn2(t[p, :nt[p]], q[p, :nq[p]], norm)
                assert np.array_equal(idx[p, :nq[p]], oi) and np.array_equal(dist[p, :nq[p]], od), (cap_q, p)
                assert np.array_equal(bi[p, :nt[p]], ri[:, 0]) and np.array_equal(bd[p, :nt[p]], rd[:, 0]), (cap_q, p)
            else:
                assert (idx[p, :nq[p]] == -1).all() and (bi[p, :nt[p]] == -1).all()


@pytest.mark.parametrize("norm", [6, 7])
@pytest.mark.parametrize("tie", [False, True])
def test_knn2_mutual_small_problems_persistent(ctx, orc, norm, tie):
    """both sides <= 128 rows: the persistent small tensor-core kernel (NORM_HAMMING2 as a K = 384 product) next to the
    integer kernel that takes the tiny / empty problems — 700 ragged problems (more than the persistent grid), every
    border of the split (40 / 41 rows, 1 row, empty sides, 16-column chunk borders, full 128)"""
    rng = np.random.default_rng(97 + norm + 2 * tie)
    mk = (lambda n: synth.tie_stress_descriptors(rng, n)) if tie else (lambda n: rng.integers(0, 256, (n, 32), dtype=np.uint8))
    fixed = [(128, 128), (41, 1), (1, 41), (40, 41), (41, 40), (40, 40), (0, 50), (50, 0), (97, 113), (33, 128), (128, 33),
             (48, 48), (49, 47), (112, 16), (17, 127), (2, 64), (64, 2)]
    B = 700
    nq = rng.integers(0, 129, B).astype(np.int32); nt = rng.integers(0, 129, B).astype(np.int32)
    for p, (x, y) in enumerate(fixed):
        nq[p], nt[p] = x, y
    q = np.zeros((B, 128, 32), np.uint8); t = np.zeros((B, 128, 32), np.uint8)
    for p in range(B):
        q[p, :nq[p]] = mk(nq[p]); t[p, :nt[p]] = mk(nt[p])
        m = min(nq[p], nt[p]) // 2
        t[p, :m] = q[p, :m] ^ np.packbits(rng.random((m, 256)) < 0.04, axis=1)
    idx, dist, bi, bd = ctx.knn2_mutual(q, nq, t, nt, norm)
    for p in range(B):
        if nq[p] and nt[p] >= 2:
            oi, od = orc.knn2(q[p, :nq[p]], t[p, :nt[p]], norm)
            assert np.array_equal(idx[p, :nq[p]], oi) and np.array_equal(dist[p, :nq[p]], od), (p, nq[p], nt[p])
        elif nq[p] and nt[p] == 1:
            assert (idx[p, :nq[p], 0] == 0).all() and (idx[p, :nq[p], 1] == -1).all() and np.isinf(dist[p, :nq[p], 1]).all()
        else:
            assert (idx[p, :nq[p]] == -1).all()
        if nt[p] and nq[p] >= 2:
            ri, rd = orc.knn2(t[p, :nt[p]], q[p, :nq[p]], norm)
            assert np.array_equal(bi[p, :nt[p]], ri[:, 0]) and np.array_equal(bd[p, :nt[p]], rd[:, 0]), (p, nq[p], nt[p])
        elif nt[p] and nq[p] == 1:
            assert (bi[p, :nt[p]] == 0).all()
        else:
            assert (bi[p, :nt[p]] == -1).all()


def test_integer_pipe_paths_with_tensor_path_disabled():
    """GFPL_NO_TC=1 routes point matching to the integer-pipe kernels (knn2_sym_kernel, f2f_hybrid_points_kernel):
    the same parity tests must pass there (the switch is read once per process, hence the subprocess)"""
    import subprocess, sys
    env = dict(os.environ, GFPL_NO_TC="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "test_stereo_match or test_f2f_match or test_knn2_mutual or test_track_whole or test_kf_match or test_map_match"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "deselected" in r.stdout
