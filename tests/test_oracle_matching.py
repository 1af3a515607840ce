"""Oracle vs the committed cv2.BFMatcher golden vectors (tests/golden, made by oracle/make_golden.py) and the
robust-threshold helpers vs plain numpy restatements of the reference lines."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bfmatcher_*.npz")))


def test_golden_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[10:-4] for p in GOLDEN])
@pytest.mark.parametrize("norm", [6, 7])
def test_knn2_matches_cv2(orc, path, norm):
    g = np.load(path)
    idx, dist = orc.knn2(g["q"], g["t"], norm)
    assert (idx == g[f"knn_idx_{norm}"]).all()          # bit-exact indices incl. tie order
    assert (dist == g[f"knn_dist_{norm}"]).all()        # bit-exact distances


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[10:-4] for p in GOLDEN])
@pytest.mark.parametrize("norm", [6, 7])
def test_radius_matches_cv2(orc, path, norm):
    g = np.load(path)
    rc, off, idx, dist = orc.radius(g["q"], g["t"], float(g["radius"]), norm)
    assert rc == 0
    assert (off == g[f"rad_off_{norm}"]).all()
    assert (idx == g[f"rad_idx_{norm}"]).all() and (dist == g[f"rad_dist_{norm}"]).all()


def test_radius_capacity(orc):
    g = np.load(GOLDEN[0])
    rc, off, _, _ = orc.radius(g["q"], g["t"], 300.0, 6, cap=10)
    assert rc == -5 and off[-1] == len(g["q"]) * len(g["t"])


def test_knn2_rejects_single_train_row(orc):
    q = np.zeros((3, 32), np.uint8)
    with pytest.raises(ValueError):
        orc.knn2(q, q[:1])


def test_line_descriptor_mad(orc):
    rng = np.random.default_rng(1)
    d = np.sort(rng.integers(0, 120, (101, 2)).astype(np.float32), axis=1)
    # stereoFrame.cpp:1306-1311 with nn_dist_median := U
    for U in (0.0, 3.5):
        v = np.abs((d[:, 1] - d[:, 0]).astype(np.float64) - U).astype(np.float32)
        ref = 1.4826 * float(np.sort(v)[len(v) // 2])
        assert orc.line_descriptor_mad(d, U) == ref


def test_budget_thres(orc):
    rng = np.random.default_rng(2)
    d = np.sort(rng.integers(0, 120, (77, 2)).astype(np.float32), axis=1)
    for m in (300, 50, 1):
        ref = float(np.sort(d[:, 0])[min(m, len(d)) - 1])  # stereoFrame.cpp:1322-1325
        assert orc.budget_thres(d, m) == ref


def test_vector_stdv_mad(orc):
    rng = np.random.default_rng(3)
    r = rng.random(200) * 3
    s = np.sort(r)
    med = s[len(s) // 2]
    a = np.sort(np.abs((r - med).astype(np.float32)).astype(np.float64))   # fabsf: float truncation
    assert orc.vector_stdv_mad(r) == 1.4826 * a[len(a) // 2]
    assert orc.vector_stdv_mad(np.zeros(0)) == 0.0


# ---- the identities the tensor-core kernels rest on, checked against the oracle's distances (no GPU) ------------------
def _pm64(bits):
    """descriptor bits -> int8 operand of the tensor-core kernels: set -> +64, clear -> -64"""
    return np.where(bits > 0, 64, -64).astype(np.int64)


def test_tensor_core_products_are_the_distances(orc):
    """NORM_HAMMING: with +-64 operands (query side negated) a K = 256 product is (hamming << 13) - 2^20.
    NORM_HAMMING2: appending the 128 products of the two bits of every 2-bit group (XOR in +-1 arithmetic) gives, at
    K = 384 and equal signs, 4096 * 384 - (distance << 14) — the accumulators knn2_tc_kernel / knn2_tc_small_kernel read."""
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (37, 32), dtype=np.uint8)
    t[:10] = q[:10] ^ np.packbits(rng.random((10, 256)) < 0.05, axis=1)
    t[10] = q[10]                                                        # distance 0
    bq = np.unpackbits(q, axis=1, bitorder="little").astype(np.int64)   # bit k of the descriptor
    bt = np.unpackbits(t, axis=1, bitorder="little").astype(np.int64)
    ham = (bq[:, None, :] != bt[None, :, :]).sum(-1)
    acc = (-_pm64(bq)) @ _pm64(bt).T                                     # query side negated
    assert np.array_equal(acc, (ham << 13) - (1 << 20))
    # the oracle's NORM_HAMMING distances are these Hamming distances
    oi, od = orc.knn2(q, t, 6)
    assert np.array_equal(od[:, 0], ham.min(1).astype(np.float32))
    # NORM_HAMMING2: groups of two bits (2k, 2k + 1)
    g_q = bq[:, 0::2] ^ bq[:, 1::2]
    g_t = bt[:, 0::2] ^ bt[:, 1::2]
    ham2 = ((bq[:, None, 0::2] != bt[None, :, 0::2]) | (bq[:, None, 1::2] != bt[None, :, 1::2])).sum(-1)
    a384 = np.concatenate([_pm64(bq), _pm64(g_q)], axis=1)
    b384 = np.concatenate([_pm64(bt), _pm64(g_t)], axis=1)
    acc2 = a384 @ b384.T
    assert np.array_equal(acc2, 4096 * 384 - (ham2 << 14))
    assert np.abs(acc2).max() < 2 ** 31 and np.abs(acc).max() < 2 ** 31
    oi2, od2 = orc.knn2(q, t, 7)
    assert np.array_equal(od2[:, 0], ham2.min(1).astype(np.float32))
    # the byte permutes of the expansion reorder K inside a 32-bit word, identically on both sides: a dot product is blind to it
    perm = np.concatenate([np.array([s + 4 * j for s in range(4) for j in range(4)] + [16 + s + 4 * j for s in range(4) for j in range(4)]) + 32 * w
                           for w in range(8)])
    assert sorted(perm.tolist()) == list(range(256))
    assert np.array_equal((-_pm64(bq))[:, perm] @ _pm64(bt)[:, perm].T, acc)


def test_endpoint_covariance_eigenvalue_bounds():
    """stereo_accept_kernel decides the line-covariance gate by Weyl bounds: the end-point covariance of
    stereoFrame.cpp:956-998 is k (2 v v^T + diag(d^2, d^2, 0)), so its largest eigenvalue lies in
    [2 k |v|^2, k (2 |v|^2 + d^2)] — checked against numpy's eigenvalues over the image and the disparity range."""
    rng = np.random.default_rng(6)
    f, cx, cy, b = 458.654, 367.215, 248.375, 0.110078
    for _ in range(2000):
        px, py, d = rng.uniform(0, 752), rng.uniform(0, 480), rng.uniform(1.0, 60.0)
        x, y = px - cx, py - cy
        c = np.array([[d * d + 2 * x * x, 2 * x * y, 2 * f * x],
                      [2 * x * y, d * d + 2 * y * y, 2 * f * y],
                      [2 * f * x, 2 * f * y, 2 * f * f]]) * b * b / d ** 4
        lam = np.linalg.eigvalsh(c)[-1]
        k = b * b / d ** 4
        v2 = x * x + y * y + f * f
        lo, hi = 2 * k * v2, k * (2 * v2 + d * d)
        assert lo * (1 - 1e-12) <= lam <= hi * (1 + 1e-12)
        assert (hi - lo) / lo < 1e-2            # (< 1e-4 for disparities up to ~6 px, where the gate's threshold sits)
