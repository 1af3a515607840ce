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
