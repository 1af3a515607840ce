"""Generates tests/golden/bfmatcher_*.npz: inputs + outputs of cv2.BFMatcher (OpenCV-Python 4.13 in the build
container) for the matching semantics the reference gets from cv::BFMatcher 3.4.1.

Run from the repo root:  python oracle/make_golden.py
The fixtures are committed; the GPU box never needs cv2 or /root/reference.
radiusMatch lists are stored re-sorted by (distance, trainIdx): OpenCV's order inside equal distances is
unspecified (unstable std::sort), the candidate SET and distances are what is pinned.
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf-pl-slam_b200", "python"))
from gfpl import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NORMS = {6: cv2.NORM_HAMMING, 7: cv2.NORM_HAMMING2}


def knn(q, t, norm):
    m = cv2.BFMatcher(NORMS[norm], False).knnMatch(q, t, 2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), np.inf, np.float32)
    for i, row in enumerate(m):
        assert all(x.queryIdx == i for x in row)
        for k, x in enumerate(row):
            idx[i, k], dist[i, k] = x.trainIdx, x.distance
    return idx, dist


def radius(q, t, norm, r):
    m = cv2.BFMatcher(NORMS[norm], False).radiusMatch(q, t, r)
    off = np.zeros(len(q) + 1, np.int32)
    idx, dist = [], []
    for i, row in enumerate(m):
        row = sorted((int(x.distance), x.trainIdx) for x in row)
        off[i + 1] = off[i] + len(row)
        dist += [d for d, _ in row]
        idx += [j for _, j in row]
    return off, np.array(idx, np.int32), np.array(dist, np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.Generator(np.random.PCG64(20181018))
    cases = {}
    cases["random_300x310"] = (rng.integers(0, 256, (300, 32), dtype=np.uint8),
                               rng.integers(0, 256, (310, 32), dtype=np.uint8), 105.0)
    cases["tie_300x310"] = (synth.tie_stress_descriptors(rng, 300), synth.tie_stress_descriptors(rng, 310), 6.0)
    base = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    cases["duplicates_64x80"] = (np.concatenate([base, base[:24]]), np.concatenate([base, base]), 1.0)
    cases["nt2_17x2"] = (rng.integers(0, 256, (17, 32), dtype=np.uint8), rng.integers(0, 256, (2, 32), dtype=np.uint8), 128.0)
    noisy = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    flips = (rng.random((500, 256)) < 0.05)
    cases["matched_500x500"] = (noisy, noisy[rng.permutation(500)] ^ np.packbits(flips, axis=1), 50.0)
    for name, (q, t, r) in cases.items():
        out = {"q": q, "t": t, "radius": np.float32(r)}
        for norm in (6, 7):
            i, d = knn(q, t, norm)
            off, ri, rd = radius(q, t, norm, r)
            out[f"knn_idx_{norm}"], out[f"knn_dist_{norm}"] = i, d
            # the reverse call of a mutual check (knnMatch(t, q), stereoFrame.cpp:808): pins the direction the
            # tensor-core kernels take from the transposed product
            out[f"rknn_idx_{norm}"], out[f"rknn_dist_{norm}"] = knn(t, q, norm)
            out[f"rad_off_{norm}"], out[f"rad_idx_{norm}"], out[f"rad_dist_{norm}"] = off, ri, rd
        np.savez_compressed(os.path.join(OUT, f"bfmatcher_{name}.npz"), **out)
        print(name, q.shape, t.shape, "radius entries", int(out["rad_off_6"][-1]), int(out["rad_off_7"][-1]))


if __name__ == "__main__":
    main()
