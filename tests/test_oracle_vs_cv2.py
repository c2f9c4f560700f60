"""CPU: pins the oracle against cv2 live (the reference's arithmetic owner) on fresh seeds.

Skipped when cv2 is not importable; the committed fixtures (test_oracle_golden.py) cover that
case.  This is where "the reference's own implementation" enters: the reference cannot be
compiled in this image (OpenCV C++ dev files absent), and its hot path is these OpenCV calls.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import c_oracle, synth
from oracle.gen_golden import cv_knn2


@pytest.mark.parametrize("seed", [11, 12])
def test_l2_int_and_float_bit_exact(seed):
    for q, t in (synth.sift_pair(400, 500, seed), synth.float_pair(300, 350, seed)):
        gi, gd, _ = cv_knn2(q, t, cv2.NORM_L2)
        idx, dist = c_oracle.l2_knn2(q, t)
        assert np.array_equal(idx, gi)
        assert np.array_equal(dist.view(np.int32), gd.view(np.int32))


def test_hamming_bit_exact():
    q, t = synth.orb_pair(500, 600, 21)
    gi, gd, _ = cv_knn2(q, t, cv2.NORM_HAMMING)
    idx, dist = c_oracle.hamming_knn2(q, t)
    assert np.array_equal(idx, gi) and np.array_equal(dist, gd)


def test_good_matches_equal_reference_loop():
    """matchFeatures end to end: knnMatch + the getGoodMatches loop written on cv2's own lists."""
    q, t = synth.sift_pair(500, 600, 31)
    res = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, 2)
    good = [r[0] for r in res if len(r) and r[0].distance < 0.7 * r[1].distance]
    got = c_oracle.match_features(0, q, t, 0.7)
    assert [m.queryIdx for m in good] == list(got["queryIdx"])
    assert [m.trainIdx for m in good] == list(got["trainIdx"])
    assert np.array_equal(np.array([m.distance for m in good], np.float32), got["distance"])
    assert len(good) > 50


@pytest.mark.parametrize("seed", [9100, 9101, 9102])
def test_find_essential_mat_mask(seed):
    p1, p2, _, _ = synth.two_view(1200, seed, noise_px=1.5)
    K4 = np.array(synth.SAMSUNG_HV_4K)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    E, mask = cv2.findEssentialMat(p1, p2, Kmat, cv2.RANSAC, 0.999, 5.0)
    E = np.asarray(E, np.float64).reshape(-1, 9)[:1]
    counts, best, m, _ = c_oracle.score_essential(p1, p2, K4, E, 5.0)
    assert best == 0 and np.array_equal(m, mask.reshape(-1))
