"""GPU parity tests for the next row SURVEY.md 8f-3 (ORB half): descriptors of given keypoints through
the C ABI, bit-identical to the CPU oracle (pinned to cv2.ORB.compute) and to cv2 itself.
Mirrors extractDescriptor(frame, features, ORB_BF, desc) at featureMatchingCPU.cpp:45-66."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, synth
from slam_indoor_code_b200 import orb_descriptors as od
from slam_indoor_code_b200.feature_matching import MatcherType


def _frame(h, w, seed, channels=3):
    return synth.textured_frame(h, w, seed, channels)


def _grid_keypoints(h, w, n, seed, angle=-1.0):
    rng = np.random.default_rng(seed)
    k = np.zeros((n, 3), np.float32)
    k[:, 0] = rng.integers(0, w, n)
    k[:, 1] = rng.integers(0, h, n)
    k[:, 2] = angle if angle is not None else rng.uniform(0, 360, n)
    return k


@pytest.mark.parametrize("h,w,ch,n,angle,seed", [(480, 640, 3, 5000, -1.0, 1), (720, 1280, 1, 12000, -1.0, 2),
                                                 (333, 517, 3, 3000, None, 3), (2160, 3840, 3, 12000, -1.0, 4),
                                                 (70, 90, 1, 2000, None, 5)])
def test_orb_descriptors_vs_oracle(ctx, h, w, ch, n, angle, seed):
    frame = _frame(h, w, seed, ch)
    kps = _grid_keypoints(h, w, n, seed + 10, angle)
    keep, desc, res = od.extractDescriptorORB(ctx, frame, kps, want_resident=True)
    rkeep, rdesc = c_oracle.orb_compute(frame, kps)
    assert np.array_equal(keep, rkeep.astype(bool)) and keep.sum() > 0
    assert np.array_equal(desc, rdesc)
    assert res.n == len(rdesc)
    # the resident set feeds the matcher without an upload: matching it against itself finds itself
    idx, dist = ctx.knnMatch(MatcherType.ORB_BF, res, res)
    oi, odist = c_oracle.hamming_knn2(rdesc, rdesc)
    assert np.array_equal(idx, oi) and np.array_equal(dist, odist)
    res.free()


def test_orb_descriptors_vs_cv2_fast_keypoints(ctx):
    """The reference's own sequence: FAST(10, true, TYPE_9_16) keypoints, ORB::compute."""
    cv2 = pytest.importorskip("cv2")
    frame = _frame(720, 1280, 21, 3)
    kps = cv2.FastFeatureDetector_create(10, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(frame)
    assert len(kps) > 1000
    kept_cv, desc_cv = cv2.ORB_create().compute(frame, kps)
    keep, desc, _ = od.extractDescriptorORB(ctx, frame, kps)
    assert keep.sum() == len(kept_cv)
    assert all(kps[i].pt == k.pt for i, k in zip(np.nonzero(keep)[0], kept_cv))
    assert np.array_equal(desc, desc_cv)
    # oriented keypoints too (ORB's own detector sets angles)
    kps2 = cv2.ORB_create(2000).detect(frame)
    kps2 = [k for k in kps2 if k.octave == 0]
    kept_cv, desc_cv = cv2.ORB_create().compute(frame, kps2)
    keep, desc, _ = od.extractDescriptorORB(ctx, frame, kps2)
    assert keep.sum() == len(kept_cv) and np.array_equal(desc, desc_cv)


def test_orb_descriptors_edges(ctx):
    frame = _frame(200, 300, 31, 3)
    # every keypoint too close to the border: nothing kept; no keypoints at all
    k = np.array([[5, 5, -1], [299, 100, -1], [150, 199, -1], [30.4, 100, -1], [100, 168.6, -1]], np.float32)
    keep, desc, res = od.extractDescriptorORB(ctx, frame, k, want_resident=True)
    assert not keep.any() and desc.shape == (0, 32) and res.n == 0
    keep, desc, _ = od.extractDescriptorORB(ctx, frame, np.zeros((0, 3), np.float32))
    assert keep.shape == (0,) and desc.shape == (0, 32)
    # exactly on the inclusive / exclusive limits, sub-pixel positions (cvRound to the centre pixel)
    # (the border test is on the ROUNDED position: Rect::contains(Point(pt)))
    k = np.array([[31, 31, -1], [268.4, 168.4, -1], [268.6, 100, -1], [100.5, 101.5, 17.0], [30.6, 100.51, 200.0]],
                 np.float32)
    keep, desc, _ = od.extractDescriptorORB(ctx, frame, k)
    rkeep, rdesc = c_oracle.orb_compute(frame, k)
    assert np.array_equal(keep, rkeep.astype(bool)) and list(keep) == [True, True, False, True, True]
    assert np.array_equal(desc, rdesc)
    # row pitch (cv::Mat::step) and a gray frame
    wide = np.zeros((200, 400, 3), np.uint8)
    wide[:, :300] = frame
    keep, desc, _ = od.extractDescriptorORB(ctx, wide[:, :300], k)
    assert np.array_equal(desc, rdesc)
    with pytest.raises(TypeError):
        od.extractDescriptorORB(ctx, frame.astype(np.float32), k)


def test_orb_descriptors_golden(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "orb_desc.npz"))
    keep, desc, _ = od.extractDescriptorORB(ctx, g["frame"], g["kps"])
    assert np.array_equal(g["kps"][keep, :2], g["kept_xy"]) and np.array_equal(desc, g["desc"])
