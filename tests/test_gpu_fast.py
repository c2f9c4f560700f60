"""FAST keypoints on the B200: the reference's fastExtractor
(featureExtraction/fastExtractor.cpp:7-13) through the C ABI against the oracle and the committed
cv2 fixture -- positions, order and responses bit for bit."""
import os

import numpy as np
import pytest

from oracle import c_oracle, synth
from slam_indoor_code_b200 import fast_extractor as fe
from slam_indoor_code_b200 import orb_descriptors as od

pytestmark = pytest.mark.gpu


def test_fast_golden(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "fast.npz"))
    for thr, nms in ((10, True), (10, False), (25, True)):
        got = fe.fastExtractor(ctx, g["frame"], thr, nms)
        assert np.array_equal(got, g[f"kp_t{thr}_n{int(nms)}"])


@pytest.mark.parametrize("h,w,ch,seed,thr,nms", [(480, 640, 3, 9600, 10, True), (301, 457, 1, 9601, 10, False),
                                                  (240, 320, 3, 9602, 0, True), (255, 257, 3, 9603, 35, True),
                                                  (7, 9, 3, 9604, 10, True), (6, 64, 1, 9605, 10, True),
                                                  (8, 300, 1, 9606, 3, True), (200, 200, 3, 9607, 255, True),
                                                  (1080, 1920, 3, 9608, 10, True)])
def test_fast_seeded(ctx, h, w, ch, seed, thr, nms):
    frame = synth.textured_frame(h, w, seed, ch)
    got = fe.fastExtractor(ctx, frame, thr, nms)
    ref = c_oracle.fast_detect(frame, thr, nms)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_fast_row_pitch_small_buffer_and_4k(ctx):
    # a view with padding between rows (cv::Mat::step > cols * channels)
    big = synth.textured_frame(300, 500, 9610, 3)
    view = big[:, :401]
    img = np.ascontiguousarray(view)
    ref = c_oracle.fast_detect(img, 10, True)
    import ctypes
    from slam_indoor_code_b200._capi import check, ptr
    kps = np.zeros((len(ref) + 10, 3), np.float32)
    n = ctypes.c_int(0)
    check(ctx._lib.slamb200_fast_detect(ctx._h, ptr(big), 300, 401, 3, big.strides[0], 10, 1, ptr(kps), len(kps),
                                        ctypes.byref(n)))
    assert n.value == len(ref) and np.array_equal(kps[: n.value], ref)
    # an output buffer that is too small: the count is still the frame's, the first rows are written
    small = fe.fastExtractor(ctx, img, 10, True, max_points=100)
    assert np.array_equal(small, ref[:100])
    # 4K frame, the default buffer grows if it has to
    frame = synth.textured_frame(2160, 3840, 9611, 3)
    got = fe.fastExtractor(ctx, frame, 10, True)
    ref = c_oracle.fast_detect(frame, 10, True)
    assert np.array_equal(got, ref) and len(ref) > 100000


def test_fast_then_orb_descriptors(ctx):
    """The reference's front end for useFM-ORB: fastExtractor then extractDescriptor."""
    frame = synth.textured_frame(480, 640, 9620, 3)
    pts = fe.fastExtractor(ctx, frame, 10, True)
    kps = fe.to_orb_keypoints(pts)
    keep, desc, _ = od.extractDescriptorORB(ctx, frame, kps)
    rkeep, rdesc = c_oracle.orb_compute(frame, fe.to_orb_keypoints(c_oracle.fast_detect(frame, 10, True)))
    assert np.array_equal(np.asarray(keep, bool), rkeep.astype(bool)) and np.array_equal(desc, rdesc)


@pytest.mark.parametrize("h,w,ch,seed", [(480, 640, 3, 9630), (300, 333, 1, 9631), (70, 70, 3, 9632), (40, 400, 3, 9633)])
def test_fast_and_orb_in_one_call(ctx, h, w, ch, seed):
    """slamb200_fast_orb_compute = fastExtractor + extractDescriptor(ORB) with one frame upload; the
    resident descriptor set matches like an uploaded one."""
    from slam_indoor_code_b200.feature_matching import MatcherType
    frame = synth.textured_frame(h, w, seed, ch)
    pts, keep, desc, res = fe.fastExtractorAndDescribeORB(ctx, frame, 10, True, want_resident=True)
    ref_pts = c_oracle.fast_detect(frame, 10, True)
    rkeep, rdesc = c_oracle.orb_compute(frame, fe.to_orb_keypoints(ref_pts))
    assert np.array_equal(pts, ref_pts)
    assert np.array_equal(keep, rkeep.astype(bool)) and np.array_equal(desc, rdesc)
    assert res.n == len(rdesc)
    if len(rdesc) >= 2:
        T = ctx.upload(rdesc)
        assert np.array_equal(ctx.matchFeatures(res, T, MatcherType.ORB_BF, 0.99),
                              c_oracle.match_features(2, rdesc, rdesc, 0.99))
        T.free()
    res.free()
