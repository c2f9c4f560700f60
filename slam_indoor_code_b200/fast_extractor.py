"""Host mirror of the reference's fastExtractor (featureExtraction/fastExtractor.cpp:7-13):

    fastExtractor(srcImage, points, threshold = 10, suppression = true, TYPE_9_16)

FAST-9/16 keypoints of a BGR (or gray) frame on the B200 through the C ABI
(slamb200_fast_detect): same keypoints, same order, same responses as
cv::FastFeatureDetector.  No CPU fallback.
"""
import ctypes

import numpy as np

from ._capi import check, ptr


def fastExtractor(ctx, srcImage, threshold=10, suppression=True, max_points=None):
    """Returns [n, 3] float32 rows {x, y, response}: cv::KeyPoint(x, y, 7, -1, response) each, in
    OpenCV's order.  `max_points` bounds the output buffer (default: every pixel could be one)."""
    img = np.ascontiguousarray(srcImage, np.uint8)
    if img.ndim not in (2, 3):
        raise ValueError("fastExtractor: a rows x cols (x channels) uint8 image is expected")
    rows, cols = img.shape[:2]
    channels = 1 if img.ndim == 2 else img.shape[2]
    cap = int(max_points) if max_points is not None else max(rows * cols // 4, 1)
    while True:
        kps = np.zeros((max(cap, 1), 3), np.float32)
        n = ctypes.c_int(0)
        check(ctx._lib.slamb200_fast_detect(ctx._h, ptr(img), rows, cols, channels, img.strides[0],
                                            int(threshold), 1 if suppression else 0, ptr(kps), cap,
                                            ctypes.byref(n)))
        if n.value <= cap or max_points is not None:
            return kps[: min(n.value, cap)].copy()
        cap = n.value           # more corners than the default buffer: once more with the exact size


def to_orb_keypoints(points):
    """{x, y, response} rows -> {x, y, angle} rows for extractDescriptorORB: FAST keypoints carry
    angle -1 (cv::KeyPoint's default), which cv::ORB::compute uses as it is."""
    out = np.asarray(points, np.float32).copy()
    out[:, 2] = -1.0
    return out


def fastExtractorAndDescribeORB(ctx, frame, threshold=10, suppression=True, want_host=True,
                                want_resident=False):
    """fastExtractor(frame, points, threshold) followed by extractDescriptor(frame, points, ORB_BF,
    desc) in one call (slamb200_fast_orb_compute): the frame is uploaded once.  Returns
    (points [n, 3] {x, y, response}, keep mask [n] bool -- the points extractDescriptor leaves in
    `features`, descriptors [n_kept, 32] uint8 or None, resident DescriptorSet or None)."""
    from . import _capi
    from .feature_matching import DescriptorSet
    img = np.ascontiguousarray(frame, np.uint8)
    rows, cols = img.shape[:2]
    channels = 1 if img.ndim == 2 else img.shape[2]
    cap = max(rows * cols // 16 + 1024, 1)
    for _ in range(2):
        kps = np.zeros((cap, 3), np.float32)
        keep = np.zeros(cap, np.uint8)
        desc = np.zeros((cap, 32), np.uint8) if want_host else None
        n, k = ctypes.c_int(0), ctypes.c_int(0)
        h = ctypes.c_void_p()
        rc = ctx._lib.slamb200_fast_orb_compute(ctx._h, ptr(img), rows, cols, channels, img.strides[0],
                                                int(threshold), 1 if suppression else 0, ptr(kps), cap,
                                                ctypes.byref(n), ptr(keep), ptr(desc) if want_host else None,
                                                ctypes.byref(k), ctypes.byref(h) if want_resident else None)
        if rc != _capi.OK and n.value > cap:
            cap = n.value            # the frame has more corners than the first guess
            continue
        check(rc)
        break
    return (kps[: n.value].copy(), keep[: n.value].astype(bool), desc[: k.value].copy() if want_host else None,
            DescriptorSet(ctx, h, k.value, _capi.DESC_U8X32) if want_resident else None)
