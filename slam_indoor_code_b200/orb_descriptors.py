"""Host-side mirror of extractDescriptor's ORB branch over the libslamb200 C ABI.

Reference: src/mainModule/featureMatching/featureMatchingCPU.cpp:45-66 --
    extractDescriptor(Mat& frame, vector<KeyPoint>& features, int matcherType, Mat& desc)
with ORB_BF runs cv::ORB::create()->compute(frame, features, desc) on the FAST keypoints of
fastExtractor.cpp:7-13; compute() erases the keypoints closer than 31 px to the border from
`features`.  Here the descriptors are computed on the B200 and can stay there as a resident set.
"""
import ctypes

import numpy as np

from ._capi import check, ptr
from . import _capi
from .feature_matching import DescriptorSet


def keypoints_to_array(features):
    """cv2.KeyPoint list (or an [n, 3] array of x, y, angle) -> float32 [n, 3]."""
    if isinstance(features, np.ndarray):
        return np.ascontiguousarray(features, np.float32).reshape(-1, 3)
    return np.array([[k.pt[0], k.pt[1], k.angle] for k in features], np.float32).reshape(-1, 3)


def extractDescriptorORB(ctx, frame, features, want_host=True, want_resident=False):
    """Returns (keep mask [n] bool -- the rows of `features` compute() keeps, in order,
    descriptors [n_kept, 32] uint8 or None, resident DescriptorSet or None)."""
    frame = np.asarray(frame)
    if frame.dtype != np.uint8 or frame.ndim not in (2, 3) or (frame.ndim == 3 and frame.shape[2] not in (1, 3)):
        raise TypeError("frame must be CV_8UC1 or CV_8UC3")
    if frame.strides[-1] != 1 or (frame.ndim == 3 and frame.strides[1] != frame.shape[2]):
        frame = np.ascontiguousarray(frame)
    rows, cols = frame.shape[:2]
    ch = 1 if frame.ndim == 2 else frame.shape[2]
    kps = keypoints_to_array(features)
    n = kps.shape[0]
    keep = np.zeros(max(n, 1), np.uint8)
    desc = np.zeros((max(n, 1), 32), np.uint8) if want_host else None
    n_kept = ctypes.c_int(0)
    h = ctypes.c_void_p()
    check(ctx._lib.slamb200_orb_compute(ctx._h, ptr(frame), rows, cols, ch, frame.strides[0], ptr(kps), n,
                                        ptr(keep), ptr(desc), ctypes.byref(n_kept),
                                        ctypes.byref(h) if want_resident else None))
    k = n_kept.value
    return keep[:n].astype(bool), (desc[:k].copy() if want_host else None), \
        (DescriptorSet(ctx, h, k, _capi.DESC_U8X32) if want_resident else None)
