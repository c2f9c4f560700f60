"""Host-side mirror of extractDescriptor's SIFT branch over the libslamb200 C ABI.

Reference: src/mainModule/featureMatching/featureMatchingCPU.cpp:45-66 --
    extractDescriptor(Mat& frame, vector<KeyPoint>& features, int matcherType, Mat& desc)
with SIFT_BF / SIFT_FLANN runs cv::SIFT::create()->compute(frame, features, desc) on the FAST
keypoints of fastExtractor.cpp:7-13 (size 7, angle -1, octave 0).  Here the descriptors are computed
on the B200 and can stay there as a resident set -- no descriptor upload.  Held to a tolerance, not
to bit-exactness (include/slamb200.h): every element within 1 of OpenCV's, >= 99.9 % equal.
"""
import ctypes

import numpy as np

from ._capi import check, ptr
from . import _capi
from .feature_matching import DescriptorSet


def keypoints_to_array(features):
    """cv2.KeyPoint list (octave 0) or an [n, 4] array of x, y, size, angle -> float32 [n, 4]."""
    if isinstance(features, np.ndarray):
        return np.ascontiguousarray(features, np.float32).reshape(-1, 4)
    for k in features:
        if (k.octave & 255) != 0 or ((k.octave >> 8) & 255) != 0:
            raise ValueError("only octave-0 / layer-0 keypoints (what fastExtractor produces) are supported")
    return np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in features], np.float32).reshape(-1, 4)


def extractDescriptorSIFT(ctx, frame, features, want_host=True, want_resident=False):
    """Returns (descriptors [n, 128] float32 (integer valued) or None, resident DescriptorSet or None);
    `features` is left as it is (cv::SIFT::compute drops no keypoint)."""
    frame = np.asarray(frame)
    if frame.dtype != np.uint8 or frame.ndim not in (2, 3) or (frame.ndim == 3 and frame.shape[2] not in (1, 3)):
        raise TypeError("frame must be CV_8UC1 or CV_8UC3")
    if frame.strides[-1] != 1 or (frame.ndim == 3 and frame.strides[1] != frame.shape[2]):
        frame = np.ascontiguousarray(frame)
    rows, cols = frame.shape[:2]
    ch = 1 if frame.ndim == 2 else frame.shape[2]
    kps = keypoints_to_array(features)
    n = kps.shape[0]
    desc = np.zeros((max(n, 1), 128), np.float32) if want_host else None
    h = ctypes.c_void_p()
    check(ctx._lib.slamb200_sift_compute(ctx._h, ptr(frame), rows, cols, ch, frame.strides[0], ptr(kps), n,
                                         ptr(desc), ctypes.byref(h) if want_resident else None))
    return (desc[:n] if want_host else None), \
        (DescriptorSet(ctx, h, n, _capi.DESC_F32X128) if want_resident else None)
