"""Host-side mirror of the RANSAC essential-matrix scoring seam over the libslamb200 C ABI.

The reference's seam is the single call
    findEssentialMat(points1, points2, K, RANSAC, RPRANSACProb, RPRANSACThreshold, mask)
at src/mainModule/translation/cameraTranslation.cpp:41-46.  The 5-point minimal solver and
recoverPose stay on the CPU (OpenCV); what moves to the GPU is the data-parallel inside of the
RANSAC loop: scoring every candidate essential matrix against every match and producing the
winner's N x 1 uchar mask (the one the reference logs at :53-54).
"""
import ctypes

import numpy as np

from ._capi import check, ptr


def scoreEssentialHypotheses(ctx, points1, points2, K, E, RPRANSACThreshold=5.0,
                             want_all_masks=False):
    """counts[H], best index (-1: none above 4 inliers), best mask[M] (uint8 0/1), all masks."""
    p1 = np.ascontiguousarray(points1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(points2, np.float32).reshape(-1, 2)
    if p1.shape != p2.shape:
        raise ValueError("points1 and points2 differ in size")
    K = np.asarray(K, np.float64)
    K4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], np.float64) if K.shape == (3, 3) \
        else np.ascontiguousarray(K.reshape(4), np.float64)
    E = np.ascontiguousarray(E, np.float64).reshape(-1, 9)
    M, H = p1.shape[0], E.shape[0]
    counts = np.zeros(max(H, 1), np.int32)
    best = ctypes.c_int32(-1)
    mask = np.zeros(max(M, 1), np.uint8)
    allm = np.zeros((max(H, 1), max(M, 1)), np.uint8) if want_all_masks else None
    check(ctx._lib.slamb200_score_essential(ctx._h, ptr(p1), ptr(p2), M, ptr(K4), ptr(E), H,
                                            float(RPRANSACThreshold), ptr(counts),
                                            ctypes.byref(best), ptr(mask), ptr(allm)))
    return counts[:H], int(best.value), mask[:M], (allm[:H, :M] if allm is not None else None)


def scoreEssentialBatch(ctx, points1_list, points2_list, K4, E, RPRANSACThreshold=5.0):
    """P pairs at once: ragged match lists, E of shape [P, H, 9]."""
    P = len(points1_list)
    E = np.ascontiguousarray(E, np.float64).reshape(P, -1, 9)
    H = E.shape[1]
    m_off = np.zeros(P + 1, np.int32)
    for p in range(P):
        m_off[p + 1] = m_off[p] + len(points1_list[p])
    tot = int(m_off[P])
    p1 = np.concatenate([np.asarray(a, np.float32).reshape(-1, 2) for a in points1_list]) \
        if tot else np.zeros((0, 2), np.float32)
    p2 = np.concatenate([np.asarray(a, np.float32).reshape(-1, 2) for a in points2_list]) \
        if tot else np.zeros((0, 2), np.float32)
    p1, p2 = np.ascontiguousarray(p1), np.ascontiguousarray(p2)
    K4 = np.ascontiguousarray(K4, np.float64)
    counts = np.zeros((P, max(H, 1)), np.int32)
    best = np.zeros(P, np.int32)
    mask = np.zeros(max(tot, 1), np.uint8)
    check(ctx._lib.slamb200_score_essential_batch(ctx._h, P, ptr(p1), ptr(p2), ptr(m_off),
                                                  ptr(K4), ptr(E), H, float(RPRANSACThreshold),
                                                  ptr(counts), ptr(best), ptr(mask)))
    masks = [mask[m_off[p]: m_off[p + 1]].copy() for p in range(P)]
    return counts[:, :H], best, masks


def scoreBatchEnqueue(ctx, query_kps, train_kps, K4, E, RPRANSACThreshold=5.0, stream=None,
                      E_device_ptr=None, H=None):
    """Chains onto the last matchBatchEnqueue: gathers the matched keypoints on the device
    (getKeyPointCoordsFromFramePair) and scores H hypotheses per pair.  Enqueue-only.  E is a host
    array [P, H, 9], or (E_device_ptr, H) names P*H*9 doubles already resident on the device."""
    P = len(train_kps)
    K4 = np.ascontiguousarray(K4, np.float64)
    arr = (ctypes.c_void_p * max(P, 1))(*[t._h for t in train_kps])
    if E_device_ptr is not None:
        e_ptr = ctypes.c_void_p(E_device_ptr)
    else:
        E = np.ascontiguousarray(E, np.float64).reshape(P, -1, 9)
        H = E.shape[1]
        e_ptr = ptr(E)
    check(ctx._lib.slamb200_score_batch_enqueue(ctx._h, query_kps._h, arr, ptr(K4), e_ptr, int(H),
                                                float(RPRANSACThreshold),
                                                ctypes.c_void_p(stream or 0)))
    ctx._last_score = (P, int(H))


def batchScoresFetch(ctx, stream=None):
    P, H = ctx._last_score
    cap = ctx._last[1]
    counts = np.zeros((P, H), np.int32)
    best = np.zeros(P, np.int32)
    mask = np.zeros((P, cap), np.uint8)
    check(ctx._lib.slamb200_batch_scores_fetch(ctx._h, ptr(counts), ptr(best), ptr(mask), cap,
                                               ctypes.c_void_p(stream or 0)))
    return counts, best, mask


def _cv2_five_point(K4):
    """The CPU minimal solver: cv::findEssentialMat on exactly five matches returns the stacked
    3k x 3 candidates of the 5-point algorithm (it stays on the CPU)."""
    import cv2
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]], np.float64)

    def solve(p1, p2):
        E = cv2.findEssentialMat(p1, p2, Kmat, cv2.RANSAC, 0.999, 1.0)[0]
        return np.zeros((0, 9)) if E is None else np.asarray(E, np.float64).reshape(-1, 9)
    return solve


def findEssentialMat(ctx, points1, points2, K, RPRANSACProb=0.999, RPRANSACThreshold=5.0,
                     five_point_fn=None, chunk=32):
    """Drop-in for findEssentialMat(points1, points2, K, RANSAC, prob, threshold, mask)
    (cameraTranslation.cpp:41-46): returns (E 3x3 or None, mask N x 1 uint8 or None), bit-identical
    to OpenCV's.  The RANSAC control runs on the host (ransac_host.py), the 5-point solver on the
    CPU, every candidate model is scored on the B200."""
    from . import ransac_host
    K = np.asarray(K, np.float64)
    K4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], np.float64) if K.shape == (3, 3) \
        else np.ascontiguousarray(K.reshape(4), np.float64)
    p1 = np.ascontiguousarray(points1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(points2, np.float32).reshape(-1, 2)
    solve = five_point_fn or _cv2_five_point(K4)

    def score(models):
        return scoreEssentialHypotheses(ctx, p1, p2, K4, models, RPRANSACThreshold)[0]

    E, mask, _, _ = ransac_host.ransac_essential(p1, p2, K4, RPRANSACProb, RPRANSACThreshold, solve,
                                                 score, chunk=chunk)
    if E is None:
        return None, None
    if mask is None:
        mask = scoreEssentialHypotheses(ctx, p1, p2, K4, E.reshape(1, 9), RPRANSACThreshold)[2]
    return E.reshape(3, 3) if E.size == 9 else E.reshape(-1, 3), mask.reshape(-1, 1)
