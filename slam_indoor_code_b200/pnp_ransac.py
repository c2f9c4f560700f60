"""Host-side mirror of the solvePnPRansac seam over the libslamb200 C ABI.

The reference's call is
    solvePnPRansac(oldSpatialPointsForNewFrame, newFrameFeatureCoords, calibrationMatrix,
                   distortionCoeffs, rotationVector, motion)
at src/mainModule/cycleProcessing/mainCycle.cpp:155-159 -- vector<Point3f>, vector<Point2f>, 3x3 and
1x5 CV_64F Mats, every other argument at its OpenCV default (useExtrinsicGuess=false,
iterationsCount=100, reprojectionError=8.0, confidence=0.99, SOLVEPNP_ITERATIVE).  It runs every
frame of the steady-state loop.  The minimal solver (EPnP on 5 points; P3P when exactly 4 points are
given) and the final Levenberg-Marquardt refit stay on the CPU (OpenCV); the data-parallel inside --
projecting every 3-D point with every candidate pose and counting reprojection inliers -- runs on
the B200 (slamb200_score_pnp).  The control around it (fixed-seed cv::RNG subsets, the
count > max(best, modelPoints-1) update rule, RANSACUpdateNumIters) is ransac_host.ransac_run, so the
returned rvec, tvec and inlier list are bit-identical to cv::solvePnPRansac's -- checked in the tests.
"""
import ctypes

import numpy as np

from ._capi import check, ptr
from . import ransac_host


def _k4(K):
    K = np.asarray(K, np.float64)
    return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], np.float64) if K.shape == (3, 3) \
        else np.ascontiguousarray(K.reshape(4), np.float64)


def _dist(distCoeffs):
    return np.zeros(0, np.float64) if distCoeffs is None else \
        np.ascontiguousarray(distCoeffs, np.float64).reshape(-1)


def scorePnPHypotheses(ctx, objectPoints, imagePoints, K, distCoeffs, poses, reprojectionError=8.0,
                       model_points=5, want_all_masks=False):
    """poses: [H, 12] = rotation matrix (row-major) + tvec.  Returns counts[H], best index (-1:
    no model above model_points-1 inliers), best mask[M] (uint8 0/1), all masks or None."""
    obj = np.ascontiguousarray(objectPoints, np.float32).reshape(-1, 3)
    img = np.ascontiguousarray(imagePoints, np.float32).reshape(-1, 2)
    if obj.shape[0] != img.shape[0]:
        raise ValueError("objectPoints and imagePoints differ in size")
    K4, d = _k4(K), _dist(distCoeffs)
    poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 12)
    M, H = obj.shape[0], poses.shape[0]
    counts = np.zeros(max(H, 1), np.int32)
    best = ctypes.c_int32(-1)
    mask = np.zeros(max(M, 1), np.uint8)
    allm = np.zeros((max(H, 1), max(M, 1)), np.uint8) if want_all_masks else None
    check(ctx._lib.slamb200_score_pnp(ctx._h, ptr(obj), ptr(img), M, ptr(K4), ptr(d), d.size,
                                      ptr(poses), H, float(reprojectionError), int(model_points),
                                      ptr(counts), ctypes.byref(best), ptr(mask), ptr(allm)))
    return counts[:H], int(best.value), mask[:M], (allm[:H, :M] if allm is not None else None)


def scorePnPBatch(ctx, object_list, image_list, K, distCoeffs, poses, reprojectionError=8.0,
                  model_points=5):
    """P frames at once: ragged correspondence lists, poses of shape [P, H, 12]."""
    P = len(object_list)
    poses = np.ascontiguousarray(poses, np.float64).reshape(P, -1, 12)
    H = poses.shape[1]
    m_off = np.zeros(P + 1, np.int32)
    for p in range(P):
        m_off[p + 1] = m_off[p] + len(object_list[p])
    tot = int(m_off[P])
    obj = np.ascontiguousarray(np.concatenate(
        [np.asarray(a, np.float32).reshape(-1, 3) for a in object_list]) if tot else np.zeros((0, 3), np.float32))
    img = np.ascontiguousarray(np.concatenate(
        [np.asarray(a, np.float32).reshape(-1, 2) for a in image_list]) if tot else np.zeros((0, 2), np.float32))
    K4, d = _k4(K), _dist(distCoeffs)
    counts = np.zeros((P, max(H, 1)), np.int32)
    best = np.zeros(P, np.int32)
    mask = np.zeros(max(tot, 1), np.uint8)
    check(ctx._lib.slamb200_score_pnp_batch(ctx._h, P, ptr(obj), ptr(img), ptr(m_off), ptr(K4), ptr(d),
                                            d.size, ptr(poses), H, float(reprojectionError),
                                            int(model_points), ptr(counts), ptr(best), ptr(mask)))
    return counts[:, :H], best, [mask[m_off[p]: m_off[p + 1]].copy() for p in range(P)]


def _cv2_minimal_solver(obj, img, Kmat, dist, method):
    """The CPU minimal solver of PnPRansacCallback::runKernel: cv::solvePnP(EPNP | P3P) on the
    sampled correspondences.  Returns (solve(idx) -> [k, 12], lookup pose-bytes -> (rvec, tvec))."""
    import cv2
    seen = {}

    def solve(idx):
        ok, r, t = cv2.solvePnP(obj[idx], img[idx], Kmat, dist, flags=method)
        if not ok:
            return np.zeros((0, 12))
        R = cv2.Rodrigues(r)[0]
        m = np.concatenate([R.reshape(-1), t.reshape(-1)])
        seen[m.tobytes()] = (r, t)
        return m[None]
    return solve, seen


def solvePnPRansac(ctx, objectPoints, imagePoints, cameraMatrix, distCoeffs, iterationsCount=100,
                   reprojectionError=8.0, confidence=0.99, chunk=8, solver=None):
    """Drop-in for cv::solvePnPRansac(objectPoints, imagePoints, K, dist, rvec, tvec) as the
    reference calls it: returns (ok, rvec 3x1, tvec 3x1, inliers n x 1 int32)."""
    import cv2
    obj = np.ascontiguousarray(objectPoints, np.float32).reshape(-1, 3)
    img = np.ascontiguousarray(imagePoints, np.float32).reshape(-1, 2)
    Kmat = np.asarray(cameraMatrix, np.float64).reshape(3, 3)
    d = _dist(distCoeffs)
    dist = d if d.size else None
    count = obj.shape[0]
    if count < 4 or count != img.shape[0]:
        raise ValueError("solvePnPRansac needs >= 4 paired points")  # CV_Assert in the reference
    mp, method = (4, cv2.SOLVEPNP_P3P) if count == 4 else (5, cv2.SOLVEPNP_EPNP)
    solve, seen = solver or _cv2_minimal_solver(obj, img, Kmat, dist, method)

    def score(models):
        return scorePnPHypotheses(ctx, obj, img, Kmat, d, models, reprojectionError, mp)[0]

    if count == mp:  # cv::solvePnPRansac's short-cut: the minimal solver on all points, no refit
        ok, r, t = cv2.solvePnP(obj, img, Kmat, dist, flags=method)
        return (True, r, t, np.arange(count, dtype=np.int32).reshape(-1, 1)) if ok \
            else (False, None, None, None)
    best, _, _ = ransac_host.ransac_run(count, mp, confidence, iterationsCount, solve, score, chunk)
    mask = None if best is None else \
        scorePnPHypotheses(ctx, obj, img, Kmat, d, best[None], reprojectionError, mp, True)[3][0]
    if best is None:
        return False, None, None, None
    inl = np.nonzero(mask)[0]
    r0, t0 = seen[best.tobytes()]
    # the refit on the inliers starts from the RANSAC model (useExtrinsicGuess is forced on for
    # SOLVEPNP_ITERATIVE inside cv::solvePnPRansac)
    ok, r, t = cv2.solvePnP(obj[inl].astype(np.float64), img[inl].astype(np.float64), Kmat, dist,
                            r0.copy(), t0.copy(), True, cv2.SOLVEPNP_ITERATIVE)
    if not ok:
        return False, r, t, inl.astype(np.int32).reshape(-1, 1)
    return True, r, t, inl.astype(np.int32).reshape(-1, 1)
