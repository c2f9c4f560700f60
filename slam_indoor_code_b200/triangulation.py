"""Host-side mirror of the triangulation seam over the libslamb200 C ABI.

Reference: src/mainModule/triangulation/triangulate.cpp -- triangulationWrapper (:57-72, the
per-point DLT + SVD of reconstructPointsFor3D :17-55), reconstruct (:74-89) and
convertHomogeneousPointsMatrixToSpatialPointsVector (:91-108).
"""
import numpy as np

from ._capi import check, ptr


def triangulationWrapper(ctx, projPoints1, projPoints2, matr1, matr2, want_spatial=False):
    """points4D [4, N] (homogeneous, one column per match) like the reference's OutputArray;
    with want_spatial also the N x 3 points (X, Y, Z) * (1 / W)."""
    p1 = np.ascontiguousarray(projPoints1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(projPoints2, np.float32).reshape(-1, 2)
    if p1.shape != p2.shape:
        raise ValueError("projPoints1 and projPoints2 differ in size")
    P1 = np.ascontiguousarray(matr1, np.float64).reshape(3, 4)
    P2 = np.ascontiguousarray(matr2, np.float64).reshape(3, 4)
    M = p1.shape[0]
    X4 = np.zeros((4, max(M, 1)), np.float64)
    X3 = np.zeros((max(M, 1), 3), np.float64)
    check(ctx._lib.slamb200_triangulate(ctx._h, ptr(P1), ptr(P2), ptr(p1), ptr(p2), M, ptr(X4), ptr(X3)))
    if M == 0:
        X4, X3 = X4[:, :0], X3[:0]
    return (X4, X3) if want_spatial else X4


def reconstruct(ctx, calibration, rotation1, transition1, rotation2, transition2, points1, points2):
    """triangulate.cpp:74-89: projection matrices K*[R|t] on the host (two 3x3 by 3x4 products),
    every match triangulated on the device; returns the N x 3 spatial points."""
    K = np.asarray(calibration, np.float64).reshape(3, 3)
    P1 = K @ np.hstack([np.asarray(rotation1, np.float64).reshape(3, 3),
                        np.asarray(transition1, np.float64).reshape(3, 1)])
    P2 = K @ np.hstack([np.asarray(rotation2, np.float64).reshape(3, 3),
                        np.asarray(transition2, np.float64).reshape(3, 1)])
    return triangulationWrapper(ctx, points1, points2, P1, P2, want_spatial=True)[1]
