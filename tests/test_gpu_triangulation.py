"""GPU parity tests for the next row SURVEY.md 8f-4: linear triangulation through the C ABI.

Checker: the CPU oracle (OpenCV's one-sided Jacobi sequence, pinned to cv2.SVDecomp /
cv2.triangulatePoints within 1e-14).  Floating point: the tolerance is 1e-12 relative on the
homogeneous vector (unit norm) and 1e-10 relative on X/W for well-conditioned points -- libm and
CUDA differ in the last bit of hypot/sqrt, and the Jacobi sweeps amplify that by a few ulps.
Mirrors reconstructPointsFor3D at triangulation/triangulate.cpp:17-55.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, synth
from slam_indoor_code_b200 import triangulation as tri

K4 = np.array(synth.SAMSUNG_HV_4K)
K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])


def _scene(m, seed, outliers=0.0):
    p1, p2, R, t = synth.two_view(m, seed, outliers=outliers)
    P1 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = K @ np.hstack([R, t.reshape(3, 1)])
    return p1, p2, P1, P2, R, t


def _same_direction(a, b, tol):
    """Unit 4-vectors up to sign (the SVD's sign is arbitrary; it cancels in X/W)."""
    s = np.sign(np.sum(a * b, axis=0))
    return np.max(np.abs(a - b * s)) <= tol


@pytest.mark.parametrize("m,seed,outl", [(5000, 31, 0.0), (1237, 32, 0.3), (1, 33, 0.0), (129, 34, 0.0)])
def test_triangulate_vs_oracle(ctx, m, seed, outl):
    p1, p2, P1, P2, _, _ = _scene(m, seed, outl)
    X4, X3 = tri.triangulationWrapper(ctx, p1, p2, P1, P2, want_spatial=True)
    R4, R3 = c_oracle.triangulate(P1, P2, p1, p2)
    assert X4.shape == (4, m) and X3.shape == (m, 3)
    assert _same_direction(X4, R4, 1e-12)
    well = np.abs(R4[3]) > 1e-6
    assert well.mean() > 0.5
    assert np.allclose(X3[well], R3[well], rtol=1e-10, atol=0)


def test_triangulate_vs_cv2_and_ground_truth(ctx):
    cv2 = pytest.importorskip("cv2")
    p1, p2, P1, P2, R, t = _scene(3000, 41)
    X4 = tri.triangulationWrapper(ctx, p1, p2, P1, P2)
    ref = cv2.triangulatePoints(P1, P2, p1.T.astype(np.float64), p2.T.astype(np.float64))
    assert _same_direction(X4, ref, 1e-12)
    # the reconstructed points reproject onto the observations (0.7 px noise in the scene)
    X = tri.reconstruct(ctx, K, np.eye(3), np.zeros(3), R, t, p1, p2)
    uv = (K @ X.T).T
    uv = uv[:, :2] / uv[:, 2:]
    assert np.median(np.linalg.norm(uv - p1, axis=1)) < 2.0


def test_triangulate_edges(ctx):
    p1, p2, P1, P2, _, _ = _scene(10, 51)
    X4 = tri.triangulationWrapper(ctx, p1[:0], p2[:0], P1, P2)
    assert X4.shape == (4, 0)
    # identical views: the system is rank deficient; the result still equals the CPU's direction
    # whenever the two smallest singular values are separated
    X4, X3 = tri.triangulationWrapper(ctx, p1, p1, P1, P1, want_spatial=True)
    assert np.all(np.isfinite(X4))
    assert np.allclose(np.linalg.norm(X4, axis=0), 1.0, atol=1e-12)
    with pytest.raises(ValueError):
        tri.triangulationWrapper(ctx, p1, p2[:5], P1, P2)
