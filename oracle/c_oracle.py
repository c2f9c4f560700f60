"""ctypes binding of oracle/liboracle.so (corr_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"),
                   ("distance", "<f4")])


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "corr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_threads(n):
    return lib().oracle_set_threads(int(n))


def l2_knn2(Q, T):
    Q = np.ascontiguousarray(Q, np.float32).reshape(-1, 128) if Q.size else np.zeros((0, 128), np.float32)
    T = np.ascontiguousarray(T, np.float32).reshape(-1, 128) if T.size else np.zeros((0, 128), np.float32)
    idx = np.full((Q.shape[0], 2), -1, np.int32)
    dist = np.zeros((Q.shape[0], 2), np.float32)
    rc = lib().oracle_l2_knn2(_p(Q), Q.shape[0], _p(T), T.shape[0], 128, _p(idx), _p(dist))
    assert rc == 0
    return idx, dist


def l1_knn2(Q, T):
    Q = np.ascontiguousarray(Q, np.float32).reshape(-1, 128) if Q.size else np.zeros((0, 128), np.float32)
    T = np.ascontiguousarray(T, np.float32).reshape(-1, 128) if T.size else np.zeros((0, 128), np.float32)
    idx = np.full((Q.shape[0], 2), -1, np.int32)
    dist = np.zeros((Q.shape[0], 2), np.float32)
    rc = lib().oracle_l1_knn2(_p(Q), Q.shape[0], _p(T), T.shape[0], 128, _p(idx), _p(dist))
    assert rc == 0
    return idx, dist


def hamming_knn2(Q, T):
    Q = np.ascontiguousarray(Q, np.uint8).reshape(-1, 32) if Q.size else np.zeros((0, 32), np.uint8)
    T = np.ascontiguousarray(T, np.uint8).reshape(-1, 32) if T.size else np.zeros((0, 32), np.uint8)
    idx = np.full((Q.shape[0], 2), -1, np.int32)
    dist = np.zeros((Q.shape[0], 2), np.float32)
    rc = lib().oracle_hamming_knn2(_p(Q), Q.shape[0], _p(T), T.shape[0], 32, _p(idx), _p(dist))
    assert rc == 0
    return idx, dist


def ratio_test(idx, dist, ratio):
    idx = np.ascontiguousarray(idx, np.int32)
    dist = np.ascontiguousarray(dist, np.float32)
    nq = idx.shape[0]
    out = np.zeros(max(nq, 1), DMATCH)
    n = ctypes.c_int(0)
    rc = lib().oracle_ratio_test(_p(idx), _p(dist), nq, ctypes.c_double(ratio), _p(out),
                                 ctypes.byref(n))
    assert rc == 0
    return out[: n.value].copy()


def match_features(matcher, Q, T, ratio):
    """matchFeatures restatement: returns the good matches as a DMATCH structured array."""
    if matcher in (0, 1):
        idx, dist = l2_knn2(Q, T)
    elif matcher == 2:
        idx, dist = hamming_knn2(Q, T)
    elif matcher == 3:
        idx, dist = l1_knn2(Q, T)
    else:
        raise ValueError("bad matcher type")  # reference: throw std::exception()
    return ratio_test(idx, dist, ratio)


def gather_points(prev_xy, next_xy, matches):
    prev_xy = np.ascontiguousarray(prev_xy, np.float32)
    next_xy = np.ascontiguousarray(next_xy, np.float32)
    m = np.ascontiguousarray(matches, DMATCH)
    p1 = np.zeros((len(m), 2), np.float32)
    p2 = np.zeros((len(m), 2), np.float32)
    lib().oracle_gather_points(_p(prev_xy), _p(next_xy), _p(m), len(m), _p(p1), _p(p2))
    return p1, p2


def score_essential(pts1, pts2, K4, E, threshold_px, want_all_masks=False):
    pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    K4 = np.ascontiguousarray(K4, np.float64)
    E = np.ascontiguousarray(E, np.float64).reshape(-1, 9)
    M, H = pts1.shape[0], E.shape[0]
    counts = np.zeros(H, np.int32)
    best = ctypes.c_int32(-1)
    best_mask = np.zeros(M, np.uint8)
    allm = np.zeros((H, M), np.uint8) if want_all_masks else None
    rc = lib().oracle_score_essential(_p(pts1), _p(pts2), M, _p(K4), _p(E), H,
                                      ctypes.c_double(threshold_px), _p(counts),
                                      ctypes.byref(best), _p(best_mask),
                                      _p(allm) if allm is not None else None)
    assert rc == 0
    return counts, int(best.value), best_mask, allm


def sampson_errors(pts1, pts2, K4, E):
    pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    K4 = np.ascontiguousarray(K4, np.float64)
    E = np.ascontiguousarray(E, np.float64).reshape(-1, 9)
    err = np.zeros((E.shape[0], pts1.shape[0]), np.float32)
    rc = lib().oracle_sampson_errors(_p(pts1), _p(pts2), pts1.shape[0], _p(K4), _p(E),
                                     E.shape[0], _p(err))
    assert rc == 0
    return err


def _dist(dist):
    d = np.zeros(0, np.float64) if dist is None else np.ascontiguousarray(dist, np.float64).reshape(-1)
    if d.size > 12:
        assert not d[12:].any(), "tilt coefficients are outside the oracle"
        d = d[:12].copy()
    return d


def project_points(obj, K4, dist, pose):
    """cv::projectPoints restatement: pose = R (row-major 9) + t (3)."""
    obj = np.ascontiguousarray(obj, np.float32).reshape(-1, 3)
    K4 = np.ascontiguousarray(K4, np.float64)
    d = _dist(dist)
    pose = np.ascontiguousarray(pose, np.float64).reshape(12)
    uv = np.zeros((obj.shape[0], 2), np.float32)
    rc = lib().oracle_project_points(_p(obj), obj.shape[0], _p(K4), _p(d), d.size, _p(pose), _p(uv))
    assert rc == 0
    return uv


def score_pnp(obj, img, K4, dist, poses, reproj_err, model_points=5, want_all_masks=False):
    obj = np.ascontiguousarray(obj, np.float32).reshape(-1, 3)
    img = np.ascontiguousarray(img, np.float32).reshape(-1, 2)
    K4 = np.ascontiguousarray(K4, np.float64)
    d = _dist(dist)
    poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 12)
    M, H = obj.shape[0], poses.shape[0]
    counts = np.zeros(H, np.int32)
    best = ctypes.c_int32(-1)
    best_mask = np.zeros(M, np.uint8)
    allm = np.zeros((H, M), np.uint8) if want_all_masks else None
    rc = lib().oracle_score_pnp(_p(obj), _p(img), M, _p(K4), _p(d), d.size, _p(poses), H,
                                ctypes.c_double(reproj_err), int(model_points), _p(counts),
                                ctypes.byref(best), _p(best_mask),
                                _p(allm) if allm is not None else None)
    assert rc == 0
    return counts, int(best.value), best_mask, allm


def svd4_null_vector(A):
    A = np.ascontiguousarray(A, np.float64).reshape(4, 4)
    v, w = np.zeros(4), np.zeros(4)
    lib().oracle_svd4_null_vector(_p(A), _p(v), _p(w))
    return v, w


def triangulate(P1, P2, pts1, pts2):
    """reconstructPointsFor3D: returns points4d [4, M] and points3d [M, 3]."""
    P1 = np.ascontiguousarray(P1, np.float64).reshape(3, 4)
    P2 = np.ascontiguousarray(P2, np.float64).reshape(3, 4)
    pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    M = pts1.shape[0]
    X4 = np.zeros((4, M))
    X3 = np.zeros((M, 3))
    rc = lib().oracle_triangulate(_p(P1), _p(P2), _p(pts1), _p(pts2), M, _p(X4), _p(X3))
    assert rc == 0
    return X4, X3


def orb_blur(image):
    """ORB's working image: gray conversion (BGR input) + the float 7x7 Gaussian, as u8."""
    image = np.ascontiguousarray(image, np.uint8)
    rows, cols = image.shape[:2]
    ch = 1 if image.ndim == 2 else image.shape[2]
    gray = np.zeros((rows, cols), np.uint8)
    blur = np.zeros((rows, cols), np.uint8)
    rc = lib().oracle_orb_blur(_p(image), rows, cols, ch, ctypes.c_size_t(image.strides[0]), _p(gray), _p(blur))
    assert rc == 0
    return gray, blur


def orb_compute(image, kps):
    """cv::ORB::compute on given keypoints: kps [n, 3] = x, y, angle (degrees; FAST gives -1).
    Returns (keep mask [n] uint8, descriptors [n_kept, 32] uint8)."""
    image = np.ascontiguousarray(image, np.uint8)
    rows, cols = image.shape[:2]
    ch = 1 if image.ndim == 2 else image.shape[2]
    kps = np.ascontiguousarray(kps, np.float32).reshape(-1, 3)
    n = kps.shape[0]
    keep = np.zeros(max(n, 1), np.uint8)
    desc = np.zeros((max(n, 1), 32), np.uint8)
    kept = lib().oracle_orb_compute(_p(image), rows, cols, ch, ctypes.c_size_t(image.strides[0]), _p(kps), n,
                                    _p(keep), _p(desc))
    assert kept >= 0
    return keep[:n], desc[:kept].copy()


def fast_detect(image, threshold=10, nonmax=True, want_scores=False):
    """FastFeatureDetector::create(threshold, nonmax, TYPE_9_16)->detect(image): [n, 3] float32 rows
    of x, y, response in OpenCV's order (row by row, left to right)."""
    image = np.ascontiguousarray(image, np.uint8)
    rows, cols = image.shape[:2]
    ch = 1 if image.ndim == 2 else image.shape[2]
    cap = max(rows * cols, 1)
    kp = np.zeros((cap, 3), np.float32)
    score = np.zeros((rows, cols), np.uint8)
    n = lib().oracle_fast_detect(_p(image), rows, cols, ch, ctypes.c_size_t(image.strides[0]), int(threshold),
                                 1 if nonmax else 0, _p(kp), cap, _p(score))
    assert n >= 0
    return (kp[:n].copy(), score) if want_scores else kp[:n].copy()


def sift_base(image):
    """SIFT's working image for octave-0 keypoints: gray -> float -> 13-tap Gaussian (sigma
    sqrt(1.6^2 - 0.5^2)); equals cv2.GaussianBlur on the float gray frame bit for bit."""
    image = np.ascontiguousarray(image, np.uint8)
    rows, cols = image.shape[:2]
    ch = 1 if image.ndim == 2 else image.shape[2]
    base = np.zeros((rows, cols), np.float32)
    rc = lib().oracle_sift_base(_p(image), rows, cols, ch, ctypes.c_size_t(image.strides[0]), _p(base))
    assert rc == 0
    return base


def sift_compute(image, kps, raw=False):
    """cv::SIFT::compute on given octave-0 keypoints: kps [n, 4] = x, y, size, angle (FAST gives
    size 7, angle -1).  Returns descriptors [n, 128] float32 (integer valued unless raw)."""
    image = np.ascontiguousarray(image, np.uint8)
    rows, cols = image.shape[:2]
    ch = 1 if image.ndim == 2 else image.shape[2]
    kps = np.ascontiguousarray(kps, np.float32).reshape(-1, 4)
    n = kps.shape[0]
    desc = np.zeros((max(n, 1), 128), np.float32)
    rc = lib().oracle_sift_compute(_p(image), rows, cols, ch, ctypes.c_size_t(image.strides[0]), _p(kps), n,
                                   _p(desc), 1 if raw else 0)
    assert rc == 0
    return desc[:n]
