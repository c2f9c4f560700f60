"""CPU: pins the oracle against cv2 live (the reference's arithmetic owner) on fresh seeds.

Skipped when cv2 is not importable; the committed fixtures (test_oracle_golden.py) cover that
case.  This is where "the reference's own implementation" enters: the reference cannot be
compiled in this image (OpenCV C++ dev files absent), and its hot path is these OpenCV calls.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import c_oracle, synth
from oracle.gen_golden import cv_knn2


@pytest.mark.parametrize("seed", [11, 12])
def test_l2_int_and_float_bit_exact(seed):
    for q, t in (synth.sift_pair(400, 500, seed), synth.float_pair(300, 350, seed)):
        gi, gd, _ = cv_knn2(q, t, cv2.NORM_L2)
        idx, dist = c_oracle.l2_knn2(q, t)
        assert np.array_equal(idx, gi)
        assert np.array_equal(dist.view(np.int32), gd.view(np.int32))


def test_hamming_bit_exact():
    q, t = synth.orb_pair(500, 600, 21)
    gi, gd, _ = cv_knn2(q, t, cv2.NORM_HAMMING)
    idx, dist = c_oracle.hamming_knn2(q, t)
    assert np.array_equal(idx, gi) and np.array_equal(dist, gd)


def test_good_matches_equal_reference_loop():
    """matchFeatures end to end: knnMatch + the getGoodMatches loop written on cv2's own lists."""
    q, t = synth.sift_pair(500, 600, 31)
    res = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, 2)
    good = [r[0] for r in res if len(r) and r[0].distance < 0.7 * r[1].distance]
    got = c_oracle.match_features(0, q, t, 0.7)
    assert [m.queryIdx for m in good] == list(got["queryIdx"])
    assert [m.trainIdx for m in good] == list(got["trainIdx"])
    assert np.array_equal(np.array([m.distance for m in good], np.float32), got["distance"])
    assert len(good) > 50


@pytest.mark.parametrize("seed", [9100, 9101, 9102])
def test_find_essential_mat_mask(seed):
    p1, p2, _, _ = synth.two_view(1200, seed, noise_px=1.5)
    K4 = np.array(synth.SAMSUNG_HV_4K)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    E, mask = cv2.findEssentialMat(p1, p2, Kmat, cv2.RANSAC, 0.999, 5.0)
    E = np.asarray(E, np.float64).reshape(-1, 9)[:1]
    counts, best, m, _ = c_oracle.score_essential(p1, p2, K4, E, 5.0)
    assert best == 0 and np.array_equal(m, mask.reshape(-1))


@pytest.mark.parametrize("seed,dv", [(9200, 0), (9201, 1), (9202, 2), (9203, 3)])
def test_project_points_bit_exact(seed, dv):
    """cv::projectPoints (the call inside PnPRansacCallback::computeError) restated."""
    dist = (np.array(synth.REF_DIST5), None, np.zeros(5),
            np.array([0.2, -0.4, 0.002, -0.001, 0.1, 0.02, -0.03, 0.004, 2e-3, -1e-3, 5e-4, 2e-4]))[dv]
    obj, img, R, t = synth.pnp_scene(3000, seed)
    K4 = np.array(synth.SAMSUNG_HV_4K)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    rng = np.random.default_rng(seed)
    for _ in range(4):
        rvec, tvec = rng.normal(0, 0.3, 3), rng.normal(0, 0.5, 3)
        want = cv2.projectPoints(obj, rvec, tvec, Kmat, dist)[0].reshape(-1, 2)
        pose = np.concatenate([cv2.Rodrigues(rvec)[0].reshape(-1), tvec])
        got = c_oracle.project_points(obj, K4, dist, pose)
        assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))


@pytest.mark.parametrize("m,seed", [(700, 9300), (60, 9301), (5, 9302), (4, 9303), (2500, 9304)])
def test_solve_pnp_ransac_control_and_scoring(m, seed):
    """cv2.solvePnPRansac rebuilt from its parts -- cv::RNG subsets, EPnP/P3P minimal solver (cv2),
    the oracle's inlier counts, the update rule, the refit -- returns cv2's rvec, tvec, inliers."""
    from slam_indoor_code_b200 import ransac_host
    obj, img, _, _ = synth.pnp_scene(m, seed, outliers=0.3 if m > 10 else 0.0)
    K4 = np.array(synth.SAMSUNG_HV_4K)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    dist = np.array(synth.REF_DIST5)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(obj, img, Kmat, dist)
    mp, method = (4, cv2.SOLVEPNP_P3P) if m == 4 else (5, cv2.SOLVEPNP_EPNP)
    seen = {}

    def solve(idx):
        k, r, t = cv2.solvePnP(obj[idx], img[idx], Kmat, dist, flags=method)
        if not k:
            return np.zeros((0, 12))
        model = np.concatenate([cv2.Rodrigues(r)[0].reshape(-1), t.reshape(-1)])
        seen[model.tobytes()] = (r, t)
        return model[None]

    if m == mp:  # solvePnPRansac short-cut: the minimal solver on all points, no RANSAC, no refit
        k, r, t = cv2.solvePnP(obj, img, Kmat, dist, flags=method)
        assert ok and k and np.array_equal(r, rvec) and np.array_equal(t, tvec)
        assert np.array_equal(inl.reshape(-1), np.arange(m))
        return
    else:
        best, _, _ = ransac_host.ransac_run(m, mp, 0.99, 100, solve,
                                            lambda mod: c_oracle.score_pnp(obj, img, K4, dist, mod, 8.0, mp)[0])
        mask = c_oracle.score_pnp(obj, img, K4, dist, best[None], 8.0, mp, True)[3][0]
    got_inl = np.nonzero(mask)[0]
    assert ok and np.array_equal(got_inl, inl.reshape(-1))
    r0, t0 = seen[best.tobytes()]
    k, r, t = cv2.solvePnP(obj[got_inl].astype(np.float64), img[got_inl].astype(np.float64), Kmat, dist,
                           r0.copy(), t0.copy(), True, cv2.SOLVEPNP_ITERATIVE)
    assert k and np.array_equal(r, rvec) and np.array_equal(t, tvec)


@pytest.mark.parametrize("seed", [21, 22])
def test_l1_int_and_float_bit_exact(seed):
    """BFMatcher(NORM_L1): s += |d0| + |d1| + |d2| + |d3| per group of four, in float."""
    for q, t in (synth.sift_pair(257, 513, seed), synth.float_pair(257, 513, seed)):
        idx, dist, _ = cv_knn2(q, t, cv2.NORM_L1)
        oi, od = c_oracle.l1_knn2(q, t)
        assert np.array_equal(idx, oi) and np.array_equal(dist.view(np.int32), od.view(np.int32))


def test_triangulation_oracle_vs_cv2():
    """OpenCV's 4x4 Jacobi SVD (row 3 of Vt) and cv2.triangulatePoints, within 1e-14: the rotation
    sequence is restated, libm's hypot/sqrt and the build's contraction choices are not."""
    rng = np.random.default_rng(5)
    for i in range(300):
        A = rng.normal(size=(4, 4)) * rng.choice([1, 100, 1e-3])
        w, u, vt = cv2.SVDecomp(A)
        v, wo = c_oracle.svd4_null_vector(A)
        assert np.allclose(wo, w.reshape(-1), rtol=1e-13, atol=1e-300)
        if w[2, 0] - w[3, 0] > 1e-3 * w[0, 0]:
            assert np.max(np.abs(v - vt[3])) < 1e-13
    K4 = np.array(synth.SAMSUNG_HV_4K)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    p1, p2, R, t = synth.two_view(2000, 11, outliers=0.0)
    P1 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = K @ np.hstack([R, t.reshape(3, 1)])
    X4, X3 = c_oracle.triangulate(P1, P2, p1, p2)
    ref = cv2.triangulatePoints(P1, P2, p1.T.astype(np.float64), p2.T.astype(np.float64))
    assert np.max(np.abs(ref - X4)) < 1e-14
    assert np.allclose(X3, (ref[:3] / ref[3]).T, rtol=1e-11)


@pytest.mark.parametrize("h,w,ch,seed", [(480, 640, 3, 9400), (301, 457, 1, 9401)])
def test_orb_compute_bit_exact(h, w, ch, seed):
    """cv2.ORB.compute on FAST keypoints (the reference's sequence) and on oriented keypoints:
    gray conversion, ORB's float blur, the recovered comparison pattern, rotation, border filter."""
    frame = synth.textured_frame(h, w, seed, ch)
    kps = cv2.FastFeatureDetector_create(10, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(frame)[:6000]
    rng = np.random.default_rng(seed)
    for k in kps[::2]:
        k.angle = float(rng.uniform(0, 360))
    kept, desc = cv2.ORB_create().compute(frame, kps)
    keep, got = c_oracle.orb_compute(frame, np.array([[k.pt[0], k.pt[1], k.angle] for k in kps], np.float32))
    assert keep.sum() == len(kept) > 500
    assert np.array_equal(got, desc)
    gray, blur = c_oracle.orb_blur(frame)
    if ch == 3:
        assert np.array_equal(gray, cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY))
    k32 = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    ref = cv2.sepFilter2D(gray, cv2.CV_8U, k32, k32, borderType=cv2.BORDER_REFLECT_101)
    # the float blur depends on the FMA dispatch of the OpenCV build: identical here, and never
    # more than one level apart on the rare pixel whose sum sits on a rounding boundary
    assert np.count_nonzero(blur != ref) <= 2 and np.abs(blur.astype(int) - ref).max() <= 1


@pytest.mark.parametrize("h,w,ch,seed,thr,nms", [(480, 640, 3, 9500, 10, True), (301, 457, 1, 9501, 10, False),
                                                  (240, 320, 3, 9502, 0, True), (240, 320, 3, 9503, 35, True),
                                                  (7, 9, 3, 9504, 10, True), (6, 64, 1, 9505, 10, True),
                                                  (200, 200, 3, 9506, 255, True)])
def test_fast_detect_bit_exact(h, w, ch, seed, thr, nms):
    """The reference's fastExtractor (fastExtractor.cpp:7-13): positions, order and responses of
    cv2's FAST-9/16 keypoints, BGR frames included (the detector converts them itself)."""
    frame = synth.textured_frame(h, w, seed, ch)
    kps = cv2.FastFeatureDetector_create(thr, nms, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(frame)
    ref = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.float32).reshape(-1, 3)
    got = c_oracle.fast_detect(frame, thr, nms)
    assert np.array_equal(got, ref)
    assert all(k.size == 7.0 and k.angle == -1.0 and k.octave == 0 for k in kps[:50])


def test_sift_base_image_and_descriptors_vs_live_cv2():
    """SIFT extraction on FAST keypoints against live cv2: the working image bit for bit on frames
    of awkward sizes, the descriptors within the stated tolerance (cv::hal::exp32f / magnitude32f
    are IPP routines in the wheel and cannot be pinned to the last ulp)."""
    sigma = float(np.sqrt(np.float32(np.float32(1.6) * np.float32(1.6) - np.float32(0.25))))
    for rows, cols, ch in ((97, 131, 1), (64, 135, 3), (50, 37, 3), (33, 7, 1), (240, 320, 3)):
        frame = synth.textured_frame(rows, cols, 6400 + cols, ch)
        gray = (cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if ch == 3 else frame).astype(np.float32)
        assert np.array_equal(c_oracle.sift_base(frame), cv2.GaussianBlur(gray, (0, 0), sigma)), (rows, cols)
    frame = synth.textured_frame(480, 640, 6401, 3)
    kps = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(frame)
    assert len(kps) > 2000
    kept, want = cv2.SIFT_create().compute(frame, kps)
    assert len(kept) == len(kps)                      # SIFT drops no keypoint
    arr = np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in kps], np.float32)
    got = c_oracle.sift_compute(frame, arr)
    assert np.abs(got - want).max() <= 1.0
    assert np.mean(got == want) >= 0.999
