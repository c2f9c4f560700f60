"""Regenerates tests/golden/*.npz from cv2 -- the reference's own arithmetic owner.

The reference has no tests, fixtures or golden vectors and cannot be compiled
in this image (OpenCV C++ dev files absent), so the golden outputs come from the very OpenCV
entry points the reference calls, through the cv2 wheel:

  featureMatchingCPU.cpp:26-40   DescriptorMatcher BRUTEFORCE / BRUTEFORCE_HAMMING, knnMatch k=2
  cameraTranslation.cpp:41-46    findEssentialMat(p1, p2, K, RANSAC, prob, threshold, mask)
  featureMatchingCUDA.cpp:28     NORM_L1 for useFM-SIFT-BF (CPU BFMatcher as the owner)
  featureMatchingCPU.cpp:45-66   ORB::create()->compute(frame, features, desc)     
  mainCycle.cpp:155-159          solvePnPRansac(obj, img, K, dist, rvec, tvec)   

Run from the repo root:   python -m oracle.gen_golden          (all fixtures)
                          python -m oracle.gen_golden pnp|l1   (only pnp.npz / sift_l1.npz)
Everything is seeded; outputs are committed so the GPU box (no /root/reference, possibly another
cv2 dispatch path) checks against exactly these bytes.
"""
import os

import cv2
import numpy as np

import synth_inputs as synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cv_knn2(Q, T, norm):
    """BFMatcher(norm).knnMatch(Q, T, 2) flattened to idx[nq,2] (-1 = absent) / dist[nq,2]."""
    nq = Q.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.zeros((nq, 2), np.float32)
    if nq == 0:
        return idx, dist, []
    if T.shape[0] == 0:
        # cv2 cannot take an empty train Mat through the Python binding (it asserts on type);
        # the C++ behaviour (Q empty lists) was verified separately.
        return idx, dist, [0] * nq
    res = cv2.BFMatcher(norm).knnMatch(Q, T, 2)
    lens = []
    for q, row in enumerate(res):
        lens.append(len(row))
        for k, m in enumerate(row):
            assert m.queryIdx == q and m.imgIdx == 0
            idx[q, k] = m.trainIdx
            dist[q, k] = m.distance
    return idx, dist, lens


def five_point_hypotheses(p1, p2, K4, n_subsets, seed):
    """E candidates of cv2's 5-point solver on seeded 5-subsets (stacked 3k x 3 per call)."""
    rng = np.random.default_rng(seed)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float64)
    out = []
    for _ in range(n_subsets):
        s = rng.choice(p1.shape[0], 5, replace=False)
        E = cv2.findEssentialMat(p1[s], p2[s], Kmat, cv2.RANSAC, 0.999, 1.0)[0]
        if E is None:
            continue
        E = np.asarray(E, np.float64).reshape(-1, 3, 3)
        out.extend(E.reshape(-1, 9))
    return np.array(out, np.float64)


def main():
    os.makedirs(OUT, exist_ok=True)
    cv2.setNumThreads(1)

    # --- SIFT-like integer descriptors (what cv::SIFT emits), planted correspondences
    q, t = synth.sift_pair(300, 400, 1001)
    idx, dist, _ = cv_knn2(q, t, cv2.NORM_L2)
    np.savez_compressed(os.path.join(OUT, "sift_int.npz"), q=q.astype(np.uint8),
                        t=t.astype(np.uint8), idx=idx, dist=dist)

    # --- general floats: the cv2 summation order matters for the low bits of the distances
    q, t = synth.float_pair(200, 300, 1002)
    idx, dist, _ = cv_knn2(q, t, cv2.NORM_L2)
    np.savez_compressed(os.path.join(OUT, "sift_float.npz"), q=q, t=t, idx=idx, dist=dist)

    # --- exact ties: duplicated train rows, duplicated best and second best, zero distances
    q, t = synth.sift_pair(64, 96, 1003, planted=0.0)
    t[10] = t[3]; t[50] = t[3]; t[51] = t[3]        # three-way tie wherever row 3 is near
    t[20] = q[5]; t[7] = q[5]                       # two exact (distance 0) copies of query 5
    t[60] = q[9]                                    # one exact copy: d0 = 0, ratio test 0 < r*d1
    q[11] = q[12]                                   # identical queries get identical rows
    idx, dist, _ = cv_knn2(q, t, cv2.NORM_L2)
    np.savez_compressed(os.path.join(OUT, "sift_ties.npz"), q=q.astype(np.uint8),
                        t=t.astype(np.uint8), idx=idx, dist=dist)

    # --- ragged / tiny shapes
    for name, nq, nt in (("sift_t1", 5, 1), ("sift_t2", 7, 2), ("sift_q1", 1, 33)):
        q, t = synth.sift_pair(nq, nt, 1100 + nq * 10 + nt, planted=0.0)
        idx, dist, lens = cv_knn2(q, t, cv2.NORM_L2)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), q=q.astype(np.uint8),
                            t=t.astype(np.uint8), idx=idx, dist=dist, lens=np.array(lens))

    # --- ORB 256-bit
    q, t = synth.orb_pair(300, 400, 2001)
    t[17] = t[4]; t[200] = t[4]                     # ties
    t[33] = q[8]                                    # zero distance
    idx, dist, _ = cv_knn2(q, t, cv2.NORM_HAMMING)
    np.savez_compressed(os.path.join(OUT, "orb.npz"), q=q, t=t, idx=idx, dist=dist)
    q, t = synth.orb_pair(6, 1, 2002, planted=0.0)
    idx, dist, lens = cv_knn2(q, t, cv2.NORM_HAMMING)
    np.savez_compressed(os.path.join(OUT, "orb_t1.npz"), q=q, t=t, idx=idx, dist=dist,
                        lens=np.array(lens))

    # --- RANSAC essential: cv2's winning model + its mask, and a 5-point hypothesis list
    K4 = np.array(synth.SAMSUNG_HV_4K, np.float64)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float64)
    cases = {}
    for c, (m, seed) in enumerate(((600, 5000), (1500, 5001), (257, 5002), (40, 5003))):
        p1, p2, _, _ = synth.two_view(m, seed)
        E, mask = cv2.findEssentialMat(p1, p2, Kmat, cv2.RANSAC, 0.999, 5.0)
        hyp = five_point_hypotheses(p1, p2, K4, 24, seed + 100)
        cases[f"p1_{c}"] = p1
        cases[f"p2_{c}"] = p2
        cases[f"E_cv_{c}"] = np.asarray(E, np.float64).reshape(-1, 9)[:1]
        cases[f"mask_cv_{c}"] = mask.reshape(-1).astype(np.uint8)
        cases[f"hyp_{c}"] = hyp
    cases["K4"] = K4
    cases["threshold_px"] = np.float64(5.0)
    np.savez_compressed(os.path.join(OUT, "ransac.npz"), **cases)

    gen_l1()
    gen_pnp()
    gen_orb_desc()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} fixtures, {total/1024:.0f} KiB, cv2 {cv2.__version__}")


def gen_l1():
    """NORM_L1: cv2.BFMatcher(NORM_L1) on integer rows, general floats, ties and
    the ragged shapes -- the inputs of the existing SIFT fixtures, so only the results are stored."""
    out = {}
    for name in ("sift_int", "sift_float", "sift_ties", "sift_t1", "sift_t2", "sift_q1"):
        g = np.load(os.path.join(OUT, name + ".npz"))
        idx, dist, _ = cv_knn2(g["q"].astype(np.float32), g["t"].astype(np.float32), cv2.NORM_L1)
        out[name + "_idx"], out[name + "_dist"] = idx, dist
    np.savez_compressed(os.path.join(OUT, "sift_l1.npz"), **out)


def gen_orb_desc():
    """extractDescriptor's ORB branch: FAST keypoints (angle -1) and oriented keypoints on a small
    textured frame; cv2.ORB.compute's kept keypoints and descriptors."""
    frame = synth.textured_frame(200, 260, 6100, 3)
    fast = cv2.FastFeatureDetector_create(10, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(frame)[:400]
    rng = np.random.default_rng(6101)
    kps = np.array([[k.pt[0], k.pt[1], k.angle] for k in fast], np.float32)
    extra = np.stack([rng.uniform(0, 260, 200), rng.uniform(0, 200, 200), rng.uniform(0, 360, 200)], 1).astype(np.float32)
    kps = np.concatenate([kps, extra])
    cvk = [cv2.KeyPoint(float(x), float(y), 7.0, float(a), 0.0, 0) for x, y, a in kps]
    kept, desc = cv2.ORB_create().compute(frame, cvk)
    kept_xy = np.array([k.pt for k in kept], np.float32)
    np.savez_compressed(os.path.join(OUT, "orb_desc.npz"), frame=frame, kps=kps, kept_xy=kept_xy, desc=desc)


def gen_pnp():
    """solvePnPRansac: cv2's own result, plus per-hypothesis inlier masks built from
    cv2.projectPoints (the call PnPRansacCallback::computeError makes) for EPnP poses of seeded
    5-subsets."""
    K4 = np.array(synth.SAMSUNG_HV_4K, np.float64)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float64)
    dists = (np.array(synth.REF_DIST5), np.zeros(5),
             np.array([0.11, -0.23, 0.0012, -0.0007, 0.09, 0.01, -0.02, 0.003, 1e-3, -2e-3, 3e-4, 1e-4]))
    cases = {"K4": K4, "reproj": np.float64(8.0)}
    t = np.float32(8.0 * 8.0)
    for c, (m, seed, dv) in enumerate(((800, 7000, 0), (2000, 7001, 0), (300, 7002, 1), (500, 7003, 2),
                                       (33, 7004, 0))):
        dist = dists[dv]
        obj, img, _, _ = synth.pnp_scene(m, seed, dist=tuple(dist))
        rng = np.random.default_rng(seed + 100)
        poses, masks = [], []
        while len(poses) < 16:
            idx = rng.choice(m, 5, replace=False)
            ok, r, tv = cv2.solvePnP(obj[idx], img[idx], Kmat, dist, flags=cv2.SOLVEPNP_EPNP)
            if not ok:
                continue
            uv = cv2.projectPoints(obj, r, tv, Kmat, dist)[0].reshape(-1, 2).astype(np.float32)
            d = img - uv                                   # float32, as Point2f - Point2f
            err = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]    # float accumulator: s = dx*dx; s += dy*dy
            poses.append(np.concatenate([cv2.Rodrigues(r)[0].reshape(-1), tv.reshape(-1)]))
            masks.append((err <= t).astype(np.uint8))
        ok, rvec, tvec, inl = cv2.solvePnPRansac(obj, img, Kmat, dist)
        cases[f"obj_{c}"], cases[f"img_{c}"], cases[f"dist_{c}"] = obj, img, dist
        cases[f"poses_{c}"] = np.array(poses)
        cases[f"masks_{c}"] = np.packbits(np.array(masks), axis=1)
        cases[f"cv_ok_{c}"] = np.bool_(ok)
        cases[f"cv_rvec_{c}"], cases[f"cv_tvec_{c}"] = rvec.reshape(3), tvec.reshape(3)
        cases[f"cv_inliers_{c}"] = inl.reshape(-1).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "pnp.npz"), **cases)


def gen_fast():
    """fastExtractor (fastExtractor.cpp:7-13): cv2's FAST-9/16 keypoints {x, y, response} of a small
    BGR frame, with and without suppression, at the reference's default threshold and another."""
    frame = synth.textured_frame(120, 168, 6200, 3)
    out = {"frame": frame}
    for thr, nms in ((10, True), (10, False), (25, True)):
        kps = cv2.FastFeatureDetector_create(thr, nms, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(frame)
        out[f"kp_t{thr}_n{int(nms)}"] = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.float32).reshape(-1, 3)
    np.savez_compressed(os.path.join(OUT, "fast.npz"), **out)


def gen_sift_desc():
    """extractDescriptor's SIFT branch: FAST keypoints (size 7, angle -1) plus keypoints of other sizes
    and orientations on a small textured frame whose width is NOT a multiple of 8 (the scalar tails of
    OpenCV's separable filter are part of the pin): cv2's float base image (GaussianBlur of the gray
    frame with createInitialImage's sigma) as a checksum-free slice, and cv2.SIFT.compute's rows."""
    frame = synth.textured_frame(150, 205, 6300, 3)
    fast = cv2.FastFeatureDetector_create(15, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(frame)[:300]
    rng = np.random.default_rng(6301)
    kps = np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in fast], np.float32)
    extra = np.stack([rng.uniform(0, 205, 120), rng.uniform(0, 150, 120), rng.uniform(2, 12, 120),
                      rng.uniform(0, 360, 120)], 1).astype(np.float32)
    kps = np.concatenate([kps, extra])
    cvk = [cv2.KeyPoint(float(x), float(y), float(s), float(a), 0.0, 0) for x, y, s, a in kps]
    kept, desc = cv2.SIFT_create().compute(frame, cvk)
    assert len(kept) == len(cvk)
    gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY).astype(np.float32)
    sigma = float(np.sqrt(np.float32(np.float32(1.6) * np.float32(1.6) - np.float32(0.25))))
    base = cv2.GaussianBlur(gray, (0, 0), sigma)
    np.savez_compressed(os.path.join(OUT, "sift_desc.npz"), frame=frame, kps=kps, desc=desc.astype(np.uint8),
                        base_rows=base[[0, 1, 74, 148, 149]], base_cols=base[:, [0, 1, 199, 200, 203, 204]])


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "sift_desc":
        gen_sift_desc()
    elif len(sys.argv) > 1 and sys.argv[1] == "pnp":
        gen_pnp()
    elif len(sys.argv) > 1 and sys.argv[1] == "l1":
        gen_l1()
    elif len(sys.argv) > 1 and sys.argv[1] == "orb_desc":
        gen_orb_desc()
    elif len(sys.argv) > 1 and sys.argv[1] == "fast":
        gen_fast()
    else:
        main()
