"""Regenerates tests/golden/*.npz from cv2 -- the reference's own arithmetic owner.

The reference has no tests, fixtures or golden vectors and cannot be compiled
in this image (OpenCV C++ dev files absent), so the golden outputs come from the very OpenCV
entry points the reference calls, through the cv2 wheel:

  featureMatchingCPU.cpp:26-40   DescriptorMatcher BRUTEFORCE / BRUTEFORCE_HAMMING, knnMatch k=2
  cameraTranslation.cpp:41-46    findEssentialMat(p1, p2, K, RANSAC, prob, threshold, mask)

Run from the repo root:   python -m oracle.gen_golden
Everything is seeded; outputs are committed so the GPU box (no /root/reference, possibly another
cv2 dispatch path) checks against exactly these bytes.
"""
import os

import cv2
import numpy as np

from . import synth

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

    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} fixtures, {total/1024:.0f} KiB, cv2 {cv2.__version__}")


if __name__ == "__main__":
    main()
