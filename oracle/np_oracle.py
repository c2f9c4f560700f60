"""NumPy restatement of the correspondence hot path.  TEST INFRASTRUCTURE ONLY.

An independent second statement of oracle/corr_oracle.c (same reference citations), written with
whole-array float32/float64 operations.  NumPy never fuses a multiply with an add, so every
operation is rounded separately exactly as in the SSE-baseline OpenCV build the reference calls
(featureMatchingCPU.cpp:40, cameraTranslation.cpp:41-46).
"""
import numpy as np

DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"),
                   ("distance", "<f4")])


def l2_dist_matrix(Q, T, block=32):
    """sqrtf(hal::normL2Sqr_(q, t, 128)) for every pair, in cv2's summation order."""
    Q = np.asarray(Q, np.float32).reshape(-1, 128)
    T = np.asarray(T, np.float32).reshape(-1, 128)
    out = np.empty((Q.shape[0], T.shape[0]), np.float32)
    for s in range(0, Q.shape[0], block):
        q = Q[s:s + block]
        diff = q[:, None, :] - T[None, :, :]
        sq = (diff * diff).reshape(q.shape[0], T.shape[0], 8, 4, 4)  # [i][a][l]
        acc = np.zeros((q.shape[0], T.shape[0], 4, 4), np.float32)
        for i in range(8):
            acc = acc + sq[:, :, i]
        S = ((acc[:, :, 0] + acc[:, :, 1]) + acc[:, :, 2]) + acc[:, :, 3]
        d2 = (S[:, :, 0] + S[:, :, 2]) + (S[:, :, 1] + S[:, :, 3])
        out[s:s + block] = np.sqrt(d2)
    return out


_POP8 = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming_dist_matrix(Q, T):
    Q = np.asarray(Q, np.uint8).reshape(-1, 32)
    T = np.asarray(T, np.uint8).reshape(-1, 32)
    x = Q[:, None, :] ^ T[None, :, :]
    return _POP8[x].sum(axis=2).astype(np.int32)


def knn2_from_matrix(D):
    """BatchDistInvoker's K=2 selection: ascending distance, ties by ascending train index."""
    nq, nt = D.shape
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.zeros((nq, 2), np.float32)
    if nt == 0:
        return idx, dist
    order = np.argsort(D, axis=1, kind="stable")[:, :2]
    k = order.shape[1]
    idx[:, :k] = order
    dist[:, :k] = np.take_along_axis(D, order, axis=1).astype(np.float32)
    return idx, dist


def ratio_test(idx, dist, ratio):
    """getGoodMatches (featureMatchingCommon.cpp:37-50); single-element rows are rejected."""
    rows = []
    r = float(ratio)
    for q in range(idx.shape[0]):
        if idx[q, 0] < 0 or idx[q, 1] < 0:
            continue
        if float(dist[q, 0]) < r * float(dist[q, 1]):
            rows.append((q, int(idx[q, 0]), 0, dist[q, 0]))
    return np.array(rows, DMATCH) if rows else np.zeros(0, DMATCH)


def normalize_points(pts, fx, fy, cx, cy):
    """findEssentialMat's (col - c) / f, folded by MatExpr into u*(1/f) + (-c*(1/f))."""
    p = np.asarray(pts, np.float32).reshape(-1, 2).astype(np.float64)
    ax, ay = np.float64(1.0) / np.float64(fx), np.float64(1.0) / np.float64(fy)
    bx, by = -np.float64(cx) * ax, -np.float64(cy) * ay
    return np.stack([p[:, 0] * ax + bx, p[:, 1] * ay + by], axis=1)


def sampson_errors(pts1, pts2, K4, E):
    """EMEstimatorCallback::computeError for every (hypothesis, match): float32 [H, M]."""
    fx, fy, cx, cy = [np.float64(v) for v in K4]
    n1 = normalize_points(pts1, fx, fy, cx, cy)
    n2 = normalize_points(pts2, fx, fy, cx, cy)
    E = np.asarray(E, np.float64).reshape(-1, 3, 3)
    x1, y1 = n1[None, :, 0], n1[None, :, 1]
    x2, y2 = n2[None, :, 0], n2[None, :, 1]
    e = lambda r, c: E[:, r, c][:, None]
    zero = np.float64(0.0)
    Ex1 = [((zero + e(r, 0) * x1) + e(r, 1) * y1) + e(r, 2) * 1.0 for r in range(3)]
    Etx2 = [((zero + e(0, r) * x2) + e(1, r) * y2) + e(2, r) * 1.0 for r in range(2)]
    s = ((zero + x2 * Ex1[0]) + y2 * Ex1[1]) + 1.0 * Ex1[2]
    a, b = Ex1[0] * Ex1[0], Ex1[1] * Ex1[1]
    c, d = Etx2[0] * Etx2[0], Etx2[1] * Etx2[1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return (s * s / (((a + b) + c) + d)).astype(np.float32)


def score_essential(pts1, pts2, K4, E, threshold_px):
    """RANSAC scoring of a fixed hypothesis list (ptsetreg.cpp findInliers + run())."""
    fx, fy = np.float64(K4[0]), np.float64(K4[1])
    thr = np.float64(threshold_px) / ((fx + fy) / np.float64(2))
    t = np.float32(thr * thr)
    err = sampson_errors(pts1, pts2, K4, E)
    masks = (err <= t).astype(np.uint8)
    counts = masks.sum(axis=1).astype(np.int32)
    best, bc = -1, 0
    for h in range(counts.shape[0]):
        if counts[h] > max(bc, 4):
            best, bc = h, int(counts[h])
    best_mask = masks[best] if best >= 0 else np.zeros(masks.shape[1], np.uint8)
    return counts, best, best_mask, masks
