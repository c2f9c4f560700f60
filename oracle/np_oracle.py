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


# ---- the "next" rows of SURVEY.md 8f, restated a second time ---------------------------------------
def l1_dist_matrix(Q, T, block=64):
    """cv::BFMatcher(NORM_L1): s += |d0| + |d1| + |d2| + |d3| over groups of four, in float
    (featureMatchingCUDA.cpp:28 is the mode; the CPU matcher owns the arithmetic)."""
    Q = np.asarray(Q, np.float32).reshape(-1, 128)
    T = np.asarray(T, np.float32).reshape(-1, 128)
    out = np.empty((Q.shape[0], T.shape[0]), np.float32)
    for s in range(0, Q.shape[0], block):
        a = np.abs(Q[s:s + block, None, :] - T[None, :, :]).reshape(-1, T.shape[0], 32, 4)
        acc = np.zeros(a.shape[:2], np.float32)
        for g in range(32):
            acc = acc + (((a[:, :, g, 0] + a[:, :, g, 1]) + a[:, :, g, 2]) + a[:, :, g, 3])
        out[s:s + block] = acc
    return out


def project_points(obj, K4, dist, pose):
    """cv::projectPoints in cvProjectPoints2Internal's operation order (mainCycle.cpp:155-159 ->
    PnPRansacCallback::computeError); pose = R row-major (9) + t (3)."""
    k = np.zeros(12)
    if dist is not None:
        d = np.asarray(dist, np.float64).reshape(-1)
        k[:min(len(d), 12)] = d[:12]
    fx, fy, cx, cy = (float(v) for v in K4)
    R = np.asarray(pose, np.float64)[:9].reshape(3, 3)
    t = np.asarray(pose, np.float64)[9:]
    X = np.asarray(obj, np.float32).reshape(-1, 3).astype(np.float64)
    x = R[0, 0] * X[:, 0] + R[0, 1] * X[:, 1] + R[0, 2] * X[:, 2] + t[0]
    y = R[1, 0] * X[:, 0] + R[1, 1] * X[:, 1] + R[1, 2] * X[:, 2] + t[1]
    z = R[2, 0] * X[:, 0] + R[2, 1] * X[:, 1] + R[2, 2] * X[:, 2] + t[2]
    with np.errstate(all="ignore"):
        z = np.where(z != 0, 1.0 / z, 1.0)
        x, y = x * z, y * z
        r2 = x * x + y * y
        r4 = r2 * r2
        r6 = r4 * r2
        a1 = 2 * x * y
        a2 = r2 + 2 * x * x
        a3 = r2 + 2 * y * y
        cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6
        icdist2 = 1.0 / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6)
        xd0 = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4
        yd0 = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4
        return np.stack([xd0 * fx + cx, yd0 * fy + cy], 1).astype(np.float32)


def score_pnp(obj, img, K4, dist, poses, reproj_err, model_points=5):
    """counts[H], best (-1: none above model_points-1), best mask, all masks [H, M]."""
    img = np.asarray(img, np.float32).reshape(-1, 2)
    poses = np.asarray(poses, np.float64).reshape(-1, 12)
    t = np.float32(reproj_err * reproj_err)
    masks = np.zeros((len(poses), len(img)), np.uint8)
    with np.errstate(all="ignore"):
        for h, pose in enumerate(poses):
            d = img - project_points(obj, K4, dist, pose)        # float32
            err = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]          # float accumulator
            masks[h] = err <= t
    counts = masks.sum(1).astype(np.int32)
    best, bc = -1, 0
    for h, c in enumerate(counts):
        if c > max(bc, model_points - 1):
            best, bc = h, int(c)
    return counts, best, (masks[best] if best >= 0 else np.zeros(len(img), np.uint8)), masks


def orb_compute(image, kps, pattern):
    """cv::ORB::compute on given keypoints (featureMatchingCPU.cpp:45-66): `pattern` is the
    [256, 4] table of csrc/orb_pattern.h.  Needs cv2 only for the reflect-101 padding."""
    import cv2
    image = np.asarray(image, np.uint8)
    if image.ndim == 3:
        b, g, r = (image[..., i].astype(np.int64) for i in range(3))
        gray = ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)
    else:
        gray = image
    H, W = gray.shape
    v = np.exp(-0.125 * (np.arange(7) - 3.0) ** 2)
    k = (v * (1.0 / v.sum())).astype(np.float32)

    def fma(a, b, c):   # exact product in float64, one rounding of the sum (float32 operands)
        return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)

    P = cv2.copyMakeBorder(gray, 3, 3, 3, 3, cv2.BORDER_REFLECT_101).astype(np.float32)
    s = (P[:, 0:W] * k[0]).astype(np.float32)
    for j in range(1, 7):
        s = fma(P[:, j:j + W], k[j], s)
    c = (s[3:3 + H] * k[3]).astype(np.float32)
    for j in range(1, 4):
        c = fma((s[3 + j:3 + j + H] + s[3 - j:3 - j + H]).astype(np.float32), k[3 + j], c)
    blur = np.clip(np.rint(c), 0, 255).astype(np.uint8)
    kps = np.asarray(kps, np.float32).reshape(-1, 3)
    cx, cy = np.rint(kps[:, 0]).astype(np.int64), np.rint(kps[:, 1]).astype(np.int64)
    keep = (cx >= 31) & (cx < W - 31) & (cy >= 31) & (cy < H - 31)
    ang = (kps[keep, 2] * np.float32(np.pi / 180.0)).astype(np.float32)
    a, b = np.cos(ang.astype(np.float64)).astype(np.float32), np.sin(ang.astype(np.float64)).astype(np.float32)
    pat = np.asarray(pattern, np.float32)
    bits = []
    for e in range(2):
        px, py = pat[None, :, 2 * e], pat[None, :, 2 * e + 1]
        xr = (px * a[:, None]).astype(np.float32) - (py * b[:, None]).astype(np.float32)
        yr = (px * b[:, None]).astype(np.float32) + (py * a[:, None]).astype(np.float32)
        bits.append(blur[cy[keep, None] + np.rint(yr).astype(np.int64), cx[keep, None] + np.rint(xr).astype(np.int64)])
    return keep.astype(np.uint8), np.packbits(bits[0] < bits[1], axis=1, bitorder="little")


# ---- FAST-9/16 keypoints (fastExtractor.cpp:7-13), an array statement independent of the C loops ----
_FAST_CIRCLE = ((0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3),
                (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3))


def fast_detect(image, threshold=10, nonmax=True):
    """cv::FastFeatureDetector(threshold, nonmax, TYPE_9_16)::detect as whole-image array operations:
    the sixteen circle differences as shifted planes, the sixteen 9-pixel arcs as min / max over
    planes, suppression as a strict comparison with the 3x3 neighbourhood.  Returns [n, 3] float32
    rows {x, y, response} in row-major order."""
    img = np.asarray(image, np.uint8)
    if img.ndim == 3:
        b, g, r = (img[..., c].astype(np.int64) for c in range(3))
        img = ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)   # cvtColor BGR2GRAY
    rows, cols = img.shape
    if rows < 7 or cols < 7:
        return np.zeros((0, 3), np.float32)
    t = min(max(int(threshold), 0), 255)
    v = img[3:rows - 3, 3:cols - 3].astype(np.int32)
    d = np.stack([v - img[3 + dy:rows - 3 + dy, 3 + dx:cols - 3 + dx].astype(np.int32) for dx, dy in _FAST_CIRCLE])
    arcs = [[(s + j) % 16 for j in range(9)] for s in range(16)]
    arc_min = np.stack([d[a].min(0) for a in arcs])   # darker side: every pixel of the arc has d > t
    arc_max = np.stack([d[a].max(0) for a in arcs])   # brighter side: every pixel has d < -t
    corner = (arc_min > t).any(0) | (arc_max < -t).any(0)
    a0 = np.maximum(t, arc_min.max(0))
    b0 = np.minimum(-a0, arc_max.min(0))
    score = np.where(corner, -b0 - 1, 0).astype(np.int32)
    full = np.zeros((rows, cols), np.int32)
    full[3:rows - 3, 3:cols - 3] = score
    keep = np.zeros((rows, cols), bool)
    keep[3:rows - 3, 3:cols - 3] = corner
    if nonmax:
        pad = np.pad(full, 1)
        neigh = np.stack([pad[1 + dy:rows + 1 + dy, 1 + dx:cols + 1 + dx]
                          for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)]).max(0)
        keep &= full > neigh
    ys, xs = np.nonzero(keep)
    resp = full[ys, xs] if nonmax else np.zeros(len(ys))
    return np.stack([xs, ys, resp], 1).astype(np.float32).reshape(-1, 3)
