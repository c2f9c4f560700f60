"""Seeded synthetic inputs for the five BASELINE.json configs.

Input generator shared by tests/, bench.py and smoke() (not part of the checker in oracle/).  Pure NumPy; the
cv2-based hypothesis generator lives in gen_golden.py because cv2 is only the pinning reference.
"""
import numpy as np

# K of config/samsung-hv-4k.xml:3-9 in the reference (fx, fy, cx, cy); 3840x2160 frames.
SAMSUNG_HV_4K = (3.4412136432617754e+03, 3.4539293801685226e+03,
                 2.0099312767871520e+03, 1.1306070992635357e+03)


def sift_like(n, seed):
    """Integer-valued fp32 rows in [0,255] with ||row|| ~ 512, like cv::SIFT output."""
    rng = np.random.default_rng(seed)
    x = np.abs(rng.standard_normal((n, 128))).astype(np.float64)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = np.minimum(x, 0.2)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.clip(np.rint(x * 512.0), 0, 255).astype(np.float32)


def sift_pair(nq, nt, seed, planted=0.3, noise=6):
    """Query/train SIFT-like sets; `planted` of the queries get a noisy copy in the train set."""
    q = sift_like(nq, seed)
    t = sift_like(nt, seed + 7919)
    rng = np.random.default_rng(seed + 104729)
    k = int(min(nq, nt) * planted)
    if k > 0:
        qi = rng.choice(nq, k, replace=False)
        ti = rng.choice(nt, k, replace=False)
        jitter = rng.integers(-noise, noise + 1, (k, 128))
        t[ti] = np.clip(q[qi] + jitter, 0, 255).astype(np.float32)
    return q, t


def sift_train_from_query(q, nt, seed, planted=0.3, noise=6):
    """A train set for a fixed query set (the 1 x framesBatchSize window of batch.cpp:120-148)."""
    t = sift_like(nt, seed)
    rng = np.random.default_rng(seed + 104729)
    k = int(min(q.shape[0], nt) * planted)
    if k > 0:
        qi = rng.choice(q.shape[0], k, replace=False)
        ti = rng.choice(nt, k, replace=False)
        t[ti] = np.clip(q[qi] + rng.integers(-noise, noise + 1, (k, 128)), 0, 255).astype(np.float32)
    return t


def float_pair(nq, nt, seed):
    """General-float stress descriptors: uniform [0,255) fp32 (no structure, worst case)."""
    rng = np.random.default_rng(seed)
    return (rng.random((nq, 128), np.float32) * 255).astype(np.float32), \
           (rng.random((nt, 128), np.float32) * 255).astype(np.float32)


def orb_pair(nq, nt, seed, planted=0.3, max_flips=40):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    k = int(min(nq, nt) * planted)
    if k > 0:
        qi = rng.choice(nq, k, replace=False)
        ti = rng.choice(nt, k, replace=False)
        rows = q[qi].copy()
        for r in range(k):
            nf = int(rng.integers(0, max_flips + 1))
            bits = rng.choice(256, nf, replace=False)
            for b in bits:
                rows[r, b >> 3] ^= np.uint8(1 << (b & 7))
        t[ti] = rows
    return q, t


def _rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def _skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def two_view(m, seed, K4=SAMSUNG_HV_4K, noise_px=0.7, outliers=0.3, size=(3840, 2160)):
    """M matched pixel pairs of a random rigid two-view scene (+ Gaussian noise, + outliers).

    Returns pts1, pts2 (float32 [M,2]) and the true (R, t) with x2 ~ R x1 + t."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K4
    R = _rodrigues(rng.normal(0, 0.08, 3))
    t = rng.normal(0, 1.0, 3)
    t /= np.linalg.norm(t)
    u = rng.uniform(0, size[0], m)
    v = rng.uniform(0, size[1], m)
    z = rng.uniform(4.0, 20.0, m)
    X1 = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], axis=1)
    X2 = X1 @ R.T + t
    p1 = np.stack([u, v], axis=1)
    p2 = np.stack([X2[:, 0] / X2[:, 2] * fx + cx, X2[:, 1] / X2[:, 2] * fy + cy], axis=1)
    p1 = p1 + rng.normal(0, noise_px, p1.shape)
    p2 = p2 + rng.normal(0, noise_px, p2.shape)
    n_out = int(m * outliers)
    if n_out:
        oi = rng.choice(m, n_out, replace=False)
        p2[oi] = np.stack([rng.uniform(0, size[0], n_out), rng.uniform(0, size[1], n_out)], axis=1)
    return p1.astype(np.float32), p2.astype(np.float32), R, t


def pose_hypotheses(h, R, t, seed, good_frac=0.25):
    """H essential matrices [t]x R around (and away from) the true pose, as RANSAC's minimal
    solver would propose: a `good_frac` share are small perturbations of the truth, the rest are
    poses fitted to contaminated samples (large perturbations)."""
    rng = np.random.default_rng(seed)
    E = np.empty((h, 9), np.float64)
    for i in range(h):
        s = 0.002 if rng.random() < good_frac else 0.3
        Ri = _rodrigues(rng.normal(0, s, 3)) @ R
        ti = t + rng.normal(0, s, 3)
        ti /= np.linalg.norm(ti)
        Ei = _skew(ti) @ Ri
        E[i] = (Ei / np.linalg.norm(Ei) * np.sqrt(2.0)).reshape(9)
    return E


REF_DIST5 = (0.11, -0.23, 0.0012, -0.0007, 0.09)  # k1 k2 p1 p2 k3: the reference's 1x5 "DC" Mat


def pnp_scene(m, seed, K4=SAMSUNG_HV_4K, dist=REF_DIST5, noise_px=0.7, outliers=0.3,
              size=(3840, 2160)):
    """M 3D-2D correspondences of a random camera pose (the solvePnPRansac input of
    mainCycle.cpp:155-159): object points (float32 [M,3]), image points (float32 [M,2], Gaussian
    noise + uniform outliers), and the true (R, t).  The image points come from this module's own
    fp64 pinhole + Brown model (no OpenCV needed)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K4
    k = np.zeros(12)
    k[:len(dist)] = dist
    R = _rodrigues(rng.normal(0, 0.15, 3))
    t = rng.normal(0, 0.3, 3)
    u = rng.uniform(0, size[0], m)
    v = rng.uniform(0, size[1], m)
    z = rng.uniform(4.0, 20.0, m)
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], axis=1)   # camera frame
    X = ((Xc - t) @ R).astype(np.float32)                              # world: Xc = R X + t
    Xc = X.astype(np.float64) @ R.T + t
    x, y = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
    r2 = x * x + y * y
    cd = (1 + k[0] * r2 + k[1] * r2 ** 2 + k[4] * r2 ** 3) / (1 + k[5] * r2 + k[6] * r2 ** 2 + k[7] * r2 ** 3)
    xd = x * cd + 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 ** 2
    yd = y * cd + k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 ** 2
    p = np.stack([xd * fx + cx, yd * fy + cy], axis=1) + rng.normal(0, noise_px, (m, 2))
    n_out = int(m * outliers)
    if n_out:
        oi = rng.choice(m, n_out, replace=False)
        p[oi] = np.stack([rng.uniform(0, size[0], n_out), rng.uniform(0, size[1], n_out)], axis=1)
    return X, p.astype(np.float32), R, t


def pnp_hypotheses(h, R, t, seed, good_frac=0.25):
    """H candidate poses [R | t] as 12 doubles each (rotation row-major, then t), a `good_frac`
    share close to the truth (a continuum of perturbation sizes, so that many points sit near
    the reprojection threshold), the rest as fitted to contaminated samples."""
    rng = np.random.default_rng(seed)
    out = np.empty((h, 12), np.float64)
    for i in range(h):
        s = 10 ** rng.uniform(-4.5, -3) if rng.random() < good_frac else 10 ** rng.uniform(-3.3, -1.5)
        out[i, :9] = (_rodrigues(rng.normal(0, s, 3)) @ R).reshape(9)
        out[i, 9:] = t + rng.normal(0, s * 3, 3)
    return out


def textured_frame(h, w, seed, channels=3):
    """A frame with structure at several scales (blocks, gradients, noise) so that FAST fires and
    blurred intensities differ between nearby pixels; uint8 [h, w, channels] or [h, w]."""
    rng = np.random.default_rng(seed)
    shape = (h, w, channels) if channels > 1 else (h, w)
    img = np.zeros(shape, np.float32)
    for cell, amp in ((64, 90.0), (16, 70.0), (4, 50.0), (1, 30.0)):
        gh, gw = (h + cell - 1) // cell, (w + cell - 1) // cell
        g = rng.random((gh, gw) + shape[2:], np.float32)
        g = np.repeat(np.repeat(g, cell, axis=0), cell, axis=1)[:h, :w]
        img += amp * g
    img += np.linspace(0, 15, w, dtype=np.float32).reshape((1, w) + (1,) * (len(shape) - 2))
    return np.clip(img, 0, 255).astype(np.uint8)


def sift_train_plants(nq, nt, seed, planted=0.3):
    """The (query row, train row) pairs sift_train_from_query(q, nt, seed) plants (same RNG draws)."""
    rng = np.random.default_rng(seed + 104729)
    k = int(min(nq, nt) * planted)
    if k <= 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    qi = rng.choice(nq, k, replace=False)
    ti = rng.choice(nt, k, replace=False)
    return qi, ti


def window_geometry(nq, nt, train_seeds, seed, K4=SAMSUNG_HV_4K, noise_px=0.7, size=(3840, 2160)):
    """Keypoint coordinates that are geometrically consistent with the planted correspondences of a
    1 x len(train_seeds) window (BASELINE cfg3: matching + RANSAC essential scoring per pair).

    The query frame sees nq random 3D points; train frame p (descriptors from
    sift_train_from_query(q, nt, train_seeds[p])) views the same scene from its own pose (R_p, t_p):
    the train row planted as the copy of query row i carries the projection of query point i in
    view p (+ Gaussian pixel noise), every other train row a uniform random position.  Returns
    kq float32 [nq,2], [kt_p float32 [nt,2]], [(R_p, t_p)]."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K4
    u = rng.uniform(0, size[0], nq)
    v = rng.uniform(0, size[1], nq)
    z = rng.uniform(4.0, 20.0, nq)
    X1 = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], axis=1)
    kq = (np.stack([u, v], axis=1) + rng.normal(0, noise_px, (nq, 2))).astype(np.float32)
    kts, poses = [], []
    for s in train_seeds:
        R = _rodrigues(rng.normal(0, 0.08, 3))
        t = rng.normal(0, 1.0, 3)
        t /= np.linalg.norm(t)
        kt = np.stack([rng.uniform(0, size[0], nt), rng.uniform(0, size[1], nt)], axis=1)
        qi, ti = sift_train_plants(nq, nt, s)
        X2 = X1[qi] @ R.T + t
        kt[ti] = np.stack([X2[:, 0] / X2[:, 2] * fx + cx, X2[:, 1] / X2[:, 2] * fy + cy], axis=1) + \
            rng.normal(0, noise_px, (len(qi), 2))
        kts.append(kt.astype(np.float32))
        poses.append((R, t))
    return kq, kts, poses


def _rodrigues_batch(w):
    """Rotation matrices of h rotation vectors [h,3] -> [h,3,3] (vectorised _rodrigues)."""
    th = np.linalg.norm(w, axis=1)
    k = w / np.maximum(th, 1e-300)[:, None]
    Kx = np.zeros((len(w), 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2] = -k[:, 2], k[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = k[:, 2], -k[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 1], k[:, 0]
    I = np.eye(3)[None]
    R = I + np.sin(th)[:, None, None] * Kx + (1 - np.cos(th))[:, None, None] * (Kx @ Kx)
    R[th < 1e-12] = np.eye(3)
    return R


def pose_hypotheses_fast(h, R, t, seed, good_frac=0.25):
    """pose_hypotheses without the Python loop (same distribution, not the same draws): [h, 9]."""
    rng = np.random.default_rng(seed)
    s = np.where(rng.random(h) < good_frac, 0.002, 0.3)
    Ri = _rodrigues_batch(rng.normal(0, 1.0, (h, 3)) * s[:, None]) @ R
    ti = t[None] + rng.normal(0, 1.0, (h, 3)) * s[:, None]
    ti /= np.linalg.norm(ti, axis=1, keepdims=True)
    Tx = np.zeros((h, 3, 3))
    Tx[:, 0, 1], Tx[:, 0, 2] = -ti[:, 2], ti[:, 1]
    Tx[:, 1, 0], Tx[:, 1, 2] = ti[:, 2], -ti[:, 0]
    Tx[:, 2, 0], Tx[:, 2, 1] = -ti[:, 1], ti[:, 0]
    E = Tx @ Ri
    E /= np.linalg.norm(E.reshape(h, 9), axis=1)[:, None, None] / np.sqrt(2.0)
    return np.ascontiguousarray(E.reshape(h, 9))
