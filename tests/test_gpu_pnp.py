"""GPU parity tests for the next row SURVEY.md 8f-2: solvePnPRansac inlier scoring through the C ABI.

Checker: the CPU oracle (pinned to cv2.projectPoints / cv2.solvePnPRansac) -- counts, winner and
masks bit-exact.  Mirrors solvePnPRansac(obj, img, K, dist, rvec, tvec) at
cycleProcessing/mainCycle.cpp:155-159.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, synth
from slam_indoor_code_b200 import pnp_ransac as pr

K4 = np.array(synth.SAMSUNG_HV_4K)
DIST12 = np.array([0.11, -0.23, 0.0012, -0.0007, 0.09, 0.01, -0.02, 0.003, 1e-3, -2e-3, 3e-4, 1e-4])


def test_golden_masks_from_cv2_project_points(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "pnp.npz"))
    for c in range(5):
        obj, img, dist, poses = g[f"obj_{c}"], g[f"img_{c}"], g[f"dist_{c}"], g[f"poses_{c}"]
        want = np.unpackbits(g[f"masks_{c}"], axis=1)[:, :len(obj)]
        counts, best, mask, allm = pr.scorePnPHypotheses(ctx, obj, img, g["K4"], dist, poses,
                                                         float(g["reproj"]), want_all_masks=True)
        assert np.array_equal(allm, want) and np.array_equal(counts, want.sum(1))
        rc, rb, rm, _ = c_oracle.score_pnp(obj, img, g["K4"], dist, poses, float(g["reproj"]))
        assert best == rb and np.array_equal(mask, rm)


@pytest.mark.parametrize("M,H,seed,dist", [(5000, 2048, 7100, synth.REF_DIST5), (1237, 300, 7101, None),
                                           (4096, 129, 7102, DIST12), (333, 1, 7103, synth.REF_DIST5),
                                           (2000, 512, 7104, np.zeros(5)), (800, 64, 7105, DIST12[:8]),
                                           (800, 64, 7106, (0, 0, 0, 0, 0, 0, 0, 0, 1e-3, 0, 0, 0, 0, 0))])
def test_synthetic_hypotheses(ctx, M, H, seed, dist):
    obj, img, R, t = synth.pnp_scene(M, seed, dist=tuple(dist)[:12] if dist is not None else ())
    poses = synth.pnp_hypotheses(H, R, t, seed + 1)
    counts, best, mask, allm = pr.scorePnPHypotheses(ctx, obj, img, K4, dist, poses, 8.0, want_all_masks=True)
    rc, rb, rm, rall = c_oracle.score_pnp(obj, img, K4, dist, poses, 8.0, want_all_masks=True)
    assert np.array_equal(counts, rc) and best == rb
    assert np.array_equal(mask, rm) and np.array_equal(allm, rall)
    if H >= 64:
        assert rc.max() > M // 2 and 0 < np.count_nonzero((rc > 0) & (rc < rc.max()))   # a real spread


def test_first_best_wins_and_model_points_floor(ctx):
    obj, img, R, t = synth.pnp_scene(300, 7200)
    poses = synth.pnp_hypotheses(8, R, t, 7201)
    both = np.concatenate([poses, poses])
    counts, best, _, _ = pr.scorePnPHypotheses(ctx, obj, img, K4, synth.REF_DIST5, both, 8.0)
    assert best == int(np.argmax(counts)) and best < 8
    # four correspondences can never exceed the "> model_points - 1" floor of the EPnP kernel ...
    _, best, mask, _ = pr.scorePnPHypotheses(ctx, obj[:4], img[:4], K4, synth.REF_DIST5, poses, 1e6, 5)
    assert best == -1 and not mask.any()
    # ... but they do pass the P3P floor (model_points = 4) when all four are inliers
    _, best, mask, _ = pr.scorePnPHypotheses(ctx, obj[:4], img[:4], K4, synth.REF_DIST5, poses, 1e6, 4)
    rc, rb, rm, _ = c_oracle.score_pnp(obj[:4], img[:4], K4, synth.REF_DIST5, poses, 1e6, 4)
    assert best == rb and np.array_equal(mask, rm)
    # empty inputs
    counts, best, mask, _ = pr.scorePnPHypotheses(ctx, obj[:0], img[:0], K4, None, poses, 8.0)
    assert best == -1 and len(mask) == 0 and not counts.any()
    counts, best, mask, _ = pr.scorePnPHypotheses(ctx, obj, img, K4, None, poses[:0], 8.0)
    assert best == -1 and len(counts) == 0 and not mask.any()


def test_degenerate_poses_and_points(ctx):
    """Points on / behind the camera plane (z == 0 takes the 'z ? 1/z : 1' branch), NaN and huge
    poses, coordinates whose r^6 overflows: the same side of the threshold as the CPU."""
    obj, img, R, t = synth.pnp_scene(400, 7300)
    poses = synth.pnp_hypotheses(8, R, t, 7301)
    poses[1] = 0.0                      # z == 0 for every point
    poses[2] = np.nan
    poses[3, 9:] *= 1e200
    poses[4, :9] *= 1e-200
    poses[5, 11] = -obj[7, 2]           # with R ~ I this puts point 7 near z = 0
    poses[5, :9] = np.eye(3).reshape(9)
    obj[11] = (1e30, -1e30, 1e-30)
    obj[12] = (0, 0, 0)
    for dist in (None, synth.REF_DIST5, DIST12):
        counts, best, mask, allm = pr.scorePnPHypotheses(ctx, obj, img, K4, dist, poses, 8.0, want_all_masks=True)
        rc, rb, rm, rall = c_oracle.score_pnp(obj, img, K4, dist, poses, 8.0, want_all_masks=True)
        assert np.array_equal(counts, rc) and best == rb and np.array_equal(allm, rall)


def test_threshold_boundary(ctx):
    """Thresholds a hair either side of observed reprojection errors."""
    obj, img, R, t = synth.pnp_scene(3000, 7400, noise_px=3.0)
    poses = synth.pnp_hypotheses(16, R, t, 7401)
    for h, i in ((0, 5), (3, 100), (7, 2500)):
        uv = c_oracle.project_points(obj, K4, synth.REF_DIST5, poses[h])
        d = img[i] - uv[i]
        e = np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])
        thr = float(np.sqrt(np.float64(e)))
        for scale in (1.0, 1.0 - 1e-8, 1.0 + 1e-8):
            counts, _, _, allm = pr.scorePnPHypotheses(ctx, obj, img, K4, synth.REF_DIST5, poses,
                                                       thr * scale, want_all_masks=True)
            rc, _, _, rall = c_oracle.score_pnp(obj, img, K4, synth.REF_DIST5, poses, thr * scale,
                                                want_all_masks=True)
            assert np.array_equal(counts, rc) and np.array_equal(allm, rall)


def test_ragged_batch(ctx):
    sizes = [1200, 0, 5, 3000, 64]
    H = 96
    objs, imgs, poses = [], [], []
    for i, m in enumerate(sizes):
        obj, img, R, t = synth.pnp_scene(max(m, 1), 7500 + i)
        objs.append(obj[:m]); imgs.append(img[:m])
        poses.append(synth.pnp_hypotheses(H, R, t, 7600 + i))
    poses = np.stack(poses)
    counts, best, masks = pr.scorePnPBatch(ctx, objs, imgs, K4, synth.REF_DIST5, poses, 8.0)
    for p, m in enumerate(sizes):
        rc, rb, rm, _ = c_oracle.score_pnp(objs[p], imgs[p], K4, synth.REF_DIST5, poses[p], 8.0)
        assert np.array_equal(counts[p], rc) and best[p] == rb and np.array_equal(masks[p], rm)


def test_tilted_sensor_model_is_refused(ctx):
    from slam_indoor_code_b200._capi import Slamb200Error
    obj, img, R, t = synth.pnp_scene(50, 7700)
    poses = synth.pnp_hypotheses(4, R, t, 7701)
    with pytest.raises(Slamb200Error):
        pr.scorePnPHypotheses(ctx, obj, img, K4, list(DIST12) + [0.01, 0.0], poses, 8.0)
    # zero tilt is the identity and is accepted
    counts, _, _, _ = pr.scorePnPHypotheses(ctx, obj, img, K4, list(DIST12) + [0.0, 0.0], poses, 8.0)
    assert np.array_equal(counts, c_oracle.score_pnp(obj, img, K4, DIST12, poses, 8.0)[0])


@pytest.mark.parametrize("c", range(5))
def test_solve_pnp_ransac_drop_in_vs_golden(ctx, golden_dir, c):
    """The whole call the reference makes.  The inlier list (GPU scoring + control) equals the
    committed cv2 result exactly; rvec / tvec come out of OpenCV's CPU Levenberg-Marquardt refit,
    whose last bits depend on the host's SIMD dispatch, so against fixtures made on another machine
    they are held to 1e-7 (the live test below holds them bit-exact against cv2 on the same host)."""
    pytest.importorskip("cv2")
    g = np.load(os.path.join(golden_dir, "pnp.npz"))
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    ok, rvec, tvec, inl = pr.solvePnPRansac(ctx, g[f"obj_{c}"], g[f"img_{c}"], K, g[f"dist_{c}"])
    assert ok == bool(g[f"cv_ok_{c}"])
    assert np.array_equal(inl.reshape(-1), g[f"cv_inliers_{c}"])
    assert np.allclose(rvec.reshape(3), g[f"cv_rvec_{c}"], rtol=1e-7, atol=1e-9)
    assert np.allclose(tvec.reshape(3), g[f"cv_tvec_{c}"], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("m,outl,seed,chunk", [(1500, 0.3, 7800, 8), (5000, 0.5, 7801, 32), (40, 0.2, 7802, 1),
                                               (5, 0.0, 7803, 8), (4, 0.0, 7804, 8), (600, 0.7, 7805, 100)])
def test_solve_pnp_ransac_drop_in_vs_cv2_live(ctx, m, outl, seed, chunk):
    cv2 = pytest.importorskip("cv2")
    obj, img, _, _ = synth.pnp_scene(m, seed, outliers=outl)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    dist = np.array(synth.REF_DIST5).reshape(1, 5)                # the reference's 1x5 CV_64F Mat
    ok_cv, rvec_cv, tvec_cv, inl_cv = cv2.solvePnPRansac(obj, img, K, dist)
    ok, rvec, tvec, inl = pr.solvePnPRansac(ctx, obj, img, K, dist, chunk=chunk)
    assert ok == ok_cv
    if ok:
        assert np.array_equal(inl.reshape(-1), inl_cv.reshape(-1))
        assert np.array_equal(rvec, rvec_cv) and np.array_equal(tvec, tvec_cv)
