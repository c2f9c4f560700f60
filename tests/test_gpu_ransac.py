"""GPU parity tests for hot path B: RANSAC essential-matrix inlier scoring through the C ABI.

Checker: the CPU oracle (pinned to cv2.findEssentialMat's own masks) -- counts, winner and masks
bit-exact.  Mirrors findEssentialMat(p1, p2, K, RANSAC, prob, thr, mask) at
cameraTranslation.cpp:41-46.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, synth
from slam_indoor_code_b200 import camera_translation as ct
from slam_indoor_code_b200.feature_matching import MatcherType

K4 = np.array(synth.SAMSUNG_HV_4K)


def test_cv2_winning_model_mask(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "ransac.npz"))
    for c in range(4):
        counts, best, mask, _ = ct.scoreEssentialHypotheses(ctx, g[f"p1_{c}"], g[f"p2_{c}"], g["K4"],
                                                            g[f"E_cv_{c}"], float(g["threshold_px"]))
        assert best == 0 and np.array_equal(mask, g[f"mask_cv_{c}"])
        assert counts[0] == g[f"mask_cv_{c}"].sum()


def test_five_point_hypothesis_lists(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "ransac.npz"))
    for c in range(4):
        p1, p2, hyp = g[f"p1_{c}"], g[f"p2_{c}"], g[f"hyp_{c}"]
        counts, best, mask, allm = ct.scoreEssentialHypotheses(ctx, p1, p2, g["K4"], hyp, 5.0,
                                                               want_all_masks=True)
        rc, rb, rm, rall = c_oracle.score_essential(p1, p2, g["K4"], hyp, 5.0, want_all_masks=True)
        assert np.array_equal(counts, rc) and best == rb
        assert np.array_equal(mask, rm) and np.array_equal(allm, rall)


@pytest.mark.parametrize("M,H,seed", [(5000, 2048, 5000), (777, 130, 5001), (4, 16, 5002), (513, 1, 5003)])
def test_synthetic_hypotheses(ctx, M, H, seed):
    p1, p2, R, t = synth.two_view(M, seed)
    E = synth.pose_hypotheses(H, R, t, seed + 1)
    for thr in (5.0, 0.5):
        counts, best, mask, _ = ct.scoreEssentialHypotheses(ctx, p1, p2, K4, E, thr)
        rc, rb, rm, _ = c_oracle.score_essential(p1, p2, K4, E, thr)
        assert np.array_equal(counts, rc) and best == rb and np.array_equal(mask, rm)


def test_first_best_wins_and_min_count(ctx):
    p1, p2, R, t = synth.two_view(300, 61)
    E = synth.pose_hypotheses(40, R, t, 62)
    E2 = np.concatenate([E, E, E])
    counts, best, mask, _ = ct.scoreEssentialHypotheses(ctx, p1, p2, K4, E2, 5.0)
    assert best == int(np.argmax(counts)) and best < 40
    counts, best, mask, _ = ct.scoreEssentialHypotheses(ctx, p1[:4], p2[:4], K4, E, 5.0)
    assert best == -1 and not mask.any()
    counts, best, mask, _ = ct.scoreEssentialHypotheses(ctx, p1[:0], p2[:0], K4, E, 5.0)
    assert best == -1 and len(mask) == 0 and not counts.any()


def test_degenerate_models_and_points(ctx):
    """Zero / NaN / huge models and coincident points go through the exact-division path."""
    p1, p2, R, t = synth.two_view(200, 71)
    E = synth.pose_hypotheses(6, R, t, 72)
    E[1] = 0.0
    E[2] = np.nan
    E[3] *= 1e200
    E[4] *= 1e-200
    p1[10] = p2[10] = (K4[2], K4[3])
    counts, best, mask, allm = ct.scoreEssentialHypotheses(ctx, p1, p2, K4, E, 5.0, want_all_masks=True)
    rc, rb, rm, rall = c_oracle.score_essential(p1, p2, K4, E, 5.0, want_all_masks=True)
    assert np.array_equal(counts, rc) and best == rb and np.array_equal(allm, rall)


def test_threshold_boundary(ctx):
    """Matches whose error sits within a few ulps of (float)(thr^2) must fall on the same side."""
    p1, p2, R, t = synth.two_view(4000, 81, noise_px=3.0)
    E = synth.pose_hypotheses(32, R, t, 82)
    err = c_oracle.sampson_errors(p1, p2, K4, E)
    fx, fy = K4[0], K4[1]
    # pick thresholds that land exactly on observed float errors
    for h, i in ((0, 5), (3, 100), (7, 2500)):
        t_f = err[h, i]
        thr_px = float(np.sqrt(np.float64(t_f))) * ((fx + fy) / 2)
        for scale in (1.0, 1.0 - 1e-8, 1.0 + 1e-8):
            counts, best, mask, allm = ct.scoreEssentialHypotheses(ctx, p1, p2, K4, E, thr_px * scale,
                                                                   want_all_masks=True)
            rc, rb, rm, rall = c_oracle.score_essential(p1, p2, K4, E, thr_px * scale,
                                                        want_all_masks=True)
            assert np.array_equal(counts, rc) and np.array_equal(allm, rall)


def test_ragged_batch(ctx):
    sizes = [1200, 0, 5, 3000, 64]
    H = 96
    p1s, p2s, Es = [], [], []
    for i, m in enumerate(sizes):
        p1, p2, R, t = synth.two_view(max(m, 1), 90 + i)
        p1s.append(p1[:m]); p2s.append(p2[:m]); Es.append(synth.pose_hypotheses(H, R, t, 190 + i))
    counts, best, masks = ct.scoreEssentialBatch(ctx, p1s, p2s, K4, np.stack(Es), 5.0)
    for i in range(len(sizes)):
        rc, rb, rm, _ = c_oracle.score_essential(p1s[i], p2s[i], K4, Es[i], 5.0)
        assert np.array_equal(counts[i], rc) and best[i] == rb and np.array_equal(masks[i], rm)


def test_match_then_score_chain(ctx):
    """cfg3 shape: batch matching, device-side getKeyPointCoordsFromFramePair, scoring."""
    rng = np.random.default_rng(11)
    nq, P, H = 1500, 3, 64
    q = synth.sift_like(nq, 1300)
    kq = rng.uniform(0, 3840, (nq, 2)).astype(np.float32)
    trains, kts = [], []
    for p in range(P):
        nt = 1400 + 100 * p
        trains.append(synth.sift_train_from_query(q, nt, 1301 + p))
        kts.append(rng.uniform(0, 2160, (nt, 2)).astype(np.float32))
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    KQ = ctx.upload_keypoints(kq)
    KTs = [ctx.upload_keypoints(k) for k in kts]
    _, _, R, t = synth.two_view(10, 1310)
    E = np.stack([synth.pose_hypotheses(H, R, t, 1320 + p) for p in range(P)])
    ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7)
    ct.scoreBatchEnqueue(ctx, KQ, KTs, K4, E, 50.0)
    matches, n_out = ctx.batchFetch()
    counts, best, mask = ct.batchScoresFetch(ctx)
    for p in range(P):
        ref = c_oracle.match_features(0, q, trains[p], 0.7)
        assert np.array_equal(matches[p], ref)
        p1, p2 = c_oracle.gather_points(kq, kts[p], ref)
        rc, rb, rm, _ = c_oracle.score_essential(p1, p2, K4, E[p], 50.0)
        assert np.array_equal(counts[p], rc) and best[p] == rb
        assert np.array_equal(mask[p, : len(ref)], rm)


@pytest.mark.parametrize("m,noise,outl,seed", [(600, 0.7, 0.3, 7000), (5000, 0.7, 0.3, 7003),
                                               (3000, 3.0, 0.7, 7005)])
def test_find_essential_mat_drop_in(ctx, m, noise, outl, seed):
    """findEssentialMat(p1, p2, K, RANSAC, 0.999, 5.0, mask) end to end: host RANSAC control +
    CPU 5-point solver + GPU scoring == cv2.findEssentialMat, E and mask bit for bit."""
    cv2 = pytest.importorskip("cv2")
    p1, p2, _, _ = synth.two_view(m, seed, noise_px=noise, outliers=outl)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    Ecv, mcv = cv2.findEssentialMat(p1, p2, Kmat, cv2.RANSAC, 0.999, 5.0)
    E, mask = ct.findEssentialMat(ctx, p1, p2, Kmat, 0.999, 5.0)
    assert np.array_equal(np.asarray(Ecv, np.float64).reshape(3, 3), E)
    assert np.array_equal(mcv, mask)


def test_cfg5_full_batch_properties(ctx):
    """BASELINE cfg5 at full size: 210 pairs x 2048 hypotheses x 5000 matches in one call; sampled
    pairs against the oracle, and the size-independent properties on all of them (the winner has
    the maximal count, first best wins, its mask sums to its count)."""
    P, H, M = 210, 2048, 5000
    p1s, p2s, Es = [], [], []
    base = [synth.two_view(M, 5000 + k) for k in range(6)]
    hyp = [synth.pose_hypotheses(H, b[2], b[3], 5100 + k) for k, b in enumerate(base)]
    for p in range(P):
        p1s.append(base[p % 6][0]); p2s.append(base[p % 6][1]); Es.append(hyp[p % 6])
    counts, best, masks = ct.scoreEssentialBatch(ctx, p1s, p2s, K4, np.stack(Es), 5.0)
    assert counts.shape == (P, H)
    for p in (0, 1, 5, 100, 209):
        rc, rb, rm, _ = c_oracle.score_essential(p1s[p], p2s[p], K4, Es[p], 5.0)
        assert np.array_equal(counts[p], rc) and best[p] == rb and np.array_equal(masks[p], rm)
    for p in range(P):
        b = int(best[p])
        assert b == int(np.argmax(counts[p])) and counts[p, b] > 4
        assert int(masks[p].sum()) == int(counts[p, b])
        assert np.array_equal(counts[p], counts[p % 6])          # identical inputs, identical counts
