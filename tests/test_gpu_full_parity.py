"""GPU parity at BASELINE.json's FULL sizes against cv2 itself -- every row, not a sample.

The reference's arithmetic owner is OpenCV (featureMatchingCPU.cpp:27-40 creates the matcher and
calls knnMatch; featureMatchingCommon.cpp:43-49 is the ratio loop; cameraTranslation.cpp:41-46 the
RANSAC call).  The GPU box runs the same image as the build container, so cv2 is importable there
and these tests compare the C ABI's results with live cv2 results on the same seeded inputs:
train indices, float distance BIT PATTERNS and ratio-test verdicts of every query row.

  cfg1  single 10k x 10k SIFT pair, integer-valued rows and general floats     whole pair
  cfg2  single 10k x 10k ORB pair                                              whole pair
  cfg3  1 x 210 window of 10k-row frames: all 210 pairs whole;
        and the chain match -> device gather -> Sampson scoring on all 210 pairs with keypoints
        that are geometrically consistent with the planted correspondences
  cfg4  8 frames x 50,000 rows, all 28 pairs: one whole pair + 512 rows of each of the other 27
  cfg5  2048 hypotheses from cv2's own 5-point solver x 5000 matches: 20 whole pairs (counts,
        winner, mask) inside one 210-pair call
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip("cv2")

import synth_inputs as synth
from oracle import c_oracle
from slam_indoor_code_b200 import camera_translation as ct
from slam_indoor_code_b200._capi import DMATCH
from slam_indoor_code_b200.feature_matching import MatcherType

K4 = synth.SAMSUNG_HV_4K
RATIO = 0.7


def cv_knn2(q, t, norm):
    """cv2.BFMatcher(norm).knnMatch(q, t, 2) as idx[nq,2] / dist[nq,2] arrays."""
    res = cv2.BFMatcher(norm).knnMatch(q, t, 2)
    idx = np.full((len(res), 2), -1, np.int32)
    dist = np.zeros((len(res), 2), np.float32)
    for i, row in enumerate(res):
        for k, m in enumerate(row):
            idx[i, k] = m.trainIdx
            dist[i, k] = m.distance
    return idx, dist


def get_good_matches(idx, dist, ratio=RATIO):
    """getGoodMatches (featureMatchingCommon.cpp:37-50): float promoted to double, strict '<'."""
    keep = (idx[:, 1] >= 0) & (dist[:, 0].astype(np.float64) < ratio * dist[:, 1].astype(np.float64))
    out = np.zeros(int(keep.sum()), DMATCH)
    out["queryIdx"] = np.nonzero(keep)[0]
    out["trainIdx"] = idx[keep, 0]
    out["distance"] = dist[keep, 0]
    return out


def assert_same_matches(got, want):
    assert len(got) == len(want)
    assert np.array_equal(got["queryIdx"], want["queryIdx"])
    assert np.array_equal(got["trainIdx"], want["trainIdx"])
    assert np.array_equal(got["distance"].view(np.int32), want["distance"].view(np.int32))
    assert np.all(got["imgIdx"] == 0)


def whole_pair(ctx, matcher, norm, q, t, ratio=RATIO):
    Q, T = ctx.upload(q), ctx.upload(t)
    idx, dist = ctx.knnMatch(matcher, Q, T)
    good = ctx.matchFeatures(Q, T, matcher, ratio)
    Q.free(); T.free()
    ridx, rdist = cv_knn2(q, t, norm)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(dist.view(np.int32), rdist.view(np.int32))
    assert_same_matches(good, get_good_matches(ridx, rdist, ratio))
    return good


def test_cfg1_sift_integer_whole_pair(ctx):
    q, t = synth.sift_pair(10000, 10000, 1001)
    good = whole_pair(ctx, MatcherType.SIFT_BF, cv2.NORM_L2, q, t)
    assert 2500 < len(good) < 4000
    # the FLANN flag answers with the exact search: identical to the BF ground truth (recall 1.0)
    Q, T = ctx.upload(q), ctx.upload(t)
    assert_same_matches(ctx.matchFeatures(Q, T, MatcherType.SIFT_FLANN, RATIO), good)


def test_cfg1_sift_general_float_whole_pair(ctx):
    q, t = synth.float_pair(10000, 10000, 1002)
    whole_pair(ctx, MatcherType.SIFT_BF, cv2.NORM_L2, q, t, ratio=0.95)


def test_cfg1_rootsift_whole_pair(ctx):
    """RootSIFT (the realistic non-integer case): sqrt of the L1-normalised rows."""
    qi, ti = synth.sift_pair(10000, 10000, 1003)
    root = lambda d: np.sqrt(d / np.maximum(d.sum(1, keepdims=True), 1)).astype(np.float32)
    good = whole_pair(ctx, MatcherType.SIFT_BF, cv2.NORM_L2, root(qi), root(ti))
    assert len(good) > 2000


def test_cfg2_orb_whole_pair(ctx):
    q, t = synth.orb_pair(10000, 10000, 2001)
    good = whole_pair(ctx, MatcherType.ORB_BF, cv2.NORM_HAMMING, q, t)
    assert len(good) > 2000
    ctx.debug_orb_kernel(tensor_cores=False)       # the XOR/POPC kernel the north star names
    try:
        whole_pair(ctx, MatcherType.ORB_BF, cv2.NORM_HAMMING, q, t)
    finally:
        ctx.debug_orb_kernel(tensor_cores=True)


@pytest.fixture(scope="module")
def cfg3(ctx):
    q = synth.sift_like(10000, 3000)
    trains = [synth.sift_train_from_query(q, 10000, 3001 + p) for p in range(210)]
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    # cv2 on every one of the 210 pairs (0.1-0.15 s each on the box's cores): the whole window's
    # ground truth, not a sample
    want = [get_good_matches(*cv_knn2(q, t, cv2.NORM_L2)) for t in trains]
    yield q, trains, Q, Ts, want
    for h in Ts + [Q]:
        h.free()


def test_cfg3_window_all_210_whole_pairs(ctx, cfg3):
    q, trains, Q, Ts, want = cfg3
    res = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, RATIO)
    assert len(res) == 210
    for p in range(210):
        assert_same_matches(res[p], want[p])
        assert len(want[p]) > 2000
    # the device-resident form the bench times returns the same lists
    ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, RATIO)
    got, n_out = ctx.batchFetch()
    for p in range(210):
        assert_same_matches(got[p], want[p])


def test_cfg3_chain_match_gather_score_all_pairs(ctx, cfg3):
    """BASELINE cfg3 as stated: matching + RANSAC essential scoring over the whole window.  The
    keypoints are consistent with the planted correspondences (a real two-view geometry per pair),
    so the winning hypothesis explains most accepted matches.  Checker: the oracle's gather +
    Sampson scoring on cv2's match lists."""
    q, trains, Q, Ts, want = cfg3
    H = 256
    kq, kts, poses = synth.window_geometry(10000, 10000, [3001 + p for p in range(210)], 3500)
    KQ = ctx.upload_keypoints(kq)
    KTs = [ctx.upload_keypoints(k) for k in kts]
    E = np.stack([synth.pose_hypotheses(H, R, t, 3600 + p) for p, (R, t) in enumerate(poses)])
    ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, RATIO)
    ct.scoreBatchEnqueue(ctx, KQ, KTs, K4, E, 5.0)
    matches, n_out = ctx.batchFetch()
    counts, best, mask = ct.batchScoresFetch(ctx)
    for p in range(210):
        ref = want[p]
        assert_same_matches(matches[p], ref)
        p1, p2 = c_oracle.gather_points(kq, kts[p], ref)
        rc, rb, rm, _ = c_oracle.score_essential(p1, p2, K4, E[p], 5.0)
        assert np.array_equal(counts[p], rc), p
        assert best[p] == rb, p
        assert np.array_equal(mask[p, : len(ref)], rm), p
        assert rc[rb] > 0.8 * len(ref), (p, rc[rb], len(ref))     # the geometry is real
    for h in KTs + [KQ]:
        h.free()


def test_cfg4_window_8_frames_50k(ctx):
    base = synth.sift_like(50000, 4000)
    frames = [base]
    for f in range(1, 8):
        fr = synth.sift_like(50000, 4000 + f)
        rng = np.random.default_rng(4100 + f)
        rows = rng.choice(50000, 12000, replace=False)
        fr[rows] = np.clip(base[rows] + rng.integers(-5, 6, (12000, 128)), 0, 255).astype(np.float32)
        frames.append(fr)
    sets = [ctx.upload(f) for f in frames]
    res = ctx.matchWindow(sets, MatcherType.SIFT_BF, RATIO)
    assert len(res) == 28
    rows = np.sort(np.random.default_rng(41).choice(50000, 512, replace=False))
    for (i, j), m in res.items():
        if (i, j) == (2, 5):
            ridx, rdist = cv_knn2(frames[i], frames[j], cv2.NORM_L2)        # the whole 50k x 50k pair
            assert_same_matches(m, get_good_matches(ridx, rdist))
            idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, sets[i], sets[j])
            assert np.array_equal(idx, ridx) and np.array_equal(dist.view(np.int32), rdist.view(np.int32))
            continue
        ridx, rdist = cv_knn2(frames[i][rows], frames[j], cv2.NORM_L2)
        want = get_good_matches(ridx, rdist)
        want["queryIdx"] = rows[want["queryIdx"]]
        assert_same_matches(m[np.isin(m["queryIdx"], rows)], want)
        assert np.all(np.diff(m["queryIdx"]) > 0)
    assert len(res[(0, 3)]) > 8000
    for s in sets:
        s.free()


def five_point_hypotheses(p1, p2, h, seed):
    """H essential-matrix candidates from cv2's own 5-point solver on seeded 5-subsets (a 5-point
    call returns every candidate stacked), padded / truncated to h rows."""
    rng = np.random.default_rng(seed)
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float64)
    out = []
    n = 0
    while n < h:
        s = rng.choice(p1.shape[0], 5, replace=False)
        E = cv2.findEssentialMat(p1[s], p2[s], Kmat, cv2.RANSAC, 0.999, 1.0)[0]
        if E is None:
            continue
        E = np.asarray(E, np.float64).reshape(-1, 9)
        out.append(E)
        n += len(E)
    return np.concatenate(out)[:h].copy()


def test_cfg5_five_point_hypotheses_whole_pairs(ctx):
    P, H, M, D = 210, 2048, 5000, 20
    scenes = [synth.two_view(M, 5000 + k) for k in range(D)]
    hyps = [five_point_hypotheses(s[0], s[1], H, 5200 + k) for k, s in enumerate(scenes)]
    p1s = [scenes[p % D][0] for p in range(P)]
    p2s = [scenes[p % D][1] for p in range(P)]
    counts, best, masks = ct.scoreEssentialBatch(ctx, p1s, p2s, K4, np.stack([hyps[p % D] for p in range(P)]), 5.0)
    assert counts.shape == (P, H)
    for k in range(D):
        rc, rb, rm, _ = c_oracle.score_essential(scenes[k][0], scenes[k][1], K4, hyps[k], 5.0)
        # minimal-sample models: a good share of them explains the inlier set
        assert rc.max() > 0.55 * M
        for p in (k, k + 100, k + 180):
            assert np.array_equal(counts[p], rc) and best[p] == rb and np.array_equal(masks[p], rm), (k, p)
    # one pair with every (hypothesis, match) verdict
    c1, b1, m1, allm = ct.scoreEssentialHypotheses(ctx, scenes[3][0], scenes[3][1], K4, hyps[3], 5.0,
                                                   want_all_masks=True)
    rc, rb, rm, rall = c_oracle.score_essential(scenes[3][0], scenes[3][1], K4, hyps[3], 5.0, want_all_masks=True)
    assert np.array_equal(c1, rc) and b1 == rb and np.array_equal(m1, rm) and np.array_equal(allm, rall)
