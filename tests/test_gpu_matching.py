"""GPU parity tests for hot path A (kNN-2 + Lowe ratio) -- all through the C ABI.

Checker: the CPU oracle (oracle/, pinned to cv2) and the committed cv2 fixtures.  Bit-exact:
train indices, accept/reject, float distance bit patterns.  Mirrors the reference's call shape
matchFeatures(prevDesc, curDesc, matches, type) (featureMatchingCPU.cpp:17-43).
"""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, synth
from slam_indoor_code_b200 import _capi
from slam_indoor_code_b200.feature_matching import (MatcherType, getMatcherTypeIndex,
                                                    matchFramesPairFeatures)

SIFT_GOLD = ["sift_int", "sift_float", "sift_ties", "sift_t1", "sift_t2", "sift_q1"]


def _f32(a):
    return a if a.dtype == np.float32 else a.astype(np.float32)


def _check_knn(idx, dist, ridx, rdist):
    assert np.array_equal(idx, ridx)
    valid = ridx >= 0
    assert np.array_equal(dist[valid].view(np.int32), rdist[valid].view(np.int32))


def _knn_and_matches(ctx, matcher, q, t, ratio=0.7):
    Q, T = ctx.upload(q), ctx.upload(t)
    idx, dist = ctx.knnMatch(matcher, Q, T)
    good = ctx.matchFeatures(Q, T, matcher, ratio)
    Q.free(); T.free()
    return idx, dist, good


@pytest.mark.parametrize("name", SIFT_GOLD)
@pytest.mark.parametrize("matcher", [MatcherType.SIFT_BF, MatcherType.SIFT_FLANN])
def test_sift_golden(ctx, golden_dir, name, matcher):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    q, t = _f32(g["q"]), _f32(g["t"])
    idx, dist, good = _knn_and_matches(ctx, matcher, q, t)
    _check_knn(idx, dist, g["idx"], g["dist"])
    assert np.array_equal(good, c_oracle.ratio_test(g["idx"], g["dist"], 0.7))


@pytest.mark.parametrize("name", ["orb", "orb_t1"])
def test_orb_golden(ctx, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    idx, dist, good = _knn_and_matches(ctx, MatcherType.ORB_BF, g["q"], g["t"])
    _check_knn(idx, dist, g["idx"], g["dist"])
    assert np.array_equal(good, c_oracle.ratio_test(g["idx"], g["dist"], 0.7))


@pytest.mark.parametrize("nq,nt,seed", [(1000, 1500, 101), (2049, 777, 102), (130, 4100, 103)])
def test_sift_int_seeded(ctx, nq, nt, seed):
    q, t = synth.sift_pair(nq, nt, seed)
    idx, dist, good = _knn_and_matches(ctx, MatcherType.SIFT_BF, q, t)
    ridx, rdist = c_oracle.l2_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    ref = c_oracle.ratio_test(ridx, rdist, 0.7)
    assert np.array_equal(good, ref) and len(ref) > 0.2 * min(nq, nt)


@pytest.mark.parametrize("nq,nt,seed", [(700, 900, 201), (257, 2300, 202)])
def test_sift_float_seeded(ctx, nq, nt, seed):
    """General floats: the low bits of the distances depend on cv2's summation order."""
    q, t = synth.float_pair(nq, nt, seed)
    idx, dist, good = _knn_and_matches(ctx, MatcherType.SIFT_BF, q, t, ratio=0.95)
    ridx, rdist = c_oracle.l2_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.95))


def test_sift_mixed_exactness(ctx):
    """An integer-valued query against a float train set (and the reverse) takes the exact path."""
    qi, ti = synth.sift_pair(300, 400, 301)
    qf, tf = synth.float_pair(300, 400, 302)
    for q, t in ((qi, tf), (qf, ti)):
        idx, dist, _ = _knn_and_matches(ctx, MatcherType.SIFT_BF, q, t)
        _check_knn(idx, dist, *c_oracle.l2_knn2(q, t))


def test_sift_large_norm_integers(ctx):
    """Integer rows whose squared norm exceeds 2^20 are not 'exact mode' but must still match."""
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (300, 128)).astype(np.float32)
    t = rng.integers(0, 256, (500, 128)).astype(np.float32)
    t[17] = q[3]
    Q, T = ctx.upload(q), ctx.upload(t)
    assert Q.exact_mode == 0
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T)
    _check_knn(idx, dist, *c_oracle.l2_knn2(q, t))
    assert ctx.upload(synth.sift_like(64, 1)).exact_mode == 1


@pytest.mark.parametrize("nq,nt,seed", [(1500, 2500, 401), (4097, 300, 402), (100, 5000, 403)])
def test_orb_seeded(ctx, nq, nt, seed):
    q, t = synth.orb_pair(nq, nt, seed)
    idx, dist, good = _knn_and_matches(ctx, MatcherType.ORB_BF, q, t)
    ridx, rdist = c_oracle.hamming_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.7))


def test_orb_many_ties(ctx):
    """Few distinct descriptors -> masses of equal distances: lowest train index must win."""
    rng = np.random.default_rng(6)
    base = rng.integers(0, 256, (5, 32), dtype=np.uint8)
    q = base[rng.integers(0, 5, 400)]
    t = base[rng.integers(0, 5, 3000)]
    idx, dist, _ = _knn_and_matches(ctx, MatcherType.ORB_BF, q, t)
    _check_knn(idx, dist, *c_oracle.hamming_knn2(q, t))


def test_sift_many_ties(ctx):
    rng = np.random.default_rng(7)
    base = synth.sift_like(6, 8)
    q = base[rng.integers(0, 6, 300)]
    t = base[rng.integers(0, 6, 2600)]
    idx, dist, _ = _knn_and_matches(ctx, MatcherType.SIFT_BF, q, t)
    _check_knn(idx, dist, *c_oracle.l2_knn2(q, t))


@pytest.mark.parametrize("matcher", [MatcherType.SIFT_BF, MatcherType.ORB_BF])
def test_empty_and_single_row_sets(ctx, matcher):
    if matcher == MatcherType.ORB_BF:
        q, t = synth.orb_pair(9, 5, 501, planted=0)
        e = np.zeros((0, 32), np.uint8)
    else:
        q, t = synth.sift_pair(9, 5, 501, planted=0)
        e = np.zeros((0, 128), np.float32)
    Q, T, E = ctx.upload(q), ctx.upload(t), ctx.upload(e)
    # empty query -> no rows, no matches (cv2: empty result)
    idx, dist = ctx.knnMatch(matcher, E, T)
    assert idx.shape == (0, 2)
    assert len(ctx.matchFeatures(E, T, matcher)) == 0
    # empty train -> Q empty lists
    idx, dist = ctx.knnMatch(matcher, Q, E)
    assert np.all(idx == -1)
    assert len(ctx.matchFeatures(Q, E, matcher)) == 0
    # one train row -> one-element lists; getGoodMatches' read of [1] is UB: defined as reject
    T1 = ctx.upload(t[:1])
    idx, dist = ctx.knnMatch(matcher, Q, T1)
    assert np.all(idx[:, 0] == 0) and np.all(idx[:, 1] == -1)
    assert len(ctx.matchFeatures(Q, T1, matcher)) == 0


def test_row_pitch_is_honoured(ctx):
    """cv::Mat::step: rows 640 B apart (SIFT) / 48 B apart (ORB)."""
    q, t = synth.sift_pair(200, 300, 601)
    wide = np.zeros((300, 160), np.float32)
    wide[:, :128] = t
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, ctx.upload(q), ctx.upload(wide[:, :128]))
    _check_knn(idx, dist, *c_oracle.l2_knn2(q, t))
    q, t = synth.orb_pair(200, 300, 602)
    wide = np.zeros((300, 48), np.uint8)
    wide[:, :32] = t
    idx, dist = ctx.knnMatch(MatcherType.ORB_BF, ctx.upload(q), ctx.upload(wide[:, :32]))
    _check_knn(idx, dist, *c_oracle.hamming_knn2(q, t))


def test_ratio_values(ctx):
    q, t = synth.sift_pair(600, 700, 701)
    ridx, rdist = c_oracle.l2_knn2(q, t)
    Q, T = ctx.upload(q), ctx.upload(t)
    for r in (0.0, 0.5, 0.7, 0.8, 1.0, 1.5):
        assert np.array_equal(ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, r),
                              c_oracle.ratio_test(ridx, rdist, r))


def test_bad_matcher_type_and_kind_mismatch(ctx):
    """featureMatchingCPU.cpp:36-37 throws on an unknown type; a Mat of the wrong type asserts."""
    q, t = synth.sift_pair(10, 10, 801)
    Q, T = ctx.upload(q), ctx.upload(t)
    with pytest.raises(RuntimeError):
        ctx.matchFeatures(Q, T, 7)
    out = np.zeros(16, _capi.DMATCH)
    n = np.zeros(1, np.int32)
    rc = ctx._lib.slamb200_match_pair(ctx._h, 7, Q._h, T._h, 0.7, _capi.ptr(out), 16, _capi.ptr(n))
    assert rc == _capi.ERR_MATCHER
    rc = ctx._lib.slamb200_match_pair(ctx._h, 2, Q._h, T._h, 0.7, _capi.ptr(out), 16, _capi.ptr(n))
    assert rc == _capi.ERR_KIND
    rc = ctx._lib.slamb200_match_pair(ctx._h, 0, Q._h, T._h, 0.7, _capi.ptr(out), 4, _capi.ptr(n))
    assert rc == _capi.ERR_INVALID                       # capacity below rows(query)
    assert getMatcherTypeIndex({"useFM-SIFT-BF": True, "useFM-ORB": True}) == MatcherType.SIFT_BF
    assert getMatcherTypeIndex({"useFM-SIFT-FLANN": True, "useFM-ORB": True}) == MatcherType.SIFT_FLANN
    with pytest.raises(RuntimeError):
        getMatcherTypeIndex({})


@pytest.mark.parametrize("matcher", [MatcherType.SIFT_BF, MatcherType.ORB_BF])
def test_batch_window_equals_pairs(ctx, matcher):
    """batch.cpp:120-148: one query frame against a ragged batch of train frames."""
    sizes = [700, 1, 0, 1300, 256, 2]
    if matcher == MatcherType.ORB_BF:
        q = synth.orb_pair(900, 8, 900)[0]
        trains = [synth.orb_pair(8, n, 901 + i)[1] if n else np.zeros((0, 32), np.uint8)
                  for i, n in enumerate(sizes)]
    else:
        q = synth.sift_like(900, 900)
        trains = [synth.sift_train_from_query(q, n, 901 + i) if n else np.zeros((0, 128), np.float32)
                  for i, n in enumerate(sizes)]
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    got = ctx.matchBatch(Q, Ts, matcher, 0.7)
    for t, g in zip(trains, got):
        assert np.array_equal(g, c_oracle.match_features(int(matcher), q, t, 0.7))
    ctx.matchBatchEnqueue(Q, Ts, matcher, 0.7)
    got2, n_out = ctx.batchFetch()
    for a, b in zip(got, got2):
        assert np.array_equal(a, b)
    assert len(ctx.matchBatch(Q, [], matcher, 0.7)) == 0


def test_frame_window_all_pairs(ctx):
    frames = [synth.sift_like(300 + 50 * i, 1000 + i) for i in range(4)]
    for i in range(1, 4):                                   # plant overlap with frame 0
        frames[i][:100] = frames[0][:100]
    sets = [ctx.upload(f) for f in frames]
    res = ctx.matchWindow(sets, MatcherType.SIFT_BF, 0.7)
    assert sorted(res) == [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    for (i, j), m in res.items():
        assert np.array_equal(m, c_oracle.match_features(0, frames[i], frames[j], 0.7))


def test_concurrent_host_threads_share_query(ctx):
    """batch.cpp:181-201: threadsCount std::threads, one shared read-only query descriptor."""
    q = synth.sift_like(800, 1100)
    trains = [synth.sift_train_from_query(q, 900 + 10 * i, 1101 + i) for i in range(8)]
    Q = ctx.upload(q)
    res, errs = [None] * 8, []

    def work(i):
        try:
            T = ctx.upload(trains[i])
            res[i] = ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7)
            T.free()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for i in range(8):
        assert np.array_equal(res[i], c_oracle.match_features(0, q, trains[i], 0.7))


def test_match_frames_pair_features_entry_point(ctx):
    """featureMatching.h:47-53 shape: descriptors in, good matches out."""
    q, t = synth.sift_pair(500, 600, 1201)
    got = matchFramesPairFeatures(ctx, q, t, MatcherType.SIFT_FLANN, 0.7)
    assert np.array_equal(got, c_oracle.match_features(1, q, t, 0.7))
    q, t = synth.orb_pair(500, 600, 1202)
    got = matchFramesPairFeatures(ctx, q, t, MatcherType.ORB_BF, 0.7)
    assert np.array_equal(got, c_oracle.match_features(2, q, t, 0.7))


# ---- BASELINE.json full sizes: size-independent properties --------------------------------------
def _properties(ctx, matcher, q, t):
    Q, T = ctx.upload(q), ctx.upload(t)
    idx, dist = ctx.knnMatch(matcher, Q, T)
    good = ctx.matchFeatures(Q, T, matcher, 0.7)
    n = q.shape[0]
    assert idx.shape == (n, 2) and idx.min() >= 0 and idx.max() < t.shape[0]
    assert np.all(dist[:, 0] <= dist[:, 1]) and np.all(idx[:, 0] != idx[:, 1])
    assert np.all(np.diff(good["queryIdx"]) > 0)                       # ascending queryIdx
    keep = dist[:, 0].astype(np.float64) < 0.7 * dist[:, 1].astype(np.float64)
    assert np.array_equal(good["queryIdx"], np.nonzero(keep)[0])
    assert np.array_equal(good["trainIdx"], idx[keep, 0])
    # spot-check 64 rows against the oracle (full rows, all 10k trains)
    rows = np.random.default_rng(1).choice(n, 64, replace=False)
    ridx, rdist = (c_oracle.hamming_knn2 if matcher == MatcherType.ORB_BF else c_oracle.l2_knn2)(q[rows], t)
    _check_knn(idx[rows], dist[rows], ridx, rdist)
    # train-set permutation: distances invariant, indices mapped
    perm = np.random.default_rng(2).permutation(t.shape[0])
    Tp = ctx.upload(np.ascontiguousarray(t[perm]))
    idx2, dist2 = ctx.knnMatch(matcher, Q, Tp)
    assert np.array_equal(dist2.view(np.int32), dist.view(np.int32))
    untied = dist[:, 0] != dist[:, 1]
    assert np.array_equal(perm[idx2[untied, 0]], idx[untied, 0])
    # self match: every row finds itself at distance 0 (lowest index among duplicates)
    idx3, dist3 = ctx.knnMatch(matcher, T, T)
    assert np.all(dist3[:, 0] == 0) and np.all(idx3[:, 0] <= np.arange(t.shape[0]))
    return good


def test_cfg1_sift_10k_properties(ctx):
    q, t = synth.sift_pair(10000, 10000, 1001)
    good = _properties(ctx, MatcherType.SIFT_BF, q, t)
    assert 2500 < len(good) < 4000                                     # ~30 % planted


def test_cfg2_orb_10k_properties(ctx):
    q, t = synth.orb_pair(10000, 10000, 2001)
    good = _properties(ctx, MatcherType.ORB_BF, q, t)
    assert len(good) > 2000


def test_cfg1_float_10k_spot(ctx):
    q, t = synth.float_pair(10000, 10000, 1002)
    Q, T = ctx.upload(q), ctx.upload(t)
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T)
    rows = np.random.default_rng(3).choice(10000, 48, replace=False)
    _check_knn(idx[rows], dist[rows], *c_oracle.l2_knn2(q[rows], t))


def test_pinned_pipelined_upload(ctx):
    """slamb200_upload_desc_pinned: rows in page-locked host memory are read by the prep kernel
    directly (no staging copy); results must not change, with and without a row pitch."""
    import torch
    q, t = synth.sift_pair(1500, 2100, 1401)
    hq = torch.empty(q.shape, dtype=torch.float32).pin_memory()
    hq.numpy()[...] = q
    wide = torch.zeros((t.shape[0], 160), dtype=torch.float32).pin_memory()
    wide.numpy()[:, :128] = t
    Q = ctx.upload_pinned(hq.numpy())
    T = ctx.upload_pinned(wide.numpy()[:, :128])
    got = ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7)
    assert np.array_equal(got, c_oracle.match_features(0, q, t, 0.7))
    # pageable memory through the same entry point falls back to a copy
    Q2 = ctx.upload_pinned(q.copy())
    ctx.synchronize()
    assert np.array_equal(ctx.matchFeatures(Q2, T, MatcherType.SIFT_BF, 0.7), got)


def test_cfg4_large_frames_window(ctx):
    """cfg4 shape: 50,000 descriptors per frame (196 train tiles, 196 query blocks), all pairs of
    a 3-frame window; property checks + oracle spot rows."""
    frames = [synth.sift_like(50000, 4000 + f) for f in range(3)]
    frames[1][:2000] = np.clip(frames[0][:2000] + 3, 0, 255)
    frames[2][1000:4000] = frames[0][1000:4000]
    sets = [ctx.upload(f) for f in frames]
    res = ctx.matchWindow(sets, MatcherType.SIFT_BF, 0.7)
    rows = np.random.default_rng(9).choice(50000, 24, replace=False)
    rows[:4] = [0, 1999, 1000, 3999]
    for (i, j), m in res.items():
        assert np.all(np.diff(m["queryIdx"]) > 0)
        ridx, rdist = c_oracle.l2_knn2(frames[i][rows], frames[j])
        ref = c_oracle.ratio_test(ridx, rdist, 0.7)
        got = m[np.isin(m["queryIdx"], rows)]
        order = np.argsort(rows)
        exp_q = rows[order][np.isin(np.arange(len(rows))[order], ref["queryIdx"])]
        assert sorted(got["queryIdx"].tolist()) == sorted(rows[ref["queryIdx"]].tolist())
        for r in ref:
            g = got[got["queryIdx"] == rows[r["queryIdx"]]][0]
            assert g["trainIdx"] == r["trainIdx"] and g["distance"] == r["distance"]
    assert len(res[(0, 2)]) >= 2500
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, sets[0], sets[1])
    _check_knn(idx[rows], dist[rows], *c_oracle.l2_knn2(frames[0][rows], frames[1]))


def test_general_float_certificate_and_fallback(ctx):
    """General floats go through the two-term-split tensor-core kernel; its answer is only used
    when a rigorous error bound certifies it.  Near-duplicate clusters make hundreds of train rows
    (spread over many chunks) indistinguishable within that bound, so those query rows must take
    the exact full-row fallback -- and everything must still equal cv2's result bit for bit."""
    import ctypes
    rng = np.random.default_rng(77)
    q, t = synth.float_pair(600, 6000, 1501)
    # 40 queries each get 150 near-copies scattered over the whole train set
    for k in range(40):
        rows = rng.choice(6000, 150, replace=False)
        t[rows] = q[k] + rng.normal(0, 2e-3, (150, 128)).astype(np.float32)
    # exact duplicates too (ties -> lowest index)
    t[100] = t[4000] = t[5999] = q[50]
    Q, T = ctx.upload(q), ctx.upload(t)
    assert Q.exact_mode == 0
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T)
    ridx, rdist = c_oracle.l2_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    good = ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.9)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.9))
    # the fallback did run (device-resident batch so that the debug counter refers to this call)
    ctx.matchBatchEnqueue(Q, [T], MatcherType.SIFT_BF, 0.9)
    lib = ctx._lib
    lib.slamb200_dbg_last_fallback_rows.argtypes = [ctypes.c_void_p]
    n_fb = lib.slamb200_dbg_last_fallback_rows(ctx._h)
    assert 30 <= n_fb <= 600, n_fb
    got, _ = ctx.batchFetch()
    assert np.array_equal(got[0], c_oracle.ratio_test(ridx, rdist, 0.9))


def test_general_float_scales_and_signs(ctx):
    """Negative values, tiny and huge magnitudes, unit-norm rows (RootSIFT-like)."""
    rng = np.random.default_rng(78)
    base = rng.standard_normal((900, 128)).astype(np.float32)
    for scale in (1e-3, 1.0, 3e4):
        q = (base[:300] * scale).astype(np.float32)
        t = (base[200:] * scale + rng.standard_normal((700, 128)).astype(np.float32) * scale * 0.05).astype(np.float32)
        idx, dist, _ = _knn_and_matches(ctx, MatcherType.SIFT_BF, q, t)
        _check_knn(idx, dist, *c_oracle.l2_knn2(q, t))
    qi, ti = synth.sift_pair(500, 800, 1502)
    rs = lambda x: np.sqrt(x / np.maximum(x.sum(1, keepdims=True), 1e-9)).astype(np.float32)
    idx, dist, good = _knn_and_matches(ctx, MatcherType.SIFT_BF, rs(qi), rs(ti))
    ridx, rdist = c_oracle.l2_knn2(rs(qi), rs(ti))
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.7))


# ---- NORM_L1 mode: useFM-SIFT-BF as the reference's OpenCV-CUDA build reads it --
@pytest.mark.parametrize("name", SIFT_GOLD)
def test_sift_l1_golden(ctx, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    l1 = np.load(os.path.join(golden_dir, "sift_l1.npz"))
    idx, dist, good = _knn_and_matches(ctx, MatcherType.SIFT_BF_L1, _f32(g["q"]), _f32(g["t"]))
    _check_knn(idx, dist, l1[name + "_idx"], l1[name + "_dist"])
    assert np.array_equal(good, c_oracle.ratio_test(l1[name + "_idx"], l1[name + "_dist"], 0.7))


@pytest.mark.parametrize("nq,nt,seed,kind", [(1000, 1500, 111, "int"), (2049, 777, 112, "int"), (130, 4100, 113, "int"),
                                             (700, 900, 114, "float"), (255, 2049, 115, "float")])
def test_sift_l1_seeded(ctx, nq, nt, seed, kind):
    q, t = synth.sift_pair(nq, nt, seed) if kind == "int" else synth.float_pair(nq, nt, seed)
    idx, dist, good = _knn_and_matches(ctx, MatcherType.SIFT_BF_L1, q, t)
    ridx, rdist = c_oracle.l1_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.7))


def test_sift_l1_ties_mixed_batch_and_edges(ctx):
    # heavy ties: equal L1 distances must keep the lowest train index
    rng = np.random.default_rng(17)
    base = synth.sift_like(6, 18)
    q, t = base[rng.integers(0, 6, 300)], base[rng.integers(0, 6, 2600)]
    idx, dist, _ = _knn_and_matches(ctx, MatcherType.SIFT_BF_L1, q, t)
    _check_knn(idx, dist, *c_oracle.l1_knn2(q, t))
    # one batch mixing integer-valued and general-float train sets, ragged sizes, T = 0 / 1 / 2
    qi = synth.sift_like(500, 19)
    trains = [synth.sift_like(700, 20), synth.float_pair(1, 300, 21)[1], synth.sift_like(1, 22),
              np.zeros((0, 128), np.float32), synth.sift_like(2, 23), synth.sift_like(1300, 24) * 0.5]
    Q = ctx.upload(qi)
    Ts = [ctx.upload(t) for t in trains]
    res = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF_L1, 0.8)
    for t, got in zip(trains, res):
        assert np.array_equal(got, c_oracle.match_features(3, qi, t, 0.8))
    # general-float query against integer trains
    qf = synth.float_pair(300, 1, 25)[0]
    res = ctx.matchBatch(ctx.upload(qf), Ts, MatcherType.SIFT_BF_L1, 0.8)
    for t, got in zip(trains, res):
        assert np.array_equal(got, c_oracle.match_features(3, qf, t, 0.8))


def test_sift_l1_10k_properties(ctx):
    """BASELINE cfg1 size: against the oracle on a row sample + planted correspondences found."""
    q, t = synth.sift_pair(10000, 10000, 1001)
    Q, T = ctx.upload(q), ctx.upload(t)
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF_L1, Q, T)
    rows = np.random.default_rng(5).choice(10000, 300, replace=False)
    ridx, rdist = c_oracle.l1_knn2(q[rows], t)
    _check_knn(idx[rows], dist[rows], ridx, rdist)
    good = ctx.matchFeatures(Q, T, MatcherType.SIFT_BF_L1, 0.7)
    assert 2000 < len(good) <= 10000 and np.all(np.diff(good["queryIdx"]) > 0)


# ---- host-narrowed upload (slamb200_upload_desc_packed) ------------------------------------------
def test_packed_upload_equals_plain_upload(ctx):
    q, t = synth.sift_pair(3000, 4100, 901)
    ridx, rdist = c_oracle.l2_knn2(q, t)
    Q, T = ctx.upload_packed(q), ctx.upload_packed(t)
    assert Q.exact_mode == 1 and T.exact_mode == 1
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T)
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7), c_oracle.ratio_test(ridx, rdist, 0.7))
    # mixed with a plainly uploaded set, and in the L1 mode (which reads the u8 and f32 copies)
    assert np.array_equal(ctx.matchFeatures(ctx.upload(q), T, MatcherType.SIFT_BF, 0.7),
                          c_oracle.ratio_test(ridx, rdist, 0.7))
    assert np.array_equal(ctx.matchFeatures(Q, T, MatcherType.SIFT_BF_L1, 0.7), c_oracle.match_features(3, q, t, 0.7))
    # row pitch (cv::Mat::step) and the caller's buffer being reusable right after the call
    wide = np.zeros((4100, 160), np.float32)
    wide[:, :128] = t
    T2 = ctx.upload_packed(wide[:, :128])
    wide[:] = 7.0
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T2)
    _check_knn(idx, dist, ridx, rdist)


def test_packed_upload_falls_back_for_other_data(ctx):
    # general floats: not packable, same results as the plain path (bit-exact distances)
    q, t = synth.float_pair(300, 500, 902)
    Q, T = ctx.upload_packed(q), ctx.upload_packed(t)
    assert Q.exact_mode == 0
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T)
    _check_knn(idx, dist, *c_oracle.l2_knn2(q, t))
    # a single non-integer element; integers whose norm breaks the exact-mode bound
    q, t = synth.sift_pair(400, 600, 903)
    t[77, 5] += 0.25
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, ctx.upload_packed(q), ctx.upload_packed(t))
    _check_knn(idx, dist, *c_oracle.l2_knn2(q, t))
    big = np.full((50, 128), 200.0, np.float32)
    big[:, ::3] = np.arange(50, dtype=np.float32)[:, None]
    B = ctx.upload_packed(big)
    assert B.exact_mode == 0
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, ctx.upload_packed(q), B)
    _check_knn(idx, dist, *c_oracle.l2_knn2(q, big))
    # ORB and empty sets pass through
    qo, to = synth.orb_pair(200, 300, 904)
    idx, dist = ctx.knnMatch(MatcherType.ORB_BF, ctx.upload_packed(qo), ctx.upload_packed(to))
    _check_knn(idx, dist, *c_oracle.hamming_knn2(qo, to))
    E = ctx.upload_packed(np.zeros((0, 128), np.float32))
    assert len(ctx.matchFeatures(ctx.upload_packed(q), E, MatcherType.SIFT_BF)) == 0


def test_packed_upload_from_many_threads(ctx):
    """More concurrent callers than staging buffers in flight at first: the pool grows and recycles."""
    import threading
    q = synth.sift_like(2000, 905)
    trains = [synth.sift_train_from_query(q, 2000 + 10 * i, 906 + i) for i in range(12)]
    want = [c_oracle.match_features(0, q, t, 0.7) for t in trains]
    Q = ctx.upload_packed(q)
    got = [None] * 12

    def work(i):
        for _ in range(5):
            T = ctx.upload_packed(trains[i])
            got[i] = ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7)
            T.free()
    th = [threading.Thread(target=work, args=(i,)) for i in range(12)]
    [t.start() for t in th]
    [t.join() for t in th]
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_sift_l1_train_set_beyond_packed_key_range(ctx):
    """More than 2^17 train rows: the byte-wise kernel's packed (distance, index) key no longer
    fits and the fp32 kernel takes the pair -- same results."""
    q = synth.sift_like(48, 931)
    t = synth.sift_like(140000, 932)
    t[139999] = q[7]                      # an exact copy at the very end of the range
    t[131072] = q[9]
    idx, dist, good = _knn_and_matches(ctx, MatcherType.SIFT_BF_L1, q, t)
    ridx, rdist = c_oracle.l1_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    assert idx[7, 0] == 139999 and idx[9, 0] == 131072
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.7))


def test_cfg3_full_window_properties(ctx):
    """BASELINE cfg3 at full size: one 10k-row query frame against 210 train frames in one call.
    Size-independent properties on every pair, the oracle on sampled pairs and rows."""
    q = synth.sift_like(10000, 3000)
    trains = [synth.sift_train_from_query(q, 10000, 3001 + p) for p in range(210)]
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    res = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
    assert len(res) == 210
    for p, m in enumerate(res):
        assert 2000 < len(m) <= 10000, p                       # the planted correspondences are found
        assert np.all(np.diff(m["queryIdx"]) > 0) and np.all(m["imgIdx"] == 0)
        assert m["trainIdx"].min() >= 0 and m["trainIdx"].max() < 10000
    rows = np.random.default_rng(9).choice(10000, 250, replace=False)
    for p in (0, 104, 209):
        ridx, rdist = c_oracle.l2_knn2(q[rows], trains[p])
        want = c_oracle.ratio_test(ridx, rdist, 0.7)
        got = res[p][np.isin(res[p]["queryIdx"], rows)]
        order = {int(r): k for k, r in enumerate(rows)}
        got_local = np.array(sorted((order[int(g["queryIdx"])], int(g["trainIdx"]), float(g["distance"])) for g in got))
        want_local = np.array(sorted((int(w["queryIdx"]), int(w["trainIdx"]), float(w["distance"])) for w in want))
        assert np.array_equal(got_local, want_local), p
        assert np.array_equal(res[p], ctx.matchFeatures(Q, Ts[p], MatcherType.SIFT_BF, 0.7))
    for t in Ts:
        t.free()


# ---- ORB: the two kernels (tcgen05 on e4m3 0/1 bytes, XOR/POPC) against the same oracle -------------
@pytest.fixture(params=["tcgen05", "xor_popc"])
def orb_kernel(ctx, request):
    ctx.debug_orb_kernel(tensor_cores=request.param == "tcgen05")
    yield request.param
    ctx.debug_orb_kernel(tensor_cores=True)


@pytest.mark.gpu
@pytest.mark.parametrize("nq,nt,seed", [(1500, 2500, 411), (4097, 300, 412), (100, 5000, 413), (257, 9, 414),
                                        (3, 2, 415), (700, 10007, 416)])
def test_orb_both_kernels_seeded(ctx, orb_kernel, nq, nt, seed):
    q, t = synth.orb_pair(nq, nt, seed)
    idx, dist, good = _knn_and_matches(ctx, MatcherType.ORB_BF, q, t)
    ridx, rdist = c_oracle.hamming_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.7))


@pytest.mark.gpu
def test_orb_both_kernels_ties_and_extremes(ctx, orb_kernel):
    """Masses of equal distances (lowest train index wins), all-zero / all-one rows (popcount 0 and
    256: the ends of the augmentation encoding), identical sets (distance 0)."""
    rng = np.random.default_rng(16)
    base = rng.integers(0, 256, (7, 32), dtype=np.uint8)
    base[0] = 0
    base[1] = 255
    q = base[rng.integers(0, 7, 600)]
    t = base[rng.integers(0, 7, 4100)]
    idx, dist, good = _knn_and_matches(ctx, MatcherType.ORB_BF, q, t)
    ridx, rdist = c_oracle.hamming_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.7))
    # every popcount from 0 to 256 on both sides
    rows = np.zeros((257, 32), np.uint8)
    for k in range(257):
        bits = np.zeros(256, np.uint8)
        bits[rng.permutation(256)[:k]] = 1
        rows[k] = np.packbits(bits)
    idx, dist, good = _knn_and_matches(ctx, MatcherType.ORB_BF, rows, rows[::-1].copy(), ratio=0.9)
    ridx, rdist = c_oracle.hamming_knn2(rows, rows[::-1].copy())
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.9))


@pytest.mark.gpu
def test_orb_both_kernels_batch_and_ratios(ctx, orb_kernel):
    q, _ = synth.orb_pair(900, 10, 421)
    trains = [synth.orb_pair(10, n, 422 + i)[1] for i, n in enumerate((1200, 1, 0, 333, 2048))]
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    for ratio in (0.0, 0.7, 1.0, 1.5):
        got = ctx.matchBatch(Q, Ts, MatcherType.ORB_BF, ratio)
        for g, t in zip(got, trains):
            assert np.array_equal(g, c_oracle.match_features(2, q, t, ratio))


@pytest.mark.gpu
def test_sift_equal_best_and_second_with_ratio_above_one(ctx):
    """Duplicated train rows: d0 == d1 exactly.  With a ratio above 1 such rows are kept, and the
    match must still carry the lower of the two train indices (cv::BFMatcher's strict-'<' insert)."""
    q, t = synth.sift_pair(500, 900, 731)
    t = np.concatenate([t, t[::-1]]).copy()       # every train row twice
    ridx, rdist = c_oracle.l2_knn2(q, t)
    assert np.all(rdist[:, 0] == rdist[:, 1])
    Q, T = ctx.upload(q), ctx.upload(t)
    for r in (0.7, 1.0, 1.25, 2.0):
        assert np.array_equal(ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, r),
                              c_oracle.ratio_test(ridx, rdist, r))
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T)
    _check_knn(idx, dist, ridx, rdist)


@pytest.mark.gpu
def test_general_float_two_nearest_in_one_chunk(ctx):
    """The certified rerank first reads only the best 8-column group of each of the four best
    32-column chunks.  Here the two nearest train rows of every query sit in the SAME chunk but in
    different groups, so that first stage cannot certify and the whole-chunk stage must."""
    rng = np.random.default_rng(91)
    q, t = synth.float_pair(512, 4096, 1510)
    for k in range(512):
        c = int(rng.integers(0, 4096 // 32))
        g = rng.permutation(4)[:2]
        a, b = 32 * c + 8 * int(g[0]) + int(rng.integers(0, 8)), 32 * c + 8 * int(g[1]) + int(rng.integers(0, 8))
        t[a] = q[k] + rng.normal(0, 0.5, 128).astype(np.float32)
        t[b] = q[k] + rng.normal(0, 0.9, 128).astype(np.float32)
    idx, dist, good = _knn_and_matches(ctx, MatcherType.SIFT_BF, q, t, ratio=0.8)
    ridx, rdist = c_oracle.l2_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    assert np.array_equal(good, c_oracle.ratio_test(ridx, rdist, 0.8))
    # the construction did what it says for most rows (later plants may overwrite earlier ones)
    same_chunk = (ridx[:, 0] // 32 == ridx[:, 1] // 32) & (ridx[:, 0] // 8 != ridx[:, 1] // 8)
    assert same_chunk.mean() > 0.6


@pytest.mark.gpu
def test_general_float_many_fallback_rows(ctx):
    """More uncertifiable rows than the fallback grid has blocks: the unsplit full-row scan."""
    import ctypes
    rng = np.random.default_rng(92)
    q, t = synth.float_pair(800, 30000, 1511)
    for k in range(700):
        rows = rng.choice(30000, 40, replace=False)
        t[rows] = q[k] + rng.normal(0, 2e-3, (40, 128)).astype(np.float32)
    Q, T = ctx.upload(q), ctx.upload(t)
    idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T)
    ridx, rdist = c_oracle.l2_knn2(q, t)
    _check_knn(idx, dist, ridx, rdist)
    ctx.matchBatchEnqueue(Q, [T], MatcherType.SIFT_BF, 0.95)
    lib = ctx._lib
    lib.slamb200_dbg_last_fallback_rows.argtypes = [ctypes.c_void_p]
    n_fb = lib.slamb200_dbg_last_fallback_rows(ctx._h)
    assert n_fb >= 592, n_fb
    got, _ = ctx.batchFetch()
    assert np.array_equal(got[0], c_oracle.ratio_test(ridx, rdist, 0.95))


def test_general_float_batch_in_sub_batches(ctx):
    """Ten general-float pairs in one call: the batch runs as sub-batches whose certified rerank
    overlaps the next sub-batch's tcgen05 kernel on a second stream, each rerank launch with its
    own (re-zeroed) fallback bookkeeping.  Uncertifiable rows are planted in pairs of different
    sub-batches (a few, and more than the split scan's grid), one pair is integer valued."""
    rng = np.random.default_rng(93)
    q, _ = synth.float_pair(600, 8, 1520)
    trains = []
    for p in range(10):
        t = synth.float_pair(8, 3000 + 217 * p, 1521 + p)[1]
        t[:200] = q[:200] + rng.normal(0, 1.5, (200, 128)).astype(np.float32)     # true neighbours
        if p in (1, 6, 9):                                                           # near-ties: fallback rows
            n_rows = 30 if p != 6 else 500
            for k in range(n_rows):
                rows = rng.choice(len(t), 40, replace=False)
                t[rows] = q[k] + rng.normal(0, 2e-3, (40, 128)).astype(np.float32)
        trains.append(t)
    trains[4] = synth.sift_train_from_query(synth.sift_like(600, 77), 3500, 78)      # an exact-mode train set
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    got = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.9)
    for p, (g, t) in enumerate(zip(got, trains)):
        assert np.array_equal(g, c_oracle.match_features(0, q, t, 0.9)), p
