"""GPU: ordering and validation contracts of the C ABI (include/slamb200.h) that the call sites of
the reference rely on implicitly (one default stream, OpenCV Mats with arbitrary pitch)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import synth_inputs as synth
from oracle import c_oracle
from slam_indoor_code_b200 import _capi
from slam_indoor_code_b200 import camera_translation as ct
from slam_indoor_code_b200._capi import Slamb200Error
from slam_indoor_code_b200.feature_matching import MatcherType

K4 = synth.SAMSUNG_HV_4K


def test_device_upload_is_ordered_behind_the_default_stream(ctx):
    """slamb200_upload_desc_device with stream = NULL (torch's default stream): the rows are still
    being produced on the legacy default stream when the call is made; the prep kernels (on a
    non-blocking lane) must wait for them.  cfg4-sized frame so that the producer is slow."""
    import torch
    q = synth.sift_like(50000, 4400)
    t = synth.sift_like(50000, 4401)
    t[:3000] = np.clip(q[:3000] + 2, 0, 255)
    want = ctx.matchFeatures(ctx.upload(q), ctx.upload(t), MatcherType.SIFT_BF, 0.7)
    hq = torch.from_numpy(q).pin_memory()
    ht = torch.from_numpy(t).pin_memory()
    for _ in range(3):
        dq = torch.zeros((50000, 128), dtype=torch.float32, device="cuda")
        dt = torch.zeros((50000, 128), dtype=torch.float32, device="cuda")
        big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        # keep the default stream busy, then produce the rows behind that work
        for _ in range(8):
            big.fill_(1)
        dq.copy_(hq, non_blocking=True)
        dt.copy_(ht, non_blocking=True)
        assert torch.cuda.current_stream().cuda_stream == 0
        Q = ctx.upload_device(dq.data_ptr(), 50000, _capi.DESC_F32X128, 512, stream=0)
        T = ctx.upload_device(dt.data_ptr(), 50000, _capi.DESC_F32X128, 512, stream=0)
        got = ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7)
        assert np.array_equal(got, want)
        Q.free(); T.free()
        torch.cuda.synchronize()


def test_batch_enqueue_and_fetch_on_different_streams(ctx):
    """The device-resident batch keeps one result set: fetch / score calls on ANOTHER stream than
    the enqueue are ordered behind it on the device, and a second enqueue on a third stream does
    not overwrite a batch that is still running."""
    import torch
    q = synth.sift_like(10000, 3000)
    trains = [synth.sift_train_from_query(q, 10000, 3001 + p) for p in range(24)]
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    want = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
    s = [torch.cuda.Stream() for _ in range(3)]
    for it in range(4):
        ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, s[0].cuda_stream)
        got, _ = ctx.batchFetch(s[1].cuda_stream)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        # back-to-back enqueues on different streams, results of the LAST one
        ctx.matchBatchEnqueue(Q, Ts[:12], MatcherType.SIFT_BF, 0.7, s[0].cuda_stream)
        ctx.matchBatchEnqueue(Q, Ts[12:], MatcherType.SIFT_BF, 0.7, s[2].cuda_stream)
        got, _ = ctx.batchFetch(s[1].cuda_stream)
        assert len(got) == 12
        for g, w in zip(got, want[12:]):
            assert np.array_equal(g, w)
    # the scoring chain on yet another stream
    kq, kts, poses = synth.window_geometry(10000, 10000, [3001 + p for p in range(24)], 77)
    KQ = ctx.upload_keypoints(kq)
    KTs = [ctx.upload_keypoints(k) for k in kts]
    E = np.stack([synth.pose_hypotheses(64, R, t, 90 + p) for p, (R, t) in enumerate(poses)])
    ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, s[0].cuda_stream)
    ct.scoreBatchEnqueue(ctx, KQ, KTs, K4, E, 5.0, s[1].cuda_stream)
    counts, best, mask = ct.batchScoresFetch(ctx, s[2].cuda_stream)
    for p in (0, 11, 23):
        p1, p2 = c_oracle.gather_points(kq, kts[p], want[p])
        rc, rb, rm, _ = c_oracle.score_essential(p1, p2, K4, E[p], 5.0)
        assert np.array_equal(counts[p], rc) and best[p] == rb and np.array_equal(mask[p, : len(rm)], rm)
    # a train keypoint set shorter than its descriptor set is refused (the gather would read past it)
    short = ctx.upload_keypoints(kts[3][:9000])
    ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7)
    with pytest.raises(Slamb200Error):
        ct.scoreBatchEnqueue(ctx, KQ, KTs[:3] + [short] + KTs[4:], K4, E, 5.0)
    ctx.synchronize()


@pytest.mark.parametrize("how", ["plain", "packed", "pinned"])
@pytest.mark.parametrize("kind", ["int", "float"])
def test_mat_pitch_and_alignment_do_not_matter(ctx, how, kind):
    """A cv::Mat ROI: pitch that is only a multiple of 4 bytes and a base pointer that is not
    16-byte aligned.  Every upload entry point accepts it whatever the VALUES are (integer rows
    take the narrowing path, general floats the fp32 path)."""
    import torch
    q, t = (synth.sift_pair if kind == "int" else synth.float_pair)(700, 1100, 880)
    want = c_oracle.match_features(0, q, t, 0.8)
    for pitch, off in ((129, 1), (131, 3), (132, 0), (160, 5)):
        def place(a):
            n = a.shape[0]
            if how == "pinned":
                buf = torch.zeros(n * pitch + 8, dtype=torch.float32).pin_memory().numpy()
            else:
                buf = np.zeros(n * pitch + 8, np.float32)
            v = buf[off: off + n * pitch].reshape(n, pitch)[:, :128]
            v[...] = a
            return v
        vq, vt = place(q), place(t)
        up = {"plain": ctx.upload, "packed": ctx.upload_packed, "pinned": ctx.upload_pinned}[how]
        Q, T = up(vq), up(vt)
        got = ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.8)
        ctx.synchronize()
        assert np.array_equal(got, want), (pitch, off)
        Q.free(); T.free()


def test_launch_counter_counts_concurrent_lanes(ctx):
    import threading
    q, t = synth.sift_pair(600, 900, 881)
    Q, T = ctx.upload(q), ctx.upload(t)
    ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7)
    n0 = ctx.launch_count()
    ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7)
    per_call = ctx.launch_count() - n0
    assert per_call > 0
    n0 = ctx.launch_count()
    th = [threading.Thread(target=lambda: [ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7) for _ in range(50)])
          for _ in range(6)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert ctx.launch_count() - n0 == 300 * per_call


def test_match_batch_host_equals_resident_batch(ctx):
    """slamb200_match_batch_host: the whole window from pageable host Mats in one call (uploads and
    matching pipelined inside the library) == upload everything + slamb200_match_batch; ragged
    window, a pitched Mat, an empty Mat, general floats inside the window, more pairs than a chunk."""
    q = synth.sift_like(1500, 9300)
    sizes = [1400, 1, 0, 2600] + [900 + 13 * i for i in range(37)]
    trains = [synth.sift_train_from_query(q, max(s, 1), 9301 + i)[:s] for i, s in enumerate(sizes)]
    wide = np.zeros((len(trains[4]), 160), np.float32)
    wide[:, :128] = trains[4]
    trains[4] = wide[:, :128]                                  # a cv::Mat ROI: pitch 640 bytes
    trains[7] = (trains[7] * np.float32(0.37)).astype(np.float32)   # not integer valued
    got = ctx.matchBatchHost(q, trains, MatcherType.SIFT_BF, 0.7)
    Q = ctx.upload(q)
    Ts = [ctx.upload(np.ascontiguousarray(t)) for t in trains]
    want = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
    assert len(got) == len(want) == len(trains)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert np.array_equal(got[0], c_oracle.match_features(0, q, np.ascontiguousarray(trains[0]), 0.7))
    assert np.array_equal(got[7], c_oracle.match_features(0, q, trains[7], 0.7))
    # ORB, and an empty window
    qo, _ = synth.orb_pair(1200, 10, 9400)
    to = [synth.orb_pair(10, 700 + 50 * i, 9401 + i)[1] for i in range(5)]
    for g, t in zip(ctx.matchBatchHost(qo, to, MatcherType.ORB_BF, 0.7), to):
        assert np.array_equal(g, c_oracle.match_features(2, qo, t, 0.7))
    assert ctx.matchBatchHost(q, [], MatcherType.SIFT_BF, 0.7) == []
    # a Mat of the wrong type is refused with the library's error, nothing leaks or hangs
    with pytest.raises((Slamb200Error, ValueError)):
        ctx.matchBatchHost(q, [trains[0], qo], MatcherType.SIFT_BF, 0.7)


def test_match_batch_host_sliced_narrowing_edges(ctx):
    """The narrowing stage works on 1k-row slices of a Mat from several threads: a value that is
    not an integer in a LATE slice sends the whole Mat down the fp32 path (its other slices were
    already narrowed), and two calls at once -- plus a plain upload from a third thread while the
    pool is taken -- give the single-threaded results."""
    q = synth.sift_like(2100, 9500)
    trains = [synth.sift_train_from_query(q, 5000 + 300 * i, 9501 + i) for i in range(9)]
    trains[3][4700, 17] += np.float32(0.5)      # slice 4 of 5
    trains[6][0, 0] = np.float32(300.0)         # out of the byte range, slice 0
    want = [c_oracle.match_features(0, q, t, 0.7) for t in trains]
    res = {}

    def call(k):
        res[k] = ctx.matchBatchHost(q, trains, MatcherType.SIFT_BF, 0.7)

    def upload(k):
        T = ctx.upload(trains[0])
        Q = ctx.upload(q)
        res[k] = ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7)
        Q.free(); T.free()

    th = [threading.Thread(target=call, args=(0,)), threading.Thread(target=call, args=(1,)),
          threading.Thread(target=upload, args=(2,))]
    [x.start() for x in th]
    [x.join() for x in th]
    for k in (0, 1):
        assert len(res[k]) == len(want)
        for g, w in zip(res[k], want):
            assert np.array_equal(g, w)
    assert np.array_equal(res[2], want[0])
