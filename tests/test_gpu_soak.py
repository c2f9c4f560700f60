"""GPU: a few seconds of mixed traffic from many host threads through one context -- packed, pinned
and plain uploads, pair / batch / window matches in all matcher modes, scoring, ORB descriptors,
frees in every order -- every result checked against the oracle.  Exercises the lane classes, the
staging pool, the batched frees and the slab cache under contention."""
import os
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, synth
from slam_indoor_code_b200 import camera_translation as ct, orb_descriptors as od, pnp_ransac as pr
from slam_indoor_code_b200.feature_matching import MatcherType

SECONDS = float(os.environ.get("SLAMB200_SOAK_SECONDS", "6"))


def test_mixed_traffic_from_many_threads(ctx):
    rng0 = np.random.default_rng(77)
    sizes = [257, 512, 700, 1025]
    qs = {n: synth.sift_like(n, 8000 + n) for n in sizes}
    ts = {n: synth.sift_train_from_query(qs[700][: min(n, 700)], n, 8100 + n) for n in sizes}
    fq, ft = synth.float_pair(300, 400, 8200)
    oq, ot = synth.orb_pair(600, 800, 8300)
    want = {}
    for nq in sizes:
        for nt in sizes:
            want[(0, nq, nt)] = c_oracle.match_features(0, qs[nq], ts[nt], 0.7)
            want[(3, nq, nt)] = c_oracle.match_features(3, qs[nq], ts[nt], 0.7)
    want_f = c_oracle.match_features(0, fq, ft, 0.7)
    want_o = c_oracle.match_features(2, oq, ot, 0.7)
    p1, p2, R, t = synth.two_view(500, 8400)
    E = synth.pose_hypotheses(64, R, t, 8401)
    want_e = c_oracle.score_essential(p1, p2, synth.SAMSUNG_HV_4K, E, 5.0)
    obj, img, Rp, tp = synth.pnp_scene(400, 8500)
    poses = synth.pnp_hypotheses(32, Rp, tp, 8501)
    want_p = c_oracle.score_pnp(obj, img, synth.SAMSUNG_HV_4K, synth.REF_DIST5, poses, 8.0)
    frame = synth.textured_frame(160, 200, 8600, 3)
    kps = np.stack([rng0.integers(0, 200, 300), rng0.integers(0, 160, 300), np.full(300, -1.0)], 1).astype(np.float32)
    want_d = c_oracle.orb_compute(frame, kps)
    pinned = {}
    import torch
    for n in sizes:
        tt = torch.empty((n, 128), dtype=torch.float32).pin_memory()
        tt.numpy()[...] = ts[n]
        pinned[n] = tt
    errors, counts = [], [0] * 8
    stop = time.perf_counter() + SECONDS

    def worker(w):
        rng = np.random.default_rng(900 + w)
        held = []
        try:
            while time.perf_counter() < stop and not errors:
                op = int(rng.integers(0, 8))
                nq, nt = (int(x) for x in rng.choice(sizes, 2))
                up = [ctx.upload, ctx.upload_packed][int(rng.integers(0, 2))]
                if op <= 2:      # pair, L2 or L1, any upload path
                    m = 0 if op < 2 else 3
                    Q = up(qs[nq])
                    T = ctx.upload_pinned(pinned[nt].numpy()) if op == 1 else up(ts[nt])
                    got = ctx.matchFeatures(Q, T, MatcherType(m), 0.7)
                    assert np.array_equal(got, want[(m, nq, nt)]), ("pair", m, nq, nt)
                    held += [Q, T]
                elif op == 3:    # batch with ragged trains
                    Q = up(qs[nq])
                    Ts = [up(ts[n]) for n in sizes]
                    res = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
                    for n, r in zip(sizes, res):
                        assert np.array_equal(r, want[(0, nq, n)]), ("batch", nq, n)
                    held += [Q] + Ts
                elif op == 4:    # general floats and ORB
                    assert np.array_equal(ctx.matchFeatures(up(fq), up(ft), MatcherType.SIFT_BF, 0.7), want_f)
                    assert np.array_equal(ctx.matchFeatures(up(oq), up(ot), MatcherType.ORB_BF, 0.7), want_o)
                elif op == 5:    # scoring
                    c, b, mk, _ = ct.scoreEssentialHypotheses(ctx, p1, p2, synth.SAMSUNG_HV_4K, E, 5.0)
                    assert np.array_equal(c, want_e[0]) and b == want_e[1] and np.array_equal(mk, want_e[2])
                    c, b, mk, _ = pr.scorePnPHypotheses(ctx, obj, img, synth.SAMSUNG_HV_4K, synth.REF_DIST5, poses, 8.0)
                    assert np.array_equal(c, want_p[0]) and b == want_p[1] and np.array_equal(mk, want_p[2])
                elif op == 6:    # ORB descriptors, resident set matched against itself
                    keep, d, res = od.extractDescriptorORB(ctx, frame, kps, want_resident=True)
                    assert np.array_equal(keep, want_d[0].astype(bool)) and np.array_equal(d, want_d[1])
                    held.append(res)
                else:            # free what was kept around, oldest or newest first
                    order = held if rng.integers(0, 2) else held[::-1]
                    for h in order:
                        h.free()
                    held = []
                counts[w] += 1
        except Exception as e:  # noqa: BLE001
            errors.append((w, repr(e)))
        finally:
            for h in held:
                h.free()

    th = [threading.Thread(target=worker, args=(w,)) for w in range(8)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors[:3]
    assert sum(counts) > 100, counts
