"""Short driver for ncu captures: 32 SIFT pairs (10k x 10k each) device-resident, 3 passes;
optionally ORB / RANSAC passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType
from slam_indoor_code_b200 import camera_translation as ct
which = sys.argv[1] if len(sys.argv) > 1 else "sift"
torch.zeros(1, device="cuda")
ctx = Context(0)
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
if which == "sift":
    q = synth.sift_like(10000, 3000)
    Q = ctx.upload(q)
    Ts = [ctx.upload(synth.sift_train_from_query(q, 10000, 3001 + i)) for i in range(32)]
    for _ in range(3):
        ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize()
    print("ok", [len(m) for m in ctx.batchFetch(st)[0]][:4])
elif which == "single":
    # the per-pair drop-in call's kernels: pair table in the parameters, tcgen05 kernel, one tail kernel
    q = synth.sift_like(10000, 3000)
    Q = ctx.upload(q)
    T = ctx.upload(synth.sift_train_from_query(q, 10000, 3001))
    for _ in range(4):
        ctx.matchBatchEnqueue(Q, [T], MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize()
    print("ok", len(ctx.batchFetch(st)[0][0]))
elif which == "orb":
    q, t = synth.orb_pair(10000, 10000, 2001)
    Q, T = ctx.upload(q), ctx.upload(t)
    for _ in range(3):
        ctx.matchBatchEnqueue(Q, [T] * 8, MatcherType.ORB_BF, 0.7, st)
    torch.cuda.synchronize()
    print("ok", len(ctx.batchFetch(st)[0][0]))
elif which == "gen":
    # general-float SIFT (uniform noise): split-bf16 tcgen05 kernel + certified rerank
    q, t = synth.float_pair(10000, 10000, 1002)
    Q, T = ctx.upload(q), ctx.upload(t)
    for _ in range(3):
        ctx.matchBatchEnqueue(Q, [T] * 8, MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize()
    print("ok", len(ctx.batchFetch(st)[0][0]))
elif which == "ransac":
    p1, p2, R, tv = synth.two_view(5000, 5000)
    E = synth.pose_hypotheses(2048, R, tv, 5001)
    for _ in range(3):
        c, b, m = ct.scoreEssentialBatch(ctx, [p1] * 16, [p2] * 16, synth.SAMSUNG_HV_4K, np.stack([E] * 16), 5.0)
    print("ok", b[:4])
elif which == "all":
    # one pass of every hot kernel at its bench shape: 210-pair SIFT window, 16 ORB pairs,
    # 16 x (2048 x 5000) essential scoring, 16 x (2048 x 5000) PnP scoring
    from slam_indoor_code_b200 import pnp_ransac as pr
    q = synth.sift_like(10000, 3000)
    Q = ctx.upload(q)
    Ts = [ctx.upload(synth.sift_train_from_query(q, 10000, 3001 + i)) for i in range(210)]
    for _ in range(2):
        ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize()
    print("sift ok", sum(len(m) for m in ctx.batchFetch(st)[0]))
    qo, to = synth.orb_pair(10000, 10000, 2001)
    Qo, To = ctx.upload(qo), ctx.upload(to)
    for _ in range(2):
        ctx.matchBatchEnqueue(Qo, [To] * 16, MatcherType.ORB_BF, 0.7, st)
    torch.cuda.synchronize()
    print("orb ok", len(ctx.batchFetch(st)[0][0]))
    p1, p2, R, tv = synth.two_view(5000, 5000)
    E = synth.pose_hypotheses(2048, R, tv, 5001)
    for _ in range(2):
        c, b, m = ct.scoreEssentialBatch(ctx, [p1] * 16, [p2] * 16, synth.SAMSUNG_HV_4K, np.stack([E] * 16), 5.0)
    print("ransac ok", b[:4])
    obj, img, Rp, tp = synth.pnp_scene(5000, 7000)
    poses = synth.pnp_hypotheses(2048, Rp, tp, 7001)
    for _ in range(2):
        c, b, m = pr.scorePnPBatch(ctx, [obj] * 16, [img] * 16, synth.SAMSUNG_HV_4K, synth.REF_DIST5, np.stack([poses] * 16), 8.0)
    print("pnp ok", b[:4])
elif which == "score":
    from slam_indoor_code_b200 import pnp_ransac as pr
    p1, p2, R, tv = synth.two_view(5000, 5000)
    E = synth.pose_hypotheses(2048, R, tv, 5001)
    for _ in range(2):
        c, b, m = ct.scoreEssentialBatch(ctx, [p1] * 16, [p2] * 16, synth.SAMSUNG_HV_4K, np.stack([E] * 16), 5.0)
    print("ransac ok", b[:4])
    obj, img, Rp, tp = synth.pnp_scene(5000, 7000)
    poses = synth.pnp_hypotheses(2048, Rp, tp, 7001)
    for dist in (synth.REF_DIST5, None):
        c, b, m = pr.scorePnPBatch(ctx, [obj] * 16, [img] * 16, synth.SAMSUNG_HV_4K, dist, np.stack([poses] * 16), 8.0)
    print("pnp ok", b[:4])
elif which == "single":
    q = synth.sift_like(10000, 3000)
    Q = ctx.upload(q)
    T = ctx.upload(synth.sift_train_from_query(q, 10000, 3001))
    for _ in range(6):
        ctx.matchBatchEnqueue(Q, [T], MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize()
    print("ok", len(ctx.batchFetch(st)[0][0]))
elif which == "next":
    # the "next" rows at their bench shapes: NORM_L1 (16 pairs), ORB descriptors (4K frame, 12k kps),
    # triangulation (5000 matches)
    from slam_indoor_code_b200 import orb_descriptors as od, triangulation as tri
    qi, ti = synth.sift_pair(10000, 10000, 1001)
    Qi, Ti = ctx.upload(qi), ctx.upload(ti)
    for _ in range(2):
        ctx.matchBatchEnqueue(Qi, [Ti] * 16, MatcherType.SIFT_BF_L1, 0.7, st)
    torch.cuda.synchronize()
    print("l1 ok", len(ctx.batchFetch(st)[0][0]))
    frame = synth.textured_frame(2160, 3840, 6000, 3)
    rng = np.random.default_rng(6001)
    kps = np.stack([rng.integers(31, 3840 - 31, 12000), rng.integers(31, 2160 - 31, 12000), np.full(12000, -1.0)], 1).astype(np.float32)
    for _ in range(2):
        keep, d, _ = od.extractDescriptorORB(ctx, frame, kps)
    print("orb desc ok", d.shape)
    K4 = synth.SAMSUNG_HV_4K
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    p1, p2, R, t = synth.two_view(5000, 8000, outliers=0.0)
    for _ in range(2):
        X = tri.reconstruct(ctx, K, np.eye(3), np.zeros(3), R, t, p1, p2)
    print("triangulate ok", X.shape)
