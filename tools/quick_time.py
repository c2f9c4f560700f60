"""Developer timing probe (not the bench): CUDA-event times of the device-resident paths."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType
from slam_indoor_code_b200 import camera_translation as ct

def ev_time(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

torch.cuda.init(); torch.zeros(1, device="cuda")
ctx = Context(0)
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
which = sys.argv[1:] or ["orb", "sift", "siftf", "ransac", "batch"]
if "orb" in which:
    q, t = synth.orb_pair(10000, 10000, 2001)
    Q, T = ctx.upload(q), ctx.upload(t)
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, [T], MatcherType.ORB_BF, 0.7, st))
    print(f"ORB 10k x 10k: {ms*1e3:.1f} us/pair  {8e8/ms/1e9:.2f} Tpopc/s", flush=True)
if "sift" in which:
    q, t = synth.sift_pair(10000, 10000, 1001)
    Q, T = ctx.upload(q), ctx.upload(t)
    print("exact_mode", Q.exact_mode, T.exact_mode)
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, [T], MatcherType.SIFT_BF, 0.7, st), iters=10)
    print(f"SIFT int 10k x 10k: {ms*1e3:.1f} us/pair  {25.6e9/ms/1e9:.1f} TFLOP/s", flush=True)
if "siftf" in which:
    q, t = synth.float_pair(10000, 10000, 1002)
    Q, T = ctx.upload(q), ctx.upload(t)
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, [T], MatcherType.SIFT_BF, 0.7, st), iters=5)
    print(f"SIFT float 10k x 10k: {ms*1e3:.1f} us/pair  {25.6e9/ms/1e9:.1f} TFLOP/s", flush=True)
if "batch" in which:
    q = synth.sift_like(10000, 3000)
    Q = ctx.upload(q)
    Ts = [ctx.upload(synth.sift_train_from_query(q, 10000, 3001 + i)) for i in range(16)]
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st), iters=5)
    print(f"SIFT int batch16: {ms/16*1e3:.1f} us/pair  {16*25.6e9/ms/1e9:.1f} TFLOP/s", flush=True)
if "ransac" in which:
    p1, p2, R, tv = synth.two_view(5000, 5000)
    E = synth.pose_hypotheses(2048, R, tv, 5001)
    t0 = time.perf_counter()
    for _ in range(10): ct.scoreEssentialHypotheses(ctx, p1, p2, synth.SAMSUNG_HV_4K, E, 5.0)
    dt = (time.perf_counter() - t0) / 10
    print(f"RANSAC 2048x5000 host-call: {dt*1e6:.1f} us  ({2048*5000*40/dt/1e12:.2f} TFLOP/s fp64-equiv incl. copies)")
    P = 32
    p1s = [p1] * P; p2s = [p2] * P; Es = np.stack([E] * P)
    t0 = time.perf_counter()
    for _ in range(3): ct.scoreEssentialBatch(ctx, p1s, p2s, synth.SAMSUNG_HV_4K, Es, 5.0)
    dt = (time.perf_counter() - t0) / 3
    print(f"RANSAC batch {P}: {dt/P*1e6:.1f} us/pair  ({P*2048*5000*40/dt/1e12:.2f} TFLOP/s fp64-equiv incl. copies)")
