"""Developer probe: device-side floor of a single-pair call -- the enqueue sequence of
slamb200_match_batch_enqueue captured once into a CUDA graph (same Q, T) and replayed, against the
plain enqueue loop (host-bound at ~45 us per call)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType

torch.zeros(1, device="cuda")
ctx = Context(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
for name, (q, t), mt in (("sift", synth.sift_pair(10000, 10000, 1001), MatcherType.SIFT_BF),
                         ("orb", synth.orb_pair(10000, 10000, 2001), MatcherType.ORB_BF)):
    Q, T = ctx.upload(q), ctx.upload(t)
    for _ in range(10):
        ctx.matchBatchEnqueue(Q, [T], mt, 0.7, st)
    torch.cuda.synchronize()
    want = ctx.batchFetch(st)[0][0].copy()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        ctx.matchBatchEnqueue(Q, [T], mt, 0.7, st)
    b.record(); torch.cuda.synchronize()
    print(f"{name}: plain enqueue loop {a.elapsed_time(b) / 200 * 1e3:.1f} us/call", flush=True)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=ts, capture_error_mode="relaxed"):
            ctx.matchBatchEnqueue(Q, [T], mt, 0.7, st)
        for _ in range(10):
            g.replay()
        torch.cuda.synchronize()
        a.record()
        for _ in range(200):
            g.replay()
        b.record(); torch.cuda.synchronize()
        got = ctx.batchFetch(st)[0][0]
        print(f"{name}: graph replay {a.elapsed_time(b) / 200 * 1e3:.1f} us/call, results equal: {np.array_equal(got, want)}", flush=True)
    except Exception as e:
        print(f"{name}: capture failed: {e!r}", flush=True)
