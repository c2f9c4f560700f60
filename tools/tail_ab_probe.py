"""Developer probe: fused tail kernel against the separate merge / rerank / finalize kernels."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType
torch.zeros(1, device="cuda")
ctx = Context(0); lib = ctx._lib
lib.slamb200_dbg_set_fused_tail.argtypes = [ctypes.c_void_p, ctypes.c_int]
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
def ev_time(fn, iters, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
q = synth.sift_like(10000, 3000)
Q = ctx.upload(q)
Ts = [ctx.upload(synth.sift_train_from_query(q, 10000, 3001 + i)) for i in range(64)]
qo, to = synth.orb_pair(10000, 10000, 2001)
Qo = ctx.upload(qo); To = [ctx.upload(to) for _ in range(16)]
for rep in range(2):
    for fused in (2, 1):
        lib.slamb200_dbg_set_fused_tail(ctx._h, fused)
        r = []
        for P, it in ((1, 200), (16, 30), (64, 10)):
            ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, Ts[:P], MatcherType.SIFT_BF, 0.7, st), it)
            r.append(f"sift x{P}: {ms / P * 1e3:.2f}")
        for P, it in ((1, 200), (16, 30)):
            ms = ev_time(lambda: ctx.matchBatchEnqueue(Qo, To[:P], MatcherType.ORB_BF, 0.7, st), it)
            r.append(f"orb x{P}: {ms / P * 1e3:.2f}")
        ctx.profile_enable(True); ctx.profile_read()
        for _ in range(20): ctx.matchBatchEnqueue(Q, Ts[:64], MatcherType.SIFT_BF, 0.7, st)
        p = ctx.profile_read(); ctx.profile_enable(False)
        r.append("x64 kernels ms: " + ", ".join(f"{k} {v[0] / 20:.3f}" for k, v in p.items() if v[1]))
        print(f"fused={fused} us/pair  " + " | ".join(r), flush=True)
