"""Developer probe: general-float batches with the certified rerank of sub-batch k beside the tcgen05
kernel of sub-batch k+1 (SLAMB200_GEN_SUB pairs per sub-batch; 0 = one launch for the whole batch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType
torch.zeros(1, device="cuda")
ctx = Context(0)
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
qi, ti = synth.sift_pair(10000, 10000, 1001)
def rootsift(x):
    x = x / np.maximum(x.sum(1, keepdims=True), 1e-9); return np.sqrt(x).astype(np.float32)
for name, (q, t) in (("uniform", synth.float_pair(10000, 10000, 1002)), ("rootsift", (rootsift(qi), rootsift(ti)))):
    Q = ctx.upload(q)
    for P in (4, 16, 32):
        Ts = [ctx.upload(t) for _ in range(P)]
        for _ in range(2): ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
        b.record(); torch.cuda.synchronize()
        got = ctx.batchFetch(st)[0]
        print(f"GEN_SUB={os.environ.get('SLAMB200_GEN_SUB', 'default')} {name} x{P}: {a.elapsed_time(b) / 5 / P * 1e3:.1f} us/pair, matches of pair 0: {len(got[0])}", flush=True)
        for T in Ts: T.free()
