"""Step time of the 210-pair window vs the sub-batch pipelining knob."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType
torch.zeros(1, device="cuda")
ctx = Context(0); lib = ctx._lib
lib.slamb200_dbg_set_sub_batch.argtypes = [ctypes.c_void_p, ctypes.c_int]
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 210
q = synth.sift_like(10000, 3000)
Q = ctx.upload(q)
Ts = [ctx.upload(synth.sift_train_from_query(q, 10000, 3001 + i)) for i in range(NP)]
for sb in (0, 105, 53, 30, 15, 0):
    lib.slamb200_dbg_set_sub_batch(ctx._h, sb)
    for _ in range(3): ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"sub_batch={sb:4d}: {ms:.3f} ms/step  {NP/ms*1e3:.0f} pairs/s", flush=True)
