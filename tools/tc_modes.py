"""Timing experiments on the tcgen05 kernel: what each role costs (results invalid for mode>0)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType
torch.zeros(1, device="cuda")
ctx = Context(0); lib = ctx._lib
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
q = synth.sift_like(10000, 3000)
Q = ctx.upload(q)
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Ts = [ctx.upload(synth.sift_train_from_query(q, 10000, 3001 + i)) for i in range(NP)]
ctx.profile_enable(True)
# 0 full | 1 no chunk processing | 2 no TMEM reads (MMA+TMA) | 3 TMA only | 4/5 five/nine kind::i8 MMAs
# (probe) | 6 epilogue without MMAs | 7 TMEM reads only
for mode in (0, 1, 2, 3, 4, 5, 6, 7, 0):
    lib.slamb200_dbg_set_tc_mode(mode)
    for _ in range(2): ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize(); ctx.profile_read()
    for _ in range(10): ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    ms, n = ctx.profile_read()["sift_tc"]
    print(f"mode {mode}: tc kernel {ms/n*1e3/NP:.2f} us/pair  {NP*25.6e9/(ms/n)/1e9:.0f} TFLOP/s", flush=True)
