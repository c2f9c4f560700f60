"""General-float path: kernel times and fallback fraction on several data distributions."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType
torch.zeros(1, device="cuda")
ctx = Context(0); lib = ctx._lib
lib.slamb200_dbg_last_fallback_rows.argtypes = [ctypes.c_void_p]
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
def run(name, q, t, P=4):
    Q = ctx.upload(q); Ts = [ctx.upload(t) for _ in range(P)]
    ctx.profile_enable(True)
    for _ in range(2): ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize(); ctx.profile_read()
    for _ in range(5): ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    pr = ctx.profile_read()
    fb = lib.slamb200_dbg_last_fallback_rows(ctx._h)
    print(f"{name}: tc_gen {pr['sift_tc_gen'][0]/5/P*1e3:.1f} us/pair, gen_rerank+fallback {pr['sift_gen_rerank'][0]/5/P*1e3:.1f} us/pair, "
          f"fallback rows {fb} of {P*len(q)} ({100.0*fb/(P*len(q)):.2f} %)", flush=True)
rng = np.random.default_rng(1)
q, t = synth.float_pair(10000, 10000, 1002); run("uniform floats [0,255)", q, t)
qi, ti = synth.sift_pair(10000, 10000, 1001)
run("SIFT + 0.25 jitter", qi + rng.random(qi.shape, np.float32)*0.25, ti + rng.random(ti.shape, np.float32)*0.25)
def rootsift(x):
    x = x / np.maximum(x.sum(1, keepdims=True), 1e-9); return np.sqrt(x).astype(np.float32)
run("RootSIFT (unit norm)", rootsift(qi), rootsift(ti))
run("SIFT L2-normalised x 512 (floats)", (qi / np.linalg.norm(qi, axis=1, keepdims=True) * 512).astype(np.float32), (ti / np.linalg.norm(ti, axis=1, keepdims=True) * 512).astype(np.float32))
