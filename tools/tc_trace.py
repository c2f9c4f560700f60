"""Developer probe: timeline of ONE tcgen05-kernel launch (mode 8 of the MODES build: the product's
behaviour + eight %globaltimer stamps per CTA) for a single 10k x 10k pair and for a 16-pair batch."""
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
Ts = [ctx.upload(synth.sift_train_from_query(q, 10000, 3001 + i)) for i in range(16)]
lib.slamb200_dbg_tc_trace.argtypes = [ctypes.c_void_p]
names = ["entry", "setup done", "first operands landed (MMA thread)", "last MMA committed", "first accumulator ready (epilogue)",
         "epilogue loop done", "after final cluster sync"]
for NP in (1, 16):
    ctx.profile_enable(True)
    for mode in (0, 8):
        lib.slamb200_dbg_set_tc_mode(mode)
        for _ in range(5): ctx.matchBatchEnqueue(Q, Ts[:NP], MatcherType.SIFT_BF, 0.7, st)
        torch.cuda.synchronize(); ctx.profile_read()
        for _ in range(50): ctx.matchBatchEnqueue(Q, Ts[:NP], MatcherType.SIFT_BF, 0.7, st)
        torch.cuda.synchronize()
        p = ctx.profile_read()
        print(f"pairs={NP} mode {mode}:", {k: round(v[0] / max(v[1], 1) * 1e3, 2) for k, v in p.items() if v[1]}, flush=True)
    ctx.profile_enable(False)
    # one isolated launch
    torch.cuda.synchronize()
    ctx.matchBatchEnqueue(Q, Ts[:NP], MatcherType.SIFT_BF, 0.7, st)
    torch.cuda.synchronize()
    tr = np.zeros((160, 8), np.uint64)
    assert lib.slamb200_dbg_tc_trace(tr.ctypes.data) == 0
    tr = tr[:148].astype(np.int64)
    t0 = tr[:, 0].min()
    print(f"pairs={NP}: tiles issued per CTA pair min/max {tr[0::2, 7].min()} / {tr[0::2, 7].max()}")
    for k, nm in enumerate(names):
        rows = tr[0::2] if k in (2, 3) else tr
        v = (rows[:, k] - t0) / 1e3
        print(f"  {nm:45s} min {v.min():7.2f}  median {np.median(v):7.2f}  max {v.max():7.2f} us")
    mm = (tr[0::2, 3] - tr[0::2, 2]) / 1e3
    print(f"  MMA span per CTA pair: median {np.median(mm):.2f} us, per tile {np.median(mm / tr[0::2, 7]):.3f} us")
    lib.slamb200_dbg_set_tc_mode(0)
