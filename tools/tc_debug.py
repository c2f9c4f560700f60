"""Developer probe: dump the raw tcgen05 accumulators of the first tile and compare with NumPy."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth_inputs as synth, c_oracle
from slam_indoor_code_b200.feature_matching import Context, MatcherType
from slam_indoor_code_b200 import _capi

ctx = Context(0)
lib = ctx._lib
lib.slamb200_dbg_tc_tile.argtypes = [ctypes.c_void_p] * 4
lib.slamb200_dbg_tc_tile.restype = ctypes.c_int
q, t = synth.sift_pair(300, 700, 5)
Q, T = ctx.upload(q), ctx.upload(t)
print("exact", Q.exact_mode, T.exact_mode, flush=True)
out = np.zeros((256, 256), np.float32)
rc = lib.slamb200_dbg_tc_tile(ctx._h, Q._h, T._h, _capi.ptr(out))
print("rc", rc, lib.slamb200_last_error(), flush=True)
qq, tt = q[:256].astype(np.float64), t[:256].astype(np.float64)
ref = ((qq ** 2).sum(1)[:, None] + (tt ** 2).sum(1)[None, :]) / 2 - qq @ tt.T
print("max abs diff", np.abs(out - ref).max(), "n mismatched", (out != ref).sum())
if (out != ref).any():
    np.set_printoptions(linewidth=200, suppress=True)
    print("out[:4,:8]\n", out[:4, :8]); print("ref[:4,:8]\n", ref[:4, :8])
    bad = np.argwhere(out != ref)
    print("first bad", bad[:10].tolist())
    # try to explain: only dot part / only aug part
    dot = -(qq @ tt.T); aug = ((qq ** 2).sum(1)[:, None] + (tt ** 2).sum(1)[None, :]) / 2
    print("matches -dot only:", (out == dot).mean(), " matches aug only:", (out == aug).mean())
idx, dist = ctx.knnMatch(MatcherType.SIFT_BF, Q, T)
ridx, rdist = c_oracle.l2_knn2(q, t)
print("knn idx equal", np.array_equal(idx, ridx), "dist equal", np.array_equal(dist.view(np.int32), rdist.view(np.int32)))
