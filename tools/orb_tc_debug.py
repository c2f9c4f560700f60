"""Developer probe: ORB through the tcgen05 kernel -- raw accumulators of the first tile against
Hamming / 2, matches and raw k-NN against the oracle, and the time per 10k x 10k pair of both ORB
kernels."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth, c_oracle
from slam_indoor_code_b200.feature_matching import Context, MatcherType
from slam_indoor_code_b200 import _capi

torch.zeros(1, device="cuda")
ctx = Context(0)
lib = ctx._lib
lib.slamb200_dbg_tc_tile.argtypes = [ctypes.c_void_p] * 4
lib.slamb200_dbg_tc_tile.restype = ctypes.c_int
lib.slamb200_dbg_set_tc_orb.argtypes = [ctypes.c_void_p, ctypes.c_int]
q, t = synth.orb_pair(300, 700, 5)
Q, T = ctx.upload(q), ctx.upload(t)
out = np.zeros((256, 256), np.float32)
rc = lib.slamb200_dbg_tc_tile(ctx._h, Q._h, T._h, _capi.ptr(out))
print("rc", rc, lib.slamb200_last_error(), flush=True)
qb, tb = np.unpackbits(q[:256], axis=1).astype(np.int32), np.unpackbits(t[:256], axis=1).astype(np.int32)
ref = ((qb[:, None, :] != tb[None, :, :]).sum(2) / 2).astype(np.float32)
print("accumulators: max abs diff", np.abs(out - ref).max(), "mismatched", int((out != ref).sum()), flush=True)
if (out != ref).any():
    np.set_printoptions(linewidth=200, suppress=True)
    print("out[:4,:8]\n", out[:4, :8]); print("ref[:4,:8]\n", ref[:4, :8])
    dot = -(qb @ tb.T).astype(np.float32); aug = ((qb.sum(1)[:, None] + tb.sum(1)[None, :]) / 2).astype(np.float32)
    print("matches -dot only:", (out == dot).mean(), " matches aug only:", (out == aug).mean())
for nq, nt, seed in [(300, 700, 5), (1024, 1536, 8), (10000, 10000, 2001), (77, 3, 9), (5, 1, 10)]:
    q, t = synth.orb_pair(nq, nt, seed)
    Q, T = ctx.upload(q), ctx.upload(t)
    got = ctx.matchFeatures(Q, T, MatcherType.ORB_BF, 0.7)
    ref = c_oracle.match_features(2, q, t, 0.7)
    idx, dist = ctx.knnMatch(MatcherType.ORB_BF, Q, T)
    ridx, rdist = c_oracle.hamming_knn2(q, t)
    print(f"{nq}x{nt}: matches equal {np.array_equal(got, ref)} ({len(got)} vs {len(ref)}), knn idx equal "
          f"{np.array_equal(idx, ridx)}, dist equal {np.array_equal(dist, rdist)}", flush=True)
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
q, t = synth.orb_pair(10000, 10000, 2001)
Q = ctx.upload(q); Ts = [ctx.upload(t) for _ in range(16)]
for on in (1, 0):
    lib.slamb200_dbg_set_tc_orb(ctx._h, on)
    for P in (1, 16):
        for _ in range(3): ctx.matchBatchEnqueue(Q, Ts[:P], MatcherType.ORB_BF, 0.7, st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): ctx.matchBatchEnqueue(Q, Ts[:P], MatcherType.ORB_BF, 0.7, st)
        b.record(); torch.cuda.synchronize()
        print(f"tc={on} batch of {P}: {a.elapsed_time(b) / 10 / P * 1e3:.1f} us/pair", flush=True)
