"""e2e variants: single batch vs chunked vs threaded."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
import bench
from slam_indoor_code_b200.feature_matching import Context, MatcherType
from slam_indoor_code_b200._capi import DMATCH
torch.zeros(1, device="cuda")
NP = 210
q, trains = bench.make_inputs(list(range(NP)), pinned=True)
ctx = Context(0)
out_buf = np.empty((NP, 10000), DMATCH); n_buf = np.zeros(NP, np.int32)
out_buf[:] = 0

def chunk_fn(Qe, ids, log=None):
    t0 = time.perf_counter()
    Te = [ctx.upload_pinned(trains[i]) for i in ids]
    t1 = time.perf_counter()
    r = ctx.matchBatch(Qe, Te, MatcherType.SIFT_BF, 0.7, out=out_buf[ids[0]:ids[-1]+1], n_out=n_buf[ids[0]:ids[-1]+1])
    t2 = time.perf_counter()
    for t in Te: t.free()
    t3 = time.perf_counter()
    if log is not None: log.append((ids[0], t0, t1, t2, t3))
    return r

def run(workers, chunk, verbose=False):
    chunks = [list(range(i, min(i + chunk, NP))) for i in range(0, NP, chunk)]
    pool = ThreadPoolExecutor(workers) if workers > 1 else None
    best = 1e9
    for it in range(4):
        log = []
        ctx.synchronize(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        Qe = ctx.upload_pinned(q)
        if pool: list(pool.map(lambda ids: chunk_fn(Qe, ids, log), chunks))
        else:
            for ids in chunks: chunk_fn(Qe, ids, log)
        Qe.free()
        ctx.synchronize()
        dt = time.perf_counter() - t0
        best = min(best, dt)
    print(f"workers={workers} chunk={chunk}: best {best*1e3:.1f} ms/step -> {NP/best:.0f} pairs/s", flush=True)
    if verbose:
        for (i, a, b, c, d) in sorted(log)[:14]:
            print(f"   chunk@{i}: start {1e3*(a-t0):6.1f} upload {1e3*(b-a):5.1f} match {1e3*(c-b):5.1f} free {1e3*(d-c):4.1f}")

run(1, 210, True); run(1, 18, True); run(4, 18, True)
