"""Timeline of the packed e2e step: per chunk, when its uploads were issued, when its match returned."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
import bench
from slam_indoor_code_b200.feature_matching import Context, MatcherType
from slam_indoor_code_b200._capi import DMATCH
torch.zeros(1, device="cuda")
NP = 210
W, C, PT = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
q, trains = bench.make_inputs(list(range(NP)), pinned=True)
ctx = Context(0)
ctx.set_pack_threads(PT)
out_buf = np.empty((NP, 10000), DMATCH); n_buf = np.zeros(NP, np.int32); out_buf[:] = 0
chunks = [list(range(i, min(i + C, NP))) for i in range(0, NP, C)]
pool = ThreadPoolExecutor(W)
log = []
def worker(Qe, mine, t0):
    nxt = [ctx.upload_packed(trains[i]) for i in mine[0]] if mine else None
    for k, ids in enumerate(mine):
        Te = nxt
        a = time.perf_counter()
        nxt = [ctx.upload_packed(trains[i]) for i in mine[k + 1]] if k + 1 < len(mine) else None
        b = time.perf_counter()
        ctx.matchBatch(Qe, Te, MatcherType.SIFT_BF, 0.7, out=out_buf[ids[0]:ids[-1]+1], n_out=n_buf[ids[0]:ids[-1]+1])
        c = time.perf_counter()
        for t in Te: t.free()
        d = time.perf_counter()
        log.append((ids[0], a - t0, b - a, c - b, d - c))
def step():
    log.clear()
    t0 = time.perf_counter()
    Qe = ctx.upload_packed(q)
    list(pool.map(lambda w: worker(Qe, chunks[w::W], t0), range(W)))
    Qe.free()
    ctx.synchronize()
    return time.perf_counter() - t0
for _ in range(3): step()
best = min(step() for _ in range(5))
dt = step()
print(f"W={W} C={C} PT={PT}: best {best*1e3:.1f} ms, last {dt*1e3:.1f} ms")
for (i, a, up, m, f) in sorted(log, key=lambda x: x[1]):
    print(f"  chunk@{i:3d}: t={a*1e3:6.2f} upload-next {up*1e3:5.2f} match {m*1e3:5.2f} free {f*1e3:4.2f}")
# upload-only
def up_only():
    t0 = time.perf_counter()
    def one(ids):
        Te = [ctx.upload_packed(trains[i]) for i in ids]
        for t in Te: t.free()
    list(pool.map(one, chunks))
    ctx.synchronize()
    return time.perf_counter() - t0
for _ in range(2): up_only()
print(f"upload-only (packed): {min(up_only() for _ in range(4))*1e3:.1f} ms per 210 frames")
