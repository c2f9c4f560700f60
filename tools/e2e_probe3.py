"""Where does the e2e step lose PCIe time?  upload-only vs upload+match, zero-copy vs staged."""
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
GB = NP * 10000 * 512 / 1e9

def timeit(fn, n=4):
    best = 1e9
    for _ in range(n):
        ctx.synchronize(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); ctx.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best

def upload_only(workers, pinned=True):
    up = ctx.upload_pinned if pinned else ctx.upload
    def one(ids):
        Te = [up(trains[i]) for i in ids]
        for t in Te: t.free()
    chunks = [list(range(i, min(i + 18, NP))) for i in range(0, NP, 18)]
    if workers == 1:
        return lambda: [one(c) for c in chunks]
    pool = ThreadPoolExecutor(workers)
    return lambda: list(pool.map(one, chunks))

for w in (1, 2, 4, 8):
    dt = timeit(upload_only(w, True))
    print(f"upload-only zero-copy workers={w}: {dt*1e3:.1f} ms  {GB/dt:.1f} GB/s", flush=True)
for w in (1, 4):
    dt = timeit(upload_only(w, False))
    print(f"upload-only staged copy workers={w}: {dt*1e3:.1f} ms  {GB/dt:.1f} GB/s", flush=True)

def full(workers, chunk):
    chunks = [list(range(i, min(i + chunk, NP))) for i in range(0, NP, chunk)]
    pool = ThreadPoolExecutor(workers)
    def one(Qe, ids):
        Te = [ctx.upload_pinned(trains[i]) for i in ids]
        r = ctx.matchBatch(Qe, Te, MatcherType.SIFT_BF, 0.7, out=out_buf[ids[0]:ids[-1]+1], n_out=n_buf[ids[0]:ids[-1]+1])
        for t in Te: t.free()
        return r
    def step():
        Qe = ctx.upload_pinned(q)
        list(pool.map(lambda ids: one(Qe, ids), chunks))
        Qe.free()
    return step
for w, c in ((4, 18), (8, 9), (8, 14), (6, 18), (8, 27)):
    dt = timeit(full(w, c))
    print(f"full workers={w} chunk={c}: {dt*1e3:.1f} ms  {NP/dt:.0f} pairs/s  {GB/dt:.1f} GB/s", flush=True)

# raw link rate: one big pinned buffer -> device with the copy engine
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dst = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2): dst.copy_(big, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4): dst.copy_(big, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"copy engine 256 MiB H2D: {4 * (256 << 20) / dt / 1e9:.1f} GB/s")
