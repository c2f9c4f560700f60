"""Does the zero-copy upload overlap with the tcgen05 kernel?"""
import sys, os, time, threading
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
Q = ctx.upload(q)
Ts = [ctx.upload(t) for t in trains[:64]]
st = torch.cuda.Stream()

def upload_only(pinned):
    up = ctx.upload_pinned if pinned else ctx.upload
    pool = ThreadPoolExecutor(4)
    def one(ids):
        Te = [up(trains[i]) for i in ids]
        for t in Te: t.free()
    chunks = [list(range(i, min(i + 18, NP))) for i in range(0, NP, 18)]
    return lambda: list(pool.map(one, chunks))

stop = False
def gpu_load():
    n = 0
    while not stop:
        for _ in range(4):
            ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st.cuda_stream)
        st.synchronize(); n += 4
    gpu_load.n = n

for pinned in (True, False):
    fn = upload_only(pinned)
    fn(); ctx.synchronize()
    t0 = time.perf_counter(); fn(); ctx.synchronize(); alone = time.perf_counter() - t0
    stop = False
    th = threading.Thread(target=gpu_load); th.start()
    time.sleep(0.05)
    t0 = time.perf_counter()
    for _ in range(5): fn()
    ctx.synchronize()
    busy = (time.perf_counter() - t0) / 5
    stop = True; th.join()
    print(f"{'zero-copy' if pinned else 'staged   '} upload of 210 frames: alone {alone*1e3:.1f} ms ({GB/alone:.1f} GB/s), "
          f"with the tcgen05 kernel running back-to-back {busy*1e3:.1f} ms ({GB/busy:.1f} GB/s)", flush=True)

def full(workers, chunk, pinned):
    up = ctx.upload_pinned if pinned else ctx.upload
    chunks = [list(range(i, min(i + chunk, NP))) for i in range(0, NP, chunk)]
    pool = ThreadPoolExecutor(workers)
    def one(Qe, ids):
        Te = [up(trains[i]) for i in ids]
        r = ctx.matchBatch(Qe, Te, MatcherType.SIFT_BF, 0.7, out=out_buf[ids[0]:ids[-1]+1], n_out=n_buf[ids[0]:ids[-1]+1])
        for t in Te: t.free()
        return r
    def step():
        Qe = up(q)
        list(pool.map(lambda ids: one(Qe, ids), chunks))
        Qe.free()
    return step
for w, c, p in ((4, 18, True), (4, 18, False), (8, 14, False), (8, 27, False)):
    fn = full(w, c, p); fn(); fn()
    best = 1e9
    for _ in range(4):
        ctx.synchronize(); t0 = time.perf_counter(); fn(); ctx.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"full {'zero-copy' if p else 'staged'} workers={w} chunk={c}: {best*1e3:.1f} ms  {NP/best:.0f} pairs/s", flush=True)
