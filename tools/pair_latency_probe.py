"""Developer probe: where a single-pair SIFT call spends its time (fixed cost vs per-pair cost,
host enqueue cost, per-kernel device times)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType

torch.cuda.init(); torch.zeros(1, device="cuda")
ctx = Context(0)
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
q = synth.sift_like(10000, 3000)
Q = ctx.upload(q)
Ts = [ctx.upload(synth.sift_train_from_query(q, 10000, 3001 + i)) for i in range(16)]

def ev_time(fn, iters, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(iters): fn()
    b.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3, t_host / iters * 1e6

for n in (1, 2, 4, 8, 16):
    dev_us, host_us = ev_time(lambda: ctx.matchBatchEnqueue(Q, Ts[:n], MatcherType.SIFT_BF, 0.7, st), 200)
    print(f"n_pairs={n:2d}: {dev_us:8.1f} us/call device ({dev_us/n:6.1f} us/pair), host enqueue {host_us:6.1f} us/call", flush=True)
ctx.profile_enable(True); ctx.profile_read()
for _ in range(200): ctx.matchBatchEnqueue(Q, Ts[:1], MatcherType.SIFT_BF, 0.7, st)
p = ctx.profile_read()
print({k: (round(v[0] / max(v[1], 1) * 1e3, 2), v[1]) for k, v in p.items() if v[1]})
# the synchronous drop-in call (slamb200_match_pair: enqueue + D2H of the matches + sync)
t0 = time.perf_counter()
for _ in range(200): ctx.matchFeatures(Q, Ts[0], MatcherType.SIFT_BF, 0.7)
print(f"match_pair host call: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
import threading
def worker(k, n):
    for _ in range(n): ctx.matchFeatures(Q, Ts[k], MatcherType.SIFT_BF, 0.7)
for nt in (2, 4, 8):
    th = [threading.Thread(target=worker, args=(k, 200)) for k in range(nt)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"match_pair from {nt} host threads: {dt / (200 * nt) * 1e6:.1f} us/pair aggregate")
