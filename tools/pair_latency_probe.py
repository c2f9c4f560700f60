"""Developer probe: where a single-pair SIFT call spends its time (fixed cost vs per-pair cost,
host enqueue cost, per-kernel device times)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
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
for _ in range(100): ctx.matchFeatures(Q, Ts[0], MatcherType.SIFT_BF, 0.7)   # every lane warmed (pinned result buffers)
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
# the C++ drop-in unit (featureMatchingB200.cpp::matchFeatures: upload both Mats, match, fetch)
import ctypes
from slam_indoor_code_b200 import build as _b
hs = ctypes.CDLL(os.path.join(_b.LIBDIR, "libslamb200_hostshim.so"))
hs.hostshim_match_features.restype = ctypes.c_int
hs.hostshim_match_features.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
from slam_indoor_code_b200._capi import DMATCH
tr = [synth.sift_train_from_query(q, 10000, 3001 + i) for i in range(8)]
def cpp_worker(k, n, pinned=False):
    out = np.zeros(10000, DMATCH)
    for _ in range(n):
        r = hs.hostshim_match_features(q.ctypes.data, 10000, 512, tr[k].ctypes.data, 10000, 512, 0, out.ctypes.data, 10000)
        assert r > 0
cpp_worker(0, 5)
t0 = time.perf_counter(); cpp_worker(0, 100); dt = time.perf_counter() - t0
print(f"C++ matchFeatures (pageable Mats, 2 uploads + match): {dt / 100 * 1e6:.1f} us/pair, 1 thread")
for nt in (3, 6):
    th = [threading.Thread(target=cpp_worker, args=(k, 100)) for k in range(nt)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"C++ matchFeatures from {nt} host threads: {dt / (100 * nt) * 1e6:.1f} us/pair aggregate")
