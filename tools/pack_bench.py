"""Host-side narrowing throughput on this machine (csrc/host_pack.cpp) vs thread count."""
import sys, os, time, ctypes, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slam_indoor_code_b200 import _capi
lib = _capi.load()
fn = lib.slamb200_host_pack_u8
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
F = 210
rng = np.random.default_rng(0)
src = [rng.integers(0, 256, (10000, 128)).astype(np.float32) // 4 for _ in range(F)]
dst = [np.zeros((10000, 128), np.uint8) for _ in range(48)]
for nt in (1, 2, 4, 8, 12, 16, 24, 32, 48):
    def work(t):
        for f in range(t, F, nt):
            assert fn(src[f].ctypes.data, 128, 10000, dst[t].ctypes.data) == 1
    best = 1e9
    for _ in range(3):
        th = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
        t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]
        best = min(best, time.perf_counter() - t0)
    print(f"threads={nt:2d}: {best*1e3:6.1f} ms per 210 frames, {best/F*1e6:6.1f} us/frame, {F*5.12e6/best/1e9:5.1f} GB/s read", flush=True)
