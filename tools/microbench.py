"""Box probes: pinned H2D/D2H bandwidth, POPC and FP64 pipe rates."""
import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_indoor_code_b200 import _capi
lib = _capi.load()
lib.slamb200_dbg_pipe_rate.restype = ctypes.c_double
lib.slamb200_dbg_pipe_rate.argtypes = [ctypes.c_int]
torch.zeros(1, device="cuda")
print(f"POPC rate  : {lib.slamb200_dbg_pipe_rate(0):.0f} Gpopc/s")
print(f"FP64 rate  : {lib.slamb200_dbg_pipe_rate(1):.0f} G(dmul+dadd)/s")
for mb in (5, 64, 1024):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20 if mb < 1000 else 3
    a.record()
    for _ in range(n): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    up = mb * n / 1024 / (a.elapsed_time(b) / 1e3)
    a.record()
    for _ in range(n): h.copy_(d, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    dn = mb * n / 1024 / (a.elapsed_time(b) / 1e3)
    print(f"pinned copy {mb:5d} MiB: H2D {up:.1f} GiB/s  D2H {dn:.1f} GiB/s")
os.system("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv")
# 5 MiB H2D copies spread over several streams (do the per-copy setup costs overlap?)
mb = 5
hs = [torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory() for _ in range(8)]
ds = [torch.empty_like(h, device="cuda") for h in hs]
for ns in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 64
    for i in range(n):
        with torch.cuda.stream(streams[i % ns]):
            ds[i % 8].copy_(hs[i % 8], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"5 MiB H2D x{n} over {ns} stream(s): {mb*n/1024/dt:.1f} GiB/s")
# raw cudaMemcpyAsync through ctypes (no torch overhead)
cudart = ctypes.CDLL("libcudart.so.12")
st = ctypes.c_void_p()
cudart.cudaStreamCreate(ctypes.byref(st))
for mbs in (5, 20):
    h = torch.empty(mbs * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 64
    for i in range(n):
        cudart.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(h.data_ptr()), ctypes.c_size_t(mbs << 20), 1, st)
    cudart.cudaStreamSynchronize(st)
    dt = time.perf_counter() - t0
    print(f"raw cudaMemcpyAsync {mbs} MiB x{n}: {mbs*n/1024/dt:.1f} GiB/s")
