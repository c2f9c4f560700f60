"""Developer probe: host-call time of the per-frame producers (FAST, ORB, SIFT descriptors) on a 4K BGR
frame in pageable memory, every worker lane warmed (a lane's first 4K call allocates ~90 MB of
scratch).  Staging the frame through page-locked memory on the pack pool was tried and is slower
than the driver's own pageable copy (ORB 2.2 against 1.5 ms per call)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context
from slam_indoor_code_b200 import fast_extractor as fe, orb_descriptors as od, sift_descriptors as sdm
torch.zeros(1, device="cuda")
ctx = Context(0)
frame = synth.textured_frame(2160, 3840, 6000, 3)
rng = np.random.default_rng(6001)
xy = np.stack([rng.integers(31, 3840 - 31, 12000), rng.integers(31, 2160 - 31, 12000)], 1).astype(np.float32)
kp_orb = np.concatenate([xy, np.full((12000, 1), -1.0, np.float32)], 1)
kp_sift = np.concatenate([xy, np.full((12000, 1), 7.0, np.float32), np.full((12000, 1), -1.0, np.float32)], 1)
n_fast = len(fe.fastExtractor(ctx, frame, 10, True))
def t(fn, n=10, warm=6):
    for _ in range(warm): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
print(f"FAST {t(lambda: fe.fastExtractor(ctx, frame, 10, True, max_points=n_fast)):.2f} ms/call ({n_fast} keypoints back)")
print(f"ORB  {t(lambda: od.extractDescriptorORB(ctx, frame, kp_orb, want_host=False, want_resident=True)[2].free()):.2f} ms/call (resident)")
print(f"SIFT {t(lambda: sdm.extractDescriptorSIFT(ctx, frame, kp_sift, want_host=False, want_resident=True)[1].free()):.2f} ms/call (resident)")
print(f"SIFT {t(lambda: sdm.extractDescriptorSIFT(ctx, frame, kp_sift, want_host=True, want_resident=False)):.2f} ms/call (descriptors back on the host)")
