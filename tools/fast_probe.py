"""Developer probe: FAST detector timing on a 4K frame (host call and kernels) against cv2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2, numpy as np
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context
from slam_indoor_code_b200 import fast_extractor as fe
ctx = Context(0)
frame = synth.textured_frame(2160, 3840, 6000, 3)
n = len(fe.fastExtractor(ctx, frame, 10, True))
ctx.profile_enable(True); ctx.profile_read()
t0 = time.perf_counter()
for _ in range(10): fe.fastExtractor(ctx, frame, 10, True, max_points=n)
dt = (time.perf_counter() - t0) / 10
kms, kn = ctx.profile_read()["fast"]
d = cv2.FastFeatureDetector_create(10, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
d.detect(frame); t0 = time.perf_counter(); d.detect(frame); tc = time.perf_counter() - t0
print(f"4K frame, {n} keypoints: host call {dt*1e3:.2f} ms, kernels {kms/max(kn,1)*1e3:.0f} us; cv2 on {cv2.getNumThreads()} threads {tc*1e3:.1f} ms")
