python -m pytest tests/test_gpu_sift_descriptors.py tests/test_gpu_host_cpp.py tests/test_gpu_orb_descriptors.py -m gpu -x -q > gpurun_out/r02_gputest_j.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_j.log; tail -30 gpurun_out/r02_gputest_j.log
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context
from slam_indoor_code_b200 import sift_descriptors as sd
ctx = Context(0)
frame = synth.textured_frame(2160, 3840, 6000, 3)
rng = np.random.default_rng(6001)
kps = np.stack([rng.integers(40, 3800, 12000), rng.integers(40, 2120, 12000), np.full(12000, 7.0), np.full(12000, -1.0)], 1).astype(np.float32)
for _ in range(2): sd.extractDescriptorSIFT(ctx, frame, kps, want_host=False, want_resident=True)[1].free()
ctx.profile_enable(True); ctx.profile_read()
t0 = time.perf_counter()
for _ in range(5): sd.extractDescriptorSIFT(ctx, frame, kps, want_host=False, want_resident=True)[1].free()
dt = (time.perf_counter() - t0) / 5
print("sift 4k 12000 kp: host call ms", dt * 1e3, "kernels", ctx.profile_read()["sift_desc"])
PY
