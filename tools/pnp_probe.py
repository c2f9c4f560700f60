"""Developer probe: solvePnPRansac scoring kernel, 16 frames x 2048 poses x 5000 points (f-2 shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context
from slam_indoor_code_b200 import pnp_ransac as pr
torch.zeros(1, device="cuda")
ctx = Context(0)
obj, img, Rp, tp = synth.pnp_scene(5000, 7000)
poses = synth.pnp_hypotheses(2048, Rp, tp, 7001)
P = 16
run = lambda: pr.scorePnPBatch(ctx, [obj] * P, [img] * P, synth.SAMSUNG_HV_4K, synth.REF_DIST5, np.stack([poses] * P), 8.0)
ctx.profile_enable(True)
for _ in range(3): res = run()
ctx.profile_read()
for _ in range(10): res = run()
ms, n = ctx.profile_read()["pnp"]
c = res[0]
print(f"PNP_HPT2={os.environ.get('SLAMB200_PNP_HPT2', '0')}: {ms / n / P * 1e3:.2f} us per 2048x5000 frame, "
      f"{2048 * 5000 * 53 / (ms / n / P * 1e-3) / 1e12:.2f} T fp64 ops/s (53 per score); counts checksum {int(np.asarray(c).sum())}")
