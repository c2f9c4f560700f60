"""Developer probe: Sampson counting kernel, 16 pairs x 2048 hypotheses x 5000 matches (cfg5 shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context
from slam_indoor_code_b200 import camera_translation as ct
torch.zeros(1, device="cuda")
ctx = Context(0)
p1, p2, R, tv = synth.two_view(5000, 5000)
E = synth.pose_hypotheses(2048, R, tv, 5001)
P = 16
ctx.profile_enable(True)
for _ in range(3): c, b, m = ct.scoreEssentialBatch(ctx, [p1] * P, [p2] * P, synth.SAMSUNG_HV_4K, np.stack([E] * P), 5.0)
ctx.profile_read()
for _ in range(10): c, b, m = ct.scoreEssentialBatch(ctx, [p1] * P, [p2] * P, synth.SAMSUNG_HV_4K, np.stack([E] * P), 5.0)
pr = ctx.profile_read()
ms, n = pr["ransac"]
print(f"RS_HPT={os.environ.get('SLAMB200_RS_HPT', 'default')}: {ms / n / P * 1e3:.2f} us per 2048x5000 pair; 33 fp64 instr per score -> "
      f"{2048 * 5000 * 33 / (ms / n / P * 1e-3) / 1e12:.2f} T fp64 instr/s; counts checksum {int(c.sum())} best {b[:2]}")
