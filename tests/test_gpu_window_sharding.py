"""GPU: the keyframe window sharded over ranks -- NCCL all-gather of the frames'
descriptor rows, device-to-device descriptor sets, round-robin pairs, and the fused form that reads
the other ranks' prepared operands over NVLink through CUDA IPC -- equals the single-process
slamb200_match_window.  Two tests: the plumbing with ONE rank (runs on any box: the NCCL calls, the
device uploads and the pair deal, but no exchange between GPUs), and the real thing with 2-4 ranks
including a cfg4-sized window (8 frames x 50 000 rows), which is SKIPPED -- visibly, not passed --
on a box with a single GPU.  bench.py --gpus N repeats the multi-rank check at every N
(window_extras.cfg4_window_8x50k.sharded_equals_single_gpu_window)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["REPO_ROOT"])
import numpy as np, torch, torch.distributed as dist
from oracle import synth
from slam_indoor_code_b200 import window_sharding as ws
from slam_indoor_code_b200.feature_matching import Context, MatcherType
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
ctx = Context(local)
ok = True
for matcher, F, sizes in ((MatcherType.SIFT_BF, 6, [3000, 2500, 1, 2049, 0, 4100]),
                          (MatcherType.ORB_BF, 5, [1500, 700, 1300, 2, 900])):
    if matcher == MatcherType.ORB_BF:
        frames = [synth.orb_pair(max(s, 1), 1, 4100 + f, planted=0)[0][:s] for f, s in enumerate(sizes)]
    else:
        base = synth.sift_like(4100, 4200)
        frames = [synth.sift_train_from_query(base, max(s, 1), 4201 + f)[:s] for f, s in enumerate(sizes)]
    local_frames = {f: a for f, a in enumerate(frames) if ws.frame_owner(f, world) == rank}
    out, counts = ws.match_window_on_gpus(ctx, local_frames, F, matcher, 0.7, dist, dev)
    # the fused form: prepared operands read over NVLink through CUDA IPC, nothing gathered
    out2, counts2 = ws.match_window_peer(ctx, local_frames, F, matcher, 0.7, dist, dev)
    ok = ok and counts2 == counts and sorted(out2.keys()) == sorted(out.keys())
    ok = ok and all(np.array_equal(out2[p], out[p]) for p in out)
    # single-process reference on this rank's own GPU
    sets = [ctx.upload(a) for a in frames]
    ref = ctx.matchWindow(sets, matcher, 0.7)
    pairs = ws.window_pairs(F)
    ok = ok and counts == [len(ref[p]) for p in pairs]
    ok = ok and sorted(out.keys()) == ws.my_window_pairs(rank, world, F)
    ok = ok and all(np.array_equal(out[p], ref[p]) for p in out)
if world > 1 and os.environ.get("WINDOW_CFG4") == "1":
    # cfg4 at full size: 8 frames x 50 000 rows, frames owned round-robin, both exchange forms
    F, ROWS = 8, 50000
    frames = {f: synth.sift_like(ROWS, 4000 + f) for f in range(F) if ws.frame_owner(f, world) == rank}
    out, counts = ws.match_window_on_gpus(ctx, frames, F, MatcherType.SIFT_BF, 0.7, dist, dev)
    out2, counts2 = ws.match_window_peer(ctx, frames, F, MatcherType.SIFT_BF, 0.7, dist, dev)
    ok = ok and counts2 == counts and all(np.array_equal(out2[p], out[p]) for p in out)
    if rank == 0:
        sets = [ctx.upload(synth.sift_like(ROWS, 4000 + f)) for f in range(F)]
        ref = ctx.matchWindow(sets, MatcherType.SIFT_BF, 0.7)
        ok = ok and counts == [len(ref[p]) for p in ws.window_pairs(F)]
        ok = ok and all(np.array_equal(out[p], ref[p]) for p in out)
print("RESULT " + json.dumps({"rank": rank, "ok": bool(ok), "world": world}), flush=True)
dist.destroy_process_group()
'''


def _run(tmp_path, world, port, cfg4):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, WINDOW_CFG4="1" if cfg4 else "0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = [json.loads(l[len("RESULT "):]) for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert len(res) == world and all(d["ok"] and d["world"] == world for d in res), res


def test_window_plumbing_with_one_rank(tmp_path):
    _run(tmp_path, 1, 29541, False)


def test_window_sharded_over_ranks_equals_single_process(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"the sharded window needs >= 2 GPUs ({n} visible): run under gpurun --gpus N; "
                    "bench.py --gpus N repeats this check at every N")
    _run(tmp_path, min(n, 4), 29542, True)
