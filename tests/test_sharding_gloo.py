"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py -- the contiguous pair shards, the
gather of per-pair results into one array, and the reference arm's "rank 0 only" rule -- without a
GPU.  The data path has no collective; the only exchange is the result gather."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pair_shards_partition_the_window():
    import bench
    for world in (1, 2, 3, 4, 8):
        got = [p for r in range(world) for p in bench.my_pairs(r, world)]
        assert got == list(range(bench.N_PAIRS))
        sizes = [len(bench.my_pairs(r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"])
import numpy as np, torch, torch.distributed as dist
import bench
from oracle import c_oracle, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# a small window: 1 query frame x 6 train frames, sharded like bench.run_b200 does
P, rows = 6, 300
pairs = bench.my_pairs(rank, world, P)
q = synth.sift_like(rows, 3000)
n_good = torch.zeros(P, dtype=torch.int32)
for p in pairs:   # the CPU oracle stands in for the device here: this test is about the plumbing
    n_good[p] = len(c_oracle.match_features(0, q, synth.sift_train_from_query(q, rows, 3001 + p), 0.7))
dist.all_reduce(n_good, op=dist.ReduceOp.SUM)
ms = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)   # max-over-ranks timing rule
if rank == 0:
    print("RESULT", n_good.tolist(), float(ms.item()), flush=True)
dist.destroy_process_group()
'''


def test_world_size_2_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0]
    counts, ms = eval(line[len("RESULT"):].strip().replace("] ", "], ", 1))
    # every pair was matched by exactly one rank, and the gathered counts equal a single-process run
    from oracle import c_oracle, synth
    q = synth.sift_like(300, 3000)
    ref = [len(c_oracle.match_features(0, q, synth.sift_train_from_query(q, 300, 3001 + p), 0.7)) for p in range(6)]
    assert counts == ref and ms == 11.0


def test_reference_arm_runs_on_rank0_only():
    """Under torchrun the CPU arm prints its line on rank 0; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "3"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "3", "--ref-pairs", "1"], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0
    assert d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0


# ---- the keyframe window across ranks: the one real exchange step -----------
def test_window_pair_deal_is_a_balanced_partition():
    from slam_indoor_code_b200 import window_sharding as ws
    for n_frames in (2, 5, 8):
        allp = ws.window_pairs(n_frames)
        assert len(allp) == n_frames * (n_frames - 1) // 2
        for world in (1, 2, 3, 8):
            shares = [ws.my_window_pairs(r, world, n_frames) for r in range(world)]
            assert sorted(p for s in shares for p in s) == sorted(allp)
            assert max(map(len, shares)) - min(map(len, shares)) <= 1


WINDOW_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"])
import numpy as np, torch, torch.distributed as dist
from oracle import c_oracle, synth
from slam_indoor_code_b200 import window_sharding as ws
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
F = 5
sizes = [300, 1, 257, 0, 411]                      # ragged frames, an empty one, a single row
frames = {f: synth.sift_like(max(sizes[f], 1), 4000 + f)[: sizes[f]] for f in range(F)}
local = {f: a for f, a in frames.items() if ws.frame_owner(f, world) == rank}
# the CPU oracle stands in for the device: this test is about the exchange and the pair deal
out, counts, _ = ws.match_window_sharded(
    local, F, 128, np.float32, dist, "cpu",
    upload_fn=lambda t: t.numpy().copy(),
    match_fn=lambda q, ts: [c_oracle.match_features(0, q, t, 0.8) for t in ts])
ok = all(np.array_equal(m, c_oracle.match_features(0, frames[i], frames[j], 0.8)) for (i, j), m in out.items())
import json
print("RESULT " + json.dumps({"rank": rank, "keys": sorted(map(list, out.keys())), "counts": counts, "ok": bool(ok)}), flush=True)
dist.destroy_process_group()
'''


def test_window_exchange_world_size_2(tmp_path):
    script = tmp_path / "window_worker.py"
    script.write_text(WINDOW_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29534", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    from oracle import c_oracle, synth
    from slam_indoor_code_b200 import window_sharding as ws
    sizes = [300, 1, 257, 0, 411]
    frames = [synth.sift_like(max(s, 1), 4000 + f)[:s] for f, s in enumerate(sizes)]
    want = [len(c_oracle.match_features(0, frames[i], frames[j], 0.8)) for i, j in ws.window_pairs(5)]
    seen = []
    for r, (so, _) in enumerate(outs):
        line = [l for l in so.splitlines() if l.startswith("RESULT ")][0]
        d = json.loads(line[len("RESULT "):])
        assert d["rank"] == r and d["ok"] and d["counts"] == want   # all counts everywhere; own matches right
        seen += [tuple(k) for k in d["keys"]]
    assert sorted(seen) == ws.window_pairs(5)   # every pair matched by exactly one rank
