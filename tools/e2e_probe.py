"""Where does an e2e step spend its wall time? (phases separated by device synchronisation)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from slam_indoor_code_b200.feature_matching import Context, MatcherType
torch.zeros(1, device="cuda")
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 210
q, trains = bench.make_inputs(list(range(NP)), pinned=True)
ctx = Context(0)
for it in range(4):
    t0 = time.perf_counter()
    Q = ctx.upload_pinned(q); Ts = [ctx.upload_pinned(t) for t in trains]
    t1 = time.perf_counter(); ctx.synchronize(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    res = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
    t3 = time.perf_counter()
    for t in Ts: t.free()
    Q.free()
    t4 = time.perf_counter()
    print(f"iter {it}: upload calls {1e3*(t1-t0):.1f} ms, upload drain {1e3*(t2-t1):.1f} ms, "
          f"matchBatch {1e3*(t3-t2):.1f} ms, free {1e3*(t4-t3):.1f} ms, total {1e3*(t4-t0):.1f} ms", flush=True)
# matchBatch split: enqueue / fetch
Q = ctx.upload_pinned(q); Ts = [ctx.upload_pinned(t) for t in trains]; ctx.synchronize()
_ts = torch.cuda.Stream(); torch.cuda.set_stream(_ts); st = _ts.cuda_stream
for it in range(3):
    t0 = time.perf_counter()
    ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7, st)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    m, n = ctx.batchFetch(st)
    t3 = time.perf_counter()
    print(f"enqueue call {1e3*(t1-t0):.2f} ms, device {1e3*(t2-t1):.2f} ms, fetch {1e3*(t3-t2):.2f} ms")
