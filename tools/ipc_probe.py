"""Cost of each step of the peer-memory path on 2 ranks: shared upload, export, import, localize,
and a 50k x 50k pair matched with the query read over NVLink vs both sets local."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import synth_inputs as synth
from slam_indoor_code_b200.feature_matching import Context, MatcherType
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
ctx = Context(local)
ROWS = 50000
a = synth.sift_like(ROWS, 4000 + rank)
def t(fn, n=3):
    best = 1e9; r = None
    for _ in range(n):
        ctx.synchronize(); t0 = time.perf_counter(); r = fn(); ctx.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, r
ms_up, own = t(lambda: ctx.upload_shared(a), 1)
ms_up2, own_plain = t(lambda: ctx.upload(a), 1)
ms_exp, rec = t(lambda: own.export_ipc())
recs = torch.zeros((world, 128), dtype=torch.uint8, device=dev)
recs[rank] = torch.frombuffer(bytearray(rec), dtype=torch.uint8).to(dev)
dist.all_reduce(recs)
other = (rank + 1) % world
ms_imp, peer = t(lambda: ctx.import_ipc(recs[other].cpu().numpy().tobytes()), 1)
ms_loc, loc = t(lambda: ctx.localize(peer))
ms_remote_q, m1 = t(lambda: ctx.matchBatch(peer, [own], MatcherType.SIFT_BF, 0.7))
ms_local, m2 = t(lambda: ctx.matchBatch(loc, [own], MatcherType.SIFT_BF, 0.7))
ms_remote_t, m3 = t(lambda: ctx.matchBatch(own, [peer], MatcherType.SIFT_BF, 0.7), 1)
ms_local_t, m4 = t(lambda: ctx.matchBatch(own, [loc], MatcherType.SIFT_BF, 0.7), 1)
ok = np.array_equal(m1[0], m2[0]) and np.array_equal(m3[0], m4[0])
st = torch.cuda.Stream(); 
def ev(qq, tt):
    with torch.cuda.stream(st):
        for _ in range(3): ctx.matchBatchEnqueue(qq, [tt], MatcherType.SIFT_BF, 0.7, st.cuda_stream)
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(st)
        for _ in range(10): ctx.matchBatchEnqueue(qq, [tt], MatcherType.SIFT_BF, 0.7, st.cuda_stream)
        b_.record(st); st.synchronize()
    return a_.elapsed_time(b_) / 10
dev_ms = {"remote_query": ev(peer, own), "all_local": ev(loc, own)}
prof = {}
for name, (qq, tt) in {"remote_query": (peer, own), "all_local": (loc, own), "remote_train": (own, peer)}.items():
    ctx.profile_enable(True); ctx.profile_read()
    for _ in range(2): ctx.matchBatch(qq, [tt], MatcherType.SIFT_BF, 0.7)
    pr = ctx.profile_read(); ctx.profile_enable(False)
    prof[name] = {k: round(v[0] / max(v[1], 1), 3) for k, v in pr.items() if v[1]}
print(json.dumps({"rank": rank, "upload_shared_ms": ms_up, "upload_plain_ms": ms_up2, "export_ms": ms_exp, "import_ms": ms_imp,
                  "localize_ms": ms_loc, "match_remote_query_ms": ms_remote_q, "match_all_local_ms": ms_local,
                  "match_remote_train_ms": ms_remote_t, "match_local_train_ms": ms_local_t, "equal": bool(ok), "enqueue_only_device_ms": dev_ms, "kernel_ms": prof}), flush=True)
dist.barrier(device_ids=[local])
peer.free(); loc.free()
dist.barrier(device_ids=[local])
own.free()
dist.destroy_process_group()
