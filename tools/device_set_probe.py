"""Developer probe: the in-process device set on the 210-pair window -- per-member device times and
the host call, with the members driven by worker threads or by the calling thread alone."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth_inputs as synth
from slam_indoor_code_b200.device_set import DeviceSet
from slam_indoor_code_b200.feature_matching import MatcherType
n = torch.cuda.device_count()
q = synth.sift_like(10000, 3000)
trains = [synth.sift_train_from_query(q, 10000, 3001 + p) for p in range(210)]
print("mode", "threads" if os.environ.get("SLAMB200_SET_THREADS", "1") != "0" else "sequential", "devices", n, flush=True)
with DeviceSet(n) as ds:
    Q = ds.upload(q)
    Ts = [ds.upload(t, ds.owner(i, 210)) for i, t in enumerate(trains)]
    for _ in range(5):
        ds.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7); ds.batchFetch()
    te, tf, dm = [], [], []
    for _ in range(10):
        t0 = time.perf_counter()
        ds.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7)
        t1 = time.perf_counter()
        got, n_out, ms = ds.batchFetch()
        t2 = time.perf_counter()
        te.append(t1 - t0); tf.append(t2 - t1); dm.append(ms.copy())
    print("enqueue host ms", np.round(np.median(te) * 1e3, 3), "fetch host ms", np.round(np.median(tf) * 1e3, 3))
    print("per-member device ms (median)", np.round(np.median(np.array(dm), axis=0), 3))
