#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the SIFT kNN-2 + ratio hot path on the framesBatchSize window.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path: OpenCV)

Workload (BASELINE.json configs[2], the one its metric is quoted on): one query frame against a
batch of 210 train frames, 10,000 SIFT descriptors (128-d) per frame, BF L2 kNN k=2 + Lowe ratio
0.7 -- the window findGoodFramesFromBatch walks (batch.cpp:120-148).  A step is one pass over the
whole 210-pair batch.  With N GPUs the 210 pairs are split contiguously over the ranks (strong
scaling, no data-path collective); NCCL only gathers the per-pair match counts after the timed
region.

  value  = pairs/s with every descriptor set already resident in HBM, CUDA-event timed.
  e2e    = pairs/s through the C ABI with HOST buffers: every step hands the query and the 210
           train descriptor Mats (CV_32F, PAGEABLE host memory) to slamb200_match_batch_host, which
           narrows the integer-valued rows to bytes on host threads (every element verified: 1/4 of
           the bytes cross PCIe), uploads, matches chunk by chunk and copies the match lists back.
  roofline = the tcgen05 candidate kernel: 2*Q*T*128 FLOP per pair / its CUDA-event duration,
           against the measured cuBLAS bf16 peak in MEASURED_PEAKS.json.
  cpu_baseline = the same OpenCV calls the reference makes (cv2.BFMatcher.knnMatch + the
           getGoodMatches loop) on the box's host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PAIRS = 210
N_ROWS = 10000
RATIO = 0.7
FLOP_PER_PAIR = 2.0 * N_ROWS * N_ROWS * 128


_REAL_STDOUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def my_pairs(rank, world, n=N_PAIRS):
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return list(range(lo, hi))


def make_inputs(pairs, pinned):
    """Query frame + the owned train frames (seeds of SURVEY.md 8d cfg3), as float32 N x 128."""
    import torch
    import synth_inputs as synth
    q = synth.sift_like(N_ROWS, 3000)

    def hold(a):
        if not pinned:
            return a
        t = torch.empty(a.shape, dtype=torch.float32).pin_memory()
        v = t.numpy()
        v[...] = a
        return v

    keep = [hold(q)]
    trains = []
    for p in pairs:
        t = hold(synth.sift_train_from_query(q, N_ROWS, 3001 + p))
        trains.append(t)
    return keep[0], trains


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=5.0):
        """Blocks until nvidia-smi has delivered its first sample (it takes ~0.1-1 s to start)."""
        t = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.01)

    def stop(self, t0, t1, t_load0=None):
        """Samples inside the timed region [t0, t1]; when the region is too short to hold three
        (N=8: 20 steps last ~10 ms) the window is widened to the warm-up steps that ran the same
        kernels immediately before it, and the result says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + 0.02]
        window = "timed region"
        if len(rows) < 3 and t_load0 is not None:
            rows = [r for r in self.rows if t_load0 <= r[0] <= t1 + 0.02]
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, \
        "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the OpenCV calls the reference makes (featureMatchingCPU.cpp:27-42)
# ------------------------------------------------------------------------------------------------
def cpu_match_pair(q, t):
    """Returns the number of good matches.  cv2 when importable (the reference's own arithmetic
    owner), else the C oracle port."""
    try:
        import cv2
    except Exception:
        cv2 = None
    if cv2 is not None:
        res = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, 2)
        n = 0
        for r in res:           # getGoodMatches, featureMatchingCommon.cpp:43-49
            if len(r) >= 2 and r[0].distance < RATIO * r[1].distance:
                n += 1
        return n
    from oracle import c_oracle
    return len(c_oracle.match_features(0, q, t, RATIO))


def cpu_threads():
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        return int(cv2.getNumThreads()), "reference", \
            f"cv2 {cv2.__version__} BFMatcher(NORM_L2).knnMatch k=2 + getGoodMatches loop"
    except Exception:
        from oracle import c_oracle
        return int(c_oracle.set_threads(os.cpu_count() or 1)), "port", \
            "oracle/corr_oracle.c (OpenMP) knn2 + ratio"


def run_reference(args, rank, world):
    """--impl reference: rank 0 times the CPU path on a bounded sample of the same batch."""
    if rank != 0:
        return
    cores, kind, what = cpu_threads()
    n_sample = max(1, args.ref_pairs)
    q, trains = make_inputs(list(range(n_sample)), pinned=False)
    for _ in range(args.warmup):
        cpu_match_pair(q, trains[0])
    t0 = time.perf_counter()
    good = 0
    for _ in range(args.steps):
        for t in trains:
            good += cpu_match_pair(q, t)
    dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    sample = (f"{n_sample} of the {N_PAIRS} pairs per step ({what}); the reference binary itself "
              "cannot be built here (no OpenCV C++ dev files), these are the calls it makes")
    line = {
        "impl": "reference", "metric": "SIFT 10k x 10k kNN+ratio frame-pairs/s", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "tflops": value * FLOP_PER_PAIR / 1e12,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(world):
    cfg = {"workload": "cfg3: framesBatchSize=210 window, 1 query frame x 210 train frames, "
                       "10000 SIFT 128-d descriptors per frame, BF L2 kNN k=2 + ratio 0.7",
           "pairs": N_PAIRS, "rows_per_frame": N_ROWS, "ratio": RATIO,
           "sharding": f"{N_PAIRS} pairs split contiguously over {world} rank(s); no data-path collective",
           "l2": "inputs larger than L2 (211 resident descriptor sets, 0.6 GB of bf16 operands per step)"}
    return cfg


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from slam_indoor_code_b200.feature_matching import Context, MatcherType

    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a host-side barrier for the section in which rank 0 alone drives every GPU: an NCCL barrier
        # would leave a spinning kernel on the other ranks' GPUs for as long as they wait
        cpu_group = dist.new_group(backend="gloo")
    dev = torch.device("cuda", local)
    torch.zeros(1, device=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    pairs = my_pairs(rank, world)
    t_gen = time.perf_counter()
    # PAGEABLE host Mats, as cv::SIFT::compute leaves them (a cv::Mat is never page-locked); only
    # --e2e-upload pinned (an explicit experiment) allocates them page-locked
    q, trains = make_inputs(pairs, pinned=args.e2e_upload == "pinned")
    log(f"[rank {rank}] generated {len(trains)} train frames in {time.perf_counter() - t_gen:.1f}s")

    ctx = Context(local)
    # host cores this rank may use: the ranks of a node share them.  The library's pack pool is
    # sized before the first upload (every upload narrows integer-valued Mats on the host).
    cores = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    if args.e2e_uploaders > 0:
        os.environ["SLAMB200_HOST_NARROWERS"] = str(args.e2e_uploaders)   # read once by the library, before its first host call
    if args.e2e_upload == "auto":
        args.e2e_upload = "packed"
    # host threads of this rank that narrow Mats inside the library (its persistent pack pool): its
    # cores minus one -- the submitting and the matching thread mostly wait, but with every core
    # narrowing they are scheduled late and the step takes twice as long (measured: 15 of 16 is the
    # best split at N = 1); with several ranks on the box, one more core per rank is left alone
    pack_threads = args.e2e_pack_threads if args.e2e_pack_threads >= 0 else \
        max(1, min(cores, 25) - (1 if world == 1 else 2))
    ctx.set_pack_threads(pack_threads)
    # A real (non-default) stream: the C ABI treats a NULL stream as "the lane's own stream", so
    # the kernels and the CUDA events that time them must share an explicit stream handle.
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    assert Q.exact_mode == 1 and all(t.exact_mode == 1 for t in Ts[:2]), \
        "synthetic SIFT rows must take the tcgen05 path"

    def step():
        ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, RATIO, stream)

    # ---- device-resident throughput (the headline `value`) ------------------------------------
    sampler = ClockSampler(local)
    sampler.wait_first()
    t_load0 = time.perf_counter()
    # the contract's W warm-up steps, repeated until at least 0.25 s of the same load has run so
    # that the clock sampler sees the GPU under this kernel even when K steps last milliseconds
    warm_run = 0
    # the warm-up steps carry events around EVERY kernel class (the tail's and the compaction's
    # durations come from them); the timed region only around the dominant kernel, whose duration
    # the roofline line needs live -- an event between two kernels is a boundary the device
    # drains to, and six of them per step cost a 0.5 ms step of the 8-rank run several per cent
    ctx.profile_enable(True)
    ctx.profile_read()
    while True:
        for _ in range(args.warmup):
            step()
        warm_run += args.warmup
        torch.cuda.synchronize()
        if time.perf_counter() - t_load0 >= 0.25 or args.warmup == 0:
            break
    prof_warm = ctx.profile_read()
    barrier()
    ctx.profile_enable(True, kinds=None if args.events_all else ["sift_tc"])
    ctx.profile_read()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect()
    gc.disable()            # as timeit does: no cyclic-GC pause while the host feeds the device
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    gc.enable()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    launches = torch.tensor([ctx.launch_count() - launches0], device=dev, dtype=torch.float64)
    clocks = sampler.stop(t0, t1, t_load0)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = N_PAIRS * args.steps / (ms_total / 1e3)

    matches, n_out = ctx.batchFetch(stream)
    n_good = torch.zeros(N_PAIRS, device=dev, dtype=torch.int32)
    if pairs:
        n_good[pairs[0]: pairs[-1] + 1] = torch.from_numpy(np.ascontiguousarray(n_out)).to(dev)
    if world > 1:
        dist.all_reduce(n_good, op=dist.ReduceOp.SUM)   # NCCL: gather the per-pair results
    n_good = n_good.cpu().numpy()

    # ---- roofline of the dominant kernel (tcgen05 candidates), this rank ------------------------
    peaks, peak_src = measured_peaks()
    tc_ms, tc_n = prof["sift_tc"]
    roof = None
    if tc_n > 0:
        # every timed step launches the kernel over this rank's pairs (possibly in sub-batches)
        achieved = FLOP_PER_PAIR * len(pairs) * args.steps / (tc_ms / 1e3) / 1e12
        # Which measured peak applies follows from the timing regime: MEASURED_PEAKS.json's sustained
        # figure is cuBLAS run back to back for seconds (SM clock settled at ~1.3 GHz under the power
        # cap), its burst figure the best of ten isolated GEMMs.  A timed region shorter than one
        # second never reaches the sustained regime (the clocks line shows it), so it is held to
        # the BURST peak; a run of a second or more to the sustained one.
        burst = float(peaks.get("bf16_tflops"))
        sustained = float(peaks.get("bf16_tflops_sustained", burst))
        timed_s = ms_total / 1e3
        use_burst = timed_s < 1.0
        peak = burst if use_burst else sustained
        roof = {"bound": "tensor", "kernel": "sift_tc_kernel (tcgen05 bf16 candidates)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": peak_src + (f", burst bf16 (timed region {timed_s:.3f} s < 1 s: the clocks never "
                                           "settle into the sustained regime)" if use_burst else
                                           f", sustained bf16 (timed region {timed_s:.1f} s)"),
                "frac_of_burst": achieved / burst, "frac_of_sustained": achieved / sustained,
                "kernel_ms_per_step": tc_ms / args.steps, "kernel_launches_per_step": tc_n / args.steps,
                "algorithmic_flop_per_step": FLOP_PER_PAIR * len(pairs),
                "other_kernels_ms_per_step": {k: v[0] / max(warm_run, 1) for k, v in prof_warm.items()
                                              if k != "sift_tc" and v[1] > 0},
                "other_kernels_note": "measured over the warm-up steps (the timed region carries events around the "
                                      "dominant kernel only); sift_rerank = the tail (slot merge + best-group rerank + "
                                      "ratio test; with up to 32 pairs per rank also the ordered compaction), "
                                      "finalize = compact_kernel (ordered compaction)",
                "kernel_share_of_step": tc_ms / ms_total
                if world == 1 else None,
                "traffic": TRAFFIC_BYTES_PER_LAUNCH_N1 if world == 1 else None,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one 210-pair launch, ncu --set full "
                                "(profiles/r02/r02_ncu_final_sift_tc_tail_compact.txt): 617.0 MB read + 93.6 MB written; "
                                "algorithmic bytes 574 MB (211 bf16 sets 540 MB + 33.6 MB of results) + 103 MB of "
                                "slot records the tail consumes",
                "tensor_pipe_active_pct": 88.2 if world == 1 else None,
                "tensor_pipe_note": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed of the same "
                                    "launch (profiles/r02/r02_ncu_tensor_pipe.txt)"}

    # ---- e2e through the C ABI with host buffers -------------------------------------------------
    for t in Ts:
        t.free()
    Q.free()
    e2e_steps = max(0, min(args.steps, args.e2e_steps))
    d2h = 0

    # One call per step through the C ABI: slamb200_match_batch_host takes the query Mat and this
    # rank's train Mats from (pageable) host memory, narrows them on the library's pack pool, queues
    # the prep kernels, matches every chunk as soon as its last Mat is resident and copies the
    # match lists back -- the call a search of the reference makes when it hands its descriptors
    # over (batch.cpp:120-148); uploads never wait for a match and vice versa, and no Python runs
    # between the Mats.
    from slam_indoor_code_b200._capi import DMATCH
    out_buf = np.empty((max(len(trains), 1), N_ROWS), DMATCH)      # caller-owned, reused per step
    n_buf = np.zeros(max(len(trains), 1), np.int32)
    if args.e2e_upload != "packed":
        log("--e2e-upload pinned is an upload experiment of the old Python pipeline; the e2e call narrows on the host")
    # bytes that cross PCIe per step: the verified byte images of the fp32 Mats
    h2d = (len(trains) + 1) * N_ROWS * 128

    def e2e_step():
        nonlocal d2h
        res = ctx.matchBatchHost(q, trains, MatcherType.SIFT_BF, RATIO, out=out_buf, n_out=n_buf)
        d2h = len(res) * 4 + sum(len(r) for r in res) * 16
        return res

    res = matches
    # warm-up until the pipeline is in steady state (staging pool, slab cache, pack threads, lazily
    # loaded kernels, first-touch effects of a fresh box): at least 4 steps, at most 16, until the
    # last three are within 1.5x of the best seen
    warm_t = []
    while e2e_steps and len(warm_t) < 16:
        ts0 = time.perf_counter()
        e2e_step()
        warm_t.append(time.perf_counter() - ts0)
        if len(warm_t) >= 4 and max(warm_t[-3:]) < 1.5 * min(warm_t):
            break
    # like timeit: no cyclic-GC pass inside the timed region (a full collection over torch's object
    # graph costs ~100 ms, several whole steps)
    import gc
    gc.collect()
    gc.disable()
    barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        ts0 = time.perf_counter()
        res = e2e_step()
        log(f"[rank {rank}] e2e step {1e3 * (time.perf_counter() - ts0):.1f} ms")
    barrier()
    te = torch.tensor([max(time.perf_counter() - te0, 1e-9)], device=dev, dtype=torch.float64)
    gc.enable()
    io = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(io, op=dist.ReduceOp.SUM)
    e2e_value = N_PAIRS * e2e_steps / float(te.item()) if e2e_steps else None
    same = all(np.array_equal(a, b) for a, b in zip(res, matches))
    # What bounds e2e on this box: every step reads the rank's fp32 Mats (5.12 MB each, far beyond
    # the CPU caches) out of host DRAM once to narrow them.  The same narrowing alone -- no GPU, no
    # PCIe, the threads this rank may use -- gives the host-side floor of the step.
    narrow = None
    if e2e_steps and trains:
        import ctypes
        fn = ctx._lib.slamb200_host_pack_u8
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        nthr = max(1, min(cores, 16))
        dst = [np.zeros((N_ROWS, 128), np.uint8) for _ in range(nthr)]

        def nwork(t):
            for f in range(t, len(trains), nthr):
                fn(trains[f].ctypes.data, 128, N_ROWS, dst[t].ctypes.data)
        best_n = 1e9
        for _ in range(3):
            barrier()
            t0n = time.perf_counter()
            th = [threading.Thread(target=nwork, args=(t,)) for t in range(nthr)]
            [x.start() for x in th]
            [x.join() for x in th]
            best_n = min(best_n, time.perf_counter() - t0n)
        tn = torch.tensor([best_n], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        narrow = {"ms_per_step_narrowing_alone": float(tn.item()) * 1e3, "threads_per_rank": nthr,
                  "host_read_GB_per_s_all_ranks": N_PAIRS * N_ROWS * 512 / float(tn.item()) / 1e9,
                  "pairs_per_s_floor": N_PAIRS / float(tn.item()),
                  "note": "slamb200_host_pack_u8 over the same pageable Mats, no GPU work: the host-DRAM-bound "
                          "floor of a step whose inputs are fp32 Mats in pageable memory"}

    shared = None
    if not args.no_extras:
        try:
            shared = multi_rank_extras(ctx, stream, rank, world, local, dev, pairs, q, trains, matches, barrier,
                                       cpu_group)
        except Exception as e:  # pragma: no cover
            shared = {"error": repr(e)}
            log(f"[rank {rank}] multi-rank extras failed: {e!r}")

    if rank == 0:
        line = {
            "metric": "SIFT 10k x 10k kNN+ratio frame-pairs/s", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": warm_run,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "tflops": value * FLOP_PER_PAIR / 1e12,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(io[0].item()),
                    "d2h_bytes_per_step": int(io[1].item()), "steps": e2e_steps,
                    "warmup_steps_run": len(warm_t), "upload": args.e2e_upload,
                    "host_buffers": "page-locked" if args.e2e_upload == "pinned" else "pageable (numpy arrays)",
                    "host_mat_bytes_per_step": (N_PAIRS + world) * N_ROWS * 512,
                    "call": "slamb200_match_batch_host (one C-ABI call per step: narrowing, uploads, matching and "
                            "result copies pipelined inside the library)",
                    "host_threads": {"narrowing": int(os.environ.get("SLAMB200_HOST_NARROWERS", pack_threads)),
                                     "submit": 1, "match_and_copy_out": 1},
                    "timing": "host wall clock between device synchronisations, max over ranks",
                    "results_equal_device_resident_run": bool(same), "host_floor": narrow},
            "gpu_launches": int(launches.item()),
            "roofline": roof,
            "checksum": {"good_matches_total": int(n_good.sum()), "pairs": int((n_good > 0).sum())},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(q, trains, matches, args.cpu_seconds)
            if not args.no_extras:
                try:   # side measurements never cost the headline line
                    line["cpu_baseline"]["other_configs"] = cpu_extras()
                except Exception as e:  # pragma: no cover
                    line["cpu_baseline"]["other_configs"] = {"error": repr(e)}
        if shared is not None:
            line["window_extras"] = shared
        if world == 1 and not args.no_extras:
            try:
                line["extras"] = extras(ctx, stream)
            except Exception as e:  # pragma: no cover
                line["extras"] = {"error": repr(e)}
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------
# Side measurements that every rank takes part in (N >= 1): BASELINE cfg3 as stated (matching +
# RANSAC essential scoring over the rank's share of the window, keypoints consistent with the
# planted correspondences), cfg4 (8 x 50k-row keyframe window) through both exchange forms with
# a self-check against the single-GPU result, and the in-process device set on rank 0.
def multi_rank_extras(ctx, stream, rank, world, local, dev, pairs, q, trains, matches, barrier, cpu_group=None):
    import torch
    import torch.distributed as dist
    import synth_inputs as synth
    from slam_indoor_code_b200 import camera_translation as ct
    from slam_indoor_code_b200 import window_sharding as ws
    from slam_indoor_code_b200.feature_matching import MatcherType
    out = {}

    def reduce_max(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_min(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    # ---- cfg3 chain: match -> device gather (getKeyPointCoordsFromFramePair) -> 2048 hypotheses ----
    H = 2048
    kq, kts_all, poses = synth.window_geometry(N_ROWS, N_ROWS, [3001 + p for p in range(N_PAIRS)], 3500)
    kts = [kts_all[p] for p in pairs]
    E = np.stack([synth.pose_hypotheses_fast(H, *poses[p], 3600 + p) for p in pairs]) if pairs else np.zeros((0, H, 9))
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    KQ = ctx.upload_keypoints(kq)
    KTs = [ctx.upload_keypoints(k) for k in kts]
    Ed = torch.from_numpy(E).to(dev)

    def chain():
        ctx.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, RATIO, stream)
        if pairs:
            ct.scoreBatchEnqueue(ctx, KQ, KTs, synth.SAMSUNG_HV_4K, None, 5.0, stream, E_device_ptr=Ed.data_ptr(), H=H)

    for _ in range(3):
        chain()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        chain()
    b.record()
    barrier()
    ms = reduce_max(a.elapsed_time(b) / 10)
    ok = True
    if pairs:
        counts, best, mask = ct.batchScoresFetch(ctx, stream)
        got, _ = ctx.batchFetch(stream)
        ok = all(np.array_equal(g, m) for g, m in zip(got, matches))
        # real geometry: the winning hypothesis explains the accepted matches
        ok = ok and all(best[i] >= 0 and counts[i, best[i]] > 0.8 * len(got[i]) for i in range(len(pairs)))
    out["cfg3_chain_match_gather_score"] = {
        "ms_per_window": ms, "us_per_pair": ms * 1e3 / N_PAIRS * world if world else None,
        "pairs_per_s": N_PAIRS / (ms / 1e3), "hypotheses_per_pair": H,
        "note": "whole 210-pair window split over the ranks; us_per_pair is per-GPU time per pair; keypoints "
                "consistent with the planted correspondences (synth_inputs.window_geometry)",
        "matches_equal_and_winner_explains_80pct": bool(reduce_min(1.0 if ok else 0.0) > 0.5)}
    for h in KTs + [KQ] + Ts + [Q]:
        h.free()
    del Ed

    # ---- cfg4: 8 frames x 50,000 rows, all 28 pairs, frames owned round-robin by the ranks ----------
    F, ROWS = 8, 50000
    frames = {f: synth.sift_like(ROWS, 4000 + f) for f in range(F) if ws.frame_owner(f, world) == rank}
    flop = 28 * 2.0 * ROWS * ROWS * 128
    best_t = 1e9
    out_a = counts_a = None
    for it in range(4):
        barrier()
        t0 = time.perf_counter()
        out_a, counts_a = ws.match_window_on_gpus(ctx, frames, F, MatcherType.SIFT_BF, RATIO, dist, dev) \
            if world > 1 else _window_single(ctx, frames, F)
        torch.cuda.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        if it >= 1:
            best_t = min(best_t, dt)
    res4 = {"n_gpus": world, "pairs": 28,
            "nccl_allgather_rows" if world > 1 else "single_gpu_host_call": {"ms_per_window": best_t * 1e3, "tflops": flop / best_t / 1e12}}
    if world > 1:
        w = ws.PeerWindow(ctx, dist, dev)
        barrier()
        t0 = time.perf_counter()
        w.publish(frames, F)
        w.match(F, MatcherType.SIFT_BF, RATIO)
        torch.cuda.synchronize()
        first = reduce_max(time.perf_counter() - t0)
        best_p = 1e9
        out_b = None
        for it in range(4):
            barrier()
            t0 = time.perf_counter()
            out_b, _ = w.match(F, MatcherType.SIFT_BF, RATIO, gather_counts=False)
            torch.cuda.synchronize()
            best_p = min(best_p, reduce_max(time.perf_counter() - t0))
        same = all(np.array_equal(out_b[p], out_a[p]) for p in out_a)
        res4["ipc_peer_operands_over_nvlink"] = {"ms_per_window_steady": best_p * 1e3,
                                                 "first_window_incl_publish_ms": first * 1e3,
                                                 "tflops": flop / best_p / 1e12,
                                                 "equal_to_nccl_form": bool(reduce_min(1.0 if same else 0.0) > 0.5)}
        w.close()
        # self-check against the single-process window on rank 0's GPU (every frame regenerated there)
        check = 1.0
        if rank == 0:
            allf = [ctx.upload(synth.sift_like(ROWS, 4000 + f)) for f in range(F)]
            ref = ctx.matchWindow(allf, MatcherType.SIFT_BF, RATIO)
            check = 1.0 if counts_a == [len(ref[p]) for p in ws.window_pairs(F)] and \
                all(np.array_equal(out_a[p], ref[p]) for p in out_a) else 0.0
            for f in allf:
                f.free()
        res4["sharded_equals_single_gpu_window"] = bool(reduce_min(check) > 0.5)
        res4["ranks_that_ran"] = world
    out["cfg4_window_8x50k"] = res4

    # ---- one process, all GPUs: the device set (rank 0 drives every GPU; the other ranks wait on the
    # HOST, their GPUs idle) ------------------------------------------------------------------------
    barrier()
    if rank == 0:
        try:
            from slam_indoor_code_b200.device_set import DeviceSet
            import synth_inputs as synth2
            n_dev = max(1, min(world, torch.cuda.device_count()))
            with DeviceSet(n_dev) as ds:
                full_q = q
                full_trains = [synth2.sift_train_from_query(full_q, N_ROWS, 3001 + p) for p in range(N_PAIRS)] \
                    if world > 1 else trains
                Qs = ds.upload(full_q)
                Tss = [ds.upload(t, ds.owner(i, N_PAIRS)) for i, t in enumerate(full_trains)]
                # the caller's result buffers, allocated once (a fresh 34 MB array per call would be
                # timed as page faults of the allocator, not as the library)
                from slam_indoor_code_b200._capi import DMATCH as _DM
                ds_out = np.zeros((N_PAIRS, N_ROWS), _DM)
                ds_n = np.zeros(N_PAIRS, np.int32)
                for _ in range(3):
                    ds.matchBatchEnqueue(Qs, Tss, MatcherType.SIFT_BF, RATIO)
                    got, n_out, dms = ds.batchFetch(ds_out, ds_n)
                best_dev, best_host = 1e9, 1e9
                for _ in range(8):
                    t0 = time.perf_counter()
                    ds.matchBatchEnqueue(Qs, Tss, MatcherType.SIFT_BF, RATIO)
                    got, n_out, dms = ds.batchFetch(ds_out, ds_n)
                    best_host = min(best_host, time.perf_counter() - t0)
                    best_dev = min(best_dev, float(dms.max()))
                out["inprocess_device_set"] = {
                    "devices": n_dev, "pairs_per_s_device_time_max_over_devices": N_PAIRS / (best_dev / 1e3),
                    "ms_per_window_device": best_dev,
                    "pairs_per_s_host_call_incl_result_copies": N_PAIRS / best_host,
                    "ms_per_window_host_call": best_host * 1e3,
                    "good_matches_total": int(n_out.sum()),
                    "note": "slamb200_set_*: one process drives every GPU (query replicated by peer copies of the "
                            "prepared set, trains resident on their owners, one enqueue per device, results gathered once)"}
                for h in Tss + [Qs]:
                    h.free()
        except Exception as e:  # pragma: no cover
            out["inprocess_device_set"] = {"error": repr(e)}
    if cpu_group is not None:
        dist.barrier(group=cpu_group)
    barrier()
    return out


def _window_single(ctx, frames, F):
    from slam_indoor_code_b200.feature_matching import MatcherType
    sets = [ctx.upload(frames[f]) for f in range(F)]
    res = ctx.matchWindow(sets, MatcherType.SIFT_BF, RATIO)
    for s_ in sets:
        s_.free()
    return res, [len(res[p]) for p in sorted(res)]

# dram__bytes_read.sum + dram__bytes_write.sum of one tcgen05-kernel launch over the 210-pair window
# (N=1), from the committed `ncu --set full` capture profiles/r02/r02_ncu_final_sift_tc_tail_compact.txt:
# 616.95 MB read + 93.59 MB written.  Compulsory bytes of that launch: 211 descriptor sets x 2.9 MB
# of bf16 operands (two 1.28 MB k-blocks + augmentation per set) + 103 MB of slot records
# (one per share and row) = 0.71 GB, i.e. no re-reads.
TRAFFIC_BYTES_PER_LAUNCH_N1 = 710_546_688


def cpu_baseline(q, trains, gpu_matches, seconds):
    """Rank 0, N=1: the OpenCV CPU path on a bounded sample of the same batch, result-checked."""
    cores, kind, what = cpu_threads()
    t0 = time.perf_counter()
    n0 = cpu_match_pair(q, trains[0])
    one = time.perf_counter() - t0
    n = int(max(1, min(len(trains) - 1, seconds / max(one, 1e-3))))
    t0 = time.perf_counter()
    agree = n0 == len(gpu_matches[0])
    for i in range(1, n + 1):
        agree = agree and (cpu_match_pair(q, trains[i]) == len(gpu_matches[i]))
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "cores": cores, "kind": kind,
            "sample": f"pairs 1..{n} of the {N_PAIRS} ({what}); good-match counts equal the GPU's: {agree}",
            "tflops": n / dt * FLOP_PER_PAIR / 1e12}


def extras(ctx, stream):
    """Short device-resident probes of the other BASELINE.json configs (not the headline)."""
    import torch
    import synth_inputs as synth
    from slam_indoor_code_b200 import camera_translation as ct
    from slam_indoor_code_b200.feature_matching import MatcherType
    out = {}

    def ev_time(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    q, t = synth.sift_pair(N_ROWS, N_ROWS, 1001)
    Q, T = ctx.upload(q), ctx.upload(t)
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, [T], MatcherType.SIFT_BF, RATIO, stream), 50)
    out["cfg1_sift_single_pair"] = {"us_per_pair": ms * 1e3, "pairs_per_s": 1e3 / ms,
                                    "tflops": FLOP_PER_PAIR / ms / 1e9}
    # cfg1 (ii): float descriptors that are not integer valued -> two-term bf16 split + certified rerank.
    # RootSIFT (sqrt of L1-normalised SIFT) is the realistic case, uniform noise the worst case.
    qi, ti = synth.sift_pair(N_ROWS, N_ROWS, 1002)
    root = lambda d: np.sqrt(d / np.maximum(d.sum(1, keepdims=True), 1)).astype(np.float32)
    for name, (q, t) in (("rootsift", (root(qi), root(ti))), ("uniform_noise", synth.float_pair(N_ROWS, N_ROWS, 1002))):
        Qf, Tf = ctx.upload(q), ctx.upload(t)
        ms = ev_time(lambda: ctx.matchBatchEnqueue(Qf, [Tf] * 16, MatcherType.SIFT_BF, RATIO, stream), 10)
        out["cfg1_general_float_16_pairs_" + name] = {"us_per_pair": ms / 16 * 1e3, "pairs_per_s": 16e3 / ms,
                                                      "tflops": 16 * FLOP_PER_PAIR / ms / 1e9}
        Qf.free()
        Tf.free()
    # next row 8f-4: NORM_L1 (useFM-SIFT-BF of the reference's OpenCV-CUDA build) on cfg1's integer rows:
    # Q*T*32 byte-wise SAD instructions per pair
    Qi, Ti = ctx.upload(qi), ctx.upload(ti)
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Qi, [Ti] * 16, MatcherType.SIFT_BF_L1, RATIO, stream), 5)
    out["f4_sift_l1_16_pairs"] = {"us_per_pair": ms / 16 * 1e3, "pairs_per_s": 16e3 / ms,
                                  "tsad4_per_s": 16 * 32.0 * N_ROWS * N_ROWS / ms / 1e9}
    Qi.free()
    Ti.free()
    q, t = synth.orb_pair(N_ROWS, N_ROWS, 2001)
    Q, T = ctx.upload(q), ctx.upload(t)
    # cfg2 through the XOR/POPC kernel (bound: the POPC pipe) ...
    ctx.debug_orb_kernel(tensor_cores=False)
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, [T] * 16, MatcherType.ORB_BF, RATIO, stream), 10)
    out["cfg2_orb_16_pairs"] = {"kernel": "orb_knn2_kernel (XOR + carry-save + POPC)",
                                "us_per_pair": ms / 16 * 1e3, "pairs_per_s": 16e3 / ms,
                                "tpopc_per_s": 16 * 8e8 / ms / 1e9}
    # ... and through the tcgen05 kernel on the bits spread to e4m3 0/1 bytes (the default route):
    # Hamming = squared L2 of the bit vectors, 2*Q*T*256 fp8 flop per pair on the tensor pipe
    ctx.debug_orb_kernel(tensor_cores=True)
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, [T] * 16, MatcherType.ORB_BF, RATIO, stream), 20)
    out["cfg2_orb_16_pairs_tcgen05"] = {"kernel": "sift_tc_kernel, kind::f8f6f4 (default ORB route)",
                                        "us_per_pair": ms / 16 * 1e3, "pairs_per_s": 16e3 / ms,
                                        "fp8_tflops": 16 * 2.0 * N_ROWS * N_ROWS * 256 / ms / 1e9,
                                        "popc_equivalent_T_per_s": 16 * 8e8 / ms / 1e9}
    ms = ev_time(lambda: ctx.matchBatchEnqueue(Q, [T], MatcherType.ORB_BF, RATIO, stream), 50)
    out["cfg2_orb_single_pair_tcgen05"] = {"us_per_pair": ms * 1e3, "pairs_per_s": 1e3 / ms}
    # cfg5: 2048 hypotheses x 5000 matches, 32 pairs per launch, host-call timing incl. copies
    p1, p2, R, tv = synth.two_view(5000, 5000)
    E = synth.pose_hypotheses(2048, R, tv, 5001)
    P = 32
    ct.scoreEssentialBatch(ctx, [p1] * P, [p2] * P, synth.SAMSUNG_HV_4K, np.stack([E] * P), 5.0)  # warm-up
    ctx.profile_enable(True)
    ctx.profile_read()
    for _ in range(3):
        ct.scoreEssentialBatch(ctx, [p1] * P, [p2] * P, synth.SAMSUNG_HV_4K, np.stack([E] * P), 5.0)
    kms, kn = ctx.profile_read()["ransac"]
    ctx.profile_enable(False)
    out["cfg5_ransac_2048x5000"] = {"kernel_us_per_pair": kms / kn / P * 1e3,
                                    "fp64_tflops_40flop_convention": P * 2048 * 5000 * 40 / (kms / kn) / 1e9,
                                    "dp_instr_per_s_T": P * 2048 * 5000 * 36 / (kms / kn) / 1e9}
    # next row 8f-2: solvePnPRansac scoring, 2048 poses x 5000 correspondences, 32 frames per launch,
    # the reference's five distortion coefficients
    from slam_indoor_code_b200 import pnp_ransac as pr
    obj, img, Rp, tp = synth.pnp_scene(5000, 7000)
    poses = synth.pnp_hypotheses(2048, Rp, tp, 7001)
    pr.scorePnPBatch(ctx, [obj] * P, [img] * P, synth.SAMSUNG_HV_4K, synth.REF_DIST5, np.stack([poses] * P), 8.0)
    ctx.profile_enable(True)
    ctx.profile_read()
    for _ in range(3):
        pr.scorePnPBatch(ctx, [obj] * P, [img] * P, synth.SAMSUNG_HV_4K, synth.REF_DIST5, np.stack([poses] * P), 8.0)
    kms, kn = ctx.profile_read()["pnp"]
    ctx.profile_enable(False)
    # 53 explicit fp64 operations per (pose, point) with five coefficients (the reciprocal counted once)
    out["f2_pnp_2048x5000"] = {"kernel_us_per_frame": kms / kn / P * 1e3,
                               "fp64_ops_T_per_s": P * 2048 * 5000 * 53 / (kms / kn) / 1e9}
    # next row 8f-3 (ORB half): descriptors of 12 000 FAST-like keypoints on a 4K BGR frame, resident
    # output (no descriptor upload); host call incl. the 24.9 MB frame copy, and the kernels alone
    from slam_indoor_code_b200 import orb_descriptors as od
    frame4k = synth.textured_frame(2160, 3840, 6000, 3)
    rngk = np.random.default_rng(6001)
    kp4k = np.stack([rngk.integers(31, 3840 - 31, 12000), rngk.integers(31, 2160 - 31, 12000),
                     np.full(12000, -1.0)], 1).astype(np.float32)
    for _ in range(6):   # every worker lane of the context has its frame buffers (calls rotate over five lanes)
        od.extractDescriptorORB(ctx, frame4k, kp4k, want_host=False, want_resident=True)[2].free()
    ctx.profile_enable(True)
    ctx.profile_read()
    t0 = time.perf_counter()
    for _ in range(5):
        od.extractDescriptorORB(ctx, frame4k, kp4k, want_host=False, want_resident=True)[2].free()
    dt = (time.perf_counter() - t0) / 5
    kms, kn = ctx.profile_read()["orb_desc"]
    ctx.profile_enable(False)
    out["f3_orb_compute_4k_12000kp"] = {"ms_per_host_call": dt * 1e3, "kernels_us": kms / max(kn, 1) * 1e3}
    # next row 8f-3, the detector in front of it: fastExtractor (FAST-9/16, threshold 10, suppression)
    # on the same 4K BGR frame; the host call includes the 24.9 MB pageable frame copy and the
    # keypoint read-back
    from slam_indoor_code_b200 import fast_extractor as fe
    n_fast = len(fe.fastExtractor(ctx, frame4k, 10, True))
    for _ in range(5):
        fe.fastExtractor(ctx, frame4k, 10, True, max_points=n_fast)
    ctx.profile_enable(True)
    ctx.profile_read()
    t0 = time.perf_counter()
    for _ in range(5):
        fe.fastExtractor(ctx, frame4k, 10, True, max_points=n_fast)
    dt = (time.perf_counter() - t0) / 5
    kms, kn = ctx.profile_read()["fast"]
    ctx.profile_enable(False)
    out["f3_fast_4k"] = {"ms_per_host_call": dt * 1e3, "kernels_us": kms / max(kn, 1) * 1e3, "keypoints": n_fast}
    # next row 8f-3, second half: SIFT descriptors of 12 000 FAST-like keypoints (size 7, angle -1) on the
    # same 4K frame, resident output; tolerance-pinned against cv2 (tests/test_gpu_sift_descriptors.py)
    from slam_indoor_code_b200 import sift_descriptors as sdm
    kp4s = np.concatenate([kp4k[:, :2], np.full((12000, 1), 7.0, np.float32), np.full((12000, 1), -1.0, np.float32)], 1)
    for _ in range(6):
        sdm.extractDescriptorSIFT(ctx, frame4k, kp4s, want_host=False, want_resident=True)[1].free()
    ctx.profile_enable(True)
    ctx.profile_read()
    t0 = time.perf_counter()
    for _ in range(5):
        sdm.extractDescriptorSIFT(ctx, frame4k, kp4s, want_host=False, want_resident=True)[1].free()
    dt = (time.perf_counter() - t0) / 5
    kms, kn = ctx.profile_read()["sift_desc"]
    ctx.profile_enable(False)
    out["f3_sift_compute_4k_12000kp"] = {"ms_per_host_call": dt * 1e3, "kernels_us": kms / max(kn, 1) * 1e3}
    # next row 8f-4: linear triangulation of 5000 matches (one host call incl. copies)
    from slam_indoor_code_b200 import triangulation as tri
    K4 = synth.SAMSUNG_HV_4K
    Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    tp1, tp2, tR, tt = synth.two_view(5000, 8000, outliers=0.0)
    for _ in range(3):
        tri.reconstruct(ctx, Km, np.eye(3), np.zeros(3), tR, tt, tp1, tp2)
    t0 = time.perf_counter()
    for _ in range(20):
        tri.reconstruct(ctx, Km, np.eye(3), np.zeros(3), tR, tt, tp1, tp2)
    out["f4_triangulate_5000"] = {"us_per_host_call": (time.perf_counter() - t0) / 20 * 1e6}
    # pipe-rate denominators measured on this box (csrc/microbench.cu)
    try:
        import ctypes
        lib = ctx._lib
        lib.slamb200_dbg_pipe_rate.restype = ctypes.c_double
        lib.slamb200_dbg_pipe_rate.argtypes = [ctypes.c_int]
        popc, fp64 = lib.slamb200_dbg_pipe_rate(0), lib.slamb200_dbg_pipe_rate(1)
        sad4 = lib.slamb200_dbg_pipe_rate(2)
        out["pipe_rates"] = {"popc_G_per_s": popc, "fp64_dmul_dadd_G_per_s": fp64, "vabsdiff4_G_per_s": sad4}
        out["f4_sift_l1_16_pairs"]["frac_of_vabsdiff4_pipe"] = out["f4_sift_l1_16_pairs"]["tsad4_per_s"] * 1e3 / sad4
        out["cfg2_orb_16_pairs"]["frac_of_popc_pipe"] = out["cfg2_orb_16_pairs"]["tpopc_per_s"] * 1e3 / popc
        out["cfg5_ransac_2048x5000"]["frac_of_fp64_pipe"] = \
            out["cfg5_ransac_2048x5000"]["dp_instr_per_s_T"] * 1e3 / fp64
        out["f2_pnp_2048x5000"]["frac_of_fp64_pipe_lower_bound"] = \
            out["f2_pnp_2048x5000"]["fp64_ops_T_per_s"] * 1e3 / fp64
    except Exception as e:  # pragma: no cover
        out["pipe_rates"] = {"error": str(e)}
    return out


def cpu_extras():
    """CPU OpenCV timings of the other configs, bounded to a few seconds."""
    out = {}
    try:
        import cv2
    except Exception:
        return {"unavailable": "cv2 not importable"}
    import synth_inputs as synth
    cv2.setNumThreads(os.cpu_count() or 1)
    out["cores"] = int(cv2.getNumThreads())

    def best(fn, n=3):
        b = 1e9
        for _ in range(n):
            t0 = time.perf_counter()
            r = fn()
            b = min(b, time.perf_counter() - t0)
        return b, r

    q, t = synth.orb_pair(N_ROWS, N_ROWS, 2001)
    dt, _ = best(lambda: cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2))
    out["cfg2_orb_bf_hamming_s_per_pair"] = dt
    q, t = synth.sift_pair(N_ROWS, N_ROWS, 1001)
    dt_bf, bf = best(lambda: cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, 2), 2)
    dt_fl, fl = best(lambda: cv2.FlannBasedMatcher().knnMatch(q, t, 2), 2)
    top1 = np.mean([a[0].trainIdx == b[0].trainIdx for a, b in zip(bf, fl)])
    good_bf = {(m[0].queryIdx, m[0].trainIdx) for m in bf if m[0].distance < RATIO * m[1].distance}
    good_fl = {(m[0].queryIdx, m[0].trainIdx) for m in fl if m[0].distance < RATIO * m[1].distance}
    out["cfg1_sift_bf_s_per_pair"] = dt_bf
    out["flann"] = {"s_per_pair": dt_fl, "top1_recall_vs_bf": float(top1),
                    "accepted_match_recall_vs_bf": len(good_bf & good_fl) / max(len(good_bf), 1),
                    "note": "the reference's useFM-SIFT-FLANN path (approximate); this repo answers "
                            "that flag with the exact search, recall 1.0 by construction"}
    p1, p2, _, _ = synth.two_view(5000, 5000)
    K4 = synth.SAMSUNG_HV_4K
    Kmat = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    dt, _ = best(lambda: cv2.findEssentialMat(p1, p2, Kmat, cv2.RANSAC, 0.999, 5.0), 5)
    out["cfg5_findEssentialMat_s_per_pair"] = dt
    # the "next" rows: NORM_L1 matcher, solvePnPRansac, triangulation
    dt, _ = best(lambda: cv2.BFMatcher(cv2.NORM_L1).knnMatch(q, t, 2), 2)
    out["f4_sift_bf_l1_s_per_pair"] = dt
    obj, img, _, _ = synth.pnp_scene(5000, 7000)
    dist = np.array(synth.REF_DIST5)
    dt, _ = best(lambda: cv2.solvePnPRansac(obj, img, Kmat, dist), 5)
    out["f2_solvePnPRansac_s_per_frame"] = dt
    frame4k = synth.textured_frame(2160, 3840, 6000, 3)
    rngk = np.random.default_rng(6001)
    cvk = [cv2.KeyPoint(float(x), float(y), 7.0, -1.0, 0.0, 0) for x, y in
           zip(rngk.integers(31, 3840 - 31, 12000), rngk.integers(31, 2160 - 31, 12000))]
    orb = cv2.ORB_create()
    dt, _ = best(lambda: orb.compute(frame4k, cvk), 3)
    out["f3_orb_compute_4k_12000kp_s"] = dt
    sift = cv2.SIFT_create()
    dt, _ = best(lambda: sift.compute(frame4k, cvk), 3)
    out["f3_sift_compute_4k_12000kp_s"] = dt
    fast = cv2.FastFeatureDetector_create(10, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    dt, _ = best(lambda: fast.detect(frame4k), 3)
    out["f3_fast_4k_s"] = dt
    tp1, tp2, tR, tt = synth.two_view(5000, 8000, outliers=0.0)
    P1 = Kmat @ np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = Kmat @ np.hstack([tR, tt.reshape(3, 1)])
    a, b = tp1.T.astype(np.float64), tp2.T.astype(np.float64)
    dt, _ = best(lambda: cv2.triangulatePoints(P1, P2, a, b), 5)
    out["f4_triangulatePoints_5000_s"] = dt
    return out


def main():
    # Exactly one JSON line may reach stdout: route fd 1 to stderr for everything else (NCCL prints
    # its version banner to stdout when NCCL_DEBUG=VERSION is set on a box) and keep the real
    # stdout for the final line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--events-all", action="store_true",
                    help="developer A/B: events around every kernel class in the timed region too")
    ap.add_argument("--e2e-uploaders", type=int, default=-1,
                    help="narrowing threads inside slamb200_match_batch_host (-1: pack threads + 1)")
    ap.add_argument("--e2e-pack-threads", type=int, default=-1,
                    help="library threads sharing the narrowing of each Mat (-1: min(cores, 16) - uploaders)")
    ap.add_argument("--e2e-upload", default="auto", choices=["auto", "packed", "pinned"],
                    help="packed (default): pageable host Mats, rows narrowed to bytes on the host threads (verified "
                         "lossless) before PCIe; pinned: allocates the Mats page-locked (experiment)")
    ap.add_argument("--ref-pairs", type=int, default=8, help="pairs per step of the CPU arm (a bounded sample "
                    "of the 210-pair window: ~1 s per step on 16 cores)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        log("warmup < 3 requested; the timing rules ask for >= 3")
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        log(f"--gpus {args.gpus} needs torchrun (one rank per GPU); running this process as 1 rank")
    run_b200(args, rank, world, local)


if __name__ == "__main__":
    main()
