"""CPU: the C++ RANSAC control (host/ransac_control.h, used by cameraTranslationB200.cpp) draws the
same subsets and takes the same decisions as the Python control, which is pinned bit-for-bit to
cv2.findEssentialMat (test_ransac_host_logic.py).  Solver and scorer are scripted."""
import ctypes
import os

import numpy as np
import pytest

from slam_indoor_code_b200 import build, ransac_host


@pytest.fixture(scope="module")
def host():
    build.build()
    lib = ctypes.CDLL(os.path.join(build.LIBDIR, "libslamb200_hostshim.so"))
    lib.hostshim_ransac_control.restype = ctypes.c_int
    lib.hostshim_ransac_control.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("seed,count,chunk", [(1, 600, 32), (2, 75, 1), (3, 5000, 7), (4, 40, 32)])
def test_cpp_control_equals_python_control(host, seed, count, chunk):
    rng = np.random.default_rng(seed)
    n_models = rng.integers(0, 11, 1200).astype(np.int32)          # 0..10 candidates per sample
    scores = rng.integers(0, max(count // 3, 6), 12000).astype(np.int32)
    # a few strong models so that the iteration budget shrinks along the way
    for pos in rng.choice(4000, 6, replace=False):
        scores[pos] = int(count * rng.uniform(0.4, 0.9))
    subsets = np.full(5 * 1200, -1, np.int32)
    best = ctypes.c_int(-2)
    iters = host.hostshim_ransac_control(count, 0.999, chunk, n_models.ctypes.data, len(n_models),
                                         scores.ctypes.data, len(scores), subsets.ctypes.data, len(subsets),
                                         ctypes.byref(best))
    # the Python control with the same script
    state = {"sample": 0, "issued": 0}
    drawn = []

    def solve(p1, p2):
        k = int(n_models[state["sample"]]) if state["sample"] < len(n_models) else 1
        state["sample"] += 1
        out = np.zeros((k, 9))
        out[:, 0] = np.arange(state["issued"], state["issued"] + k)
        state["issued"] += k
        return out

    def score(models):
        return [int(scores[int(m[0])]) if int(m[0]) < len(scores) else 0 for m in models]

    # points only serve as placeholders for the sampled rows: record the indices through them
    pts = np.arange(count, dtype=np.float32).repeat(2).reshape(count, 2)

    def solve_rec(p1, p2):
        drawn.extend(int(v) for v in p1[:, 0])
        return solve(p1, p2)

    E, _, py_iters, _ = ransac_host.ransac_essential(pts, pts, None, 0.999, 5.0, solve_rec, score, chunk=chunk)
    py_best = -1 if E is None else int(E[0])
    assert best.value == py_best
    assert iters == py_iters
    n = min(len(drawn), len(subsets))
    assert n > 0 and list(subsets[:n]) == drawn[:n]


@pytest.mark.parametrize("seed,count,mp,max_iters,chunk", [(11, 900, 5, 100, 8), (12, 30, 4, 100, 1), (13, 6, 5, 100, 8),
                                                           (14, 4000, 5, 100, 100)])
def test_cpp_control_pnp_instance(host, seed, count, mp, max_iters, chunk):
    """The solvePnPRansac instance of the control (one 12-double model per sample, maxIters 100,
    confidence 0.99, floor modelPoints-1) against ransac_host.ransac_run."""
    host.hostshim_ransac_run.restype = ctypes.c_int
    host.hostshim_ransac_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    rng = np.random.default_rng(seed)
    n_models = rng.integers(0, 2, 200).astype(np.int32)            # EPnP: one model, or a failure
    scores = rng.integers(0, max(count // 2, 6), 400).astype(np.int32)
    scores[rng.choice(60, 3, replace=False)] = int(count * 0.8)
    subsets = np.full(mp * 200, -1, np.int32)
    best = ctypes.c_int(-2)
    iters = host.hostshim_ransac_run(count, mp, 12, 0.99, max_iters, chunk, n_models.ctypes.data, len(n_models),
                                     scores.ctypes.data, len(scores), subsets.ctypes.data, len(subsets),
                                     ctypes.byref(best))
    state = {"sample": 0, "issued": 0}
    drawn = []

    def solve(idx):
        drawn.extend(idx)
        k = int(n_models[state["sample"]]) if state["sample"] < len(n_models) else 1
        state["sample"] += 1
        out = np.zeros((k, 12))
        out[:, 0] = np.arange(state["issued"], state["issued"] + k)
        state["issued"] += k
        return out

    def score(models):
        return [int(scores[int(m[0])]) if int(m[0]) < len(scores) else 0 for m in models]

    model, py_iters, _ = ransac_host.ransac_run(count, mp, 0.99, max_iters, solve, score, chunk)
    assert best.value == (-1 if model is None else int(model[0]))
    assert iters == py_iters
    n = min(len(drawn), len(subsets))
    assert n > 0 and list(subsets[:n]) == drawn[:n]
