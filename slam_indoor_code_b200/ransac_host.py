"""Host-side control of cv::findEssentialMat's RANSAC loop around the GPU scorer.

The reference's call (src/mainModule/translation/cameraTranslation.cpp:41-46) runs OpenCV's
RANSACPointSetRegistrator::run: a fixed-seed RNG draws 5-point subsets, the 5-point solver turns
each into <= 10 candidate essential matrices, every candidate is scored against all matches, a
candidate replaces the best iff count > max(best, 4), and the iteration budget shrinks with the
best inlier ratio.  Only the scoring is data-parallel; it runs on the B200
(slamb200_score_essential).  This module restates the *control*: OpenCV's RNG (multiply-with-carry,
seed 2^64-1), getSubset, RANSACUpdateNumIters and the update rule, so that the result (E and the
N x 1 mask) is bit-identical to cv::findEssentialMat -- verified against cv2 in the tests.

Iterations are scored speculatively in chunks (the loop is sequential only through `niters`):
hypotheses of `chunk` iterations go to the GPU in one call and the update rule is replayed on the
host in order; work past the final `niters` is discarded.
"""
import math
import sys

import numpy as np


class CvRNG:
    """cv::RNG: state = (uint32)state * 4164903690 + (state >> 32)."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else self.next() % (b - a) + a


def get_subset(rng, count, model_points=5):
    """RANSACPointSetRegistrator::getSubset: distinct indices, redrawn on collision."""
    idx = []
    for _ in range(model_points):
        v = rng.uniform(0, count)
        while v in idx:
            v = rng.uniform(0, count)
        idx.append(v)
    return idx


def update_num_iters(p, ep, model_points, max_iters):
    """cv::RANSACUpdateNumIters."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, sys.float_info.min)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < sys.float_info.min:
        return 0
    num, denom = math.log(num), math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))  # cvRound: to nearest, ties to even


def ransac_run(count, model_points, prob, max_iters, solve_fn, score_fn, chunk=32):
    """RANSACPointSetRegistrator::run for count > model_points: returns (best model or None,
    iterations run, hypotheses scored).

    solve_fn(idx list[model_points]) -> array [k, D] of candidate models (k may be 0);
    score_fn(models [H, D]) -> counts [H]."""
    rng = CvRNG()
    niters, it, scored = max_iters, 0, 0
    best, max_good = None, 0
    while it < niters:
        # speculate: hypotheses of the next `chunk` iterations (never past the current budget)
        n_it = min(chunk, niters - it)
        models, owner = [], []
        for j in range(n_it):
            idx = get_subset(rng, count, model_points)
            cand = np.asarray(solve_fn(idx), np.float64)
            for m in cand.reshape(len(cand), -1) if cand.size else ():
                models.append(m)
                owner.append(it + j)
        if models:
            counts = score_fn(np.array(models))
            scored += len(models)
            for h, c in enumerate(counts):
                if owner[h] >= niters:      # the budget shrank below this iteration: discard
                    break
                if c > max(max_good, model_points - 1):
                    max_good, best = int(c), models[h].copy()
                    niters = update_num_iters(prob, (count - max_good) / count, model_points, niters)
        it += n_it
    # the RNG drew subsets for speculated iterations past the final budget; they are discarded,
    # and cv's loop ends at the same model because the update rule was replayed in order
    return best, (min(it, niters) if best is not None else it), scored


def ransac_essential(points1, points2, K4, prob, threshold, five_point_fn, score_fn, max_iters=1000,
                     chunk=32):
    """Returns (E [9] or None, mask [N] uint8 or None, iterations run, hypotheses scored).

    five_point_fn(p1[5,2], p2[5,2]) -> array [k, 9] of candidate models (k may be 0);
    score_fn(E [H,9]) -> counts [H] -- see camera_translation."""
    p1 = np.ascontiguousarray(points1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(points2, np.float32).reshape(-1, 2)
    count = p1.shape[0]
    if count < 5:
        return None, None, 0, 0
    if count == 5:  # run(): exactly the minimal set -> the solver's models, mask of ones
        E = five_point_fn(p1, p2)
        return (E.reshape(-1) if len(E) else None), np.ones(count, np.uint8), 0, 0
    best, iters, scored = ransac_run(count, 5, prob, max_iters,
                                     lambda idx: five_point_fn(p1[idx], p2[idx]), score_fn, chunk)
    return best, None, iters, scored
