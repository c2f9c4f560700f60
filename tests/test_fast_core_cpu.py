"""CPU: the per-pixel FAST-9/16 arithmetic the CUDA kernels use (csrc/fast_core.h, compiled here for
the host by g++) against the oracle's restatement of cv::FAST_t<16> / cornerScore<16>
(oracle_fast_detect, itself pinned to cv2.FastFeatureDetector): the corner verdict and the score of
EVERY pixel of textured and random frames, thresholds from 0 to 255 -- no GPU involved.
Reference: src/mainModule/featureExtraction/fastExtractor.cpp:7-13."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import synth_inputs as synth
from oracle import c_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <stddef.h>
#include <stdint.h>
#include "fast_core.h"
// score map of a gray frame: 0 where the pixel is not a corner (or within 3 px of a border), else
// its score -- what oracle_fast_detect returns in score_out; corner[] is the verdict itself (a
// corner at threshold 0 can have score 0)
extern "C" void fast_core_map(const uint8_t* gray, int rows, int cols, int t, uint8_t* score, uint8_t* corner) {
  static const int circle[16][2] = FAST_CIRCLE_INIT;
  for (int y = 3; y < rows - 3; y++)
    for (int x = 3; x < cols - 3; x++) {
      int d[16];
      const int v = gray[(size_t)y * cols + x];
      for (int k = 0; k < 16; k++) d[k] = v - gray[(size_t)(y + circle[k][1]) * cols + x + circle[k][0]];
      const int s = fast9_score(d, t);
      corner[(size_t)y * cols + x] = s >= 0;
      score[(size_t)y * cols + x] = s >= 0 ? (uint8_t)s : 0;
    }
}
'''


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    d = tmp_path_factory.mktemp("fast_core")
    src = d / "fast_core_host.cpp"
    src.write_text(SRC)
    so = d / "libfast_core_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I",
                    os.path.join(ROOT, "slam_indoor_code_b200", "csrc"), str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.fast_core_map.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p]
    lib.fast_core_map.restype = None
    return lib


def _maps(core, gray, t):
    score = np.zeros(gray.shape, np.uint8)
    corner = np.zeros(gray.shape, np.uint8)
    core.fast_core_map(gray.ctypes.data, gray.shape[0], gray.shape[1], t, score.ctypes.data, corner.ctypes.data)
    return score, corner


@pytest.mark.parametrize("threshold", [0, 1, 10, 40, 120, 254, 255])
def test_score_map_equals_oracle(core, threshold):
    rng = np.random.default_rng(700 + threshold)
    frames = [np.ascontiguousarray(synth.textured_frame(90, 130, 701, 1).reshape(90, 130)),
              rng.integers(0, 256, (64, 77), dtype=np.uint8),                     # noise: corners everywhere
              (rng.integers(0, 2, (50, 60)) * 255).astype(np.uint8),             # extremes: differences of +-255
              np.full((20, 20), 128, np.uint8)]                                   # flat: none
    for gray in frames:
        score, corner = _maps(core, gray, threshold)
        kp, ref_score = c_oracle.fast_detect(gray, threshold, False, want_scores=True)
        assert np.array_equal(score, ref_score)
        # the verdicts: the oracle lists every corner (no suppression), row by row
        ys, xs = np.nonzero(corner)
        assert np.array_equal(np.stack([xs, ys], 1).astype(np.float32), kp[:, :2])


def test_frames_smaller_than_the_circle(core):
    for shape in ((6, 6), (7, 6), (3, 40), (7, 7)):
        gray = np.random.default_rng(9).integers(0, 256, shape, dtype=np.uint8)
        score, corner = _maps(core, gray, 5)
        kp, ref_score = c_oracle.fast_detect(gray, 5, False, want_scores=True)
        assert np.array_equal(score, ref_score) and int(corner.sum()) == len(kp)
