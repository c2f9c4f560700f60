// fast_core.h -- per-pixel arithmetic of FAST-9/16 (corner test and corner score), shared by the
// CUDA kernels (fast_detect.cu) and a host-compiled unit check (tests/test_fast_core_cpu.py).
//
// Restates cv::FAST_t<16> / cornerScore<16> as FastFeatureDetector::create(threshold, suppression,
// TYPE_9_16)->detect runs them for the reference's fastExtractor
// (src/mainModule/featureExtraction/fastExtractor.cpp:7-13): pixel v is a corner iff at least 9
// contiguous pixels of the 16-pixel circle of radius 3 are all < v - t or all > v + t; its score
// is the largest t for which it still is one.  The score below is the closed form of OpenCV's
// incremental min/max loop: max over the 16 arcs of 9 of the arc's minimum (darker side) or of
// minus its maximum (brighter side), minus 1 -- computed here with a running window instead of
// OpenCV's pairwise loop, same integers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FAST_HD __host__ __device__ __forceinline__
#else
#define FAST_HD inline
#endif

// circle offsets (dx, dy), OpenCV's order (fast_score.cpp makeOffsets, patternSize 16)
#define FAST_CIRCLE_INIT                                                                         \
  {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, \
   {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}}

// d[k] = v - circle[k] for the 16 circle pixels.  Returns -1 when the pixel is not a corner at
// threshold t (clamped to 0..255 by the caller), else its score (>= t).
FAST_HD int fast9_score(const int (&d)[16], int t) {
  // longest runs of d > t (darker circle pixels) and d < -t (brighter), circularly
  int dark = 0, bright = 0, is_corner = 0;
#pragma unroll
  for (int k = 0; k < 25; k++) {
    const int x = d[k & 15];
    dark = x > t ? dark + 1 : 0;
    bright = x < -t ? bright + 1 : 0;
    is_corner |= (dark > 8) | (bright > 8);
  }
  if (!is_corner) return -1;
  // score = max(t, max over arcs of min(d), max over arcs of min(-d)) over the 16 arcs of 9 ... -1
  // OpenCV: a0 = max(t, arc minima of d); b0 = min(-a0, arc maxima of d); score = -b0 - 1
  int a0 = t;
#pragma unroll
  for (int s = 0; s < 16; s++) {
    int m = d[s];
#pragma unroll
    for (int j = 1; j < 9; j++) m = m < d[(s + j) & 15] ? m : d[(s + j) & 15];
    a0 = a0 > m ? a0 : m;
  }
  int b0 = -a0;
#pragma unroll
  for (int s = 0; s < 16; s++) {
    int m = d[s];
#pragma unroll
    for (int j = 1; j < 9; j++) m = m > d[(s + j) & 15] ? m : d[(s + j) & 15];
    b0 = b0 < m ? b0 : m;
  }
  return -b0 - 1;
}
