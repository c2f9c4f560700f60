// fast_detect.cu -- FAST-9/16 keypoints of a frame on the device.
//
// Replaces FastFeatureDetector::create(threshold, suppression, TYPE_9_16)->detect(frame, points)
// as the reference's fastExtractor calls it (src/mainModule/featureExtraction/fastExtractor.cpp:
// 7-13; callers cycleProcessing/batch.cpp:245 and mainCycleInternals.cpp:144 with
// featureExtractingThreshold).  Same keypoints as OpenCV, in OpenCV's order (row by row, left to
// right), with the same response:
//   gray     cvtColor(BGR2GRAY): (B*3735 + G*19235 + R*9798 + 2^14) >> 15        (orb_desc.cu)
//   score    fast_core.h per pixel, 3 px away from every border; -1 = not a corner
//   keep     without suppression every corner; with it, corners whose score is strictly greater
//            than the scores of all 8 neighbours (non-corners count as 0)
//   order    blocks of 256 pixels of one row, counted, prefix-summed, written in place
// Byte-wide integer work, bound by the reads of the gray image (16 neighbours per pixel, served by
// L1/L2: the frame is read from HBM once).
#include "common.cuh"
#include "fast_core.h"

#define FAST_THREADS 256

__global__ void __launch_bounds__(FAST_THREADS)
fast_score_kernel(const uint8_t* __restrict__ gray, int rows, int cols, int t, int16_t* __restrict__ score) {
  const int x = blockIdx.x * FAST_THREADS + threadIdx.x, y = blockIdx.y;
  if (x >= cols) return;
  int s = -1;
  if (y >= 3 && y < rows - 3 && x >= 3 && x < cols - 3) {
    constexpr int C[16][2] = FAST_CIRCLE_INIT;
    const uint8_t* p = gray + (size_t)y * cols + x;
    const int v = p[0];
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - (int)p[C[k][1] * cols + C[k][0]];
    s = fast9_score(d, t);
  }
  score[(size_t)y * cols + x] = (int16_t)s;
}

// response of pixel (x, y) if it is a keypoint, else -1
__device__ __forceinline__ int fast_keep(const int16_t* __restrict__ score, int cols, int x, int y, int nonmax) {
  const int16_t* p = score + (size_t)y * cols + x;
  const int s = p[0];
  if (s < 0) return -1;
  if (!nonmax) return 0;   // OpenCV reports response 0 without suppression
  // corners are at least 3 px inside the image: all eight neighbours exist
  int m = max((int)p[-1], (int)p[1]);
  m = max(m, max(max((int)p[-cols - 1], (int)p[-cols]), (int)p[-cols + 1]));
  m = max(m, max(max((int)p[cols - 1], (int)p[cols]), (int)p[cols + 1]));
  m = max(m, 0);           // a non-corner neighbour scores 0
  return s > m ? s : -1;
}

__global__ void __launch_bounds__(FAST_THREADS)
fast_count_kernel(const int16_t* __restrict__ score, int rows, int cols, int nonmax, int32_t* __restrict__ cnt) {
  const int x = blockIdx.x * FAST_THREADS + threadIdx.x, y = blockIdx.y;
  const int keep = x < cols && fast_keep(score, cols, x, y, nonmax) >= 0;
  const int n = __syncthreads_count(keep);
  if (threadIdx.x == 0) cnt[y * gridDim.x + blockIdx.x] = n;
}

// exclusive prefix sum of cnt[0..n) in place, cnt[n] = total; one block
__global__ void __launch_bounds__(1024) fast_scan_kernel(int32_t* __restrict__ cnt, int n) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += o;
      }
      warp_sum[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int before = carry + (warp > 0 ? warp_sum[warp - 1] : 0) + incl - v;
    if (i < n) cnt[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt[n] = carry_s;
}

__global__ void __launch_bounds__(FAST_THREADS)
fast_write_kernel(const int16_t* __restrict__ score, int rows, int cols, int nonmax,
                  const int32_t* __restrict__ offs, float* __restrict__ kp, int cap) {
  __shared__ int warp_sum[FAST_THREADS / 32];
  const int x = blockIdx.x * FAST_THREADS + threadIdx.x, y = blockIdx.y;
  const int resp = x < cols ? fast_keep(score, cols, x, y, nonmax) : -1;
  const int keep = resp >= 0;
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sum[warp] = __popc(bal);
  __syncthreads();
  if (!keep) return;
  int at = offs[y * gridDim.x + blockIdx.x];
  for (int w = 0; w < warp; w++) at += warp_sum[w];
  at += __popc(bal & ((1u << lane) - 1));
  if (at < cap) {
    kp[3 * (size_t)at] = (float)x;
    kp[3 * (size_t)at + 1] = (float)y;
    kp[3 * (size_t)at + 2] = (float)resp;
  }
}

int fast_blocks(int rows, int cols) { return rows * ((cols + FAST_THREADS - 1) / FAST_THREADS); }

// gray: rows x cols bytes.  cnt: fast_blocks() + 1 ints (cnt[last] = number of keypoints found).
void launch_fast_detect(const uint8_t* gray, int rows, int cols, int threshold, int nonmax,
                        int16_t* score, int32_t* cnt, float* kp, int cap, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return;
  threshold = threshold < 0 ? 0 : threshold > 255 ? 255 : threshold;
  const dim3 grid((cols + FAST_THREADS - 1) / FAST_THREADS, rows);
  fast_score_kernel<<<grid, FAST_THREADS, 0, s>>>(gray, rows, cols, threshold, score);
  COUNT_LAUNCH();
  fast_count_kernel<<<grid, FAST_THREADS, 0, s>>>(score, rows, cols, nonmax, cnt);
  COUNT_LAUNCH();
  fast_scan_kernel<<<1, 1024, 0, s>>>(cnt, fast_blocks(rows, cols));
  COUNT_LAUNCH();
  fast_write_kernel<<<grid, FAST_THREADS, 0, s>>>(score, rows, cols, nonmax, cnt, kp, cap);
  COUNT_LAUNCH();
}
