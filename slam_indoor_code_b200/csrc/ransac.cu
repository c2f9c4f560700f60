// ransac.cu -- batched Sampson-error inlier scoring of essential-matrix hypotheses.
//
// Restates the data-parallel inside of cv::findEssentialMat(p1, p2, K, RANSAC, prob, thr, mask)
// as the reference calls it at src/mainModule/translation/cameraTranslation.cpp:41-46 (OpenCV
// calib3d: five-point.cpp EMEstimatorCallback::computeError, ptsetreg.cpp findInliers / run):
//   * points -> double, x = u*(1/fx) + (-cx*(1/fx)) (the MatExpr-folded "(col - c) / f"),
//   * thr = RPRANSACThreshold / ((fx+fy)/2), t = (float)(thr*thr),
//   * Ex1 = E*x1, Etx2 = E^T*x2, s = x2 . Ex1, all "s = 0; s += a*b" left-to-right fp64 sums,
//     err = (float)(s*s / (Ex1[0]^2 + Ex1[1]^2 + Etx2[0]^2 + Etx2[1]^2)), inlier iff err <= t,
//   * a model replaces the best iff count > max(best_count, 4): the first best wins.
// Every fp64 operation is an explicit round-to-nearest intrinsic (no FMA contraction), so counts
// and masks are bit-exact against the CPU.
//
// The fp64 division is the expensive instruction, so the inlier test is decided without it
// whenever the quotient is provably on one side of the float rounding boundary: with
// mid = midpoint(t, nextafterf(t)) (exact in fp64), (float)(num/den) <= t  <=>  RN(num/den) lies
// below mid (or on it when t's mantissa is even).  num < mid_lo*den (mid_lo = mid*(1-2^-49))
// proves "inlier", num > mid_hi*den proves "outlier"; only the 2^-48-wide sliver in between
// takes the exact division.
//
// Shape: a thread owns one hypothesis (E in registers) and walks the pair's matches, which are
// staged in shared memory as pre-normalised double4 {x1,y1,x2,y2} and read by all lanes at the
// same address (broadcast).  Blocks split the matches; per-hypothesis counts are reduced with
// warp-aggregated integer atomics (exact and order independent).  FP64-pipe bound.
#include "common.cuh"

#define RS_THREADS 128
#define RS_PTS 512  // matches per shared-memory chunk

struct Emat { double e[9]; };

__device__ __forceinline__ int sampson_inlier(const Emat& E, const double4 p, const ScoreParams& sp) {
  const double x1 = p.x, y1 = p.y, x2 = p.z, y2 = p.w;
  // Matx33d * Vec3d(x1, y1, 1): s = 0; s += e0*x; s += e1*y; s += e2*1
  const double a0 = __dadd_rn(__dadd_rn(__dmul_rn(E.e[0], x1), __dmul_rn(E.e[1], y1)), E.e[2]);
  const double a1 = __dadd_rn(__dadd_rn(__dmul_rn(E.e[3], x1), __dmul_rn(E.e[4], y1)), E.e[5]);
  const double a2 = __dadd_rn(__dadd_rn(__dmul_rn(E.e[6], x1), __dmul_rn(E.e[7], y1)), E.e[8]);
  // E^T * Vec3d(x2, y2, 1), rows 0 and 1
  const double b0 = __dadd_rn(__dadd_rn(__dmul_rn(E.e[0], x2), __dmul_rn(E.e[3], y2)), E.e[6]);
  const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(E.e[1], x2), __dmul_rn(E.e[4], y2)), E.e[7]);
  // x2 . Ex1
  const double s = __dadd_rn(__dadd_rn(__dmul_rn(x2, a0), __dmul_rn(y2, a1)), a2);
  const double num = __dmul_rn(s, s);
  const double den = __dadd_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(b0, b0)),
      __dmul_rn(b1, b1));
  if (num < __dmul_rn(sp.mid_lo, den)) return 1;
  if (num > __dmul_rn(sp.mid_hi, den)) return 0;
  return (float)__ddiv_rn(num, den) <= sp.t ? 1 : 0;  // boundary sliver, NaN, 0/0
}

__global__ void normalize_points_kernel(const float2* __restrict__ p1,
                                        const float2* __restrict__ p2, int total, ScoreParams sp,
                                        double4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float2 a = p1[i], b = p2[i];
  double4 o;
  o.x = __dadd_rn(__dmul_rn((double)a.x, sp.ax), sp.bx);
  o.y = __dadd_rn(__dmul_rn((double)a.y, sp.ay), sp.by);
  o.z = __dadd_rn(__dmul_rn((double)b.x, sp.ax), sp.bx);
  o.w = __dadd_rn(__dmul_rn((double)b.y, sp.ay), sp.by);
  out[i] = o;
}

// getKeyPointCoordsFromFramePair (featureMatchingCommon.cpp:23-33) fused with the normalisation:
// pair p, match i -> {prev[queryIdx].pt, next[trainIdx].pt}.
__global__ void gather_normalize_kernel(const float2* __restrict__ q_xy,
                                        const float2* const* __restrict__ t_xy,
                                        const slamb200_dmatch* __restrict__ matches, int cap,
                                        const int32_t* __restrict__ n_match, ScoreParams sp,
                                        double4* __restrict__ out) {
  const int pair = blockIdx.y;
  const int n = n_match[pair];
  const float2* __restrict__ txy = t_xy[pair];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const slamb200_dmatch m = matches[(size_t)pair * cap + i];
    const float2 a = q_xy[m.queryIdx], b = txy[m.trainIdx];
    double4 o;
    o.x = __dadd_rn(__dmul_rn((double)a.x, sp.ax), sp.bx);
    o.y = __dadd_rn(__dmul_rn((double)a.y, sp.ay), sp.by);
    o.z = __dadd_rn(__dmul_rn((double)b.x, sp.ax), sp.bx);
    o.w = __dadd_rn(__dmul_rn((double)b.y, sp.ay), sp.by);
    out[(size_t)pair * cap + i] = o;
  }
}

// Matches of pair p live at npts[m_off[p] .. ) when m_off != nullptr, else at npts[p*m_stride ..);
// their count is m_cnt[p] when m_cnt != nullptr, else m_off[p+1]-m_off[p].
__device__ __forceinline__ void pair_range(const int32_t* m_off, const int32_t* m_cnt,
                                           int m_stride, int pair, size_t& base, int& cnt) {
  base = m_off ? (size_t)m_off[pair] : (size_t)pair * m_stride;
  cnt = m_cnt ? m_cnt[pair] : (m_off[pair + 1] - m_off[pair]);
}

__global__ void __launch_bounds__(RS_THREADS)
score_counts_kernel(const double4* __restrict__ npts, const int32_t* __restrict__ m_off,
                    const int32_t* __restrict__ m_cnt, int m_stride,
                    const double* __restrict__ E, int H, ScoreParams sp,
                    int32_t* __restrict__ counts) {
  __shared__ double4 pts[RS_PTS];
  const int pair = blockIdx.z;
  size_t base; int M;
  pair_range(m_off, m_cnt, m_stride, pair, base, M);
  const int h = blockIdx.x * RS_THREADS + threadIdx.x;
  Emat Eh;
  {
    const double* src = E + ((size_t)pair * H + min(h, H - 1)) * 9;
#pragma unroll
    for (int k = 0; k < 9; k++) Eh.e[k] = src[k];
  }
  int cnt = 0;
  // this block's slice of the matches: chunks blockIdx.y, blockIdx.y + gridDim.y, ...
  for (int c0 = blockIdx.y * RS_PTS; c0 < M; c0 += gridDim.y * RS_PTS) {
    const int n = min(RS_PTS, M - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += RS_THREADS) pts[i] = npts[base + c0 + i];
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < n; i++) cnt += sampson_inlier(Eh, pts[i], sp);
  }
  if (h < H && cnt) atomicAdd(&counts[(size_t)pair * H + h], cnt);
}

// RANSACPointSetRegistrator::run's update rule over a fixed list: first index with the maximum
// count, provided that count exceeds 4 (modelPoints - 1); else -1.
__global__ void score_best_kernel(const int32_t* __restrict__ counts, int H,
                                  int32_t* __restrict__ best) {
  __shared__ unsigned long long red[32];
  const int pair = blockIdx.x;
  unsigned long long key = 0;  // (count << 32) | (0xFFFFFFFF - h): max => highest count, lowest h
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const unsigned long long k =
        ((unsigned long long)(uint32_t)counts[(size_t)pair * H + h] << 32) |
        (unsigned long long)(0xFFFFFFFFu - (uint32_t)h);
    key = k > key ? k : key;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
    key = o > key ? o : key;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) key = red[w] > key ? red[w] : key;
    const int32_t c = (int32_t)(key >> 32);
    best[pair] = (H > 0 && c > 4) ? (int32_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu)) : -1;
  }
}

__global__ void score_mask_kernel(const double4* __restrict__ npts,
                                  const int32_t* __restrict__ m_off,
                                  const int32_t* __restrict__ m_cnt, int m_stride,
                                  const double* __restrict__ E, int H,
                                  const int32_t* __restrict__ best, ScoreParams sp,
                                  uint8_t* __restrict__ mask) {
  const int pair = blockIdx.y;
  size_t base; int M;
  pair_range(m_off, m_cnt, m_stride, pair, base, M);
  const int b = best[pair];
  Emat Eh;
  if (b >= 0) {
#pragma unroll
    for (int k = 0; k < 9; k++) Eh.e[k] = E[((size_t)pair * H + b) * 9 + k];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x)
    mask[base + i] = b >= 0 ? (uint8_t)sampson_inlier(Eh, npts[base + i], sp) : (uint8_t)0;
}

// Parity aid: every (hypothesis, match) flag as a byte, one thread per match, hypotheses along y.
__global__ void score_all_masks_kernel(const double4* __restrict__ npts, int M,
                                       const double* __restrict__ E, int H, ScoreParams sp,
                                       uint8_t* __restrict__ masks) {
  const int h = blockIdx.y;
  Emat Eh;
#pragma unroll
  for (int k = 0; k < 9; k++) Eh.e[k] = E[(size_t)h * 9 + k];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x)
    masks[(size_t)h * M + i] = (uint8_t)sampson_inlier(Eh, npts[i], sp);
}

void launch_normalize_points(const float2* p1, const float2* p2, int total, ScoreParams sp,
                             double4* out, cudaStream_t s) {
  if (total <= 0) return;
  normalize_points_kernel<<<(total + 255) / 256, 256, 0, s>>>(p1, p2, total, sp, out);
  COUNT_LAUNCH();
}

void launch_gather_normalize(const float2* q_xy, const float2* const* t_xy,
                             const slamb200_dmatch* matches, int cap, const int32_t* n_match,
                             int n_pairs, ScoreParams sp, double4* out, cudaStream_t s) {
  if (n_pairs <= 0 || cap <= 0) return;
  dim3 grid(min((cap + 255) / 256, 64), n_pairs);
  gather_normalize_kernel<<<grid, 256, 0, s>>>(q_xy, t_xy, matches, cap, n_match, sp, out);
  COUNT_LAUNCH();
}

void launch_score_counts(const double4* npts, const int32_t* m_off, const int32_t* m_cnt,
                         int m_stride, const double* E, int H, int P, ScoreParams sp,
                         int32_t* counts, cudaStream_t s) {
  if (P <= 0 || H <= 0) return;
  cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)P * H, s);
  const int hb = (H + RS_THREADS - 1) / RS_THREADS;
  // enough match-splits to fill the machine for a few waves when P*hb alone cannot
  int ms = (148 * 8 + hb * P - 1) / (hb * P);
  ms = max(1, min(ms, 16));
  dim3 grid(hb, ms, P);
  score_counts_kernel<<<grid, RS_THREADS, 0, s>>>(npts, m_off, m_cnt, m_stride, E, H, sp, counts);
  COUNT_LAUNCH();
}

void launch_score_best(const int32_t* counts, int H, int P, int32_t* best, cudaStream_t s) {
  if (P <= 0) return;
  score_best_kernel<<<P, 256, 0, s>>>(counts, H, best);
  COUNT_LAUNCH();
}

void launch_score_mask(const double4* npts, const int32_t* m_off, const int32_t* m_cnt,
                       int m_stride, const double* E, int H, int P, const int32_t* best,
                       ScoreParams sp, uint8_t* mask, cudaStream_t s) {
  if (P <= 0) return;
  dim3 grid(8, P);
  score_mask_kernel<<<grid, 256, 0, s>>>(npts, m_off, m_cnt, m_stride, E, H, best, sp, mask);
  COUNT_LAUNCH();
}

void launch_score_all_masks(const double4* npts, int M, const double* E, int H, ScoreParams sp,
                            uint8_t* masks, cudaStream_t s) {
  if (M <= 0 || H <= 0) return;
  dim3 grid(min((M + 255) / 256, 32), H);
  score_all_masks_kernel<<<grid, 256, 0, s>>>(npts, M, E, H, sp, masks);
  COUNT_LAUNCH();
}
