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

// A scorer policy names the model (doubles per hypothesis), the launch parameters and the
// per-(hypothesis, correspondence) inlier predicate; the kernels below are shared.
struct SampsonScorer {
  static constexpr int ND = 9;
  typedef ScoreParams Params;
  __device__ static __forceinline__ int inlier(const double (&e)[9], const double4 p,
                                               const ScoreParams& sp) {
    const double x1 = p.x, y1 = p.y, x2 = p.z, y2 = p.w;
    // Matx33d * Vec3d(x1, y1, 1): s = 0; s += e0*x; s += e1*y; s += e2*1
    const double a0 = __dadd_rn(__dadd_rn(__dmul_rn(e[0], x1), __dmul_rn(e[1], y1)), e[2]);
    const double a1 = __dadd_rn(__dadd_rn(__dmul_rn(e[3], x1), __dmul_rn(e[4], y1)), e[5]);
    const double a2 = __dadd_rn(__dadd_rn(__dmul_rn(e[6], x1), __dmul_rn(e[7], y1)), e[8]);
    // E^T * Vec3d(x2, y2, 1), rows 0 and 1
    const double b0 = __dadd_rn(__dadd_rn(__dmul_rn(e[0], x2), __dmul_rn(e[3], y2)), e[6]);
    const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(e[1], x2), __dmul_rn(e[4], y2)), e[7]);
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
  // The counting kernel's form of the same test, branch-free: 1 = inlier, 0 = outlier, 2 = inside
  // the sliver (the caller then takes inlier()).  num and m = RN(mid * den) are non-negative
  // doubles, whose order is the order of their bit patterns as integers: "num is more than 32
  // ulps below m" implies num < mid*den*(1 - 2^-49) (an ulp is at least 2^-53 of the value, m is
  // within 2^-53 of mid*den), likewise above -- the proofs of inlier() with a wider margin.  One
  // DMUL and two 64-bit integer compares (ALU pipe) replace two DMULs and two DSETPs on the FP64
  // pipe, and without branches the compiler interleaves the dependent chains of several matches.
  // NaN (0/0, inf/inf, NaN inputs) lands in the sliver or on the outlier side, where inlier()
  // and this test agree: a NaN pattern compares above every finite value, and a NaN den makes
  // num NaN as well.
  static constexpr bool HAS_CLASSIFY = true;
  __device__ static __forceinline__ int classify(const double (&e)[9], const double4 p,
                                                 const ScoreParams& sp) {
    const double x1 = p.x, y1 = p.y, x2 = p.z, y2 = p.w;
    const double a0 = __dadd_rn(__dadd_rn(__dmul_rn(e[0], x1), __dmul_rn(e[1], y1)), e[2]);
    const double a1 = __dadd_rn(__dadd_rn(__dmul_rn(e[3], x1), __dmul_rn(e[4], y1)), e[5]);
    const double a2 = __dadd_rn(__dadd_rn(__dmul_rn(e[6], x1), __dmul_rn(e[7], y1)), e[8]);
    const double b0 = __dadd_rn(__dadd_rn(__dmul_rn(e[0], x2), __dmul_rn(e[3], y2)), e[6]);
    const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(e[1], x2), __dmul_rn(e[4], y2)), e[7]);
    const double s = __dadd_rn(__dadd_rn(__dmul_rn(x2, a0), __dmul_rn(y2, a1)), a2);
    const double num = __dmul_rn(s, s);
    const double den = __dadd_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(b0, b0)),
        __dmul_rn(b1, b1));
    const unsigned long long nb = (unsigned long long)__double_as_longlong(num);
    const unsigned long long mb = (unsigned long long)__double_as_longlong(__dmul_rn(sp.mid, den));
    const int in = nb + 32ull < mb ? 1 : 0;
    const int out = nb > mb + 32ull ? 1 : 0;
    return in ? 1 : (out ? 0 : 2);
  }
};

// solvePnPRansac's error (reference call cycleProcessing/mainCycle.cpp:155-159):
// PnPRansacCallback::computeError = cv::projectPoints (cvProjectPoints2Internal: double
// arithmetic, float result) then err = (float)norm(Matx21f(image - projected), NORM_L2SQR) with a
// float accumulator.  The model is R (row-major) then t; a correspondence is {X, Y, Z, (u, v)
// as two floats in the fourth double}.  DIST: 0 = all distortion coefficients zero (the
// polynomial collapses to x*1*1 + 0 + 0 ..., skipped), 1 = k1 k2 p1 p2 k3 only (rational and
// thin-prism terms are 1 and +0), 2 = the literal twelve-coefficient formula.  The skipped forms
// differ from the literal one only where r^6 overflows, and there both sides reject the point
// (literal: 0*inf = NaN; skipped: |u| beyond float range -> err = inf).
template <int DIST>
struct ReprojScorer {
  static constexpr int ND = 12;
  typedef PnpParams Params;
  static constexpr bool HAS_CLASSIFY = false;
  __device__ static __forceinline__ int classify(const double (&m)[12], const double4 p, const PnpParams& pp) {
    return inlier(m, p, pp);
  }
  __device__ static __forceinline__ int inlier(const double (&m)[12], const double4 p,
                                               const PnpParams& pp) {
    const double X = p.x, Y = p.y, Z = p.z;
    double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], X), __dmul_rn(m[1], Y)), __dmul_rn(m[2], Z)), m[9]);
    double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[3], X), __dmul_rn(m[4], Y)), __dmul_rn(m[5], Z)), m[10]);
    double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[6], X), __dmul_rn(m[7], Y)), __dmul_rn(m[8], Z)), m[11]);
    z = z != 0.0 ? __drcp_rn(z) : 1.0;  // z ? 1./z : 1
    x = __dmul_rn(x, z);
    y = __dmul_rn(y, z);
    double xd = x, yd = y;
    if (DIST > 0) {
      const double* k = pp.k;
      const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
      const double r4 = __dmul_rn(r2, r2);
      const double r6 = __dmul_rn(r4, r2);
      const double x2 = __dmul_rn(2.0, x), y2 = __dmul_rn(2.0, y);
      const double a1 = __dmul_rn(x2, y);                 // 2*x*y
      const double a2 = __dadd_rn(r2, __dmul_rn(x2, x));  // r2 + 2*x*x
      const double a3 = __dadd_rn(r2, __dmul_rn(y2, y));
      const double cdist = __dadd_rn(__dadd_rn(__dadd_rn(1.0, __dmul_rn(k[0], r2)), __dmul_rn(k[1], r4)),
                                     __dmul_rn(k[4], r6));
      double xc = __dmul_rn(x, cdist), yc = __dmul_rn(y, cdist);
      if (DIST > 1) {
        const double icdist2 = __drcp_rn(__dadd_rn(
            __dadd_rn(__dadd_rn(1.0, __dmul_rn(k[5], r2)), __dmul_rn(k[6], r4)), __dmul_rn(k[7], r6)));
        xc = __dmul_rn(xc, icdist2);
        yc = __dmul_rn(yc, icdist2);
      }
      xd = __dadd_rn(__dadd_rn(xc, __dmul_rn(k[2], a1)), __dmul_rn(k[3], a2));
      yd = __dadd_rn(__dadd_rn(yc, __dmul_rn(k[2], a3)), __dmul_rn(k[3], a1));
      if (DIST > 1) {
        xd = __dadd_rn(__dadd_rn(xd, __dmul_rn(k[8], r2)), __dmul_rn(k[9], r4));
        yd = __dadd_rn(__dadd_rn(yd, __dmul_rn(k[10], r2)), __dmul_rn(k[11], r4));
      }
    }
    const float u = __double2float_rn(__dadd_rn(__dmul_rn(xd, pp.fx), pp.cx));
    const float v = __double2float_rn(__dadd_rn(__dmul_rn(yd, pp.fy), pp.cy));
    const float2 uv = *reinterpret_cast<const float2*>(&p.w);
    const float dx = __fsub_rn(uv.x, u), dy = __fsub_rn(uv.y, v);
    const float err = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // s = 0; s += dx*dx; s += dy*dy
    return err <= pp.t ? 1 : 0;
  }
};

// {X, Y, Z} float -> double (exact) and the image point packed into the fourth double.
__global__ void pack_pnp_points_kernel(const float* __restrict__ obj, const float2* __restrict__ img,
                                       int total, double4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double4 o;
  o.x = (double)obj[3 * (size_t)i];
  o.y = (double)obj[3 * (size_t)i + 1];
  o.z = (double)obj[3 * (size_t)i + 2];
  *reinterpret_cast<float2*>(&o.w) = img[i];
  out[i] = o;
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

// HPT hypotheses per thread (h, h + RS_THREADS, ...): every match read from shared memory feeds HPT
// independent dependency chains (more instruction-level parallelism on the FP64 pipe, fewer
// shared-memory reads per scored pair).
template <class S, int HPT>
__global__ void __launch_bounds__(RS_THREADS)
score_counts_kernel(const double4* __restrict__ npts, const int32_t* __restrict__ m_off,
                    const int32_t* __restrict__ m_cnt, int m_stride,
                    const double* __restrict__ E, int H, const typename S::Params sp,
                    int32_t* __restrict__ counts) {
  __shared__ double4 pts[RS_PTS];
  const int pair = blockIdx.z;
  size_t base; int M;
  pair_range(m_off, m_cnt, m_stride, pair, base, M);
  const int h0 = blockIdx.x * RS_THREADS * HPT + threadIdx.x;
  double Eh[HPT][S::ND];
#pragma unroll
  for (int v = 0; v < HPT; v++) {
    const double* src = E + ((size_t)pair * H + min(h0 + v * RS_THREADS, H - 1)) * S::ND;
#pragma unroll
    for (int k = 0; k < S::ND; k++) Eh[v][k] = src[k];
  }
  int cnt[HPT];
#pragma unroll
  for (int v = 0; v < HPT; v++) cnt[v] = 0;
  // this block's slice of the matches: chunks blockIdx.y, blockIdx.y + gridDim.y, ...
  for (int c0 = blockIdx.y * RS_PTS; c0 < M; c0 += gridDim.y * RS_PTS) {
    const int n = min(RS_PTS, M - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += RS_THREADS) pts[i] = npts[base + c0 + i];
    __syncthreads();
    if (S::HAS_CLASSIFY) {
      // U matches x HPT hypotheses at a time, branch-free; the rare boundary cases are redone
      // exactly afterwards
      constexpr int U = HPT > 1 ? 2 : 4;
      int i = 0;
      for (; i + U <= n; i += U) {
        int c[U][HPT];
        int any = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const double4 p = pts[i + u];
#pragma unroll
          for (int v = 0; v < HPT; v++) {
            c[u][v] = S::classify(Eh[v], p, sp);
            cnt[v] += c[u][v] & 1;
            any |= c[u][v];
          }
        }
        if ((any & 2) != 0) {
#pragma unroll
          for (int u = 0; u < U; u++)
#pragma unroll
            for (int v = 0; v < HPT; v++)
              if (c[u][v] == 2) cnt[v] += S::inlier(Eh[v], pts[i + u], sp);
        }
      }
      for (; i < n; i++)
#pragma unroll
        for (int v = 0; v < HPT; v++) cnt[v] += S::inlier(Eh[v], pts[i], sp);
    } else {
#pragma unroll 2
      for (int i = 0; i < n; i++)
#pragma unroll
        for (int v = 0; v < HPT; v++) cnt[v] += S::inlier(Eh[v], pts[i], sp);
    }
  }
#pragma unroll
  for (int v = 0; v < HPT; v++) {
    const int h = h0 + v * RS_THREADS;
    if (h < H && cnt[v]) atomicAdd(&counts[(size_t)pair * H + h], cnt[v]);
  }
}

// RANSACPointSetRegistrator::run's update rule over a fixed list: first index with the maximum
// count, provided that count exceeds min_count (modelPoints - 1); else -1.
__global__ void score_best_kernel(const int32_t* __restrict__ counts, int H, int min_count,
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
    best[pair] = (H > 0 && c > min_count) ? (int32_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu)) : -1;
  }
}

template <class S>
__global__ void score_mask_kernel(const double4* __restrict__ npts,
                                  const int32_t* __restrict__ m_off,
                                  const int32_t* __restrict__ m_cnt, int m_stride,
                                  const double* __restrict__ E, int H,
                                  const int32_t* __restrict__ best, const typename S::Params sp,
                                  uint8_t* __restrict__ mask) {
  const int pair = blockIdx.y;
  size_t base; int M;
  pair_range(m_off, m_cnt, m_stride, pair, base, M);
  const int b = best[pair];
  double Eh[S::ND];
#pragma unroll
  for (int k = 0; k < S::ND; k++) Eh[k] = b >= 0 ? E[((size_t)pair * H + b) * S::ND + k] : 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x)
    mask[base + i] = b >= 0 ? (uint8_t)S::inlier(Eh, npts[base + i], sp) : (uint8_t)0;
}

// Parity aid: every (hypothesis, match) flag as a byte, one thread per match, hypotheses along y.
template <class S>
__global__ void score_all_masks_kernel(const double4* __restrict__ npts, int M,
                                       const double* __restrict__ E, int H,
                                       const typename S::Params sp, uint8_t* __restrict__ masks) {
  const int h = blockIdx.y;
  double Eh[S::ND];
#pragma unroll
  for (int k = 0; k < S::ND; k++) Eh[k] = E[(size_t)h * S::ND + k];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x)
    masks[(size_t)h * M + i] = (uint8_t)S::inlier(Eh, npts[i], sp);
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

template <class S>
static void score_counts_t(const double4* npts, const int32_t* m_off, const int32_t* m_cnt,
                           int m_stride, const double* E, int H, int P,
                           const typename S::Params& sp, int32_t* counts, cudaStream_t s) {
  if (P <= 0 || H <= 0) return;
  cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)P * H, s);
  // two hypotheses per thread once there are enough of them to fill the blocks that way
  const int hpt = (S::HAS_CLASSIFY && H >= 2 * RS_THREADS) ? 2 : 1;
  const int hb = (H + RS_THREADS * hpt - 1) / (RS_THREADS * hpt);
  // enough match-splits to fill the machine for a few waves when P*hb alone cannot
  int ms = (148 * 8 + hb * P - 1) / (hb * P);
  ms = max(1, min(ms, 16));
  dim3 grid(hb, ms, P);
  if (hpt == 2) score_counts_kernel<S, 2><<<grid, RS_THREADS, 0, s>>>(npts, m_off, m_cnt, m_stride, E, H, sp, counts);
  else score_counts_kernel<S, 1><<<grid, RS_THREADS, 0, s>>>(npts, m_off, m_cnt, m_stride, E, H, sp, counts);
  COUNT_LAUNCH();
}

template <class S>
static void score_mask_t(const double4* npts, const int32_t* m_off, const int32_t* m_cnt,
                         int m_stride, const double* E, int H, int P, const int32_t* best,
                         const typename S::Params& sp, uint8_t* mask, cudaStream_t s) {
  if (P <= 0) return;
  dim3 grid(8, P);
  score_mask_kernel<S><<<grid, 256, 0, s>>>(npts, m_off, m_cnt, m_stride, E, H, best, sp, mask);
  COUNT_LAUNCH();
}

template <class S>
static void score_all_masks_t(const double4* npts, int M, const double* E, int H,
                              const typename S::Params& sp, uint8_t* masks, cudaStream_t s) {
  if (M <= 0 || H <= 0) return;
  dim3 grid(min((M + 255) / 256, 32), H);
  score_all_masks_kernel<S><<<grid, 256, 0, s>>>(npts, M, E, H, sp, masks);
  COUNT_LAUNCH();
}

void launch_score_counts(const double4* npts, const int32_t* m_off, const int32_t* m_cnt,
                         int m_stride, const double* E, int H, int P, ScoreParams sp,
                         int32_t* counts, cudaStream_t s) {
  score_counts_t<SampsonScorer>(npts, m_off, m_cnt, m_stride, E, H, P, sp, counts, s);
}

void launch_score_best(const int32_t* counts, int H, int P, int min_count, int32_t* best,
                       cudaStream_t s) {
  if (P <= 0) return;
  score_best_kernel<<<P, 256, 0, s>>>(counts, H, min_count, best);
  COUNT_LAUNCH();
}

void launch_score_mask(const double4* npts, const int32_t* m_off, const int32_t* m_cnt,
                       int m_stride, const double* E, int H, int P, const int32_t* best,
                       ScoreParams sp, uint8_t* mask, cudaStream_t s) {
  score_mask_t<SampsonScorer>(npts, m_off, m_cnt, m_stride, E, H, P, best, sp, mask, s);
}

void launch_score_all_masks(const double4* npts, int M, const double* E, int H, ScoreParams sp,
                            uint8_t* masks, cudaStream_t s) {
  score_all_masks_t<SampsonScorer>(npts, M, E, H, sp, masks, s);
}

// ---- solvePnPRansac scoring --------------------------------------------------------------------
void launch_pack_pnp_points(const float* obj, const float2* img, int total, double4* out,
                            cudaStream_t s) {
  if (total <= 0) return;
  pack_pnp_points_kernel<<<(total + 255) / 256, 256, 0, s>>>(obj, img, total, out);
  COUNT_LAUNCH();
}

void launch_pnp_counts(const double4* pts, const int32_t* m_off, const double* poses, int H, int P,
                       const PnpParams& pp, int32_t* counts, cudaStream_t s) {
  if (pp.dist_level == 0)
    score_counts_t<ReprojScorer<0>>(pts, m_off, nullptr, 0, poses, H, P, pp, counts, s);
  else if (pp.dist_level == 1)
    score_counts_t<ReprojScorer<1>>(pts, m_off, nullptr, 0, poses, H, P, pp, counts, s);
  else
    score_counts_t<ReprojScorer<2>>(pts, m_off, nullptr, 0, poses, H, P, pp, counts, s);
}

void launch_pnp_mask(const double4* pts, const int32_t* m_off, const double* poses, int H, int P,
                     const int32_t* best, const PnpParams& pp, uint8_t* mask, cudaStream_t s) {
  if (pp.dist_level == 0)
    score_mask_t<ReprojScorer<0>>(pts, m_off, nullptr, 0, poses, H, P, best, pp, mask, s);
  else if (pp.dist_level == 1)
    score_mask_t<ReprojScorer<1>>(pts, m_off, nullptr, 0, poses, H, P, best, pp, mask, s);
  else
    score_mask_t<ReprojScorer<2>>(pts, m_off, nullptr, 0, poses, H, P, best, pp, mask, s);
}

void launch_pnp_all_masks(const double4* pts, int M, const double* poses, int H,
                          const PnpParams& pp, uint8_t* masks, cudaStream_t s) {
  if (pp.dist_level == 0) score_all_masks_t<ReprojScorer<0>>(pts, M, poses, H, pp, masks, s);
  else if (pp.dist_level == 1) score_all_masks_t<ReprojScorer<1>>(pts, M, poses, H, pp, masks, s);
  else score_all_masks_t<ReprojScorer<2>>(pts, M, poses, H, pp, masks, s);
}
