// sift_exact.cu -- exact fp32 L2 k=2 nearest neighbours in OpenCV's own summation order.
//
// This is the general-float SIFT path (and the device-side ground truth of the tcgen05 path):
// it reproduces cv::BFMatcher(NORM_L2)::knnMatch(query, train, 2) as called at
// src/mainModule/featureMatching/featureMatchingCPU.cpp:27-40 bit for bit, distances included.
// OpenCV's hal::normL2Sqr_ (SSE baseline) keeps 4 x 4 accumulators: element j = 16*i + 4*a + l
// goes to acc[a][l] with the multiply and the add rounded separately, then
// S[l] = ((acc0+acc1)+acc2)+acc3, d2 = (S0+S2)+(S1+S3), dist = sqrtf(d2).  Every operation below
// is an explicit round-to-nearest intrinsic so the compiler cannot contract anything into an FMA.
//
// NORM = 1 is the NORM_L1 twin (the OpenCV-CUDA build's useFM-SIFT-BF, featureMatchingCUDA.cpp:28;
// SURVEY.md 8f-4), in the order of cv::BFMatcher(NORM_L1) on the CPU: one accumulator,
// s += ((|d0| + |d1|) + |d2|) + |d3| over groups of four elements, no square root.
//
// Shape: a work item is (pair, 16 query rows, one of n_split train ranges); a 256-thread block
// holds the query rows in shared memory and streams 64-row train tiles through it.  Each thread
// owns a 2 x 2 block of (query, train) pairs = 64 accumulators; lanes run along the train rows
// (row pitch 132 floats -> conflict-free LDS.128), warps along the query rows (broadcast loads).
#include "common.cuh"

#define SE_THREADS 256
#define SE_QT 16     // query rows per block
#define SE_TT 64     // train rows per tile
#define SE_PITCH 132 // floats per shared row (128 + 4: 16-byte skew per row)

template <int NORM, int NV, int NL>
__device__ __forceinline__ uint32_t dist_key(const float (&acc)[NV][NL]) {
  if (NORM == 1) return __float_as_uint(acc[0][0]);
  float S[4];
#pragma unroll
  for (int l = 0; l < 4; l++)
    S[l] = __fadd_rn(__fadd_rn(__fadd_rn(acc[0][l % NL], acc[1 % NV][l % NL]), acc[2 % NV][l % NL]), acc[3 % NV][l % NL]);
  const float d2 = __fadd_rn(__fadd_rn(S[0], S[2]), __fadd_rn(S[1], S[3]));
  return __float_as_uint(sqrtf(d2));
}

// lexicographic (key, idx) top-2 insert; callers feed ascending idx per thread but the merge
// across threads needs the full comparison.
__device__ __forceinline__ bool lt_ki(uint32_t ka, uint32_t ia, uint32_t kb, uint32_t ib) {
  return ka < kb || (ka == kb && ia < ib);
}
__device__ __forceinline__ void top2_insert(uint4& r, uint32_t k, uint32_t i) {
  if (lt_ki(k, i, r.z, r.w)) {
    if (lt_ki(k, i, r.x, r.y)) { r.z = r.x; r.w = r.y; r.x = k; r.y = i; }
    else { r.z = k; r.w = i; }
  }
}

template <int NORM>
__global__ void __launch_bounds__(SE_THREADS)
sift_exact_knn2_kernel(const float* __restrict__ q, const int32_t* __restrict__ q_flags, int nq,
                       const PairArgs* __restrict__ pairs, int n_pairs, int n_split, long long n_items,
                       uint4* __restrict__ part, int force) {
  extern __shared__ __align__(16) float smem[];
  float* sq = smem;                      // [SE_QT][SE_PITCH]
  float* st = smem + SE_QT * SE_PITCH;   // [SE_TT][SE_PITCH]

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;  // 8 warps -> query rows 2*warp, 2*warp+1
  const int q_blocks = (nq + SE_QT - 1) / SE_QT;
  const bool q_exact = q_flags[0] == 0;
  // one parallel sweep over the pair flags: a batch that belongs entirely to the tcgen05 path
  // costs a single memory round trip here
  if (!force) {
    int mine = 0;
    if (!q_exact) mine = 1;
    else
      for (int p = threadIdx.x; p < n_pairs; p += SE_THREADS)
        if (pairs[p].t_flags == nullptr || pairs[p].t_flags[0] != 0) mine = 1;
    if (__syncthreads_or(mine) == 0) return;
  }

  // persistent walk over (pair, split, query block) work items: a batch whose pairs all belong
  // to the tcgen05 path costs a few flag reads per block, not one block launch per item
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
  const int qb = (int)(item % q_blocks);
  const int split = (int)((item / q_blocks) % n_split);
  const int pair = (int)(item / ((long long)q_blocks * n_split));
  const PairArgs pa = pairs[pair];
  if (!force && q_exact && pa.t_flags != nullptr && pa.t_flags[0] == 0)
    continue;  // exact-mode pair: the tcgen05 path owns it
  const float* __restrict__ t = reinterpret_cast<const float*>(pa.t_rows);
  const int per = (pa.t_n + n_split - 1) / n_split;
  const int t_begin = split * per;
  const int t_end = min(pa.t_n, t_begin + per);
  const int q_base = qb * SE_QT;
  __syncthreads();  // the previous item is done with the shared tiles

  // query tile
  for (int i = threadIdx.x; i < SE_QT * 32; i += SE_THREADS) {
    const int r = i >> 5, c = i & 31;
    const int row = min(q_base + r, nq - 1);
    const float4 v = reinterpret_cast<const float4*>(q + (size_t)row * 128)[c];
    *reinterpret_cast<float4*>(sq + r * SE_PITCH + 4 * c) = v;
  }

  uint4 best[2];
  best[0] = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
  best[1] = best[0];

  for (int tb = t_begin; tb < t_end; tb += SE_TT) {
    __syncthreads();
    for (int i = threadIdx.x; i < SE_TT * 32; i += SE_THREADS) {
      const int r = i >> 5, c = i & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tb + r < t_end) v = reinterpret_cast<const float4*>(t + (size_t)(tb + r) * 128)[c];
      *reinterpret_cast<float4*>(st + r * SE_PITCH + 4 * c) = v;
    }
    __syncthreads();

    constexpr int NV = NORM == 0 ? 4 : 1, NL = NORM == 0 ? 4 : 1;
    float acc[2][2][NV][NL];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
          for (int l = 0; l < NL; l++) acc[a][b][v][l] = 0.f;

    const float* q0 = sq + (2 * warp) * SE_PITCH;
    const float* q1 = q0 + SE_PITCH;
    const float* t0 = st + lane * SE_PITCH;
    const float* t1 = st + (lane + 32) * SE_PITCH;
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const float4 a0 = *reinterpret_cast<const float4*>(q0 + 16 * i + 4 * v);
        const float4 a1 = *reinterpret_cast<const float4*>(q1 + 16 * i + 4 * v);
        const float4 b0 = *reinterpret_cast<const float4*>(t0 + 16 * i + 4 * v);
        const float4 b1 = *reinterpret_cast<const float4*>(t1 + 16 * i + 4 * v);
        const float qa[2][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}};
        const float tb4[2][4] = {{b0.x, b0.y, b0.z, b0.w}, {b1.x, b1.y, b1.z, b1.w}};
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int b = 0; b < 2; b++) {
            if (NORM == 0) {
#pragma unroll
              for (int l = 0; l < 4; l++) {
                const float d = __fsub_rn(qa[a][l], tb4[b][l]);
                acc[a][b][v % NV][l % NL] = __fadd_rn(acc[a][b][v % NV][l % NL], __fmul_rn(d, d));
              }
            } else {
              const float d0 = fabsf(__fsub_rn(qa[a][0], tb4[b][0])), d1 = fabsf(__fsub_rn(qa[a][1], tb4[b][1]));
              const float d2 = fabsf(__fsub_rn(qa[a][2], tb4[b][2])), d3 = fabsf(__fsub_rn(qa[a][3], tb4[b][3]));
              acc[a][b][0][0] = __fadd_rn(acc[a][b][0][0], __fadd_rn(__fadd_rn(__fadd_rn(d0, d1), d2), d3));
            }
          }
      }
    }
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int trow = tb + lane + 32 * b;
      if (trow < t_end) {
#pragma unroll
        for (int a = 0; a < 2; a++) top2_insert(best[a], dist_key<NORM>(acc[a][b]), (uint32_t)trow);
      }
    }
  }

  // merge the 32 lanes of each warp (they hold disjoint train rows of the same two queries)
#pragma unroll
  for (int a = 0; a < 2; a++) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      uint4 o;
      o.x = __shfl_xor_sync(0xffffffffu, best[a].x, off);
      o.y = __shfl_xor_sync(0xffffffffu, best[a].y, off);
      o.z = __shfl_xor_sync(0xffffffffu, best[a].z, off);
      o.w = __shfl_xor_sync(0xffffffffu, best[a].w, off);
      top2_insert(best[a], o.x, o.y);
      top2_insert(best[a], o.z, o.w);
    }
    const int qrow = q_base + 2 * warp + a;
    if (lane == 0 && qrow < nq)
      part[((size_t)pair * n_split + split) * nq + qrow] = best[a];
  }
  }  // work items
}

void launch_sift_exact_knn2(const float* q, const int32_t* q_flags, int nq, const PairArgs* pairs,
                            int n_pairs, int n_split, uint4* part, int force, int norm_l1,
                            cudaStream_t s) {
  if (nq <= 0 || n_pairs <= 0) return;
  const size_t smem = (size_t)(SE_QT + SE_TT) * SE_PITCH * sizeof(float);
  static PerDeviceOnce attr_once;   // per-device attribute
  attr_once.run([smem] {
    return cudaFuncSetAttribute(sift_exact_knn2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
           cudaFuncSetAttribute(sift_exact_knn2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
  });
  const long long n_items = (long long)((nq + SE_QT - 1) / SE_QT) * n_split * n_pairs;
  const int grid = (int)(n_items < 148 * 8 ? n_items : 148 * 8);
  if (norm_l1)
    sift_exact_knn2_kernel<1><<<grid, SE_THREADS, smem, s>>>(q, q_flags, nq, pairs, n_pairs, n_split,
                                                             n_items, part, force);
  else
    sift_exact_knn2_kernel<0><<<grid, SE_THREADS, smem, s>>>(q, q_flags, nq, pairs, n_pairs, n_split,
                                                             n_items, part, force);
  COUNT_LAUNCH();
}
