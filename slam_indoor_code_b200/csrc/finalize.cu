// finalize.cu -- merge partial top-2 records, Lowe ratio test, ordered compaction.
//
// Restates getGoodMatches (src/mainModule/featureMatching/featureMatchingCommon.cpp:37-50) on the
// device: rows in ascending queryIdx; a row with an empty k-NN list is skipped (:45-46); row q is
// kept iff (double)d0 < ratio * (double)d1, ratio being the knnMatcherDistance double (:42,:47),
// strict '<', in double.  The reference reads [1] unchecked when the list has a single entry
// (train set of one row): undefined there, defined here as "reject" (SURVEY.md appendix A.5).
// The output records are cv::DMatch-compatible {queryIdx, trainIdx, imgIdx = 0, distance}.
#include "common.cuh"

#define FIN_THREADS 256

__device__ __forceinline__ bool lt_ki(uint32_t ka, uint32_t ia, uint32_t kb, uint32_t ib) {
  return ka < kb || (ka == kb && ia < ib);
}
__device__ __forceinline__ void top2_insert(uint4& r, uint32_t k, uint32_t i) {
  if (lt_ki(k, i, r.z, r.w)) {
    if (lt_ki(k, i, r.x, r.y)) { r.z = r.x; r.w = r.y; r.x = k; r.y = i; }
    else { r.z = k; r.w = i; }
  }
}

__global__ void __launch_bounds__(FIN_THREADS)
finalize_rows_kernel(const uint4* __restrict__ part, int nq, int n_split, int hamming,
                     double ratio, int32_t* __restrict__ knn_idx, float* __restrict__ knn_dist,
                     uint8_t* __restrict__ flags, int32_t* __restrict__ chunk_cnt) {
  const int pair = blockIdx.y;
  const int q = blockIdx.x * FIN_THREADS + threadIdx.x;
  int keep = 0;
  PDL_TRIGGER();
  PDL_WAIT();
  if (q < nq) {
    uint4 r = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
    for (int s = 0; s < n_split; s++) {
      const uint4 p = part[((size_t)pair * n_split + s) * nq + q];
      if (p.y != 0xFFFFFFFFu) top2_insert(r, p.x, p.y);
      if (p.w != 0xFFFFFFFFu) top2_insert(r, p.z, p.w);
    }
    const bool has0 = r.y != 0xFFFFFFFFu, has1 = r.w != 0xFFFFFFFFu;
    const float d0 = hamming ? (float)r.x : __uint_as_float(r.x);
    const float d1 = hamming ? (float)r.z : __uint_as_float(r.z);
    const size_t o = ((size_t)pair * nq + q) * 2;
    knn_idx[o] = has0 ? (int32_t)r.y : -1;
    knn_idx[o + 1] = has1 ? (int32_t)r.w : -1;
    knn_dist[o] = has0 ? d0 : 0.f;
    knn_dist[o + 1] = has1 ? d1 : 0.f;
    keep = (has0 && has1 && (double)d0 < __dmul_rn(ratio, (double)d1)) ? 1 : 0;
    flags[(size_t)pair * nq + q] = (uint8_t)keep;
  }
  const int n = __syncthreads_count(keep);
  if (threadIdx.x == 0) chunk_cnt[pair * gridDim.x + blockIdx.x] = n;
}

__global__ void __launch_bounds__(FIN_THREADS)
compact_kernel(const int32_t* __restrict__ knn_idx, const float* __restrict__ knn_dist,
               const uint8_t* __restrict__ flags, const int32_t* __restrict__ chunk_cnt, int nq,
               slamb200_dmatch* __restrict__ out, int cap, int32_t* __restrict__ n_out) {
  __shared__ int warp_sum[FIN_THREADS / 32];
  __shared__ int base_s;
  const int pair = blockIdx.y;
  const int chunk = blockIdx.x;
  const int q = chunk * FIN_THREADS + threadIdx.x;
  PDL_WAIT();
  if (threadIdx.x < 32) {
    int acc = 0;
    for (int c = threadIdx.x; c < chunk; c += 32) acc += chunk_cnt[pair * gridDim.x + c];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (threadIdx.x == 0) base_s = acc;
  }
  const int keep = (q < nq) ? flags[(size_t)pair * nq + q] : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sum[warp] = __popc(bal);
  __syncthreads();
  int off = base_s;
  for (int w = 0; w < warp; w++) off += warp_sum[w];
  off += __popc(bal & ((1u << lane) - 1));
  if (keep && off < cap) {
    const size_t o = ((size_t)pair * nq + q) * 2;
    slamb200_dmatch m;
    m.queryIdx = q;
    m.trainIdx = knn_idx[o];
    m.imgIdx = 0;
    m.distance = knn_dist[o];
    out[(size_t)pair * cap + off] = m;
  }
  if (chunk == gridDim.x - 1 && threadIdx.x == 0) {
    int tot = base_s;
    for (int w = 0; w < FIN_THREADS / 32; w++) tot += warp_sum[w];
    n_out[pair] = tot;
  }
}

int finalize_chunks(int nq) { return (nq + FIN_THREADS - 1) / FIN_THREADS; }

void launch_finalize(const uint4* part, int nq, const PairArgs* pairs, int n_pairs, int n_split,
                     int hamming, double ratio, int32_t* knn_idx, float* knn_dist,
                     uint8_t* flags, int32_t* chunk_cnt, slamb200_dmatch* out, int cap,
                     int32_t* n_out, cudaStream_t s) {
  (void)pairs;
  if (n_pairs <= 0) return;
  if (nq <= 0) {
    cudaMemsetAsync(n_out, 0, sizeof(int32_t) * n_pairs, s);
    return;
  }
  dim3 grid(finalize_chunks(nq), n_pairs);
  launch_pdl(finalize_rows_kernel, grid, dim3(FIN_THREADS), 0, s, part, nq, n_split, hamming, ratio, knn_idx,
             knn_dist, flags, chunk_cnt);
  COUNT_LAUNCH();
  launch_pdl(compact_kernel, grid, dim3(FIN_THREADS), 0, s, (const int32_t*)knn_idx, (const float*)knn_dist,
             (const uint8_t*)flags, (const int32_t*)chunk_cnt, nq, out, cap, n_out);
  COUNT_LAUNCH();
}

// Ordered compaction alone, behind a kernel that has already produced flags / chunk counts / best
// index and distance per row (the fused tail of the tensor-core match path, sift_tc.cu).
void launch_compact(int nq, int n_pairs, const int32_t* knn_idx, const float* knn_dist,
                    const uint8_t* flags, const int32_t* chunk_cnt, slamb200_dmatch* out, int cap,
                    int32_t* n_out, cudaStream_t s) {
  if (n_pairs <= 0) return;
  if (nq <= 0) {
    cudaMemsetAsync(n_out, 0, sizeof(int32_t) * n_pairs, s);
    return;
  }
  dim3 grid(finalize_chunks(nq), n_pairs);
  launch_pdl(compact_kernel, grid, dim3(FIN_THREADS), 0, s, knn_idx, knn_dist, flags, chunk_cnt, nq, out, cap,
             n_out);
  COUNT_LAUNCH();
}
