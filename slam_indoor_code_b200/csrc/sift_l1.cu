// sift_l1.cu -- NORM_L1 k=2 nearest neighbours for integer-valued SIFT rows.
//
// The reference's OpenCV-CUDA build answers useFM-SIFT-BF with
// cv::cuda::DescriptorMatcher::createBFMatcher(NORM_L1) (featureMatchingCUDA.cpp:28; the CPU build
// uses NORM_L2, featureMatchingCPU.cpp:28).  This is that mode for users who relied on it.  For
// the rows cv::SIFT emits (integers 0..255) every partial sum of |a-b| is an integer below 2^15,
// so the float sum OpenCV forms is exact in any order: the distance is computed on the u8 copy of
// the rows with the byte-wise sum-of-absolute-differences instruction (VABSDIFF4.U8.ACC: four
// |a-b| and the accumulate in one issue), 32 per (query, train).  Ties keep the lowest train index
// (strict-'<' insert of BFMatcher).  Non-integer sets take sift_exact_knn2_kernel<1>.
//
// Shape (as orb_knn.cu): a work item is (pair, block of L1_QB query rows, one of n_split train
// ranges); a thread keeps two query rows (2 x 32 words) in registers, train tiles stream through
// shared memory and are read as broadcast uint4; (dist << 17 | train) goes through a
// three-instruction running top-2.  Bound: the integer pipe that issues VABSDIFF4.
#include "common.cuh"

#define L1_THREADS 128
#define L1_QPT 2
#define L1_QB (L1_THREADS * L1_QPT)
#define L1_TT 64            // train rows per shared-memory tile (8 KB)
#define L1_IDX_BITS 17      // train index bits in the packed key: dist < 2^15, T <= 131072

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(b));
  return c;
}

__global__ void __launch_bounds__(L1_THREADS)
sift_l1_u8_kernel(const uint4* __restrict__ q, const int32_t* __restrict__ q_flags, int nq,
                  const PairArgs* __restrict__ pairs, int n_split, uint4* __restrict__ part) {
  __shared__ uint4 tile[2][L1_TT * 8];
  const int pair = blockIdx.z;
  const int split = blockIdx.y;
  const PairArgs pa = pairs[pair];
  // integer-valued pairs only; the others belong to the exact fp32 kernel
  if (q_flags[0] != 0 || pa.t_flags == nullptr || pa.t_flags[0] != 0) return;
  const uint4* __restrict__ t = reinterpret_cast<const uint4*>(pa.t_u8);
  const int per = (pa.t_n + n_split - 1) / n_split;
  const int t_begin = split * per;
  const int t_end = min(pa.t_n, t_begin + per);

  uint32_t qr[L1_QPT][32];
  int qrow[L1_QPT];
#pragma unroll
  for (int j = 0; j < L1_QPT; j++) {
    qrow[j] = blockIdx.x * L1_QB + j * L1_THREADS + threadIdx.x;
    const int r = min(qrow[j], nq - 1);
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const uint4 v = q[8 * (size_t)r + w];
      qr[j][4 * w] = v.x; qr[j][4 * w + 1] = v.y; qr[j][4 * w + 2] = v.z; qr[j][4 * w + 3] = v.w;
    }
  }
  uint32_t m1[L1_QPT], m2[L1_QPT];
#pragma unroll
  for (int j = 0; j < L1_QPT; j++) m1[j] = m2[j] = ABSENT_KEY;

  const int n_tiles = (t_end - t_begin + L1_TT - 1) / L1_TT;
  auto load_tile = [&](int tile_i, int buf) {
    const int base = t_begin + tile_i * L1_TT;
    for (int i = threadIdx.x; i < L1_TT * 8; i += L1_THREADS) {
      const int row = base + (i >> 3);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < t_end) v = t[8 * (size_t)row + (i & 7)];
      tile[buf][i] = v;
    }
  };
  if (n_tiles > 0) load_tile(0, 0);
  __syncthreads();
  for (int ti = 0; ti < n_tiles; ti++) {
    const int buf = ti & 1;
    if (ti + 1 < n_tiles) load_tile(ti + 1, buf ^ 1);
    const int base = t_begin + ti * L1_TT;
    const int cnt = min(L1_TT, t_end - base);
#pragma unroll 2
    for (int k = 0; k < cnt; k++) {
      uint32_t d[L1_QPT];
#pragma unroll
      for (int j = 0; j < L1_QPT; j++) d[j] = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) {
        const uint4 v = tile[buf][8 * k + w];
#pragma unroll
        for (int j = 0; j < L1_QPT; j++) {
          d[j] = sad4(qr[j][4 * w], v.x, d[j]);
          d[j] = sad4(qr[j][4 * w + 1], v.y, d[j]);
          d[j] = sad4(qr[j][4 * w + 2], v.z, d[j]);
          d[j] = sad4(qr[j][4 * w + 3], v.w, d[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < L1_QPT; j++) {
        const uint32_t key = (d[j] << L1_IDX_BITS) + (uint32_t)(base + k);
        m2[j] = min(m2[j], max(m1[j], key));
        m1[j] = min(m1[j], key);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < L1_QPT; j++) {
    if (qrow[j] < nq) {
      const uint32_t mask = (1u << L1_IDX_BITS) - 1;
      uint4 o;  // key = fp32 bits of the distance, as the float kernels write it
      o.x = m1[j] == ABSENT_KEY ? ABSENT_KEY : __float_as_uint((float)(m1[j] >> L1_IDX_BITS));
      o.y = m1[j] == ABSENT_KEY ? 0xFFFFFFFFu : (m1[j] & mask);
      o.z = m2[j] == ABSENT_KEY ? ABSENT_KEY : __float_as_uint((float)(m2[j] >> L1_IDX_BITS));
      o.w = m2[j] == ABSENT_KEY ? 0xFFFFFFFFu : (m2[j] & mask);
      part[((size_t)pair * n_split + split) * nq + qrow[j]] = o;
    }
  }
}

int sift_l1_max_train_rows() { return 1 << L1_IDX_BITS; }

void launch_sift_l1_u8(const uint8_t* q_u8, const int32_t* q_flags, int nq, const PairArgs* pairs,
                       int n_pairs, int n_split, uint4* part, cudaStream_t s) {
  if (nq <= 0 || n_pairs <= 0) return;
  dim3 grid((nq + L1_QB - 1) / L1_QB, n_split, n_pairs);
  sift_l1_u8_kernel<<<grid, L1_THREADS, 0, s>>>(reinterpret_cast<const uint4*>(q_u8), q_flags, nq,
                                                pairs, n_split, part);
  COUNT_LAUNCH();
}
