// orb_knn.cu -- 256-bit Hamming k=2 nearest neighbours (useFM-ORB).
//
// Replaces cv::BFMatcher(NORM_HAMMING)::knnMatch(query, train, 2) as the reference calls it at
// src/mainModule/featureMatching/featureMatchingCPU.cpp:33-40 (OpenCV-CUDA twin:
// featureMatchingCUDA.cpp:34-44).  Distances are exact integers; ties keep the lowest train
// index (BatchDistInvoker's strict-'<' insert).
//
// Shape: a work item is (pair, block of ORB_QB query rows, one of n_split train ranges).  Each
// thread keeps ORB_QPT query descriptors (8 x u32 each) in registers; the train range streams
// through shared memory in 128-bit vectorised tiles (uint4, two per descriptor) that every lane
// of a warp reads at the same address (broadcast, conflict free).  Per (query, train): 8 LOP3
// (xor) + 3 carry-save adders (6 LOP3) + 5 POPC + adds, then one packed key
// (dist << 22 | train)  goes through a 3-instruction running top-2 (min / max / min), so ties
// resolve to the lowest index for free.  Bound by the POPC / integer pipes, not by bytes
// (0.8 MB per 10k x 10k pair).
#include "common.cuh"

#define ORB_THREADS 128
#define ORB_QPT 2                       // query rows per thread
#define ORB_QB (ORB_THREADS * ORB_QPT)  // query rows per block
#define ORB_TT 256                      // train rows per shared-memory tile
#define ORB_IDX_BITS 22                 // train index bits in the packed key (T < 4M)

__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// 256-bit Hamming distance.  The POPC pipe issues 16 results/clk/SM (measured 4.49 Tpopc/s), a
// quarter of the LOP3 rate, so three carry-save adders (sum = a^b^c, carry = maj(a,b,c), one
// LOP3 each) first fold the eight XOR words into five: popc(s3) + popc(x7) + 2*(popc(c1) +
// popc(c2) + popc(c3)).  5 POPC + 14 LOP3 instead of 8 POPC + 8 LOP3: the two pipes balance.
__device__ __forceinline__ uint32_t ham256(const uint32_t (&q)[8], const uint4 a, const uint4 b) {
  const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
  const uint32_t x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
  const uint32_t s1 = lop3_xor3(x0, x1, x2), c1 = lop3_maj(x0, x1, x2);
  const uint32_t s2 = lop3_xor3(x3, x4, x5), c2 = lop3_maj(x3, x4, x5);
  const uint32_t s3 = lop3_xor3(s1, s2, x6), c3 = lop3_maj(s1, s2, x6);
  const uint32_t twos = __popc(c1) + __popc(c2) + __popc(c3);
  return __popc(s3) + __popc(x7) + 2u * twos;
}

__global__ void __launch_bounds__(ORB_THREADS)
orb_knn2_kernel(const uint4* __restrict__ q, int nq, const PairArgs* __restrict__ pairs,
                int n_split, uint4* __restrict__ part) {
  __shared__ uint4 tile[2][ORB_TT * 2];
  const int pair = blockIdx.z;
  const int split = blockIdx.y;
  const PairArgs pa = pairs[pair];
  const uint4* __restrict__ t = reinterpret_cast<const uint4*>(pa.t_rows);
  const int per = (pa.t_n + n_split - 1) / n_split;
  const int t_begin = split * per;
  const int t_end = min(pa.t_n, t_begin + per);

  uint32_t qr[ORB_QPT][8];
  int qrow[ORB_QPT];
#pragma unroll
  for (int j = 0; j < ORB_QPT; j++) {
    qrow[j] = blockIdx.x * ORB_QB + j * ORB_THREADS + threadIdx.x;
    const int r = min(qrow[j], nq - 1);
    const uint4 a = q[2 * r], b = q[2 * r + 1];
    qr[j][0] = a.x; qr[j][1] = a.y; qr[j][2] = a.z; qr[j][3] = a.w;
    qr[j][4] = b.x; qr[j][5] = b.y; qr[j][6] = b.z; qr[j][7] = b.w;
  }
  uint32_t m1[ORB_QPT], m2[ORB_QPT];
#pragma unroll
  for (int j = 0; j < ORB_QPT; j++) m1[j] = m2[j] = ABSENT_KEY;

  const int n_tiles = (t_end - t_begin + ORB_TT - 1) / ORB_TT;
  // prologue: tile 0
  auto load_tile = [&](int tile_i, int buf) {
    const int base = t_begin + tile_i * ORB_TT;
    for (int i = threadIdx.x; i < ORB_TT * 2; i += ORB_THREADS) {
      const int row = base + (i >> 1);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < t_end) v = t[2 * (size_t)row + (i & 1)];
      tile[buf][i] = v;
    }
  };
  if (n_tiles > 0) load_tile(0, 0);
  __syncthreads();
  for (int ti = 0; ti < n_tiles; ti++) {
    const int buf = ti & 1;
    if (ti + 1 < n_tiles) load_tile(ti + 1, buf ^ 1);
    const int base = t_begin + ti * ORB_TT;
    const int cnt = min(ORB_TT, t_end - base);
    if (cnt == ORB_TT) {
#pragma unroll 4
      for (int k = 0; k < ORB_TT; k++) {
        const uint4 a = tile[buf][2 * k], b = tile[buf][2 * k + 1];
#pragma unroll
        for (int j = 0; j < ORB_QPT; j++) {
          const uint32_t key = (ham256(qr[j], a, b) << ORB_IDX_BITS) + (uint32_t)(base + k);
          m2[j] = min(m2[j], max(m1[j], key));
          m1[j] = min(m1[j], key);
        }
      }
    } else {
      for (int k = 0; k < cnt; k++) {
        const uint4 a = tile[buf][2 * k], b = tile[buf][2 * k + 1];
#pragma unroll
        for (int j = 0; j < ORB_QPT; j++) {
          const uint32_t key = (ham256(qr[j], a, b) << ORB_IDX_BITS) + (uint32_t)(base + k);
          m2[j] = min(m2[j], max(m1[j], key));
          m1[j] = min(m1[j], key);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < ORB_QPT; j++) {
    if (qrow[j] < nq) {
      uint4 o;
      const uint32_t mask = (1u << ORB_IDX_BITS) - 1;
      o.x = m1[j] == ABSENT_KEY ? ABSENT_KEY : (m1[j] >> ORB_IDX_BITS);
      o.y = m1[j] == ABSENT_KEY ? 0xFFFFFFFFu : (m1[j] & mask);
      o.z = m2[j] == ABSENT_KEY ? ABSENT_KEY : (m2[j] >> ORB_IDX_BITS);
      o.w = m2[j] == ABSENT_KEY ? 0xFFFFFFFFu : (m2[j] & mask);
      part[((size_t)pair * n_split + split) * nq + qrow[j]] = o;
    }
  }
}

void launch_orb_knn2(const uint8_t* q, int nq, const PairArgs* pairs, int n_pairs, int n_split,
                     uint4* part, cudaStream_t s) {
  if (nq <= 0 || n_pairs <= 0) return;
  dim3 grid((nq + ORB_QB - 1) / ORB_QB, n_split, n_pairs);
  orb_knn2_kernel<<<grid, ORB_THREADS, 0, s>>>(reinterpret_cast<const uint4*>(q), nq, pairs,
                                               n_split, part);
  COUNT_LAUNCH();
}
