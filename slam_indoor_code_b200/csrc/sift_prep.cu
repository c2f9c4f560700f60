// sift_prep.cu -- one pass over a freshly uploaded SIFT frame ("upload once per frame").
//
// Input: the caller's CV_32F N x 128 rows as extractDescriptor produces them
// (src/mainModule/featureMatching/featureMatchingCPU.cpp:45-66).  Output, all resident in HBM:
//   f32   contiguous fp32 copy (exact fp32 path, pitch removed)
//   bf16  tcgen05 operand; exact for cv::SIFT's integer-valued rows
//   augq / augt  the 16-column K-augmentation blocks that make the MMA accumulate
//         (|q|^2 + |t|^2)/2 - q.t = d^2/2 directly:  query side [h m l 1 1 1 0..], train side
//         [1 1 1 h m l 0..] with h+m+l = |row|^2/2 split into three bf16 pieces (24 bits, exact)
//   u8    integer copy for the dp4a rerank;  nrm2  integer squared norms
//   flags[0] != 0 when some value is not an integer in [0,255] or some |row|^2 >= 2^20
// Padding rows [n, n_pad) get zero descriptors and a 2^30 augmentation so they are never
// candidates.  One warp per row, one float4 per lane.
#include "common.cuh"

__global__ void __launch_bounds__(256)
sift_prep_kernel(const float* __restrict__ src, size_t src_stride, int n, int n_pad,
                 float* __restrict__ f32, __nv_bfloat16* __restrict__ bf16,
                 __nv_bfloat16* __restrict__ augq, __nv_bfloat16* __restrict__ augt,
                 uint8_t* __restrict__ u8, int32_t* __restrict__ nrm2,
                 int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < n) v = reinterpret_cast<const float4*>(src + (size_t)row * src_stride)[lane];
  if (src != f32 || row >= n) reinterpret_cast<float4*>(f32 + (size_t)row * 128)[lane] = v;
  const float x[4] = {v.x, v.y, v.z, v.w};
  int bad = 0, ss = 0;
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const bool ok = (x[k] >= 0.f) && (x[k] <= 255.f) && (x[k] == rintf(x[k]));
    bad |= !ok;
    const int iv = ok ? (int)x[k] : 0;
    ss += iv * iv;
    packed |= (uint32_t)iv << (8 * k);
  }
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  reinterpret_cast<uint2*>(bf16 + (size_t)row * 128)[lane] = pk;
  reinterpret_cast<uint32_t*>(u8 + (size_t)row * 128)[lane] = packed;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, off);
    bad |= __shfl_xor_sync(0xffffffffu, bad, off);
  }
  if (ss >= (1 << 20)) bad = 1;
  if (lane == 0) {
    nrm2[row] = ss;
    if (bad && row < n) atomicOr(&flags[0], 1);
  }
  // augmentation blocks: lanes 0..15 write column `lane` of each
  if (lane < 16) {
    float h, m, l;
    if (row < n) {
      const float half = 0.5f * (float)ss;  // exact: ss < 2^24
      const __nv_bfloat16 bh = __float2bfloat16_rn(half);
      const float r1 = half - __bfloat162float(bh);
      const __nv_bfloat16 bm = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(bm);
      h = __bfloat162float(bh); m = __bfloat162float(bm); l = r2;
    } else {
      h = 1073741824.f; m = 0.f; l = 0.f;  // 2^30: padding rows are infinitely far
    }
    const float one = 1.f;
    float aq = 0.f, at = 0.f;
    if (lane == 0) { aq = h; at = one; }
    if (lane == 1) { aq = m; at = one; }
    if (lane == 2) { aq = l; at = one; }
    if (lane == 3) { aq = one; at = h; }
    if (lane == 4) { aq = one; at = m; }
    if (lane == 5) { aq = one; at = l; }
    // UMMA no-swizzle K-major core-matrix order: 8-row group -> [k 0..7 of 8 rows][k 8..15 ...]
    const size_t o = (size_t)(row >> 3) * 128 + (size_t)(lane >> 3) * 64 + (row & 7) * 8 + (lane & 7);
    augq[o] = __float2bfloat16_rn(aq);
    augt[o] = __float2bfloat16_rn(at);
  }
}

void launch_sift_prep(const float* src, size_t src_stride_floats, int n, int n_pad, float* f32,
                      __nv_bfloat16* bf16, __nv_bfloat16* augq, __nv_bfloat16* augt, uint8_t* u8,
                      int32_t* nrm2, int32_t* flags, cudaStream_t s) {
  if (n_pad <= 0) return;
  const int rows_per_block = 8;
  sift_prep_kernel<<<(n_pad + rows_per_block - 1) / rows_per_block, 256, 0, s>>>(
      src, src_stride_floats, n, n_pad, f32, bf16, augq, augt, u8, nrm2, flags);
  COUNT_LAUNCH();
}
