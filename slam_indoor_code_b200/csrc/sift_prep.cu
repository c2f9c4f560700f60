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
//   bf16lo  bf16(v - bf16(v)): the low half of a two-term bf16 split, used by the general-float
//         tensor-core path (zero for integer-valued rows);  nrmf  |row|^2 as fp32 (fp64 sum)
//   flags[0] != 0 when some value is not an integer in [0,255] or some |row|^2 >= 2^20
//   flags[2] = max |row|^2 of the set (fp32 bits; error bound of the general path)
// Padding rows [n, n_pad) get zero descriptors and a +inf augmentation (the accumulator becomes
// +inf, never NaN: the partner's entry is 1) so they are never candidates.  One warp per row, one float4 per lane.
#include "common.cuh"
#include <cuda_fp8.h>

// Everything derived from one row's 128 values (lane l holds values 4l .. 4l+3 in v).
__device__ __forceinline__ void prep_row(const float4 v, const int row, const int lane, const int n,
                                         __nv_bfloat16* __restrict__ bf16, __nv_bfloat16* __restrict__ augq,
                                         __nv_bfloat16* __restrict__ augt, uint8_t* __restrict__ u8,
                                         int32_t* __restrict__ nrm2, __nv_bfloat16* __restrict__ bf16lo,
                                         float* __restrict__ nrmf, int32_t* __restrict__ flags) {
  const float x[4] = {v.x, v.y, v.z, v.w};
  int bad = 0, ss = 0;
  uint32_t packed = 0;
  double sd = (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
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
  {
    // low half of the two-term split: v - bf16(v) is exact in fp32, then rounded to bf16
    const float2 h0 = __bfloat1622float2(lo), h1 = __bfloat1622float2(hi);
    __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - h0.x, v.y - h0.y);
    __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - h1.x, v.w - h1.y);
    uint2 pl;
    pl.x = *reinterpret_cast<uint32_t*>(&l0);
    pl.y = *reinterpret_cast<uint32_t*>(&l1);
    reinterpret_cast<uint2*>(bf16lo + (size_t)row * 128)[lane] = pl;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, off);
    bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    sd += __shfl_xor_sync(0xffffffffu, sd, off);
  }
  if (ss >= (1 << 20)) bad = 1;
  const float nf = (float)sd;  // == ss exactly for integer-valued rows
  if (lane == 0) {
    nrm2[row] = ss;
    nrmf[row] = nf;
    if (row < n) {
      if (bad) atomicOr(&flags[0], 1);
      atomicMax(&flags[2], __float_as_int(nf));
    }
  }
  // augmentation blocks: lanes 0..15 write column `lane` of each
  if (lane < 16) {
    float h, m, l;
    if (row < n) {
      const float half = 0.5f * nf;  // three bf16 pieces hold all 24 bits
      const __nv_bfloat16 bh = __float2bfloat16_rn(half);
      const float r1 = half - __bfloat162float(bh);
      const __nv_bfloat16 bm = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(bm);
      h = __bfloat162float(bh); m = __bfloat162float(bm); l = r2;
    } else {
      h = __int_as_float(0x7f800000); m = 0.f; l = 0.f;  // +inf: padding rows are infinitely far
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

// fp32 source rows (device memory, or page-locked host memory read over PCIe): one warp per row,
// one float4 per lane -- a 512-byte request per warp.
__global__ void __launch_bounds__(256)
sift_prep_kernel(const float* __restrict__ src, size_t src_stride, int n, int n_pad,
                 float* __restrict__ f32, __nv_bfloat16* __restrict__ bf16,
                 __nv_bfloat16* __restrict__ augq, __nv_bfloat16* __restrict__ augt,
                 uint8_t* __restrict__ u8, int32_t* __restrict__ nrm2,
                 __nv_bfloat16* __restrict__ bf16lo, float* __restrict__ nrmf,
                 int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < n) v = reinterpret_cast<const float4*>(src + (size_t)row * src_stride)[lane];
  if (src != f32 || row >= n) reinterpret_cast<float4*>(f32 + (size_t)row * 128)[lane] = v;
  prep_row(v, row, lane, n, bf16, augq, augt, u8, nrm2, bf16lo, nrmf, flags);
}

// Rows already narrowed to bytes (slamb200_upload_desc_packed: the host verified that they are
// integers in [0,255]); src points at n x 128 bytes, typically page-locked host memory read
// straight over PCIe.  A warp takes FOUR consecutive rows with one 512-byte request (lane l the
// 16 bytes at 16 l: a row per eight lanes) and then prepares them one after the other, the bytes
// spread by shuffles -- with one row (128 bytes) per request the link ran at half its rate.
__global__ void __launch_bounds__(256)
sift_prep_u8_kernel(const uint8_t* __restrict__ src, int n, int n_pad,
                    float* __restrict__ f32, __nv_bfloat16* __restrict__ bf16,
                    __nv_bfloat16* __restrict__ augq, __nv_bfloat16* __restrict__ augt,
                    uint8_t* __restrict__ u8, int32_t* __restrict__ nrm2,
                    __nv_bfloat16* __restrict__ bf16lo, float* __restrict__ nrmf,
                    int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4;
  if (row0 >= n_pad) return;
  uint4 w = make_uint4(0, 0, 0, 0);
  if (row0 + (lane >> 3) < n) w = reinterpret_cast<const uint4*>(src + (size_t)row0 * 128)[lane];
#pragma unroll
  for (int rr = 0; rr < 4; rr++) {
    const int row = row0 + rr;
    if (row >= n_pad) break;   // warp-uniform
    const int from = rr * 8 + (lane >> 2);
    const uint32_t w0 = __shfl_sync(0xffffffffu, w.x, from), w1 = __shfl_sync(0xffffffffu, w.y, from);
    const uint32_t w2 = __shfl_sync(0xffffffffu, w.z, from), w3 = __shfl_sync(0xffffffffu, w.w, from);
    const uint32_t word = (lane & 3) == 0 ? w0 : (lane & 3) == 1 ? w1 : (lane & 3) == 2 ? w2 : w3;
    const float4 v = make_float4((float)(word & 255u), (float)((word >> 8) & 255u), (float)((word >> 16) & 255u),
                                 (float)(word >> 24));
    reinterpret_cast<float4*>(f32 + (size_t)row * 128)[lane] = v;
    prep_row(v, row, lane, n, bf16, augq, augt, u8, nrm2, bf16lo, nrmf, flags);
  }
}

void launch_sift_prep(const float* src, size_t src_stride_floats, int n, int n_pad, float* f32,
                      __nv_bfloat16* bf16, __nv_bfloat16* augq, __nv_bfloat16* augt, uint8_t* u8,
                      int32_t* nrm2, __nv_bfloat16* bf16lo, float* nrmf, int32_t* flags,
                      cudaStream_t s) {
  if (n_pad <= 0) return;
  const int rows_per_block = 8;
  sift_prep_kernel<<<(n_pad + rows_per_block - 1) / rows_per_block, 256, 0, s>>>(
      src, src_stride_floats, n, n_pad, f32, bf16, augq, augt, u8, nrm2, bf16lo, nrmf, flags);
  COUNT_LAUNCH();
}

void launch_sift_prep_u8(const uint8_t* src_u8, int n, int n_pad, float* f32, __nv_bfloat16* bf16,
                         __nv_bfloat16* augq, __nv_bfloat16* augt, uint8_t* u8, int32_t* nrm2,
                         __nv_bfloat16* bf16lo, float* nrmf, int32_t* flags, cudaStream_t s) {
  if (n_pad <= 0) return;
  const int rows_per_block = 32;   // 8 warps x 4 rows
  sift_prep_u8_kernel<<<(n_pad + rows_per_block - 1) / rows_per_block, 256, 0, s>>>(
      src_u8, n, n_pad, f32, bf16, augq, augt, u8, nrm2, bf16lo, nrmf, flags);
  COUNT_LAUNCH();
}


// ---- ORB rows as tcgen05 operands -------------------------------------------------------------
// Hamming(q, t) over 256 bits = |q - t|^2 over 256 values in {0, 1}, so the SIFT candidate kernel
// applies as it stands once the bits are spread to bytes: e4m3 (0x00 / 0x38 = 1.0) for
// tcgen05.mma.kind::f8f6f4, 256 bytes per row -- byte for byte the shape of a bf16 SIFT row, so
// the tensor maps, the shared-memory stages and the UMMA descriptors are shared.  The
// augmentation blocks carry popcount/2 as three e4m3 pieces (16 k + j + 0.5 b: four significant
// bits each, exact): query side [h m l 1 1 1 0..], train side [1 1 1 h m l 0..], 32 bytes per
// row in the no-swizzle core-matrix order (8 rows x 16 bytes, the two K halves 128 bytes apart).
// e4m3 has no infinity: padding rows get 448 (real accumulators are <= 128), which the merge pass
// treats as "no such column".  One warp per row, one descriptor byte per lane.
__device__ __forceinline__ uint8_t e4m3_small(int v) {   // integers 0..15 and 16 * (0..8), exact
  return (uint8_t)__nv_cvt_float_to_fp8((float)v, __NV_SATFINITE, __NV_E4M3);
}
__global__ void __launch_bounds__(256)
orb_tc_prep_kernel(const uint8_t* __restrict__ u8, int n, int n_pad, uint8_t* __restrict__ e4,
                   uint8_t* __restrict__ augq, uint8_t* __restrict__ augt) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  const uint32_t b = row < n ? u8[(size_t)row * 32 + lane] : 0u;
  uint2 o;
  o.x = ((b & 1u) ? 0x38u : 0u) | ((b & 2u) ? 0x3800u : 0u) | ((b & 4u) ? 0x380000u : 0u) |
        ((b & 8u) ? 0x38000000u : 0u);
  o.y = ((b & 16u) ? 0x38u : 0u) | ((b & 32u) ? 0x3800u : 0u) | ((b & 64u) ? 0x380000u : 0u) |
        ((b & 128u) ? 0x38000000u : 0u);
  reinterpret_cast<uint2*>(e4 + (size_t)row * 256)[lane] = o;
  int pc = __popc(b);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, off);
  uint8_t h, m, l;
  if (row < n) {
    h = e4m3_small((pc >> 5) << 4);   // popcount / 2 = 16 * (pc >> 5) + ((pc >> 1) & 15) + 0.5 * (pc & 1)
    m = e4m3_small((pc >> 1) & 15);
    l = (pc & 1) ? 0x30 : 0x00;       // 0.5
  } else {
    h = 0x7E; m = 0; l = 0;           // 448, the largest e4m3 value
  }
  const uint8_t one = 0x38;
  uint8_t aq = 0, at = 0;
  if (lane == 0) { aq = h; at = one; }
  if (lane == 1) { aq = m; at = one; }
  if (lane == 2) { aq = l; at = one; }
  if (lane == 3) { aq = one; at = h; }
  if (lane == 4) { aq = one; at = m; }
  if (lane == 5) { aq = one; at = l; }
  const size_t at_off = (size_t)(row >> 3) * 256 + (size_t)(lane >> 4) * 128 + (row & 7) * 16 + (lane & 15);
  augq[at_off] = aq;
  augt[at_off] = at;
}

void launch_orb_tc_prep(const uint8_t* u8, int n, int n_pad, uint8_t* e4, uint8_t* augq,
                        uint8_t* augt, cudaStream_t s) {
  if (n_pad <= 0) return;
  orb_tc_prep_kernel<<<(n_pad + 7) / 8, 256, 0, s>>>(u8, n, n_pad, e4, augq, augt);
  COUNT_LAUNCH();
}
