// orb_desc.cu -- ORB descriptors of given keypoints on the device.
//
// Replaces cv::ORB::create()->compute(frame, features, desc) as extractDescriptor calls it
// (src/mainModule/featureMatching/featureMatchingCPU.cpp:45-66) on FAST keypoints
// (fastExtractor.cpp:7-13): the descriptors are produced in HBM, where the matcher wants them,
// instead of on the CPU followed by an upload.  Bit-identical to OpenCV (tests):
//   gray   (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   blur   7x7 Gaussian sigma 2, BORDER_REFLECT_101, the float path of sepFilter2D in the order of
//          its FMA code: rows s = k0*p0, s = fma(kj, pj, s); columns c = k3*s3,
//          c = fma(k(3+j), s(3+j) + s(3-j), c); round half even to u8
//   bit k  blur[c + R(p0_k)] < blur[c + R(p1_k)] over the 256 pairs of orb_pattern.h, R = rotation
//          by the keypoint angle in float (x*a - y*b, x*b + y*a, separately rounded products),
//          cvRound to the pixel grid
// The border filter (keypoints whose rounded position is within 31 px of the edge are dropped) and cosf/sinf of the angle run
// on the host (api.cu): they are per-keypoint scalars, and libm's cosf is what OpenCV calls.
#include "common.cuh"
#include "orb_pattern.h"

// global (not __constant__): every lane reads its own eight entries, which the constant cache would
// serialise 32 ways
__device__ __align__(16) int8_t g_orb_pattern[256][4];

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

__global__ void orb_gray_kernel(const uint8_t* __restrict__ img, int rows, int cols, int channels,
                                size_t step, uint8_t* __restrict__ gray) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cols) return;
  const uint8_t* s = img + (size_t)y * step;
  gray[(size_t)y * cols + x] =
      channels == 1 ? s[x]
                    : (uint8_t)((s[3 * x] * 3735 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 9798 + (1 << 14)) >> 15);
}

struct Gauss7 { float k[7]; };

__global__ void orb_blur_rows_kernel(const uint8_t* __restrict__ gray, int rows, int cols, Gauss7 g,
                                     float* __restrict__ rowf) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cols) return;
  const uint8_t* r = gray + (size_t)y * cols;
  float s = __fmul_rn(g.k[0], (float)r[reflect101(x - 3, cols)]);
#pragma unroll
  for (int j = 1; j < 7; j++) s = __fmaf_rn(g.k[j], (float)r[reflect101(x - 3 + j, cols)], s);
  rowf[(size_t)y * cols + x] = s;
}

__global__ void orb_blur_cols_kernel(const float* __restrict__ rowf, int rows, int cols, Gauss7 g,
                                     uint8_t* __restrict__ blur) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cols) return;
  float c = __fmul_rn(g.k[3], rowf[(size_t)y * cols + x]);
#pragma unroll
  for (int j = 1; j < 4; j++)
    c = __fmaf_rn(g.k[3 + j],
                  __fadd_rn(rowf[(size_t)reflect101(y + j, rows) * cols + x],
                            rowf[(size_t)reflect101(y - j, rows) * cols + x]), c);
  const float r = rintf(c);
  blur[(size_t)y * cols + x] = (uint8_t)(r < 0.f ? 0.f : r > 255.f ? 255.f : r);
}

// one warp per keypoint, one descriptor byte per lane
__global__ void __launch_bounds__(256)
orb_desc_kernel(const uint8_t* __restrict__ blur, int cols, const OrbKeypoint* __restrict__ kps,
                int n, int n_pad, uint8_t* __restrict__ desc) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n_pad) return;
  if (i >= n) {  // padding rows of the resident set
    desc[(size_t)i * 32 + lane] = 0;
    return;
  }
  const OrbKeypoint kp = kps[i];
  const uint8_t* center = blur + (size_t)kp.cy * cols + kp.cx;
  // this lane's eight point pairs: 32 bytes, two 16-byte loads
  const uint4 pa = __ldg(reinterpret_cast<const uint4*>(&g_orb_pattern[8 * lane][0]));
  const uint4 pb = __ldg(reinterpret_cast<const uint4*>(&g_orb_pattern[8 * lane + 4][0]));
  const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
  int v = 0;
#pragma unroll
  for (int bit = 0; bit < 8; bit++) {
    const int8_t p[4] = {(int8_t)(pw[bit] & 255u), (int8_t)((pw[bit] >> 8) & 255u),
                         (int8_t)((pw[bit] >> 16) & 255u), (int8_t)(pw[bit] >> 24)};
    int t[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const float px = (float)p[2 * e], py = (float)p[2 * e + 1];
      const float xr = __fsub_rn(__fmul_rn(px, kp.a), __fmul_rn(py, kp.b));
      const float yr = __fadd_rn(__fmul_rn(px, kp.b), __fmul_rn(py, kp.a));
      t[e] = center[__float2int_rn(yr) * cols + __float2int_rn(xr)];
    }
    v |= (t[0] < t[1]) << bit;
  }
  desc[(size_t)i * 32 + lane] = (uint8_t)v;
}

static void orb_gauss7(Gauss7& g) {
  // getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) in double, normalised, cast to float
  double v[7], sum = 0;
  for (int i = 0; i < 7; i++) { const double x = i - 3; v[i] = exp(-0.125 * x * x); sum += v[i]; }
  for (int i = 0; i < 7; i++) g.k[i] = (float)(v[i] * (1. / sum));
}

int orb_pattern_upload() {
  static PerDeviceOnce once;   // a __device__ symbol has one instance per device
  return once.run([] { return cudaMemcpyToSymbol(g_orb_pattern, kOrbPattern, sizeof(kOrbPattern)) == cudaSuccess; }) ? 0 : -1;
}

// cvtColor(BGR2GRAY) alone (the FAST detector works on the gray frame)
void launch_orb_gray(const uint8_t* img, int rows, int cols, int channels, size_t step, uint8_t* gray,
                     cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return;
  dim3 grid((cols + 255) / 256, rows);
  orb_gray_kernel<<<grid, 256, 0, s>>>(img, rows, cols, channels, step, gray);
  COUNT_LAUNCH();
}

// ORB's blur of a frame that is already gray on the device (the FAST detector made it)
void launch_orb_blur_gray(const uint8_t* gray, int rows, int cols, float* rowf, uint8_t* blur, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return;
  Gauss7 g;
  orb_gauss7(g);
  dim3 grid((cols + 255) / 256, rows);
  orb_blur_rows_kernel<<<grid, 256, 0, s>>>(gray, rows, cols, g, rowf);
  orb_blur_cols_kernel<<<grid, 256, 0, s>>>(rowf, rows, cols, g, blur);
  COUNT_LAUNCH(); COUNT_LAUNCH();
}

void launch_orb_blur(const uint8_t* img, int rows, int cols, int channels, size_t step, uint8_t* gray,
                     float* rowf, uint8_t* blur, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return;
  Gauss7 g;
  orb_gauss7(g);
  dim3 grid((cols + 255) / 256, rows);
  orb_gray_kernel<<<grid, 256, 0, s>>>(img, rows, cols, channels, step, gray);
  orb_blur_rows_kernel<<<grid, 256, 0, s>>>(gray, rows, cols, g, rowf);
  orb_blur_cols_kernel<<<grid, 256, 0, s>>>(rowf, rows, cols, g, blur);
  COUNT_LAUNCH(); COUNT_LAUNCH(); COUNT_LAUNCH();
}

void launch_orb_desc(const uint8_t* blur, int cols, const OrbKeypoint* kps, int n, int n_pad,
                     uint8_t* desc, cudaStream_t s) {
  if (n_pad <= 0) return;
  orb_desc_kernel<<<(n_pad + 7) / 8, 256, 0, s>>>(blur, cols, kps, n, n_pad, desc);
  COUNT_LAUNCH();
}
