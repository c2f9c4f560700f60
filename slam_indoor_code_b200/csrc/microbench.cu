// microbench.cu -- pipe-rate probes used for the ORB (POPC) and RANSAC (FP64) roofline
// denominators in bench.py / DESIGN.md.  Debug exports, not part of the public header.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void popc_rate_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8];
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = seed + threadIdx.x * 8 + j;
  uint32_t acc = 0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      a[j] = __popc(a[j] ^ (uint32_t)i) + a[j];   // 1 POPC + 1 LOP3 + 1 IADD per step
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) acc += a[j];
  if (acc == 0x12345678u) out[0] = acc;
}

__global__ void fp64_rate_kernel(double* out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = seed + threadIdx.x * 8 + j;
  const double m = 1.0000000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = __dadd_rn(__dmul_rn(a[j], m), c);  // 1 DMUL + 1 DADD
  }
  double acc = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) acc += a[j];
  if (acc == 0.123456789) out[0] = acc;
}

__global__ void sad4_rate_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { a[j] = seed + threadIdx.x * 8 + j; b[j] = seed * 2654435761u + j; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++)   // 1 VABSDIFF4.U8.ACC per step, eight independent chains
      asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[j]) : "r"(b[j]), "r"((uint32_t)i));
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) acc += a[j];
  if (acc == 0x12345678u) out[0] = acc;
}

// returns giga-ops/s: which = 0 POPC (popc instructions), 1 FP64 (DMUL + DADD instructions),
// 2 byte-wise SAD (VABSDIFF4.U8.ACC instructions)
extern "C" double slamb200_dbg_pipe_rate(int which) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  void* buf = nullptr;
  cudaMalloc(&buf, 64);
  const int iters = 4096, threads = 256, blocks = sms * 8;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(a);
    if (which == 0) popc_rate_kernel<<<blocks, threads>>>((uint32_t*)buf, iters, 17u + rep);
    else if (which == 2) sad4_rate_kernel<<<blocks, threads>>>((uint32_t*)buf, iters, 17u + rep);
    else fp64_rate_kernel<<<blocks, threads>>>((double*)buf, iters, 1.0 + rep);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(buf);
  const double ops = (double)blocks * threads * iters * 8 * (which == 1 ? 2 : 1);
  return ops / (best * 1e-3) / 1e9;
}
