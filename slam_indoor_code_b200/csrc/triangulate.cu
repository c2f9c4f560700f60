// triangulate.cu -- batched linear (DLT) triangulation.
//
// Restates reconstructPointsFor3D (src/mainModule/triangulation/triangulate.cpp:17-55): per match
// the 4x4 system  A = [x1*P1[2,:] - P1[0,:]; y1*P1[2,:] - P1[1,:]; x2*P2[2,:] - P2[0,:];
// y2*P2[2,:] - P2[1,:]]  and the right singular vector of its smallest singular value (row 3 of
// Vt from cv::SVD::compute), then convertHomogeneousPointsMatrixToSpatialPointsVector (:91-108):
// (X, Y, Z) * (1 / W).  OpenCV decomposes a 4x4 double matrix with its own one-sided Jacobi
// (JacobiSVDImpl_ on A^T, eps = 10*DBL_EPSILON, <= 30 sweeps, then a descending sort); the same
// rotation sequence runs here, one thread per point, A^T and Vt in registers (fully unrolled),
// every operation an explicit round-to-nearest intrinsic.  hypot/sqrt are CUDA's; OpenCV's come
// from libm, so the last bits can differ: parity is held to 1e-12 relative (tests), not bit-exact.
// The reference runs this once per accepted frame on ~5 000 points (4-6 ms on the CPU); here it
// is one launch, latency bound.
#include "common.cuh"

__device__ __forceinline__ void givens4(double (&a)[4], double (&b)[4], double c, double s,
                                        double& na, double& nb) {
  na = 0; nb = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double t0 = __dadd_rn(__dmul_rn(c, a[k]), __dmul_rn(s, b[k]));
    const double t1 = __dadd_rn(__dmul_rn(-s, a[k]), __dmul_rn(c, b[k]));
    a[k] = t0; b[k] = t1;
    na = __dadd_rn(na, __dmul_rn(t0, t0));
    nb = __dadd_rn(nb, __dmul_rn(t1, t1));
  }
}

__global__ void __launch_bounds__(128)
triangulate_kernel(const float2* __restrict__ pts1, const float2* __restrict__ pts2, int M,
                   TriParams tp, double* __restrict__ points4d, double* __restrict__ points3d) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M) return;
  double At[4][4], Vt[4][4], W[4];
  {
    const float2 a = pts1[p], b = pts2[p];
    const double x[2] = {(double)a.x, (double)b.x}, y[2] = {(double)a.y, (double)b.y};
#pragma unroll
    for (int v = 0; v < 2; v++)
#pragma unroll
      for (int c = 0; c < 4; c++) {
        At[c][2 * v] = __dsub_rn(__dmul_rn(x[v], tp.P[v][8 + c]), tp.P[v][c]);
        At[c][2 * v + 1] = __dsub_rn(__dmul_rn(y[v], tp.P[v][8 + c]), tp.P[v][4 + c]);
      }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double sd = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      sd = __dadd_rn(sd, __dmul_rn(At[i][k], At[i][k]));
      Vt[i][k] = i == k ? 1.0 : 0.0;
    }
    W[i] = sd;
  }
  const double eps = 2.220446049250313e-16 * 10;
  for (int iter = 0; iter < 30; iter++) {
    bool changed = false;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i + 1; j < 4; j++) {
        double a = W[i], b = W[j], pq = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) pq = __dadd_rn(pq, __dmul_rn(At[i][k], At[j][k]));
        if (fabs(pq) <= __dmul_rn(eps, sqrt(__dmul_rn(a, b)))) continue;
        pq = __dmul_rn(pq, 2.0);
        const double beta = __dsub_rn(a, b), gamma = hypot(pq, beta);
        double c, s;
        if (beta < 0) {
          const double delta = __dmul_rn(__dsub_rn(gamma, beta), 0.5);
          s = sqrt(__ddiv_rn(delta, gamma));
          c = __ddiv_rn(pq, __dmul_rn(__dmul_rn(gamma, s), 2.0));
        } else {
          c = sqrt(__ddiv_rn(__dadd_rn(gamma, beta), __dmul_rn(gamma, 2.0)));
          s = __ddiv_rn(pq, __dmul_rn(__dmul_rn(gamma, c), 2.0));
        }
        givens4(At[i], At[j], c, s, a, b);
        W[i] = a; W[j] = b;
        double u0, u1;
        givens4(Vt[i], Vt[j], c, s, u0, u1);
        changed = true;
      }
    if (!changed) break;
  }
  // the right singular vector of the smallest singular value: OpenCV sorts descending with a
  // selection sort that keeps the first of equal values, so row 3 ends up holding the LAST minimum
  // ... of the column norms; find it without moving the rows
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double sd = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) sd = __dadd_rn(sd, __dmul_rn(At[i][k], At[i][k]));
    W[i] = sqrt(sd);
  }
  // replay the selection sort on an index permutation (4 elements: cheap and exactly OpenCV's ties)
  int perm[4] = {0, 1, 2, 3};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    int j = i;
#pragma unroll
    for (int k = i + 1; k < 4; k++)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      const double t = W[i]; W[i] = W[j]; W[j] = t;
      const int ti = perm[i]; perm[i] = perm[j]; perm[j] = ti;
    }
  }
  double v[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    v[k] = perm[3] == 0 ? Vt[0][k] : perm[3] == 1 ? Vt[1][k] : perm[3] == 2 ? Vt[2][k] : Vt[3][k];
#pragma unroll
  for (int k = 0; k < 4; k++) points4d[(size_t)k * M + p] = v[k];
  if (points3d) {
    const double iw = __ddiv_rn(1.0, v[3]);
#pragma unroll
    for (int k = 0; k < 3; k++) points3d[3 * (size_t)p + k] = __dmul_rn(v[k], iw);
  }
}

void launch_triangulate(const float2* pts1, const float2* pts2, int M, const TriParams& tp,
                        double* points4d, double* points3d, cudaStream_t s) {
  if (M <= 0) return;
  triangulate_kernel<<<(M + 127) / 128, 128, 0, s>>>(pts1, pts2, M, tp, points4d, points3d);
  COUNT_LAUNCH();
}
