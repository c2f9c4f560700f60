/*
 * corr_oracle.c -- CPU restatement of the reference's correspondence hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * file's shared object, and there only as the checker or the timed CPU arm.  The product path
 * (libslamb200.so) never links, loads or calls it and has no CPU fallback.
 *
 * What is restated (reference file:line are relative to /root/reference):
 *   - matchFeatures            src/mainModule/featureMatching/featureMatchingCPU.cpp:17-43
 *       DescriptorMatcher BRUTEFORCE (NORM_L2) / BRUTEFORCE_HAMMING, knnMatch(query=prev,
 *       train=cur, k=2).  The arithmetic lives in OpenCV 4.8.0 (README.md:16, un-vendored):
 *       cv::batchDistance -> hal::normL2Sqr_ / normHamming, then the strict-'<' top-K insert.
 *   - getGoodMatches           src/mainModule/featureMatching/featureMatchingCommon.cpp:37-50
 *   - getKeyPointCoordsFromFramePair  featureMatchingCommon.cpp:23-33
 *   - findEssentialMat(RANSAC) scoring as called at
 *                              src/mainModule/translation/cameraTranslation.cpp:41-46
 *       (OpenCV calib3d five-point.cpp EMEstimatorCallback::computeError +
 *        ptsetreg.cpp RANSACPointSetRegistrator::findInliers / run()).
 *   - the "next" rows of SURVEY.md 8f, each introduced by its own comment block below:
 *       solvePnPRansac scoring      src/mainModule/cycleProcessing/mainCycle.cpp:155-159
 *       NORM_L1 kNN                 src/mainModule/featureMatching/featureMatchingCUDA.cpp:28
 *       linear triangulation        src/mainModule/triangulation/triangulate.cpp:17-55, :91-108
 *       ORB descriptors (compute)   src/mainModule/featureMatching/featureMatchingCPU.cpp:45-66
 *       FAST keypoints (detect)     src/mainModule/featureExtraction/fastExtractor.cpp:7-13
 *       SIFT descriptors (compute)  src/mainModule/featureMatching/featureMatchingCPU.cpp:45-66
 *           (working image pinned bit for bit; descriptors pinned to a TOLERANCE, see its block)
 *
 * Parity pin: the reference ships no tests or golden vectors, so this
 * oracle is pinned against the reference's own arithmetic owner, OpenCV, through the cv2 wheel
 * (tests/test_oracle_vs_cv2.py, and the committed cv2-generated fixtures in tests/golden/ made by
 * oracle/gen_golden.py).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction anywhere in here; the one
 * places OpenCV's own code fuses -- ORB's and SIFT's float blurs -- call fmaf explicitly).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float distance;
} oracle_dmatch; /* bit-compatible with cv::DMatch */

int oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ---- L2: hal::normL2Sqr_ of the SSE-baseline OpenCV build -------------------------------
 * Four 4-lane accumulators; element j = 16*i + 4*a + l is added to acc[a][l] with the multiply
 * and the add rounded separately (v_muladd has no FMA on the SSE3 baseline); the accumulators are
 * combined as ((acc0+acc1)+acc2)+acc3 lane-wise and the lanes as (S0+S2)+(S1+S3)
 * (v_reduce_sum); a scalar left-to-right tail handles n % 16. */
static float l2sqr_cv(const float* a, const float* b, int n) {
    float acc[4][4];
    int j = 0;
    float d = 0.f;
    memset(acc, 0, sizeof(acc));
    for (; j <= n - 16; j += 16) {
        for (int v = 0; v < 4; v++)
            for (int l = 0; l < 4; l++) {
                float t = a[j + 4 * v + l] - b[j + 4 * v + l];
                float p = t * t;
                acc[v][l] = acc[v][l] + p;
            }
    }
    {
        float S[4];
        for (int l = 0; l < 4; l++) S[l] = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l];
        d = (S[0] + S[2]) + (S[1] + S[3]);
    }
    for (; j < n; j++) {
        float t = a[j] - b[j];
        d += t * t;
    }
    return d;
}

static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

/* BatchDistInvoker's top-K insert for K=2 on int-compared keys (positive floats compare like
 * ints): strict '<' against the current K-th, shift while strictly greater -> equal distances
 * keep the lower train index first. */
static inline void top2_insert(int32_t d, int32_t j, int32_t* dk, int32_t* ik) {
    if (d < dk[1]) {
        if (dk[0] > d) {
            dk[1] = dk[0]; ik[1] = ik[0];
            dk[0] = d;     ik[0] = j;
        } else {
            dk[1] = d;     ik[1] = j;
        }
    }
}

/* BFMatcher(NORM_L2).knnMatch(Q, T, 2) raw result: idx[q][k] = -1 when absent, dist as float. */
int oracle_l2_knn2(const float* Q, int nq, const float* T, int nt, int dim,
                   int32_t* idx /* nq*2 */, float* dist /* nq*2 */) {
    if (nq < 0 || nt < 0 || dim <= 0) return -1;
#pragma omp parallel for schedule(static)
    for (int q = 0; q < nq; q++) {
        int32_t dk[2] = {f2i(FLT_MAX), f2i(FLT_MAX)};
        int32_t ik[2] = {-1, -1};
        const float* a = Q + (size_t)q * dim;
        for (int t = 0; t < nt; t++) {
            float d = sqrtf(l2sqr_cv(a, T + (size_t)t * dim, dim));
            top2_insert(f2i(d), t, dk, ik);
        }
        idx[2 * q] = ik[0]; idx[2 * q + 1] = ik[1];
        dist[2 * q] = i2f(dk[0]); dist[2 * q + 1] = i2f(dk[1]);
    }
    return 0;
}

/* NORM_L1 (what useFM-SIFT-BF selects in the reference's OpenCV-CUDA build,
 * featureMatchingCUDA.cpp:28).  OpenCV-CUDA cannot run here; the CPU BFMatcher(NORM_L1) is the
 * arithmetic owner: normL1<float,float> sums groups of four,
 * s += |d0| + |d1| + |d2| + |d3| (left to right), in float -- pinned against cv2 bit for bit
 * (tests/test_oracle_vs_cv2.py).  For the integer-valued rows cv::SIFT emits every order is exact. */
static float l1_cv(const float* a, const float* b, int n) {
    float s = 0.f;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        float v0 = a[i] - b[i], v1 = a[i + 1] - b[i + 1], v2 = a[i + 2] - b[i + 2],
              v3 = a[i + 3] - b[i + 3];
        s += fabsf(v0) + fabsf(v1) + fabsf(v2) + fabsf(v3);
    }
    for (; i < n; i++) s += fabsf(a[i] - b[i]);
    return s;
}

int oracle_l1_knn2(const float* Q, int nq, const float* T, int nt, int dim,
                   int32_t* idx /* nq*2 */, float* dist /* nq*2 */) {
    if (nq < 0 || nt < 0 || dim <= 0) return -1;
#pragma omp parallel for schedule(static)
    for (int q = 0; q < nq; q++) {
        int32_t dk[2] = {f2i(FLT_MAX), f2i(FLT_MAX)};
        int32_t ik[2] = {-1, -1};
        const float* a = Q + (size_t)q * dim;
        for (int t = 0; t < nt; t++) top2_insert(f2i(l1_cv(a, T + (size_t)t * dim, dim)), t, dk, ik);
        idx[2 * q] = ik[0]; idx[2 * q + 1] = ik[1];
        dist[2 * q] = i2f(dk[0]); dist[2 * q + 1] = i2f(dk[1]);
    }
    return 0;
}

/* BFMatcher(NORM_HAMMING).knnMatch: int distances = popcount(a ^ b), converted to float in the
 * DMatch. */
int oracle_hamming_knn2(const uint8_t* Q, int nq, const uint8_t* T, int nt, int nbytes,
                        int32_t* idx, float* dist) {
    if (nq < 0 || nt < 0 || nbytes <= 0) return -1;
#pragma omp parallel for schedule(static)
    for (int q = 0; q < nq; q++) {
        int32_t dk[2] = {INT_MAX, INT_MAX};
        int32_t ik[2] = {-1, -1};
        const uint8_t* a = Q + (size_t)q * nbytes;
        for (int t = 0; t < nt; t++) {
            const uint8_t* b = T + (size_t)t * nbytes;
            int32_t d = 0;
            int j = 0;
            for (; j + 8 <= nbytes; j += 8) {
                uint64_t x, y;
                memcpy(&x, a + j, 8); memcpy(&y, b + j, 8);
                d += __builtin_popcountll(x ^ y);
            }
            for (; j < nbytes; j++) d += __builtin_popcount((unsigned)(a[j] ^ b[j]));
            top2_insert(d, t, dk, ik);
        }
        idx[2 * q] = ik[0]; idx[2 * q + 1] = ik[1];
        dist[2 * q] = ik[0] >= 0 ? (float)dk[0] : FLT_MAX;
        dist[2 * q + 1] = ik[1] >= 0 ? (float)dk[1] : FLT_MAX;
    }
    return 0;
}

/* getGoodMatches (featureMatchingCommon.cpp:37-50): rows in ascending queryIdx, skip empty lists,
 * keep [0] iff (double)d0 < ratio * (double)d1 (strict, in double).  The reference reads [1]
 * unchecked when the list has one element (T == 1): undefined behaviour there; defined here as
 * "reject the row" (SURVEY.md appendix A.5). */
int oracle_ratio_test(const int32_t* idx, const float* dist, int nq, double ratio,
                      oracle_dmatch* out, int* n_out) {
    int n = 0;
    for (int q = 0; q < nq; q++) {
        if (idx[2 * q] < 0) continue;      /* empty list */
        if (idx[2 * q + 1] < 0) continue;  /* single-element list: reference UB, defined reject */
        if ((double)dist[2 * q] < ratio * (double)dist[2 * q + 1]) {
            out[n].queryIdx = q;
            out[n].trainIdx = idx[2 * q];
            out[n].imgIdx = 0;
            out[n].distance = dist[2 * q];
            n++;
        }
    }
    *n_out = n;
    return 0;
}

/* getKeyPointCoordsFromFramePair (featureMatchingCommon.cpp:23-33): gather pt of the matched
 * keypoints; kps are (x, y) float pairs. */
int oracle_gather_points(const float* prev_xy, const float* next_xy, const oracle_dmatch* m,
                         int n, float* pts1, float* pts2) {
    for (int i = 0; i < n; i++) {
        pts1[2 * i] = prev_xy[2 * m[i].queryIdx];
        pts1[2 * i + 1] = prev_xy[2 * m[i].queryIdx + 1];
        pts2[2 * i] = next_xy[2 * m[i].trainIdx];
        pts2[2 * i + 1] = next_xy[2 * m[i].trainIdx + 1];
    }
    return 0;
}

/* findEssentialMat's point normalisation (five-point.cpp): points converted to double, then the
 * Mat expression (col - c) / f, which OpenCV's MatExpr algebra folds into one scaled convert:
 * x = u * (1/f) + (-c * (1/f)), every operation rounded separately in double. */
void oracle_normalize_points(const float* pts, int n, double f_x, double f_y, double c_x,
                             double c_y, double* out) {
    double ax = 1.0 / f_x, ay = 1.0 / f_y;
    double bx = -c_x * ax, by = -c_y * ay;
    for (int i = 0; i < n; i++) {
        double u = (double)pts[2 * i], v = (double)pts[2 * i + 1];
        double pu = u * ax, pv = v * ay;
        out[2 * i] = pu + bx;
        out[2 * i + 1] = pv + by;
    }
}

/* EMEstimatorCallback::computeError for one (model, match): Matx33d * Vec3d and Vec3d::dot are
 * "s = 0; s += a*b" left-to-right sums; the error is rounded to float. */
static inline float sampson_cv(const double* E, double x1, double y1, double x2, double y2) {
    double Ex1[3], Etx2[3];
    for (int r = 0; r < 3; r++) {
        double s = 0;
        s += E[3 * r + 0] * x1;
        s += E[3 * r + 1] * y1;
        s += E[3 * r + 2] * 1.0;
        Ex1[r] = s;
    }
    for (int r = 0; r < 3; r++) {
        double s = 0;
        s += E[0 + r] * x2;
        s += E[3 + r] * y2;
        s += E[6 + r] * 1.0;
        Etx2[r] = s;
    }
    double x2tEx1;
    {
        double s = 0;
        s += x2 * Ex1[0];
        s += y2 * Ex1[1];
        s += 1.0 * Ex1[2];
        x2tEx1 = s;
    }
    double a = Ex1[0] * Ex1[0];
    double b = Ex1[1] * Ex1[1];
    double c = Etx2[0] * Etx2[0];
    double d = Etx2[1] * Etx2[1];
    return (float)(x2tEx1 * x2tEx1 / (a + b + c + d));
}

/* Scores H fixed essential-matrix hypotheses against M matches the way the RANSAC loop of
 * findEssentialMat does (cameraTranslation.cpp:41-46 -> ptsetreg.cpp): threshold_px is
 * RPRANSACThreshold, divided by (fx+fy)/2; inlier iff err <= (float)(thr*thr); a model replaces
 * the best iff count > max(best, 4) (strict: the first best wins).  best = -1 when no model has
 * more than 4 inliers.  all_masks (H*M) is optional. */
int oracle_score_essential(const float* pts1, const float* pts2, int M, const double* K4,
                           const double* E, int H, double threshold_px, int32_t* counts,
                           int32_t* best, uint8_t* best_mask, uint8_t* all_masks) {
    if (M < 0 || H < 0) return -1;
    double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    double* n1 = (double*)malloc(sizeof(double) * 2 * (size_t)(M > 0 ? M : 1));
    double* n2 = (double*)malloc(sizeof(double) * 2 * (size_t)(M > 0 ? M : 1));
    if (!n1 || !n2) { free(n1); free(n2); return -2; }
    oracle_normalize_points(pts1, M, fx, fy, cx, cy, n1);
    oracle_normalize_points(pts2, M, fx, fy, cx, cy, n2);
    double thr = threshold_px / ((fx + fy) / 2);
    float t = (float)(thr * thr);
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        const double* Eh = E + 9 * (size_t)h;
        int32_t nz = 0;
        for (int i = 0; i < M; i++) {
            float err = sampson_cv(Eh, n1[2 * i], n1[2 * i + 1], n2[2 * i], n2[2 * i + 1]);
            int f = err <= t;
            if (all_masks) all_masks[(size_t)h * M + i] = (uint8_t)f;
            nz += f;
        }
        counts[h] = nz;
    }
    int32_t bi = -1, bc = 0;
    for (int h = 0; h < H; h++) {
        int32_t lim = bc > 4 ? bc : 4;
        if (counts[h] > lim) { bc = counts[h]; bi = h; }
    }
    *best = bi;
    if (best_mask) {
        if (bi >= 0) {
            const double* Eh = E + 9 * (size_t)bi;
            for (int i = 0; i < M; i++) {
                float err = sampson_cv(Eh, n1[2 * i], n1[2 * i + 1], n2[2 * i], n2[2 * i + 1]);
                best_mask[i] = (uint8_t)(err <= t);
            }
        } else {
            memset(best_mask, 0, (size_t)M);
        }
    }
    free(n1); free(n2);
    return 0;
}

/* Raw per-(hypothesis, match) float errors, for tests that want to look at the margin to the
 * threshold. */
int oracle_sampson_errors(const float* pts1, const float* pts2, int M, const double* K4,
                          const double* E, int H, float* err /* H*M */) {
    double* n1 = (double*)malloc(sizeof(double) * 2 * (size_t)(M > 0 ? M : 1));
    double* n2 = (double*)malloc(sizeof(double) * 2 * (size_t)(M > 0 ? M : 1));
    if (!n1 || !n2) { free(n1); free(n2); return -2; }
    oracle_normalize_points(pts1, M, K4[0], K4[1], K4[2], K4[3], n1);
    oracle_normalize_points(pts2, M, K4[0], K4[1], K4[2], K4[3], n2);
    for (int h = 0; h < H; h++)
        for (int i = 0; i < M; i++)
            err[(size_t)h * M + i] =
                sampson_cv(E + 9 * (size_t)h, n1[2 * i], n1[2 * i + 1], n2[2 * i], n2[2 * i + 1]);
    free(n1); free(n2);
    return 0;
}

/* matchFeatures end to end (featureMatchingCPU.cpp:17-43): matcher 0 SIFT_BF and 1 SIFT_FLANN
 * -> exact L2 (FLANN is approximate in the reference; its ground truth is BF), 2 ORB_BF ->
 * Hamming.  Returns the good-match count through n_out. */
int oracle_match_features(int matcher, const void* q, int nq, const void* t, int nt, double ratio,
                          oracle_dmatch* out, int* n_out) {
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(nq > 0 ? nq : 1));
    float* dist = (float*)malloc(sizeof(float) * 2 * (size_t)(nq > 0 ? nq : 1));
    int rc;
    if (!idx || !dist) { free(idx); free(dist); return -2; }
    if (matcher == 0 || matcher == 1)
        rc = oracle_l2_knn2((const float*)q, nq, (const float*)t, nt, 128, idx, dist);
    else if (matcher == 2)
        rc = oracle_hamming_knn2((const uint8_t*)q, nq, (const uint8_t*)t, nt, 32, idx, dist);
    else if (matcher == 3) /* SIFT_BF as the OpenCV-CUDA build reads it: NORM_L1 */
        rc = oracle_l1_knn2((const float*)q, nq, (const float*)t, nt, 128, idx, dist);
    else
        rc = -3; /* reference: throw std::exception() (featureMatchingCPU.cpp:37) */
    if (rc == 0) rc = oracle_ratio_test(idx, dist, nq, ratio, out, n_out);
    free(idx); free(dist);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * solvePnPRansac scoring (reference call: cycleProcessing/mainCycle.cpp:155-159).
 *
 * OpenCV's PnPRansacCallback::computeError projects every object point with the candidate pose
 * (cv::projectPoints -> cvProjectPoints2Internal: all arithmetic in double, result stored as
 * float) and takes err = (float)norm(Matx21f(image - projected), NORM_L2SQR), whose accumulator
 * type for float is float: s = 0; s += dx*dx; s += dy*dy.  RANSACPointSetRegistrator::findInliers
 * keeps err <= (float)(reprojectionError^2).  The rotation enters as the 3x3 matrix (Rodrigues of
 * rvec is done by the caller: it needs libm's sin/cos).  dist holds up to 12 coefficients in
 * OpenCV order (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4); the tilt terms (tauX, tauY) must be zero,
 * for which the tilt matrix is the identity and "matTilt * (xd0, yd0, 1)" returns xd0, yd0.
 * Pinned against cv2.projectPoints / cv2.solvePnPRansac (tests/test_oracle_vs_cv2.py).
 * ---------------------------------------------------------------------------------------------- */
static void project_cv(const double* R, const double* t, const double* k, double fx, double fy,
                       double cx, double cy, double X, double Y, double Z, float* u, float* v) {
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z ? 1. / z : 1;
    x *= z;
    y *= z;
    double r2 = x * x + y * y;
    double r4 = r2 * r2;
    double r6 = r4 * r2;
    double a1 = 2 * x * y;
    double a2 = r2 + 2 * x * x;
    double a3 = r2 + 2 * y * y;
    double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
    double icdist2 = 1. / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
    double xd0 = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
    double yd0 = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
    *u = (float)(xd0 * fx + cx);
    *v = (float)(yd0 * fy + cy);
}

static float reproj_err_cv(const double* pose, const double* k, const double* K4, const float* o,
                           const float* m) {
    float u, v;
    project_cv(pose, pose + 9, k, K4[0], K4[1], K4[2], K4[3], (double)o[0], (double)o[1],
               (double)o[2], &u, &v);
    float dx = m[0] - u, dy = m[1] - v;
    float s = 0;
    s += dx * dx;
    s += dy * dy;
    return s;
}

/* Projects M object points with one pose (R row-major 9 doubles then t 3 doubles). */
int oracle_project_points(const float* obj, int M, const double* K4, const double* dist,
                          int n_dist, const double* pose, float* uv /* M*2 */) {
    if (M < 0 || n_dist < 0 || n_dist > 12) return -1;
    double k[12] = {0};
    for (int i = 0; i < n_dist; i++) k[i] = dist[i];
    for (int i = 0; i < M; i++)
        project_cv(pose, pose + 9, k, K4[0], K4[1], K4[2], K4[3], (double)obj[3 * i],
                   (double)obj[3 * i + 1], (double)obj[3 * i + 2], &uv[2 * i], &uv[2 * i + 1]);
    return 0;
}

/* Scores H pose hypotheses (H x 12 doubles) against M 3D-2D correspondences.  A model replaces
 * the best iff count > max(best, model_points - 1) (ptsetreg.cpp run()). */
int oracle_score_pnp(const float* obj, const float* img, int M, const double* K4,
                     const double* dist, int n_dist, const double* poses, int H,
                     double reproj_err, int model_points, int32_t* counts, int32_t* best,
                     uint8_t* best_mask, uint8_t* all_masks) {
    if (M < 0 || H < 0 || n_dist < 0 || n_dist > 12) return -1;
    double k[12] = {0};
    for (int i = 0; i < n_dist; i++) k[i] = dist[i];
    float t = (float)(reproj_err * reproj_err);
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        int32_t nz = 0;
        for (int i = 0; i < M; i++) {
            int f = reproj_err_cv(poses + 12 * (size_t)h, k, K4, obj + 3 * i, img + 2 * i) <= t;
            if (all_masks) all_masks[(size_t)h * M + i] = (uint8_t)f;
            nz += f;
        }
        counts[h] = nz;
    }
    int32_t bi = -1, bc = 0;
    for (int h = 0; h < H; h++) {
        int32_t lim = bc > model_points - 1 ? bc : model_points - 1;
        if (counts[h] > lim) { bc = counts[h]; bi = h; }
    }
    *best = bi;
    if (best_mask) {
        for (int i = 0; i < M; i++)
            best_mask[i] = bi < 0 ? 0 : (uint8_t)(reproj_err_cv(poses + 12 * (size_t)bi, k, K4,
                                                               obj + 3 * i, img + 2 * i) <= t);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Linear triangulation (reference: src/mainModule/triangulation/triangulate.cpp:17-55
 * reconstructPointsFor3D, :91-108 convertHomogeneousPointsMatrixToSpatialPointsVector).
 *
 * Per point the reference fills A (4x4): rows x*P[2,:] - P[0,:], y*P[2,:] - P[1,:] for the two views,
 * and takes row 3 of Vt from cv::SVD::compute(A, W, U, Vt).  For a 4x4 double matrix OpenCV runs its
 * own one-sided Jacobi (JacobiSVDImpl_ on At = A^T, eps = 10*DBL_EPSILON, at most 30 sweeps), then
 * sorts the singular values in descending order; restated below and pinned against cv2.SVDecomp.
 * ---------------------------------------------------------------------------------------------- */
static void jacobi_svd4_vt(double At[4][4], double W[4], double Vt[4][4]) {
    const int n = 4, m = 4;
    const double eps = DBL_EPSILON * 10;
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
        W[i] = sd;
        for (int k = 0; k < n; k++) Vt[i][k] = 0;
        Vt[i][i] = 1;
    }
    for (int iter = 0; iter < 30; iter++) {
        int changed = 0;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < m; k++) p += At[i][k] * At[j][k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < m; k++) {
                    double t0 = c * At[i][k] + s * At[j][k];
                    double t1 = -s * At[i][k] + c * At[j][k];
                    At[i][k] = t0; At[j][k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = 1;
                for (int k = 0; k < n; k++) {
                    double t0 = c * Vt[i][k] + s * Vt[j][k];
                    double t1 = -s * Vt[i][k] + c * Vt[j][k];
                    Vt[i][k] = t0; Vt[j][k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < n - 1; i++) {
        int j = i;
        for (int k = i + 1; k < n; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            for (int k = 0; k < 4; k++) {
                t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t;
                t = Vt[i][k]; Vt[i][k] = Vt[j][k]; Vt[j][k] = t;
            }
        }
    }
}

/* Row 3 of Vt of a 4x4 matrix A (row-major), for pinning against cv2.SVDecomp. */
int oracle_svd4_null_vector(const double* A, double* v4, double* w4) {
    double At[4][4], W[4], Vt[4][4];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) At[c][r] = A[4 * r + c];
    jacobi_svd4_vt(At, W, Vt);
    for (int k = 0; k < 4; k++) { v4[k] = Vt[3][k]; if (w4) w4[k] = W[k]; }
    return 0;
}

/* P1, P2: 3x4 row-major doubles (K*[R|t]); pts: M x 2 floats (vector<Point2f>, widened to double as
 * the reference does with convertTo).  points4d: 4 x M row-major (the reference's Mat layout);
 * points3d (optional): M x 3 = (X, Y, Z) * (1 / W), the "pointCol /= w" of the reference. */
int oracle_triangulate(const double* P1, const double* P2, const float* pts1, const float* pts2,
                       int M, double* points4d, double* points3d) {
    if (M < 0) return -1;
    const double* P[2] = {P1, P2};
#pragma omp parallel for schedule(static)
    for (int p = 0; p < M; p++) {
        double At[4][4], W[4], Vt[4][4];
        for (int v = 0; v < 2; v++) {
            const float* pt = v == 0 ? pts1 : pts2;
            double x = (double)pt[2 * p], y = (double)pt[2 * p + 1];
            for (int c = 0; c < 4; c++) {
                At[c][v * 2] = x * P[v][8 + c] - P[v][c];
                At[c][v * 2 + 1] = y * P[v][8 + c] - P[v][4 + c];
            }
        }
        jacobi_svd4_vt(At, W, Vt);
        for (int k = 0; k < 4; k++) points4d[(size_t)k * M + p] = Vt[3][k];
        if (points3d) {
            double iw = 1. / Vt[3][3];
            for (int k = 0; k < 3; k++) points3d[3 * (size_t)p + k] = Vt[3][k] * iw;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * ORB descriptors of given keypoints (reference: extractDescriptor,
 * featureMatchingCPU.cpp:45-66 -> cv::ORB::create()->compute(frame, features, desc) on FAST
 * keypoints, fastExtractor.cpp:7-13).  Restated from measurements of cv2 (tools/
 * recover_orb_pattern.py, tests/test_oracle_vs_cv2.py), OpenCV itself being absent as source:
 *   gray  = (B*3735 + G*19235 + R*9798 + 2^14) >> 15                       (cvtColor BGR2GRAY, 8u)
 *   keep  = keypoints with 31 <= cvRound(x) < cols-31, 31 <= cvRound(y) < rows-31   (runByImageBorder:
 *           Rect::contains(Point(pt)) rounds the Point2f)
 *   blur  = 7x7 Gaussian, sigma 2, BORDER_REFLECT_101, in float: the kernel sum is not 1 within
 *           FLT_EPSILON, so sepFilter2D takes its float path -- rows  s = k0*p0; s = fma(kj, pj, s),
 *           columns  c = k3*s3; c = fma(k(3+j), s(3+j) + s(3-j), c), then round-half-even to u8
 *           (the order of the FMA-dispatched code of the cv2 wheel; 0 differing pixels in 12M)
 *   bit k = blur[c + R(p0_k)] < blur[c + R(p1_k)],  R = rotation by kpt.angle with float
 *           a = cosf, b = sinf:  x' = cvRound(x*a - y*b), y' = cvRound(x*b + y*a); c = cvRound(pt)
 * ---------------------------------------------------------------------------------------------- */
#include "../slam_indoor_code_b200/csrc/orb_pattern.h"

static void orb_gauss7(float k[7]) {
    /* getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) in double, normalised, to float */
    double v[7], sum = 0;
    for (int i = 0; i < 7; i++) { double x = i - 3; v[i] = exp(-0.125 * x * x); sum += v[i]; }
    for (int i = 0; i < 7; i++) k[i] = (float)(v[i] * (1. / sum));
}

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

/* image: rows x cols, channels 1 (gray) or 3 (BGR), `step` bytes per row.  blur_out: rows*cols. */
int oracle_orb_blur(const uint8_t* image, int rows, int cols, int channels, size_t step,
                    uint8_t* gray_out /* optional */, uint8_t* blur_out) {
    if (rows <= 0 || cols <= 0 || (channels != 1 && channels != 3)) return -1;
    uint8_t* gray = (uint8_t*)malloc((size_t)rows * cols);
    float* rowf = (float*)malloc(sizeof(float) * (size_t)rows * cols);
    if (!gray || !rowf) { free(gray); free(rowf); return -2; }
    for (int y = 0; y < rows; y++) {
        const uint8_t* s = image + (size_t)y * step;
        for (int x = 0; x < cols; x++)
            gray[(size_t)y * cols + x] = channels == 1 ? s[x]
                : (uint8_t)((s[3 * x] * 3735 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 9798 + (1 << 14)) >> 15);
    }
    if (gray_out) memcpy(gray_out, gray, (size_t)rows * cols);
    float k[7];
    orb_gauss7(k);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            float s = k[0] * (float)gray[(size_t)y * cols + reflect101(x - 3, cols)];
            for (int j = 1; j < 7; j++)
                s = fmaf(k[j], (float)gray[(size_t)y * cols + reflect101(x - 3 + j, cols)], s);
            rowf[(size_t)y * cols + x] = s;
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            float c = k[3] * rowf[(size_t)y * cols + x];
            for (int j = 1; j < 4; j++)
                c = fmaf(k[3 + j], rowf[(size_t)reflect101(y + j, rows) * cols + x] +
                                       rowf[(size_t)reflect101(y - j, rows) * cols + x], c);
            float r = nearbyintf(c);
            blur_out[(size_t)y * cols + x] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
    free(gray); free(rowf);
    return 0;
}

/* kps: n x {x, y, angle_deg} floats.  keep[n] = 1 for the keypoints compute() keeps (order
 * preserved); desc: one 32-byte row per KEPT keypoint.  Returns the kept count or < 0. */
int oracle_orb_compute(const uint8_t* image, int rows, int cols, int channels, size_t step,
                       const float* kps, int n, uint8_t* keep, uint8_t* desc) {
    if (n < 0) return -1;
    uint8_t* blur = (uint8_t*)malloc((size_t)(rows > 0 ? rows : 1) * (cols > 0 ? cols : 1));
    if (!blur) return -2;
    int rc = oracle_orb_blur(image, rows, cols, channels, step, NULL, blur);
    if (rc) { free(blur); return rc; }
    int kept = 0;
    for (int i = 0; i < n; i++) {
        const float x = kps[3 * i], y = kps[3 * i + 1];
        /* Rect(31, 31, cols-62, rows-62).contains(Point(pt)): the Point2f is rounded first */
        const int rx = (int)lrintf(x), ry = (int)lrintf(y);
        const int ok = rx >= 31 && rx < cols - 31 && ry >= 31 && ry < rows - 31;
        keep[i] = (uint8_t)ok;
        if (!ok) continue;
        float angle = kps[3 * i + 2];
        angle *= (float)(3.141592653589793 / 180.f);
        const float a = cosf(angle), b = sinf(angle);
        const int cx = (int)lrintf(x), cy = (int)lrintf(y);
        uint8_t* d = desc + 32 * (size_t)kept++;
        for (int byte = 0; byte < 32; byte++) {
            int v = 0;
            for (int bit = 0; bit < 8; bit++) {
                const int8_t* p = kOrbPattern[8 * byte + bit];
                int t[2];
                for (int e = 0; e < 2; e++) {
                    const float px = (float)p[2 * e], py = (float)p[2 * e + 1];
                    const float xr = px * a - py * b, yr = px * b + py * a;
                    t[e] = blur[(size_t)(cy + (int)lrintf(yr)) * cols + cx + (int)lrintf(xr)];
                }
                v |= (t[0] < t[1]) << bit;
            }
            d[byte] = (uint8_t)v;
        }
    }
    free(blur);
    return kept;
}

/* ------------------------------------------------------------------------------------------------
 * FAST keypoints (reference: featureExtraction/fastExtractor.cpp:7-13 --
 * FastFeatureDetector::create(threshold, suppression, TYPE_9_16)->detect(frame, points), called
 * from cycleProcessing/batch.cpp:245 and mainCycleInternals.cpp:144 with
 * featureExtractingThreshold).
 *
 * OpenCV features2d (fast.cpp FAST_t<16>, fast_score.cpp cornerScore<16>): a BGR frame is
 * converted with cvtColor(BGR2GRAY); pixel (x, y), 3 <= x < cols-3, 3 <= y < rows-3, is a corner
 * iff at least 9 contiguous pixels of the 16-pixel Bresenham circle are all darker than v - t or
 * all brighter than v + t (strict); its score is the largest threshold for which it stays a corner
 * (min/max over the nine-pixel arcs, below); with suppression a corner survives iff its score is
 * strictly greater than the scores of its 8 neighbours (non-corners score 0).  Keypoints come out
 * row by row, left to right: KeyPoint(x, y, size 7, angle -1, response = score; 0 without
 * suppression).  Pinned against cv2.FastFeatureDetector (tests/test_oracle_vs_cv2.py).
 * ---------------------------------------------------------------------------------------------- */
static const int kFastCircle[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                       {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

static int fast_is_corner(const uint8_t* p, const int* off, int t) {
    const int v = p[0];
    int dark = 0, bright = 0;
    for (int k = 0; k < 25; k++) {   /* the circle, wrapped by 9: every arc of 9 is seen */
        const int x = p[off[k & 15]];
        dark = x < v - t ? dark + 1 : 0;
        bright = x > v + t ? bright + 1 : 0;
        if (dark > 8 || bright > 8) return 1;
    }
    return 0;
}

static int fast_corner_score(const uint8_t* p, const int* off, int threshold) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 25; k++) d[k] = v - p[off[k & 15]];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        a = a < d[k + 3] ? a : d[k + 3];
        if (a <= a0) continue;
        for (int j = 4; j <= 8; j++) a = a < d[k + j] ? a : d[k + j];
        int m = a < d[k] ? a : d[k];
        a0 = a0 > m ? a0 : m;
        m = a < d[k + 9] ? a : d[k + 9];
        a0 = a0 > m ? a0 : m;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        for (int j = 3; j <= 5; j++) b = b > d[k + j] ? b : d[k + j];
        if (b >= b0) continue;
        for (int j = 6; j <= 8; j++) b = b > d[k + j] ? b : d[k + j];
        int m = b > d[k] ? b : d[k];
        b0 = b0 < m ? b0 : m;
        m = b > d[k + 9] ? b : d[k + 9];
        b0 = b0 < m ? b0 : m;
    }
    return -b0 - 1;
}

/* image: rows x cols, channels 1 (gray) or 3 (BGR), `step` bytes per row.  kp: up to cap rows of
 * {x, y, response}.  Returns the number of keypoints found (may exceed cap: only cap are written)
 * or < 0.  score_out (optional, rows*cols): the score map before suppression (0 = not a corner). */
int oracle_fast_detect(const uint8_t* image, int rows, int cols, int channels, size_t step,
                       int threshold, int nonmax, float* kp, int cap, uint8_t* score_out) {
    if (rows < 0 || cols < 0 || (channels != 1 && channels != 3)) return -1;
    const size_t npx = (size_t)(rows > 0 ? rows : 1) * (cols > 0 ? cols : 1);
    uint8_t* gray = (uint8_t*)malloc(npx);
    uint8_t* score = (uint8_t*)calloc(npx, 1);
    uint8_t* corner = (uint8_t*)calloc(npx, 1);
    if (!gray || !score || !corner) { free(gray); free(score); free(corner); return -2; }
    for (int y = 0; y < rows; y++) {
        const uint8_t* s = image + (size_t)y * step;
        for (int x = 0; x < cols; x++)
            gray[(size_t)y * cols + x] = channels == 1 ? s[x]
                : (uint8_t)((s[3 * x] * 3735 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 9798 + (1 << 14)) >> 15);
    }
    threshold = threshold < 0 ? 0 : threshold > 255 ? 255 : threshold;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = kFastCircle[k][0] + kFastCircle[k][1] * cols;
#pragma omp parallel for schedule(static)
    for (int y = 3; y < rows - 3; y++)
        for (int x = 3; x < cols - 3; x++) {
            const uint8_t* p = gray + (size_t)y * cols + x;
            if (fast_is_corner(p, off, threshold)) {
                corner[(size_t)y * cols + x] = 1;
                score[(size_t)y * cols + x] = (uint8_t)fast_corner_score(p, off, threshold);
            }
        }
    int n = 0;
    for (int y = 3; y < rows - 3; y++)
        for (int x = 3; x < cols - 3; x++) {
            const size_t o = (size_t)y * cols + x;
            if (!corner[o]) continue;
            const int s = score[o];
            int keep = 1;
            if (nonmax)
                keep = s > score[o - 1] && s > score[o + 1] && s > score[o - cols - 1] && s > score[o - cols] &&
                       s > score[o - cols + 1] && s > score[o + cols - 1] && s > score[o + cols] &&
                       s > score[o + cols + 1];
            if (keep) {
                if (n < cap) { kp[3 * n] = (float)x; kp[3 * n + 1] = (float)y; kp[3 * n + 2] = nonmax ? (float)s : 0.f; }
                n++;
            }
        }
    if (score_out) memcpy(score_out, score, (size_t)rows * cols);
    free(gray); free(score); free(corner);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * SIFT descriptors of given keypoints (reference: extractDescriptor,
 * featureMatchingCPU.cpp:45-66 -> cv::SIFT::create()->compute(frame, features, desc) on FAST
 * keypoints, fastExtractor.cpp:7-13: KeyPoint(x, y, size 7, angle -1, octave 0)).  Restated from
 * OpenCV's published algorithm (SIFT_Impl::detectAndCompute with useProvidedKeypoints,
 * createInitialImage, calcSIFTDescriptor) and pinned to the cv2 wheel by measurement:
 *   gray  = the integer cvtColor formula (as for ORB / FAST), converted to float (scale 1)
 *   base  = GaussianBlur(gray, sigma = sqrtf(1.6^2 - 0.5^2)), 13 taps (cvRound(sigma*8 + 1) | 1),
 *           BORDER_REFLECT_101, float: rows  s = k0*p0; s = fma(kj, pj, s),  columns
 *           c = k6*s; c = fma(k(6+j), s(+j) + s(-j), c)  -- the order of the cv2 wheel's
 *           sepFilter2D (scalar, unfused tails beyond the last whole vector of a row): 0
 *           differing pixels against cv2.GaussianBlur on float frames of any size
 *   every keypoint of octave 0 / layer 0 (the only kind FAST produces) reads that base image:
 *           firstOctave = 0, no image doubling, no further pyramid level is touched
 *   desc  = calcSIFTDescriptor(base, pt, 360 - angle, size/2, d = 4, n = 8): samples in a
 *           (2 radius + 1)^2 window, gradient by central differences, orientation by the degree-7
 *           fastAtan2 polynomial, magnitude, Gaussian weight, trilinear vote into the
 *           6 x 6 x 10 histogram in sample order; circular wrap; L2 norm, clip at 0.2, renormalise
 *           to 512, saturate_cast<uchar>.
 * cv::hal::magnitude32f / exp32f are IPP routines in the wheel and differ from sqrtf / expf in the
 * last ulp on a fraction of inputs, and OpenCV's AVX2 code sums the squared norm in another order:
 * this restatement equals cv2.SIFT.compute on 99.998 % of the ELEMENTS (the rest differ by 1) --
 * a tolerance pin, not a bit-exact one (tests/test_oracle_vs_cv2.py states the bound).
 * ---------------------------------------------------------------------------------------------- */
void oracle_sift_gauss13(float k[13], float* sigma_out) {
    const float sigma = sqrtf(fmaxf(1.6f * 1.6f - 0.5f * 0.5f, 0.01f));
    double v[13], sum = 0;
    for (int i = 0; i < 13; i++) { double x = i - 6; v[i] = exp(-x * x / (2.0 * (double)sigma * (double)sigma)); sum += v[i]; }
    for (int i = 0; i < 13; i++) k[i] = (float)(v[i] * (1. / sum));
    if (sigma_out) *sigma_out = sigma;
}

/* image: rows x cols, channels 1 or 3 (BGR), `step` bytes per row -> base (rows*cols floats) */
int oracle_sift_base(const uint8_t* image, int rows, int cols, int channels, size_t step, float* base) {
    if (rows <= 0 || cols <= 0 || (channels != 1 && channels != 3)) return -1;
    float* gray = (float*)malloc(sizeof(float) * (size_t)rows * cols);
    float* rowf = (float*)malloc(sizeof(float) * (size_t)rows * cols);
    if (!gray || !rowf) { free(gray); free(rowf); return -2; }
    for (int y = 0; y < rows; y++) {
        const uint8_t* s = image + (size_t)y * step;
        for (int x = 0; x < cols; x++)
            gray[(size_t)y * cols + x] = (float)(channels == 1 ? s[x]
                : (uint8_t)((s[3 * x] * 3735 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 9798 + (1 << 14)) >> 15));
    }
    float k[13];
    oracle_sift_gauss13(k, NULL);
    /* The wheel's row filter is 4 pixels wide, its column filter 8; the pixels beyond the last
     * whole vector of a row go through scalar code that multiplies and adds separately (measured:
     * with these two widths 0 pixels differ on frames of any width, with any other pair some do). */
    const int body_r = cols & ~3, body_c = cols & ~7;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            float s = k[0] * gray[(size_t)y * cols + reflect101(x - 6, cols)];
            for (int j = 1; j < 13; j++) {
                const float p = gray[(size_t)y * cols + reflect101(x - 6 + j, cols)];
                s = x < body_r ? fmaf(k[j], p, s) : s + k[j] * p;
            }
            rowf[(size_t)y * cols + x] = s;
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            float c = k[6] * rowf[(size_t)y * cols + x];
            for (int j = 1; j <= 6; j++) {
                const float p = rowf[(size_t)reflect101(y + j, rows) * cols + x] +
                                rowf[(size_t)reflect101(y - j, rows) * cols + x];
                c = x < body_c ? fmaf(k[6 + j], p, c) : c + k[6 + j] * p;
            }
            base[(size_t)y * cols + x] = c;
        }
    free(gray); free(rowf);
    return 0;
}

static float sift_fast_atan2(float y, float x) {   /* cv::fastAtan2, degrees */
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI), p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI), p7 = -0.04432655554792128f * (float)(180 / M_PI);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON); c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON); c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* kps: n x {x, y, size, angle_deg} floats (octave 0, layer 0).  desc: n x 128 floats (integer
 * valued, as cv::SIFT emits them); raw != 0 skips the final rounding (the values before
 * saturate_cast, for tolerance analysis). */
int oracle_sift_compute(const uint8_t* image, int rows, int cols, int channels, size_t step,
                        const float* kps, int n, float* desc, int raw) {
    if (n < 0) return -1;
    float* base = (float*)malloc(sizeof(float) * (size_t)(rows > 0 ? rows : 1) * (cols > 0 ? cols : 1));
    if (!base) return -2;
    int rc = oracle_sift_base(image, rows, cols, channels, step, base);
    if (rc) { free(base); return rc; }
    const int d = 4, nb = 8;
#pragma omp parallel for schedule(dynamic, 16)
    for (int q = 0; q < n; q++) {
        const float px = kps[4 * q], py = kps[4 * q + 1], size = kps[4 * q + 2], kang = kps[4 * q + 3];
        float angle = 360.f - kang;
        if (fabsf(angle - 360.f) < FLT_EPSILON) angle = 0.f;
        const float ori = angle, scl = size * 0.5f;
        const int ptx = (int)lrintf(px), pty = (int)lrintf(py);
        float cos_t = cosf(ori * (float)(M_PI / 180)), sin_t = sinf(ori * (float)(M_PI / 180));
        const float bins_per_rad = nb / 360.f, exp_scale = -1.f / (d * d * 0.5f), hist_width = 3.0f * scl;
        int radius = (int)lrintf(hist_width * 1.4142135623730951f * (d + 1) * 0.5f);
        const int maxr = (int)sqrt((double)cols * cols + (double)rows * rows);
        if (radius > maxr) radius = maxr;
        cos_t /= hist_width;
        sin_t /= hist_width;
        float hist[6 * 6 * 10];
        memset(hist, 0, sizeof(hist));
        for (int i = -radius; i <= radius; i++)
            for (int j = -radius; j <= radius; j++) {
                const float c_rot = j * cos_t - i * sin_t, r_rot = j * sin_t + i * cos_t;
                float rbin = r_rot + d / 2 - 0.5f, cbin = c_rot + d / 2 - 0.5f;
                const int r = pty + i, c = ptx + j;
                if (!(rbin > -1 && rbin < d && cbin > -1 && cbin < d && r > 0 && r < rows - 1 && c > 0 && c < cols - 1))
                    continue;
                const float dx = base[(size_t)r * cols + c + 1] - base[(size_t)r * cols + c - 1];
                const float dy = base[(size_t)(r - 1) * cols + c] - base[(size_t)(r + 1) * cols + c];
                const float w = expf((c_rot * c_rot + r_rot * r_rot) * exp_scale);
                float obin = (sift_fast_atan2(dy, dx) - ori) * bins_per_rad;
                const float mag = sqrtf(dx * dx + dy * dy) * w;
                const int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin);
                int o0 = (int)floorf(obin);
                rbin -= r0; cbin -= c0; obin -= o0;
                if (o0 < 0) o0 += nb;
                if (o0 >= nb) o0 -= nb;
                const float v_r1 = mag * rbin, v_r0 = mag - v_r1;
                const float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11, v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
                const float v111 = v_rc11 * obin, v110 = v_rc11 - v111, v101 = v_rc10 * obin, v100 = v_rc10 - v101;
                const float v011 = v_rc01 * obin, v010 = v_rc01 - v011, v001 = v_rc00 * obin, v000 = v_rc00 - v001;
                const int idx = ((r0 + 1) * (d + 2) + c0 + 1) * (nb + 2) + o0;
                hist[idx] += v000; hist[idx + 1] += v001;
                hist[idx + (nb + 2)] += v010; hist[idx + (nb + 3)] += v011;
                hist[idx + (d + 2) * (nb + 2)] += v100; hist[idx + (d + 2) * (nb + 2) + 1] += v101;
                hist[idx + (d + 3) * (nb + 2)] += v110; hist[idx + (d + 3) * (nb + 2) + 1] += v111;
            }
        float v[128];
        for (int i = 0; i < d; i++)
            for (int j = 0; j < d; j++) {
                const int idx = ((i + 1) * (d + 2) + (j + 1)) * (nb + 2);
                hist[idx] += hist[idx + nb];
                hist[idx + 1] += hist[idx + nb + 1];
                for (int k = 0; k < nb; k++) v[(i * d + j) * nb + k] = hist[idx + k];
            }
        float nrm2 = 0;
        for (int k = 0; k < 128; k++) nrm2 += v[k] * v[k];
        const float thr = sqrtf(nrm2) * 0.2f;
        nrm2 = 0;
        for (int k = 0; k < 128; k++) { const float t = v[k] < thr ? v[k] : thr; v[k] = t; nrm2 += t * t; }
        nrm2 = 512.f / fmaxf(sqrtf(nrm2), FLT_EPSILON);
        for (int k = 0; k < 128; k++) {
            const float t = v[k] * nrm2;
            if (raw) { desc[(size_t)q * 128 + k] = t; continue; }
            long iv = lrintf(t);                       /* saturate_cast<uchar>(float): cvRound, clamp */
            desc[(size_t)q * 128 + k] = (float)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        }
    }
    free(base);
    return 0;
}
