// poseEstimationB200.cpp -- the reference-side replacement of the call
//     solvePnPRansac(oldSpatialPointsForNewFrame, newFrameFeatureCoords, calibrationMatrix,
//                    distortionCoeffs, rotationVector, motion)
// at src/mainModule/cycleProcessing/mainCycle.cpp:155-159.  Every other argument
// is at its OpenCV default there: useExtrinsicGuess=false, iterationsCount=100,
// reprojectionError=8.0, confidence=0.99, SOLVEPNP_ITERATIVE.
//
// The minimal solver (EPnP on five correspondences; P3P when exactly four points are given) and
// the final Levenberg-Marquardt refit on the inliers stay OpenCV-CPU; the RANSAC control is
// ransac_control.h; every candidate pose is scored against all correspondences on the B200
// (slamb200_score_pnp).  Result: the rvec, tvec and inlier list of OpenCV's own loop -- the Python
// twin of this unit (slam_indoor_code_b200/pnp_ransac.py) is held bit-exact to cv2.solvePnPRansac
// in tests/test_gpu_pnp.py.
//
// Built against the real OpenCV (calib3d) inside the reference tree, or against host/cv_shim.h
// (-DSLAMB200_CV_SHIM), where cv::solvePnP and cv::Rodrigues are injected by the test harness
// (tests/test_gpu_host_cpp.py plugs in cv2's).
#ifdef SLAMB200_CV_SHIM
#include "cv_shim.h"
#else
#include <opencv2/calib3d.hpp>
#include <opencv2/core.hpp>
#endif

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ransac_control.h"
#include "slamb200.h"

bool solvePnPRansacB200(slamb200_ctx* ctx, const std::vector<cv::Point3f>& objectPoints,
                        const std::vector<cv::Point2f>& imagePoints, const cv::Mat& cameraMatrix,
                        const cv::Mat& distCoeffs, cv::Mat& rvec, cv::Mat& tvec,
                        std::vector<int>* inliers = nullptr, int iterationsCount = 100,
                        float reprojectionError = 8.0f, double confidence = 0.99) {
  const int N = (int)objectPoints.size();
  CV_Assert(N >= 4 && N == (int)imagePoints.size());
  cv::Mat Kd, dist;
  cameraMatrix.convertTo(Kd, CV_64F);
  if (!distCoeffs.empty()) distCoeffs.reshape(1, 1).convertTo(dist, CV_64F);
  const double K4[4] = {Kd.at<double>(0, 0), Kd.at<double>(1, 1), Kd.at<double>(0, 2), Kd.at<double>(1, 2)};
  const int nDist = dist.empty() ? 0 : dist.cols;
  const double* distPtr = nDist ? dist.ptr<double>() : nullptr;
  const int modelPoints = N == 4 ? 4 : 5;
  const int method = N == 4 ? cv::SOLVEPNP_P3P : cv::SOLVEPNP_EPNP;
  if (N == modelPoints) {  // cv::solvePnPRansac's short-cut: the minimal solver on everything
    if (!cv::solvePnP(objectPoints, imagePoints, Kd, dist, rvec, tvec, false, method)) return false;
    if (inliers) { inliers->resize(N); for (int i = 0; i < N; i++) (*inliers)[i] = i; }
    return true;
  }
  // every candidate keeps its rvec next to the rotation matrix the GPU scores with
  struct Cand { double pose[12]; double r[3]; };
  std::vector<Cand> seen;
  auto solve = [&](const int* idx, std::vector<double>& models) {
    std::vector<cv::Point3f> o(modelPoints);
    std::vector<cv::Point2f> m(modelPoints);
    for (int i = 0; i < modelPoints; i++) { o[i] = objectPoints[idx[i]]; m[i] = imagePoints[idx[i]]; }
    cv::Mat r, t, R;
    if (!cv::solvePnP(o, m, Kd, dist, r, t, false, method)) return;
    cv::Rodrigues(r, R);
    Cand c;
    for (int i = 0; i < 9; i++) c.pose[i] = R.at<double>(i / 3, i % 3);
    for (int i = 0; i < 3; i++) { c.pose[9 + i] = t.at<double>(i); c.r[i] = r.at<double>(i); }
    seen.push_back(c);
    models.insert(models.end(), c.pose, c.pose + 12);
  };
  auto score = [&](const double* models, int H, int32_t* counts) {
    int32_t best = -1;
    const int rc = slamb200_score_pnp(ctx, (const float*)objectPoints.data(), (const float*)imagePoints.data(),
                                      N, K4, distPtr, nDist, models, H, reprojectionError, modelPoints,
                                      counts, &best, nullptr, nullptr);
    if (rc != SLAMB200_OK) throw std::runtime_error(std::string("slamb200_score_pnp: ") + slamb200_last_error());
  };
  double best[12];
  if (!slamb200::ransacRun(N, modelPoints, 12, confidence, iterationsCount, 8, solve, score, best)) return false;
  std::vector<uchar> mask(N);
  int32_t c = 0, b = -1;
  if (slamb200_score_pnp(ctx, (const float*)objectPoints.data(), (const float*)imagePoints.data(), N, K4,
                         distPtr, nDist, best, 1, reprojectionError, modelPoints, &c, &b, mask.data(),
                         nullptr) != SLAMB200_OK)
    throw std::runtime_error(std::string("slamb200_score_pnp: ") + slamb200_last_error());
  std::vector<cv::Point3d> oi;
  std::vector<cv::Point2d> mi;
  std::vector<int> inl;
  for (int i = 0; i < N; i++)
    if (mask[i]) { oi.emplace_back(objectPoints[i]); mi.emplace_back(imagePoints[i]); inl.push_back(i); }
  // the refit starts from the RANSAC model (cv::solvePnPRansac forces useExtrinsicGuess on for
  // SOLVEPNP_ITERATIVE)
  const Cand* win = nullptr;
  for (const Cand& cd : seen)
    if (!std::memcmp(cd.pose, best, sizeof(best))) { win = &cd; break; }
  rvec = (cv::Mat_<double>(3, 1) << win->r[0], win->r[1], win->r[2]);
  tvec = (cv::Mat_<double>(3, 1) << best[9], best[10], best[11]);
  const bool ok = cv::solvePnP(oi, mi, Kd, dist, rvec, tvec, true, cv::SOLVEPNP_ITERATIVE);
  if (inliers) inliers->swap(inl);
  return ok;
}
