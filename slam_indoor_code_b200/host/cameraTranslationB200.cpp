// cameraTranslationB200.cpp -- the reference-side replacement of the single call
//     findEssentialMat(points1, points2, K, RANSAC, RPRANSACProb, RPRANSACThreshold, mask)
// at src/mainModule/translation/cameraTranslation.cpp:41-46.  The 5-point minimal solver stays
// OpenCV-CPU (cv::findEssentialMat on exactly five matches returns the stacked candidates), the
// RANSAC control is ransac_control.h, every candidate is scored on the B200.  Result: the same E
// and the same N x 1 uchar mask as OpenCV's own loop.
//
// Built against the real OpenCV (calib3d) inside the reference tree, or against host/cv_shim.h
// (-DSLAMB200_CV_SHIM), where the 5-point solver is injected by the test harness
// (tests/test_gpu_host_cpp.py plugs in cv2's): the whole call then returns cv2.findEssentialMat's
// E and mask bit for bit.
#ifdef SLAMB200_CV_SHIM
#include "cv_shim.h"
#else
#include <opencv2/calib3d.hpp>
#include <opencv2/core.hpp>
#endif

#include <stdexcept>
#include <string>
#include <vector>

#include "ransac_control.h"
#include "slamb200.h"

cv::Mat findEssentialMatB200(slamb200_ctx* ctx, const std::vector<cv::Point2f>& points1,
                             const std::vector<cv::Point2f>& points2, const cv::Mat& K, double prob,
                             double threshold, cv::Mat& mask) {
  const int N = (int)points1.size();
  cv::Mat Kd;
  K.convertTo(Kd, CV_64F);
  const double K4[4] = {Kd.at<double>(0, 0), Kd.at<double>(1, 1), Kd.at<double>(0, 2), Kd.at<double>(1, 2)};
  auto solve = [&](const int* idx, std::vector<double>& models) {
    std::vector<cv::Point2f> a(5), b(5);
    for (int i = 0; i < 5; i++) { a[i] = points1[idx[i]]; b[i] = points2[idx[i]]; }
    cv::Mat E = cv::findEssentialMat(a, b, Kd, cv::RANSAC, 0.999, 1.0);  // 3k x 3 stacked candidates
    if (E.empty()) return;
    for (int r = 0; r + 2 < E.rows; r += 3)
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) models.push_back(E.at<double>(r + i, j));
  };
  auto score = [&](const double* models, int H, int32_t* counts) {
    int32_t best = -1;
    const int rc = slamb200_score_essential(ctx, (const float*)points1.data(), (const float*)points2.data(),
                                            N, K4, models, H, threshold, counts, &best, nullptr, nullptr);
    if (rc != SLAMB200_OK) throw std::runtime_error(std::string("slamb200_score_essential: ") + slamb200_last_error());
  };
  double best[9];
  if (!slamb200::ransacEssential(N, prob, 1000, 32, solve, score, best)) return cv::Mat();
  mask.create(N, 1, CV_8U);
  int32_t c = 0, b = -1;
  slamb200_score_essential(ctx, (const float*)points1.data(), (const float*)points2.data(), N, K4, best, 1,
                           threshold, &c, &b, mask.data, nullptr);
  return cv::Mat(3, 3, CV_64F, best).clone();
}
