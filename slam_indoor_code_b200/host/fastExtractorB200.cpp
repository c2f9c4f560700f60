// fastExtractorB200.cpp -- the reference-side replacement of fastExtractor
// (src/mainModule/featureExtraction/fastExtractor.cpp:7-13, declared in fastExtractor.h:19-21;
// SURVEY.md 8f-3): same signature, same keypoints -- cv::KeyPoint(x, y, 7, -1, response) in
// OpenCV's order -- with the detector running on the B200 through slamb200_fast_detect.  Replace
// fastExtractor.cpp by this file in SOURCES; the callers (cycleProcessing/batch.cpp:245,
// mainCycleInternals.cpp:144) stay as they are.
//
// Built against the real OpenCV inside the reference tree, or against host/cv_shim.h
// (-DSLAMB200_CV_SHIM) for the tests of this repository.
#ifdef SLAMB200_CV_SHIM
#include "cv_shim.h"
namespace cv { struct FastFeatureDetector { enum DetectorType { TYPE_5_8 = 0, TYPE_7_12 = 1, TYPE_9_16 = 2 }; }; }
#else
#include <opencv2/opencv.hpp>
#include "fastExtractor.h"
#endif

#include <stdexcept>
#include <string>
#include <vector>

#include "slamb200.h"

slamb200_ctx* slamb200HostContext();  // the process-wide context of featureMatchingB200.cpp

void fastExtractor(cv::Mat& srcImage, std::vector<cv::KeyPoint>& points, int threshold, bool suppression,
                   cv::FastFeatureDetector::DetectorType type) {
  // the reference only ever asks for the default neighbourhood (fastExtractor.h:21)
  if (type != cv::FastFeatureDetector::TYPE_9_16)
    throw std::invalid_argument("fastExtractor (B200): only TYPE_9_16 is implemented");
  points.clear();
  if (srcImage.empty()) return;
  if (srcImage.depth() != CV_8U || (srcImage.channels() != 1 && srcImage.channels() != 3))
    throw std::invalid_argument("fastExtractor (B200): CV_8UC1 or CV_8UC3 frame expected");
  // first guess for the buffer: one pixel in sixteen; the call reports the real count
  int cap = srcImage.rows * srcImage.cols / 16 + 1024;
  std::vector<float> kp;
  int found = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    kp.resize((size_t)cap * 3);
    const int rc = slamb200_fast_detect(slamb200HostContext(), srcImage.data, srcImage.rows, srcImage.cols,
                                        srcImage.channels(), srcImage.step, threshold, suppression ? 1 : 0,
                                        kp.data(), cap, &found);
    if (rc != SLAMB200_OK) throw std::runtime_error(std::string("slamb200_fast_detect: ") + slamb200_last_error());
    if (found <= cap) break;
    cap = found;
  }
  points.resize((size_t)found);
  for (int i = 0; i < found; i++) {
    cv::KeyPoint& k = points[(size_t)i];
    k.pt.x = kp[3 * (size_t)i];
    k.pt.y = kp[3 * (size_t)i + 1];
    k.size = 7.f;
    k.angle = -1.f;
    k.response = kp[3 * (size_t)i + 2];
    k.octave = 0;
    k.class_id = -1;
  }
}
