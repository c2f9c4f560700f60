// featureMatchingB200.cpp -- the reference-side binding of libslamb200: a third translation unit
// behind src/mainModule/featureMatching/featureMatching.h, next to featureMatchingCPU.cpp and
// featureMatchingCUDA.cpp (the reference's CMakeLists.txt:56-67 compiles exactly one of them; a
// USE_B200 option adds this one).  Same three symbols, same signatures, same
// behaviour: extractDescriptor keeps SIFT on OpenCV-CPU and computes ORB descriptors on the B200,
// matchFeatures is one C-ABI call that runs knnMatch(k=2) + getGoodMatches on the B200.
//
// Build against the real OpenCV inside the reference tree, or against host/cv_shim.h
// (-DSLAMB200_CV_SHIM) where OpenCV's headers are not installed -- the shim declares exactly the
// members this file touches.
#ifdef SLAMB200_CV_SHIM
#include "cv_shim.h"
#else
#include <opencv2/core.hpp>
#include <opencv2/opencv.hpp>
#include "../../config/config.h"
#include "featureMatching.h"
#include "featureMatchingCommon.h"
// knnMatcherDistance exactly as getGoodMatches reads it (featureMatchingCommon.cpp:42)
static double knnMatcherDistance() {
  return configService.getValue<double>(ConfigFieldEnum::FM_KNN_DISTANCE);
}
#endif

#include <cstring>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "slamb200.h"

using namespace cv;

namespace {

// One context per process, created on first use on device 0 (the reference only checks that a
// CUDA device exists, main.cpp:30-38).  The C ABI is re-entrant, so the `threadsCount` matcher
// threads of batch.cpp:181-201 share it.
slamb200_ctx* context() {
  static slamb200_ctx* ctx = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (slamb200_init(0, &ctx) != SLAMB200_OK)
      throw std::runtime_error(std::string("slamb200_init: ") + slamb200_last_error());
  });
  return ctx;
}

struct DescHandle {
  slamb200_desc* d = nullptr;
  ~DescHandle() { if (d) slamb200_free_desc(context(), d); }
};

int descKindOf(int matcherType) {
  switch (matcherType) {
    case SIFT_BF:
    case SIFT_FLANN:
      return SLAMB200_DESC_F32X128;
    case ORB_BF:
      return SLAMB200_DESC_U8X32;
    default:
      throw std::exception();  // featureMatchingCPU.cpp:36-37
  }
}

// The reference's two builds disagree on useFM-SIFT-BF: NORM_L2 on the CPU (featureMatchingCPU.cpp:28),
// NORM_L1 in the OpenCV-CUDA build (featureMatchingCUDA.cpp:28).  The CPU meaning is the default;
// -DSLAMB200_CUDA_BUILD_COMPAT selects the CUDA build's.
inline int abiMatcher(int matcherType) {
#ifdef SLAMB200_CUDA_BUILD_COMPAT
  if (matcherType == SIFT_BF) return SLAMB200_SIFT_BF_L1;
#endif
  return matcherType;
}

void upload(const Mat& desc, int kind, DescHandle& h) {
  const int want = kind == SLAMB200_DESC_F32X128 ? CV_32F : CV_8U;
  const int cols = kind == SLAMB200_DESC_F32X128 ? 128 : 32;
  if (!desc.empty() && (desc.type() != want || desc.cols != cols))
    throw std::runtime_error("descriptor Mat type/width does not fit the matcher");  // cv::Exception in OpenCV
  const int rc = slamb200_upload_desc_packed(context(), kind, desc.empty() ? nullptr : desc.data,
                                      desc.empty() ? 0 : desc.rows, desc.empty() ? 0 : (size_t)desc.step,
                                      &h.d);
  if (rc != SLAMB200_OK) throw std::runtime_error(std::string("slamb200_upload_desc: ") + slamb200_last_error());
}

}  // namespace

// the other B200 units of the drop-in (cameraTranslationB200.cpp, poseEstimationB200.cpp,
// triangulateB200.cpp) share the process-wide context
slamb200_ctx* slamb200HostContext() { return context(); }

/*
 * @param prevDesc [in]  query descriptors (previous frame)
 * @param curDesc [in]   train descriptors (candidate frame)
 * @param matches [out]  good matches, ascending queryIdx (cleared first, like getGoodMatches)
 * @param matcherType [in] 0 - sift_bf, 1 - sift_flann, 2 - orb_bf
 * Replaces featureMatchingCPU.cpp:17-43.
 */
static void matchFeatures(Mat& prevDesc, Mat& curDesc, std::vector<DMatch>& matches, int extractorType) {
  const int kind = descKindOf(extractorType);
  DescHandle q, t;
  upload(prevDesc, kind, q);
  upload(curDesc, kind, t);
  matches.clear();
  const int cap = prevDesc.empty() ? 0 : prevDesc.rows;
  if (cap == 0) return;
  static_assert(sizeof(DMatch) == sizeof(slamb200_dmatch), "cv::DMatch layout");
  matches.resize((size_t)cap);
  int n = 0;
  const int rc = slamb200_match_pair(context(), abiMatcher(extractorType), q.d, t.d, knnMatcherDistance(),
                                     reinterpret_cast<slamb200_dmatch*>(matches.data()), cap, &n);
  if (rc != SLAMB200_OK) {
    matches.clear();
    throw std::runtime_error(std::string("slamb200_match_pair: ") + slamb200_last_error());
  }
  matches.resize((size_t)n);
}

// cv::ORB::create()->compute(frame, features, desc) on the B200: same descriptors,
// same pruning of `features` (keypoints within 31 px of the border are erased, order kept).
static void computeOrbB200(Mat& frame, std::vector<KeyPoint>& features, Mat& desc) {
  if (frame.empty() || features.empty()) {  // Feature2D::compute: nothing to describe
    desc.release();
    if (frame.empty()) features.clear();
    return;
  }
  if (frame.depth() != CV_8U || (frame.channels() != 1 && frame.channels() != 3))
    throw std::runtime_error("ORB: CV_8UC1 or CV_8UC3 frame expected");  // cv::Exception in OpenCV
  const int n = (int)features.size();
  std::vector<float> k((size_t)3 * n);
  for (int i = 0; i < n; i++) {
    k[3 * i] = features[i].pt.x;
    k[3 * i + 1] = features[i].pt.y;
    k[3 * i + 2] = features[i].angle;
  }
  std::vector<unsigned char> keep((size_t)n), rows((size_t)n * 32);
  int kept = 0;
  const int rc = slamb200_orb_compute(context(), frame.data, frame.rows, frame.cols, frame.channels(),
                                      (size_t)frame.step, k.data(), n, keep.data(), rows.data(), &kept,
                                      nullptr);
  if (rc != SLAMB200_OK) throw std::runtime_error(std::string("slamb200_orb_compute: ") + slamb200_last_error());
  int w = 0;
  for (int i = 0; i < n; i++)
    if (keep[i]) features[w++] = features[i];
  features.resize((size_t)w);
  if (kept == 0) {
    desc.release();
    return;
  }
  desc.create(kept, 32, CV_8U);
  for (int r = 0; r < kept; r++) memcpy(desc.data + (size_t)r * desc.step, rows.data() + (size_t)r * 32, 32);
}

// featureMatchingCPU.cpp:45-66.  SIFT stays OpenCV-CPU (its float image pipeline has no exact
// definition across OpenCV builds); ORB runs on the device.
void extractDescriptor(Mat& frame, std::vector<KeyPoint>& features, int matcherType, Mat& desc) {
  cv::Ptr<cv::DescriptorExtractor> extractor;
  switch (matcherType) {
    case SIFT_BF:
    case SIFT_FLANN:
      extractor = cv::SIFT::create();
      break;
    case ORB_BF:
#ifndef SLAMB200_ORB_ON_CPU
      computeOrbB200(frame, features, desc);
      return;
#else
      extractor = cv::ORB::create();
      break;
#endif
    default:
      throw std::exception();
  }
  extractor->compute(frame, features, desc);
}

// featureMatching.h:29-36
void matchFramesPairFeatures(Mat& firstFrame, Mat& secondFrame, std::vector<KeyPoint>& firstFeatures,
                             std::vector<KeyPoint>& secondFeatures, int matcherType,
                             std::vector<DMatch>& matches) {
  Mat firstDescriptor;
  extractDescriptor(firstFrame, firstFeatures, matcherType, firstDescriptor);
  matchFramesPairFeatures(firstDescriptor, secondFrame, secondFeatures, matcherType, matches);
}

// featureMatching.h:47-53
void matchFramesPairFeatures(Mat& firstFrameDescriptor, Mat& secondFrame,
                             std::vector<KeyPoint>& secondFeatures, int matcherType,
                             std::vector<DMatch>& matches) {
  Mat secondDescriptor;
  extractDescriptor(secondFrame, secondFeatures, matcherType, secondDescriptor);
  matchFeatures(firstFrameDescriptor, secondDescriptor, matches, matcherType);
}

// Optional fast path for the batch window (batch.cpp:101-226): one query
// descriptor against every batch element's descriptor in ONE call.  The per-frame descriptor sets
// are uploaded once; `allMatches[i]` receives element i's good matches.
void matchFramesBatchFeatures(Mat& firstFrameDescriptor, std::vector<Mat>& batchDescriptors,
                              int matcherType, std::vector<std::vector<DMatch>>& allMatches) {
  const int kind = descKindOf(matcherType);
  DescHandle q;
  upload(firstFrameDescriptor, kind, q);
  std::vector<DescHandle> t(batchDescriptors.size());
  std::vector<const slamb200_desc*> tp(batchDescriptors.size());
  for (size_t i = 0; i < batchDescriptors.size(); i++) {
    upload(batchDescriptors[i], kind, t[i]);
    tp[i] = t[i].d;
  }
  const int P = (int)batchDescriptors.size();
  const int cap = firstFrameDescriptor.empty() ? 0 : firstFrameDescriptor.rows;
  allMatches.assign((size_t)P, std::vector<DMatch>());
  if (P == 0 || cap == 0) return;
  std::vector<slamb200_dmatch> out((size_t)P * cap);
  std::vector<int> n((size_t)P, 0);
  const int rc = slamb200_match_batch(context(), abiMatcher(matcherType), q.d, tp.data(), P, knnMatcherDistance(),
                                      out.data(), cap, n.data());
  if (rc != SLAMB200_OK) throw std::runtime_error(std::string("slamb200_match_batch: ") + slamb200_last_error());
  for (int p = 0; p < P; p++) {
    const DMatch* src = reinterpret_cast<const DMatch*>(out.data() + (size_t)p * cap);
    allMatches[(size_t)p].assign(src, src + n[(size_t)p]);
  }
}

// The rule by which both search variants of the reference pick the new good frame from a matched
// batch (batch.cpp:120-148 single-threaded, :283-307 after the matcher threads): walk the batch
// from its last element down to skipFramesFromBatchHead; an element is good when it has at least
// requiredMatchedPointsCount matches and at least as many as the good element so far; with
// useFirstFitInBatch the first good element ends the walk.  Returns the index or FRAME_NOT_FOUND
// (-1, batch.h:6).  The comparisons are the reference's: size_t against int, i.e. a negative
// requirement converts to a huge unsigned value and nothing is good.
int selectGoodFrameFromMatchCounts(const std::vector<size_t>& matched, int requiredMatchedPointsCount,
                                   bool useFirstFitInBatch, int skipFramesFromBatchHead) {
  int goodIndex = -1;
  size_t goodSize = 0;   // goodMatches.size() of an empty vector
  for (int batchIndex = (int)matched.size() - 1; batchIndex >= skipFramesFromBatchHead; batchIndex--) {
    if (batchIndex < 0) break;   // a negative skip count would read before the batch in the reference
    const size_t m = matched[(size_t)batchIndex];
    if (m >= (size_t)requiredMatchedPointsCount && m >= goodSize) {
      goodIndex = batchIndex;
      goodSize = m;
      if (useFirstFitInBatch) break;
    }
  }
  return goodIndex;
}

// The whole search of findGoodFramesFromBatch* (batch.cpp:101-226) on descriptors in ONE matcher
// call: every element of the batch is matched against the previous frame's descriptor
// (speculatively, like the reference's multi-threaded variant), then the reference's rule picks
// the good frame.  allMatches[i] = element i's matches (BatchElement::matches).
int findGoodFrameFromBatchDescriptors(Mat& previousDescriptor, std::vector<Mat>& batchDescriptors,
                                      int matcherType, int requiredMatchedPointsCount,
                                      bool useFirstFitInBatch, int skipFramesFromBatchHead,
                                      std::vector<std::vector<DMatch>>& allMatches) {
  matchFramesBatchFeatures(previousDescriptor, batchDescriptors, matcherType, allMatches);
  std::vector<size_t> matched(allMatches.size());
  for (size_t i = 0; i < allMatches.size(); i++) matched[i] = allMatches[i].size();
  return selectGoodFrameFromMatchCounts(matched, requiredMatchedPointsCount, useFirstFitInBatch,
                                        skipFramesFromBatchHead);
}
