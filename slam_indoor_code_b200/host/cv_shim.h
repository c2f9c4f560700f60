// cv_shim.h -- the handful of OpenCV declarations featureMatchingB200.cpp touches, for building
// and testing that translation unit in an image without OpenCV's C++ headers.  Layouts follow
// OpenCV 4.x (cv::DMatch, cv::KeyPoint, cv::Mat's data/rows/cols/step/type members).  With the
// real OpenCV this header is not used.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_8UC3 16  // CV_MAKETYPE(CV_8U, 3)

namespace cv {

struct Point2f { float x = 0, y = 0; };

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};

struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 3.402823466e+38f;
};

// A non-owning (or vector-backed) 2-D matrix view with cv::Mat's public field names.
struct Mat {
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  size_t step = 0;
  int type_ = CV_8U;
  std::shared_ptr<std::vector<unsigned char>> owner;
  Mat() {}
  Mat(int r, int c, int type, void* d, size_t s) : rows(r), cols(c), data((unsigned char*)d), step(s), type_(type) {}
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  void release() { rows = cols = 0; data = nullptr; step = 0; owner.reset(); }
  void create(int r, int c, int type) {
    const size_t esz = (size_t)((type & 7) == CV_32F ? 4 : 1) * (size_t)((type >> 3) + 1);
    owner = std::make_shared<std::vector<unsigned char>>((size_t)r * c * esz);
    rows = r; cols = c; type_ = type; step = (size_t)c * esz;
    data = owner->empty() ? nullptr : owner->data();
  }
};

template <class T> using Ptr = std::shared_ptr<T>;

// extractDescriptor's producers: not available without OpenCV -- the shim build keeps the symbol
// (so the TU links) and reports the missing dependency if it is ever called.
struct DescriptorExtractor {
  virtual ~DescriptorExtractor() {}
  virtual void compute(Mat&, std::vector<KeyPoint>&, Mat&) {
    throw std::runtime_error("cv::Feature2D::compute needs the real OpenCV (shim build)");
  }
};
struct SIFT : DescriptorExtractor { static Ptr<DescriptorExtractor> create() { return std::make_shared<SIFT>(); } };
struct ORB : DescriptorExtractor { static Ptr<DescriptorExtractor> create() { return std::make_shared<ORB>(); } };

}  // namespace cv

// featureMatchingCommon.h:8-12
enum MatcherType { SIFT_BF, SIFT_FLANN, ORB_BF };

// prototypes of featureMatching.h:12-53
void extractDescriptor(cv::Mat& frame, std::vector<cv::KeyPoint>& features, int extractorType, cv::Mat& desc);
void matchFramesPairFeatures(cv::Mat& firstFrame, cv::Mat& secondFrame, std::vector<cv::KeyPoint>& firstFeatures,
                             std::vector<cv::KeyPoint>& secondFeatures, int matcherType,
                             std::vector<cv::DMatch>& matches);
void matchFramesPairFeatures(cv::Mat& firstFrameDescriptor, cv::Mat& secondFrame,
                             std::vector<cv::KeyPoint>& secondFeatures, int matcherType,
                             std::vector<cv::DMatch>& matches);

// configService.getValue<double>(ConfigFieldEnum::FM_KNN_DISTANCE) (featureMatchingCommon.cpp:42);
// the shim build takes it from a settable global.
double knnMatcherDistance();
