// cv_shim.h -- the OpenCV declarations the B200 host units touch (featureMatchingB200.cpp,
// fastExtractorB200.cpp, cameraTranslationB200.cpp, poseEstimationB200.cpp, triangulateB200.cpp),
// for building and testing those translation units in an image without OpenCV's C++ headers.
// Layouts follow OpenCV 4.x (cv::DMatch, cv::KeyPoint, cv::Point*, cv::Mat's data / rows / cols /
// step / type members).  The three CPU solvers the units call (the 5-point essential-matrix solver
// behind cv::findEssentialMat on five matches, cv::solvePnP, cv::Rodrigues) are INJECTED: the shim
// forwards them to function pointers the test harness sets (tests/test_gpu_host_cpp.py plugs in
// cv2's own implementations), so the C++ control flow runs unchanged around the real arithmetic.
// With the real OpenCV this header is not used.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC3 16  // CV_MAKETYPE(CV_8U, 3)
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)

typedef unsigned char uchar;

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point3f { float x = 0, y = 0, z = 0; Point3f() {} Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {} };
struct Point2d { double x = 0, y = 0; Point2d() {} Point2d(const Point2f& p) : x(p.x), y(p.y) {} };
struct Point3d { double x = 0, y = 0, z = 0; Point3d() {} Point3d(const Point3f& p) : x(p.x), y(p.y), z(p.z) {} };

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};

struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 3.402823466e+38f;
};

template <class T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8U }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };

// A 2-D matrix with cv::Mat's public field names: a view of the caller's memory or a reference
// counted allocation of its own (owner != nullptr, the counterpart of cv::Mat::u).
struct Mat {
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  size_t step = 0;
  int type_ = CV_8U;
  std::shared_ptr<std::vector<unsigned char>> owner;
  Mat() {}
  Mat(int r, int c, int type, void* d, size_t s = 0) : rows(r), cols(c), data((unsigned char*)d), step(s), type_(type) {
    if (step == 0) step = (size_t)c * elemSize();
  }
  Mat(int r, int c, int type) { create(r, c, type); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize1() const { return depth() == CV_64F ? 8 : (depth() == CV_32F ? 4 : 1); }
  size_t elemSize() const { return elemSize1() * (size_t)channels(); }
  bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
  void release() { rows = cols = 0; data = nullptr; step = 0; owner.reset(); }
  void create(int r, int c, int type) {
    if (owner && rows == r && cols == c && type_ == type) return;   // cv::Mat::create: same shape, same buffer
    type_ = type;
    owner = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elemSize());
    rows = r; cols = c; step = (size_t)c * elemSize();
    data = owner->empty() ? nullptr : owner->data();
  }
  template <class T> T& at(int r, int c = 0) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> const T& at(int r, int c = 0) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; r++) memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
    return m;
  }
  // single-channel view with `r` rows over the same elements (continuous matrices only)
  Mat reshape(int cn, int r) const {
    CV_Assert(cn == 1 && isContinuous() && r > 0);
    Mat m = *this;
    const size_t total = (size_t)rows * cols * channels();
    m.type_ = depth();
    m.rows = r;
    m.cols = (int)(total / (size_t)r);
    m.step = (size_t)m.cols * m.elemSize();
    return m;
  }
  void convertTo(Mat& dst, int rtype) const {
    Mat out;
    if (!empty()) {
      out.create(rows, cols, (rtype & 7) | (type_ & ~7));
      const int n = cols * channels();
      for (int r = 0; r < rows; r++)
        for (int c = 0; c < n; c++) {
          double v;
          const unsigned char* s = data + (size_t)r * step;
          if (depth() == CV_64F) v = reinterpret_cast<const double*>(s)[c];
          else if (depth() == CV_32F) v = reinterpret_cast<const float*>(s)[c];
          else v = s[c];
          unsigned char* d = out.data + (size_t)r * out.step;
          if (out.depth() == CV_64F) reinterpret_cast<double*>(d)[c] = v;
          else if (out.depth() == CV_32F) reinterpret_cast<float*>(d)[c] = (float)v;
          else d[c] = (unsigned char)v;
        }
    }
    dst = out;
  }
};

template <class T> struct Mat_ : Mat {
  Mat_(int r, int c) { create(r, c, DataType<T>::type); }
};
template <class T> struct MatCommaInitializer_ {
  Mat_<T> m;
  size_t i = 0;
  MatCommaInitializer_(const Mat_<T>& m_, T v) : m(m_) { put(v); }
  MatCommaInitializer_& operator,(T v) { put(v); return *this; }
  void put(T v) { reinterpret_cast<T*>(m.data)[i++] = v; }
  operator Mat() const { return m; }
};
template <class T> MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T v) { return MatCommaInitializer_<T>(m, v); }

struct _InputArray {
  const Mat* m;
  _InputArray(const Mat& mm) : m(&mm) {}
  Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
struct _OutputArray {
  Mat* m;
  _OutputArray(Mat& mm) : m(&mm) {}
  void create(int r, int c, int t) const { m->create(r, c, t); }
  Mat getMat() const { return *m; }
};
typedef const _OutputArray& OutputArray;

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

// ---- injected CPU solvers (calib3d) -----------------------------------------------------------
enum { RANSAC = 8 };
enum { SOLVEPNP_ITERATIVE = 0, SOLVEPNP_EPNP = 1, SOLVEPNP_P3P = 2 };
namespace shim {
// five matches (2 x 5 x {x, y} floats), K row-major 3x3 -> up to 10 candidates of 9 doubles; returns their count
typedef int (*FivePointFn)(const float* a, const float* b, const double* K9, double* out90);
// n correspondences as doubles (obj n x 3, img n x 2), K 3x3, dist (n_dist) -> rvec[3], tvec[3] (in/out when
// use_guess); returns 1 on success
typedef int (*SolvePnPFn)(const double* obj, const double* img, int n, const double* K9, const double* dist,
                          int n_dist, double* rvec, double* tvec, int use_guess, int method);
typedef void (*RodriguesFn)(const double* r3, double* R9);
inline FivePointFn& five_point() { static FivePointFn f = nullptr; return f; }
inline SolvePnPFn& solve_pnp() { static SolvePnPFn f = nullptr; return f; }
inline RodriguesFn& rodrigues() { static RodriguesFn f = nullptr; return f; }
}  // namespace shim

// cv::findEssentialMat on exactly five matches: the stacked 3k x 3 candidates of the 5-point solver
inline Mat findEssentialMat(const std::vector<Point2f>& a, const std::vector<Point2f>& b, const Mat& K, int, double, double) {
  CV_Assert(a.size() == 5 && b.size() == 5 && shim::five_point() != nullptr);
  Mat Kd;
  K.convertTo(Kd, CV_64F);
  double K9[9], out[90];
  for (int i = 0; i < 9; i++) K9[i] = Kd.at<double>(i / 3, i % 3);
  const int k = shim::five_point()(&a[0].x, &b[0].x, K9, out);
  Mat E;
  if (k <= 0) return E;
  E.create(3 * k, 3, CV_64F);
  memcpy(E.data, out, sizeof(double) * 9 * (size_t)k);
  return E;
}
template <class P3, class P2>
inline bool solvePnP(const std::vector<P3>& obj, const std::vector<P2>& img, const Mat& K, const Mat& dist, Mat& rvec,
                     Mat& tvec, bool useExtrinsicGuess, int method) {
  CV_Assert(obj.size() == img.size() && shim::solve_pnp() != nullptr);
  const int n = (int)obj.size();
  std::vector<double> o((size_t)3 * n), m((size_t)2 * n);
  for (int i = 0; i < n; i++) {
    o[3 * i] = obj[i].x; o[3 * i + 1] = obj[i].y; o[3 * i + 2] = obj[i].z;
    m[2 * i] = img[i].x; m[2 * i + 1] = img[i].y;
  }
  double K9[9], r[3] = {0, 0, 0}, t[3] = {0, 0, 0};
  for (int i = 0; i < 9; i++) K9[i] = K.at<double>(i / 3, i % 3);
  if (useExtrinsicGuess)
    for (int i = 0; i < 3; i++) { r[i] = rvec.at<double>(i); t[i] = tvec.at<double>(i); }
  const int ok = shim::solve_pnp()(o.data(), m.data(), n, K9, dist.empty() ? nullptr : dist.ptr<double>(),
                                   dist.empty() ? 0 : dist.cols, r, t, useExtrinsicGuess ? 1 : 0, method);
  if (!ok) return false;
  rvec = (Mat_<double>(3, 1) << r[0], r[1], r[2]);
  tvec = (Mat_<double>(3, 1) << t[0], t[1], t[2]);
  return true;
}
inline void Rodrigues(const Mat& r, Mat& R) {
  CV_Assert(shim::rodrigues() != nullptr);
  double r3[3] = {r.at<double>(0), r.at<double>(1), r.at<double>(2)}, R9[9];
  shim::rodrigues()(r3, R9);
  R.create(3, 3, CV_64F);
  memcpy(R.data, R9, sizeof(R9));
}

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

// misc/ChronoTimer.h and misc/IOmisc.h (logStreams.timeStream): the per-call timing lines of
// featureMatchingCUDA.cpp:101,107 go to a stream the harness can read back.
#include <chrono>
#include <sstream>
class ChronoTimer {
  std::chrono::high_resolution_clock::time_point start, lastPoint;
 public:
  ChronoTimer() : start(std::chrono::high_resolution_clock::now()), lastPoint(start) {}
  void updateLastPoint() { lastPoint = std::chrono::high_resolution_clock::now(); }
  void printLastPointDelta(const std::string& message, std::ostream& stream) {
    stream << message
           << std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - lastPoint).count()
           << std::endl;
  }
};
struct LogFilesStreams { std::ostringstream timeStream; };
extern LogFilesStreams logStreams;
