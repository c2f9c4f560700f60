// triangulateB200.cpp -- the reference-side replacement of triangulationWrapper
// (src/mainModule/triangulation/triangulate.cpp:57-72): same signature, the
// per-point loop of reconstructPointsFor3D (:17-55) runs on the B200 through slamb200_triangulate.
// reconstruct() (:74-89) and the rest of triangulate.cpp stay as they are and call this.
//
// Built against the real OpenCV inside the reference tree, or against host/cv_shim.h
// (-DSLAMB200_CV_SHIM) for the tests of this repository.
#ifdef SLAMB200_CV_SHIM
#include "cv_shim.h"
#else
#include <opencv2/core.hpp>
#endif

#include <stdexcept>
#include <string>
#include <vector>

#include "slamb200.h"

slamb200_ctx* slamb200HostContext();  // the process-wide context of featureMatchingB200.cpp

void triangulationWrapper(cv::InputArray projPoints1, cv::InputArray projPoints2, const cv::Mat& matr1,
                          const cv::Mat& matr2, cv::OutputArray points4D) {
  // the reference hands N x 2 CV_64F Mats made from vector<Point2f> (triangulate.cpp:82-85): the
  // values are floats widened to double, so narrowing them back is exact
  cv::Mat p1, p2, P1, P2;
  projPoints1.getMat().convertTo(p1, CV_32F);
  projPoints2.getMat().convertTo(p2, CV_32F);
  matr1.convertTo(P1, CV_64F);
  matr2.convertTo(P2, CV_64F);
  CV_Assert(p1.rows == p2.rows && p1.cols == 2 && p2.cols == 2 && P1.isContinuous() && P2.isContinuous());
  if (!p1.isContinuous()) p1 = p1.clone();
  if (!p2.isContinuous()) p2 = p2.clone();
  points4D.create(4, p1.rows, CV_64F);  // four rows: X, Y, Z, W (triangulate.cpp:66)
  cv::Mat out = points4D.getMat();
  const int rc = slamb200_triangulate(slamb200HostContext(), P1.ptr<double>(), P2.ptr<double>(),
                                      p1.ptr<float>(), p2.ptr<float>(), p1.rows, out.ptr<double>(), nullptr);
  if (rc != SLAMB200_OK) throw std::runtime_error(std::string("slamb200_triangulate: ") + slamb200_last_error());
}
