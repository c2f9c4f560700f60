// ransac_control.h -- host-side control of OpenCV's RANSAC loop (cv::findEssentialMat,
// cv::solvePnPRansac) around the GPU scorers, for the reference-side translation units
// (cameraTranslationB200.cpp, poseEstimationB200.cpp).  Pure C++17, no OpenCV types: the minimal
// solver and the scorer are injected.
//
// The reference's call (src/mainModule/translation/cameraTranslation.cpp:41-46) runs OpenCV's
// RANSACPointSetRegistrator::run with modelPoints = 5, maxIters = 1000: cv::RNG seeded with
// 2^64-1 draws subsets (getSubset: distinct indices, redrawn on collision), every candidate model
// is scored against all matches, a candidate replaces the best iff count > max(best, 4), and
// RANSACUpdateNumIters shrinks the budget.  This header restates exactly that control flow; the
// Python twin (slam_indoor_code_b200/ransac_host.py) is verified bit-for-bit against
// cv2.findEssentialMat (tests/test_ransac_host_logic.py), and tests/test_gpu_host_cpp.py checks
// that both twins draw the same subsets and take the same decisions.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <functional>
#include <vector>

namespace slamb200 {

struct CvRNG {  // cv::RNG (multiply-with-carry)
  uint64_t state = 0xFFFFFFFFFFFFFFFFull;
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a)) + a; }
};

inline void getSubset(CvRNG& rng, int count, int* idx, int modelPoints = 5) {
  for (int i = 0; i < modelPoints; i++) {
    int v;
    bool dup;
    do {
      v = rng.uniform(0, count);
      dup = false;
      for (int j = 0; j < i; j++) dup = dup || idx[j] == v;
    } while (dup);
    idx[i] = v;
  }
}

inline int RANSACUpdateNumIters(double p, double ep, int modelPoints, int maxIters) {
  p = std::fmax(p, 0.); p = std::fmin(p, 1.);
  ep = std::fmax(ep, 0.); ep = std::fmin(ep, 1.);
  double num = std::fmax(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, modelPoints);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : (int)std::nearbyint(num / denom);
}

// RANSACPointSetRegistrator::run for count > modelPoints, any model size (ND doubles per model):
//   solve(idx[modelPoints], models)   appends the candidate models of one minimal sample
//   score(models, H, counts)          inlier counts of H models against all points (the GPU call)
// Iterations are scored speculatively `chunk` at a time; the update rule is replayed in order and
// work past the final budget is discarded, so the outcome equals the sequential loop's.
// Returns the best model in `best` (ND doubles); false when no model has > modelPoints-1 inliers.
// `subsets` (optional) receives every drawn subset, for tests.
inline bool ransacRun(int count, int modelPoints, int ND, double prob, int maxIters, int chunk,
                      const std::function<void(const int*, std::vector<double>&)>& solve,
                      const std::function<void(const double*, int, int32_t*)>& score,
                      double* best, int* iterations = nullptr,
                      std::vector<int>* subsets = nullptr) {
  if (count < modelPoints || modelPoints > 16) return false;
  CvRNG rng;
  int niters = maxIters, it = 0, maxGood = 0;
  bool found = false;
  std::vector<double> models;
  std::vector<int> owner;
  std::vector<int32_t> counts;
  while (it < niters) {
    const int n_it = chunk < niters - it ? chunk : niters - it;
    models.clear();
    owner.clear();
    for (int j = 0; j < n_it; j++) {
      int idx[16];
      getSubset(rng, count, idx, modelPoints);
      if (subsets) subsets->insert(subsets->end(), idx, idx + modelPoints);
      const size_t before = models.size() / ND;
      solve(idx, models);
      for (size_t k = before; k < models.size() / ND; k++) owner.push_back(it + j);
    }
    const int H = (int)(models.size() / ND);
    if (H > 0) {
      counts.assign((size_t)H, 0);
      score(models.data(), H, counts.data());
      for (int h = 0; h < H; h++) {
        if (owner[(size_t)h] >= niters) break;  // the budget shrank below this iteration
        const int floor_ = modelPoints - 1;
        if (counts[(size_t)h] > (maxGood > floor_ ? maxGood : floor_)) {
          maxGood = counts[(size_t)h];
          for (int k = 0; k < ND; k++) best[k] = models[(size_t)h * ND + k];
          found = true;
          niters = RANSACUpdateNumIters(prob, (double)(count - maxGood) / count, modelPoints, niters);
        }
      }
    }
    it += n_it;
  }
  if (iterations) *iterations = it < niters ? it : niters;
  return found;
}

// findEssentialMat's instance: five-point samples, 9-double models, maxIters 1000.
inline bool ransacEssential(int count, double prob, int maxIters, int chunk,
                            const std::function<void(const int*, std::vector<double>&)>& solve,
                            const std::function<void(const double*, int, int32_t*)>& score,
                            double best[9], int* iterations = nullptr,
                            std::vector<int>* subsets = nullptr) {
  return ransacRun(count, 5, 9, prob, maxIters, chunk, solve, score, best, iterations, subsets);
}

}  // namespace slamb200
