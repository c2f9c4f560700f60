// host_shim_test.cpp -- C entry points over the shim build of featureMatchingB200.cpp so that the
// Python tests can drive the C++ host layer (the static matchFeatures and the batch fast path)
// exactly as the reference's callers would (batch.cpp:127-130).
#include "featureMatchingB200.cpp"
#include "fastExtractorB200.cpp"
#include "ransac_control.h"

static double g_ratio = 0.7;
double knnMatcherDistance() { return g_ratio; }

extern "C" {

void hostshim_set_ratio(double r) { g_ratio = r; }

// matchFeatures(prevDesc, curDesc, matches, type): returns the match count or -1 on an exception.
int hostshim_match_features(const void* q, int nq, size_t q_step, const void* t, int nt, size_t t_step,
                            int type, cv::DMatch* out, int cap) {
  try {
    const bool orb = type == ORB_BF;
    cv::Mat prev(nq, orb ? 32 : 128, orb ? CV_8U : CV_32F, const_cast<void*>(q), q_step);
    cv::Mat cur(nt, orb ? 32 : 128, orb ? CV_8U : CV_32F, const_cast<void*>(t), t_step);
    std::vector<cv::DMatch> m;
    m.push_back(cv::DMatch());  // must be cleared by the callee
    matchFeatures(prev, cur, m, type);
    if ((int)m.size() > cap) return -2;
    for (size_t i = 0; i < m.size(); i++) out[i] = m[i];
    return (int)m.size();
  } catch (...) {
    return -1;
  }
}

int hostshim_match_batch(const void* q, int nq, const void* const* t, const int* nt, int P, int type,
                         cv::DMatch* out, int cap, int* n_out) {
  try {
    const bool orb = type == ORB_BF;
    const int w = orb ? 32 : 128, ty = orb ? CV_8U : CV_32F;
    const size_t step = orb ? 32 : 512;
    cv::Mat prev(nq, w, ty, const_cast<void*>(q), step);
    std::vector<cv::Mat> batch;
    for (int p = 0; p < P; p++) batch.emplace_back(nt[p], w, ty, const_cast<void*>(t[p]), step);
    std::vector<std::vector<cv::DMatch>> all;
    matchFramesBatchFeatures(prev, batch, type, all);
    for (int p = 0; p < P; p++) {
      n_out[p] = (int)all[p].size();
      for (size_t i = 0; i < all[p].size(); i++) out[(size_t)p * cap + i] = all[p][i];
    }
    return 0;
  } catch (...) {
    return -1;
  }
}

// extractDescriptor(frame, features, ORB_BF, desc): returns the number of keypoints left in
// `features` (their x, y go to kept_xy), the descriptor rows to desc_out; -1 on an exception.
int hostshim_extract_orb(const void* frame, int rows, int cols, int channels, size_t step, const float* kps,
                         int n, float* kept_xy, unsigned char* desc_out, int cap) {
  try {
    cv::Mat f(rows, cols, channels == 3 ? CV_8UC3 : CV_8U, const_cast<void*>(frame), step);
    std::vector<cv::KeyPoint> features((size_t)n);
    for (int i = 0; i < n; i++) {
      features[i].pt.x = kps[3 * i];
      features[i].pt.y = kps[3 * i + 1];
      features[i].angle = kps[3 * i + 2];
    }
    cv::Mat desc;
    extractDescriptor(f, features, ORB_BF, desc);
    if ((int)features.size() > cap || desc.rows != (desc.empty() ? 0 : (int)features.size())) return -2;
    for (size_t i = 0; i < features.size(); i++) {
      kept_xy[2 * i] = features[i].pt.x;
      kept_xy[2 * i + 1] = features[i].pt.y;
      memcpy(desc_out + 32 * i, desc.data + i * desc.step, 32);
    }
    return (int)features.size();
  } catch (...) {
    return -1;
  }
}

// The C++ RANSAC control with a scripted solver and scorer: `n_models[i]` models come out of the
// i-th minimal sample, model h of the run scores `scores[h]`.  Returns the iterations run; writes
// every drawn subset and the index of the winning model (or -1).
int hostshim_ransac_run(int count, int model_points, int nd, double prob, int max_iters, int chunk,
                        const int* n_models, int n_samples, const int32_t* scores, int n_scores,
                        int* subsets_out, int subsets_cap, int* best_model) {
  int sample = 0, issued = 0;
  auto solve = [&](const int*, std::vector<double>& models) {
    const int k = sample < n_samples ? n_models[sample] : 1;
    sample++;
    for (int j = 0; j < k; j++) {
      std::vector<double> m((size_t)nd, 0.0);
      m[0] = (double)issued++;   // the model's global ordinal, so the winner can be identified
      models.insert(models.end(), m.begin(), m.end());
    }
  };
  auto score = [&](const double* models, int H, int32_t* counts) {
    for (int h = 0; h < H; h++) {
      const int ord = (int)models[(size_t)h * nd];
      counts[h] = ord < n_scores ? scores[ord] : 0;
    }
  };
  std::vector<double> best((size_t)nd, 0.0);
  int iters = 0;
  std::vector<int> subsets;
  const bool ok = slamb200::ransacRun(count, model_points, nd, prob, max_iters, chunk, solve, score,
                                      best.data(), &iters, &subsets);
  *best_model = ok ? (int)best[0] : -1;
  for (size_t i = 0; i < subsets.size() && (int)i < subsets_cap; i++) subsets_out[i] = subsets[i];
  return iters;
}

int hostshim_ransac_control(int count, double prob, int chunk, const int* n_models, int n_samples,
                            const int32_t* scores, int n_scores, int* subsets_out, int subsets_cap,
                            int* best_model) {
  return hostshim_ransac_run(count, 5, 9, prob, 1000, chunk, n_models, n_samples, scores, n_scores,
                             subsets_out, subsets_cap, best_model);
}
}

extern "C" {
// fastExtractor(frame, points, threshold, suppression): returns the number of keypoints, their
// {x, y, response, size, angle} rows in out (cap rows); -1 on an exception.
int hostshim_fast_extractor(const void* frame, int rows, int cols, int channels, size_t step, int threshold,
                            int suppression, float* out, int cap) {
  try {
    cv::Mat f(rows, cols, channels == 3 ? CV_8UC3 : CV_8U, const_cast<void*>(frame), step);
    std::vector<cv::KeyPoint> points(3);   // must be replaced by the callee
    fastExtractor(f, points, threshold, suppression != 0, cv::FastFeatureDetector::TYPE_9_16);
    for (size_t i = 0; i < points.size() && (int)i < cap; i++) {
      out[5 * i] = points[i].pt.x; out[5 * i + 1] = points[i].pt.y; out[5 * i + 2] = points[i].response;
      out[5 * i + 3] = points[i].size; out[5 * i + 4] = points[i].angle;
    }
    return (int)points.size();
  } catch (...) {
    return -1;
  }
}
}

extern "C" {
// selectGoodFrameFromMatchCounts over plain arrays (host logic only: no GPU needed)
int hostshim_select_good_frame(const long long* matched, int n, int required, int first_fit, int skip_head) {
  std::vector<size_t> m((size_t)n);
  for (int i = 0; i < n; i++) m[(size_t)i] = (size_t)matched[i];
  return selectGoodFrameFromMatchCounts(m, required, first_fit != 0, skip_head);
}

// findGoodFrameFromBatchDescriptors: returns the good index, -2 on an exception; n_out[p] = matches of element p
int hostshim_find_good_frame(const void* q, int nq, const void* const* t, const int* nt, int P, int type,
                             int required, int first_fit, int skip_head, int* n_out) {
  try {
    const bool orb = type == ORB_BF;
    const int w = orb ? 32 : 128, ty = orb ? CV_8U : CV_32F;
    const size_t step = orb ? 32 : 512;
    cv::Mat prev(nq, w, ty, const_cast<void*>(q), step);
    std::vector<cv::Mat> batch;
    for (int p = 0; p < P; p++) batch.emplace_back(nt[p], w, ty, const_cast<void*>(t[p]), step);
    std::vector<std::vector<cv::DMatch>> all;
    const int good = findGoodFrameFromBatchDescriptors(prev, batch, type, required, first_fit != 0, skip_head, all);
    for (int p = 0; p < P; p++) n_out[p] = (int)all[(size_t)p].size();
    return good;
  } catch (...) {
    return -2;
  }
}
}
