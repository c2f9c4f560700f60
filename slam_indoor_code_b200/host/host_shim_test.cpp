// host_shim_test.cpp -- C entry points over the shim build of featureMatchingB200.cpp so that the
// Python tests can drive the C++ host layer (the static matchFeatures and the batch fast path)
// exactly as the reference's callers would (batch.cpp:127-130).
#include "featureMatchingB200.cpp"
#include "fastExtractorB200.cpp"
#include "cameraTranslationB200.cpp"
#include "poseEstimationB200.cpp"
#include "triangulateB200.cpp"
#include "ransac_control.h"

LogFilesStreams logStreams;   // misc/IOmisc.h:19 in the reference

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

// extractDescriptor(frame, features, SIFT_BF, desc): returns the number of keypoints left in `features`
// (SIFT drops none), the descriptor rows (n x 128 floats) in desc_out; -1 on an exception.
int hostshim_extract_sift(const void* frame, int rows, int cols, int channels, size_t step, const float* kps,
                          int n, float* desc_out, int cap) {
  try {
    cv::Mat f(rows, cols, channels == 3 ? CV_8UC3 : CV_8U, const_cast<void*>(frame), step);
    std::vector<cv::KeyPoint> features((size_t)n);
    for (int i = 0; i < n; i++) {
      features[i].pt.x = kps[4 * i];
      features[i].pt.y = kps[4 * i + 1];
      features[i].size = kps[4 * i + 2];
      features[i].angle = kps[4 * i + 3];
    }
    cv::Mat desc;
    extractDescriptor(f, features, SIFT_BF, desc);
    if ((int)features.size() > cap || desc.rows != (desc.empty() ? 0 : (int)features.size()) ||
        (!desc.empty() && (desc.cols != 128 || desc.type() != CV_32F)))
      return -2;
    for (int i = 0; i < desc.rows; i++) memcpy(desc_out + (size_t)128 * i, desc.data + (size_t)i * desc.step, 512);
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

// ---- the other B200 units of the drop-in, with the CPU solvers injected by the harness ----------
extern "C" {
void hostshim_set_solvers(cv::shim::FivePointFn f5, cv::shim::SolvePnPFn pnp, cv::shim::RodriguesFn rod) {
  cv::shim::five_point() = f5;
  cv::shim::solve_pnp() = pnp;
  cv::shim::rodrigues() = rod;
}

// findEssentialMatB200 (cameraTranslation.cpp:41-46): 1 = model found (E_out, mask_out filled), 0 = none, -1 = exception
int hostshim_find_essential_mat(const float* p1, const float* p2, int N, const double* K9, double prob, double thr,
                                double* E_out, unsigned char* mask_out) {
  try {
    std::vector<cv::Point2f> a((size_t)N), b((size_t)N);
    for (int i = 0; i < N; i++) { a[i] = cv::Point2f(p1[2 * i], p1[2 * i + 1]); b[i] = cv::Point2f(p2[2 * i], p2[2 * i + 1]); }
    cv::Mat K(3, 3, CV_64F, const_cast<double*>(K9)), mask;
    cv::Mat E = findEssentialMatB200(slamb200HostContext(), a, b, K, prob, thr, mask);
    if (E.empty()) return 0;
    for (int i = 0; i < 9; i++) E_out[i] = E.at<double>(i / 3, i % 3);
    for (int i = 0; i < N; i++) mask_out[i] = mask.at<unsigned char>(i, 0);
    return 1;
  } catch (...) {
    return -1;
  }
}

// solvePnPRansacB200 (mainCycle.cpp:155-159): returns 1/0 like cv::solvePnPRansac, -1 on an exception
int hostshim_solve_pnp_ransac(const float* obj, const float* img, int N, const double* K9, const double* dist,
                              int n_dist, double* rvec_out, double* tvec_out, int* inliers_out, int* n_inliers) {
  try {
    std::vector<cv::Point3f> o((size_t)N);
    std::vector<cv::Point2f> m((size_t)N);
    for (int i = 0; i < N; i++) { o[i] = cv::Point3f(obj[3 * i], obj[3 * i + 1], obj[3 * i + 2]); m[i] = cv::Point2f(img[2 * i], img[2 * i + 1]); }
    cv::Mat K(3, 3, CV_64F, const_cast<double*>(K9)), D, rvec, tvec;
    if (n_dist > 0) D = cv::Mat(1, n_dist, CV_64F, const_cast<double*>(dist));
    std::vector<int> inl;
    const bool ok = solvePnPRansacB200(slamb200HostContext(), o, m, K, D, rvec, tvec, &inl);
    if (!ok) return 0;
    for (int i = 0; i < 3; i++) { rvec_out[i] = rvec.at<double>(i); tvec_out[i] = tvec.at<double>(i); }
    *n_inliers = (int)inl.size();
    for (size_t i = 0; i < inl.size(); i++) inliers_out[i] = inl[i];
    return 1;
  } catch (...) {
    return -1;
  }
}

// triangulationWrapper (triangulate.cpp:57-72): N x 2 CV_64F point Mats, 3 x 4 projection matrices -> 4 x N
int hostshim_triangulate(const double* p1, const double* p2, int N, const double* P1, const double* P2, double* out4xN) {
  try {
    cv::Mat a(N, 2, CV_64F, const_cast<double*>(p1)), b(N, 2, CV_64F, const_cast<double*>(p2));
    cv::Mat m1(3, 4, CV_64F, const_cast<double*>(P1)), m2(3, 4, CV_64F, const_cast<double*>(P2)), X;
    triangulationWrapper(a, b, m1, m2, X);
    if (X.rows != 4 || X.cols != N) return -2;
    for (int r = 0; r < 4; r++) memcpy(out4xN + (size_t)r * N, X.ptr<double>(r), sizeof(double) * (size_t)N);
    return 0;
  } catch (...) {
    return -1;
  }
}

// ---- caller-owned Mats that persist across calls (what BatchElement / the previous frame hold) ----
void* hostshim_mat_create(const void* rows, int n, int cols, int type) {
  cv::Mat* m = new cv::Mat();
  m->create(n, cols, type);
  if (n > 0) memcpy(m->data, rows, (size_t)n * m->step);
  return m;
}
void hostshim_mat_free(void* m) { delete (cv::Mat*)m; }
void hostshim_mat_write(void* m, const void* rows) {   // in-place rewrite of the same buffer
  cv::Mat* mm = (cv::Mat*)m;
  memcpy(mm->data, rows, (size_t)mm->rows * mm->step);
}
int hostshim_match_mats(void* q, void* t, int type, cv::DMatch* out, int cap) {
  try {
    std::vector<cv::DMatch> m;
    matchFeatures(*(cv::Mat*)q, *(cv::Mat*)t, m, type);
    if ((int)m.size() > cap) return -2;
    for (size_t i = 0; i < m.size(); i++) out[i] = m[i];
    return (int)m.size();
  } catch (...) {
    return -1;
  }
}
void hostshim_cache_stats(long long* hits, long long* misses) {
  size_t h = 0, m = 0;
  slamb200HostCacheStats(&h, &m);
  *hits = (long long)h;
  *misses = (long long)m;
}
void hostshim_cache_clear() { slamb200HostCacheClear(); }

// matchFramesPairFeatures(firstFrameDescriptor, secondFrame, secondFeatures, ORB_BF, matches)
// (featureMatching.h:47-53) on a persistent query Mat and a BGR / gray frame with FAST-like keypoints
int hostshim_match_frames_pair_orb(void* q_desc, const void* frame, int rows, int cols, int channels, size_t step,
                                   const float* kps, int n, cv::DMatch* out, int cap, int* n_features_left) {
  try {
    cv::Mat f(rows, cols, channels == 3 ? CV_8UC3 : CV_8U, const_cast<void*>(frame), step);
    std::vector<cv::KeyPoint> features((size_t)n);
    for (int i = 0; i < n; i++) { features[i].pt.x = kps[3 * i]; features[i].pt.y = kps[3 * i + 1]; features[i].angle = kps[3 * i + 2]; }
    std::vector<cv::DMatch> m;
    matchFramesPairFeatures(*(cv::Mat*)q_desc, f, features, ORB_BF, m);
    *n_features_left = (int)features.size();
    if ((int)m.size() > cap) return -2;
    for (size_t i = 0; i < m.size(); i++) out[i] = m[i];
    return (int)m.size();
  } catch (...) {
    return -1;
  }
}
// the text written to logStreams.timeStream so far (and clears it)
int hostshim_time_log(char* buf, int cap) {
  const std::string s = logStreams.timeStream.str();
  logStreams.timeStream.str("");
  const int n = (int)s.size() < cap - 1 ? (int)s.size() : cap - 1;
  memcpy(buf, s.data(), (size_t)n);
  buf[n] = 0;
  return n;
}
}
