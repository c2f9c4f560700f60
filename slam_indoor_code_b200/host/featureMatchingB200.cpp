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
#include "../../misc/IOmisc.h"
#include "../../misc/ChronoTimer.h"
#include "featureMatching.h"
#include "featureMatchingCommon.h"
// knnMatcherDistance exactly as getGoodMatches reads it (featureMatchingCommon.cpp:42)
static double knnMatcherDistance() {
  return configService.getValue<double>(ConfigFieldEnum::FM_KNN_DISTANCE);
}
#endif

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
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

// Every GPU of the box behind the batch entry point (one process, as the reference
// is): with more than one sm_100 device visible -- and unless SLAMB200_MULTI_GPU=0 -- the batch
// search spreads its window over a device set (include/slamb200.h).  nullptr = single device.
slamb200_set* deviceSet() {
  static slamb200_set* set = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* e = getenv("SLAMB200_MULTI_GPU");
    if (e && atoi(e) == 0) return;
    slamb200_set* s = nullptr;
    if (slamb200_set_init(e ? atoi(e) > 1 ? atoi(e) : 0 : 0, &s) != SLAMB200_OK) return;
    if (slamb200_set_devices(s) > 1) set = s;
    else slamb200_set_shutdown(s);
  });
  return set;
}

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

// ---- resident sets of caller-owned descriptor Mats ---------------------------------------------
// The reference keeps no descriptor handle anywhere (BatchElement has no such field,
// mainCycleStructures.h:59-64) and its CUDA build re-uploads both Mats on every call
// (featureMatchingCUDA.cpp:98-99): one search uploads the same previous-frame descriptor up to
// framesBatchSize times.  Here a Mat that owns its buffer is uploaded ONCE: the cache keeps a
// reference to the Mat (so its buffer cannot be freed and handed out again while the entry lives:
// the data pointer identifies the allocation) next to the resident set, and a signature of sampled
// rows guards against a caller that rewrites the buffer in place.  The matcher threads of
// batch.cpp:181-201 share the previous frame's descriptor: the first one uploads, the others wait
// for that upload instead of repeating it.  SLAMB200_DESC_CACHE=<entries> (default 64, 0 = off).
inline bool ownsBuffer(const Mat& m) {
#ifdef SLAMB200_CV_SHIM
  return m.owner != nullptr;
#else
  return m.u != nullptr;
#endif
}

uint64_t sampleSignature(const Mat& m) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const unsigned char* p, size_t n) {
    for (size_t i = 0; i + 8 <= n; i += 8) {
      uint64_t v;
      memcpy(&v, p + i, 8);
      h = (h ^ v) * 1099511628211ull;
    }
  };
  const size_t row_bytes = (size_t)m.cols * m.elemSize();
  const int n = m.rows, take = n < 16 ? n : 16;
  for (int k = 0; k < take; k++) {
    const int r = take > 1 ? (int)((long long)k * (n - 1) / (take - 1)) : 0;
    mix(m.data + (size_t)r * m.step, row_bytes);
  }
  return h ^ ((uint64_t)n << 32) ^ (uint64_t)m.step;
}

class DescCache {
  struct Slot {
    std::mutex m;
    std::condition_variable cv;
    bool done = false;
    DescHandle set;
    std::exception_ptr err;
  };
  struct Entry {
    Mat keep;
    int rows, cols, type, kind;
    size_t step;
    uint64_t sig, tick;
    std::shared_ptr<Slot> slot;
  };
  std::mutex mu;
  std::unordered_map<const void*, Entry> map;
  uint64_t clock = 0;
  size_t capacity;

 public:
  size_t hits = 0, misses = 0;
  DescCache() {
    const char* e = getenv("SLAMB200_DESC_CACHE");
    capacity = e ? (size_t)atol(e) : 64;
  }
  // The resident set of `desc` (uploaded now if it is not cached).  The returned pointer keeps the
  // set alive for the caller even if the entry is evicted meanwhile.
  std::shared_ptr<void> get(const Mat& desc, int kind, const slamb200_desc** out) {
    if (capacity == 0 || desc.empty() || !ownsBuffer(desc)) {
      auto own = std::make_shared<DescHandle>();
      upload(desc, kind, *own);
      *out = own->d;
      return own;
    }
    std::shared_ptr<Slot> slot;
    bool uploader = false;
    {
      std::lock_guard<std::mutex> lk(mu);
      const uint64_t sig = sampleSignature(desc);
      auto it = map.find(desc.data);
      if (it != map.end()) {
        Entry& e = it->second;
        if (e.rows == desc.rows && e.cols == desc.cols && e.type == desc.type() && e.step == (size_t)desc.step &&
            e.kind == kind && e.sig == sig) {
          e.tick = ++clock;
          slot = e.slot;
          hits++;
        } else {
          map.erase(it);   // same address, other content: the old set is stale
        }
      }
      if (!slot) {
        misses++;
        if (map.size() >= capacity) {   // evict the least recently used entry
          auto lru = map.begin();
          for (auto k = map.begin(); k != map.end(); ++k)
            if (k->second.tick < lru->second.tick) lru = k;
          map.erase(lru);
        }
        slot = std::make_shared<Slot>();
        map[desc.data] = Entry{desc, desc.rows, desc.cols, desc.type(), kind, (size_t)desc.step, sig, ++clock, slot};
        uploader = true;
      }
    }
    if (uploader) {
      std::exception_ptr err;
      try {
        upload(desc, kind, slot->set);
      } catch (...) {
        err = std::current_exception();
      }
      {
        std::lock_guard<std::mutex> lk(slot->m);
        slot->err = err;
        slot->done = true;
      }
      slot->cv.notify_all();
      if (err) {
        std::lock_guard<std::mutex> lk(mu);
        auto it = map.find(desc.data);
        if (it != map.end() && it->second.slot == slot) map.erase(it);
        std::rethrow_exception(err);
      }
    } else {
      std::unique_lock<std::mutex> lk(slot->m);
      slot->cv.wait(lk, [&] { return slot->done; });
      if (slot->err) std::rethrow_exception(slot->err);
    }
    *out = slot->set.d;
    return slot;
  }
  void clear() {
    std::lock_guard<std::mutex> lk(mu);
    map.clear();
  }
};

DescCache& descCache() {
  static DescCache c;
  return c;
}

}  // namespace

// the other B200 units of the drop-in (cameraTranslationB200.cpp, poseEstimationB200.cpp,
// triangulateB200.cpp) share the process-wide context
slamb200_ctx* slamb200HostContext() { return context(); }
// cache statistics for the tests (uploads avoided / performed) and a way to drop every resident set
void slamb200HostCacheStats(size_t* hits, size_t* misses) { *hits = descCache().hits; *misses = descCache().misses; }
void slamb200HostCacheClear() { descCache().clear(); }

/*
 * @param prevDesc [in]  query descriptors (previous frame)
 * @param curDesc [in]   train descriptors (candidate frame)
 * @param matches [out]  good matches, ascending queryIdx (cleared first, like getGoodMatches)
 * @param matcherType [in] 0 - sift_bf, 1 - sift_flann, 2 - orb_bf
 * Replaces featureMatchingCPU.cpp:17-43.
 */
static void matchFeatures(Mat& prevDesc, Mat& curDesc, std::vector<DMatch>& matches, int extractorType) {
  const int kind = descKindOf(extractorType);
  const slamb200_desc *q = nullptr, *t = nullptr;
  // the previous frame's descriptor is the same Mat for every pair of a search: uploaded once
  const std::shared_ptr<void> qh = descCache().get(prevDesc, kind, &q);
  const std::shared_ptr<void> th = descCache().get(curDesc, kind, &t);
  matches.clear();
  const int cap = prevDesc.empty() ? 0 : prevDesc.rows;
  if (cap == 0) return;
  static_assert(sizeof(DMatch) == sizeof(slamb200_dmatch), "cv::DMatch layout");
  matches.resize((size_t)cap);
  int n = 0;
  const int rc = slamb200_match_pair(context(), abiMatcher(extractorType), q, t, knnMatcherDistance(),
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

// cv::SIFT::create()->compute(frame, features, desc) on the B200 for octave-0 / layer-0 keypoints
// (what fastExtractor produces): no keypoint is dropped, desc is n x 128 CV_32F with
// integer values.  Held to a tolerance against OpenCV (every element within 1, >= 99.9 % equal:
// include/slamb200.h), not to bit-exactness -- compile with -DSLAMB200_SIFT_ON_CPU to keep cv::SIFT.
// Returns false when a keypoint lives on another pyramid level (the caller then runs cv::SIFT).
static bool computeSiftB200(Mat& frame, std::vector<KeyPoint>& features, Mat& desc) {
  if (frame.empty() || features.empty()) {  // Feature2D::compute: nothing to describe
    desc.release();
    if (frame.empty()) features.clear();
    return true;
  }
  if (frame.depth() != CV_8U || (frame.channels() != 1 && frame.channels() != 3)) return false;
  const int n = (int)features.size();
  std::vector<float> k((size_t)4 * n);
  for (int i = 0; i < n; i++) {
    if ((features[i].octave & 0xFFFF) != 0) return false;   // octave and layer of unpackOctave
    k[4 * i] = features[i].pt.x;
    k[4 * i + 1] = features[i].pt.y;
    k[4 * i + 2] = features[i].size;
    k[4 * i + 3] = features[i].angle;
  }
  desc.create(n, 128, CV_32F);
  const int rc = slamb200_sift_compute(context(), frame.data, frame.rows, frame.cols, frame.channels(),
                                       (size_t)frame.step, k.data(), n, (float*)desc.data, nullptr);
  if (rc != SLAMB200_OK) throw std::runtime_error(std::string("slamb200_sift_compute: ") + slamb200_last_error());
  return true;
}

// featureMatchingCPU.cpp:45-66.  ORB descriptors are computed on the device (bit-identical to
// cv::ORB); SIFT descriptors of FAST keypoints too (tolerance-pinned, see computeSiftB200), unless
// the unit is compiled with -DSLAMB200_SIFT_ON_CPU.
void extractDescriptor(Mat& frame, std::vector<KeyPoint>& features, int matcherType, Mat& desc) {
  cv::Ptr<cv::DescriptorExtractor> extractor;
  switch (matcherType) {
    case SIFT_BF:
    case SIFT_FLANN:
#ifndef SLAMB200_SIFT_ON_CPU
      if (computeSiftB200(frame, features, desc)) return;
#endif
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
  // the per-call timing lines of the reference's CUDA unit (featureMatchingCUDA.cpp:94-107), same
  // text, same stream: the log parsers of docs/cuda keep working
  ChronoTimer timer;
  Mat secondDescriptor;
  extractDescriptor(secondFrame, secondFeatures, matcherType, secondDescriptor);
  timer.printLastPointDelta("Descriptors extracting: ", logStreams.timeStream);
  timer.updateLastPoint();
  matchFeatures(firstFrameDescriptor, secondDescriptor, matches, matcherType);
  timer.printLastPointDelta("Matching: ", logStreams.timeStream);
}

// Optional fast path for the batch window (batch.cpp:101-226): one query
// descriptor against every batch element's descriptor in ONE call.  The per-frame descriptor sets
// are uploaded once; `allMatches[i]` receives element i's good matches.
void matchFramesBatchFeatures(Mat& firstFrameDescriptor, std::vector<Mat>& batchDescriptors,
                              int matcherType, std::vector<std::vector<DMatch>>& allMatches) {
  const int kind = descKindOf(matcherType);
  if (slamb200_set* set = deviceSet()) {
    // several GPUs: the previous frame's descriptor replicated over NVLink, batch element i on
    // device i*G/n, one enqueue per device, results gathered once
    const int P = (int)batchDescriptors.size();
    const int cap = firstFrameDescriptor.empty() ? 0 : firstFrameDescriptor.rows;
    allMatches.assign((size_t)P, std::vector<DMatch>());
    if (P == 0 || cap == 0) return;
    auto up = [&](const Mat& m, int member, slamb200_mdesc** out) {
      const int want = kind == SLAMB200_DESC_F32X128 ? CV_32F : CV_8U;
      if (!m.empty() && (m.type() != want || m.cols != (kind == SLAMB200_DESC_F32X128 ? 128 : 32)))
        throw std::runtime_error("descriptor Mat type/width does not fit the matcher");
      if (slamb200_set_upload(set, member, kind, m.empty() ? nullptr : m.data, m.empty() ? 0 : m.rows,
                              m.empty() ? 0 : (size_t)m.step, out) != SLAMB200_OK)
        throw std::runtime_error(std::string("slamb200_set_upload: ") + slamb200_last_error());
    };
    struct Guard {
      slamb200_set* s;
      std::vector<slamb200_mdesc*> d;
      ~Guard() { for (slamb200_mdesc* m : d) slamb200_set_free_desc(s, m); }
    } g{set, {}};
    g.d.assign((size_t)P + 1, nullptr);
    up(firstFrameDescriptor, -1, &g.d[0]);
    for (int i = 0; i < P; i++) up(batchDescriptors[(size_t)i], slamb200_set_owner(set, i, P), &g.d[(size_t)i + 1]);
    std::vector<slamb200_dmatch> out((size_t)P * cap);
    std::vector<int> n((size_t)P, 0);
    if (slamb200_set_match_batch(set, abiMatcher(matcherType), g.d[0], g.d.data() + 1, P, knnMatcherDistance(),
                                 out.data(), cap, n.data()) != SLAMB200_OK)
      throw std::runtime_error(std::string("slamb200_set_match_batch: ") + slamb200_last_error());
    for (int p = 0; p < P; p++) {
      const DMatch* src = reinterpret_cast<const DMatch*>(out.data() + (size_t)p * cap);
      allMatches[(size_t)p].assign(src, src + n[(size_t)p]);
    }
    return;
  }
  // every Mat that owns its buffer is resident after its first use: a batch element that stays
  // in the window over several searches (batch.cpp:112-114 re-describes it each time in the
  // reference) is uploaded once in its lifetime
  const slamb200_desc* qd = nullptr;
  const std::shared_ptr<void> qh = descCache().get(firstFrameDescriptor, kind, &qd);
  std::vector<std::shared_ptr<void>> th(batchDescriptors.size());
  std::vector<const slamb200_desc*> tp(batchDescriptors.size());
  for (size_t i = 0; i < batchDescriptors.size(); i++) th[i] = descCache().get(batchDescriptors[i], kind, &tp[i]);
  const int P = (int)batchDescriptors.size();
  const int cap = firstFrameDescriptor.empty() ? 0 : firstFrameDescriptor.rows;
  allMatches.assign((size_t)P, std::vector<DMatch>());
  if (P == 0 || cap == 0) return;
  std::vector<slamb200_dmatch> out((size_t)P * cap);
  std::vector<int> n((size_t)P, 0);
  const int rc = slamb200_match_batch(context(), abiMatcher(matcherType), qd, tp.data(), P, knnMatcherDistance(),
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
