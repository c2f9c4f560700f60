/*
 * slamb200.h -- C ABI of libslamb200.so, the B200-native (sm_100a) correspondence hot path of
 * FIT-2023-SLAM-indoor/slam-indoor-code.
 *
 * This is the drop-in boundary.  Everything the reference computes between
 *     matchFramesPairFeatures(...)   src/mainModule/featureMatching/featureMatching.h:29-53
 *     findEssentialMat(..., RANSAC)  src/mainModule/translation/cameraTranslation.cpp:41-46
 * and their outputs (std::vector<cv::DMatch>, the uchar inlier mask) is reachable through the
 * plain-C entry points below: plain pointers and sizes, no C++ / OpenCV / torch types, integer
 * status codes, no exceptions.  The reference-side binding (featureMatchingB200.cpp, the third
 * translation unit behind featureMatching.h next to featureMatchingCPU.cpp /
 * featureMatchingCUDA.cpp) is shown in INTEGRATION.md and shipped in
 * slam_indoor_code_b200/host/.
 *
 * There is no CPU fallback: every compute entry point runs hand-written CUDA kernels on the
 * context's device and fails with SLAMB200_ERR_CUDA when no sm_100 device is present.
 *
 * Thread safety: a context may be used from several host threads at once (the reference calls
 * matchFramesPairFeatures from `threadsCount` std::threads sharing one read-only query
 * descriptor, batch.cpp:181-201).  Descriptor handles are immutable after upload and may be
 * shared between threads; each call borrows one of the context's internal lanes (stream +
 * scratch).
 */
#ifndef SLAMB200_H
#define SLAMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLAMB200_VERSION 100 /* 0.1.0 */

/* ---- status codes ---------------------------------------------------------------------- */
#define SLAMB200_OK 0
#define SLAMB200_ERR_INVALID (-1)  /* null pointer, negative size, capacity too small ...     */
#define SLAMB200_ERR_CUDA (-2)     /* CUDA runtime error or no sm_100 device                  */
#define SLAMB200_ERR_NOMEM (-3)    /* device or host allocation failed                        */
#define SLAMB200_ERR_MATCHER (-4)  /* matcher type not 0/1/2: the reference throws
                                      std::exception() there (featureMatchingCPU.cpp:37)      */
#define SLAMB200_ERR_KIND (-5)     /* descriptor kind does not fit the matcher (OpenCV asserts
                                      on a type mismatch in batchDistance)                    */
#define SLAMB200_ERR_INTERNAL (-6) /* a device-side self check failed (never expected)        */

/* ---- matcher types: enum MatcherType, featureMatchingCommon.h:8-12 --------------------- */
#define SLAMB200_SIFT_BF 0    /* useFM-SIFT-BF    -> BFMatcher NORM_L2 (featureMatchingCPU.cpp:27) */
#define SLAMB200_SIFT_FLANN 1 /* useFM-SIFT-FLANN -> exact L2 as well: the reference's FLANN
                                 (featureMatchingCPU.cpp:30) is approximate, its ground truth is
                                 the BF result, so recall vs BF is 1.0 by construction        */
#define SLAMB200_ORB_BF 2     /* useFM-ORB        -> BFMatcher NORM_HAMMING (:33)              */
#define SLAMB200_SIFT_BF_L1 3 /* not a reference enum value: what useFM-SIFT-BF means in the reference's
                                 OpenCV-CUDA build, cuda BFMatcher NORM_L1 (featureMatchingCUDA.cpp:28);
                                 SURVEY.md 8f-4.  Results as cv::BFMatcher(NORM_L1) on the CPU        */

/* ---- descriptor kinds: what extractDescriptor emits (featureMatchingCPU.cpp:45-66) ------ */
#define SLAMB200_DESC_F32X128 0 /* cv::SIFT: CV_32F, 128 columns */
#define SLAMB200_DESC_U8X32 1   /* cv::ORB:  CV_8U,  32 columns  */

typedef struct slamb200_ctx slamb200_ctx;
typedef struct slamb200_desc slamb200_desc; /* one frame's descriptor set, resident in HBM */
typedef struct slamb200_pts slamb200_pts;   /* one frame's keypoint coordinates (x,y) in HBM */

/* Bit-compatible with cv::DMatch {int queryIdx; int trainIdx; int imgIdx; float distance;}. */
typedef struct slamb200_dmatch {
  int32_t queryIdx;
  int32_t trainIdx;
  int32_t imgIdx;
  float distance;
} slamb200_dmatch;

/* ---- context ---------------------------------------------------------------------------- */
/* Binds a context to CUDA device `device` (main.cpp:30-38 only checks that a device exists). */
int slamb200_init(int device, slamb200_ctx** out);
int slamb200_shutdown(slamb200_ctx* ctx);
int slamb200_version(void);
/* Last error text of the calling thread (never NULL). */
const char* slamb200_last_error(void);
/* Kernels this library has launched on the context since init (the bench's gpu_launches). */
int64_t slamb200_launch_count(const slamb200_ctx* ctx);
/* Blocks until all work queued by the context has finished. */
int slamb200_synchronize(slamb200_ctx* ctx);

/* ---- descriptor sets: replaces the per-call cuda::GpuMat uploads of
 *      featureMatchingCUDA.cpp:98-99 with "upload once per frame" ----------------------- */
/* rows: host pointer to n rows of the kind's type, row_stride bytes apart (cv::Mat::step; any
 * multiple of the element size, no alignment requirement on `rows` for any upload entry point).
 * The rows are consumed when the call returns.  (Implemented by slamb200_upload_desc_packed.) */
int slamb200_upload_desc(slamb200_ctx* ctx, int kind, const void* rows, int n, size_t row_stride,
                         slamb200_desc** out);
/* Same, but `rows` is a device pointer on the context's device (bytes already in HBM);
 * `stream` (a cudaStream_t, may be NULL) is the stream the bytes were produced on. */
int slamb200_upload_desc_device(slamb200_ctx* ctx, int kind, const void* rows, int n,
                                size_t row_stride, void* stream, slamb200_desc** out);
/* Pipelined upload for callers that keep `rows` in page-locked host memory: returns as soon as
 * the copy is queued; `rows` must stay valid and unmodified until slamb200_synchronize(ctx) or
 * until a matching call that uses the set has returned its results. */
int slamb200_upload_desc_pinned(slamb200_ctx* ctx, int kind, const void* rows, int n,
                                size_t row_stride, slamb200_desc** out);
/* Upload that narrows on the host first: cv::SIFT's CV_32F rows are integers in [0,255], so the
 * calling thread packs them to bytes into page-locked staging -- checking every element and every
 * row norm -- and the GPU reads a quarter of the bytes over PCIe.  `rows` (pageable or pinned) is
 * fully consumed when the call returns; the device work is only queued.  A Mat that is not
 * integer valued, and every ORB Mat, takes slamb200_upload_desc's path.  Results are identical
 * to slamb200_upload_desc in every case. */
int slamb200_upload_desc_packed(slamb200_ctx* ctx, int kind, const void* rows, int n,
                                size_t row_stride, slamb200_desc** out);
/* ---- descriptor sets across processes (one process per GPU, keyframe window) ----- */
/* A set uploaded with slamb200_upload_desc_shared lives in an allocation that other processes on
 * the node can map (CUDA IPC).  slamb200_desc_export fills a plain 128-byte record that may be
 * sent to the other ranks by any means (it waits for the set's prep kernel first);
 * slamb200_desc_import maps the exporter's prepared slab -- bf16 operands, norms, byte copy, flags
 * -- into this process and returns an ordinary handle: every entry point accepts it, the tcgen05
 * kernel then TMA-loads that frame's tiles over NVLink while it computes (no all-gather of rows,
 * no second prep pass).  A peer-mapped QUERY set costs one pass over its operand per frame pair;
 * a peer-mapped TRAIN set would be re-read once per query block, so copy it first with
 * slamb200_desc_localize (one device-to-device transfer of the prepared slab).  The exporter must
 * keep its set alive until every importer has freed its mapping. */
typedef struct slamb200_desc_ipc {
  unsigned char handle[64]; /* cudaIpcMemHandle_t */
  int kind, n, n_pad, device, exact, reserved;
  unsigned long long slab_bytes;
  unsigned char pad[32];
} slamb200_desc_ipc;
int slamb200_upload_desc_shared(slamb200_ctx* ctx, int kind, const void* rows, int n,
                                size_t row_stride, slamb200_desc** out);
int slamb200_desc_export(slamb200_ctx* ctx, const slamb200_desc* d, slamb200_desc_ipc* out);
int slamb200_desc_import(slamb200_ctx* ctx, const slamb200_desc_ipc* in, slamb200_desc** out);
int slamb200_desc_localize(slamb200_ctx* ctx, const slamb200_desc* src, slamb200_desc** out);

/* Host threads that share the narrowing of one Mat in slamb200_upload_desc_packed (row slices; the
 * caller's thread takes one slice too).  Default min(hardware threads, 16); 0 = callers only.  Must
 * be called before the first packed upload. */
int slamb200_set_pack_threads(slamb200_ctx* ctx, int n);
int slamb200_free_desc(slamb200_ctx* ctx, slamb200_desc* d);
int slamb200_desc_rows(const slamb200_desc* d);
int slamb200_desc_kind(const slamb200_desc* d);
/* 1 when the set is integer-valued in [0,255] with small norms (what cv::SIFT emits) and so
 * takes the tcgen05 candidate path; 0 for general floats (exact fp32 path); -1 for ORB. */
int slamb200_desc_exact_mode(const slamb200_desc* d);

/* ---- hot path A: matchFeatures (featureMatchingCPU.cpp:17-43) --------------------------- */
/* knnMatch(query, train, k=2) raw result, the shape cv::BFMatcher returns before the ratio test:
 * idx[2*q+k] = train index of the k-th neighbour of query row q or -1 when the train set has
 * fewer than k+1 rows; dist likewise (sqrt-L2 as float, or the Hamming count as float). */
int slamb200_knn2(slamb200_ctx* ctx, int matcher, const slamb200_desc* query,
                  const slamb200_desc* train, int32_t* idx, float* dist);

/* knnMatch + getGoodMatches (featureMatchingCommon.cpp:37-50): out receives the accepted
 * matches in ascending queryIdx, *n_out their count; cap must be >= rows(query).  ratio is
 * the knnMatcherDistance config value; the comparison is (double)d0 < ratio * (double)d1. */
int slamb200_match_pair(slamb200_ctx* ctx, int matcher, const slamb200_desc* query,
                        const slamb200_desc* train, double ratio, slamb200_dmatch* out, int cap,
                        int* n_out);

/* The batch window of batch.cpp:120-148 / :181-201 in one call: one query frame against
 * n_pairs train frames.  out is n_pairs slabs of `cap` matches, n_out[p] the count of pair p. */
int slamb200_match_batch(slamb200_ctx* ctx, int matcher, const slamb200_desc* query,
                         const slamb200_desc* const* trains, int n_pairs, double ratio,
                         slamb200_dmatch* out, int cap, int* n_out);

/* The same window from HOST descriptor Mats that are not resident yet, in one call: query rows
 * (nq x row type, q_stride bytes apart) and n_pairs train Mats (t_rows[p], t_n[p] rows, t_stride[p]
 * bytes apart; t_stride may be NULL = dense).  Uploads and matching overlap inside the library
 * (host-side narrowing on the pack pool, chunks matched as their last Mat arrives); nothing stays
 * resident afterwards.  Results as slamb200_match_batch.  This is the end-to-end call of a search
 * whose descriptors live in host memory (bench.py `e2e`). */
int slamb200_match_batch_host(slamb200_ctx* ctx, int matcher, const void* q_rows, int nq,
                              size_t q_stride, const void* const* t_rows, const int* t_n,
                              const size_t* t_stride, int n_pairs, double ratio, slamb200_dmatch* out,
                              int cap, int* n_out);

/* All i<j pairs of a frame window (BAMaxFramesCnt window, config 4): pair order is
 * (0,1),(0,2)...(0,n-1),(1,2)...; frame i is the query, frame j the train. */
int slamb200_match_window(slamb200_ctx* ctx, int matcher, const slamb200_desc* const* frames,
                          int n_frames, double ratio, slamb200_dmatch* out, int cap, int* n_out);

/* Device-resident variants: enqueue the batch on `stream` (cudaStream_t; NULL = the lane's own
 * stream), keep the results in HBM inside the context, return without synchronising.
 * slamb200_batch_fetch copies the results of the last enqueue to the host.
 * The context keeps ONE result set for these four entry points (enqueue / fetch / score enqueue /
 * scores fetch): each call is ordered on the device behind the previous one whatever stream it
 * was given, and a new enqueue overwrites the results of the last.  n_pairs <= 65534. */
int slamb200_match_batch_enqueue(slamb200_ctx* ctx, int matcher, const slamb200_desc* query,
                                 const slamb200_desc* const* trains, int n_pairs, double ratio,
                                 void* stream);
int slamb200_batch_fetch(slamb200_ctx* ctx, slamb200_dmatch* out, int cap, int* n_out,
                         void* stream);

/* ---- one process, every GPU of the box ------------------------------------
 * The reference is ONE process that walks the framesBatchSize window from `threadsCount` host
 * threads (batch.cpp:162-226).  A device set gives that process all its GPUs behind the same call
 * shape: the query frame's prepared descriptor set is replicated over NVLink (one upload, peer
 * copies of the prepared slab, no second prep pass), every train frame lives on ONE device (the
 * one that will match it: frame i of an n-frame window on device i*G/n, slamb200_set_owner), a
 * batch call enqueues each device's share on that device without waiting and gathers the match
 * lists once.  No data-path collective: pairs are independent.  Results are those of
 * slamb200_match_batch on a single device, pair for pair. */
typedef struct slamb200_set slamb200_set;      /* contexts on n devices of this process        */
typedef struct slamb200_mdesc slamb200_mdesc;  /* a descriptor set resident on one or all of them */
/* n_devices <= 0: every visible sm_100 device; otherwise devices 0 .. n_devices-1. */
int slamb200_set_init(int n_devices, slamb200_set** out);
int slamb200_set_shutdown(slamb200_set* set);
int slamb200_set_devices(const slamb200_set* set);
/* The single-device context of member i (every slamb200_* entry point may be used on it). */
slamb200_ctx* slamb200_set_ctx(slamb200_set* set, int i);
/* Device that owns item i of n under the contiguous split (i*G/n). */
int slamb200_set_owner(const slamb200_set* set, int i, int n);
/* Host rows -> a set resident on member `member`, or (member < 0) replicated on every member. */
int slamb200_set_upload(slamb200_set* set, int member, int kind, const void* rows, int n,
                        size_t row_stride, slamb200_mdesc** out);
int slamb200_set_free_desc(slamb200_set* set, slamb200_mdesc* d);
int slamb200_mdesc_rows(const slamb200_mdesc* d);
/* One query frame against n_pairs train frames, each matched on the device it lives on (a
 * replicated train set on the member with the least work).  out / cap / n_out as
 * slamb200_match_batch.  The query must be replicated. */
int slamb200_set_match_batch(slamb200_set* set, int matcher, const slamb200_mdesc* query,
                             const slamb200_mdesc* const* trains, int n_pairs, double ratio,
                             slamb200_dmatch* out, int cap, int* n_out);
/* Device-resident form: enqueue on every member, then fetch.  device_ms (may be NULL, n members)
 * receives each member's device time of the last enqueue (CUDA events around its share). */
int slamb200_set_match_batch_enqueue(slamb200_set* set, int matcher, const slamb200_mdesc* query,
                                     const slamb200_mdesc* const* trains, int n_pairs, double ratio);
int slamb200_set_batch_fetch(slamb200_set* set, slamb200_dmatch* out, int cap, int* n_out,
                             float* device_ms);

/* ---- hot path B: RANSAC essential-matrix inlier scoring (cameraTranslation.cpp:41-46) --- */
/* Scores H candidate essential matrices (row-major 3x3 doubles, in normalised coordinates, as
 * the 5-point solver returns them) against M matches exactly as cv::findEssentialMat's RANSAC
 * loop does: points normalised with K = {fx, fy, cx, cy}, threshold_px (RPRANSACThreshold)
 * divided by (fx+fy)/2, Sampson error in fp64 rounded to float, inlier iff err <= (float)thr^2,
 * a model replaces the best iff count > max(best_count, 4).  counts[H]; *best = winning index or
 * -1; best_mask[M] = 0/1 (the N x 1 uchar mask the reference logs, cameraTranslation.cpp:53);
 * all_masks (H*M bytes) may be NULL. */
int slamb200_score_essential(slamb200_ctx* ctx, const float* pts1, const float* pts2, int M,
                             const double K[4], const double* E, int H, double threshold_px,
                             int32_t* counts, int32_t* best, uint8_t* best_mask,
                             uint8_t* all_masks);

/* P independent pairs in one launch sequence.  Matches are ragged: pair p owns
 * pts[m_off[p] .. m_off[p+1]) (m_off has P+1 entries); E is P*H*9 doubles; counts P*H;
 * best P; best_mask m_off[P] bytes. */
int slamb200_score_essential_batch(slamb200_ctx* ctx, int P, const float* pts1,
                                   const float* pts2, const int32_t* m_off, const double K[4],
                                   const double* E, int H, double threshold_px, int32_t* counts,
                                   int32_t* best, uint8_t* best_mask);

/* ---- next row: solvePnPRansac inlier scoring (cycleProcessing/mainCycle.cpp:155-159) --- */
/* Scores H candidate poses against M 3D-2D correspondences exactly as cv::solvePnPRansac's loop
 * does (PnPRansacCallback::computeError + RANSACPointSetRegistrator::findInliers): obj M x 3
 * floats (Point3f), img M x 2 floats (Point2f), K = {fx, fy, cx, cy}, dist = n_dist (0..14)
 * coefficients in OpenCV order k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 tauX tauY (the reference passes
 * five; non-zero tauX/tauY -> SLAMB200_ERR_INVALID), poses = H x 12 doubles, the rotation MATRIX
 * row-major followed by t (Rodrigues of rvec stays with the caller).  Every point is projected in
 * fp64 in cvProjectPoints2's operation order, rounded to float, err = dx*dx + dy*dy in float,
 * inlier iff err <= (float)(reproj_err^2); a model replaces the best iff
 * count > max(best_count, model_points - 1) (model_points: 5, or 4 when M == 4).  Outputs as
 * slamb200_score_essential. */
int slamb200_score_pnp(slamb200_ctx* ctx, const float* obj, const float* img, int M,
                       const double K[4], const double* dist, int n_dist, const double* poses,
                       int H, double reproj_err, int model_points, int32_t* counts,
                       int32_t* best, uint8_t* best_mask, uint8_t* all_masks);

/* P independent frames in one launch sequence, ragged like slamb200_score_essential_batch:
 * poses P*H*12 doubles, counts P*H, best P, best_mask m_off[P] bytes. */
int slamb200_score_pnp_batch(slamb200_ctx* ctx, int P, const float* obj, const float* img,
                             const int32_t* m_off, const double K[4], const double* dist,
                             int n_dist, const double* poses, int H, double reproj_err,
                             int model_points, int32_t* counts, int32_t* best,
                             uint8_t* best_mask);

/* ---- next row: ORB descriptors of given keypoints (featureMatchingCPU.cpp:45-66) --- */
/* cv::ORB::create()->compute(frame, features, desc) for the ORB_BF matcher: image = rows x cols
 * CV_8UC3 (BGR) or CV_8UC1, `step` bytes per row; kps = n x {x, y, angle in degrees} floats
 * (cv::KeyPoint::pt and ::angle; FAST keypoints carry angle -1, octave 0).  Keypoints within 31 px
 * of the border are dropped exactly as compute() drops them: keep[n] (may be NULL) flags the
 * survivors in order, *n_kept counts them -- the caller erases the others from its vector, which
 * is what the reference observes of `features` after the call.  desc (may be NULL) receives
 * n_kept x 32 bytes on the host; *resident (may be NULL) receives the same rows as a descriptor
 * set already in HBM, ready for slamb200_match_*: no descriptor upload.  Bit-identical to OpenCV's
 * descriptors (integer gray conversion, its float 7x7 Gaussian in FMA order, float rotation of the
 * 256-pair pattern, cvRound). */
int slamb200_orb_compute(slamb200_ctx* ctx, const uint8_t* image, int rows, int cols, int channels,
                         size_t step, const float* kps, int n, uint8_t* keep, uint8_t* desc,
                         int* n_kept, slamb200_desc** resident);

/* ---- next row: SIFT descriptors of given keypoints (featureMatchingCPU.cpp:45-66) --- */
/* cv::SIFT::create()->compute(frame, features, desc) for the SIFT_BF / SIFT_FLANN matchers on
 * keypoints of octave 0 / layer 0 -- the only kind the reference produces (fastExtractor.cpp:7-13:
 * FAST keypoints, size 7, angle -1).  image as in slamb200_orb_compute; kps = n x {x, y, size,
 * angle in degrees} floats.  No keypoint is dropped (cv::SIFT::compute keeps them all).  desc (may
 * be NULL) receives n x 128 floats on the host, integer valued like cv::SIFT's; *resident (may be
 * NULL) the same rows as a descriptor set already in HBM, ready for slamb200_match_*.
 * TOLERANCE, not bit-exactness: the working image (gray -> float -> 13-tap Gaussian) equals
 * OpenCV's bit for bit; the descriptors follow calcSIFTDescriptor operation by operation but
 * OpenCV's exp / magnitude are IPP routines and its histogram sum runs in sample order, so a
 * scaled value that lands within ~1e-4 of a rounding boundary may round the other way: every
 * element is within 1 of OpenCV's and >= 99.9 % are equal (measured 99.998 %). */
int slamb200_sift_compute(slamb200_ctx* ctx, const uint8_t* image, int rows, int cols, int channels,
                          size_t step, const float* kps, int n, float* desc, slamb200_desc** resident);

/* ---- next row: FAST keypoints (featureExtraction/fastExtractor.cpp:7-13) --- */
/* FastFeatureDetector::create(threshold, nonmax, TYPE_9_16)->detect(image, points), the reference's
 * fastExtractor (callers cycleProcessing/batch.cpp:245, mainCycleInternals.cpp:144 with
 * featureExtractingThreshold): image = rows x cols CV_8UC3 (BGR, converted like cvtColor) or
 * CV_8UC1, `step` bytes per row.  kps receives up to cap rows of {x, y, response} floats in
 * OpenCV's order (row by row, left to right) -- KeyPoint(x, y, size 7, angle -1, response);
 * response is the corner score with suppression and 0 without, as in OpenCV.  *n_found is the
 * number of keypoints in the frame; when it exceeds cap only the first cap are written.  Same
 * keypoints and responses as OpenCV. */
int slamb200_fast_detect(slamb200_ctx* ctx, const uint8_t* image, int rows, int cols, int channels,
                         size_t step, int threshold, int nonmax, float* kps, int cap, int* n_found);

/* slamb200_fast_detect followed by slamb200_orb_compute on the same frame in one call (the
 * reference's front end for useFM-ORB: fastExtractor at batch.cpp:245, then extractDescriptor,
 * featureMatchingCPU.cpp:45-66): the frame crosses PCIe and is converted to gray once.  kps / cap /
 * n_found as in slamb200_fast_detect, except that cap < *n_found is an error (SLAMB200_ERR_INVALID,
 * *n_found set: retry with a larger buffer); keep (may be NULL, n_found bytes), desc (may be NULL),
 * n_kept and resident as in slamb200_orb_compute.  Same results as the two calls. */
int slamb200_fast_orb_compute(slamb200_ctx* ctx, const uint8_t* image, int rows, int cols, int channels,
                              size_t step, int threshold, int nonmax, float* kps, int cap,
                              int* n_found, uint8_t* keep, uint8_t* desc, int* n_kept,
                              slamb200_desc** resident);

/* ---- next row: linear triangulation (triangulation/triangulate.cpp:17-55, :91-108) --- */
/* reconstructPointsFor3D for M matches: P1, P2 = 3x4 row-major projection matrices (K*[R|t], as
 * reconstruct() forms them at triangulate.cpp:74-80), pts = M x 2 floats (vector<Point2f>).  Per
 * match: the 4x4 DLT system and row 3 of Vt of its SVD (OpenCV's one-sided Jacobi sequence).
 * points4d (4 x M row-major, the reference's homogeneous Mat; may be NULL) and points3d (M x 3 =
 * (X,Y,Z)*(1/W), what convertHomogeneousPointsMatrixToSpatialPointsVector returns; may be NULL).
 * Floating point: equal to OpenCV within 1e-12 relative (libm vs CUDA hypot/sqrt), not bit-exact. */
int slamb200_triangulate(slamb200_ctx* ctx, const double P1[12], const double P2[12],
                         const float* pts1, const float* pts2, int M, double* points4d,
                         double* points3d);

/* Keypoint coordinates of a frame (cv::KeyPoint::pt as x,y float pairs, `stride` bytes apart)
 * kept in HBM so that getKeyPointCoordsFromFramePair (featureMatchingCommon.cpp:23-33) can run
 * on the device between matching and scoring. */
int slamb200_upload_pts(slamb200_ctx* ctx, const float* xy, int n, size_t stride,
                        slamb200_pts** out);
int slamb200_free_pts(slamb200_ctx* ctx, slamb200_pts* p);

/* Chained batch: score, for every pair of the last slamb200_match_batch_enqueue, H hypotheses
 * (E: n_pairs*H*9 doubles, host memory or already resident on the context's device) against that pair's accepted matches, gathering the coordinates
 * on the device (query_pts / train_pts[p]).  train_pts holds one set per pair of that batch, each
 * with at least as many points as the pair's train descriptor set has rows (checked:
 * SLAMB200_ERR_INVALID).  Results stay in HBM until slamb200_batch_scores_fetch.  Enqueue-only. */
int slamb200_score_batch_enqueue(slamb200_ctx* ctx, const slamb200_pts* query_pts,
                                 const slamb200_pts* const* train_pts, const double K[4],
                                 const double* E, int H, double threshold_px, void* stream);
int slamb200_batch_scores_fetch(slamb200_ctx* ctx, int32_t* counts, int32_t* best,
                                uint8_t* best_mask, int mask_cap, void* stream);

/* ---- kernel timing (CUDA events recorded around the hot kernels inside the enqueue calls) -- */
#define SLAMB200_K_SIFT_TC 0    /* tcgen05 candidate kernel          */
#define SLAMB200_K_SIFT_EXACT 1 /* exact fp32 kernel (general floats) */
#define SLAMB200_K_ORB 2        /* Hamming kernel                    */
#define SLAMB200_K_RANSAC 3     /* Sampson counting kernel           */
#define SLAMB200_K_SIFT_RERANK 4 /* exact dp4a rerank of the candidates */
#define SLAMB200_K_FINALIZE 5   /* ratio test + ordered compaction    */
#define SLAMB200_K_SIFT_TC_GEN 6 /* tcgen05 kernel, general floats     */
#define SLAMB200_K_SIFT_GEN_RERANK 7 /* certified fp32 rerank + fallback */
#define SLAMB200_K_PNP 8        /* reprojection-error counting kernel */
#define SLAMB200_K_SIFT_L1 9    /* NORM_L1 on integer-valued rows (byte-wise SAD) */
#define SLAMB200_K_ORB_DESC 10  /* ORB gray + blur + descriptor kernels */
#define SLAMB200_K_FAST 11      /* gray + FAST score, suppression, ordered output */
#define SLAMB200_K_SIFT_DESC 12 /* SIFT base image blur + descriptor kernel */
#define SLAMB200_K_COUNT 13
int slamb200_profile_enable(slamb200_ctx* ctx, int on);
/* The same for a subset of the kernel classes (bit k = SLAMB200_K_*; 0 = off): a timed region that
 * wants the dominant kernel's duration pays for two events per launch of THAT kernel only -- an
 * event between two kernels is a boundary the device must drain to. */
int slamb200_profile_enable_kinds(slamb200_ctx* ctx, unsigned kinds);
/* Synchronises, then returns the summed device time (ms) and the launch count of each kernel
 * class since the last read; ms and launches have SLAMB200_K_COUNT entries. */
int slamb200_profile_read(slamb200_ctx* ctx, double* ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* SLAMB200_H */
