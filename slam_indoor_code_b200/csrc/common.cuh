// common.cuh -- internal declarations shared by the kernels and the C ABI of libslamb200.
//
// Data layout in HBM (DESIGN.md "Layout"):
//   SIFT frame (n rows, padded to n_pad = multiple of 256 rows):
//     f32   [n_pad][128] fp32   the caller's rows (exact path, fp32 rerank)
//     bf16  [n_pad][128] bf16   tcgen05 operand, same buffer for the query and the train role
//     augq  [n_pad][16]  bf16   K-augmentation block when the frame is the query  (A operand)
//     augt  [n_pad][16]  bf16   K-augmentation block when the frame is the train  (B operand)
//                               both stored in UMMA's no-swizzle core-matrix order: per 8-row
//                               group 256 B = [8 rows x k 0..7][8 rows x k 8..15]
//     u8    [n_pad][128] u8     integer copy for the dp4a rerank (exact mode only)
//     nrm2  [n_pad]      int32  squared norms (exact mode)
//     flags [4]          int32  [0] = 0 iff every value is an integer in [0,255] and every
//                               squared norm < 2^20 ("exact mode"); non-zero = general floats
//   ORB frame: u8 [n_pad][32].
//   Partial top-2 records: uint4 {key0, idx0, key1, idx1} per (pair, split, query row); key is the
//   fp32 bit pattern of the distance (L2) or the integer Hamming distance; absent = {~0u, -1}.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/slamb200.h"

#define SLAMB200_TILE_PAD 256

struct slamb200_desc {
  int kind;
  int n;
  int n_pad;
  float* f32;
  __nv_bfloat16* bf16;
  __nv_bfloat16* augq;
  __nv_bfloat16* augt;
  uint8_t* u8;
  int32_t* nrm2;
  __nv_bfloat16* bf16lo;  // low half of the two-term bf16 split (general-float tensor-core path)
  float* nrmf;            // |row|^2 as fp32
  int32_t* flags;      // device: [0] general-float flag, [2] max |row|^2 (fp32 bits)
  int host_exact;      // -2 unknown, else 1 when flags[0] == 0
  int ready_seen;      // the ready event has been observed complete (no more stream waits)
  cudaEvent_t ready;   // recorded after the prep kernels
  void* slab;          // the one device allocation all the pointers above live in
  size_t slab_bytes;
  int device;          // CUDA device the slab lives on
  int shared;          // 1: the slab is a plain cudaMalloc allocation (exportable over CUDA IPC)
  int imported;        // 1: the slab is another process's allocation mapped here (peer memory)
  alignas(64) unsigned char tmaps[512];  // host copies of 4 CUtensorMaps: main, augq, augt, lo
};

struct slamb200_pts {
  int n;
  float2* xy;  // device
  cudaEvent_t ready;
};

// One pair of a batch as the kernels see it (device array of these).
struct PairArgs {
  const void* t_rows;      // train rows: float* (SIFT exact) or uint8_t* (ORB)
  const int32_t* t_flags;  // train exact-mode flag (SIFT) or nullptr
  const uint8_t* t_u8;     // SIFT: the u8 copy of the rows (valid in exact mode); ORB: t_rows
  int t_n;                 // train row count
  int t_pad;
};

#define ABSENT_KEY 0xFFFFFFFFu

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be scheduled
// while its predecessor in the stream is still running; it must execute PDL_WAIT() before it
// touches anything the predecessor wrote (the wait returns once the predecessor grid has completed
// and its writes are visible).  PDL_TRIGGER() in the predecessor lets the dependent grid start
// being scheduled early; without it the trigger is implicit at grid exit.  Launched without the
// attribute both instructions are no-ops.  Used on the short kernels behind the tcgen05 kernel,
// where launch latency is a visible part of a single-pair call.
#define PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#ifdef __CUDACC__
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// One-time setup that CUDA keeps PER DEVICE (function attributes, __device__ / __constant__
// symbols): a process may hold contexts on several devices (device_set.cu), so "done once" is
// tracked for each of them.  run(f) executes f (returning true on success) the first time it is
// called with the current device; concurrent callers wait for that first setup to finish.
#ifdef __CUDACC__
#include <mutex>
struct PerDeviceOnce {
  std::mutex mu;
  bool done[64] = {};
  template <class F>
  bool run(F f) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return false;
    std::lock_guard<std::mutex> lk(mu);
    if (done[d]) return true;
    if (!f()) return false;
    done[d] = true;
    return true;
  }
};
#endif

// ---- kernel launchers (each in its own .cu) ------------------------------------------------
// ORB: partial top-2 per (pair, split, query).
void launch_orb_knn2(const uint8_t* q, int nq, const PairArgs* pairs, int n_pairs, int n_split,
                     uint4* part, cudaStream_t s);
// SIFT exact fp32 (cv2 summation order); skips pairs whose query and train are both in exact
// mode (flags[0] == 0; the tcgen05 path owns those) unless force != 0.
void launch_sift_exact_knn2(const float* q, const int32_t* q_flags, int nq, const PairArgs* pairs,
                            int n_pairs, int n_split, uint4* part, int force, int norm_l1,
                            cudaStream_t s);
// NORM_L1 on the u8 copy of integer-valued SIFT rows (train sets up to sift_l1_max_train_rows()).
void launch_sift_l1_u8(const uint8_t* q_u8, const int32_t* q_flags, int nq, const PairArgs* pairs,
                       int n_pairs, int n_split, uint4* part, cudaStream_t s);
int sift_l1_max_train_rows();
// merge partials -> raw knn + ratio flags + per-chunk counts, then ordered compaction.
void launch_finalize(const uint4* part, int nq, const PairArgs* pairs, int n_pairs, int n_split,
                     int hamming, double ratio, int32_t* knn_idx, float* knn_dist,
                     uint8_t* flags, int32_t* chunk_cnt, slamb200_dmatch* out, int cap,
                     int32_t* n_out, cudaStream_t s);
int finalize_chunks(int nq);

// RANSAC essential scoring.
struct ScoreParams {
  double ax, ay, bx, by;  // normalisation: x = u*ax + bx
  double mid_lo, mid_hi;  // fast accept / reject bounds on num/den
  double mid;             // midpoint(t, nextafterf(t)) itself (the counting kernel's integer-domain test)
  float t;                // (float)(thr*thr)
};
void launch_normalize_points(const float2* p1, const float2* p2, int total, ScoreParams sp,
                             double4* out, cudaStream_t s);
void launch_gather_normalize(const float2* q_xy, const float2* const* t_xy,
                             const slamb200_dmatch* matches, int cap, const int32_t* n_match,
                             int n_pairs, ScoreParams sp, double4* out, cudaStream_t s);
void launch_score_counts(const double4* npts, const int32_t* m_off, const int32_t* m_cnt,
                         int m_stride, const double* E, int H, int P, ScoreParams sp,
                         int32_t* counts, cudaStream_t s);
void launch_score_best(const int32_t* counts, int H, int P, int min_count, int32_t* best,
                       cudaStream_t s);
void launch_score_mask(const double4* npts, const int32_t* m_off, const int32_t* m_cnt,
                       int m_stride, const double* E, int H, int P, const int32_t* best,
                       ScoreParams sp, uint8_t* mask, cudaStream_t s);
void launch_score_all_masks(const double4* npts, int M, const double* E, int H, ScoreParams sp,
                            uint8_t* masks, cudaStream_t s);

// solvePnPRansac scoring (reprojection error of pose hypotheses, R row-major then t).
struct PnpParams {
  double fx, fy, cx, cy;
  double k[12];    // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4
  float t;         // (float)(reprojectionError^2)
  int dist_level;  // 0 none, 1 first five only, 2 all twelve
};
void launch_pack_pnp_points(const float* obj, const float2* img, int total, double4* out,
                            cudaStream_t s);
void launch_pnp_counts(const double4* pts, const int32_t* m_off, const double* poses, int H, int P,
                       const PnpParams& pp, int32_t* counts, cudaStream_t s);
void launch_pnp_mask(const double4* pts, const int32_t* m_off, const double* poses, int H, int P,
                     const int32_t* best, const PnpParams& pp, uint8_t* mask, cudaStream_t s);
void launch_pnp_all_masks(const double4* pts, int M, const double* poses, int H,
                          const PnpParams& pp, uint8_t* masks, cudaStream_t s);

// ORB descriptors of given keypoints (orb_desc.cu): centre pixel and cosf/sinf of the angle.
struct OrbKeypoint { int cx, cy; float a, b; };
int orb_pattern_upload();
void launch_orb_gray(const uint8_t* img, int rows, int cols, int channels, size_t step, uint8_t* gray,
                     cudaStream_t s);
void launch_orb_blur_gray(const uint8_t* gray, int rows, int cols, float* rowf, uint8_t* blur, cudaStream_t s);
// FAST-9/16 keypoints of a gray frame (fast_detect.cu): score map, suppression, ordered output
int fast_blocks(int rows, int cols);
void launch_fast_detect(const uint8_t* gray, int rows, int cols, int threshold, int nonmax,
                        int16_t* score, int32_t* cnt, float* kp, int cap, cudaStream_t s);
void launch_orb_blur(const uint8_t* img, int rows, int cols, int channels, size_t step, uint8_t* gray,
                     float* rowf, uint8_t* blur, cudaStream_t s);
void launch_orb_desc(const uint8_t* blur, int cols, const OrbKeypoint* kps, int n, int n_pad,
                     uint8_t* desc, cudaStream_t s);

// SIFT descriptors of given octave-0 keypoints (sift_desc.cu): the rounded centre, the window
// radius, cos / sin of the orientation divided by the histogram width, the orientation in degrees
// -- everything calcSIFTDescriptor derives from the keypoint before its sample loop (host, libm).
struct SiftKeypoint { int ptx, pty, radius; float cos_t, sin_t, ori; };
int sift_gauss_upload();
void launch_sift_base(const uint8_t* gray, int rows, int cols, float* rowf, float* base, cudaStream_t s);
int launch_sift_desc(const float* base, int rows, int cols, const SiftKeypoint* kps, int n, int n_pad,
                     float* desc, cudaStream_t s);

// Linear triangulation (triangulate.cpp:17-55): P[v] = 3x4 projection matrix of view v, row-major.
struct TriParams { double P[2][12]; };
void launch_triangulate(const float2* pts1, const float2* pts2, int M, const TriParams& tp,
                        double* points4d, double* points3d, cudaStream_t s);

// SIFT prep: fp32 rows -> bf16 / aug / u8 / norms / exact flag.  n_pad rows are written
// (padding rows get an "infinitely far" augmentation so they never become candidates).
void launch_sift_prep(const float* src, size_t src_stride_floats, int n, int n_pad, float* f32,
                      __nv_bfloat16* bf16, __nv_bfloat16* augq, __nv_bfloat16* augt, uint8_t* u8,
                      int32_t* nrm2, __nv_bfloat16* bf16lo, float* nrmf, int32_t* flags,
                      cudaStream_t s);
void launch_sift_prep_u8(const uint8_t* src_u8, int n, int n_pad, float* f32, __nv_bfloat16* bf16,
                         __nv_bfloat16* augq, __nv_bfloat16* augt, uint8_t* u8, int32_t* nrm2,
                         __nv_bfloat16* bf16lo, float* nrmf, int32_t* flags, cudaStream_t s);
// host_pack.cpp: lossless fp32 -> u8 narrowing with full verification (1 = exact-mode rows).
extern "C" int slamb200_host_pack_u8(const float* src, size_t stride_floats, int n, uint8_t* dst);

// SIFT tcgen05 candidate kernel + dp4a rerank (sift_tc.cu).
struct alignas(64) TcPair {
  unsigned char tmap[384];  // the train frame's CUtensorMaps {main, aug (train role), lo}, read by TMA
  const float* t_f32;       // fp32 rows and norms (general-float path)
  const float* t_nrmf;
  const uint8_t* t_u8;
  const int32_t* t_nrm2;
  const int32_t* t_flags;
  int t_n;
  int t_pad;
};
int tc_encode_tmaps(const void* bf16_dev, const void* augq_dev, const void* augt_dev,
                    const void* bf16lo_dev, int n_pad, void* host_out_512B);
size_t tc_smem_bytes();
int tc_slots(int n_cb_max, int total_tiles, int n_cta);
int launch_sift_tc_candidates(const void* q_tmaps_host_384B, const int32_t* q_flags, int nq,
                              const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                              int total_tiles, int n_cta_pairs, int n_slots, uint4* cand,
                              int32_t* err_flag, float* dbg, int gen, cudaStream_t s, int fp8 = 0,
                              int kinds_known = 0, int wide = 0, const TcPair* inl_pairs_host = nullptr,
                              const int32_t* inl_prefix_host = nullptr);
void launch_sift_gen_rerank(const int32_t* q_flags, const float* q_f32, const float* q_nrmf, int nq,
                            const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                            int n_cta_pairs, int n_slots, int n_split, const uint4* cand, uint4* part,
                            uint2* fb_list, int32_t* fb_count, unsigned long long* fb_part,
                            int32_t* fb_done, cudaStream_t s, int wide = 0);
// scratch of the split fallback scan: fb_part holds SIFT_GEN_FB_ITEMS x 2 keys, fb_done
// SIFT_GEN_FB_ITEMS counters that must be zero when the kernels start
constexpr int SIFT_GEN_FB_ITEMS = 148 * 4;
void launch_sift_rerank(const int32_t* q_flags, const uint8_t* q_u8, const int32_t* q_nrm2, int nq,
                        const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                        int n_cta_pairs, int n_slots, int n_split, const uint4* cand, uint4* part,
                        uint4* work, float2* work_v0, int32_t* work_n, int32_t* err_flag, int prune,
                        double ratio, cudaStream_t s, int orb = 0, int wide = 0);
// merge + best-group rerank + ratio test in one kernel (match output only); then launch_compact
void launch_tc_tail_fused(const int32_t* q_flags, const uint8_t* q_u8, const int32_t* q_nrm2, int nq,
                          const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                          int n_cta_pairs, int n_slots, int n_split, const uint4* cand, uint4* part,
                          int32_t* err_flag, double ratio, int orb, int32_t* knn_idx, float* knn_dist,
                          uint8_t* flags, int32_t* chunk_cnt, cudaStream_t s, int wide = 0);
void launch_compact(int nq, int n_pairs, const int32_t* knn_idx, const float* knn_dist,
                    const uint8_t* flags, const int32_t* chunk_cnt, slamb200_dmatch* out, int cap,
                    int32_t* n_out, cudaStream_t s);
// candidates AND tail in one kernel (exact-mode / ORB match output, no empty shares, kinds known):
// scan and seg_done hold one entry per (pair, 256-row block, CTA of the pair), zeroed when allocated
int launch_sift_tc_match(const void* q_tmaps_host_384B, const int32_t* q_flags, const uint8_t* q_u8,
                         const int32_t* q_nrm2, int nq, const TcPair* pairs_dev,
                         const int32_t* tile_prefix_dev, int n_pairs, int total_tiles, int n_cta_pairs,
                         int n_slots, uint4* cand, int32_t* err_flag, double ratio, int fp8, int wide,
                         unsigned long long* scan, uint32_t epoch, int32_t* seg_done, slamb200_dmatch* out,
                         int cap, int32_t* n_out, cudaStream_t s, const TcPair* inl_pairs_host = nullptr,
                         const int32_t* inl_prefix_host = nullptr);
// batches of up to tc_inline_max() pairs may pass their tables on the host (inl_*_host): they travel
// in the kernel parameters and pairs_dev / tile_prefix_dev are not read
int tc_inline_max();
// the same tail and the ordered compaction (decoupled look-back) in ONE kernel: match lists land in
// out / n_out.  scan: one 64-bit word per (pair, 256-row block), zeroed when allocated; epoch in
// 1 .. 2^30 - 1, different for every launch that shares `scan`
void launch_tc_tail_compact(const int32_t* q_flags, const uint8_t* q_u8, const int32_t* q_nrm2, int nq,
                            const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                            int n_cta_pairs, int n_slots, int n_split, const uint4* cand, const uint4* part,
                            int32_t* err_flag, double ratio, int orb, unsigned long long* scan,
                            uint32_t epoch, slamb200_dmatch* out, int cap, int32_t* n_out, cudaStream_t s,
                            int wide = 0, const TcPair* inl_pairs_host = nullptr,
                            const int32_t* inl_prefix_host = nullptr);
// ORB rows -> tcgen05 operands (sift_prep.cu): e4m3 0/1 bytes [n_pad][256] and the two
// augmentation blocks [n_pad/8][256 B]
void launch_orb_tc_prep(const uint8_t* u8, int n, int n_pad, uint8_t* e4, uint8_t* augq,
                        uint8_t* augt, cudaStream_t s);

void count_launch();
int set_error(int code, const char* msg);
// lets `peer_device` read the context's stream-ordered allocations (descriptor slabs) directly
int ctx_grant_peer_access(slamb200_ctx* c, int peer_device);
int ctx_device(const slamb200_ctx* c);   // fills slamb200_last_error() of the calling thread
#define COUNT_LAUNCH() count_launch()
