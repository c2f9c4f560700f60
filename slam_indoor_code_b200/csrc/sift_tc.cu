// sift_tc.cu -- SIFT L2 k=2 candidates on the 5th-generation tensor cores (tcgen05 + TMEM + TMA),
// followed by the exact integer rerank.  sm_100a only.
//
// Replaces the Q x T x 128 distance computation of cv::BFMatcher(NORM_L2)::knnMatch(query, train,
// 2) (src/mainModule/featureMatching/featureMatchingCPU.cpp:27-40) for descriptor sets in "exact
// mode" (integer-valued rows in [0,255], |row|^2 < 2^20: what cv::SIFT emits).  For those, bf16
// holds every value exactly and every partial sum stays below 2^24, so the fp32 accumulators in
// TMEM are exact and the contraction itself yields d^2/2:
//
//     acc(q,t) = -(q . t)                       8 x tcgen05.mma 256x256x16, A negated
//              + (|q|^2/2) * 1 + 1 * (|t|^2/2)  1 x tcgen05.mma on the 16-column K augmentation
//
// CTA pairs (cluster of 2, tcgen05 cta_group::2): a pair owns a 256 x 256 tile -- each CTA holds
// 128 query rows (its half of M, accumulated in its own TMEM) and loads 128 of the tile's 256
// train rows (its half of N), so every CTA streams only half of each train tile from L2 and the
// tensor core reads 64 B/clk of shared memory per SM instead of 96 (a single-CTA 128 x 256 MMA
// is shared-memory-bandwidth bound at ~66 % of the tensor rate; measured).  One persistent pair
// per two SMs walks a contiguous range of tiles (pair-major, then query block, then train tile).
// Warp roles per CTA: warp 0 = TMA producer (query block resident per segment in 2 stages, train
// half-tiles in a 4-stage ring; both CTAs signal the leader's "full" barriers), warp 1 = MMA
// issuer (leader CTA only, one elected thread; accumulators double-buffered in the 512 TMEM
// columns; completion multicast to both CTAs with tcgen05.commit), warp 2 = TMEM allocator,
// warps 4..11 = epilogue.
//
// Epilogue: a thread owns one query row (TMEM lane) and half of the tile's columns.  Per 32-column
// tcgen05.ld ("chunk") it reduces four groups of 8 columns with 3-input FMNMX trees, takes the
// chunk minimum together with the group that holds it, and pushes (value, group) through a
// branch-free running top-2 over chunks (~0.9 ALU instructions per accumulator, no divergence,
// data-independent timing).  The best element of a row is the minimum of its best chunk; the
// second best is either in that same chunk or is the minimum of the second-best chunk (ties
// resolve to the lowest index at every level).  So the rerank evaluates 32 + 8 = 40 candidates per
// row exactly -- integer dp4a on the u8 copies -- and emits (sqrtf(d^2), index) records that the
// shared finalize kernels turn into the ratio-tested, ordered match list.  A device-side self
// check compares each rerank minimum with the tensor-core value and raises err_flag on any
// mismatch (never expected).
//
// Also in this file:
//  * the match-output tail fused into one kernel (tc_tail_fused_kernel: slot merge, exact
//    ratio-test pruning, best-group rerank, ratio test; compaction follows in finalize.cu);
//  * general-float sets (GEN instantiation: two-term bf16 split, 25 MMAs per tile) with the
//    certified fp32 rerank in OpenCV's summation order and the exact fallback scan;
//  * ORB: the same exact-mode kernel on e4m3 0/1 bytes (kind::f8f6f4, TcParams::fp8) -- Hamming
//    distance is the squared L2 distance of the bit vectors -- with XOR/POPC rerank kernels
//    (featureMatchingCPU.cpp:33-35: BRUTEFORCE_HAMMING).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;             // query rows per CTA (TMEM lanes); the CTA pair covers 256
constexpr int BN = 256;             // train rows per tile (TMEM columns per accumulator stage)
constexpr int BNH = 128;            // train rows each CTA of the pair loads
constexpr int KBLK = 128 * 128;     // bytes of one 64-wide bf16 k-block of 128 rows
constexpr int AUG_OFF = 2 * KBLK;   // K-augmentation block behind the two k-blocks
constexpr int STAGE = 2 * KBLK + 128 * 32;  // 36864: one 128-row operand stage (A or B half)
constexpr int N_ASTAGE = 2;
constexpr int N_BSTAGE = 4;
constexpr int SMEM_BARS = 3072;     // mbarriers, the TMEM base slot, and the 2 KB record-exchange area
constexpr int SMEM_BYTES = (N_ASTAGE + N_BSTAGE) * STAGE + SMEM_BARS + 1024;
// with the tail inside the kernel (TAIL instantiation): the scratch of one 128-row unit behind the
// barrier area, and a ring of finished segments between the epilogue and the tail warps
constexpr int SMEM_TAIL = 3584;
constexpr int SMEM_BYTES_TAIL = SMEM_BYTES + SMEM_TAIL;
constexpr int JOB_RING = 8;
constexpr int BAR_TAIL = 5;         // named barrier of the two tail warps (1..4: the epilogue's quarter pairs)
// general-float variant: operands carry a hi and a lo bf16 half (two-term split), 24 + 1 MMAs per
// tile: hi.hi + hi.lo + lo.hi + augmentation
constexpr int STAGE_G = 4 * KBLK + 128 * 32;   // 69632
constexpr int AUG_OFF_G = 4 * KBLK;
constexpr int N_ASTAGE_G = 1;
constexpr int N_BSTAGE_G = 2;
constexpr int SMEM_BYTES_G = (N_ASTAGE_G + N_BSTAGE_G) * STAGE_G + SMEM_BARS + 1024;
constexpr int EPI_WARP0 = 4;
#ifndef SLAMB200_EPI_WARPS
#define SLAMB200_EPI_WARPS 8   // 16 was measured slower (19.0 vs 17.2 us/pair): register cap + barrier traffic
#endif
constexpr int N_EPI_WARPS = SLAMB200_EPI_WARPS;       // 8 or 16
constexpr int TC_THREADS = (EPI_WARP0 + N_EPI_WARPS) * 32;
constexpr int COL_SPLITS = N_EPI_WARPS / 4;            // column ranges per row (2 halves / 4 quarters)
constexpr int COLS_PER_WARP = BN / COL_SPLITS;         // 128 or 64
constexpr int CHUNKS = COLS_PER_WARP / 32;             // 32-column tcgen05.ld chunks per tile per warp
constexpr int GROUP = 8;            // columns per candidate group
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared address

struct alignas(64) TcParams {
  CUtensorMap q_tmap[3];       // query maps: [0] main (hi), [1] aug (query role), [2] lo
  const TcPair* pairs;         // per pair: train maps / aug / sizes
  const int32_t* tile_prefix;  // [P+1] tiles before pair p
  const int32_t* q_flags;
  int n_pairs;
  int n_rb;                    // query row blocks of 256 rows (one per CTA pair)
  int total_tiles;
  int n_slots;                 // candidate slots per (pair, row) = 2 * max segments per row block
  int nq_pad;                  // n_rb * 256
  uint4* cand;                 // [pair][slot][nq_pad]
  float* dbg;                  // optional raw accumulator dump of pair 0's first tile [256][256]
  int32_t* err_flag;
  int mode;                    // 0 = product; 1..7 = timing experiments (results invalid; MODES builds only)
  int fp8;                     // operands are e4m3 bytes (ORB bits expanded to 0/1): kind::f8f6f4, K = 32
  int kinds_known;             // the host knows every pair of this launch is of this kernel's kind
                               // (integer-valued / general floats): no flag loads in the tile walk
  int wide;                    // tiles x CTA pairs can exceed 2^32: 64-bit share arithmetic
};

// Small batches (the per-pair drop-in call above all) carry their pair table in the kernel
// parameters instead of a device buffer filled by a copy in front of the launch: the tensor maps are
// then read from parameter space, as the query's always are.
constexpr int TC_INLINE_MAX = 4;
struct alignas(64) InlineTables {
  TcPair pairs[TC_INLINE_MAX];
  int32_t prefix[TC_INLINE_MAX + 1];
  int n;                       // 0: the tables are in device memory (TcParams / RerankParams pointers)
};

// Timeline probe (MODES builds, mode 8 = the product's behaviour + eight time stamps per CTA; read
// back by slamb200_dbg_tc_trace, tools/tc_trace.py): where a short launch spends its time.
__device__ unsigned long long g_tc_trace[160 * 8];
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(k) do { if (trace) g_tc_trace[(blockIdx.x % 160) * 8 + (k)] = global_ns(); } while (0)

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Waits with a watchdog: a barrier that does not flip within ~2 s (a lost TMA transaction, a bad
// descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("slamb200: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// 2-CTA TMA load: lands in this CTA's shared memory, signals the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  // completion of all prior MMAs of this thread -> arrive on `bar` in both CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 with 8-bit operands: plain bytes in shared memory, K = 32 per instruction
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// K-major SWIZZLE_128B operand: rows of 128 B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                   // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // SBO
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}
// K-major no-swizzle ("interleave") operand of K = 16: core matrices of 8 rows x 16 B; the two
// K halves are LBO = 128 B apart, 8-row groups SBO = 256 B apart.
__device__ __forceinline__ uint64_t desc_interleave(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::i8: u8 x u8 -> s32 (c_format 2), K = 32 per instruction
__host__ __device__ constexpr uint32_t idesc_u8(int m, int n) {
  return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::f8f6f4: a/b format 0 = E4M3, fp32 accumulators
__host__ __device__ constexpr uint32_t idesc_e4m3(int m, int n, int a_neg) {
  return (1u << 4) | ((uint32_t)a_neg << 13) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_neg) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_neg << 13) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// ---- tile walk: every role iterates the same sequence -----------------------------------------
// Tile walk.  Every CTA pair visits every frame pair in order and takes a contiguous share of
// THAT frame pair's tiles (row-block major): at any moment the whole grid is streaming the same
// train set, which therefore stays L2-resident (2.9 MB) instead of every CTA pair dragging its own
// train set through HBM; contiguous shares keep the segments (same query block) long.  The share
// owner rotates with the frame-pair index so that rounding does not always favour the same CTAs.
// Share arithmetic.  begin(c) = floor(total * c / n); the owner of tile x is the share c with
// begin(c) <= x < begin(c + 1), i.e. the largest c with total * c < (x + 1) * n:
// c = floor(((x + 1) * n - 1) / total).  32-bit unless the launch says the products can overflow
// (`wide`, uniform): these run on the single threads that feed the TMA unit and the tensor core
// at every frame-pair boundary, where a 64-bit division subroutine is a visible bubble.
__device__ __forceinline__ int share_begin(int total, int c, int n, bool wide) {
  return wide ? (int)(((unsigned long long)(unsigned)total * (unsigned)c) / (unsigned)n)
              : (int)(((unsigned)total * (unsigned)c) / (unsigned)n);
}
__device__ __forceinline__ int owner_cta(int total, int n, int x, bool wide) {
  return wide ? (int)(((unsigned long long)(unsigned)(x + 1) * (unsigned)n - 1ull) / (unsigned)total)
              : (int)(((unsigned)(x + 1) * (unsigned)n - 1u) / (unsigned)total);
}

struct TileIter {
  int pair, rb, cb, n_cb;
  int tile, end;            // local tile index inside the current frame pair, and the share's end
  int slot_c;               // this CTA pair's (rotated) position for the current frame pair
  int n_tiles;              // tiles of the current frame pair
  int cta, n_cta;
  int t_n;                  // rows of the current train set (the last column tile may be partial)
  int scan_p, scan_slot;    // next frame pair the walk looks at, and this CTA pair's position there
  int pfx0, pfx1, pfx2;     // tile_prefix[scan_p], [scan_p + 1], [scan_p + 2] (read one pair ahead)
  const CUtensorMap* tmap;  // train maps: [0] main (hi), [1] aug (train role), [2] lo
  const TcPair* pairs;      // the launch's pair table and tile prefix (device memory or parameters)
  const int32_t* prefix;
  bool gen;                 // which kind of frame pair this walk visits
  // The pair table (and the tensor maps inside it) is rewritten by the host before every launch:
  // make the TMA unit's descriptor reads observe those generic-proxy writes.
  __device__ void acquire_maps() const {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap + 1) : "memory");
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap + 2) : "memory");
  }
  // positions on the first non-empty share at or after frame pair scan_p; false when none is left
  __device__ bool seek(const TcParams& P) {
    for (; scan_p < P.n_pairs;) {
      const int p = scan_p;
      const int nt = pfx1 - pfx0;
      const int slot = scan_slot;
      // advance the scan state first: the prefix of the pair after next is requested now and is
      // not needed before the next seek (one whole share of MMAs later)
      scan_p = p + 1;
      scan_slot = slot + 1 == n_cta ? 0 : slot + 1;
      pfx0 = pfx1;
      pfx1 = pfx2;
      if (p + 3 <= P.n_pairs) pfx2 = prefix[p + 3];
      if (nt == 0) continue;
      // exact-mode pairs (integer-valued query and train) and general-float pairs are walked by
      // different instantiations of the kernel
      if (!P.kinds_known && ((P.q_flags[0] | pairs[p].t_flags[0]) != 0) != gen) continue;
      const bool wide = P.wide != 0;
      const int t0 = share_begin(nt, slot, n_cta, wide);
      const int t1 = share_begin(nt, slot + 1, n_cta, wide);
      if (t0 >= t1) continue;
      pair = p;
      slot_c = slot;
      n_tiles = nt;
      tile = t0;
      end = t1;
      n_cb = (int)((unsigned)nt / (unsigned)P.n_rb);
      rb = (int)((unsigned)t0 / (unsigned)n_cb);
      cb = t0 - rb * n_cb;
      t_n = pairs[p].t_n;
      tmap = reinterpret_cast<const CUtensorMap*>(pairs[p].tmap);
      return true;
    }
    pair = P.n_pairs;
    return false;
  }
  __device__ void init(const TcParams& P, const TcPair* pairs_, const int32_t* prefix_, int cta_, int n_cta_,
                       bool gen_) {
    cta = cta_; n_cta = n_cta_; gen = gen_;
    pairs = pairs_; prefix = prefix_;
    pair = 0; n_cb = 1; rb = 0; cb = 0; tile = 0; end = 0; slot_c = 0; n_tiles = 0; t_n = 0;
    tmap = nullptr;
    scan_p = 0;
    scan_slot = cta_;     // (cta + p) % n_cta at p = 0, kept incrementally
    pfx0 = prefix[0];
    pfx1 = P.n_pairs >= 1 ? prefix[1] : pfx0;
    pfx2 = P.n_pairs >= 2 ? prefix[2] : pfx1;
    seek(P);
  }
  __device__ bool valid() const { return tile < end; }
  // valid columns of the current tile rounded up to whole 32-column chunks (the rows up to the
  // set's 256-row padding exist and carry an unreachable augmentation): the MMA's N
  __device__ int n_eff() const {
    const int left = t_n - cb * BN;
    return left >= BN ? BN : ((left + 31) & ~31);
  }
  // returns true when the next tile starts a new (pair, row block) segment
  __device__ bool next(const TcParams& P) {
    tile++;
    if (tile >= end) {
      if (!seek(P)) { tile = 0; end = 0; }
      return true;
    }
    if (++cb < n_cb) return false;
    cb = 0;
    ++rb;
    return true;
  }
};

// Running top-2 over chunks, branch-free.  (m1, i1) best chunk minimum and the group that holds
// it, s1 = the second-smallest group minimum INSIDE that best chunk, (m2, i2) second-best chunk.
__device__ __forceinline__ void top2_chunk_insert(float g, int gid, float g2nd, float& m1, int& i1,
                                                  float& s1, float& m2, int& i2) {
  const bool lt1 = g < m1;
  const bool lt2 = g < m2;
  m2 = lt1 ? m1 : (lt2 ? g : m2);
  i2 = lt1 ? i1 : (lt2 ? gid : i2);
  m1 = lt1 ? g : m1;
  i1 = lt1 ? gid : i1;
  s1 = lt1 ? g2nd : s1;
}

#define TMEM_LD32(taddr, v)                                                                     \
  asm volatile(                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                 \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                 \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),     \
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), \
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),           \
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),           \
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])            \
      : "r"(taddr))

// tcgen05.wait::ld with the loaded registers as in/out operands: the compiler cannot move any
// use of them above the wait.
#define TMEM_WAIT32(v)                                                                          \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                 \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]),        \
                 "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]),      \
                 "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]),  \
                 "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),  \
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),  \
                 "+r"(v[30]), "+r"(v[31])::"memory")

// One 32-column chunk of a row: the chunk minimum, the 8-column group that holds it (lowest group
// on ties) and the second-smallest of the four group minima go through the running top-2 over
// chunks.  gid = column / 8 of the minimum's group, so chunk = gid >> 2.  Strict '<' keeps the
// earlier chunk on equal values.
// General-float variant: the accumulators are approximations, so the record keeps the best FOUR
// chunks, the fifth-best chunk minimum (a lower bound on every column outside those four) and
// the smallest second group minimum of any chunk (with the former: a lower bound on every column
// outside the best 8-column group of each kept chunk), which the rerank needs to certify its
// answer after reading 32, or else 128, candidate rows.
// Epilogue running state of one thread (row x column range).  Exact mode uses m[0..1], i[0..1]
// and s (second group minimum inside the best chunk); the general-float variant keeps the best
// FOUR chunks (m[0..3], i[0..3]), s = the fifth-best chunk minimum and s2 = the smallest second
// group minimum over all chunks.
struct EpiState {
  float m[4];
  int i[4];
  float s, s2;
  __device__ __forceinline__ void reset() {
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int k = 0; k < 4; k++) { m[k] = inf; i[k] = -1; }
    s = inf;
    s2 = inf;
  }
};

template <bool GEN>
__device__ __forceinline__ void process_chunk(const uint32_t (&v)[32], int gid0, EpiState& st) {
  float g[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float a = fmin3(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]),
                          __uint_as_float(v[8 * j + 2]));
    const float b = fmin3(__uint_as_float(v[8 * j + 3]), __uint_as_float(v[8 * j + 4]),
                          __uint_as_float(v[8 * j + 5]));
    g[j] = fmin3(a, b, fminf(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
  }
  const float m01 = fminf(g[0], g[1]), m23 = fminf(g[2], g[3]);
  const int j01 = g[1] < g[0] ? gid0 + 1 : gid0;
  const int j23 = g[3] < g[2] ? gid0 + 3 : gid0 + 2;
  const float cm = fminf(m01, m23);
  const int gid = m23 < m01 ? j23 : j01;
  const float M01 = fmaxf(g[0], g[1]), M23 = fmaxf(g[2], g[3]);
  const float c2 = fmin3(fmaxf(m01, m23), M01, M23);  // second smallest of the four
  if (GEN) {
    // branch-free sorted insert of the chunk into the best four; the displaced fourth feeds the
    // bound (s >= m[3] always: s is a minimum over chunks that were not among the four smallest).
    // s2 = the smallest "second group minimum" of any chunk: together with s it bounds every
    // column outside the best GROUP of each of the four kept chunks.
    const bool l0 = cm < st.m[0], l1 = cm < st.m[1], l2 = cm < st.m[2], l3 = cm < st.m[3];
    st.s = l3 ? st.m[3] : fminf(st.s, cm);
    st.s2 = fminf(st.s2, c2);
    st.m[3] = l2 ? st.m[2] : (l3 ? cm : st.m[3]);
    st.i[3] = l2 ? st.i[2] : (l3 ? gid : st.i[3]);
    st.m[2] = l1 ? st.m[1] : (l2 ? cm : st.m[2]);
    st.i[2] = l1 ? st.i[1] : (l2 ? gid : st.i[2]);
    st.m[1] = l0 ? st.m[0] : (l1 ? cm : st.m[1]);
    st.i[1] = l0 ? st.i[0] : (l1 ? gid : st.i[1]);
    st.m[0] = l0 ? cm : st.m[0];
    st.i[0] = l0 ? gid : st.i[0];
  } else {
    top2_chunk_insert(cm, gid, c2, st.m[0], st.i[0], st.s, st.m[1], st.i[1]);
  }
}

// ---- rerank: merge the slot records, prune by the ratio test, evaluate the survivors exactly ---
struct RerankParams {
  const uint8_t* q_u8;
  const int32_t* q_nrm2;
  const int32_t* q_flags;
  const TcPair* pairs;
  const int32_t* tile_prefix;
  const uint4* cand;
  int nq, nq_pad, n_slots, n_pairs, n_split, n_cta;
  int wide;          // 64-bit share arithmetic (as TcParams::wide)
  int prune;         // 1: rows that cannot pass the ratio test skip the exact evaluation
  int orb;           // ORB sets: the accumulators are Hamming/2, q_u8 / t_u8 are the 32-byte rows
  double ratio;
  uint4* part;       // [pair][n_split][nq]: split 0 gets the record (pre-set to "absent")
  uint4* work;       // survivors: {pair, row, gid of the best chunk, gid of the second chunk}
  float2* work_v0;   // {tensor-core minimum (self check), smallest d^2/2 outside the best group}
  int32_t* work_n;   // number of survivors
  int32_t* err_flag;
};

constexpr float ORB_PAD_THRESHOLD = 200.0f;
__device__ __forceinline__ bool lt_fi(float va, int ia, float vb, int ib) {
  return va < vb || (va == vb && ia < ib);
}

__device__ __forceinline__ void part_top2_insert(uint4& r, uint32_t k, uint32_t i) {
  const bool lt2 = k < r.z || (k == r.z && i < r.w);
  if (lt2) {
    const bool lt1 = k < r.x || (k == r.x && i < r.y);
    if (lt1) { r.z = r.x; r.w = r.y; r.x = k; r.y = i; }
    else { r.z = k; r.w = i; }
  }
}


// ================= tail + ordered compaction as ONE unit of work =================
// tail_unit: everything between the slot records of ROWS consecutive query rows of one frame pair
// and their entries in the pair's match list -- slot merge, exact ratio-test pruning, best-group
// rerank, ratio test (the arithmetic of tc_tail_fused_kernel above, operation for operation) and
// the ORDERED compaction (ascending queryIdx, getGoodMatches' loop order,
// featureMatchingCommon.cpp:43-49), done here by a decoupled look-back over the units of the pair
// instead of a second kernel: a unit publishes its kept count, sums the counts of the units before
// it back to the nearest published prefix, publishes its own prefix and writes its matches behind
// it.  One 64-bit word per (pair, unit): {launch epoch : 30, state : 2, count : 32}; the epoch
// makes words of earlier launches read as "not there yet", so the array is never cleared.
// NT threads work on a unit (a thread block of 256 on 256 rows in tc_tail_compact_kernel; the two
// tail warps of a tcgen05 CTA on the CTA's 128 rows when the tail runs inside sift_tc_kernel).
template <int ROWS>
struct TailSmem {
  float v0[ROWS], L[ROWS], d0[ROWS], d1[ROWS];
  int g0[ROWS], idx[ROWS];
  uint32_t keep_mask[ROWS / 32];
  int keep_off[ROWS / 32];
  uint16_t list[ROWS];
  int n_surv, base;
  int flag;                   // in-kernel tail: "this CTA flushed the unit's last segment"
};

struct CompactArgs {
  unsigned long long* scan;   // [pair][units per pair] look-back words
  uint32_t epoch;             // 1 .. 2^30 - 1, different for every launch that shares `scan`
  slamb200_dmatch* out;       // [pair][cap]
  int cap;
  int32_t* n_out;             // [pair]
  // two-kernel form (COMPACT = false) instead: what compact_kernel consumes
  int32_t* knn_idx;           // [pair][nq][2]
  float* knn_dist;            // [pair][nq][2]
  uint8_t* flags;             // [pair][nq]
  int32_t* chunk_cnt;         // [pair][units per pair]
};

__device__ __forceinline__ unsigned long long scan_ld(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void scan_st(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr uint32_t SCAN_AGG = 1, SCAN_PREFIX = 2;
__device__ __forceinline__ unsigned long long scan_word(uint32_t epoch, uint32_t state, uint32_t v) {
  return ((unsigned long long)((epoch << 2) | state) << 32) | v;
}

template <int BAR, int NT>
__device__ __forceinline__ void group_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}

// Exclusive prefix of unit `unit`'s count over the units of its pair (one warp, all 32 lanes).
// Units before it are either published already or will be without waiting for this one (see the
// callers for why), so the spin terminates.
__device__ __forceinline__ int scan_lookback(unsigned long long* scan_pair, uint32_t epoch, int unit, int n,
                                             int lane) {
  if (unit == 0) {
    if (lane == 0) scan_st(scan_pair, scan_word(epoch, SCAN_PREFIX, (uint32_t)n));
    return 0;
  }
  if (lane == 0) scan_st(scan_pair + unit, scan_word(epoch, SCAN_AGG, (uint32_t)n));
  int base = 0;
  for (int j = unit - 1;;) {
    const int k = j - lane;   // lane 0 looks at the nearest unit
    unsigned long long w = scan_word(epoch, SCAN_PREFIX, 0);   // before unit 0: an empty prefix
    if (k >= 0) w = scan_ld(scan_pair + k);
    const uint32_t hi = (uint32_t)(w >> 32);
    const bool ready = (hi >> 2) == epoch && (hi & 3u) != 0;
    const unsigned rm = __ballot_sync(0xffffffffu, ready);
    const unsigned pm = __ballot_sync(0xffffffffu, ready && (hi & 3u) == SCAN_PREFIX);
    const int first = pm ? __ffs(pm) - 1 : 32;                 // nearest published prefix in the window
    const unsigned need = first >= 32 ? 0xffffffffu : ((2u << first) - 1u);
    if ((rm & need) != need) {
      __nanosleep(64);
      continue;
    }
    int v = lane <= first ? (int)(uint32_t)w : 0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    base += v;
    if (first < 32) break;
    j -= 32;
  }
  if (lane == 0) scan_st(scan_pair + unit, scan_word(epoch, SCAN_PREFIX, (uint32_t)(base + n)));
  return base;
}

template <bool ORB, int NT, int ROWS, int BAR, int SU, bool COMPACT = true>
__device__ __forceinline__ void tail_unit(const RerankParams& R, const CompactArgs& C, const TcPair* pairs_tab,
                                          const int32_t* prefix_tab, int pair, int rb, int q0, int unit,
                                          int units_per_pair, bool gen_pair, TailSmem<ROWS>& sm, int tid) {
  static_assert(ROWS % NT == 0 && NT % 32 == 0 && ROWS / 32 <= 32, "unit geometry");
  const int lane = tid & 31, warp = tid >> 5;
  const TcPair* pr = pairs_tab + pair;
  const int t_n = pr->t_n;
  const uint8_t* __restrict__ t_u8 = pr->t_u8;
  const int32_t* __restrict__ t_nrm2 = pr->t_nrm2;
  const float INF = __int_as_float(0x7f800000);
  if (tid == 0) sm.n_surv = 0;
  for (int r = tid; r < ROWS; r += NT) { sm.idx[r] = -1; sm.d0[r] = 0.f; sm.d1[r] = -1.f; }
  group_sync<BAR, NT>();
  if (gen_pair) {
    // general-float pair: its records come from the certified rerank; finalize them
    for (int r = tid; r < ROWS; r += NT) {
      const int q = q0 + r;
      if (q >= R.nq) continue;
      uint4 rr = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
      for (int sp = 0; sp < R.n_split; sp++) {
        const uint4 p = R.part[((size_t)pair * R.n_split + sp) * R.nq + q];
        if (p.y != 0xFFFFFFFFu) part_top2_insert(rr, p.x, p.y);
        if (p.w != 0xFFFFFFFFu) part_top2_insert(rr, p.z, p.w);
      }
      if (rr.y != 0xFFFFFFFFu) {
        sm.idx[r] = (int)rr.y;
        sm.d0[r] = __uint_as_float(rr.x);
        // (a NaN second distance fails `d1 >= 0` below exactly as it fails the ratio test)
        if (rr.w != 0xFFFFFFFFu) sm.d1[r] = __uint_as_float(rr.z);
      }
    }
  } else {
    // slots the tcgen05 kernel wrote for this (pair, query block): one merged record per share
    int n_valid = 0;
    {
      const int n_tiles = prefix_tab[pair + 1] - prefix_tab[pair];
      if (n_tiles > 0) {
        const int n_rb = R.nq_pad / 256;
        const int n_cb = n_tiles / n_rb;
        const int first = owner_cta(n_tiles, R.n_cta, rb * n_cb, R.wide != 0);
        const int last = owner_cta(n_tiles, R.n_cta, (rb + 1) * n_cb - 1, R.wide != 0);
        n_valid = last - first + 1;
        if (n_valid > R.n_slots) n_valid = R.n_slots;
      }
    }
    for (int r = tid; r < ROWS; r += NT) {
      const int q = q0 + r;
      bool survive = false;
      if (q < R.nq) {
        float v0 = INF, v1 = INF, s0 = INF;
        int g0 = 0xFFFF, g1 = 0xFFFF;
        for (int sb = 0; sb < n_valid; sb += 4) {
          uint4 recs[4];
#pragma unroll
          for (int j = 0; j < 4; j++)   // independent loads first (L2: the records may have been
                                        // written by other SMs during this very kernel)
            recs[j] = sb + j < n_valid ? __ldcg(R.cand + ((size_t)pair * R.n_slots + sb + j) * R.nq_pad + q)
                                       : make_uint4(0x7f800000u, 0x7f800000u, 0x7f800000u, 0xFFFFFFFFu);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint4 rec = recs[j];
            const float a = __uint_as_float(rec.x), b = __uint_as_float(rec.y);
            float sa = __uint_as_float(rec.z);
            int ia = (int)(rec.w & 0xFFFFu), ib = (int)(rec.w >> 16);
            if (ORB) {
              if (a > ORB_PAD_THRESHOLD) ia = 0xFFFF;
              if (b > ORB_PAD_THRESHOLD) ib = 0xFFFF;
              if (sa > ORB_PAD_THRESHOLD) sa = INF;
            }
            if (ia != 0xFFFF) {
              if (lt_fi(a, ia, v0, g0)) { v1 = v0; g1 = g0; v0 = a; g0 = ia; s0 = sa; }
              else if (lt_fi(a, ia, v1, g1)) { v1 = a; g1 = ia; }
            }
            if (ib != 0xFFFF) {
              if (lt_fi(b, ib, v1, g1)) { v1 = b; g1 = ib; }
            }
          }
        }
        const bool has0 = g0 != 0xFFFF;
        survive = has0;
        const float L = fminf(s0, v1);
        if (R.prune && has0 && L < INF) {
          const float d0 = ORB ? 2.0f * v0 : sqrtf(2.0f * v0), D1 = ORB ? 2.0f * L : sqrtf(2.0f * L);
          if (!((double)d0 < __dmul_rn(R.ratio, (double)D1))) survive = false;   // exact pruning
        }
        sm.v0[r] = v0; sm.L[r] = L; sm.g0[r] = g0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, survive);
      int at = 0;
      if (lane == 0 && bal) at = atomicAdd(&sm.n_surv, __popc(bal));
      at = __shfl_sync(0xffffffffu, at, 0);
      if (survive) sm.list[at + __popc(bal & ((1u << lane) - 1))] = (uint16_t)r;
    }
    group_sync<BAR, NT>();
    const int n_surv = sm.n_surv;
    // survivors: eight lanes per row, lane c ending up with candidate c of the row's best group
    // (coalesced 16-byte slices + reduce-scatter: the group's eight train rows are 1 KB of consecutive
    // bytes, so each load instruction takes ONE 128-byte row per survivor row, lane c its 16-byte slice;
    // a lane reading its own candidate row touched 32 lines per instruction and the kernel sat on the
    // L1 tag stage).  SU rows per lane group
    // and pass, all their loads issued before the first is used: a block of a single-pair call is
    // a chain of dependent memory round trips, and SU = 3 turns its three survivor passes into one.
    const int cand = lane & 7;
    for (int i0 = 0; i0 < n_surv; i0 += SU * (NT / 8)) {
      bool have[SU], ok[SU];
      int rr[SU], gg[SU];
      uint4 qv[SU], tv[SU][ORB ? 2 : 8], qw[SU];
      uint32_t nn[SU];
#pragma unroll
      for (int u = 0; u < SU; u++) {
        const int i = i0 + u * (NT / 8) + (tid >> 3);
        have[u] = i < n_surv;
        rr[u] = have[u] ? sm.list[i] : 0;
        gg[u] = have[u] ? sm.g0[rr[u]] : 0;
        const int col0 = gg[u] * GROUP;
        const int col = col0 + cand;
        ok[u] = have[u] && col < t_n;
        const int qq = q0 + rr[u];
        if (ORB) {
          const int cc = ok[u] ? col : 0;
          const uint4* tp = reinterpret_cast<const uint4*>(t_u8 + (size_t)cc * 32);
          const uint4* qp = reinterpret_cast<const uint4*>(R.q_u8 + (size_t)qq * 32);
          tv[u][0] = tp[0]; tv[u][1] = tp[1]; qv[u] = qp[0]; qw[u] = qp[1];
          nn[u] = 0;
        } else {
          // (rows up to the set's 256-row padding exist: a group never leaves the allocation)
          const uint4* tp = reinterpret_cast<const uint4*>(t_u8 + (size_t)col0 * 128) + cand;
          qv[u] = *(reinterpret_cast<const uint4*>(R.q_u8 + (size_t)qq * 128) + cand);
          qw[u] = qv[u];
#pragma unroll
          for (int k = 0; k < 8; k++) tv[u][ORB ? 0 : k] = tp[8 * k];
          nn[u] = (uint32_t)t_nrm2[ok[u] ? col : 0] + (uint32_t)R.q_nrm2[qq];
        }
      }
#pragma unroll
      for (int u = 0; u < SU; u++) {
        uint32_t dist;   // exact integer distance of this lane's candidate: d^2 (SIFT) or Hamming (ORB)
        if (ORB) {
          const uint4 t0 = tv[u][0], t1 = tv[u][ORB ? 1 : 0], qa = qv[u], qb = qw[u];
          dist = (uint32_t)(__popc(qa.x ^ t0.x) + __popc(qa.y ^ t0.y) + __popc(qa.z ^ t0.z) + __popc(qa.w ^ t0.w) +
                            __popc(qb.x ^ t1.x) + __popc(qb.y ^ t1.y) + __popc(qb.z ^ t1.z) + __popc(qb.w ^ t1.w));
        } else {
          uint32_t pd[8];   // pd[k]: this lane's slice of q . (candidate row k)
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint4 t = tv[u][ORB ? 0 : k];
            uint32_t d = __dp4a(qv[u].x, t.x, 0u);
            d = __dp4a(qv[u].y, t.y, d);
            d = __dp4a(qv[u].z, t.z, d);
            pd[k] = __dp4a(qv[u].w, t.w, d);
          }
          const bool b4 = (cand & 4) != 0, b2 = (cand & 2) != 0, b1 = (cand & 1) != 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t send = b4 ? pd[j] : pd[j + 4], keep = b4 ? pd[j + 4] : pd[j];
            pd[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const uint32_t send = b2 ? pd[j] : pd[j + 2], keep = b2 ? pd[j + 2] : pd[j];
            pd[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          {
            const uint32_t send = b1 ? pd[0] : pd[1], keep = b1 ? pd[1] : pd[0];
            pd[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
          dist = nn[u] - 2u * pd[0];
        }
        uint32_t k0 = ok[u] ? ((dist << 3) | (uint32_t)cand) : 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
#pragma unroll
        for (int off = 1; off <= 4; off <<= 1) {
          const uint32_t o0 = __shfl_xor_sync(0xffffffffu, k0, off);
          const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, off);
          const uint32_t hi = max(k0, o0);
          k0 = min(k0, o0);
          k1 = min(hi, min(k1, o1));
        }
        if (have[u] && cand == 0 && k0 != 0xFFFFFFFFu) {
          const int r = rr[u];
          const uint32_t x0 = k0 >> 3;
          // self check: the group's exact minimum must equal twice the tensor-core value
          if ((float)x0 != 2.0f * sm.v0[r]) atomicOr(R.err_flag, 1);
          // second distance: inside the group, or the bound from outside it (exact values both)
          const float Lr = sm.L[r];
          float x1 = Lr < INF ? 2.0f * Lr : -1.0f;
          if (k1 != 0xFFFFFFFFu) {
            const float x2 = (float)(k1 >> 3);
            x1 = (x1 < 0.0f || x2 < x1) ? x2 : x1;
          }
          sm.idx[r] = gg[u] * GROUP + (int)(k0 & 7u);
          sm.d0[r] = ORB ? (float)x0 : sqrtf((float)x0);
          sm.d1[r] = ORB ? x1 : (x1 >= 0.0f ? sqrtf(x1) : -1.0f);
        }
      }
    }
  }
  group_sync<BAR, NT>();
  // ratio test (double, strict), 32 consecutive rows per warp and pass
  for (int r = tid; r < ROWS; r += NT) {
    const int q = q0 + r;
    const float d1 = sm.d1[r];
    const bool keep = q < R.nq && sm.idx[r] >= 0 && d1 >= 0.0f &&
                      (double)sm.d0[r] < __dmul_rn(R.ratio, (double)d1);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sm.keep_mask[r >> 5] = bal;
  }
  group_sync<BAR, NT>();
  if (!COMPACT) {
    // two-kernel form: per-row verdicts and the unit's kept count for compact_kernel
    for (int r = tid; r < ROWS; r += NT) {
      const int q = q0 + r;
      if (q >= R.nq) continue;
      const size_t o = ((size_t)pair * R.nq + q) * 2;
      C.knn_idx[o] = sm.idx[r];
      C.knn_dist[o] = sm.d0[r];
      C.flags[(size_t)pair * R.nq + q] = (uint8_t)((sm.keep_mask[r >> 5] >> (r & 31)) & 1u);
    }
    if (warp == 0) {
      int c = lane < ROWS / 32 ? __popc(sm.keep_mask[lane]) : 0;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
      if (lane == 0) C.chunk_cnt[pair * units_per_pair + unit] = c;
    }
    return;
  }
  if (warp == 0) {
    const int c = lane < ROWS / 32 ? __popc(sm.keep_mask[lane]) : 0;
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    const int n = __shfl_sync(0xffffffffu, incl, 31);
    if (lane < ROWS / 32) sm.keep_off[lane] = incl - c;
    const int base = scan_lookback(C.scan + (size_t)pair * units_per_pair, C.epoch, unit, n, lane);
    if (lane == 0) {
      sm.base = base;
      if (unit == units_per_pair - 1) C.n_out[pair] = base + n;
    }
  }
  group_sync<BAR, NT>();
  {
    const int base = sm.base;
    int4* outp = reinterpret_cast<int4*>(C.out + (size_t)pair * C.cap);
    for (int r = tid; r < ROWS; r += NT) {
      const uint32_t m = sm.keep_mask[r >> 5];
      if (!((m >> (r & 31)) & 1u)) continue;
      const int off = base + sm.keep_off[r >> 5] + __popc(m & ((1u << (r & 31)) - 1u));
      if (off < C.cap) outp[off] = make_int4(q0 + r, sm.idx[r], 0, __float_as_int(sm.d0[r]));
    }
  }
}

// What the tail warps of the TAIL instantiation need beside TcParams.
struct TcTail {
  RerankParams R;
  CompactArgs C;              // scan: one word per (pair, row block, CTA of the pair)
  int32_t* seg_done;          // [pair][row block][CTA of the pair]: segments flushed so far; zero
                              // before the first launch, left at zero by every launch
};
static_assert(sizeof(TailSmem<128>) <= SMEM_TAIL, "tail scratch does not fit its shared-memory area");

// MODES = false is the product: `mode` is the constant 0 and none of the timing experiments below
// exists in the instruction stream.  MODES = true (tools/tc_modes.py through
// slamb200_dbg_set_tc_mode) selects role ablations at run time; their results are void.
// TAIL = true (exact-mode match output, every share non-empty): warps 2 and 3 turn the slot records
// of a finished (pair, row block) into its entries of the match list while the other roles go on
// (tail_unit above) -- no tail kernel behind this one, and the records are read back while they
// are still in L2.  The epilogue hands every flushed segment to them through a ring in shared
// memory (mbarrier full / empty pairs, like the operand stages); they count the segments of a
// row block in a global counter, and the CTA that flushes the last one owns the unit.  Its ring is
// in walk order, so is every other CTA's: the smallest unit that has not published its count is
// never behind a larger one, which is why the look-back's spin always ends.
template <bool DBG, bool GEN, bool MODES, bool TAIL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
sift_tc_kernel(const __grid_constant__ TcParams P, const __grid_constant__ TcTail TT,
               const __grid_constant__ InlineTables IT) {
  static_assert(!TAIL || (!GEN && !DBG && !MODES), "the in-kernel tail exists for the exact-mode product only");
  constexpr int STG = GEN ? STAGE_G : STAGE;
  constexpr int NA = GEN ? N_ASTAGE_G : N_ASTAGE;
  constexpr int NB = GEN ? N_BSTAGE_G : N_BSTAGE;
  constexpr int AUG = GEN ? AUG_OFF_G : AUG_OFF;
  const TcPair* pairs_tab = IT.n ? IT.pairs : P.pairs;
  const int32_t* prefix_tab = IT.n ? IT.prefix : P.tile_prefix;
  const int mode = MODES ? (P.mode == 8 ? 0 : P.mode) : 0;
  const bool trace = MODES && P.mode == 8;
  if (threadIdx.x == 0) TC_STAMP(0);
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the SWIZZLE_128B atoms (same offset in both CTAs of the pair)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + NA * STG;
  const uint32_t bar_base = b_base + NB * STG;
  // barrier slots (8 B each)
  const uint32_t a_full = bar_base, a_empty = bar_base + 16;          // 2 + 2
  const uint32_t b_full = bar_base + 32, b_empty = bar_base + 64;     // 4 + 4
  const uint32_t t_full = bar_base + 96, t_empty = bar_base + 112;    // 2 + 2
  const uint32_t tmem_slot = bar_base + 128;
  const uint32_t job_full = bar_base + 192, job_empty = bar_base + 256;   // JOB_RING + JOB_RING (TAIL)
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (NA + NB) * STG + 128);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader (issues the MMAs)
  const int n_pairs_cta = gridDim.x >> 1;        // CTA pairs in the grid
  const int pair_id = blockIdx.x >> 1;

  // One parallel sweep over the pair flags: a launch with no frame pair of this kernel's kind
  // (all integer-valued, or all general-float) costs a single memory round trip.
  if (!P.kinds_known) {
    const int qg = P.q_flags[0] != 0;
    int mine = 0;
    for (int p = threadIdx.x; p < P.n_pairs; p += TC_THREADS)
      if (((qg | (pairs_tab[p].t_flags[0] != 0)) != 0) == GEN &&
          prefix_tab[p + 1] != prefix_tab[p])
        mine = 1;
    if (__syncthreads_or(mine) == 0) return;  // both CTAs of the pair reach the same verdict
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < NA; s++) {
      mbar_init(a_full + 8 * s, 1);    // leader producer's arrive.expect_tx (bytes of both CTAs)
      mbar_init(a_empty + 8 * s, 1);   // tcgen05.commit multicast
    }
    for (int s = 0; s < NB; s++) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(t_full + 8 * s, 1);                  // tcgen05.commit multicast
      mbar_init(t_empty + 8 * s, 2 * N_EPI_WARPS);   // epilogue warps of both CTAs (leader's copy)
    }
    if (TAIL) {
      for (int s = 0; s < JOB_RING; s++) {
        mbar_init(job_full + 8 * s, 4);    // the four epilogue warps that store a segment's records
        mbar_init(job_empty + 8 * s, 1);   // the tail warps have read the entry
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) TC_STAMP(1);
  PDL_TRIGGER();   // the merge kernel behind this one may be scheduled now; it waits for our exit

  if (warp == 0) {
    // ================= TMA producer (both CTAs; each loads its own halves) =================
    if (elect_one()) {
      TileIter it;
      it.init(P, pairs_tab, prefix_tab, pair_id, n_pairs_cta, GEN);
      int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0;
      int maps_of_pair = -1;
      bool new_seg = true;
      while (it.valid()) {
        if (maps_of_pair != it.pair) {
          if (!IT.n) it.acquire_maps();   // (maps in the parameters need no proxy fence)
          maps_of_pair = it.pair;
        }
        if (new_seg) {
          mbar_wait(a_empty + 8 * a_stage, a_phase ^ 1);
          const uint32_t dst = a_base + a_stage * STG;
          const uint32_t bar = a_full + 8 * a_stage;
          const int row0 = it.rb * 2 * BM + (int)rank * BM;
          if (rank == 0) mbar_expect_tx(bar, 2 * STG);
          tma_load_2d(dst, P.q_tmap, bar, 0, row0);
          tma_load_2d(dst + KBLK, P.q_tmap, bar, 64, row0);
          if (GEN) {
            tma_load_2d(dst + 2 * KBLK, P.q_tmap + 2, bar, 0, row0);
            tma_load_2d(dst + 3 * KBLK, P.q_tmap + 2, bar, 64, row0);
          }
          tma_load_2d(dst + AUG, P.q_tmap + 1, bar, 0, row0 >> 3);
          if (++a_stage == NA) { a_stage = 0; a_phase ^= 1; }
        }
        mbar_wait(b_empty + 8 * b_stage, b_phase ^ 1);
        {
          const uint32_t dst = b_base + b_stage * STG;
          const uint32_t bar = b_full + 8 * b_stage;
          // the MMA takes the first N/2 rows of each CTA's stage as this CTA's half of the N
          // columns: on a partial last tile (N < 256) the second CTA starts N/2 rows in, not 128
          const int row0 = it.cb * BN + (int)rank * (it.n_eff() >> 1);
          if (rank == 0) mbar_expect_tx(bar, 2 * STG);
          tma_load_2d(dst, it.tmap, bar, 0, row0);
          tma_load_2d(dst + KBLK, it.tmap, bar, 64, row0);
          if (GEN) {
            tma_load_2d(dst + 2 * KBLK, it.tmap + 2, bar, 0, row0);
            tma_load_2d(dst + 3 * KBLK, it.tmap + 2, bar, 64, row0);
          }
          tma_load_2d(dst + AUG, it.tmap + 1, bar, 0, row0 >> 3);
          if (++b_stage == NB) { b_stage = 0; b_phase ^= 1; }
        }
        new_seg = it.next(P);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0 && elect_one()) {
      TileIter it;
      it.init(P, pairs_tab, prefix_tab, pair_id, n_pairs_cta, GEN);
      int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0, t_stage = 0, t_phase = 0;
      int cur_a = 0;
      int n_issued = 0;
      bool new_seg = true;
      while (it.valid()) {
        if (new_seg) {
          mbar_wait(a_full + 8 * a_stage, a_phase);
          cur_a = a_stage;
          if (++a_stage == NA) { a_stage = 0; a_phase ^= 1; }
        }
        mbar_wait(b_full + 8 * b_stage, b_phase);
        mbar_wait(t_empty + 8 * t_stage, t_phase ^ 1);
        tc_fence_after();
        if (MODES && n_issued++ == 0) TC_STAMP(2);
        const uint32_t a_addr = a_base + cur_a * STG;
        const uint32_t b_addr = b_base + b_stage * STG;
        const uint32_t d_tmem = tmem_base + t_stage * BN;
        // N of this tile's instructions: 256, or the valid columns of a partial last tile in whole
        // chunks (the instruction descriptor's N field is bits 17-22 = N >> 3)
        const uint32_t n_bits = (uint32_t)(it.n_eff() >> 3) << 17;
        if (!GEN && P.fp8) {
          // ORB bits as e4m3 0/1 bytes: a stage row is 256 bytes = 256 elements, 32 per MMA; the
          // descriptor arithmetic is byte for byte that of the bf16 operand
          const uint32_t f8_neg = idesc_e4m3(2 * BM, 0, 1) | n_bits, f8_pos = idesc_e4m3(2 * BM, 0, 0) | n_bits;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint64_t ad = desc_sw128(a_addr + (k >> 2) * KBLK + (k & 3) * 32);
            const uint64_t bd = desc_sw128(b_addr + (k >> 2) * KBLK + (k & 3) * 32);
            tc_mma_f8(d_tmem, ad, bd, f8_neg, k > 0 ? 1u : 0u);
          }
          tc_mma_f8(d_tmem, desc_interleave(a_addr + AUG), desc_interleave(b_addr + AUG), f8_pos, 1u);
        } else if (mode < 3) {
          // exact mode: -(q.t); general floats: -(qh.th + qh.tl + ql.th), hi at k-blocks 0-1 and lo
          // at k-blocks 2-3 of the stage
          const uint32_t idesc_neg = idesc_bf16(2 * BM, 0, 1) | n_bits, idesc_pos = idesc_bf16(2 * BM, 0, 0) | n_bits;
          constexpr int N_TERMS = GEN ? 3 : 1;
#pragma unroll
          for (int term = 0; term < N_TERMS; term++) {
            const uint32_t ao = a_addr + (term == 2 ? 2 * KBLK : 0);
            const uint32_t bo = b_addr + (term == 1 ? 2 * KBLK : 0);
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const uint64_t ad = desc_sw128(ao + (k >> 2) * KBLK + (k & 3) * 32);
              const uint64_t bd = desc_sw128(bo + (k >> 2) * KBLK + (k & 3) * 32);
              tc_mma(d_tmem, ad, bd, idesc_neg, (term | k) > 0 ? 1u : 0u);
            }
          }
          tc_mma(d_tmem, desc_interleave(a_addr + AUG), desc_interleave(b_addr + AUG), idesc_pos, 1u);
        } else if (MODES && mode >= 4) {
          // timing probe: the same tile with kind::i8 MMAs (K = 32 bytes each) on whatever bytes the
          // stage holds; mode 4 issues the 5 a byte-wide kernel would need, mode 5 issues 9
          const int n_i8 = mode == 4 ? 5 : 9;
          for (int k = 0; k < n_i8; k++) {
            const uint64_t ad = desc_sw128(a_addr + ((k >> 2) & 1) * KBLK + (k & 3) * 32);
            const uint64_t bd = desc_sw128(b_addr + ((k >> 2) & 1) * KBLK + (k & 3) * 32);
            tc_mma_i8(d_tmem, ad, bd, idesc_u8(2 * BM, BN), k > 0 ? 1u : 0u);
          }
        }
        tc_commit(b_empty + 8 * b_stage);
        tc_commit(t_full + 8 * t_stage);
        if (++b_stage == NB) { b_stage = 0; b_phase ^= 1; }
        if (++t_stage == 2) { t_stage = 0; t_phase ^= 1; }
        new_seg = it.next(P);
        if (new_seg) tc_commit(a_empty + 8 * cur_a);  // the segment's MMAs are done with A
      }
      if (MODES) { TC_STAMP(3); if (trace) g_tc_trace[(blockIdx.x % 160) * 8 + 7] = (unsigned long long)n_issued; }
    }
  } else if (warp >= EPI_WARP0) {
    // ================= epilogue (both CTAs; each drains its own TMEM) =================
    const int ew = warp - EPI_WARP0;
    const int quarter = ew & 3;   // TMEM lane quarter this warp may read
    const int half = ew >> 2;     // which column range of the tile (COLS_PER_WARP wide)
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int row_in_tile = (int)rank * BM + quarter * 32 + lane;  // within the pair's 256 rows
    const bool do_ld = !MODES || mode < 2 || mode >= 6;
    const bool do_proc = !MODES || mode < 1 || mode == 6;
    TileIter it;
    it.init(P, pairs_tab, prefix_tab, pair_id, n_pairs_cta, GEN);
    int t_stage = 0, t_phase = 0;
    EpiState st;
    st.reset();
    int seg_pair = -1, seg_rb = 0;
    bool new_seg = true;
    bool first_tile = true;
    int seg_ntiles = 1, seg_ncb = 1, seg_c = 0;
    int seg_i = 0;   // TAIL: segments this CTA has flushed
    volatile int2* jobs = reinterpret_cast<volatile int2*>(smem_gen + (NA + NB) * STG + 320);
    while (it.valid()) {
      if (new_seg) {
        seg_pair = it.pair; seg_rb = it.rb;
        seg_ntiles = it.n_tiles; seg_ncb = it.n_cb; seg_c = it.slot_c;
        st.reset();
      }
      // whole chunks of this warp's column range that hold columns of the tile (4 except on a
      // partial last tile, where the MMA wrote only the first n_eff columns of the stage)
      int n_chunks = (it.n_eff() - half * COLS_PER_WARP) >> 5;
      n_chunks = n_chunks < 0 ? 0 : (n_chunks > CHUNKS ? CHUNKS : n_chunks);
      mbar_wait(t_full + 8 * t_stage, t_phase);
      tc_fence_after();
      if (MODES && first_tile && ew == 0 && lane == 0) TC_STAMP(4);
      const uint32_t t_addr = tmem_base + lane_addr + t_stage * BN + half * COLS_PER_WARP;
      const int gid_tile = (it.cb * BN + half * COLS_PER_WARP) / GROUP;
      const bool dump = DBG && first_tile && pair_id == 0;
      uint32_t va[32], vb[32];
      if (MODES) {   // (modes without TMEM reads process whatever the registers hold)
#pragma unroll
        for (int j = 0; j < 32; j++) { va[j] = 0; vb[j] = 0; }
      }
      if (n_chunks == CHUNKS) {
        if (do_ld) TMEM_LD32(t_addr, va);
#pragma unroll
        for (int c = 0; c < CHUNKS; c += 2) {
          if (do_ld) {
            TMEM_WAIT32(va);
            TMEM_LD32(t_addr + 32 * (c + 1), vb);
          }
          if (dump) {
#pragma unroll
            for (int j = 0; j < 32; j++)
              P.dbg[(size_t)row_in_tile * BN + half * COLS_PER_WARP + 32 * c + j] = __uint_as_float(va[j]);
          }
          if (do_proc) process_chunk<GEN>(va, gid_tile + 4 * c, st);
          if (do_ld) {
            TMEM_WAIT32(vb);
            if (c + 2 < CHUNKS) TMEM_LD32(t_addr + 32 * (c + 2), va);
          }
          if (c + 2 >= CHUNKS) {
            // all TMEM reads of this accumulator stage are complete: hand it back to the MMA warp
            // (the leader's barrier counts the epilogue warps of both CTAs)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(t_empty + 8 * t_stage, 0);
          }
          if (dump) {
#pragma unroll
            for (int j = 0; j < 32; j++)
              P.dbg[(size_t)row_in_tile * BN + half * COLS_PER_WARP + 32 * (c + 1) + j] = __uint_as_float(vb[j]);
          }
          if (do_proc) process_chunk<GEN>(vb, gid_tile + 4 * (c + 1), st);
        }
      } else {
        // partial last tile of a train set whose row count is not a multiple of 256: only the
        // chunks the MMA wrote (the rest of the stage holds another tile's accumulators)
#pragma unroll 1
        for (int c = 0; c < n_chunks; c++) {
          if (do_ld) {
            TMEM_LD32(t_addr + 32 * c, va);
            TMEM_WAIT32(va);
          }
          if (dump) {
#pragma unroll
            for (int j = 0; j < 32; j++)
              P.dbg[(size_t)row_in_tile * BN + half * COLS_PER_WARP + 32 * c + j] = __uint_as_float(va[j]);
          }
          if (do_proc) process_chunk<GEN>(va, gid_tile + 4 * c, st);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(t_empty + 8 * t_stage, 0);
      }
      first_tile = false;
      if (++t_stage == 2) { t_stage = 0; t_phase ^= 1; }
      new_seg = it.next(P);
      if (new_seg) {
        // flush this segment's per-row record
        // ordinal of this share among the shares that cut the row block (<= n_slots / 2)
        int ord = seg_c - owner_cta(seg_ntiles, n_pairs_cta, seg_rb * seg_ncb, P.wide != 0);
        if (ord < 0 || COL_SPLITS * ord + COL_SPLITS - 1 >= P.n_slots) {
          if (lane == 0) atomicOr(P.err_flag, 2);
          ord = 0;
        }
        const int slot = GEN ? COL_SPLITS * ord + half : ord;   // exact mode: the halves are merged below
        const size_t at = ((size_t)seg_pair * P.n_slots + slot) * P.nq_pad + seg_rb * 2 * BM + row_in_tile;
        if (GEN) {
          // two 16-byte records: the four chunk minima; their group indices and the two bounds
          P.cand[2 * at] = make_uint4(__float_as_uint(st.m[0]), __float_as_uint(st.m[1]),
                                      __float_as_uint(st.m[2]), __float_as_uint(st.m[3]));
          P.cand[2 * at + 1] = make_uint4(((uint32_t)st.i[0] & 0xFFFFu) | ((uint32_t)st.i[1] << 16),
                                          ((uint32_t)st.i[2] & 0xFFFFu) | ((uint32_t)st.i[3] << 16),
                                          __float_as_uint(st.s), __float_as_uint(st.s2));
        } else {
          // The two warps that share these rows (column halves of every tile) combine their
          // records through shared memory: ONE record per (share, row) reaches HBM instead of two
          // (half the slot-record traffic of the kernel and of the tail behind it).
          // {best chunk min, second chunk min, second group min inside the best chunk,
          //  gid of the best | gid of the second << 16}; 0xFFFF = absent (gids fit: T <= 524k rows)
          uint4* xch = reinterpret_cast<uint4*>(smem_gen + (NA + NB) * STG + 1024) + quarter * 32 + lane;
          if (half == 1)
            *xch = make_uint4(__float_as_uint(st.m[0]), __float_as_uint(st.m[1]), __float_as_uint(st.s),
                              ((uint32_t)st.i[0] & 0xFFFFu) | ((uint32_t)st.i[1] << 16));
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
          if (half == 0) {
            const uint4 o = *xch;
            float am0 = st.m[0], am1 = st.m[1], as = st.s;
            int ag0 = st.i[0] & 0xFFFF, ag1 = st.i[1] & 0xFFFF;
            const float bm0 = __uint_as_float(o.x), bm1 = __uint_as_float(o.y), bs = __uint_as_float(o.z);
            const int bg0 = (int)(o.w & 0xFFFFu), bg1 = (int)(o.w >> 16);
            // (value, group) lexicographic order; an absent entry is (+inf, 0xFFFF) and loses every tie
            const bool b_first = bm0 < am0 || (bm0 == am0 && bg0 < ag0);
            float m0, m1, sv;
            int g0, g1;
            if (b_first) {
              m0 = bm0; g0 = bg0; sv = bs;
              const bool own = bm1 < am0 || (bm1 == am0 && bg1 < ag0);   // winner's second vs loser's best
              m1 = own ? bm1 : am0; g1 = own ? bg1 : ag0;
            } else {
              m0 = am0; g0 = ag0; sv = as;
              const bool own = am1 < bm0 || (am1 == bm0 && ag1 < bg0);
              m1 = own ? am1 : bm0; g1 = own ? ag1 : bg0;
            }
            P.cand[at] = make_uint4(__float_as_uint(m0), __float_as_uint(m1), __float_as_uint(sv),
                                    (uint32_t)g0 | ((uint32_t)g1 << 16));
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");   // the exchange slot may be rewritten
          if (TAIL && half == 0) {
            // hand the segment to the tail warps: the entry, then one arrival per storing warp (the
            // arrival releases this warp's record stores to whoever waits on the barrier)
            const int slot_i = seg_i % JOB_RING;
            if (ew == 0 && lane == 0) {
              mbar_wait(job_empty + 8 * slot_i, ((seg_i / JOB_RING) & 1) ^ 1);
              jobs[slot_i].x = seg_pair;
              jobs[slot_i].y = seg_rb;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(job_full + 8 * slot_i);
            seg_i++;
          }
        }
      }
    }
    if (TAIL && half == 0) {   // end of this CTA's walk
      const int slot_i = seg_i % JOB_RING;
      if (ew == 0 && lane == 0) {
        mbar_wait(job_empty + 8 * slot_i, ((seg_i / JOB_RING) & 1) ^ 1);
        jobs[slot_i].x = -1;
        jobs[slot_i].y = 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(job_full + 8 * slot_i);
    }
    if (MODES && ew == 0 && lane == 0) TC_STAMP(5);
  } else if (TAIL && (warp == 2 || warp == 3)) {
    // ================= tail (both CTAs; each finishes units of its own 128 rows) =================
    const int tid = (warp - 2) * 32 + lane;
    TailSmem<128>& sm = *reinterpret_cast<TailSmem<128>*>(smem_gen + (NA + NB) * STG + SMEM_BARS);
    volatile int2* jobs = reinterpret_cast<volatile int2*>(smem_gen + (NA + NB) * STG + 320);
    for (int s = 0;; s++) {
      const int slot_i = s % JOB_RING;
      mbar_wait(job_full + 8 * slot_i, (s / JOB_RING) & 1);
      const int pair = jobs[slot_i].x, rb = jobs[slot_i].y;
      group_sync<BAR_TAIL, 64>();                       // both warps hold the entry
      if (tid == 0) mbar_arrive(job_empty + 8 * slot_i);
      if (pair < 0) break;
      if (tid == 0) {
        // one more share has flushed its records of this row block; the last one owns the unit
        const int n_tiles = prefix_tab[pair + 1] - prefix_tab[pair];
        const int n_cb = n_tiles / P.n_rb;
        const int first = owner_cta(n_tiles, n_pairs_cta, rb * n_cb, P.wide != 0);
        const int last = owner_cta(n_tiles, n_pairs_cta, (rb + 1) * n_cb - 1, P.wide != 0);
        int32_t* cnt = TT.seg_done + ((size_t)pair * P.n_rb + rb) * 2 + rank;
        __threadfence();                                // (cumulative: covers the epilogue's stores)
        const int old = atomicAdd(cnt, 1);
        const int mine = old == last - first;
        if (mine) *cnt = 0;                             // nobody else touches it in this launch
        __threadfence();
        sm.flag = mine;
      }
      group_sync<BAR_TAIL, 64>();
      if (sm.flag) {
        const int q0 = rb * 2 * BM + (int)rank * BM;
        if (P.fp8) tail_unit<true, 64, 128, BAR_TAIL, 1>(TT.R, TT.C, pairs_tab, prefix_tab, pair, rb, q0, rb * 2 + (int)rank, P.n_rb * 2, false, sm, tid);
        else tail_unit<false, 64, 128, BAR_TAIL, 1>(TT.R, TT.C, pairs_tab, prefix_tab, pair, rb, q0, rb * 2 + (int)rank, P.n_rb * 2, false, sm, tid);
      }
    }
  }

  // both CTAs must be done with each other's shared memory / barriers before either exits;
  // reconverge each warp first (the elected producer / MMA lanes come back from their loops):
  // barrier.cluster.*.aligned needs the whole warp
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 0) TC_STAMP(6);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                 : "memory");
  }
}

// ---- rerank kernels (RerankParams and the shared tail are defined above the tcgen05 kernel) ---
// Pass 1, one thread per (pair, row): merges the row's slot records into its best two chunks.
// Lanes run along the rows, so every slot read is a coalesced 512 B per warp.
//
// Ratio-test pruning (exact, not a heuristic): v0 is the row's exact best d^2/2, and the second
// best element is no farther than L = min(second group minimum inside the best chunk, second
// chunk's minimum) -- both exact distances of real elements.  sqrtf, float->double and the
// multiplication by a non-negative ratio are monotone, so if even the upper bound
// D1 = sqrtf(2*L) fails  d0 < ratio * D1  (getGoodMatches' own comparison,
// featureMatchingCommon.cpp:47), the true second distance fails it too: the row is rejected
// without touching its candidates, and its record carries (d0, D1) so that the finalize kernel
// reaches the same verdict with the same arithmetic.  Rows that survive go to the work list.
// General-float pairs have no slot records (the tcgen05 kernel skipped them) and are left alone.
__global__ void __launch_bounds__(256) sift_merge_kernel(const RerankParams R) {
  __shared__ int n_valid_s;
  const int pair = blockIdx.y;
  const int q = blockIdx.x * 256 + threadIdx.x;   // one block = one 256-row query block
  PDL_TRIGGER();
  PDL_WAIT();
  // general-float pair: its records belong to the exact fp32 kernel (block-uniform exit)
  if (R.q_flags[0] != 0 || R.pairs[pair].t_flags[0] != 0) return;
  if (threadIdx.x == 0) {
    // Which slots did the tcgen05 kernel write for this (pair, query block)?  Slot 2*ord + half,
    // ord = position of the writing share among the shares that cut the block's tile range --
    // the same arithmetic as the kernel's flush, so no slot needs to be pre-cleared.
    const int n_tiles = R.tile_prefix[pair + 1] - R.tile_prefix[pair];
    int nv = 0;
    if (n_tiles > 0) {
      const int n_rb = R.nq_pad / 256;
      const int n_cb = n_tiles / n_rb;
      const int first = owner_cta(n_tiles, R.n_cta, blockIdx.x * n_cb, R.wide != 0);
      const int last = owner_cta(n_tiles, R.n_cta, (blockIdx.x + 1) * n_cb - 1, R.wide != 0);
      nv = last - first + 1;               // one merged record per share (the kernel's flush)
      if (nv > R.n_slots) nv = R.n_slots;  // cannot happen (host sizing); the self check would trip
    }
    n_valid_s = nv;
  }
  __syncthreads();
  const int n_valid = n_valid_s;
  bool survive = false;
  const float INF = __int_as_float(0x7f800000);
  float v0 = INF, v1 = INF, s0 = INF;   // best chunk min, second chunk min, 2nd group min in best
  int g0 = 0xFFFF, g1 = 0xFFFF;
  if (q < R.nq) {
    for (int s = 0; s < n_valid; s++) {
      const uint4 rec = R.cand[((size_t)pair * R.n_slots + s) * R.nq_pad + q];
      const float a = __uint_as_float(rec.x), b = __uint_as_float(rec.y);
      float sa = __uint_as_float(rec.z);
      int ia = (int)(rec.w & 0xFFFFu), ib = (int)(rec.w >> 16);
      if (R.orb) {
        // e4m3 has no infinity: padding rows carry 448 in their augmentation, real values are
        // Hamming / 2 <= 128
        if (a > ORB_PAD_THRESHOLD) ia = 0xFFFF;
        if (b > ORB_PAD_THRESHOLD) ib = 0xFFFF;
        if (sa > ORB_PAD_THRESHOLD) sa = INF;
      }
      if (ia != 0xFFFF) {
        if (lt_fi(a, ia, v0, g0)) { v1 = v0; g1 = g0; v0 = a; g0 = ia; s0 = sa; }
        else if (lt_fi(a, ia, v1, g1)) { v1 = a; g1 = ia; }
      }
      if (ib != 0xFFFF) {
        // a slot's second entry can never be the global best chunk (its own first entry beats it)
        if (lt_fi(b, ib, v1, g1)) { v1 = b; g1 = ib; }
      }
    }
    const bool has0 = g0 != 0xFFFF;
    const uint4 none = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
    for (int sp = 1; sp < R.n_split; sp++) R.part[((size_t)pair * R.n_split + sp) * R.nq + q] = none;
    if (!has0) R.part[((size_t)pair * R.n_split) * R.nq + q] = none;  // empty train set
    survive = has0;
    // L = the smallest d^2/2 outside the best group: the second group of the best chunk or the
    // second-best chunk.  +inf comes from padding columns, i.e. "no such element".
    const float L = fminf(s0, v1);
    if (R.prune && has0 && L < INF) {
      // ORB: the distance is the Hamming count itself (int -> float), its record key the count
      const float d0 = R.orb ? 2.0f * v0 : sqrtf(2.0f * v0), D1 = R.orb ? 2.0f * L : sqrtf(2.0f * L);
      if (!((double)d0 < __dmul_rn(R.ratio, (double)D1))) {
        R.part[((size_t)pair * R.n_split) * R.nq + q] =
            R.orb ? make_uint4((uint32_t)d0, 0u, (uint32_t)D1, 0u)
                  : make_uint4(__float_as_uint(d0), 0u, __float_as_uint(D1), 0u);
        survive = false;
      }
    }
    v1 = L;  // what the survivors need downstream
  }
  // warp-aggregated append to the work list
  const unsigned bal = __ballot_sync(0xffffffffu, survive);
  if (bal) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(R.work_n, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (survive) {
      const int at = base + __popc(bal & ((1u << lane) - 1));
      R.work[at] = make_uint4((uint32_t)pair, (uint32_t)q, (uint32_t)g0, (uint32_t)g1);
      R.work_v0[at] = make_float2(v0, v1);
    }
  }
}

// Pass 2 (match output), persistent warps over the work list, one warp per surviving row.
// Only the best group's 8 columns are evaluated: that yields the best match (index and exact
// d^2) and the second smallest of the group; every element outside the group is bounded below by
// L (an exact distance of a real element, from the merge pass), so the second distance is
// exactly min(second of the group, L) -- which is all getGoodMatches needs
// (featureMatchingCommon.cpp:47 compares distances only; DMatch carries the best index).
// 4 lanes x 32 B per candidate row, dp4a, two shuffles to reduce, three to select.
__global__ void __launch_bounds__(256) sift_rerank_lite_kernel(const RerankParams R) {
  const int lane = threadIdx.x & 31;
  const int part = lane & 3, cand = lane >> 2;
  PDL_TRIGGER();
  PDL_WAIT();
  const int n_work = *R.work_n;
  const int warps = gridDim.x * 8;
  for (int w = blockIdx.x * 8 + (threadIdx.x >> 5); w < n_work; w += warps) {
    const uint4 wk = R.work[w];
    const int pair = (int)wk.x, q = (int)wk.y, g0 = (int)wk.z;
    const float2 vl = R.work_v0[w];
    const TcPair* pr = R.pairs + pair;
    const int col = g0 * GROUP + cand;
    const bool ok = col < pr->t_n;
    const int cc = ok ? col : 0;
    const uint4* tp = reinterpret_cast<const uint4*>(pr->t_u8 + (size_t)cc * 128 + part * 32);
    const uint4* qp = reinterpret_cast<const uint4*>(R.q_u8 + (size_t)q * 128 + part * 32);
    const uint4 t0 = tp[0], t1 = tp[1], q0 = qp[0], q1 = qp[1];
    const uint32_t nt2 = (uint32_t)pr->t_nrm2[cc], nq2 = (uint32_t)R.q_nrm2[q];
    uint32_t dot = 0;
    dot = __dp4a(q0.x, t0.x, dot); dot = __dp4a(q0.y, t0.y, dot);
    dot = __dp4a(q0.z, t0.z, dot); dot = __dp4a(q0.w, t0.w, dot);
    dot = __dp4a(q1.x, t1.x, dot); dot = __dp4a(q1.y, t1.y, dot);
    dot = __dp4a(q1.z, t1.z, dot); dot = __dp4a(q1.w, t1.w, dot);
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    unsigned long long k0 = ok ? (((unsigned long long)(nq2 + nt2 - 2u * dot) << 32) | (uint32_t)col) : ~0ull;
    unsigned long long k1 = ~0ull;
#pragma unroll
    for (int off = 4; off <= 16; off <<= 1) {
      const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, off);
      const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, off);
      const unsigned long long lo = k0 < o0 ? k0 : o0, hi = k0 < o0 ? o0 : k0;
      const unsigned long long s2 = k1 < o1 ? k1 : o1;
      k0 = lo;
      k1 = hi < s2 ? hi : s2;
    }
    if (lane == 0) {
      uint4 rec = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
      if (k0 != ~0ull) {
        const uint32_t d2 = (uint32_t)(k0 >> 32);
        // self check: the group's exact minimum must equal twice the tensor-core value
        if ((float)d2 != 2.0f * vl.x) atomicOr(R.err_flag, 1);
        rec.x = __float_as_uint(sqrtf((float)d2));
        rec.y = (uint32_t)(k0 & 0xFFFFFFFFu);
        // second distance: inside the group, or the bound from outside it (exact values both)
        float d1sq = vl.y < __int_as_float(0x7f800000) ? 2.0f * vl.y : -1.0f;
        if (k1 != ~0ull) {
          const float x2 = (float)(uint32_t)(k1 >> 32);
          d1sq = (d1sq < 0.0f || x2 < d1sq) ? x2 : d1sq;
        }
        if (d1sq >= 0.0f) {
          rec.z = __float_as_uint(sqrtf(d1sq));
          // the second neighbour's index is not part of the match output; the placeholder must
          // sort behind every real index, or an equal second distance would overtake the best
          // entry in the finalize merge (visible with ratios above 1)
          rec.w = 0xFFFFFFFEu;
        }
      }
      R.part[((size_t)pair * R.n_split) * R.nq + q] = rec;
    }
  }
}

// Pass 2 (raw k-NN output: both indices are needed), one warp per row.  Candidates = all 32
// columns of the best chunk + the 8 columns of the second chunk's group (40): the true top-2
// columns are always among them.  Each candidate row is read by 8 lanes x 16 B (one 128 B line
// per candidate), dp4a'd against the matching 16 B of the query row and reduced over the 8 lanes.
__global__ void __launch_bounds__(256) sift_rerank_kernel(const RerankParams R) {
  const int lane = threadIdx.x & 31;
  const int part = lane & 7, sub = lane >> 3;
  PDL_TRIGGER();
  PDL_WAIT();
  const int n_work = *R.work_n;
  const int warps = gridDim.x * 8;
  for (int w = blockIdx.x * 8 + (threadIdx.x >> 5); w < n_work; w += warps) {
    const uint4 wk = R.work[w];
    const int pair = (int)wk.x, q = (int)wk.y, g0 = (int)wk.z, g1 = (int)wk.w;
    const float v0 = R.work_v0[w].x;
    const TcPair* pr = R.pairs + pair;
    const uint8_t* t_u8 = pr->t_u8;
    const int32_t* t_nrm2 = pr->t_nrm2;
    const int t_n = pr->t_n;
    const bool has1 = g1 != 0xFFFF;
    const int col_a = (g0 >> 2) * 32;          // first column of the best chunk
    const int col_b = has1 ? g1 * GROUP : 0;   // first column of the second chunk's group

    const uint4 qv = *reinterpret_cast<const uint4*>(R.q_u8 + (size_t)q * 128 + part * 16);
    const uint32_t nq2 = (uint32_t)R.q_nrm2[q];
    unsigned long long k0 = ~0ull, k1 = ~0ull;
    uint32_t mn_a = 0xFFFFFFFFu;  // minimum d^2 inside the best chunk (self check)
#pragma unroll 5
    for (int it = 0; it < 10; it++) {
      const int k = it * 4 + sub;
      const int col = k < 32 ? col_a + k : col_b + (k - 32);
      const bool ok = (k < 32 || has1) && col < t_n;
      const int cc = ok ? col : 0;
      const uint4 tv = *reinterpret_cast<const uint4*>(t_u8 + (size_t)cc * 128 + part * 16);
      const uint32_t nt2 = (uint32_t)t_nrm2[cc];
      uint32_t dot = 0;
      dot = __dp4a(qv.x, tv.x, dot);
      dot = __dp4a(qv.y, tv.y, dot);
      dot = __dp4a(qv.z, tv.z, dot);
      dot = __dp4a(qv.w, tv.w, dot);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      if (ok) {
        const uint32_t d2 = nq2 + nt2 - 2u * dot;
        const unsigned long long key = ((unsigned long long)d2 << 32) | (uint32_t)col;
        if (k < 32) mn_a = min(mn_a, d2);
        if (key < k1) {
          if (key < k0) { k1 = k0; k0 = key; } else { k1 = key; }
        }
      }
    }
    // merge the four candidate lane groups (the 8 lanes of a group hold identical values)
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
      const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, off);
      const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, off);
      const unsigned long long lo = k0 < o0 ? k0 : o0, hi = k0 < o0 ? o0 : k0;
      const unsigned long long s2 = k1 < o1 ? k1 : o1;
      k0 = lo;
      k1 = hi < s2 ? hi : s2;
      mn_a = min(mn_a, __shfl_xor_sync(0xffffffffu, mn_a, off));
    }
    if (lane == 0) {
      // self check: the best chunk's exact minimum must equal twice the tensor-core value
      if (mn_a != 0xFFFFFFFFu && (float)mn_a != 2.0f * v0) atomicOr(R.err_flag, 1);
      uint4 rec = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
      if (k0 != ~0ull) {
        rec.x = __float_as_uint(sqrtf((float)(uint32_t)(k0 >> 32)));
        rec.y = (uint32_t)(k0 & 0xFFFFFFFFu);
      }
      if (k1 != ~0ull) {
        rec.z = __float_as_uint(sqrtf((float)(uint32_t)(k1 >> 32)));
        rec.w = (uint32_t)(k1 & 0xFFFFFFFFu);
      }
      R.part[((size_t)pair * R.n_split) * R.nq + q] = rec;
    }
  }
}

// ---- ORB through the tensor cores: Hamming(q, t) = |q - t|^2 over the 256 bits as 0/1 values ----
// The tcgen05 kernel runs unchanged on e4m3 0/1 bytes (kind::f8f6f4, fp32 accumulators hold
// Hamming / 2 exactly), so the candidate logic, its tie rules and the ratio-test pruning are the
// exact-mode SIFT ones; only the exact evaluation of the candidates differs: XOR + POPC on the
// original 32-byte rows, as cv::BFMatcher(NORM_HAMMING) does
// (src/mainModule/featureMatching/featureMatchingCPU.cpp:33-35, :40).
// Match output: the best group's 8 columns, 4 lanes x 8 bytes per candidate row.
__global__ void __launch_bounds__(256) orb_rerank_lite_kernel(const RerankParams R) {
  const int lane = threadIdx.x & 31;
  const int part = lane & 3, cand = lane >> 2;
  PDL_TRIGGER();
  PDL_WAIT();
  const int n_work = *R.work_n;
  const int warps = gridDim.x * 8;
  const float INF = __int_as_float(0x7f800000);
  for (int w = blockIdx.x * 8 + (threadIdx.x >> 5); w < n_work; w += warps) {
    const uint4 wk = R.work[w];
    const int pair = (int)wk.x, q = (int)wk.y, g0 = (int)wk.z;
    const float2 vl = R.work_v0[w];
    const TcPair* pr = R.pairs + pair;
    const int col = g0 * GROUP + cand;
    const bool ok = col < pr->t_n;
    const int cc = ok ? col : 0;
    const unsigned long long tv = *reinterpret_cast<const unsigned long long*>(pr->t_u8 + (size_t)cc * 32 + part * 8);
    const unsigned long long qv = *reinterpret_cast<const unsigned long long*>(R.q_u8 + (size_t)q * 32 + part * 8);
    uint32_t h = (uint32_t)__popcll(qv ^ tv);
    h += __shfl_xor_sync(0xffffffffu, h, 1);
    h += __shfl_xor_sync(0xffffffffu, h, 2);
    unsigned long long k0 = ok ? (((unsigned long long)h << 32) | (uint32_t)col) : ~0ull;
    unsigned long long k1 = ~0ull;
#pragma unroll
    for (int off = 4; off <= 16; off <<= 1) {
      const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, off);
      const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, off);
      const unsigned long long lo = k0 < o0 ? k0 : o0, hi = k0 < o0 ? o0 : k0;
      const unsigned long long s2 = k1 < o1 ? k1 : o1;
      k0 = lo;
      k1 = hi < s2 ? hi : s2;
    }
    if (lane == 0) {
      uint4 rec = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
      if (k0 != ~0ull) {
        const uint32_t h0 = (uint32_t)(k0 >> 32);
        // self check: the group's exact minimum must equal twice the tensor-core value
        if ((float)h0 != 2.0f * vl.x) atomicOr(R.err_flag, 1);
        rec.x = h0;
        rec.y = (uint32_t)(k0 & 0xFFFFFFFFu);
        // second distance: inside the group, or the bound from outside it (exact values both)
        uint32_t h1 = vl.y < INF ? (uint32_t)(2.0f * vl.y) : 0xFFFFFFFFu;
        if (k1 != ~0ull) h1 = min(h1, (uint32_t)(k1 >> 32));
        if (h1 != 0xFFFFFFFFu) {
          rec.z = h1;
          // the second neighbour's index is not part of the match output; the placeholder must
          // sort behind every real index, or an equal second distance would overtake the best
          // entry in the finalize merge (visible with ratios above 1)
          rec.w = 0xFFFFFFFEu;
        }
      }
      R.part[((size_t)pair * R.n_split) * R.nq + q] = rec;
    }
  }
}

// Raw k-NN output (both indices): the best chunk's 32 columns + the 8 columns of the second
// chunk's group, 4 lanes per candidate row, 8 candidates per pass.
__global__ void __launch_bounds__(256) orb_rerank_kernel(const RerankParams R) {
  const int lane = threadIdx.x & 31;
  const int part = lane & 3, cand = lane >> 2;
  PDL_TRIGGER();
  PDL_WAIT();
  const int n_work = *R.work_n;
  const int warps = gridDim.x * 8;
  for (int w = blockIdx.x * 8 + (threadIdx.x >> 5); w < n_work; w += warps) {
    const uint4 wk = R.work[w];
    const int pair = (int)wk.x, q = (int)wk.y, g0 = (int)wk.z, g1 = (int)wk.w;
    const float v0 = R.work_v0[w].x;
    const TcPair* pr = R.pairs + pair;
    const int t_n = pr->t_n;
    const bool has1 = g1 != 0xFFFF;
    const int col_a = (g0 >> 2) * 32;          // first column of the best chunk
    const int col_b = has1 ? g1 * GROUP : 0;   // first column of the second chunk's group
    const unsigned long long qv = *reinterpret_cast<const unsigned long long*>(R.q_u8 + (size_t)q * 32 + part * 8);
    unsigned long long k0 = ~0ull, k1 = ~0ull;
    uint32_t mn_a = 0xFFFFFFFFu;  // minimum inside the best chunk (self check)
#pragma unroll
    for (int it = 0; it < 5; it++) {
      const int k = it * 8 + cand;
      const int col = k < 32 ? col_a + k : col_b + (k - 32);
      const bool ok = (k < 32 || has1) && col < t_n;
      const int cc = ok ? col : 0;
      const unsigned long long tv = *reinterpret_cast<const unsigned long long*>(pr->t_u8 + (size_t)cc * 32 + part * 8);
      uint32_t h = (uint32_t)__popcll(qv ^ tv);
      h += __shfl_xor_sync(0xffffffffu, h, 1);
      h += __shfl_xor_sync(0xffffffffu, h, 2);
      if (ok) {
        const unsigned long long key = ((unsigned long long)h << 32) | (uint32_t)col;
        if (k < 32) mn_a = min(mn_a, h);
        if (key < k1) {
          if (key < k0) { k1 = k0; k0 = key; } else { k1 = key; }
        }
      }
    }
    // merge the eight candidate lane groups (the 4 lanes of a group hold identical values)
#pragma unroll
    for (int off = 4; off <= 16; off <<= 1) {
      const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, off);
      const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, off);
      const unsigned long long lo = k0 < o0 ? k0 : o0, hi = k0 < o0 ? o0 : k0;
      const unsigned long long s2 = k1 < o1 ? k1 : o1;
      k0 = lo;
      k1 = hi < s2 ? hi : s2;
      mn_a = min(mn_a, __shfl_xor_sync(0xffffffffu, mn_a, off));
    }
    if (lane == 0) {
      if (mn_a != 0xFFFFFFFFu && (float)mn_a != 2.0f * v0) atomicOr(R.err_flag, 1);
      uint4 rec = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
      if (k0 != ~0ull) { rec.x = (uint32_t)(k0 >> 32); rec.y = (uint32_t)(k0 & 0xFFFFFFFFu); }
      if (k1 != ~0ull) { rec.z = (uint32_t)(k1 >> 32); rec.w = (uint32_t)(k1 & 0xFFFFFFFFu); }
      R.part[((size_t)pair * R.n_split) * R.nq + q] = rec;
    }
  }
}

// ================= fused tail of the match path, two-kernel form =================
// merge -> rerank (best group only) -> ratio test for one block of 256 query rows of one frame
// pair (tail_unit with COMPACT = false).  Output is what compact_kernel (finalize.cu) consumes: the
// kept flag, the best index and distance per row, and the block's kept count.  General-float
// pairs (their records come from the certified rerank) take the plain finalize branch.  No block
// waits for another here, which is why large windows use this form (see tail_unit).
template <bool ORB, int SU>
__global__ void __launch_bounds__(256)
tc_tail_fused_kernel(const RerankParams R, const CompactArgs C, const __grid_constant__ InlineTables IT) {
  __shared__ TailSmem<256> sm;
  PDL_TRIGGER();
  PDL_WAIT();
  const TcPair* pairs_tab = IT.n ? IT.pairs : R.pairs;
  const int32_t* prefix_tab = IT.n ? IT.prefix : R.tile_prefix;
  const int pair = blockIdx.y;
  const bool gen_pair = R.q_flags[0] != 0 || pairs_tab[pair].t_flags[0] != 0;
  tail_unit<ORB, 256, 256, 0, SU, false>(R, C, pairs_tab, prefix_tab, pair, blockIdx.x, blockIdx.x * 256, blockIdx.x,
                                         gridDim.x, gen_pair, sm, threadIdx.x);
}

// The match path's tail as one kernel behind the tcgen05 kernel: a block per (256 query rows,
// frame pair).  The look-back waits only on blocks of the same pair with a smaller index, which the
// hardware dispatches first.
template <bool ORB, int SU>
__global__ void __launch_bounds__(256)
tc_tail_compact_kernel(const RerankParams R, const CompactArgs C, const __grid_constant__ InlineTables IT) {
  __shared__ TailSmem<256> sm;
  PDL_TRIGGER();
  PDL_WAIT();
  const TcPair* pairs_tab = IT.n ? IT.pairs : R.pairs;
  const int32_t* prefix_tab = IT.n ? IT.prefix : R.tile_prefix;
  const int pair = blockIdx.y;
  const bool gen_pair = R.q_flags[0] != 0 || pairs_tab[pair].t_flags[0] != 0;
  tail_unit<ORB, 256, 256, 0, SU>(R, C, pairs_tab, prefix_tab, pair, blockIdx.x, blockIdx.x * 256, blockIdx.x, gridDim.x, gen_pair, sm,
                              threadIdx.x);
}

// ================= general-float pairs: certify-or-fallback rerank =================
struct GenParams {
  const float* q_f32;
  const float* q_nrmf;
  const int32_t* q_flags;
  const TcPair* pairs;
  const int32_t* tile_prefix;
  const uint4* cand;
  int nq, nq_pad, n_slots, n_pairs, n_split, n_cta;
  int wide;
  uint4* part;
  uint2* fb_list;     // rows whose answer could not be certified: {pair, row}
  int32_t* fb_count;
  unsigned long long* fb_part;   // [FB_GRID][2] partial top-2 keys of the split fallback scan
  int32_t* fb_done;              // [FB_GRID] finished segments per listed row (zero on entry)
};

// sqrtf(hal::normL2Sqr_(q, t, 128)) in OpenCV's own summation order (see sift_exact.cu): 4 x 4
// accumulators, separately rounded multiply and add, ((a0+a1)+a2)+a3 per lane, (S0+S2)+(S1+S3).
__device__ __forceinline__ float l2_cv_order(const float* __restrict__ q, const float* __restrict__ t) {
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int l = 0; l < 4; l++) acc[a][l] = 0.f;
#pragma unroll 2
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const float4 qv = *reinterpret_cast<const float4*>(q + 16 * i + 4 * a);
      const float4 tv = *reinterpret_cast<const float4*>(t + 16 * i + 4 * a);
      float d;
      d = __fsub_rn(qv.x, tv.x); acc[a][0] = __fadd_rn(acc[a][0], __fmul_rn(d, d));
      d = __fsub_rn(qv.y, tv.y); acc[a][1] = __fadd_rn(acc[a][1], __fmul_rn(d, d));
      d = __fsub_rn(qv.z, tv.z); acc[a][2] = __fadd_rn(acc[a][2], __fmul_rn(d, d));
      d = __fsub_rn(qv.w, tv.w); acc[a][3] = __fadd_rn(acc[a][3], __fmul_rn(d, d));
    }
  }
  float S[4];
#pragma unroll
  for (int l = 0; l < 4; l++)
    S[l] = __fadd_rn(__fadd_rn(__fadd_rn(acc[0][l], acc[1][l]), acc[2][l]), acc[3][l]);
  return sqrtf(__fadd_rn(__fadd_rn(S[0], S[2]), __fadd_rn(S[1], S[3])));
}

__device__ __forceinline__ void key_top2(unsigned long long key, unsigned long long& k0,
                                         unsigned long long& k1) {
  if (key < k1) {
    if (key < k0) { k1 = k0; k0 = key; } else { k1 = key; }
  }
}
__device__ __forceinline__ void warp_top2(unsigned long long& k0, unsigned long long& k1) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, off);
    const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, off);
    const unsigned long long lo = k0 < o0 ? k0 : o0, hi = k0 < o0 ? o0 : k0;
    const unsigned long long s2 = k1 < o1 ? k1 : o1;
    k0 = lo;
    k1 = hi < s2 ? hi : s2;
  }
}

// One warp per (general-float pair, query row).  The tcgen05 accumulators are approximations of
// d^2/2 (two-term bf16 split): |2*acc - d^2| <= E = 2^-13 (|q|^2 + max|t|^2), a worst-case bound
// (dropped lo.lo and residual terms: 2^-16.4; 400 truncating fp32 accumulations of terms whose
// magnitudes sum to <= 1.01 (|q|^2+|t|^2): 2^-13.3).  Stage 0 evaluates the best 8-column group
// of each of the row's best four chunks (32 columns, one per lane) exactly, in OpenCV's summation
// order; every other column is bounded below by the fifth-best chunk minimum or by the smallest
// second group minimum of a chunk.  If that bound, minus E, clears the exact second distance, the
// answer is provably the one cv::BFMatcher gives (ties included).  Otherwise (the two nearest
// rows share a chunk: ~0.3 % of rows) stage 1 evaluates the four chunks completely (128 columns)
// against the fifth-chunk bound, and what is still uncertified goes to the exact full-row
// fallback.  Nothing is ever returned uncertified.
__global__ void __launch_bounds__(256) sift_gen_rerank_kernel(const GenParams G) {
  const int pair = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  const TcPair* pr = G.pairs + pair;
  if ((G.q_flags[0] | pr->t_flags[0]) == 0) return;  // exact-mode pair: not ours (block-uniform)
  if (q >= G.nq) return;                             // warp-uniform
  const int n_tiles = G.tile_prefix[pair + 1] - G.tile_prefix[pair];
  int n_valid = 0;
  if (n_tiles > 0) {
    const int n_rb = G.nq_pad / 256, n_cb = n_tiles / n_rb, rb = q >> 8;
    const int first = owner_cta(n_tiles, G.n_cta, rb * n_cb, G.wide != 0);
    const int last = owner_cta(n_tiles, G.n_cta, (rb + 1) * n_cb - 1, G.wide != 0);
    n_valid = min(COL_SPLITS * (last - first + 1), G.n_slots);
  }
  const float INF = __int_as_float(0x7f800000);
  // order-preserving key: the approximations can be slightly negative, so flip like a radix sort
  auto fkey = [](float v, int g) -> unsigned long long {
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (uint32_t)g;
  };
  // this lane's slot record: four sorted (value, gid) entries and the bounds for the rest of its
  // slot.  rest: columns outside the kept chunks; rest2: also the kept chunks' other groups.
  unsigned long long ek[4] = {~0ull, ~0ull, ~0ull, ~0ull};
  float ev[4] = {INF, INF, INF, INF};
  float rest = INF, rest2 = INF;
  for (int s = lane; s < n_valid; s += 32) {
    const size_t at = ((size_t)pair * G.n_slots + s) * G.nq_pad + q;
    const uint4 ra = G.cand[2 * at], rb2 = G.cand[2 * at + 1];
    const float v4[4] = {__uint_as_float(ra.x), __uint_as_float(ra.y), __uint_as_float(ra.z),
                         __uint_as_float(ra.w)};
    const int g4[4] = {(int)(rb2.x & 0xFFFFu), (int)(rb2.x >> 16), (int)(rb2.y & 0xFFFFu),
                       (int)(rb2.y >> 16)};
    rest = fminf(rest, __uint_as_float(rb2.z));
    rest2 = fminf(rest2, __uint_as_float(rb2.w));
    if (s < 32) {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (g4[k] != 0xFFFF) { ek[k] = fkey(v4[k], g4[k]); ev[k] = v4[k]; }
    } else {  // more than 32 slots: the extra records only tighten nothing, they bound
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (g4[k] != 0xFFFF) rest = fminf(rest, v4[k]);
    }
  }
  // global best four chunks: four rounds of "warp minimum of the lanes' heads, owner pops".
  // Lane r (mod 4) remembers the group (column / 8) holding the minimum of the r-th chunk.
  int gsel = -1;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const unsigned long long mine = ek[0];
    unsigned long long best = mine;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
      best = o < best ? o : best;
    }
    const bool has = best != ~0ull;
    if (has && (lane & 3) == r) gsel = (int)(uint32_t)best;
    // keys are unique (distinct chunks): exactly one lane pops
    if (has && mine == best) {
      ek[0] = ek[1]; ek[1] = ek[2]; ek[2] = ek[3]; ek[3] = ~0ull;
      ev[0] = ev[1]; ev[1] = ev[2]; ev[2] = ev[3]; ev[3] = INF;
    }
  }
  // everything this lane still holds bounds the columns that are not evaluated
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (ek[k] != ~0ull) rest = fminf(rest, ev[k]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    rest = fminf(rest, __shfl_xor_sync(0xffffffffu, rest, off));
    rest2 = fminf(rest2, __shfl_xor_sync(0xffffffffu, rest2, off));
  }

  const float* qrow = G.q_f32 + (size_t)q * 128;
  const double E = ((double)G.q_nrmf[q] + (double)__int_as_float(pr->t_flags[2])) * (1.0 / 8192.0);
  auto certified_against = [=](float bound, unsigned long long second) -> bool {
    if (!(bound < INF)) return true;            // +inf (or NaN): only padding is left
    if (second == ~0ull) return false;          // fewer than two evaluated columns, more exist
    const float d1 = __uint_as_float((uint32_t)(second >> 32));
    const double D2 = (double)d1 * (double)d1 * (1.0 + 4.8e-7);       // d^2 of the 2nd best, rounded up
    const double lower = (2.0 * (double)bound - E) * (1.0 - 3.9e-6);  // oracle d^2 of any other column
    return lower > D2;
  };
  // stage 0: the best group of each of the four chunks.  A group is 8 consecutive train rows
  // (4 KB); four lanes share a row -- lane a of the quad owns OpenCV's accumulator a (elements
  // 16 i + 4 a .. + 3), so every load instruction reads whole 64-byte pieces of 8 rows -- and the
  // quad then combines the accumulators in OpenCV's order.
  unsigned long long e0 = ~0ull, e1 = ~0ull;
  {
    const int a = lane & 3, j = lane >> 2, quad0 = lane & ~3;
    float4 q4[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q4[i] = *reinterpret_cast<const float4*>(qrow + 16 * i + 4 * a);
#pragma unroll 2
    for (int r = 0; r < 4; r++) {
      const int g = __shfl_sync(0xffffffffu, gsel, r);
      if (g < 0) continue;   // warp-uniform
      const int col = g * GROUP + j;
      const float* trow = pr->t_f32 + (size_t)col * 128 + 4 * a;   // padding rows exist up to n_pad
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float4 tv = *reinterpret_cast<const float4*>(trow + 16 * i);
        float d;
        d = __fsub_rn(q4[i].x, tv.x); acc.x = __fadd_rn(acc.x, __fmul_rn(d, d));
        d = __fsub_rn(q4[i].y, tv.y); acc.y = __fadd_rn(acc.y, __fmul_rn(d, d));
        d = __fsub_rn(q4[i].z, tv.z); acc.z = __fadd_rn(acc.z, __fmul_rn(d, d));
        d = __fsub_rn(q4[i].w, tv.w); acc.w = __fadd_rn(acc.w, __fmul_rn(d, d));
      }
      float S[4];
      const float av[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int l = 0; l < 4; l++) {
        const float a0 = __shfl_sync(0xffffffffu, av[l], quad0 + 0);
        const float a1 = __shfl_sync(0xffffffffu, av[l], quad0 + 1);
        const float a2 = __shfl_sync(0xffffffffu, av[l], quad0 + 2);
        const float a3 = __shfl_sync(0xffffffffu, av[l], quad0 + 3);
        S[l] = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
      }
      const float d = sqrtf(__fadd_rn(__fadd_rn(S[0], S[2]), __fadd_rn(S[1], S[3])));
      if (a == 0 && col < pr->t_n)
        key_top2(((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)col, e0, e1);
    }
    warp_top2(e0, e1);
  }
  bool certified = certified_against(fminf(rest, rest2), e1);
  if (!certified) {   // warp-uniform
    // stage 1: the four chunks completely, one column per lane per chunk
    e0 = ~0ull; e1 = ~0ull;
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
      const int g = __shfl_sync(0xffffffffu, gsel, r);
      const int col = (g >> 2) * 32 + lane;
      if (g >= 0 && col < pr->t_n) {
        const float d = l2_cv_order(qrow, pr->t_f32 + (size_t)col * 128);
        key_top2(((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)col, e0, e1);
      }
    }
    warp_top2(e0, e1);
    certified = certified_against(rest, e1);
  }
  if (lane == 0) {
    uint4 rec = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
    if (e0 != ~0ull) { rec.x = (uint32_t)(e0 >> 32); rec.y = (uint32_t)e0; }
    if (e1 != ~0ull) { rec.z = (uint32_t)(e1 >> 32); rec.w = (uint32_t)e1; }
    G.part[((size_t)pair * G.n_split) * G.nq + q] = rec;
    const uint4 none = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
    for (int sp = 1; sp < G.n_split; sp++) G.part[((size_t)pair * G.n_split + sp) * G.nq + q] = none;
    // not certified even after all four chunks: exact full-row fallback
    if (!certified) {
      const int at = atomicAdd(G.fb_count, 1);
      G.fb_list[at] = make_uint2((uint32_t)pair, (uint32_t)q);
    }
  }
}

// Exact full-row scan for the rows the certificate could not clear, OpenCV's arithmetic
// throughout.  A work item is (listed row, segment of the train set): with few rows listed (the
// usual case: a handful per pair) each row is cut into up to 32 segments so that the scan is
// spread over the whole grid instead of running ~40 dependent distance evaluations per thread
// in a few blocks; the block that finishes a row's last segment merges the partial top-2 keys
// (a key is distance bits : train index, so the merge is order independent).
constexpr int FB_GRID = SIFT_GEN_FB_ITEMS;
__global__ void __launch_bounds__(256) sift_gen_fallback_kernel(const GenParams G) {
  __shared__ __align__(16) float qs[128];
  __shared__ unsigned long long red[2][8];
  const int n = *G.fb_count;
  if (n <= 0) return;
  const int S = n >= FB_GRID ? 1 : min(32, FB_GRID / n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int item = blockIdx.x; item < n * S; item += gridDim.x) {
    const int i = item / S, seg = item - i * S;
    const uint2 w = G.fb_list[i];
    const int pair = (int)w.x, q = (int)w.y;
    const TcPair* pr = G.pairs + pair;
    __syncthreads();
    if (threadIdx.x < 128) qs[threadIdx.x] = G.q_f32[(size_t)q * 128 + threadIdx.x];
    __syncthreads();
    const int len = (pr->t_n + S - 1) / S;
    const int t_end = min(pr->t_n, (seg + 1) * len);
    unsigned long long k0 = ~0ull, k1 = ~0ull;
    for (int t = seg * len + threadIdx.x; t < t_end; t += 256) {
      const float d = l2_cv_order(qs, pr->t_f32 + (size_t)t * 128);
      key_top2(((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)t, k0, k1);
    }
    warp_top2(k0, k1);
    if (lane == 0) { red[0][warp] = k0; red[1][warp] = k1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b0 = ~0ull, b1 = ~0ull;
      for (int k = 0; k < 8; k++) { key_top2(red[0][k], b0, b1); key_top2(red[1][k], b0, b1); }
      bool last = true;
      if (S > 1) {
        // n * S <= FB_GRID work items: one scratch slot each
        G.fb_part[2 * item] = b0;
        G.fb_part[2 * item + 1] = b1;
        __threadfence();
        last = atomicAdd(G.fb_done + i, 1) == S - 1;
        if (last) {
          __threadfence();
          b0 = ~0ull; b1 = ~0ull;
          for (int k = 0; k < S; k++) {
            key_top2(__ldcg(G.fb_part + 2 * (i * S + k)), b0, b1);
            key_top2(__ldcg(G.fb_part + 2 * (i * S + k) + 1), b0, b1);
          }
        }
      }
      if (last) {
        uint4 rec = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
        if (b0 != ~0ull) { rec.x = (uint32_t)(b0 >> 32); rec.y = (uint32_t)b0; }
        if (b1 != ~0ull) { rec.z = (uint32_t)(b1 >> 32); rec.w = (uint32_t)b1; }
        G.part[((size_t)pair * G.n_split) * G.nq + q] = rec;
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace

int g_tc_mode = 0;
extern "C" int slamb200_dbg_set_tc_mode(int m) { g_tc_mode = m; return 0; }
// the time stamps of the last mode-8 launch: out[160][8] (ns, %globaltimer; [7] = tiles issued)
extern "C" int slamb200_dbg_tc_trace(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(unsigned long long) * 160 * 8) == cudaSuccess ? 0 : -1;
}

// Encodes a frame's tensor maps into host_out (4 x 128 B): [0] main (hi), [1] aug (query role),
// [2] aug (train role), [3] lo half.
//   main / lo: bf16 [n_pad][128] row-major, box 64 x 128, SWIZZLE_128B
//   aug : the interleaved K-augmentation block viewed as bytes [n_pad/8][256], box 256 x 16
//         (= 128 rows, a dense 4 KB copy), no swizzle
int tc_encode_tmaps(const void* bf16_dev, const void* augq_dev, const void* augt_dev,
                    const void* bf16lo_dev, int n_pad, void* host_out_512B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  CUtensorMap* out = reinterpret_cast<CUtensorMap*>(host_out_512B);
  const void* mains[2] = {bf16_dev, bf16lo_dev};
  for (int i = 0; i < 2; i++) {
    cuuint64_t dims[2] = {128, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&out[i == 0 ? 0 : 3], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(mains[i]),
                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -2;
  }
  const void* aug[2] = {augq_dev, augt_dev};
  for (int i = 0; i < 2; i++) {
    cuuint64_t dims[2] = {256, (cuuint64_t)(n_pad / 8)};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {256, 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&out[1 + i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(aug[i]), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -3;
  }
  return 0;
}

size_t tc_smem_bytes() { return SMEM_BYTES; }

int tc_slots(int n_cb_max, int total_tiles, int n_cta) {
  const int tpc = total_tiles / n_cta > 0 ? total_tiles / n_cta : 1;
  int segs = (n_cb_max - 1) / tpc + 2;
  if (segs > n_cb_max) segs = n_cb_max;
  if (segs < 1) segs = 1;
  return COL_SPLITS * segs;
}

static bool tc_attrs_once() {
  static PerDeviceOnce attr_once;   // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute
  return attr_once.run([] {
    return cudaFuncSetAttribute(sift_tc_kernel<false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(sift_tc_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SMEM_BYTES_TAIL) == cudaSuccess &&
           cudaFuncSetAttribute(sift_tc_kernel<true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(sift_tc_kernel<false, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(sift_tc_kernel<false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SMEM_BYTES_G) == cudaSuccess &&
           cudaFuncSetAttribute(sift_tc_kernel<true, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SMEM_BYTES_G) == cudaSuccess;
  });
}

// pairs_host != nullptr: the table travels in the parameters (n_pairs <= TC_INLINE_MAX)
static void tc_fill_inline(InlineTables& IT, const TcPair* pairs_host, const int32_t* prefix_host, int n_pairs) {
  IT.n = 0;
  if (!pairs_host || n_pairs > TC_INLINE_MAX) return;
  IT.n = n_pairs;
  memcpy(IT.pairs, pairs_host, sizeof(TcPair) * (size_t)n_pairs);
  memcpy(IT.prefix, prefix_host, sizeof(int32_t) * (size_t)(n_pairs + 1));
}
int tc_inline_max() { return TC_INLINE_MAX; }

static void tc_fill_params(TcParams& P, const void* q_tmaps_host_384B, const int32_t* q_flags, int nq,
                           const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                           int total_tiles, int n_slots, uint4* cand, int32_t* err_flag, float* dbg, int fp8,
                           int kinds_known, int wide) {
  const int n_rb = (nq + 2 * BM - 1) / (2 * BM);
  memcpy(P.q_tmap, q_tmaps_host_384B, 384);   // {main, aug (query role), lo}
  P.pairs = pairs_dev;
  P.tile_prefix = tile_prefix_dev;
  P.q_flags = q_flags;
  P.n_pairs = n_pairs;
  P.n_rb = n_rb;
  P.total_tiles = total_tiles;
  P.n_slots = n_slots;
  P.nq_pad = n_rb * 2 * BM;
  P.cand = cand;
  P.dbg = dbg;
  P.err_flag = err_flag;
  P.mode = g_tc_mode;
  P.fp8 = fp8;
  P.kinds_known = kinds_known;
  P.wide = wide;
}

int launch_sift_tc_candidates(const void* q_tmaps_host_384B, const int32_t* q_flags, int nq,
                              const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                              int total_tiles, int n_cta_pairs, int n_slots, uint4* cand,
                              int32_t* err_flag, float* dbg, int gen, cudaStream_t s, int fp8,
                              int kinds_known, int wide, const TcPair* inl_pairs_host,
                              const int32_t* inl_prefix_host) {
  if (!tc_attrs_once()) return -1;
  if (total_tiles <= 0 || nq <= 0) return 0;
  InlineTables IT;
  tc_fill_inline(IT, inl_pairs_host, inl_prefix_host, n_pairs);
  TcParams P;
  tc_fill_params(P, q_tmaps_host_384B, q_flags, nq, pairs_dev, tile_prefix_dev, n_pairs, total_tiles, n_slots,
                 cand, err_flag, dbg, fp8, kinds_known, wide);
  TcTail TT;
  memset(&TT, 0, sizeof(TT));
  const dim3 grid(2 * n_cta_pairs);
  if (gen) {
    if (dbg) sift_tc_kernel<true, true, false, false><<<grid, TC_THREADS, SMEM_BYTES_G, s>>>(P, TT, IT);
    else sift_tc_kernel<false, true, false, false><<<grid, TC_THREADS, SMEM_BYTES_G, s>>>(P, TT, IT);
  } else {
    if (dbg) sift_tc_kernel<true, false, false, false><<<grid, TC_THREADS, SMEM_BYTES, s>>>(P, TT, IT);
    else if (g_tc_mode != 0) sift_tc_kernel<false, false, true, false><<<grid, TC_THREADS, SMEM_BYTES, s>>>(P, TT, IT);   // role ablations
    else sift_tc_kernel<false, false, false, false><<<grid, TC_THREADS, SMEM_BYTES, s>>>(P, TT, IT);
  }
  COUNT_LAUNCH();
  return 0;
}

// The whole match path of a batch of integer-valued (or ORB) pairs in ONE kernel: candidates on the
// tensor cores, and the tail (slot merge, pruning, best-group rerank, ratio test, ordered
// compaction) on two otherwise idle warps of every CTA while the tiles go on.  Requires every pair
// to have at least as many tiles as there are CTA pairs (no empty shares) and the kinds known on
// the host.  scan: one word per (pair, row block, CTA of the pair), zeroed when allocated;
// seg_done: one int32 per (pair, row block, CTA of the pair), zeroed when allocated.
int launch_sift_tc_match(const void* q_tmaps_host_384B, const int32_t* q_flags, const uint8_t* q_u8,
                         const int32_t* q_nrm2, int nq, const TcPair* pairs_dev,
                         const int32_t* tile_prefix_dev, int n_pairs, int total_tiles, int n_cta_pairs,
                         int n_slots, uint4* cand, int32_t* err_flag, double ratio, int fp8, int wide,
                         unsigned long long* scan, uint32_t epoch, int32_t* seg_done, slamb200_dmatch* out,
                         int cap, int32_t* n_out, cudaStream_t s, const TcPair* inl_pairs_host,
                         const int32_t* inl_prefix_host) {
  if (!tc_attrs_once()) return -1;
  if (total_tiles <= 0 || nq <= 0) return 0;
  InlineTables IT;
  tc_fill_inline(IT, inl_pairs_host, inl_prefix_host, n_pairs);
  TcParams P;
  tc_fill_params(P, q_tmaps_host_384B, q_flags, nq, pairs_dev, tile_prefix_dev, n_pairs, total_tiles, n_slots,
                 cand, err_flag, nullptr, fp8, 1, wide);
  P.mode = 0;
  TcTail TT;
  memset(&TT, 0, sizeof(TT));
  RerankParams& R = TT.R;
  R.wide = wide;
  R.orb = fp8;
  R.q_u8 = q_u8; R.q_nrm2 = q_nrm2; R.q_flags = q_flags; R.pairs = pairs_dev; R.cand = cand;
  R.tile_prefix = tile_prefix_dev; R.n_cta = n_cta_pairs;
  R.nq = nq; R.nq_pad = P.nq_pad; R.n_slots = n_slots; R.n_pairs = n_pairs;
  R.n_split = 1; R.part = nullptr; R.err_flag = err_flag;
  R.prune = (ratio >= 0.0 && ratio < 1e300) ? 1 : 0; R.ratio = ratio;
  TT.C.scan = scan; TT.C.epoch = epoch; TT.C.out = out; TT.C.cap = cap; TT.C.n_out = n_out;
  TT.seg_done = seg_done;
  const dim3 grid(2 * n_cta_pairs);
  sift_tc_kernel<false, false, false, true><<<grid, TC_THREADS, SMEM_BYTES_TAIL, s>>>(P, TT, IT);
  COUNT_LAUNCH();
  return 0;
}

void launch_sift_gen_rerank(const int32_t* q_flags, const float* q_f32, const float* q_nrmf, int nq,
                            const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                            int n_cta_pairs, int n_slots, int n_split, const uint4* cand, uint4* part,
                            uint2* fb_list, int32_t* fb_count, unsigned long long* fb_part,
                            int32_t* fb_done, cudaStream_t s, int wide) {
  if (nq <= 0 || n_pairs <= 0) return;
  GenParams G;
  G.wide = wide;
  G.q_f32 = q_f32; G.q_nrmf = q_nrmf; G.q_flags = q_flags; G.pairs = pairs_dev;
  G.tile_prefix = tile_prefix_dev; G.cand = cand;
  G.nq = nq; G.nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM); G.n_slots = n_slots;
  G.n_pairs = n_pairs; G.n_split = n_split; G.n_cta = n_cta_pairs;
  G.part = part; G.fb_list = fb_list; G.fb_count = fb_count;
  G.fb_part = fb_part; G.fb_done = fb_done;
  dim3 grid((nq + 7) / 8, n_pairs);
  sift_gen_rerank_kernel<<<grid, 256, 0, s>>>(G);
  COUNT_LAUNCH();
  sift_gen_fallback_kernel<<<FB_GRID, 256, 0, s>>>(G);
  COUNT_LAUNCH();
}

void launch_sift_rerank(const int32_t* q_flags, const uint8_t* q_u8, const int32_t* q_nrm2, int nq,
                        const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                        int n_cta_pairs, int n_slots, int n_split, const uint4* cand, uint4* part,
                        uint4* work, float2* work_v0, int32_t* work_n, int32_t* err_flag, int prune,
                        double ratio, cudaStream_t s, int orb, int wide) {
  if (nq <= 0 || n_pairs <= 0) return;
  RerankParams R;
  R.wide = wide;
  R.orb = orb;
  R.q_u8 = q_u8; R.q_nrm2 = q_nrm2; R.q_flags = q_flags; R.pairs = pairs_dev; R.cand = cand;
  R.tile_prefix = tile_prefix_dev; R.n_cta = n_cta_pairs;
  R.nq = nq; R.nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM); R.n_slots = n_slots; R.n_pairs = n_pairs;
  R.n_split = n_split; R.part = part; R.err_flag = err_flag;
  R.work = work; R.work_v0 = work_v0; R.work_n = work_n;
  R.prune = (prune && ratio >= 0.0 && ratio < 1e300) ? 1 : 0; R.ratio = ratio;
  dim3 grid((nq + 255) / 256, n_pairs);
  launch_pdl(sift_merge_kernel, grid, dim3(256), 0, s, R);
  COUNT_LAUNCH();
  const long long rows = (long long)nq * n_pairs;
  const int blocks = (int)((rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8);
  if (orb) {
    if (prune) launch_pdl(orb_rerank_lite_kernel, dim3(blocks), dim3(256), 0, s, R);
    else launch_pdl(orb_rerank_kernel, dim3(blocks), dim3(256), 0, s, R);
  } else if (prune) {
    launch_pdl(sift_rerank_lite_kernel, dim3(blocks), dim3(256), 0, s, R);   // match output: distances + best index
  } else {
    launch_pdl(sift_rerank_kernel, dim3(blocks), dim3(256), 0, s, R);        // raw k-NN output: both indices
  }
  COUNT_LAUNCH();
}

// The fused tail of the match path: fills knn_idx[.][0], knn_dist[.][0], flags and chunk_cnt for
// compact_kernel (launch_compact, finalize.cu).
void launch_tc_tail_fused(const int32_t* q_flags, const uint8_t* q_u8, const int32_t* q_nrm2, int nq,
                          const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                          int n_cta_pairs, int n_slots, int n_split, const uint4* cand, uint4* part,
                          int32_t* err_flag, double ratio, int orb, int32_t* knn_idx, float* knn_dist,
                          uint8_t* flags, int32_t* chunk_cnt, cudaStream_t s, int wide) {
  if (nq <= 0 || n_pairs <= 0) return;
  RerankParams R;
  R.wide = wide;
  R.orb = orb;
  R.q_u8 = q_u8; R.q_nrm2 = q_nrm2; R.q_flags = q_flags; R.pairs = pairs_dev; R.cand = cand;
  R.tile_prefix = tile_prefix_dev; R.n_cta = n_cta_pairs;
  R.nq = nq; R.nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM); R.n_slots = n_slots; R.n_pairs = n_pairs;
  R.n_split = n_split; R.part = part; R.err_flag = err_flag;
  R.work = nullptr; R.work_v0 = nullptr; R.work_n = nullptr;
  R.prune = (ratio >= 0.0 && ratio < 1e300) ? 1 : 0; R.ratio = ratio;
  CompactArgs C;
  memset(&C, 0, sizeof(C));
  C.knn_idx = knn_idx; C.knn_dist = knn_dist; C.flags = flags; C.chunk_cnt = chunk_cnt;
  InlineTables IT;
  IT.n = 0;
  // (one survivor row per lane group and pass: with thousands of blocks in flight the occupancy of
  // the 64-register form hides the latency; two rows in flight measured 3 % slower)
  dim3 grid((nq + 255) / 256, n_pairs);
  if (orb) launch_pdl(tc_tail_fused_kernel<true, 1>, grid, dim3(256), 0, s, R, C, IT);
  else launch_pdl(tc_tail_fused_kernel<false, 1>, grid, dim3(256), 0, s, R, C, IT);
  COUNT_LAUNCH();
}

// Tail + ordered compaction in one kernel (tc_tail_compact_kernel): the match lists land in
// out[pair][cap] / n_out[pair] directly.  `scan` holds one 64-bit word per (pair, 256-row block),
// zeroed once when allocated; `epoch` (1 .. 2^30 - 1) must differ between launches that share it.
void launch_tc_tail_compact(const int32_t* q_flags, const uint8_t* q_u8, const int32_t* q_nrm2, int nq,
                            const TcPair* pairs_dev, const int32_t* tile_prefix_dev, int n_pairs,
                            int n_cta_pairs, int n_slots, int n_split, const uint4* cand, const uint4* part,
                            int32_t* err_flag, double ratio, int orb, unsigned long long* scan,
                            uint32_t epoch, slamb200_dmatch* out, int cap, int32_t* n_out, cudaStream_t s,
                            int wide, const TcPair* inl_pairs_host, const int32_t* inl_prefix_host) {
  if (nq <= 0 || n_pairs <= 0) return;
  InlineTables IT;
  tc_fill_inline(IT, inl_pairs_host, inl_prefix_host, n_pairs);
  RerankParams R;
  R.wide = wide;
  R.orb = orb;
  R.q_u8 = q_u8; R.q_nrm2 = q_nrm2; R.q_flags = q_flags; R.pairs = pairs_dev; R.cand = cand;
  R.tile_prefix = tile_prefix_dev; R.n_cta = n_cta_pairs;
  R.nq = nq; R.nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM); R.n_slots = n_slots; R.n_pairs = n_pairs;
  R.n_split = n_split; R.part = const_cast<uint4*>(part); R.err_flag = err_flag;
  R.work = nullptr; R.work_v0 = nullptr; R.work_n = nullptr;
  R.prune = (ratio >= 0.0 && ratio < 1e300) ? 1 : 0; R.ratio = ratio;
  CompactArgs C;
  memset(&C, 0, sizeof(C));
  C.scan = scan; C.epoch = epoch; C.out = out; C.cap = cap; C.n_out = n_out;
  // The tcgen05 kernel in front of this one (and behind it, in a loop of calls) runs with the
  // largest shared-memory carve-out; asking for the same one here spares the SMs a reconfiguration
  // of their L1 / shared split between the two kernels of a call.
  static PerDeviceOnce carve_once;
  carve_once.run([] {
    if (getenv("SLAMB200_NO_CARVEOUT")) return true;
    cudaFuncSetAttribute(tc_tail_compact_kernel<true, 1>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(tc_tail_compact_kernel<false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(tc_tail_compact_kernel<true, 3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(tc_tail_compact_kernel<false, 3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    return true;
  });
  dim3 grid((nq + 255) / 256, n_pairs);
  // a call of a few pairs is a chain of memory round trips (three survivor rows per lane group in
  // flight: one pass instead of three); a batch wants the occupancy of the 64-register form
  if (n_pairs <= TC_INLINE_MAX) {
    if (orb) launch_pdl(tc_tail_compact_kernel<true, 3>, grid, dim3(256), 0, s, R, C, IT);
    else launch_pdl(tc_tail_compact_kernel<false, 3>, grid, dim3(256), 0, s, R, C, IT);
  } else {
    if (orb) launch_pdl(tc_tail_compact_kernel<true, 1>, grid, dim3(256), 0, s, R, C, IT);
    else launch_pdl(tc_tail_compact_kernel<false, 1>, grid, dim3(256), 0, s, R, C, IT);
  }
  COUNT_LAUNCH();
}
