// sift_tc.cu -- SIFT L2 k=2 candidates on the 5th-generation tensor cores (tcgen05 + TMEM + TMA),
// followed by the exact integer rerank.  sm_100a only.
//
// Replaces the Q x T x 128 distance computation of cv::BFMatcher(NORM_L2)::knnMatch(query, train,
// 2) (src/mainModule/featureMatching/featureMatchingCPU.cpp:27-40) for descriptor sets in "exact
// mode" (integer-valued rows in [0,255], |row|^2 < 2^20: what cv::SIFT emits).  For those, bf16
// holds every value exactly and every partial sum stays below 2^24, so the fp32 accumulators in
// TMEM are exact and the contraction itself yields d^2/2:
//
//     acc(q,t) = -(q . t)                       8 x tcgen05.mma 128x256x16, A negated
//              + (|q|^2/2) * 1 + 1 * (|t|^2/2)  1 x tcgen05.mma on the 16-column K augmentation
//
// One persistent CTA per SM walks a contiguous range of 128 x 256 tiles (pair-major, then query
// block, then train tile).  Warp roles: warp 0 = TMA producer (query block resident per
// segment, train tiles double-buffered), warp 1 = MMA issuer (one elected thread, accumulators
// double-buffered in the 512 TMEM columns), warp 2 = TMEM allocator, warps 4..11 = epilogue.
//
// Epilogue: a thread owns one query row (TMEM lane) and half of the tile's columns.  Per 32-column
// tcgen05.ld it reduces four groups of 8 columns with 3-input FMNMX trees and only when some
// lane of the warp sees a group minimum below its running second-best does the warp fall into the
// insert path, which keeps the best two *groups* (value, group index) per row.  The true top-2
// columns of a row always lie inside its top-2 groups (ties resolve to the lowest index at every
// level), so the rerank evaluates 16 candidates per row exactly -- integer dp4a on the u8 copies --
// and emits (sqrtf(d^2), index) records that the shared finalize kernels turn into the ratio-
// tested, ordered match list.  A device-side self check compares each rerank minimum with the
// tensor-core value and raises err_flag on any mismatch (never expected).
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;             // query rows per tile (TMEM lanes)
constexpr int BN = 256;             // train rows per tile (TMEM columns per accumulator stage)
constexpr int A_KBLK = BM * 128;    // bytes of one 64-wide bf16 k-block of the query tile
constexpr int B_KBLK = BN * 128;
constexpr int A_AUG_OFF = 2 * A_KBLK;
constexpr int B_AUG_OFF = 2 * B_KBLK;
constexpr int A_STAGE = 2 * A_KBLK + BM * 32;  // 36864
constexpr int B_STAGE = 2 * B_KBLK + BN * 32;  // 73728
constexpr int N_ASTAGE = 2;
constexpr int N_BSTAGE = 2;
constexpr int SMEM_BARS = 1024;
constexpr int SMEM_BYTES = N_ASTAGE * A_STAGE + N_BSTAGE * B_STAGE + SMEM_BARS + 1024;
constexpr int TC_THREADS = 384;     // 12 warps
constexpr int EPI_WARP0 = 4;
constexpr int GROUP = 8;            // columns per candidate group

struct TcParams {
  const CUtensorMap* q_tmap;   // query main map (device memory)
  const uint8_t* q_aug;        // query aug block, interleaved layout
  const TcPair* pairs;         // per pair: train maps / aug / sizes
  const int32_t* tile_prefix;  // [P+1] tiles before pair p
  const int32_t* q_flags;
  int n_pairs;
  int n_rb;                    // query row blocks
  int total_tiles;
  int n_slots;                 // candidate slots per (pair, row) = 2 * max segments per row block
  int nq_pad;                  // n_rb * 128
  uint4* cand;                 // [pair][slot][nq_pad]
  float* dbg;                  // optional raw accumulator dump of the CTA-0 first tile
  int32_t* err_flag;
};

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
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
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_neg) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_neg << 13) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// ---- tile walk: every role iterates the same sequence -----------------------------------------
struct TileIter {
  int pair, rb, cb, n_cb;
  int tile, end;
  __device__ void init(const TcParams& P, int cta, int n_cta) {
    tile = (int)(((long long)P.total_tiles * cta) / n_cta);
    end = (int)(((long long)P.total_tiles * (cta + 1)) / n_cta);
    pair = 0;
    if (tile < end) {
      while (P.tile_prefix[pair + 1] <= tile) pair++;
      n_cb = (P.tile_prefix[pair + 1] - P.tile_prefix[pair]) / P.n_rb;
      const int local = tile - P.tile_prefix[pair];
      rb = local / n_cb;
      cb = local - rb * n_cb;
    } else {
      n_cb = 1; rb = 0; cb = 0;
    }
  }
  __device__ bool valid() const { return tile < end; }
  // returns true when the next tile starts a new (pair, row block) segment
  __device__ bool next(const TcParams& P) {
    tile++;
    if (tile >= end) return true;
    if (++cb < n_cb) return false;
    cb = 0;
    if (++rb < P.n_rb) return true;
    rb = 0;
    do { pair++; } while (P.tile_prefix[pair + 1] == P.tile_prefix[pair]);
    n_cb = (P.tile_prefix[pair + 1] - P.tile_prefix[pair]) / P.n_rb;
    return true;
  }
};

__device__ __forceinline__ int cta_range_begin(int total, int c, int n) {
  return (int)(((long long)total * c) / n);
}
__device__ int owner_cta(int total, int n, int x) {
  int c = (int)(((long long)x * n) / total);
  if (c >= n) c = n - 1;
  while (c + 1 < n && cta_range_begin(total, c + 1, n) <= x) c++;
  while (c > 0 && cta_range_begin(total, c, n) > x) c--;
  return c;
}

__device__ __forceinline__ void top2_group_insert(float g, int gid, float& m1, int& i1, float& m2,
                                                  int& i2) {
  const bool lt1 = g < m1;
  const bool lt2 = g < m2;
  m2 = lt1 ? m1 : (lt2 ? g : m2);
  i2 = lt1 ? i1 : (lt2 ? gid : i2);
  m1 = lt1 ? g : m1;
  i1 = lt1 ? gid : i1;
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

__device__ __forceinline__ void process_chunk(const uint32_t (&v)[32], int gid0, float& m1, int& i1,
                                              float& m2, int& i2) {
  float g[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float a = fmin3(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]),
                          __uint_as_float(v[8 * j + 2]));
    const float b = fmin3(__uint_as_float(v[8 * j + 3]), __uint_as_float(v[8 * j + 4]),
                          __uint_as_float(v[8 * j + 5]));
    g[j] = fmin3(a, b, fminf(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
  }
  const float cmin = fmin3(g[0], g[1], fminf(g[2], g[3]));
  if (__any_sync(0xffffffffu, cmin < m2)) {
#pragma unroll
    for (int j = 0; j < 4; j++) top2_group_insert(g[j], gid0 + j, m1, i1, m2, i2);
  }
}

template <bool DBG>
__global__ void __launch_bounds__(TC_THREADS, 1) sift_tc_kernel(const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the SWIZZLE_128B atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + N_ASTAGE * A_STAGE;
  const uint32_t bar_base = b_base + N_BSTAGE * B_STAGE;
  // barrier slots (8 B each)
  const uint32_t a_full = bar_base, a_empty = bar_base + 16;
  const uint32_t b_full = bar_base + 32, b_empty = bar_base + 48;
  const uint32_t t_full = bar_base + 64, t_empty = bar_base + 80;
  const uint32_t tmem_slot = bar_base + 96;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + N_ASTAGE * A_STAGE + N_BSTAGE * B_STAGE + 96);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_cta = gridDim.x;
  const int cta = blockIdx.x;

  if (P.q_flags[0] != 0) return;  // general-float query: the exact fp32 kernel owns this batch

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(a_full + 8 * s, 1);
      mbar_init(a_empty + 8 * s, 1);
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
      mbar_init(t_full + 8 * s, 1);
      mbar_init(t_empty + 8 * s, 8);  // one elected lane of each of the 8 epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      TileIter it;
      it.init(P, cta, n_cta);
      int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0;
      bool new_seg = true;
      while (it.valid()) {
        const TcPair pr = P.pairs[it.pair];
        if (pr.t_flags[0] != 0) {  // general-float train set: skipped here (exact kernel)
          new_seg = it.next(P) || new_seg;
          continue;
        }
        if (new_seg) {
          mbar_wait(a_empty + 8 * a_stage, a_phase ^ 1);
          const uint32_t dst = a_base + a_stage * A_STAGE;
          const uint32_t bar = a_full + 8 * a_stage;
          mbar_expect_tx(bar, A_STAGE);
          tma_load_2d(dst, P.q_tmap, bar, 0, it.rb * BM);
          tma_load_2d(dst + A_KBLK, P.q_tmap, bar, 64, it.rb * BM);
          bulk_load_1d(dst + A_AUG_OFF, P.q_aug + (size_t)it.rb * BM * 32, BM * 32, bar);
          if (++a_stage == N_ASTAGE) { a_stage = 0; a_phase ^= 1; }
        }
        mbar_wait(b_empty + 8 * b_stage, b_phase ^ 1);
        {
          const uint32_t dst = b_base + b_stage * B_STAGE;
          const uint32_t bar = b_full + 8 * b_stage;
          const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(pr.tmap_main);
          const int row0 = it.cb * BN;
          mbar_expect_tx(bar, B_STAGE);
          tma_load_2d(dst, tm, bar, 0, row0);
          tma_load_2d(dst + BM * 128, tm, bar, 0, row0 + 128);
          tma_load_2d(dst + B_KBLK, tm, bar, 64, row0);
          tma_load_2d(dst + B_KBLK + BM * 128, tm, bar, 64, row0 + 128);
          bulk_load_1d(dst + B_AUG_OFF, pr.t_aug + (size_t)row0 * 32, BN * 32, bar);
          if (++b_stage == N_BSTAGE) { b_stage = 0; b_phase ^= 1; }
        }
        new_seg = it.next(P);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      TileIter it;
      it.init(P, cta, n_cta);
      int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0, t_stage = 0, t_phase = 0;
      int cur_a = 0;
      bool new_seg = true;
      constexpr uint32_t IDESC_NEG = idesc_bf16(BM, BN, 1);
      constexpr uint32_t IDESC_POS = idesc_bf16(BM, BN, 0);
      while (it.valid()) {
        const TcPair pr = P.pairs[it.pair];
        if (pr.t_flags[0] != 0) {
          new_seg = it.next(P) || new_seg;
          continue;
        }
        if (new_seg) {
          mbar_wait(a_full + 8 * a_stage, a_phase);
          cur_a = a_stage;
          if (++a_stage == N_ASTAGE) { a_stage = 0; a_phase ^= 1; }
        }
        mbar_wait(b_full + 8 * b_stage, b_phase);
        mbar_wait(t_empty + 8 * t_stage, t_phase ^ 1);
        tc_fence_after();
        const uint32_t a_addr = a_base + cur_a * A_STAGE;
        const uint32_t b_addr = b_base + b_stage * B_STAGE;
        const uint32_t d_tmem = tmem_base + t_stage * BN;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint64_t ad = desc_sw128(a_addr + (k >> 2) * A_KBLK + (k & 3) * 32);
          const uint64_t bd = desc_sw128(b_addr + (k >> 2) * B_KBLK + (k & 3) * 32);
          tc_mma(d_tmem, ad, bd, IDESC_NEG, k > 0 ? 1u : 0u);
        }
        tc_mma(d_tmem, desc_interleave(a_addr + A_AUG_OFF), desc_interleave(b_addr + B_AUG_OFF),
               IDESC_POS, 1u);
        tc_commit(b_empty + 8 * b_stage);
        tc_commit(t_full + 8 * t_stage);
        if (++b_stage == N_BSTAGE) { b_stage = 0; b_phase ^= 1; }
        if (++t_stage == 2) { t_stage = 0; t_phase ^= 1; }
        new_seg = it.next(P);
        if (new_seg) tc_commit(a_empty + 8 * cur_a);  // the segment's MMAs are done with A
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================= epilogue =================
    const int ew = warp - EPI_WARP0;
    const int quarter = ew & 3;   // TMEM lane quarter this warp may read
    const int half = ew >> 2;     // which 128 columns of the tile
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int row_in_tile = quarter * 32 + lane;
    TileIter it;
    it.init(P, cta, n_cta);
    int t_stage = 0, t_phase = 0;
    float m1 = __int_as_float(0x7f800000), m2 = m1;
    int i1 = -1, i2 = -1;
    int seg_pair = -1, seg_rb = 0, seg_first_tile = 0;
    bool new_seg = true;
    bool first_tile_of_cta = true;
    while (it.valid()) {
      const TcPair pr = P.pairs[it.pair];
      if (pr.t_flags[0] != 0) {
        new_seg = it.next(P) || new_seg;
        continue;
      }
      if (new_seg) {
        seg_pair = it.pair; seg_rb = it.rb; seg_first_tile = it.tile;
        m1 = m2 = __int_as_float(0x7f800000);
        i1 = i2 = -1;
      }
      mbar_wait(t_full + 8 * t_stage, t_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + lane_addr + t_stage * BN + half * 128;
      const int gid_tile = (it.cb * BN + half * 128) / GROUP;
      uint32_t va[32], vb[32];
      TMEM_LD32(t_addr, va);
      TMEM_WAIT32(va);
      TMEM_LD32(t_addr + 32, vb);
      if (DBG && first_tile_of_cta && cta == 0) {
#pragma unroll
        for (int j = 0; j < 32; j++)
          P.dbg[(size_t)row_in_tile * BN + half * 128 + j] = __uint_as_float(va[j]);
      }
      process_chunk(va, gid_tile + 0, m1, i1, m2, i2);
      TMEM_WAIT32(vb);
      TMEM_LD32(t_addr + 64, va);
      if (DBG && first_tile_of_cta && cta == 0) {
#pragma unroll
        for (int j = 0; j < 32; j++)
          P.dbg[(size_t)row_in_tile * BN + half * 128 + 32 + j] = __uint_as_float(vb[j]);
      }
      process_chunk(vb, gid_tile + 4, m1, i1, m2, i2);
      TMEM_WAIT32(va);
      TMEM_LD32(t_addr + 96, vb);
      if (DBG && first_tile_of_cta && cta == 0) {
#pragma unroll
        for (int j = 0; j < 32; j++)
          P.dbg[(size_t)row_in_tile * BN + half * 128 + 64 + j] = __uint_as_float(va[j]);
      }
      process_chunk(va, gid_tile + 8, m1, i1, m2, i2);
      TMEM_WAIT32(vb);
      // all TMEM reads of this accumulator stage are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + 8 * t_stage);
      if (DBG && first_tile_of_cta && cta == 0) {
#pragma unroll
        for (int j = 0; j < 32; j++)
          P.dbg[(size_t)row_in_tile * BN + half * 128 + 96 + j] = __uint_as_float(vb[j]);
      }
      process_chunk(vb, gid_tile + 12, m1, i1, m2, i2);
      first_tile_of_cta = false;
      if (++t_stage == 2) { t_stage = 0; t_phase ^= 1; }
      new_seg = it.next(P);
      if (new_seg) {
        // flush this segment's per-row record
        const int rb_first = P.tile_prefix[seg_pair] +
                             seg_rb * ((P.tile_prefix[seg_pair + 1] - P.tile_prefix[seg_pair]) / P.n_rb);
        int ord = cta - owner_cta(P.total_tiles, n_cta, rb_first);
        (void)seg_first_tile;
        if (ord < 0 || 2 * ord + 1 >= P.n_slots) {
          if (lane == 0) atomicOr(P.err_flag, 2);
          ord = 0;
        }
        const int slot = 2 * ord + half;
        uint4 rec;
        rec.x = __float_as_uint(m1); rec.y = (uint32_t)i1;
        rec.z = __float_as_uint(m2); rec.w = (uint32_t)i2;
        P.cand[((size_t)seg_pair * P.n_slots + slot) * P.nq_pad + seg_rb * BM + row_in_tile] = rec;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                 : "memory");
  }
}

// ---- rerank: exact integer distances of the 16 candidates of every row ------------------------
struct RerankParams {
  const uint8_t* q_u8;
  const int32_t* q_nrm2;
  const int32_t* q_flags;
  const TcPair* pairs;
  const uint4* cand;
  int nq, nq_pad, n_slots, n_pairs, n_split;
  uint4* part;       // [pair][n_split][nq]: split 0 gets the record, the others "absent"
  int32_t* err_flag;
};

__device__ __forceinline__ bool lt_fi(float va, int ia, float vb, int ib) {
  return va < vb || (va == vb && ia < ib);
}

__global__ void __launch_bounds__(256) sift_rerank_kernel(const RerankParams R) {
  const int pair = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= R.nq) return;
  const TcPair pr = R.pairs[pair];
  if (R.q_flags[0] != 0 || pr.t_flags[0] != 0) return;  // not an exact-mode pair

  // 1. best two candidate groups over all slots (each lane reads one slot record)
  float v0 = __int_as_float(0x7f800000), v1 = v0;
  int g0 = 0x7fffffff, g1 = 0x7fffffff;
  for (int s = lane; s < R.n_slots; s += 32) {
    const uint4 rec = R.cand[((size_t)pair * R.n_slots + s) * R.nq_pad + q];
    const float a = __uint_as_float(rec.x), b = __uint_as_float(rec.z);
    const int ia = (int)rec.y, ib = (int)rec.w;
    if (ia >= 0) {
      if (lt_fi(a, ia, v0, g0)) { v1 = v0; g1 = g0; v0 = a; g0 = ia; }
      else if (lt_fi(a, ia, v1, g1)) { v1 = a; g1 = ia; }
    }
    if (ib >= 0) {
      if (lt_fi(b, ib, v0, g0)) { v1 = v0; g1 = g0; v0 = b; g0 = ib; }
      else if (lt_fi(b, ib, v1, g1)) { v1 = b; g1 = ib; }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const float ov0 = __shfl_xor_sync(0xffffffffu, v0, off), ov1 = __shfl_xor_sync(0xffffffffu, v1, off);
    const int og0 = __shfl_xor_sync(0xffffffffu, g0, off), og1 = __shfl_xor_sync(0xffffffffu, g1, off);
    if (lt_fi(ov0, og0, v0, g0)) {
      // other's best wins: second = min(mine best, other's second)
      if (lt_fi(v0, g0, ov1, og1)) { v1 = v0; g1 = g0; } else { v1 = ov1; g1 = og1; }
      v0 = ov0; g0 = og0;
    } else if (!(ov0 == v0 && og0 == g0)) {
      if (lt_fi(ov0, og0, v1, g1)) { v1 = ov0; g1 = og0; }
    } else {
      // identical best (same record seen by both): second = min of the seconds
      if (lt_fi(ov1, og1, v1, g1)) { v1 = ov1; g1 = og1; }
    }
  }

  // 2. exact integer d^2 for the 16 candidate columns: lanes 0..7 -> group g0, 8..15 -> group g1
  const int grp = lane < 8 ? g0 : g1;
  const bool has_grp = lane < 16 && grp != 0x7fffffff;
  const int col = has_grp ? grp * GROUP + (lane & 7) : -1;
  uint32_t d2 = 0xFFFFFFFFu;
  if (has_grp && col < pr.t_n) {
    const uint4* qp = reinterpret_cast<const uint4*>(R.q_u8 + (size_t)q * 128);
    const uint4* tp = reinterpret_cast<const uint4*>(pr.t_u8 + (size_t)col * 128);
    uint32_t dot = 0;  // u8 x u8 products: unsigned dp4a
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint4 a = qp[k], b = tp[k];
      dot = __dp4a(a.x, b.x, dot);
      dot = __dp4a(a.y, b.y, dot);
      dot = __dp4a(a.z, b.z, dot);
      dot = __dp4a(a.w, b.w, dot);
    }
    d2 = (uint32_t)(R.q_nrm2[q] + pr.t_nrm2[col]) - 2u * dot;
  }
  // self check: the minimum of group g0 must equal twice the tensor-core value
  {
    uint32_t mn = lane < 8 ? d2 : 0xFFFFFFFFu;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    if (lane == 0 && g0 != 0x7fffffff && mn != 0xFFFFFFFFu) {
      if ((float)mn != 2.0f * v0) atomicOr(R.err_flag, 1);
    }
  }
  // 3. top-2 by (d2, col) across the 16 lanes
  unsigned long long key = d2 == 0xFFFFFFFFu ? ~0ull : (((unsigned long long)d2 << 32) | (uint32_t)col);
  unsigned long long k0 = key;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k0, off);
    k0 = o < k0 ? o : k0;
  }
  unsigned long long k1 = key == k0 ? ~0ull : key;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k1, off);
    k1 = o < k1 ? o : k1;
  }
  if (lane == 0) {
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
    const uint4 none = make_uint4(ABSENT_KEY, 0xFFFFFFFFu, ABSENT_KEY, 0xFFFFFFFFu);
    for (int sp = 1; sp < R.n_split; sp++) R.part[((size_t)pair * R.n_split + sp) * R.nq + q] = none;
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

// Encodes the frame's main tensor map: bf16 [n_pad][128] row-major, box 64 x 128, SWIZZLE_128B.
int tc_encode_tmap(const void* bf16_dev, int n_pad, void* host_out_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {128, (cuuint64_t)n_pad};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(host_out_128B), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(bf16_dev), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

size_t tc_smem_bytes() { return SMEM_BYTES; }

int tc_slots(int n_cb_max, int total_tiles, int n_cta) {
  const int tpc = total_tiles / n_cta > 0 ? total_tiles / n_cta : 1;
  int segs = (n_cb_max - 1) / tpc + 2;
  if (segs > n_cb_max) segs = n_cb_max;
  if (segs < 1) segs = 1;
  return 2 * segs;
}

int launch_sift_tc(const void* q_tmap_dev, const uint8_t* q_aug, const int32_t* q_flags,
                   const uint8_t* q_u8, const int32_t* q_nrm2, int nq, const TcPair* pairs_dev,
                   const int32_t* tile_prefix_dev, int n_pairs, int total_tiles, int n_cta,
                   int n_slots, int n_split, uint4* cand, uint4* part, int32_t* err_flag,
                   float* dbg, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(sift_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(sift_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM_BYTES) != cudaSuccess)
      return -1;
    attr_done = true;
  }
  const int n_rb = (nq + BM - 1) / BM;
  if (total_tiles > 0 && nq > 0) {
    TcParams P;
    P.q_tmap = reinterpret_cast<const CUtensorMap*>(q_tmap_dev);
    P.q_aug = q_aug;
    P.pairs = pairs_dev;
    P.tile_prefix = tile_prefix_dev;
    P.q_flags = q_flags;
    P.n_pairs = n_pairs;
    P.n_rb = n_rb;
    P.total_tiles = total_tiles;
    P.n_slots = n_slots;
    P.nq_pad = n_rb * BM;
    P.cand = cand;
    P.dbg = dbg;
    P.err_flag = err_flag;
    if (dbg)
      sift_tc_kernel<true><<<n_cta, TC_THREADS, SMEM_BYTES, s>>>(P);
    else
      sift_tc_kernel<false><<<n_cta, TC_THREADS, SMEM_BYTES, s>>>(P);
    COUNT_LAUNCH();
  }
  if (nq > 0 && n_pairs > 0) {
    RerankParams R;
    R.q_u8 = q_u8; R.q_nrm2 = q_nrm2; R.q_flags = q_flags; R.pairs = pairs_dev; R.cand = cand;
    R.nq = nq; R.nq_pad = n_rb * BM; R.n_slots = n_slots; R.n_pairs = n_pairs; R.n_split = n_split;
    R.part = part; R.err_flag = err_flag;
    dim3 grid((nq + 7) / 8, n_pairs);
    sift_rerank_kernel<<<grid, 256, 0, s>>>(R);
    COUNT_LAUNCH();
  }
  return 0;
}
