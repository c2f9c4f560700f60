// api.cu -- the C ABI of libslamb200 (include/slamb200.h): context, lanes, HBM-resident
// descriptor sets, batch orchestration.  All arithmetic is in the kernels; this file only moves
// bytes and sequences launches.  No CPU fallback exists anywhere in this library.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
// kernels launched by this library since load: bumped from concurrently running lanes
static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// last-error text for the other translation units of the library (device_set.cu)
int set_error(int code, const char* msg) { return fail(code, "%s", msg); }


#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail(e__ == cudaErrorMemoryAllocation ? SLAMB200_ERR_NOMEM                \
                                                   : SLAMB200_ERR_CUDA,                \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,   \
                  __LINE__);                                                           \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

#define N_LANES 10
#define N_UPLOAD_LANES 4  // the last lanes: descriptor uploads never queue in front of a match

struct Lane {
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;      // rerank / finalize side of the sub-batch pipeline
  std::vector<cudaEvent_t> sub_ev;
  bool busy = false;
  // matching scratch (device)
  DevBuf pairs, tcpairs, tile_prefix, part, cand, cand_g, work, work_v0, fb_list, knn_idx, knn_dist, flags, chunk_cnt, out, n_out, err_flag, scan, seg_done, status_fixed;
  uint32_t scan_epoch = 0;           // launches that have used `scan` (look-back words carry it)
  // scoring scratch (device)
  DevBuf npts, E, counts, best, mask, m_off, p1, p2, txy, all_masks;
  DevBuf orb_img, orb_gray, orb_rowf, orb_blur, orb_kp, orb_desc;
  DevBuf fast_score, fast_cnt, fast_kp;
  DevBuf sift_rowf, sift_base, sift_kp, sift_desc;
  // pinned staging
  // pinned staging ring: the host fills slot k+1 while the copy out of slot k may still be
  // queued behind the previous step's kernels (an enqueue never waits for the device)
  struct StageSlot { void* p = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; } stage[4];
  int stage_i = 0;
  void* h_stage = nullptr;           // current slot
  cudaEvent_t stage_free = nullptr;  // current slot's event: record after the copy out of it
  cudaEvent_t done = nullptr;        // recorded after the last work queued through this lane
  int32_t* h_small = nullptr;        // pinned, N_SMALL ints (counts read-back)
  void* h_out = nullptr;             // pinned staging for the match lists on their way out
  size_t h_out_cap = 0;
  // state of the last enqueued batch
  int b_matcher = -1, b_nq = 0, b_pairs = 0, b_cap = 0;
  std::vector<int> b_tn;      // train rows of every pair of the last batch (bounds of trainIdx)
  cudaStream_t b_stream = nullptr;   // stream of the last lane-0 call (same stream: already ordered)
  bool b_stream_set = false;
  int s_H = 0, s_pairs = 0;
  int32_t* status = nullptr;  // device status words of the last batch (inside the table block)
  float* dbg = nullptr;  // debug: raw accumulator dump target of the next tcgen05 launch
};
#define N_SMALL 65536

struct ProfEvent {
  cudaEvent_t a, b;
  int kind;
};

struct slamb200_ctx {
  int device = 0;
  int profile = 0;
  unsigned profile_kinds = 0xFFFFFFFFu;   // kernel classes that get events while profile != 0
  std::mutex prof_mu;
  std::vector<ProfEvent> prof;
  cudaMemPool_t pool = nullptr;
  std::mutex mu;
  std::condition_variable cv;
  Lane lanes[N_LANES];  // lane 0 is the batch (enqueue/fetch) lane
  std::mutex batch_mu;
  std::mutex free_mu;
  int n_sm = 148;
  int next_lane = 1;
  int next_upload_lane = N_LANES - N_UPLOAD_LANES;
  cudaStream_t free_stream = nullptr;  // frees are stream-ordered here behind every lane's work
  // Descriptor-set slabs are recycled: a freed slab waits here with the event that marks the end
  // of all work that could still read it; the next upload of the same size makes its stream wait
  // on that event (on the device) and reuses the memory -- no allocator call in steady state.
  // Frees are batched: a freed slab first waits in `slab_pending` at no CUDA cost; a flush orders
  // the whole batch behind every lane's work with one event (shared, reference counted).
  struct SharedEv { cudaEvent_t ev; int refs; };
  struct CachedSlab { void* p; size_t bytes; SharedEv* ev; };
  std::vector<CachedSlab> slab_cache;
  std::vector<CachedSlab> slab_pending;
  size_t slab_cache_bytes = 0;
  std::vector<cudaEvent_t> event_pool;  // recycled cudaEventDisableTiming events (under free_mu)
  // Page-locked staging for slamb200_upload_desc_packed: a buffer is reusable once the event that
  // follows its prep kernel has completed.
  struct PinBuf { void* p; size_t cap; cudaEvent_t ev; bool busy; };
  std::vector<PinBuf*> pin_pool;
  std::mutex pin_mu;
  // Host threads that share the narrowing of one Mat (row slices), so that a few caller threads
  // can still use every core of the box; started on first use, size via slamb200_set_pack_threads.
  struct PackPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    std::atomic<bool> stop{false};
    std::atomic<int> queued{0}, sleepers{0};
    void start(int n) {
      for (int i = 0; i < n; i++)
        th.emplace_back([this] {
          for (;;) {
            std::function<void()> f;
            // a slice takes ~60 us and slices arrive in bursts: poll for a while before sleeping
            // (a futex wake-up costs about as much as the slice itself)
            const auto t_idle = std::chrono::steady_clock::now();
            while (queued.load(std::memory_order_acquire) == 0 && !stop &&
                   std::chrono::steady_clock::now() - t_idle < std::chrono::microseconds(300))
              __builtin_ia32_pause();
            {
              std::unique_lock<std::mutex> lk(mu);
              if (q.empty()) {
                sleepers++;
                cv.wait(lk, [this] { return stop.load() || !q.empty(); });
                sleepers--;
              }
              if (stop && q.empty()) return;
              f = std::move(q.front());
              q.pop_front();
              queued--;
            }
            f();
          }
        });
    }
    void submit(std::function<void()> f) {
      {
        std::lock_guard<std::mutex> lk(mu);
        q.push_back(std::move(f));
        queued++;
      }
      if (sleepers.load() > 0) cv.notify_one();
    }
    void shutdown() {
      {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
      }
      cv.notify_all();
      for (auto& t : th) t.join();
      th.clear();
    }
  } pack_pool;
  std::once_flag pack_once;
  int pack_threads = -1;  // -1: min(hardware threads, 16)
  bool pack_started = false;
  std::atomic<int> pool_busy{0};   // long-running narrowing workers of slamb200_match_batch_host on the pool
  int sub_batch = 0;  // debug: pairs per tcgen05 launch when pipelining against the rerank
  int use_tc = 1;  // debug switch (slamb200_dbg_set_tc): 0 routes exact-mode pairs to the fp32 kernel
  int fused_tail = -1;  // form of the match path's tail (debug switch slamb200_dbg_set_fused_tail):
                        // 0 = separate merge / rerank / finalize kernels, 1 = one tail kernel + compaction kernel,
                        // 2 = tail and compaction in one kernel (look-back), 3 = experimental: the tail inside the
                        // tcgen05 kernel where the batch allows it (else as 2) -- correct, measured slower;
                        // -1 (default) = by batch size: 2 up to 32 pairs (one launch and ~11 us less per call),
                        // 1 beyond (the look-back keeps blocks resident: +0.07 ms per 210-pair window)
  int use_tc_orb = 1;  // debug switch (slamb200_dbg_set_tc_orb): 0 routes ORB pairs to the XOR/POPC kernel
};

static int dev_alloc(slamb200_ctx* c, void** p, size_t bytes, cudaStream_t s) {
  if (bytes == 0) bytes = 16;
  CU(cudaMallocFromPoolAsync(p, bytes, c->pool, s));
  return SLAMB200_OK;
}

static int buf_reserve(slamb200_ctx* c, DevBuf& b, size_t bytes, cudaStream_t s) {
  if (bytes <= b.cap && b.p) return SLAMB200_OK;
  if (b.p) CU(cudaFreeAsync(b.p, s));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  int rc = dev_alloc(c, &b.p, want, s);
  if (rc) return rc;
  b.cap = want;
  return SLAMB200_OK;
}

static int stage_reserve(Lane& L, size_t bytes) {
  L.stage_i = (L.stage_i + 1) & 3;
  Lane::StageSlot& sl = L.stage[L.stage_i];
  if (!sl.ev) CU(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
  // the copy that last read this slot (four reservations ago) must have been consumed
  CU(cudaEventSynchronize(sl.ev));
  if (bytes > sl.cap) {
    if (sl.p) CU(cudaFreeHost(sl.p));
    sl.p = nullptr;
    sl.cap = 0;
    const size_t want = bytes * 2 + 4096;
    CU(cudaMallocHost(&sl.p, want));
    sl.cap = want;
  }
  L.h_stage = sl.p;
  L.stage_free = sl.ev;
  return SLAMB200_OK;
}

struct LaneGuard {
  slamb200_ctx* c;
  int idx;
  // Worker lanes are split in two classes so that a match (compute lane) is never enqueued behind
  // descriptor uploads another host thread queued for *later* work (upload lanes): with one shared
  // pool every thread's uploads and matches moved in a convoy and the PCIe link idled while all of
  // them matched.  Lane 0 is the batch (enqueue/fetch) lane.
  LaneGuard(slamb200_ctx* c_, bool upload = false) : c(c_), idx(-1) {
    const int lo = upload ? N_LANES - N_UPLOAD_LANES : 1;
    const int cnt = upload ? N_UPLOAD_LANES : N_LANES - N_UPLOAD_LANES - 1;
    int& next = upload ? c->next_upload_lane : c->next_lane;
    std::unique_lock<std::mutex> lk(c->mu);
    for (;;) {
      // round-robin so that back-to-back calls land on different streams
      for (int k = 1; k <= cnt; k++) {
        const int i = lo + (next - lo + k) % cnt;
        if (!c->lanes[i].busy) { idx = i; break; }
      }
      if (idx >= 0) break;
      c->cv.wait(lk);
    }
    next = idx;
    c->lanes[idx].busy = true;
  }
  ~LaneGuard() {
    {
      std::lock_guard<std::mutex> lk(c->mu);
      c->lanes[idx].busy = false;
    }
    c->cv.notify_all();
  }
  Lane& lane() { return c->lanes[idx]; }
};

// ---------------------------------------------------------------------------------------------
extern "C" int slamb200_version(void) { return SLAMB200_VERSION; }
extern "C" const char* slamb200_last_error(void) { return g_err; }
extern "C" int64_t slamb200_launch_count(const slamb200_ctx*) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int slamb200_init(int device, slamb200_ctx** out) {
  if (!out) return fail(SLAMB200_ERR_INVALID, "slamb200_init: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(SLAMB200_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                cudaGetErrorString(e));
  if (device < 0 || device >= n)
    return fail(SLAMB200_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(SLAMB200_ERR_CUDA, "device %d is sm_%d%d; libslamb200 is built for sm_100a only",
                device, prop.major, prop.minor);
  CU(cudaSetDevice(device));
  slamb200_ctx* c = new (std::nothrow) slamb200_ctx();
  if (!c) return fail(SLAMB200_ERR_NOMEM, "host allocation failed");
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  if (const char* e = getenv("SLAMB200_TAIL_FORM")) c->fused_tail = atoi(e);   // developer switch, see fused_tail
  cudaMemPoolProps pp;
  memset(&pp, 0, sizeof(pp));
  pp.allocType = cudaMemAllocationTypePinned;
  pp.handleTypes = cudaMemHandleTypeNone;
  pp.location.type = cudaMemLocationTypeDevice;
  pp.location.id = device;
  CU(cudaMemPoolCreate(&c->pool, &pp));
  uint64_t thr = UINT64_MAX;
  CU(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
  for (int i = 0; i < N_LANES; i++) {
    CU(cudaStreamCreateWithFlags(&c->lanes[i].stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->lanes[i].stream2, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->lanes[i].done, cudaEventDisableTiming));
    CU(cudaMallocHost((void**)&c->lanes[i].h_small, sizeof(int32_t) * N_SMALL));
  }
  CU(cudaStreamCreateWithFlags(&c->free_stream, cudaStreamNonBlocking));
  *out = c;
  return SLAMB200_OK;
}

int ctx_device(const slamb200_ctx* c) { return c->device; }
int ctx_grant_peer_access(slamb200_ctx* c, int peer_device) {
  cudaMemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = cudaMemLocationTypeDevice;
  acc.location.id = peer_device;
  acc.flags = cudaMemAccessFlagsProtReadWrite;
  if (cudaMemPoolSetAccess(c->pool, &acc, 1) != cudaSuccess) {
    cudaGetLastError();
    return SLAMB200_ERR_CUDA;
  }
  return SLAMB200_OK;
}

extern "C" int slamb200_synchronize(slamb200_ctx* c) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  for (int i = 0; i < N_LANES; i++) CU(cudaStreamSynchronize(c->lanes[i].stream));
  CU(cudaStreamSynchronize(c->free_stream));
  return SLAMB200_OK;
}

extern "C" int slamb200_shutdown(slamb200_ctx* c) {
  if (!c) return SLAMB200_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < N_LANES; i++) {
    Lane& L = c->lanes[i];
    DevBuf* bufs[] = {&L.pairs, &L.tcpairs, &L.tile_prefix, &L.work, &L.work_v0, &L.fb_list, &L.cand_g, &L.part, &L.cand, &L.knn_idx, &L.knn_dist, &L.flags,
                      &L.chunk_cnt, &L.out, &L.n_out, &L.err_flag, &L.scan, &L.seg_done, &L.status_fixed, &L.npts, &L.E, &L.counts,
                      &L.best, &L.mask, &L.m_off, &L.p1, &L.p2, &L.txy, &L.all_masks,
                      &L.orb_img, &L.orb_gray, &L.orb_rowf, &L.orb_blur, &L.orb_kp, &L.orb_desc,
                      &L.fast_score, &L.fast_cnt, &L.fast_kp, &L.sift_rowf, &L.sift_base, &L.sift_kp, &L.sift_desc};
    for (DevBuf* b : bufs)
      if (b->p) cudaFreeAsync(b->p, L.stream);
    cudaStreamSynchronize(L.stream);
    for (auto& sl : L.stage) {
      if (sl.p) cudaFreeHost(sl.p);
      if (sl.ev) cudaEventDestroy(sl.ev);
    }
    if (L.h_small) cudaFreeHost(L.h_small);
    if (L.h_out) cudaFreeHost(L.h_out);
    if (L.done) cudaEventDestroy(L.done);
    for (cudaEvent_t e : L.sub_ev)
      if (e) cudaEventDestroy(e);
    cudaStreamDestroy(L.stream2);
    cudaStreamDestroy(L.stream);
  }
  for (auto& e : c->slab_pending) cudaFree(e.p);
  for (auto& e : c->slab_cache) {
    cudaFree(e.p);
    if (--e.ev->refs == 0) {
      cudaEventDestroy(e.ev->ev);
      delete e.ev;
    }
  }
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  c->pack_pool.shutdown();
  for (auto* b : c->pin_pool) {
    cudaFreeHost(b->p);
    cudaEventDestroy(b->ev);
    delete b;
  }
  if (c->free_stream) cudaStreamDestroy(c->free_stream);
  if (c->pool) cudaMemPoolDestroy(c->pool);
  delete c;
  return SLAMB200_OK;
}

// ---- kernel timing ---------------------------------------------------------------------------
struct ProfScope {
  slamb200_ctx* c;
  cudaStream_t s;
  ProfEvent ev;
  bool on;
  ProfScope(slamb200_ctx* c_, cudaStream_t s_, int kind)
      : c(c_), s(s_), on(c_->profile != 0 && ((c_->profile_kinds >> kind) & 1u) != 0) {
    if (!on) return;
    ev.kind = kind;
    if (cudaEventCreate(&ev.a) != cudaSuccess || cudaEventCreate(&ev.b) != cudaSuccess) { on = false; return; }
    cudaEventRecord(ev.a, s);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(ev.b, s);
    std::lock_guard<std::mutex> lk(c->prof_mu);
    c->prof.push_back(ev);
  }
};

extern "C" int slamb200_profile_enable(slamb200_ctx* c, int on) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  c->profile = on ? 1 : 0;
  c->profile_kinds = 0xFFFFFFFFu;
  return SLAMB200_OK;
}

extern "C" int slamb200_profile_enable_kinds(slamb200_ctx* c, unsigned kinds) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  c->profile = kinds ? 1 : 0;
  c->profile_kinds = kinds;
  return SLAMB200_OK;
}

extern "C" int slamb200_profile_read(slamb200_ctx* c, double* ms, int64_t* launches) {
  if (!c || !ms || !launches) return fail(SLAMB200_ERR_INVALID, "profile_read: NULL argument");
  CU(cudaSetDevice(c->device));
  for (int k = 0; k < SLAMB200_K_COUNT; k++) { ms[k] = 0; launches[k] = 0; }
  std::lock_guard<std::mutex> lk(c->prof_mu);
  for (ProfEvent& e : c->prof) {
    float t = 0.f;
    if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) {
      ms[e.kind] += t;
      launches[e.kind] += 1;
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  c->prof.clear();
  return SLAMB200_OK;
}

// ---- descriptor sets ------------------------------------------------------------------------
static void* slab_from_cache(slamb200_ctx* c, size_t bytes, cudaStream_t s);
static cudaEvent_t event_get(slamb200_ctx* c);
static int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- page-locked staging pool -----------------------------------------------------------------
#define PIN_POOL_MAX 48
static slamb200_ctx::PinBuf* pin_acquire(slamb200_ctx* c, size_t bytes) {
  for (int spin = 0;; spin++) {
    {
      std::lock_guard<std::mutex> lk(c->pin_mu);
      for (auto* b : c->pin_pool)
        if (!b->busy && b->cap >= bytes && cudaEventQuery(b->ev) == cudaSuccess) {
          b->busy = true;
          return b;
        }
      cudaGetLastError();  // cudaErrorNotReady from the queries above
      // grow freely up to 16 buffers; beyond that only after waiting a little for one in flight
      // (cudaMallocHost costs milliseconds and stalls the device queue)
      if ((int)c->pin_pool.size() < 16 || ((int)c->pin_pool.size() < PIN_POOL_MAX && spin > 50)) {
        auto* b = new slamb200_ctx::PinBuf{nullptr, 0, nullptr, true};
        const size_t cap = (bytes + ((size_t)1 << 21) - 1) >> 21 << 21;  // 2 MiB granules
        if (cudaMallocHost(&b->p, cap) != cudaSuccess ||
            cudaEventCreateWithFlags(&b->ev, cudaEventDisableTiming) != cudaSuccess) {
          if (b->p) cudaFreeHost(b->p);
          delete b;
          cudaGetLastError();
          return nullptr;
        }
        b->cap = cap;
        c->pin_pool.push_back(b);
        return b;
      }
    }
    // every buffer is in flight: the GPU drains them in tens of microseconds
    std::this_thread::sleep_for(std::chrono::microseconds(20));
    if (spin > 500000) return nullptr;
  }
}
static void pin_release(slamb200_ctx* c, slamb200_ctx::PinBuf* b, cudaStream_t after) {
  if (after) cudaEventRecord(b->ev, after);
  std::lock_guard<std::mutex> lk(c->pin_mu);
  b->busy = false;
}

// one SIFT slab: f32 | bf16 | bf16lo | augq | augt | u8 | nrm2 | nrmf | flags  (256-byte aligned)
static size_t sift_slab_bytes(size_t np) { return np * (512 + 256 + 256 + 32 + 32 + 128 + 4 + 4) + 256; }
static void sift_slab_point(slamb200_desc* d) {
  const size_t np = (size_t)d->n_pad;
  const size_t o_f32 = 0, o_bf16 = o_f32 + np * 512, o_lo = o_bf16 + np * 256,
               o_augq = o_lo + np * 256, o_augt = o_augq + np * 32, o_u8 = o_augt + np * 32,
               o_nrm = o_u8 + np * 128, o_nrmf = o_nrm + np * 4, o_flags = o_nrmf + np * 4;
  char* base = (char*)d->slab;
  d->f32 = (float*)(base + o_f32);
  d->bf16 = (__nv_bfloat16*)(base + o_bf16);
  d->augq = (__nv_bfloat16*)(base + o_augq);
  d->augt = (__nv_bfloat16*)(base + o_augt);
  d->u8 = (uint8_t*)(base + o_u8);
  d->nrm2 = (int32_t*)(base + o_nrm);
  d->bf16lo = (__nv_bfloat16*)(base + o_lo);
  d->nrmf = (float*)(base + o_nrmf);
  d->flags = (int32_t*)(base + o_flags);
}

// one ORB slab: u8 rows | e4m3 0/1 bytes (tcgen05 operand) | augq | augt | flags (zero: the
// tensor-core kernels read them as "exact mode")
static size_t orb_slab_bytes(size_t np) { return np * (32 + 256 + 32 + 32) + 256; }
static void orb_slab_point(slamb200_desc* d) {
  const size_t np = (size_t)d->n_pad;
  char* base = (char*)d->slab;
  d->u8 = (uint8_t*)base;
  d->bf16 = (__nv_bfloat16*)(base + np * 32);            // the e4m3 bytes, 256 per row like a bf16 SIFT row
  d->augq = (__nv_bfloat16*)(base + np * (32 + 256));
  d->augt = (__nv_bfloat16*)(base + np * (32 + 256 + 32));
  d->flags = (int32_t*)(base + np * (32 + 256 + 32 + 32));
}

static int desc_create(slamb200_ctx* c, int kind, const void* rows, int n, size_t row_stride,
                       bool src_on_device, cudaStream_t producer, bool no_sync,
                       slamb200_desc** out, slamb200_ctx::PinBuf* packed = nullptr,
                       bool shared = false) {
  if (!c || !out) return fail(SLAMB200_ERR_INVALID, "upload_desc: NULL argument");
  *out = nullptr;
  if (kind != SLAMB200_DESC_F32X128 && kind != SLAMB200_DESC_U8X32)
    return fail(SLAMB200_ERR_KIND, "upload_desc: unknown descriptor kind %d", kind);
  if (n < 0 || (n > 0 && !rows)) return fail(SLAMB200_ERR_INVALID, "upload_desc: bad rows/n");
  const size_t row_bytes = kind == SLAMB200_DESC_F32X128 ? 512 : 32;
  if (row_stride == 0) row_stride = row_bytes;
  if (row_stride < row_bytes || (kind == SLAMB200_DESC_F32X128 && (row_stride % 4)))
    return fail(SLAMB200_ERR_INVALID, "upload_desc: row_stride %zu unsupported", row_stride);
  CU(cudaSetDevice(c->device));
  slamb200_desc* d =
      (slamb200_desc*)aligned_alloc(64, (sizeof(slamb200_desc) + 63) / 64 * 64);  // CUtensorMap: 64 B
  if (!d) return fail(SLAMB200_ERR_NOMEM, "host allocation failed");
  memset(d, 0, sizeof(slamb200_desc));
  d->kind = kind;
  d->n = n;
  d->n_pad = round_up(n > 0 ? n : 1, SLAMB200_TILE_PAD);
  d->host_exact = -2;
  d->device = c->device;
  LaneGuard g(c, /*upload=*/true);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  int rc = SLAMB200_OK;
  cudaEvent_t ev = nullptr;
#define DCU(call)                                                                   \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      rc = fail(SLAMB200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
      goto done;                                                                    \
    }                                                                               \
  } while (0)
  if (src_on_device) {
    // The bytes were produced on `producer`; NULL names the legacy default stream, which the
    // (non-blocking) upload lanes do not synchronise with implicitly, so it is recorded like any
    // other stream.
    DCU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    DCU(cudaEventRecord(ev, producer));
    DCU(cudaStreamWaitEvent(s, ev, 0));
  }
  {
    std::lock_guard<std::mutex> lk(c->free_mu);
    d->ready = event_get(c);
  }
  if (!d->ready) { rc = fail(SLAMB200_ERR_CUDA, "cudaEventCreate failed"); goto done; }
  if (kind == SLAMB200_DESC_U8X32) {
    d->slab_bytes = orb_slab_bytes(d->n_pad);
    if (shared) {
      DCU(cudaMalloc(&d->slab, d->slab_bytes));
      d->shared = 1;
    } else if (!(d->slab = slab_from_cache(c, d->slab_bytes, s))) {
      if ((rc = dev_alloc(c, &d->slab, d->slab_bytes, s))) goto done;
    }
    orb_slab_point(d);
    DCU(cudaMemsetAsync(d->flags, 0, 16, s));
    if (d->n_pad > n) DCU(cudaMemsetAsync(d->u8 + (size_t)n * 32, 0, (size_t)(d->n_pad - n) * 32, s));
    if (n > 0) {
      const cudaMemcpyKind k = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
      if (row_stride == 32) DCU(cudaMemcpyAsync(d->u8, rows, (size_t)n * 32, k, s));
      else DCU(cudaMemcpy2DAsync(d->u8, 32, rows, row_stride, 32, n, k, s));
    }
    // the bits spread to e4m3 bytes + popcount augmentation: operands of the tcgen05 kernel
    launch_orb_tc_prep(d->u8, n, d->n_pad, (uint8_t*)d->bf16, (uint8_t*)d->augq, (uint8_t*)d->augt, s);
    DCU(cudaGetLastError());
    if (tc_encode_tmaps(d->bf16, d->augq, d->augt, d->bf16, d->n_pad, d->tmaps) != 0) {
      rc = fail(SLAMB200_ERR_CUDA, "cuTensorMapEncodeTiled failed");
      goto done;
    }
  } else {
    const size_t total = sift_slab_bytes(d->n_pad);
    d->slab_bytes = total;
    if (shared) {
      // exportable over CUDA IPC: a plain allocation of its own (pool memory cannot be exported
      // with cudaIpcGetMemHandle)
      DCU(cudaMalloc(&d->slab, total));
      d->shared = 1;
    } else if (!(d->slab = slab_from_cache(c, total, s))) {
      if ((rc = dev_alloc(c, &d->slab, total, s))) goto done;
    }
    sift_slab_point(d);
    DCU(cudaMemsetAsync(d->flags, 0, 16, s));
    const float* prep_src = d->f32;
    size_t prep_stride = 128;
    if (packed) {
      // rows already narrowed to bytes in page-locked staging (verified exact on the host).  A
      // copy engine brings them into the slab's byte part and the prep kernel reads HBM: its
      // blocks then never sit on an SM waiting for PCIe while a match kernel wants every SM
      // (a window from host Mats: 0.90 of the host's narrowing floor against 0.78).
      // SLAMB200_UPLOAD_DMA=0: the prep kernel reads the staging over PCIe itself (one pass, no
      // copy-engine setup: the better choice for a lone upload).
      static const int dma = [] { const char* e = getenv("SLAMB200_UPLOAD_DMA"); return e ? atoi(e) : 1; }();
      const uint8_t* src8 = (const uint8_t*)packed->p;
      if (dma && n > 0) {
        DCU(cudaMemcpyAsync(d->u8, packed->p, (size_t)n * 128, cudaMemcpyHostToDevice, s));
        src8 = d->u8;
      }
      launch_sift_prep_u8(src8, n, d->n_pad, d->f32, d->bf16, d->augq, d->augt,
                          d->u8, d->nrm2, d->bf16lo, d->nrmf, d->flags, s);
      d->host_exact = 1;
      d->ready_seen = 0;
    } else if (n > 0) {
      // Page-locked, device-mapped host rows (the pipelined upload): the prep kernel reads them
      // straight over PCIe -- no staging copy, no per-copy setup cost, one pass over the data.
      const void* mapped = nullptr;
      // (the prep kernel reads float4: base and pitch must be 16-byte aligned, anything else is
      // copied first)
      if (!src_on_device && no_sync && row_stride % 16 == 0 && ((uintptr_t)rows & 15) == 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, rows) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
            pa.devicePointer != nullptr)
          mapped = pa.devicePointer;
        cudaGetLastError();
      }
      if (mapped) {
        prep_src = (const float*)mapped;
        prep_stride = row_stride / 4;
      } else {
        // the caller's rows land straight in the fp32 part (pitch removed by the copy itself)
        const cudaMemcpyKind k = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        if (row_stride == 512) DCU(cudaMemcpyAsync(d->f32, rows, (size_t)n * 512, k, s));
        else DCU(cudaMemcpy2DAsync(d->f32, 512, rows, row_stride, 512, n, k, s));
      }
    }
    if (!packed)
      launch_sift_prep(prep_src, prep_stride, n, d->n_pad, d->f32, d->bf16, d->augq, d->augt, d->u8,
                       d->nrm2, d->bf16lo, d->nrmf, d->flags, s);
    DCU(cudaGetLastError());
    if (tc_encode_tmaps(d->bf16, d->augq, d->augt, d->bf16lo, d->n_pad, d->tmaps) != 0) {
      rc = fail(SLAMB200_ERR_CUDA, "cuTensorMapEncodeTiled failed");
      goto done;
    }
  }
  DCU(cudaEventRecord(d->ready, s));
  DCU(cudaEventRecord(L.done, s));
  if (packed) DCU(cudaEventRecord(packed->ev, s));
  if (!src_on_device && !no_sync) {
    // the caller may reuse `rows` on return; the same synchronisation brings the exact-mode flag back
    const bool sift = kind == SLAMB200_DESC_F32X128;
    if (sift) DCU(cudaMemcpyAsync(L.h_small, d->flags, 4, cudaMemcpyDeviceToHost, s));
    DCU(cudaStreamSynchronize(s));
    if (sift) d->host_exact = L.h_small[0] == 0 ? 1 : 0;
    d->ready_seen = 1;
  }
done:
  if (ev) cudaEventDestroy(ev);
  if (rc != SLAMB200_OK) {
    slamb200_free_desc(c, d);
    return rc;
  }
  *out = d;
  return SLAMB200_OK;
#undef DCU
}

extern "C" int slamb200_upload_desc_packed(slamb200_ctx* c, int kind, const void* rows, int n,
                                           size_t row_stride, slamb200_desc** out);
// The plain upload IS the narrowing upload: same contract (rows consumed on return), same
// results, a quarter of the PCIe bytes for the Mats cv::SIFT produces, and page-locked staging
// instead of a pageable copy; anything that does not narrow losslessly takes the fp32 copy.
extern "C" int slamb200_upload_desc(slamb200_ctx* c, int kind, const void* rows, int n,
                                    size_t row_stride, slamb200_desc** out) {
  return slamb200_upload_desc_packed(c, kind, rows, n, row_stride, out);
}

extern "C" int slamb200_upload_desc_pinned(slamb200_ctx* c, int kind, const void* rows, int n,
                                           size_t row_stride, slamb200_desc** out) {
  return desc_create(c, kind, rows, n, row_stride, false, nullptr, true, out);
}

static void pack_pool_start(slamb200_ctx* c) {
  std::call_once(c->pack_once, [c] {
    int nthr = c->pack_threads;
    if (nthr < 0) {
      nthr = (int)std::thread::hardware_concurrency();
      nthr = nthr > 24 ? 24 : nthr;
    }
    c->pack_pool.start(nthr > 0 ? nthr : 0);
    c->pack_started = true;
  });
}

// The caller's rows are consumed before the call returns (narrowed into page-locked staging on the
// calling thread), the GPU work is only enqueued.
extern "C" int slamb200_upload_desc_packed(slamb200_ctx* c, int kind, const void* rows, int n,
                                           size_t row_stride, slamb200_desc** out) {
  if (!c || !out) return fail(SLAMB200_ERR_INVALID, "upload_desc: NULL argument");
  if (kind != SLAMB200_DESC_F32X128 || n <= 0 || !rows)
    return desc_create(c, kind, rows, n, row_stride, false, nullptr, false, out);
  if (row_stride == 0) row_stride = 512;
  if (row_stride < 512 || (row_stride % 4))
    return fail(SLAMB200_ERR_INVALID, "upload_desc: row_stride %zu unsupported", row_stride);
  CU(cudaSetDevice(c->device));
  slamb200_ctx::PinBuf* b = pin_acquire(c, (size_t)n * 128);
  if (!b) return desc_create(c, kind, rows, n, row_stride, false, nullptr, false, out);
  // narrow + verify: row slices on the pack pool, the last slice on the calling thread
  pack_pool_start(c);
  int exact = 1;
  {
    // (while a slamb200_match_batch_host call occupies the pool, narrow on the calling thread)
    const int pool = c->pool_busy.load() > 0 ? 0 : (int)c->pack_pool.th.size();
    int slices = n / 1024;                       // at least ~1k rows (0.5 MB) per slice
    slices = slices > 8 ? 8 : slices;
    slices = slices > pool + 1 ? pool + 1 : slices;
    if (slices <= 1) {
      exact = slamb200_host_pack_u8((const float*)rows, row_stride / 4, n, (uint8_t*)b->p);
    } else {
      std::atomic<int> pending(slices - 1), all_ok(1);
      std::mutex mu;
      std::condition_variable cv;
      const int per = (n + slices - 1) / slices;
      auto run = [&](int k) {
        const int r0 = k * per, r1 = r0 + per < n ? r0 + per : n;
        if (r1 > r0 && !slamb200_host_pack_u8((const float*)((const char*)rows + (size_t)r0 * row_stride),
                                              row_stride / 4, r1 - r0, (uint8_t*)b->p + (size_t)r0 * 128))
          all_ok.store(0);
      };
      for (int k = 0; k < slices - 1; k++)
        c->pack_pool.submit([&, k] {
          run(k);
          if (pending.fetch_sub(1) == 1) {
            std::lock_guard<std::mutex> lk(mu);
            cv.notify_one();
          }
        });
      run(slices - 1);
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return pending.load() == 0; });
      exact = all_ok.load();
    }
  }
  if (!exact) {
    pin_release(c, b, nullptr);  // not integer valued: the fp32 path
    return desc_create(c, kind, rows, n, row_stride, false, nullptr, false, out);
  }
  const int rc = desc_create(c, kind, rows, n, row_stride, false, nullptr, true, out, b);
  if (rc != SLAMB200_OK) cudaDeviceSynchronize();  // a prep kernel may still be reading the staging
  pin_release(c, b, nullptr);  // desc_create recorded b->ev behind the prep kernel
  return rc;
}

// ---- descriptor sets across processes (one process per GPU): CUDA IPC over NVLink --------------
static_assert(sizeof(slamb200_desc_ipc) == 128, "slamb200_desc_ipc is a 128-byte wire record");

extern "C" int slamb200_upload_desc_shared(slamb200_ctx* c, int kind, const void* rows, int n,
                                           size_t row_stride, slamb200_desc** out) {
  return desc_create(c, kind, rows, n, row_stride, false, nullptr, false, out, nullptr, true);
}

extern "C" int slamb200_desc_export(slamb200_ctx* c, const slamb200_desc* d, slamb200_desc_ipc* out) {
  if (!c || !d || !out) return fail(SLAMB200_ERR_INVALID, "desc_export: NULL argument");
  if (!d->shared) return fail(SLAMB200_ERR_INVALID, "desc_export: the set was not created with slamb200_upload_desc_shared");
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(d->ready));   // the importer has no event to wait on
  memset(out, 0, sizeof(*out));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, d->slab));
  static_assert(sizeof(h) == sizeof(out->handle), "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out->handle, &h, sizeof(h));
  out->kind = d->kind;
  out->n = d->n;
  out->n_pad = d->n_pad;
  out->device = c->device;
  out->slab_bytes = (unsigned long long)d->slab_bytes;
  out->exact = slamb200_desc_exact_mode(d);
  return SLAMB200_OK;
}

extern "C" int slamb200_desc_import(slamb200_ctx* c, const slamb200_desc_ipc* in, slamb200_desc** out) {
  if (!c || !in || !out) return fail(SLAMB200_ERR_INVALID, "desc_import: NULL argument");
  *out = nullptr;
  if (in->kind != SLAMB200_DESC_F32X128 && in->kind != SLAMB200_DESC_U8X32)
    return fail(SLAMB200_ERR_KIND, "desc_import: unknown descriptor kind %d", in->kind);
  if (in->n < 0 || in->n_pad < in->n || in->n_pad % SLAMB200_TILE_PAD)
    return fail(SLAMB200_ERR_INVALID, "desc_import: bad sizes");
  const size_t want = in->kind == SLAMB200_DESC_F32X128 ? sift_slab_bytes((size_t)in->n_pad) : orb_slab_bytes((size_t)in->n_pad);
  if (in->slab_bytes != want) return fail(SLAMB200_ERR_INVALID, "desc_import: slab size mismatch");
  CU(cudaSetDevice(c->device));
  slamb200_desc* d = (slamb200_desc*)aligned_alloc(64, (sizeof(slamb200_desc) + 63) / 64 * 64);
  if (!d) return fail(SLAMB200_ERR_NOMEM, "host allocation failed");
  memset(d, 0, sizeof(slamb200_desc));
  d->kind = in->kind;
  d->n = in->n;
  d->n_pad = in->n_pad;
  d->slab_bytes = want;
  d->imported = 1;
  d->device = in->device;
  d->host_exact = in->kind == SLAMB200_DESC_F32X128 ? (in->exact == 1 ? 1 : 0) : -2;
  d->ready_seen = 1;   // the exporter synchronised on its prep kernels
  cudaIpcMemHandle_t h;
  memcpy(&h, in->handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(&d->slab, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    free(d);
    return fail(SLAMB200_ERR_CUDA, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
  }
  if (d->kind == SLAMB200_DESC_U8X32) orb_slab_point(d);
  else sift_slab_point(d);
  {
    if (tc_encode_tmaps(d->bf16, d->augq, d->augt, d->kind == SLAMB200_DESC_U8X32 ? d->bf16 : d->bf16lo,
                        d->n_pad, d->tmaps) != 0) {
      cudaIpcCloseMemHandle(d->slab);
      free(d);
      return fail(SLAMB200_ERR_CUDA, "cuTensorMapEncodeTiled failed on the peer mapping");
    }
  }
  {
    std::lock_guard<std::mutex> lk(c->free_mu);
    d->ready = event_get(c);   // never recorded: queries as complete
  }
  *out = d;
  return SLAMB200_OK;
}

// A local copy of a (peer-mapped) set: one device-to-device transfer of the prepared slab over
// NVLink, no second prep pass.
extern "C" int slamb200_desc_localize(slamb200_ctx* c, const slamb200_desc* src, slamb200_desc** out) {
  if (!c || !src || !out) return fail(SLAMB200_ERR_INVALID, "desc_localize: NULL argument");
  *out = nullptr;
  CU(cudaSetDevice(c->device));
  slamb200_desc* d = (slamb200_desc*)aligned_alloc(64, (sizeof(slamb200_desc) + 63) / 64 * 64);
  if (!d) return fail(SLAMB200_ERR_NOMEM, "host allocation failed");
  memcpy(d, src, sizeof(slamb200_desc));
  d->imported = 0;
  d->shared = 0;
  d->device = c->device;
  d->slab = nullptr;
  d->ready = nullptr;
  d->ready_seen = 0;
  int rc = SLAMB200_OK;
  {
    LaneGuard g(c, /*upload=*/true);
    Lane& L = g.lane();
    cudaStream_t s = L.stream;
    if (!(d->slab = slab_from_cache(c, d->slab_bytes, s))) rc = dev_alloc(c, &d->slab, d->slab_bytes, s);
    if (rc == SLAMB200_OK) {
      if (!src->ready_seen && cudaStreamWaitEvent(s, src->ready, 0) != cudaSuccess) rc = SLAMB200_ERR_CUDA;
      // a set of another device of THIS process (device set) travels by a peer copy -- direct over
      // NVLink when the owner's pool grants this device access (ctx_grant_peer_access), staged by
      // the driver otherwise; a mapping of another process's set (CUDA IPC) is an ordinary
      // device pointer here
      cudaError_t ce = (!src->imported && src->device != c->device)
                           ? cudaMemcpyPeerAsync(d->slab, c->device, src->slab, src->device, d->slab_bytes, s)
                           : cudaMemcpyAsync(d->slab, src->slab, d->slab_bytes, cudaMemcpyDeviceToDevice, s);
      if (ce != cudaSuccess) rc = fail(SLAMB200_ERR_CUDA, "desc_localize: peer copy failed: %s", cudaGetErrorString(ce));
      {
        std::lock_guard<std::mutex> lk(c->free_mu);
        d->ready = event_get(c);
      }
      if (rc == SLAMB200_OK && (!d->ready || cudaEventRecord(d->ready, s) != cudaSuccess)) rc = SLAMB200_ERR_CUDA;
      cudaEventRecord(L.done, s);
    }
  }
  if (rc == SLAMB200_OK) {
    if (d->kind == SLAMB200_DESC_U8X32) orb_slab_point(d);
    else sift_slab_point(d);
    if (tc_encode_tmaps(d->bf16, d->augq, d->augt, d->kind == SLAMB200_DESC_U8X32 ? d->bf16 : d->bf16lo,
                        d->n_pad, d->tmaps) != 0)
      rc = fail(SLAMB200_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  }
  if (rc != SLAMB200_OK) {
    slamb200_free_desc(c, d);
    return rc;
  }
  *out = d;
  return SLAMB200_OK;
}

extern "C" int slamb200_set_pack_threads(slamb200_ctx* c, int n) {
  if (!c || n < 0 || n > 256) return fail(SLAMB200_ERR_INVALID, "set_pack_threads: bad argument");
  if (!c->pack_pool.th.empty() || c->pack_started)
    return (int)c->pack_pool.th.size() == n
               ? SLAMB200_OK
               : fail(SLAMB200_ERR_INVALID, "set_pack_threads: the pool is already running with %d threads",
                      (int)c->pack_pool.th.size());
  c->pack_threads = n;
  return SLAMB200_OK;
}

extern "C" int slamb200_upload_desc_device(slamb200_ctx* c, int kind, const void* rows, int n,
                                           size_t row_stride, void* stream, slamb200_desc** out) {
  return desc_create(c, kind, rows, n, row_stride, true, (cudaStream_t)stream, false, out);
}

static cudaEvent_t event_get(slamb200_ctx* c) {  // free_mu held
  if (!c->event_pool.empty()) {
    cudaEvent_t e = c->event_pool.back();
    c->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return e;
}
static void event_put(slamb200_ctx* c, cudaEvent_t e) {  // free_mu held
  if (!e) return;
  if (c->event_pool.size() < 1024) c->event_pool.push_back(e);
  else cudaEventDestroy(e);
}

// Orders the pending frees behind everything the context has queued so far without blocking the
// host: the free stream waits (on the device) for every lane's latest work -- a descriptor set's
// own prep kernel included, its lane's `done` event was recorded behind it -- then one event marks
// the whole batch reusable.  free_mu held.
#define SLAB_PENDING_MAX 16
static void flush_pending(slamb200_ctx* c) {
  if (c->slab_pending.empty()) return;
  for (int i = 0; i < N_LANES; i++) cudaStreamWaitEvent(c->free_stream, c->lanes[i].done, 0);
  auto* sev = new slamb200_ctx::SharedEv{event_get(c), 0};
  const bool ok = sev->ev && cudaEventRecord(sev->ev, c->free_stream) == cudaSuccess;
  const size_t kCacheLimit = (size_t)8 << 30;
  for (auto& e : c->slab_pending) {
    if (ok && e.bytes > 0 && c->slab_cache_bytes + e.bytes <= kCacheLimit) {
      sev->refs++;
      c->slab_cache.push_back({e.p, e.bytes, sev});
      c->slab_cache_bytes += e.bytes;
    } else {
      cudaFreeAsync(e.p, c->free_stream);
    }
  }
  c->slab_pending.clear();
  if (sev->refs == 0) {
    event_put(c, sev->ev);
    delete sev;
  }
}

static void free_behind_lanes(slamb200_ctx* c, void* p, size_t cache_bytes) {
  std::lock_guard<std::mutex> lk(c->free_mu);
  c->slab_pending.push_back({p, cache_bytes, nullptr});
  if (c->slab_pending.size() >= SLAB_PENDING_MAX) flush_pending(c);
}

// A recycled slab of exactly `bytes`, ordered behind its previous users on stream s; or nullptr.
static void* slab_from_cache(slamb200_ctx* c, size_t bytes, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(c->free_mu);
  for (int pass = 0; pass < 2; pass++) {
    for (size_t i = c->slab_cache.size(); i-- > 0;) {
      if (c->slab_cache[i].bytes == bytes) {
        slamb200_ctx::CachedSlab e = c->slab_cache[i];
        c->slab_cache.erase(c->slab_cache.begin() + (long)i);
        c->slab_cache_bytes -= bytes;
        cudaStreamWaitEvent(s, e.ev->ev, 0);
        if (--e.ev->refs == 0) {
          event_put(c, e.ev->ev);
          delete e.ev;
        }
        return e.p;
      }
    }
    // nothing ready: a slab of this size may be waiting in the pending batch
    bool have = false;
    for (auto& e : c->slab_pending) have = have || e.bytes == bytes;
    if (!have) break;
    flush_pending(c);
  }
  return nullptr;
}

extern "C" int slamb200_free_desc(slamb200_ctx* c, slamb200_desc* d) {
  if (!d) return SLAMB200_OK;
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  cudaSetDevice(c->device);
  // Work this context queued that may still read the set drains first (stream-ordered, the host
  // does not wait).  Work the caller queued on its own streams through the *_enqueue entry
  // points must have been recorded by them (it is: every enqueue records the lane's event).
  if (d->slab && (d->imported || d->shared)) {
    // peer mappings and exportable allocations are released synchronously, behind everything the
    // context has queued (rare: once per frame of a sharded window)
    for (int i = 0; i < N_LANES; i++) cudaStreamSynchronize(c->lanes[i].stream);
    if (d->imported) cudaIpcCloseMemHandle(d->slab);
    else cudaFree(d->slab);
  } else if (d->slab) {
    free_behind_lanes(c, d->slab, d->slab_bytes);
  }
  if (d->ready) {
    std::lock_guard<std::mutex> lk(c->free_mu);
    event_put(c, d->ready);
  }
  free(d);
  return SLAMB200_OK;
}

extern "C" int slamb200_desc_rows(const slamb200_desc* d) { return d ? d->n : -1; }
extern "C" int slamb200_desc_kind(const slamb200_desc* d) { return d ? d->kind : -1; }
extern "C" int slamb200_desc_exact_mode(const slamb200_desc* d) {
  if (!d) return -1;
  if (d->kind != SLAMB200_DESC_F32X128) return -1;
  slamb200_desc* m = const_cast<slamb200_desc*>(d);
  if (m->host_exact == -2) {
    int32_t f = 1;
    if (cudaEventSynchronize(d->ready) != cudaSuccess) return -1;
    if (cudaMemcpy(&f, d->flags, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    m->host_exact = f == 0 ? 1 : 0;
  }
  return m->host_exact;
}

// ---- matching ---------------------------------------------------------------------------------
static int pick_splits(int q_blocks, int n_pairs, int t_max, int min_chunk) {
  // enough (q-block, split, pair) work items for ~4 waves of 148 SMs
  long long items = (long long)q_blocks * n_pairs;
  int want = (int)((148LL * 4 + items - 1) / (items > 0 ? items : 1));
  int max_split = t_max / min_chunk;
  if (max_split < 1) max_split = 1;
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  return want;
}

static int check_matcher(int matcher, const slamb200_desc* q, const slamb200_desc* const* t,
                         int n_pairs) {
  if (matcher != SLAMB200_SIFT_BF && matcher != SLAMB200_SIFT_FLANN && matcher != SLAMB200_ORB_BF &&
      matcher != SLAMB200_SIFT_BF_L1)
    return fail(SLAMB200_ERR_MATCHER,
                "matcher type %d is not 0 (SIFT_BF), 1 (SIFT_FLANN), 2 (ORB_BF) or 3 (SIFT_BF_L1)", matcher);
  const int kind = matcher == SLAMB200_ORB_BF ? SLAMB200_DESC_U8X32 : SLAMB200_DESC_F32X128;
  if (!q) return fail(SLAMB200_ERR_INVALID, "query descriptor set is NULL");
  if (q->kind != kind) return fail(SLAMB200_ERR_KIND, "query descriptor kind does not fit the matcher");
  for (int p = 0; p < n_pairs; p++) {
    if (!t || !t[p]) return fail(SLAMB200_ERR_INVALID, "train descriptor set %d is NULL", p);
    if (t[p]->kind != kind)
      return fail(SLAMB200_ERR_KIND, "train descriptor kind of pair %d does not fit the matcher", p);
  }
  return SLAMB200_OK;
}

// Queues one (query x n_pairs trains) batch on stream s using lane L's scratch.  Results:
// L.knn_idx / L.knn_dist [P][nq][2], L.out [P][cap] dmatch, L.n_out [P].
//
// Everything the kernels need from the host travels in ONE pinned block and one copy:
//   [ PairArgs[P] | TcPair[P] | tile_prefix[P+1] | status words (zeroed) ]
// so a step costs one H2D copy plus the kernels themselves (no memsets, no per-table copies).
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static_assert(1024 + SIFT_GEN_FB_ITEMS <= 8192 / 4, "status area too small");
constexpr size_t FB_PART_BYTES = sizeof(unsigned long long) * 2 * SIFT_GEN_FB_ITEMS;
#define STATUS_BYTES 8192  // [0] self-check flag, [1+k] work-list counts, [1002] fallback rows,
                           // [1024 + i] finished segments of fallback row i (split scan)

static int enqueue_batch(slamb200_ctx* c, Lane& L, cudaStream_t s, int matcher,
                         const slamb200_desc* q, const slamb200_desc* const* trains, int n_pairs,
                         double ratio, bool want_knn = false) {
  const int nq = q->n;
  const int cap = nq > 0 ? nq : 1;
  // a batch whose counts could not be fetched (fetch_batch reads them through h_small) is
  // rejected before any GPU work is queued
  if (n_pairs + 1 > N_SMALL) return fail(SLAMB200_ERR_INVALID, "too many pairs in one batch (%d)", n_pairs);
  L.b_matcher = matcher; L.b_nq = nq; L.b_pairs = n_pairs; L.b_cap = cap;
  L.b_tn.resize((size_t)n_pairs);
  for (int p = 0; p < n_pairs; p++) L.b_tn[(size_t)p] = trains[p]->n;
  if (n_pairs == 0) return SLAMB200_OK;
  int rc;
  const bool orb = matcher == SLAMB200_ORB_BF;
  const bool l1 = matcher == SLAMB200_SIFT_BF_L1;
  int t_max = 0;
  for (int p = 0; p < n_pairs; p++) t_max = trains[p]->n > t_max ? trains[p]->n : t_max;
  if (orb && t_max >= (1 << 22))
    return fail(SLAMB200_ERR_INVALID, "ORB train set of %d rows exceeds the 2^22-row key range", t_max);
  // the candidate records pack two 16-bit group indices: train sets beyond 524k rows take the
  // exact fp32 kernel instead
  // ORB takes the same tcgen05 candidate kernel on e4m3 0/1 bytes (Hamming = squared L2 of the
  // bit vectors); the XOR/POPC kernel remains for huge train sets and behind the debug switch
  const bool tc = !l1 && (orb ? c->use_tc_orb : c->use_tc) && nq > 0 && t_max <= 65535 * 8;
  const int q_blocks = orb ? (nq + 255) / 256 : (nq + 15) / 16;
  const int n_split = pick_splits(q_blocks > 0 ? q_blocks : 1, n_pairs, t_max, orb ? 1024 : 512);

  // the exact fp32 kernel is not even launched when every set is known (on the host) to be in
  // exact mode; sets whose flag has not been read back yet leave the decision to the device
  bool all_exact_known = orb || q->host_exact == 1;
  for (int p = 0; p < n_pairs && all_exact_known && !orb; p++) all_exact_known = trains[p]->host_exact == 1;
  // Small batches of known kind on the one-kernel tail (the per-pair drop-in call): the pair table
  // travels in the kernel parameters and the status words live in a buffer that stays zero, so the
  // call queues no copy at all in front of its two kernels.
  const int tail_form = c->fused_tail >= 0 ? c->fused_tail : (n_pairs <= 32 ? 2 : 1);
  const bool inline_tables = tc && n_pairs <= tc_inline_max() && all_exact_known && !want_knn &&
                             tail_form >= 2 && !(c->sub_batch > 0 && c->sub_batch < n_pairs);

  // ---- host tables -----------------------------------------------------------------------
  const size_t off_pairs = 0;
  const size_t off_tc = align_up(sizeof(PairArgs) * (size_t)n_pairs, 64);
  const size_t off_pre = off_tc + (tc ? sizeof(TcPair) * (size_t)n_pairs : 0);
  const size_t off_status = align_up(off_pre + (tc ? sizeof(int32_t) * (size_t)(n_pairs + 1) : 0), 64);
  const size_t table_bytes = off_status + STATUS_BYTES;
  alignas(64) char inline_buf[4096];
  static_assert(sizeof(inline_buf) >= 64 + (sizeof(PairArgs) + sizeof(TcPair) + 4) * 4 + 128, "inline table buffer");
  char* hb = inline_buf;
  if (!inline_tables) {
    if ((rc = stage_reserve(L, table_bytes))) return rc;
    hb = (char*)L.h_stage;
    memset(hb + off_status, 0, STATUS_BYTES);
  }
  PairArgs* hp = (PairArgs*)(hb + off_pairs);
  TcPair* tp = (TcPair*)(hb + off_tc);
  int32_t* pre = (int32_t*)(hb + off_pre);
  for (int p = 0; p < n_pairs; p++) {
    hp[p].t_rows = orb ? (const void*)trains[p]->u8 : (const void*)trains[p]->f32;
    hp[p].t_flags = orb ? nullptr : trains[p]->flags;
    hp[p].t_u8 = trains[p]->u8;
    hp[p].t_n = trains[p]->n;
    hp[p].t_pad = trains[p]->n_pad;
  }
  // tcgen05 path geometry: one query block of 256 rows per CTA pair, 256-column train tiles
  const int n_rb = (nq + 255) / 256;
  int n_cta = 1, n_slots = 2, wide = 0;
  int min_cb = 1 << 30;   // column tiles of the smallest train set of the batch
  bool holes = false;
  if (tc) {
    long long total = 0;
    const int max_pairs = c->n_sm / 2;  // one CTA pair per two SMs
    int max_tiles = 0;
    for (int p = 0; p < n_pairs; p++) {
      const int n_cb = (trains[p]->n + 255) / 256;
      max_tiles = n_cb * n_rb > max_tiles ? n_cb * n_rb : max_tiles;
    }
    n_cta = max_tiles < max_pairs ? (max_tiles > 0 ? max_tiles : 1) : max_pairs;
    // the kernels' share arithmetic multiplies a tile index by the number of CTA pairs
    wide = (unsigned long long)max_tiles * (unsigned long long)n_cta >= (1ull << 32) ? 1 : 0;
    for (int p = 0; p < n_pairs; p++) {
      memcpy(tp[p].tmap, trains[p]->tmaps, 128);             // main (hi)
      memcpy(tp[p].tmap + 128, trains[p]->tmaps + 256, 128);  // aug, train role
      memcpy(tp[p].tmap + 256, trains[p]->tmaps + 384, 128);  // lo
      tp[p].t_f32 = trains[p]->f32;
      tp[p].t_nrmf = trains[p]->nrmf;
      tp[p].t_u8 = trains[p]->u8;
      tp[p].t_nrm2 = trains[p]->nrm2;
      tp[p].t_flags = trains[p]->flags;
      tp[p].t_n = trains[p]->n;
      tp[p].t_pad = trains[p]->n_pad;
      const int n_cb = (trains[p]->n + 255) / 256;
      min_cb = n_cb < min_cb ? n_cb : min_cb;
      pre[p] = (int32_t)total;
      total += (long long)n_cb * n_rb;
      // every CTA pair takes a contiguous share of this frame pair's tiles: how many shares can
      // cut one row block of n_cb tiles
      if (n_cb > 0) {
        const int slots = tc_slots(n_cb, n_cb * n_rb, n_cta);
        n_slots = slots > n_slots ? slots : n_slots;
        // fewer tiles than CTA pairs: empty shares leave holes between written slot records
        if (n_cb * n_rb < n_cta) holes = true;
      }
    }
    pre[n_pairs] = (int32_t)total;
    if (total > 0x7fffffffLL) return fail(SLAMB200_ERR_INVALID, "batch too large (tile count)");
  }

  // ---- device scratch ----------------------------------------------------------------------
  const size_t rows = (size_t)n_pairs * cap;
  const size_t cand_bytes = tc ? sizeof(uint4) * (size_t)n_pairs * n_slots * (size_t)n_rb * 256 : 0;
  if (inline_tables) {
    if (!L.status_fixed.p) {
      if ((rc = buf_reserve(c, L.status_fixed, STATUS_BYTES, s))) return rc;
      CU(cudaMemsetAsync(L.status_fixed.p, 0, L.status_fixed.cap, s));
    }
  } else if ((rc = buf_reserve(c, L.pairs, table_bytes, s))) {
    return rc;
  }
  if ((rc = buf_reserve(c, L.part, sizeof(uint4) * rows * n_split, s))) return rc;
  if ((rc = buf_reserve(c, L.knn_idx, sizeof(int32_t) * rows * 2, s))) return rc;
  if ((rc = buf_reserve(c, L.knn_dist, sizeof(float) * rows * 2, s))) return rc;
  if ((rc = buf_reserve(c, L.flags, rows, s))) return rc;
  if ((rc = buf_reserve(c, L.chunk_cnt, sizeof(int32_t) * (size_t)n_pairs * (finalize_chunks(cap) + 1), s))) return rc;
  if ((rc = buf_reserve(c, L.out, sizeof(slamb200_dmatch) * rows, s))) return rc;
  if ((rc = buf_reserve(c, L.n_out, sizeof(int32_t) * (size_t)n_pairs, s))) return rc;
  if (tc) {
    if ((rc = buf_reserve(c, L.cand, cand_bytes, s))) return rc;
    if ((rc = buf_reserve(c, L.work, sizeof(uint4) * rows, s))) return rc;
    if ((rc = buf_reserve(c, L.work_v0, sizeof(float2) * rows, s))) return rc;
    // head: partial keys of the split fallback scan, then the list of fallback rows
    if ((rc = buf_reserve(c, L.fb_list, FB_PART_BYTES + sizeof(uint2) * rows, s))) return rc;
  }
  char* db = (char*)L.pairs.p;
  const PairArgs* d_pairs = inline_tables ? nullptr : (const PairArgs*)(db + off_pairs);
  const TcPair* d_tc = inline_tables ? nullptr : (const TcPair*)(db + off_tc);
  const int32_t* d_pre = inline_tables ? nullptr : (const int32_t*)(db + off_pre);
  int32_t* d_status = inline_tables ? (int32_t*)L.status_fixed.p : (int32_t*)(db + off_status);
  const TcPair* inl_tc = inline_tables ? tp : nullptr;
  const int32_t* inl_pre = inline_tables ? pre : nullptr;
  L.status = d_status;
  if (!inline_tables) {
    CU(cudaMemcpyAsync(db, hb, table_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaEventRecord(L.stage_free, s));
  }

  // the descriptor sets must have finished their prep kernels
  {
    auto wait_ready = [&](const slamb200_desc* d) -> cudaError_t {
      slamb200_desc* m = const_cast<slamb200_desc*>(d);
      if (m->ready_seen) return cudaSuccess;
      if (cudaEventQuery(d->ready) == cudaSuccess) { m->ready_seen = 1; return cudaSuccess; }
      cudaGetLastError();
      return cudaStreamWaitEvent(s, d->ready, 0);
    };
    CU(wait_ready(q));
    for (int p = 0; p < n_pairs; p++) CU(wait_ready(trains[p]));
  }
  if (orb && !tc) {
    ProfScope ps(c, s, SLAMB200_K_ORB);
    launch_orb_knn2(q->u8, nq, d_pairs, n_pairs, n_split, (uint4*)L.part.p, s);
  } else {
    // Without the tensor-core path (debug switch, or train sets beyond 524k rows) the exact fp32
    // kernel does every pair.  With it: integer-valued pairs (what cv::SIFT emits) take the exact-
    // mode tcgen05 kernel + dp4a rerank; general-float pairs take the two-term-split tcgen05 kernel
    // + certified fp32 rerank (+ exact fallback rows).  The kernels pick their pairs from the
    // flags on the device, so no host synchronisation is needed to route.
    if (l1) {
      // NORM_L1 (the OpenCV-CUDA build's useFM-SIFT-BF): integer-valued pairs on the u8 copy with
      // the byte-wise SAD instruction, everything else in fp32 in cv::BFMatcher(NORM_L1)'s order;
      // both kernels route on the device flags
      const bool u8_ok = t_max <= sift_l1_max_train_rows();
      if (u8_ok) {
        ProfScope ps(c, s, SLAMB200_K_SIFT_L1);
        launch_sift_l1_u8(q->u8, q->flags, nq, d_pairs, n_pairs, n_split, (uint4*)L.part.p, s);
      }
      if (!u8_ok || !all_exact_known) {
        ProfScope ps(c, s, SLAMB200_K_SIFT_EXACT);
        launch_sift_exact_knn2(q->f32, q->flags, nq, d_pairs, n_pairs, n_split, (uint4*)L.part.p,
                               u8_ok ? 0 : 1, 1, s);
      }
    } else if (!tc) {
      ProfScope ps(c, s, SLAMB200_K_SIFT_EXACT);
      launch_sift_exact_knn2(q->f32, q->flags, nq, d_pairs, n_pairs, n_split, (uint4*)L.part.p, 1, 0, s);
    }
    if (tc) {
      // Slot records are normally all written by the kernel (the merge pass computes which ones
      // exist); only the hole case needs them cleared (small or ragged batches).
      if (holes) CU(cudaMemsetAsync(L.cand.p, 0xFF, cand_bytes, s));
      alignas(64) unsigned char qmaps[384];
      memcpy(qmaps, q->tmaps, 256);              // main (hi), aug (query role)
      memcpy(qmaps + 256, q->tmaps + 384, 128);  // lo
      if (!all_exact_known) {
        // general-float pairs (skipped entirely when every set is known to be integer-valued);
        // their slot records are two 16-byte words wide and live in their own buffer
        if ((rc = buf_reserve(c, L.cand_g, 2 * cand_bytes, s))) return rc;
        if (holes) CU(cudaMemsetAsync(L.cand_g.p, 0xFF, 2 * cand_bytes, s));
        // Sub-batches of general-float pairs: the tcgen05 kernel of sub-batch k+1 on `s` runs
        // beside the certified rerank of sub-batch k on the lane's second stream.  The GEN kernel
        // is bound by the tensor pipe (25 MMAs per tile) and leaves registers, 14 KB of shared
        // memory and most issue slots of every SM free; the rerank is bound by L2 reads -- unlike
        // the exact-mode pair of kernels (below) the two do not compete.
        static const int gen_sub_env = [] { const char* e = getenv("SLAMB200_GEN_SUB"); return e ? atoi(e) : 4; }();
        const int gsub = (gen_sub_env > 0 && n_pairs >= 2 * gen_sub_env) ? gen_sub_env : n_pairs;
        const int g_nsub = (n_pairs + gsub - 1) / gsub;
        if ((int)L.sub_ev.size() < g_nsub + 1) {
          const size_t old_n = L.sub_ev.size();
          L.sub_ev.resize(g_nsub + 1, nullptr);
          for (size_t i = old_n; i < L.sub_ev.size(); i++)
            CU(cudaEventCreateWithFlags(&L.sub_ev[i], cudaEventDisableTiming));
        }
        cudaStream_t gs2 = g_nsub > 1 ? L.stream2 : s;
        const size_t cand_per_pair_g = 2 * (size_t)n_slots * n_rb * 256;
        for (int k = 0; k < g_nsub; k++) {
          const int p0 = k * gsub;
          const int np = n_pairs - p0 < gsub ? n_pairs - p0 : gsub;
          const int tiles_k = pre[p0 + np] - pre[p0];
          uint4* cand_gk = (uint4*)L.cand_g.p + (size_t)p0 * cand_per_pair_g;
          uint4* part_k = (uint4*)L.part.p + (size_t)p0 * n_split * nq;
          int grc;
          {
            ProfScope ps(c, s, SLAMB200_K_SIFT_TC_GEN);
            grc = launch_sift_tc_candidates(qmaps, q->flags, nq, d_tc + p0, d_pre + p0, np, tiles_k, n_cta,
                                            n_slots, cand_gk, d_status, nullptr, 1, s, 0,
                                            q->host_exact == 0 ? 1 : 0, wide);
          }
          if (grc != 0) return fail(SLAMB200_ERR_CUDA, "tcgen05 kernel configuration failed");
          if (gs2 != s) {
            CU(cudaEventRecord(L.sub_ev[k], s));
            CU(cudaStreamWaitEvent(gs2, L.sub_ev[k], 0));
          }
          // the fallback bookkeeping (row count, per-row segment counters) starts from zero for
          // every rerank launch: the table copy cleared it for the first one
          if (k > 0)
            CU(cudaMemsetAsync(d_status + 1002, 0, sizeof(int32_t) * (size_t)(1024 - 1002 + SIFT_GEN_FB_ITEMS), gs2));
          ProfScope ps(c, gs2, SLAMB200_K_SIFT_GEN_RERANK);
          launch_sift_gen_rerank(q->flags, q->f32, q->nrmf, nq, d_tc + p0, d_pre + p0, np, n_cta, n_slots, n_split,
                                 (const uint4*)cand_gk, part_k,
                                 (uint2*)((char*)L.fb_list.p + FB_PART_BYTES), d_status + 1002,
                                 (unsigned long long*)L.fb_list.p, d_status + 1024, gs2, wide);
        }
        if (gs2 != s) {
          CU(cudaEventRecord(L.sub_ev[g_nsub], gs2));
          CU(cudaStreamWaitEvent(s, L.sub_ev[g_nsub], 0));
        }
      }
      // Sub-batch pipeline (debug knob, default one sub-batch): the tcgen05 kernel of sub-batch
      // k+1 on `s` against the rerank / finalize kernels of sub-batch k on the lane's second
      // stream.  Measured: no gain (both sides contend for the same SM issue slots).
      const int sub = (c->sub_batch > 0 && c->sub_batch < n_pairs) ? c->sub_batch : n_pairs;
      const int n_sub = (n_pairs + sub - 1) / sub;
      if ((int)L.sub_ev.size() < n_sub + 1) {
        const size_t old_n = L.sub_ev.size();
        L.sub_ev.resize(n_sub + 1, nullptr);
        for (size_t i = old_n; i < L.sub_ev.size(); i++)
          CU(cudaEventCreateWithFlags(&L.sub_ev[i], cudaEventDisableTiming));
      }
      cudaStream_t s2 = n_sub > 1 ? L.stream2 : s;
      const size_t cand_per_pair = (size_t)n_slots * n_rb * 256;
      const int chunks = finalize_chunks(cap);
      // a query set known (on the host) to be general floats makes every pair a general-float
      // pair: the exact-mode kernels would find nothing to do
      const bool none_exact = !orb && q->host_exact == 0;
      for (int k = 0; k < n_sub; k++) {
        const int p0 = k * sub;
        const int np = n_pairs - p0 < sub ? n_pairs - p0 : sub;
        const TcPair* tcp = d_tc + p0;
        const int32_t* pre_k = d_pre + p0;
        const int tiles_k = pre[p0 + np] - pre[p0];
        uint4* cand_k = (uint4*)L.cand.p + (size_t)p0 * cand_per_pair;
        uint4* part_k = (uint4*)L.part.p + (size_t)p0 * n_split * nq;
        // look-back words of the ordered compaction (both one-kernel forms of the tail): one per
        // (pair, row block, CTA of the pair), cleared only when (re)allocated -- the epoch tells
        // the launches apart
        if (tail_form >= 2 && !want_knn) {
          const size_t scan_bytes = sizeof(unsigned long long) * (size_t)n_pairs * chunks * 2;
          if (scan_bytes > L.scan.cap || !L.scan.p || L.scan_epoch >= (1u << 30) - 2) {
            if ((rc = buf_reserve(c, L.scan, scan_bytes, s))) return rc;
            CU(cudaMemsetAsync(L.scan.p, 0, L.scan.cap, s));
            L.scan_epoch = 0;
          }
        }
        // The whole match path in one kernel: every pair integer-valued (or ORB) and known to be so
        // on the host, no empty shares, and row blocks long enough (>= 12 column tiles) for the two
        // tail warps of a CTA to keep up with its epilogue.
        if (tail_form >= 3 && !want_knn && !none_exact && all_exact_known && !holes && min_cb >= 12 &&
            n_sub == 1) {
          const size_t done_bytes = sizeof(int32_t) * (size_t)n_pairs * chunks * 2;
          if (done_bytes > L.seg_done.cap || !L.seg_done.p) {
            if ((rc = buf_reserve(c, L.seg_done, done_bytes, s))) return rc;
            CU(cudaMemsetAsync(L.seg_done.p, 0, L.seg_done.cap, s));
          }
          ProfScope ps(c, s, orb ? SLAMB200_K_ORB : SLAMB200_K_SIFT_TC);
          const int mrc = launch_sift_tc_match(qmaps, q->flags, q->u8, q->nrm2, nq, tcp, pre_k, np, tiles_k, n_cta,
                                               n_slots, cand_k, d_status, ratio, orb ? 1 : 0, wide,
                                               (unsigned long long*)L.scan.p, ++L.scan_epoch,
                                               (int32_t*)L.seg_done.p, (slamb200_dmatch*)L.out.p, cap,
                                               (int32_t*)L.n_out.p, s, inl_tc, inl_pre);
          if (mrc != 0) return fail(SLAMB200_ERR_CUDA, "tcgen05 kernel configuration failed");
          continue;
        }
        int trc = 0;
        if (!none_exact) {
          ProfScope ps(c, s, orb ? SLAMB200_K_ORB : SLAMB200_K_SIFT_TC);
          trc = launch_sift_tc_candidates(qmaps, q->flags, nq, tcp, pre_k, np, tiles_k, n_cta,
                                          n_slots, cand_k, d_status, k == 0 ? L.dbg : nullptr, 0, s,
                                          orb ? 1 : 0, all_exact_known ? 1 : 0, wide, inl_tc, inl_pre);
        }
        if (trc != 0) return fail(SLAMB200_ERR_CUDA, "tcgen05 kernel configuration failed");
        if (s2 != s) {
          CU(cudaEventRecord(L.sub_ev[k], s));
          CU(cudaStreamWaitEvent(s2, L.sub_ev[k], 0));
        }
        if (tail_form >= 2 && !want_knn) {
          // match output: merge, best-group rerank, ratio test AND the ordered compaction in one
          // kernel (look-back over the 256-row blocks of a pair); general-float pairs of the batch
          // are finalized by the same kernel from their records
          ProfScope ps(c, s2, orb ? SLAMB200_K_ORB : SLAMB200_K_SIFT_RERANK);
          launch_tc_tail_compact(q->flags, q->u8, q->nrm2, nq, tcp, pre_k, np, n_cta, n_slots, n_split,
                                 cand_k, part_k, d_status, ratio, orb ? 1 : 0,
                                 (unsigned long long*)L.scan.p + (size_t)p0 * chunks, ++L.scan_epoch,
                                 (slamb200_dmatch*)L.out.p + (size_t)p0 * cap, cap, (int32_t*)L.n_out.p + p0, s2,
                                 wide, inl_tc, inl_pre);
          continue;
        }
        if (tail_form >= 1 && !want_knn) {
          // match output: merge, best-group rerank and ratio test in one kernel, then compaction
          // (general-float pairs of the batch are finalized by the same kernel from their records)
          {
            ProfScope ps(c, s2, orb ? SLAMB200_K_ORB : SLAMB200_K_SIFT_RERANK);
            launch_tc_tail_fused(q->flags, q->u8, q->nrm2, nq, tcp, pre_k, np, n_cta, n_slots, n_split,
                                 cand_k, part_k, d_status, ratio, orb ? 1 : 0,
                                 (int32_t*)L.knn_idx.p + (size_t)p0 * nq * 2,
                                 (float*)L.knn_dist.p + (size_t)p0 * nq * 2,
                                 (uint8_t*)L.flags.p + (size_t)p0 * nq,
                                 (int32_t*)L.chunk_cnt.p + (size_t)p0 * chunks, s2, wide);
          }
          ProfScope psf(c, s2, SLAMB200_K_FINALIZE);
          launch_compact(nq, np, (const int32_t*)L.knn_idx.p + (size_t)p0 * nq * 2,
                         (const float*)L.knn_dist.p + (size_t)p0 * nq * 2,
                         (const uint8_t*)L.flags.p + (size_t)p0 * nq,
                         (const int32_t*)L.chunk_cnt.p + (size_t)p0 * chunks,
                         (slamb200_dmatch*)L.out.p + (size_t)p0 * cap, cap, (int32_t*)L.n_out.p + p0, s2);
          continue;
        }
        if (!none_exact) {
          ProfScope ps(c, s2, orb ? SLAMB200_K_ORB : SLAMB200_K_SIFT_RERANK);
          launch_sift_rerank(q->flags, q->u8, q->nrm2, nq, tcp, pre_k, np, n_cta, n_slots, n_split,
                             cand_k, part_k, (uint4*)L.work.p + (size_t)p0 * nq,
                             (float2*)L.work_v0.p + (size_t)p0 * nq, d_status + 1 + (k < 1000 ? k : 1000),
                             d_status, want_knn ? 0 : 1, ratio, s2, orb ? 1 : 0, wide);
        }
        ProfScope psf(c, s2, SLAMB200_K_FINALIZE);
        launch_finalize(part_k, nq, d_pairs + p0, np, n_split, orb ? 1 : 0, ratio,
                        (int32_t*)L.knn_idx.p + (size_t)p0 * nq * 2,
                        (float*)L.knn_dist.p + (size_t)p0 * nq * 2, (uint8_t*)L.flags.p + (size_t)p0 * nq,
                        (int32_t*)L.chunk_cnt.p + (size_t)p0 * chunks,
                        (slamb200_dmatch*)L.out.p + (size_t)p0 * cap, cap, (int32_t*)L.n_out.p + p0, s2);
      }
      CU(cudaGetLastError());
      if (s2 != s) {
        CU(cudaEventRecord(L.sub_ev[n_sub], s2));
        CU(cudaStreamWaitEvent(s, L.sub_ev[n_sub], 0));
      }
      CU(cudaEventRecord(L.done, s));
      return SLAMB200_OK;
    }
  }
  CU(cudaGetLastError());
  launch_finalize((const uint4*)L.part.p, nq, d_pairs, n_pairs, n_split, orb ? 1 : 0, ratio,
                  (int32_t*)L.knn_idx.p, (float*)L.knn_dist.p, (uint8_t*)L.flags.p,
                  (int32_t*)L.chunk_cnt.p, (slamb200_dmatch*)L.out.p, cap, (int32_t*)L.n_out.p, s);
  CU(cudaGetLastError());
  CU(cudaEventRecord(L.done, s));
  return SLAMB200_OK;
}

// Copies the last batch of lane L to the host: out = P slabs of out_cap matches, n_out[P].
// The device->host transfer goes through the lane's page-locked staging area at full PCIe speed
// (the caller's buffers are usually pageable: std::vector<cv::DMatch>), then a host memcpy.
static int fetch_batch(Lane& L, cudaStream_t s, slamb200_dmatch* out, int out_cap, int* n_out) {
  const int P = L.b_pairs;
  if (P == 0) return SLAMB200_OK;
  if (!n_out) return fail(SLAMB200_ERR_INVALID, "n_out is NULL");
  if (L.b_nq == 0) {
    for (int p = 0; p < P; p++) n_out[p] = 0;
    return SLAMB200_OK;
  }
  if (out_cap < L.b_nq) return fail(SLAMB200_ERR_INVALID, "cap %d < query rows %d", out_cap, L.b_nq);
  if (!out) return fail(SLAMB200_ERR_INVALID, "out is NULL");
  auto ensure_h_out = [&](size_t bytes) -> int {
    if (L.h_out_cap >= bytes) return SLAMB200_OK;
    if (L.h_out) CU(cudaFreeHost(L.h_out));
    L.h_out = nullptr;
    L.h_out_cap = 0;
    CU(cudaMallocHost(&L.h_out, bytes * 2));
    L.h_out_cap = bytes * 2;
    return SLAMB200_OK;
  };
  // Small results (the per-pair drop-in call: 160 KB for 10k rows) come back speculatively in the
  // same round trip as their counts: one synchronisation instead of two.
  const size_t full = sizeof(slamb200_dmatch) * (size_t)L.b_cap * P;
  const bool speculative = full <= (size_t)1 << 20;
  if (speculative) {
    int rc = ensure_h_out(full);
    if (rc) return rc;
    CU(cudaMemcpyAsync(L.h_out, L.out.p, full, cudaMemcpyDeviceToHost, s));
  }
  CU(cudaMemcpyAsync(L.h_small + 1, L.n_out.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(L.h_small, L.status, 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (L.h_small[0] != 0) cudaMemsetAsync(L.status, 0, 4, s);   // (the status words of small batches persist)
  if (L.h_small[0] != 0)
    return fail(SLAMB200_ERR_INTERNAL, "device self-check failed (flag %d): tensor-core candidates "
                "disagree with the exact rerank", L.h_small[0]);
  int mx = 0;
  for (int p = 0; p < P; p++) {
    n_out[p] = L.h_small[1 + p];
    mx = n_out[p] > mx ? n_out[p] : mx;
  }
  if (mx > 0 && speculative) {
    for (int p = 0; p < P; p++)
      memcpy(out + (size_t)p * out_cap, (const char*)L.h_out + sizeof(slamb200_dmatch) * (size_t)L.b_cap * p,
             sizeof(slamb200_dmatch) * (size_t)n_out[p]);
  } else if (mx > 0) {
    const size_t row = sizeof(slamb200_dmatch) * (size_t)mx;
    int rc = ensure_h_out(row * P);
    if (rc) return rc;
    CU(cudaMemcpy2DAsync(L.h_out, row, L.out.p, sizeof(slamb200_dmatch) * (size_t)L.b_cap, row, P,
                         cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int p = 0; p < P; p++)
      memcpy(out + (size_t)p * out_cap, (const char*)L.h_out + row * p,
             sizeof(slamb200_dmatch) * (size_t)n_out[p]);
  }
  return SLAMB200_OK;
}

extern "C" int slamb200_match_batch(slamb200_ctx* c, int matcher, const slamb200_desc* q,
                                    const slamb200_desc* const* trains, int n_pairs, double ratio,
                                    slamb200_dmatch* out, int cap, int* n_out) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  if (n_pairs < 0) return fail(SLAMB200_ERR_INVALID, "n_pairs < 0");
  int rc = check_matcher(matcher, q, trains, n_pairs);
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  LaneGuard g(c);
  Lane& L = g.lane();
  if ((rc = enqueue_batch(c, L, L.stream, matcher, q, trains, n_pairs, ratio))) return rc;
  return fetch_batch(L, L.stream, out, cap, n_out);
}

extern "C" int slamb200_match_pair(slamb200_ctx* c, int matcher, const slamb200_desc* q,
                                   const slamb200_desc* t, double ratio, slamb200_dmatch* out,
                                   int cap, int* n_out) {
  const slamb200_desc* tt[1] = {t};
  return slamb200_match_batch(c, matcher, q, tt, 1, ratio, out, cap, n_out);
}

extern "C" int slamb200_knn2(slamb200_ctx* c, int matcher, const slamb200_desc* q,
                             const slamb200_desc* t, int32_t* idx, float* dist) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  const slamb200_desc* tt[1] = {t};
  int rc = check_matcher(matcher, q, tt, 1);
  if (rc) return rc;
  if (q->n > 0 && (!idx || !dist)) return fail(SLAMB200_ERR_INVALID, "idx/dist is NULL");
  CU(cudaSetDevice(c->device));
  LaneGuard g(c);
  Lane& L = g.lane();
  if ((rc = enqueue_batch(c, L, L.stream, matcher, q, tt, 1, 0.0, true))) return rc;
  if (q->n > 0) {
    CU(cudaMemcpyAsync(idx, L.knn_idx.p, sizeof(int32_t) * 2 * (size_t)q->n, cudaMemcpyDeviceToHost, L.stream));
    CU(cudaMemcpyAsync(dist, L.knn_dist.p, sizeof(float) * 2 * (size_t)q->n, cudaMemcpyDeviceToHost, L.stream));
  }
  CU(cudaMemcpyAsync(L.h_small, L.status, 4, cudaMemcpyDeviceToHost, L.stream));
  CU(cudaStreamSynchronize(L.stream));
  if (L.h_small[0] != 0)
    return fail(SLAMB200_ERR_INTERNAL, "device self-check failed (flag %d)", L.h_small[0]);
  return SLAMB200_OK;
}

extern "C" int slamb200_match_window(slamb200_ctx* c, int matcher,
                                     const slamb200_desc* const* frames, int n_frames,
                                     double ratio, slamb200_dmatch* out, int cap, int* n_out) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  if (n_frames < 0 || (n_frames > 0 && !frames)) return fail(SLAMB200_ERR_INVALID, "bad frames");
  int pair0 = 0;
  for (int i = 0; i + 1 < n_frames; i++) {
    const int np = n_frames - 1 - i;
    if (frames[i] && cap < frames[i]->n) return fail(SLAMB200_ERR_INVALID, "cap too small");
    int rc = slamb200_match_batch(c, matcher, frames[i], frames + i + 1, np, ratio,
                                  out ? out + (size_t)pair0 * cap : nullptr, cap, n_out + pair0);
    if (rc) return rc;
    pair0 += np;
  }
  return SLAMB200_OK;
}

// ---- the whole window from HOST Mats in one call ------------------------------------------------
// What a search of the reference does with its descriptors (batch.cpp:120-148, :181-201: one
// previous-frame descriptor against every batch element) when none of them is resident yet.  A
// three-stage pipeline inside the library:
//   narrow  : `SLAMB200_HOST_NARROWERS` threads (default: hardware threads - 3, at most 24) take
//             train Mats in window order and narrow whole Mats to bytes into page-locked staging
//             (no CUDA submission from these threads: they only stream host memory);
//   submit  : ONE thread queues the prep kernel of every narrowed Mat (the driver serialises
//             submissions anyway: a dozen threads in the driver contend for its lock);
//   match   : the calling thread matches every chunk of `SLAMB200_HOST_CHUNK` (14) pairs as soon
//             as its last Mat is in, and copies its match lists out.
// The step is bound by the narrowing, i.e. by reading the fp32 Mats out of host DRAM once.  Same
// results as uploading everything and calling slamb200_match_batch.
extern "C" int slamb200_match_batch_host(slamb200_ctx* c, int matcher, const void* q_rows, int nq,
                                         size_t q_stride, const void* const* t_rows, const int* t_n,
                                         const size_t* t_stride, int n_pairs, double ratio,
                                         slamb200_dmatch* out, int cap, int* n_out) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  if (n_pairs < 0 || nq < 0) return fail(SLAMB200_ERR_INVALID, "negative size");
  if (matcher != SLAMB200_SIFT_BF && matcher != SLAMB200_SIFT_FLANN && matcher != SLAMB200_ORB_BF &&
      matcher != SLAMB200_SIFT_BF_L1)
    return fail(SLAMB200_ERR_MATCHER, "matcher type %d is not 0 (SIFT_BF), 1 (SIFT_FLANN), 2 (ORB_BF) or 3 (SIFT_BF_L1)", matcher);
  if (n_pairs > 0 && (!t_rows || !t_n || !n_out)) return fail(SLAMB200_ERR_INVALID, "NULL argument");
  if (nq > 0 && n_pairs > 0 && (!out || cap < nq)) return fail(SLAMB200_ERR_INVALID, "out is NULL or cap < query rows");
  const int kind = matcher == SLAMB200_ORB_BF ? SLAMB200_DESC_U8X32 : SLAMB200_DESC_F32X128;
  static const int narrow_env = [] { const char* e = getenv("SLAMB200_HOST_NARROWERS"); return e ? atoi(e) : 0; }();
  static const int chunk_env = [] { const char* e = getenv("SLAMB200_HOST_CHUNK"); return e ? atoi(e) : 14; }();
  const int chunk = chunk_env > 0 ? chunk_env : 14;
  // chunks of `chunk` pairs, the last ones halving (14, 7, 4, 3): what runs behind the last narrowed
  // Mat is one short match call, not a whole chunk
  std::vector<int> chunk_start{0};
  {
    int rem = n_pairs;
    while (rem > 2 * chunk) { chunk_start.push_back(chunk_start.back() + chunk); rem -= chunk; }
    while (rem > 0) {
      int t = (rem + 1) / 2 < 3 ? 3 : (rem + 1) / 2;
      t = t > rem ? rem : (t > chunk ? chunk : t);
      chunk_start.push_back(chunk_start.back() + t);
      rem -= t;
    }
  }
  const int n_chunks = (int)chunk_start.size() - 1;
  std::vector<int> chunk_of((size_t)(n_pairs > 0 ? n_pairs : 1), 0);
  for (int k = 0; k < n_chunks; k++)
    for (int i = chunk_start[(size_t)k]; i < chunk_start[(size_t)k + 1]; i++) chunk_of[(size_t)i] = k;
  // developer trace (SLAMB200_HOST_TRACE=1): where a step spends its wall time
  static const int trace = [] { const char* e = getenv("SLAMB200_HOST_TRACE"); return e ? atoi(e) : 0; }();
  const auto T0 = std::chrono::steady_clock::now();
  auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count(); };
  std::atomic<int> narrowed(0);
  double t_q = 0, t_spawn = 0, t_first_narrow = 0, t_last_narrow = 0, t_last_submit = 0, t_first_chunk_ready = 0, t_wait_chunks = 0, t_match = 0;
  CU(cudaSetDevice(c->device));
  slamb200_desc* qd = nullptr;
  int rc = slamb200_upload_desc_packed(c, kind, q_rows, nq, q_stride, &qd);
  if (rc != SLAMB200_OK) return rc;
  t_q = since();
  std::vector<slamb200_desc*> td((size_t)n_pairs, nullptr);
  std::vector<int> left((size_t)(n_chunks > 0 ? n_chunks : 1));
  for (int k = 0; k < n_chunks; k++) left[(size_t)k] = chunk_start[(size_t)k + 1] - chunk_start[(size_t)k];
  struct Job { int i; slamb200_ctx::PinBuf* b; };
  std::mutex mu;
  std::condition_variable cv_jobs, cv_chunks;
  std::deque<Job> jobs;
  int first_err = SLAMB200_OK;
  char err_text[512] = "";
  auto stride_of = [&](int i) -> size_t {
    const size_t st = t_stride ? t_stride[i] : 0;
    return st ? st : (kind == SLAMB200_DESC_F32X128 ? 512 : 32);
  };
  // Narrowing, Mat after Mat IN WINDOW ORDER with every thread on the same Mat: a Mat is cut into
  // slices of ~1k rows and one counter hands out (Mat, slice) in order, so Mats become resident one
  // by one (every ~50 us with 16 threads) instead of a thread count of them together every ~800 us
  // -- the first chunk can be matched after 14 Mats' worth of narrowing, and behind the LAST byte
  // narrowed only one Mat's prep and one short chunk remain.
  constexpr int SLICE_ROWS = 1024;
  struct MatState { std::atomic<int> pending; std::atomic<int> ok; std::atomic<slamb200_ctx::PinBuf*> b; std::atomic<int> claimed; };
  std::vector<MatState> ms((size_t)(n_pairs > 0 ? n_pairs : 1));
  std::vector<int> slice0((size_t)n_pairs + 1, 0);
  for (int i = 0; i < n_pairs; i++) {
    const size_t st = stride_of(i);
    const bool packable = kind == SLAMB200_DESC_F32X128 && t_n[i] > 0 && t_rows[i] && st >= 512 && st % 4 == 0;
    const int ns = packable ? (t_n[i] + SLICE_ROWS - 1) / SLICE_ROWS : 1;
    slice0[(size_t)i + 1] = slice0[(size_t)i] + ns;
    ms[(size_t)i].pending.store(ns);
    ms[(size_t)i].ok.store(packable ? 1 : 0);
    ms[(size_t)i].b.store(nullptr);
    ms[(size_t)i].claimed.store(0);
  }
  const int total_slices = slice0[(size_t)n_pairs];
  std::atomic<int> next(0);
  auto narrower = [&] {
    cudaSetDevice(c->device);
    int i = 0;
    for (;;) {
      const int g = next.fetch_add(1);
      if (g >= total_slices) return;
      while (slice0[(size_t)i + 1] <= g) i++;   // (g only grows on this thread)
      MatState& M = ms[(size_t)i];
      const int k = g - slice0[(size_t)i];
      const size_t st = stride_of(i);
      if (M.ok.load(std::memory_order_relaxed)) {
        // the thread that draws a Mat's first slice gets its staging buffer; the others wait for it
        slamb200_ctx::PinBuf* b = nullptr;
        if (k == 0) {
          b = pin_acquire(c, (size_t)t_n[i] * 128);
          if (!b) M.ok.store(0);
          M.b.store(b, std::memory_order_release);
          M.claimed.store(1, std::memory_order_release);
        } else {
          while (!M.claimed.load(std::memory_order_acquire)) __builtin_ia32_pause();
          b = M.b.load(std::memory_order_acquire);
        }
        if (b && M.ok.load(std::memory_order_relaxed)) {
          const int r0 = k * SLICE_ROWS, r1 = r0 + SLICE_ROWS < t_n[i] ? r0 + SLICE_ROWS : t_n[i];
          if (!slamb200_host_pack_u8((const float*)((const char*)t_rows[i] + (size_t)r0 * st), st / 4, r1 - r0,
                                     (uint8_t*)b->p + (size_t)r0 * 128))
            M.ok.store(0);   // not integer valued: the whole Mat takes the fp32 path
        }
      }
      if (M.pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        slamb200_ctx::PinBuf* b = M.b.load(std::memory_order_acquire);
        if (b && !M.ok.load()) {
          pin_release(c, b, nullptr);
          b = nullptr;
        }
        {
          std::lock_guard<std::mutex> lk(mu);
          jobs.push_back({i, b});
          const int kk = narrowed.fetch_add(1) + 1;
          if (kk == 1) t_first_narrow = since();
          if (kk == n_pairs) t_last_narrow = since();
        }
        cv_jobs.notify_one();
      }
    }
  };
  auto submitter = [&] {
    cudaSetDevice(c->device);
    for (int done = 0; done < n_pairs; done++) {
      Job j;
      int err;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_jobs.wait(lk, [&] { return !jobs.empty(); });
        j = jobs.front();
        jobs.pop_front();
        err = first_err;
      }
      int r = err;
      if (err == SLAMB200_OK) {
        r = j.b ? desc_create(c, kind, t_rows[j.i], t_n[j.i], stride_of(j.i), false, nullptr, true, &td[(size_t)j.i], j.b)
                : desc_create(c, kind, t_rows[j.i], t_n[j.i], t_stride ? t_stride[j.i] : 0, false, nullptr, false,
                              &td[(size_t)j.i]);
        if (r != SLAMB200_OK && j.b) cudaDeviceSynchronize();   // a prep kernel may still be reading the staging
      }
      if (j.b) pin_release(c, j.b, nullptr);   // desc_create recorded its event behind the prep kernel
      std::lock_guard<std::mutex> lk(mu);
      if (r != SLAMB200_OK && first_err == SLAMB200_OK) {
        first_err = r;
        snprintf(err_text, sizeof(err_text), "%s", g_err);
      }
      if (--left[(size_t)chunk_of[(size_t)j.i]] == 0) cv_chunks.notify_all();
      if (done == n_pairs - 1) t_last_submit = since();
    }
  };
  // narrowing threads: the context's persistent pack pool (slamb200_set_pack_threads; default: the
  // hardware threads, at most 24), all of it unless SLAMB200_HOST_NARROWERS says fewer; one thread
  // of this call queues the prep kernels
  pack_pool_start(c);
  int n_narrow = (int)c->pack_pool.th.size();
  {
    // one hardware thread stays free for the submitting and the matching thread: with every core
    // narrowing they are scheduled late and the step takes twice as long (measured)
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 1 && n_narrow > hw - 1) n_narrow = hw - 1;
  }
  if (narrow_env > 0 && narrow_env < n_narrow) n_narrow = narrow_env;
  if (n_narrow > total_slices) n_narrow = total_slices;
  std::atomic<int> workers_left(n_narrow);
  std::mutex done_mu;
  std::condition_variable done_cv;
  c->pool_busy.fetch_add(1);
  for (int k = 0; k < n_narrow; k++)
    c->pack_pool.submit([&] {
      narrower();
      if (workers_left.fetch_sub(1) == 1) {
        std::lock_guard<std::mutex> lk(done_mu);
        done_cv.notify_all();
      }
    });
  std::vector<std::thread> th;
  if (n_pairs > 0) th.emplace_back(submitter);
  if (n_narrow == 0 && n_pairs > 0) narrower();   // a context without pool threads: narrow here
  t_spawn = since();
  for (int k = 0; k < n_chunks; k++) {
    {
      const double tw = since();
      std::unique_lock<std::mutex> lk(mu);
      cv_chunks.wait(lk, [&] { return left[(size_t)k] == 0; });
      t_wait_chunks += since() - tw;
      if (k == 0) t_first_chunk_ready = since();
      if (first_err != SLAMB200_OK) break;
    }
    const int p0 = chunk_start[(size_t)k], np = chunk_start[(size_t)k + 1] - p0;
    const double tm = since();
    const int r = slamb200_match_batch(c, matcher, qd, td.data() + p0, np, ratio,
                                       out ? out + (size_t)p0 * cap : nullptr, cap, n_out + p0);
    t_match += since() - tm;
    if (r != SLAMB200_OK) {
      std::lock_guard<std::mutex> lk(mu);
      if (first_err == SLAMB200_OK) {
        first_err = r;
        snprintf(err_text, sizeof(err_text), "%s", g_err);
      }
      break;
    }
    for (int p = p0; p < p0 + np; p++) {
      slamb200_free_desc(c, td[(size_t)p]);
      td[(size_t)p] = nullptr;
    }
  }
  const double t_chunks_done = since();
  {   // (an error may have ended the loop early: the workers still hold references to this frame)
    std::unique_lock<std::mutex> lk(done_mu);
    done_cv.wait(lk, [&] { return workers_left.load() <= 0; });
  }
  c->pool_busy.fetch_sub(1);
  for (auto& t : th) t.join();
  for (slamb200_desc* d : td)
    if (d) slamb200_free_desc(c, d);
  slamb200_free_desc(c, qd);
  if (trace)
    fprintf(stderr, "[match_batch_host] q %.2f spawn %.2f first narrow %.2f first chunk ready %.2f last narrow %.2f "
            "last submit %.2f chunks done %.2f end %.2f ms | waiting for chunks %.2f, in match calls %.2f (%d narrowers)\n",
            t_q, t_spawn, t_first_narrow, t_first_chunk_ready, t_last_narrow, t_last_submit, t_chunks_done, since(),
            t_wait_chunks, t_match, n_narrow);
  if (first_err != SLAMB200_OK) return fail(first_err, "%s", err_text);
  return SLAMB200_OK;
}

extern "C" int slamb200_match_batch_enqueue(slamb200_ctx* c, int matcher, const slamb200_desc* q,
                                            const slamb200_desc* const* trains, int n_pairs,
                                            double ratio, void* stream) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  if (n_pairs < 0) return fail(SLAMB200_ERR_INVALID, "n_pairs < 0");
  int rc = check_matcher(matcher, q, trains, n_pairs);
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> lk(c->batch_mu);
  Lane& L = c->lanes[0];
  cudaStream_t s = stream ? (cudaStream_t)stream : L.stream;
  // Lane 0 owns ONE scratch set; the previous batch may have been queued on another stream and
  // may still be running or unfetched there: order this stream behind it (nothing to do in the
  // usual same-stream case).
  if (!L.b_stream_set || L.b_stream != s) CU(cudaStreamWaitEvent(s, L.done, 0));
  L.b_stream = s;
  L.b_stream_set = true;
  return enqueue_batch(c, L, s, matcher, q, trains, n_pairs, ratio);
}

extern "C" int slamb200_batch_fetch(slamb200_ctx* c, slamb200_dmatch* out, int cap, int* n_out,
                                    void* stream) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> lk(c->batch_mu);
  Lane& L = c->lanes[0];
  cudaStream_t s = stream ? (cudaStream_t)stream : L.stream;
  if (!L.b_stream_set || L.b_stream != s) CU(cudaStreamWaitEvent(s, L.done, 0));   // enqueued on another stream
  return fetch_batch(L, s, out, cap, n_out);
}

// ---- RANSAC essential scoring -----------------------------------------------------------------
static int make_score_params(const double K[4], double threshold_px, ScoreParams* sp) {
  if (!K) return fail(SLAMB200_ERR_INVALID, "K is NULL");
  const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
  sp->ax = 1.0 / fx;
  sp->ay = 1.0 / fy;
  sp->bx = -cx * sp->ax;
  sp->by = -cy * sp->ay;
  const double thr = threshold_px / ((fx + fy) / 2);
  const float t = (float)(thr * thr);
  const float u = nextafterf(t, INFINITY);
  const double mid = ((double)t + (double)u) * 0.5;
  sp->t = t;
  sp->mid = mid;
  sp->mid_lo = mid * (1.0 - ldexp(1.0, -49));
  sp->mid_hi = mid * (1.0 + ldexp(1.0, -49));
  return SLAMB200_OK;
}

extern "C" int slamb200_score_essential_batch(slamb200_ctx* c, int P, const float* pts1,
                                              const float* pts2, const int32_t* m_off,
                                              const double K[4], const double* E, int H,
                                              double threshold_px, int32_t* counts, int32_t* best,
                                              uint8_t* best_mask) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  if (P < 0 || H < 0) return fail(SLAMB200_ERR_INVALID, "P or H negative");
  if (P == 0) return SLAMB200_OK;
  if (!m_off) return fail(SLAMB200_ERR_INVALID, "m_off is NULL");
  for (int p = 0; p < P; p++)
    if (m_off[p + 1] < m_off[p]) return fail(SLAMB200_ERR_INVALID, "m_off not monotone");
  const int total = m_off[P] - m_off[0];
  if (m_off[0] != 0) return fail(SLAMB200_ERR_INVALID, "m_off[0] must be 0");
  if ((total > 0 && (!pts1 || !pts2)) || (H > 0 && !E)) return fail(SLAMB200_ERR_INVALID, "NULL input");
  if ((H > 0 && !counts) || !best) return fail(SLAMB200_ERR_INVALID, "NULL output");
  ScoreParams sp;
  int rc = make_score_params(K, threshold_px, &sp);
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  const size_t tot = total > 0 ? total : 1;
  if ((rc = buf_reserve(c, L.p1, tot * 8, s))) return rc;
  if ((rc = buf_reserve(c, L.p2, tot * 8, s))) return rc;
  if ((rc = buf_reserve(c, L.npts, tot * 32, s))) return rc;
  if ((rc = buf_reserve(c, L.m_off, sizeof(int32_t) * (size_t)(P + 1), s))) return rc;
  if ((rc = buf_reserve(c, L.E, sizeof(double) * 9 * (size_t)P * (H > 0 ? H : 1), s))) return rc;
  if ((rc = buf_reserve(c, L.counts, sizeof(int32_t) * (size_t)P * (H > 0 ? H : 1), s))) return rc;
  if ((rc = buf_reserve(c, L.best, sizeof(int32_t) * (size_t)P, s))) return rc;
  if ((rc = buf_reserve(c, L.mask, tot, s))) return rc;
  if (total > 0) {
    CU(cudaMemcpyAsync(L.p1.p, pts1, (size_t)total * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(L.p2.p, pts2, (size_t)total * 8, cudaMemcpyHostToDevice, s));
  }
  CU(cudaMemcpyAsync(L.m_off.p, m_off, sizeof(int32_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, s));
  if (H > 0) CU(cudaMemcpyAsync(L.E.p, E, sizeof(double) * 9 * (size_t)P * H, cudaMemcpyHostToDevice, s));
  launch_normalize_points((const float2*)L.p1.p, (const float2*)L.p2.p, total, sp, (double4*)L.npts.p, s);
  {
    ProfScope ps(c, s, SLAMB200_K_RANSAC);
    launch_score_counts((const double4*)L.npts.p, (const int32_t*)L.m_off.p, nullptr, 0,
                        (const double*)L.E.p, H, P, sp, (int32_t*)L.counts.p, s);
  }
  launch_score_best((const int32_t*)L.counts.p, H, P, 4, (int32_t*)L.best.p, s);
  if (best_mask)
    launch_score_mask((const double4*)L.npts.p, (const int32_t*)L.m_off.p, nullptr, 0,
                      (const double*)L.E.p, H, P, (const int32_t*)L.best.p, sp, (uint8_t*)L.mask.p, s);
  CU(cudaGetLastError());
  if (H > 0) CU(cudaMemcpyAsync(counts, L.counts.p, sizeof(int32_t) * (size_t)P * H, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(best, L.best.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, s));
  if (best_mask && total > 0) CU(cudaMemcpyAsync(best_mask, L.mask.p, (size_t)total, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SLAMB200_OK;
}

extern "C" int slamb200_score_essential(slamb200_ctx* c, const float* pts1, const float* pts2,
                                        int M, const double K[4], const double* E, int H,
                                        double threshold_px, int32_t* counts, int32_t* best,
                                        uint8_t* best_mask, uint8_t* all_masks) {
  if (M < 0) return fail(SLAMB200_ERR_INVALID, "M negative");
  int32_t off[2] = {0, M};
  int rc = slamb200_score_essential_batch(c, 1, pts1, pts2, off, K, E, H, threshold_px, counts,
                                          best, best_mask);
  if (rc || !all_masks || M == 0 || H == 0) return rc;
  // parity aid: every (hypothesis, match) flag
  ScoreParams sp;
  if ((rc = make_score_params(K, threshold_px, &sp))) return rc;
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  if ((rc = buf_reserve(c, L.p1, (size_t)M * 8, s))) return rc;
  if ((rc = buf_reserve(c, L.p2, (size_t)M * 8, s))) return rc;
  if ((rc = buf_reserve(c, L.npts, (size_t)M * 32, s))) return rc;
  if ((rc = buf_reserve(c, L.E, sizeof(double) * 9 * (size_t)H, s))) return rc;
  if ((rc = buf_reserve(c, L.all_masks, (size_t)H * M, s))) return rc;
  CU(cudaMemcpyAsync(L.p1.p, pts1, (size_t)M * 8, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(L.p2.p, pts2, (size_t)M * 8, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(L.E.p, E, sizeof(double) * 9 * (size_t)H, cudaMemcpyHostToDevice, s));
  launch_normalize_points((const float2*)L.p1.p, (const float2*)L.p2.p, M, sp, (double4*)L.npts.p, s);
  launch_score_all_masks((const double4*)L.npts.p, M, (const double*)L.E.p, H, sp, (uint8_t*)L.all_masks.p, s);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(all_masks, L.all_masks.p, (size_t)H * M, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SLAMB200_OK;
}

// ---- solvePnPRansac scoring ---------------------------------------------------
static int make_pnp_params(const double K[4], const double* dist, int n_dist, double reproj_err,
                           PnpParams* pp) {
  if (!K) return fail(SLAMB200_ERR_INVALID, "K is NULL");
  if (n_dist < 0 || (n_dist > 0 && !dist)) return fail(SLAMB200_ERR_INVALID, "bad dist/n_dist");
  if (n_dist > 14) return fail(SLAMB200_ERR_INVALID, "more than 14 distortion coefficients");
  pp->fx = K[0]; pp->fy = K[1]; pp->cx = K[2]; pp->cy = K[3];
  for (int i = 0; i < 12; i++) pp->k[i] = i < n_dist ? dist[i] : 0.0;
  for (int i = 12; i < n_dist; i++)
    if (dist[i] != 0.0) return fail(SLAMB200_ERR_INVALID, "tilted sensor model (tauX, tauY) is not supported");
  pp->t = (float)(reproj_err * reproj_err);
  pp->dist_level = 0;
  for (int i = 0; i < 5; i++) if (pp->k[i] != 0.0) pp->dist_level = 1;
  for (int i = 5; i < 12; i++) if (pp->k[i] != 0.0) pp->dist_level = 2;
  return SLAMB200_OK;
}

extern "C" int slamb200_score_pnp_batch(slamb200_ctx* c, int P, const float* obj, const float* img,
                                        const int32_t* m_off, const double K[4],
                                        const double* dist, int n_dist, const double* poses, int H,
                                        double reproj_err, int model_points, int32_t* counts,
                                        int32_t* best, uint8_t* best_mask) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  if (P < 0 || H < 0) return fail(SLAMB200_ERR_INVALID, "P or H negative");
  if (model_points < 1) return fail(SLAMB200_ERR_INVALID, "model_points < 1");
  if (P == 0) return SLAMB200_OK;
  if (!m_off) return fail(SLAMB200_ERR_INVALID, "m_off is NULL");
  if (m_off[0] != 0) return fail(SLAMB200_ERR_INVALID, "m_off[0] must be 0");
  for (int p = 0; p < P; p++)
    if (m_off[p + 1] < m_off[p]) return fail(SLAMB200_ERR_INVALID, "m_off not monotone");
  const int total = m_off[P];
  if ((total > 0 && (!obj || !img)) || (H > 0 && !poses)) return fail(SLAMB200_ERR_INVALID, "NULL input");
  if ((H > 0 && !counts) || !best) return fail(SLAMB200_ERR_INVALID, "NULL output");
  PnpParams pp;
  int rc = make_pnp_params(K, dist, n_dist, reproj_err, &pp);
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  const size_t tot = total > 0 ? total : 1;
  const size_t hh = (size_t)P * (H > 0 ? H : 1);
  if ((rc = buf_reserve(c, L.p1, tot * 12, s))) return rc;
  if ((rc = buf_reserve(c, L.p2, tot * 8, s))) return rc;
  if ((rc = buf_reserve(c, L.npts, tot * 32, s))) return rc;
  if ((rc = buf_reserve(c, L.m_off, sizeof(int32_t) * (size_t)(P + 1), s))) return rc;
  if ((rc = buf_reserve(c, L.E, sizeof(double) * 12 * hh, s))) return rc;
  if ((rc = buf_reserve(c, L.counts, sizeof(int32_t) * hh, s))) return rc;
  if ((rc = buf_reserve(c, L.best, sizeof(int32_t) * (size_t)P, s))) return rc;
  if ((rc = buf_reserve(c, L.mask, tot, s))) return rc;
  if (total > 0) {
    CU(cudaMemcpyAsync(L.p1.p, obj, (size_t)total * 12, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(L.p2.p, img, (size_t)total * 8, cudaMemcpyHostToDevice, s));
  }
  CU(cudaMemcpyAsync(L.m_off.p, m_off, sizeof(int32_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, s));
  if (H > 0) CU(cudaMemcpyAsync(L.E.p, poses, sizeof(double) * 12 * (size_t)P * H, cudaMemcpyHostToDevice, s));
  launch_pack_pnp_points((const float*)L.p1.p, (const float2*)L.p2.p, total, (double4*)L.npts.p, s);
  {
    ProfScope ps(c, s, SLAMB200_K_PNP);
    launch_pnp_counts((const double4*)L.npts.p, (const int32_t*)L.m_off.p, (const double*)L.E.p, H,
                      P, pp, (int32_t*)L.counts.p, s);
  }
  launch_score_best((const int32_t*)L.counts.p, H, P, model_points - 1, (int32_t*)L.best.p, s);
  if (best_mask)
    launch_pnp_mask((const double4*)L.npts.p, (const int32_t*)L.m_off.p, (const double*)L.E.p, H, P,
                    (const int32_t*)L.best.p, pp, (uint8_t*)L.mask.p, s);
  CU(cudaGetLastError());
  if (H > 0) CU(cudaMemcpyAsync(counts, L.counts.p, sizeof(int32_t) * (size_t)P * H, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(best, L.best.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, s));
  if (best_mask && total > 0) CU(cudaMemcpyAsync(best_mask, L.mask.p, (size_t)total, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SLAMB200_OK;
}

extern "C" int slamb200_score_pnp(slamb200_ctx* c, const float* obj, const float* img, int M,
                                  const double K[4], const double* dist, int n_dist,
                                  const double* poses, int H, double reproj_err, int model_points,
                                  int32_t* counts, int32_t* best, uint8_t* best_mask,
                                  uint8_t* all_masks) {
  if (M < 0) return fail(SLAMB200_ERR_INVALID, "M negative");
  int32_t off[2] = {0, M};
  int rc = slamb200_score_pnp_batch(c, 1, obj, img, off, K, dist, n_dist, poses, H, reproj_err,
                                    model_points, counts, best, best_mask);
  if (rc || !all_masks || M == 0 || H == 0) return rc;
  PnpParams pp;
  if ((rc = make_pnp_params(K, dist, n_dist, reproj_err, &pp))) return rc;
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  if ((rc = buf_reserve(c, L.p1, (size_t)M * 12, s))) return rc;
  if ((rc = buf_reserve(c, L.p2, (size_t)M * 8, s))) return rc;
  if ((rc = buf_reserve(c, L.npts, (size_t)M * 32, s))) return rc;
  if ((rc = buf_reserve(c, L.E, sizeof(double) * 12 * (size_t)H, s))) return rc;
  if ((rc = buf_reserve(c, L.all_masks, (size_t)H * M, s))) return rc;
  CU(cudaMemcpyAsync(L.p1.p, obj, (size_t)M * 12, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(L.p2.p, img, (size_t)M * 8, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(L.E.p, poses, sizeof(double) * 12 * (size_t)H, cudaMemcpyHostToDevice, s));
  launch_pack_pnp_points((const float*)L.p1.p, (const float2*)L.p2.p, M, (double4*)L.npts.p, s);
  launch_pnp_all_masks((const double4*)L.npts.p, M, (const double*)L.E.p, H, pp, (uint8_t*)L.all_masks.p, s);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(all_masks, L.all_masks.p, (size_t)H * M, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SLAMB200_OK;
}

// ---- ORB descriptors of given keypoints ---------------------------------------
extern "C" int slamb200_orb_compute(slamb200_ctx* c, const uint8_t* image, int rows, int cols,
                                    int channels, size_t step, const float* kps, int n,
                                    uint8_t* keep, uint8_t* desc, int* n_kept,
                                    slamb200_desc** resident) {
  if (!c || !n_kept) return fail(SLAMB200_ERR_INVALID, "orb_compute: NULL argument");
  *n_kept = 0;
  if (resident) *resident = nullptr;
  if (rows <= 0 || cols <= 0 || !image) return fail(SLAMB200_ERR_INVALID, "orb_compute: bad image");
  if (channels != 1 && channels != 3)
    return fail(SLAMB200_ERR_KIND, "orb_compute: %d-channel image (CV_8UC1 or CV_8UC3 expected)", channels);
  if (step == 0) step = (size_t)cols * channels;
  if (step < (size_t)cols * channels) return fail(SLAMB200_ERR_INVALID, "orb_compute: step too small");
  if (n < 0 || (n > 0 && !kps)) return fail(SLAMB200_ERR_INVALID, "orb_compute: bad keypoints");
  // KeyPointsFilter::runByImageBorder(keypoints, size, edgeThreshold = 31), order preserved; the
  // centre pixel (cvRound) and the rotation (float cosf / sinf of angle * pi/180) per kept keypoint
  std::vector<OrbKeypoint> hk;
  hk.reserve((size_t)n);
  for (int i = 0; i < n; i++) {
    const float x = kps[3 * i], y = kps[3 * i + 1];
    const int rx = (int)lrintf(x), ry = (int)lrintf(y);  // Rect::contains(Point(pt)) rounds first
    const bool ok = rx >= 31 && rx < cols - 31 && ry >= 31 && ry < rows - 31;
    if (keep) keep[i] = ok ? 1 : 0;
    if (!ok) continue;
    float angle = kps[3 * i + 2];
    angle *= (float)(3.141592653589793 / 180.f);
    hk.push_back({rx, ry, cosf(angle), sinf(angle)});
  }
  const int kept = (int)hk.size();
  *n_kept = kept;
  CU(cudaSetDevice(c->device));
  if (orb_pattern_upload() != 0) return fail(SLAMB200_ERR_CUDA, "orb_compute: pattern upload failed");
  int rc;
  const int n_pad = round_up(kept > 0 ? kept : 1, SLAMB200_TILE_PAD);
  {
    LaneGuard g(c);
    Lane& L = g.lane();
    cudaStream_t s = L.stream;
    const size_t px = (size_t)rows * cols;
    if ((rc = buf_reserve(c, L.orb_img, (size_t)rows * step, s))) return rc;
    if ((rc = buf_reserve(c, L.orb_gray, px, s))) return rc;
    if ((rc = buf_reserve(c, L.orb_rowf, px * sizeof(float), s))) return rc;
    if ((rc = buf_reserve(c, L.orb_blur, px, s))) return rc;
    if ((rc = buf_reserve(c, L.orb_kp, sizeof(OrbKeypoint) * (size_t)(kept > 0 ? kept : 1), s))) return rc;
    if ((rc = buf_reserve(c, L.orb_desc, (size_t)n_pad * 32, s))) return rc;
    CU(cudaMemcpyAsync(L.orb_img.p, image, (size_t)rows * step, cudaMemcpyHostToDevice, s));
    if (kept > 0)
      CU(cudaMemcpyAsync(L.orb_kp.p, hk.data(), sizeof(OrbKeypoint) * (size_t)kept, cudaMemcpyHostToDevice, s));
    {
      ProfScope ps(c, s, SLAMB200_K_ORB_DESC);
      launch_orb_blur((const uint8_t*)L.orb_img.p, rows, cols, channels, step, (uint8_t*)L.orb_gray.p,
                      (float*)L.orb_rowf.p, (uint8_t*)L.orb_blur.p, s);
      launch_orb_desc((const uint8_t*)L.orb_blur.p, cols, (const OrbKeypoint*)L.orb_kp.p, kept, n_pad,
                      (uint8_t*)L.orb_desc.p, s);
    }
    CU(cudaGetLastError());
    if (desc && kept > 0)
      CU(cudaMemcpyAsync(desc, L.orb_desc.p, (size_t)kept * 32, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(L.done, s));
    // the host vector and the caller's image must outlive the copies; the resident set below is
    // built from the device rows on the same stream
    if (resident) {
      // desc_create takes an upload lane of its own; order it behind this lane's work
      rc = desc_create(c, SLAMB200_DESC_U8X32, L.orb_desc.p, kept, 32, true, s, false, resident);
      if (rc) return rc;
      // the lane's descriptor buffer may be overwritten by the next call only after that copy
      CU(cudaStreamWaitEvent(s, (*resident)->ready, 0));
    }
    CU(cudaStreamSynchronize(s));
  }
  return SLAMB200_OK;
}

// ---- SIFT descriptors of given keypoints (second half) ----------------------------
extern "C" int slamb200_sift_compute(slamb200_ctx* c, const uint8_t* image, int rows, int cols,
                                     int channels, size_t step, const float* kps, int n, float* desc,
                                     slamb200_desc** resident) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "sift_compute: ctx is NULL");
  if (resident) *resident = nullptr;
  if (rows <= 0 || cols <= 0 || !image) return fail(SLAMB200_ERR_INVALID, "sift_compute: bad image");
  if (channels != 1 && channels != 3)
    return fail(SLAMB200_ERR_KIND, "sift_compute: %d-channel image (CV_8UC1 or CV_8UC3 expected)", channels);
  if (step == 0) step = (size_t)cols * channels;
  if (step < (size_t)cols * channels) return fail(SLAMB200_ERR_INVALID, "sift_compute: step too small");
  if (n < 0 || (n > 0 && !kps)) return fail(SLAMB200_ERR_INVALID, "sift_compute: bad keypoints");
  // what calcSIFTDescriptor derives from a keypoint before its sample loop (float, libm's cosf / sinf)
  std::vector<SiftKeypoint> hk((size_t)n);
  const int max_radius = (int)sqrt((double)cols * cols + (double)rows * rows);
  for (int i = 0; i < n; i++) {
    const float x = kps[4 * (size_t)i], y = kps[4 * (size_t)i + 1], size = kps[4 * (size_t)i + 2];
    float angle = 360.f - kps[4 * (size_t)i + 3];
    if (fabsf(angle - 360.f) < 1.1920928955078125e-07f) angle = 0.f;
    const float scl = size * 0.5f, hist_width = 3.0f * scl;
    SiftKeypoint k;
    k.ptx = (int)lrintf(x);
    k.pty = (int)lrintf(y);
    k.ori = angle;
    k.cos_t = cosf(angle * (float)(3.14159265358979323846 / 180)) / hist_width;
    k.sin_t = sinf(angle * (float)(3.14159265358979323846 / 180)) / hist_width;
    int radius = (int)lrintf(hist_width * 1.4142135623730951f * 5 * 0.5f);
    k.radius = radius < max_radius ? radius : max_radius;
    if (!(size > 0.f) || k.radius > 4096)
      return fail(SLAMB200_ERR_INVALID, "sift_compute: keypoint %d has size %g", i, (double)size);
    hk[(size_t)i] = k;
  }
  CU(cudaSetDevice(c->device));
  if (sift_gauss_upload() != 0) return fail(SLAMB200_ERR_CUDA, "sift_compute: kernel table upload failed");
  int rc;
  const int n_pad = round_up(n > 0 ? n : 1, SLAMB200_TILE_PAD);
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  const size_t px = (size_t)rows * cols;
  if ((rc = buf_reserve(c, L.orb_img, (size_t)rows * step, s))) return rc;
  if ((rc = buf_reserve(c, L.orb_gray, px, s))) return rc;
  if ((rc = buf_reserve(c, L.sift_rowf, px * sizeof(float), s))) return rc;
  if ((rc = buf_reserve(c, L.sift_base, px * sizeof(float), s))) return rc;
  if ((rc = buf_reserve(c, L.sift_kp, sizeof(SiftKeypoint) * (size_t)(n > 0 ? n : 1), s))) return rc;
  if ((rc = buf_reserve(c, L.sift_desc, sizeof(float) * 128 * (size_t)n_pad, s))) return rc;
  CU(cudaMemcpyAsync(L.orb_img.p, image, (size_t)rows * step, cudaMemcpyHostToDevice, s));
  if (n > 0) CU(cudaMemcpyAsync(L.sift_kp.p, hk.data(), sizeof(SiftKeypoint) * (size_t)n, cudaMemcpyHostToDevice, s));
  {
    ProfScope ps(c, s, SLAMB200_K_SIFT_DESC);
    launch_orb_gray((const uint8_t*)L.orb_img.p, rows, cols, channels, step, (uint8_t*)L.orb_gray.p, s);
    launch_sift_base((const uint8_t*)L.orb_gray.p, rows, cols, (float*)L.sift_rowf.p, (float*)L.sift_base.p, s);
    if (launch_sift_desc((const float*)L.sift_base.p, rows, cols, (const SiftKeypoint*)L.sift_kp.p, n, n_pad,
                         (float*)L.sift_desc.p, s) != 0)
      return fail(SLAMB200_ERR_CUDA, "sift_compute: kernel configuration failed");
  }
  CU(cudaGetLastError());
  if (desc && n > 0)
    CU(cudaMemcpyAsync(desc, L.sift_desc.p, sizeof(float) * 128 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(L.done, s));
  if (resident) {
    // the rows are already in HBM: the resident set is prepared from them on an upload lane,
    // ordered behind this lane's work; this lane's buffer may be reused only after that
    rc = desc_create(c, SLAMB200_DESC_F32X128, L.sift_desc.p, n, 512, true, s, false, resident);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(s, (*resident)->ready, 0));
  }
  CU(cudaStreamSynchronize(s));   // hk and the caller's image must outlive the copies
  return SLAMB200_OK;
}

// ---- FAST keypoints (fastExtractor.cpp:7-13) --------------------------------------
extern "C" int slamb200_fast_detect(slamb200_ctx* c, const uint8_t* image, int rows, int cols,
                                    int channels, size_t step, int threshold, int nonmax,
                                    float* kps, int cap, int* n_found) {
  if (!c || !n_found) return fail(SLAMB200_ERR_INVALID, "fast_detect: NULL argument");
  *n_found = 0;
  if (rows < 0 || cols < 0 || ((rows > 0 && cols > 0) && !image))
    return fail(SLAMB200_ERR_INVALID, "fast_detect: bad image");
  if (channels != 1 && channels != 3)
    return fail(SLAMB200_ERR_KIND, "fast_detect: %d-channel image (CV_8UC1 or CV_8UC3 expected)", channels);
  if (cap < 0 || (cap > 0 && !kps)) return fail(SLAMB200_ERR_INVALID, "fast_detect: bad output buffer");
  if (rows > 65535) return fail(SLAMB200_ERR_INVALID, "fast_detect: more than 65535 rows");
  if (rows < 7 || cols < 7) return SLAMB200_OK;   // no pixel is 3 px away from every border
  if (step == 0) step = (size_t)cols * channels;
  if (step < (size_t)cols * channels) return fail(SLAMB200_ERR_INVALID, "fast_detect: step too small");
  CU(cudaSetDevice(c->device));
  int rc;
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  const size_t px = (size_t)rows * cols;
  const int nb = fast_blocks(rows, cols);
  if ((rc = buf_reserve(c, L.orb_img, (size_t)rows * step, s))) return rc;
  if ((rc = buf_reserve(c, L.orb_gray, px, s))) return rc;
  if ((rc = buf_reserve(c, L.fast_score, px * sizeof(int16_t), s))) return rc;
  if ((rc = buf_reserve(c, L.fast_cnt, sizeof(int32_t) * ((size_t)nb + 1), s))) return rc;
  if ((rc = buf_reserve(c, L.fast_kp, sizeof(float) * 3 * (size_t)(cap > 0 ? cap : 1), s))) return rc;
  CU(cudaMemcpyAsync(L.orb_img.p, image, (size_t)rows * step, cudaMemcpyHostToDevice, s));
  {
    ProfScope ps(c, s, SLAMB200_K_FAST);
    launch_orb_gray((const uint8_t*)L.orb_img.p, rows, cols, channels, step, (uint8_t*)L.orb_gray.p, s);
    launch_fast_detect((const uint8_t*)L.orb_gray.p, rows, cols, threshold, nonmax ? 1 : 0,
                       (int16_t*)L.fast_score.p, (int32_t*)L.fast_cnt.p, (float*)L.fast_kp.p, cap, s);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(L.h_small, (int32_t*)L.fast_cnt.p + nb, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(L.done, s));
  CU(cudaStreamSynchronize(s));
  const int found = L.h_small[0];
  *n_found = found;
  const int n_copy = found < cap ? found : cap;
  if (n_copy > 0) {
    CU(cudaMemcpyAsync(kps, L.fast_kp.p, sizeof(float) * 3 * (size_t)n_copy, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  return SLAMB200_OK;
}

// fastExtractor followed by extractDescriptor(ORB) on the same frame in one call: the frame is
// uploaded and converted to gray once, the keypoints go to the host (the reference keeps them in
// BatchElement::features) and the descriptors of those that survive ORB's border filter stay in
// HBM as a resident set.  Equal to slamb200_fast_detect + slamb200_orb_compute.
extern "C" int slamb200_fast_orb_compute(slamb200_ctx* c, const uint8_t* image, int rows, int cols,
                                         int channels, size_t step, int threshold, int nonmax,
                                         float* kps, int cap, int* n_found, uint8_t* keep,
                                         uint8_t* desc, int* n_kept, slamb200_desc** resident) {
  if (!c || !n_found || !n_kept) return fail(SLAMB200_ERR_INVALID, "fast_orb_compute: NULL argument");
  *n_found = 0;
  *n_kept = 0;
  if (resident) *resident = nullptr;
  if (rows <= 0 || cols <= 0 || !image) return fail(SLAMB200_ERR_INVALID, "fast_orb_compute: bad image");
  if (channels != 1 && channels != 3)
    return fail(SLAMB200_ERR_KIND, "fast_orb_compute: %d-channel image (CV_8UC1 or CV_8UC3 expected)", channels);
  if (cap < 0 || (cap > 0 && !kps)) return fail(SLAMB200_ERR_INVALID, "fast_orb_compute: bad output buffer");
  if (rows > 65535) return fail(SLAMB200_ERR_INVALID, "fast_orb_compute: more than 65535 rows");
  if (step == 0) step = (size_t)cols * channels;
  if (step < (size_t)cols * channels) return fail(SLAMB200_ERR_INVALID, "fast_orb_compute: step too small");
  CU(cudaSetDevice(c->device));
  if (orb_pattern_upload() != 0) return fail(SLAMB200_ERR_CUDA, "fast_orb_compute: pattern upload failed");
  int rc;
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  const size_t px = (size_t)rows * cols;
  int found = 0;
  if ((rc = buf_reserve(c, L.orb_img, (size_t)rows * step, s))) return rc;
  if ((rc = buf_reserve(c, L.orb_gray, px, s))) return rc;
  CU(cudaMemcpyAsync(L.orb_img.p, image, (size_t)rows * step, cudaMemcpyHostToDevice, s));
  launch_orb_gray((const uint8_t*)L.orb_img.p, rows, cols, channels, step, (uint8_t*)L.orb_gray.p, s);
  if (rows >= 7 && cols >= 7) {
    const int nb = fast_blocks(rows, cols);
    if ((rc = buf_reserve(c, L.fast_score, px * sizeof(int16_t), s))) return rc;
    if ((rc = buf_reserve(c, L.fast_cnt, sizeof(int32_t) * ((size_t)nb + 1), s))) return rc;
    if ((rc = buf_reserve(c, L.fast_kp, sizeof(float) * 3 * (size_t)(cap > 0 ? cap : 1), s))) return rc;
    {
      ProfScope ps(c, s, SLAMB200_K_FAST);
      launch_fast_detect((const uint8_t*)L.orb_gray.p, rows, cols, threshold, nonmax ? 1 : 0,
                         (int16_t*)L.fast_score.p, (int32_t*)L.fast_cnt.p, (float*)L.fast_kp.p, cap, s);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(L.h_small, (int32_t*)L.fast_cnt.p + nb, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    found = L.h_small[0];
  }
  *n_found = found;
  if (found > cap)
    return fail(SLAMB200_ERR_INVALID, "fast_orb_compute: %d keypoints, buffer holds %d (retry with cap >= *n_found)",
                found, cap);
  if (found > 0) {
    CU(cudaMemcpyAsync(kps, L.fast_kp.p, sizeof(float) * 3 * (size_t)found, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  // KeyPointsFilter::runByImageBorder(.., 31) on the host, as in slamb200_orb_compute; every FAST
  // keypoint has angle -1 degree, rotated with libm's float cosf / sinf like cv::ORB does
  float angle = -1.0f;
  angle *= (float)(3.141592653589793 / 180.f);
  const float ca = cosf(angle), sa = sinf(angle);
  std::vector<OrbKeypoint> hk;
  hk.reserve((size_t)found);
  for (int i = 0; i < found; i++) {
    const int rx = (int)lrintf(kps[3 * (size_t)i]), ry = (int)lrintf(kps[3 * (size_t)i + 1]);
    const bool ok = rx >= 31 && rx < cols - 31 && ry >= 31 && ry < rows - 31;
    if (keep) keep[i] = ok ? 1 : 0;
    if (ok) hk.push_back({rx, ry, ca, sa});
  }
  const int kept = (int)hk.size();
  *n_kept = kept;
  const int n_pad = round_up(kept > 0 ? kept : 1, SLAMB200_TILE_PAD);
  if ((rc = buf_reserve(c, L.orb_rowf, px * sizeof(float), s))) return rc;
  if ((rc = buf_reserve(c, L.orb_blur, px, s))) return rc;
  if ((rc = buf_reserve(c, L.orb_kp, sizeof(OrbKeypoint) * (size_t)(kept > 0 ? kept : 1), s))) return rc;
  if ((rc = buf_reserve(c, L.orb_desc, (size_t)n_pad * 32, s))) return rc;
  if (kept > 0)
    CU(cudaMemcpyAsync(L.orb_kp.p, hk.data(), sizeof(OrbKeypoint) * (size_t)kept, cudaMemcpyHostToDevice, s));
  {
    ProfScope ps(c, s, SLAMB200_K_ORB_DESC);
    launch_orb_blur_gray((const uint8_t*)L.orb_gray.p, rows, cols, (float*)L.orb_rowf.p, (uint8_t*)L.orb_blur.p, s);
    launch_orb_desc((const uint8_t*)L.orb_blur.p, cols, (const OrbKeypoint*)L.orb_kp.p, kept, n_pad,
                    (uint8_t*)L.orb_desc.p, s);
  }
  CU(cudaGetLastError());
  if (desc && kept > 0)
    CU(cudaMemcpyAsync(desc, L.orb_desc.p, (size_t)kept * 32, cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(L.done, s));
  if (resident) {
    rc = desc_create(c, SLAMB200_DESC_U8X32, L.orb_desc.p, kept, 32, true, s, false, resident);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(s, (*resident)->ready, 0));
  }
  CU(cudaStreamSynchronize(s));
  return SLAMB200_OK;
}

// ---- linear triangulation -----------------------------------------------------
extern "C" int slamb200_triangulate(slamb200_ctx* c, const double P1[12], const double P2[12],
                                    const float* pts1, const float* pts2, int M, double* points4d,
                                    double* points3d) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  if (M < 0) return fail(SLAMB200_ERR_INVALID, "M negative");
  if (!P1 || !P2) return fail(SLAMB200_ERR_INVALID, "projection matrix is NULL");
  if (M == 0) return SLAMB200_OK;
  if (!pts1 || !pts2) return fail(SLAMB200_ERR_INVALID, "NULL input");
  if (!points4d && !points3d) return fail(SLAMB200_ERR_INVALID, "NULL output");
  TriParams tp;
  memcpy(tp.P[0], P1, sizeof(double) * 12);
  memcpy(tp.P[1], P2, sizeof(double) * 12);
  CU(cudaSetDevice(c->device));
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  int rc;
  if ((rc = buf_reserve(c, L.p1, (size_t)M * 8, s))) return rc;
  if ((rc = buf_reserve(c, L.p2, (size_t)M * 8, s))) return rc;
  if ((rc = buf_reserve(c, L.npts, (size_t)M * 56, s))) return rc;   // 4 x M + M x 3 doubles
  double* d4 = (double*)L.npts.p;
  double* d3 = d4 + (size_t)4 * M;
  CU(cudaMemcpyAsync(L.p1.p, pts1, (size_t)M * 8, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(L.p2.p, pts2, (size_t)M * 8, cudaMemcpyHostToDevice, s));
  launch_triangulate((const float2*)L.p1.p, (const float2*)L.p2.p, M, tp, d4, d3, s);
  CU(cudaGetLastError());
  if (points4d) CU(cudaMemcpyAsync(points4d, d4, sizeof(double) * 4 * (size_t)M, cudaMemcpyDeviceToHost, s));
  if (points3d) CU(cudaMemcpyAsync(points3d, d3, sizeof(double) * 3 * (size_t)M, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SLAMB200_OK;
}

// ---- keypoint coordinates + chained batch scoring ---------------------------------------------
extern "C" int slamb200_upload_pts(slamb200_ctx* c, const float* xy, int n, size_t stride,
                                   slamb200_pts** out) {
  if (!c || !out) return fail(SLAMB200_ERR_INVALID, "upload_pts: NULL argument");
  *out = nullptr;
  if (n < 0 || (n > 0 && !xy)) return fail(SLAMB200_ERR_INVALID, "upload_pts: bad xy/n");
  if (stride == 0) stride = 8;
  if (stride < 8) return fail(SLAMB200_ERR_INVALID, "upload_pts: stride < 8");
  CU(cudaSetDevice(c->device));
  slamb200_pts* p = (slamb200_pts*)calloc(1, sizeof(slamb200_pts));
  if (!p) return fail(SLAMB200_ERR_NOMEM, "host allocation failed");
  p->n = n;
  LaneGuard g(c);
  cudaStream_t s = g.lane().stream;
  int rc = dev_alloc(c, (void**)&p->xy, (size_t)(n > 0 ? n : 1) * 8, s);
  if (rc) { free(p); return rc; }
  cudaError_t e = cudaSuccess;
  if (n > 0) e = cudaMemcpy2DAsync(p->xy, 8, xy, stride, 8, n, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(p->ready, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    cudaFreeAsync(p->xy, s);
    free(p);
    return fail(SLAMB200_ERR_CUDA, "upload_pts: %s", cudaGetErrorString(e));
  }
  *out = p;
  return SLAMB200_OK;
}

extern "C" int slamb200_free_pts(slamb200_ctx* c, slamb200_pts* p) {
  if (!p) return SLAMB200_OK;
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  cudaSetDevice(c->device);
  if (p->xy) free_behind_lanes(c, p->xy, 0);  // the upload was synchronised: only readers remain
  if (p->ready) cudaEventDestroy(p->ready);
  free(p);
  return SLAMB200_OK;
}

extern "C" int slamb200_score_batch_enqueue(slamb200_ctx* c, const slamb200_pts* query_pts,
                                            const slamb200_pts* const* train_pts,
                                            const double K[4], const double* E, int H,
                                            double threshold_px, void* stream) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  if (H <= 0 || !E) return fail(SLAMB200_ERR_INVALID, "H <= 0 or E is NULL");
  ScoreParams sp;
  int rc = make_score_params(K, threshold_px, &sp);
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> lk(c->batch_mu);
  Lane& L = c->lanes[0];
  cudaStream_t s = stream ? (cudaStream_t)stream : L.stream;
  const int P = L.b_pairs, cap = L.b_cap;
  if (P <= 0) return fail(SLAMB200_ERR_INVALID, "no batch has been enqueued on this context");
  if (!query_pts || !train_pts) return fail(SLAMB200_ERR_INVALID, "NULL keypoint sets");
  if (query_pts->n < L.b_nq) return fail(SLAMB200_ERR_INVALID, "query keypoints < query rows");
  // train_pts must hold one set per pair of the batch (b_pairs entries), each covering every
  // trainIdx the matcher can have produced: the gather indexes it with trainIdx < rows(train p)
  for (int p = 0; p < P; p++) {
    if (!train_pts[p]) return fail(SLAMB200_ERR_INVALID, "train keypoint set %d is NULL", p);
    if (train_pts[p]->n < L.b_tn[(size_t)p])
      return fail(SLAMB200_ERR_INVALID, "train keypoint set %d has %d points, its descriptor set %d rows",
                  p, train_pts[p]->n, L.b_tn[(size_t)p]);
  }
  if (!L.b_stream_set || L.b_stream != s) CU(cudaStreamWaitEvent(s, L.done, 0));   // the match batch ran on another stream
  L.b_stream = s;   // the next lane-0 call is ordered behind THIS stream's work (L.done is re-recorded below)
  L.b_stream_set = true;
  // E (unless it is already resident on this device) and the train keypoint pointer table go
  // through the pinned staging area
  const size_t e_bytes = sizeof(double) * 9 * (size_t)P * H;
  const size_t tab_bytes = sizeof(void*) * (size_t)P;
  bool e_on_device = false;
  {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, E) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
        pa.device == c->device)
      e_on_device = true;
    cudaGetLastError();
  }
  if ((rc = stage_reserve(L, (e_on_device ? 0 : e_bytes) + tab_bytes))) return rc;
  const size_t tab_off = e_on_device ? 0 : e_bytes;
  if (!e_on_device) memcpy(L.h_stage, E, e_bytes);
  const float2** tab = (const float2**)((char*)L.h_stage + tab_off);
  for (int p = 0; p < P; p++) {
    if (!train_pts[p]) return fail(SLAMB200_ERR_INVALID, "train keypoint set %d is NULL", p);
    tab[p] = train_pts[p]->xy;
  }
  const double* E_dev = E;
  if (!e_on_device) {
    if ((rc = buf_reserve(c, L.E, e_bytes, s))) return rc;
    E_dev = (const double*)L.E.p;
  }
  if ((rc = buf_reserve(c, L.txy, tab_bytes, s))) return rc;
  if ((rc = buf_reserve(c, L.npts, (size_t)P * cap * 32, s))) return rc;
  if ((rc = buf_reserve(c, L.counts, sizeof(int32_t) * (size_t)P * H, s))) return rc;
  if ((rc = buf_reserve(c, L.best, sizeof(int32_t) * (size_t)P, s))) return rc;
  if ((rc = buf_reserve(c, L.mask, (size_t)P * cap, s))) return rc;
  if (!e_on_device) CU(cudaMemcpyAsync(L.E.p, L.h_stage, e_bytes, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(L.txy.p, tab, tab_bytes, cudaMemcpyHostToDevice, s));
  CU(cudaEventRecord(L.stage_free, s));
  CU(cudaStreamWaitEvent(s, query_pts->ready, 0));
  for (int p = 0; p < P; p++) CU(cudaStreamWaitEvent(s, train_pts[p]->ready, 0));
  launch_gather_normalize(query_pts->xy, (const float2* const*)L.txy.p,
                          (const slamb200_dmatch*)L.out.p, cap, (const int32_t*)L.n_out.p, P, sp,
                          (double4*)L.npts.p, s);
  {
    ProfScope ps(c, s, SLAMB200_K_RANSAC);
    launch_score_counts((const double4*)L.npts.p, nullptr, (const int32_t*)L.n_out.p, cap, E_dev, H,
                        P, sp, (int32_t*)L.counts.p, s);
  }
  launch_score_best((const int32_t*)L.counts.p, H, P, 4, (int32_t*)L.best.p, s);
  launch_score_mask((const double4*)L.npts.p, nullptr, (const int32_t*)L.n_out.p, cap, E_dev, H, P,
                    (const int32_t*)L.best.p, sp, (uint8_t*)L.mask.p, s);
  CU(cudaGetLastError());
  CU(cudaEventRecord(L.done, s));
  L.s_H = H;
  L.s_pairs = P;
  return SLAMB200_OK;
}

extern "C" int slamb200_batch_scores_fetch(slamb200_ctx* c, int32_t* counts, int32_t* best,
                                           uint8_t* best_mask, int mask_cap, void* stream) {
  if (!c) return fail(SLAMB200_ERR_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> lk(c->batch_mu);
  Lane& L = c->lanes[0];
  cudaStream_t s = stream ? (cudaStream_t)stream : L.stream;
  const int P = L.s_pairs, H = L.s_H;
  if (P <= 0) return fail(SLAMB200_ERR_INVALID, "no scores have been enqueued on this context");
  if (best_mask && mask_cap < L.b_cap) return fail(SLAMB200_ERR_INVALID, "mask_cap %d < %d", mask_cap, L.b_cap);
  CU(cudaStreamWaitEvent(s, L.done, 0));
  if (counts) CU(cudaMemcpyAsync(counts, L.counts.p, sizeof(int32_t) * (size_t)P * H, cudaMemcpyDeviceToHost, s));
  if (best) CU(cudaMemcpyAsync(best, L.best.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, s));
  if (best_mask) {
    CU(cudaMemcpy2DAsync(best_mask, mask_cap, L.mask.p, L.b_cap, L.b_cap, P, cudaMemcpyDeviceToHost, s));
  }
  CU(cudaStreamSynchronize(s));
  return SLAMB200_OK;
}

// ---- debug hooks (not part of the public header) ----------------------------------------------
// SIFT's working image of a frame (gray -> float -> 13-tap Gaussian) as the descriptor kernel reads
// it: out = rows x cols floats.  For the bit-exactness test against cv2.GaussianBlur.
extern "C" int slamb200_dbg_sift_base(slamb200_ctx* c, const uint8_t* image, int rows, int cols, int channels,
                                      size_t step, float* out) {
  if (!c || !image || !out || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3)) return SLAMB200_ERR_INVALID;
  if (step == 0) step = (size_t)cols * channels;
  CU(cudaSetDevice(c->device));
  if (sift_gauss_upload() != 0) return fail(SLAMB200_ERR_CUDA, "kernel table upload failed");
  LaneGuard g(c);
  Lane& L = g.lane();
  cudaStream_t s = L.stream;
  const size_t px = (size_t)rows * cols;
  int rc;
  if ((rc = buf_reserve(c, L.orb_img, (size_t)rows * step, s))) return rc;
  if ((rc = buf_reserve(c, L.orb_gray, px, s))) return rc;
  if ((rc = buf_reserve(c, L.sift_rowf, px * sizeof(float), s))) return rc;
  if ((rc = buf_reserve(c, L.sift_base, px * sizeof(float), s))) return rc;
  CU(cudaMemcpyAsync(L.orb_img.p, image, (size_t)rows * step, cudaMemcpyHostToDevice, s));
  launch_orb_gray((const uint8_t*)L.orb_img.p, rows, cols, channels, step, (uint8_t*)L.orb_gray.p, s);
  launch_sift_base((const uint8_t*)L.orb_gray.p, rows, cols, (float*)L.sift_rowf.p, (float*)L.sift_base.p, s);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, L.sift_base.p, px * sizeof(float), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SLAMB200_OK;
}

extern "C" int slamb200_dbg_set_sub_batch(slamb200_ctx* c, int pairs) {
  if (!c) return SLAMB200_ERR_INVALID;
  c->sub_batch = pairs;
  return SLAMB200_OK;
}

extern "C" int slamb200_dbg_set_tc(slamb200_ctx* c, int on) {
  if (!c) return SLAMB200_ERR_INVALID;
  c->use_tc = on ? 1 : 0;
  return SLAMB200_OK;
}

extern "C" int slamb200_dbg_set_fused_tail(slamb200_ctx* c, int on) {
  if (!c) return SLAMB200_ERR_INVALID;
  c->fused_tail = on;   // see slamb200_ctx::fused_tail (-1 = by batch size)
  return SLAMB200_OK;
}

extern "C" int slamb200_dbg_set_tc_orb(slamb200_ctx* c, int on) {
  if (!c) return SLAMB200_ERR_INVALID;
  c->use_tc_orb = on ? 1 : 0;
  return SLAMB200_OK;
}

// Raw fp32 accumulators (d^2/2) of the first 256 x 256 tile of (query, train): out[256*256].
extern "C" int slamb200_dbg_tc_tile(slamb200_ctx* c, const slamb200_desc* q, const slamb200_desc* t,
                                    float* out) {
  if (!c || !q || !t || !out) return SLAMB200_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  LaneGuard g(c);
  Lane& L = g.lane();
  float* d = nullptr;
  int rc = dev_alloc(c, (void**)&d, 256 * 256 * 4, L.stream);
  if (rc) return rc;
  CU(cudaMemsetAsync(d, 0, 256 * 256 * 4, L.stream));
  L.dbg = d;
  const slamb200_desc* tt[1] = {t};
  rc = enqueue_batch(c, L, L.stream, q->kind == SLAMB200_DESC_U8X32 ? SLAMB200_ORB_BF : SLAMB200_SIFT_BF, q, tt, 1, 0.7);
  L.dbg = nullptr;
  if (rc == SLAMB200_OK) {
    cudaError_t e = cudaMemcpyAsync(out, d, 256 * 256 * 4, cudaMemcpyDeviceToHost, L.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(L.stream);
    if (e != cudaSuccess) rc = fail(SLAMB200_ERR_CUDA, "dbg_tc_tile: %s", cudaGetErrorString(e));
  }
  cudaFreeAsync(d, L.stream);
  return rc;
}

// Rows of the last enqueued batch (lane 0) that the general-float certificate sent to the exact
// fallback.  Synchronises.
extern "C" int slamb200_dbg_last_fallback_rows(slamb200_ctx* c) {
  if (!c) return -1;
  Lane& L = c->lanes[0];
  if (!L.status) return -1;
  int32_t v = -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(&v, L.status + 1002, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}
