// device_set.cu -- one process, every GPU of the box (include/slamb200.h "device set";
// SURVEY.md 8e).  The reference walks the framesBatchSize window from host threads of ONE
// process (src/mainModule/cycleProcessing/batch.cpp:162-226); a device set puts all GPUs of the
// box behind that process: per-device contexts, the query frame's prepared set replicated by peer
// copies over NVLink, train frames resident on the device that matches them, one enqueue per
// device without waiting, results gathered once.  Built on the single-device entry points only.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

#define SET_MAX 16

// One host thread per member (the reference's search strides its batch with `threadsCount` host
// threads, batch.cpp:181-201): the members' enqueues and fetches run side by side instead of one
// after the other on the calling thread.
struct SetWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = true, stop = false;
  int rc = 0;
  void start() {
    th = std::thread([this] {
      for (;;) {
        std::function<int()> f;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [this] { return has_job || stop; });
          if (stop) return;
          f = std::move(job);
          has_job = false;
        }
        const int r = f();
        {
          std::lock_guard<std::mutex> lk(mu);
          rc = r;
          done = true;
        }
        cv.notify_all();
      }
    });
  }
  void submit(std::function<int()> f) {
    {
      std::lock_guard<std::mutex> lk(mu);
      job = std::move(f);
      has_job = true;
      done = false;
    }
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [this] { return done; });
    return rc;
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    if (th.joinable()) th.join();
  }
};

struct slamb200_set {
  int n = 0;
  SetWorker worker[SET_MAX];   // members 1 .. n-1 (member 0 runs on the calling thread)
  int dev[SET_MAX];
  slamb200_ctx* ctx[SET_MAX];
  cudaStream_t stream[SET_MAX];
  cudaEvent_t ev0[SET_MAX], ev1[SET_MAX];
  std::mutex mu;
  // the last enqueued batch: which global pairs each member holds, in its local order
  std::vector<int> part[SET_MAX];
  int b_pairs = 0, b_nq = 0;
  std::vector<slamb200_dmatch> tmp[SET_MAX];   // scatter buffers for members whose pairs are not contiguous
  std::vector<int> tmp_n[SET_MAX];
  std::string err[SET_MAX];
};

struct slamb200_mdesc {
  int member;                 // >= 0: resident on that member only; -1: replicated
  int rows, kind;
  slamb200_desc* d[SET_MAX];  // per member (NULL where not resident)
};

static int sfail(int code, const char* msg) { return set_error(code, msg); }

extern "C" int slamb200_set_init(int n_devices, slamb200_set** out) {
  if (!out) return SLAMB200_ERR_INVALID;
  *out = nullptr;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) return SLAMB200_ERR_CUDA;
  int n = n_devices <= 0 ? have : n_devices;
  if (n > have || n > SET_MAX) return sfail(SLAMB200_ERR_INVALID, "set_init: more devices requested than present");
  slamb200_set* s = new (std::nothrow) slamb200_set();
  if (!s) return SLAMB200_ERR_NOMEM;
  for (int i = 0; i < n; i++) {
    s->dev[i] = i;
    int rc = slamb200_init(i, &s->ctx[i]);
    if (rc == SLAMB200_OK && (cudaSetDevice(i) != cudaSuccess ||
                              cudaStreamCreateWithFlags(&s->stream[i], cudaStreamNonBlocking) != cudaSuccess ||
                              cudaEventCreate(&s->ev0[i]) != cudaSuccess || cudaEventCreate(&s->ev1[i]) != cudaSuccess))
      rc = SLAMB200_ERR_CUDA;
    if (rc != SLAMB200_OK) {
      for (int k = 0; k <= i; k++)
        if (s->ctx[k]) slamb200_shutdown(s->ctx[k]);
      delete s;
      return rc;
    }
    s->n = i + 1;
  }
  // direct peer copies (NVLink / NVSwitch) between every pair of members; where the topology does
  // not allow it the runtime stages the copy instead, results are the same
  for (int i = 0; i < n; i++) {
    cudaSetDevice(i);
    for (int k = 0; k < n; k++) {
      int can = 0;
      if (k != i && cudaDeviceCanAccessPeer(&can, i, k) == cudaSuccess && can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(k, 0);
        if (e != cudaSuccess) cudaGetLastError();   // already enabled (by the host application)
        // descriptor slabs are stream-ordered pool allocations: device i may read member k's pool
        ctx_grant_peer_access(s->ctx[k], i);
      }
    }
  }
  for (int i = 1; i < n; i++) s->worker[i].start();
  *out = s;
  return SLAMB200_OK;
}

// runs f(member) for every member, members 1.. on their worker threads, member 0 here
template <class F>
static int for_each_member(slamb200_set* s, F f) {
  static const bool threads = [] { const char* e = getenv("SLAMB200_SET_THREADS"); return !e || atoi(e) != 0; }();
  if (!threads) {   // everything on the calling thread, member after member
    int rc = SLAMB200_OK;
    for (int i = 0; i < s->n; i++) {
      const int r = f(i);
      if (rc == SLAMB200_OK) rc = r;
    }
    return rc;
  }
  for (int i = 1; i < s->n; i++)
    s->worker[i].submit([s, f, i] {
      const int r = f(i);
      if (r != SLAMB200_OK) s->err[i] = slamb200_last_error();   // the worker thread's message
      return r;
    });
  int rc = f(0);
  for (int i = 1; i < s->n; i++) {
    const int r = s->worker[i].wait();
    if (rc == SLAMB200_OK && r != SLAMB200_OK) rc = set_error(r, s->err[i].c_str());
  }
  return rc;
}

extern "C" int slamb200_set_shutdown(slamb200_set* s) {
  if (!s) return SLAMB200_OK;
  for (int i = 1; i < s->n; i++) s->worker[i].shutdown();
  for (int i = 0; i < s->n; i++) {
    cudaSetDevice(s->dev[i]);
    cudaStreamSynchronize(s->stream[i]);
    cudaEventDestroy(s->ev0[i]);
    cudaEventDestroy(s->ev1[i]);
    cudaStreamDestroy(s->stream[i]);
    slamb200_shutdown(s->ctx[i]);
  }
  delete s;
  return SLAMB200_OK;
}

extern "C" int slamb200_set_devices(const slamb200_set* s) { return s ? s->n : 0; }
extern "C" slamb200_ctx* slamb200_set_ctx(slamb200_set* s, int i) {
  return (s && i >= 0 && i < s->n) ? s->ctx[i] : nullptr;
}
extern "C" int slamb200_set_owner(const slamb200_set* s, int i, int n) {
  if (!s || n <= 0 || i < 0) return 0;
  if (i >= n) i = n - 1;
  return (int)(((long long)i * s->n) / n);
}

extern "C" int slamb200_set_upload(slamb200_set* s, int member, int kind, const void* rows, int n,
                                   size_t row_stride, slamb200_mdesc** out) {
  if (!s || !out) return SLAMB200_ERR_INVALID;
  *out = nullptr;
  if (member >= s->n) return sfail(SLAMB200_ERR_INVALID, "set_upload: no such member");
  slamb200_mdesc* m = (slamb200_mdesc*)calloc(1, sizeof(slamb200_mdesc));
  if (!m) return SLAMB200_ERR_NOMEM;
  m->member = member < 0 ? -1 : member;
  m->rows = n;
  m->kind = kind;
  const int first = member < 0 ? 0 : member;
  int rc = slamb200_upload_desc(s->ctx[first], kind, rows, n, row_stride, &m->d[first]);
  // replication: the PREPARED slab (tensor-core operands, norms, byte copy, flags) travels
  // device to device; no second pass over the host rows, no second prep kernel
  for (int i = 0; rc == SLAMB200_OK && member < 0 && i < s->n; i++)
    if (i != first) rc = slamb200_desc_localize(s->ctx[i], m->d[first], &m->d[i]);
  if (rc != SLAMB200_OK) {
    for (int i = 0; i < s->n; i++)
      if (m->d[i]) slamb200_free_desc(s->ctx[i], m->d[i]);
    free(m);
    return rc;
  }
  *out = m;
  return SLAMB200_OK;
}

extern "C" int slamb200_set_free_desc(slamb200_set* s, slamb200_mdesc* m) {
  if (!m) return SLAMB200_OK;
  if (!s) return SLAMB200_ERR_INVALID;
  for (int i = 0; i < s->n; i++)
    if (m->d[i]) slamb200_free_desc(s->ctx[i], m->d[i]);
  free(m);
  return SLAMB200_OK;
}

extern "C" int slamb200_mdesc_rows(const slamb200_mdesc* m) { return m ? m->rows : -1; }

static int set_enqueue(slamb200_set* s, int matcher, const slamb200_mdesc* q,
                       const slamb200_mdesc* const* trains, int n_pairs, double ratio) {
  if (!q || q->member >= 0) return sfail(SLAMB200_ERR_INVALID, "set_match_batch: the query set must be replicated");
  if (n_pairs < 0 || (n_pairs > 0 && !trains)) return SLAMB200_ERR_INVALID;
  for (int i = 0; i < s->n; i++) s->part[i].clear();
  // a train frame is matched where it lives; a replicated one on the member with the least work
  for (int p = 0; p < n_pairs; p++) {
    if (!trains[p]) return sfail(SLAMB200_ERR_INVALID, "set_match_batch: NULL train set");
    if (trains[p]->member >= 0) s->part[trains[p]->member].push_back(p);
  }
  for (int p = 0; p < n_pairs; p++) {
    if (trains[p]->member >= 0) continue;
    int best = 0;
    for (int i = 1; i < s->n; i++)
      if (s->part[i].size() < s->part[best].size()) best = i;
    s->part[best].push_back(p);
  }
  s->b_pairs = n_pairs;
  s->b_nq = q->rows;
  // enqueue only, every member from its own host thread: all of them start on their share before
  // the first one is fetched
  return for_each_member(s, [=](int i) -> int {
    std::vector<const slamb200_desc*> local;
    for (int p : s->part[i]) local.push_back(trains[p]->d[i]);
    cudaSetDevice(s->dev[i]);
    cudaEventRecord(s->ev0[i], s->stream[i]);
    const int rc = slamb200_match_batch_enqueue(s->ctx[i], matcher, q->d[i], local.data(), (int)local.size(), ratio,
                                                (void*)s->stream[i]);
    cudaEventRecord(s->ev1[i], s->stream[i]);
    return rc;
  });
}

static int set_fetch(slamb200_set* s, slamb200_dmatch* out, int cap, int* n_out, float* device_ms) {
  if (s->b_pairs > 0 && !n_out) return SLAMB200_ERR_INVALID;
  return for_each_member(s, [=](int i) -> int {
    const std::vector<int>& part = s->part[i];
    if (device_ms) device_ms[i] = 0.f;
    if (part.empty()) return SLAMB200_OK;
    cudaSetDevice(s->dev[i]);
    bool contiguous = true;
    for (size_t k = 1; k < part.size(); k++) contiguous = contiguous && part[k] == part[k - 1] + 1;
    int rc;
    if (contiguous) {   // the usual placement: the member's results land where they belong
      rc = slamb200_batch_fetch(s->ctx[i], out ? out + (size_t)part[0] * cap : nullptr, cap, n_out + part[0],
                                (void*)s->stream[i]);
    } else {
      std::vector<slamb200_dmatch>& tmp = s->tmp[i];
      std::vector<int>& tmp_n = s->tmp_n[i];
      tmp.resize(part.size() * (size_t)cap);
      tmp_n.resize(part.size());
      rc = slamb200_batch_fetch(s->ctx[i], tmp.data(), cap, tmp_n.data(), (void*)s->stream[i]);
      for (size_t k = 0; rc == SLAMB200_OK && k < part.size(); k++) {
        n_out[part[k]] = tmp_n[k];
        if (out) memcpy(out + (size_t)part[k] * cap, tmp.data() + k * (size_t)cap, sizeof(slamb200_dmatch) * (size_t)tmp_n[k]);
      }
    }
    if (rc != SLAMB200_OK) return rc;
    if (device_ms) {
      cudaEventSynchronize(s->ev1[i]);
      cudaEventElapsedTime(&device_ms[i], s->ev0[i], s->ev1[i]);
    }
    return SLAMB200_OK;
  });
}

extern "C" int slamb200_set_match_batch_enqueue(slamb200_set* s, int matcher, const slamb200_mdesc* q,
                                                const slamb200_mdesc* const* trains, int n_pairs, double ratio) {
  if (!s) return SLAMB200_ERR_INVALID;
  std::lock_guard<std::mutex> lk(s->mu);
  return set_enqueue(s, matcher, q, trains, n_pairs, ratio);
}

extern "C" int slamb200_set_batch_fetch(slamb200_set* s, slamb200_dmatch* out, int cap, int* n_out, float* device_ms) {
  if (!s) return SLAMB200_ERR_INVALID;
  std::lock_guard<std::mutex> lk(s->mu);
  return set_fetch(s, out, cap, n_out, device_ms);
}

extern "C" int slamb200_set_match_batch(slamb200_set* s, int matcher, const slamb200_mdesc* q,
                                        const slamb200_mdesc* const* trains, int n_pairs, double ratio,
                                        slamb200_dmatch* out, int cap, int* n_out) {
  if (!s) return SLAMB200_ERR_INVALID;
  std::lock_guard<std::mutex> lk(s->mu);
  int rc = set_enqueue(s, matcher, q, trains, n_pairs, ratio);
  if (rc != SLAMB200_OK) return rc;
  return set_fetch(s, out, cap, n_out, nullptr);
}
