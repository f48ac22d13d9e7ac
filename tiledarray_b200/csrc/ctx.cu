// ctx.cu — context, streams, stream-ordered device memory pool, staging, error plumbing.
// B200-native replacement for TiledArray's device::Env (reference: src/TiledArray/external/
// device.h:422-441 streams, :536-549 rank->device, :566-605 Umpire pools): tiles live in plain
// device memory served by a CUDA stream-ordered pool; unified memory is not used anywhere.
#include "common.h"

static thread_local char g_err[1024] = "";

void tadev_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* tadev_last_error(void) { return g_err; }

GemmTimingHook& tadev_gemm_timing_hook() {
  static thread_local GemmTimingHook hook;
  return hook;
}
extern "C" const char* tadev_version(void) { return "tadev 0.2 (sm_100a)"; }

extern "C" int tadev_device_count(int* n) {
  TADEV_REQUIRE(n, "tadev_device_count: null out");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  *n = c;
  return TADEV_OK;
}

extern "C" int tadev_init(int device, size_t pool_bytes, tadev_ctx** out) {
  TADEV_REQUIRE(out, "tadev_init: null out");
  int ndev = 0;
  tadev_device_count(&ndev);
  if (ndev == 0) {
    tadev_set_error("tadev_init: no CUDA device visible; libtadev has no CPU path");
    return TADEV_ENODEVICE;
  }
  TADEV_REQUIRE(device >= 0 && device < ndev, "tadev_init: device %d out of range [0,%d)", device, ndev);
  TADEV_CHECK_CUDA(cudaSetDevice(device));
  tadev_ctx* ctx = new tadev_ctx();
  ctx->device = device;
  TADEV_CHECK_CUDA(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device));
  TADEV_CHECK_CUDA(cudaDeviceGetDefaultMemPool(&ctx->pool, device));
  // keep freed tiles cached in the pool (a tile pool, not malloc/free per tile)
  uint64_t thresh = pool_bytes ? (uint64_t)pool_bytes : UINT64_MAX;
  TADEV_CHECK_CUDA(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  int nstreams = 3;  // TA_DEVICE_NUM_STREAMS default (external/device.h:422-441)
  if (const char* e = getenv("TA_DEVICE_NUM_STREAMS")) nstreams = atoi(e) > 0 ? atoi(e) : nstreams;
  if (const char* e = getenv("TADEV_GEMM_GENERIC")) ctx->force_generic_gemm = atoi(e) != 0;
  int lo = 0, hi = 0;
  TADEV_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  ctx->streams.resize(nstreams);
  for (int i = 0; i < nstreams; ++i)
    TADEV_CHECK_CUDA(cudaStreamCreateWithPriority(&ctx->streams[i], cudaStreamNonBlocking, lo));
  for (int i = 0; i < 2; ++i)
    TADEV_CHECK_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream[i], cudaStreamNonBlocking, hi));
  TADEV_CHECK_CUDA(cudaStreamCreateWithPriority(&ctx->desc_stream, cudaStreamNonBlocking, hi));
  *out = ctx;
  return TADEV_OK;
}

extern "C" int tadev_finalize(tadev_ctx* ctx) {
  if (!ctx) return TADEV_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  tadev_comm_destroy(ctx);
  tadev_tmap_cache_destroy(ctx);
  for (auto& pr : ctx->staging) {
    for (int i = 0; i < StagingRing::kSlots; ++i) {
      if (pr.second->h[i]) cudaFreeHost(pr.second->h[i]);
      if (pr.second->d[i]) cudaFree(pr.second->d[i]);
      if (pr.second->done[i]) cudaEventDestroy(pr.second->done[i]);
      if (pr.second->uploaded[i]) cudaEventDestroy(pr.second->uploaded[i]);
    }
  }
  for (auto s : ctx->streams) cudaStreamDestroy(s);
  for (int i = 0; i < 2; ++i)
    if (ctx->comm_stream[i]) cudaStreamDestroy(ctx->comm_stream[i]);
  if (ctx->desc_stream) cudaStreamDestroy(ctx->desc_stream);
  delete ctx;
  return TADEV_OK;
}

extern "C" int tadev_num_streams(tadev_ctx* ctx, int* n) {
  TADEV_REQUIRE(ctx && n, "tadev_num_streams: null");
  *n = (int)ctx->streams.size();
  return TADEV_OK;
}
extern "C" int tadev_get_stream(tadev_ctx* ctx, int i, tadev_stream* out) {
  TADEV_REQUIRE(ctx && out, "tadev_get_stream: null");
  TADEV_REQUIRE(i >= 0 && i < (int)ctx->streams.size(), "tadev_get_stream: index %d out of range", i);
  *out = (tadev_stream)ctx->streams[i];
  return TADEV_OK;
}
extern "C" int tadev_stream_for(tadev_ctx* ctx, uint64_t ordinal, tadev_stream* out) {
  TADEV_REQUIRE(ctx && out, "tadev_stream_for: null");
  *out = (tadev_stream)ctx->streams[ordinal % ctx->streams.size()];
  return TADEV_OK;
}
extern "C" int tadev_stream_sync(tadev_ctx* ctx, tadev_stream s) {
  TADEV_REQUIRE(ctx, "tadev_stream_sync: null ctx");
  TADEV_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)s));
  return TADEV_OK;
}

extern "C" int tadev_alloc(tadev_ctx* ctx, size_t bytes, void** d_ptr, tadev_stream s) {
  TADEV_REQUIRE(ctx && d_ptr, "tadev_alloc: null");
  if (bytes == 0) {
    *d_ptr = nullptr;
    return TADEV_OK;
  }
  cudaError_t e = cudaMallocFromPoolAsync(d_ptr, bytes, ctx->pool, (cudaStream_t)s);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    tadev_set_error("tadev_alloc: out of device memory (%zu bytes)", bytes);
    return TADEV_ENOMEM;
  }
  TADEV_CHECK_CUDA(e);
  return TADEV_OK;
}
extern "C" int tadev_free(tadev_ctx* ctx, void* d_ptr, tadev_stream s) {
  TADEV_REQUIRE(ctx, "tadev_free: null ctx");
  if (d_ptr) TADEV_CHECK_CUDA(cudaFreeAsync(d_ptr, (cudaStream_t)s));
  return TADEV_OK;
}
extern "C" int tadev_memcpy_h2d(tadev_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, tadev_stream s) {
  TADEV_REQUIRE(ctx, "tadev_memcpy_h2d: null ctx");
  if (bytes) TADEV_CHECK_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)s));
  return TADEV_OK;
}
extern "C" int tadev_memcpy_d2h(tadev_ctx* ctx, void* h_dst, const void* d_src, size_t bytes, tadev_stream s) {
  TADEV_REQUIRE(ctx, "tadev_memcpy_d2h: null ctx");
  if (bytes) TADEV_CHECK_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)s));
  return TADEV_OK;
}
extern "C" int tadev_memset(tadev_ctx* ctx, void* d_dst, int byte, size_t bytes, tadev_stream s) {
  TADEV_REQUIRE(ctx, "tadev_memset: null ctx");
  if (bytes) TADEV_CHECK_CUDA(cudaMemsetAsync(d_dst, byte, bytes, (cudaStream_t)s));
  return TADEV_OK;
}
extern "C" int tadev_host_alloc(size_t bytes, void** h_ptr) {
  TADEV_REQUIRE(h_ptr, "tadev_host_alloc: null");
  TADEV_CHECK_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 1));
  return TADEV_OK;
}
extern "C" int tadev_host_free(void* h_ptr) {
  if (h_ptr) TADEV_CHECK_CUDA(cudaFreeHost(h_ptr));
  return TADEV_OK;
}

extern "C" int tadev_event_create(tadev_ctx* ctx, void** ev) {
  TADEV_REQUIRE(ctx && ev, "tadev_event_create: null");
  cudaEvent_t e;
  TADEV_CHECK_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return TADEV_OK;
}
extern "C" int tadev_event_record(tadev_ctx* ctx, void* ev, tadev_stream s) {
  TADEV_REQUIRE(ctx && ev, "tadev_event_record: null");
  TADEV_CHECK_CUDA(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)s));
  return TADEV_OK;
}
extern "C" int tadev_event_elapsed_ms(tadev_ctx* ctx, void* ev_start, void* ev_stop, float* ms) {
  TADEV_REQUIRE(ctx && ev_start && ev_stop && ms, "tadev_event_elapsed_ms: null");
  TADEV_CHECK_CUDA(cudaEventSynchronize((cudaEvent_t)ev_stop));
  TADEV_CHECK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)ev_start, (cudaEvent_t)ev_stop));
  return TADEV_OK;
}
extern "C" int tadev_event_destroy(tadev_ctx* ctx, void* ev) {
  TADEV_REQUIRE(ctx, "tadev_event_destroy: null ctx");
  if (ev) TADEV_CHECK_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return TADEV_OK;
}

// ---- asynchronous completion / cross-stream ordering for the tile plug-in --------------------------------------
// (reference contract: tile ops enqueue work and return; the runtime completes the task from a host function on
// the stream, external/device.h:847-875 sync_madness_task_with, madness::add_device_task; reduce_task.h:460-486)
extern "C" int tadev_sync_event_create(tadev_ctx* ctx, void** ev) {
  TADEV_REQUIRE(ctx && ev, "tadev_sync_event_create: null");
  cudaEvent_t e;
  TADEV_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *ev = (void*)e;
  return TADEV_OK;
}
extern "C" int tadev_stream_wait_event(tadev_ctx* ctx, tadev_stream s, void* ev) {
  TADEV_REQUIRE(ctx && ev, "tadev_stream_wait_event: null");
  TADEV_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)ev, 0));
  return TADEV_OK;
}
extern "C" int tadev_event_query(tadev_ctx* ctx, void* ev, int* done) {
  TADEV_REQUIRE(ctx && ev && done, "tadev_event_query: null");
  cudaError_t e = cudaEventQuery((cudaEvent_t)ev);
  if (e == cudaErrorNotReady) { *done = 0; return TADEV_OK; }
  TADEV_CHECK_CUDA(e);
  *done = 1;
  return TADEV_OK;
}
extern "C" int tadev_event_sync(tadev_ctx* ctx, void* ev) {
  TADEV_REQUIRE(ctx && ev, "tadev_event_sync: null");
  TADEV_CHECK_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return TADEV_OK;
}
extern "C" int tadev_stream_add_callback(tadev_ctx* ctx, tadev_stream s, tadev_host_fn fn, void* user) {
  TADEV_REQUIRE(ctx && fn, "tadev_stream_add_callback: null");
  TADEV_CHECK_CUDA(cudaLaunchHostFunc((cudaStream_t)s, fn, user));
  return TADEV_OK;
}
extern "C" int tadev_memcpy_d2d(tadev_ctx* ctx, void* d_dst, const void* d_src, size_t bytes, tadev_stream s) {
  TADEV_REQUIRE(ctx, "tadev_memcpy_d2d: null ctx");
  if (bytes) TADEV_CHECK_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
  return TADEV_OK;
}

extern "C" int tadev_launch_count(tadev_ctx* ctx, int64_t* n) {
  TADEV_REQUIRE(ctx && n, "tadev_launch_count: null");
  *n = ctx->launches.load();
  return TADEV_OK;
}

int tadev_stage_upload(tadev_ctx* ctx, cudaStream_t s, void* d, const void* h, size_t bytes, cudaEvent_t uploaded) {
  TADEV_CHECK_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->desc_stream));
  TADEV_CHECK_CUDA(cudaEventRecord(uploaded, ctx->desc_stream));
  TADEV_CHECK_CUDA(cudaStreamWaitEvent(s, uploaded, 0));
  return TADEV_OK;
}

int StageLease::acquire(tadev_ctx* ctx_, cudaStream_t s_, size_t bytes) {
  release();
  ctx = ctx_; s = s_;
  int i = -1;
  {
    std::unique_lock<std::mutex> lk(ctx->mu);
    for (auto& pr : ctx->staging)
      if (pr.first == s) ring = pr.second.get();
    if (!ring) {
      ctx->staging.emplace_back(s, std::unique_ptr<StagingRing>(new StagingRing()));
      ring = ctx->staging.back().second.get();
    }
    StagingRing* rg = ring;
    auto free_slot = [rg]() {
      for (int n = 0; n < StagingRing::kSlots; ++n) {
        const int c = (rg->next + n) % StagingRing::kSlots;
        if (!rg->busy[c]) return c;
      }
      return -1;
    };
    ctx->stage_cv.wait(lk, [&] { return free_slot() >= 0; });
    i = free_slot();
    ring->busy[i] = true;
    ring->next = (i + 1) % StagingRing::kSlots;
  }
  slot = i;  // from here on the slot is ours: no lock needed
  if (!ring->done[i]) TADEV_CHECK_CUDA(cudaEventCreateWithFlags(&ring->done[i], cudaEventDisableTiming));
  else TADEV_CHECK_CUDA(cudaEventSynchronize(ring->done[i]));  // the kernel that last read this slot finished
  if (!ring->uploaded[i]) TADEV_CHECK_CUDA(cudaEventCreateWithFlags(&ring->uploaded[i], cudaEventDisableTiming));
  if (ring->cap[i] < bytes) {
    size_t cap = bytes + bytes / 2 + 4096;
    if (ring->h[i]) TADEV_CHECK_CUDA(cudaFreeHost(ring->h[i]));
    if (ring->d[i]) TADEV_CHECK_CUDA(cudaFree(ring->d[i]));
    ring->h[i] = ring->d[i] = nullptr;
    ring->cap[i] = 0;
    TADEV_CHECK_CUDA(cudaMallocHost(&ring->h[i], cap));
    TADEV_CHECK_CUDA(cudaMalloc(&ring->d[i], cap));
    ring->cap[i] = cap;
  }
  h = ring->h[i];
  d = ring->d[i];
  done = ring->done[i];
  uploaded = ring->uploaded[i];
  return TADEV_OK;
}

void StageLease::release() {
  if (slot < 0 || !ring) return;
  if (ring->done[slot]) cudaEventRecord(ring->done[slot], s);  // after the consuming kernel in stream order
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    ring->busy[slot] = false;
  }
  ctx->stage_cv.notify_all();
  slot = -1;
  ring = nullptr;
}
