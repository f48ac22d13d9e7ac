// common.h — internal declarations shared by the libtadev translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/tadev.h"

struct ncclComm;

void tadev_set_error(const char* fmt, ...);

#define TADEV_CHECK_CUDA(expr)                                                            \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      tadev_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return TADEV_ECUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define TADEV_REQUIRE(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      tadev_set_error(__VA_ARGS__); \
      return TADEV_EINVAL;         \
    }                              \
  } while (0)

// A grow-only device/pinned staging buffer pair used to ship descriptor lists to the device.
// One ring per stream slot so concurrent callers on distinct streams never share a buffer.
struct StagingRing {
  static constexpr int kSlots = 4;
  void* h[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  void* d[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap[kSlots] = {0, 0, 0, 0};
  cudaEvent_t done[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t uploaded[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  bool busy[kSlots] = {false, false, false, false};  // leased to a caller that has not recorded `done` yet
  int next = 0;
};

struct tadev_ctx {
  int device = -1;
  int num_sms = 0;
  cudaMemPool_t pool = nullptr;
  std::vector<cudaStream_t> streams;  // compute streams (non-blocking)
  cudaStream_t comm_stream[2] = {nullptr, nullptr};  // row / column panel broadcasts (high priority)
  cudaStream_t desc_stream = nullptr;  // descriptor uploads: never queued behind a caller stream's pending work
  std::atomic<int64_t> launches{0};
  std::mutex mu;  // guards staging rings
  std::condition_variable stage_cv;  // a leased staging slot was released
  std::vector<std::pair<cudaStream_t, std::unique_ptr<StagingRing>>> staging;
  // communicators
  ncclComm* world = nullptr;
  ncclComm* row_comm = nullptr;  // ranks sharing my grid row   (A column-panels travel here)
  ncclComm* col_comm = nullptr;  // ranks sharing my grid column (B row-panels travel here)
  int rank = 0, nranks = 1, Pr = 1, Pc = 1, my_r = 0, my_c = 0;
  // grouped-GEMM launch policy
  int gemm_sm_reserve = 0;          // SMs the persistent kernel leaves free right now (for NCCL's CTAs)
  int sm_reserve_thin = 0;          // ... for contractions whose panel traffic is small next to their GEMM work (dense)
  int sm_reserve_wide = 0;          // ... for traffic-heavy (block-sparse) contractions: wide communicators, more CTAs
  ncclComm* row_comm_wide = nullptr;
  ncclComm* col_comm_wide = nullptr;
  bool force_generic_gemm = false;  // TADEV_GEMM_GENERIC=1: always use the cp.async kernel
  void* tmap_cache = nullptr;       // TmapCache (gemm_f64_ws.cu): per-tile CUtensorMaps in device memory
};

// A leased staging slot of at least `bytes` for stream s: pinned host + device buffers for one
// descriptor block. acquire() waits (under the ctx lock) for a slot that no other thread holds, then
// — outside the lock — for the kernel that last consumed that slot. release() (also run by the
// destructor, so every early return is covered) records `done` on the stream AFTER the consuming
// kernel was enqueued and only then hands the slot back: concurrent callers that share a stream
// (MADWorld pool threads, stream_for(ordinal)) can therefore never overwrite a block that another
// thread is still filling or has not launched yet.
struct StageLease {
  tadev_ctx* ctx = nullptr;
  cudaStream_t s = nullptr;
  StagingRing* ring = nullptr;
  int slot = -1;
  void* h = nullptr;
  void* d = nullptr;
  cudaEvent_t done = nullptr, uploaded = nullptr;
  int acquire(tadev_ctx* ctx, cudaStream_t s, size_t bytes);
  void release();
  ~StageLease() { release(); }
  StageLease() = default;
  StageLease(const StageLease&) = delete;
  StageLease& operator=(const StageLease&) = delete;
};
// Upload a staged descriptor block on the ctx's descriptor stream and make `s` wait for it. Enqueued
// on `s` itself the copy would only be issued when the previous kernel of `s` finishes — the moment the
// SUMMA driver's next multi-GB panel upload grabs the H2D copy engine — and the next GEMM would start a
// whole panel-copy late (measured: 25 ms per window with host-resident operands).
int tadev_stage_upload(tadev_ctx* ctx, cudaStream_t s, void* d, const void* h, size_t bytes, cudaEvent_t uploaded);

// kernels' internal launchers (device-resident descriptors)
int launch_gemm_grouped_f64(tadev_ctx* ctx, cudaStream_t s, int opA, int opB, double alpha,
                            const tadev_gemm_group* d_groups, int ngroups, const tadev_gemm_task* d_tasks,
                            const int32_t* d_tile_prefix, int total_cta_tiles, bool aligned16);

// fast path (host descriptors; stages them, resolves tensor maps, launches the persistent kernel)
int launch_gemm_grouped_f64_ws(tadev_ctx* ctx, cudaStream_t s, int opA, int opB, double alpha,
                               const tadev_gemm_group* h_groups, int ngroups, const tadev_gemm_task* h_tasks,
                               int ntasks, int total_cta_tiles);
void tadev_tmap_cache_destroy(tadev_ctx* ctx);

// device-side task of the fast (TMA) kernel: the public task + the tensor maps of its k-contiguous operands
struct TadevWsTask {
  const double* A;
  const double* B;
  int32_t k;
  int32_t pad;
  const void* mapA;  // CUtensorMap*
  const void* mapB;
};

// Optional per-launch timing hook of the grouped GEMM: when set (by the SUMMA driver, on the calling thread) the
// next launch records ev[0] immediately before and ev[1] immediately after the kernel on its stream, so the
// kernel's own duration is measured without the descriptor upload / panel waits that precede it.
struct GemmTimingHook { cudaEvent_t before = nullptr, after = nullptr; };
GemmTimingHook& tadev_gemm_timing_hook();

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// CTA tile of the grouped DGEMM kernel (shared by host tile counting and the kernel)
constexpr int kGemmBM = 128;
constexpr int kGemmBN = 128;
