// gemm_f64_ws.cu — the fast path of the grouped FP64 tile GEMM: persistent, warp-specialised,
// TMA-staged, mbarrier-pipelined DMMA kernel for sm_100a.
//
// Same contract as gemm_f64.cu (reference chain: contract_reduce.h:409-453 -> tensor.h:3132 ->
// kernels.h:92-231 -> math/blas.h:171-177), used when every operand row is 16-byte aligned
// (even leading dimensions — the common case).
//
// Structure (one CTA per SM, 384 threads = 3 warpgroups):
//   warp 8      producer (warps 9-11 idle, their registers go to the consumers): pulls CTA tiles from a global atomic counter (dynamic scheduling evens
//               out block-sparse groups of different K), finds the owning group by a
//               warp-cooperative 32-ary search, and streams operand slabs of 16 k-values into a
//               3-stage shared-memory ring (2 x 16 k-values per stage) through the TMA engine, completing on the stage's
//               "full" mbarrier (expect-tx). It runs ahead across tile boundaries, so the next
//               tile's operands land while the consumers are still in their epilogue.
//                 * operand stored k-contiguous ([outer][k]: A/N, B/T): ONE tiled-mode TMA
//                   (cp.async.bulk.tensor.2d, box 16 x 128, SWIZZLE_128B, OOB zero-fill) from a
//                   per-tile CUtensorMap kept in a ctx-owned device cache;
//                 * operand stored outer-contiguous ([k][outer]: A/T, B/N): 16 bulk row copies of
//                   1 KiB (cp.async.bulk) into rows padded to 132 doubles.
//               (First attempt used 128-byte bulk row copies for the k-contiguous case: the TMA
//               engine retires ~1 small copy per 30 clk and the kernel ran at 17.7 TF; with 1 KiB
//               rows / tiled boxes it runs at 35.6 TF = 96% of the DMMA peak.)
//   warps 0..7  consumers: 2(m) x 4(n) layout, each owns a 64x32 block of the 128x128 CTA tile as
//               8x4 DMMA.8x8x4 fragments (64 accumulator doubles per lane); wait on "full", LDS
//               fragments, issue DMMAs, arrive on "empty". No CTA-wide barrier in the loop.
// Shared-memory bank conflicts: the contraction index is a dummy, so lane t of k4-step s uses
//   kk(s,t) = (10*(t>>1) + (t&1)) ^ (2*s)
// instead of 4*s+t. Bits (3,0) of kk enumerate t  => conflict-free reads of the 128B-swizzled
// k-contiguous slab; bits (1,0) enumerate t => conflict-free reads of the padded
// outer-contiguous slab. Both operands use the same kk, so products still pair up correctly.
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint
#include <cudaTypedefs.h>

#include <algorithm>
#include <unordered_map>

#include "common.h"
#include "tilelist.h"

namespace {

#ifndef TADEV_WS_SUB
#define TADEV_WS_SUB 2   // 16-wide k sub-slabs per pipeline stage (one mbarrier round trip per stage)
#endif
#ifndef TADEV_WS_STAGES
#define TADEV_WS_STAGES (TADEV_WS_SUB == 1 ? 4 : 3)
#endif
constexpr int BM = kGemmBM, BN = kGemmBN, BK = 16, SUB = TADEV_WS_SUB, STAGES = TADEV_WS_STAGES;
constexpr int NCONS = 8;  // consumer warps
// 12 warps = 3 warpgroups: the register file is carved per warpgroup, so the producer group
// (warp 8 works, 9-11 idle) hands its registers to the two consumer groups via setmaxnreg.
constexpr int NTHREADS = (NCONS + 4) * 32;
constexpr int LD_OMAJOR = BM + 4;                // 132 doubles: [k 16][outer 128] padded rows
constexpr int SLAB_BYTES = 17408;                // >= 16*132*8 (16896) and >= 128*128 (16384); 17 KiB keeps 1 KiB alignment
constexpr int SLAB_DOUBLES = SLAB_BYTES / 8;
constexpr int SUBSTAGE_DOUBLES = 2 * SLAB_DOUBLES;          // [A slab][B slab] of one 16-wide sub-slab
constexpr int STAGE_DOUBLES = SUB * SUBSTAGE_DOUBLES;
constexpr int SMEM_DATA_BYTES = STAGES * SUB * 2 * SLAB_BYTES;  // 3 x 2 x 34816 = 208896
constexpr int LAST_FLAG = 0x100;

using WsTask = TadevWsTask;  // device-side task: public task + tensor maps of its k-contiguous operands

struct Ctrl {  // lives after the data ring in dynamic smem
  unsigned long long full[STAGES];
  unsigned long long empty[STAGES];
  unsigned long long sched_full[2];
  unsigned long long sched_empty[2];
  int kb[STAGES];      // valid k-values in the slab | LAST_FLAG
  int sched_group[2];  // group index or -1 (no more work)
  int sched_m0[2];
  int sched_n0[2];
};
constexpr int SMEM_BYTES = SMEM_DATA_BYTES + (int)sizeof(Ctrl);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_2d_g2s(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                           unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tensormap_acquire(const CUtensorMap* map) {
  // the descriptor was written to global memory by a host copy: order it for the tensormap proxy
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;\n" ::"l"(map) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

template <int OPA, int OPB>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_grouped_f64_ws_kernel(const tadev_gemm_group* __restrict__ groups, int ngroups,
                           const WsTask* __restrict__ tasks, const int2* __restrict__ items,
                           const int* __restrict__ total_ptr, int* __restrict__ tile_counter, double alpha, int static_sched,
                           int wave_sync) {
  constexpr bool A_KIN = (OPA == TADEV_OP_N);  // A stored [m][k]
  constexpr bool B_KIN = (OPB == TADEV_OP_T);  // B stored [n][k]
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw + SMEM_DATA_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&ctrl->full[s], 1); mbar_init(&ctrl->empty[s], NCONS); }
    for (int s = 0; s < 2; ++s) { mbar_init(&ctrl->sched_full[s], 1); mbar_init(&ctrl->sched_empty[s], NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCONS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (warp != NCONS) return;
    // =============================== producer warp ===============================
    // number of work items: host-built lists stage it next to the scheduler counters, device-built lists
    // (tilelist.cu) leave it in their counter block
    const int total_tiles = __ldg(total_ptr);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0;; ++it) {
      const int slot = it & 1;
      mbar_wait(&ctrl->sched_empty[slot], ((it >> 1) & 1) ^ 1);
      int w = it * (int)gridDim.x + (int)blockIdx.x;
      if (!static_sched) {
        if (lane == 0) w = atomicAdd(tile_counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
      }
      if (w >= total_tiles) {
        if (lane == 0) { ctrl->sched_group[slot] = -1; mbar_arrive(&ctrl->sched_full[slot]); }
        break;
      }
      // Re-alignment: CTAs that share operand panels through L2 must stay within a few dozen k-slabs
      // of each other (126 MB L2 vs 148 streams of 32 KB per slab); start-time jitter accumulates
      // over hundreds of waves, so every `wave_sync` waves all CTAs wait for the previous wave.
      if (wave_sync > 0) {
        const int wave = w / (int)gridDim.x;
        if (wave > 0 && wave % wave_sync == 0) {
          const int need = wave * (int)gridDim.x;
          if (lane == 0) {
            while (*reinterpret_cast<volatile int*>(tile_counter + 1) < need) __nanosleep(200);
          }
          __syncwarp();
        }
      }
      // work item w of the rasterised order (host-built, see raster_items): group + 128x128 block
      const int2 item = __ldg(items + w);
      const int gi = item.x;
      const tadev_gemm_group grp = groups[gi];
      const int m0 = (int)((unsigned)item.y >> 16) * BM, n0 = (item.y & 0xffff) * BN;
      if (lane == 0) {
        ctrl->sched_group[slot] = gi; ctrl->sched_m0[slot] = m0; ctrl->sched_n0[slot] = n0;
        mbar_arrive(&ctrl->sched_full[slot]);
      }
      const int M = grp.m, N = grp.n;
      const int rowsA = min(BM, M - m0), colsB = min(BN, N - n0);
      // the LAST flag rides on the final slab of the last task with k > 0
      int last_task = -1;
      for (int ti = grp.task_end - 1; ti >= grp.task_begin; --ti)
        if (tasks[ti].k > 0) { last_task = ti; break; }
      if (last_task < 0) {  // nothing to contract: publish an empty terminal slab
        mbar_wait(&ctrl->empty[stage], phase ^ 1);
        if (lane == 0) { ctrl->kb[stage] = LAST_FLAG; mbar_arrive(&ctrl->full[stage]); }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (wave_sync > 0 && lane == 0) atomicAdd(tile_counter + 1, 1);
        continue;
      }
      for (int ti = grp.task_begin; ti <= last_task; ++ti) {
        const WsTask T = tasks[ti];
        const int K = T.k;
        if (K <= 0) continue;
        if (lane == 0) {
          if (A_KIN) tensormap_acquire(static_cast<const CUtensorMap*>(T.mapA));
          if (B_KIN) tensormap_acquire(static_cast<const CUtensorMap*>(T.mapB));
        }
        __syncwarp();
        for (int k0 = 0; k0 < K; k0 += BK * SUB) {
          const int kb = min(BK * SUB, K - k0);
          const bool last = (ti == last_task) && (k0 + BK * SUB >= K);
          mbar_wait(&ctrl->empty[stage], phase ^ 1);
          double* sStage = smem + stage * STAGE_DOUBLES;
          if (lane == 0) {
            ctrl->kb[stage] = kb | (last ? LAST_FLAG : 0);
            uint32_t bytes = 0;
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {
              const int kbs = min(BK, kb - sub * BK);
              if (kbs <= 0) break;
              bytes += A_KIN ? (uint32_t)(BM * BK * 8) : (uint32_t)(rowsA * kbs * 8);
              bytes += B_KIN ? (uint32_t)(BN * BK * 8) : (uint32_t)(colsB * kbs * 8);
            }
            mbar_arrive_expect_tx(&ctrl->full[stage], bytes);
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {
              if (kb - sub * BK <= 0) break;
              double* sA = sStage + sub * SUBSTAGE_DOUBLES;
              if (A_KIN) tma_2d_g2s(sA, static_cast<const CUtensorMap*>(T.mapA), k0 + sub * BK, m0, &ctrl->full[stage]);
              if (B_KIN) tma_2d_g2s(sA + SLAB_DOUBLES, static_cast<const CUtensorMap*>(T.mapB), k0 + sub * BK, n0, &ctrl->full[stage]);
            }
          }
          __syncwarp();
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
            const int kbs = min(BK, kb - sub * BK);
            if (kbs <= 0) break;
            double* sA = sStage + sub * SUBSTAGE_DOUBLES;
            double* sB = sA + SLAB_DOUBLES;
            const int kr = k0 + sub * BK + lane;
            if (!A_KIN && lane < kbs) bulk_g2s(sA + lane * LD_OMAJOR, T.A + (size_t)kr * M + m0, rowsA * 8, &ctrl->full[stage]);
            if (!B_KIN && lane < kbs) bulk_g2s(sB + lane * LD_OMAJOR, T.B + (size_t)kr * N + n0, colsB * 8, &ctrl->full[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (wave_sync > 0 && lane == 0) atomicAdd(tile_counter + 1, 1);  // all loads of this item are issued
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    // =============================== consumer warps ===============================
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    const int kt = (t >> 1) * 10 + (t & 1);  // kk(s,t) = kt ^ (2*s)
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0;; ++it) {
      const int slot = it & 1;
      mbar_wait(&ctrl->sched_full[slot], (it >> 1) & 1);
      const int gi = ctrl->sched_group[slot];
      const int m0 = ctrl->sched_m0[slot], n0 = ctrl->sched_n0[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->sched_empty[slot]);
      if (gi < 0) break;
      const tadev_gemm_group grp = groups[gi];

      double acc[8][4][2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

      for (;;) {
        mbar_wait(&ctrl->full[stage], phase);
        const int flags = ctrl->kb[stage];
        const int kb_stage = flags & 0xff;
#pragma unroll
        for (int sub = 0; sub < SUB; ++sub) {
          const int kb = min(BK, kb_stage - sub * BK);  // valid k-values of this sub-slab
          if (kb <= 0) break;
          const double* sA = smem + stage * STAGE_DOUBLES + sub * SUBSTAGE_DOUBLES;
          const double* sB = sA + SLAB_DOUBLES;
          // fragment offsets. k-contiguous slab: row R holds 16 doubles; its 16-byte chunk c sits at
          // chunk (c ^ (R & 7)) (SWIZZLE_128B); rows wm+i*8+g have R & 7 == g.
          auto kstep = [&](int s, bool masked) {
            const int kk = kt ^ (2 * s);
            const int kin_off = ((((kk >> 1) ^ g) << 1) | (kk & 1));
            const bool valid = !masked || (kk < kb);
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = wm + i * 8 + g;
              const double v = A_KIN ? sA[row * BK + kin_off] : sA[kk * LD_OMAJOR + row];
              a[i] = valid ? v : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = wn + j * 8 + g;
              const double v = B_KIN ? sB[col * BK + kin_off] : sB[kk * LD_OMAJOR + col];
              b[j] = valid ? v : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          };
          if (kb == BK) {
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep(s, false);
          } else {
            // K tail: stale rows of an outer-contiguous slab are masked to zero on both operands
#pragma unroll 1
            for (int s = 0; s < 4; ++s) kstep(s, true);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (flags & LAST_FLAG) break;
      }

      // ---- epilogue (identical mapping to the generic kernel)
      const int M = grp.m, N = grp.n;
      double* __restrict__ C = grp.C;
      const bool beta1 = grp.accumulate != 0;
      const bool vec_ok = ((N & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = m0 + wm + i * 8 + g;
        if (row >= M) continue;
        double* crow = C + (size_t)row * N;
        if (vec_ok) {
          double2 old[4];
          if (beta1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = n0 + wn + j * 8 + 2 * t;
              old[j] = (col < N) ? *reinterpret_cast<const double2*>(crow + col) : make_double2(0.0, 0.0);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn + j * 8 + 2 * t;
            if (col < N) {
              double2 v = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
              if (beta1) { v.x += old[j].x; v.y += old[j].y; }
              *reinterpret_cast<double2*>(crow + col) = v;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn + j * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (col + e < N) {
                double v = alpha * acc[i][j][e];
                if (beta1) v += crow[col + e];
                crow[col + e] = v;
              }
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + launcher

struct TmapKey {
  const void* ptr; int32_t outer, k;
  bool operator==(const TmapKey& o) const { return ptr == o.ptr && outer == o.outer && k == o.k; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& x) const {
    return std::hash<const void*>()(x.ptr) ^ (std::hash<uint64_t>()(((uint64_t)(uint32_t)x.outer << 32) | (uint32_t)x.k) * 0x9E3779B97F4A7C15ull);
  }
};

struct TmapCache {
  static constexpr int kChunk = 4096;          // maps per device chunk (chunks are never reallocated)
  static constexpr size_t kMaxMaps = 1u << 20; // 128 MiB of descriptors: beyond that the cache is recycled
  std::mutex mu;
  std::unordered_map<TmapKey, const CUtensorMap*, TmapKeyHash> index;
  std::vector<CUtensorMap*> chunks;
  int cur_chunk = -1, used_in_cur = kChunk;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  cudaStream_t upload = nullptr;
  CUtensorMap* h_stage = nullptr;  // pinned
  int h_cap = 0;
  struct Pending { CUtensorMap* dst; int chunk; };
  std::vector<Pending> pending;    // created, not yet uploaded (h_stage[n] <-> pending[n])
};

int tmap_cache_get(tadev_ctx* ctx, TmapCache** out) {
  if (!ctx->tmap_cache) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->tmap_cache) {
      TmapCache* c = new TmapCache();
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      TADEV_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      TADEV_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from the driver");
      c->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
      TADEV_CHECK_CUDA(cudaStreamCreateWithFlags(&c->upload, cudaStreamNonBlocking));
      ctx->tmap_cache = c;
    }
  }
  *out = (TmapCache*)ctx->tmap_cache;
  return TADEV_OK;
}

// encode the tiled map of one k-contiguous operand tile ([outer][k] doubles, box 16 x 128, SWIZZLE_128B) on the host
int encode_map(TmapCache* c, CUtensorMap* hm, const double* ptr, int outer, int k) {
  static const int promo_env = getenv("TADEV_TMAP_L2PROMO") ? atoi(getenv("TADEV_TMAP_L2PROMO")) : 3;
  const CUtensorMapL2promotion l2promo = promo_env == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : promo_env == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : promo_env == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)outer};
  const cuuint64_t gstr[1] = {(cuuint64_t)k * 8};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = c->encode(hm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2promo,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    tadev_set_error("cuTensorMapEncodeTiled failed (%d) for tile %p [%d x %d]", (int)cr, (const void*)ptr, outer, k);
    return TADEV_ECUDA;
  }
  return TADEV_OK;
}

// upload the maps created since the last flush on the cache's private stream and wait for them (c->mu held)
int flush_pending_locked(TmapCache* c) {
  auto& pending = c->pending;
  if (pending.empty()) return TADEV_OK;
  size_t i = 0;
  while (i < pending.size()) {  // contiguous runs inside a chunk are uploaded with one copy each
    size_t j = i + 1;
    // (two chunks may be adjacent in the address space: one copy must not span two allocations)
    while (j < pending.size() && pending[j].chunk == pending[j - 1].chunk && pending[j].dst == pending[j - 1].dst + 1) ++j;
    TADEV_CHECK_CUDA(cudaMemcpyAsync(pending[i].dst, c->h_stage + i, sizeof(CUtensorMap) * (j - i), cudaMemcpyHostToDevice, c->upload));
    i = j;
  }
  TADEV_CHECK_CUDA(cudaStreamSynchronize(c->upload));
  pending.clear();
  return TADEV_OK;
}

// cached device map of (ptr, outer, k); created (encoded, queued for upload) on a miss (c->mu held)
int lookup_locked(TmapCache* c, const double* ptr, int outer, int k, const CUtensorMap** res, bool* created) {
  TmapKey key{ptr, outer, k};
  auto itf = c->index.find(key);
  if (itf != c->index.end()) { *res = itf->second; return TADEV_OK; }
  if (c->index.size() >= TmapCache::kMaxMaps) {
    // bounded cache: long runs with changing tile addresses would otherwise grow it without limit. Nothing may
    // still read a descriptor that is about to be overwritten, hence the device-wide sync (a rare event).
    int rc = flush_pending_locked(c);
    if (rc) return rc;
    TADEV_CHECK_CUDA(cudaDeviceSynchronize());
    c->index.clear();
    c->cur_chunk = c->chunks.empty() ? -1 : 0;
    c->used_in_cur = c->chunks.empty() ? TmapCache::kChunk : 0;
  }
  if (c->used_in_cur == TmapCache::kChunk) {
    if (c->cur_chunk + 1 < (int)c->chunks.size()) ++c->cur_chunk;
    else {
      CUtensorMap* chunk = nullptr;
      TADEV_CHECK_CUDA(cudaMalloc(&chunk, sizeof(CUtensorMap) * TmapCache::kChunk));
      c->chunks.push_back(chunk);
      c->cur_chunk = (int)c->chunks.size() - 1;
    }
    c->used_in_cur = 0;
  }
  CUtensorMap* dst = c->chunks[c->cur_chunk] + c->used_in_cur++;
  if ((int)c->pending.size() == c->h_cap) {
    const int ncap = c->h_cap ? c->h_cap * 2 : 1024;
    CUtensorMap* nh = nullptr;
    TADEV_CHECK_CUDA(cudaMallocHost(&nh, sizeof(CUtensorMap) * ncap));
    if (c->h_stage) { memcpy(nh, c->h_stage, sizeof(CUtensorMap) * c->pending.size()); cudaFreeHost(c->h_stage); }
    c->h_stage = nh; c->h_cap = ncap;
  }
  int rc = encode_map(c, c->h_stage + c->pending.size(), ptr, outer, k);
  if (rc) { --c->used_in_cur; return rc; }
  c->pending.push_back({dst, c->cur_chunk});
  c->index.emplace(key, dst);
  *res = dst;
  if (created) *created = true;
  return TADEV_OK;
}

// Resolve (and create on demand) the tensor maps of all k-contiguous operands of a batch.
// New maps are encoded on the host and uploaded on a private stream which is synchronised
// before returning, so any stream may use them afterwards.
int resolve_maps(tadev_ctx* ctx, int opA, int opB, const tadev_gemm_group* groups, int ngroups,
                 const tadev_gemm_task* tasks, WsTask* out) {
  const bool a_kin = opA == TADEV_OP_N, b_kin = opB == TADEV_OP_T;
  TmapCache* c = nullptr;
  int rc = tmap_cache_get(ctx, &c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  for (int gi = 0; gi < ngroups; ++gi) {
    const tadev_gemm_group& G = groups[gi];
    for (int ti = G.task_begin; ti < G.task_end; ++ti) {
      const tadev_gemm_task& T = tasks[ti];
      WsTask& W = out[ti];
      W.A = T.A; W.B = T.B; W.k = T.k; W.pad = 0; W.mapA = nullptr; W.mapB = nullptr;
      if (T.k <= 0 || G.m <= 0 || G.n <= 0) continue;
      const CUtensorMap* m = nullptr;
      if (a_kin) { rc = lookup_locked(c, T.A, G.m, T.k, &m, nullptr); if (rc) return rc; W.mapA = m; }
      if (b_kin) { rc = lookup_locked(c, T.B, G.n, T.k, &m, nullptr); if (rc) return rc; W.mapB = m; }
    }
  }
  return flush_pending_locked(c);
}

template <int OPA, int OPB>
int launch_ws_variant(cudaStream_t s, int grid, const tadev_gemm_group* d_groups, int ngroups, const WsTask* d_tasks,
                      const int2* d_items, const int* d_total, int* d_counter, double alpha, int static_sched, int wave_sync) {
  auto kern = gemm_grouped_f64_ws_kernel<OPA, OPB>;
  // the attribute is per device (context): set it on every launch (a cheap host-side call) rather than once per
  // process, so a process driving several devices gets the opt-in on each of them
  TADEV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  GemmTimingHook& hook = tadev_gemm_timing_hook();
  if (hook.before) TADEV_CHECK_CUDA(cudaEventRecord(hook.before, s));
  kern<<<grid, NTHREADS, SMEM_BYTES, s>>>(d_groups, ngroups, d_tasks, d_items, d_total, d_counter, alpha, static_sched, wave_sync);
  TADEV_CHECK_CUDA(cudaGetLastError());
  if (hook.after) TADEV_CHECK_CUDA(cudaEventRecord(hook.after, s));
  hook.before = hook.after = nullptr;
  return TADEV_OK;
}

}  // namespace

void tadev_tmap_cache_destroy(tadev_ctx* ctx) {
  TmapCache* c = (TmapCache*)ctx->tmap_cache;
  if (!c) return;
  for (auto p : c->chunks) cudaFree(p);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->upload) cudaStreamDestroy(c->upload);
  delete c;
  ctx->tmap_cache = nullptr;
}

namespace {

// Work-item order of the persistent kernel. Item = (group, 128x128 block of its result tile); the
// CTAs pull items in this order from an atomic counter, so the ~num_SMs items in flight at any time
// are `grid` consecutive entries. Groups that carry a raster hint (tadev_gemm_group::raster: origin
// of the result tile in a global grid of 128x128 blocks) are ordered in bands of S = floor(sqrt(grid))
// block rows, column-major inside a band: a wave then covers ~S x S blocks, i.e. S block-rows of A
// and S block-columns of B are shared through L2 instead of 8 x 8 per result tile (DRAM traffic of
// the N=32768 launch: 1.19 TB group-major). Without hints the order is group-major.
void raster_items(const tadev_gemm_group* groups, int ngroups, int grid, int2* items, int total) {
  bool hinted = ngroups > 1 && !(getenv("TADEV_RASTER_S") && atoi(getenv("TADEV_RASTER_S")) == 0);
  for (int gi = 0; gi < ngroups && hinted; ++gi) hinted = groups[gi].raster != 0 || groups[gi].m == 0 || groups[gi].n == 0;
  int w = 0;
  if (!hinted) {
    for (int gi = 0; gi < ngroups; ++gi) {
      const int tm = (int)ceil_div64(groups[gi].m, BM), tn = (int)ceil_div64(groups[gi].n, BN);
      for (int a = 0; a < tm; ++a)
        for (int b = 0; b < tn; ++b) items[w++] = make_int2(gi, (a << 16) | b);
    }
    return;
  }
  int S = 1;
  while ((S + 1) * (S + 1) <= grid) ++S;
  static const int env_S = getenv("TADEV_RASTER_S") ? atoi(getenv("TADEV_RASTER_S")) : -1;
  static const bool row_major = getenv("TADEV_RASTER_ROWMAJOR") && atoi(getenv("TADEV_RASTER_ROWMAJOR"));
  if (env_S > 0) S = env_S;
  int rows = 0, cols = 0;
  for (int gi = 0; gi < ngroups; ++gi) {
    if (!groups[gi].raster) continue;
    const int r0 = (int)((uint32_t)groups[gi].raster >> 16) - 1, c0 = (groups[gi].raster & 0xffff) - 1;
    rows = std::max(rows, r0 + (int)ceil_div64(groups[gi].m, BM));
    cols = std::max(cols, c0 + (int)ceil_div64(groups[gi].n, BN));
  }
  // key = (band, global block column, row inside the band): dense integers -> counting sort
  const int64_t nkeys = row_major ? (int64_t)ceil_div64(cols, S) * rows * S : (int64_t)ceil_div64(rows, S) * cols * S;
  auto key_of = [&](int gr, int gc) {
    return row_major ? ((int64_t)(gc / S) * rows + gr) * S + gc % S : ((int64_t)(gr / S) * cols + gc) * S + gr % S;
  };
  if (nkeys <= 16 * (int64_t)total + 4096) {
    std::vector<int32_t> start((size_t)nkeys + 1, 0);
    for (int gi = 0; gi < ngroups; ++gi) {
      if (!groups[gi].raster) continue;
      const int r0 = (int)((uint32_t)groups[gi].raster >> 16) - 1, c0 = (groups[gi].raster & 0xffff) - 1;
      const int tm = (int)ceil_div64(groups[gi].m, BM), tn = (int)ceil_div64(groups[gi].n, BN);
      for (int a = 0; a < tm; ++a)
        for (int b = 0; b < tn; ++b) ++start[(size_t)key_of(r0 + a, c0 + b) + 1];
    }
    for (int64_t k = 0; k < nkeys; ++k) start[k + 1] += start[k];
    for (int gi = 0; gi < ngroups; ++gi) {
      if (!groups[gi].raster) continue;
      const int r0 = (int)((uint32_t)groups[gi].raster >> 16) - 1, c0 = (groups[gi].raster & 0xffff) - 1;
      const int tm = (int)ceil_div64(groups[gi].m, BM), tn = (int)ceil_div64(groups[gi].n, BN);
      for (int a = 0; a < tm; ++a)
        for (int b = 0; b < tn; ++b) items[start[key_of(r0 + a, c0 + b)]++] = make_int2(gi, (a << 16) | b);
    }
    return;
  }
  std::vector<std::pair<int64_t, int2>> keyed;
  keyed.reserve(total);
  for (int gi = 0; gi < ngroups; ++gi) {
    if (!groups[gi].raster) continue;
    const int r0 = (int)((uint32_t)groups[gi].raster >> 16) - 1, c0 = (groups[gi].raster & 0xffff) - 1;
    const int tm = (int)ceil_div64(groups[gi].m, BM), tn = (int)ceil_div64(groups[gi].n, BN);
    for (int a = 0; a < tm; ++a)
      for (int b = 0; b < tn; ++b) keyed.push_back({key_of(r0 + a, c0 + b), make_int2(gi, (a << 16) | b)});
  }
  std::stable_sort(keyed.begin(), keyed.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
  for (auto& kv : keyed) items[w++] = kv.second;
}

}  // namespace

// Host-descriptor entry of the fast path: builds device descriptors (+ tensor maps) and launches.
int launch_gemm_grouped_f64_ws(tadev_ctx* ctx, cudaStream_t s, int opA, int opB, double alpha,
                               const tadev_gemm_group* h_groups, int ngroups, const tadev_gemm_task* h_tasks,
                               int ntasks, int total_cta_tiles) {
  if (ngroups == 0 || total_cta_tiles == 0) return TADEV_OK;
  const size_t gb = sizeof(tadev_gemm_group) * (size_t)ngroups;
  const size_t tb = sizeof(WsTask) * (size_t)ntasks;
  const size_t pb = sizeof(int2) * (size_t)total_cta_tiles;
  const size_t off_t = (gb + 15) & ~size_t(15);
  const size_t off_p = (off_t + tb + 15) & ~size_t(15);
  const size_t off_c = (off_p + pb + 15) & ~size_t(15);
  StageLease L;
  int rc = L.acquire(ctx, s, off_c + 16);
  if (rc) return rc;
  void *h = L.h, *d = L.d;
  cudaEvent_t uploaded = L.uploaded;
  memcpy(h, h_groups, gb);
  rc = resolve_maps(ctx, opA, opB, h_groups, ngroups, h_tasks, (WsTask*)((char*)h + off_t));
  if (rc) return rc;
  int grid = ctx->num_sms - ctx->gemm_sm_reserve;
  if (grid < 1) grid = 1;
  if (grid > total_cta_tiles) grid = total_cta_tiles;
  raster_items(h_groups, ngroups, grid, (int2*)((char*)h + off_p), total_cta_tiles);
  static const int static_sched = getenv("TADEV_SCHED_STATIC") ? atoi(getenv("TADEV_SCHED_STATIC")) : 0;
  // wave re-alignment (see the kernel) pays off when all work items take the same time (dense
  // contractions: every result tile has the same total K); with ragged K it would idle CTAs.
  static const int wave_sync_env = getenv("TADEV_WAVE_SYNC") ? atoi(getenv("TADEV_WAVE_SYNC")) : -1;
  int wave_sync = wave_sync_env;
  if (wave_sync < 0) {
    bool uniform = ngroups > 1;
    int64_t k0 = -1;
    for (int gi = 0; gi < ngroups && uniform; ++gi) {
      if (h_groups[gi].m == 0 || h_groups[gi].n == 0) continue;
      int64_t kk = 0;
      for (int ti = h_groups[gi].task_begin; ti < h_groups[gi].task_end; ++ti) kk += h_tasks[ti].k;
      if (k0 < 0) k0 = kk;
      uniform = (kk == k0) && h_groups[gi].raster != 0;
    }
    wave_sync = uniform ? 1 : 0;
  }
  memset((char*)h + off_c, 0, 16);
  ((int32_t*)((char*)h + off_c))[2] = total_cta_tiles;  // [work counter, wave counter, number of work items, -]
  rc = tadev_stage_upload(ctx, s, d, h, off_c + 16, uploaded);
  if (rc) return rc;
  ctx->launches++;
  const tadev_gemm_group* dg = (const tadev_gemm_group*)d;
  const WsTask* dt = (const WsTask*)((char*)d + off_t);
  const int2* dp = (const int2*)((char*)d + off_p);
  int* dc = (int*)((char*)d + off_c);
  switch ((opA << 1) | opB) {
    case 0: rc = launch_ws_variant<0, 0>(s, grid, dg, ngroups, dt, dp, dc + 2, dc, alpha, static_sched, wave_sync); break;
    case 1: rc = launch_ws_variant<0, 1>(s, grid, dg, ngroups, dt, dp, dc + 2, dc, alpha, static_sched, wave_sync); break;
    case 2: rc = launch_ws_variant<1, 0>(s, grid, dg, ngroups, dt, dp, dc + 2, dc, alpha, static_sched, wave_sync); break;
    case 3: rc = launch_ws_variant<1, 1>(s, grid, dg, ngroups, dt, dp, dc + 2, dc, alpha, static_sched, wave_sync); break;
    default: tadev_set_error("launch_gemm_grouped_f64_ws: bad op flags %d %d", opA, opB); rc = TADEV_EINVAL;
  }
  return rc;  // ~StageLease records `done` after the launch and returns the slot
}

// Fast-path launch with DEVICE-resident lists (tilelist.cu): groups, tasks (with tensor maps), rasterised work items
// and their count were produced by kernels on `s`; the host knows only upper bounds, so the full persistent grid is
// launched and CTAs that find no work exit at once.
int launch_gemm_ws_devlists(tadev_ctx* ctx, cudaStream_t s, int opA, int opB, double alpha, const tadev_gemm_group* d_groups,
                            int ngroups, const void* d_wstasks, const int2* d_items, const int32_t* d_total, int32_t* d_sched,
                            int wave_sync) {
  int grid = ctx->num_sms - ctx->gemm_sm_reserve;
  if (grid < 1) grid = 1;
  static const int static_sched = getenv("TADEV_SCHED_STATIC") ? atoi(getenv("TADEV_SCHED_STATIC")) : 0;
  ctx->launches++;
  const WsTask* dt = static_cast<const WsTask*>(d_wstasks);
  switch ((opA << 1) | opB) {
    case 0: return launch_ws_variant<0, 0>(s, grid, d_groups, ngroups, dt, d_items, d_total, d_sched, alpha, static_sched, wave_sync);
    case 1: return launch_ws_variant<0, 1>(s, grid, d_groups, ngroups, dt, d_items, d_total, d_sched, alpha, static_sched, wave_sync);
    case 2: return launch_ws_variant<1, 0>(s, grid, d_groups, ngroups, dt, d_items, d_total, d_sched, alpha, static_sched, wave_sync);
    case 3: return launch_ws_variant<1, 1>(s, grid, d_groups, ngroups, dt, d_items, d_total, d_sched, alpha, static_sched, wave_sync);
  }
  tadev_set_error("launch_gemm_ws_devlists: bad op flags %d %d", opA, opB);
  return TADEV_EINVAL;
}

int tadev_ws_cached_map(tadev_ctx* ctx, const double* ptr, int outer, int k, const void** dev_map, bool* created) {
  TmapCache* c = nullptr;
  int rc = tmap_cache_get(ctx, &c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  const CUtensorMap* m = nullptr;
  rc = lookup_locked(c, ptr, outer, k, &m, created);
  *dev_map = m;
  return rc;
}

int tadev_ws_flush_new_maps(tadev_ctx* ctx) {
  TmapCache* c = nullptr;
  int rc = tmap_cache_get(ctx, &c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  return flush_pending_locked(c);
}

int tadev_ws_encode_map(tadev_ctx* ctx, void* h_dst128, const double* ptr, int outer, int k) {
  TmapCache* c = nullptr;
  int rc = tmap_cache_get(ctx, &c);
  if (rc) return rc;
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
  return encode_map(c, static_cast<CUtensorMap*>(h_dst128), ptr, outer, k);
}
