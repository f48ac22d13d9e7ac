// gemm_f64_ws.cu — the fast path of the grouped FP64 tile GEMM: persistent, warp-specialised,
// bulk-async-copy (TMA engine, UBLKCP) staged, mbarrier-pipelined DMMA kernel for sm_100a.
//
// Same contract as gemm_f64.cu (reference chain: contract_reduce.h:409-453 -> tensor.h:3132 ->
// kernels.h:92-231 -> math/blas.h:171-177), used when every operand row is 16-byte aligned and
// every contracted extent is a multiple of 4 (the common case: even tile extents).
//
// Structure (one CTA per SM, 288 threads):
//   warp 8      producer: pulls CTA tiles from a global atomic counter (dynamic scheduling evens
//               out block-sparse groups of different K), finds the owning group by a
//               warp-cooperative 32-ary search, and streams 128x16 / 16x128 operand slabs into a
//               4-stage shared-memory ring with cp.async.bulk row copies that complete on the
//               stage's "full" mbarrier (expect-tx). It runs ahead across tile boundaries, so the
//               next tile's operands land while the consumers are still in their epilogue.
//   warps 0..7  consumers: 2(m) x 4(n) layout, each owns a 64x32 block of the 128x128 CTA tile as
//               8x4 DMMA.8x8x4 fragments (64 accumulator doubles per lane); wait on "full",
//               LDS fragments, issue DMMAs, arrive on "empty". No CTA-wide barrier in the loop,
//               so the warps drift apart and the tensor pipe sees a steady instruction stream
//               (v1's per-slab __syncthreads aligned all warps' load phases: 84% pipe-active).
// Ragged M/N edges: rows past the edge are simply not copied; stale shared memory only feeds
// accumulators whose results are never stored. K tails (k % 16 in {4,8,12}) shorten the slab.
#include "common.h"

namespace {

constexpr int BM = kGemmBM, BN = kGemmBN, BK = 16, STAGES = 4;
constexpr int NCONS = 8;                 // consumer warps
constexpr int NTHREADS = (NCONS + 1) * 32;
constexpr int LD_KMAJOR = BK + 4;        // 20  doubles: [outer 128][k 16]
constexpr int LD_OMAJOR = BM + 4;        // 132 doubles: [k 16][outer 128]
constexpr int SLAB_DOUBLES = BM * LD_KMAJOR;  // 2560
constexpr int STAGE_DOUBLES = 2 * SLAB_DOUBLES;
constexpr int SMEM_DATA_BYTES = STAGES * STAGE_DOUBLES * 8;  // 163840
constexpr int LAST_FLAG = 0x100;

struct Ctrl {  // lives after the data ring in dynamic smem
  unsigned long long full[STAGES];
  unsigned long long empty[STAGES];
  unsigned long long sched_full[2];
  unsigned long long sched_empty[2];
  int nk4[STAGES];       // k4-steps in the slab | LAST_FLAG
  int sched_group[2];    // group index or -1 (no more work)
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
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

template <int OPA, int OPB>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_grouped_f64_ws_kernel(const tadev_gemm_group* __restrict__ groups, int ngroups,
                           const tadev_gemm_task* __restrict__ tasks, const int32_t* __restrict__ tile_prefix,
                           int total_tiles, int* __restrict__ tile_counter, double alpha) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw + SMEM_DATA_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&ctrl->full[s], 1); mbar_init(&ctrl->empty[s], NCONS); }
    for (int s = 0; s < 2; ++s) { mbar_init(&ctrl->sched_full[s], 1); mbar_init(&ctrl->sched_empty[s], NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == NCONS) {
    // =============================== producer warp ===============================
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0;; ++it) {
      const int slot = it & 1;
      mbar_wait(&ctrl->sched_empty[slot], ((it >> 1) & 1) ^ 1);
      int w = 0;
      if (lane == 0) w = atomicAdd(tile_counter, 1);
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w >= total_tiles) {
        if (lane == 0) { ctrl->sched_group[slot] = -1; mbar_arrive(&ctrl->sched_full[slot]); }
        break;
      }
      // warp-cooperative 32-ary search: largest g with tile_prefix[g] <= w
      int lo = 0, hi = ngroups;  // prefix[lo] <= w < prefix[hi]
      while (hi - lo > 1) {
        const int span = hi - lo;
        const int step = (span + 31) / 32;
        const int probe = lo + lane * step;
        const bool le = (probe < hi) && (__ldg(tile_prefix + probe) <= w);
        const unsigned m = __ballot_sync(0xffffffffu, le);
        const int last = 31 - __clz(m);  // lane 0 always satisfies (prefix[lo] <= w)
        const int nlo = lo + last * step;
        int nhi = nlo + step;
        if (nhi > hi) nhi = hi;
        lo = nlo; hi = nhi;
      }
      const int gi = lo;
      const tadev_gemm_group grp = groups[gi];
      const int local = w - __ldg(tile_prefix + gi);
      const int tiles_n = (grp.n + BN - 1) / BN;
      const int m0 = (local / tiles_n) * BM, n0 = (local % tiles_n) * BN;
      if (lane == 0) {
        ctrl->sched_group[slot] = gi; ctrl->sched_m0[slot] = m0; ctrl->sched_n0[slot] = n0;
        mbar_arrive(&ctrl->sched_full[slot]);
      }
      const int M = grp.m, N = grp.n;
      const int rowsA = min(BM, M - m0), colsB = min(BN, N - n0);
      // find the last task with k > 0 so the LAST flag rides on a real slab when possible
      int last_task = -1;
      for (int ti = grp.task_end - 1; ti >= grp.task_begin; --ti)
        if (tasks[ti].k > 0) { last_task = ti; break; }
      if (last_task < 0) {  // nothing to contract: publish an empty terminal slab
        mbar_wait(&ctrl->empty[stage], phase ^ 1);
        if (lane == 0) { ctrl->nk4[stage] = LAST_FLAG; mbar_arrive(&ctrl->full[stage]); }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        continue;
      }
      for (int ti = grp.task_begin; ti <= last_task; ++ti) {
        const tadev_gemm_task T = tasks[ti];
        const int K = T.k;
        const int lda = OPA == TADEV_OP_N ? K : M;
        const int ldb = OPB == TADEV_OP_N ? N : K;
        for (int k0 = 0; k0 < K; k0 += BK) {
          const int kb = min(BK, K - k0);
          const bool last = (ti == last_task) && (k0 + BK >= K);
          mbar_wait(&ctrl->empty[stage], phase ^ 1);
          double* sA = smem + stage * STAGE_DOUBLES;
          double* sB = sA + SLAB_DOUBLES;
          if (lane == 0) {
            ctrl->nk4[stage] = (kb >> 2) | (last ? LAST_FLAG : 0);
            mbar_arrive_expect_tx(&ctrl->full[stage], (uint32_t)((rowsA + colsB) * kb * 8));
          }
          __syncwarp();
          if (OPA == TADEV_OP_N) {
            for (int rr = lane; rr < rowsA; rr += 32)
              bulk_g2s(sA + rr * LD_KMAJOR, T.A + (size_t)(m0 + rr) * lda + k0, kb * 8, &ctrl->full[stage]);
          } else {
            if (lane < kb) bulk_g2s(sA + lane * LD_OMAJOR, T.A + (size_t)(k0 + lane) * lda + m0, rowsA * 8, &ctrl->full[stage]);
          }
          if (OPB == TADEV_OP_N) {
            if (lane < kb) bulk_g2s(sB + lane * LD_OMAJOR, T.B + (size_t)(k0 + lane) * ldb + n0, colsB * 8, &ctrl->full[stage]);
          } else {
            for (int rr = lane; rr < colsB; rr += 32)
              bulk_g2s(sB + rr * LD_KMAJOR, T.B + (size_t)(n0 + rr) * ldb + k0, kb * 8, &ctrl->full[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =============================== consumer warps ===============================
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
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
        const int flags = ctrl->nk4[stage];
        const int nk = flags & 0xff;
        const double* sA = smem + stage * STAGE_DOUBLES;
        const double* sB = sA + SLAB_DOUBLES;
        auto kstep = [&](int s) {
          const int kk = s * 4 + t;
          double a[8], b[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = wm + i * 8 + g;
            a[i] = (OPA == TADEV_OP_N) ? sA[row * LD_KMAJOR + kk] : sA[kk * LD_OMAJOR + row];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = wn + j * 8 + g;
            b[j] = (OPB == TADEV_OP_N) ? sB[kk * LD_OMAJOR + col] : sB[col * LD_KMAJOR + kk];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        };
        if (nk == 4) {
#pragma unroll
          for (int s = 0; s < 4; ++s) kstep(s);
        } else {
          for (int s = 0; s < nk; ++s) kstep(s);
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

template <int OPA, int OPB>
int launch_ws_variant(cudaStream_t s, int grid, const tadev_gemm_group* d_groups, int ngroups,
                      const tadev_gemm_task* d_tasks, const int32_t* d_tile_prefix, int total_tiles, int* d_counter,
                      double alpha) {
  auto kern = gemm_grouped_f64_ws_kernel<OPA, OPB>;
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    TADEV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  kern<<<grid, NTHREADS, SMEM_BYTES, s>>>(d_groups, ngroups, d_tasks, d_tile_prefix, total_tiles, d_counter, alpha);
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}

}  // namespace

int launch_gemm_grouped_f64_ws(tadev_ctx* ctx, cudaStream_t s, int opA, int opB, double alpha,
                               const tadev_gemm_group* d_groups, int ngroups, const tadev_gemm_task* d_tasks,
                               const int32_t* d_tile_prefix, int total_cta_tiles, int* d_counter, int sm_reserve) {
  if (ngroups == 0 || total_cta_tiles == 0) return TADEV_OK;
  ctx->launches++;
  int grid = ctx->num_sms - sm_reserve;
  if (grid < 1) grid = 1;
  if (grid > total_cta_tiles) grid = total_cta_tiles;
  switch ((opA << 1) | opB) {
    case 0: return launch_ws_variant<0, 0>(s, grid, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, d_counter, alpha);
    case 1: return launch_ws_variant<0, 1>(s, grid, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, d_counter, alpha);
    case 2: return launch_ws_variant<1, 0>(s, grid, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, d_counter, alpha);
    case 3: return launch_ws_variant<1, 1>(s, grid, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, d_counter, alpha);
  }
  tadev_set_error("launch_gemm_grouped_f64_ws: bad op flags %d %d", opA, opB);
  return TADEV_EINVAL;
}
