// gemm_f64.cu — grouped FP64 tile GEMM on the DMMA tensor pipe (sm_100a).
//
// Replaces, for device-resident tiles, the per-pair BLAS call chain of the reference
//   ContractReduce::operator() (tile_op/contract_reduce.h:409-453) -> Tensor::gemm
//   (tensor/tensor.h:3132-3219) -> detail::gemm (tensor/kernels.h:92-231) ->
//   math::blas::gemm (math/blas.h:171-177) / device::btas::gemm -> cuBLAS (device/btas.h:52-222)
// and the add_to merge of partial results (contract_reduce.h:386-400) by ONE launch per batch:
// a work item is a 128x128 block of a result tile; the K loop runs over the concatenation of all
// (left,right) tile pairs that contribute to that result tile (a "group"), accumulating in
// registers, and the result is written once (beta = 0) or added in place (beta = 1).
//
// FP64 has no tcgen05/UMMA kind on Blackwell; the tensor-core path for doubles is the warp-level
// DMMA.8x8x4 (every mma.sync f64 shape lowers to it on sm_100a — checked with cuobjdump). The
// pipe retires 64 FMA/clk/SM, so the kernel is DMMA-issue bound by a wide margin: operand
// staging needs only ~16 B/clk/SM. Staging is a 4-stage cp.async (LDGSTS) ring with zero-fill
// predication, which handles ragged tile extents (prime-sized tiles in the reference's tests)
// without a separate edge kernel.
//
// Layouts: all tiles row-major with natural leading dimensions (tensor/kernels.h:146-158):
//   opA == N: A is m x k (lda = k)      opA == T: A is k x m (lda = m)
//   opB == N: B is k x n (ldb = n)      opB == T: B is n x k (ldb = k)
//   C is m x n (ldc = n).
#include "common.h"

namespace {

constexpr int BM = kGemmBM;  // 128
constexpr int BN = kGemmBN;  // 128
constexpr int BK = 16;
constexpr int STAGES = 4;
constexpr int NTHREADS = 256;
// smem row strides (in doubles), chosen == 4 (mod 16) so that the 16 lanes of a half-warp
// (g = 0..3, t = 0..3) hit 16 distinct 8-byte bank pairs for both fragment patterns.
constexpr int LD_KMAJOR = BK + 4;   // 20 : slab stored [outer 128][k 16]   (A/N, B/T)
constexpr int LD_OMAJOR = BM + 4;   // 132: slab stored [k 16][outer 128]   (A/T, B/N)
constexpr int SLAB_DOUBLES = BM * LD_KMAJOR;  // 2560 >= BK*LD_OMAJOR (2112)
constexpr int STAGE_DOUBLES = 2 * SLAB_DOUBLES;
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;  // 163840

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// Load one operand slab. KMAJOR_GLOBAL: the global tile is stored [outer][k] (k contiguous),
// else [k][outer]. `outer0` is the first outer index of this CTA tile, `outer_ext` the tile's
// outer extent, k0 the first k of the slab, k_ext the pair's contracted extent.
template <bool KCONTIG, int VEC>
__device__ __forceinline__ void load_slab(double* smem, const double* __restrict__ g, int ld,
                                          int outer0, int outer_ext, int k0, int k_ext, int tid) {
  if (KCONTIG) {
    // smem [128][LD_KMAJOR]; global row = outer index, contiguous along k
    constexpr int CHUNKS_PER_ROW = BK / VEC;
    constexpr int TOTAL = BM * CHUNKS_PER_ROW;
#pragma unroll
    for (int c = tid; c < TOTAL; c += NTHREADS) {
      int row = c / CHUNKS_PER_ROW;
      int kc = (c % CHUNKS_PER_ROW) * VEC;
      bool p = (outer0 + row < outer_ext) && (k0 + kc < k_ext);
      const double* src = p ? g + (size_t)(outer0 + row) * ld + (k0 + kc) : g;
      if (VEC == 2) cp_async_16(smem + row * LD_KMAJOR + kc, src, p);
      else cp_async_8(smem + row * LD_KMAJOR + kc, src, p);
    }
  } else {
    // smem [16][LD_OMAJOR]; global row = k index, contiguous along outer
    constexpr int CHUNKS_PER_ROW = BM / VEC;
    constexpr int TOTAL = BK * CHUNKS_PER_ROW;
#pragma unroll
    for (int c = tid; c < TOTAL; c += NTHREADS) {
      int kr = c / CHUNKS_PER_ROW;
      int oc = (c % CHUNKS_PER_ROW) * VEC;
      bool p = (k0 + kr < k_ext) && (outer0 + oc < outer_ext);
      const double* src = p ? g + (size_t)(k0 + kr) * ld + (outer0 + oc) : g;
      if (VEC == 2) cp_async_16(smem + kr * LD_OMAJOR + oc, src, p);
      else cp_async_8(smem + kr * LD_OMAJOR + oc, src, p);
    }
  }
}

template <int OPA, int OPB, int VEC>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_grouped_f64_kernel(const tadev_gemm_group* __restrict__ groups, int ngroups,
                        const tadev_gemm_task* __restrict__ tasks,
                        const int32_t* __restrict__ tile_prefix, double alpha) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 2) * 64;  // warp tile origin inside the CTA tile
  const int wn = (warp & 3) * 32;

  // ---- locate my work item: binary search for the group that owns CTA tile blockIdx.x
  const int w = blockIdx.x;
  if (w >= __ldg(tile_prefix + ngroups)) return;  // device-built lists launch an upper bound of CTAs (tilelist.cu)
  int lo = 0, hi = ngroups;  // invariant: prefix[lo] <= w < prefix[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(tile_prefix + mid) <= w) lo = mid; else hi = mid;
  }
  const tadev_gemm_group grp = groups[lo];
  const int local = w - __ldg(tile_prefix + lo);
  const int tiles_n = (grp.n + BN - 1) / BN;
  const int m0 = (local / tiles_n) * BM;
  const int n0 = (local % tiles_n) * BN;
  const int M = grp.m, N = grp.n;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // ---- total slab count of the chained K loop
  int total_slabs = 0;
  for (int ti = grp.task_begin; ti < grp.task_end; ++ti) total_slabs += (tasks[ti].k + BK - 1) / BK;

  // producer cursor (runs STAGES-1 slabs ahead of the consumer, across task boundaries)
  int p_task = grp.task_begin;
  int p_k0 = 0;
  const double* pA = nullptr;
  const double* pB = nullptr;
  int p_k = 0;
  if (p_task < grp.task_end) {
    pA = tasks[p_task].A; pB = tasks[p_task].B; p_k = tasks[p_task].k;
  }
  auto produce = [&](int stage) {
    // skip exhausted (or zero-k) tasks
    while (p_task < grp.task_end && p_k0 >= p_k) {
      ++p_task; p_k0 = 0;
      if (p_task < grp.task_end) { pA = tasks[p_task].A; pB = tasks[p_task].B; p_k = tasks[p_task].k; }
    }
    if (p_task < grp.task_end) {
      double* sA = smem + stage * STAGE_DOUBLES;
      double* sB = sA + SLAB_DOUBLES;
      load_slab<OPA == TADEV_OP_N, VEC>(sA, pA, OPA == TADEV_OP_N ? p_k : M, m0, M, p_k0, p_k, tid);
      load_slab<OPB == TADEV_OP_T, VEC>(sB, pB, OPB == TADEV_OP_N ? N : p_k, n0, N, p_k0, p_k, tid);
      p_k0 += BK;
    }
    cp_async_commit();
  };

#pragma unroll
  for (int st = 0; st < STAGES - 1; ++st) produce(st);

  for (int it = 0; it < total_slabs; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    produce((it + STAGES - 1) % STAGES);
    const double* sA = smem + (it % STAGES) * STAGE_DOUBLES;
    const double* sB = sA + SLAB_DOUBLES;
#pragma unroll
    for (int s = 0; s < BK / 4; ++s) {
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
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: C = alpha*acc (+ C). Lane owns C[row g][cols 2t, 2t+1] of each 8x8 fragment.
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

template <int OPA, int OPB, int VEC>
int launch_variant(cudaStream_t s, const tadev_gemm_group* d_groups, int ngroups,
                   const tadev_gemm_task* d_tasks, const int32_t* d_tile_prefix, int total_cta_tiles,
                   double alpha) {
  auto kern = gemm_grouped_f64_kernel<OPA, OPB, VEC>;
  // per-device attribute: set on every launch (see gemm_f64_ws.cu)
  TADEV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  GemmTimingHook& hook = tadev_gemm_timing_hook();
  if (hook.before) TADEV_CHECK_CUDA(cudaEventRecord(hook.before, s));
  kern<<<total_cta_tiles, NTHREADS, SMEM_BYTES, s>>>(d_groups, ngroups, d_tasks, d_tile_prefix, alpha);
  TADEV_CHECK_CUDA(cudaGetLastError());
  if (hook.after) TADEV_CHECK_CUDA(cudaEventRecord(hook.after, s));
  hook.before = hook.after = nullptr;
  return TADEV_OK;
}

}  // namespace

int launch_gemm_grouped_f64(tadev_ctx* ctx, cudaStream_t s, int opA, int opB, double alpha,
                            const tadev_gemm_group* d_groups, int ngroups, const tadev_gemm_task* d_tasks,
                            const int32_t* d_tile_prefix, int total_cta_tiles, bool aligned16) {
  if (ngroups == 0 || total_cta_tiles == 0) return TADEV_OK;
  ctx->launches++;
  const int key = (opA << 2) | (opB << 1) | (aligned16 ? 1 : 0);
  switch (key) {
    case 0: return launch_variant<0, 0, 1>(s, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, alpha);
    case 1: return launch_variant<0, 0, 2>(s, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, alpha);
    case 2: return launch_variant<0, 1, 1>(s, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, alpha);
    case 3: return launch_variant<0, 1, 2>(s, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, alpha);
    case 4: return launch_variant<1, 0, 1>(s, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, alpha);
    case 5: return launch_variant<1, 0, 2>(s, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, alpha);
    case 6: return launch_variant<1, 1, 1>(s, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, alpha);
    case 7: return launch_variant<1, 1, 2>(s, d_groups, ngroups, d_tasks, d_tile_prefix, total_cta_tiles, alpha);
  }
  tadev_set_error("launch_gemm_grouped_f64: bad op flags %d %d", opA, opB);
  return TADEV_EINVAL;
}

// Decide whether every operand of the batch satisfies the 16-byte cp.async contract.
static bool batch_aligned16(int opA, int opB, const tadev_gemm_group* groups, int ngroups,
                            const tadev_gemm_task* tasks) {
  for (int gi = 0; gi < ngroups; ++gi) {
    const tadev_gemm_group& G = groups[gi];
    for (int ti = G.task_begin; ti < G.task_end; ++ti) {
      const tadev_gemm_task& T = tasks[ti];
      const int lda = opA == TADEV_OP_N ? T.k : G.m;
      const int ldb = opB == TADEV_OP_N ? G.n : T.k;
      if ((lda & 1) || (ldb & 1)) return false;
      if ((reinterpret_cast<uintptr_t>(T.A) & 15) || (reinterpret_cast<uintptr_t>(T.B) & 15)) return false;
    }
  }
  return true;
}

extern "C" int tadev_gemm_grouped_f64(tadev_ctx* ctx, tadev_stream s_, int opA, int opB, double alpha,
                                      const tadev_gemm_group* h_groups, int ngroups,
                                      const tadev_gemm_task* h_tasks, int ntasks) {
  TADEV_REQUIRE(ctx, "tadev_gemm_grouped_f64: null ctx");
  TADEV_REQUIRE((opA == 0 || opA == 1) && (opB == 0 || opB == 1), "tadev_gemm_grouped_f64: bad op flags");
  TADEV_REQUIRE(ngroups >= 0 && ntasks >= 0, "tadev_gemm_grouped_f64: negative counts");
  if (ngroups == 0) return TADEV_OK;
  TADEV_REQUIRE(h_groups && (ntasks == 0 || h_tasks), "tadev_gemm_grouped_f64: null descriptor arrays");
  cudaStream_t s = (cudaStream_t)s_;
  // validate + count CTA tiles
  std::vector<int32_t> prefix(ngroups + 1);
  int64_t total = 0;
  for (int gi = 0; gi < ngroups; ++gi) {
    const tadev_gemm_group& G = h_groups[gi];
    TADEV_REQUIRE(G.m >= 0 && G.n >= 0, "group %d: negative extent", gi);
    TADEV_REQUIRE(G.task_begin >= 0 && G.task_begin <= G.task_end && G.task_end <= ntasks,
                  "group %d: task range [%d,%d) outside [0,%d)", gi, G.task_begin, G.task_end, ntasks);
    TADEV_REQUIRE(G.C || G.m == 0 || G.n == 0, "group %d: null result tile", gi);
    for (int ti = G.task_begin; ti < G.task_end; ++ti) {
      TADEV_REQUIRE(h_tasks[ti].k >= 0, "task %d: negative k", ti);
      TADEV_REQUIRE((h_tasks[ti].A && h_tasks[ti].B) || h_tasks[ti].k == 0 || G.m == 0 || G.n == 0,
                    "task %d: null argument tile", ti);
    }
    prefix[gi] = (int32_t)total;
    total += ceil_div64(G.m, kGemmBM) * ceil_div64(G.n, kGemmBN);
    TADEV_REQUIRE(total < (1ll << 31), "tadev_gemm_grouped_f64: too many CTA tiles");
  }
  prefix[ngroups] = (int32_t)total;
  if (total == 0) return TADEV_OK;
  const bool al = batch_aligned16(opA, opB, h_groups, ngroups, h_tasks);
  if (al && !ctx->force_generic_gemm)  // fast path: persistent warp-specialised TMA kernel
    return launch_gemm_grouped_f64_ws(ctx, s, opA, opB, alpha, h_groups, ngroups, h_tasks, ntasks, (int)total);
  const size_t gb = sizeof(tadev_gemm_group) * (size_t)ngroups;
  const size_t tb = sizeof(tadev_gemm_task) * (size_t)ntasks;
  const size_t pb = sizeof(int32_t) * (size_t)(ngroups + 1);
  const size_t off_t = (gb + 15) & ~size_t(15);
  const size_t off_p = (off_t + tb + 15) & ~size_t(15);
  StageLease L;
  int rc = L.acquire(ctx, s, off_p + pb);
  if (rc) return rc;
  void *h = L.h, *d = L.d;
  memcpy(h, h_groups, gb);
  if (tb) memcpy((char*)h + off_t, h_tasks, tb);
  memcpy((char*)h + off_p, prefix.data(), pb);
  TADEV_CHECK_CUDA(cudaMemcpyAsync(d, h, off_p + pb, cudaMemcpyHostToDevice, s));
  rc = launch_gemm_grouped_f64(ctx, s, opA, opB, alpha, (const tadev_gemm_group*)d, ngroups,
                               (const tadev_gemm_task*)((char*)d + off_t), (const int32_t*)((char*)d + off_p),
                               (int)total, al);
  return rc;  // ~StageLease records `done` after the launch and returns the slot
}

extern "C" int tadev_gemm_grouped_f64_dev(tadev_ctx* ctx, tadev_stream s, int opA, int opB, double alpha,
                                          const tadev_gemm_group* d_groups, int ngroups,
                                          const tadev_gemm_task* d_tasks, const int32_t* d_tile_prefix,
                                          int total_cta_tiles) {
  TADEV_REQUIRE(ctx, "tadev_gemm_grouped_f64_dev: null ctx");
  TADEV_REQUIRE((opA == 0 || opA == 1) && (opB == 0 || opB == 1), "tadev_gemm_grouped_f64_dev: bad op flags");
  // device-resident descriptors cannot be inspected on the host: use the 8-byte staging path,
  // which has no alignment contract.
  return launch_gemm_grouped_f64(ctx, (cudaStream_t)s, opA, opB, alpha, d_groups, ngroups, d_tasks,
                                 d_tile_prefix, total_cta_tiles, false);
}

extern "C" int tadev_gemm_f64(tadev_ctx* ctx, tadev_stream s, int opA, int opB, int m, int n, int k,
                              double alpha, const double* d_A, const double* d_B, double beta, double* d_C) {
  TADEV_REQUIRE(ctx, "tadev_gemm_f64: null ctx");
  TADEV_REQUIRE(beta == 0.0 || beta == 1.0,
                "tadev_gemm_f64: beta must be 0 or 1 (ContractReduce only seeds or accumulates)");
  TADEV_REQUIRE(m >= 0 && n >= 0 && k >= 0, "tadev_gemm_f64: negative extent");
  tadev_gemm_task T{d_A, d_B, k, 0};
  tadev_gemm_group G{d_C, m, n, 0, 1, beta == 1.0 ? 1 : 0, 0};
  return tadev_gemm_grouped_f64(ctx, s, opA, opB, alpha, &G, 1, &T, 1);
}
