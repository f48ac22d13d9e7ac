// permute.cu — out-of-place tile permutation, HBM-bandwidth bound.
//
// Replaces detail::permute (reference: src/TiledArray/tensor/permute.h:118-209; dimension
// fusion follows the idea of fuse_dimensions, :51-94) and the per-call plan/execute/destroy of
// LibreTT (external/librett.h:81-111). Semantics are TiledArray's image convention
// (permutation.h:69-79): out.extent[perm[i]] = in.extent[i] and out[p(idx)] = in[idx].
//
// Planning is a few integer ops on the host per call (no plan objects): drop unit extents, fuse
// input modes that stay adjacent and ordered in the output, then pick one of three kernels
//   1. identity after fusion           -> cudaMemcpyAsync D2D
//   2. last mode stays last            -> row-copy kernel (16-byte vectorised when aligned)
//   3. otherwise                       -> shared-memory tiled transpose over (a, b) where a is the
//      input-fastest mode and b the output-fastest mode; global reads are coalesced along a,
//      global writes along b, smem rows are padded by one element (conflict-free both ways).
#include <algorithm>
#include <cstdlib>

#include "common.h"

namespace {

constexpr int MAXR = 8;

struct PermParams {
  int R;                 // reduced rank
  int64_t ext[MAXR];     // input extents (reduced)
  int64_t sin[MAXR];     // input strides  (elements)
  int64_t sout[MAXR];    // output stride of the output mode each input mode maps to
  int a, b;              // input-fastest mode (R-1) and output-fastest mode (as input-mode index)
  int TA, TB;            // tile extents along a and b (powers of two)
  int64_t nTa, nTb;      // tiles along a and b
  int nother;            // number of remaining modes
  int other[MAXR];       // their input-mode indices, fastest varying first
  int64_t total;         // total elements
  int64_t row_len;       // row copy: elements per preserved row (already divided by the vector width)
  int64_t nrows;
};

// blockIdx.y selects the tile of a batch (all tiles of a batch share extents and permutation).
template <typename T>
__global__ void __launch_bounds__(256) transpose_tiled_kernel(const T* __restrict__ in0, T* __restrict__ out0,
                                                              const void* const* __restrict__ ins,
                                                              void* const* __restrict__ outs, PermParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const T* __restrict__ in = ins ? static_cast<const T*>(ins[blockIdx.y]) : in0;
  T* __restrict__ out = outs ? static_cast<T*>(outs[blockIdx.y]) : out0;
  const int TA = p.TA, TB = p.TB;
  const int ldt = TA + 1;
  int64_t bid = blockIdx.x;
  const int64_t ta = bid % p.nTa; bid /= p.nTa;
  const int64_t tb = bid % p.nTb; bid /= p.nTb;
  int64_t base_in = 0, base_out = 0;
  for (int d = 0; d < p.nother; ++d) {
    const int m = p.other[d];
    const int64_t i = bid % p.ext[m];
    bid /= p.ext[m];
    base_in += i * p.sin[m];
    base_out += i * p.sout[m];
  }
  const int64_t a0 = ta * TA, b0 = tb * TB;
  const int ra = (int)min((int64_t)TA, p.ext[p.a] - a0);  // valid extent of this tile along a / b
  const int rb = (int)min((int64_t)TB, p.ext[p.b] - b0);
  const int64_t sin_b = p.sin[p.b], sout_a = p.sout[p.a];
  const T* __restrict__ src = in + base_in + b0 * sin_b + a0;
  T* __restrict__ dst = out + base_out + a0 * sout_a + b0;
  const int la = 31 - __clz(TA), lb = 31 - __clz(TB);
  const int n = TA * TB;
  // read: consecutive threads walk the input-fastest mode a
#pragma unroll 8
  for (int idx = threadIdx.x; idx < n; idx += 256) {
    const int ia = idx & (TA - 1), ib = idx >> la;
    if (ia < ra && ib < rb) tile[ib * ldt + ia] = src[(int64_t)ib * sin_b + ia];
  }
  __syncthreads();
  // write: consecutive threads walk the output-fastest mode b
#pragma unroll 8
  for (int idx = threadIdx.x; idx < n; idx += 256) {
    const int ib = idx & (TB - 1), ia = idx >> lb;
    if (ia < ra && ib < rb) dst[(int64_t)ia * sout_a + ib] = tile[ib * ldt + ia];
  }
}

// Fast path of the tiled transpose for 8-byte elements and full-size 64 x 32 tiles (the shape the
// planner picks whenever both modes are long): ncu on the generic kernel showed 71 warp
// instructions per 32 elements (64-bit index math and predicates per element, 64-bit divisions per
// block) and the issue slots, not DRAM, as the limiter (issue 60 %, dram 60 %). Here every thread
// keeps a fixed (column, first row) and walks rows by a constant stride: one add per access,
// block decode in 32-bit, all loads of a thread in flight before the first shared-memory store.
// VEC: 16-byte global loads/stores (two elements along a on the way in, two along b on the way out).
template <int TA, int TB, bool VEC>
__global__ void __launch_bounds__(256) transpose_fast_kernel(const uint64_t* __restrict__ in0, uint64_t* __restrict__ out0,
                                                             const void* const* __restrict__ ins,
                                                             void* const* __restrict__ outs, PermParams p) {
  constexpr int LDT = TA + 1;
  __shared__ uint64_t tile[TB * LDT];
  const uint64_t* __restrict__ in = ins ? static_cast<const uint64_t*>(ins[blockIdx.y]) : in0;
  uint64_t* __restrict__ out = outs ? static_cast<uint64_t*>(outs[blockIdx.y]) : out0;
  uint32_t bid = blockIdx.x;
  const uint32_t nTa = (uint32_t)p.nTa, nTb = (uint32_t)p.nTb;
  const uint32_t ta = bid % nTa; bid /= nTa;
  const uint32_t tb = bid % nTb; bid /= nTb;
  int64_t base_in = 0, base_out = 0;
  for (int d = 0; d < p.nother; ++d) {
    const int m = p.other[d];
    const uint32_t e = (uint32_t)p.ext[m];
    const uint32_t i = bid % e;
    bid /= e;
    base_in += (int64_t)i * p.sin[m];
    base_out += (int64_t)i * p.sout[m];
  }
  const int64_t a0 = (int64_t)ta * TA, b0 = (int64_t)tb * TB;
  const int ra = (int)min((int64_t)TA, p.ext[p.a] - a0), rb = (int)min((int64_t)TB, p.ext[p.b] - b0);
  const bool full = (ra == TA) && (rb == TB);
  const int64_t sin_b = p.sin[p.b], sout_a = p.sout[p.a];
  const uint64_t* __restrict__ src = in + base_in + b0 * sin_b + a0;
  uint64_t* __restrict__ dst = out + base_out + a0 * sout_a + b0;
  const int tid = threadIdx.x;
  if (VEC) {
    // in: TA/2 threads per row, 256/(TA/2) rows per pass
    constexpr int TPR = TA / 2, RPP = 256 / TPR, NP = TB / RPP;
    const int ia = (tid % TPR) * 2, ib0 = tid / TPR;
    ulonglong2 v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int ib = ib0 + i * RPP;
      if (full || (ia < ra && ib < rb)) v[i] = *reinterpret_cast<const ulonglong2*>(src + (int64_t)ib * sin_b + ia);  // ra is even
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int ib = ib0 + i * RPP;
      tile[ib * LDT + ia] = v[i].x;
      tile[ib * LDT + ia + 1] = v[i].y;
    }
    __syncthreads();
    // out: TB/2 threads per output row, 256/(TB/2) rows per pass
    constexpr int TPO = TB / 2, OPP = 256 / TPO, NO = TA / OPP;
    const int ob = (tid % TPO) * 2, oa0 = tid / TPO;
#pragma unroll
    for (int i = 0; i < NO; ++i) {
      const int oa = oa0 + i * OPP;
      ulonglong2 w;
      w.x = tile[ob * LDT + oa];
      w.y = tile[(ob + 1) * LDT + oa];
      if (full || (oa < ra && ob < rb)) *reinterpret_cast<ulonglong2*>(dst + (int64_t)oa * sout_a + ob) = w;  // rb is even
    }
  } else {
    constexpr int RPP = 256 / TA, NP = TB / RPP;
    const int ia = tid % TA, ib0 = tid / TA;
    uint64_t v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int ib = ib0 + i * RPP;
      if (full || (ia < ra && ib < rb)) v[i] = src[(int64_t)ib * sin_b + ia];
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) tile[(ib0 + i * RPP) * LDT + ia] = v[i];
    __syncthreads();
    constexpr int OPP = 256 / TB, NO = TA / OPP;
    const int ob = tid % TB, oa0 = tid / TB;
#pragma unroll
    for (int i = 0; i < NO; ++i) {
      const int oa = oa0 + i * OPP;
      const uint64_t w = tile[ob * LDT + oa];
      if (full || (oa < ra && ob < rb)) dst[(int64_t)oa * sout_a + ob] = w;
    }
  }
}

// Rows (the preserved, contiguous last mode) are copied whole; V is the vector type, I the index
// type (32-bit whenever the tile has < 2^31 vector elements: 64-bit div/mod costs ~10x more).
// Each thread moves 4 independent chunks per iteration (loads first, then stores).
template <typename V, typename I>
__global__ void __launch_bounds__(256) rowcopy_kernel(const V* __restrict__ in0, V* __restrict__ out0,
                                                      const void* const* __restrict__ ins,
                                                      void* const* __restrict__ outs, PermParams p) {
  const V* __restrict__ in = ins ? static_cast<const V*>(ins[blockIdx.y]) : in0;
  V* __restrict__ out = outs ? static_cast<V*>(outs[blockIdx.y]) : out0;
  const I row_len = (I)p.row_len;
  const I total = (I)(p.nrows * p.row_len);
  const I stride = (I)gridDim.x * (I)blockDim.x;
  I ext[MAXR], sin[MAXR];
#pragma unroll
  for (int d = 0; d < MAXR; ++d) { const int m = d < p.nother ? p.other[d] : 0; ext[d] = (I)p.ext[m]; sin[d] = (I)p.sin[m]; }
  const int nother = p.nother;
  auto src_of = [&](I q) -> I {
    I row = q / row_len;
    const I c = q - row * row_len;
    // q enumerates the OUTPUT in order: decompose the output row ordinal into output-mode
    // indices. other[] lists the non-row modes by increasing output stride.
    I off_in = c;
#pragma unroll
    for (int d = 0; d < MAXR; ++d) {
      if (d < nother) {
        const I nxt = row / ext[d];
        off_in += (row - nxt * ext[d]) * sin[d];
        row = nxt;
      }
    }
    return off_in;
  };
  for (I q = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; q < total; q += 4 * stride) {
    V v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const I qq = q + (I)u * stride;
      if (qq < total) v[u] = in[src_of(qq)];  // sin[] pre-divided by the vector width for the row mode
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const I qq = q + (I)u * stride;
      if (qq < total) out[qq] = v[u];
    }
  }
}

int64_t pow2_ceil(int64_t x) {
  int64_t r = 1;
  while (r < x) r *= 2;
  return r;
}

enum PlanKind { kCopy = 0, kRowCopy = 1, kTiled = 2 };
struct Plan {
  PlanKind kind;
  PermParams p;
  int64_t total;
  int elem_bytes;
  bool row_vec16_possible;  // row bytes % 16 == 0 (pointer alignment is checked per tile)
};

// Host planning shared by the single-tile and the batched entry.
int plan_permute(int rank, const int64_t* extent, const int32_t* perm, int elem_bytes, Plan* plan) {
  TADEV_REQUIRE(rank >= 0 && rank <= 16, "tadev_permute: rank %d unsupported", rank);
  TADEV_REQUIRE(elem_bytes == 4 || elem_bytes == 8 || elem_bytes == 16, "tadev_permute: elem_bytes %d", elem_bytes);
  TADEV_REQUIRE(rank == 0 || (extent && perm), "tadev_permute: null extent/perm");
  {
    uint32_t seen = 0;
    for (int i = 0; i < rank; ++i) {
      TADEV_REQUIRE(perm[i] >= 0 && perm[i] < rank && !(seen >> perm[i] & 1), "tadev_permute: not a permutation");
      seen |= 1u << perm[i];
      TADEV_REQUIRE(extent[i] >= 0, "tadev_permute: negative extent");
    }
  }
  int64_t total = 1;
  for (int i = 0; i < rank; ++i) total *= extent[i];
  plan->total = total;
  plan->elem_bytes = elem_bytes;
  plan->kind = kCopy;
  plan->row_vec16_possible = false;
  if (total == 0) return TADEV_OK;
  // output strides per output mode
  int64_t out_ext[16], out_stride[16];
  for (int i = 0; i < rank; ++i) out_ext[perm[i]] = extent[i];
  {
    int64_t st = 1;
    for (int j = rank - 1; j >= 0; --j) { out_stride[j] = st; st *= out_ext[j]; }
  }
  // reduce: drop unit modes; fuse input modes that stay adjacent and ordered in the output
  int R = 0;
  int64_t ext[16], sout[16];
  for (int i = 0; i < rank; ++i) {
    if (extent[i] == 1) continue;
    if (R > 0 && sout[R - 1] == out_stride[perm[i]] * extent[i]) {
      ext[R - 1] *= extent[i];
      sout[R - 1] = out_stride[perm[i]];
    } else {
      ext[R] = extent[i]; sout[R] = out_stride[perm[i]]; ++R;
    }
  }
  if (R <= 1) return TADEV_OK;  // identity
  TADEV_REQUIRE(R <= MAXR, "tadev_permute: reduced rank %d > %d", R, MAXR);
  PermParams& p = plan->p;
  p = PermParams{};
  p.R = R;
  p.total = total;
  {
    int64_t st = 1;
    for (int i = R - 1; i >= 0; --i) { p.ext[i] = ext[i]; p.sout[i] = sout[i]; p.sin[i] = st; st *= ext[i]; }
  }
  if (p.sout[R - 1] == 1) {
    // last mode preserved: row copy. Enumerate rows in OUTPUT order: sort the other modes by sout.
    plan->kind = kRowCopy;
    int idx[MAXR], n = 0;
    for (int i = 0; i < R - 1; ++i) idx[n++] = i;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j)
        if (p.sout[idx[j]] < p.sout[idx[i]]) { int tswap = idx[i]; idx[i] = idx[j]; idx[j] = tswap; }
    p.nother = n;
    for (int i = 0; i < n; ++i) p.other[i] = idx[i];
    p.row_len = p.ext[R - 1];
    p.nrows = total / p.row_len;
    plan->row_vec16_possible = (p.row_len * elem_bytes) % 16 == 0;
    return TADEV_OK;
  }
  // general: tiled transpose over a = R-1 and b = the mode with sout == 1
  plan->kind = kTiled;
  p.a = R - 1;
  p.b = -1;
  for (int i = 0; i < R; ++i)
    if (p.sout[i] == 1) p.b = i;
  TADEV_REQUIRE(p.b >= 0 && p.b != p.a, "tadev_permute: internal planning error");
  // tile: up to 4096 elements (<= 32 KiB of 8-byte elements); 64 x 64 when both modes are long,
  // otherwise the short mode gets its full (power-of-two padded) extent and the other widens
  int64_t budget = 2048;
  static const char* env = getenv("TADEV_PERM_TILE");  // "TAxTB" override for experiments
  int64_t TA = 64, TB = 32;  // measured best on B200: 4.85 TB/s vs 4.3 for 64x64 (profiles/r01_permute_*.json)
  if (env && sscanf(env, "%ldx%ld", &TA, &TB) == 2) { budget = TA * TB; }
  TB = std::min<int64_t>(pow2_ceil(p.ext[p.b]), TB);
  TA = std::min<int64_t>(pow2_ceil(p.ext[p.a]), budget / TB);
  if (TA * TB < budget) TB = std::min<int64_t>(pow2_ceil(p.ext[p.b]), budget / TA);
  p.TA = (int)TA; p.TB = (int)TB;
  p.nTa = ceil_div64(p.ext[p.a], TA);
  p.nTb = ceil_div64(p.ext[p.b], TB);
  p.nother = 0;
  for (int i = R - 2; i >= 0; --i)
    if (i != p.b) p.other[p.nother++] = i;
  return TADEV_OK;
}

template <typename T>
int launch_tiled(tadev_ctx* ctx, cudaStream_t s, const PermParams& p, const void* in, void* out,
                 const void* const* d_ins, void* const* d_outs, int ntiles, bool ptrs_al16) {
  int64_t blocks = p.nTa * p.nTb;
  for (int d = 0; d < p.nother; ++d) blocks *= p.ext[p.other[d]];
  TADEV_REQUIRE(blocks < (1ll << 31), "tadev_permute: tile too large for one launch");
  static const int fast_env = getenv("TADEV_PERM_FAST") ? atoi(getenv("TADEV_PERM_FAST")) : 2;  // 0 generic, 1 scalar, 2 vector
  bool ext32 = true;
  for (int d = 0; d < p.R; ++d) ext32 = ext32 && p.ext[d] < (1ll << 31);
  const bool fast_shape = (p.TA * p.TB == 2048) && p.TA >= 8 && p.TA <= 256;
  if (sizeof(T) == 8 && fast_shape && fast_env > 0 && ext32) {
    // 16-byte accesses need every row start 16-byte aligned on both sides: even extents along a and
    // b make all strides even (the other modes' strides are multiples of those extents)
    bool vec = fast_env > 1 && ptrs_al16 && (p.ext[p.a] % 2 == 0) && (p.ext[p.b] % 2 == 0);
    for (int d = 0; d < p.R && vec; ++d) {
      if (d != p.a && (p.sin[d] & 1)) vec = false;
      if (d != p.b && (p.sout[d] & 1)) vec = false;
    }
    const dim3 grid((unsigned)blocks, (unsigned)ntiles);
    const uint64_t* i8 = (const uint64_t*)in;
    uint64_t* o8 = (uint64_t*)out;
#define TADEV_FAST_CASE(TA_, TB_)                                                                        \
  case TA_:                                                                                              \
    if (vec) transpose_fast_kernel<TA_, TB_, true><<<grid, 256, 0, s>>>(i8, o8, d_ins, d_outs, p);       \
    else transpose_fast_kernel<TA_, TB_, false><<<grid, 256, 0, s>>>(i8, o8, d_ins, d_outs, p);         \
    break;
    switch (p.TA) {
      TADEV_FAST_CASE(8, 256)
      TADEV_FAST_CASE(16, 128)
      TADEV_FAST_CASE(32, 64)
      TADEV_FAST_CASE(64, 32)
      TADEV_FAST_CASE(128, 16)
      TADEV_FAST_CASE(256, 8)
      default: break;
    }
#undef TADEV_FAST_CASE
    ctx->launches++;
    TADEV_CHECK_CUDA(cudaGetLastError());
    return TADEV_OK;
  }
  const size_t smem = sizeof(T) * (size_t)(p.TA + 1) * p.TB;
  auto kern = transpose_tiled_kernel<T>;
  if (smem > 48 * 1024) TADEV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)blocks, (unsigned)ntiles), 256, smem, s>>>((const T*)in, (T*)out, d_ins, d_outs, p);
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}

template <typename V>
int launch_rowcopy(tadev_ctx* ctx, cudaStream_t s, const PermParams& p, const void* in, void* out,
                   const void* const* d_ins, void* const* d_outs, int ntiles) {
  const int64_t total = p.nrows * p.row_len;
  int64_t blocks = ceil_div64(total, 256 * 4);
  const int64_t maxb = std::max<int64_t>(1, (int64_t)ctx->num_sms * 16 / ntiles);
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  if (total + 4 * blocks * 256 < (int64_t(1) << 31))
    rowcopy_kernel<V, uint32_t><<<dim3((unsigned)blocks, (unsigned)ntiles), 256, 0, s>>>((const V*)in, (V*)out, d_ins, d_outs, p);
  else
    rowcopy_kernel<V, int64_t><<<dim3((unsigned)blocks, (unsigned)ntiles), 256, 0, s>>>((const V*)in, (V*)out, d_ins, d_outs, p);
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}

int launch_plan(tadev_ctx* ctx, cudaStream_t s, Plan& plan, bool ptrs_al16, const void* in, void* out,
                const void* const* d_ins, void* const* d_outs, int ntiles) {
  PermParams p = plan.p;
  const int eb = plan.elem_bytes;
  if (plan.kind == kRowCopy) {
    if (plan.row_vec16_possible && ptrs_al16) {
      const int vec = 16 / eb;
      p.row_len /= vec;
      for (int i = 0; i < p.R - 1; ++i) p.sin[i] /= vec;
      return launch_rowcopy<uint4>(ctx, s, p, in, out, d_ins, d_outs, ntiles);
    }
    if (eb == 4) return launch_rowcopy<uint32_t>(ctx, s, p, in, out, d_ins, d_outs, ntiles);
    if (eb == 8) return launch_rowcopy<uint64_t>(ctx, s, p, in, out, d_ins, d_outs, ntiles);
    return launch_rowcopy<uint4>(ctx, s, p, in, out, d_ins, d_outs, ntiles);
  }
  if (eb == 4) return launch_tiled<uint32_t>(ctx, s, p, in, out, d_ins, d_outs, ntiles, ptrs_al16);
  if (eb == 8) return launch_tiled<uint64_t>(ctx, s, p, in, out, d_ins, d_outs, ntiles, ptrs_al16);
  return launch_tiled<uint4>(ctx, s, p, in, out, d_ins, d_outs, ntiles, ptrs_al16);
}

}  // namespace

extern "C" int tadev_permute(tadev_ctx* ctx, tadev_stream s_, int rank, const int64_t* extent,
                             const int32_t* perm, int elem_bytes, const void* d_in, void* d_out) {
  TADEV_REQUIRE(ctx, "tadev_permute: null ctx");
  cudaStream_t s = (cudaStream_t)s_;
  Plan plan;
  int rc = plan_permute(rank, extent, perm, elem_bytes, &plan);
  if (rc) return rc;
  if (plan.total == 0) return TADEV_OK;
  TADEV_REQUIRE(d_in && d_out, "tadev_permute: null tile");
  TADEV_REQUIRE(d_in != d_out, "tadev_permute: in-place permutation is not supported");
  if (plan.kind == kCopy) {
    TADEV_CHECK_CUDA(cudaMemcpyAsync(d_out, d_in, (size_t)plan.total * elem_bytes, cudaMemcpyDeviceToDevice, s));
    return TADEV_OK;
  }
  const bool al16 = ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
  return launch_plan(ctx, s, plan, al16, d_in, d_out, nullptr, nullptr, 1);
}

// Many tiles with identical extents and permutation in ONE launch (the argument-tile permutes of a
// contraction, dist_eval/array_eval.h:42,170: every tile of an operand gets the same permutation).
extern "C" int tadev_permute_batched(tadev_ctx* ctx, tadev_stream s_, int rank, const int64_t* extent,
                                     const int32_t* perm, int elem_bytes, int ntiles, const void* const* h_in,
                                     void* const* h_out) {
  TADEV_REQUIRE(ctx, "tadev_permute_batched: null ctx");
  TADEV_REQUIRE(ntiles >= 0, "tadev_permute_batched: negative tile count");
  if (ntiles == 0) return TADEV_OK;
  TADEV_REQUIRE(h_in && h_out, "tadev_permute_batched: null pointer arrays");
  cudaStream_t s = (cudaStream_t)s_;
  Plan plan;
  int rc = plan_permute(rank, extent, perm, elem_bytes, &plan);
  if (rc) return rc;
  if (plan.total == 0) return TADEV_OK;
  bool al16 = true;
  for (int i = 0; i < ntiles; ++i) {
    TADEV_REQUIRE(h_in[i] && h_out[i] && h_in[i] != h_out[i], "tadev_permute_batched: tile %d null or in place", i);
    al16 = al16 && (((reinterpret_cast<uintptr_t>(h_in[i]) | reinterpret_cast<uintptr_t>(h_out[i])) & 15) == 0);
  }
  if (plan.kind == kCopy) {
    for (int i = 0; i < ntiles; ++i)
      TADEV_CHECK_CUDA(cudaMemcpyAsync(h_out[i], h_in[i], (size_t)plan.total * elem_bytes, cudaMemcpyDeviceToDevice, s));
    return TADEV_OK;
  }
  for (int first = 0; first < ntiles; first += 32768) {  // gridDim.y limit
    const int n = std::min(32768, ntiles - first);
    StageLease L;
    rc = L.acquire(ctx, s, sizeof(void*) * 2 * (size_t)n);
    if (rc) return rc;
    void *h = L.h, *d = L.d;
    memcpy(h, h_in + first, sizeof(void*) * n);
    memcpy((char*)h + sizeof(void*) * n, h_out + first, sizeof(void*) * n);
    TADEV_CHECK_CUDA(cudaMemcpyAsync(d, h, sizeof(void*) * 2 * (size_t)n, cudaMemcpyHostToDevice, s));
    rc = launch_plan(ctx, s, plan, al16, nullptr, nullptr, (const void* const*)d, (void* const*)((char*)d + sizeof(void*) * n), n);
    if (rc) return rc;
  }  // ~StageLease records `done` after the launch and returns the slot
  return TADEV_OK;
}

// ---- small elementwise helpers on the contraction path -------------------------------------
namespace {
__global__ void add_to_kernel(double* __restrict__ r, const double* __restrict__ a, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) r[i] += a[i];
}
__global__ void scale_kernel(double* __restrict__ x, size_t n, double f) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] *= f;
}
// Squared Frobenius norms, two deterministic stages (bound: HBM): stage 1 reduces one chunk of kSqChunk doubles per
// CTA (16-byte loads, 8 in flight per thread, fixed intra-chunk order), stage 2 adds a tile's chunk partials in
// ascending order. The association depends only on the tile's own size, never on how many tiles share the launch,
// so a tile's norm - and the float shape built from it - is reproducible bit for bit.
constexpr int kSqChunk = 16384;
__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const double* const* __restrict__ ptrs, const int64_t* __restrict__ sizes,
                                                             double* __restrict__ partial, int nchunks_max) {
  const int tile = blockIdx.y;
  const int64_t n = sizes[tile];
  const int64_t lo = (int64_t)blockIdx.x * kSqChunk;
  if (lo >= n) return;
  const double* __restrict__ p = ptrs[tile] + lo;
  const int len = (int)min((int64_t)kSqChunk, n - lo);
  double acc = 0.0;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const double2* __restrict__ p2 = reinterpret_cast<const double2*>(p);
    const int len2 = len >> 1;
    int i = threadIdx.x;
    for (; i + 7 * 256 < len2; i += 8 * 256) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = p2[i + u * 256];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u].x * v[u].x + v[u].y * v[u].y;
    }
    for (; i < len2; i += 256) { const double2 v = p2[i]; acc += v.x * v.x + v.y * v.y; }
    if ((len & 1) && threadIdx.x == 0) { const double v = p[len - 1]; acc += v * v; }
  } else {
    for (int i = threadIdx.x; i < len; i += 256) { const double v = p[i]; acc += v * v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w];
    partial[(size_t)tile * nchunks_max + blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) sqnorm_final_kernel(const int64_t* __restrict__ sizes, const double* __restrict__ partial,
                                                           int nchunks_max, int ntiles, double* __restrict__ out) {
  const int tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= ntiles) return;
  const int64_t n = sizes[tile];
  const int nch = (int)((n + kSqChunk - 1) / kSqChunk);
  double t = 0.0;
  for (int c = 0; c < nch; ++c) t += partial[(size_t)tile * nchunks_max + c];
  out[tile] = t;
}
// splitmix64-based counter RNG: value depends only on (seed, global element offset)
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void fill_uniform_kernel(double* __restrict__ x, size_t n, uint64_t seed, uint64_t offset) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + (offset + i));
    x[i] = (double)(r >> 11) * (2.0 / 9007199254740992.0) - 1.0;  // [-1, 1)
  }
}
int grid_for(tadev_ctx* ctx, size_t n) {
  size_t b = (n + 1023) / 1024;
  size_t maxb = (size_t)ctx->num_sms * 16;
  if (b > maxb) b = maxb;
  return b ? (int)b : 1;
}
}  // namespace

extern "C" int tadev_add_to_f64(tadev_ctx* ctx, tadev_stream s, size_t n, double* d_result, const double* d_arg) {
  TADEV_REQUIRE(ctx, "tadev_add_to_f64: null ctx");
  if (!n) return TADEV_OK;
  TADEV_REQUIRE(d_result && d_arg, "tadev_add_to_f64: null tile");
  add_to_kernel<<<grid_for(ctx, n), 256, 0, (cudaStream_t)s>>>(d_result, d_arg, n);
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}
extern "C" int tadev_scale_f64(tadev_ctx* ctx, tadev_stream s, size_t n, double* d_x, double factor) {
  TADEV_REQUIRE(ctx, "tadev_scale_f64: null ctx");
  if (!n) return TADEV_OK;
  TADEV_REQUIRE(d_x, "tadev_scale_f64: null tile");
  scale_kernel<<<grid_for(ctx, n), 256, 0, (cudaStream_t)s>>>(d_x, n, factor);
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}
extern "C" int tadev_tile_sqnorms_f64(tadev_ctx* ctx, tadev_stream s, int ntiles, const double* const* d_ptrs,
                                      const int64_t* d_sizes, int64_t max_elems, double* d_out) {
  TADEV_REQUIRE(ctx, "tadev_tile_sqnorms_f64: null ctx");
  if (ntiles <= 0) return TADEV_OK;
  TADEV_REQUIRE(d_ptrs && d_sizes && d_out && max_elems >= 0, "tadev_tile_sqnorms_f64: bad arguments");
  TADEV_REQUIRE(ntiles <= 65535, "tadev_tile_sqnorms_f64: at most 65535 tiles per call");
  const int64_t nch = std::max<int64_t>(1, (max_elems + kSqChunk - 1) / kSqChunk);
  TADEV_REQUIRE(nch < (1ll << 31), "tadev_tile_sqnorms_f64: tile too large");
  double* partial = nullptr;
  int rc = tadev_alloc(ctx, (size_t)ntiles * (size_t)nch * 8, (void**)&partial, s);
  if (rc) return rc;
  sqnorm_partial_kernel<<<dim3((unsigned)nch, (unsigned)ntiles), 256, 0, (cudaStream_t)s>>>(d_ptrs, d_sizes, partial, (int)nch);
  sqnorm_final_kernel<<<(ntiles + 255) / 256, 256, 0, (cudaStream_t)s>>>(d_sizes, partial, (int)nch, ntiles, d_out);
  ctx->launches += 2;
  cudaError_t e = cudaGetLastError();
  tadev_free(ctx, partial, s);
  TADEV_CHECK_CUDA(e);
  return TADEV_OK;
}
// one tile, scalar handed to the host (Tensor::squared_norm): stages (ptr, size) on the device, reduces, copies back
extern "C" int tadev_sqnorm_f64(tadev_ctx* ctx, tadev_stream s, size_t n, const double* d_x, double* h_out) {
  TADEV_REQUIRE(ctx && h_out, "tadev_sqnorm_f64: null");
  *h_out = 0.0;
  if (!n) return TADEV_OK;
  TADEV_REQUIRE(d_x, "tadev_sqnorm_f64: null tile");
  struct { const double* p; int64_t n; double out; } h{d_x, (int64_t)n, 0.0};
  char* d = nullptr;
  int rc = tadev_alloc(ctx, sizeof(h), (void**)&d, s);
  if (rc) return rc;
  rc = tadev_memcpy_h2d(ctx, d, &h, 16, s);  // pageable source: staged before the call returns
  if (!rc) rc = tadev_tile_sqnorms_f64(ctx, s, 1, (const double* const*)d, (const int64_t*)(d + 8), (int64_t)n, (double*)(d + 16));
  if (!rc) rc = tadev_memcpy_d2h(ctx, h_out, d + 16, 8, s);
  if (!rc) rc = tadev_stream_sync(ctx, s);
  tadev_free(ctx, d, s);
  return rc;
}
extern "C" int tadev_fill_uniform_f64(tadev_ctx* ctx, tadev_stream s, double* d_x, size_t n, uint64_t seed,
                                      uint64_t offset) {
  TADEV_REQUIRE(ctx, "tadev_fill_uniform_f64: null ctx");
  if (!n) return TADEV_OK;
  TADEV_REQUIRE(d_x, "tadev_fill_uniform_f64: null");
  fill_uniform_kernel<<<grid_for(ctx, n), 256, 0, (cudaStream_t)s>>>(d_x, n, seed, offset);
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}
