// summa.cu — communicators and the SUMMA contraction driver.
//
// Replaces detail::Summa (reference: src/TiledArray/dist_eval/contraction_eval.h:55-2027):
//   internal_eval :1907-2025, StepTask::run :1559-1620, get_col/get_row :655-676,
//   bcast_col/bcast_row :779-806 (world.gop.bcast of serialized tiles over MPI),
//   contract :1311-1384 (+ ReducePairTask, reduce_task.h), finalize :1180-1269.
//
// B200 process model: one process per GPU on a Pr x Pc grid (ProcGrid); shapes are replicated, so
// every rank derives the same panel contents and no sizes are ever exchanged. Per K step the
// non-zero tiles A(i,k), i = r (mod Pr), form a packed panel that is ncclBroadcast along the grid
// row from column k % Pc; B(k,j), j = c (mod Pc), travels along the grid column from row k % Pr.
// Broadcasts run on a high-priority communication stream into a ring of panel buffers; the
// contraction of a window of steps is ONE grouped DMMA launch on the compute stream (all pairs
// of all steps of the window, chained per result tile and accumulated in place), so the
// broadcasts of window w+1 overlap the GEMM batch of window w. With P == 1 the whole contraction
// is a single launch and no panel buffers exist.
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>
#include <string>

#include "common.h"
#include "summa_schedule.h"
#include "tilelist.h"

#define TADEV_CHECK_NCCL(expr)                                                                   \
  do {                                                                                           \
    ncclResult_t r__ = (expr);                                                                   \
    if (r__ != ncclSuccess) {                                                                    \
      tadev_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(r__));    \
      return TADEV_ENCCL;                                                                        \
    }                                                                                            \
  } while (0)

extern "C" int tadev_comm_unique_id(void* out128) {
  TADEV_REQUIRE(out128, "tadev_comm_unique_id: null");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  TADEV_CHECK_NCCL(ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return TADEV_OK;
}

extern "C" int tadev_comm_init(tadev_ctx* ctx, const void* unique_id128, int rank, int nranks, int Pr, int Pc) {
  TADEV_REQUIRE(ctx && unique_id128, "tadev_comm_init: null");
  TADEV_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "tadev_comm_init: bad rank %d of %d", rank, nranks);
  TADEV_REQUIRE(Pr >= 1 && Pc >= 1 && Pr * Pc <= nranks, "tadev_comm_init: grid %dx%d does not fit %d ranks", Pr, Pc, nranks);
  TADEV_REQUIRE(!ctx->world, "tadev_comm_init: communicators already initialised");
  TADEV_CHECK_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  // The persistent GEMM leaves `gemm_sm_reserve` SMs free and NCCL is capped to as many CTAs, so the
  // panel broadcasts of window w+1 are always resident next to the GEMM of window w.
  ctx->gemm_sm_reserve = (Pr * Pc > 1) ? 4 : 0;
  if (const char* e = getenv("TADEV_SM_RESERVE")) ctx->gemm_sm_reserve = atoi(e);
  ctx->sm_reserve_thin = ctx->gemm_sm_reserve;
  ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
  if (ctx->gemm_sm_reserve > 0) { cfg.minCTAs = 1; cfg.maxCTAs = ctx->gemm_sm_reserve; }
  // NCCL's default on sm_90+ launches its CTAs as thread-block clusters of 4, which must be co-scheduled inside
  // one GPC: the few SMs the persistent GEMM leaves free are spread over the GPCs, so a clustered broadcast
  // kernel could not start before the GEMM's CTAs exited and every other window's panel traffic was exposed
  // (profiles/r02_n4_trace_cga4.log: bcast_done of windows 2n, 2n+1 7 ms after gemm_done of window 2n-1).
  cfg.cgaClusterSize = 1;
  if (const char* e = getenv("TADEV_NCCL_CGA")) cfg.cgaClusterSize = atoi(e);
  TADEV_CHECK_NCCL(ncclCommInitRankConfig(&ctx->world, nranks, id, rank, &cfg));
  ctx->rank = rank; ctx->nranks = nranks; ctx->Pr = Pr; ctx->Pc = Pc;
  const bool in_grid = rank < Pr * Pc;
  ctx->my_r = in_grid ? rank / Pc : -1;  // proc_grid.h: rank_row = rank / proc_cols
  ctx->my_c = in_grid ? rank % Pc : -1;
  TADEV_CHECK_NCCL(ncclCommSplit(ctx->world, in_grid ? ctx->my_r : NCCL_SPLIT_NOCOLOR, ctx->my_c, &ctx->row_comm, &cfg));
  TADEV_CHECK_NCCL(ncclCommSplit(ctx->world, in_grid ? ctx->my_c : NCCL_SPLIT_NOCOLOR, ctx->my_r, &ctx->col_comm, &cfg));
  // A second, "wide" pair of communicators for contractions whose panel traffic is large next to their GEMM work
  // (block-sparse config 3 at 8 GPUs receives 1.7 GB per rank against 20 ms of GEMM; 4 CTAs sustain 90-150 GB/s and
  // left the GEMM waiting: share of the step 0.62). Those contractions leave `sm_reserve_wide` SMs to NCCL instead.
  ctx->sm_reserve_wide = ctx->sm_reserve_thin > 0 ? 12 : 0;
  if (const char* e = getenv("TADEV_SM_RESERVE_WIDE")) ctx->sm_reserve_wide = atoi(e);
  if (ctx->sm_reserve_wide > ctx->sm_reserve_thin) {
    ncclConfig_t wcfg = cfg;
    wcfg.minCTAs = 1; wcfg.maxCTAs = ctx->sm_reserve_wide;
    TADEV_CHECK_NCCL(ncclCommSplit(ctx->world, in_grid ? ctx->my_r : NCCL_SPLIT_NOCOLOR, ctx->my_c, &ctx->row_comm_wide, &wcfg));
    TADEV_CHECK_NCCL(ncclCommSplit(ctx->world, in_grid ? ctx->my_c : NCCL_SPLIT_NOCOLOR, ctx->my_r, &ctx->col_comm_wide, &wcfg));
  }
  return TADEV_OK;
}

extern "C" int tadev_comm_destroy(tadev_ctx* ctx) {
  if (!ctx) return TADEV_OK;
  if (ctx->row_comm_wide) { ncclCommDestroy(ctx->row_comm_wide); ctx->row_comm_wide = nullptr; }
  if (ctx->col_comm_wide) { ncclCommDestroy(ctx->col_comm_wide); ctx->col_comm_wide = nullptr; }
  if (ctx->row_comm) { ncclCommDestroy(ctx->row_comm); ctx->row_comm = nullptr; }
  if (ctx->col_comm) { ncclCommDestroy(ctx->col_comm); ctx->col_comm = nullptr; }
  if (ctx->world) { ncclCommDestroy(ctx->world); ctx->world = nullptr; }
  ctx->rank = 0; ctx->nranks = 1; ctx->Pr = ctx->Pc = 1; ctx->my_r = ctx->my_c = 0;
  ctx->gemm_sm_reserve = ctx->sm_reserve_thin = ctx->sm_reserve_wide = 0;
  return TADEV_OK;
}

extern "C" int tadev_exchange_tiles(tadev_ctx* ctx, tadev_stream s, int nsend, const void* const* h_src, const size_t* h_sbytes,
                                    const int32_t* h_dst_rank, int nrecv, void* const* h_dst, const size_t* h_rbytes,
                                    const int32_t* h_src_rank) {
  TADEV_REQUIRE(ctx && nsend >= 0 && nrecv >= 0, "tadev_exchange_tiles: bad args");
  if (nsend == 0 && nrecv == 0) return TADEV_OK;
  TADEV_REQUIRE(ctx->world, "tadev_exchange_tiles: communicators not initialised");
  TADEV_REQUIRE((nsend == 0 || (h_src && h_sbytes && h_dst_rank)) && (nrecv == 0 || (h_dst && h_rbytes && h_src_rank)), "tadev_exchange_tiles: null arrays");
  TADEV_CHECK_NCCL(ncclGroupStart());
  for (int i = 0; i < nsend; ++i) {
    TADEV_REQUIRE(h_dst_rank[i] >= 0 && h_dst_rank[i] < ctx->nranks && h_dst_rank[i] != ctx->rank, "tadev_exchange_tiles: bad destination rank");
    if (h_sbytes[i]) TADEV_CHECK_NCCL(ncclSend(h_src[i], h_sbytes[i], ncclChar, h_dst_rank[i], ctx->world, (cudaStream_t)s));
  }
  for (int i = 0; i < nrecv; ++i) {
    TADEV_REQUIRE(h_src_rank[i] >= 0 && h_src_rank[i] < ctx->nranks && h_src_rank[i] != ctx->rank, "tadev_exchange_tiles: bad source rank");
    if (h_rbytes[i]) TADEV_CHECK_NCCL(ncclRecv(h_dst[i], h_rbytes[i], ncclChar, h_src_rank[i], ctx->world, (cudaStream_t)s));
  }
  TADEV_CHECK_NCCL(ncclGroupEnd());
  return TADEV_OK;
}

extern "C" int tadev_bcast_panel(tadev_ctx* ctx, tadev_stream s, int which, int root, void* d_buf, size_t bytes) {
  TADEV_REQUIRE(ctx && (which == 0 || which == 1), "tadev_bcast_panel: bad args");
  ncclComm* comm = which == 0 ? ctx->row_comm : ctx->col_comm;
  TADEV_REQUIRE(comm, "tadev_bcast_panel: communicators not initialised (or rank outside the grid)");
  if (bytes == 0) return TADEV_OK;
  TADEV_CHECK_NCCL(ncclBroadcast(d_buf, d_buf, bytes, ncclChar, root, comm, (cudaStream_t)s));
  return TADEV_OK;
}

// Shape replication (SparseShape ctor with a World: world.gop.max over the tile norms, sparse_shape.h:416):
// every rank contributes the norms of its own tiles (zeros elsewhere); afterwards all ranks hold all norms.
extern "C" int tadev_shape_allreduce_max_f32(tadev_ctx* ctx, tadev_stream s, float* d_norms, int64_t n) {
  TADEV_REQUIRE(ctx && n >= 0, "tadev_shape_allreduce_max_f32: bad args");
  if (n == 0 || ctx->nranks == 1) return TADEV_OK;
  TADEV_REQUIRE(ctx->world && d_norms, "tadev_shape_allreduce_max_f32: communicators not initialised");
  TADEV_CHECK_NCCL(ncclAllReduce(d_norms, d_norms, (size_t)n, ncclFloat, ncclMax, ctx->world, (cudaStream_t)s));
  return TADEV_OK;
}

// ---- built-in tile providers for lazy operands ---------------------------------------------------
extern "C" int tadev_provider_uniform(void* user, tadev_stream s, int ntiles, const uint64_t* tokens,
                                      double* const* d_dst, const size_t* elems) {
  const tadev_uniform_source* src = static_cast<const tadev_uniform_source*>(user);
  TADEV_REQUIRE(src && src->ctx && (ntiles == 0 || (tokens && d_dst && elems)), "tadev_provider_uniform: null");
  for (int n = 0; n < ntiles; ++n) {
    TADEV_REQUIRE(tokens[n] != 0, "tadev_provider_uniform: zero token");
    int rc = tadev_fill_uniform_f64(src->ctx, s, d_dst[n], elems[n], src->seed, (tokens[n] - 1) << 32);
    if (rc) return rc;
  }
  return TADEV_OK;
}

// Argument-tile permutation performed when the SUMMA window needs the tile (ArrayEvalImpl /
// LazyArrayTile with a permuting op, dist_eval/array_eval.h:42,170): tiles with identical extents
// share one batched launch.
extern "C" int tadev_provider_permute(void* user, tadev_stream s, int ntiles, const uint64_t* tokens,
                                      double* const* d_dst, const size_t* elems) {
  const tadev_permute_source* src = static_cast<const tadev_permute_source*>(user);
  TADEV_REQUIRE(src && src->ctx && src->extents && src->rank >= 0 && src->rank <= 16, "tadev_provider_permute: bad source");
  TADEV_REQUIRE(src->src_memory == TADEV_MEM_LAZY ? src->ordinals != nullptr : src->src != nullptr, "tadev_provider_permute: bad source tables");
  TADEV_REQUIRE(ntiles == 0 || (tokens && d_dst && elems), "tadev_provider_permute: null");
  const int R = src->rank;
  // host-resident / lazy sources are first brought into a stream-ordered scratch buffer (one per call)
  double* scratch = nullptr;
  std::vector<size_t> soff((size_t)ntiles, 0);
  if (src->src_memory != TADEV_MEM_DEVICE && ntiles > 0) {
    size_t tot = 0;
    for (int n = 0; n < ntiles; ++n) { soff[n] = tot; tot += (elems[n] + 1) & ~size_t(1); }
    int rc = tadev_alloc(src->ctx, std::max<size_t>(tot, 2) * 8, (void**)&scratch, s);
    if (rc) return rc;
    for (int n = 0; n < ntiles && !rc; ++n) {
      if (tokens[n] == 0) { rc = TADEV_EINVAL; tadev_set_error("tadev_provider_permute: zero token"); break; }
      if (src->src_memory == TADEV_MEM_HOST) rc = tadev_memcpy_h2d(src->ctx, scratch + soff[n], src->src[tokens[n] - 1], elems[n] * 8, s);
      else rc = tadev_fill_uniform_f64(src->ctx, s, scratch + soff[n], elems[n], src->lazy_seed, (uint64_t)src->ordinals[tokens[n] - 1] << 32);
    }
    if (rc) { tadev_free(src->ctx, scratch, s); return rc; }
  }
  std::vector<char> done((size_t)ntiles, 0);
  std::vector<const void*> ins;
  std::vector<void*> outs;
  int rc = TADEV_OK;
  for (int n = 0; n < ntiles && !rc; ++n) {
    if (done[n]) continue;
    if (tokens[n] == 0) { tadev_set_error("tadev_provider_permute: zero token"); rc = TADEV_EINVAL; break; }
    const int64_t* ext = src->extents + (size_t)(tokens[n] - 1) * R;
    ins.clear(); outs.clear();
    for (int q = n; q < ntiles; ++q) {
      if (done[q] || tokens[q] == 0) continue;
      const int64_t* eq = src->extents + (size_t)(tokens[q] - 1) * R;
      if (q != n && memcmp(eq, ext, sizeof(int64_t) * R) != 0) continue;
      size_t vol = 1;
      for (int d = 0; d < R; ++d) vol *= (size_t)eq[d];
      if (vol != elems[q]) { tadev_set_error("tadev_provider_permute: tile %d has %zu elements, the driver expects %zu", q, vol, elems[q]); rc = TADEV_EINVAL; break; }
      ins.push_back(scratch ? (const void*)(scratch + soff[q]) : src->src[tokens[q] - 1]);
      outs.push_back(d_dst[q]);
      done[q] = 1;
    }
    if (!rc) rc = tadev_permute_batched(src->ctx, s, R, ext, src->perm, 8, (int)ins.size(), ins.data(), outs.data());
  }
  if (scratch) tadev_free(src->ctx, scratch, s);
  return rc;
}

namespace {

inline size_t pad2(size_t n) { return (n + 1) & ~size_t(1); }  // keep every tile 16-byte aligned

struct Bcast { void* ptr; size_t bytes; int root; };

// A SUMMA step restricted to one row block of the result.
struct BlockStep {
  const SummaStep* st;
  std::vector<int> a_rows;  // st->a_rows ∩ block rows
  bool bcast_a, bcast_b, compute;
  size_t a_elems, b_elems;  // padded panel sizes (doubles)
};

struct Window {
  std::vector<int> steps;  // indices into the block's BlockStep list
  size_t bytes = 0;        // ring bytes needed
};

}  // namespace

namespace {

// Everything the driver derives from the plan before touching the device: the step schedule of this
// grid position, the row blocks, the B cache decision, and the (block, window) structure. Host-only,
// shared by tadev_summa_f64 and tadev_summa_comm_trace.
struct SummaWindows {
  SummaSchedule S;
  std::vector<int> my_rows;
  int nb = 1, W = 1, D = 2;
  bool b_cache = false;
  std::vector<std::vector<BlockStep>> bsteps;
  std::vector<std::vector<Window>> bwins;
  std::vector<std::pair<int, int>> brange;  // [first, last) positions in my_rows
  size_t max_bytes = 0, b_cache_elems = 0;
  std::vector<size_t> b_cache_off;
};

void build_summa_windows(int Pr, int Pc, int r, int c, const tadev_summa_plan& P, SummaWindows& X) {
  const int Mt = P.Mt, Nt = P.Nt, Kt = P.Kt;
  const bool multi = (Pr * Pc > 1);
  const bool a_host = (P.flags & TADEV_SUMMA_A_ON_HOST) != 0, b_host = (P.flags & TADEV_SUMMA_B_ON_HOST) != 0,
             c_host = (P.flags & TADEV_SUMMA_C_ON_HOST) != 0;
  const bool a_lazy = (P.flags & TADEV_SUMMA_A_LAZY) != 0, b_lazy = (P.flags & TADEV_SUMMA_B_LAZY) != 0;
  const bool a_stg = a_host || a_lazy, b_stg = b_host || b_lazy;
  SummaSchedule& S = X.S;
  std::vector<int>& my_rows = X.my_rows;
  int& nb = X.nb;
  int& W = X.W;
  bool& b_cache = X.b_cache;
  auto& bsteps = X.bsteps;
  auto& bwins = X.bwins;
  auto& brange = X.brange;
  size_t& max_bytes = X.max_bytes;
  size_t& b_cache_elems = X.b_cache_elems;
  auto& b_cache_off = X.b_cache_off;
  S = make_summa_schedule(Pr, Pc, r, c, Mt, Nt, Kt, P.a_norms, P.b_norms, P.c_norms, P.threshold);

  // ---- row blocks (deterministic from global quantities: every rank must agree on the count
  //      because B panels are re-broadcast per block unless they are cached)
  for (int i = r; i < Mt; i += Pr) my_rows.push_back(i);
  nb = 1;
  if (!c_host && (P.row_blocks > 0 || a_lazy)) {
    // device-resident result in row blocks: bounds the A panel of one step when A is generated on
    // the fly (each block needs only its own rows of the panel). Same count on every rank.
    const int max_rows = (Mt + Pr - 1) / Pr;
    nb = P.row_blocks;
    if (nb <= 0) {
      double panel = 0, kmax = 0;
      for (int k = 0; k < Kt; ++k) kmax = std::max(kmax, (double)P.k_ext[k]);
      for (int rr = 0; rr < Pr; ++rr) { double t = 0; for (int i = rr; i < Mt; i += Pr) t += (double)P.m_ext[i]; panel = std::max(panel, t * kmax * 8.0); }
      nb = (int)std::ceil(panel / (2.0 * 1073741824.0));
    }
    nb = std::max(1, std::min(nb, max_rows));
  }
  if (c_host) {
    const int max_rows = (Mt + Pr - 1) / Pr;
    double m_sum = 0, n_sum = 0;
    for (int i = 0; i < Mt; ++i) m_sum += (double)P.m_ext[i];
    for (int j = 0; j < Nt; ++j) n_sum += (double)P.n_ext[j];
    const double c_bytes = (m_sum / Pr) * (n_sum / Pc) * 8.0;
    nb = P.row_blocks > 0 ? P.row_blocks : (int)std::ceil(c_bytes / (2.0 * 1073741824.0));
    nb = std::max(nb, std::min(4, max_rows));
    nb = std::max(1, std::min(nb, max_rows));
  }
  // B staged through the device (received by broadcast or uploaded from the host) is kept for all
  // row blocks when it fits the cache budget; the dense upper bound is the same on every rank.
  const bool b_staged = b_stg || (multi && Pr > 1);
  b_cache = false;
  if (nb > 1 && b_staged) {
    double k_sum = 0, n_max = 0;
    for (int k = 0; k < Kt; ++k) k_sum += (double)P.k_ext[k];
    for (int cc = 0; cc < Pc; ++cc) { double t = 0; for (int j = cc; j < Nt; j += Pc) t += (double)P.n_ext[j]; n_max = std::max(n_max, t); }
    double limit = 48.0 * 1073741824.0;
    if (const char* e = getenv("TADEV_B_CACHE_GB")) limit = atof(e) * 1073741824.0;
    b_cache = k_sum * n_max * 8.0 <= limit;
  }

  // Pipeline policy. The reference keeps `depth` SUMMA iterations in flight (contraction_eval.h:1925-1977: at least 2,
  // boosted by the sparsity of the arguments, bounded by TA_SUMMA_MAX_DEPTH and by TA_SUMMA_MAX_MEMORY through
  // mem_bound_depth, :225-263). Here D ring slots of W steps each are in flight (one grouped GEMM launch per window):
  // W = ceil(4096 contracted elements / average k extent), multiplied by the reference's sparse boost
  // 1 - 1.35638 log2((1 - min(sA, 0.9)) (1 - min(sB, 0.9))); D * W <= TA_SUMMA_MAX_DEPTH; D * window bytes <=
  // TA_SUMMA_MAX_MEMORY (same syntax as the reference: "<number> [kB|KiB|MB|MiB|GB|GiB]", at least 100 MiB).
  // ring depth: 2 windows in flight (broadcast of n+1 under the GEMM of n); staged operands (host-resident / lazy)
  // that are also broadcast have three stages per window - upload or generate, broadcast, GEMM - and need 3 slots,
  // else every other window waits for a slot (8 GPUs, host operands: 12.7 ms per 8.5 ms window with depth 2)
  X.D = P.depth > 0 ? std::max(2, P.depth) : ((multi && (a_stg || b_stg)) ? 3 : 2);
  size_t kMaxWindowBytes = size_t(5) << 30;
  if (const char* e = getenv("TA_SUMMA_MAX_MEMORY")) {
    char unit[16] = "";
    double mem = 0.0;
    if (sscanf(e, "%lf %15s", &mem, unit) >= 1 && mem > 0.0) {
      const std::string u(unit);
      if (u == "KB" || u == "kB") mem *= 1e3; else if (u == "KiB" || u == "kiB") mem *= 1024.0;
      else if (u == "MB") mem *= 1e6; else if (u == "MiB") mem *= 1048576.0;
      else if (u == "GB") mem *= 1e9; else if (u == "GiB") mem *= 1073741824.0;
      mem = std::max(mem, 104857600.0);
      kMaxWindowBytes = std::min<size_t>(kMaxWindowBytes, (size_t)(mem / X.D));
    }
  }
  W = P.steps_per_launch;
  if (W <= 0) {
    if (!multi && !a_stg && !b_stg) W = std::max(1, Kt);
    else {
      double avgk = 0;
      for (int k = 0; k < Kt; ++k) avgk += (double)P.k_ext[k];
      avgk = Kt ? avgk / Kt : 1.0;
      double w = std::max(1.0, std::ceil(4096.0 / std::max(1.0, avgk)));
      if (P.a_norms && P.b_norms && Mt > 0 && Nt > 0 && Kt > 0) {
        size_t za = 0, zb = 0;
        for (size_t x = 0; x < (size_t)Mt * Kt; ++x) za += P.a_norms[x] < P.threshold;
        for (size_t x = 0; x < (size_t)Kt * Nt; ++x) zb += P.b_norms[x] < P.threshold;
        const float sa = (float)za / ((float)Mt * Kt), sb = (float)zb / ((float)Kt * Nt);
        const float frac = (1.0f - std::min(sa, 0.9f)) * (1.0f - std::min(sb, 0.9f));
        w = w * (1.0 - 1.35638 * std::log2((double)frac)) + 0.5;
      }
      W = (int)std::min<double>(64.0, std::max(1.0, std::floor(w)));
      if (const char* e = getenv("TADEV_SUMMA_W")) if (atoi(e) > 0) W = atoi(e);
    }
    if (const char* e = getenv("TA_SUMMA_MAX_DEPTH")) if (atoi(e) > 0) W = std::max(1, std::min(W, atoi(e) / X.D));
  }
  // ---- per-block step views and windows
  auto tile_a_elems = [&](int i, int k) { return (size_t)P.m_ext[i] * (size_t)P.k_ext[k]; };
  auto tile_b_elems = [&](int k, int j) { return (size_t)P.k_ext[k] * (size_t)P.n_ext[j]; };
  bsteps.assign(nb, {});
  bwins.assign(nb, {});
  brange.assign(nb, {0, 0});
  max_bytes = 0; b_cache_elems = 0;
  b_cache_off.assign(S.steps.size(), 0);
  // ---- global windows. NCCL may reorder the operations inside one group, so two ranks of a
  // communicator must put the same broadcasts into the same group: window boundaries are therefore
  // derived from replicated data only (globally active steps, a byte bound that is the maximum
  // over all grid positions), never from this rank's own compute pattern.
  // Two maps: block 0 ramps its windows up (pipeline fill); later row blocks of a host-resident / lazy-operand
  // contraction start with their first panels already prefetched during the previous block, so they use full-size
  // windows from the start (the ramp cost ~15 % per block: profiles/r02_trace_C2_n8_e2e_depth2.log).
  double a_block_frac = 1.0;  // largest share of this grid row's tile rows that one row block holds
  if (nb > 1 && !my_rows.empty()) {
    const int L = (int)my_rows.size();
    const int64_t wtot = c_host && nb >= 2 ? 2 * (int64_t)nb - 1 : nb, wsc = c_host && nb >= 2 ? 2 : 1;
    int most = 1;
    for (int b = 0; b < nb; ++b) {
      const int lo = (int)((int64_t)L * std::min<int64_t>(wsc * b, wtot) / wtot), hi = (int)((int64_t)L * std::min<int64_t>(wsc * (b + 1), wtot) / wtot);
      most = std::max(most, hi - lo);
    }
    // every grid row cuts its rows the same way; one extra row of slack for rounding between grid rows
    a_block_frac = std::min(1.0, (double)(most + 1) / (double)std::max(1, (Mt + Pr - 1) / Pr));
  }
  std::vector<int> win_of_k(std::max(Kt, 1), 0), win_of_k_steady(std::max(Kt, 1), 0);
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<int>& wmap = pass == 0 ? win_of_k : win_of_k_steady;
    auto a_nz = [&](int i, int k) { return !P.a_norms || P.a_norms[(size_t)i * Kt + k] >= P.threshold; };
    auto b_nz = [&](int k, int j) { return !P.b_norms || P.b_norms[(size_t)k * Nt + j] >= P.threshold; };
    int nwin = 0, cnt = 0;
    size_t gbytes = 0;
    std::vector<size_t> per_r(Pr), per_c(Pc);
    for (int k = 0; k < Kt; ++k) {
      std::fill(per_r.begin(), per_r.end(), 0);
      std::fill(per_c.begin(), per_c.end(), 0);
      bool any_a = false, any_b = false;
      for (int i = 0; i < Mt; ++i) if (a_nz(i, k)) { any_a = true; per_r[i % Pr] += pad2((size_t)P.m_ext[i] * P.k_ext[k]) * 8; }
      for (int j = 0; j < Nt; ++j) if (b_nz(k, j)) { any_b = true; per_c[j % Pc] += pad2((size_t)P.k_ext[k] * P.n_ext[j]) * 8; }
      wmap[k] = nwin;
      if (!(any_a && any_b)) continue;  // no rank computes or broadcasts in this step
      size_t need = 0;
      // (with row blocks the ring holds only one block's rows of an A panel: BASELINE config 4 has 22.7 GB A panels but
      // 2.3 GB per row block - without this factor its windows degenerated to single steps, K = 4096 per work item)
      if (a_stg || Pc > 1) need += (size_t)((double)*std::max_element(per_r.begin(), per_r.end()) * a_block_frac);
      if ((b_stg || Pr > 1) && !b_cache) need += *std::max_element(per_c.begin(), per_c.end());
      // windows ramp up geometrically (1, 2, 4, ... W steps): the panels of window n+1 travel while window n computes,
      // so only the first, single-step window is exposed; a full-size second window had its whole broadcast in the
      // open (block-sparse config 3 on 4 GPUs: 11 of 55 ms, profiles/r02_trace_C3_n4_before_ramp.log)
      const int wcap = (pass == 0 && nwin < 30) ? std::min(W, 1 << nwin) : W;
      if (cnt > 0 && (cnt >= wcap || gbytes + need > kMaxWindowBytes)) { ++nwin; cnt = 0; gbytes = 0; wmap[k] = nwin; }
      ++cnt; gbytes += need;
    }
  }
  // B cache layout: the panels one root row sends in one window are contiguous (they travel as ONE broadcast):
  // order (global window, root row k % Pr, step)
  if (b_cache) {
    int nwin_total = 0;
    for (int k = 0; k < Kt; ++k) nwin_total = std::max(nwin_total, win_of_k[k] + 1);
    std::vector<std::vector<size_t>> by_slot((size_t)std::max(nwin_total, 1) * Pr);
    for (size_t si = 0; si < S.steps.size(); ++si)
      if (S.steps[si].bcast_b || b_stg) by_slot[(size_t)win_of_k[S.steps[si].k] * Pr + S.steps[si].k % Pr].push_back(si);
    for (auto& v : by_slot)
      for (size_t si : v) {
        size_t e = 0;
        for (int j : S.steps[si].b_cols) e += pad2(tile_b_elems(S.steps[si].k, j));
        b_cache_off[si] = b_cache_elems;
        b_cache_elems += e;
      }
  }
  for (int b = 0; b < nb; ++b) {
    const int L = (int)my_rows.size();
    // host-resident result: the download of the LAST block is not hidden by any GEMM, so that block gets
    // half the rows of the others (weights 2,...,2,1)
    const int64_t wtot = c_host && nb >= 2 ? 2 * (int64_t)nb - 1 : nb, wsc = c_host && nb >= 2 ? 2 : 1;
    const int lo = (int)((int64_t)L * std::min<int64_t>(wsc * b, wtot) / wtot), hi = (int)((int64_t)L * std::min<int64_t>(wsc * (b + 1), wtot) / wtot);
    brange[b] = {lo, hi};
    const int row_lo = lo < L ? my_rows[lo] : Mt, row_hi = hi < L ? my_rows[hi] : Mt;  // global row bounds
    for (size_t si = 0; si < S.steps.size(); ++si) {
      const SummaStep& st = S.steps[si];
      BlockStep bs;
      bs.st = &st;
      for (int i : st.a_rows) if (i >= row_lo && i < row_hi) bs.a_rows.push_back(i);
      bs.a_elems = 0;
      for (int i : bs.a_rows) bs.a_elems += pad2(tile_a_elems(i, st.k));
      bs.b_elems = 0;
      for (int j : st.b_cols) bs.b_elems += pad2(tile_b_elems(st.k, j));
      // a panel travels iff the (block-restricted) panel is non-empty and some rank can use it;
      // every member of the row group shares the same rows, so the decision is group-consistent
      bs.bcast_a = st.bcast_a && !bs.a_rows.empty();
      bs.bcast_b = st.bcast_b && !(b_cache && b > 0);
      bs.compute = !bs.a_rows.empty() && !st.b_cols.empty();
      if (bs.compute || bs.bcast_a || bs.bcast_b) bsteps[b].push_back(std::move(bs));
    }
    // windows are GLOBAL k-ranges (win_of_k): every rank of a row/column communicator then issues
    // identically composed NCCL groups (same broadcasts in the same group) whatever its own sparsity
    Window cur;
    int cur_win = -1;
    for (int x = 0; x < (int)bsteps[b].size(); ++x) {
      const BlockStep& bs = bsteps[b][x];
      size_t need = 0;
      if (bs.bcast_a || (a_stg && bs.compute)) need += bs.a_elems * 8;
      const bool b_to_cache = b_cache && (bs.st->bcast_b || b_stg);
      if (!b_to_cache && (bs.bcast_b || (b_stg && bs.compute))) need += bs.b_elems * 8;
      const int gw = (b == 0 ? win_of_k : win_of_k_steady)[bs.st->k];
      if (!cur.steps.empty() && gw != cur_win) { bwins[b].push_back(std::move(cur)); cur = Window(); }
      cur_win = gw;
      cur.steps.push_back(x); cur.bytes += need;
    }
    if (!cur.steps.empty()) bwins[b].push_back(std::move(cur));
    for (auto& w : bwins[b]) max_bytes = std::max(max_bytes, w.bytes);
  }
  if (!multi && !a_stg && !b_stg) {  // P == 1, device-resident: undo the short first window
    // (no staging at all, so one launch for everything is best)
    for (int b = 0; b < nb; ++b) {
      Window all;
      for (int x = 0; x < (int)bsteps[b].size(); ++x) all.steps.push_back(x);
      bwins[b].clear();
      if (!all.steps.empty()) {
        if (P.steps_per_launch > 0) {
          Window cur; int ncomp = 0;
          for (int x : all.steps) {
            if (!cur.steps.empty() && ncomp >= P.steps_per_launch) { bwins[b].push_back(std::move(cur)); cur = Window(); ncomp = 0; }
            cur.steps.push_back(x);
            if (bsteps[b][x].compute) ++ncomp;
          }
          if (!cur.steps.empty()) bwins[b].push_back(std::move(cur));
        } else bwins[b].push_back(std::move(all));
      }
    }
    max_bytes = 0;
  }

}

}  // namespace

// [host] the panel broadcasts the driver issues for this plan at grid position (r, c), in issue order
extern "C" int tadev_summa_comm_trace(int Pr, int Pc, int r, int c, const tadev_summa_plan* plan, int32_t* comm,
                                      int32_t* group, int32_t* k, int32_t* root, int64_t* bytes, int64_t capacity,
                                      int64_t* n_out) {
  TADEV_REQUIRE(plan && n_out, "tadev_summa_comm_trace: null");
  TADEV_REQUIRE(Pr >= 1 && Pc >= 1 && r >= 0 && r < Pr && c >= 0 && c < Pc, "tadev_summa_comm_trace: bad grid position");
  TADEV_REQUIRE(plan->m_ext && plan->n_ext && (plan->Kt == 0 || plan->k_ext), "tadev_summa_comm_trace: null extent arrays");
  SummaWindows X;
  build_summa_windows(Pr, Pc, r, c, *plan, X);
  int64_t n = 0;
  int32_t g = 0;
  auto emit = [&](int cm, int kk, int rt, size_t by) {
    if (n < capacity && comm && group && k && root && bytes) { comm[n] = cm; group[n] = g; k[n] = kk; root[n] = rt; bytes[n] = (int64_t)by; }
    ++n;
  };
  // one broadcast per (window, communicator, root): the root's panels of all steps of the window travel together
  for (int b = 0; b < X.nb; ++b)
    for (const Window& win : X.bwins[b]) {
      for (int which = 0; which < 2; ++which) {
        const int nroots = which == 0 ? Pc : Pr;
        bool any = false;
        for (int root = 0; root < nroots; ++root) {
          size_t by = 0;
          int first_k = -1;
          for (int x : win.steps) {
            const BlockStep& bs = X.bsteps[b][x];
            if (!(which == 0 ? bs.bcast_a : bs.bcast_b) || bs.st->k % nroots != root) continue;
            by += (which == 0 ? bs.a_elems : bs.b_elems) * 8;
            if (first_k < 0) first_k = bs.st->k;
          }
          if (by) { emit(which, first_k, root, by); any = true; }
        }
        if (any) ++g;
      }
    }
  *n_out = n;
  TADEV_REQUIRE(n <= capacity || !comm, "tadev_summa_comm_trace: capacity %lld < %lld", (long long)capacity, (long long)n);
  return TADEV_OK;
}

// -----------------------------------------------------------------------------------------------
// The driver. Loops: row blocks of the result (1 unless the result lives in host memory) ->
// windows of K steps -> [H2D staging of host-resident panels | NCCL panel broadcasts] overlapped
// with ONE grouped GEMM launch per window; finished result blocks are copied to the host while
// the next block computes. See the file header for the reference mapping.
extern "C" int tadev_summa_f64(tadev_ctx* ctx, const tadev_summa_plan* plan, tadev_summa_stats* stats) {
  TADEV_REQUIRE(ctx && plan, "tadev_summa_f64: null");
  const tadev_summa_plan& P = *plan;
  TADEV_REQUIRE(P.Mt >= 0 && P.Nt >= 0 && P.Kt >= 0, "tadev_summa_f64: negative tile-grid extents");
  TADEV_REQUIRE((P.opA == 0 || P.opA == 1) && (P.opB == 0 || P.opB == 1), "tadev_summa_f64: bad op flags");
  TADEV_REQUIRE(P.m_ext && P.n_ext && (P.Kt == 0 || P.k_ext), "tadev_summa_f64: null extent arrays");
  TADEV_REQUIRE(P.a_tiles && P.b_tiles && P.c_tiles, "tadev_summa_f64: null tile tables");
  if (stats) memset(stats, 0, sizeof(*stats));
  const int Pr = ctx->Pr, Pc = ctx->Pc, r = ctx->my_r, c = ctx->my_c;
  if (r < 0 || c < 0) return TADEV_OK;  // outside the process grid: nothing to do
  const bool multi = (Pr * Pc > 1);
  TADEV_REQUIRE(!multi || (ctx->row_comm && ctx->col_comm), "tadev_summa_f64: communicators not initialised");
  const int Mt = P.Mt, Nt = P.Nt, Kt = P.Kt;
  const bool a_host = (P.flags & TADEV_SUMMA_A_ON_HOST) != 0, b_host = (P.flags & TADEV_SUMMA_B_ON_HOST) != 0,
             c_host = (P.flags & TADEV_SUMMA_C_ON_HOST) != 0;
  const bool a_lazy = (P.flags & TADEV_SUMMA_A_LAZY) != 0, b_lazy = (P.flags & TADEV_SUMMA_B_LAZY) != 0;
  TADEV_REQUIRE(!(a_lazy && a_host) && !(b_lazy && b_host), "tadev_summa_f64: an operand cannot be both lazy and host-resident");
  TADEV_REQUIRE((!a_lazy || P.a_provider) && (!b_lazy || P.b_provider), "tadev_summa_f64: lazy operand without a tile provider");
  // "staged" operands have no device-resident tiles: every panel is materialised in the ring (or the
  // B cache) by an upload or by the provider, on the staging stream
  const bool b_stg = b_host || b_lazy;
  TADEV_CHECK_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s0 = ctx->streams[0];                       // compute
  cudaStream_t sd = ctx->streams[ctx->streams.size() > 1 ? 1 : 0];  // result download
  cudaStream_t sc = ctx->comm_stream[0];                   // NCCL panel broadcasts
  cudaStream_t sh = ctx->comm_stream[1];                   // host -> device panel staging

  const auto host_t0 = std::chrono::steady_clock::now();
  auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
  // ---- which communicators: contractions whose panel traffic is large next to their GEMM work (block-sparse ones)
  //      use the wide pair and leave more SMs to NCCL; the decision uses replicated data only (every rank agrees).
  //      Estimate: bytes the busiest rank receives at ~100 GB/s (4 CTAs) vs its GEMM work at ~35 TFLOP/s.
  bool wide = false;
  if (multi && ctx->row_comm_wide && ctx->col_comm_wide) {
    double a_bytes = 0, b_bytes = 0, gflop = 0;
    for (int k = 0; k < Kt; ++k) {
      double na = 0, nbk = 0;
      for (int i = 0; i < Mt; ++i) if (!P.a_norms || P.a_norms[(size_t)i * Kt + k] >= P.threshold) na += (double)P.m_ext[i];
      for (int j = 0; j < Nt; ++j) if (!P.b_norms || P.b_norms[(size_t)k * Nt + j] >= P.threshold) nbk += (double)P.n_ext[j];
      a_bytes += na * (double)P.k_ext[k] * 8.0; b_bytes += nbk * (double)P.k_ext[k] * 8.0;
      gflop += 2.0 * na * nbk * (double)P.k_ext[k];
    }
    const double recv = (a_bytes / Pr) * (Pc - 1) / Pc + (b_bytes / Pc) * (Pr - 1) / Pr;
    const double t_comm = recv / 100e9, t_gemm = gflop / ((double)Pr * Pc * 35e12);
    wide = t_comm > 0.5 * t_gemm;
    if (const char* e = getenv("TADEV_WIDE_COMM")) wide = atoi(e) != 0;
  }
  ncclComm* const row_comm = wide ? ctx->row_comm_wide : ctx->row_comm;
  ncclComm* const col_comm = wide ? ctx->col_comm_wide : ctx->col_comm;
  struct ReserveGuard {  // the GEMM launchers of this contraction leave the matching number of SMs free
    tadev_ctx* ctx; int saved;
    ~ReserveGuard() { ctx->gemm_sm_reserve = saved; }
  } reserve_guard{ctx, ctx->gemm_sm_reserve};
  if (multi) ctx->gemm_sm_reserve = wide ? ctx->sm_reserve_wide : ctx->sm_reserve_thin;

  SummaWindows X;
  build_summa_windows(Pr, Pc, r, c, P, X);
  SummaSchedule& S = X.S;
  const std::vector<int>& my_rows = X.my_rows;
  const int nb = X.nb, D = X.D;
  const bool b_cache = X.b_cache;
  auto& bsteps = X.bsteps;
  auto& bwins = X.bwins;
  auto& brange = X.brange;
  const size_t max_bytes = X.max_bytes, b_cache_elems = X.b_cache_elems;
  auto& b_cache_off = X.b_cache_off;
  auto tile_a_elems = [&](int i, int k) { return (size_t)P.m_ext[i] * (size_t)P.k_ext[k]; };
  auto tile_b_elems = [&](int k, int j) { return (size_t)P.k_ext[k] * (size_t)P.n_ext[j]; };

  // ---- up-front validation: every tile this rank must supply exists (nothing may fail for a local reason once
  //      the first collective has been enqueued: the peers would wait for this rank forever)
  int local_rc = TADEV_OK;
  for (int b = 0; b < nb && !local_rc; ++b)
    for (const BlockStep& bs : bsteps[b]) {
      const int k = bs.st->k;
      if ((bs.bcast_a || bs.compute) && c == k % Pc)
        for (int i : bs.a_rows)
          if (!P.a_tiles[(size_t)i * Kt + k]) { tadev_set_error("tadev_summa_f64: A tile (%d,%d) is owned by this rank but has no data", i, k); local_rc = TADEV_EINVAL; }
      if ((bs.bcast_b || bs.compute) && r == k % Pr)
        for (int j : bs.st->b_cols)
          if (!P.b_tiles[(size_t)k * Nt + j]) { tadev_set_error("tadev_summa_f64: B tile (%d,%d) is owned by this rank but has no data", k, j); local_rc = TADEV_EINVAL; }
      if (bs.compute)
        for (int64_t pp = bs.st->pair_begin; pp < bs.st->pair_end; ++pp)
          if (!P.c_tiles[(size_t)S.pair_i[pp] * Nt + S.pair_j[pp]]) {
            tadev_set_error("tadev_summa_f64: result tile (%d,%d) is non-zero and local but has no storage", S.pair_i[pp], S.pair_j[pp]);
            local_rc = TADEV_EINVAL;
          }
      if (local_rc) break;
    }

  // ---- resources: events and device buffers are owned by `R`; its destructor releases them on EVERY exit path
  //      (on an error path it first drains the four streams so that nothing still reads a buffer being freed)
  struct Resources {
    tadev_ctx* ctx; cudaStream_t s0, sc, sh, sd;
    std::vector<cudaEvent_t> events;
    std::vector<void*> buffers;
    bool ok = false;
    int event(cudaEvent_t* e, unsigned flags) {
      TADEV_CHECK_CUDA(cudaEventCreateWithFlags(e, flags));
      events.push_back(*e);
      return TADEV_OK;
    }
    int alloc(size_t bytes, double** p) {
      int rc = tadev_alloc(ctx, bytes, (void**)p, (tadev_stream)s0);
      if (!rc && *p) buffers.push_back(*p);
      return rc;
    }
    ~Resources() {
      if (!ok) { for (cudaStream_t st : {s0, sc, sh, sd}) cudaStreamSynchronize(st); cudaGetLastError(); }
      for (void* b : buffers) tadev_free(ctx, b, (tadev_stream)s0);
      for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
  } R{ctx, s0, sc, sh, sd};
  cudaEvent_t ev_start, ev_end, ev_aux;
  int rc0 = R.event(&ev_start, cudaEventDefault);
  if (!rc0) rc0 = R.event(&ev_end, cudaEventDefault);
  if (!rc0) rc0 = R.event(&ev_aux, cudaEventDisableTiming);
  if (rc0) return rc0;
  const bool need_ring = max_bytes > 0;
  std::vector<double*> ring(D, nullptr);
  std::vector<cudaEvent_t> panel_ready(D), buf_free(D), h2d_done(D);
  std::vector<char> buf_used(D, 0);
  for (int d = 0; d < D && !local_rc; ++d) {
    if (need_ring) local_rc = R.alloc(max_bytes, &ring[d]);
    if (!local_rc) local_rc = R.event(&panel_ready[d], cudaEventDisableTiming);
    if (!local_rc) local_rc = R.event(&buf_free[d], cudaEventDisableTiming);
    if (!local_rc) local_rc = R.event(&h2d_done[d], cudaEventDisableTiming);
  }
  double* bcache = nullptr;
  if (!local_rc && b_cache && b_cache_elems) local_rc = R.alloc(b_cache_elems * 8, &bcache);
  // result blocks staged on the device when the result lives in host memory (double buffered)
  std::vector<size_t> cblock_elems(nb, 0);
  auto c_local = [&](int i, int j) {
    return (j % Pc == c) && (!P.c_norms || P.c_norms[(size_t)i * Nt + j] >= P.threshold) && P.c_tiles[(size_t)i * Nt + j] != nullptr;
  };
  size_t cmax = 0;
  if (c_host) {
    for (int b = 0; b < nb; ++b) {
      for (int x = brange[b].first; x < brange[b].second; ++x)
        for (int j = c; j < Nt; j += Pc)
          if (c_local(my_rows[x], j)) cblock_elems[b] += pad2((size_t)P.m_ext[my_rows[x]] * P.n_ext[j]);
      cmax = std::max(cmax, cblock_elems[b]);
    }
  }
  double* carena[2] = {nullptr, nullptr};
  cudaEvent_t c_done[2], d2h_done[2];
  bool c_used[2] = {false, false};
  for (int x = 0; x < 2 && !local_rc; ++x) {
    if (c_host && cmax) local_rc = R.alloc(cmax * 8, &carena[x]);
    if (!local_rc) local_rc = R.event(&c_done[x], cudaEventDisableTiming);
    if (!local_rc) local_rc = R.event(&d2h_done[x], cudaEventDisableTiming);
  }
  // ---- agree on success across the grid before the first panel broadcast: a rank that failed validation or an
  //      allocation returns its error and every other rank returns TADEV_ENCCL instead of blocking in NCCL
  if (multi) {
    int32_t* flag = nullptr;
    if (cudaMallocHost((void**)&flag, sizeof(int32_t)) != cudaSuccess) { cudaGetLastError(); flag = nullptr; }
    double* dflag = nullptr;
    int arc = flag ? R.alloc(16, &dflag) : TADEV_ENOMEM;
    if (!arc) {
      *flag = local_rc;
      TADEV_CHECK_CUDA(cudaMemcpyAsync(dflag, flag, sizeof(int32_t), cudaMemcpyHostToDevice, sc));
      // every rank of the grid is in one row and one column communicator: max over both = max over the grid
      ncclResult_t n1 = ncclAllReduce(dflag, dflag, 1, ncclInt32, ncclMax, row_comm, sc);
      ncclResult_t n2 = n1 == ncclSuccess ? ncclAllReduce(dflag, dflag, 1, ncclInt32, ncclMax, col_comm, sc) : n1;
      if (n2 != ncclSuccess) { cudaFreeHost(flag); tadev_set_error("tadev_summa_f64: status all-reduce failed: %s", ncclGetErrorString(n2)); return TADEV_ENCCL; }
      TADEV_CHECK_CUDA(cudaMemcpyAsync(flag, dflag, sizeof(int32_t), cudaMemcpyDeviceToHost, sc));
      TADEV_CHECK_CUDA(cudaStreamSynchronize(sc));
      const int32_t grid_rc = *flag;
      cudaFreeHost(flag);
      if (local_rc) return local_rc;
      if (grid_rc) { tadev_set_error("tadev_summa_f64: another rank of the process grid failed before the first broadcast (status %d)", grid_rc); return TADEV_ENCCL; }
    } else {
      if (flag) cudaFreeHost(flag);
      return local_rc ? local_rc : arc;
    }
  } else if (local_rc) return local_rc;
  static const bool trace_host = getenv("TADEV_SUMMA_TRACE") && atoi(getenv("TADEV_SUMMA_TRACE"));
  if (trace_host) fprintf(stderr, "[tadev summa rank %d] host: schedule+windows+resources ready at %.3f ms (%s communicators, %d SMs left to NCCL)\n",
                          ctx->rank, host_ms(), wide ? "wide" : "thin", ctx->gemm_sm_reserve);
  TADEV_CHECK_CUDA(cudaEventRecord(ev_start, s0));
  TADEV_CHECK_CUDA(cudaStreamWaitEvent(sc, ev_start, 0));
  TADEV_CHECK_CUDA(cudaStreamWaitEvent(sh, ev_start, 0));
  TADEV_CHECK_CUDA(cudaStreamWaitEvent(sd, ev_start, 0));

  // origin of every local result tile in this rank's grid of 128x128 blocks (L2 rasterisation hint)
  std::vector<int32_t> blk_row0(Mt, 0), blk_col0(Nt, 0);
  {
    int64_t o = 0;
    for (int i = r; i < Mt; i += Pr) { blk_row0[i] = (int32_t)o; o += ceil_div64(P.m_ext[i], kGemmBM); }
    bool ok = o < 65535;
    o = 0;
    for (int j = c; j < Nt; j += Pc) { blk_col0[j] = (int32_t)o; o += ceil_div64(P.n_ext[j], kGemmBN); }
    if (!ok || o >= 65535) { blk_row0.clear(); blk_col0.clear(); }  // too large to encode: no hints
  }
  // TADEV_SUMMA_TRACE=1: per-window timeline (ms since the start) of staging, GEMM and download
  static const bool trace = getenv("TADEV_SUMMA_TRACE") && atoi(getenv("TADEV_SUMMA_TRACE"));
  struct Mark { const char* what; int block, window; cudaEvent_t ev; };
  std::vector<Mark> marks;
  auto mark = [&](const char* what, int blk, int win, cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    if (R.event(&e, cudaEventDefault) != TADEV_OK) return;
    cudaEventRecord(e, st);
    marks.push_back({what, blk, win, e});
  };
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;  // around every GEMM launch (stats->gemm_ms)
  std::vector<char> touched((size_t)Mt * Nt, 0);
  int64_t npairs = 0, nlaunches = 0, bcast_bytes = 0, h2d_bytes = 0, d2h_bytes = 0, lazy_tiles = 0;
  double flops = 0.0;
  struct Contribution { int64_t key; const double* A; const double* B; int k; };
  std::vector<Contribution> contrib;
  std::vector<tadev_gemm_group> groups;
  std::vector<tadev_gemm_task> tasks;
  std::vector<double*> c_dev((size_t)Mt * Nt, nullptr);  // device address of each local result tile
  int64_t wcount = 0;                                     // global window counter -> ring slot
  std::vector<char> b_cached(S.steps.size(), 0);

  std::vector<uint64_t> prov_tok;
  std::vector<double*> prov_dst;

  // ---- tile lists are built on the device (tilelist.cu): the host stages O(rows + cols) panel-tile tables per
  //      step, the O(pairs) enumeration / chaining / rasterisation runs as kernels on the compute stream.
  //      TADEV_HOST_LISTS=1 selects the host-built lists (A/B comparison; the two paths give identical launches).
  const bool host_lists = getenv("TADEV_HOST_LISTS") && atoi(getenv("TADEV_HOST_LISTS"));
  const bool use_tl = !host_lists;
  TileListBuilder TL;
  const int ncl = c < Nt ? (Nt - c + Pc - 1) / Pc : 0;
  float list_ms = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> list_events;
  if (use_tl) {
    size_t max_tasks = 1;
    for (int b = 0; b < nb; ++b)
      for (const Window& win : bwins[b]) {
        size_t t = 0;
        for (int x : win.steps) { const BlockStep& bs = bsteps[b][x]; if (bs.compute) t += bs.a_rows.size() * bs.st->b_cols.size(); }
        max_tasks = std::max(max_tasks, t);
      }
    int rc = TL.init(ctx, s0, Pr, Pc, r, c, Mt, Nt, Kt, P.m_ext, P.n_ext, P.k_ext, P.a_norms, P.b_norms, P.c_norms, P.threshold,
                     P.accumulate, max_tasks);
    if (TL.d_base) R.buffers.push_back(TL.d_base);
    if (rc) return rc;
  }
  const bool a_kin = P.opA == TADEV_OP_N, b_kin = P.opB == TADEV_OP_T;  // operand stored k-contiguous: needs a tensor map

  if (trace_host) fprintf(stderr, "[tadev summa rank %d] host: list builder ready at %.3f ms\n", ctx->rank, host_ms());
  for (int b = 0; b < nb; ++b) {
    bool block_listed = false;  // the block's result-tile table is on the device (first listed window)
    // ---- device addresses of this block's result tiles
    const int cslot = b & 1;
    if (c_host) {
      if (c_used[cslot]) TADEV_CHECK_CUDA(cudaStreamWaitEvent(s0, d2h_done[cslot], 0));  // block b-2 downloaded
      size_t o = 0;
      for (int x = brange[b].first; x < brange[b].second; ++x)
        for (int j = c; j < Nt; j += Pc) {
          const int i = my_rows[x];
          if (!c_local(i, j)) continue;
          c_dev[(size_t)i * Nt + j] = carena[cslot] + o;
          const size_t e = (size_t)P.m_ext[i] * P.n_ext[j];
          if (P.accumulate) {  // C += : bring the previous contents
            TADEV_CHECK_CUDA(cudaMemcpyAsync(carena[cslot] + o, P.c_tiles[(size_t)i * Nt + j], e * 8, cudaMemcpyHostToDevice, s0));
            h2d_bytes += (int64_t)e * 8;
          }
          o += pad2(e);
        }
    } else {
      for (int x = brange[b].first; x < brange[b].second; ++x)
        for (int j = c; j < Nt; j += Pc) c_dev[(size_t)my_rows[x] * Nt + j] = P.c_tiles[(size_t)my_rows[x] * Nt + j];
    }

    for (int wi = 0; wi < (int)bwins[b].size(); ++wi, ++wcount) {
      const Window& win = bwins[b][wi];
      const int d = (int)(wcount % D);
      double* buf = ring[d];
      size_t cursor = 0;
      bool any_bcast = false, any_h2d = false, ring_touched = false;
      std::vector<std::vector<const double*>> a_ptrs(win.steps.size()), b_ptrs(win.steps.size());
      std::vector<char> a_trans(win.steps.size(), 0), b_trans(win.steps.size(), 0);  // panel lives in the ring / B cache
      std::vector<Bcast> row_bcasts, col_bcasts;
      if (need_ring && win.bytes > 0 && buf_used[d]) {  // the GEMM that last read this slot is done
        TADEV_CHECK_CUDA(cudaStreamWaitEvent(sc, buf_free[d], 0));
        TADEV_CHECK_CUDA(cudaStreamWaitEvent(sh, buf_free[d], 0));
      }
      // ---- panels of this window. Panels that travel are merged per (communicator, root): the root's panels of
      //      all steps of the window form ONE contiguous byte range (in its arena when the array was laid out panel by
      //      panel, else packed into the ring) and go out as ONE ncclBroadcast. Block-sparse panels are small (config 3:
      //      ~26 MB per step); one NCCL operation per step ran at ~25 GB/s (one channel each), one per root and window
      //      uses all channels (profiles/r02_c3_n2_*.json).
      struct Panel {
        size_t wsi, si; bool is_a, bcast, mine, to_cache; int root; size_t elems;
        std::vector<const double*> src; std::vector<size_t> el;
      };
      std::vector<Panel> panels;
      for (size_t wsi = 0; wsi < win.steps.size(); ++wsi) {
        const BlockStep& bs = bsteps[b][win.steps[wsi]];
        const SummaStep& st = *bs.st;
        const int k = st.k;
        const size_t si = (size_t)(bs.st - &S.steps[0]);
        if (bs.bcast_a || bs.compute) {  // A panel (travels along my grid row; root column k % Pc)
          Panel p{wsi, si, true, bs.bcast_a, c == k % Pc, false, k % Pc, bs.a_elems, {}, {}};
          for (int i : bs.a_rows) {
            const double* tp = P.a_tiles[(size_t)i * Kt + k];
            TADEV_REQUIRE(!p.mine || tp, "tadev_summa_f64: A tile (%d,%d) is owned by this rank but has no data", i, k);
            p.src.push_back(tp); p.el.push_back(tile_a_elems(i, k));
          }
          panels.push_back(std::move(p));
        }
        if (bs.bcast_b || bs.compute) {  // B panel (travels along my grid column; root row k % Pr); cached across row blocks
          const bool to_cache = b_cache && (st.bcast_b || b_stg);
          if (to_cache && b_cached[si]) {
            size_t o = 0;
            for (int j : st.b_cols) { b_ptrs[wsi].push_back(bcache + b_cache_off[si] + o); o += pad2(tile_b_elems(k, j)); }
            b_trans[wsi] = 1;
            continue;
          }
          Panel p{wsi, si, false, bs.bcast_b, r == k % Pr, to_cache, k % Pr, bs.b_elems, {}, {}};
          for (int j : st.b_cols) {
            const double* tp = P.b_tiles[(size_t)k * Nt + j];
            TADEV_REQUIRE(!p.mine || tp, "tadev_summa_f64: B tile (%d,%d) is owned by this rank but has no data", k, j);
            p.src.push_back(tp); p.el.push_back(tile_b_elems(k, j));
          }
          panels.push_back(std::move(p));
        }
      }
      // travelling panels first, grouped by (operand, root), steps ascending inside a group
      std::stable_sort(panels.begin(), panels.end(), [](const Panel& x, const Panel& y) {
        const int kx = (x.is_a ? 0 : 2) + (x.bcast ? 0 : 1), ky = (y.is_a ? 0 : 2) + (y.bcast ? 0 : 1);
        if (kx != ky) return kx < ky;
        if (x.bcast && x.root != y.root) return x.root < y.root;
        return x.wsi < y.wsi;
      });
      for (size_t p0 = 0; p0 < panels.size();) {
        size_t p1 = p0 + 1;
        if (panels[p0].bcast)
          while (p1 < panels.size() && panels[p1].bcast && panels[p1].is_a == panels[p0].is_a && panels[p1].root == panels[p0].root) ++p1;
        const Panel& first = panels[p0];
        const bool on_host = first.is_a ? a_host : b_host;
        const tadev_tile_provider provider = first.is_a ? (a_lazy ? P.a_provider : nullptr) : (b_lazy ? P.b_provider : nullptr);
        void* user = first.is_a ? P.a_user : P.b_user;
        auto& ptrs_of = first.is_a ? a_ptrs : b_ptrs;
        auto& trans_of = first.is_a ? a_trans : b_trans;
        size_t total = 0;
        for (size_t q = p0; q < p1; ++q) total += panels[q].elems;
        // in place: device-resident source tiles that need no transport, or a root whose tiles are one contiguous range
        bool inplace = !on_host && !provider && (!first.bcast || first.mine);
        if (inplace && first.bcast) {
          const double* base = nullptr;
          size_t off = 0;
          for (size_t q = p0; q < p1 && inplace; ++q)
            for (size_t n = 0; n < panels[q].src.size(); ++n) {
              if (!base) base = panels[q].src[n];
              if (panels[q].src[n] != base + off || (reinterpret_cast<uintptr_t>(panels[q].src[n]) & 15)) { inplace = false; break; }
              off += pad2(panels[q].el[n]);
            }
        }
        double* dest = nullptr;
        if (!inplace || first.to_cache) dest = first.to_cache ? bcache + b_cache_off[first.si] : (buf ? buf + cursor : nullptr);
        size_t doff = 0;
        const double* first_ptr = nullptr;
        for (size_t q = p0; q < p1; ++q) {
          const Panel& pn = panels[q];
          auto& out = ptrs_of[pn.wsi];
          out.assign(pn.src.size(), nullptr);
          if (inplace) {
            for (size_t n = 0; n < pn.src.size(); ++n) out[n] = pn.src[n];
          } else {
            TADEV_REQUIRE(dest || pn.src.empty(), "tadev_summa_f64: internal: no staging buffer for a panel");
            double* panel = dest + doff;
            if (pn.mine || !pn.bcast) {  // I hold the source: materialise the panel
              if (provider) {  // lazy tiles: src[] holds opaque tokens; the provider fills the panel on the staging stream
                prov_tok.clear(); prov_dst.clear();
                size_t o = 0;
                for (size_t n = 0; n < pn.src.size(); ++n) {
                  prov_tok.push_back((uint64_t)reinterpret_cast<uintptr_t>(pn.src[n]));
                  prov_dst.push_back(panel + o);
                  o += pad2(pn.el[n]);
                }
                if (!pn.src.empty()) {
                  int rc = provider(user, (tadev_stream)sh, (int)pn.src.size(), prov_tok.data(), prov_dst.data(), pn.el.data());
                  if (rc) return rc;
                  lazy_tiles += (int64_t)pn.src.size();
                }
                any_h2d = true;
              } else if (on_host) {
                size_t o = 0;
                for (size_t n = 0; n < pn.src.size(); ++n) {
                  TADEV_CHECK_CUDA(cudaMemcpyAsync(panel + o, pn.src[n], pn.el[n] * 8, cudaMemcpyHostToDevice, sh));
                  h2d_bytes += (int64_t)pn.el[n] * 8;
                  o += pad2(pn.el[n]);
                }
                any_h2d = true;
              } else {  // device-resident but scattered: pack
                size_t o = 0;
                for (size_t n = 0; n < pn.src.size(); ++n) {
                  TADEV_CHECK_CUDA(cudaMemcpyAsync(panel + o, pn.src[n], pn.el[n] * 8, cudaMemcpyDeviceToDevice, sc));
                  o += pad2(pn.el[n]);
                }
              }
            }
            size_t o = 0;
            for (size_t n = 0; n < pn.src.size(); ++n) { out[n] = panel + o; o += pad2(pn.el[n]); }
            trans_of[pn.wsi] = 1;
          }
          if (!first_ptr && !out.empty()) first_ptr = out[0];
          doff += pn.elems;
          if (pn.to_cache) b_cached[pn.si] = 1;
        }
        if (!inplace && !first.to_cache) { cursor += total; ring_touched = true; }
        if (inplace && first.to_cache && total && first_ptr) {
          // an in-place broadcast source lives outside the cache: mirror it so later blocks find it
          TADEV_CHECK_CUDA(cudaMemcpyAsync(dest, first_ptr, total * 8, cudaMemcpyDeviceToDevice, sc));
          ring_touched = true;  // (a copy on sc the GEMM must wait for)
        }
        if (first.bcast && total) {
          (first.is_a ? row_bcasts : col_bcasts).push_back({const_cast<double*>(first_ptr), total * 8, first.root});
          bcast_bytes += (int64_t)total * 8;
          any_bcast = true;
        }
        p0 = p1;
      }
      // ---- ordering: H2D (sh) -> broadcasts (sc) -> GEMM (s0)
      if (any_h2d) {
        mark("h2d_done", b, wi, sh);
        TADEV_CHECK_CUDA(cudaEventRecord(h2d_done[d], sh));
        TADEV_CHECK_CUDA(cudaStreamWaitEvent(any_bcast ? sc : s0, h2d_done[d], 0));
      }
      if (trace) {
        fprintf(stderr, "[tadev summa rank %d] block %d window %d slot %d: %zu steps, h2d %d, ring_touched %d\n", ctx->rank, b, wi, d,
                win.steps.size(), (int)any_h2d, (int)ring_touched);
        for (auto& bc : row_bcasts) fprintf(stderr, "[tadev summa rank %d]   row bcast root %d bytes %zu\n", ctx->rank, bc.root, bc.bytes);
        for (auto& bc : col_bcasts) fprintf(stderr, "[tadev summa rank %d]   col bcast root %d bytes %zu\n", ctx->rank, bc.root, bc.bytes);
      }
      if (any_bcast) {
        // same (communicator, k) order on every rank of a group => no cross-communicator deadlock
        if (!row_bcasts.empty()) {
          TADEV_CHECK_NCCL(ncclGroupStart());
          for (auto& bc : row_bcasts) TADEV_CHECK_NCCL(ncclBroadcast(bc.ptr, bc.ptr, bc.bytes, ncclChar, bc.root, row_comm, sc));
          TADEV_CHECK_NCCL(ncclGroupEnd());
        }
        if (!col_bcasts.empty()) {
          TADEV_CHECK_NCCL(ncclGroupStart());
          for (auto& bc : col_bcasts) TADEV_CHECK_NCCL(ncclBroadcast(bc.ptr, bc.ptr, bc.bytes, ncclChar, bc.root, col_comm, sc));
          TADEV_CHECK_NCCL(ncclGroupEnd());
        }
      }
      if (any_bcast || ring_touched) {  // D2D packs also run on sc
        mark("bcast_done", b, wi, sc);
        TADEV_CHECK_CUDA(cudaEventRecord(panel_ready[d], sc));
        TADEV_CHECK_CUDA(cudaStreamWaitEvent(s0, panel_ready[d], 0));
      }

      // ---- tile lists of this window, built on the device from staged panel-tile tables
      if (use_tl) {
        const int li0 = brange[b].first, li1 = brange[b].second, nrows = li1 - li0;
        std::vector<int> wsteps;
        for (size_t wsi = 0; wsi < win.steps.size(); ++wsi) if (bsteps[b][win.steps[wsi]].compute) wsteps.push_back((int)wsi);
        if (!wsteps.empty() && nrows > 0 && ncl > 0) {
          const int nws = (int)wsteps.size();
          // fast (TMA) kernel needs 16-byte aligned operand rows: even leading dimensions, aligned tile addresses
          bool fast = !ctx->force_generic_gemm;
          size_t nmaps = 0;
          for (int wsi : wsteps) {
            const BlockStep& bs = bsteps[b][win.steps[wsi]];
            const int64_t kx = P.k_ext[bs.st->k];
            for (size_t ai = 0; ai < bs.a_rows.size() && fast; ++ai)
              fast = !(((a_kin ? kx : P.m_ext[bs.a_rows[ai]]) & 1) || (reinterpret_cast<uintptr_t>(a_ptrs[wsi][ai]) & 15));
            for (size_t bj = 0; bj < bs.st->b_cols.size() && fast; ++bj)
              fast = !(((b_kin ? kx : P.n_ext[bs.st->b_cols[bj]]) & 1) || (reinterpret_cast<uintptr_t>(b_ptrs[wsi][bj]) & 15));
            if (a_kin && a_trans[wsi]) nmaps += bs.a_rows.size();
            if (b_kin && b_trans[wsi]) nmaps += bs.st->b_cols.size();
          }
          auto al = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
          const size_t o_k = 0, o_a = al((size_t)nws * 4, 16), o_b = o_a + (size_t)nws * nrows * sizeof(TlTile);
          const size_t o_c = o_b + (size_t)nws * ncl * sizeof(TlTile);
          const size_t o_m = al(o_c + (block_listed ? 0 : (size_t)nrows * ncl * 8), 128), total_b = o_m + (fast ? nmaps : 0) * 128;
          StageLease L;
          int rc = L.acquire(ctx, s0, total_b + 128);
          if (rc) return rc;
          // (the staging buffers come from cudaMalloc / cudaMallocHost: 256-byte aligned, as tensor maps require)
          char* h = static_cast<char*>(L.h);
          char* dblk = static_cast<char*>(L.d);
          int32_t* hk = reinterpret_cast<int32_t*>(h + o_k);
          TlTile* ha = reinterpret_cast<TlTile*>(h + o_a);
          TlTile* hb = reinterpret_cast<TlTile*>(h + o_b);
          memset(ha, 0, (size_t)nws * nrows * sizeof(TlTile));
          memset(hb, 0, (size_t)nws * ncl * sizeof(TlTile));
          size_t nm = 0;
          bool created = false;
          for (int x = 0; x < nws; ++x) {
            const int wsi = wsteps[x];
            const BlockStep& bs = bsteps[b][win.steps[wsi]];
            const int k = bs.st->k;
            const int kx = (int)P.k_ext[k];
            hk[x] = k;
            for (size_t ai = 0; ai < bs.a_rows.size(); ++ai) {
              const int i = bs.a_rows[ai];
              TlTile& t = ha[(size_t)x * nrows + ((i - r) / Pr - li0)];
              t.ptr = a_ptrs[wsi][ai];
              if (fast && a_kin && kx > 0) {
                if (a_trans[wsi]) {
                  rc = tadev_ws_encode_map(ctx, h + o_m + nm * 128, t.ptr, (int)P.m_ext[i], kx);
                  t.map = dblk + o_m + nm * 128; ++nm;
                } else rc = tadev_ws_cached_map(ctx, t.ptr, (int)P.m_ext[i], kx, &t.map, &created);
                if (rc) return rc;
              }
            }
            for (size_t bj = 0; bj < bs.st->b_cols.size(); ++bj) {
              const int j = bs.st->b_cols[bj];
              TlTile& t = hb[(size_t)x * ncl + (j - c) / Pc];
              t.ptr = b_ptrs[wsi][bj];
              if (fast && b_kin && kx > 0) {
                if (b_trans[wsi]) {
                  rc = tadev_ws_encode_map(ctx, h + o_m + nm * 128, t.ptr, (int)P.n_ext[j], kx);
                  t.map = dblk + o_m + nm * 128; ++nm;
                } else rc = tadev_ws_cached_map(ctx, t.ptr, (int)P.n_ext[j], kx, &t.map, &created);
                if (rc) return rc;
              }
            }
          }
          double* const* d_cstaged = nullptr;
          if (!block_listed) {
            double** hc = reinterpret_cast<double**>(h + o_c);
            for (int lr = 0; lr < nrows; ++lr)
              for (int lj = 0; lj < ncl; ++lj) hc[(size_t)lr * ncl + lj] = c_dev[(size_t)my_rows[li0 + lr] * Nt + (c + lj * Pc)];
            d_cstaged = reinterpret_cast<double* const*>(dblk + o_c);
          }
          if (created && (rc = tadev_ws_flush_new_maps(ctx))) return rc;
          rc = tadev_stage_upload(ctx, s0, L.d, L.h, total_b, L.uploaded);
          if (rc) return rc;
          cudaEvent_t l0 = nullptr, l1 = nullptr, g0 = nullptr, g1 = nullptr;
          if (R.event(&l0, cudaEventDefault) == TADEV_OK && R.event(&l1, cudaEventDefault) == TADEV_OK) list_events.push_back({l0, l1});
          if (R.event(&g0, cudaEventDefault) == TADEV_OK && R.event(&g1, cudaEventDefault) == TADEV_OK) gemm_events.push_back({g0, g1});
          mark("tables_up", b, wi, s0);
          rc = TL.build_and_launch(P.opA, P.opB, P.alpha, li0, li1, nws, reinterpret_cast<const int32_t*>(dblk + o_k),
                                   reinterpret_cast<const TlTile*>(dblk + o_a), reinterpret_cast<const TlTile*>(dblk + o_b), d_cstaged,
                                   fast, l0, l1, g0, g1);
          if (rc) return rc;
          block_listed = true;
          ++nlaunches;
          if (trace_host) fprintf(stderr, "[tadev summa rank %d] host: block %d window %d lists+gemm enqueued at %.3f ms\n", ctx->rank, b, wi, host_ms());
          mark("gemm_done", b, wi, s0);
        }
      }
      // ---- (TADEV_HOST_LISTS=1) grouped GEMM descriptors of this window built on the host: chain contributions per result tile
      contrib.clear();
      if (!use_tl)
      for (size_t wsi = 0; wsi < win.steps.size(); ++wsi) {
        const BlockStep& bs = bsteps[b][win.steps[wsi]];
        if (!bs.compute) continue;
        const SummaStep& st = *bs.st;
        size_t ai = 0;
        for (int64_t pp = st.pair_begin; pp < st.pair_end; ++pp) {
          const int i = S.pair_i[pp], j = S.pair_j[pp];
          if (bs.a_rows.empty() || i < bs.a_rows.front() || i > bs.a_rows.back()) continue;  // other row block
          while (bs.a_rows[ai] != i) ++ai;  // pairs are row-major: rows appear in a_rows order
          const size_t bj = std::lower_bound(st.b_cols.begin(), st.b_cols.end(), j) - st.b_cols.begin();
          contrib.push_back({(int64_t)i * Nt + j, a_ptrs[wsi][ai], b_ptrs[wsi][bj], (int)P.k_ext[st.k]});
          flops += 2.0 * (double)P.m_ext[i] * (double)P.n_ext[j] * (double)P.k_ext[st.k];
        }
      }
      npairs += (int64_t)contrib.size();
      if (!contrib.empty()) {
        std::stable_sort(contrib.begin(), contrib.end(), [](const Contribution& x, const Contribution& y) { return x.key < y.key; });
        groups.clear(); tasks.clear();
        for (size_t n = 0; n < contrib.size(); ++n) {
          if (n == 0 || contrib[n].key != contrib[n - 1].key) {
            const int i = (int)(contrib[n].key / Nt), j = (int)(contrib[n].key % Nt);
            double* ct = c_dev[contrib[n].key];
            TADEV_REQUIRE(ct, "tadev_summa_f64: result tile (%d,%d) is non-zero and local but has no storage", i, j);
            if (!groups.empty()) groups.back().task_end = (int32_t)tasks.size();
            tadev_gemm_group G{ct, (int32_t)P.m_ext[i], (int32_t)P.n_ext[j], (int32_t)tasks.size(), 0,
                               (P.accumulate || touched[contrib[n].key]) ? 1 : 0,
                               blk_row0.empty() ? 0 : (int32_t)((uint32_t)(blk_row0[i] + 1) << 16 | (uint32_t)(blk_col0[j] + 1))};
            touched[contrib[n].key] = 1;
            groups.push_back(G);
          }
          tasks.push_back({contrib[n].A, contrib[n].B, contrib[n].k, 0});
        }
        groups.back().task_end = (int32_t)tasks.size();
        GemmTimingHook& hook = tadev_gemm_timing_hook();
        cudaEvent_t g0 = nullptr, g1 = nullptr;
        if (R.event(&g0, cudaEventDefault) == TADEV_OK && R.event(&g1, cudaEventDefault) == TADEV_OK) {
          hook.before = g0; hook.after = g1;
          gemm_events.push_back({g0, g1});
        }
        int rc = tadev_gemm_grouped_f64(ctx, s0, P.opA, P.opB, P.alpha, groups.data(), (int)groups.size(), tasks.data(),
                                        (int)tasks.size());
        hook.before = hook.after = nullptr;
        if (rc) return rc;
        ++nlaunches;
        mark("gemm_done", b, wi, s0);
      }
      if (need_ring && win.bytes > 0) {
        TADEV_CHECK_CUDA(cudaEventRecord(buf_free[d], s0));
        buf_used[d] = 1;
      }
    }

    // ---- block epilogue: zero-fill untouched non-zero tiles, then hand the block to the host
    if (use_tl && block_listed) {
      int rc = TL.zero_untouched(brange[b].first, brange[b].second);
      if (rc) return rc;
    } else
    for (int x = brange[b].first; x < brange[b].second; ++x)
      for (int j = c; j < Nt; j += Pc) {
        const size_t key = (size_t)my_rows[x] * Nt + j;
        if (!P.accumulate && c_dev[key] && !touched[key] && (!P.c_norms || P.c_norms[key] >= P.threshold))
          TADEV_CHECK_CUDA(cudaMemsetAsync(c_dev[key], 0, (size_t)P.m_ext[my_rows[x]] * P.n_ext[j] * 8, s0));
      }
    if (c_host && cblock_elems[b]) {
      TADEV_CHECK_CUDA(cudaEventRecord(c_done[cslot], s0));
      TADEV_CHECK_CUDA(cudaStreamWaitEvent(sd, c_done[cslot], 0));
      for (int x = brange[b].first; x < brange[b].second; ++x)
        for (int j = c; j < Nt; j += Pc) {
          const size_t key = (size_t)my_rows[x] * Nt + j;
          if (!c_local(my_rows[x], j)) continue;
          const size_t e = (size_t)P.m_ext[my_rows[x]] * P.n_ext[j];
          TADEV_CHECK_CUDA(cudaMemcpyAsync(P.c_tiles[key], c_dev[key], e * 8, cudaMemcpyDeviceToHost, sd));
          d2h_bytes += (int64_t)e * 8;
        }
      mark("d2h_done", b, -1, sd);
      TADEV_CHECK_CUDA(cudaEventRecord(d2h_done[cslot], sd));
      c_used[cslot] = true;
    }
  }

  // ---- join all streams on s0, time, release
  TADEV_CHECK_CUDA(cudaEventRecord(ev_aux, sc));
  TADEV_CHECK_CUDA(cudaStreamWaitEvent(s0, ev_aux, 0));
  TADEV_CHECK_CUDA(cudaEventRecord(ev_aux, sh));
  TADEV_CHECK_CUDA(cudaStreamWaitEvent(s0, ev_aux, 0));
  if (sd != s0) {
    TADEV_CHECK_CUDA(cudaEventRecord(ev_aux, sd));
    TADEV_CHECK_CUDA(cudaStreamWaitEvent(s0, ev_aux, 0));
  }
  TADEV_CHECK_CUDA(cudaEventRecord(ev_end, s0));
  if (trace_host) fprintf(stderr, "[tadev summa rank %d] host: everything enqueued at %.3f ms\n", ctx->rank, host_ms());
  TADEV_CHECK_CUDA(cudaEventSynchronize(ev_end));
  TADEV_CHECK_CUDA(cudaGetLastError());
  float ms = 0, gemm_ms = 0;
  TADEV_CHECK_CUDA(cudaEventElapsedTime(&ms, ev_start, ev_end));
  for (auto& ge : gemm_events) {
    float t = 0;
    if (cudaEventElapsedTime(&t, ge.first, ge.second) == cudaSuccess) gemm_ms += t; else cudaGetLastError();
  }
  for (auto& mk : marks) {
    float t = 0;
    cudaEventSynchronize(mk.ev);
    cudaEventElapsedTime(&t, ev_start, mk.ev);
    fprintf(stderr, "[tadev summa rank %d] %-9s block %d window %2d  t=%9.3f ms\n", ctx->rank, mk.what, mk.block, mk.window, t);
  }
  for (auto& le : list_events) {
    float t = 0;
    if (cudaEventElapsedTime(&t, le.first, le.second) == cudaSuccess) list_ms += t; else cudaGetLastError();
  }
  if (use_tl) {
    unsigned long long np = 0;
    int rc = TL.read_counters(&np, &flops);
    if (rc) return rc;
    npairs = (int64_t)np;
  }
  R.ok = true;  // all streams were joined on s0 and s0 has drained: plain stream-ordered release
  if (stats) {
    stats->nsteps = (int64_t)S.steps.size();
    stats->nsteps_skipped = S.nskipped;
    stats->npairs = npairs;
    stats->nlaunches = nlaunches;
    stats->flops = flops;
    stats->bcast_bytes = bcast_bytes;
    stats->device_ms = ms;
    stats->h2d_bytes = h2d_bytes;
    stats->d2h_bytes = d2h_bytes;
    stats->row_blocks = nb;
    stats->lazy_tiles = lazy_tiles;
    stats->gemm_ms = gemm_ms;
    stats->list_ms = list_ms;
  }
  return TADEV_OK;
}
