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
#include <numeric>

#include "common.h"
#include "summa_schedule.h"

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
  TADEV_CHECK_NCCL(ncclCommInitRank(&ctx->world, nranks, id, rank));
  ctx->rank = rank; ctx->nranks = nranks; ctx->Pr = Pr; ctx->Pc = Pc;
  // leave a few SMs to NCCL's broadcast CTAs so panel traffic overlaps the persistent GEMM
  ctx->gemm_sm_reserve = (Pr * Pc > 1) ? 4 : 0;
  if (const char* e = getenv("TADEV_SM_RESERVE")) ctx->gemm_sm_reserve = atoi(e);
  const bool in_grid = rank < Pr * Pc;
  ctx->my_r = in_grid ? rank / Pc : -1;  // proc_grid.h: rank_row = rank / proc_cols
  ctx->my_c = in_grid ? rank % Pc : -1;
  TADEV_CHECK_NCCL(ncclCommSplit(ctx->world, in_grid ? ctx->my_r : NCCL_SPLIT_NOCOLOR, ctx->my_c, &ctx->row_comm, nullptr));
  TADEV_CHECK_NCCL(ncclCommSplit(ctx->world, in_grid ? ctx->my_c : NCCL_SPLIT_NOCOLOR, ctx->my_r, &ctx->col_comm, nullptr));
  return TADEV_OK;
}

extern "C" int tadev_comm_destroy(tadev_ctx* ctx) {
  if (!ctx) return TADEV_OK;
  if (ctx->row_comm) { ncclCommDestroy(ctx->row_comm); ctx->row_comm = nullptr; }
  if (ctx->col_comm) { ncclCommDestroy(ctx->col_comm); ctx->col_comm = nullptr; }
  if (ctx->world) { ncclCommDestroy(ctx->world); ctx->world = nullptr; }
  ctx->rank = 0; ctx->nranks = 1; ctx->Pr = ctx->Pc = 1; ctx->my_r = ctx->my_c = 0;
  ctx->gemm_sm_reserve = 0;
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

namespace {

struct PanelTile { int idx; size_t off; size_t elems; };  // tile row/col index, offset (doubles) in panel

inline size_t pad2(size_t n) { return (n + 1) & ~size_t(1); }  // keep every tile 16-byte aligned

struct Window {
  std::vector<int> steps;  // indices into schedule.steps
  size_t bytes = 0;        // panel-buffer bytes needed
};

}  // namespace

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
  TADEV_CHECK_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s0 = ctx->streams[0];
  cudaStream_t sc = ctx->comm_stream[0];

  SummaSchedule S = make_summa_schedule(Pr, Pc, r, c, Mt, Nt, Kt, P.a_norms, P.b_norms, P.c_norms, P.threshold);

  // ---- windows of steps (one grouped-GEMM launch each)
  int W = P.steps_per_launch;
  if (W <= 0) {
    if (!multi) W = std::max(1, Kt);
    else {
      double avgk = 0;
      for (int k = 0; k < Kt; ++k) avgk += (double)P.k_ext[k];
      avgk = Kt ? avgk / Kt : 1.0;
      W = (int)std::min<double>(64.0, std::max(1.0, std::ceil(4096.0 / std::max(1.0, avgk))));
    }
  }
  const size_t kMaxWindowBytes = size_t(3) << 30;
  auto a_panel_elems = [&](const SummaStep& st) { size_t e = 0; for (int i : st.a_rows) e += pad2((size_t)P.m_ext[i] * P.k_ext[st.k]); return e; };
  auto b_panel_elems = [&](const SummaStep& st) { size_t e = 0; for (int j : st.b_cols) e += pad2((size_t)P.k_ext[st.k] * P.n_ext[j]); return e; };
  std::vector<Window> windows;
  {
    Window cur;
    int ncomp = 0;
    for (int si = 0; si < (int)S.steps.size(); ++si) {
      const SummaStep& st = S.steps[si];
      size_t need = 0;
      if (st.bcast_a) need += a_panel_elems(st) * 8;
      if (st.bcast_b) need += b_panel_elems(st) * 8;
      if (!cur.steps.empty() && (ncomp >= W || cur.bytes + need > kMaxWindowBytes)) {
        windows.push_back(std::move(cur)); cur = Window(); ncomp = 0;
      }
      cur.steps.push_back(si); cur.bytes += need;
      if (st.compute) ++ncomp;
    }
    if (!cur.steps.empty()) windows.push_back(std::move(cur));
  }
  size_t max_bytes = 0;
  for (auto& w : windows) max_bytes = std::max(max_bytes, w.bytes);
  const int D = std::max(2, P.depth > 0 ? P.depth : 2);

  // ---- resources
  cudaEvent_t ev_start, ev_end, ev_comm_done;
  TADEV_CHECK_CUDA(cudaEventCreate(&ev_start));
  TADEV_CHECK_CUDA(cudaEventCreate(&ev_end));
  TADEV_CHECK_CUDA(cudaEventCreateWithFlags(&ev_comm_done, cudaEventDisableTiming));
  std::vector<double*> ring(D, nullptr);
  std::vector<cudaEvent_t> panel_ready(D), buf_free(D);
  std::vector<char> buf_used(D, 0);
  const bool need_ring = multi && max_bytes > 0;
  if (need_ring) {
    for (int d = 0; d < D; ++d) {
      int rc = tadev_alloc(ctx, max_bytes, (void**)&ring[d], s0);
      if (rc) return rc;
      TADEV_CHECK_CUDA(cudaEventCreateWithFlags(&panel_ready[d], cudaEventDisableTiming));
      TADEV_CHECK_CUDA(cudaEventCreateWithFlags(&buf_free[d], cudaEventDisableTiming));
    }
  }
  TADEV_CHECK_CUDA(cudaEventRecord(ev_start, s0));
  if (need_ring) TADEV_CHECK_CUDA(cudaStreamWaitEvent(sc, ev_start, 0));

  std::vector<char> touched((size_t)Mt * Nt, 0);
  int64_t npairs = 0, nlaunches = 0, bcast_bytes = 0;
  double flops = 0.0;

  struct Contribution { int64_t key; const double* A; const double* B; int k; };
  std::vector<Contribution> contrib;
  std::vector<tadev_gemm_group> groups;
  std::vector<tadev_gemm_task> tasks;

  for (int wi = 0; wi < (int)windows.size(); ++wi) {
    const Window& win = windows[wi];
    const int d = wi % D;
    double* buf = need_ring ? ring[d] : nullptr;
    size_t cursor = 0;  // doubles
    bool any_comm = false;
    // per-step resolved tile pointers for the GEMM tasks
    std::vector<std::vector<const double*>> a_ptrs(win.steps.size()), b_ptrs(win.steps.size());

    if (need_ring && win.bytes > 0 && buf_used[d]) TADEV_CHECK_CUDA(cudaStreamWaitEvent(sc, buf_free[d], 0));

    struct Bcast { void* ptr; size_t bytes; int root; };
    std::vector<Bcast> row_bcasts, col_bcasts;

    for (size_t wsi = 0; wsi < win.steps.size(); ++wsi) {
      const SummaStep& st = S.steps[win.steps[wsi]];
      const int k = st.k;
      // ---- A panel (travels along my grid row)
      {
        auto& ptrs = a_ptrs[wsi];
        ptrs.resize(st.a_rows.size(), nullptr);
        const int root = k % Pc;
        if (st.bcast_a) {
          const size_t elems = a_panel_elems(st);
          double* panel = buf + cursor;
          bool inplace = false;
          if (c == root) {
            // owner: tiles adjacent in panel order can be sent from where they live
            inplace = true;
            const double* first = P.a_tiles[(size_t)st.a_rows[0] * Kt + k];
            size_t off = 0;
            for (size_t n = 0; n < st.a_rows.size(); ++n) {
              const double* tp = P.a_tiles[(size_t)st.a_rows[n] * Kt + k];
              TADEV_REQUIRE(tp, "tadev_summa_f64: A tile (%d,%d) is owned by this rank but has no data", st.a_rows[n], k);
              if (tp != first + off || (reinterpret_cast<uintptr_t>(tp) & 15)) inplace = false;
              off += pad2((size_t)P.m_ext[st.a_rows[n]] * P.k_ext[k]);
            }
            if (inplace) panel = const_cast<double*>(first);
            else {
              size_t o = 0;
              for (size_t n = 0; n < st.a_rows.size(); ++n) {
                const size_t e = (size_t)P.m_ext[st.a_rows[n]] * P.k_ext[k];
                TADEV_CHECK_CUDA(cudaMemcpyAsync(panel + o, P.a_tiles[(size_t)st.a_rows[n] * Kt + k], e * 8,
                                                 cudaMemcpyDeviceToDevice, sc));
                o += pad2(e);
              }
            }
          }
          if (!inplace) cursor += elems;
          size_t o = 0;
          for (size_t n = 0; n < st.a_rows.size(); ++n) {
            ptrs[n] = panel + o;
            o += pad2((size_t)P.m_ext[st.a_rows[n]] * P.k_ext[k]);
          }
          row_bcasts.push_back({panel, elems * 8, root});
          bcast_bytes += (int64_t)elems * 8;
          any_comm = true;
        } else if (st.compute) {
          for (size_t n = 0; n < st.a_rows.size(); ++n) {
            ptrs[n] = P.a_tiles[(size_t)st.a_rows[n] * Kt + k];
            TADEV_REQUIRE(ptrs[n], "tadev_summa_f64: A tile (%d,%d) has no data on this rank", st.a_rows[n], k);
          }
        }
      }
      // ---- B panel (travels along my grid column)
      {
        auto& ptrs = b_ptrs[wsi];
        ptrs.resize(st.b_cols.size(), nullptr);
        const int root = k % Pr;
        if (st.bcast_b) {
          const size_t elems = b_panel_elems(st);
          double* panel = buf + cursor;
          bool inplace = false;
          if (r == root) {
            inplace = true;
            const double* first = P.b_tiles[(size_t)k * Nt + st.b_cols[0]];
            size_t off = 0;
            for (size_t n = 0; n < st.b_cols.size(); ++n) {
              const double* tp = P.b_tiles[(size_t)k * Nt + st.b_cols[n]];
              TADEV_REQUIRE(tp, "tadev_summa_f64: B tile (%d,%d) is owned by this rank but has no data", k, st.b_cols[n]);
              if (tp != first + off || (reinterpret_cast<uintptr_t>(tp) & 15)) inplace = false;
              off += pad2((size_t)P.k_ext[k] * P.n_ext[st.b_cols[n]]);
            }
            if (inplace) panel = const_cast<double*>(first);
            else {
              size_t o = 0;
              for (size_t n = 0; n < st.b_cols.size(); ++n) {
                const size_t e = (size_t)P.k_ext[k] * P.n_ext[st.b_cols[n]];
                TADEV_CHECK_CUDA(cudaMemcpyAsync(panel + o, P.b_tiles[(size_t)k * Nt + st.b_cols[n]], e * 8,
                                                 cudaMemcpyDeviceToDevice, sc));
                o += pad2(e);
              }
            }
          }
          if (!inplace) cursor += elems;
          size_t o = 0;
          for (size_t n = 0; n < st.b_cols.size(); ++n) {
            ptrs[n] = panel + o;
            o += pad2((size_t)P.k_ext[k] * P.n_ext[st.b_cols[n]]);
          }
          col_bcasts.push_back({panel, elems * 8, root});
          bcast_bytes += (int64_t)elems * 8;
          any_comm = true;
        } else if (st.compute) {
          for (size_t n = 0; n < st.b_cols.size(); ++n) {
            ptrs[n] = P.b_tiles[(size_t)k * Nt + st.b_cols[n]];
            TADEV_REQUIRE(ptrs[n], "tadev_summa_f64: B tile (%d,%d) has no data on this rank", k, st.b_cols[n]);
          }
        }
      }
    }
    if (any_comm) {
      // same (comm, k) order on every rank of a group => no cross-communicator deadlock
      if (!row_bcasts.empty()) {
        TADEV_CHECK_NCCL(ncclGroupStart());
        for (auto& bc : row_bcasts) TADEV_CHECK_NCCL(ncclBroadcast(bc.ptr, bc.ptr, bc.bytes, ncclChar, bc.root, ctx->row_comm, sc));
        TADEV_CHECK_NCCL(ncclGroupEnd());
      }
      if (!col_bcasts.empty()) {
        TADEV_CHECK_NCCL(ncclGroupStart());
        for (auto& bc : col_bcasts) TADEV_CHECK_NCCL(ncclBroadcast(bc.ptr, bc.ptr, bc.bytes, ncclChar, bc.root, ctx->col_comm, sc));
        TADEV_CHECK_NCCL(ncclGroupEnd());
      }
      TADEV_CHECK_CUDA(cudaEventRecord(panel_ready[d], sc));
      TADEV_CHECK_CUDA(cudaStreamWaitEvent(s0, panel_ready[d], 0));
    }

    // ---- grouped GEMM descriptors of this window: chain contributions per result tile
    contrib.clear();
    for (size_t wsi = 0; wsi < win.steps.size(); ++wsi) {
      const SummaStep& st = S.steps[win.steps[wsi]];
      if (!st.compute) continue;
      // index of a global row/col inside this step's panel lists
      size_t ai = 0;
      for (int64_t pp = st.pair_begin; pp < st.pair_end; ++pp) {
        const int i = S.pair_i[pp], j = S.pair_j[pp];
        while (st.a_rows[ai] != i) ++ai;  // pairs are row-major: rows appear in a_rows order
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
          double* ct = P.c_tiles[contrib[n].key];
          TADEV_REQUIRE(ct, "tadev_summa_f64: result tile (%d,%d) is non-zero and local but has no storage", i, j);
          if (!groups.empty()) groups.back().task_end = (int32_t)tasks.size();
          tadev_gemm_group G{ct, (int32_t)P.m_ext[i], (int32_t)P.n_ext[j], (int32_t)tasks.size(), 0,
                             (P.accumulate || touched[contrib[n].key]) ? 1 : 0, 0};
          touched[contrib[n].key] = 1;
          groups.push_back(G);
        }
        tasks.push_back({contrib[n].A, contrib[n].B, contrib[n].k, 0});
      }
      groups.back().task_end = (int32_t)tasks.size();
      int rc = tadev_gemm_grouped_f64(ctx, s0, P.opA, P.opB, P.alpha, groups.data(), (int)groups.size(), tasks.data(),
                                      (int)tasks.size());
      if (rc) return rc;
      ++nlaunches;
    }
    if (need_ring && win.bytes > 0) {
      TADEV_CHECK_CUDA(cudaEventRecord(buf_free[d], s0));
      buf_used[d] = 1;
    }
  }

  // result tiles that are non-zero in the result shape but received no contribution
  if (!P.accumulate) {
    for (int i = r; i < Mt; i += Pr)
      for (int j = c; j < Nt; j += Pc) {
        const size_t key = (size_t)i * Nt + j;
        const bool nz = !P.c_norms || P.c_norms[key] >= P.threshold;
        if (nz && !touched[key] && P.c_tiles[key])
          TADEV_CHECK_CUDA(cudaMemsetAsync(P.c_tiles[key], 0, (size_t)P.m_ext[i] * P.n_ext[j] * 8, s0));
      }
  }
  if (need_ring) {
    TADEV_CHECK_CUDA(cudaEventRecord(ev_comm_done, sc));
    TADEV_CHECK_CUDA(cudaStreamWaitEvent(s0, ev_comm_done, 0));
  }
  TADEV_CHECK_CUDA(cudaEventRecord(ev_end, s0));
  if (need_ring)
    for (int d = 0; d < D; ++d) { int rc = tadev_free(ctx, ring[d], s0); if (rc) return rc; }
  TADEV_CHECK_CUDA(cudaEventSynchronize(ev_end));
  TADEV_CHECK_CUDA(cudaGetLastError());
  float ms = 0;
  TADEV_CHECK_CUDA(cudaEventElapsedTime(&ms, ev_start, ev_end));
  if (stats) {
    stats->nsteps = (int64_t)S.steps.size();
    stats->nsteps_skipped = S.nskipped;
    stats->npairs = npairs;
    stats->nlaunches = nlaunches;
    stats->flops = flops;
    stats->bcast_bytes = bcast_bytes;
    stats->device_ms = ms;
  }
  cudaEventDestroy(ev_start); cudaEventDestroy(ev_end); cudaEventDestroy(ev_comm_done);
  if (need_ring)
    for (int d = 0; d < D; ++d) { cudaEventDestroy(panel_ready[d]); cudaEventDestroy(buf_free[d]); }
  return TADEV_OK;
}
