// tilelist.h — device-side construction of the grouped-GEMM work lists of a SUMMA window (internal).
//
// Replaces Summa::contract (reference: src/TiledArray/dist_eval/contraction_eval.h:1311-1384): for every step k of
// a window and every local result tile (i,j) the pair (A(i,k), B(k,j)) is scheduled iff
//   a_norm(i,k) >= thr && b_norm(k,j) >= thr && c_norm(i,j) >= thr        (dense arrays: always)
// The host uploads only O(rows + cols) panel-tile tables per step; the O(pairs) enumeration, the per-result-tile
// chaining (ContractReduce groups), the first-touch beta flags and the L2-rasterised 128x128 work-item order are
// produced by kernels on the compute stream and consumed by the GEMM kernel directly from device memory.
#pragma once
#include "common.h"

// one staged operand tile of a panel table (nullptr ptr = absent)
struct TlTile {
  const double* ptr;
  const void* map;  // device CUtensorMap of a k-contiguous operand tile (fast path), else nullptr
};

struct TlCounters {       // device-resident, read back by the driver at the end of the contraction
  unsigned long long npairs;
  double flops;
  unsigned long long error; // sticky: a panel table lacked a tile the shapes call non-zero
  int32_t total_items;    // work items (128x128 blocks) of the current launch
  int32_t total_tasks;
  int32_t sched[2];       // the persistent kernel's work counter and wave-sync counter (zeroed per launch)
  int32_t total_prefix;   // == total_items (generic kernel's tile_prefix[ngroups])
  int32_t pad;
};

struct TileListBuilder {
  tadev_ctx* ctx = nullptr;
  cudaStream_t s = nullptr;
  int Pr = 1, Pc = 1, r = 0, c = 0, Mt = 0, Nt = 0, Kt = 0;
  int nrl = 0, ncl = 0;                 // local tile rows / cols of the result (i = r + li*Pr, j = c + lj*Pc)
  bool sparse = false, raster = false;
  float thr = 0.f;
  int accumulate = 0;
  // device arrays (owned; one allocation)
  char* d_base = nullptr;
  size_t d_bytes = 0;
  float *d_an = nullptr, *d_bn = nullptr, *d_cn = nullptr;
  int32_t *d_mloc = nullptr, *d_nloc = nullptr, *d_kext = nullptr;
  int32_t *d_brow0 = nullptr, *d_bcol0 = nullptr, *d_brow2li = nullptr, *d_bcol2lj = nullptr;
  double** d_cptr = nullptr;            // [nrl * ncl] result tile addresses (owned copy, refreshed per row block)
  uint8_t* d_touched = nullptr;         // [nrl * ncl]
  int32_t *d_cnt = nullptr, *d_tbegin = nullptr, *d_nblk = nullptr, *d_bprefix = nullptr;  // [ngroups + 1]
  int32_t *d_kflag = nullptr, *d_kpos = nullptr, *d_scan_tmp = nullptr;
  tadev_gemm_group* d_groups = nullptr;
  void* d_tasks = nullptr;              // TadevWsTask[] or tadev_gemm_task[]
  int2* d_items = nullptr;
  TlCounters* d_counters = nullptr;
  size_t cap_tasks = 0, cap_items = 0, cap_keys = 0;
  std::vector<int32_t> h_brow0, h_bcol0;  // block-row/col origin of every local tile row / col (+ total)
  int total_brows = 0, total_bcols = 0;

  // max_tasks / max_items: upper bounds over all windows of the contraction (host knows them from panel sizes)
  int init(tadev_ctx* ctx, cudaStream_t s, int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt, const int64_t* m_ext,
           const int64_t* n_ext, const int64_t* k_ext, const float* a_norms, const float* b_norms, const float* c_norms,
           float thr, int accumulate, size_t max_tasks);
  // Build the lists of one window and launch the GEMM. d_ksteps: [nws] global k of each step; d_atab: [nws][li1-li0],
  // d_btab: [nws][ncl] (device, part of a staged block); d_cptr_staged: result-tile addresses of local rows
  // [li0, li1) ([nrows][ncl], staged with the first window of a row block; nullptr = unchanged).
  // fast = every operand row is 16-byte aligned (TMA kernel).
  int build_and_launch(int opA, int opB, double alpha, int li0, int li1, int nws, const int32_t* d_ksteps,
                       const TlTile* d_atab, const TlTile* d_btab, double* const* d_cptr_staged, bool fast,
                       cudaEvent_t ev_list0, cudaEvent_t ev_list1, cudaEvent_t ev_gemm0, cudaEvent_t ev_gemm1);
  // zero-fill result tiles of rows [li0, li1) that no window touched (beta = 0 contractions only)
  int zero_untouched(int li0, int li1);
  int read_counters(unsigned long long* npairs, double* flops);  // synchronises the stream
  void destroy();
};

// fast-path launcher with device-resident lists (gemm_f64_ws.cu)
int launch_gemm_ws_devlists(tadev_ctx* ctx, cudaStream_t s, int opA, int opB, double alpha, const tadev_gemm_group* d_groups,
                            int ngroups, const void* d_wstasks, const int2* d_items, const int32_t* d_total,
                            int32_t* d_sched, int wave_sync);
// tensor maps of k-contiguous operand tiles: cached (persistent tiles) or encoded into caller memory (transient ones)
int tadev_ws_cached_map(tadev_ctx* ctx, const double* ptr, int outer, int k, const void** dev_map, bool* created);
int tadev_ws_flush_new_maps(tadev_ctx* ctx);  // upload maps created since the last flush (host-synchronous)
int tadev_ws_encode_map(tadev_ctx* ctx, void* h_dst128, const double* ptr, int outer, int k);
