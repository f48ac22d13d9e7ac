// tilelist.cu — device-side tile-pair lists of a SUMMA window (see tilelist.h).
//
// Reference semantics: Summa::contract (dist_eval/contraction_eval.h:1311-1384) — for step k, every non-zero
// A(i,k) of my column panel times every non-zero B(k,j) of my row panel whose result tile C(i,j) is non-zero is
// one ReducePairTask contribution (reduce_task.h:961-1026), accumulated per result tile in step order
// (contract_reduce.h:409-453). Here one thread owns one local result tile: it counts its contributions over all
// steps of the window, an exclusive scan assigns the task ranges, a second pass writes the chained tasks
// (ascending k: the accumulation order of the reference's sequential per-tile reduction) and the group record
// with its first-touch beta flag; finally the 128x128 work items of all active groups are emitted in the
// L2-rasterised order of gemm_f64_ws.cu by a flag + scan + scatter over the dense key space of the block grid.
#include "tilelist.h"

#include <algorithm>
#include <cmath>

namespace {

constexpr int SCAN_THREADS = 1024, SCAN_ITEMS = 4, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

// exclusive scan of one 4096-element block; block_sums[blockIdx.x] = block total. `in` and `out` may be the same
// array (the callers scan in place), hence no __restrict__: every element is read and later written by one thread.
__global__ void __launch_bounds__(SCAN_THREADS) tl_scan_block_kernel(const int32_t* in, int32_t* out, int32_t* block_sums, int64_t n) {
  __shared__ int32_t warp_sums[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (base + i < n) ? in[base + i] : 0;
  int32_t run = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { const int32_t t = v[i]; v[i] = run; run += t; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    const int32_t ws = warp_sums[lane];
    int32_t wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    warp_sums[lane] = wi - ws;
    if (lane == 31 && block_sums) block_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  const int32_t off = warp_sums[w] + inc - run;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) out[base + i] = v[i] + off;
}

__global__ void __launch_bounds__(SCAN_THREADS) tl_scan_add_kernel(int32_t* __restrict__ data, const int32_t* __restrict__ block_off, int64_t n) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  const int32_t off = block_off[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) data[base + i] += off;
}

// in-place exclusive scan of data[0..n); *total = sum. tmp: >= ceil(n / 4096) + 1 ints.
int tl_scan(tadev_ctx* ctx, cudaStream_t s, int32_t* data, int64_t n, int32_t* tmp, int32_t* total) {
  if (n <= 0) return TADEV_OK;
  const int64_t nb = ceil_div64(n, SCAN_BLOCK);
  TADEV_REQUIRE(nb <= SCAN_BLOCK, "device tile-list builder: %lld elements exceed the two-level scan", (long long)n);
  ctx->launches += nb == 1 ? 1 : 3;
  if (nb == 1) {
    tl_scan_block_kernel<<<1, SCAN_THREADS, 0, s>>>(data, data, total, n);
  } else {
    tl_scan_block_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(data, data, tmp, n);
    tl_scan_block_kernel<<<1, SCAN_THREADS, 0, s>>>(tmp, tmp, total, nb);
    tl_scan_add_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(data, tmp, n);
  }
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}

struct TlArgs {
  int nws, li0, nrows, ncl, Pr, Pc, r, c, Nt, Kt, accumulate, fast;
  float thr;
  const int32_t* ksteps;
  const float *an, *bn, *cn;
  const int32_t *mloc, *nloc, *kext;
  const TlTile *atab, *btab;
  double* const* cptr;     // [nrows * ncl]
  uint8_t* touched;        // [nrl * ncl] (global local index)
  int32_t *cnt, *nblk, *tbegin;
  tadev_gemm_group* groups;
  void* tasks;
  TlCounters* counters;
};

__device__ __forceinline__ bool tl_pair(const TlArgs& P, int i, int j, int k) {
  return (!P.an || P.an[(size_t)i * P.Kt + k] >= P.thr) && (!P.bn || P.bn[(size_t)k * P.Nt + j] >= P.thr);
}

// pass 1: contributions per local result tile
__global__ void __launch_bounds__(256) tl_count_kernel(const __grid_constant__ TlArgs P) {
  const int ng = P.nrows * P.ncl;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long np = 0;
  double fl = 0.0;
  if (g < ng) {
    const int lr = g / P.ncl, lj = g % P.ncl, li = P.li0 + lr;
    const int i = P.r + li * P.Pr, j = P.c + lj * P.Pc;
    int count = 0;
    long long ksum = 0;
    const bool shape_only = P.atab == nullptr;  // tadev_build_tile_lists: lists from the shapes alone
    const bool cnz = (!P.cn || P.cn[(size_t)i * P.Nt + j] >= P.thr) && (shape_only || P.cptr[g] != nullptr);
    if (cnz) {
      for (int w = 0; w < P.nws; ++w) {
        const int k = P.ksteps[w];
        if (!tl_pair(P, i, j, k)) continue;
        // the panel tables must hold every tile the shapes call non-zero
        if (!shape_only && (!P.atab[(size_t)w * P.nrows + lr].ptr || !P.btab[(size_t)w * P.ncl + lj].ptr)) {
          if (P.an || P.bn) P.counters->error = 1;  // sparse: inconsistent tables (reported by the driver)
          continue;                                 // dense: the tile is simply not part of this window's panels
        }
        ++count;
        ksum += P.kext[k];
      }
    }
    const int m = P.mloc[li], n = P.nloc[lj];
    P.cnt[g] = count;
    P.nblk[g] = count > 0 ? ((m + kGemmBM - 1) / kGemmBM) * ((n + kGemmBN - 1) / kGemmBN) : 0;
    np = (unsigned long long)count;
    fl = 2.0 * (double)m * (double)n * (double)ksum;
  }
  if (g == ng) { P.cnt[ng] = 0; P.nblk[ng] = 0; }  // scans run over ng + 1 entries (last = total)
  // block reduction of the statistics
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    np += __shfl_down_sync(0xffffffffu, np, o);
    fl += __shfl_down_sync(0xffffffffu, fl, o);
  }
  __shared__ unsigned long long s_np[8];
  __shared__ double s_fl[8];
  if ((threadIdx.x & 31) == 0) { s_np[threadIdx.x >> 5] = np; s_fl[threadIdx.x >> 5] = fl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { np += s_np[w]; fl += s_fl[w]; }
    if (np) { atomicAdd(&P.counters->npairs, np); atomicAdd(&P.counters->flops, fl); }
  }
}

// pass 2 of tadev_build_tile_lists: the contracted tile index k of every contribution, chained per result tile
__global__ void __launch_bounds__(256) tl_fill_k_kernel(const __grid_constant__ TlArgs P, int32_t* __restrict__ task_k, int64_t capacity) {
  const int ng = P.nrows * P.ncl;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng || P.cnt[g] == 0) return;
  const int lr = g / P.ncl, lj = g % P.ncl;
  const int i = P.r + (P.li0 + lr) * P.Pr, j = P.c + lj * P.Pc;
  int64_t t = P.tbegin[g];
  for (int w = 0; w < P.nws; ++w) {
    const int k = P.ksteps[w];
    if (!tl_pair(P, i, j, k)) continue;
    if (t < capacity) task_k[t] = k;
    ++t;
  }
}

// pass 2: group records + chained tasks (ascending step order)
__global__ void __launch_bounds__(256) tl_fill_kernel(const __grid_constant__ TlArgs P) {
  const int ng = P.nrows * P.ncl;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  const int lr = g / P.ncl, lj = g % P.ncl, li = P.li0 + lr;
  const int i = P.r + li * P.Pr, j = P.c + lj * P.Pc;
  const int count = P.cnt[g];  // (cnt was scanned into tbegin; cnt itself is untouched)
  int t = P.tbegin[g];
  const size_t tg = (size_t)li * P.ncl + lj;
  tadev_gemm_group G;
  G.C = P.cptr[g];
  G.m = P.mloc[li]; G.n = P.nloc[lj];
  G.task_begin = t; G.task_end = t + count;
  G.accumulate = (P.accumulate || P.touched[tg]) ? 1 : 0;
  G.raster = 0;
  P.groups[g] = G;
  if (count == 0) return;
  P.touched[tg] = 1;
  for (int w = 0; w < P.nws; ++w) {
    const int k = P.ksteps[w];
    if (!tl_pair(P, i, j, k)) continue;
    const TlTile a = P.atab[(size_t)w * P.nrows + lr], b = P.btab[(size_t)w * P.ncl + lj];
    if (!a.ptr || !b.ptr) continue;
    if (P.fast) {
      TadevWsTask T;
      T.A = a.ptr; T.B = b.ptr; T.k = P.kext[k]; T.pad = 0; T.mapA = a.map; T.mapB = b.map;
      static_cast<TadevWsTask*>(P.tasks)[t] = T;
    } else {
      tadev_gemm_task T;
      T.A = a.ptr; T.B = b.ptr; T.k = P.kext[k]; T.reserved = 0;
      static_cast<tadev_gemm_task*>(P.tasks)[t] = T;
    }
    ++t;
  }
}

struct TlItemArgs {
  int S, rows_blk, cols_blk, brow_origin, li0, ncl;
  int64_t nkeys;
  const int32_t *brow2li, *bcol2lj, *brow0, *bcol0, *cnt;
  int32_t *kflag;
  int2* items;
};

// key q of the rasterised order = (band of S block rows, global block column, row inside the band)
__device__ __forceinline__ bool tl_key_decode(const TlItemArgs& P, int64_t q, int& g, int& a, int& b) {
  const int64_t per_band = (int64_t)P.cols_blk * P.S;
  const int band = (int)(q / per_band);
  const int rem = (int)(q % per_band);
  const int gc = rem / P.S, gr = band * P.S + rem % P.S;
  if (gr >= P.rows_blk) return false;
  const int li = P.brow2li[P.brow_origin + gr], lj = P.bcol2lj[gc];
  g = (li - P.li0) * P.ncl + lj;
  a = P.brow_origin + gr - P.brow0[li];
  b = gc - P.bcol0[lj];
  return true;
}

__global__ void __launch_bounds__(256) tl_keyflag_kernel(const __grid_constant__ TlItemArgs P) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q > P.nkeys) return;
  int g, a, b, f = 0;
  if (q < P.nkeys && tl_key_decode(P, q, g, a, b)) f = P.cnt[g] > 0 ? 1 : 0;
  P.kflag[q] = f;  // entry nkeys = 0: the scan's last element is the total
}

__global__ void __launch_bounds__(256) tl_items_kernel(const __grid_constant__ TlItemArgs P, const int32_t* __restrict__ kpos) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P.nkeys) return;
  int g, a, b;
  if (!tl_key_decode(P, q, g, a, b) || P.cnt[g] <= 0) return;
  P.items[kpos[q]] = make_int2(g, (a << 16) | b);
}

// group-major item order (no raster hints): block t of group g sits at bprefix[g] + t
__global__ void __launch_bounds__(256) tl_items_groupmajor_kernel(int ng, const int32_t* __restrict__ nblk, const int32_t* __restrict__ bprefix,
                                                                  const tadev_gemm_group* __restrict__ groups, int2* __restrict__ items) {
  const int g = blockIdx.x;
  if (g >= ng) return;
  const int nb = nblk[g];
  if (nb == 0) return;
  const int tn = (groups[g].n + kGemmBN - 1) / kGemmBN, base = bprefix[g];
  for (int t = threadIdx.x; t < nb; t += blockDim.x) items[base + t] = make_int2(g, ((t / tn) << 16) | (t % tn));
}

__global__ void tl_copy_total_kernel(TlCounters* c, const int32_t* total_prefix) { c->total_items = *total_prefix; }

__global__ void __launch_bounds__(256) tl_zero_untouched_kernel(int li0, int nrows, int ncl, int Pr, int Pc, int r, int c, int Nt, float thr,
                                                               const float* __restrict__ cn, double* const* __restrict__ cptr,
                                                               const uint8_t* __restrict__ touched, const int32_t* __restrict__ mloc,
                                                               const int32_t* __restrict__ nloc) {
  const int g = blockIdx.x;
  if (g >= nrows * ncl) return;
  const int lr = g / ncl, lj = g % ncl, li = li0 + lr;
  if (touched[(size_t)li * ncl + lj]) return;
  const int i = r + li * Pr, j = c + lj * Pc;
  if (cn && cn[(size_t)i * Nt + j] < thr) return;
  double* C = cptr[g];
  if (!C) return;
  const size_t n = (size_t)mloc[li] * nloc[lj];
  for (size_t x = threadIdx.x; x < n; x += blockDim.x) C[x] = 0.0;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

int TileListBuilder::init(tadev_ctx* ctx_, cudaStream_t s_, int Pr_, int Pc_, int r_, int c_, int Mt_, int Nt_, int Kt_,
                          const int64_t* m_ext, const int64_t* n_ext, const int64_t* k_ext, const float* a_norms,
                          const float* b_norms, const float* c_norms, float thr_, int accumulate_, size_t max_tasks) {
  ctx = ctx_; s = s_; Pr = Pr_; Pc = Pc_; r = r_; c = c_; Mt = Mt_; Nt = Nt_; Kt = Kt_; thr = thr_; accumulate = accumulate_;
  nrl = r < Mt ? (Mt - r + Pr - 1) / Pr : 0;
  ncl = c < Nt ? (Nt - c + Pc - 1) / Pc : 0;
  sparse = a_norms || b_norms || c_norms;
  const size_t ng = (size_t)nrl * ncl;
  std::vector<int32_t> mloc(nrl), nloc(ncl), kext(Kt);
  h_brow0.assign(nrl + 1, 0); h_bcol0.assign(ncl + 1, 0);
  for (int li = 0; li < nrl; ++li) {
    TADEV_REQUIRE(m_ext[r + li * Pr] < (1ll << 31), "tile extent too large");
    mloc[li] = (int32_t)m_ext[r + li * Pr];
    h_brow0[li + 1] = h_brow0[li] + (int32_t)ceil_div64(mloc[li], kGemmBM);
  }
  for (int lj = 0; lj < ncl; ++lj) {
    TADEV_REQUIRE(n_ext[c + lj * Pc] < (1ll << 31), "tile extent too large");
    nloc[lj] = (int32_t)n_ext[c + lj * Pc];
    h_bcol0[lj + 1] = h_bcol0[lj] + (int32_t)ceil_div64(nloc[lj], kGemmBN);
  }
  for (int k = 0; k < Kt; ++k) { TADEV_REQUIRE(k_ext[k] < (1ll << 31), "tile extent too large"); kext[k] = (int32_t)k_ext[k]; }
  total_brows = h_brow0[nrl]; total_bcols = h_bcol0[ncl];
  std::vector<int32_t> brow2li(std::max(total_brows, 1)), bcol2lj(std::max(total_bcols, 1));
  for (int li = 0; li < nrl; ++li) for (int x = h_brow0[li]; x < h_brow0[li + 1]; ++x) brow2li[x] = li;
  for (int lj = 0; lj < ncl; ++lj) for (int x = h_bcol0[lj]; x < h_bcol0[lj + 1]; ++x) bcol2lj[x] = lj;
  int grid = ctx->num_sms - ctx->gemm_sm_reserve;
  if (grid < 1) grid = 1;
  int S = 1;
  while ((S + 1) * (S + 1) <= grid) ++S;
  if (const char* e = getenv("TADEV_RASTER_S")) { if (atoi(e) > 0) S = atoi(e); }
  const bool raster_off = getenv("TADEV_RASTER_S") && atoi(getenv("TADEV_RASTER_S")) == 0;
  const int64_t nkeys = (int64_t)ceil_div64(std::max(total_brows, 1), S) * S * std::max(total_bcols, 1);
  // blocks of a tile are encoded (a << 16) | b in a work item
  raster = !raster_off && ng > 1 && total_brows < 65535 && total_bcols < 65535 && nkeys + 1 <= (int64_t)SCAN_BLOCK * SCAN_BLOCK;
  cap_keys = raster ? (size_t)nkeys + 1 : 0;
  cap_items = (size_t)total_brows * total_bcols;
  cap_tasks = std::max<size_t>(max_tasks, 1);
  TADEV_REQUIRE(ng + 1 <= (size_t)SCAN_BLOCK * SCAN_BLOCK && cap_items < (size_t)1 << 31 && cap_tasks < (size_t)1 << 31,
                "device tile-list builder: tile grid too large");

  // one device allocation, carved up
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_an = carve(a_norms ? (size_t)Mt * Kt * 4 : 0), o_bn = carve(b_norms ? (size_t)Kt * Nt * 4 : 0), o_cn = carve(c_norms ? (size_t)Mt * Nt * 4 : 0);
  const size_t o_ml = carve((size_t)nrl * 4), o_nl = carve((size_t)ncl * 4), o_ke = carve((size_t)Kt * 4);
  const size_t o_br0 = carve((size_t)(nrl + 1) * 4), o_bc0 = carve((size_t)(ncl + 1) * 4);
  const size_t o_b2l = carve(brow2li.size() * 4), o_c2l = carve(bcol2lj.size() * 4);
  const size_t host_part = off;  // everything above is initialised from the host
  const size_t o_touch = carve(std::max<size_t>(ng, 1));
  const size_t o_cp = carve(std::max<size_t>(ng, 1) * 8);
  const size_t o_cnt = carve((ng + 1) * 4), o_tb = carve((ng + 1) * 4), o_nb = carve((ng + 1) * 4), o_bp = carve((ng + 1) * 4);
  const size_t o_kf = carve(cap_keys * 4), o_kp = carve(cap_keys * 4), o_st = carve((size_t)(SCAN_BLOCK + 2) * 4);
  const size_t o_gr = carve(std::max<size_t>(ng, 1) * sizeof(tadev_gemm_group));
  const size_t o_ta = carve(cap_tasks * sizeof(TadevWsTask));
  const size_t o_it = carve(std::max<size_t>(cap_items, 1) * sizeof(int2));
  const size_t o_ct = carve(sizeof(TlCounters));
  d_bytes = off;
  int rc = tadev_alloc(ctx, d_bytes, (void**)&d_base, (tadev_stream)s);
  if (rc) return rc;
  std::vector<char> h(host_part, 0);
  if (a_norms) memcpy(h.data() + o_an, a_norms, (size_t)Mt * Kt * 4);
  if (b_norms) memcpy(h.data() + o_bn, b_norms, (size_t)Kt * Nt * 4);
  if (c_norms) memcpy(h.data() + o_cn, c_norms, (size_t)Mt * Nt * 4);
  memcpy(h.data() + o_ml, mloc.data(), (size_t)nrl * 4);
  memcpy(h.data() + o_nl, nloc.data(), (size_t)ncl * 4);
  memcpy(h.data() + o_ke, kext.data(), (size_t)Kt * 4);
  memcpy(h.data() + o_br0, h_brow0.data(), (size_t)(nrl + 1) * 4);
  memcpy(h.data() + o_bc0, h_bcol0.data(), (size_t)(ncl + 1) * 4);
  memcpy(h.data() + o_b2l, brow2li.data(), brow2li.size() * 4);
  memcpy(h.data() + o_c2l, bcol2lj.data(), bcol2lj.size() * 4);
  // pageable source: the runtime stages the copy before returning, `h` may go out of scope afterwards
  if (host_part) TADEV_CHECK_CUDA(cudaMemcpyAsync(d_base, h.data(), host_part, cudaMemcpyHostToDevice, s));
  TADEV_CHECK_CUDA(cudaMemsetAsync(d_base + o_touch, 0, std::max<size_t>(ng, 1), s));
  TADEV_CHECK_CUDA(cudaMemsetAsync(d_base + o_ct, 0, sizeof(TlCounters), s));
  d_an = a_norms ? (float*)(d_base + o_an) : nullptr;
  d_bn = b_norms ? (float*)(d_base + o_bn) : nullptr;
  d_cn = c_norms ? (float*)(d_base + o_cn) : nullptr;
  d_mloc = (int32_t*)(d_base + o_ml); d_nloc = (int32_t*)(d_base + o_nl); d_kext = (int32_t*)(d_base + o_ke);
  d_brow0 = (int32_t*)(d_base + o_br0); d_bcol0 = (int32_t*)(d_base + o_bc0);
  d_brow2li = (int32_t*)(d_base + o_b2l); d_bcol2lj = (int32_t*)(d_base + o_c2l);
  d_touched = (uint8_t*)(d_base + o_touch);
  d_cptr = (double**)(d_base + o_cp);
  d_cnt = (int32_t*)(d_base + o_cnt); d_tbegin = (int32_t*)(d_base + o_tb); d_nblk = (int32_t*)(d_base + o_nb); d_bprefix = (int32_t*)(d_base + o_bp);
  d_kflag = (int32_t*)(d_base + o_kf); d_kpos = (int32_t*)(d_base + o_kp); d_scan_tmp = (int32_t*)(d_base + o_st);
  d_groups = (tadev_gemm_group*)(d_base + o_gr);
  d_tasks = d_base + o_ta;
  d_items = (int2*)(d_base + o_it);
  d_counters = (TlCounters*)(d_base + o_ct);
  return TADEV_OK;
}

void TileListBuilder::destroy() {
  if (d_base) tadev_free(ctx, d_base, (tadev_stream)s);
  d_base = nullptr;
}

int TileListBuilder::build_and_launch(int opA, int opB, double alpha, int li0, int li1, int nws, const int32_t* d_ksteps,
                                      const TlTile* d_atab, const TlTile* d_btab, double* const* d_cptr_staged, bool fast,
                                      cudaEvent_t ev_list0, cudaEvent_t ev_list1, cudaEvent_t ev_gemm0, cudaEvent_t ev_gemm1) {
  const int nrows = li1 - li0;
  const int ng = nrows * ncl;
  if (ng <= 0 || nws <= 0) return TADEV_OK;
  if (ev_list0) TADEV_CHECK_CUDA(cudaEventRecord(ev_list0, s));
  if (d_cptr_staged)
    TADEV_CHECK_CUDA(cudaMemcpyAsync(d_cptr + (size_t)li0 * ncl, d_cptr_staged, (size_t)ng * 8, cudaMemcpyDeviceToDevice, s));
  // per-launch counters (everything after npairs/flops)
  TADEV_CHECK_CUDA(cudaMemsetAsync(&d_counters->total_items, 0, sizeof(TlCounters) - offsetof(TlCounters, total_items), s));
  TlArgs A{};
  A.nws = nws; A.li0 = li0; A.nrows = nrows; A.ncl = ncl; A.Pr = Pr; A.Pc = Pc; A.r = r; A.c = c; A.Nt = Nt; A.Kt = Kt;
  A.accumulate = accumulate; A.fast = fast ? 1 : 0; A.thr = thr;
  A.ksteps = d_ksteps; A.an = d_an; A.bn = d_bn; A.cn = d_cn; A.mloc = d_mloc; A.nloc = d_nloc; A.kext = d_kext;
  A.atab = d_atab; A.btab = d_btab; A.cptr = d_cptr + (size_t)li0 * ncl; A.touched = d_touched;
  A.cnt = d_cnt; A.nblk = d_nblk; A.tbegin = d_tbegin; A.groups = d_groups; A.tasks = d_tasks; A.counters = d_counters;
  const unsigned gb = (unsigned)ceil_div64((int64_t)ng + 1, 256);
  tl_count_kernel<<<gb, 256, 0, s>>>(A);
  TADEV_CHECK_CUDA(cudaMemcpyAsync(d_tbegin, d_cnt, (size_t)(ng + 1) * 4, cudaMemcpyDeviceToDevice, s));
  TADEV_CHECK_CUDA(cudaMemcpyAsync(d_bprefix, d_nblk, (size_t)(ng + 1) * 4, cudaMemcpyDeviceToDevice, s));
  int rc = tl_scan(ctx, s, d_tbegin, (int64_t)ng + 1, d_scan_tmp, &d_counters->total_tasks);
  if (!rc) rc = tl_scan(ctx, s, d_bprefix, (int64_t)ng + 1, d_scan_tmp, &d_counters->total_prefix);
  if (rc) return rc;
  tl_fill_kernel<<<gb, 256, 0, s>>>(A);
  ctx->launches += 2;
  const int rows_blk = h_brow0[li1] - h_brow0[li0];
  const int64_t max_items = (int64_t)rows_blk * total_bcols;
  if (raster && fast) {
    int grid = ctx->num_sms - ctx->gemm_sm_reserve;
    if (grid < 1) grid = 1;
    int S = 1;
    while ((S + 1) * (S + 1) <= grid) ++S;
    if (const char* e = getenv("TADEV_RASTER_S")) { if (atoi(e) > 0) S = atoi(e); }
    TlItemArgs I{};
    I.S = S; I.rows_blk = rows_blk; I.cols_blk = total_bcols; I.brow_origin = h_brow0[li0]; I.li0 = li0; I.ncl = ncl;
    I.nkeys = (int64_t)ceil_div64(std::max(rows_blk, 1), S) * S * std::max(total_bcols, 1);
    TADEV_REQUIRE((size_t)I.nkeys + 1 <= cap_keys, "device tile-list builder: key space exceeds its allocation");
    I.brow2li = d_brow2li; I.bcol2lj = d_bcol2lj; I.brow0 = d_brow0; I.bcol0 = d_bcol0; I.cnt = d_cnt; I.kflag = d_kflag; I.items = d_items;
    const unsigned kb = (unsigned)ceil_div64(I.nkeys + 1, 256);
    tl_keyflag_kernel<<<kb, 256, 0, s>>>(I);
    TADEV_CHECK_CUDA(cudaMemcpyAsync(d_kpos, d_kflag, (size_t)(I.nkeys + 1) * 4, cudaMemcpyDeviceToDevice, s));
    rc = tl_scan(ctx, s, d_kpos, I.nkeys + 1, d_scan_tmp, &d_counters->total_items);
    if (rc) return rc;
    tl_items_kernel<<<kb, 256, 0, s>>>(I, d_kpos);
    ctx->launches += 2;
  } else if (fast) {
    tl_items_groupmajor_kernel<<<ng, 256, 0, s>>>(ng, d_nblk, d_bprefix, d_groups, d_items);
    tl_copy_total_kernel<<<1, 1, 0, s>>>(d_counters, &d_counters->total_prefix);
    ctx->launches += 2;
  }
  TADEV_CHECK_CUDA(cudaGetLastError());
  if (ev_list1) TADEV_CHECK_CUDA(cudaEventRecord(ev_list1, s));
  GemmTimingHook& hook = tadev_gemm_timing_hook();
  hook.before = ev_gemm0; hook.after = ev_gemm1;
  if (fast) {
    // wave re-alignment pays off when every work item takes the same time (dense: all result tiles see the same K)
    static const int wave_sync_env = getenv("TADEV_WAVE_SYNC") ? atoi(getenv("TADEV_WAVE_SYNC")) : -1;
    const int wave_sync = wave_sync_env >= 0 ? wave_sync_env : ((!sparse && raster) ? 1 : 0);
    rc = launch_gemm_ws_devlists(ctx, s, opA, opB, alpha, d_groups, ng, d_tasks, d_items, &d_counters->total_items,
                                 d_counters->sched, wave_sync);
  } else {
    rc = launch_gemm_grouped_f64(ctx, s, opA, opB, alpha, d_groups, ng, static_cast<const tadev_gemm_task*>(d_tasks), d_bprefix,
                                 (int)std::max<int64_t>(max_items, 1), false);
  }
  hook.before = hook.after = nullptr;
  return rc;
}

int TileListBuilder::zero_untouched(int li0, int li1) {
  const int ng = (li1 - li0) * ncl;
  if (ng <= 0 || accumulate) return TADEV_OK;
  tl_zero_untouched_kernel<<<ng, 256, 0, s>>>(li0, li1 - li0, ncl, Pr, Pc, r, c, Nt, thr, d_cn, d_cptr + (size_t)li0 * ncl, d_touched,
                                              d_mloc, d_nloc);
  TADEV_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return TADEV_OK;
}

int TileListBuilder::read_counters(unsigned long long* npairs, double* flops) {
  TlCounters hc;
  TADEV_CHECK_CUDA(cudaMemcpyAsync(&hc, d_counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
  TADEV_CHECK_CUDA(cudaStreamSynchronize(s));
  *npairs = hc.npairs; *flops = hc.flops;
  TADEV_REQUIRE(hc.error == 0, "device tile-list builder: a panel table lacks a tile the shapes call non-zero");
  return TADEV_OK;
}

// The list kernel of the SUMMA driver as a stand-alone entry (tests, callers that want the lists): multi-step,
// multi-CTA successor of tadev_build_pairlist.
extern "C" int tadev_build_tile_lists(tadev_ctx* ctx, tadev_stream s_, int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt,
                                      const int32_t* d_ksteps, int nsteps, const float* d_a, const float* d_b, const float* d_c,
                                      float threshold, int32_t* d_group_begin, int32_t* d_task_k, int64_t capacity,
                                      int32_t* d_ntasks) {
  TADEV_REQUIRE(ctx, "tadev_build_tile_lists: null ctx");
  TADEV_REQUIRE(Pr >= 1 && Pc >= 1 && r >= 0 && r < Pr && c >= 0 && c < Pc, "tadev_build_tile_lists: bad grid position");
  TADEV_REQUIRE(Mt >= 0 && Nt >= 0 && Kt >= 0 && nsteps >= 0 && capacity >= 0, "tadev_build_tile_lists: bad extents");
  TADEV_REQUIRE(d_group_begin && d_ntasks && (nsteps == 0 || d_ksteps) && (capacity == 0 || d_task_k), "tadev_build_tile_lists: null arrays");
  cudaStream_t s = (cudaStream_t)s_;
  const int nrl = r < Mt ? (Mt - r + Pr - 1) / Pr : 0, ncl = c < Nt ? (Nt - c + Pc - 1) / Pc : 0;
  const int64_t ng64 = (int64_t)nrl * ncl;
  TADEV_REQUIRE(ng64 + 1 <= (int64_t)SCAN_BLOCK * SCAN_BLOCK, "tadev_build_tile_lists: tile grid too large");
  const int ng = (int)ng64;
  // scratch: cnt[ng+1], nblk[ng+1], mloc[nrl], nloc[ncl], kext[Kt] (unit extents: only counts matter), scan tmp, counters
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_cnt = carve((size_t)(ng + 1) * 4), o_nb = carve((size_t)(ng + 1) * 4), o_ml = carve((size_t)std::max(nrl, 1) * 4),
               o_nl = carve((size_t)std::max(ncl, 1) * 4), o_ke = carve((size_t)std::max(Kt, 1) * 4), o_st = carve((size_t)(SCAN_BLOCK + 2) * 4),
               o_ct = carve(sizeof(TlCounters));
  char* base = nullptr;
  int rc = tadev_alloc(ctx, off, (void**)&base, s_);
  if (rc) return rc;
  TADEV_CHECK_CUDA(cudaMemsetAsync(base, 0, off, s));
  TlArgs A{};
  A.nws = nsteps; A.li0 = 0; A.nrows = nrl; A.ncl = ncl; A.Pr = Pr; A.Pc = Pc; A.r = r; A.c = c; A.Nt = Nt; A.Kt = Kt; A.thr = threshold;
  A.ksteps = d_ksteps; A.an = d_a; A.bn = d_b; A.cn = d_c;
  A.mloc = (int32_t*)(base + o_ml); A.nloc = (int32_t*)(base + o_nl); A.kext = (int32_t*)(base + o_ke);
  A.cnt = (int32_t*)(base + o_cnt); A.nblk = (int32_t*)(base + o_nb); A.tbegin = d_group_begin;
  A.counters = (TlCounters*)(base + o_ct);
  const unsigned gb = (unsigned)ceil_div64((int64_t)ng + 1, 256);
  tl_count_kernel<<<gb, 256, 0, s>>>(A);
  TADEV_CHECK_CUDA(cudaMemcpyAsync(d_group_begin, A.cnt, (size_t)(ng + 1) * 4, cudaMemcpyDeviceToDevice, s));
  rc = tl_scan(ctx, s, d_group_begin, (int64_t)ng + 1, (int32_t*)(base + o_st), d_ntasks);
  if (!rc && ng > 0) tl_fill_k_kernel<<<gb, 256, 0, s>>>(A, d_task_k, capacity);
  ctx->launches += 2;
  cudaError_t e = cudaGetLastError();
  tadev_free(ctx, base, s_);
  TADEV_CHECK_CUDA(e);
  return rc;
}
