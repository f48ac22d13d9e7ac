// cont_engine.cpp — the contraction engine above the SUMMA driver: everything TiledArray's
// expression layer decides between `c("m,n") = a("m,k") * b("k,n")` and the Summa evaluator.
//
// Restates (for device-resident / host-resident / lazy arrays described through the C ABI):
//   ContEngine::perm_indices + init_indices_ (expressions/cont_engine.h:176-352,
//       binary_engine.h:101-178, permopt.h:254-376)         -> tadev_plan_contraction(_opt)
//   ContEngine::init_struct  (cont_engine.h:354-529)         -> op flags, permuted operand structure,
//       result trange (make_trange :593-637) and result shape (make_shape :642-660 ->
//       SparseShape::gemm, sparse_shape.h:1589-1691, evaluated by the device screening kernel)
//   ContEngine::init_distribution (cont_engine.h:537-587)    -> ProcGrid check, cyclic result map
//   ContEngine::make_dist_eval (cont_engine.h:662-677)       -> tadev_summa_f64
//   argument-tile permutation (dist_eval/array_eval.h:42,170) -> up-front batched permutes or the
//       just-in-time permute provider; result-tile permutation (contract_reduce.h:370-378)
// Host code only: every tile operation is a call into the kernels of this library.
#include <algorithm>
#include <cmath>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "common.h"

namespace {

struct TRange {  // TiledRange: per-dimension tile boundaries
  std::vector<std::vector<int64_t>> b;
  int rank() const { return (int)b.size(); }
  int64_t ntiles(int d) const { return (int64_t)b[d].size() - 1; }
  int64_t ext(int d, int64_t t) const { return b[d][t + 1] - b[d][t]; }
  std::vector<int64_t> tiles_shape() const {
    std::vector<int64_t> s(rank());
    for (int d = 0; d < rank(); ++d) s[d] = ntiles(d);
    return s;
  }
  int64_t total() const {
    int64_t n = 1;
    for (int d = 0; d < rank(); ++d) n *= ntiles(d);
    return n;
  }
};

int64_t ravel(const std::vector<int64_t>& idx, const std::vector<int64_t>& shape) {
  int64_t o = 0;
  for (size_t d = 0; d < shape.size(); ++d) o = o * shape[d] + idx[d];
  return o;
}
void unravel(int64_t o, const std::vector<int64_t>& shape, std::vector<int64_t>& idx) {
  idx.resize(shape.size());
  for (int d = (int)shape.size() - 1; d >= 0; --d) { idx[d] = o % shape[d]; o /= shape[d]; }
}

// out[perm(idx)] = in[idx] over a small row-major float tensor (SparseShape::perm, sparse_shape.h:1222)
std::vector<float> permute_norms(const std::vector<float>& in, const std::vector<int64_t>& shape, const std::vector<int>& perm) {
  const int R = (int)shape.size();
  std::vector<int64_t> oshape(R), idx, oidx(R);
  for (int i = 0; i < R; ++i) oshape[perm[i]] = shape[i];
  std::vector<float> out(in.size());
  for (int64_t o = 0; o < (int64_t)in.size(); ++o) {
    unravel(o, shape, idx);
    for (int i = 0; i < R; ++i) oidx[perm[i]] = idx[i];
    out[ravel(oidx, oshape)] = in[o];
  }
  return out;
}

// recursive_outer_product of tile-extent vectors in fp32 (sparse_shape.h:105-132): the association
// order is part of the bit-exact screening spec.
std::vector<float> recursive_outer(const TRange& tr, int lo, int hi) {
  const int dim = hi - lo;
  if (dim == 1) {
    std::vector<float> v((size_t)tr.ntiles(lo));
    for (int64_t t = 0; t < tr.ntiles(lo); ++t) v[t] = (float)tr.ext(lo, t);
    return v;
  }
  const int middle = (dim >> 1) + (dim & 1);
  const std::vector<float> l = recursive_outer(tr, lo, lo + middle), r = recursive_outer(tr, lo + middle, hi);
  std::vector<float> out(l.size() * r.size());
  for (size_t i = 0; i < l.size(); ++i)
    for (size_t j = 0; j < r.size(); ++j) out[i * r.size() + j] = l[i] * r[j];
  return out;
}

// fused (row-major) element extents of the tile grid spanned by dims [lo, hi)
std::vector<int64_t> fused_ext(const TRange& tr, int lo, int hi) {
  std::vector<int64_t> e(1, 1);
  for (int d = lo; d < hi; ++d) {
    std::vector<int64_t> n;
    n.reserve(e.size() * tr.ntiles(d));
    for (int64_t x : e)
      for (int64_t t = 0; t < tr.ntiles(d); ++t) n.push_back(x * tr.ext(d, t));
    e.swap(n);
  }
  return e;
}

struct Operand {
  tadev_array_desc d{};               // as given (pointers into the copies below)
  TRange tr;                          // original tiling
  std::vector<int64_t> bounds_flat;
  std::vector<int32_t> ntiles;
  std::vector<float> norms;           // empty = dense
  std::vector<const void*> tiles;     // empty: lazy array whose non-zero tiles are all local
  std::vector<int32_t> owners;        // the array's process map (empty: unknown)
  double* redist_arena = nullptr;     // tiles received by the redistribution (live during eval only)
  std::vector<const void*> tiles_before;  // the caller's tile table, restored after eval
  bool redistributed = false;
  std::vector<int> perm;              // explicit permutation (empty = none)
  TRange ptr;                         // permuted tiling
  std::vector<float> pnorms;          // permuted norms
  bool dense() const { return norms.empty(); }
};

int copy_operand(const tadev_array_desc* a, Operand& o, const char* who, bool need_tiles = true) {
  TADEV_REQUIRE(a && a->rank >= 0 && a->rank <= 16, "%s: bad array descriptor", who);
  TADEV_REQUIRE(a->rank == 0 || (a->bounds && a->ntiles), "%s: null tiling", who);
  TADEV_REQUIRE(a->memory >= 0 && a->memory <= 2, "%s: bad memory kind %d", who, a->memory);
  o.d = *a;
  size_t off = 0;
  o.tr.b.resize(a->rank);
  for (int d = 0; d < a->rank; ++d) {
    TADEV_REQUIRE(a->ntiles[d] >= 1, "%s: dimension %d has no tiles", who, d);
    o.tr.b[d].assign(a->bounds + off, a->bounds + off + a->ntiles[d] + 1);
    for (int t = 0; t < a->ntiles[d]; ++t) TADEV_REQUIRE(o.tr.b[d][t + 1] > o.tr.b[d][t], "%s: tile boundaries must increase", who);
    off += (size_t)a->ntiles[d] + 1;
  }
  const int64_t n = o.tr.total();
  if (a->norms) o.norms.assign(a->norms, a->norms + n);
  if (a->tiles) o.tiles.assign(a->tiles, a->tiles + n);
  if (a->owners) o.owners.assign(a->owners, a->owners + n);
  TADEV_REQUIRE(!need_tiles || a->memory == TADEV_MEM_LAZY || a->tiles, "%s: null tile table", who);
  return TADEV_OK;
}

}  // namespace

namespace {
std::vector<std::string> split_idx(const char* s);
}

struct tadev_contraction {
  tadev_ctx* ctx = nullptr;
  tadev_contract_options opt{};
  tadev_contraction_plan plan{};
  int32_t swapped = 0;
  double factor = 1.0;
  Operand L, R;  // after the optional exchange: L is the GEMM's left operand
  int lo[2], li[2], ro[2], ri[2];
  std::vector<int64_t> m_ext, n_ext, k_ext;
  int Mt = 0, Nt = 0, Kt = 0;
  // general (fused + contracted + free indices) products: the leading `nh` modes of both operands and
  // of the result are fused (batch) modes; Ht fused tile slabs, h_ext[h] batch elements in slab h
  bool general = false;
  int nh = 0, Ht = 1;
  std::vector<int64_t> h_ext;
  TRange tr_gemm, tr_target;
  std::vector<int> perm_res;             // GEMM order -> target order (empty = identity)
  bool sparse = false;
  std::vector<float> a_n, b_n, c_n;      // 2-D scaled norms in GEMM orientation [Mt,Kt] [Kt,Nt] [Mt,Nt]
  std::vector<float> target_norms;       // result shape in target order
  uint64_t nzero = 0;
  int Pr = 1, Pc = 1, r = 0, c = 0;
  // local non-zero result tiles, in GEMM (i,j) row-major order
  std::vector<int64_t> keys, elems, offs, target_ord;
  int64_t arena_elems = 0;
  // flattened target tiling for the info struct
  std::vector<int64_t> tb_flat;
  std::vector<int32_t> tn;
};

static int build_operand_structure(Operand& o, const int32_t* perm_arr, int rank) {
  o.perm.clear();
  if (perm_arr[0] >= 0) o.perm.assign(perm_arr, perm_arr + rank);
  if (o.perm.empty()) { o.ptr = o.tr; o.pnorms = o.norms; return TADEV_OK; }
  o.ptr.b.resize(rank);
  for (int i = 0; i < rank; ++i) o.ptr.b[o.perm[i]] = o.tr.b[i];
  if (!o.dense()) o.pnorms = permute_norms(o.norms, o.tr.tiles_shape(), o.perm);
  return TADEV_OK;
}

// SparseShape::gemm on the device: out[Mt,Nt] = |factor| * (a .* ksz) (b .* ksz), hard-zero + count
// followed (optionally) by SparseShape::mask (sparse_shape.h:653-676) with a user mask given in the same order
static int device_shape_gemm(tadev_ctx* ctx, int Mt, int Nt, int Kt, const std::vector<float>& a, const std::vector<float>& b,
                             const std::vector<float>& ksz, float abs_factor, float thr, std::vector<float>& out, uint64_t* nzero,
                             const float* mask = nullptr, float mask_thr = 0.0f) {
  tadev_stream s = (tadev_stream)ctx->streams[0];
  float *d_a = nullptr, *d_b = nullptr, *d_k = nullptr, *d_o = nullptr, *d_m = nullptr;
  uint64_t* d_z = nullptr;
  const size_t na = a.size() * 4, nb = b.size() * 4, nk = ksz.size() * 4, no = (size_t)Mt * Nt * 4;
  int rc = tadev_alloc(ctx, std::max<size_t>(na, 4), (void**)&d_a, s);
  if (!rc) rc = tadev_alloc(ctx, std::max<size_t>(nb, 4), (void**)&d_b, s);
  if (!rc) rc = tadev_alloc(ctx, std::max<size_t>(nk, 4), (void**)&d_k, s);
  if (!rc) rc = tadev_alloc(ctx, std::max<size_t>(no, 4), (void**)&d_o, s);
  if (!rc) rc = tadev_alloc(ctx, 8, (void**)&d_z, s);
  if (!rc) rc = tadev_memcpy_h2d(ctx, d_a, a.data(), na, s);
  if (!rc) rc = tadev_memcpy_h2d(ctx, d_b, b.data(), nb, s);
  if (!rc && nk) rc = tadev_memcpy_h2d(ctx, d_k, ksz.data(), nk, s);
  if (!rc) rc = tadev_memset(ctx, d_z, 0, 8, s);
  if (!rc) rc = tadev_shape_gemm_f32(ctx, s, Mt, Nt, Kt, d_a, d_b, Kt ? d_k : nullptr, abs_factor, thr, d_o, d_z);
  if (!rc && mask) {
    rc = tadev_alloc(ctx, std::max<size_t>(no, 4), (void**)&d_m, s);
    if (!rc) rc = tadev_memcpy_h2d(ctx, d_m, mask, no, s);
    if (!rc) rc = tadev_shape_mask_f32(ctx, s, (int64_t)Mt * Nt, d_o, d_m, thr, mask_thr, d_z);
  }
  out.resize((size_t)Mt * Nt);
  if (!rc) rc = tadev_memcpy_d2h(ctx, out.data(), d_o, no, s);
  if (!rc) rc = tadev_memcpy_d2h(ctx, nzero, d_z, 8, s);
  if (!rc) rc = tadev_stream_sync(ctx, s);
  tadev_free(ctx, d_a, s); tadev_free(ctx, d_b, s); tadev_free(ctx, d_k, s); tadev_free(ctx, d_o, s); tadev_free(ctx, d_z, s);
  tadev_free(ctx, d_m, s);
  return rc;
}

extern "C" int tadev_contract_options_default(tadev_contract_options* o) {
  TADEV_REQUIRE(o, "tadev_contract_options_default: null");
  memset(o, 0, sizeof(*o));
  o->exchange_operands = 1;
  o->stream_permutes = -1;
  o->stream_permute_bytes = (int64_t)8 << 30;
  o->threshold = 1.1920928955078125e-07f;  // SparseShape<float> default: FLT_EPSILON (sparse_shape.h:1941)
  o->mask_threshold = o->threshold;
  return TADEV_OK;
}

// [host] GeneralPermutationOptimizer (expressions/permopt.h; tests/general_product.cpp:96-135): indices
// present in both arguments AND in the target are fused (Hadamard / batch) indices. Canonical layouts:
//   left (fused..., left external..., contracted...)   right (fused..., contracted..., right external...)
//   result (fused..., left external..., right external...)
// with the fused and external classes ordered as in the TARGET (so a target already laid out that way
// needs no result permutation) and the contracted class as in the left argument. *nfused = 0 and an
// untouched plan mean "not a general product" (pure contraction); a pure Hadamard product is an error
// here (it belongs to the element-wise engine).
extern "C" int tadev_plan_general_product(const char* target, const char* left, const char* right,
                                          tadev_contraction_plan* out, int32_t* nfused) {
  TADEV_REQUIRE(target && left && right && out && nfused, "tadev_plan_general_product: null");
  *nfused = 0;
  const auto T = split_idx(target), Li = split_idx(left), Ri = split_idx(right);
  auto has = [](const std::vector<std::string>& v, const std::string& x) { return std::find(v.begin(), v.end(), x) != v.end(); };
  std::vector<std::string> Hh, oL, Kk, oR;
  for (auto& x : T) if (has(Li, x) && has(Ri, x)) Hh.push_back(x);
  if (Hh.empty()) return TADEV_OK;
  for (auto& x : T) { if (has(Li, x) && !has(Ri, x)) oL.push_back(x); if (has(Ri, x) && !has(Li, x)) oR.push_back(x); }
  for (auto& x : Li) if (has(Ri, x) && !has(T, x)) Kk.push_back(x);
  TADEV_REQUIRE(!(Kk.empty() && oL.empty() && oR.empty()), "a pure Hadamard product is evaluated by the element-wise engine (tadev_elementwise_create)");
  TADEV_REQUIRE(Li.size() <= 16 && Ri.size() <= 16 && T.size() <= 16, "general product: rank > 16");
  // every index of an argument must be fused, contracted or kept: no implicit reductions
  for (auto& x : Li) TADEV_REQUIRE(has(Ri, x) || has(T, x), "general product: left index '%s' appears in neither the right argument nor the target", x.c_str());
  for (auto& x : Ri) TADEV_REQUIRE(has(Li, x) || has(T, x), "general product: right index '%s' appears in neither the left argument nor the target", x.c_str());
  for (size_t i = 0; i < T.size(); ++i) {
    TADEV_REQUIRE(has(Li, T[i]) || has(Ri, T[i]), "general product: target index '%s' is in neither argument", T[i].c_str());
    for (size_t j = 0; j < i; ++j) TADEV_REQUIRE(T[i] != T[j], "general product: repeated target index '%s'", T[i].c_str());
  }
  std::vector<std::string> cL = Hh, cR = Hh, cRes = Hh;
  cL.insert(cL.end(), oL.begin(), oL.end()); cL.insert(cL.end(), Kk.begin(), Kk.end());
  cR.insert(cR.end(), Kk.begin(), Kk.end()); cR.insert(cR.end(), oR.begin(), oR.end());
  cRes.insert(cRes.end(), oL.begin(), oL.end()); cRes.insert(cRes.end(), oR.begin(), oR.end());
  TADEV_REQUIRE(cL.size() == Li.size() && cR.size() == Ri.size() && cRes.size() == T.size(), "general product: repeated index in an argument");
  tadev_contraction_plan G;
  memset(&G, 0, sizeof(G));
  G.left_rank = (int32_t)Li.size(); G.right_rank = (int32_t)Ri.size(); G.result_rank = (int32_t)cRes.size();
  G.inner_rank = (int32_t)Kk.size(); G.opA = G.opB = TADEV_OP_N; G.left_permtype = G.right_permtype = 1;
  for (int i = 0; i < 16; ++i) G.perm_left[i] = G.perm_right[i] = G.perm_result[i] = -1;
  auto image = [](const std::vector<std::string>& from, const std::vector<std::string>& to, int32_t* perm) {
    bool ident = true;
    for (size_t i = 0; i < from.size(); ++i) ident = ident && from[i] == to[i];
    if (ident) return;
    for (size_t i = 0; i < from.size(); ++i) perm[i] = (int32_t)(std::find(to.begin(), to.end(), from[i]) - to.begin());
  };
  image(Li, cL, G.perm_left); image(Ri, cR, G.perm_right); image(cRes, T, G.perm_result);
  if (G.perm_left[0] >= 0) G.left_permtype = 3;
  if (G.perm_right[0] >= 0) G.right_permtype = 3;
  auto join = [](const std::vector<std::string>& v, char* dst, size_t cap) {
    std::string j;
    for (size_t i = 0; i < v.size(); ++i) { if (i) j += ","; j += v[i]; }
    snprintf(dst, cap, "%s", j.c_str());
  };
  join(cL, G.left_target, sizeof(G.left_target)); join(cR, G.right_target, sizeof(G.right_target)); join(cRes, G.result_gemm, sizeof(G.result_gemm));
  *out = G;
  *nfused = (int32_t)Hh.size();
  return TADEV_OK;
}

// Everything ContEngine::init_struct decides from the index lists and the tilings alone (no device work):
// plan (+ optional operand exchange), operand structures, GemmHelper ranges, fused extents, result tiling.
static int plan_structure(tadev_contraction* Ep, int nranks, const char* target, const char* left_idx, const char* right_idx,
                          const tadev_array_desc* left, const tadev_array_desc* right, bool need_tiles) {
  tadev_contraction* E = Ep;
  int rc;
  {
    int32_t nfused = 0;
    rc = tadev_plan_general_product(target, left_idx, right_idx, &E->plan, &nfused);
    if (rc) return rc;
    if (nfused > 0) { E->general = true; E->nh = nfused; }
    else if (E->opt.exchange_operands) rc = tadev_plan_contraction_opt(target, left_idx, right_idx, &E->plan, &E->swapped);
    else rc = tadev_plan_contraction(target, left_idx, right_idx, &E->plan);
  }
  if (rc) return rc;
  const tadev_contraction_plan& P = E->plan;
  if ((rc = copy_operand(E->swapped ? right : left, E->L, "tadev_contraction_create(left)", need_tiles))) return rc;
  if ((rc = copy_operand(E->swapped ? left : right, E->R, "tadev_contraction_create(right)", need_tiles))) return rc;
  TADEV_REQUIRE(E->L.tr.rank() == P.left_rank && E->R.tr.rank() == P.right_rank, "index list rank does not match the array");
  // (host-resident and lazy operands that need an explicit permutation are permuted tile by tile when a SUMMA window
  // asks for them: build_view configures the permute provider for their memory kind)
  build_operand_structure(E->L, P.perm_left, P.left_rank);
  build_operand_structure(E->R, P.perm_right, P.right_rank);

  // GemmHelper ranges (math/gemm_helper.h:62-98)
  const int lr = P.left_rank, rr = P.right_rank, nc = P.inner_rank;
  if (P.opA == TADEV_OP_N) { E->lo[0] = 0; E->lo[1] = E->li[0] = lr - nc; E->li[1] = lr; }
  else { E->li[0] = 0; E->li[1] = E->lo[0] = nc; E->lo[1] = lr; }
  if (P.opB == TADEV_OP_N) { E->ri[0] = 0; E->ri[1] = E->ro[0] = nc; E->ro[1] = rr; }
  else { E->ro[0] = 0; E->ro[1] = E->ri[0] = rr - nc; E->ri[1] = rr; }
  const TRange &trA = E->L.ptr, &trB = E->R.ptr;
  if (E->general) {
    // ranges in the canonical layouts: left (H, outer, K), right (H, K, outer)
    const int nh = E->nh;
    TADEV_REQUIRE(E->L.d.memory == TADEV_MEM_DEVICE && E->R.d.memory == TADEV_MEM_DEVICE, "general products take device-resident arrays");
    (void)nranks;  // multi-rank general products: every fused slab is a SUMMA on the shared 2-d grid (see eval)
    E->lo[0] = nh; E->lo[1] = E->li[0] = lr - nc; E->li[1] = lr;
    E->ri[0] = nh; E->ri[1] = E->ro[0] = nh + nc; E->ro[1] = rr;
    for (int d = 0; d < nh; ++d) TADEV_REQUIRE(trA.b[d] == trB.b[d], "general product: the fused tiled ranges are not congruent");
    E->h_ext = fused_ext(trA, 0, nh);
    E->Ht = (int)E->h_ext.size();
    for (int d = 0; d < nh; ++d) E->tr_gemm.b.push_back(trA.b[d]);
  }
  for (int d = 0; d < nc; ++d)
    TADEV_REQUIRE(trA.b[E->li[0] + d] == trB.b[E->ri[0] + d], "contraction: inner tiled ranges are not congruent");
  E->m_ext = fused_ext(trA, E->lo[0], E->lo[1]);
  E->n_ext = fused_ext(trB, E->ro[0], E->ro[1]);
  E->k_ext = fused_ext(trA, E->li[0], E->li[1]);
  TADEV_REQUIRE(E->m_ext.size() < (1u << 30) && E->n_ext.size() < (1u << 30) && E->k_ext.size() < (1u << 30), "tile grid too large");
  E->Mt = (int)E->m_ext.size(); E->Nt = (int)E->n_ext.size(); E->Kt = (int)E->k_ext.size();
  for (int d = E->lo[0]; d < E->lo[1]; ++d) E->tr_gemm.b.push_back(trA.b[d]);
  for (int d = E->ro[0]; d < E->ro[1]; ++d) E->tr_gemm.b.push_back(trB.b[d]);
  if (P.perm_result[0] >= 0) E->perm_res.assign(P.perm_result, P.perm_result + P.result_rank);
  if (E->perm_res.empty()) E->tr_target = E->tr_gemm;
  else {
    E->tr_target.b.resize(P.result_rank);
    for (int i = 0; i < P.result_rank; ++i) E->tr_target.b[E->perm_res[i]] = E->tr_gemm.b[i];
  }
  return TADEV_OK;
}

// [host] Where SUMMA wants the operand tiles of this expression: the process grid ProcGrid chooses for the
// result and, for every tile of each USER operand (original tiling, row-major ordinal), its position in the
// fused GEMM-side tile grid. The tile belongs on rank (frow % Pr) * Pc + fcol % Pc (proc_grid.h:566-597); an
// arena ordered by (fcol, frow) for the GEMM's left operand and by (frow, fcol) for the right one makes every
// SUMMA panel a contiguous byte range. role[x] = 0: operand x is the GEMM's left operand A(i,k), 1: right B(k,j)
// (the engine may exchange the operands). Arrays created this way need no redistribution.
extern "C" int tadev_contraction_layout(const char* target, const char* left_idx, const char* right_idx,
                                        const tadev_array_desc* left, const tadev_array_desc* right,
                                        const tadev_contract_options* options, int nranks, tadev_contraction_layout_info* info,
                                        int32_t* left_frow, int32_t* left_fcol, int32_t* right_frow, int32_t* right_fcol) {
  TADEV_REQUIRE(target && left_idx && right_idx && left && right && info && nranks >= 1, "tadev_contraction_layout: bad args");
  std::unique_ptr<tadev_contraction> E(new tadev_contraction());
  if (options) E->opt = *options; else tadev_contract_options_default(&E->opt);
  int rc = plan_structure(E.get(), nranks, target, left_idx, right_idx, left, right, false);
  if (rc) return rc;
  memset(info, 0, sizeof(*info));
  info->swapped = E->swapped; info->Mt = E->Mt; info->Nt = E->Nt; info->Kt = E->Kt;
  info->opA = E->plan.opA; info->opB = E->plan.opB;
  info->left_role = E->swapped ? 1 : 0; info->right_role = E->swapped ? 0 : 1;
  info->Pr = info->Pc = 1;
  if (nranks > 1) {
    tadev_proc_grid g;
    const int64_t msum = std::accumulate(E->m_ext.begin(), E->m_ext.end(), (int64_t)0), nsum = std::accumulate(E->n_ext.begin(), E->n_ext.end(), (int64_t)0);
    if ((rc = tadev_proc_grid_make(0, nranks, E->Mt, E->Nt, msum, nsum, &g))) return rc;
    info->Pr = g.proc_rows; info->Pc = g.proc_cols;
  }
  for (int side = 0; side < 2; ++side) {
    const Operand& o = side == 0 ? E->L : E->R;  // GEMM side
    const bool user_left = (side == 0) != (E->swapped != 0);
    int32_t* frow = user_left ? left_frow : right_frow;
    int32_t* fcol = user_left ? left_fcol : right_fcol;
    if (!frow || !fcol) continue;
    const bool op_n = (side == 0 ? E->plan.opA : E->plan.opB) == TADEV_OP_N;
    // general products: the fused slabs share ONE 2-d grid, a tile's owner does not depend on its slab
    // (contraction_eval.h:96-104: "nh_ independent SUMMA slabs sharing one process grid"): positions are slab-local
    const int64_t rows = side == 0 ? E->Mt : E->Kt, cols = side == 0 ? E->Kt : E->Nt;
    const int R = o.tr.rank();
    const std::vector<int64_t> tshape = o.tr.tiles_shape(), pshape = o.ptr.tiles_shape();
    std::vector<int64_t> idx, pidx(R);
    for (int64_t ord = 0; ord < o.tr.total(); ++ord) {
      int64_t po = ord;
      if (!o.perm.empty()) {
        unravel(ord, tshape, idx);
        for (int i = 0; i < R; ++i) pidx[o.perm[i]] = idx[i];
        po = ravel(pidx, pshape);
      }
      if (E->general) po %= rows * cols;
      const int64_t pos = op_n ? po : (po % rows) * cols + po / rows;  // == fused_pos()
      frow[ord] = (int32_t)(pos / cols); fcol[ord] = (int32_t)(pos % cols);
    }
  }
  return TADEV_OK;
}

extern "C" int tadev_contraction_create(tadev_ctx* ctx, const char* target, const char* left_idx, const char* right_idx,
                                        const tadev_array_desc* left, const tadev_array_desc* right, double factor,
                                        const tadev_contract_options* options, tadev_contraction** out) {
  TADEV_REQUIRE(ctx && target && left_idx && right_idx && left && right && out, "tadev_contraction_create: null");
  std::unique_ptr<tadev_contraction> E(new tadev_contraction());
  E->ctx = ctx;
  E->factor = factor;
  if (options) E->opt = *options; else tadev_contract_options_default(&E->opt);
  int rc = plan_structure(E.get(), ctx->nranks, target, left_idx, right_idx, left, right, true);
  if (rc) return rc;
  const tadev_contraction_plan& P = E->plan;
  const int nc = P.inner_rank;
  const TRange& trA = E->L.ptr;

  // result shape (make_shape): sparse x sparse -> device screening
  E->sparse = !(E->L.dense() && E->R.dense());
  const float thr = E->opt.threshold;
  TADEV_REQUIRE(E->sparse || !E->opt.mask_norms, "set_shape: a result mask needs block-sparse operands (the dense policy has no shape override)");
  if (E->sparse) {
    // (the reference static_asserts that both arguments use the same policy, expressions/mult_engine.h:112-115)
    TADEV_REQUIRE(!E->L.dense() && !E->R.dense(), "mixed dense/sparse contraction is not supported");
    const int Mt = E->Mt, Nt = E->Nt, Kt = E->Kt;
    // 2-D views in GEMM orientation
    E->a_n.resize((size_t)E->Ht * Mt * Kt);
    E->b_n.resize((size_t)E->Ht * Kt * Nt);
    if (P.opA == TADEV_OP_N) E->a_n = E->L.pnorms;
    else for (int k = 0; k < Kt; ++k) for (int i = 0; i < Mt; ++i) E->a_n[(size_t)i * Kt + k] = E->L.pnorms[(size_t)k * Mt + i];
    if (P.opB == TADEV_OP_N) E->b_n = E->R.pnorms;
    else for (int j = 0; j < Nt; ++j) for (int k = 0; k < Kt; ++k) E->b_n[(size_t)k * Nt + j] = E->R.pnorms[(size_t)j * Kt + k];
    std::vector<float> ksz;
    if (nc > 0) ksz = recursive_outer(trA, E->li[0], E->li[1]);
    // user result mask (Expr::set_shape, expressions/expr.h:116; applied after make_shape, cont_engine.h:526-528):
    // given over the TARGET tile grid, brought to GEMM order for the device kernel
    std::vector<float> mask_g;
    if (E->opt.mask_norms) {
      const int64_t nt = E->tr_target.total();
      mask_g.assign(E->opt.mask_norms, E->opt.mask_norms + nt);
      if (!E->perm_res.empty()) {
        std::vector<int> inv(E->perm_res.size());
        for (size_t a = 0; a < E->perm_res.size(); ++a) inv[E->perm_res[a]] = (int)a;
        mask_g = permute_norms(mask_g, E->tr_target.tiles_shape(), inv);
      }
      E->opt.mask_norms = nullptr;  // the caller's array is not kept
    }
    if (!E->general) {
      rc = device_shape_gemm(ctx, Mt, Nt, nc > 0 ? Kt : 0, E->a_n, E->b_n, ksz, (float)std::fabs(factor), thr, E->c_n, &E->nzero,
                             mask_g.empty() ? nullptr : mask_g.data(), E->opt.mask_threshold);
      if (rc) return rc;
    } else {
      // SparseShape::gemm_batched (sparse_shape.h:1707-1900): every fused-index slab is a norm GEMM with the
      // same contracted-size scaling (or a per-slab outer product when nothing is contracted)
      E->c_n.assign((size_t)E->Ht * Mt * Nt, 0.0f);
      E->nzero = 0;
      std::vector<float> as((size_t)Mt * Kt), bs((size_t)Kt * Nt), cs;
      for (int h = 0; h < E->Ht; ++h) {
        std::copy(E->a_n.begin() + (size_t)h * Mt * Kt, E->a_n.begin() + (size_t)(h + 1) * Mt * Kt, as.begin());
        std::copy(E->b_n.begin() + (size_t)h * Kt * Nt, E->b_n.begin() + (size_t)(h + 1) * Kt * Nt, bs.begin());
        uint64_t nz = 0;
        rc = device_shape_gemm(ctx, Mt, Nt, nc > 0 ? Kt : 0, as, bs, ksz, (float)std::fabs(factor), thr, cs, &nz,
                               mask_g.empty() ? nullptr : mask_g.data() + (size_t)h * Mt * Nt, E->opt.mask_threshold);
        if (rc) return rc;
        std::copy(cs.begin(), cs.end(), E->c_n.begin() + (size_t)h * Mt * Nt);
        E->nzero += nz;
      }
    }
    E->target_norms = E->perm_res.empty() ? E->c_n : permute_norms(E->c_n, E->tr_gemm.tiles_shape(), E->perm_res);
  }

  // distribution (init_distribution): the grid the communicators were built for must be the one
  // ProcGrid chooses for this result
  E->Pr = ctx->Pr; E->Pc = ctx->Pc; E->r = ctx->my_r; E->c = ctx->my_c;
  if (ctx->nranks > 1) {
    tadev_proc_grid g;
    const int64_t msum = std::accumulate(E->m_ext.begin(), E->m_ext.end(), (int64_t)0), nsum = std::accumulate(E->n_ext.begin(), E->n_ext.end(), (int64_t)0);
    if ((rc = tadev_proc_grid_make(ctx->rank, ctx->nranks, E->Mt, E->Nt, msum, nsum, &g))) return rc;
    TADEV_REQUIRE(g.proc_rows == E->Pr && g.proc_cols == E->Pc, "communicators were built for a %dx%d grid but ProcGrid chooses %dx%d",
                  E->Pr, E->Pc, g.proc_rows, g.proc_cols);
  }

  // local non-zero result tiles: C(i,j) -> rank (i % Pr, j % Pc), GEMM row-major order
  if (E->r >= 0) {
    const std::vector<int64_t> gshape = E->tr_gemm.tiles_shape(), tshape = E->tr_target.tiles_shape();
    std::vector<int64_t> gidx, tidx(gshape.size());
    int64_t off = 0;
    for (int h = 0; h < E->Ht; ++h)
    for (int i = E->r; i < E->Mt; i += E->Pr)
      for (int j = E->c; j < E->Nt; j += E->Pc) {
        const int64_t key = ((int64_t)h * E->Mt + i) * E->Nt + j;  // ordinal in the GEMM-order tile grid (H, M, N)
        if (E->sparse && E->c_n[key] < thr) continue;
        const int64_t e = (E->general ? E->h_ext[h] : 1) * E->m_ext[i] * E->n_ext[j];
        E->keys.push_back(key); E->elems.push_back(e); E->offs.push_back(off);
        off += (e + 1) & ~(int64_t)1;
        if (E->perm_res.empty()) E->target_ord.push_back(key);
        else {
          unravel(key, gshape, gidx);
          for (size_t a = 0; a < gidx.size(); ++a) tidx[E->perm_res[a]] = gidx[a];
          E->target_ord.push_back(ravel(tidx, tshape));
        }
      }
    E->arena_elems = std::max<int64_t>(off, 2);
  } else E->arena_elems = 2;
  for (int d = 0; d < E->tr_target.rank(); ++d) {
    E->tn.push_back((int32_t)E->tr_target.ntiles(d));
    E->tb_flat.insert(E->tb_flat.end(), E->tr_target.b[d].begin(), E->tr_target.b[d].end());
  }
  *out = E.release();
  return TADEV_OK;
}

extern "C" int tadev_contraction_info_get(const tadev_contraction* E, tadev_contraction_info* info) {
  TADEV_REQUIRE(E && info, "tadev_contraction_info_get: null");
  memset(info, 0, sizeof(*info));
  info->rank = E->tr_target.rank();
  info->swapped = E->swapped;
  info->bounds = E->tb_flat.data();
  info->ntiles = E->tn.data();
  info->norms = E->sparse ? E->target_norms.data() : nullptr;
  info->nzero = E->nzero;
  info->nlocal = (int64_t)E->keys.size();
  info->ordinals = E->target_ord.data();
  info->elems = E->elems.data();
  info->offsets = E->offs.data();
  info->arena_elems = E->arena_elems;
  info->Pr = E->Pr; info->Pc = E->Pc; info->Mt = E->Mt; info->Nt = E->Nt; info->Kt = E->Kt;
  info->opA = E->plan.opA; info->opB = E->plan.opB;
  info->needs_result_permute = E->perm_res.empty() ? 0 : 1;
  return TADEV_OK;
}

// owner of a result tile given its ordinal in the TARGET tiling: the cyclic map of the GEMM-order
// grid (result pmap = grid pmap, cont_engine.h:584)
extern "C" int tadev_contraction_owner(const tadev_contraction* E, int64_t target_ordinal, int* owner) {
  TADEV_REQUIRE(E && owner, "tadev_contraction_owner: null");
  const std::vector<int64_t> tshape = E->tr_target.tiles_shape();
  TADEV_REQUIRE(target_ordinal >= 0 && target_ordinal < E->tr_target.total(), "tadev_contraction_owner: ordinal out of range");
  std::vector<int64_t> tidx, gidx(tshape.size());
  unravel(target_ordinal, tshape, tidx);
  if (E->perm_res.empty()) gidx = tidx;
  else for (size_t a = 0; a < tidx.size(); ++a) gidx[a] = tidx[E->perm_res[a]];
  const int64_t go = ravel(gidx, E->tr_gemm.tiles_shape());
  *owner = (int)((((go / E->Nt) % E->Mt) % E->Pr) * E->Pc + (go % E->Nt) % E->Pc);  // (% Mt: the owner does not depend on the fused slab)
  return TADEV_OK;
}

extern "C" int tadev_contraction_destroy(tadev_contraction* E) {
  if (E) {
    for (Operand* o : {&E->L, &E->R})
      if (o->redist_arena) tadev_free(E->ctx, o->redist_arena, (tadev_stream)E->ctx->streams[0]);
  }
  delete E;
  return TADEV_OK;
}

namespace {

// The operand as the SUMMA driver sees it: a table over the fused tile grid holding pointers or
// provider tokens, plus (for just-in-time permutes) the provider's source tables.
struct View {
  std::vector<const double*> table;
  bool lazy = false;
  tadev_tile_provider provider = nullptr;
  tadev_uniform_source usrc{};
  tadev_permute_source psrc{};
  std::vector<int64_t> p_ext;      // [ntok][rank]
  std::vector<const void*> p_src;  // [ntok]
  std::vector<int64_t> p_ord;      // [ntok] original tile ordinals (lazy sources)
  double* tmp_arena = nullptr;     // up-front permuted copy
  void* user() { return provider == tadev_provider_uniform ? (void*)&usrc : (void*)&psrc; }
};

// position of the tile with PERMUTED ordinal `po` in the fused [rows x cols] table of the GEMM
inline size_t fused_pos(int64_t po, bool op_n, int64_t rows, int64_t cols) {
  // op N: the permuted tile grid is already [rows][cols]; op T: it is stored [cols][rows]
  return op_n ? (size_t)po : (size_t)((po % rows) * cols + po / rows);
}

// Redistribution: bring every non-zero tile of the operand to the rank SUMMA's cyclic maps want it on
// (A(i,k) -> (i % Pr, k % Pc), B(k,j) -> (k % Pr, j % Pc) of the permuted, fused tile grid). Every rank
// derives the same plan from the replicated shape + process map; tiles travel in ascending ordinal order.
int redistribute_operand(tadev_contraction* E, Operand& o, bool is_left) {
  tadev_ctx* ctx = E->ctx;
  if (ctx->nranks == 1 || o.owners.empty() || o.redistributed || o.d.memory == TADEV_MEM_LAZY) return TADEV_OK;
  const bool op_n = (is_left ? E->plan.opA : E->plan.opB) == TADEV_OP_N;
  const int64_t rows = is_left ? E->Mt : E->Kt, cols = is_left ? E->Kt : E->Nt;
  const int64_t n = o.tr.total();
  const float thr = E->opt.threshold;
  const int R = o.tr.rank(), me = ctx->rank;
  const std::vector<int64_t> tshape = o.tr.tiles_shape(), pshape = o.ptr.tiles_shape();
  std::vector<int64_t> idx, pidx(R);
  std::vector<const void*> src;
  std::vector<void*> dst;
  std::vector<size_t> sbytes, rbytes;
  std::vector<int32_t> to, from;
  std::vector<int64_t> recv_ord, recv_off;
  int64_t off = 0;
  bool any_move = false;
  std::vector<int32_t> need((size_t)n, -1);
  for (int64_t ord = 0; ord < n; ++ord) {
    if (!o.dense() && o.norms[ord] < thr) continue;
    int64_t po = ord;
    unravel(ord, tshape, idx);
    if (!o.perm.empty()) { for (int i = 0; i < R; ++i) pidx[o.perm[i]] = idx[i]; po = ravel(pidx, pshape); }
    const size_t pos = fused_pos(E->general ? po % (rows * cols) : po, op_n, rows, cols);  // general: slab-local position
    const int64_t fr = (int64_t)pos / cols, fc = (int64_t)pos % cols;
    need[ord] = (int32_t)((fr % E->Pr) * E->Pc + fc % E->Pc);
    TADEV_REQUIRE(o.owners[ord] >= 0 && o.owners[ord] < ctx->nranks, "redistribution: tile %lld has owner %d", (long long)ord, o.owners[ord]);
    if (need[ord] == o.owners[ord]) continue;
    any_move = true;
    size_t vol = 1;
    for (int d = 0; d < R; ++d) vol *= (size_t)o.tr.ext(d, idx[d]);
    if (o.owners[ord] == me) {
      TADEV_REQUIRE(o.tiles[ord], "redistribution: tile %lld is owned by this rank but has no data", (long long)ord);
      src.push_back(o.tiles[ord]); sbytes.push_back(vol * 8); to.push_back(need[ord]);
    } else if (need[ord] == me) {
      recv_ord.push_back(ord); recv_off.push_back(off); rbytes.push_back(vol * 8); from.push_back(o.owners[ord]);
      off += (int64_t)((vol + 1) & ~(size_t)1);
    }
  }
  o.redistributed = true;
  if (!any_move) return TADEV_OK;  // same decision on every rank: the plan is built from replicated data
  o.tiles_before = o.tiles;
  TADEV_REQUIRE(o.d.memory == TADEV_MEM_DEVICE, "redistribution moves device-resident tiles only");
  tadev_stream s = (tadev_stream)ctx->streams[0];
  if (off > 0) { int rc = tadev_alloc(ctx, (size_t)off * 8, (void**)&o.redist_arena, s); if (rc) return rc; }
  for (size_t t = 0; t < recv_ord.size(); ++t) dst.push_back(o.redist_arena + recv_off[t]);
  int rc = tadev_exchange_tiles(ctx, s, (int)src.size(), src.data(), sbytes.data(), to.data(), (int)dst.size(), dst.data(), rbytes.data(), from.data());
  if (rc) return rc;
  // the effective local tile table: tiles that stayed, tiles that arrived; tiles that left are gone
  for (int64_t ord = 0; ord < n; ++ord) if (need[ord] >= 0 && need[ord] != me) o.tiles[ord] = nullptr;
  for (size_t t = 0; t < recv_ord.size(); ++t) o.tiles[recv_ord[t]] = dst[t];
  return TADEV_OK;
}

int build_view(tadev_contraction* E, Operand& o, bool is_left, View& v, float* permute_ms_acc) {
  tadev_ctx* ctx = E->ctx;
  const bool op_n = (is_left ? E->plan.opA : E->plan.opB) == TADEV_OP_N;
  // general products: the fused slabs are the leading (slowest) part of the row index
  const int64_t rows = (int64_t)(E->general ? E->Ht : 1) * (is_left ? E->Mt : E->Kt), cols = is_left ? E->Kt : E->Nt;
  v.table.assign((size_t)std::max<int64_t>(rows * cols, 1), nullptr);
  const int64_t n = o.tr.total();
  const float thr = E->opt.threshold;
  const int R = o.tr.rank();
  const std::vector<int64_t> tshape = o.tr.tiles_shape(), pshape = o.ptr.tiles_shape();
  std::vector<int64_t> idx, pidx(R);
  auto permuted_ordinal = [&](int64_t ord) -> int64_t {
    if (o.perm.empty()) return ord;
    unravel(ord, tshape, idx);
    for (int i = 0; i < R; ++i) pidx[o.perm[i]] = idx[i];
    return ravel(pidx, pshape);
  };
  // a lazy tile can be generated anywhere: it is "local" exactly where SUMMA's cyclic maps want it, whatever table
  // the caller passed (this is the redistribution of a lazy operand: none is needed)
  auto lazy_local = [&](int64_t po) {
    if (ctx->nranks == 1) return true;
    const size_t pos = fused_pos(po, op_n, rows, cols);
    const int64_t fr = (int64_t)pos / cols, fc = (int64_t)pos % cols;
    return (int)((fr % E->Pr) * E->Pc + fc % E->Pc) == ctx->rank;
  };
  if (o.d.memory == TADEV_MEM_LAZY && o.perm.empty()) {
    v.lazy = true;
    v.provider = tadev_provider_uniform;
    v.usrc.ctx = ctx; v.usrc.seed = o.d.lazy_seed;
    for (int64_t ord = 0; ord < n; ++ord) {
      if (!o.dense() && o.norms[ord] < thr) continue;
      if (!lazy_local(ord)) continue;
      v.table[fused_pos(ord, op_n, rows, cols)] = reinterpret_cast<const double*>((uintptr_t)(ord + 1));
    }
    return TADEV_OK;
  }
  if (o.perm.empty()) {
    for (int64_t ord = 0; ord < n; ++ord)
      if (o.tiles[ord]) v.table[fused_pos(ord, op_n, rows, cols)] = static_cast<const double*>(o.tiles[ord]);
    return TADEV_OK;
  }
  // explicit permutation: gather the local tiles
  std::vector<int64_t> ords, exts;
  int64_t bytes = 0;
  for (int64_t ord = 0; ord < n; ++ord) {
    if (o.d.memory == TADEV_MEM_LAZY) {
      if ((!o.dense() && o.norms[ord] < thr) || !lazy_local(permuted_ordinal(ord))) continue;
    } else if (!o.tiles[ord]) continue;
    ords.push_back(ord);
    unravel(ord, tshape, idx);
    int64_t vol = 1;
    for (int d = 0; d < R; ++d) { exts.push_back(o.tr.ext(d, idx[d])); vol *= exts.back(); }
    bytes += vol * 8;
  }
  // host-resident and lazy sources are always permuted per SUMMA window (there is no device copy to permute up front)
  bool stream = !E->general && (o.d.memory != TADEV_MEM_DEVICE || E->opt.stream_permutes > 0 ||
                                (E->opt.stream_permutes < 0 && bytes > E->opt.stream_permute_bytes));
  TADEV_REQUIRE(o.d.memory == TADEV_MEM_DEVICE || stream, "general products take device-resident arrays");
  if (stream) {
    v.lazy = true;
    v.provider = tadev_provider_permute;
    v.p_ext = exts;
    v.p_ord = ords;
    if (o.d.memory != TADEV_MEM_LAZY) for (int64_t ord : ords) v.p_src.push_back(o.tiles[ord]);
    v.psrc.ctx = ctx; v.psrc.rank = R;
    for (int i = 0; i < R; ++i) v.psrc.perm[i] = o.perm[i];
    v.psrc.extents = v.p_ext.data(); v.psrc.src = v.p_src.empty() ? nullptr : v.p_src.data();
    v.psrc.src_memory = o.d.memory; v.psrc.lazy_seed = o.d.lazy_seed; v.psrc.ordinals = v.p_ord.data();
    for (size_t t = 0; t < ords.size(); ++t)
      v.table[fused_pos(permuted_ordinal(ords[t]), op_n, rows, cols)] = reinterpret_cast<const double*>((uintptr_t)(t + 1));
    return TADEV_OK;
  }
  // up-front: one arena, one batched launch per distinct tile extent
  tadev_stream s = (tadev_stream)ctx->streams[0];
  std::vector<int64_t> offs(ords.size());
  int64_t off = 0;
  for (size_t t = 0; t < ords.size(); ++t) {
    int64_t vol = 1;
    for (int d = 0; d < R; ++d) vol *= exts[t * R + d];
    offs[t] = off;
    off += (vol + 1) & ~(int64_t)1;
  }
  int rc = tadev_alloc(ctx, (size_t)std::max<int64_t>(off, 2) * 8, (void**)&v.tmp_arena, s);
  if (rc) return rc;
  void *e0 = nullptr, *e1 = nullptr;
  tadev_event_create(ctx, &e0); tadev_event_create(ctx, &e1);
  tadev_event_record(ctx, e0, s);
  std::vector<char> done(ords.size(), 0);
  std::vector<const void*> ins;
  std::vector<void*> outs;
  std::vector<int32_t> perm32(o.perm.begin(), o.perm.end());
  for (size_t t = 0; t < ords.size(); ++t) {
    if (done[t]) continue;
    ins.clear(); outs.clear();
    for (size_t q = t; q < ords.size(); ++q) {
      if (done[q] || memcmp(&exts[q * R], &exts[t * R], sizeof(int64_t) * R) != 0) continue;
      ins.push_back(o.tiles[ords[q]]);
      outs.push_back(v.tmp_arena + offs[q]);
      done[q] = 1;
    }
    rc = tadev_permute_batched(ctx, s, R, &exts[t * R], perm32.data(), 8, (int)ins.size(), ins.data(), outs.data());
    if (rc) return rc;
  }
  tadev_event_record(ctx, e1, s);
  float ms = 0;
  tadev_event_elapsed_ms(ctx, e0, e1, &ms);
  tadev_event_destroy(ctx, e0); tadev_event_destroy(ctx, e1);
  *permute_ms_acc += ms;
  for (size_t t = 0; t < ords.size(); ++t)
    v.table[fused_pos(permuted_ordinal(ords[t]), op_n, rows, cols)] = v.tmp_arena + offs[t];
  return TADEV_OK;
}

}  // namespace

// Evaluate into caller-owned result tiles: result_tiles[t] is the storage of local result tile t of the info
// struct (ordinals[t], elems[t]), device or pinned host memory. With accumulate != 0 the tiles hold the previous
// contents of c and the product is added (c("m,n") += a("m,k") * b("k,n")): the existing array's own tile
// pointers are used, whatever its arena layout is.
extern "C" int tadev_contraction_eval_tiles(tadev_contraction* E, void* const* result_tiles, int result_memory, int accumulate,
                                            tadev_contract_stats* stats) {
  TADEV_REQUIRE(E && (result_tiles || E->keys.empty()), "tadev_contraction_eval_tiles: null");
  TADEV_REQUIRE(result_memory == TADEV_MEM_DEVICE || result_memory == TADEV_MEM_HOST, "tadev_contraction_eval: the result must be device- or host-resident");
  const bool permuted = !E->perm_res.empty();
  TADEV_REQUIRE(!(permuted && accumulate && result_memory == TADEV_MEM_HOST),
                "accumulating into a host-resident result that needs a result permutation is not supported");
  for (size_t t = 0; t < E->keys.size(); ++t) TADEV_REQUIRE(result_tiles[t], "tadev_contraction_eval_tiles: result tile %zu has no storage", t);
  tadev_ctx* ctx = E->ctx;
  tadev_stream s = (tadev_stream)ctx->streams[0];
  if (stats) memset(stats, 0, sizeof(*stats));
  float permute_ms = 0.0f;
  View vA, vB;
  double *gemm_arena = nullptr, *perm_arena = nullptr;
  // every exit path releases the temporaries (stream-ordered frees on the compute stream)
  struct Cleanup {
    tadev_ctx* ctx; tadev_stream s; View *a, *b; double **g, **p; tadev_contraction* E;
    ~Cleanup() {
      if (a->tmp_arena) tadev_free(ctx, a->tmp_arena, s);
      if (b->tmp_arena) tadev_free(ctx, b->tmp_arena, s);
      if (*g) tadev_free(ctx, *g, s);
      if (*p) tadev_free(ctx, *p, s);
      // the received copies of redistributed operand tiles were only needed for this evaluation
      for (Operand* o : {&E->L, &E->R}) {
        if (o->redist_arena) { tadev_free(ctx, o->redist_arena, s); o->redist_arena = nullptr; }
        if (!o->tiles_before.empty()) { o->tiles.swap(o->tiles_before); o->tiles_before.clear(); }
        o->redistributed = false;
      }
    }
  } cleanup{ctx, s, &vA, &vB, &gemm_arena, &perm_arena, E};
  int rc = redistribute_operand(E, E->L, true);
  if (!rc) rc = redistribute_operand(E, E->R, false);
  if (rc) return rc;
  rc = build_view(E, E->L, true, vA, &permute_ms);
  if (!rc) rc = build_view(E, E->R, false, vB, &permute_ms);
  if (rc) return rc;

  // result tiles in GEMM order: straight into the caller's tiles, or into a temporary device arena that is
  // permuted into the caller's tiles afterwards (ContractReduce's post-process, contract_reduce.h:370-378)
  std::vector<double*> c_tab((size_t)std::max<int64_t>((int64_t)E->Ht * E->Mt * E->Nt, 1), nullptr);
  if (permuted) {
    rc = tadev_alloc(ctx, (size_t)E->arena_elems * 8, (void**)&gemm_arena, s);
    if (rc) return rc;
    for (size_t t = 0; t < E->keys.size(); ++t) c_tab[E->keys[t]] = gemm_arena + E->offs[t];
  } else {
    for (size_t t = 0; t < E->keys.size(); ++t) c_tab[E->keys[t]] = static_cast<double*>(result_tiles[t]);
  }
  const int gemm_memory = permuted ? TADEV_MEM_DEVICE : result_memory;
  const int gemm_accumulate = permuted ? 0 : (accumulate ? 1 : 0);

  tadev_summa_stats st{};
  if (E->general && ctx->nranks > 1) {
    // Multi-rank general product: the reference evaluates the nh fused slabs as independent SUMMAs that share one
    // process grid (contraction_eval.h:96-104, proc_h_ == 1). A tile (nb, m, k) holds nb contiguous row-major
    // matrices, so batch element e of slab h is an ordinary matrix SUMMA over tile pointers offset by e*m*k: the
    // driver runs once per (slab, batch element), every rank in the same order. (Panels of one batch element are
    // strided in their tiles and are packed before they travel; moving whole batched tiles once per slab is the
    // obvious next step.)
    TADEV_REQUIRE(gemm_memory == TADEV_MEM_DEVICE, "general products produce device-resident results");
    const int Mt = E->Mt, Nt = E->Nt, Kt = E->Kt;
    std::vector<const double*> at((size_t)Mt * Kt), bt((size_t)Kt * Nt);
    std::vector<double*> ct((size_t)Mt * Nt);
    for (int h = 0; h < E->Ht && !rc; ++h)
      for (int64_t e = 0; e < E->h_ext[h] && !rc; ++e) {
        for (int i = 0; i < Mt; ++i)
          for (int k = 0; k < Kt; ++k) {
            const double* p = vA.table[((size_t)h * Mt + i) * Kt + k];
            at[(size_t)i * Kt + k] = p ? p + e * E->m_ext[i] * E->k_ext[k] : nullptr;
          }
        for (int k = 0; k < Kt; ++k)
          for (int j = 0; j < Nt; ++j) {
            const double* p = vB.table[((size_t)h * Kt + k) * Nt + j];
            bt[(size_t)k * Nt + j] = p ? p + e * E->k_ext[k] * E->n_ext[j] : nullptr;
          }
        for (int i = 0; i < Mt; ++i)
          for (int j = 0; j < Nt; ++j) {
            double* p = c_tab[((size_t)h * Mt + i) * Nt + j];
            ct[(size_t)i * Nt + j] = p ? p + e * E->m_ext[i] * E->n_ext[j] : nullptr;
          }
        tadev_summa_plan sp{};
        sp.Mt = Mt; sp.Nt = Nt; sp.Kt = Kt;
        sp.m_ext = E->m_ext.data(); sp.n_ext = E->n_ext.data(); sp.k_ext = E->k_ext.data();
        sp.opA = TADEV_OP_N; sp.opB = TADEV_OP_N; sp.alpha = E->factor;
        sp.a_norms = E->sparse ? E->a_n.data() + (size_t)h * Mt * Kt : nullptr;
        sp.b_norms = E->sparse ? E->b_n.data() + (size_t)h * Kt * Nt : nullptr;
        sp.c_norms = E->sparse ? E->c_n.data() + (size_t)h * Mt * Nt : nullptr;
        sp.threshold = E->opt.threshold;
        sp.a_tiles = at.data(); sp.b_tiles = bt.data(); sp.c_tiles = ct.data();
        sp.accumulate = gemm_accumulate;
        sp.depth = E->opt.depth; sp.steps_per_launch = E->opt.steps_per_launch;
        tadev_summa_stats one{};
        rc = tadev_summa_f64(ctx, &sp, &one);
        st.nsteps += one.nsteps; st.nsteps_skipped += one.nsteps_skipped; st.nlaunches += one.nlaunches;
        st.bcast_bytes += one.bcast_bytes; st.device_ms += one.device_ms; st.gemm_ms += one.gemm_ms; st.list_ms += one.list_ms;
        st.flops += one.flops;
        if (e == 0) st.npairs += one.npairs;  // pairs are counted per tile, as on one rank
      }
    st.row_blocks = 1;
  } else if (E->general) {
    // BatchedContractReduce (tile_op/batched_contract_reduce.h) for every result tile of every fused slab,
    // in ONE grouped launch: batch element e of tile (h,i,j) is its own group — C + e*m*n accumulates
    // A(h,i,k)[e] * B(h,k,j)[e] over the contracted tiles k (a tile (nb, m, k) is nb row-major m x k matrices).
    TADEV_REQUIRE(gemm_memory == TADEV_MEM_DEVICE, "general products produce device-resident results");
    std::vector<tadev_gemm_group> groups;
    std::vector<tadev_gemm_task> tasks;
    const float thr = E->opt.threshold;
    const int Mt = E->Mt, Nt = E->Nt, Kt = E->Kt;
    for (size_t t = 0; t < E->keys.size(); ++t) {
      const int64_t key = E->keys[t];
      const int h = (int)(key / ((int64_t)Mt * Nt)), i = (int)((key / Nt) % Mt), j = (int)(key % Nt);
      const int64_t m = E->m_ext[i], n = E->n_ext[j], nb = E->h_ext[h];
      for (int64_t e = 0; e < nb; ++e) {
        tadev_gemm_group G{c_tab[key] + e * m * n, (int32_t)m, (int32_t)n, (int32_t)tasks.size(), 0, gemm_accumulate, 0};
        for (int k = 0; k < Kt; ++k) {
          const size_t ak = ((size_t)h * Mt + i) * Kt + k, bk = ((size_t)h * Kt + k) * Nt + j;
          if (E->sparse && (E->a_n[ak] < thr || E->b_n[bk] < thr)) continue;
          const double *At = vA.table[ak], *Bt = vB.table[bk];
          TADEV_REQUIRE(At && Bt, "general product: argument tile (%d,%d,%d) has no data", h, i, k);
          const int64_t kk = E->k_ext[k];
          tasks.push_back({At + e * m * kk, Bt + e * kk * n, (int32_t)kk, 0});
          if (e == 0) { ++st.npairs; st.flops += 2.0 * (double)nb * (double)m * (double)n * (double)kk; }
        }
        G.task_end = (int32_t)tasks.size();
        groups.push_back(G);
        TADEV_REQUIRE(tasks.size() < (size_t)1 << 30 && groups.size() < (size_t)1 << 30, "general product: too many batch tasks for one launch");
      }
    }
    void *e0 = nullptr, *e1 = nullptr;
    tadev_event_create(ctx, &e0); tadev_event_create(ctx, &e1);
    tadev_event_record(ctx, e0, s);
    rc = tadev_gemm_grouped_f64(ctx, s, TADEV_OP_N, TADEV_OP_N, E->factor, groups.data(), (int)groups.size(), tasks.data(), (int)tasks.size());
    tadev_event_record(ctx, e1, s);
    float ms = 0;
    tadev_event_elapsed_ms(ctx, e0, e1, &ms);
    tadev_event_destroy(ctx, e0); tadev_event_destroy(ctx, e1);
    st.nsteps = Kt; st.nlaunches = groups.empty() ? 0 : 1; st.device_ms = ms; st.gemm_ms = ms; st.row_blocks = 1;
  } else {
    tadev_summa_plan sp{};
    sp.Mt = E->Mt; sp.Nt = E->Nt; sp.Kt = E->Kt;
    sp.m_ext = E->m_ext.data(); sp.n_ext = E->n_ext.data(); sp.k_ext = E->k_ext.data();
    sp.opA = E->plan.opA; sp.opB = E->plan.opB; sp.alpha = E->factor;
    sp.a_norms = E->sparse ? E->a_n.data() : nullptr;
    sp.b_norms = E->sparse ? E->b_n.data() : nullptr;
    sp.c_norms = E->sparse ? E->c_n.data() : nullptr;
    sp.threshold = E->opt.threshold;
    sp.a_tiles = vA.table.data(); sp.b_tiles = vB.table.data(); sp.c_tiles = c_tab.data();
    sp.accumulate = gemm_accumulate;
    sp.depth = E->opt.depth; sp.steps_per_launch = E->opt.steps_per_launch; sp.row_blocks = E->opt.row_blocks;
    sp.flags = (E->L.d.memory == TADEV_MEM_HOST && !vA.lazy ? TADEV_SUMMA_A_ON_HOST : 0) | (E->R.d.memory == TADEV_MEM_HOST && !vB.lazy ? TADEV_SUMMA_B_ON_HOST : 0) |
               (gemm_memory == TADEV_MEM_HOST ? TADEV_SUMMA_C_ON_HOST : 0) | (vA.lazy ? TADEV_SUMMA_A_LAZY : 0) | (vB.lazy ? TADEV_SUMMA_B_LAZY : 0);
    if (vA.lazy) { sp.a_provider = vA.provider; sp.a_user = vA.user(); }
    if (vB.lazy) { sp.b_provider = vB.provider; sp.b_user = vB.user(); }
    rc = tadev_summa_f64(ctx, &sp, &st);
  }
  if (rc) return rc;

  // result permutation, batched per tile extent; then (+=) the permuted product is added to the existing
  // tiles, or (host-resident result) copied back tile by tile
  if (permuted) {
    void *e0 = nullptr, *e1 = nullptr;
    tadev_event_create(ctx, &e0); tadev_event_create(ctx, &e1);
    tadev_event_record(ctx, e0, s);
    const int R = E->tr_gemm.rank();
    const std::vector<int64_t> gshape = E->tr_gemm.tiles_shape();
    std::vector<int64_t> exts(E->keys.size() * (size_t)R), gidx;
    for (size_t t = 0; t < E->keys.size(); ++t) {
      unravel(E->keys[t], gshape, gidx);
      for (int d = 0; d < R; ++d) exts[t * R + d] = E->tr_gemm.ext(d, gidx[d]);
    }
    const bool direct = !accumulate && result_memory == TADEV_MEM_DEVICE;
    if (!direct) {
      rc = tadev_alloc(ctx, (size_t)E->arena_elems * 8, (void**)&perm_arena, s);
      if (rc) { tadev_event_destroy(ctx, e0); tadev_event_destroy(ctx, e1); return rc; }
    }
    std::vector<char> done(E->keys.size(), 0);
    std::vector<const void*> ins;
    std::vector<void*> outs;
    std::vector<int32_t> perm32(E->perm_res.begin(), E->perm_res.end());
    for (size_t t = 0; t < E->keys.size() && !rc; ++t) {
      if (done[t]) continue;
      ins.clear(); outs.clear();
      for (size_t q = t; q < E->keys.size(); ++q) {
        if (done[q] || memcmp(&exts[q * R], &exts[t * R], sizeof(int64_t) * R) != 0) continue;
        ins.push_back(gemm_arena + E->offs[q]);
        outs.push_back(direct ? result_tiles[q] : (void*)(perm_arena + E->offs[q]));
        done[q] = 1;
      }
      rc = tadev_permute_batched(ctx, s, R, &exts[t * R], perm32.data(), 8, (int)ins.size(), ins.data(), outs.data());
    }
    if (!rc && !direct && result_memory == TADEV_MEM_DEVICE && !E->keys.empty()) {  // c += permuted product
      std::vector<double*> out(E->keys.size());
      std::vector<const double*> x(E->keys.size()), y(E->keys.size());
      for (size_t t = 0; t < E->keys.size(); ++t) { out[t] = static_cast<double*>(result_tiles[t]); x[t] = out[t]; y[t] = perm_arena + E->offs[t]; }
      rc = tadev_tiles_binary_f64(ctx, s, TADEV_EW_AXPBY, (int)out.size(), out.data(), x.data(), y.data(), E->elems.data(), 1.0, 1.0);
    }
    if (!rc && !direct && result_memory == TADEV_MEM_HOST) {
      for (size_t t = 0; t < E->keys.size() && !rc; ++t)
        rc = tadev_memcpy_d2h(ctx, result_tiles[t], perm_arena + E->offs[t], (size_t)E->elems[t] * 8, s);
      if (!rc) rc = tadev_stream_sync(ctx, s);
    }
    tadev_event_record(ctx, e1, s);
    float ms = 0;
    tadev_event_elapsed_ms(ctx, e0, e1, &ms);
    tadev_event_destroy(ctx, e0); tadev_event_destroy(ctx, e1);
    permute_ms += ms;
    if (rc) return rc;
  }
  if (stats) { stats->summa = st; stats->permute_ms = permute_ms; }
  return TADEV_OK;
}

extern "C" int tadev_contraction_eval(tadev_contraction* E, void* result_arena, int result_memory, int accumulate,
                                      tadev_contract_stats* stats) {
  TADEV_REQUIRE(E && result_arena, "tadev_contraction_eval: null");
  std::vector<void*> tiles(E->keys.size());
  for (size_t t = 0; t < E->keys.size(); ++t) tiles[t] = static_cast<double*>(result_arena) + E->offs[t];
  return tadev_contraction_eval_tiles(E, tiles.data(), result_memory, accumulate, stats);
}

// =================================================================================================
// Element-wise engine: c(idx) = alpha * a(idx_a) [ + beta * b(idx_b) | .* b(idx_b) ]
// The reference's AddEngine / SubtEngine / ScalEngine / Hadamard MultEngine (expressions/add_engine.h,
// subt_engine.h, scal_engine.h, mult_engine.h) for arrays described through the C ABI. Operands whose
// index order differs from the target are permuted first (one batched launch per tile extent), then
// ONE tadev_tiles_binary_f64 launch produces every result tile. Result shapes follow
// SparseShape::scale / add / mult (sparse_shape.h:1243-1262, 1309-1331, 1522-1563) in fp32, operation
// by operation.
namespace {

std::vector<std::string> split_idx(const char* s) {
  std::vector<std::string> out;
  std::string cur;
  for (const char* p = s; *p; ++p) {
    if (*p == ',') { out.push_back(cur); cur.clear(); }
    else if (*p != ' ' && *p != '\t') cur.push_back(*p);
  }
  if (!cur.empty() || !out.empty()) out.push_back(cur);
  return out;
}

// hard-zero below the threshold, count zeros (screen step shared by scale / add / mult)
uint64_t screen(std::vector<float>& v, float thr) {
  uint64_t nz = 0;
  for (float& x : v) if (x < thr) { x = 0.0f; ++nz; }
  return nz;
}

}  // namespace

struct tadev_elementwise {
  tadev_ctx* ctx = nullptr;
  int op = 0;
  double alpha = 1.0, beta = 0.0;
  float thr = 0.0f;
  bool has_b = false;
  Operand A, B;
  TRange tr;                       // result tiling (target order)
  bool sparse = false;
  std::vector<float> norms;        // result shape
  uint64_t nzero = 0;
  std::vector<int64_t> ords, elems, offs;
  int64_t arena_elems = 2;
  std::vector<int64_t> tb_flat;
  std::vector<int32_t> tn;
};

static int ew_prepare_operand(Operand& o, const std::vector<std::string>& idx, const std::vector<std::string>& target, const char* who) {
  TADEV_REQUIRE((int)idx.size() == o.tr.rank(), "%s: index list rank does not match the array", who);
  TADEV_REQUIRE(idx.size() == target.size(), "%s: indices do not match the target", who);
  TADEV_REQUIRE(o.d.memory == TADEV_MEM_DEVICE, "%s: element-wise expressions take device-resident arrays", who);
  int32_t perm[16];
  bool ident = true;
  for (size_t i = 0; i < idx.size(); ++i) {
    const size_t p = std::find(target.begin(), target.end(), idx[i]) - target.begin();
    TADEV_REQUIRE(p < target.size(), "%s: index '%s' is not in the target", who, idx[i].c_str());
    for (size_t j = 0; j < i; ++j) TADEV_REQUIRE(idx[j] != idx[i], "%s: repeated index '%s'", who, idx[i].c_str());
    perm[i] = (int32_t)p;
    ident = ident && p == i;
  }
  if (ident) perm[0] = -1;
  return build_operand_structure(o, perm, (int)idx.size());
}

extern "C" int tadev_elementwise_create(tadev_ctx* ctx, int op, const char* target, double alpha, const char* a_idx,
                                        const tadev_array_desc* a, double beta, const char* b_idx, const tadev_array_desc* b,
                                        float threshold, tadev_elementwise** out) {
  TADEV_REQUIRE(ctx && target && a_idx && a && out, "tadev_elementwise_create: null");
  TADEV_REQUIRE(op == TADEV_EW_AXPBY || op == TADEV_EW_MULT, "tadev_elementwise_create: bad op %d", op);
  TADEV_REQUIRE((b == nullptr) == (b_idx == nullptr), "tadev_elementwise_create: b and b_idx go together");
  TADEV_REQUIRE(b || op == TADEV_EW_AXPBY, "tadev_elementwise_create: a Hadamard product needs two operands");
  std::unique_ptr<tadev_elementwise> E(new tadev_elementwise());
  E->ctx = ctx; E->op = op; E->alpha = alpha; E->beta = b ? beta : 0.0; E->thr = threshold; E->has_b = b != nullptr;
  const auto T = split_idx(target);
  int rc = copy_operand(a, E->A, "tadev_elementwise_create(a)");
  if (!rc) rc = ew_prepare_operand(E->A, split_idx(a_idx), T, "tadev_elementwise_create(a)");
  if (!rc && b) rc = copy_operand(b, E->B, "tadev_elementwise_create(b)");
  if (!rc && b) rc = ew_prepare_operand(E->B, split_idx(b_idx), T, "tadev_elementwise_create(b)");
  if (rc) return rc;
  TADEV_REQUIRE(ctx->nranks == 1 || (E->A.perm.empty() && (!b || E->B.perm.empty())),
                "multi-rank element-wise expressions need operands in the target's index order (no redistribution)");
  E->tr = E->A.ptr;
  if (b) {
    TADEV_REQUIRE(E->B.ptr.b == E->tr.b, "element-wise expression: operand tilings differ");
    TADEV_REQUIRE(E->A.dense() == E->B.dense(), "mixed dense/sparse element-wise expression is not supported");
    if (ctx->nranks > 1) {
      // a tile held by another rank must not be mistaken for a zero tile: both operands need the same process
      // map (the reference redistributes through the result pmap; this engine does not move element-wise operands)
      const int64_t nt = E->tr.total();
      for (int64_t o = 0; o < nt; ++o) {
        const bool za = !E->A.dense() && E->A.norms[o] < threshold, zb = !E->B.dense() && E->B.norms[o] < threshold;
        if (za || zb) continue;
        if (!E->A.owners.empty() && !E->B.owners.empty())
          TADEV_REQUIRE(E->A.owners[o] == E->B.owners[o], "multi-rank element-wise expression: tile %lld lives on rank %d in one operand and on rank %d in the other",
                        (long long)o, E->A.owners[o], E->B.owners[o]);
        TADEV_REQUIRE((E->A.tiles[o] != nullptr) == (E->B.tiles[o] != nullptr),
                      "multi-rank element-wise expression: tile %lld is local in only one operand (the operands need identical process maps)", (long long)o);
      }
    }
  }
  E->sparse = !E->A.dense();
  const int64_t n = E->tr.total();
  if (E->sparse) {
    const float fa = (float)std::fabs(alpha), fb = (float)std::fabs(beta);
    std::vector<float> r = E->A.pnorms;
    if (!b) { for (float& x : r) x *= fa; E->nzero = screen(r, threshold); }
    else if (op == TADEV_EW_AXPBY) {
      std::vector<float> sb = E->B.pnorms;
      for (float& x : r) x *= fa;
      screen(r, threshold);                              // leaf scaling, each thresholded (ScalTsrExpr shapes)
      for (float& x : sb) x *= fb;
      screen(sb, threshold);
      for (int64_t i = 0; i < n; ++i) r[i] = r[i] + sb[i];
      E->nzero = screen(r, threshold);
    } else {
      for (int64_t i = 0; i < n; ++i) r[i] = r[i] * E->B.pnorms[i];
      if (alpha != 1.0) for (float& x : r) x *= fa;
      // scale_tile_norms<ScaleBy::Volume>: rank 1 norm *= size, else norm *= (x * y) of the two outer products
      const int R = E->tr.rank();
      if (R == 1) { for (int64_t i = 0; i < n; ++i) r[i] *= (float)E->tr.ext(0, i); }
      else {
        const int middle = (R >> 1) + (R & 1);
        const std::vector<float> l = recursive_outer(E->tr, 0, middle), rr = recursive_outer(E->tr, middle, R);
        for (size_t i = 0; i < l.size(); ++i)
          for (size_t j = 0; j < rr.size(); ++j) { const float xy = l[i] * rr[j]; r[i * rr.size() + j] *= xy; }
      }
      E->nzero = screen(r, threshold);
    }
    E->norms.swap(r);
  }
  // local result tiles: non-zero in the result shape and (multi-rank) held by this rank in an operand
  const std::vector<int64_t> tshape = E->tr.tiles_shape();
  std::vector<int64_t> idx;
  int64_t off = 0;
  for (int64_t o = 0; o < n; ++o) {
    if (E->sparse && E->norms[o] < threshold) continue;
    if (ctx->nranks > 1 && !(E->A.tiles[o] || (b && E->B.tiles[o]))) continue;
    unravel(o, tshape, idx);
    int64_t e = 1;
    for (int d = 0; d < E->tr.rank(); ++d) e *= E->tr.ext(d, idx[d]);
    E->ords.push_back(o); E->elems.push_back(e); E->offs.push_back(off);
    off += (e + 1) & ~(int64_t)1;
  }
  E->arena_elems = std::max<int64_t>(off, 2);
  for (int d = 0; d < E->tr.rank(); ++d) {
    E->tn.push_back((int32_t)E->tr.ntiles(d));
    E->tb_flat.insert(E->tb_flat.end(), E->tr.b[d].begin(), E->tr.b[d].end());
  }
  *out = E.release();
  return TADEV_OK;
}

extern "C" int tadev_elementwise_info_get(const tadev_elementwise* E, tadev_contraction_info* info) {
  TADEV_REQUIRE(E && info, "tadev_elementwise_info_get: null");
  memset(info, 0, sizeof(*info));
  info->rank = E->tr.rank();
  info->bounds = E->tb_flat.data();
  info->ntiles = E->tn.data();
  info->norms = E->sparse ? E->norms.data() : nullptr;
  info->nzero = E->nzero;
  info->nlocal = (int64_t)E->ords.size();
  info->ordinals = E->ords.data();
  info->elems = E->elems.data();
  info->offsets = E->offs.data();
  info->arena_elems = E->arena_elems;
  info->Pr = E->ctx->Pr; info->Pc = E->ctx->Pc;
  return TADEV_OK;
}

namespace {

// device pointers of the operand's tiles by TARGET ordinal (nullptr = zero tile); operands in another
// index order are permuted into a temporary arena first
int ew_operand_tiles(tadev_ctx* ctx, Operand& o, std::vector<const double*>& by_target, double** tmp_arena) {
  const int64_t n = o.tr.total();
  by_target.assign((size_t)n, nullptr);
  *tmp_arena = nullptr;
  if (o.perm.empty()) {
    for (int64_t t = 0; t < n; ++t) by_target[t] = static_cast<const double*>(o.tiles[t]);
    return TADEV_OK;
  }
  const int R = o.tr.rank();
  const std::vector<int64_t> tshape = o.tr.tiles_shape(), pshape = o.ptr.tiles_shape();
  std::vector<int64_t> idx, pidx(R), ords, exts, offs;
  int64_t off = 0;
  for (int64_t t = 0; t < n; ++t) {
    if (!o.tiles[t]) continue;
    ords.push_back(t);
    unravel(t, tshape, idx);
    int64_t vol = 1;
    for (int d = 0; d < R; ++d) { exts.push_back(o.tr.ext(d, idx[d])); vol *= exts.back(); }
    offs.push_back(off);
    off += (vol + 1) & ~(int64_t)1;
  }
  tadev_stream s = (tadev_stream)ctx->streams[0];
  int rc = tadev_alloc(ctx, (size_t)std::max<int64_t>(off, 2) * 8, (void**)tmp_arena, s);
  if (rc) return rc;
  std::vector<char> done(ords.size(), 0);
  std::vector<const void*> ins;
  std::vector<void*> outs;
  std::vector<int32_t> perm32(o.perm.begin(), o.perm.end());
  for (size_t t = 0; t < ords.size(); ++t) {
    unravel(ords[t], tshape, idx);
    for (int i = 0; i < R; ++i) pidx[o.perm[i]] = idx[i];
    by_target[ravel(pidx, pshape)] = *tmp_arena + offs[t];
    if (done[t]) continue;
    ins.clear(); outs.clear();
    for (size_t q = t; q < ords.size(); ++q) {
      if (done[q] || memcmp(&exts[q * R], &exts[t * R], sizeof(int64_t) * R) != 0) continue;
      ins.push_back(o.tiles[ords[q]]);
      outs.push_back(*tmp_arena + offs[q]);
      done[q] = 1;
    }
    rc = tadev_permute_batched(ctx, s, R, &exts[t * R], perm32.data(), 8, (int)ins.size(), ins.data(), outs.data());
    if (rc) return rc;
  }
  return TADEV_OK;
}

}  // namespace

extern "C" int tadev_elementwise_eval(tadev_elementwise* E, void* result_arena, float* ms_out) {
  TADEV_REQUIRE(E && result_arena, "tadev_elementwise_eval: null");
  tadev_ctx* ctx = E->ctx;
  tadev_stream s = (tadev_stream)ctx->streams[0];
  void *e0 = nullptr, *e1 = nullptr;
  tadev_event_create(ctx, &e0); tadev_event_create(ctx, &e1);
  tadev_event_record(ctx, e0, s);
  std::vector<const double*> ta, tb;
  double *tmpA = nullptr, *tmpB = nullptr;
  int rc = ew_operand_tiles(ctx, E->A, ta, &tmpA);
  if (!rc && E->has_b) rc = ew_operand_tiles(ctx, E->B, tb, &tmpB);
  if (!rc && !E->ords.empty()) {
    const size_t n = E->ords.size();
    std::vector<double*> out(n);
    std::vector<const double*> x(n), y(n, nullptr);
    double* arena = static_cast<double*>(result_arena);
    for (size_t t = 0; t < n; ++t) {
      out[t] = arena + E->offs[t];
      x[t] = ta[E->ords[t]];
      if (E->has_b) y[t] = tb[E->ords[t]];
    }
    rc = tadev_tiles_binary_f64(ctx, s, E->op, (int)n, out.data(), x.data(), y.data(), E->elems.data(), E->alpha, E->beta);
  }
  if (tmpA) tadev_free(ctx, tmpA, s);
  if (tmpB) tadev_free(ctx, tmpB, s);
  tadev_event_record(ctx, e1, s);
  float ms = 0;
  tadev_event_elapsed_ms(ctx, e0, e1, &ms);
  tadev_event_destroy(ctx, e0); tadev_event_destroy(ctx, e1);
  if (ms_out) *ms_out = ms;
  return rc;
}

extern "C" int tadev_elementwise_destroy(tadev_elementwise* E) {
  delete E;
  return TADEV_OK;
}
