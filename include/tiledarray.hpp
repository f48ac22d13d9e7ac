// tiledarray.hpp — the TiledArray user API for the contraction path, in C++, over libtadev.
//
//   TA::TArrayD a(world, trange), b(world, trange), c;
//   a.fill(1.0); b.fill(1.0);
//   c("m,n") = a("m,k") * b("k,n");                    // examples/gemm/ta_dense.cpp:174
//   c("i,a,j,b") = 2.0 * (A("i,k,a,c") * B("j,c,k,b")); // permuted / scaled contractions
//
// The reference is a header-only C++ library on MADNESS; its own dependencies are not available to
// this repository, so this header restates the user-facing types of the path with the reference's
// names, argument meaning and error behaviour (TiledRange1 tiled_range1.h:47, TiledRange
// tiled_range.h:44, SparseShape<float> sparse_shape.h:77, DistArray dist_array.h:63, the
// expression objects tsr_expr.h:132 / mult_expr.h:192 / scal_expr.h, TiledArray::Exception
// error.h:39-83). Only metadata lives on the host: tiles are device memory (or pinned host memory,
// or lazy) and every operation is a call into the C ABI (tadev.h) — planning, screening, permutes
// and SUMMA all run in tadev_contraction_create / tadev_contraction_eval. One process drives one GPU
// (MADWorld's one-process-per-rank model); multi-GPU worlds use NCCL communicators (World::init_comm).
//
// Scope: contraction expressions only; element access is via host copies (find/set).
#pragma once
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "tadev.hpp"

namespace tadev {
namespace ta {

using Exception = ::tadev::Exception;
#define TA_TADEV_ASSERT(cond, msg) \
  do { if (!(cond)) throw ::tadev::Exception(TADEV_EINVAL, msg); } while (0)

// ---- TiledRange1 / TiledRange ---------------------------------------------------------------------
class TiledRange1 {
 public:
  TiledRange1() = default;
  TiledRange1(std::initializer_list<int64_t> bounds) : b_(bounds) { check_(); }
  template <typename It>
  TiledRange1(It first, It last) : b_(first, last) { check_(); }
  // make_uniform (tiled_range1.h:289-313): ceil(extent / tile) tiles, as uniform as possible (the first tiles
  // are one element larger), e.g. make_uniform(55, 10) == {0,10,19,28,37,46,55}
  static TiledRange1 make_uniform(int64_t extent, int64_t tile, int64_t lo = 0) {
    TA_TADEV_ASSERT(extent > 0 && tile > 0, "TiledRange1::make_uniform: positive extent and tile size required");
    const int64_t ntiles = (extent + tile - 1) / tile;
    const int64_t quot = (extent + ntiles - 1) / ntiles, rem = (extent + ntiles - 1) % ntiles;
    const int64_t avg = quot - 1, nplus = rem + 1;
    std::vector<int64_t> b;
    int64_t e = lo;
    for (int64_t i = 0; i < ntiles; ++i) { b.push_back(e); e += i < nplus ? avg + 1 : avg; }
    b.push_back(lo + extent);
    return TiledRange1(b.begin(), b.end());
  }
  int64_t tile_extent() const { return (int64_t)b_.size() - 1; }  // number of tiles (tiles_range().extent())
  int64_t ntiles() const { return tile_extent(); }
  int64_t extent() const { return b_.back() - b_.front(); }        // elements
  int64_t tile_size(int64_t t) const { return b_[t + 1] - b_[t]; }
  std::pair<int64_t, int64_t> tile(int64_t t) const { return {b_[t], b_[t + 1]}; }
  const std::vector<int64_t>& bounds() const { return b_; }
  bool operator==(const TiledRange1& o) const { return b_ == o.b_; }
  bool operator!=(const TiledRange1& o) const { return b_ != o.b_; }

 private:
  void check_() const {
    TA_TADEV_ASSERT(b_.size() >= 2, "TiledRange1: at least one tile is required");
    for (size_t i = 1; i < b_.size(); ++i) TA_TADEV_ASSERT(b_[i] > b_[i - 1], "TiledRange1: tile boundaries must be strictly increasing");
  }
  std::vector<int64_t> b_;
};

class TiledRange {
 public:
  TiledRange() = default;
  TiledRange(std::initializer_list<TiledRange1> dims) : d_(dims) {}
  template <typename It>
  TiledRange(It first, It last) : d_(first, last) {}
  unsigned rank() const { return (unsigned)d_.size(); }
  const TiledRange1& dim(unsigned i) const { return d_[i]; }
  const std::vector<TiledRange1>& data() const { return d_; }
  int64_t ntiles() const { int64_t n = 1; for (auto& d : d_) n *= d.ntiles(); return n; }   // tiles_range().volume()
  int64_t nelements() const { int64_t n = 1; for (auto& d : d_) n *= d.extent(); return n; }  // elements_range().volume()
  std::vector<int64_t> tiles_extent() const { std::vector<int64_t> s; for (auto& d : d_) s.push_back(d.ntiles()); return s; }
  std::vector<int64_t> elements_extent() const { std::vector<int64_t> s; for (auto& d : d_) s.push_back(d.extent()); return s; }
  std::vector<int64_t> tile_index(int64_t ord) const {
    std::vector<int64_t> idx(d_.size());
    for (int i = (int)d_.size() - 1; i >= 0; --i) { idx[i] = ord % d_[i].ntiles(); ord /= d_[i].ntiles(); }
    return idx;
  }
  int64_t tile_ordinal(const std::vector<int64_t>& idx) const {
    int64_t o = 0;
    for (size_t i = 0; i < d_.size(); ++i) o = o * d_[i].ntiles() + idx[i];
    return o;
  }
  // extents of tile `ord` (make_tile_range(ord).extent())
  Range tile_extent(int64_t ord) const {
    const auto idx = tile_index(ord);
    Range r(d_.size());
    for (size_t i = 0; i < d_.size(); ++i) r[i] = d_[i].tile_size(idx[i]);
    return r;
  }
  // first element of tile `ord` (make_tile_range(ord).lobound())
  std::vector<int64_t> tile_lobound(int64_t ord) const {
    const auto idx = tile_index(ord);
    std::vector<int64_t> r(d_.size());
    for (size_t i = 0; i < d_.size(); ++i) r[i] = d_[i].tile(idx[i]).first;
    return r;
  }
  bool operator==(const TiledRange& o) const { return d_ == o.d_; }
  bool operator!=(const TiledRange& o) const { return !(d_ == o.d_); }

 private:
  std::vector<TiledRange1> d_;
};

// ---- host Tensor (TA::Tensor<T>: shallow copy, row-major; tensor/tensor.h) -----------------------------
template <typename T>
class Tensor {
 public:
  using value_type = T;
  Tensor() = default;
  explicit Tensor(const Range& extent, T value = T()) : r_(extent), d_(std::make_shared<std::vector<T>>((size_t)volume(extent), value)) {}
  bool empty() const { return !d_; }
  const Range& range() const { return r_; }
  size_t size() const { return d_ ? d_->size() : 0; }
  T* data() { return d_->data(); }
  const T* data() const { return d_->data(); }
  T& operator[](size_t i) { return (*d_)[i]; }
  const T& operator[](size_t i) const { return (*d_)[i]; }
  T& operator()(const std::vector<int64_t>& idx) { return (*d_)[ord_(idx)]; }
  const T& operator()(const std::vector<int64_t>& idx) const { return (*d_)[ord_(idx)]; }
  Tensor clone() const { Tensor t; t.r_ = r_; if (d_) t.d_ = std::make_shared<std::vector<T>>(*d_); return t; }
  double norm() const { double s = 0; for (size_t i = 0; i < size(); ++i) s += (double)(*d_)[i] * (double)(*d_)[i]; return std::sqrt(s); }

 private:
  size_t ord_(const std::vector<int64_t>& idx) const { size_t o = 0; for (size_t i = 0; i < r_.size(); ++i) o = o * (size_t)r_[i] + (size_t)idx[i]; return o; }
  Range r_;
  std::shared_ptr<std::vector<T>> d_;
};

// a ready future (madness::Future analogue for find())
template <typename T>
class Future {
 public:
  explicit Future(T v) : v_(std::move(v)) {}
  const T& get() const { return v_; }
  bool probe() const { return true; }

 private:
  T v_;
};

// ---- World -----------------------------------------------------------------------------------------
// One process <-> one GPU. rank/size come from the launcher (TADEV_RANK / TADEV_SIZE or the usual
// RANK / WORLD_SIZE / LOCAL_RANK variables); size > 1 needs NCCL communicators (init_comm).
class World {
 public:
  World(int device, int rank, int size) : ctx_(new Context(device)), rank_(rank), size_(size) {}
  int rank() const { return rank_; }
  int size() const { return size_; }
  Context& context() { return *ctx_; }
  tadev_ctx* ctx() const { return ctx_->get(); }
  tadev_stream stream() const { return ctx_->stream_for(0); }
  void sync() const { ctx_->sync(); }
  std::pair<int, int> grid() const { return {pr_, pc_}; }

  tadev_proc_grid proc_grid(int64_t rows, int64_t cols, int64_t row_size, int64_t col_size) const {
    tadev_proc_grid g;
    check(tadev_proc_grid_make(rank_, size_, rows, cols, row_size, col_size, &g));
    return g;
  }
  // Build world/row/column communicators for a Pr x Pc grid. The 128-byte NCCL id is created by rank 0
  // and handed to the other ranks through `id_file` (a path on a filesystem all ranks see; default
  // $TADEV_ID_FILE) — the single-node stand-in for MADWorld's MPI bootstrap.
  void init_comm(int Pr, int Pc, std::string id_file = "") {
    pr_ = Pr; pc_ = Pc;
    if (size_ == 1) return;
    if (id_file.empty() && std::getenv("TADEV_ID_FILE")) id_file = std::getenv("TADEV_ID_FILE");
    TA_TADEV_ASSERT(!id_file.empty(), "World::init_comm: size > 1 needs an id file (TADEV_ID_FILE)");
    char id[128];
    if (rank_ == 0) {
      check(tadev_comm_unique_id(id));
      const std::string tmp = id_file + ".tmp";
      FILE* f = std::fopen(tmp.c_str(), "wb");
      TA_TADEV_ASSERT(f && std::fwrite(id, 1, 128, f) == 128, "World::init_comm: cannot write the id file");
      std::fclose(f);
      TA_TADEV_ASSERT(std::rename(tmp.c_str(), id_file.c_str()) == 0, "World::init_comm: cannot publish the id file");
    } else {
      FILE* f = nullptr;
      for (int tries = 0; tries < 600 && !(f = std::fopen(id_file.c_str(), "rb")); ++tries) usleep(100000);
      TA_TADEV_ASSERT(f && std::fread(id, 1, 128, f) == 128, "World::init_comm: cannot read the id file");
      std::fclose(f);
    }
    check(tadev_comm_init(ctx(), id, rank_, size_, Pr, Pc));
  }

 private:
  std::unique_ptr<Context> ctx_;
  int rank_, size_, pr_ = 1, pc_ = 1;
};

namespace detail {
inline std::unique_ptr<World>& default_world() { static std::unique_ptr<World> w; return w; }
inline int env_int(const char* a, const char* b, int dflt) {
  if (const char* e = std::getenv(a)) return std::atoi(e);
  if (const char* e = std::getenv(b)) return std::atoi(e);
  return dflt;
}
}  // namespace detail

// TA::initialize / finalize / get_default_world (tiledarray.cpp:84-165)
inline World& initialize(int& /*argc*/, char**& /*argv*/) {
  if (!detail::default_world()) {
    const int rank = detail::env_int("TADEV_RANK", "RANK", 0), size = detail::env_int("TADEV_SIZE", "WORLD_SIZE", 1);
    const int local = detail::env_int("TADEV_LOCAL_RANK", "LOCAL_RANK", rank);
    detail::default_world().reset(new World(local, rank, size));
  }
  return *detail::default_world();
}
inline World& get_default_world() {
  TA_TADEV_ASSERT(detail::default_world() != nullptr, "TiledArray has not been initialized");
  return *detail::default_world();
}
inline void finalize() { detail::default_world().reset(); }

// ---- shapes ----------------------------------------------------------------------------------------
struct DensePolicy {};
struct SparsePolicy {};

// SparseShape<float> (sparse_shape.h:77): per-tile Frobenius norms divided by the tile volume
// (scaled on the device, :149-217), hard-zeroed below the threshold.
template <typename T = float>
class SparseShape {
 public:
  SparseShape() = default;
  // tile_norms: unscaled Frobenius norms over the tile grid (sparse_shape.h:335-349 ctor)
  SparseShape(World& world, const Tensor<T>& tile_norms, const TiledRange& trange) : norms_(tile_norms.clone()) {
    TA_TADEV_ASSERT((int64_t)tile_norms.size() == trange.ntiles() && tile_norms.size() > 0, "SparseShape: norm tensor does not match trange");
    const unsigned dim = trange.rank();
    std::vector<float> left, right;
    if (dim == 1) { for (int64_t t = 0; t < trange.dim(0).ntiles(); ++t) left.push_back((float)trange.dim(0).tile_size(t)); }
    else {
      const unsigned middle = (dim >> 1) + (dim & 1);
      left = inv_outer_(trange, 0, middle);
      right = inv_outer_(trange, middle, dim);
    }
    tadev_ctx* ctx = world.ctx();
    tadev_stream s = world.stream();
    float *d_n = nullptr, *d_l = nullptr, *d_r = nullptr;
    uint64_t* d_z = nullptr;
    const size_t n = norms_.size();
    check(tadev_alloc(ctx, n * 4, (void**)&d_n, s));
    check(tadev_alloc(ctx, left.size() * 4, (void**)&d_l, s));
    check(tadev_alloc(ctx, std::max<size_t>(right.size(), 1) * 4, (void**)&d_r, s));
    check(tadev_alloc(ctx, 8, (void**)&d_z, s));
    check(tadev_memcpy_h2d(ctx, d_n, norms_.data(), n * 4, s));
    check(tadev_memcpy_h2d(ctx, d_l, left.data(), left.size() * 4, s));
    if (!right.empty()) check(tadev_memcpy_h2d(ctx, d_r, right.data(), right.size() * 4, s));
    check(tadev_memset(ctx, d_z, 0, 8, s));
    check(tadev_shape_scale_f32(ctx, s, d_n, d_l, (int64_t)left.size(), right.empty() ? nullptr : d_r, (int64_t)right.size(), threshold(), d_z));
    uint64_t nz = 0;
    check(tadev_memcpy_d2h(ctx, norms_.data(), d_n, n * 4, s));
    check(tadev_memcpy_d2h(ctx, &nz, d_z, 8, s));
    check(tadev_stream_sync(ctx, s));
    tadev_free(ctx, d_n, s); tadev_free(ctx, d_l, s); tadev_free(ctx, d_r, s); tadev_free(ctx, d_z, s);
    zero_tile_count_ = (int64_t)nz;
  }
  // already-scaled data (the result shape of a contraction)
  static SparseShape from_scaled(Tensor<T> norms, int64_t zero_tile_count) {
    SparseShape s; s.norms_ = std::move(norms); s.zero_tile_count_ = zero_tile_count; return s;
  }
  static float& threshold_ref() { static float t = 1.1920928955078125e-07f; return t; }
  static float threshold() { return threshold_ref(); }
  static void threshold(float t) { threshold_ref() = t; }
  bool empty() const { return norms_.empty(); }
  bool is_zero(int64_t ord) const { return norms_[(size_t)ord] < threshold(); }
  static constexpr bool is_dense() { return false; }
  float sparsity() const { return (float)zero_tile_count_ / (float)norms_.size(); }
  int64_t zero_tile_count() const { return zero_tile_count_; }
  const Tensor<T>& data() const { return norms_; }

 private:
  static std::vector<float> inv_outer_(const TiledRange& tr, unsigned lo, unsigned hi) {  // recursive_outer_product, :105-132
    const unsigned dim = hi - lo;
    if (dim == 1) {
      std::vector<float> v;
      for (int64_t t = 0; t < tr.dim(lo).ntiles(); ++t) v.push_back(1.0f / (float)tr.dim(lo).tile_size(t));
      return v;
    }
    const unsigned middle = (dim >> 1) + (dim & 1);
    const auto l = inv_outer_(tr, lo, lo + middle), r = inv_outer_(tr, lo + middle, hi);
    std::vector<float> out;
    for (float x : l) for (float y : r) out.push_back(x * y);
    return out;
  }
  Tensor<T> norms_;
  int64_t zero_tile_count_ = 0;
};

// ---- DistArray + expressions --------------------------------------------------------------------------
template <typename Tile, typename Policy> class DistArray;
template <typename A> struct MultExpr;
template <typename A> struct ScalMultExpr;

template <typename A> struct ScalTsrExpr;
template <typename A> struct AddExpr;

namespace detail {
inline std::vector<std::string> split_indices(const std::string& s) {
  std::vector<std::string> out;
  std::string cur;
  for (char ch : s) {
    if (ch == ',') { out.push_back(cur); cur.clear(); }
    else if (ch != ' ' && ch != '\t') cur.push_back(ch);
  }
  if (!cur.empty() || !out.empty()) out.push_back(cur);
  return out;
}
inline bool same_index_set(const std::string& a, const std::string& b) {
  auto x = split_indices(a), y = split_indices(b);
  std::sort(x.begin(), x.end());
  std::sort(y.begin(), y.end());
  return x == y;
}
}  // namespace detail

template <typename A>
struct TsrExpr {
  A* array;
  std::string idx;
  MultExpr<A> operator*(const TsrExpr& o) const { return MultExpr<A>{*this, o}; }
  // products: a contraction, or — when every index is shared and kept — the Hadamard product (mult_engine.h)
  void operator=(const MultExpr<A>& e) { assign_product(e.left, e.right, 1.0, false, e.mask); }
  void operator=(const ScalMultExpr<A>& e) { assign_product(e.expr.left, e.expr.right, e.factor, false, e.expr.mask); }
  void operator+=(const MultExpr<A>& e) { assign_product(e.left, e.right, 1.0, true, e.mask); }
  void operator+=(const ScalMultExpr<A>& e) { assign_product(e.expr.left, e.expr.right, e.factor, true, e.expr.mask); }
  // sums, scaling, copy / permutation (add_engine.h, subt_engine.h, scal_engine.h)
  void operator=(const AddExpr<A>& e) { array->assign_elementwise(idx, TADEV_EW_AXPBY, e.l.f, e.l.t, e.r.f, &e.r.t); }
  void operator=(const ScalTsrExpr<A>& e) { array->assign_elementwise(idx, TADEV_EW_AXPBY, e.f, e.t, 0.0, nullptr); }
  void operator=(const TsrExpr& e) { array->assign_elementwise(idx, TADEV_EW_AXPBY, 1.0, e, 0.0, nullptr); }

 private:
  void assign_product(const TsrExpr& l, const TsrExpr& r, double factor, bool accumulate, const typename A::shape_type* mask) {
    if (detail::same_index_set(l.idx, r.idx) && detail::same_index_set(l.idx, idx)) {
      TA_TADEV_ASSERT(!accumulate, "+= of a Hadamard product is not implemented");
      TA_TADEV_ASSERT(!mask, "set_shape is implemented for contractions only");
      array->assign_elementwise(idx, TADEV_EW_MULT, factor, l, 1.0, &r);
    } else array->assign_contraction(idx, l, r, factor, accumulate, mask);
  }
};
template <typename A>
struct MultExpr {
  TsrExpr<A> left, right;
  const typename A::shape_type* mask = nullptr;
  // (a("m,k") * b("k,n")).set_shape(shape): Expr::set_shape (expressions/expr.h:116) — the result shape is masked by
  // `shape` after ContEngine::make_shape (cont_engine.h:526-528); `shape` must outlive the assignment
  MultExpr& set_shape(const typename A::shape_type& shape) { mask = &shape; return *this; }
};
template <typename A>
struct ScalMultExpr { MultExpr<A> expr; double factor; };
template <typename A>
struct ScalTsrExpr { TsrExpr<A> t; double f; };
template <typename A>
struct AddExpr { ScalTsrExpr<A> l, r; };
template <typename A> ScalMultExpr<A> operator*(double f, const MultExpr<A>& e) { return {e, f}; }
template <typename A> ScalMultExpr<A> operator*(const MultExpr<A>& e, double f) { return {e, f}; }
template <typename A> ScalMultExpr<A> operator*(double f, const ScalMultExpr<A>& e) { return {e.expr, f * e.factor}; }
template <typename A> ScalMultExpr<A> operator-(const MultExpr<A>& e) { return {e, -1.0}; }
template <typename A> ScalTsrExpr<A> operator*(double f, const TsrExpr<A>& t) { return {t, f}; }
template <typename A> ScalTsrExpr<A> operator*(const TsrExpr<A>& t, double f) { return {t, f}; }
template <typename A> ScalTsrExpr<A> operator*(double f, const ScalTsrExpr<A>& t) { return {t.t, f * t.f}; }
template <typename A> ScalTsrExpr<A> operator-(const TsrExpr<A>& t) { return {t, -1.0}; }
template <typename A> AddExpr<A> operator+(const TsrExpr<A>& l, const TsrExpr<A>& r) { return {{l, 1.0}, {r, 1.0}}; }
template <typename A> AddExpr<A> operator-(const TsrExpr<A>& l, const TsrExpr<A>& r) { return {{l, 1.0}, {r, -1.0}}; }
template <typename A> AddExpr<A> operator+(const ScalTsrExpr<A>& l, const TsrExpr<A>& r) { return {l, {r, 1.0}}; }
template <typename A> AddExpr<A> operator-(const ScalTsrExpr<A>& l, const TsrExpr<A>& r) { return {l, {r, -1.0}}; }
template <typename A> AddExpr<A> operator+(const TsrExpr<A>& l, const ScalTsrExpr<A>& r) { return {{l, 1.0}, r}; }
template <typename A> AddExpr<A> operator-(const TsrExpr<A>& l, const ScalTsrExpr<A>& r) { return {{l, 1.0}, {r.t, -r.f}}; }
template <typename A> AddExpr<A> operator+(const ScalTsrExpr<A>& l, const ScalTsrExpr<A>& r) { return {l, r}; }
template <typename A> AddExpr<A> operator-(const ScalTsrExpr<A>& l, const ScalTsrExpr<A>& r) { return {l, {r.t, -r.f}}; }
template <typename A> AddExpr<A> operator*(double f, const AddExpr<A>& e) { return {{e.l.t, f * e.l.f}, {e.r.t, f * e.r.f}}; }
template <typename A> AddExpr<A> operator-(const AddExpr<A>& e) { return {{e.l.t, -e.l.f}, {e.r.t, -e.r.f}}; }

// knobs of the evaluator (TA_SUMMA_* analogues); see tadev_contract_options
struct ContractionOptions {
  static tadev_contract_options& get() {
    static tadev_contract_options o = [] { tadev_contract_options x; tadev_contract_options_default(&x); return x; }();
    return o;
  }
  static tadev_contract_stats& last_stats() { static tadev_contract_stats s{}; return s; }
};

template <typename Tile, typename Policy>
class DistArray {
  static_assert(std::is_same<typename Tile::value_type, double>::value, "the device path is FP64 (SURVEY §8)");

 public:
  using value_type = Tile;
  using shape_type = SparseShape<float>;
  static constexpr bool is_sparse = std::is_same<Policy, SparsePolicy>::value;
  enum Memory { Device = TADEV_MEM_DEVICE, Host = TADEV_MEM_HOST, Lazy = TADEV_MEM_LAZY };

  DistArray() = default;
  DistArray(World& world, const TiledRange& trange, Memory mem = Device) : st_(std::make_shared<State>()) {
    static_assert(!is_sparse || sizeof(Tile) == 0, "a sparse array needs a shape");
    st_->world = &world; st_->trange = trange; st_->mem = mem;
  }
  DistArray(World& world, const TiledRange& trange, const shape_type& shape, Memory mem = Device) : st_(std::make_shared<State>()) {
    static_assert(is_sparse, "only SparsePolicy arrays take a shape");
    TA_TADEV_ASSERT((int64_t)shape.data().size() == trange.ntiles(), "DistArray: shape does not match trange");
    st_->world = &world; st_->trange = trange; st_->shape = shape; st_->mem = mem;
  }
  // a lazy array (never stored): tiles are generated on the device, uniform(-1,1) keyed by (seed, tile ordinal)
  static DistArray make_lazy(World& world, const TiledRange& trange, uint64_t seed) {
    DistArray a(world, trange, Lazy);
    a.st_->lazy_seed = seed;
    return a;
  }

  bool is_initialized() const { return (bool)st_; }
  World& world() const { return *st_->world; }
  const TiledRange& trange() const { return st_->trange; }
  const shape_type& shape() const { return st_->shape; }
  int64_t size() const { return st_->trange.ntiles(); }
  bool is_zero(int64_t ord) const { return is_sparse && st_->shape.is_zero(ord); }
  bool is_dense() const { return !is_sparse; }
  // owner under the array's process map: single-rank arrays live on rank 0; contraction results follow
  // the cyclic map of the process grid (proc_grid.h:566-597)
  int owner(int64_t ord) const {
    if (st_->owner) return st_->owner(ord);
    if (st_->world->size() == 1) return 0;
    // default process map of a matrix in a multi-GPU world: cyclic over the process grid, the
    // distribution SUMMA consumes in place (make_row_phase_pmap / make_col_phase_pmap, proc_grid.h:566-597)
    TA_TADEV_ASSERT(st_->trange.rank() == 2, "multi-GPU worlds: only matrices have a default process map (no redistribution)");
    const auto g = st_->world->grid();
    const int64_t cols = st_->trange.dim(1).ntiles();
    return (int)(((ord / cols) % g.first) * g.second + (ord % cols) % g.second);
  }
  bool is_local(int64_t ord) const { return owner(ord) == st_->world->rank(); }

  // fill_local / fill (dist_array.h:983): every local non-zero tile gets `value`
  void fill(double value) {
    allocate_();
    std::vector<double> buf;
    for (int64_t o = 0; o < size(); ++o) {
      if (!st_->tiles[o]) continue;
      buf.assign((size_t)st_->elems[o], value);
      put_(o, buf.data());
    }
    world().sync();
  }
  // uniform(-1,1) from the counter RNG keyed by (seed, tile ordinal, offset): same data under any distribution
  void fill_random(uint64_t seed) {
    allocate_();
    TA_TADEV_ASSERT(st_->mem == Device, "fill_random: device-resident arrays only");
    for (int64_t o = 0; o < size(); ++o)
      if (st_->tiles[o]) check(tadev_fill_uniform_f64(world().ctx(), world().stream(), (double*)st_->tiles[o], (size_t)st_->elems[o], seed, (uint64_t)o << 32));
  }
  // set(ordinal, tile) (dist_array.h:937)
  void set(int64_t ord, const Tile& tile) {
    allocate_();
    TA_TADEV_ASSERT(ord >= 0 && ord < size() && is_local(ord), "DistArray::set: tile is not local");
    TA_TADEV_ASSERT(!is_zero(ord), "DistArray::set: tile is zero in the shape");
    TA_TADEV_ASSERT(tile.range() == st_->trange.tile_extent(ord), "DistArray::set: tile extent mismatch");
    put_(ord, tile.data());
    world().sync();
  }
  // init_tiles(op): tile = op(extent, lobound) for every local non-zero tile (dist_array.h:1117)
  template <typename Op>
  void init_tiles(Op&& op) {
    allocate_();
    for (int64_t o = 0; o < size(); ++o)
      if (st_->tiles[o]) { Tile t = op(st_->trange.tile_extent(o), st_->trange.tile_lobound(o)); put_(o, t.data()); }
    world().sync();
  }
  // find(ordinal): host copy of a local tile (dist_array.h:717)
  Future<Tile> find(int64_t ord) const {
    TA_TADEV_ASSERT(ord >= 0 && ord < size() && !is_zero(ord) && is_local(ord), "DistArray::find: tile is zero or not local");
    Tile t(st_->trange.tile_extent(ord));
    tadev_ctx* ctx = world().ctx();
    tadev_stream s = world().stream();
    if (st_->mem == Lazy) {
      double* tmp = nullptr;
      check(tadev_alloc(ctx, t.size() * 8, (void**)&tmp, s));
      check(tadev_fill_uniform_f64(ctx, s, tmp, t.size(), st_->lazy_seed, (uint64_t)ord << 32));
      check(tadev_memcpy_d2h(ctx, t.data(), tmp, t.size() * 8, s));
      check(tadev_stream_sync(ctx, s));
      tadev_free(ctx, tmp, s);
    } else {
      TA_TADEV_ASSERT(!st_->tiles.empty() && st_->tiles[ord], "DistArray::find: tile has not been set");
      if (st_->mem == Host) std::memcpy(t.data(), st_->tiles[ord], t.size() * 8);
      else { check(tadev_memcpy_d2h(ctx, t.data(), st_->tiles[ord], t.size() * 8, s)); check(tadev_stream_sync(ctx, s)); }
    }
    return Future<Tile>(std::move(t));
  }
  TsrExpr<DistArray> operator()(const std::string& idx) { return {this, idx}; }

  // c(target) = alpha * a(idx_a) [+ beta * b(idx_b) | .* b(idx_b)] — the element-wise engines
  void assign_elementwise(const std::string& target, int op, double alpha, const TsrExpr<DistArray>& a, double beta,
                          const TsrExpr<DistArray>* b) {
    DistArray& A = *a.array;
    TA_TADEV_ASSERT(A.st_ && (!b || b->array->st_), "element-wise expression: uninitialized argument");
    World& w = A.world();
    A.allocate_();
    if (b) b->array->allocate_();
    Desc da(A);
    std::unique_ptr<Desc> db(b ? new Desc(*b->array) : nullptr);
    tadev_elementwise* eng = nullptr;
    check(tadev_elementwise_create(w.ctx(), op, target.c_str(), alpha, a.idx.c_str(), &da.d, beta, b ? b->idx.c_str() : nullptr,
                                   b ? &db->d : nullptr, shape_type::threshold(), &eng));
    std::shared_ptr<tadev_elementwise> guard(eng, [](tadev_elementwise* e) { tadev_elementwise_destroy(e); });
    tadev_contraction_info info;
    check(tadev_elementwise_info_get(eng, &info));
    auto ns = adopt_structure_(w, info, Device);
    check(tadev_elementwise_eval(eng, ns->arena, nullptr));
    if (w.size() > 1) ns->owner = A.st_->owner;  // operands are in the target's order: same process map
    st_ = ns;  // the old tiles (possibly operands) are freed stream-ordered after the kernel
  }

  // Frobenius norms of the local tiles over the tile grid (zeros elsewhere), computed on the device
  Tensor<float> tile_norms() const {
    const State& s = *st_;
    TA_TADEV_ASSERT(s.mem == Device, "tile_norms: device-resident arrays only");
    Tensor<float> out(Range(s.trange.tiles_extent()), 0.0f);
    std::vector<const double*> ptrs;
    std::vector<int64_t> sizes, ords;
    for (int64_t o = 0; o < size(); ++o) if (!s.tiles.empty() && s.tiles[o]) { ptrs.push_back((const double*)s.tiles[o]); sizes.push_back(s.elems[o]); ords.push_back(o); }
    if (ptrs.empty()) return out;
    tadev_ctx* ctx = world().ctx();
    tadev_stream st = world().stream();
    void *d_p = nullptr, *d_s = nullptr, *d_o = nullptr;
    const size_t n = ptrs.size();
    check(tadev_alloc(ctx, n * 8, &d_p, st)); check(tadev_alloc(ctx, n * 8, &d_s, st)); check(tadev_alloc(ctx, n * 8, &d_o, st));
    check(tadev_memcpy_h2d(ctx, d_p, ptrs.data(), n * 8, st));
    check(tadev_memcpy_h2d(ctx, d_s, sizes.data(), n * 8, st));
    int64_t max_elems = 0;
    for (int64_t e : sizes) max_elems = std::max(max_elems, e);
    check(tadev_tile_sqnorms_f64(ctx, st, (int)n, (const double* const*)d_p, (const int64_t*)d_s, max_elems, (double*)d_o));
    std::vector<double> sq(n);
    check(tadev_memcpy_d2h(ctx, sq.data(), d_o, n * 8, st));
    check(tadev_stream_sync(ctx, st));
    tadev_free(ctx, d_p, st); tadev_free(ctx, d_s, st); tadev_free(ctx, d_o, st);
    for (size_t t = 0; t < n; ++t) out[(size_t)ords[t]] = (float)std::sqrt(sq[t]);
    return out;
  }
  // truncate (dist_array.h:1553, conversions/truncate.h): rebuild the shape from the true tile norms and drop
  // the tiles that fall below the threshold. Dense arrays are unchanged. (Single-rank worlds: a multi-rank
  // world needs the max-reduction of the norms over ranks, sparse_shape.h:416.)
  void truncate() {
    if (!is_sparse) return;
    TA_TADEV_ASSERT(world().size() == 1, "truncate: multi-rank worlds are not supported by the C++ layer yet");
    shape_type ns(world(), tile_norms(), st_->trange);
    for (int64_t o = 0; o < size(); ++o) if (ns.is_zero(o) && !st_->tiles.empty()) { st_->tiles[o] = nullptr; st_->elems[o] = 0; }
    st_->shape = ns;
  }

  // c(target) (+)= factor * left * right — ExprEngine hand-off (expr.h:378 eval_to)
  void assign_contraction(const std::string& target, const TsrExpr<DistArray>& l, const TsrExpr<DistArray>& r, double factor, bool accumulate,
                          const shape_type* mask = nullptr) {
    DistArray &A = *l.array, &B = *r.array;
    TA_TADEV_ASSERT(A.st_ && B.st_, "contraction: uninitialized argument");
    World& w = A.world();
    A.allocate_(); B.allocate_();
    Desc da(A), db(B);
    tadev_contraction* eng = nullptr;
    tadev_contract_options opt = ContractionOptions::get();
    if (mask) {
      TA_TADEV_ASSERT(is_sparse && !mask->empty(), "set_shape: a result mask needs the sparse policy and a non-empty shape");
      opt.mask_norms = mask->data().data();
      opt.mask_threshold = shape_type::threshold();
    }
    check(tadev_contraction_create(w.ctx(), target.c_str(), l.idx.c_str(), r.idx.c_str(), &da.d, &db.d, factor, &opt, &eng));
    std::shared_ptr<tadev_contraction> guard(eng, [](tadev_contraction* e) { tadev_contraction_destroy(e); });
    tadev_contraction_info info;
    check(tadev_contraction_info_get(eng, &info));
    const TiledRange tr = trange_of_(info);
    const Memory mem = st_ ? st_->mem : Device;
    TA_TADEV_ASSERT(mem != Lazy, "the result of a contraction cannot be a lazy array");
    std::shared_ptr<State> ns;
    if (accumulate) {
      // the product is added into the EXISTING tiles through their own pointers (whatever the arena layout is)
      TA_TADEV_ASSERT(st_ && st_->arena && st_->trange == tr, "c += a*b: the existing result must have the tiling of the product");
      std::vector<void*> tiles((size_t)info.nlocal, nullptr);
      for (int64_t t = 0; t < info.nlocal; ++t) {
        tiles[(size_t)t] = st_->tiles[info.ordinals[t]];
        TA_TADEV_ASSERT(tiles[(size_t)t] != nullptr, "c += a*b: the existing result lacks a tile of the product");
      }
      check(tadev_contraction_eval_tiles(eng, tiles.data(), mem, 1, &ContractionOptions::last_stats()));
      return;
    } else {
      TA_TADEV_ASSERT(!st_ || st_->trange.rank() == 0 || st_->trange == tr || st_->tiles.empty(), "result array tiling does not match the expression");
      if (this != &A && this != &B) st_.reset();  // give the old result back to the pool before allocating the new one
      ns = adopt_structure_(w, info, mem);
      ns->owner = [guard](int64_t ord) { int o = 0; check(tadev_contraction_owner(guard.get(), ord, &o)); return o; };
    }
    check(tadev_contraction_eval(eng, ns->arena, mem, 0, &ContractionOptions::last_stats()));
    st_ = ns;
  }

 private:
  struct State {
    World* world = nullptr;
    TiledRange trange;
    shape_type shape;
    Memory mem = Device;
    uint64_t lazy_seed = 0;
    void* arena = nullptr;
    int64_t arena_elems = 0;
    std::vector<void*> tiles;    // [ntiles] pointer into the arena; nullptr = zero / not local / not allocated
    std::vector<int64_t> elems;
    std::function<int(int64_t)> owner;
    void alloc_arena(int64_t n) {
      arena_elems = n;
      if (mem == Host) check(tadev_host_alloc((size_t)n * 8, &arena));
      else check(tadev_alloc(world->ctx(), (size_t)n * 8, &arena, world->stream()));
    }
    ~State() {
      if (!arena) return;
      if (mem == Host) tadev_host_free(arena);
      else tadev_free(world->ctx(), arena, world->stream());
    }
  };
  static TiledRange trange_of_(const tadev_contraction_info& info) {
    std::vector<TiledRange1> dims;
    size_t off = 0;
    for (int d = 0; d < info.rank; ++d) { dims.emplace_back(info.bounds + off, info.bounds + off + info.ntiles[d] + 1); off += (size_t)info.ntiles[d] + 1; }
    return TiledRange(dims.begin(), dims.end());
  }
  // a fresh State with the tiling, shape, arena and tile table an engine's info struct describes
  static std::shared_ptr<State> adopt_structure_(World& w, const tadev_contraction_info& info, Memory mem) {
    auto ns = std::make_shared<State>();
    ns->world = &w; ns->trange = trange_of_(info); ns->mem = mem;
    ns->alloc_arena(info.arena_elems);
    ns->tiles.assign((size_t)ns->trange.ntiles(), nullptr);
    ns->elems.assign((size_t)ns->trange.ntiles(), 0);
    for (int64_t t = 0; t < info.nlocal; ++t) {
      ns->tiles[info.ordinals[t]] = (char*)ns->arena + info.offsets[t] * 8;
      ns->elems[info.ordinals[t]] = info.elems[t];
    }
    if (info.norms) {
      Tensor<float> nt(Range(ns->trange.tiles_extent()));
      std::memcpy(nt.data(), info.norms, nt.size() * 4);
      ns->shape = shape_type::from_scaled(std::move(nt), (int64_t)info.nzero);
    }
    return ns;
  }
  struct Desc {  // tadev_array_desc of an array plus the storage it points to
    tadev_array_desc d{};
    std::vector<int64_t> bounds;
    std::vector<int32_t> ntiles;
    std::vector<const void*> table;
    std::vector<int32_t> owners;
    explicit Desc(const DistArray& a) {
      const State& s = *a.st_;
      for (auto& dim : s.trange.data()) { ntiles.push_back((int32_t)dim.ntiles()); bounds.insert(bounds.end(), dim.bounds().begin(), dim.bounds().end()); }
      d.rank = (int32_t)s.trange.rank();
      d.memory = (int32_t)s.mem;
      d.bounds = bounds.data(); d.ntiles = ntiles.data();
      d.norms = (is_sparse && !s.shape.empty()) ? s.shape.data().data() : nullptr;
      d.lazy_seed = s.lazy_seed;
      if (s.mem == Lazy) {
        table.assign((size_t)s.trange.ntiles(), nullptr);
        for (int64_t o = 0; o < s.trange.ntiles(); ++o) if (a.is_local(o) && !a.is_zero(o)) table[o] = (const void*)1;
      } else table.assign(s.tiles.begin(), s.tiles.end());
      d.tiles = table.data();
      if (s.world->size() > 1 && s.mem == Device) {  // the process map: lets the engine redistribute tiles
        for (int64_t o = 0; o < s.trange.ntiles(); ++o) owners.push_back((int32_t)a.owner(o));
        d.owners = owners.data();
      }
    }
  };
  void allocate_() {
    State& s = *st_;
    if (s.mem == Lazy || !s.tiles.empty()) return;
    const int64_t n = s.trange.ntiles();
    s.tiles.assign((size_t)n, nullptr);
    s.elems.assign((size_t)n, 0);
    std::vector<int64_t> offs((size_t)n, -1);
    int64_t off = 0;
    for (int64_t o = 0; o < n; ++o) {
      if (!is_local(o) || is_zero(o)) continue;
      s.elems[o] = volume(s.trange.tile_extent(o));
      offs[o] = off;
      off += (s.elems[o] + 1) & ~(int64_t)1;  // keep every tile 16-byte aligned
    }
    s.alloc_arena(std::max<int64_t>(off, 2));
    for (int64_t o = 0; o < n; ++o) if (offs[o] >= 0) s.tiles[o] = (char*)s.arena + offs[o] * 8;
  }
  void put_(int64_t ord, const double* src) {
    State& s = *st_;
    if (s.mem == Host) std::memcpy(s.tiles[ord], src, (size_t)s.elems[ord] * 8);
    else { check(tadev_memcpy_h2d(s.world->ctx(), s.tiles[ord], src, (size_t)s.elems[ord] * 8, s.world->stream())); check(tadev_stream_sync(s.world->ctx(), s.world->stream())); }
  }
  std::shared_ptr<State> st_;
};

using TArrayD = DistArray<Tensor<double>, DensePolicy>;
using TSpArrayD = DistArray<Tensor<double>, SparsePolicy>;

}  // namespace ta
}  // namespace tadev

#ifndef TADEV_NO_TA_ALIAS
namespace TA = tadev::ta;
namespace TiledArray = tadev::ta;
#endif
