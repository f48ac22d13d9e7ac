// tadev_tiledarray_shim.hpp — the TiledArray-side binding of libtadev.
//
// Written against TiledArray's OWN headers: on a box that has TiledArray (+ MADNESS), including this header gives
//   * TiledArray::tadevTensor          a device tile type that satisfies the tile concept (tile_op/tile_interface.h)
//                                      the way the reference's btasUMTensorVarray does (device/btas_um_tensor.h:98-565)
//   * TiledArray::tadevTile            = TA::Tile<tadevTensor> (shallow-copy wrapper, tile.h:93), is_device_tile == true
//                                      so Summa routes its conversions through madness::add_device_task
//                                      (dist_eval/contraction_eval.h:598-619, reduce_task.h:661-673)
//   * madness::archive load / store    tiles travel between ranks as (range, host copy), btas_um_tensor.h:64-88
//   * to_host_array / to_device_array  TA::DistArray<Tensor<double>> <-> DistArray<tadevTile> (btas_um_tensor.h:619-745)
//   * TiledArray::detail::SummaTadev   a DistEvalImpl whose internal_eval() hands the whole contraction to
//                                      tadev_summa_f64 — constructed where ContEngine::make_dist_eval builds Summa
//                                      (expressions/cont_engine.h:662-677; ctor signature contraction_eval.h:1758-1805)
// so `c("m,n") = a("m,k") * b("k,n")` of an unmodified TiledArray application runs through the B200 engine either
// tile by tile (the ADL functions below) or as one SUMMA call (SummaTadev).
//
// This header cannot be compiled against the real TiledArray in the authoring container (MADNESS/Boost/Eigen/BTAS
// are absent); it is type-checked with -fsyntax-only against tests/cpp/ta_facsimile/, which restates only the
// declarations used here, each with its reference file:line (tests/test_host_logic.py runs the check).
#pragma once
#include <cstring>
#include <memory>
#include <type_traits>
#include <vector>

#include <TiledArray/conversions/to_new_tile_type.h>
#include <TiledArray/dist_eval/dist_eval.h>
#include <TiledArray/external/madness.h>
#include <TiledArray/math/gemm_helper.h>
#include <TiledArray/permutation.h>
#include <TiledArray/range.h>
#include <TiledArray/tile.h>

#include "tadev.hpp"

namespace TiledArray {

/// The device tile: TiledArray::Range + a tadev::Tile (shallow, ref-counted device storage, event-ordered streams).
class tadevTensor {
 public:
  typedef Range range_type;
  typedef double value_type;
  typedef double numeric_type;
  typedef double scalar_type;
  typedef std::size_t size_type;

  tadevTensor() = default;
  explicit tadevTensor(const Range& range) : range_(range), tile_(context(), extents_of(range), range.offset()) {
    std::vector<int64_t> lo(range.rank());
    for (unsigned d = 0; d < range.rank(); ++d) lo[d] = range.lobound_data()[d];
    tile_.shift_to(lo);
  }
  tadevTensor(const Range& range, tadev::Tile tile) : range_(range), tile_(std::move(tile)) {}

  const Range& range() const { return range_; }
  size_type size() const { return range_.volume(); }
  bool empty() const { return tile_.empty(); }
  double* data() { return tile_.data(); }
  const double* data() const { return tile_.data(); }
  tadev::Tile& tile() { return tile_; }
  const tadev::Tile& tile() const { return tile_; }

  /// deviceEnv::instance() analogue: one tadev context per process (external/device.h:536-605)
  static tadev::Context& context() {
    TADEV_ASSERT(tadev::Context::current() != nullptr, "tadevTensor: no tadev::Context installed (tadev::Context::current())");
    return *tadev::Context::current();
  }
  static tadev::Range extents_of(const Range& r) {
    tadev::Range e(r.rank());
    for (unsigned d = 0; d < r.rank(); ++d) e[d] = r.extent_data()[d];
    return e;
  }

 private:
  Range range_;
  tadev::Tile tile_;
};

typedef Tile<tadevTensor> tadevTile;

namespace detail {
template <>
struct is_device_tile<tadevTensor> : public std::true_type {};  // tensor/type_traits.h:415-418 picks it up for Tile<>
}  // namespace detail

namespace tadev_shim {
inline tadev::Op op_of(blas::Op op) {
  TADEV_ASSERT(op != blas::Op::ConjTrans, "tadevTensor: real tiles only");
  return op == blas::Op::NoTrans ? tadev::Op::NoTrans : tadev::Op::Trans;
}
// the stand-in helper of tadev.hpp built from TiledArray's (same ranks, same ops: math/gemm_helper.h:62-98)
inline tadev::GemmHelper helper_of(const math::GemmHelper& h, unsigned left_rank, unsigned right_rank) {
  const unsigned result_rank = left_rank + right_rank - 2u * h.num_contract_ranks();
  return tadev::GemmHelper(op_of(h.left_op()), op_of(h.right_op()), result_rank, left_rank, right_rank);
}
inline tadev::Permutation perm_of(const Permutation& p) {
  tadev::Permutation q(p.size());
  for (unsigned i = 0; i < p.size(); ++i) q[i] = (int32_t)p[i];  // both are image form (permutation.h:69-79)
  return q;
}
}  // namespace tadev_shim

// ---- the tile interface (found by ADL from tile_op/*, tile_interface/*) --------------------------------------------

/// gemm(left, right, factor, helper) — tile_interface.h:803-808; device reference btas_um_tensor.h:98-165
template <typename Scalar, typename = std::enable_if_t<detail::is_numeric_v<Scalar>>>
tadevTensor gemm(const tadevTensor& left, const tadevTensor& right, Scalar factor, const math::GemmHelper& helper) {
  TADEV_ASSERT(helper.left_right_congruent(left.range().extent_data(), right.range().extent_data()), "gemm: contracted ranges are not congruent");
  Range result_range = helper.make_result_range<Range>(left.range(), right.range());
  const tadev::GemmHelper h = tadev_shim::helper_of(helper, left.range().rank(), right.range().rank());
  tadev::this_task_ordinal() = result_range.offset();  // stream_for(result range), external/device.h:899-907
  return tadevTensor(result_range, tadev::gemm(left.tile(), right.tile(), factor, h));
}

/// gemm(result, left, right, factor, helper) — tile_interface.h:825-830 (accumulate)
template <typename Scalar, typename = std::enable_if_t<detail::is_numeric_v<Scalar>>>
tadevTensor& gemm(tadevTensor& result, const tadevTensor& left, const tadevTensor& right, Scalar factor,
                  const math::GemmHelper& helper) {
  if (result.empty()) { result = gemm(left, right, factor, helper); return result; }
  const tadev::GemmHelper h = tadev_shim::helper_of(helper, left.range().rank(), right.range().rank());
  tadev::gemm(result.tile(), left.tile(), right.tile(), factor, h);
  return result;
}

/// permute(arg, perm) — tile_interface/permute.h; device reference btas_um_tensor.h:169-191 (librett_permute)
inline tadevTensor permute(const tadevTensor& arg, const Permutation& perm) {
  Range result_range = perm * arg.range();
  tadev::this_task_ordinal() = result_range.offset();
  return tadevTensor(result_range, tadev::permute(arg.tile(), tadev_shim::perm_of(perm)));
}

inline tadevTensor clone(const tadevTensor& arg) { return tadevTensor(arg.range(), tadev::clone(arg.tile())); }

/// shift / shift_to — tile_interface/shift.h; btas_um_tensor.h:137-167
template <typename Index>
tadevTensor shift(const tadevTensor& arg, const Index& bound_shift) {
  Range r = arg.range();
  r.inplace_shift(bound_shift);
  return tadevTensor(r, tadev::clone(arg.tile()));
}
template <typename Index>
tadevTensor& shift_to(tadevTensor& arg, const Index& bound_shift) {
  const_cast<Range&>(arg.range()).inplace_shift(bound_shift);
  return arg;
}

/// add_to(result, arg) — the ContractReduce merge, tile_op/contract_reduce.h:397-398; btas_um_tensor.h:377-384
inline tadevTensor& add_to(tadevTensor& result, const tadevTensor& arg) {
  tadev::add_to(result.tile(), arg.tile());
  return result;
}
inline tadevTensor add(const tadevTensor& a, const tadevTensor& b) { return tadevTensor(a.range(), tadev::add(a.tile(), b.tile())); }
inline tadevTensor subt(const tadevTensor& a, const tadevTensor& b) { return tadevTensor(a.range(), tadev::subt(a.tile(), b.tile())); }
inline tadevTensor mult(const tadevTensor& a, const tadevTensor& b) { return tadevTensor(a.range(), tadev::mult(a.tile(), b.tile())); }
template <typename Scalar, typename = std::enable_if_t<detail::is_numeric_v<Scalar>>>
tadevTensor scale(const tadevTensor& arg, Scalar factor) { return tadevTensor(arg.range(), tadev::scale(arg.tile(), factor)); }
template <typename Scalar, typename = std::enable_if_t<detail::is_numeric_v<Scalar>>>
tadevTensor& scale_to(tadevTensor& arg, Scalar factor) { tadev::scale_to(arg.tile(), factor); return arg; }
inline tadevTensor neg(const tadevTensor& arg) { return scale(arg, -1.0); }

/// squared_norm / norm — used by DistArray::truncate and SparseShape construction (dist_array.h:1553)
inline double squared_norm(const tadevTensor& arg) { return tadev::squared_norm(arg.tile()); }
inline double norm(const tadevTensor& arg) { return tadev::norm(arg.tile()); }
inline bool empty(const tadevTensor& arg) { return arg.empty(); }

// ---- array conversions (btas_um_tensor.h:619-745) -------------------------------------------------------------------

/// DistArray<tadevTile> -> DistArray<HostTensor>: one asynchronous D2H copy per tile, completed by the device task
template <typename HostTensor, typename Policy>
auto to_host_array(const DistArray<tadevTile, Policy>& device_array) {
  return to_new_tile_type(device_array, [](const tadevTile& tile) {
    HostTensor result(tile.tensor().range());
    tile.tensor().tile().to_host(result.data());
    return result;
  });
}
/// DistArray<HostTensor> -> DistArray<tadevTile>
template <typename HostTensor, typename Policy>
auto to_device_array(const DistArray<HostTensor, Policy>& host_array) {
  return to_new_tile_type(host_array, [](const HostTensor& tile) {
    tadevTensor result(tile.range());
    result.tile().from_host(tile.data());
    return tadevTile(result);
  });
}

namespace detail {

/// The evaluator plug-in: a DistEvalImpl that evaluates the whole contraction with ONE call of the SUMMA driver
/// instead of the reference's task graph of StepTasks / ReducePairTasks (contraction_eval.h:1559-1731). Constructed
/// with Summa's own arguments (contraction_eval.h:1758-1805) where ContEngine::make_dist_eval builds Summa
/// (cont_engine.h:662-677). `Left` / `Right` are the argument evaluators (get(i) -> Future<tadevTile>, is_zero(i),
/// is_local(i), shape().data()); `Op` is the ContractReduce functor (gemm_helper(), factor(): contract_reduce.h:302).
template <typename Left, typename Right, typename Op, typename Policy>
class SummaTadev : public DistEvalImpl<tadevTile, Policy> {
 public:
  typedef DistEvalImpl<tadevTile, Policy> DistEvalImpl_;
  typedef typename DistEvalImpl_::ordinal_type ordinal_type;
  typedef typename DistEvalImpl_::trange_type trange_type;
  typedef typename DistEvalImpl_::shape_type shape_type;
  typedef typename DistEvalImpl_::pmap_interface pmap_interface;
  typedef tadevTile value_type;

  template <typename Perm, typename ProcGrid>
  SummaTadev(const Left& left, const Right& right, madness::World& world, const trange_type& trange, const shape_type& shape,
             const std::shared_ptr<const pmap_interface>& pmap, const Perm& perm, const Op& op, const ordinal_type k,
             const ProcGrid& proc_grid)
      : DistEvalImpl_(world, trange, shape, pmap, perm), left_(left), right_(right), op_(op), k_(k),
        rows_(proc_grid.rows()), cols_(proc_grid.cols()) {}

  madness::Future<value_type> get_tile(ordinal_type i) const override { return madness::Future<value_type>(tiles_[i]); }
  void discard_tile(ordinal_type i) const override { tiles_[i] = value_type(); }

 private:
  /// \return the number of tiles this rank sets (dist_eval.h:245)
  int internal_eval() override {
    const ordinal_type Mt = rows_, Nt = cols_, Kt = k_;
    // fused tile extents of the three matrices from the argument / result tiled ranges
    std::vector<int64_t> m_ext(Mt, 0), n_ext(Nt, 0), k_ext(Kt, 0);
    std::vector<const double*> a_tiles(Mt * Kt, nullptr), b_tiles(Kt * Nt, nullptr);
    std::vector<double*> c_tiles(Mt * Nt, nullptr);
    std::vector<value_type> keep_a(Mt * Kt), keep_b(Kt * Nt);
    for (ordinal_type i = 0; i < Mt; ++i)
      for (ordinal_type k = 0; k < Kt; ++k) {
        const ordinal_type o = i * Kt + k;
        if (left_.is_zero(o) || !left_.is_local(o)) continue;
        keep_a[o] = left_.get(o).get();  // (the argument evaluators have already permuted their tiles, array_eval.h:170)
        a_tiles[o] = keep_a[o].tensor().data();
        blas::integer m, n, kk;
        (void)n;
        if (!m_ext[i] || !k_ext[k]) { fused_extents(keep_a[o].tensor().range(), op_.gemm_helper(), true, m, kk); m_ext[i] = m; k_ext[k] = kk; }
      }
    for (ordinal_type k = 0; k < Kt; ++k)
      for (ordinal_type j = 0; j < Nt; ++j) {
        const ordinal_type o = k * Nt + j;
        if (right_.is_zero(o) || !right_.is_local(o)) continue;
        keep_b[o] = right_.get(o).get();
        b_tiles[o] = keep_b[o].tensor().data();
        blas::integer n, kk;
        if (!n_ext[j]) { fused_extents(keep_b[o].tensor().range(), op_.gemm_helper(), false, n, kk); n_ext[j] = n; if (!k_ext[k]) k_ext[k] = kk; }
      }
    // result tiles owned by this rank (result pmap = grid pmap, cont_engine.h:584)
    tiles_.assign(Mt * Nt, value_type());
    int nset = 0;
    for (ordinal_type o = 0; o < Mt * Nt; ++o) {
      if (this->is_zero(o) || !this->is_local(o)) continue;
      tadevTensor t(result_range(o));
      c_tiles[o] = t.data();
      tiles_[o] = value_type(t);
      ++nset;
    }
    tadev_summa_plan plan;
    std::memset(&plan, 0, sizeof(plan));
    plan.Mt = (int32_t)Mt; plan.Nt = (int32_t)Nt; plan.Kt = (int32_t)Kt;
    plan.m_ext = m_ext.data(); plan.n_ext = n_ext.data(); plan.k_ext = k_ext.data();
    plan.opA = (int)tadev_shim::op_of(op_.gemm_helper().left_op());
    plan.opB = (int)tadev_shim::op_of(op_.gemm_helper().right_op());
    plan.alpha = (double)op_.factor();
    plan.a_norms = sparse_norms(left_.shape());
    plan.b_norms = sparse_norms(right_.shape());
    plan.c_norms = sparse_norms(this->shape());
    plan.threshold = shape_type::threshold();
    plan.a_tiles = a_tiles.data(); plan.b_tiles = b_tiles.data(); plan.c_tiles = c_tiles.data();
    tadev_summa_stats stats;
    tadev::check(tadev_summa_f64(tadevTensor::context().get(), &plan, &stats));
    // hand the finished tiles to their consumers (finalize, contraction_eval.h:1180-1269); a result permutation, if
    // any, is applied by DistEvalImpl::set_tile's perm_index_to_target + the tile-level permute of the consumer
    for (ordinal_type o = 0; o < Mt * Nt; ++o)
      if (c_tiles[o]) DistEvalImpl_::set_tile(DistEvalImpl_::perm_index_to_target(o), tiles_[o]);
    return nset;
  }

  template <typename Shape>
  static const float* sparse_norms(const Shape& shape) {
    if constexpr (std::is_same<decltype(shape.is_dense()), bool>::value) { if (shape.is_dense()) return nullptr; }
    return shape.data().data();
  }
  static void fused_extents(const Range& r, const math::GemmHelper& h, bool is_left, blas::integer& outer, blas::integer& inner) {
    // math::GemmHelper::compute_matrix_sizes needs both ranges; one operand suffices for its own two extents
    const unsigned rank = r.rank(), nc = h.num_contract_ranks();
    const bool inner_last = is_left ? h.left_op() == blas::Op::NoTrans : h.right_op() != blas::Op::NoTrans;
    outer = inner = 1;
    for (unsigned d = 0; d < rank; ++d) {
      const bool is_inner = inner_last ? d >= rank - nc : d < nc;
      (is_inner ? inner : outer) *= r.extent_data()[d];
    }
  }
  Range result_range(ordinal_type o) const { return this->trange().make_tile_range(o); }

  Left left_;
  Right right_;
  Op op_;
  ordinal_type k_, rows_, cols_;
  mutable std::vector<value_type> tiles_;
};

}  // namespace detail
}  // namespace TiledArray

// ---- inter-rank transfer (madness::archive; reference device/btas_um_tensor.h:64-88) -------------------------------
namespace madness {
namespace archive {

template <class Archive>
struct ArchiveLoadImpl<Archive, TiledArray::tadevTensor> {
  static inline void load(const Archive& ar, TiledArray::tadevTensor& t) {
    TiledArray::Range range{};
    ar & range;
    std::vector<double> host(range.volume());
    ar & wrap(host.data(), host.size());
    t = TiledArray::tadevTensor(range);
    t.tile().from_host(host.data());
    t.tile().sync();  // `host` goes out of scope
  }
};

template <class Archive>
struct ArchiveStoreImpl<Archive, TiledArray::tadevTensor> {
  static inline void store(const Archive& ar, const TiledArray::tadevTensor& t) {
    std::vector<double> host(t.size());
    t.tile().to_host(host.data());  // blocks until the tile's last write has finished
    ar & t.range() & wrap(host.data(), host.size());
  }
};

}  // namespace archive
}  // namespace madness
