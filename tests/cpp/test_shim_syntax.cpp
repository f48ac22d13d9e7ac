// Type-check of include/tadev_tiledarray_shim.hpp against the facsimile declarations of the reference interfaces
// (tests/cpp/ta_facsimile/README.md): every template of the shim is instantiated with mock policy / evaluator types.
//   g++ -std=c++17 -fsyntax-only -I include -I tests/cpp/ta_facsimile tests/cpp/test_shim_syntax.cpp
#include <cstddef>
#include <vector>

#include "tadev_tiledarray_shim.hpp"

namespace TiledArray {
template <typename Tile, typename Policy>
class DistArray {};
}  // namespace TiledArray

namespace mock {
struct Shape {
  static float threshold();
  bool is_dense() const;
  const std::vector<float>& data() const;
};
struct TRange { TiledArray::Range make_tile_range(std::size_t) const; };
struct Pmap {};
struct Policy {
  typedef std::size_t ordinal_type;
  typedef TRange trange_type;
  typedef Shape shape_type;
  typedef Pmap pmap_interface;
};
struct Arg {  // an argument evaluator: DistEval<...> (dist_eval.h:330)
  bool is_zero(std::size_t) const;
  bool is_local(std::size_t) const;
  madness::Future<TiledArray::tadevTile> get(std::size_t) const;
  const Shape& shape() const;
};
struct Op {  // ContractReduce (tile_op/contract_reduce.h:302)
  typedef TiledArray::tadevTile result_type;
  const TiledArray::math::GemmHelper& gemm_helper() const;
  double factor() const;
};
struct Grid { std::size_t rows() const; std::size_t cols() const; };
struct Archive {
  template <class T> const Archive& operator&(const T&) const;
  template <class T> const Archive& operator&(T&) const;
};
struct HostTensor {
  HostTensor();
  explicit HostTensor(const TiledArray::Range&);
  const TiledArray::Range& range() const;
  double* data();
  const double* data() const;
};
}  // namespace mock

template class TiledArray::detail::SummaTadev<mock::Arg, mock::Arg, mock::Op, mock::Policy>;

void instantiate_everything(const TiledArray::math::GemmHelper& h, const TiledArray::Permutation& p, mock::Archive& ar,
                            madness::World& world, const mock::TRange& tr, const mock::Shape& sh,
                            const std::shared_ptr<const mock::Pmap>& pmap, const mock::Arg& arg, const mock::Op& op,
                            const mock::Grid& grid) {
  using namespace TiledArray;
  static_assert(detail::is_device_tile<tadevTile>::value, "Tile<tadevTensor> must be a device tile");
  tadevTensor a, b;
  tadevTensor c = gemm(a, b, 2.0, h);
  gemm(c, a, b, 1, h);
  tadevTensor d = permute(c, p);
  tadevTensor e = clone(d);
  add_to(e, d);
  tadevTensor f = add(e, d), g = subt(e, d), m = mult(e, d), s = scale(e, 3.0), n = neg(e);
  scale_to(s, 0.5);
  std::vector<long> bs;
  tadevTensor sh2 = shift(e, bs);
  shift_to(sh2, bs);
  double x = squared_norm(e) + norm(e);
  (void)x; (void)empty(e);
  (void)f; (void)g; (void)m; (void)n;
  madness::archive::ArchiveStoreImpl<mock::Archive, tadevTensor>::store(ar, e);
  madness::archive::ArchiveLoadImpl<mock::Archive, tadevTensor>::load(ar, e);
  DistArray<tadevTile, mock::Policy> dev_arr;
  auto host_arr = to_host_array<mock::HostTensor>(dev_arr);
  auto dev_arr2 = to_device_array(host_arr);
  (void)dev_arr2;
  detail::SummaTadev<mock::Arg, mock::Arg, mock::Op, mock::Policy> summa(arg, arg, world, tr, sh, pmap, p, op, std::size_t(4), grid);
  (void)summa.get_tile(0);
}
