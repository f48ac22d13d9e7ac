// facsimile of src/TiledArray/dist_eval/dist_eval.h:41-330 (DistEvalImpl) and tensor_impl.h (TensorImpl accessors)
#pragma once
#include <memory>
#include "TiledArray/external/madness.h"
#include "TiledArray/permutation.h"
namespace TiledArray { namespace detail {
template <typename Policy>
class TensorImpl {  // tensor_impl.h:38-200
 public:
  typedef typename Policy::ordinal_type ordinal_type;
  typedef typename Policy::trange_type trange_type;
  typedef typename Policy::shape_type shape_type;
  typedef typename Policy::pmap_interface pmap_interface;
  madness::World& world() const;                  // tensor_impl.h:88
  const trange_type& trange() const;              // :110
  const shape_type& shape() const;                // :116
  const std::shared_ptr<const pmap_interface>& pmap() const;  // :94
  bool is_zero(ordinal_type i) const;             // :152
  bool is_local(ordinal_type i) const;            // :140
  int owner(ordinal_type i) const;                // :128
  ordinal_type size() const;                      // :122
};
template <typename Tile, typename Policy>
class DistEvalImpl : public TensorImpl<Policy> {  // dist_eval.h:41
 public:
  typedef TensorImpl<Policy> TensorImpl_;
  typedef typename TensorImpl_::ordinal_type ordinal_type;
  typedef typename TensorImpl_::trange_type trange_type;
  typedef typename TensorImpl_::shape_type shape_type;
  typedef typename TensorImpl_::pmap_interface pmap_interface;
  typedef Tile value_type;
  DistEvalImpl(madness::World& world, const trange_type& trange, const shape_type& shape,
               const std::shared_ptr<const pmap_interface>& pmap, const Permutation& perm);  // dist_eval.h:116-140
  virtual ~DistEvalImpl();
  virtual madness::Future<value_type> get_tile(ordinal_type i) const = 0;  // :161
  virtual void discard_tile(ordinal_type i) const = 0;                     // :169
  void set_tile(ordinal_type i, const value_type& value);                  // :177-185
  ordinal_type perm_index_to_target(ordinal_type index) const;             // :104
 private:
  virtual int internal_eval() = 0;                                         // :245
};
}}  // namespace TiledArray::detail
