// facsimile of src/TiledArray/conversions/to_new_tile_type.h and dist_array.h (DistArray forward)
#pragma once
#include <utility>
namespace TiledArray {
template <typename Tile, typename Policy> class DistArray;  // dist_array.h:63
// to_new_tile_type.h:80 — (an empty body only because a function template instantiated with a local lambda type must be
// defined in the translation unit that uses it; the real one converts every tile with `op`)
template <typename Tile, typename Policy, typename Op>
auto to_new_tile_type(const DistArray<Tile, Policy>&, Op&& op) -> DistArray<decltype(op(std::declval<const Tile&>())), Policy> {
  return {};
}
}  // namespace TiledArray
