// facsimile of src/TiledArray/tile.h:93-700 — the shallow-copy tile wrapper; only what the shim needs
#pragma once
#include <memory>
#include "TiledArray/tensor/type_traits.h"
namespace TiledArray {
template <typename T>
class Tile {
 public:
  typedef T tensor_type;                      // tile.h:98
  Tile() = default;
  explicit Tile(const tensor_type& tensor);   // tile.h:137
  tensor_type& tensor();                      // tile.h:213
  const tensor_type& tensor() const;          // tile.h:215
  decltype(auto) range() const;               // tile.h:260
  bool empty() const;                         // tile.h:305
};
}  // namespace TiledArray
