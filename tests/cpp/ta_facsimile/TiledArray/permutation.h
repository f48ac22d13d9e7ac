// facsimile of src/TiledArray/permutation.h:80-500 — image-form permutation (:69-79: result[perm[i]] = arg[i])
#pragma once
#include <vector>
namespace TiledArray {
class Permutation {
 public:
  typedef unsigned int index_type;           // permutation.h:85
  index_type size() const;                   // permutation.h:243
  index_type operator[](unsigned int i) const;  // permutation.h:276
  const std::vector<index_type>& data() const;  // permutation.h:429
  explicit operator bool() const;            // permutation.h:416
};
}  // namespace TiledArray
