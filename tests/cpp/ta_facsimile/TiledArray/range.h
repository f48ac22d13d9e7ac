// facsimile (declarations only) of src/TiledArray/range.h:40-1300 — what the shim uses of TiledArray::Range
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
namespace TiledArray {
class Permutation;
class Range {
 public:
  typedef long index1_type;        // range.h:48
  typedef std::size_t ordinal_type;  // range.h:52
  Range() = default;
  template <typename Extents> explicit Range(const Extents& extents);             // range.h:480 (extents ctor)
  template <typename L, typename U> Range(const L& lobound, const U& upbound);    // range.h:402
  unsigned int rank() const;                       // range.h:733
  const index1_type* lobound_data() const;         // range.h:746
  const index1_type* upbound_data() const;         // range.h:788
  const index1_type* extent_data() const;          // range.h:830
  ordinal_type volume() const;                     // range.h:903
  ordinal_type offset() const;                     // range.h:913
  template <typename Index> Range& inplace_shift(const Index& bound_shift);  // range.h:1040
  template <typename Archive> void serialize(Archive& ar);                   // range.h:1211
};
bool operator==(const Range&, const Range&);      // range.h:1318
Range operator*(const Permutation&, const Range&);  // range.h:1303
}  // namespace TiledArray
