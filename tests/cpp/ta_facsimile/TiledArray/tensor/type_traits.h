// facsimile of src/TiledArray/tensor/type_traits.h:410-425 (is_device_tile) and tile.h (Tile<T> forward)
#pragma once
#include <type_traits>
namespace TiledArray {
template <typename T> class Tile;                                             // tile.h:93
namespace detail {
template <typename T> struct is_device_tile : public std::false_type {};       // type_traits.h:415
template <typename T> struct is_device_tile<Tile<T>> : public is_device_tile<T> {};  // :418
template <typename T> constexpr bool is_numeric_v = std::is_arithmetic<T>::value;    // type_traits.h (is_numeric)
}}  // namespace TiledArray::detail
