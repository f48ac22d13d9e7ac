// Host-only C++ tests of include/tiledarray.hpp (no GPU needed): the metadata classes against the
// reference's own known answers
//   tests/tiled_range1.cpp:76-130, 352-393   TiledRange1 constructor, accessors, make_uniform
//   tests/tiled_range.cpp                     tile ordinals / extents / lower bounds (row-major)
//   tests/general_product.cpp:96-135          GeneralPermutationOptimizer through the C ABI planner
// and that constructing a World without a CUDA device fails loudly (there is no CPU path).
#include <cstdio>
#include <cstring>
#include <string>

#include "tiledarray.hpp"

static int failures = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

using TA::TiledRange;
using TA::TiledRange1;

int main() {
  // ---- TiledRange1
  CHECK(TiledRange1::make_uniform(3, 10) == (TiledRange1{0, 3}));
  CHECK(TiledRange1::make_uniform(50, 10) == (TiledRange1{0, 10, 20, 30, 40, 50}));
  CHECK(TiledRange1::make_uniform(55, 10) == (TiledRange1{0, 10, 19, 28, 37, 46, 55}));
  CHECK(TiledRange1::make_uniform(59, 10) == (TiledRange1{0, 10, 20, 30, 40, 50, 59}));
  CHECK(TiledRange1::make_uniform(55, 10, 10) == (TiledRange1{10, 20, 29, 38, 47, 56, 65}));
  CHECK(TiledRange1::make_uniform(50, 30) == (TiledRange1{0, 25, 50}));
  const TiledRange1 t{0, 2, 5, 10, 17, 28};  // the reference fixture's tiling (tests/range_fixture.h)
  CHECK(t.ntiles() == 5 && t.extent() == 28 && t.tile_size(3) == 7);
  CHECK(t.tile(2).first == 5 && t.tile(2).second == 10);
  bool threw = false;
  try { TiledRange1 bad{0, 4, 4}; (void)bad; } catch (const TA::Exception&) { threw = true; }
  CHECK(threw);  // boundaries must strictly increase

  // ---- TiledRange: row-major tile ordinals, extents, lower bounds
  const TiledRange tr{t, TiledRange1{0, 3, 7}, TiledRange1{1, 4}};
  CHECK(tr.rank() == 3 && tr.ntiles() == 10 && tr.nelements() == 28 * 7 * 3);
  CHECK(tr.tile_ordinal({3, 1, 0}) == 7);
  CHECK((tr.tile_index(7) == std::vector<int64_t>{3, 1, 0}));
  CHECK((tr.tile_extent(7) == tadev::Range{7, 4, 3}));
  CHECK((tr.tile_lobound(7) == std::vector<int64_t>{10, 3, 1}));
  CHECK(tr == (TiledRange{t, TiledRange1{0, 3, 7}, TiledRange1{1, 4}}));
  CHECK(tr != (TiledRange{t, TiledRange1{0, 3, 7}}));

  // ---- host Tensor: shallow copy, row-major element access
  TA::Tensor<double> x(tadev::Range{2, 3}, 1.5), y = x;
  y({1, 2}) = 4.0;
  CHECK(x[5] == 4.0 && x.size() == 6);  // shallow copy shares storage (TA::Tensor semantics)
  TA::Tensor<double> z = x.clone();
  z[0] = -1.0;
  CHECK(x[0] == 1.5);

  // ---- planners through the C ABI (host-only entries)
  tadev_contraction_plan p;
  int32_t nf = -1, sw = -1;
  CHECK(tadev_plan_general_product("c,b,i,k", "b,c,i,j", "c,j,b,k", &p, &nf) == TADEV_OK && nf == 2);
  CHECK(std::string(p.left_target) == "c,b,i,j" && std::string(p.right_target) == "c,b,j,k" && std::string(p.result_gemm) == "c,b,i,k");
  CHECK(tadev_plan_general_product("b,i", "b,i,j", "b,i", &p, &nf) == TADEV_EINVAL);  // implicit reduction
  CHECK(tadev_plan_contraction_opt("a,b,i,j", "c,d,i,j", "a,b,c,d", &p, &sw) == TADEV_OK && sw == 1 && p.opA == 0 && p.opB == 0 && p.perm_result[0] == -1);
  tadev_proc_grid g;
  CHECK(tadev_proc_grid_make(0, 8, 32, 32, 32768, 32768, &g) == TADEV_OK && g.proc_rows == 4 && g.proc_cols == 2);

  // ---- no device, no engine: the library must refuse, not fall back
  int ndev = 0;
  tadev_device_count(&ndev);
  if (ndev == 0) {
    threw = false;
    try { TA::World w(0, 0, 1); (void)w; } catch (const TA::Exception& e) { threw = e.code == TADEV_ENODEVICE; }
    CHECK(threw);
  }
  std::printf(failures ? "HOST API TESTS FAILED (%d)\n" : "HOST API TESTS PASSED\n", failures);
  return failures ? 1 : 0;
}
