// C++ tests of the TiledArray-mirroring API (include/tiledarray.hpp) on one GPU. They restate, for
// the device engine, the reference's own tests of this path:
//   examples/gemm/ta_dense.cpp:162-174           fill(1) * fill(1)  =>  every element == N
//   tests/dist_eval_contraction_eval.cpp:293-372  dense SUMMA vs an explicit reference product (exact, integer data)
//   tests/dist_eval_contraction_eval.cpp:375-472  sparse: zero result tiles absent, non-zero tiles exact
//   tests/expressions_impl.h:1808-2675            cont / cont_permute / scale_cont / cont_non_uniform / outer product
//   tests/sparse_shape.cpp:82-154                 SparseShape ctor: norm / tile volume, hard zero below threshold
// Build: g++ -std=c++17 -I include tests/cpp/test_tiledarray_api.cpp -L tiledarray_b200 -ltadev
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tiledarray.hpp"

static int failures = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

using TA::TiledRange;
using TA::TiledRange1;
using Tensor = TA::Tensor<double>;

static unsigned lcg_state = 12345u;
static double small_int() { lcg_state = lcg_state * 1664525u + 1013904223u; return (double)((int)((lcg_state >> 10) % 9) - 4); }

// dense host image of an array (zeros where tiles are zero)
template <typename Array>
static std::vector<double> to_host(const Array& a) {
  const auto ext = a.trange().elements_extent();
  std::vector<int64_t> stride(ext.size(), 1);
  for (int d = (int)ext.size() - 2; d >= 0; --d) stride[d] = stride[d + 1] * ext[d + 1];
  std::vector<double> out((size_t)a.trange().nelements(), 0.0);
  for (int64_t o = 0; o < a.size(); ++o) {
    if (a.is_zero(o) || !a.is_local(o)) continue;
    const Tensor t = a.find(o).get();
    const auto lo = a.trange().tile_lobound(o);
    const auto& te = t.range();
    std::vector<int64_t> idx(te.size(), 0);
    for (size_t n = 0; n < t.size(); ++n) {
      size_t off = 0;
      for (size_t d = 0; d < te.size(); ++d) off += (size_t)((lo[d] - a.trange().dim((unsigned)d).bounds().front()) + idx[d]) * (size_t)stride[d];
      out[off] = t[n];
      for (int d = (int)te.size() - 1; d >= 0; --d) { if (++idx[d] < te[d]) break; idx[d] = 0; }
    }
  }
  return out;
}

template <typename Array>
static void fill_ints(Array& a, std::vector<double>* image = nullptr) {
  a.init_tiles([](const tadev::Range& ext, const std::vector<int64_t>&) {
    Tensor t(ext);
    for (size_t i = 0; i < t.size(); ++i) t[i] = small_int();
    return t;
  });
  if (image) *image = to_host(a);
}

int main(int argc, char** argv) {
  TA::World& world = TA::initialize(argc, argv);
  world.init_comm(1, 1);

  {  // ta_dense: c("m,n") = a("m,k") * b("k,n") with constant tiles
    const TiledRange1 t = TiledRange1::make_uniform(512, 128);
    const TiledRange tr{t, t};
    TA::TArrayD a(world, tr), b(world, tr), c;
    a.fill(1.0);
    b.fill(0.5);
    c("m,n") = a("m,k") * b("k,n");
    const auto C = to_host(c);
    bool ok = C.size() == 512u * 512u;
    for (double v : C) ok = ok && v == 256.0;
    CHECK(ok);
    CHECK(c.trange() == tr);
    c("m,n") += a("m,k") * b("k,n");  // accumulate into the existing result
    CHECK(to_host(c)[777] == 512.0);
    c("m,n") = -2.0 * (a("m,k") * b("k,n"));  // scale_cont
    CHECK(to_host(c)[4242] == -512.0);
  }

  {  // non-uniform tiling, integer data, all transpose forms + result permutation, vs a host triple loop
    const TiledRange1 dm{0, 4, 10, 16}, dk{0, 6, 8}, dn{0, 2, 12, 20, 24};
    const int M = 16, K = 8, N = 24;
    TA::TArrayD a(world, TiledRange{dm, dk}), at(world, TiledRange{dk, dm}), b(world, TiledRange{dk, dn}), bt(world, TiledRange{dn, dk});
    std::vector<double> A, AT, B, BT;
    fill_ints(a, &A); fill_ints(at, &AT); fill_ints(b, &B); fill_ints(bt, &BT);
    auto ref = [&](const std::vector<double>& x, bool xt, const std::vector<double>& y, bool yt, bool res_t) {
      std::vector<double> r((size_t)M * N, 0.0);
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
          double s = 0;
          for (int k = 0; k < K; ++k) s += (xt ? x[(size_t)k * M + i] : x[(size_t)i * K + k]) * (yt ? y[(size_t)j * K + k] : y[(size_t)k * N + j]);
          r[res_t ? (size_t)j * M + i : (size_t)i * N + j] = s;
        }
      return r;
    };
    TA::TArrayD c;
    c("m,n") = a("m,k") * b("k,n");    CHECK(to_host(c) == ref(A, false, B, false, false));
    c("m,n") = at("k,m") * b("k,n");   CHECK(to_host(c) == ref(AT, true, B, false, false));
    c("m,n") = a("m,k") * bt("n,k");   CHECK(to_host(c) == ref(A, false, BT, true, false));
    c("m,n") = at("k,m") * bt("n,k");  CHECK(to_host(c) == ref(AT, true, BT, true, false));
    TA::TArrayD ct;
    ct("n,m") = a("m,k") * b("k,n");   CHECK(to_host(ct) == ref(A, false, B, false, true));
    CHECK(ct.trange() == (TiledRange{dn, dm}));
    bool threw = false;
    try { TA::TArrayD bad; bad("m,n") = a("m,k") * bt("k,n"); } catch (const TA::Exception&) { threw = true; }
    CHECK(threw);  // inner tilings not congruent
  }

  {  // 4-index permuted contraction (BASELINE config 5 in miniature): C(i,a,j,b) = A(i,k,a,c) * B(j,c,k,b)
    const TiledRange1 s{0, 2, 6}, v{0, 4, 8, 10};
    const int S = 6, V = 10;
    TA::TArrayD a(world, TiledRange{s, s, v, v}), b(world, TiledRange{s, v, s, v}), c;
    std::vector<double> A, B;
    fill_ints(a, &A); fill_ints(b, &B);
    for (int pass = 0; pass < 2; ++pass) {
      TA::ContractionOptions::get().stream_permutes = pass;  // up-front permuted copies, then just-in-time permutes
      c("i,a,j,b") = a("i,k,a,c") * b("j,c,k,b");
      const auto C = to_host(c);
      bool ok = true;
      for (int i = 0; i < S && ok; ++i) for (int x = 0; x < V; ++x) for (int j = 0; j < S; ++j) for (int y = 0; y < V; ++y) {
        double sum = 0;
        for (int k = 0; k < S; ++k) for (int z = 0; z < V; ++z)
          sum += A[(((size_t)i * S + k) * V + x) * V + z] * B[(((size_t)j * V + z) * S + k) * V + y];
        ok = ok && C[(((size_t)i * V + x) * S + j) * V + y] == sum;
      }
      CHECK(ok);
    }
    TA::ContractionOptions::get().stream_permutes = -1;
  }

  {  // sparse: shapes from true tile norms, result shape by device screening, zero tiles absent
    const TiledRange1 d = TiledRange1::make_uniform(48, 8);
    const TiledRange tr{d, d};
    auto make = [&](unsigned seed, std::vector<double>& image) {
      lcg_state = seed;
      std::vector<Tensor> tiles;
      TA::Tensor<float> norms(tadev::Range{6, 6}, 0.0f);
      for (int64_t o = 0; o < 36; ++o) {
        Tensor t(tr.tile_extent(o));
        const bool keep = (small_int() > -1.0);
        for (size_t i = 0; i < t.size(); ++i) t[i] = keep ? small_int() : 0.0;
        norms[(size_t)o] = (float)t.norm();
        tiles.push_back(t);
      }
      TA::SparseShape<float> shape(world, norms, tr);
      TA::TSpArrayD arr(world, tr, shape);
      for (int64_t o = 0; o < 36; ++o) if (!arr.is_zero(o)) arr.set(o, tiles[(size_t)o]);
      image = to_host(arr);
      // SparseShape ctor: scaled norm = ||tile|| / volume (tests/sparse_shape.cpp:82-154)
      for (int64_t o = 0; o < 36; ++o) {
        const float want = norms[(size_t)o] / 64.0f;
        CHECK(shape.data()[(size_t)o] == (want < TA::SparseShape<float>::threshold() ? 0.0f : want));
      }
      return arr;
    };
    std::vector<double> A, B;
    TA::TSpArrayD a = make(7u, A), b = make(99u, B), c;
    c("m,n") = a("m,k") * b("k,n");
    const auto C = to_host(c);
    bool ok = true;
    for (int i = 0; i < 48; ++i) for (int j = 0; j < 48; ++j) {
      double s = 0;
      for (int k = 0; k < 48; ++k) s += A[(size_t)i * 48 + k] * B[(size_t)k * 48 + j];
      ok = ok && C[(size_t)i * 48 + j] == s;  // zero result tiles read back as zeros and the reference block is zero too
    }
    CHECK(ok);
    CHECK(c.shape().zero_tile_count() >= 0 && c.shape().sparsity() < 1.0f);
    const auto& st = TA::ContractionOptions::last_stats();
    CHECK(st.summa.npairs > 0 && st.summa.npairs < 216);
  }

  {  // lazy operand + outer product + matrix-vector
    const TiledRange1 d3{0, 3, 5}, d2{0, 2, 6, 7};
    TA::TArrayD x(world, TiledRange{d3}), y(world, TiledRange{d2}), m(world, TiledRange{d2, d3}), o, mv;
    std::vector<double> X, Y, Mx;
    fill_ints(x, &X); fill_ints(y, &Y); fill_ints(m, &Mx);
    o("a,b") = x("a") * y("b");
    const auto O = to_host(o);
    bool ok = true;
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 7; ++j) ok = ok && O[(size_t)i * 7 + j] == X[i] * Y[j];
    CHECK(ok);
    mv("i") = m("i,k") * x("k");
    const auto MV = to_host(mv);
    ok = true;
    for (int i = 0; i < 7; ++i) { double s = 0; for (int k = 0; k < 5; ++k) s += Mx[(size_t)i * 5 + k] * X[k]; ok = ok && MV[i] == s; }
    CHECK(ok);
    const TiledRange1 t = TiledRange1::make_uniform(64, 16);
    TA::TArrayD lz = TA::TArrayD::make_lazy(world, TiledRange{t, t}, 42), mat(world, TiledRange{t, t}), p1, p2;
    mat.fill_random(42);  // same generator: the lazy array and the materialised one hold the same values
    TA::TArrayD id(world, TiledRange{t, t});
    id.init_tiles([](const tadev::Range& ext, const std::vector<int64_t>& lo) {
      Tensor tl(ext, 0.0);
      for (int64_t i = 0; i < ext[0]; ++i) for (int64_t j = 0; j < ext[1]; ++j) if (lo[0] + i == lo[1] + j) tl[(size_t)(i * ext[1] + j)] = 1.0;
      return tl;
    });
    p1("m,n") = lz("m,k") * id("k,n");
    p2("m,n") = mat("m,k") * id("k,n");
    CHECK(to_host(p1) == to_host(p2));
    CHECK(TA::ContractionOptions::last_stats().summa.lazy_tiles == 0);
  }

  {  // element-wise expressions (tests/expressions_impl.h add / subt / scale / mult / permute blocks) + truncate
    const TiledRange1 d1{0, 3, 10, 16}, d2{0, 5, 8, 21};
    TA::TArrayD a(world, TiledRange{d1, d2}), b(world, TiledRange{d1, d2}), bt(world, TiledRange{d2, d1}), c, ct;
    std::vector<double> A, B, BT;
    fill_ints(a, &A); fill_ints(b, &B); fill_ints(bt, &BT);
    const int R = 16, Cn = 21;
    auto check_all = [&](const TA::TArrayD& arr, auto&& f) {
      const auto X = to_host(arr);
      bool ok = X.size() == (size_t)R * Cn;
      for (int i = 0; i < R && ok; ++i) for (int j = 0; j < Cn; ++j) ok = ok && X[(size_t)i * Cn + j] == f(i, j);
      return ok;
    };
    c("i,j") = a("i,j") + b("i,j");
    CHECK(check_all(c, [&](int i, int j) { return A[i * Cn + j] + B[i * Cn + j]; }));
    c("i,j") = a("i,j") - b("i,j");
    CHECK(check_all(c, [&](int i, int j) { return A[i * Cn + j] - B[i * Cn + j]; }));
    c("i,j") = 2.0 * a("i,j") + 3.0 * bt("j,i");
    CHECK(check_all(c, [&](int i, int j) { return 2 * A[i * Cn + j] + 3 * BT[j * R + i]; }));
    c("i,j") = -(a("i,j") - 2.0 * b("i,j"));
    CHECK(check_all(c, [&](int i, int j) { return -(A[i * Cn + j] - 2 * B[i * Cn + j]); }));
    c("i,j") = a("i,j") * b("i,j");  // Hadamard
    CHECK(check_all(c, [&](int i, int j) { return A[i * Cn + j] * B[i * Cn + j]; }));
    c("i,j") = 0.5 * (a("i,j") * bt("j,i"));
    CHECK(check_all(c, [&](int i, int j) { return 0.5 * A[i * Cn + j] * BT[j * R + i]; }));
    c("i,j") = 4.0 * a("i,j");
    CHECK(check_all(c, [&](int i, int j) { return 4 * A[i * Cn + j]; }));
    c("i,j") = c("i,j") + a("i,j");  // the result is also an operand
    CHECK(check_all(c, [&](int i, int j) { return 5 * A[i * Cn + j]; }));
    ct("j,i") = a("i,j");  // pure permutation
    const auto CT = to_host(ct);
    bool ok = true;
    for (int i = 0; i < R; ++i) for (int j = 0; j < Cn; ++j) ok = ok && CT[(size_t)j * R + i] == A[i * Cn + j];
    CHECK(ok);

    // sparse: a - a keeps a's tiles in the estimated shape; truncate() finds them all zero
    const TiledRange1 d = TiledRange1::make_uniform(32, 8);
    const TiledRange tr{d, d};
    TA::Tensor<float> norms(tadev::Range{4, 4}, 0.0f);
    std::vector<Tensor> tiles;
    lcg_state = 4242u;
    for (int64_t o = 0; o < 16; ++o) {
      Tensor t(tr.tile_extent(o));
      const bool keep = (o % 3) != 1;
      for (size_t i = 0; i < t.size(); ++i) t[i] = keep ? small_int() + 5.0 : 0.0;
      norms[(size_t)o] = (float)t.norm();
      tiles.push_back(t);
    }
    TA::TSpArrayD sa(world, tr, TA::SparseShape<float>(world, norms, tr)), sc;
    for (int64_t o = 0; o < 16; ++o) if (!sa.is_zero(o)) sa.set(o, tiles[(size_t)o]);
    sc("i,j") = sa("i,j") - sa("i,j");
    CHECK(sc.shape().zero_tile_count() == sa.shape().zero_tile_count());
    bool allzero = true;
    for (double v : to_host(sc)) allzero = allzero && v == 0.0;
    CHECK(allzero);
    sc.truncate();
    CHECK(sc.shape().zero_tile_count() == 16);
    sc("i,j") = 2.0 * sa("i,j");
    sc.truncate();  // true norms of 2a == 2 * norms of a: same zero pattern
    for (int64_t o = 0; o < 16; ++o) CHECK(sc.is_zero(o) == sa.is_zero(o));
    const auto S2 = to_host(sc), S1 = to_host(sa);
    ok = true;
    for (size_t i = 0; i < S1.size(); ++i) ok = ok && S2[i] == 2 * S1[i];
    CHECK(ok);

    // c("m,n") = (a("m,k") * b("k,n")).set_shape(mask): the expression of examples/gemm/ta_sparse.cpp:190. Result tiles
    // the mask calls zero are absent; the others equal the plain product; += then adds into a DIFFERENT array (its own
    // tile pointers: an array whose arena was laid out by another expression)
    TA::TSpArrayD prod, masked;
    prod("m,n") = sa("m,k") * sa("k,n");
    TA::Tensor<float> mnorms(tadev::Range{4, 4}, 0.0f);
    for (int64_t o = 0; o < 16; ++o) mnorms[(size_t)o] = (o % 2) ? 7.0f : 0.0f;
    const TA::SparseShape<float> mask(world, mnorms, tr);
    masked("m,n") = (sa("m,k") * sa("k,n")).set_shape(mask);
    const auto P = to_host(prod), M = to_host(masked);
    ok = true;
    for (int64_t o = 0; o < 16; ++o) {
      CHECK(masked.is_zero(o) == (prod.is_zero(o) || mask.is_zero(o)));
      const auto ext = tr.tile_extent(o);
      const int64_t ti = o / 4, tj = o % 4;
      for (int64_t x = 0; x < ext[0]; ++x)
        for (int64_t y = 0; y < ext[1]; ++y) {
          const size_t at = (size_t)((ti * 8 + x) * 32 + tj * 8 + y);
          ok = ok && M[at] == (masked.is_zero(o) ? 0.0 : P[at]);
        }
    }
    CHECK(ok);
    TA::TSpArrayD acc;
    acc("n,m") = 1.0 * sa("m,n");                       // laid out by a permuting element-wise expression
    acc("n,m") += (sa("m,k") * sa("k,n")).set_shape(acc.shape());  // masked by its own shape: every product tile exists in acc
    const auto ACC = to_host(acc), SA = to_host(sa);
    ok = true;
    for (int i = 0; i < 32; ++i)
      for (int j = 0; j < 32; ++j) {
        const int64_t o = (int64_t)(j / 8) * 4 + i / 8;  // tile of acc (n = j, m = i)
        const double want = acc.is_zero(o) ? 0.0 : SA[(size_t)i * 32 + j] + P[(size_t)i * 32 + j];
        ok = ok && ACC[(size_t)j * 32 + i] == want;
      }
    CHECK(ok);
  }

  TA::finalize();
  if (failures == 0) std::printf("ALL TILEDARRAY API TESTS PASSED\n");
  return failures == 0 ? 0 : 1;
}
