// C++ test of the tile plug-in (include/tadev.hpp) against known answers restated from the
// reference's own tests:
//   tests/tile_op_contract_reduce.cpp:109-220  integer tiles, factor 3, NN/TN/NT/TT, accumulate
//   tests/librett.cpp:594-661                   rank-4 permutations {0,3,2,1} and {1,0,3,2}
//   contract_reduce.h:397-398                   add_to merge
// Build: g++ -std=c++17 -I include tests/cpp/test_tile_plugin.cpp -L tiledarray_b200 -ltadev
// Runs on a GPU box only (tests/test_gpu_cpp_plugin.py).
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "tadev.hpp"

static int failures = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

using tadev::GemmHelper;
using tadev::Op;
using tadev::Tile;

static std::vector<double> int_matrix(int rows, int cols, unsigned seed) {
  std::vector<double> v((size_t)rows * cols);
  unsigned s = seed;
  for (auto& x : v) { s = s * 1664525u + 1013904223u; x = (double)((s >> 8) % 101); }
  return v;
}
static std::vector<double> transpose(const std::vector<double>& a, int rows, int cols) {
  std::vector<double> t(a.size());
  for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) t[(size_t)j * rows + i] = a[(size_t)i * cols + j];
  return t;
}

int main() {
  tadev::Context ctx(0);
  const int m = 18, n = 36, k = 27;
  const auto A = int_matrix(m, k, 1), B = int_matrix(k, n, 2);
  const auto AT = transpose(A, m, k), BT = transpose(B, k, n);
  std::vector<double> ref((size_t)m * n, 0.0);
  for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int x = 0; x < k; ++x) s += A[(size_t)i * k + x] * B[(size_t)x * n + j]; ref[(size_t)i * n + j] = 3 * s; }

  for (int ta = 0; ta < 2; ++ta)
    for (int tb = 0; tb < 2; ++tb) {
      Tile left(ctx, ta ? tadev::Range{k, m} : tadev::Range{m, k}), right(ctx, tb ? tadev::Range{n, k} : tadev::Range{k, n});
      left.from_host(ta ? AT.data() : A.data());
      right.from_host(tb ? BT.data() : B.data());
      GemmHelper h(ta ? Op::Trans : Op::NoTrans, tb ? Op::Trans : Op::NoTrans, 2u, 2u, 2u);
      Tile result;                       // empty: first op(result, left, right) seeds it
      gemm(result, left, right, 3, h);
      std::vector<double> got((size_t)m * n);
      result.to_host(got.data());
      bool ok = true;
      for (size_t i = 0; i < got.size(); ++i) ok = ok && got[i] == ref[i];
      CHECK(ok);
      gemm(result, left, right, 3, h);   // second application accumulates
      result.to_host(got.data());
      ok = true;
      for (size_t i = 0; i < got.size(); ++i) ok = ok && got[i] == 2 * ref[i];
      CHECK(ok);
      Tile fresh = gemm(left, right, 3, h);
      add_to(fresh, result);             // partial-result merge
      fresh.to_host(got.data());
      ok = true;
      for (size_t i = 0; i < got.size(); ++i) ok = ok && got[i] == 3 * ref[i];
      CHECK(ok);
    }

  // congruence violation -> exception (TA_ASSERT analogue)
  {
    Tile l(ctx, {4, 5}), r(ctx, {6, 3});
    bool threw = false;
    try { (void)gemm(l, r, 1.0, GemmHelper(Op::NoTrans, Op::NoTrans, 2u, 2u, 2u)); } catch (const tadev::Exception&) { threw = true; }
    CHECK(threw);
  }

  // rank-4 permutations, tests/librett.cpp:594-661
  {
    const int64_t a = 2, b = 3, c = 6, d = 4;
    std::vector<double> in((size_t)(a * b * c * d));
    for (size_t i = 0; i < in.size(); ++i) in[i] = (double)i;
    Tile ta_(ctx, {a, b, c, d});
    ta_.from_host(in.data());
    {
      Tile tb_ = permute(ta_, {0, 3, 2, 1});  // b(i,l,k,j) = a(i,j,k,l)
      CHECK((tb_.range() == tadev::Range{a, d, c, b}));
      std::vector<double> out(in.size());
      tb_.to_host(out.data());
      size_t it = 0; bool ok = true;
      for (int64_t i = 0; i < a; ++i) for (int64_t j = 0; j < b; ++j) for (int64_t kk = 0; kk < c; ++kk) for (int64_t l = 0; l < d; ++l, ++it)
        ok = ok && out[(size_t)(((i * d + l) * c + kk) * b + j)] == (double)it;
      CHECK(ok);
    }
    {
      Tile tb_ = permute(ta_, {1, 0, 3, 2});  // b(j,i,l,k) = a(i,j,k,l)
      std::vector<double> out(in.size());
      tb_.to_host(out.data());
      size_t it = 0; bool ok = true;
      for (int64_t i = 0; i < a; ++i) for (int64_t j = 0; j < b; ++j) for (int64_t kk = 0; kk < c; ++kk) for (int64_t l = 0; l < d; ++l, ++it)
        ok = ok && out[(size_t)(((j * a + i) * d + l) * c + kk)] == (double)it;
      CHECK(ok);
    }
    Tile cl = clone(ta_);
    scale_to(cl, 2.0);
    std::vector<double> out(in.size());
    cl.to_host(out.data());
    bool ok = true;
    for (size_t i = 0; i < in.size(); ++i) ok = ok && out[i] == 2.0 * in[i];
    CHECK(ok);
  }
  // MADWorld-style concurrency (SURVEY §8 b / a10): pool threads call the tile ops concurrently, each
  // task on the stream of its result (stream_for(ordinal)); the ABI promises thread-safety for
  // concurrent calls, including calls that share a stream (descriptor staging ring, tensor-map cache).
  {
    const int nthreads = 8, ntasks = 24, dim = 96;  // even, 16-byte aligned rows -> the TMA fast path
    const auto X = int_matrix(dim, dim, 7), Y = int_matrix(dim, dim, 8);
    std::vector<double> want((size_t)dim * dim, 0.0);
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) { double s2 = 0; for (int x = 0; x < dim; ++x) s2 += X[(size_t)i * dim + x] * Y[(size_t)x * dim + j]; want[(size_t)i * dim + j] = s2; }
    Tile tx(ctx, {dim, dim}), ty(ctx, {dim, dim});
    tx.from_host(X.data());
    ty.from_host(Y.data());
    std::vector<int> bad(nthreads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&, t] {
        GemmHelper h(Op::NoTrans, Op::NoTrans, 2u, 2u, 2u);
        std::vector<double> got((size_t)dim * dim);
        for (int task = t; task < ntasks; task += nthreads) {
          Tile acc(ctx, {dim, dim}, (uint64_t)task);   // stream = task % nstreams: several threads share a stream
          const int reps = 1 + task % 3;
          Tile first = gemm(tx, ty, 1.0, h);
          Tile sum = clone(first);
          for (int r = 1; r < reps; ++r) gemm(sum, tx, ty, 1.0, h);  // reduce-pair accumulation
          Tile p = permute(sum, {1, 0});
          p.to_host(got.data());
          for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) if (got[(size_t)j * dim + i] != reps * want[(size_t)i * dim + j]) bad[t]++;
        }
      });
    for (auto& th : pool) th.join();
    int nbad = 0;
    for (int b : bad) nbad += b;
    CHECK(nbad == 0);
  }
  // ---- asynchrony: chains of tile ops across streams with NO host synchronisation until the very end
  // (external/device.h:847-875 contract). Every tile of a chain lives on a different stream (ordinal = step), so each
  // op depends on the previous one through the event ordering of the plug-in; inputs are dropped while consumers
  // are still queued (their storage must outlive the reads); completion is observed through callbacks.
  {
    const int nthreads = 6, nchains = 18, dim = 64, steps = 12;
    const auto X = int_matrix(dim, dim, 11);
    std::vector<double> I((size_t)dim * dim, 0.0);
    for (int i = 0; i < dim; ++i) I[(size_t)i * dim + i] = 1.0;
    Tile tI(ctx, {dim, dim}, 1);
    tI.from_host(I.data());
    std::atomic<int> callbacks{0};
    std::vector<Tile> finals(nchains);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&, t] {
        GemmHelper h(Op::NoTrans, Op::NoTrans, 2u, 2u, 2u);
        for (int chain = t; chain < nchains; chain += nthreads) {
          Tile cur(ctx, {dim, dim}, (uint64_t)chain);
          cur.from_host(X.data());
          for (int st = 0; st < steps; ++st) {
            // alternate: transpose, multiply by the identity (on another stream), clone, add to itself and halve
            Tile nxt;
            tadev::this_task_ordinal() = (uint64_t)(chain + st);  // results of this step go to another stream
            switch (st % 4) {
              case 0: nxt = permute(cur, {1, 0}); break;
              case 1: nxt = gemm(cur, tI, 1.0, h); break;
              case 2: nxt = clone(cur); break;
              default: nxt = add(cur, cur); scale_to(nxt, 0.5); break;
            }
            cur = nxt;  // the previous tile is released while its consumer may still be queued
          }
          tadev::this_task_ordinal() = 0;
          cur.on_ready([](void* u) { static_cast<std::atomic<int>*>(u)->fetch_add(1); }, &callbacks);
          finals[chain] = cur;
        }
      });
    for (auto& th : pool) th.join();
    // steps: 3 transposes in 12 steps (st = 0, 4, 8) -> odd number -> transposed once overall
    int nbad = 0;
    std::vector<double> got((size_t)dim * dim);
    for (int chain = 0; chain < nchains; ++chain) {
      finals[chain].to_host(got.data());
      for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) if (got[(size_t)j * dim + i] != X[(size_t)i * dim + j]) ++nbad;
    }
    CHECK(nbad == 0);
    ctx.sync();
    CHECK(callbacks.load() == nchains);
  }
  // squared_norm / norm / shift / subt / mult
  {
    const int dim = 300;  // > one 16384-element reduction chunk
    const auto X = int_matrix(dim, dim, 21);
    Tile tx(ctx, {dim, dim});
    tx.from_host(X.data());
    double want = 0;
    for (double v : X) want += v * v;
    CHECK(squared_norm(tx) == want);
    CHECK(norm(tx) == std::sqrt(want));
    Tile sh = shift(tx, {5, -2});
    CHECK((sh.lobound() == tadev::Range{5, -2}) && (tx.lobound() == tadev::Range{0, 0}));
    Tile d = subt(sh, tx);
    CHECK(squared_norm(d) == 0.0);
    Tile p = mult(tx, tx);
    std::vector<double> got(X.size());
    p.to_host(got.data());
    bool ok = true;
    for (size_t i = 0; i < X.size(); ++i) ok = ok && got[i] == X[i] * X[i];
    CHECK(ok);
  }
  // per-tile permute throughput (the path a TA::Tile permute takes, one launch per tile): 8 MiB tiles
  // (16,16,64,64) -> (0,2,3,1), no host sync between launches; algorithmic bytes = 2 x tile bytes
  {
    const tadev::Range ext{16, 16, 64, 64};
    const int ntiles = 64, reps = 6;
    std::vector<Tile> src;
    for (int i = 0; i < ntiles; ++i) { src.emplace_back(ctx, ext, (uint64_t)i); tadev_fill_uniform_f64(ctx.get(), src.back().stream(), src.back().data(), (size_t)src.back().size(), 5, (uint64_t)i << 32); src.back().release_write(); }
    ctx.sync();
    {  // warm-up: grow the tile pool (mapping fresh device memory costs ~100 us per tile, once)
      std::vector<Tile> warm;
      for (int i = 0; i < ntiles; ++i) warm.push_back(permute(src[i], {0, 2, 3, 1}));
      ctx.sync();
    }
    for (int nthr : {1, 3}) {
      std::vector<std::vector<Tile>> keep(nthr);
      ctx.sync();
      const auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> pool;
      for (int t = 0; t < nthr; ++t)
        pool.emplace_back([&, t] {
          for (int rp = 0; rp < reps; ++rp) {
            keep[t].clear();  // results of the previous round go back to the pool (stream-ordered)
            for (int i = t; i < ntiles; i += nthr) keep[t].push_back(permute(src[i], {0, 2, 3, 1}));
          }
        });
      for (auto& th : pool) th.join();
      ctx.sync();
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const double gbs = 2.0 * 8.0 * 16 * 16 * 64 * 64 * ntiles * reps / secs / 1e9;
      std::printf("PERMUTE_PER_TILE threads=%d tiles=%d x %d GBs=%.1f\n", nthr, ntiles, reps, gbs);
    }
  }
  ctx.sync();
  std::printf(failures ? "CPP_PLUGIN FAILED (%d)\n" : "CPP_PLUGIN OK\n", failures);
  return failures ? 1 : 0;
}
