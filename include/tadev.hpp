// tadev.hpp — C++ tile plug-in over the libtadev C ABI (include/tadev.h).
//
// This is the host-side drop-in for TiledArray's device tile (reference:
// src/TiledArray/device/btas_um_tensor.h:98-565): a shallow-copy, reference-counted tile of doubles
// in device memory plus the ADL customization points the contraction path calls
// (src/TiledArray/tile_op/tile_interface.h:803-830 gemm, tile_interface/permute.h permute,
// contract_reduce.h:397 add_to, clone, scale, squared_norm/norm, empty). Signatures, argument
// meaning and error behaviour follow the reference; only TiledArray's own types are replaced by
// minimal stand-ins (Range = extents vector, Permutation = image-form vector, GemmHelper restated
// from math/gemm_helper.h:41-278) so the header compiles without MADNESS/Boost. With TiledArray
// available, `tadev::Tile` is wrapped as `TA::Tile<tadev::Tile>` and `tadev::GemmHelper` /
// `tadev::Permutation` are replaced by the TiledArray types.
//
// Every operation is asynchronous on a stream chosen like the reference does
// (stream_for(range): ordinal % nstreams, external/device.h:899-907); call `Context::sync()` or
// `tile.sync()` before reading results on the host.
#pragma once
#include <cstdint>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "tadev.h"

namespace tadev {

struct Exception : std::runtime_error {  // TiledArray::Exception analogue (error.h:39-83)
  int code;
  Exception(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != TADEV_OK) throw Exception(rc, tadev_last_error());
}
#define TADEV_ASSERT(cond, msg) \
  do { if (!(cond)) throw ::tadev::Exception(TADEV_EINVAL, msg); } while (0)

// device::Env analogue: one context per process/GPU
class Context {
 public:
  explicit Context(int device = 0, size_t pool_bytes = 0) { check(tadev_init(device, pool_bytes, &ctx_)); check(tadev_num_streams(ctx_, &nstreams_)); }
  ~Context() { if (ctx_) tadev_finalize(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  tadev_ctx* get() const { return ctx_; }
  tadev_stream stream_for(uint64_t ordinal) const { tadev_stream s; check(tadev_stream_for(ctx_, ordinal, &s)); return s; }
  void sync() const { for (int i = 0; i < nstreams_; ++i) { tadev_stream s; check(tadev_get_stream(ctx_, i, &s)); check(tadev_stream_sync(ctx_, s)); } }
  static Context*& current() { static Context* c = nullptr; return c; }  // deviceEnv::instance() analogue

 private:
  tadev_ctx* ctx_ = nullptr;
  int nstreams_ = 0;
};

using Permutation = std::vector<int32_t>;  // image form: result[perm[i]] = arg[i] (permutation.h:69-79)
using Range = std::vector<int64_t>;        // tile extents, row-major

inline int64_t volume(const Range& r) { return std::accumulate(r.begin(), r.end(), int64_t(1), std::multiplies<int64_t>()); }

enum class Op { NoTrans = TADEV_OP_N, Trans = TADEV_OP_T };

// math::GemmHelper (math/gemm_helper.h:41-278)
class GemmHelper {
 public:
  GemmHelper(Op left_op, Op right_op, unsigned result_rank, unsigned left_rank, unsigned right_rank)
      : left_op_(left_op), right_op_(right_op), result_rank_(result_rank), left_rank_(left_rank), right_rank_(right_rank) {
    TADEV_ASSERT(((left_rank + right_rank - result_rank) % 2u) == 0u, "GemmHelper: inconsistent ranks");
    const unsigned c = num_contract_ranks();
    if (left_op == Op::NoTrans) { lo_[0] = 0; lo_[1] = li_[0] = left_rank - c; li_[1] = left_rank; }
    else { li_[0] = 0; li_[1] = lo_[0] = c; lo_[1] = left_rank; }
    if (right_op == Op::NoTrans) { ri_[0] = 0; ri_[1] = ro_[0] = c; ro_[1] = right_rank; }
    else { ro_[0] = 0; ro_[1] = ri_[0] = right_rank - c; ri_[1] = right_rank; }
  }
  unsigned num_contract_ranks() const { return (left_rank_ + right_rank_ - result_rank_) >> 1; }
  Op left_op() const { return left_op_; }
  Op right_op() const { return right_op_; }
  void compute_matrix_sizes(int64_t& m, int64_t& n, int64_t& k, const Range& left, const Range& right) const {
    TADEV_ASSERT(left.size() == left_rank_ && right.size() == right_rank_, "GemmHelper: rank mismatch");
    m = k = n = 1;
    for (unsigned i = lo_[0]; i < lo_[1]; ++i) m *= left[i];
    for (unsigned i = li_[0]; i < li_[1]; ++i) k *= left[i];
    for (unsigned i = ro_[0]; i < ro_[1]; ++i) n *= right[i];
  }
  Range make_result_range(const Range& left, const Range& right) const {
    Range r;
    for (unsigned i = lo_[0]; i < lo_[1]; ++i) r.push_back(left[i]);
    for (unsigned i = ro_[0]; i < ro_[1]; ++i) r.push_back(right[i]);
    return r;
  }
  bool left_right_congruent(const Range& left, const Range& right) const {
    for (unsigned d = 0; d < num_contract_ranks(); ++d)
      if (left[li_[0] + d] != right[ri_[0] + d]) return false;
    return true;
  }

 private:
  Op left_op_, right_op_;
  unsigned result_rank_, left_rank_, right_rank_;
  unsigned li_[2], lo_[2], ri_[2], ro_[2];
};

// The device tile: shallow copy semantics (copies share storage, like TA::Tensor / TA::Tile).
class Tile {
  struct Storage {
    Context* ctx; double* ptr; size_t bytes; tadev_stream stream;
    ~Storage() { if (ptr) tadev_free(ctx->get(), ptr, stream); }
  };

 public:
  Tile() = default;
  Tile(Context& ctx, const Range& range, uint64_t ordinal = 0) : range_(range) {
    auto st = std::make_shared<Storage>();
    st->ctx = &ctx; st->ptr = nullptr; st->bytes = sizeof(double) * (size_t)volume(range); st->stream = ctx.stream_for(ordinal);
    check(tadev_alloc(ctx.get(), st->bytes, (void**)&st->ptr, st->stream));
    storage_ = std::move(st);
  }
  bool empty() const { return !storage_; }
  const Range& range() const { return range_; }
  int64_t size() const { return empty() ? 0 : volume(range_); }
  double* data() const { return storage_ ? storage_->ptr : nullptr; }
  Context& context() const { return *storage_->ctx; }
  tadev_stream stream() const { return storage_->stream; }
  void sync() const { if (storage_) check(tadev_stream_sync(storage_->ctx->get(), storage_->stream)); }
  // host <-> device helpers (the analogue of to_host/to_device, btas_um_tensor.h:64-88)
  void from_host(const double* src) { check(tadev_memcpy_h2d(context().get(), data(), src, storage_->bytes, stream())); sync(); }
  void to_host(double* dst) const { check(tadev_memcpy_d2h(context().get(), dst, data(), storage_->bytes, stream())); sync(); }

 private:
  Range range_;
  std::shared_ptr<Storage> storage_;
};

// ---- customization points of the contraction path (found by ADL) ------------------------------

// gemm(left, right, factor, helper): new result, beta = 0 (tile_interface.h:803-808)
template <typename Scalar>
Tile gemm(const Tile& left, const Tile& right, Scalar factor, const GemmHelper& h) {
  TADEV_ASSERT(!left.empty() && !right.empty(), "gemm: empty argument");
  TADEV_ASSERT(h.left_right_congruent(left.range(), right.range()), "gemm: contracted ranges are not congruent");
  int64_t m, n, k;
  h.compute_matrix_sizes(m, n, k, left.range(), right.range());
  Tile result(left.context(), h.make_result_range(left.range(), right.range()));
  // order the inputs before the kernel on the result's stream
  left.sync(); right.sync();
  check(tadev_gemm_f64(left.context().get(), result.stream(), (int)h.left_op(), (int)h.right_op(), (int)m, (int)n, (int)k,
                       (double)factor, left.data(), right.data(), 0.0, result.data()));
  return result;
}

// gemm(result, left, right, factor, helper): accumulate, beta = 1 (tile_interface.h:825-830);
// an empty result is seeded (Tensor::gemm, tensor.h:3134-3140)
template <typename Scalar>
Tile& gemm(Tile& result, const Tile& left, const Tile& right, Scalar factor, const GemmHelper& h) {
  if (result.empty()) { result = gemm(left, right, factor, h); return result; }
  TADEV_ASSERT(h.left_right_congruent(left.range(), right.range()), "gemm: contracted ranges are not congruent");
  TADEV_ASSERT(result.range() == h.make_result_range(left.range(), right.range()), "gemm: result range mismatch");
  int64_t m, n, k;
  h.compute_matrix_sizes(m, n, k, left.range(), right.range());
  left.sync(); right.sync();
  check(tadev_gemm_f64(result.context().get(), result.stream(), (int)h.left_op(), (int)h.right_op(), (int)m, (int)n, (int)k,
                       (double)factor, left.data(), right.data(), 1.0, result.data()));
  return result;
}

// permute(arg, perm) (tile_interface/permute.h; device ref btas_um_tensor.h:169-191)
inline Tile permute(const Tile& arg, const Permutation& perm) {
  TADEV_ASSERT(!arg.empty() && perm.size() == arg.range().size(), "permute: rank mismatch");
  Range rr(arg.range().size());
  for (size_t i = 0; i < perm.size(); ++i) rr[perm[i]] = arg.range()[i];
  Tile result(arg.context(), rr);
  arg.sync();
  check(tadev_permute(arg.context().get(), result.stream(), (int)perm.size(), arg.range().data(), perm.data(), 8, arg.data(), result.data()));
  return result;
}

// add_to(result, arg): ContractReduce partial-result merge (contract_reduce.h:397-398)
inline Tile& add_to(Tile& result, const Tile& arg) {
  TADEV_ASSERT(result.range() == arg.range(), "add_to: range mismatch");
  arg.sync();
  check(tadev_add_to_f64(result.context().get(), result.stream(), (size_t)result.size(), result.data(), arg.data()));
  return result;
}

inline Tile clone(const Tile& arg) {
  Tile r(arg.context(), arg.range());
  arg.sync();
  check(tadev_memset(arg.context().get(), r.data(), 0, sizeof(double) * (size_t)r.size(), r.stream()));
  check(tadev_add_to_f64(arg.context().get(), r.stream(), (size_t)r.size(), r.data(), arg.data()));
  return r;
}

template <typename Scalar>
Tile& scale_to(Tile& arg, Scalar factor) {
  check(tadev_scale_f64(arg.context().get(), arg.stream(), (size_t)arg.size(), arg.data(), (double)factor));
  return arg;
}

inline bool empty(const Tile& t) { return t.empty(); }

}  // namespace tadev
