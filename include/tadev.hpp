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
// (stream_for(range): ordinal % nstreams, external/device.h:899-907) and ordered against the other streams by
// events, never by host synchronisation; `tile.sync()`, `to_host`, `squared_norm` are the only blocking calls,
// `tile.on_ready(fn, user)` is the completion callback a task runtime uses instead.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <mutex>
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
  ~Context() {
    if (!ctx_) return;
    for (void* e : event_pool_) tadev_event_destroy(ctx_, e);
    tadev_finalize(ctx_);
  }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  tadev_ctx* get() const { return ctx_; }
  tadev_stream stream_for(uint64_t ordinal) const { tadev_stream s; check(tadev_stream_for(ctx_, ordinal, &s)); return s; }
  void sync() const { for (int i = 0; i < nstreams_; ++i) { tadev_stream s; check(tadev_get_stream(ctx_, i, &s)); check(tadev_stream_sync(ctx_, s)); } }
  static Context*& current() { static Context* c = nullptr; return c; }  // deviceEnv::instance() analogue
  // sync events are pooled: a tile op needs one or two and creating one costs microseconds
  void* acquire_event() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!event_pool_.empty()) { void* e = event_pool_.back(); event_pool_.pop_back(); return e; }
    }
    void* e = nullptr;
    check(tadev_sync_event_create(ctx_, &e));
    return e;
  }
  void release_event(void* e) { if (e) { std::lock_guard<std::mutex> lk(mu_); event_pool_.push_back(e); } }
  // completion hook of a task that enqueued work on `s` (sync_madness_task_with + add_device_task,
  // external/device.h:847-875): fn(user) runs once that work has finished; no host thread blocks meanwhile
  void add_completion_callback(tadev_stream s, tadev_host_fn fn, void* user) const { check(tadev_stream_add_callback(ctx_, s, fn, user)); }

 private:
  tadev_ctx* ctx_ = nullptr;
  int nstreams_ = 0;
  std::mutex mu_;
  std::vector<void*> event_pool_;
};

// The stream a task's results go to. The reference keeps a thread-local "stream of the current device task"
// (detail::madness_task_stream_opt_ptr_accessor, external/device.h:822-840) and otherwise derives the stream from the
// result range's offset (stream_for, :899-907); here a task sets the ordinal its results should be keyed by.
inline uint64_t& this_task_ordinal() { static thread_local uint64_t ordinal = 0; return ordinal; }

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
//
// Asynchrony (reference contract: device/btas_um_tensor.h:98-565 with external/device.h:822-907): every operation
// ENQUEUES work on the home stream of its result (stream_for(ordinal)) and returns; nothing below blocks the host
// except the functions that hand data or a scalar to the host (to_host, squared_norm, norm, sync). Ordering across
// streams is by events: a tile carries the event of its last write; an op whose input lives on another stream makes
// its own stream wait for that event, and leaves a "read finished" event with the input so that the input's
// storage is neither overwritten nor returned to the pool before the read has run. A task runtime (MADWorld)
// completes its task with Tile::on_ready / Context::add_completion_callback instead of synchronising.
class Tile {
  struct Storage {
    Context* ctx; double* ptr; size_t bytes; tadev_stream stream;
    void* ready = nullptr;        // recorded on `stream` after the last write
    std::mutex mu;
    std::vector<void*> readers;   // reads enqueued on OTHER streams since the last write
    void wait_readers() {         // the home stream waits for every outstanding foreign read (WAR / free hazard)
      std::vector<void*> r;
      { std::lock_guard<std::mutex> lk(mu); r.swap(readers); }
      for (void* e : r) { tadev_stream_wait_event(ctx->get(), stream, e); ctx->release_event(e); }
    }
    ~Storage() {
      wait_readers();
      if (ready) ctx->release_event(ready);
      if (ptr) tadev_free(ctx->get(), ptr, stream);
    }
  };

 public:
  Tile() = default;
  Tile(Context& ctx, const Range& range) : Tile(ctx, range, this_task_ordinal()) {}
  Tile(Context& ctx, const Range& range, uint64_t ordinal) : range_(range), lobound_(range.size(), 0) {
    auto st = std::make_shared<Storage>();
    st->ctx = &ctx; st->ptr = nullptr; st->bytes = sizeof(double) * (size_t)volume(range); st->stream = ctx.stream_for(ordinal);
    check(tadev_alloc(ctx.get(), st->bytes, (void**)&st->ptr, st->stream));
    storage_ = std::move(st);
  }
  bool empty() const { return !storage_; }
  const Range& range() const { return range_; }
  const Range& lobound() const { return lobound_; }
  int64_t size() const { return empty() ? 0 : volume(range_); }
  double* data() const { return storage_ ? storage_->ptr : nullptr; }
  Context& context() const { return *storage_->ctx; }
  tadev_stream stream() const { return storage_->stream; }

  // ---- ordering (used by the tile ops below)
  // the caller is about to READ this tile on stream s: make s wait for the last write
  void acquire_read(tadev_stream s) const {
    if (s != storage_->stream && storage_->ready) check(tadev_stream_wait_event(context().get(), s, storage_->ready));
  }
  // the read enqueued on s is in place: remember it until the next write / the release of the storage
  void release_read(tadev_stream s) const {
    if (s == storage_->stream) return;
    void* e = context().acquire_event();
    check(tadev_event_record(context().get(), e, s));
    std::lock_guard<std::mutex> lk(storage_->mu);
    storage_->readers.push_back(e);
  }
  // the caller is about to WRITE this tile on its home stream / has enqueued the write
  void acquire_write() const { storage_->wait_readers(); }
  void release_write() const {
    if (!storage_->ready) storage_->ready = context().acquire_event();
    check(tadev_event_record(context().get(), storage_->ready, storage_->stream));
  }
  // ---- completion
  void sync() const {  // blocks the host until the last write has finished
    if (!storage_) return;
    if (storage_->ready) check(tadev_event_sync(context().get(), storage_->ready));
    else check(tadev_stream_sync(context().get(), storage_->stream));
  }
  bool is_ready() const {
    if (!storage_ || !storage_->ready) return true;
    int done = 0;
    check(tadev_event_query(context().get(), storage_->ready, &done));
    return done != 0;
  }
  // fn(user) runs (on a runtime thread) when everything enqueued so far on the tile's stream has finished
  void on_ready(tadev_host_fn fn, void* user) const { context().add_completion_callback(stream(), fn, user); }
  // host <-> device (the analogue of to_host/to_device, btas_um_tensor.h:64-88). from_host does not block:
  // a pageable source is staged by the runtime before the call returns.
  void from_host(const double* src) {
    acquire_write();
    check(tadev_memcpy_h2d(context().get(), data(), src, storage_->bytes, stream()));
    release_write();
  }
  void to_host(double* dst) const {
    check(tadev_memcpy_d2h(context().get(), dst, data(), storage_->bytes, stream()));
    check(tadev_stream_sync(context().get(), stream()));
  }
  // shift_to (tile_interface/shift.h): only the lower bound changes, the data stay
  Tile& shift_to(const std::vector<int64_t>& bound_shift) {
    TADEV_ASSERT(bound_shift.size() == lobound_.size(), "shift_to: rank mismatch");
    for (size_t d = 0; d < lobound_.size(); ++d) lobound_[d] += bound_shift[d];
    return *this;
  }

 private:
  Range range_;
  Range lobound_;
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
  const tadev_stream s = result.stream();
  left.acquire_read(s); right.acquire_read(s);
  check(tadev_gemm_f64(left.context().get(), s, (int)h.left_op(), (int)h.right_op(), (int)m, (int)n, (int)k,
                       (double)factor, left.data(), right.data(), 0.0, result.data()));
  left.release_read(s); right.release_read(s);
  result.release_write();
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
  const tadev_stream s = result.stream();
  result.acquire_write();
  left.acquire_read(s); right.acquire_read(s);
  check(tadev_gemm_f64(result.context().get(), s, (int)h.left_op(), (int)h.right_op(), (int)m, (int)n, (int)k,
                       (double)factor, left.data(), right.data(), 1.0, result.data()));
  left.release_read(s); right.release_read(s);
  result.release_write();
  return result;
}

// permute(arg, perm) (tile_interface/permute.h; device ref btas_um_tensor.h:169-191)
inline Tile permute(const Tile& arg, const Permutation& perm) {
  TADEV_ASSERT(!arg.empty() && perm.size() == arg.range().size(), "permute: rank mismatch");
  Range rr(arg.range().size());
  for (size_t i = 0; i < perm.size(); ++i) rr[perm[i]] = arg.range()[i];
  Tile result(arg.context(), rr);
  const tadev_stream s = result.stream();
  arg.acquire_read(s);
  check(tadev_permute(arg.context().get(), s, (int)perm.size(), arg.range().data(), perm.data(), 8, arg.data(), result.data()));
  arg.release_read(s);
  result.release_write();
  return result;
}

// add_to(result, arg): ContractReduce partial-result merge (contract_reduce.h:397-398)
inline Tile& add_to(Tile& result, const Tile& arg) {
  TADEV_ASSERT(result.range() == arg.range(), "add_to: range mismatch");
  const tadev_stream s = result.stream();
  result.acquire_write();
  arg.acquire_read(s);
  check(tadev_add_to_f64(result.context().get(), s, (size_t)result.size(), result.data(), arg.data()));
  arg.release_read(s);
  result.release_write();
  return result;
}

// clone: ONE device-to-device copy (btas_um_tensor.h:98-118)
inline Tile clone(const Tile& arg) {
  if (arg.empty()) return Tile();
  Tile r(arg.context(), arg.range());
  const tadev_stream s = r.stream();
  arg.acquire_read(s);
  check(tadev_memcpy_d2d(arg.context().get(), r.data(), arg.data(), sizeof(double) * (size_t)r.size(), s));
  arg.release_read(s);
  r.release_write();
  r.shift_to(arg.lobound());
  return r;
}

// shift (tile_interface/shift.h): a copy whose lower bound is moved
inline Tile shift(const Tile& arg, const std::vector<int64_t>& bound_shift) { Tile r = clone(arg); r.shift_to(bound_shift); return r; }
inline Tile& shift_to(Tile& arg, const std::vector<int64_t>& bound_shift) { return arg.shift_to(bound_shift); }

template <typename Scalar>
Tile& scale_to(Tile& arg, Scalar factor) {
  arg.acquire_write();
  check(tadev_scale_f64(arg.context().get(), arg.stream(), (size_t)arg.size(), arg.data(), (double)factor));
  arg.release_write();
  return arg;
}
template <typename Scalar>
Tile scale(const Tile& arg, Scalar factor) { Tile r = clone(arg); return scale_to(r, factor); }

namespace detail {
// out = alpha * x (+ beta * y | .* y) on the stream of a NEW result tile (tile_op/add.h, subt.h, mult.h)
inline Tile binary(int op, const Tile& x, const Tile& y, double alpha, double beta) {
  TADEV_ASSERT(!x.empty() && !y.empty() && x.range() == y.range(), "element-wise tile op: range mismatch");
  Tile r(x.context(), x.range());
  const tadev_stream s = r.stream();
  x.acquire_read(s); y.acquire_read(s);
  double* out = r.data();
  const double *px = x.data(), *py = y.data();
  const int64_t n = r.size();
  check(tadev_tiles_binary_f64(x.context().get(), s, op, 1, &out, &px, &py, &n, alpha, beta));
  x.release_read(s); y.release_read(s);
  r.release_write();
  return r;
}
}  // namespace detail
inline Tile add(const Tile& x, const Tile& y) { return detail::binary(TADEV_EW_AXPBY, x, y, 1.0, 1.0); }
inline Tile subt(const Tile& x, const Tile& y) { return detail::binary(TADEV_EW_AXPBY, x, y, 1.0, -1.0); }
inline Tile mult(const Tile& x, const Tile& y) { return detail::binary(TADEV_EW_MULT, x, y, 1.0, 0.0); }

// squared_norm / norm hand a scalar to the host, so they wait for the reduction (as the reference's device tile
// does: btas_um_tensor.h squared_norm -> blas dot into a host scalar)
inline double squared_norm(const Tile& arg) {
  if (arg.empty() || arg.size() == 0) return 0.0;
  double out = 0.0;
  check(tadev_sqnorm_f64(arg.context().get(), arg.stream(), (size_t)arg.size(), arg.data(), &out));
  return out;
}
inline double norm(const Tile& arg) { return std::sqrt(squared_norm(arg)); }

inline bool empty(const Tile& t) { return t.empty(); }

}  // namespace tadev
