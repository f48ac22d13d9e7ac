// elementwise.cu — batched tile-wise add / subtract / scale / Hadamard product, HBM-bandwidth bound.
//
// Replaces, for device tiles, the reference's element-wise tile ops next to the contraction path
// (SURVEY §8 f2/f4): tile_op/add.h, subt.h, scal.h, mult.h -> Tensor::add/subt/scale/mult
// (tensor/tensor.h) and the device versions device/btas_um_tensor.h:377-470 (cuBLAS axpy/scal) and
// device/kernel/thrust/mult_kernel.h (thrust::transform). One launch handles every tile of an
// expression (blockIdx.y = tile), 16-byte vector accesses, grid-stride over the tile.
// Algorithmic bytes: 8 * (inputs read + 1 written) per element.
#include <algorithm>

#include "common.h"

namespace {

struct TileOp {
  double* out;
  const double* x;  // nullptr = zero tile
  const double* y;  // nullptr = zero tile
  int64_t n;
};

template <int OP>
__device__ __forceinline__ double apply(double x, double y, double alpha, double beta) {
  return OP == TADEV_EW_AXPBY ? alpha * x + beta * y : alpha * (x * y);
}

template <int OP>
__global__ void __launch_bounds__(256) tiles_binary_kernel(const TileOp* __restrict__ ops, double alpha, double beta) {
  const TileOp t = ops[blockIdx.y];
  const int64_t n = t.n;
  const bool vec = ((reinterpret_cast<uintptr_t>(t.out) | reinterpret_cast<uintptr_t>(t.x) | reinterpret_cast<uintptr_t>(t.y)) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const int64_t n2 = n >> 1;
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(t.x);
    const double2* __restrict__ y2 = reinterpret_cast<const double2*>(t.y);
    double2* __restrict__ o2 = reinterpret_cast<double2*>(t.out);
    const double2 zero = make_double2(0.0, 0.0);
    // four independent 16-byte accesses per input stream in flight per thread (8 loads outstanding), streaming
    // cache hints on both sides: every byte is touched exactly once, nothing is worth keeping in L2
    constexpr int U = 4;
    for (int64_t i = tid; i < n2; i += U * stride) {
      double2 xv[U], yv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = i + u * stride;
        xv[u] = (x2 && j < n2) ? __ldcs(x2 + j) : zero;
        yv[u] = (y2 && j < n2) ? __ldcs(y2 + j) : zero;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = i + u * stride;
        if (j < n2) __stcs(o2 + j, make_double2(apply<OP>(xv[u].x, yv[u].x, alpha, beta), apply<OP>(xv[u].y, yv[u].y, alpha, beta)));
      }
    }
    if ((n & 1) && tid == 0) t.out[n - 1] = apply<OP>(t.x ? t.x[n - 1] : 0.0, t.y ? t.y[n - 1] : 0.0, alpha, beta);
  } else {
    for (int64_t i = tid; i < n; i += stride) t.out[i] = apply<OP>(t.x ? t.x[i] : 0.0, t.y ? t.y[i] : 0.0, alpha, beta);
  }
}

}  // namespace

extern "C" int tadev_tiles_binary_f64(tadev_ctx* ctx, tadev_stream s_, int op, int ntiles, double* const* h_out,
                                      const double* const* h_x, const double* const* h_y, const int64_t* h_elems,
                                      double alpha, double beta) {
  TADEV_REQUIRE(ctx, "tadev_tiles_binary_f64: null ctx");
  TADEV_REQUIRE(op == TADEV_EW_AXPBY || op == TADEV_EW_MULT, "tadev_tiles_binary_f64: bad op %d", op);
  TADEV_REQUIRE(ntiles >= 0, "tadev_tiles_binary_f64: negative tile count");
  if (ntiles == 0) return TADEV_OK;
  TADEV_REQUIRE(h_out && h_x && h_y && h_elems, "tadev_tiles_binary_f64: null arrays");
  cudaStream_t s = (cudaStream_t)s_;
  for (int first = 0; first < ntiles; first += 32768) {  // gridDim.y limit
    const int n = std::min(32768, ntiles - first);
    StageLease L;
    int rc = L.acquire(ctx, s, sizeof(TileOp) * (size_t)n);
    if (rc) return rc;
    void *h = L.h, *d = L.d;
    TileOp* ops = static_cast<TileOp*>(h);
    int64_t maxn = 0;
    for (int i = 0; i < n; ++i) {
      TADEV_REQUIRE(h_elems[first + i] >= 0 && (h_out[first + i] || h_elems[first + i] == 0), "tadev_tiles_binary_f64: tile %d: null result", first + i);
      ops[i] = TileOp{h_out[first + i], h_x[first + i], h_y[first + i], h_elems[first + i]};
      maxn = std::max(maxn, h_elems[first + i]);
    }
    TADEV_CHECK_CUDA(cudaMemcpyAsync(d, h, sizeof(TileOp) * (size_t)n, cudaMemcpyHostToDevice, s));
    // enough CTAs to fill the machine a few times over, never more than the largest tile needs
    int64_t bx = ceil_div64(std::max<int64_t>(maxn / 2, 1), 256 * 4);
    const int64_t cap = std::max<int64_t>(1, (int64_t)ctx->num_sms * 16 / n);
    bx = std::max<int64_t>(1, std::min(bx, cap));
    const dim3 grid((unsigned)bx, (unsigned)n);
    if (op == TADEV_EW_AXPBY) tiles_binary_kernel<TADEV_EW_AXPBY><<<grid, 256, 0, s>>>(static_cast<const TileOp*>(d), alpha, beta);
    else tiles_binary_kernel<TADEV_EW_MULT><<<grid, 256, 0, s>>>(static_cast<const TileOp*>(d), alpha, beta);
    ctx->launches++;
    TADEV_CHECK_CUDA(cudaGetLastError());
  }  // ~StageLease records `done` after the launch and returns the slot
  return TADEV_OK;
}
