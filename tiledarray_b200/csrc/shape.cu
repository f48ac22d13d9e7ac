// shape.cu — SparseShape<float> screening and SUMMA tile-pair list construction on the device.
//
// Replaces (reference, src/TiledArray/sparse_shape.h):
//   scale_tile_norms<InverseVolume> :149-217  -> tadev_shape_scale_f32
//   gemm                            :1589-1681 -> tadev_shape_gemm_f32
//   mask                            :653-676  -> tadev_shape_mask_f32
// and Summa::contract's pair enumeration (dist_eval/contraction_eval.h:1311-1384) ->
//   tadev_build_pairlist.
//
// Bit-exactness contract (shared with oracle/): every fp32 operation is an individually rounded
// IEEE operation (no FMA contraction), and the k-sum of shape gemm is sequential in k:
//   la = a[m,k]*ksz[k];  rb = b[k,n]*ksz[k];  acc = acc + la*rb  (k = 0..K-1);  out = |f| * acc.
// The reference's own low-order bits depend on its vendor SGEMM (sparse_shape.h:1654), so this
// fixed order IS the spec here; the zero/non-zero decision is what the reference's tests pin.
#include "common.h"

namespace {

__device__ __forceinline__ void count_zero(bool z, unsigned long long* nzero) {
  const unsigned m = __ballot_sync(0xffffffffu, z);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nzero, (unsigned long long)__popc(m));
}

__global__ void __launch_bounds__(256) shape_scale_kernel(float* __restrict__ norms, const float* __restrict__ left,
                                                          int64_t nleft, const float* __restrict__ right,
                                                          int64_t nright, float thr, unsigned long long* nzero) {
  const int64_t n = nright > 0 ? nleft * nright : nleft;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool z = false;
  if (i < n) {
    float v = norms[i];
    if (nright > 0) {
      const float xy = __fmul_rn(left[i / nright], right[i % nright]);  // norm *= x * y
      v = __fmul_rn(v, xy);
    } else {
      v = __fdiv_rn(v, left[i]);  // rank-1: norm /= size
    }
    if (v < thr) { v = 0.0f; z = true; }
    norms[i] = v;
  }
  count_zero(z, nzero);
}

constexpr int ST = 16;
__global__ void __launch_bounds__(ST * ST) shape_gemm_kernel(int Mt, int Nt, int Kt, const float* __restrict__ a,
                                                             const float* __restrict__ b,
                                                             const float* __restrict__ ksz, float abs_factor,
                                                             float thr, float* __restrict__ out,
                                                             unsigned long long* nzero) {
  __shared__ float sa[ST][ST + 1];
  __shared__ float sb[ST][ST + 1];
  const int tx = threadIdx.x % ST, ty = threadIdx.x / ST;
  const int m = blockIdx.y * ST + ty, n = blockIdx.x * ST + tx;
  float acc = 0.0f;
  for (int k0 = 0; k0 < Kt; k0 += ST) {
    {
      const int am = blockIdx.y * ST + ty, ak = k0 + tx;
      sa[ty][tx] = (am < Mt && ak < Kt) ? __fmul_rn(a[(size_t)am * Kt + ak], ksz[ak]) : 0.0f;
      const int bk = k0 + ty, bn = blockIdx.x * ST + tx;
      sb[ty][tx] = (bk < Kt && bn < Nt) ? __fmul_rn(b[(size_t)bk * Nt + bn], ksz[bk]) : 0.0f;
    }
    __syncthreads();
    const int kend = min(ST, Kt - k0);
    for (int kk = 0; kk < kend; ++kk) acc = __fadd_rn(acc, __fmul_rn(sa[ty][kk], sb[kk][tx]));
    __syncthreads();
  }
  bool z = false;
  if (m < Mt && n < Nt) {
    float v = __fmul_rn(abs_factor, acc);
    if (v < thr) { v = 0.0f; z = true; }
    out[(size_t)m * Nt + n] = v;
  }
  count_zero(z, nzero);
}

// Kt == 0: outer product, norm = left*right*abs_factor (sparse_shape.h:1665-1678)
__global__ void __launch_bounds__(256) shape_outer_kernel(int Mt, int Nt, const float* __restrict__ a,
                                                          const float* __restrict__ b, float abs_factor, float thr,
                                                          float* __restrict__ out, unsigned long long* nzero) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool z = false;
  if (i < (int64_t)Mt * Nt) {
    float v = __fmul_rn(__fmul_rn(a[i / Nt], b[i % Nt]), abs_factor);
    if (v < thr) { v = 0.0f; z = true; }
    out[i] = v;
  }
  count_zero(z, nzero);
}

__global__ void __launch_bounds__(256) shape_mask_kernel(int64_t n, float* __restrict__ norms,
                                                         const float* __restrict__ mask, float thr_this,
                                                         float thr_mask, unsigned long long* nzero) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool z = false;
  if (i < n) {
    float l = norms[i];
    if (l >= thr_this && mask[i] < thr_mask) { l = 0.0f; z = true; norms[i] = l; }
  }
  count_zero(z, nzero);
}

// One block; candidates are this rank's local (i,j) slots in row-major order; block-wide
// exclusive scan (warp ballots) keeps the output ordered exactly like the reference's double loop.
__global__ void __launch_bounds__(1024) pairlist_kernel(int k, int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt,
                                                        const float* __restrict__ a, const float* __restrict__ b,
                                                        const float* __restrict__ cn, float thr,
                                                        int32_t* __restrict__ pi, int32_t* __restrict__ pj,
                                                        int32_t* __restrict__ npairs) {
  const int lrows = (Mt - r + Pr - 1) / Pr, lcols = (Nt - c + Pc - 1) / Pc;
  const int total = (r < Mt && c < Nt) ? lrows * lcols : 0;
  __shared__ int warp_sums[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int start = 0; start < total; start += 1024) {
    const int q = start + threadIdx.x;
    bool valid = false;
    int i = 0, j = 0;
    if (q < total) {
      i = r + (q / lcols) * Pr;
      j = c + (q % lcols) * Pc;
      valid = (!a || a[(size_t)i * Kt + k] >= thr) && (!b || b[(size_t)k * Nt + j] >= thr) &&
              (!cn || cn[(size_t)i * Nt + j] >= thr);
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int before = __popc(m & ((1u << lane) - 1));
    if (lane == 0) warp_sums[wid] = __popc(m);
    __syncthreads();
    if (wid == 0) {
      int v = warp_sums[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      warp_sums[lane] = v;  // inclusive
    }
    __syncthreads();
    const int off = base + (wid ? warp_sums[wid - 1] : 0) + before;
    if (valid) { pi[off] = i; pj[off] = j; }
    __syncthreads();
    if (threadIdx.x == 0) base += warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *npairs = base;
}

}  // namespace

extern "C" int tadev_shape_scale_f32(tadev_ctx* ctx, tadev_stream s, float* d_norms, const float* d_left,
                                     int64_t nleft, const float* d_right, int64_t nright, float threshold,
                                     uint64_t* d_nzero) {
  TADEV_REQUIRE(ctx, "tadev_shape_scale_f32: null ctx");
  TADEV_REQUIRE(nleft >= 0 && nright >= 0, "tadev_shape_scale_f32: negative sizes");
  const int64_t n = nright > 0 ? nleft * nright : nleft;
  if (n == 0) return TADEV_OK;
  TADEV_REQUIRE(d_norms && d_left && (nright == 0 || d_right) && d_nzero, "tadev_shape_scale_f32: null arrays");
  shape_scale_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)s>>>(
      d_norms, d_left, nleft, d_right, nright, threshold, (unsigned long long*)d_nzero);
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}

extern "C" int tadev_shape_gemm_f32(tadev_ctx* ctx, tadev_stream s, int Mt, int Nt, int Kt, const float* d_a,
                                    const float* d_b, const float* d_ksz, float abs_factor, float threshold,
                                    float* d_out, uint64_t* d_nzero) {
  TADEV_REQUIRE(ctx, "tadev_shape_gemm_f32: null ctx");
  TADEV_REQUIRE(Mt >= 0 && Nt >= 0 && Kt >= 0, "tadev_shape_gemm_f32: negative extents");
  if (Mt == 0 || Nt == 0) return TADEV_OK;
  TADEV_REQUIRE(d_a && d_b && d_out && d_nzero && (Kt == 0 || d_ksz), "tadev_shape_gemm_f32: null arrays");
  if (Kt == 0) {
    shape_outer_kernel<<<(unsigned)ceil_div64((int64_t)Mt * Nt, 256), 256, 0, (cudaStream_t)s>>>(
        Mt, Nt, d_a, d_b, abs_factor, threshold, d_out, (unsigned long long*)d_nzero);
  } else {
    dim3 grid((Nt + ST - 1) / ST, (Mt + ST - 1) / ST);
    shape_gemm_kernel<<<grid, ST * ST, 0, (cudaStream_t)s>>>(Mt, Nt, Kt, d_a, d_b, d_ksz, abs_factor, threshold,
                                                            d_out, (unsigned long long*)d_nzero);
  }
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}

extern "C" int tadev_shape_mask_f32(tadev_ctx* ctx, tadev_stream s, int64_t n, float* d_norms, const float* d_mask,
                                    float thr_this, float thr_mask, uint64_t* d_nzero) {
  TADEV_REQUIRE(ctx, "tadev_shape_mask_f32: null ctx");
  if (n <= 0) return TADEV_OK;
  TADEV_REQUIRE(d_norms && d_mask && d_nzero, "tadev_shape_mask_f32: null arrays");
  shape_mask_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)s>>>(n, d_norms, d_mask, thr_this,
                                                                              thr_mask, (unsigned long long*)d_nzero);
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}

extern "C" int tadev_build_pairlist(tadev_ctx* ctx, tadev_stream s, int k, int Pr, int Pc, int r, int c, int Mt,
                                    int Nt, int Kt, const float* d_a, const float* d_b, const float* d_c,
                                    float threshold, int32_t* d_pair_i, int32_t* d_pair_j, int32_t* d_npairs) {
  TADEV_REQUIRE(ctx, "tadev_build_pairlist: null ctx");
  TADEV_REQUIRE(Pr >= 1 && Pc >= 1 && r >= 0 && r < Pr && c >= 0 && c < Pc, "tadev_build_pairlist: bad grid position");
  TADEV_REQUIRE(Mt >= 0 && Nt >= 0 && Kt >= 1 && k >= 0 && k < Kt, "tadev_build_pairlist: bad extents/step");
  TADEV_REQUIRE(d_pair_i && d_pair_j && d_npairs, "tadev_build_pairlist: null outputs");
  pairlist_kernel<<<1, 1024, 0, (cudaStream_t)s>>>(k, Pr, Pc, r, c, Mt, Nt, Kt, d_a, d_b, d_c, threshold, d_pair_i,
                                                   d_pair_j, d_npairs);
  ctx->launches++;
  TADEV_CHECK_CUDA(cudaGetLastError());
  return TADEV_OK;
}
