// probe.cu — on-box measurement of the FP64 roofline denominators (the FP64 peak is
// not in MEASURED_PEAKS.json and must be measured by the builder). Register-resident loops:
// kind 0 = DMMA.8x8x4 issue rate, kind 1 = DFMA issue rate, kind 2 = both interleaved (tells
// whether the tensor FP64 path and the vector FP64 path share one datapath).
#include "common.h"

namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int KIND>
__global__ void __launch_bounds__(256) probe_kernel(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
  double c[16][2];
  double f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = c[i][1] = 0.0; f[i] = i; }
  for (int it = 0; it < iters; ++it) {
    if (KIND == 0 || KIND == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    if (KIND == 1 || KIND == 2) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f64 %0, %1, %2, %0;\n" : "+d"(f[i]) : "d"(a), "d"(b));
    }
  }
  double sum = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) sum += c[i][0] + c[i][1] + f[i];
  if (sum == 123.456) out[0] = sum;  // keep the loop alive
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}

}  // namespace

extern "C" int tadev_probe_fp64_peak(tadev_ctx* ctx, int kind, int iters, double* tflops, float* ms_out) {
  TADEV_REQUIRE(ctx && tflops, "tadev_probe_fp64_peak: null");
  TADEV_REQUIRE(kind >= 0 && kind <= 2 && iters > 0, "tadev_probe_fp64_peak: bad kind/iters");
  cudaStream_t s = ctx->streams[0];
  double* d_out = nullptr;
  TADEV_CHECK_CUDA(cudaMalloc(&d_out, 64));
  cudaEvent_t e0, e1;
  TADEV_CHECK_CUDA(cudaEventCreate(&e0));
  TADEV_CHECK_CUDA(cudaEventCreate(&e1));
  const int grid = ctx->num_sms * 2, block = 256;
  auto launch = [&](int n) {
    if (kind == 0) probe_kernel<0><<<grid, block, 0, s>>>(d_out, n, 1.0);
    else if (kind == 1) probe_kernel<1><<<grid, block, 0, s>>>(d_out, n, 1.0);
    else probe_kernel<2><<<grid, block, 0, s>>>(d_out, n, 1.0);
    ctx->launches++;
  };
  launch(iters / 8 + 1);  // warm-up
  TADEV_CHECK_CUDA(cudaEventRecord(e0, s));
  launch(iters);
  TADEV_CHECK_CUDA(cudaEventRecord(e1, s));
  TADEV_CHECK_CUDA(cudaEventSynchronize(e1));
  TADEV_CHECK_CUDA(cudaGetLastError());
  float ms = 0;
  TADEV_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  const double warps = (double)grid * block / 32;
  double flop = 0;
  if (kind == 0 || kind == 2) flop += warps * iters * 16.0 * (2.0 * 8 * 8 * 4);
  if (kind == 1 || kind == 2) flop += warps * iters * 8.0 * 16.0 * 32 * 2.0;
  *tflops = flop / (ms * 1e-3) / 1e12;
  if (ms_out) *ms_out = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return TADEV_OK;
}

extern "C" int tadev_probe_copy_gbs(tadev_ctx* ctx, size_t bytes, int iters, double* gbs) {
  TADEV_REQUIRE(ctx && gbs && bytes >= 32 && iters > 0, "tadev_probe_copy_gbs: bad args");
  cudaStream_t s = ctx->streams[0];
  void *a = nullptr, *b = nullptr;
  TADEV_CHECK_CUDA(cudaMalloc(&a, bytes));
  TADEV_CHECK_CUDA(cudaMalloc(&b, bytes));
  TADEV_CHECK_CUDA(cudaMemsetAsync(a, 1, bytes, s));
  cudaEvent_t e0, e1;
  TADEV_CHECK_CUDA(cudaEventCreate(&e0));
  TADEV_CHECK_CUDA(cudaEventCreate(&e1));
  const size_t n = bytes / 16;
  copy_kernel<<<ctx->num_sms * 8, 512, 0, s>>>((const double2*)a, (double2*)b, n);
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) {
    TADEV_CHECK_CUDA(cudaEventRecord(e0, s));
    copy_kernel<<<ctx->num_sms * 8, 512, 0, s>>>((const double2*)a, (double2*)b, n);
    ctx->launches++;
    TADEV_CHECK_CUDA(cudaEventRecord(e1, s));
    TADEV_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    TADEV_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  TADEV_CHECK_CUDA(cudaGetLastError());
  *gbs = 2.0 * (double)(n * 16) / (best * 1e-3) / 1e9;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  return TADEV_OK;
}

// Pinned-host <-> device copy rates: h2d alone, d2h alone, and both directions at once (the e2e
// leg of bench.py streams operands in while result blocks stream out).
extern "C" int tadev_probe_pcie_gbs(tadev_ctx* ctx, size_t bytes, double* h2d, double* d2h, double* h2d_bidir,
                                    double* d2h_bidir) {
  TADEV_REQUIRE(ctx && bytes >= 4096, "tadev_probe_pcie_gbs: bad args");
  cudaStream_t s0 = ctx->comm_stream[1], s1 = ctx->streams[ctx->streams.size() > 1 ? 1 : 0];
  void *h0 = nullptr, *h1 = nullptr, *d0 = nullptr, *d1 = nullptr;
  TADEV_CHECK_CUDA(cudaMallocHost(&h0, bytes));
  TADEV_CHECK_CUDA(cudaMallocHost(&h1, bytes));
  TADEV_CHECK_CUDA(cudaMalloc(&d0, bytes));
  TADEV_CHECK_CUDA(cudaMalloc(&d1, bytes));
  memset(h0, 1, bytes);
  memset(h1, 2, bytes);
  cudaEvent_t e[4];
  for (auto& x : e) TADEV_CHECK_CUDA(cudaEventCreate(&x));
  auto timed = [&](bool up, bool down, float* t_up, float* t_down) -> int {
    TADEV_CHECK_CUDA(cudaDeviceSynchronize());
    if (up) { TADEV_CHECK_CUDA(cudaEventRecord(e[0], s0)); TADEV_CHECK_CUDA(cudaMemcpyAsync(d0, h0, bytes, cudaMemcpyHostToDevice, s0)); TADEV_CHECK_CUDA(cudaEventRecord(e[1], s0)); }
    if (down) { TADEV_CHECK_CUDA(cudaEventRecord(e[2], s1)); TADEV_CHECK_CUDA(cudaMemcpyAsync(h1, d1, bytes, cudaMemcpyDeviceToHost, s1)); TADEV_CHECK_CUDA(cudaEventRecord(e[3], s1)); }
    TADEV_CHECK_CUDA(cudaDeviceSynchronize());
    if (up) TADEV_CHECK_CUDA(cudaEventElapsedTime(t_up, e[0], e[1]));
    if (down) TADEV_CHECK_CUDA(cudaEventElapsedTime(t_down, e[2], e[3]));
    return TADEV_OK;
  };
  float tu = 0, td = 0;
  int rc = timed(true, true, &tu, &td);  // warm-up
  if (!rc) rc = timed(true, false, &tu, &td);
  if (!rc && h2d) *h2d = bytes / (tu * 1e-3) / 1e9;
  if (!rc) rc = timed(false, true, &tu, &td);
  if (!rc && d2h) *d2h = bytes / (td * 1e-3) / 1e9;
  if (!rc) rc = timed(true, true, &tu, &td);
  if (!rc && h2d_bidir) *h2d_bidir = bytes / (tu * 1e-3) / 1e9;
  if (!rc && d2h_bidir) *d2h_bidir = bytes / (td * 1e-3) / 1e9;
  for (auto& x : e) cudaEventDestroy(x);
  cudaFreeHost(h0); cudaFreeHost(h1); cudaFree(d0); cudaFree(d1);
  return rc;
}
