// fsph_diag.cu -- device micro-benchmarks behind the C ABI: the measured denominators of the fp64 / HBM rooflines.
//
// the pair sweeps are bound by the fp64 pipe, whose peak is not in MEASURED_PEAKS.json; bench.py
// measures it live with fsph_measure_fp64_tflops (independent DFMA chains, every SM busy) and reports the sweeps'
// algorithmic flop rates against it.
#include <algorithm>
#include "fsph_internal.h"

namespace fsph {

// 8 independent fused-multiply-add chains per thread: the fp64 pipe issues back to back (latency hidden by
// 8 chains x 32 warps per SM)
__global__ void __launch_bounds__(256) k_dfma(double * __restrict__ out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
  for(int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) k_copy16(const double2 * __restrict__ in, double2 * __restrict__ out, size_t n) {
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace fsph

using namespace fsph;

extern "C" {

// Sustained fp64 FMA rate of `device` in TFLOP/s (2 flops per DFMA), best of `reps` launches of ~`ms_target` ms.
double fsph_measure_fp64_tflops(int32_t device, int32_t reps) {
  if(cudaSetDevice(device) != cudaSuccess) return -1.0;
  cudaDeviceProp pr;
  if(cudaGetDeviceProperties(&pr, device) != cudaSuccess) return -1.0;
  const int grid = pr.multiProcessorCount * 8, block = 256, iters = 1 << 15;
  double * out = nullptr;
  if(cudaMalloc(&out, (size_t)grid * block * 8) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for(int r = 0; r < std::max(reps, 1) + 1; ++r) {
    cudaEventRecord(e0);
    ++g_launches;
    k_dfma<<<grid, block>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * iters * (double)grid * block / (ms * 1e-3) / 1e12;
    if(r > 0) best = std::max(best, tf);   // the first launch warms up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

// Device copy bandwidth in GB/s (read + write bytes), 128-bit accesses, `bytes` per direction.
double fsph_measure_copy_gbs(int32_t device, int64_t bytes, int32_t reps) {
  if(cudaSetDevice(device) != cudaSuccess) return -1.0;
  cudaDeviceProp pr;
  if(cudaGetDeviceProperties(&pr, device) != cudaSuccess) return -1.0;
  double2 * a = nullptr, * b = nullptr;
  const size_t n = (size_t)bytes / 16;
  if(cudaMalloc(&a, n * 16) != cudaSuccess || cudaMalloc(&b, n * 16) != cudaSuccess) { cudaFree(a); return -1.0; }
  cudaMemset(a, 0, n * 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for(int r = 0; r < std::max(reps, 1) + 1; ++r) {
    cudaEventRecord(e0);
    ++g_launches;
    k_copy16<<<pr.multiProcessorCount * 16, 256>>>(a, b, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if(r > 0) best = std::max(best, 2.0 * n * 16 / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

}  // extern "C"
