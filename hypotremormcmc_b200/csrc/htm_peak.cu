// FFMA and MUFU throughput microbenchmarks: the roofline denominators for this path.  The
// driver measures HBM and bf16 GEMM peaks only (MEASURED_PEAKS.json); the hot path is bound
// by the FP32 FMA pipe co-limited by MUFU, so we measure those here.
#include "htm_kernels.hpp"

namespace htm {

// 8 independent FFMA chains per thread, 3-register form (operands are not immediates)
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, float a, float b, int iters) {
  float v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      v0 = fmaf(v0, a, b);
      v1 = fmaf(v1, a, b);
      v2 = fmaf(v2, a, b);
      v3 = fmaf(v3, a, b);
      v4 = fmaf(v4, a, b);
      v5 = fmaf(v5, a, b);
      v6 = fmaf(v6, a, b);
      v7 = fmaf(v7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

// 8 independent MUFU.RSQ chains per thread
__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, int iters) {
  float v0 = threadIdx.x + 1.5f, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      v0 = rsqrtf(v0);
      v1 = rsqrtf(v1);
      v2 = rsqrtf(v2);
      v3 = rsqrtf(v3);
      v4 = rsqrtf(v4);
      v5 = rsqrtf(v5);
      v6 = rsqrtf(v6);
      v7 = rsqrtf(v7);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

// 8 independent DFMA chains per thread: the denominator of the float64 (validation) kernels
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, double a, double b, int iters) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      v0 = fma(v0, a, b);
      v1 = fma(v1, a, b);
      v2 = fma(v2, a, b);
      v3 = fma(v3, a, b);
      v4 = fma(v4, a, b);
      v5 = fma(v5, a, b);
      v6 = fma(v6, a, b);
      v7 = fma(v7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

cudaError_t measure_fp64_peak(int device, double* tflops) {
  cudaError_t err = cudaSetDevice(device);
  if (err != cudaSuccess) return err;
  cudaDeviceProp prop;
  err = cudaGetDeviceProperties(&prop, device);
  if (err != cudaSuccess) return err;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1024;
  double* out = nullptr;
  err = cudaMalloc(&out, sizeof(double) * blocks * threads);
  if (err != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-7, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess) return err;
  if (tflops) *tflops = static_cast<double>(blocks) * threads * iters * 8.0 * 8.0 * 2.0 / (best * 1e-3) / 1e12;
  return cudaSuccess;
}

cudaError_t measure_fp32_peak(int device, double* tflops, double* mufu_gops) {
  cudaError_t err = cudaSetDevice(device);
  if (err != cudaSuccess) return err;
  cudaDeviceProp prop;
  err = cudaGetDeviceProperties(&prop, device);
  if (err != cudaSuccess) return err;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  float* out = nullptr;
  err = cudaMalloc(&out, sizeof(float) * blocks * threads);
  if (err != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best_f = 1e30f, best_m = 1e30f;
  const int it_f = 4096, it_m = 1024;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    ffma_peak_kernel<<<blocks, threads>>>(out, 1.0000001f, 1e-7f, it_f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best_f) best_f = ms;
    cudaEventRecord(e0);
    mufu_peak_kernel<<<blocks, threads>>>(out, it_m);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best_m) best_m = ms;
  }
  err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess) return err;
  const double n_thr = static_cast<double>(blocks) * threads;
  if (tflops) *tflops = n_thr * it_f * 16.0 * 8.0 * 2.0 / (best_f * 1e-3) / 1e12;
  if (mufu_gops) *mufu_gops = n_thr * it_m * 8.0 * 8.0 / (best_m * 1e-3) / 1e9;
  return cudaSuccess;
}

}  // namespace htm
