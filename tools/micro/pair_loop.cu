// The lane kernel's packed station loop (htm_forward.cuh: forward_pairs) in isolation: cycles per station pair with
// 1-5 warps per scheduler, one or two chains per lane, against the MUFU and FMA pipe rates measured the same way.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DHTM_PK_SQRT=1] -o pair_loop pair_loop.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../../hypotremormcmc_b200/csrc/htm_forward.cuh"

using namespace htm;
constexpr int kPairs = 100, kIters = 300;

template <int NSLOT>
__global__ void __launch_bounds__(128) loop_kernel(float* out, long long* cyc, float seed) {
  __shared__ float4 s_pk[4 * kPairs];
  for (int m = threadIdx.x; m < kPairs; m += blockDim.x) {
    float4 st0 = make_float4(3.f + m, -2.f + 0.5f * m, 0.1f, 0.f), st1 = make_float4(-4.f - m, 1.f + 0.3f * m, 0.2f, 0.f);
    float4 ob0 = make_float4(0.3f * m, 100.f, -0.1f * m, 25.f), ob1 = make_float4(-0.2f * m, 80.f, 0.05f * m, 30.f);
    store_station_pair(s_pk + 4 * m, expand_station(st0, ob0, 0.5f, 0.5f), expand_station(st1, ob1, 0.5f, 0.5f));
  }
  __syncthreads();
  const Glob<float> g = make_glob<float>(3.5f * seed, 250.f);
  float hx[NSLOT], hy[NSLOT], hz[NSLOT], nct[NSLOT], nca[NSLOT], S1t[NSLOT], S1a[NSLOT], S2[NSLOT];
#pragma unroll
  for (int q = 0; q < NSLOT; ++q) {
    hx[q] = 0.01f * threadIdx.x + q;
    hy[q] = -0.02f * threadIdx.x;
    hz[q] = 20.f + q;
    nct[q] = 0.f;
    nca[q] = 0.f;
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    forward_pairs<NSLOT>(s_pk, kPairs, hx, hy, hz, g, nct, nca, S1t, S1a, S2);
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      hx[q] += 1e-9f * S2[q];
      nct[q] = 1e-9f * S1t[q];
      nca[q] = 1e-9f * S1a[q];
    }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int q = 0; q < NSLOT; ++q) acc += hx[q] + nct[q] + nca[q];
  if (acc == 12345.678f) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// MODE 0: MUFU.RSQ, 1: MUFU.LG2, 2: MUFU.SQRT (8 chains), 3: FFMA2 (16 chains), 4: FFMA (16 chains)
template <int MODE>
__global__ void __launch_bounds__(128) pipe_kernel(float* out, long long* cyc, float seed) {
  float v[16];
  float2 p[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = seed + threadIdx.x + i;
    p[i] = make_float2(v[i], -v[i]);
  }
  const float2 a = make_float2(seed, seed), b = make_float2(1e-7f * seed, 1e-7f);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters * 20; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < (MODE < 3 ? 8 : 16); ++i) {
        if (MODE == 0) v[i] = mufu_rsq(v[i]);
        if (MODE == 1) v[i] = mufu_lg2(v[i]);
        if (MODE == 2) v[i] = mufu_sqrt(v[i]);
        if (MODE == 3) p[i] = __ffma2_rn(p[i], a, b);
        if (MODE == 4) v[i] = fmaf(v[i], a.x, b.x);
      }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += v[i] + p[i].x + p[i].y;
  if (acc == 12345.678f) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
void sweep(const char* name, K kern, double units, int wmax) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("%-28s", name);
  for (int w = 1; w <= wmax; ++w) {
    kern<<<148 * w, 128>>>(out, cyc, 1.0f);
    cudaEventRecord(e0);
    kern<<<148 * w, 128>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8];
    cudaMemcpy(h, cyc, 148 * w * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0, mx = 0;
    for (int i = 0; i < 148 * w; ++i) {
      s += h[i];
      if (h[i] > mx) mx = h[i];
    }
    // cycles of one scheduler per unit: mean CTA clock64 span / (units per warp * w); the event time at 1965 MHz
    printf(" | W=%d %6.2f (max %6.2f, event %6.2f)", w, s / (148.0 * w) / units / w, mx / units / w,
           ms * 1e-3 * 1.965e9 / units / w);
  }
  printf("\n");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  printf("cycles of one scheduler per unit (unit = one station pair for all slots of a warp, or one instruction)\n");
  sweep("forward_pairs<1> per pair", loop_kernel<1>, double(kPairs) * kIters, 5);
  sweep("forward_pairs<2> per pair", loop_kernel<2>, double(kPairs) * kIters, 4);
  sweep("MUFU.RSQ", pipe_kernel<0>, 16.0 * kIters * 20, 8);
  sweep("MUFU.LG2", pipe_kernel<1>, 16.0 * kIters * 20, 8);
  sweep("MUFU.SQRT", pipe_kernel<2>, 16.0 * kIters * 20, 8);
  sweep("FFMA2", pipe_kernel<3>, 32.0 * kIters * 20, 8);
  sweep("FFMA", pipe_kernel<4>, 32.0 * kIters * 20, 8);
  return 0;
}
