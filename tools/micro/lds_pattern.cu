// Shared-memory wavefront cost of 16-byte (and 8-byte) loads for different lane -> row patterns.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_pattern lds_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kRow = 103;   // float4 per row (odd)
constexpr int kIters = 4096;

template <int PAT, int WIDTH>
__global__ void __launch_bounds__(256) k(float* out) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 32 * kRow; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int row;
  switch (PAT) {
    case 0: row = 0; break;
    case 1: row = lane; break;
    case 2: row = lane & 7; break;
    case 3: row = lane >> 2; break;
    case 4: row = lane >> 3; break;
    case 5: row = lane & 3; break;
    case 6: row = lane >> 1; break;
    case 7: row = lane & 15; break;
    default: row = 0;
  }
  const float4* base = sm + row * kRow;
  float acc = 0.f;
  unsigned addr0 = static_cast<unsigned>(__cvta_generic_to_shared(base)) + (threadIdx.x >> 5) * 16;
  unsigned addr = addr0;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (WIDTH == 16) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr + u * 16));
        acc += (v.x + v.y) + (v.z + v.w);
      } else {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr + u * 16));
        acc += v.x + v.y;
      }
    }
    addr = (it & 7) == 7 ? addr0 : addr + 128;
  }
  if (acc == 12345.678f) out[0] = acc;
}

template <int PAT, int WIDTH>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 4);
  const size_t smem = 32 * kRow * sizeof(float4);
  cudaFuncSetAttribute(k<PAT, WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<PAT, WIDTH><<<148 * 3, 256, smem>>>(out);
  cudaEventRecord(a);
  k<PAT, WIDTH><<<148 * 3, 256, smem>>>(out);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  // warp-level loads per SM: 3 CTAs * 8 warps * kIters * 8
  const double loads = 3.0 * 8 * kIters * 8;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-34s width %2d: %.3f ms, %.2f SM-cycles per warp load (at %d MHz nominal)\n", name, WIDTH, ms,
         ms * 1e-3 * clk * 1e3 / loads, clk / 1000);
  cudaFree(out);
}

int main() {
  run<0, 16>("all lanes one row");
  run<1, 16>("32 distinct rows");
  run<2, 16>("row = lane & 7");
  run<3, 16>("row = lane >> 2");
  run<4, 16>("row = lane >> 3");
  run<5, 16>("row = lane & 3");
  run<6, 16>("row = lane >> 1");
  run<7, 16>("row = lane & 15");
  run<0, 8>("all lanes one row");
  run<1, 8>("32 distinct rows");
  run<2, 8>("row = lane & 7");
  run<3, 8>("row = lane >> 2");
  run<7, 8>("row = lane & 15");
  run<6, 8>("row = lane >> 1");
  return 0;
}
