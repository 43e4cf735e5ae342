// Issue / pipe cost of the lane kernel's instruction mix on one SM sub-partition: packed FFMA2, scalar FFMA, MUFU,
// integer ALU and LDS.128 streams alone and interleaved, with 1-4 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 20000;

#define F2(acc, a, b) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b))
#define F1(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))
#define MU(y, x) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(y))
#define XR(x, y) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define LD(v, addr) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr))

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, long long* cyc, float seed) {
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, 1.f, 2.f, 3.f);
  __syncthreads();
  unsigned long long p[8], a = __double_as_longlong(1.0000001), b = __double_as_longlong(0.9999999);
  float s[16], m[4], x = seed + threadIdx.x;
  unsigned u[6], v = threadIdx.x;
  float4 l0, l1;
  const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(sm));
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = seed * i;
#pragma unroll
  for (int i = 0; i < 6; ++i) u[i] = i * 77 + threadIdx.x;
  m[0] = x; m[1] = x + 1.f; m[2] = x + 2.f; m[3] = x + 3.f;
  l0 = l1 = make_float4(0, 0, 0, 0);
  const float sa = 1.0000001f, sb = 0.5f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    if (MODE == 0) {  // 16 FFMA2
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) F2(p[i], a, b);
    } else if (MODE == 1) {  // 16 FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) F1(s[i], sa, sb);
    } else if (MODE == 2) {  // 8 FFMA2 + 8 XOR
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        F2(p[i], a, b);
        XR(u[i % 6], v);
      }
    } else if (MODE == 3) {  // 16 MUFU
#pragma unroll
      for (int i = 0; i < 16; ++i) MU(m[i & 3], x);
    } else if (MODE == 4 || MODE == 6 || MODE == 9) {  // 14 FFMA2 + 4 MUFU (+ 2 LDS.128) (+ 6 XOR)
      if (MODE == 6) LD(l0, addr);
      F2(p[0], a, b); F2(p[1], a, b); F2(p[2], a, b); F2(p[3], a, b);
      MU(m[0], x);
      if (MODE == 9) { XR(u[0], v); XR(u[1], v); }
      F2(p[4], a, b); F2(p[5], a, b); F2(p[6], a, b);
      MU(m[1], x);
      if (MODE == 6) LD(l1, addr + 16);
      if (MODE == 9) { XR(u[2], v); XR(u[3], v); }
      F2(p[7], a, b); F2(p[0], a, b); F2(p[1], a, b); F2(p[2], a, b);
      MU(m[2], x);
      if (MODE == 9) { XR(u[4], v); XR(u[5], v); }
      F2(p[3], a, b); F2(p[4], a, b); F2(p[5], a, b);
      MU(m[3], x);
    } else if (MODE == 5) {  // 28 FFMA + 4 MUFU
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int i = 0; i < 7; ++i) F1(s[(g * 7 + i) & 15], sa, sb);
        MU(m[g], x);
      }
    } else if (MODE == 7) {  // 12 FFMA2 + 4 MUFU
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        F2(p[(3 * g) & 7], a, b); F2(p[(3 * g + 1) & 7], a, b); F2(p[(3 * g + 2) & 7], a, b);
        MU(m[g], x);
      }
    } else if (MODE == 8) {  // 10 FFMA2 + 4 MUFU
      F2(p[0], a, b); F2(p[1], a, b); F2(p[2], a, b);
      MU(m[0], x);
      F2(p[3], a, b); F2(p[4], a, b);
      MU(m[1], x);
      F2(p[5], a, b); F2(p[6], a, b); F2(p[7], a, b);
      MU(m[2], x);
      F2(p[0], a, b); F2(p[1], a, b);
      MU(m[3], x);
    }
  }
  const long long t1 = clock64();
  float acc = m[0] + m[1] + m[2] + m[3] + l0.x + l1.y;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += __longlong_as_double(p[i]) > 1e300 ? 1.f : 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += s[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) acc += u[i] == 0x12345u ? 1.f : 0.f;
  if (acc == 12345.678f) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NF, int NM, int NS>
__global__ void __launch_bounds__(128) kmix(float* out, long long* cyc, float seed) {
  unsigned long long p[8], a = __double_as_longlong(1.0000001), b = __double_as_longlong(0.9999999);
  float s[16], m[8], x = seed + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = seed * i;
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = x + i;
  const float sa = 1.0000001f, sb = 0.5f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    // Bresenham interleave of the three streams over NF + NM + NS slots
    constexpr int N = NF + NM + NS;
    int f = 0, mm = 0, sc = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      // pick the stream that is furthest behind its share
      const int df = NF ? (i + 1) * NF - f * N : -1000000, dm = NM ? (i + 1) * NM - mm * N : -1000000,
                ds = NS ? (i + 1) * NS - sc * N : -1000000;
      if (dm >= df && dm >= ds) {
        MU(m[mm & 7], x);
        ++mm;
      } else if (df >= ds) {
        F2(p[f & 7], a, b);
        ++f;
      } else {
        F1(s[sc & 15], sa, sb);
        ++sc;
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += m[i] + (__longlong_as_double(p[i]) > 1e300 ? 1.f : 0.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += s[i];
  if (acc == 12345.678f) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// FFMA2 operand sourcing: PAT 0: acc += a * b with a, b loop-invariant (one lands in a uniform register);
// 1: acc_i += q_i * r_i, three distinct 64-bit register operands; 2: acc_i += q_i * q_i (two); 3: acc_i = acc_i * q_i + acc_i
template <int PAT, int NM>
__global__ void __launch_bounds__(128) kops(float* out, long long* cyc, float seed) {
  unsigned long long p[8], q[8], r[8];
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    p[i] = __double_as_longlong(1.0 + threadIdx.x * 1e-3 + i);
    q[i] = __double_as_longlong(0.5 + threadIdx.x * 1e-4 + i * seed);
    r[i] = __double_as_longlong(0.25 + threadIdx.x * 1e-5 + i * seed);
    m[i] = seed + threadIdx.x + i;
  }
  unsigned long long a = q[0], b = __double_as_longlong((double)seed);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (PAT == 0) F2(p[i & 7], a, b);
      if (PAT == 1) F2(p[i & 7], q[i & 7], r[(i + 3) & 7]);
      if (PAT == 2) F2(p[i & 7], q[i & 7], q[i & 7]);
      if (PAT == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(p[i & 7]) : "l"(q[i & 7]));
      if (NM && (i % (16 / NM)) == 0) MU(m[(i / (16 / NM)) & 7], 0);
    }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += m[i] + (__longlong_as_double(p[i]) > 1e300 ? 1.f : 0.f) + (__longlong_as_double(q[i] ^ r[i]) > 1e300 ? 1.f : 0.f);
  if (acc == 12345.678f) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int PAT, int NM>
void runops() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("16 FFMA2 operand pattern %d + %d MUFU: cycles of one scheduler per body (event time at 1965 MHz)", PAT, NM);
  for (int w = 1; w <= 6; ++w) {
    kops<PAT, NM><<<148 * w, 128>>>(out, cyc, 1.5f);
    cudaEventRecord(e0);
    kops<PAT, NM><<<148 * w, 128>>>(out, cyc, 1.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  W=%d %6.1f", w, ms * 1e-3 * 1.965e9 / kIters / w);
  }
  printf("\n");
  cudaFree(out);
  cudaFree(cyc);
}

template <int NF, int NM, int NS>
void runmix() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("%2d FFMA2 + %2d MUFU + %2d FFMA: cycles of one scheduler per body (event time at 1965 MHz)", NF, NM, NS);
  for (int w = 1; w <= 6; ++w) {
    kmix<NF, NM, NS><<<148 * w, 128>>>(out, cyc, 1.5f);
    cudaEventRecord(e0);
    kmix<NF, NM, NS><<<148 * w, 128>>>(out, cyc, 1.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  W=%d %6.1f", w, ms * 1e-3 * 1.965e9 / kIters / w);
  }
  printf("\n");
  cudaFree(out);
  cudaFree(cyc);
}

template <int MODE>
void run(const char* name, int n_instr) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, 148 * 4 * sizeof(long long));
  printf("%-34s", name);
  for (int w = 1; w <= 4; ++w) {
    k<MODE><<<148 * w, 128>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    k<MODE><<<148 * w, 128>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, cyc, 148 * w * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148 * w; ++i) s += h[i];
    const double per_body = s / (148.0 * w) / kIters;   // cycles a warp needs per body
    printf("  W=%d: %6.1f cyc/body/warp = %5.2f cyc/instr/scheduler", w, per_body, per_body / w / n_instr);
  }
  printf("\n");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  runops<0, 0>();
  runops<1, 0>();
  runops<2, 0>();
  runops<3, 0>();
  runops<0, 4>();
  runops<1, 4>();
  runops<3, 4>();

  runmix<14, 0, 0>();
  runmix<14, 1, 0>();
  runmix<14, 2, 0>();
  runmix<14, 4, 0>();
  runmix<14, 6, 0>();
  runmix<12, 4, 0>();
  runmix<10, 4, 0>();
  runmix<7, 4, 0>();
  runmix<0, 4, 0>();
  runmix<0, 4, 28>();
  runmix<0, 4, 20>();
  runmix<0, 4, 14>();
  runmix<0, 0, 28>();
  runmix<7, 4, 14>();

  run<0>("16 FFMA2", 16);
  run<1>("16 FFMA", 16);
  run<2>("8 FFMA2 + 8 XOR", 16);
  run<3>("16 MUFU", 16);
  run<4>("14 FFMA2 + 4 MUFU", 18);
  run<5>("28 FFMA + 4 MUFU", 32);
  run<6>("14 FFMA2 + 4 MUFU + 2 LDS.128", 20);
  run<7>("12 FFMA2 + 4 MUFU", 16);
  run<8>("10 FFMA2 + 4 MUFU", 14);
  run<9>("14 FFMA2 + 4 MUFU + 6 XOR", 24);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
