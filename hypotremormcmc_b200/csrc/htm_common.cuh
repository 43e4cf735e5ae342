// Device-side building blocks shared by every kernel of libhtm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/htm_b200.h"

namespace htm {

constexpr double kEps64 = 2.220446049250313e-16;  // epsilon(1.d0), src/cls_mcmc.f90:184
constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kFreq = 5.0;                     // src/cls_forward.f90:190,234
constexpr double kLog2PiHalf = 0.91893853320467274178;  // 0.5*log(2*pi), src/cls_forward.f90:5

// Philox counter word c2 ("purpose"); the test oracle uses the same numbering
enum : uint32_t { PHX_STEP = 0, PHX_SWAP = 1, PHX_INIT = 2, PHX_GLOBAL = 3, PHX_TEMP = 4 };

// ---- Philox4x32-10 (replaces mod_random in modes B and C) ---------------------------------
struct u32x4 {
  uint32_t v[4];
};
__device__ __forceinline__ u32x4 philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                               uint32_t c3) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  u32x4 o;
  o.v[0] = c0;
  o.v[1] = c1;
  o.v[2] = c2;
  o.v[3] = c3;
  return o;
}
// Same generator with the 10 round keys precomputed on the host (they live in the kernel's
// constant bank, so the key schedule costs no instructions and no registers).
struct PhiloxKeys {
  uint32_t k[20];
};
inline PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys r;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    r.k[2 * i] = k0;
    r.k[2 * i + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return r;
}
__device__ __forceinline__ u32x4 philox4x32_10(const PhiloxKeys& rk, uint32_t c0, uint32_t c1, uint32_t c2,
                                               uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ rk.k[2 * r];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ rk.k[2 * r + 1];
    c3 = lo0;
  }
  u32x4 o;
  o.v[0] = c0;
  o.v[1] = c1;
  o.v[2] = c2;
  o.v[3] = c3;
  return o;
}
// int(u*n), u = 24-bit uniform in [0,1)
__device__ __forceinline__ uint32_t below(uint32_t w, uint32_t n) {
  return static_cast<uint32_t>((static_cast<uint64_t>(w >> 8) * n) >> 24);
}

// ---- per-precision math ----------------------------------------------------------------------
// float: MUFU-based approximations (rsqrt, lg2, cos) -- the throughput path.
// double: IEEE sqrt/div and libdevice log/cos (<= 1 ulp) -- the parity path.
template <typename real>
struct M;

// raw MUFU approximations, flush-to-zero: no denormal fix-up code around them (inputs here are
// squared distances, uniforms in (0,1) and temperatures -- never denormal)
__device__ __forceinline__ float mufu_rsq(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_cos(float x) {
  float y;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <>
struct M<float> {
  typedef float4 real4;
  static constexpr float kLn2 = 0.6931471805599453f;
  static __device__ __forceinline__ float u_co(uint32_t w) { return static_cast<float>(w >> 8) * (1.0f / 16777216.0f); }
  static __device__ __forceinline__ float u_oo(uint32_t w) {
    return (static_cast<float>(w >> 9) + 0.5f) * (1.0f / 8388608.0f);
  }
  static __device__ __forceinline__ float log(float x) { return kLn2 * mufu_lg2(x); }
  static __device__ __forceinline__ float sqrt(float x) { return mufu_sqrt(x); }
  static __device__ __forceinline__ float gauss(uint32_t wa, uint32_t wb) {
    // sqrt(-2 ln u1) cos(2 pi u2) = sqrt(-2 ln2 lg2 u1) cos(2 pi u2)
    const float r = mufu_sqrt(-2.0f * kLn2 * mufu_lg2(u_oo(wa)));
    // angle 2 pi u2 with the uniform's scale folded into one multiply
    return r * mufu_cos((static_cast<float>(wb >> 9) + 0.5f) * (6.283185307179586f / 8388608.0f));
  }
  // distance and ln(distance) from the squared distance: two independent MUFU ops
  static __device__ __forceinline__ void dist(float d2, float& d, float& lnd) {
    d = d2 * mufu_rsq(d2);
    lnd = 0.34657359027997264f * mufu_lg2(d2);  // 0.5*ln2*lg2(d2)
  }
  static __device__ __forceinline__ float rcp(float x) { return mufu_rcp(x); }
  static __device__ __forceinline__ float exp(float x) { return mufu_ex2(x * 1.4426950408889634f); }
  // a / b where b's reciprocal ib is already known
  static __device__ __forceinline__ float div(float a, float /*b*/, float ib) { return a * ib; }
};

template <>
struct M<double> {
  typedef double4 real4;
  static __device__ __forceinline__ double u_co(uint32_t w) { return static_cast<double>(w >> 8) * (1.0 / 16777216.0); }
  static __device__ __forceinline__ double u_oo(uint32_t w) {
    return (static_cast<double>(w >> 9) + 0.5) * (1.0 / 8388608.0);
  }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double gauss(uint32_t wa, uint32_t wb) {
    const double pi2 = 2.0 * kPi;
    return ::sqrt(-2.0 * ::log(u_oo(wa))) * ::cos(pi2 * u_oo(wb));
  }
  static __device__ __forceinline__ void dist(double d2, double& d, double& lnd) {
    d = ::sqrt(d2);
    lnd = ::log(d);
  }
  static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  // the reference divides (src/cls_mcmc.f90:194); keep the IEEE division in the parity path
  static __device__ __forceinline__ double div(double a, double b, double /*ib*/) { return a / b; }
};

// ---- warp helpers --------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 1-D bulk TMA (cp.async.bulk global -> shared, completes on an mbarrier) -------------------
// SASS: UBLKCP + SYNCS.  Bytes must be a multiple of 16; both addresses 16-byte aligned.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
               "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
      "l"(gmem_src), "r"(bytes), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace htm
