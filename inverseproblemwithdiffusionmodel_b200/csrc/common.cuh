// Shared host/device helpers for the ipdm_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/ipdm_b200.h"

namespace ipdm {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  return 0;
}

// Once-per-device guard for cudaFuncSetAttribute and similar per-device set-up: bit d of `done` = device d is set up.
// Safe from any number of host threads (setting an attribute twice is harmless, skipping it is not).
inline bool device_needs_setup(std::atomic<unsigned long long>& done, int* dev_out = nullptr) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev_out) *dev_out = dev;
  return ((done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) == 0;
}
inline void device_setup_done(std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaGetDevice(&dev);
  done.fetch_or(1ull << (dev & 63), std::memory_order_release);
}

#define IPDM_REQUIRE(cond, code, ...)     \
  do {                                    \
    if (!(cond)) {                        \
      ::ipdm::set_error(__VA_ARGS__);     \
      return (code);                      \
    }                                     \
  } while (0)

#define IPDM_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::ipdm::set_error("%s: %s", #call, cudaGetErrorString(e__));             \
      return static_cast<int>(e__);                                            \
    }                                                                          \
  } while (0)

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }
// ELU whose result is about to be rounded to f16: exp via the SFU (absolute error ~1e-7 near 0)
__device__ __forceinline__ float elu_f16bound(float v) { return v > 0.f ? v : __expf(v) - 1.0f; }

// f32 x2 -> f16 x2 with saturation: an activation beyond the f16 range degrades to +-65504 instead of turning the
// rest of the network into inf/NaN (trained checkpoints are not available to prove the range).
__device__ __forceinline__ unsigned pack_half2_sat(float a, float b) {
  unsigned d;   // one F2FP.SATFINITE.F16.F32.PACK_AB: low half = a, high half = b, |x| > 65504 -> +-65504
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// ---- Philox4x32-10 (Salmon et al.) + Box-Muller: two N(0,1) per call -------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 23 random bits -> a float strictly inside (0,1): (k + 0.5) 2^-23 is exact for every k < 2^23 (min 2^-24, max 1 - 2^-24).
// With 24 bits the top value k + 0.5 = 2^24 - 0.5 is not a float, rounds to 2^24 and yields u = 1: log u = 0, and
// a Box-Muller radius computed as a * rsqrt(a) turns into 0 * inf = NaN about once per 1.7e7 draws.
__device__ __forceinline__ float u01_open(uint32_t r) {
  return (static_cast<float>(r >> 9) + 0.5f) * (1.0f / 8388608.0f);
}

// ---- noise streams keyed by CHAIN -------------------------------------------------------------------------------
// chain i draws from Philox(key = seed, counter = (position inside the chain, chain id, step, tag)), so a
// chain's noise does not depend on the rank it runs on, on the number of ranks or on its slot in the rank's batch
// (reference draw site: ncsn/models/ALD_optimizers.py:238-241, per-sample torch.randn_like).
struct RngArgs {
  uint64_t seed;
  const unsigned long long* seed_dev;   // optional: XORed into seed at run time (a captured graph replays with fresh streams)
  uint32_t step;                        // step counter (added to *cursor when the step comes from a device schedule)
  int chain_base;                       // chain id of sample i = chain_ids ? chain_ids[i] : chain_base + i
  const int* chain_ids;
};
__device__ __forceinline__ uint64_t rng_seed(const RngArgs& r) { return r.seed_dev ? (r.seed ^ *r.seed_dev) : r.seed; }
__device__ __forceinline__ uint32_t rng_chain(const RngArgs& r, size_t i) {
  return (uint32_t)(r.chain_ids ? r.chain_ids[i] : r.chain_base + (int)i);
}

// (n0, n1) ~ N(0,1) for the element pair `pair` of chain `chain` at step `step` under `seed` (generic Langevin kernel).
__device__ __forceinline__ float2 philox_normal2(uint64_t seed, uint32_t chain, uint64_t pair, uint32_t step) {
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(pair), chain, step, 0x1BD11BDAu ^ static_cast<uint32_t>(pair >> 32),
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const float u0 = u01_open(r[0]), u1 = u01_open(r[1]);
  const float rad = sqrtf(fmaxf(-2.0f * logf(u0), 0.0f));
  float s, c;
  sincospif(2.0f * u1, &s, &c);
  return make_float2(rad * c, rad * s);
}

// Four N(0,1) per Philox call (both Box-Muller pairs), SFU log / sincos: the noise of the complex pixels `pix` and
// `pix + W/2` of one image row -- out = (re, im, re', im').  Every fused-step kernel family (pruned, two-pass,
// Stockham) pairs pixels the same way, so the in-kernel noise of a chain does not depend on which one runs.
__device__ __forceinline__ void philox_chain_normal4(uint64_t seed, uint32_t chain, uint32_t pix, uint32_t step, float out[4]) {
  uint32_t r[4];
  philox4x32_10(pix, chain, step, 0x1BD11BDBu, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float u0 = u01_open(r[2 * p]), u1 = u01_open(r[2 * p + 1]);
    const float a2 = fmaxf(-2.0f * __logf(u0), 1e-12f);   // the SFU log of u0 = 1 - 2^-24 may come back as 0: keep rsqrt finite
    const float rad = a2 * rsqrtf(a2);
    float s, c;
    __sincosf(6.283185307179586f * u1, &s, &c);
    out[2 * p] = rad * c;
    out[2 * p + 1] = rad * s;
  }
}

}  // namespace ipdm
