// wrnn_device.cuh -- device helpers shared by the engines: memory-ordering primitives,
// the grid barrier with a watchdog, Philox4x32-10, and the two samplers
// (reference utils/distribution.py:87-123 and models/fatchord_version.py:231-237).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wrnn {

// -------------------------------------------------------------------------------------
// memory ordering
// -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Spin budget before a waiter declares the grid dead and raises the abort flag.  SM clock
// is <= ~2 GHz, so 2^32 cycles is > 2 s: far beyond any legitimate wait (a step is ~10 us)
// and far below gpurun's own limits, so a protocol bug ends in an error, not a hung box.
constexpr long long kWatchdogCycles = 1ll << 32;

// Grid-wide barrier over a monotonically increasing counter: the k-th barrier of a launch
// waits for counter >= k * gridDim.x.  Returns false when the launch must be abandoned.
// The comparison is wrap-safe (signed distance): the 32-bit counter may roll over on very long unbatched jobs
// (6 barriers x 128 CTAs per step wrap after ~5.6 M steps); CTAs are never more than one barrier apart, so the
// distance between counter and target is always far below 2^31.
__device__ __forceinline__ bool counter_behind(unsigned value, unsigned target) { return (int)(value - target) < 0; }
__device__ __forceinline__ bool grid_barrier(unsigned* counter, unsigned target, int* abort_flag) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_add_u32(counter, 1u);   // release: orders this CTA's prior writes (made visible to
                                        // thread 0 by the bar.sync above) before the increment
    int ok = 1;
    const long long t0 = clock64();
    while (counter_behind(ld_acquire_u32(counter), target)) {
      if (clock64() - t0 > kWatchdogCycles || ld_relaxed_s32(abort_flag) != 0) {
        atomicExch(abort_flag, 1);
        ok = 0;
        break;
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// -------------------------------------------------------------------------------------
// numerics
// -------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// RNE to IEEE half, saturating to +-65504 (cvt.rn.satfinite.f16.f32), back to float
__device__ __forceinline__ float f16_round(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// PyTorch GRUCell update for one hidden unit (gate order r, z, n).
__device__ __forceinline__ float gru_unit(float gi_r, float gi_z, float gi_n, float gh_r, float gh_z,
                                          float gh_n, float h) {
  const float r = sigmoidf_(gi_r + gh_r);
  const float z = sigmoidf_(gi_z + gh_z);
  const float n = tanhf(gi_n + r * gh_n);
  return (1.0f - z) * n + z * h;
}

// -------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: (seed) x (step, fold, lane, offset)
// -------------------------------------------------------------------------------------
struct Philox4 { unsigned x, y, z, w; };
__device__ __forceinline__ Philox4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                                  unsigned k0, unsigned k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
// 24 random bits -> open interval (0,1)
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// the reference's uniform_(1e-5, 1 - 1e-5): rand*(to-from)+from
__device__ __forceinline__ float u_ref_range(unsigned x) { return u01(x) * (1.0f - 2e-5f) + 1e-5f; }

// -------------------------------------------------------------------------------------
// samplers
// -------------------------------------------------------------------------------------
// utils/distribution.py:99-121.  l: 30 logits = [10 mixture | 10 means | 10 log-scales].
// `lget(i)` fetches logit i, `uget(i)` uniform i (0..9 mixture draws, 10 = logistic draw).
template <typename LGet, typename UGet>
__device__ __forceinline__ float mol_sample(LGet lget, UGet uget) {
  int best = 0;
  float bestv = -INFINITY;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float g = lget(k) - logf(-logf(uget(k)));            // :107
    if (g > bestv) { bestv = g; best = k; }                     // :108 first maximum wins
  }
  const float mean = lget(10 + best);                           // :113
  const float log_scale = fmaxf(lget(20 + best), -32.23619130191664f);   // :114 log(1e-14)
  const float u = uget(10);
  float x = mean + expf(log_scale) * (logf(u) - logf(1.0f - u));         // :119
  return fminf(fmaxf(x, -1.0f), 1.0f);                          // :121
}

}  // namespace wrnn
