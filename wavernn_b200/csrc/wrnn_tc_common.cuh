// wrnn_tc_common.cuh -- PTX wrappers and epilogue math shared by the tcgen05 engines (wrnn_tc.cu, wrnn_tcc.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "wrnn_device.cuh"

namespace wrnn {
namespace tc {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // SM100 shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
  // layout_type [61,64) = 0 (SWIZZLE_NONE / interleaved 8x16B core matrices), K-major
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t umma_idesc(int M, int N, int fmt) {
  // c_format F32 [4,6)=1 | a_format [7,10) | b_format [10,13) | K-major A,B | N>>3 [17,23) | M>>4 [24,29)
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Executed by a WHOLE converged warp with warp-uniform operands; elect.sync predicates the instruction
// onto one lane.  (Issuing from a `tid == k` branch makes ptxas wrap every UTCHMMA in an ELECT/BRA.U.ANY
// loop: ~70 cycles per instruction instead of ~38, measured with tests/probes/umma_probe.cu.)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(mbar) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (whole warp, one lane elected)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
               "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}\n"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
  v[0] = __uint_as_float(a); v[1] = __uint_as_float(b); v[2] = __uint_as_float(c); v[3] = __uint_as_float(d);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tmem_ldN(uint32_t taddr, float* v) {
  if constexpr (N == 4) tmem_ld4(taddr, v);
  else if constexpr (N == 8) tmem_ld8(taddr, v);
  else if constexpr (N == 16) tmem_ld16(taddr, v);
  else tmem_ld32(taddr, v);
}
// out[0 .. NC) = sum over the NW partial accumulators (column stride `wstride`) of NC columns at taddr.
// One wide tcgen05.ld per partial (NL = NC rounded up to 4/8/16/32 columns; the surplus columns are ignored):
// every tcgen05.ld instruction costs tens of cycles of issue on the critical path, so fewer and wider wins.
template <int NC, int NW>
__device__ __forceinline__ void tmem_ld_sum(uint32_t taddr, int wstride, float* out) {
  constexpr int NL = NC <= 4 ? 4 : NC <= 8 ? 8 : NC <= 16 ? 16 : 32;
  float part[NW][NL];
#pragma unroll
  for (int w = 0; w < NW; ++w) tmem_ldN<NL>(taddr + w * wstride, part[w]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    float acc = part[0][i];
#pragma unroll
    for (int w = 1; w < NW; ++w) acc += part[w][i];
    out[i] = acc;
  }
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

// bounded waits: a protocol bug must end in an error code, never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* abort_flag) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 1023u) == 0) {                       // rarely: has another wait already given up?
      if (ld_relaxed_s32(abort_flag) != 0) return;
      if (clock64() - t0 > kWatchdogCycles) { atomicExch(abort_flag, 2); return; }
    }
  }
}
__device__ __forceinline__ void counter_wait(const unsigned* ctr, unsigned target, int* abort_flag) {
  if (!counter_behind(ld_acquire_u32(ctr), target)) return;          // wrap-safe, see grid_barrier
  const long long t0 = clock64();
  unsigned spins = 0;
  while (counter_behind(ld_acquire_u32(ctr), target)) {
    // watchdog / abort check (a second L2 load) only every 64th poll.  Measured: no change of the step (12.10 us either
    // way) -- the exchange is bound by when the last of the 128 producers arrives, not by the polling period.
    if ((++spins & 63u) == 0 && (clock64() - t0 > kWatchdogCycles || ld_relaxed_s32(abort_flag) != 0)) { atomicExch(abort_flag, 1); return; }
  }
}

// Streamed draws (wrnn_job::uniforms_ready): returns the number of valid rows once it exceeds `need - 1`
// (0xffffffff if the job is being abandoned, so the caller stops asking).
__device__ __forceinline__ unsigned rows_wait(const unsigned* ready, unsigned need, int* abort_flag) {
  unsigned v = ld_acquire_u32(ready);
  if (v >= need) return v;
  const long long t0 = clock64();
  while ((v = ld_acquire_u32(ready)) < need) {
    if (ld_relaxed_s32(abort_flag) != 0) return 0xffffffffu;
    if (clock64() - t0 > kWatchdogCycles) { atomicExch(abort_flag, 1); return 0xffffffffu; }
    __nanosleep(200);
  }
  return v;
}

template <int FMT> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<0>(float a, float b) {      // fp16, RNE, saturating
  uint32_t r;
  asm("{\n\t.reg .b16 lo, hi;\n\tcvt.rn.satfinite.f16.f32 lo, %1;\n\tcvt.rn.satfinite.f16.f32 hi, %2;\n\tmov.b32 %0, {lo, hi};\n\t}\n"
      : "=r"(r) : "f"(a), "f"(b));
  return r;
}
template <> __device__ __forceinline__ uint32_t pack2<1>(float a, float b) {      // bf16, RNE
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// Epilogue transcendentals on the SFU (ex2/rcp/lg2 approx, rel. error ~1e-6 -- two orders below the
// fp16 operand rounding of the contractions).  The strict SIMT engine keeps the libm versions.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ float gru_unit_fast(float gi_r, float gi_z, float gi_n, float gh_r, float gh_z, float gh_n, float h) {
  const float r = fast_sigmoid(gi_r + gh_r);
  const float z = fast_sigmoid(gi_z + gh_z);
  const float n = fast_tanh(gi_n + r * gh_n);
  return (1.0f - z) * n + z * h;
}
// utils/distribution.py:99-121 on SFU logs/exp
__device__ __forceinline__ float mol_sample_fast(const float* lg, const float* u) {
  int best = 0;
  float bestv = -INFINITY;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float g = lg[k] - __logf(-__logf(u[k]));
    if (g > bestv) { bestv = g; best = k; }
  }
  float mean = lg[10], ls = lg[20];
#pragma unroll
  for (int k = 1; k < 10; ++k) { if (best == k) { mean = lg[10 + k]; ls = lg[20 + k]; } }
  ls = fmaxf(ls, -32.23619130191664f);
  const float x = mean + __expf(ls) * (__logf(u[10]) - __logf(1.0f - u[10]));
  return fminf(fmaxf(x, -1.0f), 1.0f);
}


// K-major no-swizzle operand image (host side): element (r, k) at (r/8)*(K/8)*64 + (k/8)*64 + (r%8)*8 + k%8
inline size_t img_index(int r, int k, int K) { return (size_t)(r / 8) * (K / 8) * 64 + (size_t)(k / 8) * 64 + (r % 8) * 8 + (k % 8); }

}  // namespace tc
}  // namespace wrnn
