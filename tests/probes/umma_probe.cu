// umma_probe.cu -- standalone bring-up probe for the tcgen05 building blocks used by
// wavernn_b200/csrc/wrnn_tc.cu: no-swizzle K-major shared-memory descriptors, M=64 / M=128
// accumulator layouts in TMEM, small N, "don't-care" row groups aliasing other data, and
// the issue->commit->ld latency of a K=512 chain.  Build: make -C tests/probes.  GPU only.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;       // version = 1 (Blackwell)
  return d;                     // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// whole converged warp executes; elect.sync predicates the instruction onto one lane
__device__ __forceinline__ void mma_f16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

struct Params { const __half* A; const __half* B; float* D; long long* cycles; int M, N, K, a_rows_real, mode; };

// A image: canonical K-major no-swizzle, LBO = 128 B (adjacent k8 chunks contiguous), SBO = K/8*128 B
__global__ void __launch_bounds__(128, 1) umma_probe(Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int K = p.K, M = p.M, N = p.N;
  const uint32_t a_bytes = (uint32_t)p.a_rows_real * K * 2, sbo = (uint32_t)(K / 8) * 128;
  unsigned char* sA = smem;
  unsigned char* sB = smem + 65536 * 2;     // leave room: A row groups beyond a_rows_real alias what follows
  const int tid = threadIdx.x, warp = tid >> 5;
  // fill the aliased region with a NaN pattern to prove those rows cannot leak into real rows
  for (int i = tid; i < (65536 * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x7e007e00u;
  __syncthreads();
  for (int i = tid; i < (int)a_bytes / 16; i += 128) reinterpret_cast<int4*>(sA)[i] = reinterpret_cast<const int4*>(p.A)[i];
  for (int i = tid; i < N * K * 2 / 16; i += 128) reinterpret_cast<int4*>(sB)[i] = reinterpret_cast<const int4*>(p.B)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // idesc: c=F32(1)<<4, a=b=F16(0), K-major, N>>3 at bit 17, M>>4 at bit 24
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  long long t0 = 0, t1 = 0, t2 = 0;
  for (int rep = 0; rep < 3; ++rep) {
    __syncthreads();
    t0 = clock64();
    if (p.mode == 0) {
      if (tid == 0) {
        for (int k = 0; k < K / 16; ++k) {
          const uint64_t ad = make_desc(smem_u32(sA) + k * 256, 128, sbo);
          const uint64_t bd = make_desc(smem_u32(sB) + k * 256, 128, sbo);
          mma_f16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
      }
    } else if (warp == 0) {
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll 4
      for (int k = 0; k < K / 16; ++k)
        mma_f16_elect(tmem, make_desc(a0 + k * 256, 128, sbo), make_desc(b0 + k * 256, 128, sbo), idesc, k > 0 ? 1u : 0u);
      asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(smem_u32(&mbar)) : "memory");
    }
    t1 = clock64();
    mbar_wait(smem_u32(&mbar), rep & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t2 = clock64();
  }
  // raw dump: lane x column
  uint32_t r[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const long long t3 = clock64();
  for (int c = 0; c < 32; ++c) p.D[tid * 32 + c] = __uint_as_float(r[c]);
  if (tid == 0) { p.cycles[0] = t1 - t0; p.cycles[1] = t2 - t0; p.cycles[2] = t3 - t2; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem));
}

static int run_case(int M, int N, int K, int a_rows_real, int mode) {
  std::vector<__half> A((size_t)a_rows_real * K), B((size_t)N * K);
  std::vector<float> Af((size_t)a_rows_real * K), Bf((size_t)N * K);
  srand(1234 + M + N);
  auto img = [&](int r, int k) { return (size_t)(r / 8) * (K / 8) * 64 + (size_t)(k / 8) * 64 + (r % 8) * 8 + (k % 8); };
  for (int r = 0; r < a_rows_real; ++r) for (int k = 0; k < K; ++k) { float v = (float)((rand() % 9) - 4) * 0.25f; Af[(size_t)r * K + k] = v; A[img(r, k)] = __float2half(v); }
  for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) { float v = (float)((rand() % 7) - 3) * 0.5f; Bf[(size_t)r * K + k] = v; B[img(r, k)] = __float2half(v); }
  __half *dA, *dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * 32 * 4)); CK(cudaMalloc(&dC, 64));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * 32 * 4));
  Params p{dA, dB, dD, dC, M, N, K, a_rows_real, mode};
  const int smem = 65536 * 2 + N * K * 2 + 1024;
  CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe<<<1, 128, smem>>>(p);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 32); long long cyc[3];
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cyc, dC, 24, cudaMemcpyDeviceToHost));
  // hypothesis: M=128 -> lane = row; M=64 -> lane = (row/16)*32 + row%16
  double maxerr = 0; int bad = 0;
  for (int r = 0; r < a_rows_real && r < M; ++r) {
    const int lane = (M == 128) ? r : (r / 16) * 32 + (r % 16);
    for (int c = 0; c < N; ++c) {
      double ref = 0; for (int k = 0; k < K; ++k) ref += (double)Af[(size_t)r * K + k] * Bf[(size_t)c * K + k];
      const double e = fabs(ref - D[lane * 32 + c]);
      if (!(e <= 1e-3)) { if (bad < 4) printf("   mismatch row %d col %d: got %f want %f (lane %d)\n", r, c, D[lane * 32 + c], ref, lane); ++bad; }
      if (e > maxerr) maxerr = e;
    }
  }
  printf("%s M=%3d N=%2d K=%3d real_rows=%3d : %s  maxerr=%.3g  issue=%lld cyc  issue->done=%lld cyc  ld=%lld cyc\n", mode ? "[elect-warp]" : "[tid==0   ]", M, N, K,
         a_rows_real, bad ? "MISMATCH" : "ok", maxerr, cyc[0], cyc[1], cyc[2]);
  if (bad) {   // help decoding: where did row 0 / row 17 land?
    for (int lane = 0; lane < 128; lane += 1) { bool nz = false; for (int c = 0; c < N; ++c) nz |= (D[lane * 32 + c] == D[lane * 32 + c]) && D[lane * 32 + c] != 0.f; if (nz && lane % 8 == 0) printf("   lane %d has data: %f %f\n", lane, D[lane * 32], D[lane * 32 + 1]); }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad != 0;
}

int main() {
  int fails = 0;
  const int cases[][4] = {{128, 32, 512, 128}, {64, 32, 512, 64}, {64, 16, 512, 64}, {64, 8, 512, 64}, {128, 16, 512, 128},
                          {64, 32, 512, 24}, {128, 32, 512, 24}, {64, 32, 208, 24}, {64, 32, 512, 56}};
  for (int mode = 0; mode < 2; ++mode) for (auto& c : cases) fails += run_case(c[0], c[1], c[2], c[3], mode);
  printf(fails ? "UMMA PROBE: %d case(s) FAILED\n" : "UMMA PROBE: all cases ok\n", fails);
  return fails ? 1 : 0;
}
