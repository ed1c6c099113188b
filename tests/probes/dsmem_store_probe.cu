// dsmem_store_probe.cu -- validates the cluster primitives the cluster form of wrnn_stream.cu relies on:
// st.shared::cluster.u16 / .f32 through mapa addresses into every CTA of a 4-CTA cluster, remote mbarrier.arrive
// (release.cluster) and mbarrier.try_wait.parity.acquire.cluster, with dynamic shared memory at a large offset.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t d; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(r)); return d; }
__device__ __forceinline__ uint32_t crank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
constexpr int CL = 4, NT = 128, SMEM = 200 * 1024, OFF = 100 * 1024;
__global__ void __launch_bounds__(NT, 1) k(int* result) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint16_t* img = reinterpret_cast<uint16_t*>(smem + OFF);           // [CL][NT] u16
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF + 4096);
  const uint32_t rank = crank();
  for (int i = threadIdx.x; i < CL * NT; i += NT) img[i] = 0;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "n"(CL * NT)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  // every thread writes (rank, tid) into slot [rank][tid] of ALL CTAs, then arrives on every CTA's barrier
  const uint16_t v = (uint16_t)(0x1000 * (rank + 1) + threadIdx.x);
  const uint32_t a = smem_u32(img + rank * NT + threadIdx.x);
  for (int r = 0; r < CL; ++r) asm volatile("st.shared::cluster.u16 [%0], %1;" :: "r"(mapa(a, r)), "h"(v) : "memory");
  asm volatile("fence.proxy.async;" ::: "memory");
  for (int r = 0; r < CL; ++r) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(mapa(smem_u32(bar), r)) : "memory");
  uint32_t ok = 0; long long t0 = clock64();
  while (!ok && clock64() - t0 < (1ll << 28))
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
  int bad = ok ? 0 : 1000000;
  for (int r = 0; r < CL; ++r) if (img[r * NT + threadIdx.x] != (uint16_t)(0x1000 * (r + 1) + threadIdx.x)) ++bad;
  if (bad) atomicAdd(result, bad);
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  int* d; CK(cudaMalloc(&d, 4)); CK(cudaMemset(d, 0, 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  cudaLaunchConfig_t lc{}; lc.gridDim = dim3(CL * 8); lc.blockDim = dim3(NT); lc.dynamicSmemBytes = SMEM;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  void* args[] = {&d};
  CK(cudaLaunchKernelExC(&lc, (const void*)k, args));
  CK(cudaDeviceSynchronize());
  int h = -1; CK(cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost));
  printf("dsmem store probe: %s (mismatches %d)\n", h == 0 ? "OK" : "FAIL", h);
  return h == 0 ? 0 : 1;
}
