// dsmem_probe.cu -- round-2 groundwork (NOT part of the product, NOT yet run: written after the round-1 GPU budget
// was spent).  Question it answers: what does an exchange cost when it stays inside a thread-block cluster?
//
// Plan it supports: keep h1', h2', y1 on the global L2 exchange, but give every cluster of 8
// CTAs a full copy of fc2 (64 rows per CTA) and fc3, so the y2 vector is exchanged over distributed shared memory
// instead of L2 (one of the four ~4000-cycle exchanges of a step becomes a cluster-local one).
//
// Measured here, for clusters of 2 / 4 / 8 CTAs out of 128 co-resident CTAs with ~200 KB of dynamic shared memory
// each (the engine's footprint):
//   * cudaOccupancyMaxActiveClusters (are 128 / CL clusters co-resident at all?)
//   * cycles per round of: every CTA bulk-copies its 1 KB-per-row-group slice (3 row groups = 3 KB, the y2 slice of
//     19 folds) into the same offset of all CL shared memories (cp.async.bulk.shared::cluster.shared::cta with
//     complete_tx on the destination's mbarrier), then waits until its own buffer has received all CL slices.
// Double-buffered by round parity: a CTA cannot finish round r before every peer has sent round r, so peers are at
// most one round apart.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int NT = 256, ROUNDS = 2000, SLICE = 3 * 1024, P = 128;
constexpr int SMEM = 200 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > (1ll << 28)) return false;
  }
}

template <int CL>
__global__ void __launch_bounds__(NT, 1) dsmem_exchange(long long* cycles, int* failed) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // [0, 2*CL*SLICE): receive buffers (2 parities x CL slices); then the local slice; then 2 mbarriers
  unsigned char* recv = smem;
  unsigned char* mine = smem + 2 * CL * SLICE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * CL * SLICE + SLICE);
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < SLICE / 4; i += NT) reinterpret_cast<uint32_t*>(mine)[i] = blockIdx.x * 1000 + i;
  __syncthreads();
  cluster_sync();                                        // every peer's barriers are initialised
  long long t_start = 0;
  bool ok = true;
  for (int r = 0; r < ROUNDS; ++r) {
    const int par = r & 1;
    if (r == 8 && threadIdx.x == 0) t_start = clock64();
    if (threadIdx.x == 0) {
      const uint32_t bar = smem_u32(&bars[par]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(CL * SLICE) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes of `mine` -> async-proxy reads
      const uint32_t src = smem_u32(mine);
      const uint32_t dst_local = smem_u32(recv + (size_t)par * CL * SLICE + rank * SLICE);
#pragma unroll
      for (int c = 0; c < CL; ++c) {
        const uint32_t dst = map_to_cta(dst_local, c), rbar = map_to_cta(bar, c);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "r"(src), "r"(SLICE), "r"(rbar) : "memory");
      }
      ok = ok && mbar_wait(bar, (uint32_t)((r >> 1) & 1));
    }
    __syncthreads();                                     // the fold warps would read the assembled vector here
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[0] = (clock64() - t_start) / (ROUNDS - 8); if (!ok) *failed = 1; }
  // spot check of the last round's data: slice c came from CTA (cluster base + c)
  if (threadIdx.x < CL) {
    const uint32_t got = reinterpret_cast<uint32_t*>(recv + (size_t)((ROUNDS - 1) & 1) * CL * SLICE + threadIdx.x * SLICE)[5];
    if (got != (blockIdx.x - rank + threadIdx.x) * 1000 + 5) *failed = 2;
  }
  cluster_sync();                                        // nobody leaves while a peer may still write into it
}

template <int CL>
void run(long long* d_cycles, int* d_failed) {
  CK(cudaFuncSetAttribute(dsmem_exchange<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  if (CL > 8) CK(cudaFuncSetAttribute(dsmem_exchange<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(P); lc.blockDim = dim3(NT); lc.dynamicSmemBytes = SMEM;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  lc.attrs = at; lc.numAttrs = 1;
  int n_clusters = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n_clusters, dsmem_exchange<CL>, &lc);
  printf("cluster size %2d: max active clusters %d (need %d)%s\n", CL, n_clusters, P / CL, e == cudaSuccess ? "" : " [query failed]");
  fflush(stdout);
  if (e != cudaSuccess) { cudaGetLastError(); return; }
  if (n_clusters < P / CL) return;
  lc.numAttrs = 2;
  CK(cudaMemset(d_cycles, 0, 8)); CK(cudaMemset(d_failed, 0, 4));
  void* args[] = {&d_cycles, &d_failed};
  CK(cudaLaunchKernelExC(&lc, (const void*)dsmem_exchange<CL>, args));
  CK(cudaDeviceSynchronize());
  long long cyc = 0; int failed = 0;
  CK(cudaMemcpy(&cyc, d_cycles, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&failed, d_failed, 4, cudaMemcpyDeviceToHost));
  printf("cluster size %2d: %lld cycles per DSMEM exchange round (%d B per CTA to %d CTAs)%s\n", CL, cyc, SLICE, CL,
         failed == 0 ? "" : failed == 1 ? "  [TIMEOUT]" : "  [DATA MISMATCH]");
  fflush(stdout);
}

int main() {
  long long* d_cycles; int* d_failed;
  CK(cudaMalloc(&d_cycles, 8)); CK(cudaMalloc(&d_failed, 4));
  run<2>(d_cycles, d_failed);
  run<4>(d_cycles, d_failed);
  run<8>(d_cycles, d_failed);
  run<16>(d_cycles, d_failed);
  return 0;
}
