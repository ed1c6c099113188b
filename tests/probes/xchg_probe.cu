// xchg_probe.cu -- measures the inter-SM exchange primitives the persistent kernel is built on
// (128 co-resident CTAs, L2-resident buffers):
//   A  counter barrier: red.release.gpu + ld.acquire.gpu polling by one thread per CTA
//   B  publish (19 x 8 B per CTA) + counter barrier + gather of the 24 KB image by all threads
//   C  like B but the gather is one cp.async.bulk (TMA 1-D) issued by one thread
//   D  flag-in-data: every 16 B chunk carries the step tag; consumers poll the data itself
// Prints average cycles per round seen by CTA 0.  GPU only.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int NT = 256, ROUNDS = 2000, IMG = 24576;   // 3 row groups x 8 KB

__device__ int P = 128;
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_rel(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_rel(ctr);
    long long t0 = clock64();
    while (ld_acq(ctr) < target) if (clock64() - t0 > (1ll << 28)) break;
  }
  __syncthreads();
  return true;
}

// variant: relaxed polling by 4 staggered lanes (one per warp 0..3), one acquire fence after success
__device__ __forceinline__ void barrier_staggered(unsigned* ctr, unsigned target) {
  __shared__ volatile int s_go;
  if (threadIdx.x == 0) s_go = 0;
  __syncthreads();
  if (threadIdx.x == 0) red_rel(ctr);
  if ((threadIdx.x & 31) == 0 && threadIdx.x < 128) {
    const int k = threadIdx.x >> 5;
    long long t0 = clock64();
    while (clock64() - t0 < 170 * k) {}                      // stagger the poll phases by ~170 cycles
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target && s_go == 0 && clock64() - t0 < (1ll << 28));
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    s_go = 1;
  }
  __syncthreads();
}

// variant: no release/acquire at all (red.relaxed + ld.relaxed): the bare arrival cost
__device__ __forceinline__ void barrier_relaxed(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    long long t0 = clock64();
    unsigned v;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target && clock64() - t0 < (1ll << 28));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT, 1) probe(unsigned* ctr, unsigned char* img, long long* result, int mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  const int cta = blockIdx.x, tid = threadIdx.x;
  unsigned nb = 0;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  barrier(ctr, ++nb * P);
  const long long t0 = clock64();
  unsigned acc = 0;
  int failed = 0;
  __shared__ int dbg_once_s; int* dbg_once = &dbg_once_s; if (tid == 0) dbg_once_s = 0;
  for (int r = 0; r < ROUNDS; ++r) {
    unsigned char* buf = img + (size_t)(r & 1) * IMG * 2;
    if (mode == 0) {
      barrier(ctr, ++nb * P);
    } else if (mode == 4) {
      barrier_staggered(ctr, ++nb * P);
    } else if (mode == 5) {
      barrier_relaxed(ctr, ++nb * P);
    } else if (mode == 6) {
      // flag-in-data, polite polling: one 16 B chunk per producer CTA is polled with back-off by ONE warp,
      // everything else is fetched once the probes are in
      const unsigned tag = (unsigned)r + 1;
      if (tid < 12) {
        int4 v = make_int4(cta, tid, r, (int)tag);
        asm volatile("st.global.cg.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(reinterpret_cast<int4*>(buf) + cta * 12 + tid), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      if (tid < 128) {                       // thread k watches the LAST chunk of producer k
        int4 v; long long tw = clock64();
        for (;;) {
          asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(reinterpret_cast<const int4*>(buf) + tid * 12 + 11) : "memory");
          if ((unsigned)v.w == tag || clock64() - tw > (1ll << 24)) break;
          __nanosleep(40);
        }
        if ((unsigned)v.w != tag) failed = 1;
      }
      if (__syncthreads_or(failed)) { if (cta == 0 && tid == 0) result[24] = r + 1; break; }
      for (int i = tid; i < IMG / 16; i += NT) {
        int4 v;
        long long tw = clock64();
        do {
          asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(reinterpret_cast<const int4*>(buf) + i) : "memory");
        } while ((unsigned)v.w != tag && clock64() - tw < (1ll << 24));
        if ((unsigned)v.w != tag) failed = 1;
        reinterpret_cast<int4*>(smem)[i] = v;
      }
      if (__syncthreads_or(failed)) { if (cta == 0 && tid == 0) result[24] = r + 1; break; }
      acc += reinterpret_cast<unsigned*>(smem)[(tid * 4) % (IMG / 4)];
    } else if (mode == 1 || mode == 2) {
      // publish: 19 folds x 8 bytes (4 halfs of this CTA's units)
      if (tid < 19) *reinterpret_cast<uint2*>(buf + (tid / 8) * 8192 + (cta / 2) * 128 + (tid % 8) * 16 + (cta % 2) * 8) = make_uint2(r, cta);
      barrier(ctr, ++nb * P);
      if (mode == 1) {
        for (int i = tid; i < IMG / 16; i += NT) reinterpret_cast<int4*>(smem)[i] = __ldcg(reinterpret_cast<const int4*>(buf) + i);
        __syncthreads();
      } else {
        if (tid == 0) {
          asm volatile("fence.proxy.async.global;" ::: "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(IMG) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       :: "r"(smem_u32(smem)), "l"(buf), "r"(IMG), "r"(smem_u32(&mbar)) : "memory");
        }
        {
          long long tw = clock64(); unsigned ok = 0;
          while (!ok && clock64() - tw < (1ll << 28))
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(r & 1) : "memory");
          if (!ok) { if (tid == 0 && cta == 0) result[4] = r + 1; break; }
        }
      }
      acc += reinterpret_cast<unsigned*>(smem)[(tid * 4) % (IMG / 4)];
    } else {
      // flag-in-data: each 16 B chunk = {payload x3, tag}; every CTA owns chunks cta*12 .. cta*12+11 (12*128 = 1536 = IMG/16)
      const unsigned tag = (unsigned)r + 1;
      if (tid < 12) {
        int4 v = make_int4(cta, tid, r, (int)tag);
        asm volatile("st.global.cg.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(reinterpret_cast<int4*>(buf) + cta * 12 + tid), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      for (int i = tid; i < IMG / 16; i += NT) {
        int4 v;
        long long tw = clock64();
        do {
          asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(reinterpret_cast<const int4*>(buf) + i) : "memory");
        } while ((unsigned)v.w != tag && clock64() - tw < (1ll << 24));
        if ((unsigned)v.w != tag && cta == 0 && atomicAdd(dbg_once, 1) == 0) {
          result[16] = i; result[17] = v.w; result[18] = tag; result[19] = r; result[20] = v.x; result[21] = v.y; result[22] = v.z;
        }
        if ((unsigned)v.w != tag) { failed = 1; }
        reinterpret_cast<int4*>(smem)[i] = v;
      }
      if (__syncthreads_or(failed)) break;
      acc += reinterpret_cast<unsigned*>(smem)[(tid * 4) % (IMG / 4)];
    }
  }
  const long long t1 = clock64();
  if (cta == 0 && tid == 0) { result[mode] = (t1 - t0) / ROUNDS; result[8 + mode] = acc; }
}

int main(int argc, char** argv) {
  int hostP = argc > 1 ? atoi(argv[1]) : 128;
  CK(cudaMemcpyToSymbol(P, &hostP, sizeof(int)));
  printf("== %d CTAs ==\n", hostP);
  unsigned* ctr; unsigned char* img; long long* res;
  CK(cudaMalloc(&ctr, 64)); CK(cudaMalloc(&img, IMG * 4)); CK(cudaMalloc(&res, 256));
  CK(cudaMemset(res, 0, 256));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG + 1024));
  const char* names[] = {"A counter barrier only", "B publish + barrier + LDG gather 24KB", "C publish + barrier + TMA bulk gather 24KB",
                         "D flag-in-data gather 24KB (no barrier)", "E counter barrier, 4 staggered relaxed pollers",
                         "F counter barrier, red.relaxed + ld.relaxed (no fences)", "G flag-in-data, polite probe + LDG gather 24KB"};
  for (int mode = 0; mode < 7; ++mode) {
    if (hostP != 128 && (mode == 3 || mode == 6)) continue;
    CK(cudaMemset(ctr, 0, 64)); CK(cudaMemset(img, 0, IMG * 4));
    void* args[] = {&ctr, &img, &res, &mode};
    CK(cudaLaunchCooperativeKernel((const void*)probe, dim3(hostP), dim3(NT), args, IMG + 1024, 0));
    CK(cudaDeviceSynchronize());
    long long h[32]; CK(cudaMemcpy(h, res, 256, cudaMemcpyDeviceToHost));
    if (mode == 6 && h[24]) printf("   mode G FAILED at round %lld\n", h[24] - 1);
    if (mode == 3 && h[18]) printf("   mode D first timeout: chunk %lld saw tag %lld (x=%lld y=%lld z=%lld) wanted %lld at round %lld\n", h[16], h[17], h[20], h[21], h[22], h[18], h[19]);
    printf("%-45s : %6lld cycles/round%s\n", names[mode], h[mode], (mode == 2 && h[4]) ? "  (TMA wait TIMED OUT)" : "");
    fflush(stdout);
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("(SM clock attr %d kHz)\n", clk);
  return 0;
}
