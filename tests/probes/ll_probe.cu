// ll_probe.cu -- round-2 exchange probes (NOT part of the product).
//
// (1) "LL" push exchange: instead of  st data -> release fence -> red counter -> consumers poll counter -> TMA gather
//     (xchg_probe variant C, ~3500-3900 cycles), every producer CTA stores its slice as 16-byte granules
//     {8 B payload, tag, tag} into a mailbox per consumer GROUP (SHARE consumers read the same mailbox: 1 = private,
//     8 = one per cluster-sized group, 128 = one global mailbox), and consumers poll the granules themselves:
//     data and flag travel together, no fence, no counter, no second round trip.
//     Reported: cycles per round seen by CTA 0 (2000 rounds) for NF folds.
// (2) L2 -> SM ingress: n CTAs each stream an L2-resident 8 MB buffer through a 4 x 32 KB shared-memory ring with
//     cp.async.bulk (TMA 1-D); reported as bytes / cycle / SM and aggregate B/cycle.  Bounds any design that streams
//     weights instead of keeping them resident.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int NT = 256, ROUNDS = 2000, P = 128;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mailbox of group g: [parity][producer 128][fold nf] granules of 16 B
template <int POLL_THREADS>
__global__ void __launch_bounds__(NT, 1) ll_exchange(int4* mail, long long* result, int nf, int share, int slot) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int ngroups = P / share, mygroup = cta / share;
  const size_t per_par = (size_t)P * nf, per_group = 2 * per_par;
  cooperative_groups::this_grid().sync();
  const long long t0 = clock64();
  int failed = 0;
  unsigned acc = 0;
  for (int r = 0; r < ROUNDS; ++r) {
    const int par = r & 1;
    const int tag = r + 1;
    // ---- publish: ngroups copies of this CTA's nf granules
    for (int i = tid; i < ngroups * nf; i += NT) {
      const int g = i / nf, f = i - g * nf;
      int4* dst = mail + (size_t)g * per_group + (size_t)par * per_par + (size_t)cta * nf + f;
      asm volatile("st.global.cg.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"(cta), "r"(f), "r"(tag), "r"(tag) : "memory");
    }
    // ---- gather: poll my group's mailbox (all P*nf granules), batches of 4 outstanding loads per thread
    const int4* src = mail + (size_t)mygroup * per_group + (size_t)par * per_par;
    const int total = P * nf;
    if (tid < POLL_THREADS) {
      for (int i0 = tid; i0 < total; i0 += POLL_THREADS * 4) {
        int4 v[4]; bool need[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) need[j] = (i0 + j * POLL_THREADS) < total;
        const long long tw = clock64();
        for (;;) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (need[j]) asm volatile("ld.relaxed.gpu.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(src + i0 + j * POLL_THREADS) : "memory");
          bool any = false;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (need[j]) {
              if (v[j].z == tag && v[j].w == tag) { reinterpret_cast<int2*>(smem)[i0 + j * POLL_THREADS] = make_int2(v[j].x, v[j].y); need[j] = false; }
              else any = true;
            }
          if (!any) break;
          if (clock64() - tw > (1ll << 24)) { failed = 1; break; }
        }
      }
    }
    if (__syncthreads_or(failed)) { if (cta == 0 && tid == 0) result[32 + slot] = r + 1; break; }
    acc += reinterpret_cast<unsigned*>(smem)[(tid * 2) % (total * 2)];
  }
  const long long t1 = clock64();
  if (cta == 0 && tid == 0) { result[slot] = (t1 - t0) / ROUNDS; result[16 + slot] = acc; }
}

// ---- (2) L2 -> SM ingress ------------------------------------------------------------------------------------
constexpr int CHUNK = 32768, SLOTS = 4, BUF = 8 << 20, REPS = 6;
__global__ void __launch_bounds__(128, 1) ingress(const unsigned char* buf, long long* result, int slot) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[SLOTS];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < SLOTS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nchunks = BUF / CHUNK * REPS;
  long long t0 = 0;
  unsigned acc = 0;
  // different CTAs start at different offsets so they do not all hit the same L2 lines at the same time
  const int start = (blockIdx.x * 37) % (BUF / CHUNK);
  auto issue = [&](int c) {
    const int s = c % SLOTS;
    const size_t off = (size_t)((start + c) % (BUF / CHUNK)) * CHUNK;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bars[s])), "r"(CHUNK) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem + s * CHUNK)), "l"(buf + off), "r"(CHUNK), "r"(smem_u32(&bars[s])) : "memory");
  };
  if (tid == 0) {
    // warm pass (brings the buffer into L2) is the first BUF/CHUNK chunks; timing starts after it
    for (int c = 0; c < SLOTS; ++c) issue(c);
    for (int c = 0; c < nchunks; ++c) {
      if (c == BUF / CHUNK) t0 = clock64();
      const int s = c % SLOTS; const uint32_t ph = (c / SLOTS) & 1;
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                               : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(ph) : "memory");
      acc += smem[s * CHUNK + (c & 1023)];
      if (c + SLOTS < nchunks) issue(c + SLOTS);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) { result[slot] = t1 - t0; result[16 + slot] = acc; }
  }
}

int main() {
  int4* mail; long long* res; unsigned char* buf;
  const size_t mail_bytes = (size_t)P * 2 * P * 64 * 16;       // private mailboxes, up to 64 folds
  CK(cudaMalloc(&mail, mail_bytes)); CK(cudaMalloc(&res, 1024)); CK(cudaMalloc(&buf, BUF));
  CK(cudaMemset(buf, 1, BUF));
  CK(cudaFuncSetAttribute(ll_exchange<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 64 * 8 + 1024));
  CK(cudaFuncSetAttribute(ll_exchange<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 64 * 8 + 1024));
  int slot = 0;
  const int nfs[] = {19, 24, 8, 64};
  const int shares[] = {1, 8, 32, 128};
  for (int pt = 0; pt < 2; ++pt)
    for (int nf : nfs)
      for (int share : shares) {
        if (pt == 1 && nf != 19) continue;
        CK(cudaMemset(mail, 0, mail_bytes)); CK(cudaMemset(res, 0, 1024));
        slot = 0;
        void* args[] = {&mail, &res, &nf, &share, &slot};
        const void* fn = pt == 0 ? (const void*)ll_exchange<256> : (const void*)ll_exchange<128>;
        CK(cudaLaunchCooperativeKernel(fn, dim3(P), dim3(NT), args, 128 * 64 * 8 + 1024, 0));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("LL nf=%d share=%d: %s\n", nf, share, cudaGetErrorString(e)); return 2; }
        long long h[128]; CK(cudaMemcpy(h, res, 1024, cudaMemcpyDeviceToHost));
        printf("LL push  pollers=%3d  nf=%2d  share=%3d : %6lld cycles/round%s\n", pt == 0 ? 256 : 128, nf, share, h[0], h[32] ? "  [TIMEOUT]" : "");
        fflush(stdout);
      }
  CK(cudaFuncSetAttribute(ingress, cudaFuncAttributeMaxDynamicSharedMemorySize, SLOTS * CHUNK));
  const int grids[] = {1, 8, 32, 64, 128, 148};
  for (int g : grids) {
    CK(cudaMemset(res, 0, 1024));
    slot = 0;
    ingress<<<g, 128, SLOTS * CHUNK>>>(buf, res, slot);
    CK(cudaDeviceSynchronize());
    long long h[128]; CK(cudaMemcpy(h, res, 1024, cudaMemcpyDeviceToHost));
    const double bytes = (double)BUF * (REPS - 1);
    printf("ingress  %3d CTAs : %8lld cycles for %.0f MB per CTA -> %.1f B/cycle/SM, %.0f B/cycle aggregate\n", g, h[0], bytes / 1e6,
           bytes / (double)h[0], bytes * g / (double)h[0]);
    fflush(stdout);
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("(SM clock attr %d kHz)\n", clk);
  return 0;
}
