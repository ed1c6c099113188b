// wrnn_tcc.cu -- tcgen05 "cluster-tail" engine: the persistent generate() kernel of wrnn_tc.cu with the
// fc1 / fc2 / fc3 tail replicated per 16-CTA thread-block cluster, so that only TWO of the four per-sample
// activation exchanges cross the whole GPU through L2; the other two stay inside a cluster on distributed
// shared memory.
//
//   global (L2 image + release counter + TMA gather, as in wrnn_tc.cu):   h1', h2'   -- they feed the
//       GRU matrices, which are sharded over all 128 CTAs (CTA c owns hidden units [4c, 4c+4));
//   cluster-local (st.shared::cluster pushes + remote mbarrier arrives):   y1, y2     -- fc1 and fc2 rows are
//       split 16-way INSIDE each cluster (rank r owns rows [32r, 32r+32)), every cluster computes the whole
//       tail redundantly (8x), fc3 + sampling stays replicated in every CTA (bitwise identical samples).
//
// Why: an exchange through L2 costs ~4000-5000 cycles on B200 however it is sliced
// (profiles/r01_exchange_probes.md); a DSMEM push + mbarrier costs a fraction of that.
//
// Per-CTA shared memory (fp16/bf16 operand images, K-major SWIZZLE_NONE):
//   S1F1 = [W2x(12) ; W1h(12) ; F1x rows 32r..(32)]  N=56, K=512   consumes h1' (and rows 24.. again on h2')
//   S2   = [W2h(12) ; 0(4)]                          N=16, K=512   consumes h2'
//   F2   = F2x rows 32r.. | F2a | F2a                N=32, K=576   consumes [y1 ; a4(even step) ; a4(odd step)]
//   F3   = fc3                                       N=32, K=512   consumes y2
//   Q    = folded conditioning rows gi1(12) gi2(12) fc1(32)  N=56, K=208
// Tiles of <= 24 folds (three 8-row groups); larger jobs are served by wrnn_tc.cu.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "wrnn_device.cuh"
#include "wrnn_engine.h"
#include "wrnn_tc_common.cuh"

namespace wrnn {
namespace {
using namespace tc;

constexpr int P = 128;             // CTAs
constexpr int CL = 16;             // CTAs per cluster
constexpr int U = H / P;           // 4 GRU units per CTA
constexpr int TU = H / CL;         // 32 tail rows (fc1, fc2) per CTA
constexpr int NT = 256;
constexpr int MT = 64;             // MMA M (folds); only MAXB of them are real
constexpr int MAXB = 24;           // folds per launch (3 row groups)
constexpr int NG = MAXB / 8;       // row groups of the A images
constexpr int KW = 2;              // issuing warps per chain (K halves)

constexpr int KY = H + 2 * AUXD;   // 576: y1 | a4 (even steps) | a4 (odd steps)
constexpr int SBO_H = (H / 8) * 128;      // 8192
constexpr int SBO_Y = (KY / 8) * 128;     // 9216
constexpr int KQ = CDIM / 8;              // 26
constexpr int SBO_Q = KQ * 128;           // 3328

constexpr int N_B = 56, N_S2 = 16, N_F2 = 32, N_F3 = 32, N_Q = 56;
constexpr int QROWS = 56;          // gi1 12 | gi2 12 | fc1 32
// shared memory map
constexpr int OFF_A = 0;                                   // h1' / h2' (TMA) and y2 (DSMEM pushes)
constexpr int OFF_Y = OFF_A + NG * SBO_H;                  // y1 (DSMEM pushes) + a4 columns
constexpr int OFF_C = OFF_Y + NG * SBO_Y;                  // cond image
constexpr int OFF_WB = OFF_C + NG * SBO_Q;                 // S1F1
constexpr int OFF_WS2 = OFF_WB + (N_B / 8) * SBO_H;
constexpr int OFF_WF2 = OFF_WS2 + (N_S2 / 8) * SBO_H;
constexpr int OFF_WF3 = OFF_WF2 + (N_F2 / 8) * SBO_Y;
constexpr int OFF_WQ = OFF_WF3 + (N_F3 / 8) * SBO_H;
constexpr int OFF_VEC = OFF_WQ + (N_Q / 8) * SBO_Q;        // fp32: qk[56] vq[56] b1h[12] b2h[12] bf2[32] b3[32] -> 200, pad 256
constexpr int NVEC = 256;
constexpr int OFF_BAR = OFF_VEC + NVEC * 4;
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int WEIGHT_BYTES = OFF_BAR - OFF_WB;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
// vector offsets (floats)
constexpr int V_QK = 0, V_VQ = 56, V_B1H = 112, V_B2H = 124, V_BF2 = 136, V_B3 = 168;

// TMEM columns: KW partial accumulators per chain
constexpr int TC_B = 0;                       // 2 x 64 (56 used): gi2x 0-11 | gh1n 12-23 | fc1 24-55
constexpr int TC_S2 = 128;                    // 2 x 16: gh2n 0-11
constexpr int TC_F2 = 160;                    // 2 x 32
constexpr int TC_F3 = 224;                    // 2 x 32
constexpr int TC_Q0 = 288, TC_Q1 = 352;       // 64 each (56 used)
constexpr int TMEM_COLS = 512;

struct TccParams {
  const unsigned char* blob;
  const float* mels_up; const float* aux; long long L; long long seg_stride;
  int n_seg, steps, out_pitch, seg_first;
  const float* uniforms; unsigned long long seed, offset;
  float* out; const float* x_force; float* logits_out;
  unsigned char* xch;        // [2 vectors][2 parities][NG * SBO_H]
  unsigned* counters;        // [2]
  int* abort_flag;
  long long* prof;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// mbarrier wait with cluster-scope acquire (the data was written by remote CTAs)
__device__ __forceinline__ bool mbar_try_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int* abort_flag) {
  if (mbar_try_cluster(bar, parity)) return;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try_cluster(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      if (ld_relaxed_s32(abort_flag) != 0) return;
      if (clock64() - t0 > kWatchdogCycles) { atomicExch(abort_flag, 3); return; }
    }
  }
}

template <int FMT>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) wrnn_tcc_kernel(const TccParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const float* fv = reinterpret_cast<const float*>(smem + OFF_VEC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);
  const uint32_t bar_mma = smem_u32(&bars[0]), bar_q = smem_u32(&bars[1]), bar_g = smem_u32(&bars[2]),
                 bar_y1 = smem_u32(&bars[3]), bar_y2 = smem_u32(&bars[4]);

  const int cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();               // 0..15: which 32 tail rows this CTA owns
  const int B = p.n_seg, S = p.steps, u0 = cta * U;
  const int n_groups = (B + 7) / 8;
  const uint32_t img_bytes = (uint32_t)n_groups * SBO_H;
  const size_t xch_stride = (size_t)2 * img_bytes;

  // ---- one-time setup ---------------------------------------------------------------------------
  {
    const int4* src = reinterpret_cast<const int4*>(p.blob + (size_t)cta * WEIGHT_BYTES);
    int4* dst = reinterpret_cast<int4*>(smem + OFF_WB);
    for (int i = tid; i < WEIGHT_BYTES / 16; i += NT) dst[i] = src[i];
    int4* z = reinterpret_cast<int4*>(smem);
    for (int i = tid; i < OFF_WB / 16; i += NT) z[i] = make_int4(0, 0, 0, 0);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_mma), "n"(KW));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_q));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_g));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_y1), "r"(CL * B));   // every fold thread of every peer
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_y2), "r"(CL * B));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  proxy_fence_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // every peer's barriers and zeroed images exist before any push
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // =========================================================================================
    // fold warps (folds 0-15 in warp 0, 16-23 in warp 1; warps 2-3 carry no fold at MAXB = 24)
    // =========================================================================================
    const int fold = warp * 16 + lane;
    const bool owns_fold = lane < 16 && fold < B;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const size_t pub_off = (size_t)(fold >> 3) * SBO_H + (u0 >> 3) * 128 + (fold & 7) * 16 + (u0 & 7) * 2;
    const bool profiling = (p.prof != nullptr) && cta == 0 && tid == 0;
    long long tprof[5] = {0, 0, 0, 0, 0};
    auto publish_global = [&](int v, int par, const float* val) {
      uint2 w;
      w.x = pack2<FMT>(val[0], val[1]); w.y = pack2<FMT>(val[2], val[3]);
      if (owns_fold) *reinterpret_cast<uint2*>(p.xch + (size_t)v * xch_stride + (size_t)par * img_bytes + pub_off) = w;
      tc_fence_before();
      named_bar_sync(1, 128);
      if (tid == 0) red_release_add_u32(p.counters + v, 1u);
    };
    // push this fold's 32 tail values (rows 32*rank ..) into the same place of every cluster peer's image
    auto push_cluster = [&](uint32_t img_saddr, uint32_t sbo, uint32_t bar_saddr, const float* val) {
      tc_fence_before();
      if (owns_fold) {
        uint4 c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          c[j].x = pack2<FMT>(val[8 * j + 0], val[8 * j + 1]); c[j].y = pack2<FMT>(val[8 * j + 2], val[8 * j + 3]);
          c[j].z = pack2<FMT>(val[8 * j + 4], val[8 * j + 5]); c[j].w = pack2<FMT>(val[8 * j + 6], val[8 * j + 7]);
        }
        const uint32_t local = img_saddr + (uint32_t)(fold >> 3) * sbo + (uint32_t)(4 * rank) * 128 + (uint32_t)(fold & 7) * 16;
#pragma unroll 4
        for (int pr = 0; pr < CL; ++pr) {
          const uint32_t peer = (uint32_t)((rank + pr) & (CL - 1));     // start with myself, spread the rest
          const uint32_t dst = mapa(local, peer);
#pragma unroll
          for (int j = 0; j < 4; ++j) st_cluster_v4(dst + j * 128, c[j]);
          mbar_arrive_remote(mapa(bar_saddr, peer));                    // release.cluster: orders this thread's stores
        }
      }
    };
    float h1[U] = {0.f, 0.f, 0.f, 0.f}, h2[U] = {0.f, 0.f, 0.f, 0.f};
    float x = 0.f;
    unsigned n_mma = 0;

    for (int t = 0; t < S; ++t) {
      const int par = t & 1;
      const uint32_t tq = (par ? TC_Q1 : TC_Q0);
      long long tp0 = 0;
      if (profiling) tp0 = clock64();
      float ur[11];
#pragma unroll
      for (int i = 0; i < 11; ++i) ur[i] = 0.5f;
      if (owns_fold) {
        if (p.uniforms) {
          const float* u = p.uniforms + (size_t)t * 11 * B;
#pragma unroll
          for (int i = 0; i < 10; ++i) ur[i] = __ldg(u + fold * 10 + i);
          ur[10] = __ldg(u + 10 * B + fold);
        } else {
          const unsigned g = (unsigned)(p.seg_first + fold), k0 = (unsigned)p.seed, k1 = (unsigned)(p.seed >> 32), o0 = (unsigned)p.offset;
          const Philox4 r0 = philox4x32_10((unsigned)t, g, 0u, o0, k0, k1), r1 = philox4x32_10((unsigned)t, g, 1u, o0, k0, k1),
                        r2 = philox4x32_10((unsigned)t, g, 2u, o0, k0, k1);
          ur[0] = u_ref_range(r0.x); ur[1] = u_ref_range(r0.y); ur[2] = u_ref_range(r0.z); ur[3] = u_ref_range(r0.w);
          ur[4] = u_ref_range(r1.x); ur[5] = u_ref_range(r1.y); ur[6] = u_ref_range(r1.z); ur[7] = u_ref_range(r1.w);
          ur[8] = u_ref_range(r2.x); ur[9] = u_ref_range(r2.y); ur[10] = u_ref_range(r2.z);
        }
      }
      float xf = 0.f;
      if (owns_fold && p.x_force && t > 0) xf = __ldg(p.x_force + (size_t)(t - 1) * B + fold);

      // ---- A: GRU1 ---------------------------------------------------------------------------------
      float pre2[12 + TU];                                // gi2 (12) and fc1 (32) conditioning terms, kept for B and C
      {
        mbar_wait(bar_q, (uint32_t)(t & 1), p.abort_flag);
        tc_fence_after();
        if (p.x_force && t > 0) x = xf;
        float pre[64];
        tmem_ld32(tlane + tq, pre);
        tmem_ld32(tlane + tq + 32, pre + 32);
        float gh[12];
        tmem_ld_sum<12, KW>(tlane + TC_B + 12, 64, gh);
#pragma unroll
        for (int q = 0; q < QROWS; ++q) pre[q] += fv[V_QK + q] + x * fv[V_VQ + q];
        float hv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float ghr = (t > 0 ? gh[j] : 0.f) + fv[V_B1H + j], ghz = (t > 0 ? gh[U + j] : 0.f) + fv[V_B1H + U + j],
                      ghn = (t > 0 ? gh[2 * U + j] : 0.f) + fv[V_B1H + 2 * U + j];
          h1[j] = gru_unit_fast(pre[j], pre[U + j], pre[2 * U + j], ghr, ghz, ghn, h1[j]);
          hv[j] = h1[j];
        }
#pragma unroll
        for (int q = 0; q < 12 + TU; ++q) pre2[q] = pre[12 + q];
        publish_global(0, par, hv);
      }
      if (profiling) { const long long c = clock64(); tprof[0] += c - tp0; tp0 = c; }

      // ---- B: [W2x ; W1h ; F1x_r] h1' -> GRU2 --------------------------------------------------------
      {
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float gi[12], gh[12];
        tmem_ld_sum<12, KW>(tlane + TC_B, 64, gi);
        tmem_ld_sum<12, KW>(tlane + TC_S2, N_S2, gh);
        float hv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float ghr = (t > 0 ? gh[j] : 0.f) + fv[V_B2H + j], ghz = (t > 0 ? gh[U + j] : 0.f) + fv[V_B2H + U + j],
                      ghn = (t > 0 ? gh[2 * U + j] : 0.f) + fv[V_B2H + 2 * U + j];
          h2[j] = gru_unit_fast(gi[j] + pre2[j], gi[U + j] + pre2[U + j], gi[2 * U + j] + pre2[2 * U + j], ghr, ghz, ghn, h2[j]);
          hv[j] = h2[j];
        }
        publish_global(1, par, hv);
      }
      if (profiling) { const long long c = clock64(); tprof[1] += c - tp0; tp0 = c; }

      // ---- C: F1x_r (h1' + h2') -> y1 rows of this rank, pushed to the cluster ------------------------
      {
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float acc[TU];
        tmem_ld_sum<16, KW>(tlane + TC_B + 24, 64, acc);
        tmem_ld_sum<16, KW>(tlane + TC_B + 40, 64, acc + 16);
#pragma unroll
        for (int j = 0; j < TU; ++j) acc[j] = fmaxf(acc[j] + pre2[12 + j], 0.f);
        push_cluster(smem_u32(smem + OFF_Y), SBO_Y, bar_y1, acc);
      }
      if (profiling) { const long long c = clock64(); tprof[2] += c - tp0; tp0 = c; }

      // ---- D: F2x_r y1 (+ F2a a4 through the K extension) -> y2 rows of this rank -> cluster -----------
      {
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float acc[TU];
        tmem_ld_sum<16, KW>(tlane + TC_F2, N_F2, acc);
        tmem_ld_sum<16, KW>(tlane + TC_F2 + 16, N_F2, acc + 16);
#pragma unroll
        for (int j = 0; j < TU; ++j) acc[j] = fmaxf(acc[j] + fv[V_BF2 + j], 0.f);
        push_cluster(smem_u32(smem + OFF_A), SBO_H, bar_y2, acc);
      }
      if (profiling) { const long long c = clock64(); tprof[3] += c - tp0; tp0 = c; }

      // ---- E: logits, MoL sample (replicated everywhere) ----------------------------------------------
      {
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float lg[32];
        tmem_ld_sum<16, KW>(tlane + TC_F3, N_F3, lg);
        tmem_ld_sum<16, KW>(tlane + TC_F3 + 16, N_F3, lg + 16);
#pragma unroll
        for (int i = 0; i < 30; ++i) lg[i] += fv[V_B3 + i];
        x = mol_sample_fast(lg, ur);
        if (owns_fold && cta == 0) {
          p.out[(size_t)fold * p.out_pitch + t] = x;
          if (p.logits_out) {
#pragma unroll
            for (int i = 0; i < 30; ++i) p.logits_out[((size_t)t * B + fold) * 30 + i] = lg[i];
          }
        }
      }
      if (profiling) { tprof[4] += clock64() - tp0; }
    }
    if (profiling) for (int i = 0; i < 5; ++i) p.prof[i] = tprof[i];

  } else {
    // =========================================================================================
    // warps 4-7: warps 4,5 issue the K halves of every chain; warp 4 also drives the global gathers;
    // warp 6 issues the conditioning chain; all four stage cond_{t+1} (and its a4 columns of the y1 image)
    // =========================================================================================
    const int q = warp - 4;
    const bool issuer = q < KW, leader = (q == 0);
    const uint32_t sA = smem_u32(smem + OFF_A), sY = smem_u32(smem + OFF_Y);
    const uint64_t dA = umma_desc(sA, 128, SBO_H), dY = umma_desc(sY, 128, SBO_Y), dC = umma_desc(smem_u32(smem + OFF_C), 128, SBO_Q);
    const uint64_t dWB = umma_desc(smem_u32(smem + OFF_WB), 128, SBO_H), dWF1 = umma_desc(smem_u32(smem + OFF_WB) + 3 * SBO_H, 128, SBO_H),
                   dWS2 = umma_desc(smem_u32(smem + OFF_WS2), 128, SBO_H), dWF2 = umma_desc(smem_u32(smem + OFF_WF2), 128, SBO_Y),
                   dWF3 = umma_desc(smem_u32(smem + OFF_WF3), 128, SBO_H), dWQ = umma_desc(smem_u32(smem + OFF_WQ), 128, SBO_Q);
    const uint32_t id_b = umma_idesc(MT, N_B, FMT), id_f1 = umma_idesc(MT, TU, FMT), id_s2 = umma_idesc(MT, N_S2, FMT),
                   id_f2 = umma_idesc(MT, N_F2, FMT), id_f3 = umma_idesc(MT, N_F3, FMT), id_q = umma_idesc(MT, N_Q, FMT);
    const bool profiling = (p.prof != nullptr) && cta == 0 && tid == 128;
    long long t_poll = 0, t_gather = 0, t_issue = 0;
    unsigned n_g = 0, n_y = 0;
    const int kh0 = q * (H / 16 / KW), kh1 = kh0 + H / 16 / KW;      // this warp's k-steps of a K=512 chain

    auto launch = [&](int v, unsigned target, const unsigned char* img) {      // leader only: counter -> TMA gather
      long long c0 = 0;
      if (profiling) c0 = clock64();
      if (lane == 0) counter_wait(p.counters + v, target, p.abort_flag);
      __syncwarp();
      proxy_fence_global();
      if (profiling) t_poll += clock64() - c0;
      tma_bulk_g2s(sA, img, img_bytes, bar_g);
    };
    auto mmas = [&](uint64_t da, uint64_t db, int k0, int k1, uint32_t d_col, uint32_t idesc, bool acc0) {
      for (int k = k0; k < k1; ++k) umma_f16(tmem + d_col, da + (uint64_t)(k * 16), db + (uint64_t)(k * 16), idesc, (acc0 || k > k0) ? 1u : 0u);
    };
    auto cond_chain = [&](uint32_t d_col) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < CDIM / 16; ++k) umma_f16(tmem + d_col, dC + (uint64_t)(k * 16), dWQ + (uint64_t)(k * 16), id_q, k > 0);
      umma_commit(bar_q);
    };

    // ---- conditioning staging (128 threads) -----------------------------------------------------------
    const int st = tid - 128;
    constexpr int COND_TASKS = 5;
    const int n_tasks = B * KQ;                            // <= 24 * 26 = 624 <= 5 * 128
    float4 creg[COND_TASKS][2];
    auto cond_fetch = [&](int n) {
#pragma unroll
      for (int j = 0; j < COND_TASKS; ++j) {
        const int task = st + j * 128;
        creg[j][0] = make_float4(0.f, 0.f, 0.f, 0.f); creg[j][1] = creg[j][0];
        if (task < n_tasks) {
          const int f = task / KQ, c8 = task % KQ;
          const long long row = (long long)f * p.seg_stride + n;
          if (row < p.L) {
            const float* s = (c8 < FEAT / 8) ? p.mels_up + row * FEAT + c8 * 8 : p.aux + row * (4 * AUXD) + (c8 - FEAT / 8) * 8;
            creg[j][0] = __ldg(reinterpret_cast<const float4*>(s)); creg[j][1] = __ldg(reinterpret_cast<const float4*>(s) + 1);
          }
        }
      }
    };
    auto cond_store = [&](int n) {                          // cond_n -> cond image; its a4 part also -> y1 image block (n & 1)
#pragma unroll
      for (int j = 0; j < COND_TASKS; ++j) {
        const int task = st + j * 128;
        if (task < n_tasks) {
          const int f = task / KQ, c8 = task % KQ;
          uint4 v;
          v.x = pack2<FMT>(creg[j][0].x, creg[j][0].y); v.y = pack2<FMT>(creg[j][0].z, creg[j][0].w);
          v.z = pack2<FMT>(creg[j][1].x, creg[j][1].y); v.w = pack2<FMT>(creg[j][1].z, creg[j][1].w);
          *reinterpret_cast<uint4*>(smem + OFF_C + (f >> 3) * SBO_Q + c8 * 128 + (f & 7) * 16) = v;
          if (c8 >= (CDIM - AUXD) / 8)                      // a4 = cond[176:208): chunks 22..25
            *reinterpret_cast<uint4*>(smem + OFF_Y + (f >> 3) * SBO_Y + (H / 8 + (n & 1) * (AUXD / 8) + (c8 - (CDIM - AUXD) / 8)) * 128 + (f & 7) * 16) = v;
        }
      }
      proxy_fence_smem();
    };

    cond_fetch(0);
    cond_store(0);
    named_bar_sync(2, 128);
    if (q == 2) cond_chain(TC_Q0);
    if (S > 1) cond_fetch(1);

    for (int t = 0; t < S; ++t) {
      const int par = t & 1;
      const unsigned target = (unsigned)P * (unsigned)(t + 1);
      const unsigned char* base = p.xch + (size_t)par * img_bytes;
      long long c0 = 0, c1 = 0;
      // ---- B: h1' (global) ----
      if (leader) launch(0, target, base + 0 * xch_stride);
      if (issuer) {
        if (profiling) c0 = clock64();
        mbar_wait(bar_g, n_g & 1, p.abort_flag);
        tc_fence_after();
        if (profiling) c1 = clock64();
        mmas(dA, dWB, kh0, kh1, TC_B + q * 64, id_b, false);
        umma_commit(bar_mma);
        if (profiling) { const long long c2 = clock64(); t_gather += c1 - c0; t_issue += c2 - c1; }
      }
      ++n_g;
      // ---- C: h2' (global): W2h rows fresh, F1x rows accumulate onto the h1' part ----
      if (leader) launch(1, target, base + 1 * xch_stride);
      if (issuer) {
        if (profiling) c0 = clock64();
        mbar_wait(bar_g, n_g & 1, p.abort_flag);
        tc_fence_after();
        if (profiling) c1 = clock64();
        mmas(dA, dWS2, kh0, kh1, TC_S2 + q * N_S2, id_s2, false);
        mmas(dA, dWF1, kh0, kh1, TC_B + q * 64 + 24, id_f1, true);
        umma_commit(bar_mma);
        if (profiling) { const long long c2 = clock64(); t_gather += c1 - c0; t_issue += c2 - c1; }
      }
      ++n_g;
      if (t + 1 < S) {                                      // conditioning of step t+1 (queued behind the phase-C chains)
        mbar_wait(bar_q, (uint32_t)(t & 1), p.abort_flag);
        cond_store(t + 1);
        named_bar_sync(2, 128);
        if (q == 2) cond_chain(par ? TC_Q0 : TC_Q1);
        if (t + 2 < S) cond_fetch(t + 2);
      }
      // ---- D: y1 (cluster pushes) | a4 block of this step ----
      if (issuer) {
        mbar_wait_cluster(bar_y1, n_y & 1, p.abort_flag);
        proxy_fence_smem();                                  // remote generic-proxy stores -> tensor-core (async proxy) reads
        tc_fence_after();
        if (q == 0) { mmas(dY, dWF2, 0, 17, TC_F2 + 0 * N_F2, id_f2, false); }
        else {
          mmas(dY, dWF2, 17, 32, TC_F2 + 1 * N_F2, id_f2, false);
          const int ka = H / 16 + par * (AUXD / 16);          // the a4 columns written for THIS step
          mmas(dY, dWF2, ka, ka + AUXD / 16, TC_F2 + 1 * N_F2, id_f2, true);
        }
        umma_commit(bar_mma);
      }
      // ---- E: y2 (cluster pushes into the A image) ----
      if (issuer) {
        mbar_wait_cluster(bar_y2, n_y & 1, p.abort_flag);
        proxy_fence_smem();
        tc_fence_after();
        mmas(dA, dWF3, kh0, kh1, TC_F3 + q * N_F3, id_f3, false);
        umma_commit(bar_mma);
      }
      ++n_y;
    }
    if (profiling) { p.prof[5] = t_poll; p.prof[6] = t_gather; p.prof[7] = t_issue; }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // no CTA may leave while a peer can still push into it
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS));
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
class TccEngine : public Engine {
 public:
  ~TccEngine() override {
    cudaSetDevice(device);
    cudaFree(d_blob_); cudaFree(d_scratch_); cudaFree(d_sync_);
  }
  const char* name() const override { return cfg.precision == WRNN_PREC_BF16 ? "tcgen05c-bf16" : "tcgen05c-fp16"; }
  int grid_ctas() const override { return P; }
  const void* kernel() const { return cfg.precision == WRNN_PREC_BF16 ? (const void*)wrnn_tcc_kernel<1> : (const void*)wrnn_tcc_kernel<0>; }
  bool supports(const wrnn_job& job) const override { return job.n_seg <= MAXB; }

  int init(const HostWeights& w) {
    Folded f; fold(w, f);
    const bool bf = cfg.precision == WRNN_PREC_BF16;
    auto cvt = [&](double v) -> uint16_t { return bf ? f2bf((float)v) : f2h((float)v); };
    const int W2W = H + AUXD;
    std::vector<unsigned char> blob((size_t)WEIGHT_BYTES * P, 0);
    CtaSlice s;
    for (int c = 0; c < P; ++c) {
      slice_for_cta(w, f, c, U, s);
      const int r0 = (c % CL) * TU;                          // first tail row (fc1 / fc2) of this CTA
      unsigned char* base = blob.data() + (size_t)c * WEIGHT_BYTES;
      uint16_t* wb = reinterpret_cast<uint16_t*>(base + (OFF_WB - OFF_WB));
      uint16_t* ws2 = reinterpret_cast<uint16_t*>(base + (OFF_WS2 - OFF_WB));
      uint16_t* wf2 = reinterpret_cast<uint16_t*>(base + (OFF_WF2 - OFF_WB));
      uint16_t* wf3 = reinterpret_cast<uint16_t*>(base + (OFF_WF3 - OFF_WB));
      uint16_t* wq = reinterpret_cast<uint16_t*>(base + (OFF_WQ - OFF_WB));
      float* fv = reinterpret_cast<float*>(base + (OFF_VEC - OFF_WB));
      for (int k = 0; k < H; ++k) {
        for (int r = 0; r < 6 * U; ++r) wb[img_index(r, k, H)] = cvt(s.S1[(size_t)r * H + k]);          // W2x 12 | W1h 12
        for (int r = 0; r < TU; ++r) wb[img_index(24 + r, k, H)] = cvt(w.f1w[(size_t)(r0 + r) * W2W + k]);   // F1x rows of this rank
        for (int r = 0; r < 3 * U; ++r) ws2[img_index(r, k, H)] = cvt(s.S2[(size_t)(U + r) * H + k]);   // W2h 12
        for (int r = 0; r < TU; ++r) wf2[img_index(r, k, KY)] = cvt(w.f2w[(size_t)(r0 + r) * W2W + k]);
        for (int r = 0; r < cfg.n_classes; ++r) wf3[img_index(r, k, H)] = cvt(w.f3w[(size_t)r * H + k]);
      }
      for (int k = 0; k < AUXD; ++k)
        for (int r = 0; r < TU; ++r) {
          const uint16_t v = cvt(w.f2w[(size_t)(r0 + r) * W2W + H + k]);                                 // F2a, both step-parity blocks
          wf2[img_index(r, H + k, KY)] = v; wf2[img_index(r, H + AUXD + k, KY)] = v;
        }
      // Q: gi1 12 | gi2 12 from the per-CTA slice; fc1 rows of this rank from the folded model
      for (int k = 0; k < CDIM; ++k) {
        for (int r = 0; r < 6 * U; ++r) wq[img_index(r, k, CDIM)] = cvt(s.Q[(size_t)r * CDIM + k]);
        for (int r = 0; r < TU; ++r) {
          double v = 0.0;
          if (k < F1IN) v = f.A3[(size_t)(r0 + r) * F1IN + k];
          else if (k >= F1IN + AUXD && k < F1IN + 2 * AUXD) v = w.f1w[(size_t)(r0 + r) * W2W + H + (k - F1IN - AUXD)];
          wq[img_index(24 + r, k, CDIM)] = cvt(v);
        }
      }
      for (int r = 0; r < 6 * U; ++r) { fv[V_QK + r] = s.qk[r]; fv[V_VQ + r] = s.vq[r]; }
      for (int r = 0; r < TU; ++r) { fv[V_QK + 24 + r] = (float)f.k3[r0 + r]; fv[V_VQ + 24 + r] = (float)f.v3[r0 + r]; fv[V_BF2 + r] = w.f2b[r0 + r]; }
      for (int r = 0; r < 3 * U; ++r) { fv[V_B1H + r] = s.b1h[r]; fv[V_B2H + r] = s.b2h[r]; }
      for (int r = 0; r < cfg.n_classes; ++r) fv[V_B3 + r] = w.f3b[r];
    }
    WRNN_CUDA_OK(cudaMalloc(&d_blob_, blob.size()));
    WRNN_CUDA_OK(cudaMemcpy(d_blob_, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    WRNN_CUDA_OK(cudaMalloc(&d_sync_, 256));
    WRNN_CUDA_OK(cudaMemset(d_sync_, 0, 256));
    scratch_bytes_ = (size_t)2 * 2 * NG * SBO_H;
    WRNN_CUDA_OK(cudaMalloc(&d_scratch_, scratch_bytes_));
    WRNN_CUDA_OK(cudaFuncSetAttribute(kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    WRNN_CUDA_OK(cudaFuncSetAttribute(kernel(), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // can 8 clusters of 16 CTAs (one CTA per SM) be co-resident on this part?
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(P); lc.blockDim = dim3(NT); lc.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int n_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n_clusters, kernel(), &lc);
    if (e != cudaSuccess || n_clusters < P / CL) {
      cudaGetLastError();
      set_error("cluster-tail engine: the device cannot co-schedule 8 clusters of 16 CTAs (max active clusters = " +
                std::to_string(n_clusters) + ")");
      return WRNN_E_INVALID;
    }
    return WRNN_OK;
  }

  int generate(const wrnn_job& job, cudaStream_t stream) override {
    if (job.n_seg > MAXB) { set_error("cluster-tail engine: n_seg must be <= 24"); return WRNN_E_INVALID; }
    WRNN_CUDA_OK(cudaSetDevice(device));
    WRNN_CUDA_OK(cudaMemsetAsync(d_sync_, 0, 256, stream));
    TccParams p{};
    p.blob = static_cast<const unsigned char*>(d_blob_);
    p.mels_up = job.mels_up; p.aux = job.aux; p.L = job.L; p.seg_stride = job.seg_stride;
    p.n_seg = job.n_seg; p.steps = job.steps > 0 ? job.steps : job.seg_len; p.out_pitch = p.steps;
    p.seg_first = job.seg_first;
    p.uniforms = job.uniforms; p.seed = job.philox_seed; p.offset = job.philox_offset;
    p.out = job.out; p.x_force = job.x_force; p.logits_out = job.logits_out;
    p.xch = static_cast<unsigned char*>(d_scratch_);
    p.counters = static_cast<unsigned*>(d_sync_);
    p.abort_flag = reinterpret_cast<int*>(static_cast<unsigned*>(d_sync_) + 8);
    p.prof = reinterpret_cast<long long*>(static_cast<unsigned char*>(d_sync_) + 64);
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(P); lc.blockDim = dim3(NT); lc.dynamicSmemBytes = SMEM_BYTES; lc.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;     // all 128 CTAs co-resident (spin waits on L2 counters)
    lc.attrs = at; lc.numAttrs = 2;
    void* args[] = {&p};
    WRNN_CUDA_OK(cudaLaunchKernelExC(&lc, kernel(), args));
    ++launches;
    last_steps_ = p.steps;
    return WRNN_OK;
  }

  int check() override {
    unsigned char buf[256];
    WRNN_CUDA_OK(cudaSetDevice(device));
    WRNN_CUDA_OK(cudaMemcpy(buf, d_sync_, 256, cudaMemcpyDeviceToHost));
    const int flag = reinterpret_cast<int*>(buf)[8];
    long long prof[8];
    std::memcpy(prof, buf + 64, sizeof(prof));
    if (getenv("WRNN_TC_PROF") && last_steps_ > 0) {
      const long long n = last_steps_;
      fprintf(stderr, "[wrnn_tcc prof] steps=%d | fold thread: A=%lld B=%lld C=%lld D=%lld E=%lld | issuer warp 4: poll=%lld gather=%lld issue=%lld (cycles per step)\n",
              last_steps_, prof[0] / n, prof[1] / n, prof[2] / n, prof[3] / n, prof[4] / n, prof[5] / n, prof[6] / n, prof[7] / n);
    }
    if (flag != 0) {
      set_error(flag == 3 ? "cluster-tail kernel aborted: a DSMEM exchange wait timed out"
                : flag == 2 ? "cluster-tail kernel aborted: an mbarrier wait (MMA / TMA completion) timed out"
                            : "cluster-tail kernel aborted: an inter-SM exchange wait timed out");
      return WRNN_E_WATCHDOG;
    }
    return WRNN_OK;
  }

 private:
  void *d_blob_ = nullptr, *d_scratch_ = nullptr, *d_sync_ = nullptr;
  size_t scratch_bytes_ = 0;
  int last_steps_ = 0;
};

}  // namespace

int make_tcc_engine(const wrnn_cfg& cfg, const HostWeights& w, int device, Engine** out) {
  if (cfg.mode != WRNN_MODE_MOL || cfg.n_classes != 30 || cfg.precision == WRNN_PREC_FP32) {
    set_error("cluster-tail engine serves the MoL head with fp16/bf16 operands");
    return WRNN_E_INVALID;
  }
  TccEngine* e = new TccEngine();
  e->cfg = cfg; e->device = device;
  const int rc = e->init(w);
  if (rc != WRNN_OK) { delete e; return rc; }
  *out = e;
  return WRNN_OK;
}

}  // namespace wrnn
