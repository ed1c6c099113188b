// wrnn_tc.cu -- tcgen05 engine: the generate() loop (reference models/fatchord_version.py:
// 201-241 + utils/distribution.py:87-123) as ONE persistent kernel whose dense contractions
// run on the 5th-generation tensor cores with accumulators in TMEM.
//
// Decomposition (same row ownership as the SIMT engine, wrnn_fold.h): P = 128 CTAs, CTA c
// owns hidden units [4c, 4c+4) of every layer.  Its weight slices are staged ONCE into shared
// memory as UMMA K-major (no-swizzle) operand images and stay there for the whole sequence:
//     S1 = [W2x ; W1h ; F1x] rows (N=32, K=512)   consumes h1'
//     S2 = [F1x ; W2h] rows      (N=16, K=512)   consumes h2'
//     S3 = F2x rows              (N=8,  K=512)   consumes y1
//     F3 = fc3 (all 30 rows)     (N=32, K=512)   consumes y2 (replicated in every CTA)
//     Q  = folded conditioning rows (N=32, K=208) consumes cond_t, runs one step ahead
// The folds are the M dimension ("swap-AB"): tile of M=64 folds, so in TMEM lane == fold and
// every epilogue (GRU gates, relu, MoL sampling) is thread-local: one thread owns one fold,
// reads its accumulator columns with tcgen05.ld and keeps that fold's hidden state in
// registers for the whole sequence.
//
// Per step only the four 512-wide activation vectors cross SMs.  They are exchanged through
// L2-resident buffers that are byte images of the UMMA A-operand layout, so a consumer pulls a
// whole vector with ONE TMA bulk copy (cp.async.bulk global -> shared, mbarrier completion)
// straight into the operand buffer; arrival is a release/acquire counter per vector.
// fc3 + sampling is replicated in every CTA (same inputs, same arithmetic => bitwise
// identical samples), which removes the fifth exchange of the step.
//
// Warp roles (no CTA-wide barrier inside the step loop):
//   warps 0-3  fold warps : TMEM -> registers (sum of the K-quarter partials), gates / relu / sampler, publish, signal
//   warps 4-7  issuers    : warp 4+q owns the K quarter [128q, 128q+128) of every K=512 chain (8 tcgen05.mma each,
//                           warp-uniform, elect.sync-predicated, own accumulator columns).  Warp 4 also polls the
//                           arrival counters and launches the TMA gathers; warp 5 issues the conditioning chain; all
//                           four stream cond_{t+1} HBM -> registers (a step ahead) -> fp16 operand image.
//
// One launch serves a tile of <= 64 folds (one M tile); larger jobs run tile after tile.
// RAW 9-bit head (template RAW; reference :231-237): fc3 has 512 = 4 x 128 rows, so CTA c owns classes
// [4c, 4c+4) like every other layer.  Categorical(softmax(l)).sample() == argmax_k(p_k / e_k), e ~ Exp(1),
// == argmax_k(l_k - log e_k): each CTA publishes its best (score, class) per fold, a FIFTH exchange gathers the
// 128 candidates per fold into every CTA, and every CTA reduces them identically (ties -> lowest class).
// The fp32 strict mode is served by the SIMT engine (ENGINE_AUTO falls through).
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

constexpr int P = 128;             // CTAs == weight shards
constexpr int U = H / P;           // 4 hidden units per CTA
constexpr int NT = 256;            // 8 warps, roles above
constexpr int MT = 64;             // folds per M tile
constexpr int KC = H / 8;          // 64 16-byte chunks per activation row
constexpr int SBO_H = KC * 128;    // 8192: byte stride between 8-row groups, K = 512 images
constexpr int KQ = CDIM / 8;       // 26 chunks per conditioning row
constexpr int SBO_Q = KQ * 128;    // 3328

constexpr int N_S1 = 32, N_S2 = 16, N_S3 = 8, N_F3 = 32, N_Q = 32;
// shared memory map (bytes)
constexpr int OFF_A = 0;                                  // activation A image, 8 row groups
constexpr int OFF_COND = OFF_A + 8 * SBO_H;               // conditioning A image (8 groups x 3328)
constexpr int OFF_S1 = OFF_COND + 8 * SBO_Q;
constexpr int OFF_S2 = OFF_S1 + (N_S1 / 8) * SBO_H;
constexpr int OFF_S3 = OFF_S2 + (N_S2 / 8) * SBO_H;
constexpr int OFF_F3 = OFF_S3 + (N_S3 / 8) * SBO_H;
constexpr int OFF_Q = OFF_F3 + (N_F3 / 8) * SBO_H;
constexpr int OFF_VEC = OFF_Q + (N_Q / 8) * SBO_Q;        // fp32: qk[32] vq[32] b1h[12] b2h[12] b3[32] pad -> 128 floats
constexpr int NVEC = 128;
constexpr int OFF_BAR = OFF_VEC + NVEC * 4;               // 3 mbarriers + tmem base
constexpr int SMEM_BYTES = OFF_BAR + 64;
constexpr int WEIGHT_BYTES = OFF_VEC + NVEC * 4 - OFF_S1; // per-CTA blob == smem[OFF_S1, OFF_BAR)
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// TMEM columns (fp32 accumulators; M=64 => lanes 0-15 of each lane quarter)
// Every K=512 chain is split over KW issuing warps (K quarters), each accumulating into its own columns;
// the fold threads add the KW partials when they read them (fixed order => identical in every CTA).
constexpr int KW = 4;
constexpr int TC_S1 = 0, TC_S2 = TC_S1 + KW * N_S1, TC_S3 = TC_S2 + KW * N_S2, TC_F3 = TC_S3 + KW * N_S3,
              TC_Q0 = TC_F3 + KW * N_F3, TC_Q1 = TC_Q0 + N_Q, TMEM_COLS = 512;
static_assert(TC_Q1 + N_Q <= TMEM_COLS, "TMEM budget");

struct TcParams {
  const unsigned char* blob;
  const float* mels_up; const float* aux; long long L; long long seg_stride;
  long long row_base;        // rows are numbered (global fold)*seg_stride - row_base (tile-local conditioning scratch)
  int n_seg, steps, out_pitch, seg_first;   // n_seg = folds of THIS launch's tile (<= 64)
  int f0, n_total;                         // first fold of the tile / folds of the whole job (indexing of the job-wide arrays)
  const float* uniforms; const float* expo; unsigned long long seed, offset;   // expo: RAW head, [steps, n_total, 512]
  const unsigned* uniforms_ready;   // optional: rows of `uniforms` uploaded so far (draws streamed in while the kernel runs)
  float* out; const float* x_force; float* logits_out;
  const long long* fold_row0; const long long* fold_row_end;   // optional per-fold conditioning windows (job-wide, [n_total])
  const float* mel_frames; const float* aux_frames; const float* up_taps; int hop;   // optional frame-rate conditioning
  unsigned char* xch;        // [4 vectors][2 parities][n_groups * SBO_H] activation images
  unsigned char* xch5;       // RAW: [2 parities][128 CTAs][n_groups * 8 folds] (score, class) candidates
  unsigned* counters;        // [5] monotonically increasing arrival counters
  int* abort_flag;
  long long* prof;           // cycle counters of CTA 0 (fold thread 0: [0..4], driver lane 0: [5..7])
};

// ------------------------------------------------------------------------------------------
// conditioning pre-pass (WRNN_COND_EXPAND): rows [r_lo, r_lo + n_rows) of the per-sample conditioning
// stream from the frame-rate tensors -- the whole UpsampleNetwork tail (three stretch+conv stages as a
// 5-tap table, aux as a nearest repeat; reference fatchord_version.py:73-88) as one HBM-bound pass.
// One thread per float4 of a row (20 of mel, 32 of aux), 8 rows per block pass; 832 B written per row, frames from L2.
// The tap sum uses the same fmaf order as the in-kernel staging, so both modes give identical rows.
// ------------------------------------------------------------------------------------------
constexpr int XP_ROWS = 8, XP_Q = CDIM / 4, XP_THREADS = XP_ROWS * XP_Q;      // 8 rows x 52 float4 per block pass
__global__ void __launch_bounds__(XP_THREADS) wrnn_expand_rows_kernel(const float* __restrict__ mel_frames, const float* __restrict__ aux_frames,
                                                                      const float* __restrict__ taps, int hop, long long r_lo, long long n_rows,
                                                                      float* __restrict__ m_out, float* __restrict__ a_out) {
  constexpr int QM = FEAT / 4;
  const int rr = threadIdx.x / XP_Q, q = threadIdx.x - rr * XP_Q;            // row within the pass, float4 within the row
  for (long long rl = (long long)blockIdx.x * XP_ROWS + rr; rl < n_rows; rl += (long long)gridDim.x * XP_ROWS) {
    const unsigned r = (unsigned)(r_lo + rl);
    const unsigned fr = r / (unsigned)hop, ph = r - fr * (unsigned)hop;
    if (q >= QM) {
      __stcs(reinterpret_cast<float4*>(a_out + rl * (4 * AUXD)) + (q - QM),
             __ldg(reinterpret_cast<const float4*>(aux_frames + (size_t)fr * (4 * AUXD)) + (q - QM)));
    } else {
      const float* k = taps + ph * 5;
      const float* m = mel_frames + (size_t)fr * FEAT + q * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        const float wgt = __ldg(k + d);
        const float4 x = __ldg(reinterpret_cast<const float4*>(m + d * FEAT));
        a.x = fmaf(wgt, x.x, a.x); a.y = fmaf(wgt, x.y, a.y); a.z = fmaf(wgt, x.z, a.z); a.w = fmaf(wgt, x.w, a.w);
      }
      __stcs(reinterpret_cast<float4*>(m_out + rl * FEAT) + q, a);
    }
  }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int FMT, bool FRAMES, bool RAW, bool BIG>
__global__ void __launch_bounds__(NT, 1) wrnn_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const float* fv = reinterpret_cast<const float*>(smem + OFF_VEC);
  const float* qk = fv; const float* vq = fv + 32; const float* b1h = fv + 64; const float* b2h = fv + 76;
  const float* b3 = fv + 88;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 32);
  const uint32_t bar_mma = smem_u32(&bars[0]), bar_q = smem_u32(&bars[1]), bar_g = smem_u32(&bars[2]);

  const int cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = p.n_seg, S = p.steps, u0 = cta * U;
  const int n_groups = (B + 7) / 8;                    // real 8-row groups of the A images
  const uint32_t img_bytes = (uint32_t)n_groups * SBO_H;
  const size_t xch_stride = (size_t)2 * img_bytes;     // per vector: two parities
  constexpr int GPS = RAW ? 5 : 4;                     // gathers (completions of bar_g) per step
  constexpr int NCLS = 4 * P;                          // RAW classes
  const uint32_t cand_bytes = (uint32_t)P * n_groups * 8 * 8;   // RAW: one parity of the candidate buffer

  // ---- one-time setup: weights -> smem images, barriers, TMEM --------------------------------
  {
    const int4* src = reinterpret_cast<const int4*>(p.blob + (size_t)cta * WEIGHT_BYTES);
    int4* dst = reinterpret_cast<int4*>(smem + OFF_S1);
    for (int i = tid; i < WEIGHT_BYTES / 16; i += NT) dst[i] = src[i];
    int4* z = reinterpret_cast<int4*>(smem);
    for (int i = tid; i < OFF_S1 / 16; i += NT) z[i] = make_int4(0, 0, 0, 0);   // A images start as zeros
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_mma), "n"(KW));   // one commit per issuing warp
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_q));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_g));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  proxy_fence_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // =========================================================================================
    // fold warps: M=64 accumulators put fold f in TMEM lane 32*(f/16) + f%16
    // =========================================================================================
    const int fold = warp * 16 + lane;
    const bool owns_fold = lane < 16 && fold < B;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const size_t pub_off = (size_t)(fold >> 3) * SBO_H + (u0 >> 3) * 128 + (fold & 7) * 16 + (u0 & 7) * 2;
    const bool profiling = (p.prof != nullptr) && cta == 0 && tid == 0;
    long long tprof[5] = {0, 0, 0, 0, 0};
    auto publish = [&](unsigned char* img, const float* v) {   // this fold's 4 values of this CTA's units
      uint2 w;
      w.x = pack2<FMT>(v[0], v[1]); w.y = pack2<FMT>(v[2], v[3]);
      if (owns_fold) *reinterpret_cast<uint2*>(img + pub_off) = w;
    };
    auto signal = [&](int v) {                           // all fold warps have stored: one release increment per CTA
      tc_fence_before();
      named_bar_sync(1, 128);
      if (tid == 0) red_release_add_u32(p.counters + v, 1u);
    };
    float h1[U] = {0.f, 0.f, 0.f, 0.f}, h2[U] = {0.f, 0.f, 0.f, 0.f};
    float x = 0.f;
    unsigned n_mma = 0;                                   // completed phases of bar_mma
    unsigned rows_known = p.uniforms_ready ? 0u : 0xffffffffu;   // rows of p.uniforms known to have landed

    for (int t = 0; t < S; ++t) {
      const int par = t & 1;
      const uint32_t tq = (par ? TC_Q1 : TC_Q0);
      unsigned char* img_h1 = p.xch + 0 * xch_stride + (size_t)par * img_bytes;
      unsigned char* img_h2 = p.xch + 1 * xch_stride + (size_t)par * img_bytes;
      unsigned char* img_y1 = p.xch + 2 * xch_stride + (size_t)par * img_bytes;
      unsigned char* img_y2 = p.xch + 3 * xch_stride + (size_t)par * img_bytes;
      long long tp0 = 0;
      if (profiling) tp0 = clock64();

      // draws for this step, fetched early so their latency hides under the step
      float ur[11];
#pragma unroll
      for (int i = 0; i < 11; ++i) ur[i] = 0.5f;
      float ex[4] = {1.f, 1.f, 1.f, 1.f};                 // RAW: Exp(1) draws of this CTA's 4 classes
      if constexpr (RAW) {
        if (owns_fold) {
          if (p.expo) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.expo + ((size_t)t * p.n_total + p.f0 + fold) * NCLS + u0));
            ex[0] = v.x; ex[1] = v.y; ex[2] = v.z; ex[3] = v.w;
          } else {                                        // same keying as the SIMT engine: block 16 + class/4
            const Philox4 r = philox4x32_10((unsigned)t, (unsigned)(p.seg_first + p.f0 + fold), 16u + (unsigned)cta,
                                            (unsigned)p.offset, (unsigned)p.seed, (unsigned)(p.seed >> 32));
            ex[0] = -logf(u01(r.x)); ex[1] = -logf(u01(r.y)); ex[2] = -logf(u01(r.z)); ex[3] = -logf(u01(r.w));
          }
        }
      } else if (p.uniforms) {
        if (p.uniforms_ready) {
          // Streamed draws: rows land (DMA) while the kernel runs.  All 128 CTAs read the same row, so the loads must stay
          // L1-cached (as L2 loads they cost 1 us per step: 27 k requests on a few hot lines, measured) -- which is safe
          // only if no line is ever cached before all of it has landed: a row is >= 44 B, a 128 B line spans at most four
          // rows, so row t is read once rows t .. t+3 are in (the host uploads in chunks of ~1000 rows: free).
          const unsigned need = min((unsigned)S, (unsigned)t + 4u);
          if (rows_known < need) rows_known = rows_wait(p.uniforms_ready, need, p.abort_flag);
          if (owns_fold) {
            const float* u = p.uniforms + (size_t)t * 11 * p.n_total;
#pragma unroll
            for (int i = 0; i < 10; ++i) ur[i] = __ldca(u + (p.f0 + fold) * 10 + i);
            ur[10] = __ldca(u + 10 * p.n_total + p.f0 + fold);
          }
        } else if (owns_fold) {                            // resident draws: read-only path (11 values = 2 lines per fold)
          const float* u = p.uniforms + (size_t)t * 11 * p.n_total;
#pragma unroll
          for (int i = 0; i < 10; ++i) ur[i] = __ldg(u + (p.f0 + fold) * 10 + i);
          ur[10] = __ldg(u + 10 * p.n_total + p.f0 + fold);
        }
      } else if (owns_fold) {
        {
          const unsigned g = (unsigned)(p.seg_first + p.f0 + fold), k0 = (unsigned)p.seed, k1 = (unsigned)(p.seed >> 32), o0 = (unsigned)p.offset;
          const Philox4 r0 = philox4x32_10((unsigned)t, g, 0u, o0, k0, k1), r1 = philox4x32_10((unsigned)t, g, 1u, o0, k0, k1),
                        r2 = philox4x32_10((unsigned)t, g, 2u, o0, k0, k1);
          ur[0] = u_ref_range(r0.x); ur[1] = u_ref_range(r0.y); ur[2] = u_ref_range(r0.z); ur[3] = u_ref_range(r0.w);
          ur[4] = u_ref_range(r1.x); ur[5] = u_ref_range(r1.y); ur[6] = u_ref_range(r1.z); ur[7] = u_ref_range(r1.w);
          ur[8] = u_ref_range(r2.x); ur[9] = u_ref_range(r2.y); ur[10] = u_ref_range(r2.z);
        }
      }
      float xf = 0.f;
      if (owns_fold && p.x_force && t > 0) xf = __ldg(p.x_force + (size_t)(t - 1) * p.n_total + p.f0 + fold);

      // ---- A: GRU1.  gi1 = pre_t (conditioning chain, D_Q) + x * v1 ; gh1 = W1h h1 + b1h (D_S1 of step t-1)
      float pre[32];                                      // this fold's 32 conditioning rows (+ qk + x*vq)
      {
        mbar_wait(bar_q, (uint32_t)(t & 1), p.abort_flag);           // pre_t landed in D_Q[par]
        tc_fence_after();
        if (p.x_force && t > 0) x = xf;
        tmem_ld32(tlane + tq, pre);
        float gh[12];
        tmem_ld_sum<12, KW>(tlane + TC_S1 + 12, N_S1, gh);           // W1h h1 (step t-1, phase B); also completes `pre`
#pragma unroll
        for (int q = 0; q < 32; ++q) pre[q] += qk[q] + x * vq[q];
        float hv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float ghr = (t > 0 ? gh[j] : 0.f) + b1h[j], ghz = (t > 0 ? gh[U + j] : 0.f) + b1h[U + j],
                      ghn = (t > 0 ? gh[2 * U + j] : 0.f) + b1h[2 * U + j];
          h1[j] = gru_unit_fast(pre[j], pre[U + j], pre[2 * U + j], ghr, ghz, ghn, h1[j]);
          hv[j] = h1[j];
        }
        publish(img_h1, hv);
        signal(0);
      }
      if (profiling) { const long long c = clock64(); tprof[0] += c - tp0; tp0 = c; }

      // ---- B: [W2x ; W1h ; F1x] h1'  ->  GRU2, gh1 for step t+1, fc1 partial ---------------------
      {
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float gi[12], gh[12];
        tmem_ld_sum<12, KW>(tlane + TC_S1, N_S1, gi);                // W2x h1'
        tmem_ld_sum<12, KW>(tlane + TC_S2 + 4, N_S2, gh);            // W2h h2 (step t-1, phase C)
        float hv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float ghr = (t > 0 ? gh[j] : 0.f) + b2h[j], ghz = (t > 0 ? gh[U + j] : 0.f) + b2h[U + j],
                      ghn = (t > 0 ? gh[2 * U + j] : 0.f) + b2h[2 * U + j];
          h2[j] = gru_unit_fast(gi[j] + pre[3 * U + j], gi[U + j] + pre[4 * U + j], gi[2 * U + j] + pre[5 * U + j], ghr, ghz, ghn, h2[j]);
          hv[j] = h2[j];
        }
        publish(img_h2, hv);
        signal(1);
      }
      if (profiling) { const long long c = clock64(); tprof[1] += c - tp0; tp0 = c; }

      // ---- C: [F1x ; W2h] h2'  ->  y1 = relu(fc1), gh2 for step t+1 ------------------------------
      {
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float a[4], b4[4];
        tmem_ld_sum<4, KW>(tlane + TC_S2, N_S2, a);                 // F1x h2'
        tmem_ld_sum<4, KW>(tlane + TC_S1 + 24, N_S1, b4);           // F1x h1' (phase B)
        float yv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) yv[j] = fmaxf(a[j] + b4[j] + pre[6 * U + j], 0.f);
        publish(img_y1, yv);
        signal(2);
      }
      if (profiling) { const long long c = clock64(); tprof[2] += c - tp0; tp0 = c; }

      // ---- D: F2x y1 -> y2 = relu(fc2) --------------------------------------------------------------
      {
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float a[4];
        tmem_ld_sum<4, KW>(tlane + TC_S3, N_S3, a);
        float yv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) yv[j] = fmaxf(a[j] + pre[7 * U + j], 0.f);
        publish(img_y2, yv);
        signal(3);
      }
      if (profiling) { const long long c = clock64(); tprof[3] += c - tp0; tp0 = c; }

      // ---- E: logits = F3 y2 + b3, sample ---------------------------------------------------------------
      if constexpr (!RAW) {                                 // MoL: fc3 + sampler replicated in every CTA
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float lg[32];
        tmem_ld_sum<16, KW>(tlane + TC_F3, N_F3, lg);               // two halves keep the register peak down
        tmem_ld_sum<16, KW>(tlane + TC_F3 + 16, N_F3, lg + 16);
#pragma unroll
        for (int i = 0; i < 30; ++i) lg[i] += b3[i];
        x = mol_sample_fast(lg, ur);
        if (owns_fold && cta == 0) {
          p.out[(size_t)(p.f0 + fold) * p.out_pitch + t] = x;
          if (p.logits_out) {
#pragma unroll
            for (int i = 0; i < 30; ++i) p.logits_out[((size_t)t * p.n_total + p.f0 + fold) * 30 + i] = lg[i];
          }
        }
      } else {                                              // RAW: this CTA's 4 classes, then the candidate exchange
        mbar_wait(bar_mma, n_mma & 1, p.abort_flag); ++n_mma;
        tc_fence_after();
        float lg[4];
        tmem_ld_sum<4, KW>(tlane + TC_F3, N_F3, lg);
        float best = -INFINITY; int bestk = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          lg[i] += b3[i];
          const float sc = lg[i] - __logf(ex[i]);
          if (sc > best) { best = sc; bestk = u0 + i; }
        }
        unsigned char* cand_out = p.xch5 + (size_t)par * cand_bytes;
        if (owns_fold) {
          *reinterpret_cast<uint2*>(cand_out + ((size_t)cta * n_groups * 8 + fold) * 8) = make_uint2(__float_as_uint(best), (unsigned)bestk);
          if (p.logits_out)
            *reinterpret_cast<float4*>(p.logits_out + ((size_t)t * p.n_total + p.f0 + fold) * NCLS + u0) = make_float4(lg[0], lg[1], lg[2], lg[3]);
        }
        signal(4);
        // every CTA reduces the 128 candidates of each fold the same way; the two half-warps split the CTAs
        mbar_wait(bar_g, (uint32_t)((GPS * t + 4) & 1), p.abort_flag);
        const uint2* cand = reinterpret_cast<const uint2*>(smem + OFF_A);
        const int fr = warp * 16 + (lane & 15), c0 = (lane >> 4) * (P / 2);
        best = -INFINITY; bestk = 0x7fffffff;
        if (fr < n_groups * 8) {
#pragma unroll 8
          for (int c = c0; c < c0 + P / 2; ++c) {
            const uint2 v = cand[c * n_groups * 8 + fr];
            const float sc = __uint_as_float(v.x);
            if (sc > best) { best = sc; bestk = (int)v.y; }        // ascending class order: first maximum wins
          }
        }
        {
          const float ob = __shfl_xor_sync(0xffffffffu, best, 16);
          const int ok = __shfl_xor_sync(0xffffffffu, bestk, 16);
          if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
        }
        x = 2.0f * (float)bestk / ((float)NCLS - 1.0f) - 1.0f;     // :235
        if (owns_fold && cta == 0) p.out[(size_t)(p.f0 + fold) * p.out_pitch + t] = x;
      }
      if (profiling) { tprof[4] += clock64() - tp0; }
      // no early exit on abort: every wait is bounded and fails fast once the abort flag is up, and the
      // named barriers must be executed the same number of times by all participating warps
    }
    if (profiling) for (int i = 0; i < 5; ++i) p.prof[i] = tprof[i];

  } else {
    // =========================================================================================
    // issuer warps 4-7.  Warp q = warp-4 owns the K quarter [128q, 128q+128) of every K=512 chain and
    // accumulates into its own TMEM columns.  Warp 4 ("leader") also watches the arrival counters and
    // launches the TMA gather; warp 5 issues the conditioning chain; all four stage cond_{t+1}.
    // =========================================================================================
    const int q = warp - 4;
    const bool leader = (q == 0);
    const uint32_t sA = smem_u32(smem + OFF_A);
    const uint64_t koff = (uint64_t)(q * (H / 16 / KW) * 16);             // descriptor address-field offset of this K quarter
    const uint64_t dA = umma_desc(sA, 128, SBO_H) + koff, dC = umma_desc(smem_u32(smem + OFF_COND), 128, SBO_Q);
    const uint64_t dS1 = umma_desc(smem_u32(smem + OFF_S1), 128, SBO_H) + koff, dS2 = umma_desc(smem_u32(smem + OFF_S2), 128, SBO_H) + koff,
                   dS3 = umma_desc(smem_u32(smem + OFF_S3), 128, SBO_H) + koff, dF3 = umma_desc(smem_u32(smem + OFF_F3), 128, SBO_H) + koff,
                   dQ = umma_desc(smem_u32(smem + OFF_Q), 128, SBO_Q);
    const uint32_t idesc_s1 = umma_idesc(MT, N_S1, FMT), idesc_s2 = umma_idesc(MT, N_S2, FMT),
                   idesc_s3 = umma_idesc(MT, N_S3, FMT), idesc_f3 = umma_idesc(MT, N_F3, FMT),
                   idesc_q = umma_idesc(MT, N_Q, FMT);
    const bool profiling = (p.prof != nullptr) && cta == 0 && tid == 128;
    long long t_poll = 0, t_gather = 0, t_issue = 0;
    unsigned n_g = 0;

    // D[64 folds, N] (+)= A[64, 16] * B[N, 16]^T per instruction; K advances by two core-matrix columns
    // (256 B => +16 in the descriptor's address field; no carry: every image ends below 256 KB)
    auto launch = [&](int v, unsigned target, const unsigned char* img, uint32_t bytes) {      // leader only
      long long c0 = 0;
      if (profiling) c0 = clock64();
      if (lane == 0) counter_wait(p.counters + v, target, p.abort_flag);        // acquire: all 128 producers have published
      __syncwarp();
      proxy_fence_global();                                                     // generic-proxy writes -> async-proxy (TMA) read
      if (profiling) t_poll += clock64() - c0;
      tma_bulk_g2s(sA, img, bytes, bar_g);
    };
    auto quarter = [&](uint64_t db, uint32_t d_col, uint32_t idesc) {
      long long c0 = 0, c1 = 0;
      if (profiling) c0 = clock64();
      mbar_wait(bar_g, n_g & 1, p.abort_flag); ++n_g;                           // the gathered vector is in smem
      tc_fence_after();
      if (profiling) c1 = clock64();
#pragma unroll
      for (int k = 0; k < H / 16 / KW; ++k) umma_f16(tmem + d_col, dA + (uint64_t)(k * 16), db + (uint64_t)(k * 16), idesc, k > 0);
      umma_commit(bar_mma);
      if (profiling) { const long long c2 = clock64(); t_gather += c1 - c0; t_issue += c2 - c1; }
    };
    auto cond_chain = [&](uint32_t d_col) {                                     // warp 5: pre_{n} = Q cond_n (K = 208)
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < CDIM / 16; ++k) umma_f16(tmem + d_col, dC + (uint64_t)(k * 16), dQ + (uint64_t)(k * 16), idesc_q, k > 0);
      umma_commit(bar_q);
    };

    // ---- conditioning staging (all 128 issuer threads): cond_n rows -> fp16/bf16 A image, one step ahead
    const int st = tid - 128;
    constexpr int COND_TASKS = 5;                          // (fold, 8-column chunk) tasks held in registers per thread
    const int n_tasks = B * KQ;
    // BIG = false (tile of <= 24 folds, host-selected): every task fits the per-thread registers and its loads fly a
    // whole step before use; BIG = true: up to 13 tasks per thread, loaded and stored inside the block below
    constexpr bool deferred = !BIG;
    float4 creg[COND_TASKS][2];
    // conditioning window of fold f of this tile: the strided fold, or the caller's tables (several utterances in one job)
    // strided windows: streams are indexed from the job's first fold, frame tensors from the stream's first sample
    auto row0_of = [&](int f) -> long long {
      return p.fold_row0 ? __ldg(p.fold_row0 + p.f0 + f) : (long long)(p.f0 + f + (FRAMES ? p.seg_first : 0)) * p.seg_stride - p.row_base;
    };
    auto end_of = [&](int f) -> long long { return p.fold_row_end ? __ldg(p.fold_row_end + p.f0 + f) : p.L; };
    auto src_of = [&](int c8, long long row) -> const float4* {
      const float* s = (c8 < FEAT / 8) ? p.mels_up + row * FEAT + c8 * 8 : p.aux + row * (4 * AUXD) + (c8 - FEAT / 8) * 8;
      return reinterpret_cast<const float4*>(s);
    };
    // the 8 conditioning values (fold f, columns 8*c8 .. 8*c8+7) of step n: from the upsampled streams, or built here
    // from frame-rate tensors (UpsampleNetwork's stretch+conv cascade collapsed into 5 taps per output phase)
    auto cond_chunk = [&](int f, int c8, int n, float4& a, float4& b) {
      a = make_float4(0.f, 0.f, 0.f, 0.f); b = a;
      const long long row = row0_of(f) + n;
      if (row >= end_of(f)) return;
      if constexpr (!FRAMES) { const float4* s = src_of(c8, row); a = __ldg(s); b = __ldg(s + 1); return; }
      const unsigned r = (unsigned)row, frame = r / (unsigned)p.hop, phase = r - frame * (unsigned)p.hop;
      if (c8 >= FEAT / 8) {
        const float4* s = reinterpret_cast<const float4*>(p.aux_frames + (size_t)frame * (4 * AUXD) + (c8 - FEAT / 8) * 8);
        a = __ldg(s); b = __ldg(s + 1);
        return;
      }
      const float* k = p.up_taps + phase * 5;
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        const float wgt = __ldg(k + d);
        const float4* s = reinterpret_cast<const float4*>(p.mel_frames + (size_t)(frame + d) * FEAT + c8 * 8);
        const float4 x0 = __ldg(s), x1 = __ldg(s + 1);
        a.x = fmaf(wgt, x0.x, a.x); a.y = fmaf(wgt, x0.y, a.y); a.z = fmaf(wgt, x0.z, a.z); a.w = fmaf(wgt, x0.w, a.w);
        b.x = fmaf(wgt, x1.x, b.x); b.y = fmaf(wgt, x1.y, b.y); b.z = fmaf(wgt, x1.z, b.z); b.w = fmaf(wgt, x1.w, b.w);
      }
    };
    auto store_task = [&](int f, int c8, const float4& a, const float4& b) {
      uint4 v;
      v.x = pack2<FMT>(a.x, a.y); v.y = pack2<FMT>(a.z, a.w); v.z = pack2<FMT>(b.x, b.y); v.w = pack2<FMT>(b.z, b.w);
      *reinterpret_cast<uint4*>(smem + OFF_COND + (f >> 3) * SBO_Q + c8 * 128 + (f & 7) * 16) = v;
    };
    // The (fold, chunk) tasks of a staging thread are the same for the whole launch: hoist everything that does not
    // depend on the step.  In FRAMES mode (frame, phase) of the next row are advanced incrementally -- cond_fetch is
    // called for n = 0, 1, 2, ... in order -- so the step loop has no division.
    int tk_c8[COND_TASKS]; long long tk_left[COND_TASKS];           // rows left in the fold's window: row n exists iff n < left
    const float* tk_src[COND_TASKS]; unsigned tk_frame[COND_TASKS], tk_phase[COND_TASKS];
#pragma unroll
    for (int j = 0; j < COND_TASKS; ++j) {
      const int task = st + j * 128;
      tk_c8[j] = -1; tk_left[j] = 0; tk_src[j] = nullptr; tk_frame[j] = 0; tk_phase[j] = 0;
      if (deferred && task < n_tasks) {
        const int f = task / KQ, c8 = task % KQ;
        const long long r0 = row0_of(f);
        tk_c8[j] = c8; tk_left[j] = end_of(f) - r0;
        if constexpr (FRAMES) { tk_frame[j] = (unsigned)r0 / (unsigned)p.hop; tk_phase[j] = (unsigned)r0 - tk_frame[j] * (unsigned)p.hop; }
        else tk_src[j] = (c8 < FEAT / 8) ? p.mels_up + r0 * FEAT + c8 * 8 : p.aux + r0 * (4 * AUXD) + (c8 - FEAT / 8) * 8;
      }
    }
    auto cond_fetch = [&](int n) {
      if (!deferred) return;
#pragma unroll
      for (int j = 0; j < COND_TASKS; ++j) {
        creg[j][0] = make_float4(0.f, 0.f, 0.f, 0.f); creg[j][1] = creg[j][0];
        const int c8 = tk_c8[j];
        if (c8 >= 0 && n < tk_left[j]) {
          if constexpr (!FRAMES) {
            const float4* s = reinterpret_cast<const float4*>(tk_src[j] + (size_t)n * (c8 < FEAT / 8 ? FEAT : 4 * AUXD));
            creg[j][0] = __ldg(s); creg[j][1] = __ldg(s + 1);
          } else if (c8 >= FEAT / 8) {
            const float4* s = reinterpret_cast<const float4*>(p.aux_frames + (size_t)tk_frame[j] * (4 * AUXD) + (c8 - FEAT / 8) * 8);
            creg[j][0] = __ldg(s); creg[j][1] = __ldg(s + 1);
          } else {
            const float* k = p.up_taps + tk_phase[j] * 5;
            const float* m = p.mel_frames + (size_t)tk_frame[j] * FEAT + c8 * 8;
            float4 a = creg[j][0], b = creg[j][1];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
              const float wgt = __ldg(k + d);
              const float4 x0 = __ldg(reinterpret_cast<const float4*>(m + d * FEAT)), x1 = __ldg(reinterpret_cast<const float4*>(m + d * FEAT) + 1);
              a.x = fmaf(wgt, x0.x, a.x); a.y = fmaf(wgt, x0.y, a.y); a.z = fmaf(wgt, x0.z, a.z); a.w = fmaf(wgt, x0.w, a.w);
              b.x = fmaf(wgt, x1.x, b.x); b.y = fmaf(wgt, x1.y, b.y); b.z = fmaf(wgt, x1.z, b.z); b.w = fmaf(wgt, x1.w, b.w);
            }
            creg[j][0] = a; creg[j][1] = b;
          }
        }
        if constexpr (FRAMES) { if (++tk_phase[j] == (unsigned)p.hop) { tk_phase[j] = 0; ++tk_frame[j]; } }
      }
    };
    auto cond_store = [&](int n) {
      if (deferred) {
#pragma unroll
        for (int j = 0; j < COND_TASKS; ++j) {
          const int task = st + j * 128;
          if (task < n_tasks) store_task(task / KQ, task % KQ, creg[j][0], creg[j][1]);
        }
      } else {
        // more than 24 folds: up to 13 tasks per thread.  All loads are issued before the first use so their
        // latencies overlap (one L2/HBM round trip per step instead of one per task -- this block sits between
        // the S2 and S3 chains of the issuer warps).
        constexpr int MAX_TASKS = (MT * KQ + 127) / 128;
        float4 r[MAX_TASKS][2];
#pragma unroll
        for (int j = 0; j < MAX_TASKS; ++j) {
          const int task = st + j * 128;
          r[j][0] = make_float4(0.f, 0.f, 0.f, 0.f); r[j][1] = r[j][0];
          if (task < n_tasks) cond_chunk(task / KQ, task % KQ, n, r[j][0], r[j][1]);
        }
#pragma unroll
        for (int j = 0; j < MAX_TASKS; ++j) {
          const int task = st + j * 128;
          if (task < n_tasks) store_task(task / KQ, task % KQ, r[j][0], r[j][1]);
        }
      }
      proxy_fence_smem();
    };

    // Conditioning pipeline (round 2): the image of step n is written in the shadow of the y2 exchange of step n-2 and
    // consumed by a 13-MMA chain issued after the S2 chain of step n-1, so that between two exchanges the issuing warps
    // only ever do ONE of {store an image, issue the chain} and neither sits in front of a gather they could already
    // have launched.  (Round 1 did wait + store + barrier + chain + fetch in one block between S2 and S3.)
    cond_fetch(0);
    cond_store(0);
    named_bar_sync(2, 128);
    if (q == 1) cond_chain(TC_Q0);
    if (S > 1) {
      cond_fetch(1);
      mbar_wait(bar_q, 0u, p.abort_flag);                                    // chain 0 has consumed the image
      cond_store(1);
      if (S > 2) cond_fetch(2);
    }

    for (int t = 0; t < S; ++t) {
      const int par = t & 1;
      const unsigned target = (unsigned)P * (unsigned)(t + 1);
      const unsigned char* base = p.xch + (size_t)par * img_bytes;
      if (leader) launch(0, target, base + 0 * xch_stride, img_bytes);
      quarter(dS1, TC_S1 + q * N_S1, idesc_s1);
      if (leader) launch(1, target, base + 1 * xch_stride, img_bytes);
      quarter(dS2, TC_S2 + q * N_S2, idesc_s2);
      if (t + 1 < S) {
        named_bar_sync(2, 128);                                              // every staging thread has stored image t+1
        if (q == 1) cond_chain(par ? TC_Q0 : TC_Q1);                         // pre_{t+1}, queued behind S2 in the tensor pipe
      }
      if (leader) launch(2, target, base + 2 * xch_stride, img_bytes);
      quarter(dS3, TC_S3 + q * N_S3, idesc_s3);
      if (t + 2 < S) {
        // y2 of this step is still two epilogues away: free time for the staging threads
        mbar_wait(bar_q, (uint32_t)((t + 1) & 1), p.abort_flag);             // chain t+1 has consumed the image
        cond_store(t + 2);
        if (t + 3 < S) cond_fetch(t + 3);
      }
      if (leader) launch(3, target, base + 3 * xch_stride, img_bytes);
      quarter(dF3, TC_F3 + q * N_F3, idesc_f3);
      if constexpr (RAW) {
        // fifth exchange: the per-CTA (score, class) candidates land in the (now idle) activation buffer; the fold
        // warps wait on bar_g themselves, the issuers only keep their phase count in step
        if (leader) launch(4, target, p.xch5 + (size_t)par * cand_bytes, cand_bytes);
        mbar_wait(bar_g, n_g & 1, p.abort_flag); ++n_g;
      }
    }
    if (profiling) { p.prof[5] = t_poll; p.prof[6] = t_gather; p.prof[7] = t_issue; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS));
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------

class TcEngine : public Engine {
 public:
  ~TcEngine() override {
    cudaSetDevice(device);
    cudaFree(d_blob_); cudaFree(d_scratch_); cudaFree(d_sync_); cudaFree(d_cond_);
  }
  const char* name() const override { return cfg.precision == WRNN_PREC_BF16 ? "tcgen05-bf16" : "tcgen05-fp16"; }
  int grid_ctas() const override { return P; }
  // FRAMES = conditioning rows built in the kernel from frame-rate tensors (wrnn_job::mel_frames)
  template <int FMT, bool FR, bool BIG>
  const void* kernel_of() const {
    return cfg.mode == WRNN_MODE_RAW ? (const void*)wrnn_tc_kernel<FMT, FR, true, BIG> : (const void*)wrnn_tc_kernel<FMT, FR, false, BIG>;
  }
  // frames: rows formed in the kernel from frame-rate tensors; big: tile of more than SMALL_TILE folds
  const void* kernel(bool frames, bool big) const {
    if (cfg.precision == WRNN_PREC_BF16)
      return frames ? (big ? kernel_of<1, true, true>() : kernel_of<1, true, false>()) : (big ? kernel_of<1, false, true>() : kernel_of<1, false, false>());
    return frames ? (big ? kernel_of<0, true, true>() : kernel_of<0, true, false>()) : (big ? kernel_of<0, false, true>() : kernel_of<0, false, false>());
  }
  static constexpr int SMALL_TILE = 24;     // 24 folds x 26 chunks = 624 <= 5 tasks x 128 staging threads

  int init(const HostWeights& w) {
    Folded f; fold(w, f);
    const bool bf = cfg.precision == WRNN_PREC_BF16, raw = cfg.mode == WRNN_MODE_RAW;
    auto cvt = [&](double v) -> uint16_t { return bf ? f2bf((float)v) : f2h((float)v); };
    std::vector<unsigned char> blob((size_t)WEIGHT_BYTES * P, 0);
    CtaSlice s;
    for (int c = 0; c < P; ++c) {
      slice_for_cta(w, f, c, U, s);
      unsigned char* base = blob.data() + (size_t)c * WEIGHT_BYTES;
      uint16_t* s1 = reinterpret_cast<uint16_t*>(base + (OFF_S1 - OFF_S1));
      uint16_t* s2 = reinterpret_cast<uint16_t*>(base + (OFF_S2 - OFF_S1));
      uint16_t* s3 = reinterpret_cast<uint16_t*>(base + (OFF_S3 - OFF_S1));
      uint16_t* f3 = reinterpret_cast<uint16_t*>(base + (OFF_F3 - OFF_S1));
      uint16_t* q = reinterpret_cast<uint16_t*>(base + (OFF_Q - OFF_S1));
      float* fv = reinterpret_cast<float*>(base + (OFF_VEC - OFF_S1));
      for (int k = 0; k < H; ++k) {
        for (int r = 0; r < 7 * U; ++r) s1[img_index(r, k, H)] = cvt(s.S1[(size_t)r * H + k]);
        for (int r = 0; r < 4 * U; ++r) s2[img_index(r, k, H)] = cvt(s.S2[(size_t)r * H + k]);
        for (int r = 0; r < U; ++r) s3[img_index(r, k, H)] = cvt(s.S3[(size_t)r * H + k]);
        if (raw) { for (int r = 0; r < U; ++r) f3[img_index(r, k, H)] = cvt(w.f3w[(size_t)(c * U + r) * H + k]); }
        else for (int r = 0; r < cfg.n_classes; ++r) f3[img_index(r, k, H)] = cvt(w.f3w[(size_t)r * H + k]);
      }
      for (int k = 0; k < CDIM; ++k)
        for (int r = 0; r < 8 * U; ++r) q[img_index(r, k, CDIM)] = cvt(s.Q[(size_t)r * CDIM + k]);
      for (int r = 0; r < 8 * U; ++r) { fv[r] = s.qk[r]; fv[32 + r] = s.vq[r]; }
      for (int r = 0; r < 3 * U; ++r) { fv[64 + r] = s.b1h[r]; fv[76 + r] = s.b2h[r]; }
      if (raw) { for (int r = 0; r < U; ++r) fv[88 + r] = w.f3b[c * U + r]; }
      else for (int r = 0; r < cfg.n_classes; ++r) fv[88 + r] = w.f3b[r];
    }
    WRNN_CUDA_OK(cudaMalloc(&d_blob_, blob.size()));
    WRNN_CUDA_OK(cudaMemcpy(d_blob_, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    WRNN_CUDA_OK(cudaMalloc(&d_sync_, 256));
    WRNN_CUDA_OK(cudaMemset(d_sync_, 0, 256));
    xch5_off_ = (size_t)4 * 2 * 8 * SBO_H;               // 4 vectors x 2 parities x (up to 8 row groups)
    scratch_bytes_ = xch5_off_ + (size_t)2 * P * MT * 8;   // + RAW candidates: 2 parities x 128 CTAs x 64 folds x 8 B
    WRNN_CUDA_OK(cudaMalloc(&d_scratch_, scratch_bytes_));
    for (int v = 0; v < 4; ++v)
      WRNN_CUDA_OK(cudaFuncSetAttribute(kernel((v & 1) != 0, (v & 2) != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int n_sm = 0;
    WRNN_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    if (n_sm < P) { set_error("tcgen05 engine needs >= 128 SMs for its co-resident weight shards"); return WRNN_E_NO_DEVICE; }
    return WRNN_OK;
  }

  static bool supports_cfg(const wrnn_cfg& c) {
    if (c.precision == WRNN_PREC_FP32) return false;
    return (c.mode == WRNN_MODE_MOL && c.n_classes == 30) || (c.mode == WRNN_MODE_RAW && c.n_classes == 4 * P);
  }
  bool supports(const wrnn_job&) const override { return true; }   // any fold count: tiles of 64 folds, one launch each

  int generate(const wrnn_job& job, cudaStream_t stream) override {
    WRNN_CUDA_OK(cudaSetDevice(device));
    TcParams p{};
    p.blob = static_cast<const unsigned char*>(d_blob_);
    p.mels_up = job.mels_up; p.aux = job.aux; p.L = job.L; p.seg_stride = job.seg_stride;
    p.n_seg = job.n_seg; p.steps = job.steps > 0 ? job.steps : job.seg_len; p.out_pitch = p.steps;
    p.seg_first = job.seg_first;
    p.uniforms = job.uniforms; p.expo = job.expo; p.seed = job.philox_seed; p.offset = job.philox_offset;
    p.uniforms_ready = job.uniforms_ready;
    p.out = job.out; p.x_force = job.x_force; p.logits_out = job.logits_out;
    p.fold_row0 = reinterpret_cast<const long long*>(job.fold_row0); p.fold_row_end = reinterpret_cast<const long long*>(job.fold_row_end);
    p.mel_frames = job.mel_frames; p.aux_frames = job.aux_frames; p.up_taps = job.up_taps; p.hop = job.hop;
    p.xch = static_cast<unsigned char*>(d_scratch_);
    p.xch5 = p.xch + xch5_off_;
    p.counters = static_cast<unsigned*>(d_sync_);
    p.abort_flag = reinterpret_cast<int*>(static_cast<unsigned*>(d_sync_) + 8);
    p.prof = reinterpret_cast<long long*>(static_cast<unsigned char*>(d_sync_) + 64);
    WRNN_CUDA_OK(cudaMemsetAsync(d_sync_, 0, 256, stream));
    p.n_total = job.n_seg;
    // frame-rate conditioning: rows formed by the pre-pass into tile-local scratch (stream kernel), or in the kernel
    const bool frames = job.mel_frames != nullptr;
    const bool expand = frames && (job.cond_mode == WRNN_COND_EXPAND || (job.cond_mode == WRNN_COND_AUTO && !job.fold_row0));
    if (expand && job.fold_row0) { set_error("WRNN_COND_EXPAND needs strided folds (no fold_row0 / fold_row_end tables)"); return WRNN_E_INVALID; }
    if (expand) {
      const int nt = job.n_seg < MT ? job.n_seg : MT;
      long long rows = (long long)(nt - 1) * job.seg_stride + p.steps;
      if (rows > job.L) rows = job.L;
      const size_t need = (size_t)rows * CDIM * sizeof(float);
      if (need > cond_bytes_) {
        cudaFree(d_cond_); d_cond_ = nullptr; cond_bytes_ = 0;      // (cudaFree waits for earlier launches that read it)
        WRNN_CUDA_OK(cudaMalloc(&d_cond_, need));
        cond_bytes_ = need;
      }
    }
    // In this latency-bound regime a step costs the same for 1 or 64 folds, so larger jobs run as consecutive
    // tiles of 64 folds (one persistent launch each) at the full per-tile rate.
    for (int f0 = 0; f0 < job.n_seg; f0 += MT) {
      p.f0 = f0;
      p.n_seg = job.n_seg - f0 < MT ? job.n_seg - f0 : MT;
      if (expand) {
        const long long r_lo = (long long)(job.seg_first + f0) * job.seg_stride;
        long long r_hi = r_lo + (long long)(p.n_seg - 1) * job.seg_stride + p.steps;
        if (r_hi > job.L) r_hi = job.L;
        const long long n_rows = r_hi > r_lo ? r_hi - r_lo : 0;
        float* m_out = static_cast<float*>(d_cond_);
        float* a_out = m_out + (size_t)(cond_bytes_ / (CDIM * sizeof(float))) * FEAT;
        if (n_rows > 0) {
          const long long blocks = (n_rows + XP_ROWS - 1) / XP_ROWS;
          const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
          wrnn_expand_rows_kernel<<<grid, XP_THREADS, 0, stream>>>(job.mel_frames, job.aux_frames, job.up_taps, job.hop, r_lo, n_rows, m_out, a_out);
          WRNN_CUDA_OK(cudaGetLastError());
          ++launches;
        }
        p.mels_up = m_out; p.aux = a_out; p.L = n_rows; p.row_base = (long long)f0 * job.seg_stride;
        p.mel_frames = nullptr;
      }
      WRNN_CUDA_OK(cudaMemsetAsync(d_sync_, 0, 32, stream));                       // arrival counters (the abort flag is sticky)
      void* args[] = {&p};
      WRNN_CUDA_OK(cudaLaunchCooperativeKernel(kernel(frames && !expand, p.n_seg > SMALL_TILE), dim3(P), dim3(NT), args, SMEM_BYTES, stream));
      ++launches;
    }
    last_steps_ = p.steps;
    return WRNN_OK;
  }

  int check() override {
    unsigned char buf[256];
    WRNN_CUDA_OK(cudaSetDevice(device));
    WRNN_CUDA_OK(cudaMemcpy(buf, d_sync_, 256, cudaMemcpyDeviceToHost));
    const int flag = reinterpret_cast<int*>(buf)[8];
    std::memcpy(prof, buf + 64, sizeof(prof));
    if (getenv("WRNN_TC_PROF") && last_steps_ > 0) {     // average cycles per step seen by CTA 0
      const long long n = last_steps_;
      fprintf(stderr, "[wrnn_tc prof] steps=%d | fold thread: A(gru1)=%lld B(h1'->gru2)=%lld C(h2'->y1)=%lld D(y1->y2)=%lld "
              "E(y2->sample)=%lld | driver warp: poll=%lld gather=%lld issue=%lld  (cycles per step)\n",
              last_steps_, prof[0] / n, prof[1] / n, prof[2] / n, prof[3] / n, prof[4] / n, prof[5] / n, prof[6] / n, prof[7] / n);
    }
    if (flag != 0) {
      set_error(flag == 2 ? "persistent kernel aborted: an mbarrier wait (MMA / TMA completion) timed out"
                          : "persistent kernel aborted: an inter-SM exchange wait timed out");
      return WRNN_E_WATCHDOG;
    }
    return WRNN_OK;
  }
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};

 private:
  void *d_blob_ = nullptr, *d_scratch_ = nullptr, *d_sync_ = nullptr;
  void* d_cond_ = nullptr; size_t cond_bytes_ = 0;   // WRNN_COND_EXPAND: one tile's conditioning rows ([rows, 80] then [rows, 128])
  size_t scratch_bytes_ = 0, xch5_off_ = 0;
  int last_steps_ = 0;
};

}  // namespace

// C-ABI entry of the conditioning pre-pass (also used per tile by TcEngine::generate)
int expand_conditioning(const float* mel_frames, const float* aux_frames, const float* up_taps, int hop, long long row_lo,
                        long long n_rows, float* mels_up, float* aux, cudaStream_t stream) {
  const long long blocks = (n_rows + XP_ROWS - 1) / XP_ROWS;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  wrnn_expand_rows_kernel<<<grid, XP_THREADS, 0, stream>>>(mel_frames, aux_frames, up_taps, hop, row_lo, n_rows, mels_up, aux);
  WRNN_CUDA_OK(cudaGetLastError());
  return WRNN_OK;
}

int make_tc_engine(const wrnn_cfg& cfg, const HostWeights& w, int device, Engine** out) {
  if (!TcEngine::supports_cfg(cfg)) {
    set_error("tcgen05 engine serves the MoL head (30 classes) and the 9-bit RAW head (512 classes) with fp16/bf16 operands");
    return WRNN_E_INVALID;
  }
  TcEngine* e = new TcEngine();
  e->cfg = cfg; e->device = device;
  const int rc = e->init(w);
  if (rc != WRNN_OK) { delete e; return rc; }
  *out = e;
  return WRNN_OK;
}

}  // namespace wrnn
