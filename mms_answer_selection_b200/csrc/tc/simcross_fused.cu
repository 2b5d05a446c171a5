// SimCross mode 2 forward as ONE kernel: S[n,k] = (Q_n M_k) A_n^T + B_k with T = Q_n M_k never
// leaving the SM (reference: src/caffe/layers/sim_cross_layer.cpp:146-160, two cblas_sgemm calls per
// (pair, measure) with T in the layer's measure_temp0_ blob).
//
// One tile = P consecutive QA pairs (P*Lq <= 128 token rows) x one measure k:
//   GEMM1  T[128 x N1]  = Qtile[128 x D] * M_k[D x D]     operands by TMA (Q K-major, M_k MN-major),
//                                                          fp32 accumulator T in TMEM columns [0, N1)
//   round  T -> tf32 (round to nearest) in place           tcgen05.ld -> cvt.rna -> tcgen05.st, 32 columns at a
//                                                          time; tcgen05 would otherwise TRUNCATE the operand
//   GEMM2  S[128 x N2]  = T * Atile[N2 x D]^T              A operand read straight from TMEM (tcgen05.mma
//                                                          [tmem], smem-desc), Atile = the P pairs' answer rows;
//                                                          accumulator S in TMEM columns [N1, N1 + N2)
//   store  row (p, lq) keeps columns [p*La, (p+1)*La) of S, + B_k, -> S[n0+p][k][lq][:]
// GEMM2 is block-diagonal (only the P diagonal Lq x La blocks are kept); it costs L/D of GEMM1, so
// computing the off-diagonal blocks is cheaper than a second pass over T through HBM.
//
//   warp 0     TMA producer (one lane), one ring of stages shared by both GEMMs
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-9  T rounding, then the S epilogue (TMEM lane quarter = warp % 4, two warps per quarter
//              taking alternate column chunks)
// The epilogue of tile i overlaps GEMM1 of tile i+1 (S and T occupy different TMEM columns); GEMM2 starts on
// a 32-column chunk of T as soon as that chunk is rounded.
#include <cuda.h>

#include <stdlib.h>

#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"
#include "fused_trace.cuh"

namespace {

using namespace umma;

constexpr int kThreads = 320;          // TMA warp, MMA warp, 2 x 4 rounding / epilogue warps
constexpr int kMaxStages = 6;
constexpr int kMaxChunks = 12;          // 32-column chunks of T (N1 <= 384)

struct FwdGeom {
  int N, Lq, La, D, mc;
  int P;                 // pairs per tile
  int N1, N2;            // accumulator widths of GEMM1 / GEMM2 (multiples of 16)
  int nkb;               // 32-wide k-blocks over D (both GEMMs reduce over D)
  int nboxes;            // 32-wide column boxes of M_k per k-block
  int stages, stage_bytes;
  int kb2;               // k-blocks of the answer tile per ring stage in GEMM2
  unsigned total_tiles;
  uint32_t tmem_cols;
  int vec;               // S / B rows are 16-byte aligned
};

struct FwdSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t t_full, s_full, s_empty;
  uint64_t t_ready[kMaxChunks];
  uint32_t tmem_base;
};

template <bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
simcross2_fwd_fused_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapM,
                           const __grid_constant__ CUtensorMap mapA, const float* __restrict__ Bias,
                           float* __restrict__ S, const FwdGeom g, long long* tr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  FwdSmem* sm = reinterpret_cast<FwdSmem*>(ring + g.stages * g.stage_bytes);

  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  const int stages = g.stages;
  const int nch = (g.N1 + 31) >> 5;
  if (TRACE && threadIdx.x == 0) trace_begin(tr);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
      mbar_init(&sm->t_full, 1); mbar_init(&sm->s_full, 1); mbar_init(&sm->s_empty, 8);
      for (int c = 0; c < kMaxChunks; ++c) mbar_init(&sm->t_ready[c], 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, g.tmem_cols);
    tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapM);
    tma_prefetch_desc(&mapA);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  const uint32_t tmem_S = tmem + (uint32_t)g.N1;
  if (TRACE && threadIdx.x == 0) trace(tr, 1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp walks the loops,
    {                                                            // one elected lane issues)
      const uint32_t tx1 = 16384u + (uint32_t)g.nboxes * 4096u;
      const uint32_t tx2 = (uint32_t)g.N2 * 128u;
      int it = 0;
      for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
        const int k = (int)(t % (unsigned)g.mc);
        const int n0 = (int)(t / (unsigned)g.mc) * g.P;
        for (int b = 0; b < g.nkb; ++b, ++it) {                  // GEMM1: Q k-block + M_k k-block
          const int s = it % stages;
          if (it >= stages) mbar_wait(&sm->empty[s], ((it / stages) - 1) & 1);
          uint8_t* dst = ring + s * g.stage_bytes;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&sm->full[s], tx1);
            tma_load_5d(dst, &mapQ, &sm->full[s], b * 32, n0 * g.Lq, 0, 0, 0);
            for (int x = 0; x < g.nboxes; ++x)
              tma_load_5d(dst + 16384 + x * 4096, &mapM, &sm->full[s], 32 * x, b * 32, 0, k, 0);
          }
          __syncwarp();
        }
        for (int b0 = 0; b0 < g.nkb; b0 += g.kb2, ++it) {        // GEMM2: answer rows of the P pairs, kb2 k-blocks
          const int s = it % stages;                             // per stage (they are small: keep bytes in flight)
          const int nb = min(g.kb2, g.nkb - b0);
          if (it >= stages) mbar_wait(&sm->empty[s], ((it / stages) - 1) & 1);
          uint8_t* dst = ring + s * g.stage_bytes;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&sm->full[s], tx2 * (uint32_t)nb);
            for (int j = 0; j < nb; ++j)
              tma_load_5d(dst + j * tx2, &mapA, &sm->full[s], (b0 + j) * 32, n0 * g.La, 0, 0, 0);
          }
          __syncwarp();
        }
      }
      if (TRACE && lane == 0) trace(tr, 2);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue (whole warp walks the loops,
    {                                                            // one elected lane issues)
      // GEMM1 runs as one MMA per k-step, or two of about equal N (N <= 256 each; the split is a multiple of the
      // 32-column boxes of M_k; measured: 160 + 144 costs 152 cycles per k-step, 256 + 48 costs 171).
      // The blocks below keep the uniform-datapath work per MMA small (one elect per block, descriptors advanced by
      // immediates from a per-stage base, no per-MMA predicates): a lone warp retires a dependent instruction
      // every ~5 cycles, and at 16 instructions per MMA the issue stream was slower than the MMAs.
      const int np0 = g.N1 <= 256 ? g.N1 : ((g.N1 / 2 + 31) & ~31), np1 = g.N1 - np0;
      const bool two = np1 > 0;
      const uint32_t idesc_p0 = idesc_tf32(128, np0, false, true);
      const uint32_t idesc_p1 = idesc_tf32(128, two ? np1 : 16, false, true);
      const uint32_t idesc_2 = idesc_tf32(128, g.N2, false, false);
      const uint32_t ring_base = smem_u32(ring);
      const uint32_t ring_lo_a = desc_lo_k(ring_base), ring_lo_b = desc_lo_mn(ring_base + 16384, 4096);
      const uint32_t ring_lo_2 = desc_lo_k(ring_base);
      const uint32_t stage_lo = (uint32_t)g.stage_bytes >> 4;
      const uint32_t b1_off = (uint32_t)(np0 >> 5) * (4096u >> 4);
      const uint32_t tmem_T1 = tmem + (uint32_t)np0;
      const uint32_t n2_lo = ((uint32_t)g.N2 * 128u) >> 4;
      const int nkb = g.nkb, D = g.D, kb2 = g.kb2;
      int s = 0; uint32_t ph = 0;
      int tc = 0;
      for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x, ++tc) {
        for (int b = 0; b < nkb; ++b) {
          mbar_wait(&sm->full[s], ph);
          tc_fence_after();
          if (TRACE && tc == 0 && b == 0 && lane == 0) trace(tr, 3);
          const uint32_t a_lo = ring_lo_a + (uint32_t)s * stage_lo, b_lo = ring_lo_b + (uint32_t)s * stage_lo;
          const int left = D - b * 32;
          const int nks = left >= 32 ? 4 : (left + 7) >> 3;
          const uint32_t acc0 = b > 0 ? 1u : 0u;
          if (elect_one_sync()) {
            if (two) {
              const uint32_t b1_lo = b_lo + b1_off;
              mma_tf32_ss_lh(tmem, a_lo, kDescHiK, b_lo, kDescHiMN, idesc_p0, acc0);
              mma_tf32_ss_lh(tmem_T1, a_lo, kDescHiK, b1_lo, kDescHiMN, idesc_p1, acc0);
              if (nks > 1) {
                mma_tf32_ss_lh(tmem, a_lo + 1 * kDescStepK, kDescHiK, b_lo + 1 * kDescStepMN, kDescHiMN, idesc_p0, 1u);
                mma_tf32_ss_lh(tmem_T1, a_lo + 1 * kDescStepK, kDescHiK, b1_lo + 1 * kDescStepMN, kDescHiMN, idesc_p1, 1u);
              }
              if (nks > 2) {
                mma_tf32_ss_lh(tmem, a_lo + 2 * kDescStepK, kDescHiK, b_lo + 2 * kDescStepMN, kDescHiMN, idesc_p0, 1u);
                mma_tf32_ss_lh(tmem_T1, a_lo + 2 * kDescStepK, kDescHiK, b1_lo + 2 * kDescStepMN, kDescHiMN, idesc_p1, 1u);
              }
              if (nks > 3) {
                mma_tf32_ss_lh(tmem, a_lo + 3 * kDescStepK, kDescHiK, b_lo + 3 * kDescStepMN, kDescHiMN, idesc_p0, 1u);
                mma_tf32_ss_lh(tmem_T1, a_lo + 3 * kDescStepK, kDescHiK, b1_lo + 3 * kDescStepMN, kDescHiMN, idesc_p1, 1u);
              }
            } else {
              mma_tf32_ss_lh(tmem, a_lo, kDescHiK, b_lo, kDescHiMN, idesc_p0, acc0);
              if (nks > 1) mma_tf32_ss_lh(tmem, a_lo + 1 * kDescStepK, kDescHiK, b_lo + 1 * kDescStepMN, kDescHiMN, idesc_p0, 1u);
              if (nks > 2) mma_tf32_ss_lh(tmem, a_lo + 2 * kDescStepK, kDescHiK, b_lo + 2 * kDescStepMN, kDescHiMN, idesc_p0, 1u);
              if (nks > 3) mma_tf32_ss_lh(tmem, a_lo + 3 * kDescStepK, kDescHiK, b_lo + 3 * kDescStepMN, kDescHiMN, idesc_p0, 1u);
            }
            mma_commit(&sm->empty[s]);
            if (b == nkb - 1) mma_commit(&sm->t_full);
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        if (TRACE && tc == 0 && lane == 0) trace(tr, 4);
        if (tc > 0) mbar_wait(&sm->s_empty, (uint32_t)(tc - 1) & 1u);       // the previous tile's S has been read out
        for (int b0 = 0; b0 < nkb; b0 += kb2) {
          const int nb = min(kb2, nkb - b0);
          mbar_wait(&sm->full[s], ph);
          const uint32_t base2 = ring_lo_2 + (uint32_t)s * stage_lo;
          for (int j = 0; j < nb; ++j) {
            const int b = b0 + j;
            mbar_wait(&sm->t_ready[b], (uint32_t)tc & 1u);                  // T columns [32 b, 32 b + 32) are rounded
            tc_fence_after();
            const uint32_t b_lo = base2 + (uint32_t)j * n2_lo;
            const uint32_t a_t = tmem + (uint32_t)(b * 32);
            const int left = D - b * 32;
            const int nks = left >= 32 ? 4 : (left + 7) >> 3;
            if (elect_one_sync()) {
              mma_tf32_ts_lh(tmem_S, a_t, b_lo, kDescHiK, idesc_2, b > 0 ? 1u : 0u);
              if (nks > 1) mma_tf32_ts_lh(tmem_S, a_t + 8, b_lo + 1 * kDescStepK, kDescHiK, idesc_2, 1u);
              if (nks > 2) mma_tf32_ts_lh(tmem_S, a_t + 16, b_lo + 2 * kDescStepK, kDescHiK, idesc_2, 1u);
              if (nks > 3) mma_tf32_ts_lh(tmem_S, a_t + 24, b_lo + 3 * kDescStepK, kDescHiK, idesc_2, 1u);
              if (j == nb - 1) mma_commit(&sm->empty[s]);
              if (b == nkb - 1) mma_commit(&sm->s_full);
            }
            __syncwarp();
          }
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        if (TRACE && tc == 0 && lane == 0) trace(tr, 7);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ T rounding + S epilogue
    const int quarter = warp & 3;
    const int set = (warp - 2) >> 2;                             // two warps per TMEM lane quarter share the columns
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    const int row = quarter * 32 + lane;                         // token row of the tile = TMEM lane
    const int p_lane = row / g.Lq, lq = row - p_lane * g.Lq;
    const int p_lo = (quarter * 32) / g.Lq;
    const int p_hi = min(g.P - 1, (quarter * 32 + 31) / g.Lq);
    const int nc8 = (g.La + 7) >> 3;
    int tc = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x, ++tc) {
      const int k = (int)(t % (unsigned)g.mc);
      const int n0 = (int)(t / (unsigned)g.mc) * g.P;
      mbar_wait(&sm->t_full, tc & 1);
      tc_fence_after();
      if (TRACE && tc == 0 && threadIdx.x == 64) trace(tr, 5);
      for (int c = set; c < nch; c += 2) {
        float v[32];
        const uint32_t ta = tmem + lane_bits + (uint32_t)(c * 32);
        if (c * 32 + 16 < g.N1) {
          tmem_ld32(ta, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = to_tf32(v[i]);
          tmem_st32(ta, v);
        } else {
          tmem_ld16(ta, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = to_tf32(v[i]);
          tmem_st16(ta, v);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->t_ready[c]);
      }
      if (TRACE && tc == 0 && threadIdx.x == 64) trace(tr, 6);
      mbar_wait(&sm->s_full, tc & 1);
      tc_fence_after();
      if (TRACE && tc == 0 && threadIdx.x == 64) trace(tr, 8);
      for (int p = p_lo; p <= p_hi; ++p) {
        const int n = n0 + p;
        const bool active = (p_lane == p) && (n < g.N);
        float* srow = S + (((size_t)n * g.mc + k) * g.Lq + lq) * g.La;
        const float* brow = Bias ? Bias + ((size_t)k * g.Lq + lq) * g.La : nullptr;
        for (int c8 = set; c8 < nc8; c8 += 2) {
          float v[8];
          tmem_ld8(tmem_S + lane_bits + (uint32_t)(p * g.La + c8 * 8), v);
          if (active) {
            const int col = c8 * 8;
            if (g.vec && col + 8 <= g.La) {
              if (brow) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(brow + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(brow + col + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              *reinterpret_cast<float4*>(srow + col) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(srow + col + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (col + j < g.La) srow[col + j] = v[j] + (brow ? __ldg(brow + col + j) : 0.f);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->s_empty);
      if (TRACE && tc == 0 && threadIdx.x == 64) trace(tr, 9);
    }
    if (TRACE && threadIdx.x == 64) trace(tr, 10);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, g.tmem_cols);
  if (TRACE && threadIdx.x == 0) trace_end(tr);
}

}  // namespace

// qr (N*Lq x Dp), ar (N*La x Dp), Mr (mc x D x Dp): TF32-rounded copies with 16-byte-aligned rows.
// Returns MMS_E_UNSUPPORTED for shapes the fused tile does not cover (the caller composes GEMMs instead).
int mms_tc_simcross2_forward_fused(mms_context* ctx, const float* qr, const float* ar, const float* Mr,
                                   const float* B, float* S, int N, int Lq, int La, int D, int mc, int Dp) {
  static const bool disabled = mms_dev_knob("MMS_NO_FUSED");
  if (disabled) return MMS_E_UNSUPPORTED;
  if (Lq > 128 || La > 256) return MMS_E_UNSUPPORTED;
  FwdGeom g;
  g.N = N; g.Lq = Lq; g.La = La; g.D = D; g.mc = mc;
  g.P = mms_max(1, mms_min(mms_min(128 / Lq, 256 / La), N));
  g.N1 = mms_ceil_div(D, 16) * 16;
  // the epilogue reads S in 8-column groups per pair: the last group may overhang the pair's La columns
  g.N2 = mms_ceil_div((g.P - 1) * La + mms_ceil_div(La, 8) * 8, 16) * 16;
  if (g.N1 > 32 * kMaxChunks || g.N2 > 256 || g.N1 + g.N2 > 512) return MMS_E_UNSUPPORTED;
  g.nkb = mms_ceil_div(D, 32);
  g.nboxes = mms_ceil_div(g.N1, 32);
  g.stage_bytes = mms_ceil_div(mms_max(16384 + g.nboxes * 4096, g.N2 * 128), 1024) * 1024;
  g.kb2 = mms_max(1, mms_min(4, g.stage_bytes / (g.N2 * 128)));
  int stages = kMaxStages;
  while (stages > 2 && (size_t)stages * g.stage_bytes + sizeof(FwdSmem) + 1024 > 226 * 1024) --stages;
  if ((size_t)stages * g.stage_bytes + sizeof(FwdSmem) + 1024 > 226 * 1024) return MMS_E_UNSUPPORTED;
  g.stages = stages;
  const long long total = (long long)mms_ceil_div(N, g.P) * mc;
  if (total > 0x7fffffffLL) return MMS_E_UNSUPPORTED;
  g.total_tiles = (unsigned)total;
  g.tmem_cols = umma::tmem_cols_pow2((uint32_t)(g.N1 + g.N2));
  g.vec = (La % 4 == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0) &&
          (!B || (reinterpret_cast<uintptr_t>(B) & 15) == 0);

  CUtensorMap mapQ, mapM, mapA;
  MMS_TRY(mms_tc_make_map(ctx, &mapQ, qr, Dp, false, (long long)N * Lq, D, 128, 0, 0, 0, 1, 1, 1));
  MMS_TRY(mms_tc_make_map(ctx, &mapM, Mr, Dp, true, D, D, 32, (long long)D * Dp, 0, 0, mc, 1, 1));
  MMS_TRY(mms_tc_make_map(ctx, &mapA, ar, Dp, false, (long long)N * La, D, g.N2, 0, 0, 0, 1, 1, 1));

  static bool configured = false;
  if (!configured) {
    MMS_MAX_SMEM(simcross2_fwd_fused_kernel<false>, 227 * 1024);
    MMS_MAX_SMEM(simcross2_fwd_fused_kernel<true>, 227 * 1024);
    configured = true;
  }
  const size_t smem = (size_t)stages * g.stage_bytes + sizeof(FwdSmem) + 1024;
  const unsigned grid = (unsigned)mms_min<long long>(total, ctx->sm_count);
  TraceBuf tb;
  MMS_TRY(tb.begin(grid));
  { MmsKernelScope ks_(ctx, "simcross2_fwd_fused_kernel");
    if (tb.dev) simcross2_fwd_fused_kernel<true><<<grid, kThreads, smem, ctx->stream>>>(mapQ, mapM, mapA, B, S, g, tb.dev);
    else simcross2_fwd_fused_kernel<false><<<grid, kThreads, smem, ctx->stream>>>(mapQ, mapM, mapA, B, S, g, nullptr); }
  MMS_LAUNCH_CHECK();
  static const char* const names[kTraceSlots] = {"entry", "setup", "tma_issued", "first_full", "g1_issued", "t_full",
                                                 "rounded", "g2_issued", "s_full", "epi0_done", "epi_done", "exit",
                                                 nullptr, nullptr, nullptr, nullptr,
                                                 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  char what[96];
  snprintf(what, sizeof(what), "fwd N %d L %dx%d D %d mc %d P %d tiles %u stages %d", N, Lq, La, D, mc, g.P,
           g.total_tiles, stages);
  MMS_TRY(tb.end(ctx, what, names));
  return 0;
}
