// SimCross mode 2 backward, weight gradient:  dM_k += Qall^T U_k   (reference:
// src/caffe/layers/sim_cross_layer.cpp:286-289, dM_k += (Q_n^T dS[n,k]) A_n per pair, with U_k = dS[.,k] A the
// intermediate the fused dQ kernel exports).  A reduction over ALL token rows into a D x D matrix per measure:
// tall-skinny operands, tiny output.
//
// One CTA = (measure k, column part h of dM_k, slice of the token rows).  It keeps ALL row blocks of its column
// part as accumulators in TMEM (ceil(D / 128) blocks x Nh0 columns: 3 x 160 at D = 300) and streams 32 token rows
// per ring stage:
//     A  = 32 rows of Q  (MN-major boxes of 32 d x 32 rows)
//     B  = 32 rows of U_k, columns of part h, K-major, from the BLOCKED layout the dQ kernel writes (below)
//     per 8-row k-step: one MMA per row block, all sharing the B tile
// so a stage of 60 KB feeds 12 MMAs of Nh0 columns (~62 B per tensor-pipe cycle; the generic 128 x 160 tile needs
// 112 B per cycle and ran at the TMA rate instead, 0.32 ms at C3).  The slices add into dM with vector atomics.
//
// Blocked U layout (float index): token rows in groups of 32, each group stored TRANSPOSED:
//     U[k][row][col]  at  ((k * n_groups + row / 32) * Dp + col) * 32 + row % 32
// A warp of the dQ kernel owns 32 consecutive rows (lane = row): one 4-byte store per column then writes 128
// consecutive bytes (one or two LSU wavefronts; the row-major export wrote 32 bytes per lane into 32 different
// lines, 32 wavefronts per instruction, ~1000 cycles per 64-column chunk), and for this kernel the group is a
// K-major B tile as it stands: [column][32 token rows] with 128-byte rows, one plain SWIZZLE_128B box per stage.
// (A layout with 32-byte units, {8 floats, 4 units, r rows} boxes, was tried first: TMA faults on a swizzled box
// whose inner extent is not the 128-byte swizzle span; tools/tma_blocked_test.cu.)
//
//   warp 0  TMA producer    warp 1  MMA issuer + TMEM allocator    warps 2-5  epilogue (one per TMEM lane quarter)
#include <cuda.h>

#include <stdlib.h>

#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kThreads = 6 * 32;
constexpr int kMaxStages = 8;
constexpr int kRows = 32;               // token rows per stage

struct DmGeom {
  int D, mc;
  int nmb;               // 128-row blocks of dM rows (d index)
  int nh, Nh0, N1;       // column parts of dM
  int slices;            // CTAs that share one (k, h)
  int kbt;               // 16-row blocks over all token rows
  int stages, stage_bytes, a_bytes;
  uint32_t tmem_cols;
  int vec;               // 16-byte atomics
  int blocked;           // U layout: 1 blocked (above), 0 row-major (pitch Dp)
  int a_boxes;           // 32-column boxes of Q per stage (the last row block may run past them, see below)
  int b_bytes;
};

struct DmSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
simcross2_dm_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapU,
                    float* __restrict__ dM, const DmGeom g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  DmSmem* sm = reinterpret_cast<DmSmem*>(ring + g.stages * g.stage_bytes);
  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;

  const int slice = (int)(blockIdx.x % (unsigned)g.slices);
  const int h = (int)((blockIdx.x / (unsigned)g.slices) % (unsigned)g.nh);
  const int k = (int)(blockIdx.x / (unsigned)(g.slices * g.nh));
  const int Nh = min(g.Nh0, g.N1 - h * g.Nh0);
  const int kb_lo = (int)((long long)slice * g.kbt / g.slices), kb_hi = (int)((long long)(slice + 1) * g.kbt / g.slices);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
      mbar_init(&sm->acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, g.tmem_cols);
    tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapU);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const int a_boxes = g.a_boxes, b_boxes = (Nh + 31) >> 5;
    const uint32_t tx = (uint32_t)a_boxes * 4096u + (g.blocked ? (uint32_t)g.Nh0 * 128u : (uint32_t)b_boxes * 4096u);
    int s = 0; uint32_t ph = 0;
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      mbar_wait(&sm->empty[s], ph ^ 1u);
      uint8_t* dst = ring + s * g.stage_bytes;
      const int row0 = kb * kRows;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&sm->full[s], tx);
        for (int x = 0; x < a_boxes; ++x) tma_load_5d(dst + x * 4096, &mapQ, &sm->full[s], 32 * x, row0, 0, 0, 0);
        if (g.blocked) {
          tma_load_5d(dst + g.a_bytes, &mapU, &sm->full[s], 0, h * g.Nh0, kb, k, 0);     // [Nh0 columns][32 rows]
        } else {
          for (int x = 0; x < b_boxes; ++x)
            tma_load_5d(dst + g.a_bytes + x * 4096, &mapU, &sm->full[s], h * g.Nh0 + 32 * x, row0, 0, k, 0);
        }
      }
      __syncwarp();
      if (++s == g.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue
    // Row block j reads the Q boxes 4j .. 4j+3.  Only the boxes that hold columns < D are loaded; the last block's
    // descriptor runs on into the B region of the stage: finite values that only reach accumulator rows d >= D,
    // which the epilogue never stores.
    const bool blk = g.blocked != 0;
    const uint32_t idesc = idesc_tf32(128, Nh, true, !blk);
    const uint32_t ring_a = desc_lo_mn(smem_u32(ring), 4096);
    const uint32_t ring_b = blk ? desc_lo_k(smem_u32(ring) + (uint32_t)g.a_bytes)
                                : desc_lo_mn(smem_u32(ring) + (uint32_t)g.a_bytes, 4096);
    const uint32_t b_hi = blk ? kDescHiK : kDescHiMN, b_step = blk ? kDescStepK : kDescStepMN;
    const uint32_t stage_lo = (uint32_t)g.stage_bytes >> 4;
    const int nmb = g.nmb;
    const uint32_t acc_step = (uint32_t)g.Nh0;
    int s = 0; uint32_t ph = 0;
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      mbar_wait(&sm->full[s], ph);
      tc_fence_after();
      const uint32_t a_lo = ring_a + (uint32_t)s * stage_lo, b_lo = ring_b + (uint32_t)s * stage_lo;
      const uint32_t acc = kb > kb_lo ? 1u : 0u;
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ac = ks == 0 ? acc : 1u;
          mma_tf32_ss_lh(tmem, a_lo + ks * kDescStepMN, kDescHiMN, b_lo + ks * b_step, b_hi, idesc, ac);
          if (nmb > 1)
            mma_tf32_ss_lh(tmem + acc_step, a_lo + 4 * (4096u >> 4) + ks * kDescStepMN, kDescHiMN, b_lo + ks * b_step, b_hi,
                           idesc, ac);
          if (nmb > 2)
            mma_tf32_ss_lh(tmem + 2 * acc_step, a_lo + 8 * (4096u >> 4) + ks * kDescStepMN, kDescHiMN, b_lo + ks * b_step,
                           b_hi, idesc, ac);
        }
        mma_commit(&sm->empty[s]);
        if (kb == kb_hi - 1) mma_commit(&sm->acc_full);
      }
      __syncwarp();
      if (++s == g.stages) { s = 0; ph ^= 1u; }
    }
  } else if (kb_hi > kb_lo) {
    // ------------------------------------------------------------ epilogue: slices add into dM_k
    const int quarter = warp & 3;
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    mbar_wait(&sm->acc_full, 0);
    tc_fence_after();
    for (int j = 0; j < g.nmb; ++j) {
      const int d = j * 128 + quarter * 32 + lane;               // row of dM_k
      float* drow = dM + ((size_t)k * g.D + (d < g.D ? d : 0)) * g.D + h * g.Nh0;
      for (int c = 0; c * 32 < Nh; ++c) {
        float v[32];
        const uint32_t ta = tmem + lane_bits + (uint32_t)(j * g.Nh0 + c * 32);
        const int wc = (Nh - c * 32 > 16) ? 32 : 16;
        if (wc == 32) tmem_ld32(ta, v); else tmem_ld16(ta, v);
        if (d >= g.D) continue;
        const int ncols = mms_min(wc, g.D - h * g.Nh0 - c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i * 4 < ncols) {
            if (g.vec && i * 4 + 4 <= ncols) {
              atomicAdd(reinterpret_cast<float4*>(drow + c * 32 + 4 * i),
                        make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (i * 4 + e < ncols) atomicAdd(drow + c * 32 + 4 * i + e, v[4 * i + e]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, g.tmem_cols);
}

bool dm_shape_ok(int D, int* nmb, int* nh, int* Nh0, int* N1) {
  *N1 = mms_ceil_div(D, 16) * 16;
  *nmb = mms_ceil_div(D, 128);
  *nh = *N1 <= 256 ? 1 : 2;
  *Nh0 = *nh == 1 ? *N1 : ((*N1 / 2 + 31) & ~31);
  return *nmb * *Nh0 <= 512 && *nmb <= 3 && *Nh0 <= 256;
}

}  // namespace

// Whether dM can be computed by this kernel from the blocked U layout (the dQ kernel then exports that layout).
int mms_tc_simcross2_dm_plan(int D) {
  static const bool disabled = mms_dev_knob("MMS_NO_FUSED") || mms_dev_knob("MMS_NO_FUSED_DM");
  int nmb, nh, Nh0, N1;
  return (!disabled && dm_shape_ok(D, &nmb, &nh, &Nh0, &N1)) ? 0 : MMS_E_UNSUPPORTED;
}

// dM (mc x D x D) += Qall^T U_k;  qr (rows x Dp) TF32-rounded questions, Ub the blocked export of the dQ kernel
// (mc slabs of ceil(rows / 32) * 32 * Dp floats; the rows past `rows` of the last group are zero).
int mms_tc_simcross2_dm(mms_context* ctx, const float* qr, const float* Ub, float* dM, long long rows, int D, int Dp,
                        int mc, int blocked) {
  DmGeom g;
  MMS_REQUIRE(dm_shape_ok(D, &g.nmb, &g.nh, &g.Nh0, &g.N1), MMS_E_UNSUPPORTED, "shape not covered");
  MMS_REQUIRE(Dp % 32 == 0 && rows > 0, MMS_E_INVALID, "bad layout");
  g.D = D; g.mc = mc;
  const long long n_groups = (rows + 31) / 32;
  g.kbt = (int)((rows + kRows - 1) / kRows);
  MMS_REQUIRE(n_groups <= 0x7fffffffLL, MMS_E_UNSUPPORTED, "too many rows");
  // one wave of CTAs; a slice should stream at least ~8 stages, or its fixed costs (pipeline fill, the atomic
  // epilogue over ceil(D / 128) x 128 x Nh0 accumulators) outweigh what the extra CTAs save
  g.slices = (int)mms_max<long long>(1, mms_min<long long>(ctx->sm_count / (mc * g.nh), g.kbt / 8));
  g.blocked = blocked;
  g.a_boxes = mms_ceil_div(D, 32);
  g.a_bytes = g.a_boxes * 4096;
  g.b_bytes = blocked ? g.Nh0 * 128 : mms_ceil_div(g.Nh0, 32) * 4096;
  g.stage_bytes = mms_ceil_div(mms_max(g.a_bytes + g.b_bytes, 4 * g.nmb * 4096), 1024) * 1024;
  int stages = kMaxStages;
  while (stages > 2 && (size_t)stages * g.stage_bytes + sizeof(DmSmem) + 1024 > 227 * 1024) --stages;
  g.stages = stages;
  g.tmem_cols = umma::tmem_cols_pow2((uint32_t)(g.nmb * g.Nh0));
  g.vec = (D % 4 == 0) && (g.Nh0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(dM) & 15) == 0);

  CUtensorMap mapQ, mapU;
  MMS_TRY(mms_tc_make_map(ctx, &mapQ, qr, Dp, true, D, rows, 32, 0, 0, 0, 1, 1, 1, kRows));
  if (!blocked) {
    MMS_TRY(mms_tc_make_map(ctx, &mapU, Ub, Dp, true, D, rows, 32, rows * Dp, 0, 0, mc, 1, 1, kRows));
  } else {   // K-major operand [column][32 rows]: row pitch 32 floats, batch dimensions = 32-row group, measure
    MMS_TRY(mms_tc_make_map(ctx, &mapU, Ub, 32, false, D, 32, g.Nh0, n_groups * (long long)Dp * 32, (long long)Dp * 32, 0,
                            mc, (int)n_groups, 1));
  }
  static bool configured = false;
  if (!configured) {
    MMS_MAX_SMEM(simcross2_dm_kernel, 227 * 1024);
    configured = true;
  }
  const size_t smem = (size_t)stages * g.stage_bytes + sizeof(DmSmem) + 1024;
  const unsigned grid = (unsigned)(mc * g.nh * g.slices);
  { MmsKernelScope ks_(ctx, "simcross2_dm_kernel");
    simcross2_dm_kernel<<<grid, kThreads, smem, ctx->stream>>>(mapQ, mapU, dM, g); }
  MMS_LAUNCH_CHECK();
  return 0;
}
