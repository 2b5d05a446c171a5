// SimCross mode 2 backward, bottom gradients, as fused tcgen05 kernels (reference:
// src/caffe/layers/sim_cross_layer.cpp:282-299, six cblas_sgemm calls per (pair, measure)):
//
//   dQ_n = sum_k (G_nk A_n) M_k^T          G_nk = dS[n,k]  (Lq x La)
//   dA_n = sum_k (G_nk^T Q_n) M_k          (= sum_k G_nk^T (Q_n M_k) of the reference, re-associated so that both
//                                           gradients have the same shape: small product first, D x D product second)
//
// One kernel template serves both (DA = false: dQ, DA = true: dA).  One tile = P consecutive QA pairs (their
// P*Lr output rows, Lr = Lq or La) x one column part h of the output x a range of measures:
//   for k in the tile's measures:
//     build  Gblk [128 x 128]: block-diagonal, block p = G_{n0+p,k} (dQ) or its transpose (dA), rounded to TF32,
//            written into shared memory in the K-major UMMA layout by four builder warps straight from dS
//     GEMM-A U [128 x N1] = Gblk * X      X = the P pairs' answer (dQ) / question (dA) rows, MN-major by TMA;
//                                         accumulator U in TMEM columns [0, N1)
//     round  U -> tf32 in place (tcgen05.ld / cvt.rna / tcgen05.st); the dQ kernel also writes U to global memory,
//            where the dM contraction (dM_k = Q^T U_k, a reduction over ALL pairs) picks it up
//     GEMM-B O [128 x Nh] += U * W_k      W_k = columns [h*Nh0, ..) of M_k^T (dQ, K-major boxes) or M_k (dA, MN-major
//                                         boxes); A operand read from TMEM; O accumulates over k in TMEM
//   store  O -> dq / da rows (plain stores; red.global.add when the measures of one tile are split over CTAs)
// The output is split in column parts because U (N1 columns) and O share the 512 TMEM columns: D = 300 gives
// N1 = 304 and two parts of 160 + 144 columns, with U recomputed for each part.
//
//   warp 0       TMA producer        warp 1       tcgen05.mma issuer + TMEM allocator
//   warps 2-9    rounding of U, U export, output epilogue (two warps per TMEM lane quarter)
//   warps 10-13  Gblk builders
#include <cuda.h>

#include <stdlib.h>

#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"
#include "fused_trace.cuh"

namespace {

using namespace umma;

constexpr int kThreads = 14 * 32;
constexpr int kMaxStages = 5;
constexpr int kMaxChunks = 12;          // 32-column chunks of U (N1 <= 384)
constexpr int kGblkBytes = 4 * 16384;   // 128 rows x 128 contraction columns, four 32-wide k-blocks

struct BwdGeom {
  int N, Lq, La, D, mc;
  int Lr, Lk;            // output-row side / contraction side sentence length
  int P;                 // pairs per tile
  int N1, np0;           // width of U; GEMM-A runs as MMAs of np0 and N1 - np0 columns
  int nh, Nh0;           // column parts of the output: part h = [h*Nh0, min(N1, (h+1)*Nh0))
  int nka, nkb;          // 32-wide k-blocks of GEMM-A (over P*Lk) / GEMM-B (over D)
  int nboxes;            // 32-column boxes of X per GEMM-A stage
  int kb2, b2_bytes;     // GEMM-B: k-blocks per ring slot, bytes per k-block
  int stages, stage_bytes;
  int ksplit;            // CTAs that share the measures of one (group, part)
  unsigned total_tiles;
  uint32_t tmem_cols;
  int vec_g, vec_out;
  int Dp;                // row pitch of the exported U
  long long u_rows;      // rows of one measure slab of the exported U
};

struct BwdSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t g_full, g_empty, u_full, o_full, o_empty;
  uint64_t t_ready[kMaxChunks];
  uint32_t tmem_base;
};

// one full 32-byte sector per thread and instruction (row-per-thread stores of the U export)
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

template <bool DA>
__global__ void __launch_bounds__(kThreads, 1)
simcross2_bwd_fused_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapM,
                           const float* __restrict__ dS, float* __restrict__ out, float* __restrict__ Uexp,
                           const BwdGeom g, long long* tr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* gblk = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = gblk + kGblkBytes;
  BwdSmem* sm = reinterpret_cast<BwdSmem*>(ring + g.stages * g.stage_bytes);

  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  const int stages = g.stages;
  const int nch = (g.N1 + 31) >> 5;
  const int PLk = g.P * g.Lk;
  if (threadIdx.x == 0) trace_begin(tr);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
      mbar_init(&sm->g_full, 4); mbar_init(&sm->g_empty, 1); mbar_init(&sm->u_full, 1);
      mbar_init(&sm->o_full, 1); mbar_init(&sm->o_empty, 8);
      for (int c = 0; c < kMaxChunks; ++c) mbar_init(&sm->t_ready[c], 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, g.tmem_cols);
    tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapM);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  const uint32_t tmem_O = tmem + (uint32_t)g.N1;
  if (threadIdx.x == 0) trace(tr, 1);

  // tile -> (pair group, column part, measure range)
  auto decode = [&](unsigned t, int& n0, int& h, int& k_lo, int& k_hi) {
    const int ksp = (int)(t % (unsigned)g.ksplit); t /= (unsigned)g.ksplit;
    h = (int)(t % (unsigned)g.nh);
    n0 = (int)(t / (unsigned)g.nh) * g.P;
    k_lo = ksp * g.mc / g.ksplit;
    k_hi = (ksp + 1) * g.mc / g.ksplit;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const uint32_t txa = (uint32_t)g.nboxes * 4096u;
    const int nbx2 = g.b2_bytes >> 12;                           // dA: 32-column boxes of M_k per k-block
    int it = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      int n0, h, k_lo, k_hi;
      decode(t, n0, h, k_lo, k_hi);
      for (int k = k_lo; k < k_hi; ++k) {
        for (int kb = 0; kb < g.nka; ++kb, ++it) {               // GEMM-A: 32 contraction rows of X, all N1 columns
          const int s = it % stages;
          if (it >= stages) mbar_wait(&sm->empty[s], ((it / stages) - 1) & 1);
          uint8_t* dst = ring + s * g.stage_bytes;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&sm->full[s], txa);
            for (int x = 0; x < g.nboxes; ++x)
              tma_load_5d(dst + x * 4096, &mapX, &sm->full[s], 32 * x, n0 * g.Lk + 32 * kb, 0, 0, 0);
          }
          __syncwarp();
        }
        for (int b0 = 0; b0 < g.nkb; b0 += g.kb2, ++it) {        // GEMM-B: kb2 k-blocks of W_k per slot
          const int s = it % stages;
          const int nb = min(g.kb2, g.nkb - b0);
          if (it >= stages) mbar_wait(&sm->empty[s], ((it / stages) - 1) & 1);
          uint8_t* dst = ring + s * g.stage_bytes;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&sm->full[s], (uint32_t)(nb * g.b2_bytes));
            for (int j = 0; j < nb; ++j) {
              if (!DA) {
                tma_load_5d(dst + j * g.b2_bytes, &mapM, &sm->full[s], (b0 + j) * 32, h * g.Nh0, 0, k, 0);
              } else {
                for (int x = 0; x < nbx2; ++x)
                  tma_load_5d(dst + j * g.b2_bytes + x * 4096, &mapM, &sm->full[s], h * g.Nh0 + 32 * x, (b0 + j) * 32,
                              0, k, 0);
              }
            }
          }
          __syncwarp();
        }
      }
    }
    if (lane == 0) trace(tr, 2);
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue
    const int np1 = g.N1 - g.np0;
    const uint32_t idesc_p0 = idesc_tf32(128, g.np0, false, true);
    const uint32_t idesc_p1 = idesc_tf32(128, np1 > 0 ? np1 : 16, false, true);
    const uint32_t gblk_base = smem_u32(gblk);
    int it = 0, tc = 0, itk = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x, ++tc) {
      int n0, h, k_lo, k_hi;
      decode(t, n0, h, k_lo, k_hi);
      const int Nh = min(g.Nh0, g.N1 - h * g.Nh0);
      const uint32_t idesc_b = idesc_tf32(128, Nh, false, DA);
      for (int k = k_lo; k < k_hi; ++k, ++itk) {
        long long tw = tr ? trace_now() : 0;
        mbar_wait(&sm->g_full, itk & 1);                         // Gblk of (tile, k) is in shared memory
        if (tr && lane == 0) { const long long n = trace_now(); trace_add(tr, 0, n - tw); tw = n; }
        for (int kb = 0; kb < g.nka; ++kb, ++it) {
          const int s = it % stages;
          if (tr) tw = trace_now();
          mbar_wait(&sm->full[s], (it / stages) & 1);
          if (tr && lane == 0) trace_add(tr, 1, trace_now() - tw);
          tc_fence_after();
          if (it == 0 && lane == 0) trace(tr, 3);
          const uint32_t a_lo = desc_lo_k(gblk_base + kb * 16384);
          const uint32_t b_lo = desc_lo_mn(smem_u32(ring + s * g.stage_bytes), 4096);
          const uint32_t b1_lo = b_lo + (uint32_t)(g.np0 >> 5) * (4096u >> 4);
          if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (kb * 32 + ks * 8 < PLk) {
                const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
                mma_tf32_ss_lh(tmem, a_lo + ks * kDescStepK, kDescHiK, b_lo + ks * kDescStepMN, kDescHiMN, idesc_p0, acc);
                if (np1 > 0)
                  mma_tf32_ss_lh(tmem + g.np0, a_lo + ks * kDescStepK, kDescHiK, b1_lo + ks * kDescStepMN, kDescHiMN,
                                 idesc_p1, acc);
              }
            }
            mma_commit(&sm->empty[s]);
            if (kb == g.nka - 1) { mma_commit(&sm->u_full); mma_commit(&sm->g_empty); }
          }
          __syncwarp();
        }
        if (itk == 0 && lane == 0) trace(tr, 4);
        if (tr) tw = trace_now();
        if (k == k_lo && tc > 0) mbar_wait(&sm->o_empty, (tc - 1) & 1);   // the previous tile's O has been read out
        if (tr && lane == 0) trace_add(tr, 2, trace_now() - tw);
        for (int b0 = 0; b0 < g.nkb; b0 += g.kb2, ++it) {
          const int s = it % stages;
          const int nb = min(g.kb2, g.nkb - b0);
          if (tr) tw = trace_now();
          mbar_wait(&sm->full[s], (it / stages) & 1);
          if (tr && lane == 0) trace_add(tr, 3, trace_now() - tw);
          for (int j = 0; j < nb; ++j) {
            const int b = b0 + j;
            if (tr) tw = trace_now();
            mbar_wait(&sm->t_ready[b], itk & 1);                 // U columns [32 b, 32 b + 32) are rounded
            if (tr && lane == 0) trace_add(tr, 4, trace_now() - tw);
            tc_fence_after();
            const uint32_t bb = smem_u32(ring + s * g.stage_bytes) + (uint32_t)(j * g.b2_bytes);
            const uint32_t b_lo = DA ? desc_lo_mn(bb, 4096) : desc_lo_k(bb);
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if (b * 32 + ks * 8 < g.D)
                  mma_tf32_ts_lh(tmem_O, tmem + b * 32 + ks * 8, b_lo + ks * (DA ? kDescStepMN : kDescStepK),
                                 DA ? kDescHiMN : kDescHiK, idesc_b, (k > k_lo || b > 0 || ks > 0) ? 1u : 0u);
              }
              if (j == nb - 1) mma_commit(&sm->empty[s]);
              if (b == g.nkb - 1 && k == k_hi - 1) mma_commit(&sm->o_full);
            }
            __syncwarp();
          }
        }
        if (itk == 0 && lane == 0) trace(tr, 7);
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ U rounding / export, output epilogue
    const int quarter = warp & 3;
    const int set = (warp - 2) >> 2;
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    const int row = quarter * 32 + lane;
    const int p_lane = row / g.Lr;
    int tc = 0, itk = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x, ++tc) {
      int n0, h, k_lo, k_hi;
      decode(t, n0, h, k_lo, k_hi);
      const bool valid = (p_lane < g.P) && (n0 + p_lane < g.N);
      const long long grow = (long long)n0 * g.Lr + row;         // row of dq / da / U this thread owns
      const bool exporting = !DA && Uexp != nullptr && h == 0 && valid;
      for (int k = k_lo; k < k_hi; ++k, ++itk) {
        mbar_wait(&sm->u_full, itk & 1);
        tc_fence_after();
        if (itk == 0 && threadIdx.x == 64) trace(tr, 5);
        if (threadIdx.x == 64) trace(tr, 12);                    // keeps the LAST iteration's time
        float* urow = exporting ? Uexp + ((size_t)k * g.u_rows + grow) * g.Dp : nullptr;
        for (int c = set; c < nch; c += 2) {
          float v[32];
          const uint32_t ta = tmem + lane_bits + (uint32_t)(c * 32);
          if (c * 32 + 16 < g.N1) {
            tmem_ld32(ta, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = to_tf32(v[i]);
            tmem_st32(ta, v);
            if (urow) {
#pragma unroll
              for (int i = 0; i < 4; ++i) st_global_v8(urow + c * 32 + i * 8, v + 8 * i);
            }
          } else {
            tmem_ld16(ta, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = to_tf32(v[i]);
            tmem_st16(ta, v);
            if (urow) {
#pragma unroll
              for (int i = 0; i < 2; ++i) st_global_v8(urow + c * 32 + i * 8, v + 8 * i);
            }
          }
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm->t_ready[c]);
        }
        if (itk == 0 && threadIdx.x == 64) trace(tr, 6);
      }
      mbar_wait(&sm->o_full, tc & 1);
      tc_fence_after();
      if (tc == 0 && threadIdx.x == 64) trace(tr, 8);
      const int Nh = min(g.Nh0, g.N1 - h * g.Nh0);
      float* orow = out + grow * g.D + h * g.Nh0;
      const int ncols = min(Nh, g.D - h * g.Nh0);                // columns of this part that exist in the output
      for (int c = set; c * 32 < Nh; c += 2) {
        float v[32];
        const uint32_t ta = tmem_O + lane_bits + (uint32_t)(c * 32);
        const int w = (Nh - c * 32 > 16) ? 32 : 16;
        if (w == 32) tmem_ld32(ta, v); else tmem_ld16(ta, v);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = c * 32 + i * 4;
            if (i * 4 < w && col < ncols) {
              if (g.vec_out && col + 4 <= ncols) {
                float4* dst = reinterpret_cast<float4*>(orow + col);
                const float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                if (g.ksplit > 1) atomicAdd(dst, o); else *dst = o;
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (col + j < ncols) {
                    if (g.ksplit > 1) atomicAdd(orow + col + j, v[4 * i + j]); else orow[col + j] = v[4 * i + j];
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->o_empty);
      if (tc == 0 && threadIdx.x == 64) trace(tr, 9);
    }
    if (threadIdx.x == 64) trace(tr, 10);
  } else {
    // ------------------------------------------------------------ Gblk builders: thread r owns tile row r
    const int r = (warp - 10) * 32 + lane;
    const int p = r / g.Lr, l = r - p * g.Lr;
    // off-diagonal blocks stay zero for the whole kernel: clear the row once
    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        *reinterpret_cast<float4*>(gblk + kb * 16384 + swz128(r, c4)) = make_float4(0.f, 0.f, 0.f, 0.f);
    int itk = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      int n0, h, k_lo, k_hi;
      decode(t, n0, h, k_lo, k_hi);
      const bool valid = (p < g.P) && (n0 + p < g.N);
      for (int k = k_lo; k < k_hi; ++k, ++itk) {
        // The G values of this row are loaded into registers BEFORE waiting for the tile to be free (the loads do
        // not touch shared memory), 64 at a time with all loads of a batch in flight together.
        const float* Gnk = dS + ((size_t)(n0 + (valid ? p : 0)) * g.mc + k) * g.Lq * g.La;
        const float* src = DA ? Gnk + l : Gnk + (size_t)l * g.La;   // dA: column la = l (stride La); dQ: row lq = l
        const int col0 = p * g.Lk;
        float v[64];
        for (int e0 = 0; e0 < g.Lk; e0 += 64) {
          if (valid) {
            if (!DA && g.vec_g) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (e0 + 4 * i < g.Lk) {
                  const float4 x = __ldg(reinterpret_cast<const float4*>(src + e0 + 4 * i));
                  v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i)
                if (e0 + i < g.Lk) v[i] = __ldg(src + (size_t)(e0 + i) * (DA ? g.La : 1));
            }
          }
          if (e0 == 0 && itk > 0) mbar_wait(&sm->g_empty, (itk - 1) & 1);   // GEMM-A of the previous iteration has read Gblk
          if (valid) {
            if (!DA && g.vec_g) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (e0 + 4 * i < g.Lk) {
                  const int col = col0 + e0 + 4 * i;
                  *reinterpret_cast<float4*>(gblk + (col >> 5) * 16384 + swz128(r, (col & 31) >> 2)) =
                      make_float4(to_tf32(v[4 * i]), to_tf32(v[4 * i + 1]), to_tf32(v[4 * i + 2]), to_tf32(v[4 * i + 3]));
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i) {
                if (e0 + i < g.Lk) {
                  const int col = col0 + e0 + i;
                  *reinterpret_cast<float*>(gblk + (col >> 5) * 16384 + swz128(r, (col & 31) >> 2) + (col & 3) * 4) =
                      to_tf32(v[i]);
                }
              }
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->g_full);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, g.tmem_cols);
  if (threadIdx.x == 0) trace_end(tr);
}

}  // namespace

// which = 0: dq (N*Lq x D) from dS, ar = rounded answers, Mr; exports U (mc x N*Lq x Dp) when Uexp != nullptr.
// which = 1: da (N*La x D) from dS, xr = rounded questions, Mr.
// xr (rows x Dp) and Mr (mc x D x Dp) are TF32-rounded copies with 128-byte-aligned rows; dS is read as is.
// ksplit > 1 accumulates with atomics: the caller must have zeroed `out`.  Returns MMS_E_UNSUPPORTED for shapes the
// tile does not cover.
int mms_tc_simcross2_backward_fused_plan(int which, int N, int Lq, int La, int D, int mc, int ctas, int* ksplit) {
  const int sm_count = ctas;
  static const bool disabled = getenv("MMS_NO_FUSED") != nullptr || getenv("MMS_NO_FUSED_BWD") != nullptr;
  if (disabled) return MMS_E_UNSUPPORTED;
  if (Lq > 128 || La > 128 || N < 1) return MMS_E_UNSUPPORTED;
  const int N1 = mms_ceil_div(D, 16) * 16;
  if (N1 > 32 * kMaxChunks) return MMS_E_UNSUPPORTED;
  const int nh = 2 * N1 <= 512 ? 1 : 2;
  const int Nh0 = nh == 1 ? N1 : ((N1 / 2 + 31) & ~31);
  if (N1 + Nh0 > 512 || Nh0 > 256) return MMS_E_UNSUPPORTED;
  const int P = mms_max(1, mms_min(mms_min(128 / Lq, 128 / La), N));
  const long long base = (long long)mms_ceil_div(N, P) * nh;
  int ks = 1;
  if (base * 2 <= sm_count) ks = (int)mms_min<long long>(mc, sm_count / base);
  *ksplit = mms_max(1, ks);
  (void)which;
  return 0;
}

int mms_tc_simcross2_backward_fused(mms_context* ctx, int which, const float* xr, const float* Mr, const float* dS,
                                    float* out, float* Uexp, int N, int Lq, int La, int D, int mc, int Dp,
                                    int ksplit) {
  BwdGeom g;
  {
    int unused = 1;
    MMS_TRY(mms_tc_simcross2_backward_fused_plan(which, N, Lq, La, D, mc, ctx->sm_count, &unused));
    MMS_REQUIRE(ksplit >= 1 && ksplit <= mc, MMS_E_INVALID, "bad measure split");
  }
  const bool DA = which != 0;
  g.N = N; g.Lq = Lq; g.La = La; g.D = D; g.mc = mc;
  g.Lr = DA ? La : Lq; g.Lk = DA ? Lq : La;
  g.P = mms_max(1, mms_min(mms_min(128 / Lq, 128 / La), N));
  g.N1 = mms_ceil_div(D, 16) * 16;
  g.np0 = g.N1 <= 256 ? g.N1 : ((g.N1 / 2 + 31) & ~31);
  g.nh = 2 * g.N1 <= 512 ? 1 : 2;
  g.Nh0 = g.nh == 1 ? g.N1 : ((g.N1 / 2 + 31) & ~31);
  g.nka = mms_ceil_div(g.P * g.Lk, 32);
  g.nkb = mms_ceil_div(D, 32);
  g.nboxes = mms_ceil_div(g.N1, 32);
  g.b2_bytes = DA ? mms_ceil_div(g.Nh0, 32) * 4096 : mms_ceil_div(g.Nh0 * 128, 1024) * 1024;
  g.stage_bytes = mms_ceil_div(mms_max(g.nboxes * 4096, g.b2_bytes), 1024) * 1024;
  g.kb2 = mms_max(1, mms_min(4, g.stage_bytes / g.b2_bytes));
  int stages = kMaxStages;
  while (stages > 2 && (size_t)kGblkBytes + (size_t)stages * g.stage_bytes + sizeof(BwdSmem) + 1024 > 226 * 1024) --stages;
  if ((size_t)kGblkBytes + (size_t)stages * g.stage_bytes + sizeof(BwdSmem) + 1024 > 226 * 1024) return MMS_E_UNSUPPORTED;
  g.stages = stages;
  g.ksplit = ksplit;
  const long long total = (long long)mms_ceil_div(N, g.P) * g.nh * ksplit;
  if (total > 0x7fffffffLL) return MMS_E_UNSUPPORTED;
  g.total_tiles = (unsigned)total;
  g.tmem_cols = umma::tmem_cols_pow2((uint32_t)(g.N1 + g.Nh0));
  g.vec_g = (La % 4 == 0) && ((reinterpret_cast<uintptr_t>(dS) & 15) == 0);
  g.vec_out = (D % 4 == 0) && (g.Nh0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  g.Dp = Dp;
  g.u_rows = (long long)N * g.Lr;

  CUtensorMap mapX, mapM;
  MMS_TRY(mms_tc_make_map(ctx, &mapX, xr, Dp, true, D, (long long)N * g.Lk, 32, 0, 0, 0, 1, 1, 1));
  if (!DA) MMS_TRY(mms_tc_make_map(ctx, &mapM, Mr, Dp, false, D, D, g.Nh0, (long long)D * Dp, 0, 0, mc, 1, 1));
  else MMS_TRY(mms_tc_make_map(ctx, &mapM, Mr, Dp, true, D, D, 32, (long long)D * Dp, 0, 0, mc, 1, 1));

  static bool configured = false;
  if (!configured) {
    MMS_CUDA(cudaFuncSetAttribute(simcross2_bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
    MMS_CUDA(cudaFuncSetAttribute(simcross2_bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
    configured = true;
  }
  const size_t smem = (size_t)kGblkBytes + (size_t)stages * g.stage_bytes + sizeof(BwdSmem) + 1024;
  const unsigned grid = (unsigned)mms_min<long long>(total, ctx->sm_count);
  TraceBuf tb;
  MMS_TRY(tb.begin(grid));
  { MmsKernelScope ks_(ctx, DA ? "simcross2_bwd_fused_kernel<dA>" : "simcross2_bwd_fused_kernel<dQ>");
    if (!DA) simcross2_bwd_fused_kernel<false><<<grid, kThreads, smem, ctx->stream>>>(mapX, mapM, dS, out, Uexp, g, tb.dev);
    else simcross2_bwd_fused_kernel<true><<<grid, kThreads, smem, ctx->stream>>>(mapX, mapM, dS, out, nullptr, g, tb.dev); }
  MMS_LAUNCH_CHECK();
  static const char* const names[kTraceSlots] = {"entry", "setup", "tma_issued", "first_full", "gA_issued", "u_full",
                                                 "rounded", "gB_issued", "o_full", "epi0_done", "epi_done", "exit",
                                                 "last_u_full", nullptr, nullptr, nullptr,
                                                 "W:g_full", "W:fullA", "W:o_empty", "W:fullB", "W:t_ready", nullptr, nullptr,
                                                 nullptr};
  char what[112];
  snprintf(what, sizeof(what), "bwd %s N %d L %dx%d D %d mc %d P %d nh %d ksplit %d tiles %u stages %d", DA ? "dA" : "dQ",
           N, Lq, La, D, mc, g.P, g.nh, ksplit, g.total_tiles, stages);
  MMS_TRY(tb.end(ctx, what, names));
  return 0;
}
