// Sentence convolution, forward, as ONE dedicated tcgen05 kernel:  Yt[c][r] = sum_i sum_d W[c][i][d] x[r + i][d]
// (reference ConvolutionLayer::Forward_cpu for a kernel as wide as the input, conv_layer.cpp:25-43).
//
// The generic GEMM treats the kh shifted windows as one long reduction and fetches every 32-column slab of token rows
// kh times (48 KB of operands per 512 tensor-pipe cycles: ingest-bound at 47 % tensor pipe).  Here a slab of
// 256 + 8 token rows is staged ONCE per 32 columns and the kh MMAs of that slab read it at row offsets 0 .. kh-1:
// a K-major, 128-byte-swizzled tile can be read by tcgen05.mma from any row -- start address + i * 128 B with the
// descriptor's base-offset field 0 (tools/smem_row_shift_test.cu).  Per slab: 33 KB of x + kh blocks of filters
// (C rows each) for 4 * kh MMAs of N = 256.
//   warp 0: TMA producer (x slab in two boxes, kh filter blocks)   warp 1: MMA issue   warps 2-5: epilogue
//   two stages of 34 KB + kh * Cb * 128 B, two accumulators of 256 TMEM columns (epilogue of tile t beside tile t+1)
#include <cuda.h>

#include <cstdlib>

#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"

using namespace umma;

namespace {

constexpr int kRowsTile = 256;                 // sentence rows (MMA N) per tile
constexpr int kSlabBytes = 34 * 1024;          // (256 + 8) rows x 128 B, padded to a multiple of 1024
constexpr int kStages = 2;
constexpr int kEpiLd = 36;
constexpr int kThreads = 6 * 32;

// 2-D tiled bulk tensor load global -> shared, completion signalled on `bar`
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct ConvSmem {
  uint64_t full[kStages], empty[kStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// out(m, r) = sum_i sum_k F[m][i*Kdf + k] * X[(r + i)*ldx + k],  m < Mtot, r < rows, k < Kd
struct ConvGeom {
  int kh, Kdf, Mtot, Mper, Cb, m_tiles, wblock_bytes, stage_bytes, kblocks, row_major_out;
  long long rows, ld_out;                      // rows = output rows r to compute
  unsigned tiles;                              // row tiles x m_tiles, m fastest (the tiles of a slab run together)
};

__global__ void __launch_bounds__(kThreads, 1)
sentconv_fwd_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapX8,
                    const __grid_constant__ CUtensorMap mapW, float* __restrict__ out, const ConvGeom q) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* epi = reinterpret_cast<float*>(ring + kStages * q.stage_bytes);
  ConvSmem* sm = reinterpret_cast<ConvSmem*>(ring + kStages * q.stage_bytes + 4 * 32 * kEpiLd * 4);
  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, 512);
    tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapX8); tma_prefetch_desc(&mapW);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const uint32_t tx = (uint32_t)((kRowsTile + 8) * 128 + q.kh * q.Cb * 128);
    int it = 0;
    for (unsigned t = blockIdx.x; t < q.tiles; t += gridDim.x) {
      const int r0 = (int)(t / q.m_tiles) * kRowsTile, m0 = (int)(t % q.m_tiles) * q.Mper;
      for (int kb = 0; kb < q.kblocks; ++kb, ++it) {
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(&sm->empty[s], ((it / kStages) - 1) & 1);
        uint8_t* xs = ring + s * q.stage_bytes;
        uint8_t* ws = xs + kSlabBytes;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&sm->full[s], tx);
          tma_load_2d(xs, &mapX, &sm->full[s], kb * 32, r0);                                  // rows r0 .. r0+255
          tma_load_2d(xs + kRowsTile * 128, &mapX8, &sm->full[s], kb * 32, r0 + kRowsTile);   // .. r0+263
          for (int i = 0; i < q.kh; ++i)
            tma_load_2d(ws + i * q.wblock_bytes, &mapW, &sm->full[s], i * q.Kdf + kb * 32, m0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue
    const uint32_t idesc = idesc_tf32(128, kRowsTile, false, false);
    int it = 0, tcount = 0;
    for (unsigned t = blockIdx.x; t < q.tiles; t += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      if (tcount >= 2) mbar_wait(&sm->acc_empty[buf], ((tcount >> 1) - 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem + buf * kRowsTile;
      for (int kb = 0; kb < q.kblocks; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&sm->full[s], (it / kStages) & 1);
        tc_fence_after();
        const uint32_t xbase = smem_u32(ring + s * q.stage_bytes), wbase = xbase + kSlabBytes;
        if (elect_one_sync()) {
          for (int i = 0; i < q.kh; ++i) {
            const uint32_t a_lo = desc_lo_k(wbase + i * q.wblock_bytes);     // filters of kernel row i: the MMA's M side
            const uint32_t b_lo = desc_lo_k(xbase + i * 128);                // the slab, read from token row i on
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_tf32_ss_lh(acc, a_lo + ks * kDescStepK, kDescHiK, b_lo + ks * kDescStepK, kDescHiK, idesc,
                             (kb > 0 || i > 0 || ks > 0) ? 1u : 0u);
          }
          mma_commit(&sm->empty[s]);
          if (kb == q.kblocks - 1) mma_commit(&sm->acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: lane = filter, columns = sentence rows
    const int quarter = warp & 3;
    float* stage = epi + quarter * (32 * kEpiLd);
    int tcount = 0;
    for (unsigned t = blockIdx.x; t < q.tiles; t += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      const long long r0 = (long long)(t / q.m_tiles) * kRowsTile;
      const int m0 = (int)(t % q.m_tiles) * q.Mper, m_end = min(q.Mtot, m0 + q.Mper);
      mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem + buf * kRowsTile + ((uint32_t)(quarter * 32) << 16);
      const int c_first = m0 + quarter * 32;
      if (c_first < m_end) {
        for (int c0 = 0; c0 < kRowsTile && r0 + c0 < q.rows; c0 += 32) {
          float v[32];
          tmem_ld32(acc + c0, v);
          if (q.row_major_out) {                               // out[r][m]: the 32 lanes of a warp are 32 consecutive m
            const int m = c_first + lane;
            if (m < m_end) {
              float* p = out + (size_t)(r0 + c0) * q.ld_out + m;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (r0 + c0 + j < q.rows) p[(size_t)j * q.ld_out] = v[j];
            }
            continue;
          }
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4)
            *reinterpret_cast<float4*>(stage + lane * kEpiLd + i4 * 4) =
                make_float4(v[i4 * 4], v[i4 * 4 + 1], v[i4 * 4 + 2], v[i4 * 4 + 3]);
          __syncwarp();
          const int cc = (lane & 7) * 4, rr0 = lane >> 3;
          const long long col = r0 + c0 + cc;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int row = rr0 + rr * 4, c = c_first + row;
            if (c < m_end && col < q.rows) {
              const float4 o = *reinterpret_cast<const float4*>(stage + row * kEpiLd + cc);
              float* p = out + (size_t)c * q.ld_out + col;
              if (col + 3 < q.rows) *reinterpret_cast<float4*>(p) = o;
              else { p[0] = o.x; if (col + 1 < q.rows) p[1] = o.y; if (col + 2 < q.rows) p[2] = o.z; }
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient on the same idea:  dW[c][i][d] += sum_r G[r][c] * x[(r + i)*D + d]   (conv_layer.cpp:57-60).
// The reduction index r is the row of both tiles (MN-major operands); a slab of 32 + 8 token rows of x serves the
// kh accumulators dW[:, i, d-range] by being read from k-row i on (MN-major tiles can be read from any k-row too:
// tools/smem_row_shift_test.cu).  One CTA = one range of 96 d-columns x one slice of the token rows; kh accumulators
// of 96 TMEM columns; per 32-row k-block 16 KB of G + 15 KB of x feed 4 * kh MMAs of N = 96 (960 tensor-pipe
// cycles: 33 B/clk).  The slices add into dW with red.global.add (dW accumulates across calls anyway).
namespace {

constexpr int kDwN = 96, kDwStages = 6, kDwSlabRows = 40;
constexpr int kDwABytes = 4 * 4096, kDwBBlock = kDwSlabRows * 128, kDwBBytes = (kDwN / 32) * kDwBBlock;

struct DwSmem {
  uint64_t full[kDwStages], empty[kDwStages], acc_full;
  uint32_t tmem_base;
};
struct DwGeom {
  int kh, D, C, n_dr, ksplit;
  long long kblocks;                          // 32-row blocks of the reduction
  long long ldw;                              // kh * D
};

__global__ void __launch_bounds__(kThreads, 1)
sentconv_dw_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX,
                   float* __restrict__ dW, const DwGeom q) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kStageB = 32768;                               // 16 KB of G + 15 KB of x, padded to 32 KB
  DwSmem* sm = reinterpret_cast<DwSmem*>(ring + kDwStages * kStageB);
  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  const int dr = blockIdx.x % q.n_dr, slice = blockIdx.x / q.n_dr;
  const int d0 = dr * kDwN;
  const long long per = (q.kblocks + q.ksplit - 1) / q.ksplit;
  const long long kb0 = slice * per, kb1 = min(q.kblocks, kb0 + per);
  const int nkb = (int)max(0LL, kb1 - kb0);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kDwStages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
      mbar_init(&sm->acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, 512);
    tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapG); tma_prefetch_desc(&mapX);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 0) {
    for (int it = 0; it < nkb; ++it) {
      const int s = it % kDwStages;
      if (it >= kDwStages) mbar_wait(&sm->empty[s], ((it / kDwStages) - 1) & 1);
      uint8_t* a_dst = ring + s * kStageB;
      uint8_t* b_dst = a_dst + kDwABytes;
      const int r0 = (int)((kb0 + it) * 32);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&sm->full[s], (uint32_t)(kDwABytes + kDwBBytes));
#pragma unroll
        for (int b = 0; b < 4; ++b) tma_load_2d(a_dst + b * 4096, &mapG, &sm->full[s], 32 * b, r0);        // c block b
#pragma unroll
        for (int b = 0; b < kDwN / 32; ++b) tma_load_2d(b_dst + b * kDwBBlock, &mapX, &sm->full[s], d0 + 32 * b, r0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_tf32(128, kDwN, true, true);
    for (int it = 0; it < nkb; ++it) {
      const int s = it % kDwStages;
      mbar_wait(&sm->full[s], (it / kDwStages) & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(ring + s * kStageB), b_base = a_base + kDwABytes;
      const uint32_t a_lo = desc_lo_mn(a_base, 4096);
      if (elect_one_sync()) {
        for (int i = 0; i < q.kh; ++i) {
          const uint32_t b_lo = desc_lo_mn(b_base + i * 128, kDwBBlock);       // the slab from token row i on
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_tf32_ss_lh(tmem + i * kDwN, a_lo + ks * kDescStepMN, kDescHiMN, b_lo + ks * kDescStepMN, kDescHiMN, idesc,
                           (it > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit(&sm->empty[s]);
        if (it == nkb - 1) mma_commit(&sm->acc_full);
      }
      __syncwarp();
    }
  } else if (nkb > 0) {
    // ------------------------------------------------------------ epilogue: lane = filter c, columns = d; adds into dW
    const int quarter = warp & 3, c = quarter * 32 + lane;
    mbar_wait(&sm->acc_full, 0);
    tc_fence_after();
    for (int i = 0; i < q.kh; ++i) {
      for (int c0 = 0; c0 < kDwN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + i * kDwN + c0 + ((uint32_t)(quarter * 32) << 16), v);
        if (c < q.C) {
          float* p = dW + (size_t)c * q.ldw + (size_t)i * q.D + d0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {                   // red.global.add.v4.f32: D % 4 == 0, dW 16-byte aligned
            if (d0 + c0 + j + 3 < q.D) atomicAdd(reinterpret_cast<float4*>(p + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

// G: (rows, ldg) TF32-exact gradient rows (zero where no window starts), xr: (rows_total, D) TF32-exact token rows.
int mms_tc_sentconv_dw(mms_context* ctx, const float* G, long long rows, int ldg, const float* xr, long long rows_total,
                       float* dW, int D, int C, int kh) {
  static const bool disabled = mms_dev_knob("MMS_NO_SENTCONV_KERNEL") || mms_dev_knob("MMS_NO_SENTCONV_DW");
  if (disabled || C > 128 || kh > 5 || kh < 1 || D % 4 != 0 || ldg % 4 != 0 || rows < 1) return MMS_E_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(dW)) & 15) != 0)
    return MMS_E_UNSUPPORTED;
  DwGeom q;
  q.kh = kh; q.D = D; q.C = C;
  q.n_dr = (D + kDwN - 1) / kDwN;
  q.kblocks = (rows + 31) / 32;
  q.ksplit = (int)mms_max<long long>(1, mms_min<long long>(ctx->sm_count / q.n_dr, q.kblocks));
  q.ldw = (long long)kh * D;
  CUtensorMap mapG, mapX;
  const unsigned long long dg[2] = {(unsigned long long)ldg, (unsigned long long)rows};
  const unsigned long long sg[1] = {(unsigned long long)ldg * 4};
  const unsigned bg[2] = {32, 32};
  MMS_TRY(mms_tc_make_map_raw(ctx, &mapG, G, 2, dg, sg, bg, true));
  const unsigned long long dx[2] = {(unsigned long long)D, (unsigned long long)rows_total};
  const unsigned long long sx[1] = {(unsigned long long)D * 4};
  const unsigned bx[2] = {32, (unsigned)kDwSlabRows};
  MMS_TRY(mms_tc_make_map_raw(ctx, &mapX, xr, 2, dx, sx, bx, true));
  const size_t smem = (size_t)kDwStages * 32768 + sizeof(DwSmem) + 1024;
  static bool configured = false;
  if (!configured) {
    MMS_MAX_SMEM(sentconv_dw_kernel, 227 * 1024);
    configured = true;
  }
  { MmsKernelScope ks_(ctx, "sentconv_dw_kernel");
    sentconv_dw_kernel<<<q.n_dr * q.ksplit, kThreads, smem, ctx->stream>>>(mapG, mapX, dW, q); }
  MMS_LAUNCH_CHECK();
  return 0;
}

// out(m, r) = sum_{i < kh} sum_{k < Kd} F[m*ldf + i*Kdf + k] * X[(r + i)*ldx + k]   for m < Mtot, r < rows.
//   X: (rows_total, ldx) TF32-exact rows (rows past rows_total read as zero), F: (Mtot, ldf) TF32-exact.
//   row_major_out 0: out[m*ld_out + r] (channel-major, the forward's Yt)   1: out[r*ld_out + m] (the backward's dx)
// dry_run != 0 only answers whether the shape is served (0) or not (MMS_E_UNSUPPORTED).
int mms_tc_sentconv_shifted(mms_context* ctx, const float* X, long long rows_total, long long ldx, int Kd, const float* F,
                            long long ldf, int Kdf, int Mtot, int kh, float* out, long long ld_out, long long rows,
                            int row_major_out, int dry_run) {
  static const bool disabled = mms_dev_knob("MMS_NO_SENTCONV_KERNEL");
  if (disabled || kh > 8 || kh < 1 || ldx % 4 != 0 || ldf % 4 != 0 || rows < 1 || Mtot < 1 || Kd < 1) return MMS_E_UNSUPPORTED;
  if (!row_major_out && ld_out % 4 != 0) return MMS_E_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(F)) & 15) != 0) return MMS_E_UNSUPPORTED;
  if (!row_major_out && (reinterpret_cast<uintptr_t>(out) & 15) != 0) return MMS_E_UNSUPPORTED;
  ConvGeom q;
  q.kh = kh; q.Kdf = Kdf; q.Mtot = Mtot;
  q.m_tiles = (Mtot + 127) / 128;
  q.Mper = (Mtot + q.m_tiles - 1) / q.m_tiles;               // outputs per tile, <= 128
  q.Cb = (q.Mper + 7) & ~7;
  q.wblock_bytes = q.Cb * 128;
  q.stage_bytes = kSlabBytes + kh * q.wblock_bytes;
  q.kblocks = (Kd + 31) / 32;
  q.rows = rows; q.ld_out = ld_out; q.row_major_out = row_major_out;
  const long long row_tiles = (rows + kRowsTile - 1) / kRowsTile;
  if (row_tiles * q.m_tiles > 0x7fffffffLL) return MMS_E_UNSUPPORTED;
  q.tiles = (unsigned)(row_tiles * q.m_tiles);
  // (the MMA reads 128 rows of every filter block whatever Cb is: the rows past Cb are the next block, or the epilogue
  // staging behind the last one -- garbage that only reaches accumulator lanes >= Cb, which are never stored)
  const size_t smem = (size_t)kStages * q.stage_bytes + 4 * 32 * kEpiLd * 4 + sizeof(ConvSmem) + 1024;
  if (smem > 227 * 1024) return MMS_E_UNSUPPORTED;
  if (dry_run) return 0;
  CUtensorMap mapX, mapX8, mapW;
  const unsigned long long dx[2] = {(unsigned long long)Kd, (unsigned long long)rows_total};
  const unsigned long long sx[1] = {(unsigned long long)ldx * 4};
  const unsigned bx[2] = {32, (unsigned)kRowsTile}, bx8[2] = {32, 8};
  MMS_TRY(mms_tc_make_map_raw(ctx, &mapX, X, 2, dx, sx, bx, false));
  MMS_TRY(mms_tc_make_map_raw(ctx, &mapX8, X, 2, dx, sx, bx8, false));
  const unsigned long long dw[2] = {(unsigned long long)ldf, (unsigned long long)Mtot};
  const unsigned long long sw[1] = {(unsigned long long)ldf * 4};
  const unsigned bw[2] = {32, (unsigned)q.Cb};
  MMS_TRY(mms_tc_make_map_raw(ctx, &mapW, F, 2, dw, sw, bw, false));
  static bool configured = false;
  if (!configured) {
    MMS_MAX_SMEM(sentconv_fwd_kernel, 227 * 1024);
    configured = true;
  }
  const unsigned grid = mms_min<unsigned>(q.tiles, (unsigned)ctx->sm_count);
  { MmsKernelScope ks_(ctx, row_major_out ? "sentconv_dx_kernel" : "sentconv_fwd_kernel");
    sentconv_fwd_kernel<<<grid, kThreads, smem, ctx->stream>>>(mapX, mapX8, mapW, out, q); }
  MMS_LAUNCH_CHECK();
  return 0;
}

// forward: xr (rows_total, D), Wr (C, kh*D) -> Yt[c][r], r < rows
int mms_tc_sentconv_forward(mms_context* ctx, const float* xr, long long rows_total, const float* Wr, float* Yt,
                            long long rows, int D, int C, int kh, long long ldyt) {
  return mms_tc_sentconv_shifted(ctx, xr, rows_total, D, D, Wr, (long long)kh * D, D, C, kh, Yt, ldyt, rows, 0, 0);
}
