// SimCross mode 2 backward, bottom gradients, as fused tcgen05 kernels (reference:
// src/caffe/layers/sim_cross_layer.cpp:282-299, six cblas_sgemm calls per (pair, measure)):
//
//   dQ_n = sum_k (G_nk A_n) M_k^T          G_nk = dS[n,k]  (Lq x La)
//   dA_n = sum_k (G_nk^T Q_n) M_k          (= sum_k G_nk^T (Q_n M_k) of the reference, re-associated so that both
//                                           gradients have the same shape: small product first, D x D product second)
//
// One kernel template serves both (DA = false: dQ, DA = true: dA).  One tile = P consecutive QA pairs (their
// P*Lr output rows, Lr = Lq or La) x a range of measures -- all of them for the pair groups that fill whole waves of
// the grid, a 1/ksplit share for the groups left over after the last whole wave (and for every group of a small
// batch), whose CTAs then ADD into pre-zeroed output rows; the full-width output O [128 x N1] accumulates over the
// measures in TMEM columns [0, N1).  The intermediate U = Gblk * X never exists as a whole: it is produced, rounded
// and consumed in column chunks of CW (64) columns that rotate through three TMEM buffers behind O,
//
//   chunk (tile, k, c):
//     GEMM-A  U_c [128 x CW] = Gblk * X[:, chunk c]     Gblk [128 x 128]: block-diagonal, block p = G_{n0+p,k} (dQ) or
//                                                       its transpose (dA), built in shared memory by four builder
//                                                       warps straight from the rows of dS: a K-major tile for dQ;
//                                                       for dA the same rows as an MN-major tile that GEMM-A reads
//                                                       transposed through its descriptor.  X = the P pairs' answer
//                                                       (dQ) / question (dA) rows, MN-major boxes of 32 columns x 128 rows
//     round   U_c -> tf32 in place (tcgen05.ld / cvt.rna / tcgen05.st); the dQ kernel also writes U_c to global
//             memory -- each rounding warp its own piece, from registers, after announcing the chunk -- where the dM
//             contraction (dM_k = Q^T U_k, a reduction over ALL pairs) picks it up
//     GEMM-B  O [128 x N1] += U_c * W_k[chunk c rows]   W_k = M_k^T (dQ, K-major boxes) or M_k (dA, MN-major boxes);
//                                                       A operand read from TMEM
//
// and the MMA warp issues  A(0) A(1) | B(0) A(2) | B(1) A(3) | ...  across measure and tile boundaries: while the
// rounding warps work on chunk c+1 the tensor pipe runs B(c-1), A(c+1) .. B(c), A(c+2), so neither the rounding
// nor the Gblk build of the next measure leaves it idle.  (The previous version of this kernel kept a whole U in
// TMEM, which forced two column parts of O with U recomputed for each, and serialised GEMM-A -> round -> GEMM-B
// per measure: 35-41 % tensor-pipe activity.)
//
//   warp 0       TMA producer        warp 1       GEMM-A issuer + TMEM allocator        warp 14   GEMM-B issuer
//   warps 2-9    rounding of U, U export (two warps per TMEM lane quarter, 32 columns of the chunk each)
//   warps 10-13  Gblk builders        warps 15-18  output epilogue (one warp per TMEM lane quarter)
//
// The issue loops are written for the uniform datapath: every quantity an MMA needs besides the k-step offset is
// hoisted out of the unrolled k-steps (tools/mma_rate.cu: the tensor pipe runs every shape used here at its ideal
// N/2 cycles per K=8 step, but an issue loop that rebuilds descriptors from kernel parameters takes ~150 cycles per
// iteration, more than the MMAs it issues).
#include <cuda.h>

#include <stdlib.h>

#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"
#include "fused_trace.cuh"

namespace {

using namespace umma;

constexpr int kThreads = 19 * 32;
constexpr int kMaxStages = 6;
constexpr int kUBufs = 3;               // TMEM buffers the chunks of U rotate through
constexpr int kGblkBytes = 4 * 16384;   // 128 rows x 128 contraction columns, four 32-wide k-blocks

struct BwdGeom {
  int N, Lq, La, D, mc;
  int Lr, Lk;            // output-row side / contraction side sentence length
  int P;                 // pairs per tile
  int nksA;              // k-steps (of 8) of GEMM-A: ceil(P*Lk / 8)
  int N1, np0, np1;      // width of U and O; O is written by MMAs of np0 and np1 columns
  int CW, nch, wl;       // chunk width, chunks per (tile, measure), width of the last chunk
  int nbxB;              // dA: 32-column boxes of M_k per k-block
  int stages, stage_bytes;
  int ksplit;            // CTAs that share the measures of one pair group ...
  unsigned whole;        // ... from group `whole` on: tiles [0, whole) are groups 0.. with all their measures (whole
                         // waves of the grid), the groups left over are spread by measure so that the last wave is
                         // 1/ksplit of a tile long instead of a full one
  unsigned total_tiles;
  uint32_t tmem_cols;
  int vec_g, vec_s, vec_out;   // 16-byte dS loads (dQ) / 16-byte Gblk stores / 16-byte output stores
  int Dp;                // row pitch of the exported U
  int dbg;               // timing probes (MMS_BWD_PROBES builds only; WRONG results): 1 no rounding, 2 no output stores,
                         // 4 no U export, 8 no Gblk build, 16 no epilogue TMEM loads, 32 L2 prefetch of the next tile's X,
                         // 64 / 128 GEMM-B slots filled from private lines / with half of the bytes, 256 no dS loads
  long long u_rows;      // rows of one measure slab of the exported U
  int u_blocked;         // U export layout: 0 row-major (pitch Dp), 1 blocked (tc/simcross_dm.cu)
  int u_inline;          // dQ: the eight rounding warps export their own piece from registers (no export warps)
  int a_mn;              // dA: Gblk holds G itself (rows lq), read by GEMM-A as an MN-major A operand (= G^T)
  long long u_groups;    // blocked: 32-row groups per measure slab
};

struct BwdSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t g_full, g_empty, o_full, o_empty;
  uint64_t u_full[kUBufs], u_ready[kUBufs], u_free[kUBufs];
  uint32_t tmem_base;
};

// one full 32-byte sector per thread and instruction (row-per-thread stores of the U export); streaming (evict-first)
// like every store of these kernels: U, dq and da are written once and read by a LATER kernel, and must not push the
// M_k / X lines that the TMA ring keeps re-reading out of L2
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// Walks the chunks (tile, measure, chunk) of this CTA in issue order.  Every role keeps its own copy (the MMA warp
// and the producer keep two: GEMM-A runs two chunks ahead of GEMM-B).
// tile t -> its pair group's first pair and its range of measures
__device__ __forceinline__ void tile_map(const BwdGeom& g, unsigned t, int& n0, int& k_lo, int& k_hi) {
  if (t < g.whole) { n0 = (int)t * g.P; k_lo = 0; k_hi = g.mc; return; }
  const unsigned u = t - g.whole;
  const int ksp = (int)(u % (unsigned)g.ksplit);
  n0 = (int)(g.whole + u / (unsigned)g.ksplit) * g.P;
  k_lo = ksp * g.mc / g.ksplit;
  k_hi = (ksp + 1) * g.mc / g.ksplit;
}

struct ChunkIter {
  unsigned t;
  int n0, k, k_lo, k_hi, c;
  bool live;
  __device__ __forceinline__ void load(const BwdGeom& g) {
    live = t < g.total_tiles;
    c = 0;
    if (live) { tile_map(g, t, n0, k_lo, k_hi); k = k_lo; }
  }
  __device__ __forceinline__ void start(const BwdGeom& g) { t = blockIdx.x; load(g); }
  __device__ __forceinline__ void next(const BwdGeom& g) {
    if (++c == g.nch) {
      c = 0;
      if (++k == k_hi) { t += gridDim.x; load(g); }
    }
  }
};

// ring slot / U buffer cursors: index and phase bit, advanced without divisions
struct Cursor {
  int i; uint32_t ph;
  __device__ __forceinline__ void advance(int n) { if (++i == n) { i = 0; ph ^= 1u; } }
  __device__ __forceinline__ void skip(int by, int n) { i += by; if (i >= n) { i -= n; ph ^= 1u; } }   // by < n
};

template <bool DA, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
simcross2_bwd_fused_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapM,
                           const float* __restrict__ dS, float* __restrict__ out, float* __restrict__ Uexp,
                           const BwdGeom g, long long* tr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* gblk = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = gblk + kGblkBytes;
  BwdSmem* sm = reinterpret_cast<BwdSmem*>(ring + g.stages * g.stage_bytes);

  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  if (TRACE && threadIdx.x == 0) trace_begin(tr);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
      mbar_init(&sm->g_full, 4); mbar_init(&sm->g_empty, 1);
      mbar_init(&sm->o_full, 1); mbar_init(&sm->o_empty, 4);
      for (int j = 0; j < kUBufs; ++j) { mbar_init(&sm->u_full[j], 1); mbar_init(&sm->u_ready[j], (DA || g.u_inline) ? 8 : 4); mbar_init(&sm->u_free[j], (DA || g.u_inline) ? 1 : 5); }
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
  const uint32_t tmem_O = sm->tmem_base;
  const uint32_t tmem_U = tmem_O + (uint32_t)g.N1;
  if (TRACE && threadIdx.x == 0) trace(tr, 1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const int stages = g.stages, stage_bytes = g.stage_bytes, CW = g.CW, nch = g.nch, wl = g.wl;
    const int Lk = g.Lk, np0 = g.np0, nbxB = g.nbxB;
    const bool two = g.np1 > 0;
    const uint32_t bytesB = DA ? (uint32_t)nbxB * 4096u : (uint32_t)(two ? 2 : 1) * (uint32_t)np0 * 128u;
    Cursor slot = {0, 0};
    auto load_a = [&](const ChunkIter& it) {               // X[:, chunk c]: 32-column x 128-row boxes
      const int w = it.c == nch - 1 ? wl : CW;
      const int nbx = (w + 31) >> 5;
      mbar_wait(&sm->empty[slot.i], slot.ph ^ 1u);
      uint8_t* dst = ring + slot.i * stage_bytes;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&sm->full[slot.i], (uint32_t)nbx * 16384u);
        for (int x = 0; x < nbx; ++x)
          tma_load_5d(dst + x * 16384, &mapX, &sm->full[slot.i], it.c * CW + 32 * x, it.n0 * Lk, 0, 0, 0);
        // the next tile's X rows come from HBM: pull them into L2 a whole tile ahead, so that the ring (which holds
        // about 1 us of operands) only ever waits for L2
        if (it.c == 0 && it.k == it.k_lo && (g.dbg & 32)) {
          const unsigned tn = it.t + gridDim.x;
          if (tn < g.total_tiles) {
            int n0n, kl, kh;
            tile_map(g, tn, n0n, kl, kh);
            for (int x = 0; x < nbxB; ++x) tma_prefetch_l2_5d(&mapX, 32 * x, n0n * Lk, 0, 0, 0);
          }
        }
      }
      __syncwarp();
      slot.advance(stages);
    };
    auto load_b = [&](const ChunkIter& it) {               // rows [chunk c] of W_k, one 32-row k-block per slot
      const int w = it.c == nch - 1 ? wl : CW;
      const int nkb = (w + 31) >> 5;
      for (int kb = 0; kb < nkb; ++kb) {
        const int e0 = it.c * CW + kb * 32;
        mbar_wait(&sm->empty[slot.i], slot.ph ^ 1u);
        uint8_t* dst = ring + slot.i * stage_bytes;
        if (elect_one_sync()) {
          if (g.dbg & 64) {            // timing probe: the same slot filled from this CTA's own X rows (no shared lines)
            mbar_arrive_expect_tx(&sm->full[slot.i], 32768u);
            tma_load_5d(dst, &mapX, &sm->full[slot.i], (e0 & 255), it.n0 * Lk, 0, 0, 0);
            tma_load_5d(dst + 16384, &mapX, &sm->full[slot.i], ((e0 + 32) & 255), it.n0 * Lk, 0, 0, 0);
          } else if (g.dbg & 128) {    // timing probe: half of the bytes
            mbar_arrive_expect_tx(&sm->full[slot.i], DA ? 5u * 4096u : (uint32_t)np0 * 128u);
            if (!DA) tma_load_5d(dst, &mapM, &sm->full[slot.i], e0, 0, 0, it.k, 0);
            else for (int x = 0; x < 5; ++x) tma_load_5d(dst + x * 4096, &mapM, &sm->full[slot.i], 32 * x, e0, 0, it.k, 0);
          } else {
          mbar_arrive_expect_tx(&sm->full[slot.i], bytesB);
          if (!DA) {
            tma_load_5d(dst, &mapM, &sm->full[slot.i], e0, 0, 0, it.k, 0);
            if (two) tma_load_5d(dst + np0 * 128, &mapM, &sm->full[slot.i], e0, np0, 0, it.k, 0);
          } else {
            for (int x = 0; x < nbxB; ++x)
              tma_load_5d(dst + x * 4096, &mapM, &sm->full[slot.i], 32 * x, e0, 0, it.k, 0);
          }
          }
        }
        __syncwarp();
        slot.advance(stages);
      }
    };
    ChunkIter ia, ib;
    ia.start(g); ib.start(g);
    for (int i = 0; i < 2 && ia.live; ++i) { load_a(ia); ia.next(g); }
    while (ib.live) {
      load_b(ib); ib.next(g);
      if (ia.live) { load_a(ia); ia.next(g); }
    }
    if (TRACE && lane == 0) trace(tr, 2);
  } else if (warp == 1 || warp == 14) {
    // ------------------------------------------------------------ MMA issue: warp 1 GEMM-A, warp 14 GEMM-B
    // Two issuers because a lone warp retires one dependent uniform-datapath instruction every ~5 cycles: with one
    // issuer the instruction stream of the 151 short MMAs of a (tile, measure) takes about as long as the MMAs
    // themselves (ncu: the single issue warp was busy, not waiting, 55 % of the kernel).  Both warps walk the same
    // static schedule  A(0) A(1) | B(0) A(2) | B(1) A(3) | ...  over the one TMA ring and act on their own items
    // only; what the in-order issue of a single warp used to guarantee is now a barrier: u_free[j] (GEMM-B has
    // read U buffer j) before GEMM-A overwrites it.
    const bool is_a = warp == 1;
    const int stages = g.stages, stage_bytes = g.stage_bytes, CW = g.CW, nch = g.nch, wl = g.wl;
    const int nksA = g.nksA, D = g.D, np0 = g.np0;
    const bool two = g.np1 > 0;
    const uint32_t ring_base = smem_u32(ring);
    const uint32_t stage_lo = (uint32_t)stage_bytes >> 4;
    Cursor slot = {0, 0};
    long long tw = 0;
    ChunkIter ia, ib;
    ia.start(g); ib.start(g);
    if (is_a) {
      const bool a_mn = DA && g.a_mn;
      const uint32_t idesc_a = idesc_tf32(128, CW, a_mn, true);
      const uint32_t idesc_al = idesc_tf32(128, wl, a_mn, true);
      const uint32_t gblk_lo = a_mn ? desc_lo_mn(smem_u32(gblk), 16384) : desc_lo_k(smem_u32(gblk));
      const uint32_t ring_lo_mnA = desc_lo_mn(ring_base, 16384);
      const int a_full = nksA >> 2, a_rem = nksA & 3;
      Cursor ua = {0, 0};
      uint32_t gph = 0;          // phase of g_full
      auto issue_a = [&](const ChunkIter& it) {
        if (it.c == 0) {
          if (TRACE) tw = trace_now();
          mbar_wait(&sm->g_full, gph);                       // Gblk of (tile, k) is in shared memory
          gph ^= 1u;
          if (TRACE && lane == 0) trace_add(tr, 0, trace_now() - tw);
        }
        if (TRACE) tw = trace_now();
        mbar_wait(&sm->u_free[ua.i], ua.ph ^ 1u);            // GEMM-B of the chunk three back has read this buffer
        if (TRACE && lane == 0) trace_add(tr, 5, trace_now() - tw);
        if (TRACE) tw = trace_now();
        mbar_wait(&sm->full[slot.i], slot.ph);
        if (TRACE && lane == 0) trace_add(tr, 1, trace_now() - tw);
        tc_fence_after();
        const bool last = it.c == nch - 1;
        const uint32_t d = tmem_U + (uint32_t)(ua.i * CW);
        const uint32_t idesc = last ? idesc_al : idesc_a;
        if (elect_one_sync()) {
          uint32_t al = gblk_lo, bl = ring_lo_mnA + (uint32_t)slot.i * stage_lo;
          if (a_mn) {              // A = Gblk read MN-major: 8 contraction rows (1024 bytes) per k-step, no 32-column blocks
#pragma unroll 1
            for (int kb = 0; kb < a_full; ++kb) {
              mma_tf32_ss_lh(d, al, kDescHiMN, bl, kDescHiMN, idesc, kb > 0 ? 1u : 0u);
              mma_tf32_ss_lh(d, al + 1 * kDescStepMN, kDescHiMN, bl + 1 * kDescStepMN, kDescHiMN, idesc, 1u);
              mma_tf32_ss_lh(d, al + 2 * kDescStepMN, kDescHiMN, bl + 2 * kDescStepMN, kDescHiMN, idesc, 1u);
              mma_tf32_ss_lh(d, al + 3 * kDescStepMN, kDescHiMN, bl + 3 * kDescStepMN, kDescHiMN, idesc, 1u);
              al += 4 * kDescStepMN; bl += 4 * kDescStepMN;
            }
            if (a_rem > 0) mma_tf32_ss_lh(d, al, kDescHiMN, bl, kDescHiMN, idesc, a_full > 0 ? 1u : 0u);
            if (a_rem > 1) mma_tf32_ss_lh(d, al + 1 * kDescStepMN, kDescHiMN, bl + 1 * kDescStepMN, kDescHiMN, idesc, 1u);
            if (a_rem > 2) mma_tf32_ss_lh(d, al + 2 * kDescStepMN, kDescHiMN, bl + 2 * kDescStepMN, kDescHiMN, idesc, 1u);
          } else {
#pragma unroll 1
          for (int kb = 0; kb < a_full; ++kb) {
            mma_tf32_ss_lh(d, al, kDescHiK, bl, kDescHiMN, idesc, kb > 0 ? 1u : 0u);
            mma_tf32_ss_lh(d, al + 1 * kDescStepK, kDescHiK, bl + 1 * kDescStepMN, kDescHiMN, idesc, 1u);
            mma_tf32_ss_lh(d, al + 2 * kDescStepK, kDescHiK, bl + 2 * kDescStepMN, kDescHiMN, idesc, 1u);
            mma_tf32_ss_lh(d, al + 3 * kDescStepK, kDescHiK, bl + 3 * kDescStepMN, kDescHiMN, idesc, 1u);
            al += 16384u >> 4; bl += 4 * kDescStepMN;
          }
          if (a_rem > 0) mma_tf32_ss_lh(d, al, kDescHiK, bl, kDescHiMN, idesc, a_full > 0 ? 1u : 0u);
          if (a_rem > 1) mma_tf32_ss_lh(d, al + 1 * kDescStepK, kDescHiK, bl + 1 * kDescStepMN, kDescHiMN, idesc, 1u);
          if (a_rem > 2) mma_tf32_ss_lh(d, al + 2 * kDescStepK, kDescHiK, bl + 2 * kDescStepMN, kDescHiMN, idesc, 1u);
          }
          mma_commit(&sm->empty[slot.i]);
          mma_commit(&sm->u_full[ua.i]);
          if (last) mma_commit(&sm->g_empty);
        }
        __syncwarp();
        slot.advance(stages);
        ua.advance(kUBufs);
      };
      for (int i = 0; i < 2 && ia.live; ++i) { issue_a(ia); ia.next(g); }
      if (TRACE && lane == 0) trace(tr, 4);
      while (ib.live) {
        slot.skip(((ib.c == nch - 1 ? wl : CW) + 31) >> 5, stages);     // the ring slots of B(c)
        ib.next(g);
        if (ia.live) { issue_a(ia); ia.next(g); }
      }
    } else {
      const uint32_t idesc_b0 = idesc_tf32(128, np0, false, DA);
      const uint32_t idesc_b1 = idesc_tf32(128, two ? g.np1 : 16, false, DA);
      const uint32_t b1_off = DA ? (uint32_t)(np0 >> 5) * (4096u >> 4) : ((uint32_t)np0 * 128u) >> 4;
      constexpr uint32_t kStepB = DA ? kDescStepMN : kDescStepK;
      constexpr uint32_t kHiB = DA ? kDescHiMN : kDescHiK;
      const uint32_t ring_lo_B = DA ? desc_lo_mn(ring_base, 4096) : desc_lo_k(ring_base);
      const uint32_t tmem_O1 = tmem_O + (uint32_t)np0;
      Cursor ub = {0, 0};
      int tcb = 0;               // tiles whose GEMM-B has started
      auto issue_b = [&](const ChunkIter& it) {
        const bool tile_first = it.c == 0 && it.k == it.k_lo;
        const bool tile_last = it.c == nch - 1 && it.k == it.k_hi - 1;
        if (tile_first && tcb > 0) {
          if (TRACE) tw = trace_now();
          mbar_wait(&sm->o_empty, (uint32_t)(tcb - 1) & 1u);   // the previous tile's O has been read out
          if (TRACE && lane == 0) trace_add(tr, 2, trace_now() - tw);
        }
        if (TRACE) tw = trace_now();
        mbar_wait(&sm->u_ready[ub.i], ub.ph);                  // chunk c of U is rounded
        if (TRACE && lane == 0) trace_add(tr, 4, trace_now() - tw);
        const int w = it.c == nch - 1 ? wl : CW;
        const int nkb = (w + 31) >> 5;
        for (int kb = 0; kb < nkb; ++kb) {
          if (TRACE) tw = trace_now();
          mbar_wait(&sm->full[slot.i], slot.ph);
          if (TRACE && lane == 0) trace_add(tr, 3, trace_now() - tw);
          tc_fence_after();
          const uint32_t b_lo = ring_lo_B + (uint32_t)slot.i * stage_lo;
          const uint32_t a_t = tmem_U + (uint32_t)(ub.i * CW + kb * 32);
          const int left = D - (it.c * CW + kb * 32);          // valid contraction columns from this k-block on
          const int nks = left >= 32 ? 4 : (left + 7) >> 3;
          const uint32_t acc0 = (tile_first && kb == 0) ? 0u : 1u;
          if (elect_one_sync()) {
            if (two) {
              const uint32_t b1_lo = b_lo + b1_off;
              mma_tf32_ts_lh(tmem_O, a_t, b_lo, kHiB, idesc_b0, acc0);
              mma_tf32_ts_lh(tmem_O1, a_t, b1_lo, kHiB, idesc_b1, acc0);
              if (nks > 1) {
                mma_tf32_ts_lh(tmem_O, a_t + 8, b_lo + 1 * kStepB, kHiB, idesc_b0, 1u);
                mma_tf32_ts_lh(tmem_O1, a_t + 8, b1_lo + 1 * kStepB, kHiB, idesc_b1, 1u);
              }
              if (nks > 2) {
                mma_tf32_ts_lh(tmem_O, a_t + 16, b_lo + 2 * kStepB, kHiB, idesc_b0, 1u);
                mma_tf32_ts_lh(tmem_O1, a_t + 16, b1_lo + 2 * kStepB, kHiB, idesc_b1, 1u);
              }
              if (nks > 3) {
                mma_tf32_ts_lh(tmem_O, a_t + 24, b_lo + 3 * kStepB, kHiB, idesc_b0, 1u);
                mma_tf32_ts_lh(tmem_O1, a_t + 24, b1_lo + 3 * kStepB, kHiB, idesc_b1, 1u);
              }
            } else {
              mma_tf32_ts_lh(tmem_O, a_t, b_lo, kHiB, idesc_b0, acc0);
              if (nks > 1) mma_tf32_ts_lh(tmem_O, a_t + 8, b_lo + 1 * kStepB, kHiB, idesc_b0, 1u);
              if (nks > 2) mma_tf32_ts_lh(tmem_O, a_t + 16, b_lo + 2 * kStepB, kHiB, idesc_b0, 1u);
              if (nks > 3) mma_tf32_ts_lh(tmem_O, a_t + 24, b_lo + 3 * kStepB, kHiB, idesc_b0, 1u);
            }
            mma_commit(&sm->empty[slot.i]);
            if (kb == nkb - 1) {
              mma_commit(&sm->u_free[ub.i]);
              if (tile_last) mma_commit(&sm->o_full);
            }
          }
          __syncwarp();
          slot.advance(stages);
        }
        ub.advance(kUBufs);
        if (tile_last) ++tcb;
      };
      for (int i = 0; i < 2 && ia.live; ++i) { slot.advance(stages); ia.next(g); }   // the ring slots of A(0), A(1)
      while (ib.live) {
        issue_b(ib); ib.next(g);
        if (ia.live) { slot.advance(stages); ia.next(g); }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ U rounding / export
    // Eight warps round (two per TMEM lane quarter, 32 columns of the chunk each); in the dQ kernel each then exports the
    // piece it holds (u_inline).  The older split -- warps 2-5 round the whole chunk, warps 6-9 read it back from TMEM
    // and export it -- survives for the probes build only (MMS_BWD_EXPORT_WARPS): 0.395 vs 0.386 ms.
    const int quarter = warp & 3;
    const int set = (warp - 2) >> 2;
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    const int row = quarter * 32 + lane;
    const int p_lane = row / g.Lr;
    const int CW = g.CW, nch = g.nch, wl = g.wl;
    Cursor ur = {0, 0};
    bool first = true;
    ChunkIter it;
    it.start(g);
    if (!DA && set == 1 && g.u_blocked && Uexp != nullptr && blockIdx.x == 0) {
      {
        // rows past the last token row of the last 32-row group are read by the dM kernel: they must be zero
        const int tail0 = (int)(g.u_rows & 31);
        if (tail0 != 0) {
          const int ntail = 32 - tail0, per_k = g.N1 * ntail;
          for (int e = (warp - 6) * 32 + lane; e < g.mc * per_k; e += 128) {
            const int kk = e / per_k, rem = e - kk * per_k;
            const int col = rem / ntail, r = tail0 + rem - col * ntail;
            Uexp[(((size_t)kk * g.u_groups + (size_t)(g.u_groups - 1)) * g.Dp + col) * 32 + r] = 0.f;
          }
        }
      }
    }
    if (!DA && set == 1 && !g.u_inline) {
      while (it.live) {
        const bool valid = (p_lane < g.P) && (it.n0 + p_lane < g.N) && Uexp != nullptr && !(g.dbg & 4);
        const long long grow = (long long)it.n0 * g.Lr + row;    // row of U this thread owns
        mbar_wait(&sm->u_ready[ur.i], ur.ph);
        tc_fence_after();
        const int w = it.c == nch - 1 ? wl : CW;
        // row-major: this row's 32 bytes every 8 columns (a different 128-byte line per lane: 32 LSU wavefronts per
        // instruction); blocked (tc/simcross_dm.cu): the 32 rows of a group are contiguous per column, so one 4-byte
        // store per column writes 128 consecutive bytes per warp
        float* urow = g.u_blocked
            ? Uexp + (((size_t)it.k * g.u_groups + (size_t)(grow >> 5)) * g.Dp + (size_t)(it.c * CW)) * 32 + (size_t)(grow & 31)
            : Uexp + ((size_t)it.k * g.u_rows + grow) * g.Dp + it.c * CW;
        for (int c0 = 0; c0 < w; c0 += 32) {
          float v[32];
          const uint32_t ta = tmem_U + lane_bits + (uint32_t)(ur.i * CW + c0);
          const bool wide = w - c0 > 16;
          if (wide) tmem_ld32(ta, v); else tmem_ld16(ta, v);
          if (valid) {
            if (g.u_blocked) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < 16 || wide) __stcs(urow + (size_t)(c0 + i) * 32, v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < 2 || wide) st_global_v8(urow + c0 + i * 8, v + 8 * i);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->u_free[ur.i]);
        ur.advance(kUBufs);
        it.next(g);
      }
    } else {
      // Two warps per TMEM lane quarter, 32 columns of the chunk each (dQ without u_inline: warps 2-5 alone, 64 columns).
      // dQ with u_inline: a warp's rounded piece is still in its registers when it has announced the chunk -- it goes to
      // global memory from there (blocked layout: 128 consecutive bytes per warp and instruction), after the arrival,
      // so the export never sits between GEMM-A and GEMM-B of this chunk and TMEM is read once, not twice.
      const bool pair = DA || g.u_inline;
      const int step = pair ? 64 : 32;                            // column stride between the pieces of one warp
      while (it.live) {
        mbar_wait(&sm->u_full[ur.i], ur.ph);
        tc_fence_after();
        if (TRACE && first && threadIdx.x == 64) trace(tr, 5);
        const int w = it.c == nch - 1 ? wl : CW;
        float v[32];
        int cx = -1;                                              // first column of the piece held in v
        bool wide = false;
        if (!(g.dbg & 1)) {
          for (int c0 = pair ? set * 32 : 0; c0 < w; c0 += step) {
            const uint32_t ta = tmem_U + lane_bits + (uint32_t)(ur.i * CW + c0);
            cx = c0;
            if (w - c0 > 16) {
              wide = true;
              tmem_ld32(ta, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = to_tf32(v[i]);
              tmem_st32(ta, v);
            } else {
              wide = false;
              tmem_ld16(ta, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = to_tf32(v[i]);
              tmem_st16(ta, v);
            }
          }
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->u_ready[ur.i]);
        if (TRACE && first && threadIdx.x == 64) trace(tr, 6);
        first = false;
        if (!DA && g.u_inline && cx >= 0) {
          const bool valid = (p_lane < g.P) && (it.n0 + p_lane < g.N) && Uexp != nullptr && !(g.dbg & 4);
          if (valid) {
            const long long grow = (long long)it.n0 * g.Lr + row;
            if (g.u_blocked) {
              float* urow = Uexp + (((size_t)it.k * g.u_groups + (size_t)(grow >> 5)) * g.Dp + (size_t)(it.c * CW + cx)) * 32 +
                            (size_t)(grow & 31);
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < 16 || wide) __stcs(urow + (size_t)i * 32, v[i]);
            } else {
              float* urow = Uexp + ((size_t)it.k * g.u_rows + grow) * g.Dp + it.c * CW + cx;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < 2 || wide) st_global_v8(urow + i * 8, v + 8 * i);
            }
          }
        }
        ur.advance(kUBufs);
        it.next(g);
      }
    }
    if (TRACE && threadIdx.x == 64) trace(tr, 10);
  } else if (warp >= 15) {
    // ------------------------------------------------------------ output epilogue: O -> dq / da rows
    // Row-per-thread stores are bound by the LSU (one 32-byte sector per lane and instruction): 32-byte stores where
    // the row allows (rows of D = 300 floats are 32-byte aligned every other row; the others start with one 16-byte
    // store) halve the instruction count of the 16-byte version.
    const int quarter = warp & 3;
    const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
    const int row = quarter * 32 + lane;
    const int p_lane = row / g.Lr;
    const int N1 = g.N1, D = g.D;
    int tc = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x, ++tc) {
      int n0, k_lo_, k_hi_;
      tile_map(g, t, n0, k_lo_, k_hi_);
      const bool split = g.ksplit > 1 && t >= g.whole;              // this tile's measures are shared: add, do not store
      const bool valid = (p_lane < g.P) && (n0 + p_lane < g.N) && !(g.dbg & 2);
      const long long grow = (long long)n0 * g.Lr + row;
      float* orow = out + grow * D;
      const bool shifted = (reinterpret_cast<uintptr_t>(orow) & 31) != 0;   // 16-byte aligned only
      mbar_wait(&sm->o_full, (uint32_t)tc & 1u);
      tc_fence_after();
      if (TRACE && tc == 0 && lane == 0 && quarter == 0) trace(tr, 8);
      for (int c = 0; c * 32 < N1; ++c) {
        float v[32];
        const uint32_t ta = tmem_O + lane_bits + (uint32_t)(c * 32);
        const int wc = (N1 - c * 32 > 16) ? 32 : 16;
        if (!(g.dbg & 16)) { if (wc == 32) tmem_ld32(ta, v); else tmem_ld16(ta, v); }
        if (!valid) continue;
        const int ncols = mms_min(wc, D - c * 32);               // columns of this chunk that exist in the output
        float* dst = orow + c * 32;
        if (split || !g.vec_out) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i * 4 < ncols) {
              if (g.vec_out && i * 4 + 4 <= ncols) {
                const float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                if (split) atomicAdd(reinterpret_cast<float4*>(dst + 4 * i), o);
                else __stcs(reinterpret_cast<float4*>(dst + 4 * i), o);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (i * 4 + j < ncols) {
                    if (split) atomicAdd(dst + 4 * i + j, v[4 * i + j]); else dst[4 * i + j] = v[4 * i + j];
                  }
                }
              }
            }
          }
        } else if (!shifted) {          // D % 4 == 0: ncols is a multiple of 4
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i * 8 + 8 <= ncols) st_global_v8(dst + i * 8, v + i * 8);
            else if (i * 8 + 4 <= ncols)
              __stcs(reinterpret_cast<float4*>(dst + i * 8), make_float4(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3]));
          }
        } else {
          if (ncols >= 4) __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c0 = 4 + i * 8;
            if (c0 + 8 <= ncols) st_global_v8(dst + c0, v + c0);
            else if (c0 + 4 <= ncols)
              __stcs(reinterpret_cast<float4*>(dst + c0), make_float4(v[c0], v[c0 + 1], v[c0 + 2], v[c0 + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->o_empty);
      if (TRACE && tc == 0 && lane == 0 && quarter == 0) trace(tr, 9);
    }
  } else if (warp >= 10 && warp < 14) {
    // ------------------------------------------------------------ Gblk builders: thread r owns tile row r
    // dA with a_mn: Gblk holds G itself -- thread r owns CONTRACTION row r = (pair p, lq) and writes the row of G it reads
    // with 16-byte loads (exactly what the dQ builder does) as 16-byte stores along the M direction of an MN-major
    // tile; GEMM-A reads that tile transposed through its descriptor.  (Before: one 4-byte load per element down a
    // column of G, 40 load instructions per thread and measure instead of 10.)
    const bool a_mn = DA && g.a_mn;
    const int r = (warp - 10) * 32 + lane;
    const int own = a_mn ? g.Lk : g.Lr;                    // rows per pair in the dimension this thread's index runs over
    const int p = r / own, l = r - p * own;
    // off-diagonal blocks stay zero for the whole kernel: clear the row once
    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        *reinterpret_cast<float4*>(gblk + kb * 16384 + swz128(r, c4)) = make_float4(0.f, 0.f, 0.f, 0.f);
    int itk = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      int n0, k_lo, k_hi;
      tile_map(g, t, n0, k_lo, k_hi);
      const bool valid = (p < g.P) && (n0 + p < g.N) && !(g.dbg & 8);
      for (int k = k_lo; k < k_hi; ++k, ++itk) {
        // The G values of this row are loaded into registers BEFORE waiting for the tile to be free (the loads do
        // not touch shared memory), 64 at a time with all loads of a batch in flight together.
        const float* Gnk = dS + ((size_t)(n0 + (valid ? p : 0)) * g.mc + k) * g.Lq * g.La;
        const bool by_row = !DA || a_mn;                            // this thread reads row lq = l of G
        const float* src = by_row ? Gnk + (size_t)l * g.La : Gnk + l;   // else (dA, K-major Gblk): column la = l, stride La
        const int nval = a_mn ? g.Lr : g.Lk;                        // values per thread
        const int col0 = p * nval;
        float v[64];
        for (int e0 = 0; e0 < nval; e0 += 64) {
          if (valid && !(g.dbg & 256)) {
            if (by_row && g.vec_g) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (e0 + 4 * i < nval) {
                  const float4 x = __ldcs(reinterpret_cast<const float4*>(src + e0 + 4 * i));
                  v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i)
                if (e0 + i < nval) v[i] = __ldcs(src + (size_t)(e0 + i) * (by_row ? 1 : g.La));
            }
          }
          if (e0 == 0 && itk > 0) mbar_wait(&sm->g_empty, (uint32_t)(itk - 1) & 1u);   // every GEMM-A of the previous measure has read Gblk
          // a thread's values are consecutive contraction columns of its own Gblk row (dQ: a row of G, dA: a column
          // of G): 16-byte stores whenever the pair blocks start on a multiple of four columns
          if (valid) {
            if (a_mn) {               // MN-major tile: [32-wide M block][contraction row r][32 M values], 32-byte-atom swizzle
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (e0 + 4 * i < nval) {
                  const int m = col0 + e0 + 4 * i;
                  *reinterpret_cast<float4*>(gblk + (m >> 5) * 16384 + swz128_mn(r, (m & 31) >> 2)) =
                      make_float4(to_tf32(v[4 * i]), to_tf32(v[4 * i + 1]), to_tf32(v[4 * i + 2]), to_tf32(v[4 * i + 3]));
                }
              }
            } else if (g.vec_s) {
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
  if (warp == 1) tmem_dealloc(tmem_O, g.tmem_cols);
  if (TRACE && threadIdx.x == 0) trace_end(tr);
}

// chunk width of U: three buffers behind O must fit the 512 TMEM columns
int chunk_width(int N1) { return N1 + kUBufs * 64 <= 512 ? 64 : (N1 + kUBufs * 32 <= 512 ? 32 : 0); }

}  // namespace

// which = 0: dq (N*Lq x D) from dS, ar = rounded answers, Mr; exports U (mc x N*Lq x Dp) when Uexp != nullptr.
// which = 1: da (N*La x D) from dS, xr = rounded questions, Mr.
// xr (rows x Dp) and Mr (mc x D x Dp) are TF32-rounded copies with 128-byte-aligned rows; dS is read as is.
// ksplit > 1 accumulates the rows of pairs [split_from, N) with atomics: the caller must have zeroed those rows of `out`.  Returns MMS_E_UNSUPPORTED for shapes the
// tile does not cover.
// ksplit > 1: the measures of the pair groups from pair `*split_from` on are spread over ksplit CTAs, which ADD into the
// output -- the caller zeroes the output rows of pairs [*split_from, N) first.  Small batches split every group
// (*split_from = 0); large ones only the groups left over after the last whole wave of the grid.  allow_split = 0
// (deterministic handles): never, every output row has one writer.
int mms_tc_simcross2_backward_fused_plan(int which, int N, int Lq, int La, int D, int mc, int ctas, int allow_split,
                                         int* ksplit, int* split_from) {
  const int sm_count = ctas;
  static const bool disabled = mms_dev_knob("MMS_NO_FUSED") || mms_dev_knob("MMS_NO_FUSED_BWD");
  if (disabled) return MMS_E_UNSUPPORTED;
  if (Lq > 128 || La > 128 || N < 1) return MMS_E_UNSUPPORTED;
  const int N1 = mms_ceil_div(D, 16) * 16;
  if (chunk_width(N1) == 0) return MMS_E_UNSUPPORTED;
  const int P = mms_max(1, mms_min(mms_min(128 / Lq, 128 / La), N));
  const long long base = mms_ceil_div(N, P);
  int ks = 1;
  long long whole = base;
#ifdef MMS_BWD_PROBES
  static const bool no_tail = getenv("MMS_BWD_NO_TAIL_SPLIT") != nullptr;
#else
  const bool no_tail = false;
#endif
  if (allow_split && mc > 1) {
    if (base * 2 <= sm_count) { ks = (int)mms_min<long long>(mc, sm_count / base); whole = 0; }
    else if (!no_tail) {
      const long long rest = base % sm_count;
      if (rest > 0 && rest * 2 <= sm_count) { ks = (int)mms_min<long long>(mc, sm_count / rest); whole = base - rest; }
    }
  }
  if (ks <= 1) { ks = 1; whole = base; }
  *ksplit = ks;
  if (split_from) *split_from = (int)mms_min<long long>(N, whole * P);
  (void)which;
  return 0;
}

int mms_tc_simcross2_backward_fused(mms_context* ctx, int which, const float* xr, const float* Mr, const float* dS,
                                    float* out, float* Uexp, int N, int Lq, int La, int D, int mc, int Dp,
                                    int ksplit, int split_from, int u_blocked) {
  BwdGeom g;
  {
    int unused = 1;
    MMS_TRY(mms_tc_simcross2_backward_fused_plan(which, N, Lq, La, D, mc, ctx->sm_count, 0, &unused, nullptr));
    MMS_REQUIRE(ksplit >= 1 && ksplit <= mc, MMS_E_INVALID, "bad measure split");
  }
  const bool DA = which != 0;
  g.N = N; g.Lq = Lq; g.La = La; g.D = D; g.mc = mc;
  g.Lr = DA ? La : Lq; g.Lk = DA ? Lq : La;
  g.P = mms_max(1, mms_min(mms_min(128 / Lq, 128 / La), N));
  g.nksA = mms_ceil_div(g.P * g.Lk, 8);
  g.N1 = mms_ceil_div(D, 16) * 16;
  g.np0 = g.N1 <= 256 ? g.N1 : ((g.N1 / 2 + 31) & ~31);
  g.np1 = g.N1 - g.np0;
  g.CW = chunk_width(g.N1);
  g.nch = mms_ceil_div(g.N1, g.CW);
  g.wl = g.N1 - (g.nch - 1) * g.CW;
  g.nbxB = mms_ceil_div(g.N1, 32);
  const int a_bytes = mms_ceil_div(g.CW, 32) * 16384;
  const int b_bytes = DA ? g.nbxB * 4096 : (g.np1 > 0 ? 2 : 1) * g.np0 * 128;
  g.stage_bytes = mms_ceil_div(mms_max(a_bytes, b_bytes), 1024) * 1024;
  int stages = kMaxStages;
  while (stages > 2 && (size_t)kGblkBytes + (size_t)stages * g.stage_bytes + sizeof(BwdSmem) + 1024 > 227 * 1024) --stages;
  if ((size_t)kGblkBytes + (size_t)stages * g.stage_bytes + sizeof(BwdSmem) + 1024 > 227 * 1024) return MMS_E_UNSUPPORTED;
#ifdef MMS_BWD_PROBES
  { static const char* e = getenv("MMS_BWD_STAGES"); if (e && atoi(e) >= 2 && atoi(e) < stages) stages = atoi(e); }
#endif
  g.stages = stages;
  g.ksplit = ksplit;
  const long long groups = mms_ceil_div(N, g.P);
  MMS_REQUIRE(ksplit == 1 || (split_from >= 0 && split_from % g.P == 0 && split_from <= N), MMS_E_INVALID, "bad split start");
  const long long whole = ksplit == 1 ? groups : mms_min<long long>(groups, split_from / g.P);
  const long long total = whole + (groups - whole) * ksplit;
  if (total > 0x7fffffffLL) return MMS_E_UNSUPPORTED;
  g.total_tiles = (unsigned)total;
  g.whole = (unsigned)whole;
  g.tmem_cols = umma::tmem_cols_pow2((uint32_t)(g.N1 + kUBufs * g.CW));
  g.vec_g = (La % 4 == 0) && ((reinterpret_cast<uintptr_t>(dS) & 15) == 0);
  g.vec_s = g.Lk % 4 == 0;
  g.a_mn = DA && g.vec_g && g.Lr % 4 == 0;        // rows of G by 16-byte loads, 16-byte stores along M
#ifdef MMS_BWD_PROBES
  { static const char* e = getenv("MMS_BWD_DA_KMAJOR"); if (e && atoi(e)) g.a_mn = 0; }
#endif
  g.vec_out = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  g.Dp = Dp;
  g.dbg = 0;
#ifdef MMS_BWD_PROBES   // timing probes that skip parts of the kernel (WRONG results): MMS_NVCC_EXTRA=-DMMS_BWD_PROBES builds only
  { static const char* e = getenv("MMS_BWD_DEBUG"); g.dbg = e ? atoi(e) : 0; }
#endif
  g.u_rows = (long long)N * g.Lr;
  g.u_blocked = u_blocked;
  g.u_inline = 1;                           // CW <= 64: one piece per rounding warp
#ifdef MMS_BWD_PROBES
  { static const char* e = getenv("MMS_BWD_EXPORT_WARPS"); if (e && atoi(e)) g.u_inline = 0; }
#endif
  g.u_groups = (g.u_rows + 31) / 32;

  CUtensorMap mapX, mapM;
  MMS_TRY(mms_tc_make_map(ctx, &mapX, xr, Dp, true, D, (long long)N * g.Lk, 32, 0, 0, 0, 1, 1, 1, 128));
  if (!DA) MMS_TRY(mms_tc_make_map(ctx, &mapM, Mr, Dp, false, D, D, g.np0, (long long)D * Dp, 0, 0, mc, 1, 1));
  else MMS_TRY(mms_tc_make_map(ctx, &mapM, Mr, Dp, true, D, D, 32, (long long)D * Dp, 0, 0, mc, 1, 1));

  static bool configured = false;
  if (!configured) {
    MMS_MAX_SMEM((simcross2_bwd_fused_kernel<false, false>), 227 * 1024);
    MMS_MAX_SMEM((simcross2_bwd_fused_kernel<true, false>), 227 * 1024);
    MMS_MAX_SMEM((simcross2_bwd_fused_kernel<false, true>), 227 * 1024);
    MMS_MAX_SMEM((simcross2_bwd_fused_kernel<true, true>), 227 * 1024);
    configured = true;
  }
  const size_t smem = (size_t)kGblkBytes + (size_t)stages * g.stage_bytes + sizeof(BwdSmem) + 1024;
  unsigned grid = (unsigned)mms_min<long long>(total, ctx->sm_count);
#ifdef MMS_BWD_PROBES
  { static const char* e = getenv("MMS_BWD_GRID"); if (e && atoi(e) > 0) grid = mms_min<unsigned>(grid, (unsigned)atoi(e)); }
#endif
  TraceBuf tb;
  MMS_TRY(tb.begin(grid));
  { MmsKernelScope ks_(ctx, DA ? "simcross2_bwd_fused_kernel<dA>" : "simcross2_bwd_fused_kernel<dQ>");
    if (tb.dev) {
      if (!DA) simcross2_bwd_fused_kernel<false, true><<<grid, kThreads, smem, ctx->stream>>>(mapX, mapM, dS, out, Uexp, g, tb.dev);
      else simcross2_bwd_fused_kernel<true, true><<<grid, kThreads, smem, ctx->stream>>>(mapX, mapM, dS, out, nullptr, g, tb.dev);
    } else {
      if (!DA) simcross2_bwd_fused_kernel<false, false><<<grid, kThreads, smem, ctx->stream>>>(mapX, mapM, dS, out, Uexp, g, nullptr);
      else simcross2_bwd_fused_kernel<true, false><<<grid, kThreads, smem, ctx->stream>>>(mapX, mapM, dS, out, nullptr, g, nullptr);
    } }
  MMS_LAUNCH_CHECK();
  static const char* const names[kTraceSlots] = {"entry", "setup", "tma_issued", nullptr, "prologue_issued", "u_full",
                                                 "rounded", nullptr, "o_full", "epi0_done", "epi_done", "exit",
                                                 nullptr, nullptr, nullptr, nullptr,
                                                 "W:g_full", "W:fullA", "W:o_empty", "W:fullB", "W:u_ready", "W:u_free", nullptr,
                                                 nullptr};
  char what[128];
  snprintf(what, sizeof(what), "bwd %s N %d L %dx%d D %d mc %d P %d CW %d nch %d ksplit %d tiles %u stages %d", DA ? "dA" : "dQ",
           N, Lq, La, D, mc, g.P, g.CW, g.nch, ksplit, g.total_tiles, stages);
  MMS_TRY(tb.end(ctx, what, names));
  return 0;
}
