// TMA-fed TF32 GEMM on tcgen05 (sm_100a): same contract as tc_gemm.cu, for operands that already
// hold TF32-exact values (rounded by the producing kernel) and whose strides are 16-byte multiples.
//
//   warp 0     producer: one lane issues cp.async.bulk.tensor (5-D tiled TMA, 128-byte swizzle) for
//              every stage as soon as its ring slot is free -- no registers, no generic-proxy
//              stores, `stages` loads in flight per SM; out-of-range rows / k are zero-filled by TMA
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer (accumulator double-buffered)
//   warps 2-5  epilogue: tcgen05.ld -> registers -> global (store / += / atomicAdd, fused bias,
//              optional TF32 rounding of results that feed the next contraction)
// K-major operands are one [rows x 32 k] box per stage (CU_TENSOR_MAP_SWIZZLE_128B); MN-major
// operands are [32 k x 32 mn] boxes (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, the only MN-major layout
// tcgen05 takes for 32-bit types).  Batch indices and reduction segments are tensor-map dimensions.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <map>
#include <tuple>
#include <vector>

#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kBM = 128;
constexpr int kBK = 32;
constexpr int kEpiWarps = 4;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr int kMaxStages = 6;

struct Smem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

constexpr int kEpiLd = 36;                                   // floats per staged row (32 + 4: conflict-free 16-byte accesses)
constexpr int kEpiBytes = kEpiWarps * 32 * kEpiLd * 4;        // one 32 x 32 transposition buffer per epilogue warp

struct Geometry {
  int BN, stages, b_bytes, n_tiles, m_tiles;
  int m_fast;            // tile order: 1 = m-tile index fastest (few row panels, many column panels), 0 = n-tile fastest
  unsigned total_tiles;
  uint32_t tmem_cols;
  // 0/1 multipliers: a broadcast (stride 0) batch / segment dimension has extent 1 in the tensor map
  int a_z1, a_z2, a_seg, b_z1, b_z2, b_seg;
};

struct Tile {
  int z1, z2, m0, n0, ibeg, nk;
};

// Optional per-CTA trace (MMS_TC_TRACE=1 in the environment): globaltimer stamps at 8 checkpoints.
__device__ __forceinline__ void trace(long long* tr, int slot) {
  if (!tr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  tr[(size_t)blockIdx.x * 8 + slot] = (long long)t;
}

__device__ __forceinline__ Tile decode_tile(const TcGemmArgs& g, const Geometry& q, unsigned t, int sps, int bm = kBM) {
  Tile tl;
  int n_tile, m_tile;
  if (q.m_fast) { m_tile = t % q.m_tiles; t /= q.m_tiles; n_tile = t % q.n_tiles; t /= q.n_tiles; }
  else { n_tile = t % q.n_tiles; t /= q.n_tiles; m_tile = t % q.m_tiles; t /= q.m_tiles; }
  const int split = t % g.ksplit;
  const int z = t / g.ksplit;
  tl.z1 = z / g.nb2; tl.z2 = z % g.nb2;
  tl.m0 = m_tile * bm; tl.n0 = n_tile * q.BN;
  const int total_stages = g.nseg * sps;
  const int per_split = (total_stages + g.ksplit - 1) / g.ksplit;
  tl.ibeg = split * per_split;
  tl.nk = max(0, min(total_stages, tl.ibeg + per_split) - tl.ibeg);
  return tl;
}

// Threshold epilogue (TcGemmArgs::flt_*): this thread's accumulator row against the row's current k-th list entry.
__device__ __forceinline__ void epi_filter(const TcGemmArgs& g, const float (&v)[32], int row, int col0, int ncols) {
  if (row >= g.M) return;
  const float ts = __ldg(g.flt_s + (size_t)row * g.flt_k + g.flt_k - 1);
  const long long ti = __ldg(g.flt_i + (size_t)row * g.flt_k + g.flt_k - 1);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float x = v[i];
    if (i < ncols && x >= ts) {                            // almost never: one compare per value on the common path
      const long long gi = g.flt_base + col0 + i;
      if (x > ts || gi < ti) {
        const int pos = atomicAdd(g.cand_n + row, 1);
        if (pos < g.cand_cap) { g.cand_s[(size_t)row * g.cand_cap + pos] = x; g.cand_i[(size_t)row * g.cand_cap + pos] = gi; }
      }
    }
  }
}

// Epilogue store of one 32 x 32 chunk that a warp has staged in shared memory.  Thread (lane) owns the
// 4 columns cc = 4*(lane%8) of rows r0 + 4*rr, r0 = lane/8: every warp instruction covers 4 whole
// 128-byte row segments.  MODE is resolved outside the row loop so that the loop body is a handful
// of instructions (the epilogue is instruction-bound, not bandwidth-bound, at these tile counts).
template <int MODE, bool ROUND, bool ADD>
__device__ __forceinline__ void epi_store_vec(const float* sp, float* p, long long pstep, const float* ap,
                                              long long astep, bool add_vec, int rows_left) {
  float4 o[8];
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) o[rr] = *reinterpret_cast<const float4*>(sp + rr * 4 * kEpiLd);
  if (ADD) {
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      if (rr * 4 < rows_left) {
        const float* a = ap + rr * astep;
        float4 b;
        if (add_vec) b = __ldg(reinterpret_cast<const float4*>(a));
        else b = make_float4(__ldg(a), __ldg(a + 1), __ldg(a + 2), __ldg(a + 3));
        o[rr].x += b.x; o[rr].y += b.y; o[rr].z += b.z; o[rr].w += b.w;
      }
    }
  }
  if (ROUND) {
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      o[rr].x = to_tf32(o[rr].x); o[rr].y = to_tf32(o[rr].y); o[rr].z = to_tf32(o[rr].z); o[rr].w = to_tf32(o[rr].w);
    }
  }
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    float4* dst = reinterpret_cast<float4*>(p + rr * pstep);
    if (rr * 4 < rows_left) {
      if (MODE == TC_STORE) {
        *dst = o[rr];
      } else if (MODE == TC_ACCUM) {
        float4 c = *dst;
        c.x += o[rr].x; c.y += o[rr].y; c.z += o[rr].z; c.w += o[rr].w;
        *dst = c;
      } else {
        atomicAdd(dst, o[rr]);
      }
    }
  }
}

// Ragged / unaligned tiles: element-wise, any mode.
__device__ __noinline__ void epi_store_scalar(const float* sp, float* p, long long pstep, const float* ap,
                                              long long astep, bool round_out, int rows_left, int nvalid, int mode) {
  for (int rr = 0; rr * 4 < rows_left; ++rr) {
    for (int j = 0; j < nvalid; ++j) {
      float o = sp[rr * 4 * kEpiLd + j];
      if (ap) o += __ldg(ap + rr * astep + j);
      if (round_out) o = to_tf32(o);
      float* d = p + rr * pstep + j;
      if (mode == TC_STORE) *d = o;
      else if (mode == TC_ACCUM) *d += o;
      else atomicAdd(d, o);
    }
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const TcGemmArgs g, const Geometry q, long long* tr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET from the __shared__ symbol (not by integer arithmetic on the pointer), so the
  // compiler keeps the shared address space and emits STS/LDS instead of generic ST/LD for everything derived from it
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = 16384 + q.b_bytes;
  float* epi = reinterpret_cast<float*>(ring + q.stages * stage_bytes);
  Smem* sm = reinterpret_cast<Smem*>(ring + q.stages * stage_bytes + kEpiBytes);

  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  const int sps = (g.K + kBK - 1) / kBK;
  const int BN = q.BN, stages = q.stages;
  if (threadIdx.x == 0) trace(tr, 0);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], kEpiWarps); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, q.tmem_cols);
    tmem_relinquish();
  } else if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  if (threadIdx.x == 0) trace(tr, 1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: the whole warp walks the
    {                                                            // loops, one elected lane issues (umma.cuh)
      const uint32_t tx_bytes = (uint32_t)stage_bytes;
      const int b_blocks = (BN + 31) >> 5;
      int it = 0;
      for (unsigned t = blockIdx.x; t < q.total_tiles; t += gridDim.x) {
        const Tile tl = decode_tile(g, q, t, sps);
        const int az1 = tl.z1 * q.a_z1, az2 = tl.z2 * q.a_z2;
        const int bz1 = tl.z1 * q.b_z1, bz2 = tl.z2 * q.b_z2;
        for (int i = 0; i < tl.nk; ++i, ++it) {
          const int s = it % stages;
          const int gi = tl.ibeg + i;
          const int seg = gi / sps;
          const int k0 = (gi - seg * sps) * kBK;
          if (it >= stages) mbar_wait(&sm->empty[s], ((it / stages) - 1) & 1);
          uint8_t* a_dst = ring + s * stage_bytes;
          uint8_t* b_dst = a_dst + 16384;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&sm->full[s], tx_bytes);
            if (!A_MN) {
              tma_load_5d(a_dst, &mapA, &sm->full[s], k0, tl.m0, az2, az1, seg * q.a_seg);
            } else {
#pragma unroll
              for (int b = 0; b < kBM / 32; ++b)
                tma_load_5d(a_dst + b * 4096, &mapA, &sm->full[s], tl.m0 + 32 * b, k0, az2, az1, seg * q.a_seg);
            }
            if (!B_MN) {
              tma_load_5d(b_dst, &mapB, &sm->full[s], k0, tl.n0, bz2, bz1, seg * q.b_seg);
            } else {
              for (int b = 0; b < b_blocks; ++b)
                tma_load_5d(b_dst + b * 4096, &mapB, &sm->full[s], tl.n0 + 32 * b, k0, bz2, bz1, seg * q.b_seg);
            }
          }
          __syncwarp();
        }
      }
      if (lane == 0) trace(tr, 2);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue (whole warp, one elected lane)
    {
      const uint32_t idesc = idesc_tf32(kBM, BN, A_MN, B_MN);
      int it = 0, tcount = 0;
      for (unsigned t = blockIdx.x; t < q.total_tiles; t += gridDim.x, ++tcount) {
        const Tile tl = decode_tile(g, q, t, sps);
        const int buf = tcount & 1;
        if (tcount >= 2) mbar_wait(&sm->acc_empty[buf], ((tcount >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t acc = tmem + buf * BN;
        for (int i = 0; i < tl.nk; ++i, ++it) {
          const int s = it % stages;
          mbar_wait(&sm->full[s], (it / stages) & 1);
          tc_fence_after();
          if (it == 0 && lane == 0) trace(tr, 3);
          const uint32_t a_base = smem_u32(ring + s * stage_bytes);
          const uint32_t a_lo = A_MN ? desc_lo_mn(a_base, 4096) : desc_lo_k(a_base);
          const uint32_t b_lo = B_MN ? desc_lo_mn(a_base + 16384, 4096) : desc_lo_k(a_base + 16384);
          if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < kBK / 8; ++ks)
              mma_tf32_ss_lh(acc, a_lo + ks * (A_MN ? kDescStepMN : kDescStepK), A_MN ? kDescHiMN : kDescHiK,
                             b_lo + ks * (B_MN ? kDescStepMN : kDescStepK), B_MN ? kDescHiMN : kDescHiK, idesc,
                             (i > 0 || ks > 0) ? 1u : 0u);
            mma_commit(&sm->empty[s]);
            if (i == tl.nk - 1) mma_commit(&sm->acc_full[buf]);
          }
          __syncwarp();
        }
        if (tl.nk == 0) {                                        // nothing to reduce: hand over an untouched accumulator
          if (elect_one_sync()) mma_commit(&sm->acc_full[buf]);
          __syncwarp();
        }
      }
      if (lane == 0) trace(tr, 4);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (TMEM lane quarter = warp % 4)
    const int quarter = warp & 3;
    // which specialised store loop serves this launch (-1: only the element-wise one)
    int fast_kind = -1;
    if (g.mode == TC_STORE) fast_kind = (g.c_add && g.round_out) ? -1 : g.c_add ? 2 : g.round_out ? 1 : 0;
    else if (!g.c_add && !g.round_out) fast_kind = g.mode == TC_ATOMIC ? 3 : 4;
    int tcount = 0;
    for (unsigned t = blockIdx.x; t < q.total_tiles; t += gridDim.x, ++tcount) {
      const Tile tl = decode_tile(g, q, t, sps);
      const int buf = tcount & 1;
      float* C = g.C + tl.z1 * g.sC1 + tl.z2 * g.sC2;
      const float* Cadd = g.c_add ? g.c_add + tl.z1 * g.s_add1 + tl.z2 * g.s_add2 : nullptr;
      mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1);
      tc_fence_after();
      if (tcount == 0 && threadIdx.x == 64) trace(tr, 5);
      const bool c_vec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (g.ldc % 4 == 0) && (tl.n0 % 4 == 0);
      const bool add_vec = Cadd && ((reinterpret_cast<uintptr_t>(Cadd) & 15) == 0) && (g.ld_add % 4 == 0);
      const int row_tm = tl.m0 + quarter * 32 + lane;           // the accumulator row this thread reads from TMEM
      const float rs = (g.out_rowscale && row_tm < g.M) ? __ldg(g.out_rowscale + row_tm) : 1.f;
      const uint32_t acc = tmem + buf * BN + ((uint32_t)(quarter * 32) << 16);
      const int ncols = min(BN, g.N - tl.n0);
      float* stage = epi + quarter * (32 * kEpiLd);
      // rows of this warp that exist: a warp that owns none skips its TMEM reads altogether
      const int warp_rows = min(32, g.M - (tl.m0 + quarter * 32));
      for (int c0 = 0; warp_rows > 0 && c0 < ncols; c0 += 32) {
        float v[32];
        if (tl.nk > 0) {
          if (c0 + 16 < BN) tmem_ld32(acc + c0, v);          // BN is a multiple of 16
          else tmem_ld16(acc + c0, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (g.flt_s) {                                         // threshold epilogue: nothing is stored
          epi_filter(g, v, row_tm, tl.n0 + c0, min(32, ncols - c0));
          continue;
        }
        // transpose through shared memory: TMEM hands every thread one ROW (32 columns); global memory
        // wants every warp instruction to cover whole 128-byte row segments
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4)
          *reinterpret_cast<float4*>(stage + lane * kEpiLd + i4 * 4) =
              make_float4(v[i4 * 4] * rs, v[i4 * 4 + 1] * rs, v[i4 * 4 + 2] * rs, v[i4 * 4 + 3] * rs);
        __syncwarp();
        const int cc = (lane & 7) * 4;                       // 4 columns of this thread
        const int r0 = lane >> 3;                            // first of its 8 rows (stride 4)
        const int n = tl.n0 + c0 + cc;
        const int nvalid = (c0 + cc < BN) ? max(0, min(4, g.N - n)) : 0;
        const int rows_left = warp_rows - r0;
        const long long row0 = tl.m0 + quarter * 32 + r0;
        float* p = C + row0 * g.ldc + n;
        const float* ap = Cadd ? Cadd + row0 * g.ld_add + n : nullptr;
        const float* sp = stage + r0 * kEpiLd + cc;
        const long long pstep = 4 * g.ldc, astep = 4 * g.ld_add;
        if (nvalid == 4 && c_vec && fast_kind >= 0) {
          switch (fast_kind) {
            case 0: epi_store_vec<TC_STORE, false, false>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
            case 1: epi_store_vec<TC_STORE, true, false>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
            case 2: epi_store_vec<TC_STORE, false, true>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
            case 3: epi_store_vec<TC_ATOMIC, false, false>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
            default: epi_store_vec<TC_ACCUM, false, false>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
          }
        } else if (nvalid > 0 && rows_left > 0) {
          epi_store_scalar(sp, p, pstep, ap, astep, g.round_out != 0, rows_left, nvalid, g.mode);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->acc_empty[buf]);
    }
    if (threadIdx.x == 64) trace(tr, 6);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, q.tmem_cols);
  if (threadIdx.x == 0) trace(tr, 7);
}

// ---- CTA-pair variant: 256 x BN tiles on two SMs (tcgen05 cta_group::2) -----------------------------------
// Same contract and the same roles as tc_gemm_tma_kernel, launched as clusters of two CTAs.  Each CTA stages its own
// 128 rows of A and HALF of the B tile (16 KB + BN * 64 B per k-block instead of 16 KB + BN * 128 B), keeps its
// 128 x BN half of the accumulator in its own TMEM and stores it; the leader CTA (rank 0) issues one M = 256 MMA
// per k-step for both.  What it buys: the 128 x 256 tile of the single-CTA kernel needs 94 B of TMA ingest per
// tensor-pipe cycle per SM and runs at the ~55-60 B/clk the SMs sustain (candidate scoring: 447 TFLOP/s); the pair
// needs 64.
//   full[s]      lives in the leader; both producers arrive on it (expect_tx of their own bytes) and both CTAs'
//                TMA loads complete_tx on it
//   empty[s], acc_full[b]   per CTA; the leader's tcgen05.commit multicasts its arrival to both
//   acc_empty[b] lives in the leader; the epilogue warps of both CTAs arrive on it
constexpr int kBM2 = 256;

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_gemm_tma2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const TcGemmArgs g, const Geometry q) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int BN = q.BN, stages = q.stages, BNh = BN >> 1;
  const int stage_bytes = 16384 + q.b_bytes;                     // q.b_bytes: this CTA's half of the B tile
  float* epi = reinterpret_cast<float*>(ring + stages * stage_bytes);
  Smem* sm = reinterpret_cast<Smem*>(ring + stages * stage_bytes + kEpiBytes);

  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int sps = (g.K + kBK - 1) / kBK;
  const unsigned cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) { mbar_init(&sm->full[s], 2); mbar_init(&sm->empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], 2 * kEpiWarps); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2cta(&sm->tmem_base, q.tmem_cols);
    tmem_relinquish_2cta();
  } else if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                            // the peer's barriers exist before anyone signals them
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t tx_bytes = (uint32_t)stage_bytes;
    const int b_blocks = (BNh + 31) >> 5;
    int it = 0;
    for (unsigned t = cluster_id; t < q.total_tiles; t += n_clusters) {
      const Tile tl = decode_tile(g, q, t, sps, kBM2);
      const int az1 = tl.z1 * q.a_z1, az2 = tl.z2 * q.a_z2;
      const int bz1 = tl.z1 * q.b_z1, bz2 = tl.z2 * q.b_z2;
      const int m0 = tl.m0 + (int)rank * kBM, n0 = tl.n0 + (int)rank * BNh;
      for (int i = 0; i < tl.nk; ++i, ++it) {
        const int s = it % stages;
        const int gi = tl.ibeg + i;
        const int seg = gi / sps;
        const int k0 = (gi - seg * sps) * kBK;
        if (it >= stages) mbar_wait(&sm->empty[s], ((it / stages) - 1) & 1);
        uint8_t* a_dst = ring + s * stage_bytes;
        uint8_t* b_dst = a_dst + 16384;
        const uint32_t full_leader = mapa_shared(smem_u32(&sm->full[s]), 0);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx_cluster(full_leader, tx_bytes);
          if (!A_MN) {
            tma_load_5d_2cta(a_dst, &mapA, full_leader, k0, m0, az2, az1, seg * q.a_seg);
          } else {
#pragma unroll
            for (int b = 0; b < kBM / 32; ++b)
              tma_load_5d_2cta(a_dst + b * 4096, &mapA, full_leader, m0 + 32 * b, k0, az2, az1, seg * q.a_seg);
          }
          if (!B_MN) {
            tma_load_5d_2cta(b_dst, &mapB, full_leader, k0, n0, bz2, bz1, seg * q.b_seg);
          } else {
            for (int b = 0; b < b_blocks; ++b)
              tma_load_5d_2cta(b_dst + b * 4096, &mapB, full_leader, n0 + 32 * b, k0, bz2, bz1, seg * q.b_seg);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue: the leader, for both CTAs
    if (leader) {
      const uint32_t idesc = idesc_tf32(kBM2, BN, A_MN, B_MN);
      const uint32_t ring_base = smem_u32(ring);
      const uint32_t ring_a = A_MN ? desc_lo_mn(ring_base, 4096) : desc_lo_k(ring_base);
      const uint32_t ring_b = B_MN ? desc_lo_mn(ring_base + 16384, 4096) : desc_lo_k(ring_base + 16384);
      const uint32_t stage_lo = (uint32_t)stage_bytes >> 4;
      constexpr uint32_t a_step = A_MN ? kDescStepMN : kDescStepK, a_hi = A_MN ? kDescHiMN : kDescHiK;
      constexpr uint32_t b_step = B_MN ? kDescStepMN : kDescStepK, b_hi = B_MN ? kDescHiMN : kDescHiK;
      int it = 0, tcount = 0;
      for (unsigned t = cluster_id; t < q.total_tiles; t += n_clusters, ++tcount) {
        const Tile tl = decode_tile(g, q, t, sps, kBM2);
        const int buf = tcount & 1;
        if (tcount >= 2) mbar_wait(&sm->acc_empty[buf], ((tcount >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t acc = tmem + buf * BN;
        for (int i = 0; i < tl.nk; ++i, ++it) {
          const int s = it % stages;
          mbar_wait(&sm->full[s], (it / stages) & 1);
          tc_fence_after();
          const uint32_t a_lo = ring_a + (uint32_t)s * stage_lo, b_lo = ring_b + (uint32_t)s * stage_lo;
          if (elect_one_sync()) {
            mma_tf32_ss_lh_2cta(acc, a_lo, a_hi, b_lo, b_hi, idesc, i > 0 ? 1u : 0u);
            mma_tf32_ss_lh_2cta(acc, a_lo + 1 * a_step, a_hi, b_lo + 1 * b_step, b_hi, idesc, 1u);
            mma_tf32_ss_lh_2cta(acc, a_lo + 2 * a_step, a_hi, b_lo + 2 * b_step, b_hi, idesc, 1u);
            mma_tf32_ss_lh_2cta(acc, a_lo + 3 * a_step, a_hi, b_lo + 3 * b_step, b_hi, idesc, 1u);
            mma_commit_2cta(&sm->empty[s]);
            if (i == tl.nk - 1) mma_commit_2cta(&sm->acc_full[buf]);
          }
          __syncwarp();
        }
        if (tl.nk == 0) {
          if (elect_one_sync()) mma_commit_2cta(&sm->acc_full[buf]);
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: each CTA its own 128 rows
    const int quarter = warp & 3;
    int fast_kind = -1;
    if (g.mode == TC_STORE) fast_kind = (g.c_add && g.round_out) ? -1 : g.c_add ? 2 : g.round_out ? 1 : 0;
    else if (!g.c_add && !g.round_out) fast_kind = g.mode == TC_ATOMIC ? 3 : 4;
    int tcount = 0;
    for (unsigned t = cluster_id; t < q.total_tiles; t += n_clusters, ++tcount) {
      const Tile tl = decode_tile(g, q, t, sps, kBM2);
      const int m0 = tl.m0 + (int)rank * kBM;
      const int buf = tcount & 1;
      float* C = g.C + tl.z1 * g.sC1 + tl.z2 * g.sC2;
      const float* Cadd = g.c_add ? g.c_add + tl.z1 * g.s_add1 + tl.z2 * g.s_add2 : nullptr;
      mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const bool c_vec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (g.ldc % 4 == 0) && (tl.n0 % 4 == 0);
      const bool add_vec = Cadd && ((reinterpret_cast<uintptr_t>(Cadd) & 15) == 0) && (g.ld_add % 4 == 0);
      const int row_tm = m0 + quarter * 32 + lane;
      const float rs = (g.out_rowscale && row_tm < g.M) ? __ldg(g.out_rowscale + row_tm) : 1.f;
      const uint32_t acc = tmem + buf * BN + ((uint32_t)(quarter * 32) << 16);
      const int ncols = min(BN, g.N - tl.n0);
      float* stage = epi + quarter * (32 * kEpiLd);
      const int warp_rows = min(32, g.M - (m0 + quarter * 32));
      for (int c0 = 0; warp_rows > 0 && c0 < ncols; c0 += 32) {
        float v[32];
        if (tl.nk > 0) {
          if (c0 + 16 < BN) tmem_ld32(acc + c0, v);
          else tmem_ld16(acc + c0, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (g.flt_s) {                                         // threshold epilogue: nothing is stored
          epi_filter(g, v, row_tm, tl.n0 + c0, min(32, ncols - c0));
          continue;
        }
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4)
          *reinterpret_cast<float4*>(stage + lane * kEpiLd + i4 * 4) =
              make_float4(v[i4 * 4] * rs, v[i4 * 4 + 1] * rs, v[i4 * 4 + 2] * rs, v[i4 * 4 + 3] * rs);
        __syncwarp();
        const int cc = (lane & 7) * 4;
        const int r0 = lane >> 3;
        const int n = tl.n0 + c0 + cc;
        const int nvalid = (c0 + cc < BN) ? max(0, min(4, g.N - n)) : 0;
        const int rows_left = warp_rows - r0;
        const long long row0 = m0 + quarter * 32 + r0;
        float* p = C + row0 * g.ldc + n;
        const float* ap = Cadd ? Cadd + row0 * g.ld_add + n : nullptr;
        const float* sp = stage + r0 * kEpiLd + cc;
        const long long pstep = 4 * g.ldc, astep = 4 * g.ld_add;
        if (nvalid == 4 && c_vec && fast_kind >= 0) {
          switch (fast_kind) {
            case 0: epi_store_vec<TC_STORE, false, false>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
            case 1: epi_store_vec<TC_STORE, true, false>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
            case 2: epi_store_vec<TC_STORE, false, true>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
            case 3: epi_store_vec<TC_ATOMIC, false, false>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
            default: epi_store_vec<TC_ACCUM, false, false>(sp, p, pstep, ap, astep, add_vec, rows_left); break;
          }
        } else if (nvalid > 0 && rows_left > 0) {
          epi_store_scalar(sp, p, pstep, ap, astep, g.round_out != 0, rows_left, nvalid, g.mode);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&sm->acc_empty[buf]), 0));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                            // no CTA frees TMEM or exits while its peer still works
  if (warp == 1) tmem_dealloc_2cta(tmem, q.tmem_cols);
}

// ---- tf32 rounding / repacking pass --------------------------------------------------------
struct RoundJobs {
  RoundJob j[4];
};

__global__ void __launch_bounds__(256) tf32_round_kernel(const RoundJobs jobs) {
  const RoundJob jb = jobs.j[blockIdx.y];
  const bool vec = (jb.cols % 4 == 0) && (jb.lds % 4 == 0) && (jb.ldd % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(jb.src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(jb.dst) & 15) == 0);
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const int c4n = jb.cols >> 2;
    const long long total = jb.rows * c4n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
      const long long r = e / c4n;
      const int c = (int)(e - r * c4n) * 4;
      float4 v = __ldg(reinterpret_cast<const float4*>(jb.src + r * jb.lds + c));
      const float s = jb.scale ? __ldg(jb.scale + r) : 1.f;
      v.x = to_tf32(v.x * s); v.y = to_tf32(v.y * s); v.z = to_tf32(v.z * s); v.w = to_tf32(v.w * s);
      *reinterpret_cast<float4*>(jb.dst + r * jb.ldd + c) = v;
    }
  } else {
    const long long total = jb.rows * jb.cols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
      const long long r = e / jb.cols;
      const int c = (int)(e - r * jb.cols);
      const float s = jb.scale ? __ldg(jb.scale + r) : 1.f;
      jb.dst[r * jb.ldd + c] = to_tf32(__ldg(jb.src + r * jb.lds + c) * s);
    }
  }
}

// ---- host: tensor maps -------------------------------------------------------------------
typedef std::tuple<const void*, long long, int, long long, long long, int, long long, long long, long long, int,
                   int, int>
    MapKey;

struct TcState {
  PFN_cuTensorMapEncodeTiled encode = nullptr;
  std::map<MapKey, CUtensorMap> maps;
};

TcState* state_of(mms_context* ctx) {
  if (!ctx->tc_state) ctx->tc_state = new TcState();
  return static_cast<TcState*>(ctx->tc_state);
}

// Operand X(mn, k): K-major X[mn*ld + k] or MN-major X[k*ld + mn]; batch strides s1, s2 and segment
// stride sseg in elements (0 = broadcast).  rows_box: rows per K-major box.
int make_map(mms_context* ctx, CUtensorMap* out, const float* ptr, long long ld, bool mn_major, long long MN,
             long long K, int rows_box, long long s1, long long s2, long long sseg, int nb1, int nb2, int nseg,
             int k_box = 32) {
  TcState* st = state_of(ctx);
  if (!st->encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MMS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MMS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, MMS_E_UNSUPPORTED, "cuTensorMapEncodeTiled unavailable");
    st->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  }
  const MapKey key(ptr, ld, mn_major ? k_box : 0, MN, K, rows_box, s1, s2, sseg, nb1, nb2, nseg);
  auto hit = st->maps.find(key);
  if (hit != st->maps.end()) { *out = hit->second; return 0; }
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5] = {32, 32, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  dims[0] = (cuuint64_t)(mn_major ? MN : K);
  dims[1] = (cuuint64_t)(mn_major ? K : MN);
  box[1] = (cuuint32_t)(mn_major ? k_box : rows_box);
  strides[0] = (cuuint64_t)ld * 4;
  // broadcast / singleton dimensions get extent 1; their stride only has to be a legal value
  const cuuint64_t filler = strides[0] * dims[1];
  dims[2] = s2 ? (cuuint64_t)nb2 : 1;   strides[1] = s2 ? (cuuint64_t)s2 * 4 : filler;
  dims[3] = s1 ? (cuuint64_t)nb1 : 1;   strides[2] = s1 ? (cuuint64_t)s1 * 4 : filler;
  dims[4] = sseg ? (cuuint64_t)nseg : 1; strides[3] = sseg ? (cuuint64_t)sseg * 4 : filler;
  CUtensorMap m;
  const CUresult r = st->encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(ptr), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mms_set_error("cuTensorMapEncodeTiled failed with CUresult %d (ld %lld, MN %lld, K %lld)", (int)r, ld, MN, K);
    return MMS_E_UNSUPPORTED;
  }
  if (st->maps.size() > 256) st->maps.clear();
  st->maps[key] = m;
  *out = m;
  return 0;
}

bool tma_ok(const float* p, long long ld, long long s1, long long s2, long long sseg) {
  return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0) && (s1 % 4 == 0) && (s2 % 4 == 0) &&
         (sseg % 4 == 0) && ld > 0;
}

}  // namespace

int mms_tc_make_map(mms_context* ctx, void* out, const float* ptr, long long ld, bool mn_major, long long MN,
                    long long K, int rows_box, long long s1, long long s2, long long sseg, int nb1, int nb2,
                    int nseg, int k_box) {
  return make_map(ctx, static_cast<CUtensorMap*>(out), ptr, ld, mn_major, MN, K, rows_box, s1, s2, sseg, nb1, nb2,
                  nseg, k_box);
}

// A tensor map given dimension by dimension (fp32 elements, strides in bytes for dimensions 1..rank-1, 128-byte
// swizzle: the 32-byte-atom variant the MN-major UMMA layout needs, or the plain one).  Not cached.
int mms_tc_make_map_raw(mms_context* ctx, void* out, const float* ptr, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box, bool atom32b) {
  TcState* st = state_of(ctx);
  if (!st->encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MMS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MMS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, MMS_E_UNSUPPORTED, "cuTensorMapEncodeTiled unavailable");
    st->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  }
  MMS_REQUIRE(rank >= 1 && rank <= 5, MMS_E_INVALID, "tensor map rank");
  cuuint64_t d[5], sb[4];
  cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) sb[i] = strides_bytes[i];
  const CUresult r = st->encode(static_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                                const_cast<float*>(ptr), d, sb, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mms_set_error("cuTensorMapEncodeTiled (raw, rank %d) failed with CUresult %d", rank, (int)r);
    return MMS_E_UNSUPPORTED;
  }
  return 0;
}

void mms_tc_destroy_state(mms_context* ctx) {
  delete static_cast<TcState*>(ctx->tc_state);
  ctx->tc_state = nullptr;
}

int mms_tf32_round(mms_context* ctx, const RoundJob* jobs, int njobs) {
  MMS_REQUIRE(jobs && njobs > 0 && njobs <= 4, MMS_E_INVALID, "1..4 rounding jobs per launch");
  RoundJobs js;
  long long biggest = 0;
  for (int i = 0; i < njobs; ++i) {
    js.j[i] = jobs[i];
    biggest = mms_max(biggest, jobs[i].rows * jobs[i].cols);
  }
  if (biggest == 0) return 0;
  const int gx = (int)mms_min<long long>((biggest / 4 + 255) / 256 + 1, (long long)ctx->sm_count * 8);
  { MmsKernelScope ks_(ctx, "tf32_round_kernel");
    MMS_CARVEOUT(tf32_round_kernel);
    tf32_round_kernel<<<dim3(gx, njobs), 256, 0, ctx->stream>>>(js); }
  MMS_LAUNCH_CHECK();
  return 0;
}

namespace {

// The CTA-pair kernel pays when the tile is wide (the B half it saves is large) and there are enough 256-row tiles
// to fill the 74 pairs; otherwise the single-CTA kernel's smaller tiles spread better.
int gemm_tma_pair(mms_context* ctx, const TcGemmArgs& a) {
  static const bool disabled = mms_dev_knob("MMS_NO_2CTA");
  if (disabled || a.M <= 128 || a.N < 256) return MMS_E_UNSUPPORTED;
  Geometry q;
  // column tiles of equal width, each CTA's half a multiple of 32 columns: N = 300 runs as 2 x 192 (78 % of the MMA
  // columns useful) instead of 256 + 44 (59 %)
  const int BN = mms_min(256, mms_ceil_div(mms_ceil_div(a.N, mms_ceil_div(a.N, 256)), 64) * 64);
  q.BN = BN;
  q.n_tiles = mms_ceil_div(a.N, BN);
  q.m_tiles = mms_ceil_div(a.M, kBM2);
  const long long total = (long long)q.n_tiles * q.m_tiles * a.ksplit * a.nb1 * a.nb2;
  const int pairs = ctx->sm_count / 2;
  // rows wasted by the 256-row tiles vs the 128-row tiles, and enough tiles for every pair
  if (total < pairs || total > 0x7fffffffLL) return MMS_E_UNSUPPORTED;
  if ((long long)q.m_tiles * kBM2 > (long long)mms_ceil_div(a.M, kBM) * kBM + 64) return MMS_E_UNSUPPORTED;
  q.m_fast = (q.m_tiles < q.n_tiles && q.m_tiles <= pairs) ? 1 : 0;
  q.b_bytes = a.b_mn ? mms_ceil_div(BN / 2, 32) * 4096 : (BN / 2) * 128;
  const int stage_bytes = 16384 + q.b_bytes;
  int stages = 8;
  while (stages > 2 && (size_t)stages * stage_bytes + kEpiBytes + sizeof(Smem) + 1024 > 220 * 1024) --stages;
  if (stages > kMaxStages) stages = kMaxStages;
  q.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + kEpiBytes + sizeof(Smem) + 1024;
  q.total_tiles = (unsigned)total;
  q.tmem_cols = umma::tmem_cols_pow2(2 * BN);
  q.a_z1 = a.sA1 != 0; q.a_z2 = a.sA2 != 0; q.a_seg = a.segA != 0;
  q.b_z1 = a.sB1 != 0; q.b_z2 = a.sB2 != 0; q.b_seg = a.segB != 0;
  CUtensorMap mapA, mapB;
  MMS_TRY(make_map(ctx, &mapA, a.A, a.lda, a.a_mn != 0, a.M, a.K, kBM, a.sA1, a.sA2, a.segA, a.nb1, a.nb2, a.nseg));
  MMS_TRY(make_map(ctx, &mapB, a.B, a.ldb, a.b_mn != 0, a.N, a.K, a.b_mn ? 32 : BN / 2, a.sB1, a.sB2, a.segB, a.nb1,
                   a.nb2, a.nseg));
  typedef void (*kernel_t)(const CUtensorMap, const CUtensorMap, const TcGemmArgs, const Geometry);
  static const kernel_t kernels[4] = {tc_gemm_tma2_kernel<false, false>, tc_gemm_tma2_kernel<false, true>,
                                      tc_gemm_tma2_kernel<true, false>, tc_gemm_tma2_kernel<true, true>};
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < 4; ++i)
      MMS_MAX_SMEM(kernels[i], 221 * 1024);
    configured = true;
  }
  const kernel_t kernel = kernels[(a.a_mn ? 2 : 0) + (a.b_mn ? 1 : 0)];
  const unsigned grid = 2u * (unsigned)mms_min<long long>(total, pairs);
  { MmsKernelScope ks_(ctx, "tc_gemm_tma2_kernel");
    kernel<<<grid, kThreads, smem, ctx->stream>>>(mapA, mapB, a, q); }
  MMS_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int mms_tc_gemm_tma(mms_context* ctx, const TcGemmArgs& a) {
  MMS_REQUIRE(a.A && a.B && a.C, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.nb1 > 0 && a.nb2 > 0 && a.ksplit > 0 && a.nseg > 0,
              MMS_E_INVALID, "bad size");
  MMS_REQUIRE(a.ksplit == 1 || a.mode == TC_ATOMIC, MMS_E_INVALID, "split-K needs the atomic epilogue");
  if (a.a_rowscale || a.b_rowscale) return MMS_E_UNSUPPORTED;     // row scaling needs the register pass
  if (!tma_ok(a.A, a.lda, a.sA1, a.sA2, a.segA) || !tma_ok(a.B, a.ldb, a.sB1, a.sB2, a.segB))
    return MMS_E_UNSUPPORTED;
  {
    const int rc = gemm_tma_pair(ctx, a);
    if (rc != MMS_E_UNSUPPORTED) return rc;
  }
  Geometry q;
  const int ntiles = mms_ceil_div(a.N, 256);
  int BN = mms_ceil_div(mms_ceil_div(a.N, ntiles), 16) * 16;
  if (a.b_mn) BN = mms_ceil_div(BN, 32) * 32 > 256 ? BN : mms_ceil_div(BN, 32) * 32;
  q.BN = BN;
  q.n_tiles = mms_ceil_div(a.N, BN);
  q.m_tiles = mms_ceil_div(a.M, kBM);
  // The CTAs that run together take consecutive tiles.  With n fastest they share one row panel of A and stream
  // distinct panels of B; when A has only a few row panels (candidate scoring: 1000 queries x 10^6 candidates) that
  // re-reads all of B from HBM once per row panel -- walk the row panels first instead, so that each panel of B is
  // fetched from HBM once and served to its m_tiles consumers from L2.
  q.m_fast = (q.m_tiles < q.n_tiles && q.m_tiles <= ctx->sm_count) ? 1 : 0;
  q.b_bytes = a.b_mn ? mms_ceil_div(BN, 32) * 4096 : BN * 128;
  const int stage_bytes = 16384 + q.b_bytes;
  int stages = kMaxStages;
  while (stages > 2 && (size_t)stages * stage_bytes + kEpiBytes + sizeof(Smem) + 1024 > 220 * 1024) --stages;
  q.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + kEpiBytes + sizeof(Smem) + 1024;
  const long long total = (long long)q.n_tiles * q.m_tiles * a.ksplit * a.nb1 * a.nb2;
  MMS_REQUIRE(total <= 0x7fffffffLL, MMS_E_UNSUPPORTED, "too many tiles");
  q.total_tiles = (unsigned)total;
  q.tmem_cols = umma::tmem_cols_pow2(2 * BN);
  q.a_z1 = a.sA1 != 0; q.a_z2 = a.sA2 != 0; q.a_seg = a.segA != 0;
  q.b_z1 = a.sB1 != 0; q.b_z2 = a.sB2 != 0; q.b_seg = a.segB != 0;

  CUtensorMap mapA, mapB;
  MMS_TRY(make_map(ctx, &mapA, a.A, a.lda, a.a_mn != 0, a.M, a.K, kBM, a.sA1, a.sA2, a.segA, a.nb1, a.nb2, a.nseg));
  MMS_TRY(make_map(ctx, &mapB, a.B, a.ldb, a.b_mn != 0, a.N, a.K, a.b_mn ? 32 : BN, a.sB1, a.sB2, a.segB, a.nb1,
                   a.nb2, a.nseg));

  typedef void (*kernel_t)(const CUtensorMap, const CUtensorMap, const TcGemmArgs, const Geometry, long long*);
  static const kernel_t kernels[4] = {tc_gemm_tma_kernel<false, false>, tc_gemm_tma_kernel<false, true>,
                                      tc_gemm_tma_kernel<true, false>, tc_gemm_tma_kernel<true, true>};
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < 4; ++i)
      MMS_MAX_SMEM(kernels[i], 221 * 1024);
    configured = true;
  }
  const kernel_t kernel = kernels[(a.a_mn ? 2 : 0) + (a.b_mn ? 1 : 0)];
  const unsigned grid = (unsigned)mms_min<long long>(total, ctx->sm_count);
  static const bool tracing = getenv("MMS_TC_TRACE") != nullptr;
  long long* tr = nullptr;
  if (tracing) {
    MMS_CUDA(cudaMalloc(&tr, sizeof(long long) * 8 * grid));
    MMS_CUDA(cudaMemset(tr, 0, sizeof(long long) * 8 * grid));
  }
  { MmsKernelScope ks_(ctx, "tc_gemm_tma_kernel");
    kernel<<<grid, kThreads, smem, ctx->stream>>>(mapA, mapB, a, q, tr); }
  MMS_LAUNCH_CHECK();
  if (tracing) {
    std::vector<long long> h((size_t)8 * grid);
    MMS_CUDA(cudaStreamSynchronize(ctx->stream));
    MMS_CUDA(cudaMemcpy(h.data(), tr, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(tr);
    long long t0 = h[0];
    for (unsigned b = 0; b < grid; ++b) t0 = mms_min(t0, h[(size_t)b * 8]);
    static const char* names[8] = {"entry", "setup", "tma_issued", "first_full", "mma_issued", "acc_ready",
                                   "epi_done", "exit"};
    fprintf(stderr, "[tc trace] M %d N %d K %d nseg %d batch %dx%d ksplit %d a_mn %d b_mn %d BN %d stages %d tiles %u grid %u | ns since first entry (min/avg/max over CTAs):",
            a.M, a.N, a.K, a.nseg, a.nb1, a.nb2, a.ksplit, a.a_mn, a.b_mn, BN, stages, q.total_tiles, grid);
    for (int s = 0; s < 8; ++s) {
      long long mn = 1LL << 62, mx = 0; double sum = 0;
      for (unsigned b = 0; b < grid; ++b) {
        const long long v = h[(size_t)b * 8 + s] - t0;
        mn = mms_min(mn, v); mx = mms_max(mx, v); sum += (double)v;
      }
      fprintf(stderr, " %s %lld/%.0f/%lld", names[s], mn, sum / grid, mx);
    }
    fprintf(stderr, "\n");
  }
  return 0;
}
