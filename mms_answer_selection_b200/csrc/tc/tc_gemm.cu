// Generic TF32 GEMM on tcgen05 with TMEM accumulators (sm_100a), persistent and warp-specialised.
//
//   C[z] (+)= sum_seg op(A[z][seg]) * op(B[z][seg])    M x N x K per segment, fp32 in / fp32 out
//
// Either operand may be K-major (contiguous along K) or MN-major (contiguous along M / N);
// both are staged into 128-byte-swizzled shared memory in their natural orientation -- no
// transposed copies -- and the UMMA descriptors carry the majorness.  Operands are rounded to
// TF32 with cvt.rna while they pass through registers (the MMA itself would truncate, which
// biases sums), an optional per-row scale fuses diag(ds) into the load (SimMatrix backward).
//
// One CTA per SM walks the tile list (n tile fastest, then m tile, reduction split, batch):
//   warps 0-3  epilogue   TMEM -> registers -> global (plain store / += / atomicAdd, fused bias add)
//   warp  4    TMEM allocator + single-thread tcgen05.mma issuer
//   warps 5-12 operand staging: all of a stage's 16-byte global loads are issued before the
//              first cvt/st.shared (12 loads in flight per thread), ring of up to 6 stages with
//              full/empty mbarriers (tcgen05.commit releases a stage)
// The accumulator is double-buffered in TMEM (2 x BN columns): the epilogue of tile i overlaps
// the main loop of tile i+1.
#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kBM = 128;
constexpr int kBK = 32;                 // fp32 elements per stage along K = one 128-byte swizzle row
constexpr int kEpiWarps = 4;
constexpr int kLoadWarps = 8;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kThreads = (kEpiWarps + 1 + kLoadWarps) * 32;
constexpr int kMaxStages = 6;
constexpr int kAChunks = kBM * 8 / kLoadThreads;      // 16-byte chunks of the A stage per loader thread (4)
constexpr int kBChunks = 256 * 8 / kLoadThreads;      // ... of the largest B stage (8)

struct Smem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

// ---- operand staging ---------------------------------------------------------------------
// A stage holds [rows][32 k] (K-major) or [32 k][cols] (MN-major) as 16-byte chunks.  Loader
// thread `tid` owns chunks e = tid + j*256: the chunk's position inside a 4096-byte block is the
// same for every j, so shared-memory offset, global pointer and validity are affine in j and are
// set up once per tile; a stage in the interior of K is then one predicated LDG.128 per chunk.
//   K-major : r = tid/8 + 32 j (mn index), c4 = tid%8 (4 k columns)
//   MN-major: r = tid/8 (k index), c4 = tid%8, 32-wide mn block j
// Chunks wholly outside the matrix in the mn direction are never staged: they only feed
// accumulator rows / columns that the epilogue does not read.
template <bool MN, int NCH>
struct Stager {
  const float* p0;        // chunk 0 at k = 0 of segment 0
  long long jstride;      // elements from chunk j to chunk j + 1
  long long ld;
  const float* rowscale;
  uint32_t soff0;         // shared-memory offset of chunk 0 (chunk j: + j * 4096)
  uint32_t valid, fast;   // bit j: chunk exists / may be loaded with one 16-byte load
  int r, c4;              // K-major: r = first row of the thread; MN-major: r = k row
  int mn_first;           // MN-major: first mn index of chunk 0;  K-major: mn index of chunk 0's row
  int mn_limit;

  __device__ __forceinline__ void init(int tid) {
    r = tid >> 3; c4 = tid & 7;
    soff0 = MN ? swz128_mn(r, c4) : swz128(r, c4);
  }
  __device__ __forceinline__ void tile(const float* base, long long ld_, const float* rs, int mn0, int extent,
                                       int limit, bool vec_ok) {
    ld = ld_; rowscale = rs; mn_limit = limit;
    valid = fast = 0;
    if (!MN) {
      mn_first = mn0 + r;
      p0 = base + (long long)mn_first * ld + c4 * 4;
      jstride = 32 * ld;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const bool v = (r + 32 * j < extent) && (mn_first + 32 * j < limit);
        valid |= (uint32_t)v << j;
        fast |= (uint32_t)(v && vec_ok) << j;
      }
    } else {
      mn_first = mn0 + c4 * 4;
      p0 = base + (long long)r * ld + mn_first;
      jstride = 32;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const bool v = (32 * j < extent) && (mn_first + 32 * j < limit);
        valid |= (uint32_t)v << j;
        fast |= (uint32_t)(v && vec_ok && (mn_first + 32 * j + 3 < limit)) << j;
      }
    }
  }
  // Issues every global load of the stage (k block k0 of the segment starting at `seg_off`).
  __device__ __forceinline__ void load(long long seg_off, int k0, int K, float4 (&v)[NCH]) const {
    const float* p = p0 + seg_off + (MN ? (long long)k0 * ld : (long long)k0);
    if (k0 + kBK <= K && fast == valid) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((valid >> j) & 1) v[j] = __ldg(reinterpret_cast<const float4*>(p + j * jstride));
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!((valid >> j) & 1)) continue;
      const float* pj = p + j * jstride;
      if (!MN) {
        const int gk = k0 + c4 * 4;
        if (((fast >> j) & 1) && gk + 3 < K) { v[j] = __ldg(reinterpret_cast<const float4*>(pj)); continue; }
        if (gk < K) v[j].x = __ldg(pj);
        if (gk + 1 < K) v[j].y = __ldg(pj + 1);
        if (gk + 2 < K) v[j].z = __ldg(pj + 2);
        if (gk + 3 < K) v[j].w = __ldg(pj + 3);
      } else {
        if (k0 + r >= K) continue;
        if ((fast >> j) & 1) { v[j] = __ldg(reinterpret_cast<const float4*>(pj)); continue; }
        const int gm = mn_first + 32 * j;
        v[j].x = __ldg(pj);
        if (gm + 1 < mn_limit) v[j].y = __ldg(pj + 1);
        if (gm + 2 < mn_limit) v[j].z = __ldg(pj + 2);
        if (gm + 3 < mn_limit) v[j].w = __ldg(pj + 3);
      }
    }
  }
  // cvt.rna.tf32 (+ optional row scale) and swizzled st.shared of the stage.
  __device__ __forceinline__ void store(uint8_t* dst, int k0, int K, const float4 (&v)[NCH]) const {
    float s_mn = 1.f;
    if (MN && rowscale && k0 + r < K) s_mn = __ldg(rowscale + k0 + r);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      if (!((valid >> j) & 1)) continue;
      float s = s_mn;
      if (!MN && rowscale) s = __ldg(rowscale + mn_first + 32 * j);
      float4 o;
      o.x = to_tf32(v[j].x * s); o.y = to_tf32(v[j].y * s); o.z = to_tf32(v[j].z * s); o.w = to_tf32(v[j].w * s);
      *reinterpret_cast<float4*>(dst + soff0 + j * 4096) = o;
    }
  }
};

struct Tile {
  int z1, z2, m0, n0, ibeg, nk;
};

__device__ __forceinline__ Tile decode_tile(const TcGemmArgs& g, unsigned t, int BN, int n_tiles, int m_tiles,
                                            int sps) {
  Tile tl;
  const int n_tile = t % n_tiles; t /= n_tiles;
  const int m_tile = t % m_tiles; t /= m_tiles;
  const int split = t % g.ksplit;
  const int z = t / g.ksplit;
  tl.z1 = z / g.nb2; tl.z2 = z % g.nb2;
  tl.m0 = m_tile * kBM; tl.n0 = n_tile * BN;
  // the reduction is a sequence of (segment, 32-wide k block) stages, split evenly over ksplit CTAs
  const int total_stages = g.nseg * sps;
  const int per_split = (total_stages + g.ksplit - 1) / g.ksplit;
  tl.ibeg = split * per_split;
  tl.nk = max(0, min(total_stages, tl.ibeg + per_split) - tl.ibeg);
  return tl;
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const TcGemmArgs g, const int BN, const int stages, const int b_bytes, const int n_tiles,
               const int m_tiles, const unsigned total_tiles, const uint32_t tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned operand ring first, barriers after it
  // 1024-byte alignment by an OFFSET from the __shared__ symbol (not by integer arithmetic on the pointer), so the
  // compiler keeps the shared address space and emits STS/LDS instead of generic ST/LD for everything derived from it
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = 16384 + b_bytes;
  Smem* sm = reinterpret_cast<Smem*>(ring + stages * stage_bytes);

  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  const int sps = (g.K + kBK - 1) / kBK;

  if (warp == kEpiWarps) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) { mbar_init(&sm->full[s], kLoadWarps); mbar_init(&sm->empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], kEpiWarps); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp > kEpiWarps) {
    // ------------------------------------------------------------ operand staging
    const int tid = threadIdx.x - (kEpiWarps + 1) * 32;
    Stager<A_MN, kAChunks> la;
    Stager<B_MN, kBChunks> lb;
    la.init(tid); lb.init(tid);
    int it = 0;
    for (unsigned t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const Tile tl = decode_tile(g, t, BN, n_tiles, m_tiles, sps);
      const float* A0 = g.A + tl.z1 * g.sA1 + tl.z2 * g.sA2;
      const float* B0 = g.B + tl.z1 * g.sB1 + tl.z2 * g.sB2;
      la.tile(A0, g.lda, g.a_rowscale, tl.m0, kBM, g.M,
              ((reinterpret_cast<uintptr_t>(A0) & 15) == 0) && (g.lda % 4 == 0) && (g.segA % 4 == 0) &&
                  (!A_MN || tl.m0 % 4 == 0));
      lb.tile(B0, g.ldb, g.b_rowscale, tl.n0, BN, g.N,
              ((reinterpret_cast<uintptr_t>(B0) & 15) == 0) && (g.ldb % 4 == 0) && (g.segB % 4 == 0) &&
                  (!B_MN || tl.n0 % 4 == 0));
      for (int i = 0; i < tl.nk; ++i, ++it) {
        const int s = it % stages;
        const int gi = tl.ibeg + i;
        const int seg = gi / sps;
        const int k0 = (gi - seg * sps) * kBK;
        // 1. every global load of the stage is issued before the first use
        float4 va[kAChunks], vb[kBChunks];
        la.load(seg * g.segA, k0, g.K, va);
        lb.load(seg * g.segB, k0, g.K, vb);
        // 2. the smem slot must have been drained by the MMAs that read it last
        if (it >= stages) mbar_wait(&sm->empty[s], ((it / stages) - 1) & 1);
        uint8_t* a_dst = ring + s * stage_bytes;
        la.store(a_dst, k0, g.K, va);
        lb.store(a_dst + 16384, k0, g.K, vb);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->full[s]);
      }
    }
  } else if (warp == kEpiWarps) {
    // ------------------------------------------------------------ MMA issue (whole warp, one elected lane)
    {
      const uint32_t idesc = idesc_tf32(kBM, BN, A_MN, B_MN);
      int it = 0, tcount = 0;
      for (unsigned t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tcount) {
        const Tile tl = decode_tile(g, t, BN, n_tiles, m_tiles, sps);
        const int buf = tcount & 1;
        if (tcount >= 2) mbar_wait(&sm->acc_empty[buf], ((tcount >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t acc = tmem + buf * BN;
        for (int i = 0; i < tl.nk; ++i, ++it) {
          const int s = it % stages;
          mbar_wait(&sm->full[s], (it / stages) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(ring + s * stage_bytes);
          const uint32_t b_base = a_base + 16384;
          if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < kBK / 8; ++ks) {
              const uint64_t da = A_MN ? desc_mnmajor(a_base + ks * 1024, 4096) : desc_kmajor(a_base + ks * 32);
              const uint64_t db = B_MN ? desc_mnmajor(b_base + ks * 1024, 4096) : desc_kmajor(b_base + ks * 32);
              mma_tf32_ss(acc, da, db, idesc, (i > 0 || ks > 0) ? 1u : 0u);
            }
            mma_commit(&sm->empty[s]);
            if (i == tl.nk - 1) mma_commit(&sm->acc_full[buf]);
          }
          __syncwarp();
        }
        if (tl.nk == 0) {
          if (elect_one_sync()) mma_commit(&sm->acc_full[buf]);
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue
    int tcount = 0;
    for (unsigned t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tcount) {
      const Tile tl = decode_tile(g, t, BN, n_tiles, m_tiles, sps);
      const int buf = tcount & 1;
      float* C = g.C + tl.z1 * g.sC1 + tl.z2 * g.sC2;
      const float* Cadd = g.c_add ? g.c_add + tl.z1 * g.s_add1 + tl.z2 * g.s_add2 : nullptr;
      const int row = tl.m0 + warp * 32 + lane;
      mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const bool c_vec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (g.ldc % 4 == 0) && (tl.n0 % 4 == 0);
      const float rs = (g.out_rowscale && row < g.M) ? __ldg(g.out_rowscale + row) : 1.f;
      const uint32_t acc = tmem + buf * BN + ((uint32_t)(warp * 32) << 16);
      const int ncols = min(BN, g.N - tl.n0);       // columns of this tile that exist
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        float v[32];
        if (tl.nk > 0) {
          if (c0 + 16 < BN) tmem_ld32(acc + c0, v);          // BN is a multiple of 16
          else tmem_ld16(acc + c0, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (row >= g.M) continue;
        if (g.round_out) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = to_tf32(v[i]);
        }
        float* crow = C + (long long)row * g.ldc + tl.n0 + c0;
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const int n = tl.n0 + c0 + i4 * 4;
          if (n < g.N && c0 + i4 * 4 < BN) {
            float4 o = make_float4(v[i4 * 4] * rs, v[i4 * 4 + 1] * rs, v[i4 * 4 + 2] * rs, v[i4 * 4 + 3] * rs);
            if (Cadd) {
              const float* ap = Cadd + (long long)row * g.ld_add + n;
              o.x += __ldg(ap);
              if (n + 1 < g.N) o.y += __ldg(ap + 1);
              if (n + 2 < g.N) o.z += __ldg(ap + 2);
              if (n + 3 < g.N) o.w += __ldg(ap + 3);
            }
            float* p = crow + i4 * 4;
            if (g.mode == TC_STORE) {
              if (c_vec && n + 3 < g.N) {
                *reinterpret_cast<float4*>(p) = o;
              } else {
                p[0] = o.x;
                if (n + 1 < g.N) p[1] = o.y;
                if (n + 2 < g.N) p[2] = o.z;
                if (n + 3 < g.N) p[3] = o.w;
              }
            } else if (g.mode == TC_ACCUM) {
              p[0] += o.x;
              if (n + 1 < g.N) p[1] += o.y;
              if (n + 2 < g.N) p[2] += o.z;
              if (n + 3 < g.N) p[3] += o.w;
            } else {
              atomicAdd(p, o.x);
              if (n + 1 < g.N) atomicAdd(p + 1, o.y);
              if (n + 2 < g.N) atomicAdd(p + 2, o.z);
              if (n + 3 < g.N) atomicAdd(p + 3, o.w);
            }
          }
        }
      }
      // this accumulator buffer may be overwritten by the tile after next
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace

int mms_tc_gemm_staged(mms_context* ctx, const TcGemmArgs& a) {
  MMS_REQUIRE(a.A && a.B && a.C, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.nb1 > 0 && a.nb2 > 0 && a.ksplit > 0 && a.nseg > 0,
              MMS_E_INVALID, "bad size");
  MMS_REQUIRE(a.ksplit == 1 || a.mode == TC_ATOMIC, MMS_E_INVALID, "split-K needs the atomic epilogue");
  // balanced N tiling, BN a multiple of 16 (UMMA M=128 needs N % 16 == 0), at most 256
  const int ntiles = mms_ceil_div(a.N, 256);
  int BN = mms_ceil_div(mms_ceil_div(a.N, ntiles), 16) * 16;
  if (a.b_mn) BN = mms_ceil_div(BN, 32) * 32 > 256 ? BN : mms_ceil_div(BN, 32) * 32;   // whole 32-wide blocks
  const int n_tiles = mms_ceil_div(a.N, BN);
  const int m_tiles = mms_ceil_div(a.M, kBM);
  const int b_bytes = a.b_mn ? mms_ceil_div(BN, 32) * 4096 : BN * 128;
  const int stage_bytes = 16384 + b_bytes;
  int stages = kMaxStages;
  while (stages > 2 && (size_t)stages * stage_bytes + sizeof(Smem) + 1024 > 200 * 1024) --stages;
  const size_t smem = (size_t)stages * stage_bytes + sizeof(Smem) + 1024;
  typedef void (*kernel_t)(const TcGemmArgs, const int, const int, const int, const int, const int,
                           const unsigned, const uint32_t);
  static const kernel_t kernels[4] = {tc_gemm_kernel<false, false>, tc_gemm_kernel<false, true>,
                                      tc_gemm_kernel<true, false>, tc_gemm_kernel<true, true>};
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < 4; ++i)
      MMS_MAX_SMEM(kernels[i], 201 * 1024);
    configured = true;
  }
  const kernel_t kernel = kernels[(a.a_mn ? 2 : 0) + (a.b_mn ? 1 : 0)];
  const long long total = (long long)n_tiles * m_tiles * a.ksplit * a.nb1 * a.nb2;
  MMS_REQUIRE(total <= 0x7fffffffLL, MMS_E_UNSUPPORTED, "too many tiles");
  const unsigned grid = (unsigned)mms_min<long long>(total, ctx->sm_count);
  const uint32_t tmem_cols = umma::tmem_cols_pow2(2 * BN);
  { MmsKernelScope ks_(ctx, "tc_gemm_kernel");
    kernel<<<grid, kThreads, smem, ctx->stream>>>(a, BN, stages, b_bytes, n_tiles, m_tiles, (unsigned)total,
                                                  tmem_cols); }
  MMS_LAUNCH_CHECK();
  return 0;
}

int mms_tc_gemm(mms_context* ctx, const TcGemmArgs& a) {
  if (a.operands_tf32) {
    const int rc = mms_tc_gemm_tma(ctx, a);
    if (rc != MMS_E_UNSUPPORTED) return rc;
  }
  return mms_tc_gemm_staged(ctx, a);
}
