// Generic TF32 GEMM on tcgen05 with TMEM accumulators (sm_100a).
//
//   C[z] (+)= op(A[z]) * op(B[z])      M x N x K, fp32 in / fp32 out, TF32 multiply
//
// Either operand may be K-major (contiguous along K) or MN-major (contiguous along M / N);
// both are staged into 128-byte-swizzled shared memory in their natural orientation -- no
// transposed copies -- and the UMMA descriptors carry the majorness.  An optional per-row
// scale on either operand fuses diag(ds) into the load (SimMatrix backward).
//
// One CTA computes one 128 x BN output tile for one K split:
//   warps 0-3  epilogue   TMEM -> registers -> global (plain store / += / atomicAdd)
//   warp  4    TMEM allocator + single-thread tcgen05.mma issuer
//   warps 5-12 operand staging: global -> cvt.rna.tf32 -> swizzled smem, 4-stage ring,
//              full/empty mbarriers (tcgen05.commit releases a stage)
#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kBM = 128;
constexpr int kBK = 32;                 // fp32 elements per stage along K = one 128-byte swizzle row
constexpr int kEpiWarps = 4;
constexpr int kLoadWarps = 8;
constexpr int kThreads = (kEpiWarps + 1 + kLoadWarps) * 32;
constexpr int kMaxStages = 4;

struct Smem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void stage_operand(uint8_t* dst, const float* src, long long ld, bool mn_major,
                                              int mn0, int mn_extent, int mn_limit, int k0, int k_limit,
                                              const float* rowscale, bool vec_ok, int tid, int nthreads) {
  if (!mn_major) {
    // rows = m/n index, 32 k-columns
    const int nchunks = mn_extent * 8;
    for (int e = tid; e < nchunks; e += nthreads) {
      const int r = e >> 3, c4 = e & 7;
      const int gr = mn0 + r, gc = k0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < mn_limit && gc < k_limit) {
        const float* p = src + (long long)gr * ld + gc;
        if (vec_ok && gc + 3 < k_limit) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          v.x = __ldg(p);
          if (gc + 1 < k_limit) v.y = __ldg(p + 1);
          if (gc + 2 < k_limit) v.z = __ldg(p + 2);
          if (gc + 3 < k_limit) v.w = __ldg(p + 3);
        }
        if (rowscale) { const float s = __ldg(rowscale + gr); v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
      }
      v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
      *reinterpret_cast<float4*>(dst + swz128(r, c4)) = v;
    }
  } else {
    // blocks of 32 mn-columns; rows = k index (32 per stage)
    const int nblocks = (mn_extent + 31) >> 5;
    const int nchunks = nblocks * kBK * 8;
    for (int e = tid; e < nchunks; e += nthreads) {
      const int c4 = e & 7, r = (e >> 3) & (kBK - 1), b = e >> 8;
      const int gr = k0 + r, gc = mn0 + b * 32 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < k_limit && gc < mn_limit) {
        const float* p = src + (long long)gr * ld + gc;
        if (vec_ok && gc + 3 < mn_limit) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          v.x = __ldg(p);
          if (gc + 1 < mn_limit) v.y = __ldg(p + 1);
          if (gc + 2 < mn_limit) v.z = __ldg(p + 2);
          if (gc + 3 < mn_limit) v.w = __ldg(p + 3);
        }
        if (rowscale) { const float s = __ldg(rowscale + gr); v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
      }
      v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
      *reinterpret_cast<float4*>(dst + b * 4096 + swz128_mn(r, c4)) = v;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) tc_gemm_kernel(TcGemmArgs g, int BN, int stages, int b_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned operand ring first, barriers after it
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = 16384 + b_bytes;
  Smem* sm = reinterpret_cast<Smem*>(ring + stages * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int zs = blockIdx.z;
  const int z = zs / g.ksplit, split = zs % g.ksplit;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int kchunk = ((g.K + g.ksplit - 1) / g.ksplit + kBK - 1) / kBK * kBK;
  const int kbeg = split * kchunk;
  const int kend = min(g.K, kbeg + kchunk);
  const int nk = kend > kbeg ? (kend - kbeg + kBK - 1) / kBK : 0;
  const uint32_t tmem_cols = tmem_cols_pow2(BN);

  if (warp == kEpiWarps) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) { mbar_init(&sm->full[s], kLoadWarps); mbar_init(&sm->empty[s], 1); }
      mbar_init(&sm->acc_full, 1);
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
    const int tid = threadIdx.x - (kEpiWarps + 1) * 32, nth = kLoadWarps * 32;
    const float* A = g.A + (long long)z * g.sA;
    const float* B = g.B + (long long)z * g.sB;
    const bool a_vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (g.lda % 4 == 0) && (!g.a_mn || m0 % 4 == 0);
    const bool b_vec = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (g.ldb % 4 == 0) && (!g.b_mn || n0 % 4 == 0);
    for (int i = 0; i < nk; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&sm->empty[s], ((i / stages) - 1) & 1);
      uint8_t* a_dst = ring + s * stage_bytes;
      uint8_t* b_dst = a_dst + 16384;
      const int k0 = kbeg + i * kBK;
      stage_operand(a_dst, A, g.lda, g.a_mn != 0, m0, kBM, g.M, k0, kend, g.a_rowscale, a_vec, tid, nth);
      stage_operand(b_dst, B, g.ldb, g.b_mn != 0, n0, BN, g.N, k0, kend, g.b_rowscale, b_vec, tid, nth);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->full[s]);
    }
  } else if (warp == kEpiWarps) {
    // ------------------------------------------------------------ MMA issue (one thread)
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(kBM, BN, g.a_mn != 0, g.b_mn != 0);
      for (int i = 0; i < nk; ++i) {
        const int s = i % stages;
        mbar_wait(&sm->full[s], (i / stages) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(ring + s * stage_bytes);
        const uint32_t b_base = a_base + 16384;
#pragma unroll
        for (int ks = 0; ks < kBK / 8; ++ks) {
          const uint64_t da = g.a_mn ? desc_mnmajor(a_base + ks * 1024, 4096) : desc_kmajor(a_base + ks * 32);
          const uint64_t db = g.b_mn ? desc_mnmajor(b_base + ks * 1024, 4096) : desc_kmajor(b_base + ks * 32);
          mma_tf32_ss(tmem, da, db, idesc, (i > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit(&sm->empty[s]);
      }
      mma_commit(&sm->acc_full);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue
    float* C = g.C + (long long)z * g.sC;
    const int row = m0 + warp * 32 + lane;
    if (nk > 0) {
      mbar_wait(&sm->acc_full, 0);
      tc_fence_after();
    }
    const bool c_vec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (g.ldc % 4 == 0) && (n0 % 4 == 0);
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      if (nk > 0) {
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      if (row >= g.M) continue;
      const float rs = g.out_rowscale ? __ldg(g.out_rowscale + row) : 1.f;
      float* crow = C + (long long)row * g.ldc + n0 + c0;
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const int n = n0 + c0 + i4 * 4;
        if (n >= g.N) break;
        float4 o = make_float4(v[i4 * 4] * rs, v[i4 * 4 + 1] * rs, v[i4 * 4 + 2] * rs, v[i4 * 4 + 3] * rs);
        if (g.mode == TC_STORE && c_vec && n + 3 < g.N) {
          *reinterpret_cast<float4*>(crow + i4 * 4) = o;
        } else {
          const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j >= g.N) break;
            float* p = crow + i4 * 4 + j;
            if (g.mode == TC_STORE) *p = ov[j];
            else if (g.mode == TC_ACCUM) *p += ov[j];
            else atomicAdd(p, ov[j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace

int mms_tc_gemm(mms_context* ctx, const TcGemmArgs& a) {
  MMS_REQUIRE(a.A && a.B && a.C, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.batch > 0 && a.ksplit > 0, MMS_E_INVALID, "bad size");
  MMS_REQUIRE(a.ksplit == 1 || a.mode == TC_ATOMIC, MMS_E_INVALID, "split-K needs the atomic epilogue");
  // balanced N tiling, BN a multiple of 16 (UMMA M=128 needs N % 16 == 0), at most 256
  const int ntiles = mms_ceil_div(a.N, 256);
  int BN = mms_ceil_div(mms_ceil_div(a.N, ntiles), 16) * 16;
  if (a.b_mn) BN = mms_ceil_div(BN, 32) * 32 > 256 ? BN : mms_ceil_div(BN, 32) * 32;   // whole 32-wide blocks
  const int n_tiles = mms_ceil_div(a.N, BN);
  const int b_bytes = a.b_mn ? mms_ceil_div(BN, 32) * 4096 : BN * 128;
  const int stage_bytes = 16384 + b_bytes;
  int stages = kMaxStages;
  while (stages > 2 && (size_t)stages * stage_bytes + sizeof(Smem) + 1024 > 200 * 1024) --stages;
  const size_t smem = (size_t)stages * stage_bytes + sizeof(Smem) + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    MMS_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long gz = (long long)a.batch * a.ksplit;
  MMS_REQUIRE(gz <= 65535 && mms_ceil_div(a.M, kBM) <= 65535, MMS_E_UNSUPPORTED, "grid too large");
  dim3 grid(n_tiles, mms_ceil_div(a.M, kBM), (unsigned)gz);
  { MmsKernelScope ks_(ctx, "tc_gemm_kernel");
    tc_gemm_kernel<<<grid, kThreads, smem, ctx->stream>>>(a, BN, stages, b_bytes); }
  MMS_LAUNCH_CHECK();
  return 0;
}
