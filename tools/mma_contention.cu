// Micro-benchmark: how much do concurrent TMA writes into shared memory and concurrent tcgen05.ld/st traffic slow a
// stream of tcgen05.mma kind::tf32 down?  One CTA per SM:
//   warp 0      issues the MMA sequence of the fused kernels' steady state (compile-time descriptors, elect.sync)
//   warp 1      (flag 1) streams 16 KB TMA boxes from an L2-resident buffer into a 4-slot shared-memory ring
//   warps 2-9   (flag 2) tcgen05.ld x32 -> cvt -> tcgen05.st x32 on TMEM columns the MMAs do not touch
//   warps 10-13 (flag 4) plain st.shared traffic (a Gblk-builder stand-in)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mma_contention tools/mma_contention.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../mms_answer_selection_b200/csrc/tc/umma.cuh"

using namespace umma;

__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct Out { long long mma_cycles, tma_boxes, ldst_iters; };

// MODE 0: GEMM-B like  TS 160@0 + TS 144@160, A@304, B K-major
// MODE 1: GEMM-A like  SS 64@304, B MN-major
// MODE 2: fwd GEMM1    SS 160@0 + SS 144@160, B MN-major
// MODE 3: mix per iteration: 2x (TS 160 + TS 144) + 1x SS 64   (the backward's steady state)
template <int MODE>
__global__ void __launch_bounds__(14 * 32, 1)
contention_kernel(const __grid_constant__ CUtensorMap map, int flags, int iters, int rows_total, Out* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tile = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = tile + 16384 + 40960;        // 4 x 16 KB TMA slots
  uint8_t* scratch = ring + 4 * 16384;             // 16 KB for the st.shared traffic
  __shared__ uint64_t bar, tbar[4];
  __shared__ uint32_t tmem_base;
  __shared__ volatile int done;
  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 40960) / 4; i += blockDim.x) reinterpret_cast<float*>(tile)[i] = 0.f;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); for (int s = 0; s < 4; ++s) mbar_init(&tbar[s], 1); done = 0; fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    constexpr uint32_t id160k = idesc_tf32(128, 160, false, false), id144k = idesc_tf32(128, 144, false, false);
    constexpr uint32_t id160m = idesc_tf32(128, 160, false, true), id144m = idesc_tf32(128, 144, false, true);
    constexpr uint32_t id64m = idesc_tf32(128, 64, false, true);
    const uint32_t a_lo = desc_lo_k(smem_u32(tile));
    const uint32_t bk_lo = desc_lo_k(smem_u32(tile) + 16384), bm_lo = desc_lo_mn(smem_u32(tile) + 16384, 4096);
    const uint32_t bk1_lo = bk_lo + ((160u * 128u) >> 4), bm1_lo = bm_lo + 5u * (4096u >> 4);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (MODE == 0 || MODE == 3) {
            mma_tf32_ts_lh(tmem, tmem + 304 + ks * 8, bk_lo + ks * kDescStepK, kDescHiK, id160k, 1u);
            mma_tf32_ts_lh(tmem + 160, tmem + 304 + ks * 8, bk1_lo + ks * kDescStepK, kDescHiK, id144k, 1u);
          }
          if (MODE == 3) {
            mma_tf32_ts_lh(tmem, tmem + 336 + ks * 8, bk_lo + ks * kDescStepK, kDescHiK, id160k, 1u);
            mma_tf32_ts_lh(tmem + 160, tmem + 336 + ks * 8, bk1_lo + ks * kDescStepK, kDescHiK, id144k, 1u);
          }
          if (MODE == 1 || MODE == 3)
            mma_tf32_ss_lh(tmem + 368, a_lo + ks * kDescStepK, kDescHiK, bm_lo + ks * kDescStepMN, kDescHiMN, id64m, 1u);
          if (MODE == 2) {
            mma_tf32_ss_lh(tmem, a_lo + ks * kDescStepK, kDescHiK, bm_lo + ks * kDescStepMN, kDescHiMN, id160m, 1u);
            mma_tf32_ss_lh(tmem + 160, a_lo + ks * kDescStepK, kDescHiK, bm1_lo + ks * kDescStepMN, kDescHiMN, id144m, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) { done = 1; if (blockIdx.x == 0) out->mma_cycles = t1 - t0; }
  } else if (warp == 1) {
    if (flags & 1) {
      long long n = 0;
      uint32_t ph[4] = {0, 0, 0, 0};
      int row = (blockIdx.x * 977) % (rows_total - 128);
      // prime the ring, then refill each slot as soon as its box has landed
      for (int s = 0; s < 4; ++s) {
        if (elect_one_sync()) { mbar_arrive_expect_tx(&tbar[s], 16384); tma_load_2d(ring + s * 16384, &map, &tbar[s], 0, row); }
        __syncwarp();
        row = (row + 128 * 148) % (rows_total - 128);
      }
      while (!done) {
        for (int s = 0; s < 4; ++s) {
          mbar_wait(&tbar[s], ph[s]); ph[s] ^= 1u;
          if (elect_one_sync()) { mbar_arrive_expect_tx(&tbar[s], 16384); tma_load_2d(ring + s * 16384, &map, &tbar[s], 0, row); }
          __syncwarp();
          row = (row + 128 * 148) % (rows_total - 128);
          ++n;
        }
      }
      for (int s = 0; s < 4; ++s) mbar_wait(&tbar[s], ph[s]);       // drain before the CTA exits
      if (lane == 0 && blockIdx.x == 0) out->tma_boxes = n;
    }
  } else if (warp < 10) {
    if (flags & 2) {
      const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t ta = tmem + lane_bits + 432u + (uint32_t)(((warp - 2) >> 2) * 32);
      long long n = 0;
      float v[32];
      while (!done) {
        tmem_ld32(ta, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = to_tf32(v[i]);
        tmem_st32(ta, v);
        tmem_wait_st();
        ++n;
      }
      if (threadIdx.x == 64 && blockIdx.x == 0) out->ldst_iters = n;
    }
  } else {
    if (flags & 4) {
      const int r = (warp - 10) * 32 + lane;
      while (!done) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          *reinterpret_cast<float4*>(scratch + swz128(r, c4)) = make_float4(1.f, 2.f, 3.f, 4.f);
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
static void run(const CUtensorMap& map, int rows, Out* d_out, const char* name, double ideal) {
  const int iters = 8192;
  const size_t smem = 1024 + 16384 + 40960 + 4 * 16384 + 16384;
  cudaFuncSetAttribute(contention_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int flags : {0, 1, 2, 4, 3, 7}) {
    Out h = {0, 0, 0};
    for (int r = 0; r < 2; ++r) {
      cudaMemset(d_out, 0, sizeof(Out));
      contention_kernel<MODE><<<148, 14 * 32, smem>>>(map, flags, iters, rows, d_out);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d_out, sizeof(Out), cudaMemcpyDeviceToHost);
    const double cyc = (double)h.mma_cycles / iters;
    printf("%-44s tma %d ldst %d sts %d : %7.1f cyc/k-step (ideal %5.1f)  TMA %5.1f B/clk  ld+st %5.1f B/clk  (%s)\n", name,
           flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, cyc, ideal, h.tma_boxes * 16384.0 / (double)h.mma_cycles,
           h.ldst_iters * 8.0 * 32 * 32 * 4 * 2 / (double)h.mma_cycles, cudaGetErrorString(e));
  }
}

int main() {
  const int rows = 1 << 18;                       // 2^18 rows x 128 B = 32 MB: L2-resident
  float* buf;
  cudaMalloc(&buf, (size_t)rows * 128);
  cudaMemset(buf, 0, (size_t)rows * 128);
  Out* d_out;
  cudaMalloc(&d_out, sizeof(Out));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t dims[2] = {32, (cuuint64_t)rows}, strides[1] = {128};
  cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  run<0>(map, rows, d_out, "GEMM-B   TS 160@0 + TS 144@160 (B K)", 152);
  run<1>(map, rows, d_out, "GEMM-A   SS 64@368 (B MN)", 48);
  run<2>(map, rows, d_out, "GEMM1    SS 160@0 + SS 144@160 (B MN)", 152);
  run<3>(map, rows, d_out, "bwd mix  2x(TS 160 + TS 144) + SS 64", 352);
  return 0;
}
