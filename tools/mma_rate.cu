// Micro-benchmark: execution rate of tcgen05.mma kind::tf32 (M = 128, K = 8) for several N, with the A operand in
// shared memory (SS) or in tensor memory (TS), the accumulator at several TMEM column offsets, and the MMA
// sequences the fused kernels issue per k-step.  The issue loop is walked by the whole warp and the MMA is issued
// under elect.sync, exactly like the kernels (a `threadIdx.x == 0` guard makes ptxas wrap every UTCHMMA in a
// per-thread waterfall and measures that instead).  One CTA per SM, operands are whatever the memories hold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_rate tools/mma_rate.cu && tools/bin/mma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "../mms_answer_selection_b200/csrc/tc/umma.cuh"

using namespace umma;

struct Seq {          // up to 4 MMAs per iteration
  int n;              // MMAs per iteration
  int N[4];           // MMA N
  int ts[4];          // A operand from TMEM
  int dcol[4];        // accumulator column
  int acol[4];        // TS: A operand column
  int bmn[4];         // B operand MN-major
};

__global__ void __launch_bounds__(128, 1) rate_kernel(const Seq q, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tile = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 4 * 40960) / 4; i += blockDim.x) reinterpret_cast<float*>(tile)[i] = 0.f;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
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
    uint32_t idesc[4], b_hi[4], b_step[4];
    for (int j = 0; j < 4; ++j) {
      idesc[j] = idesc_tf32(128, q.N[j], false, q.bmn[j] != 0);
      b_hi[j] = q.bmn[j] ? kDescHiMN : kDescHiK;
      b_step[j] = q.bmn[j] ? kDescStepMN : kDescStepK;
    }
    const uint32_t a_base = smem_u32(tile), b_base = a_base + 16384;
    const uint32_t a_lo = desc_lo_k(a_base);
    const uint32_t bk_lo = desc_lo_k(b_base), bmn_lo = desc_lo_mn(b_base, 4096);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int ks = i & 3;
      if (elect_one_sync()) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < q.n) {
            const uint32_t b_lo = (q.bmn[j] ? bmn_lo : bk_lo) + ks * b_step[j];
            if (q.ts[j]) mma_tf32_ts_lh(tmem + q.dcol[j], tmem + q.acol[j] + ks * 8, b_lo, b_hi[j], idesc[j], 1u);
            else mma_tf32_ss_lh(tmem + q.dcol[j], a_lo + ks * kDescStepK, kDescHiK, b_lo, b_hi[j], idesc[j], 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// Compile-time variant: everything but the k-step offset is an immediate, eight iterations unrolled, so the loop
// carries (almost) no uniform-datapath work besides the MMAs themselves.
template <int N0, int TS0, int D0, int A0, int BMN0, int N1, int TS1, int D1, int A1, int BMN1, int NM>
__global__ void __launch_bounds__(128, 1) rate_ct_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tile = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 4 * 40960) / 4; i += blockDim.x) reinterpret_cast<float*>(tile)[i] = 0.f;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
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
    constexpr uint32_t id0 = idesc_tf32(128, N0, false, BMN0 != 0), id1 = idesc_tf32(128, N1 > 0 ? N1 : 16, false, BMN1 != 0);
    const uint32_t a_lo = desc_lo_k(smem_u32(tile));
    const uint32_t b0_lo = BMN0 ? desc_lo_mn(smem_u32(tile) + 16384, 4096) : desc_lo_k(smem_u32(tile) + 16384);
    const uint32_t b1_lo = BMN1 ? desc_lo_mn(smem_u32(tile) + 16384, 4096) : desc_lo_k(smem_u32(tile) + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
      if (elect_one_sync()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int ks = u & 3;
          if (TS0) mma_tf32_ts_lh(tmem + D0, tmem + A0 + ks * 8, b0_lo + ks * (BMN0 ? kDescStepMN : kDescStepK), BMN0 ? kDescHiMN : kDescHiK, id0, 1u);
          else mma_tf32_ss_lh(tmem + D0, a_lo + ks * kDescStepK, kDescHiK, b0_lo + ks * (BMN0 ? kDescStepMN : kDescStepK), BMN0 ? kDescHiMN : kDescHiK, id0, 1u);
          if (NM > 1) {
            if (TS1) mma_tf32_ts_lh(tmem + D1, tmem + A1 + ks * 8, b1_lo + ks * (BMN1 ? kDescStepMN : kDescStepK), BMN1 ? kDescHiMN : kDescHiK, id1, 1u);
            else mma_tf32_ss_lh(tmem + D1, a_lo + ks * kDescStepK, kDescHiK, b1_lo + ks * (BMN1 ? kDescStepMN : kDescStepK), BMN1 ? kDescHiMN : kDescHiK, id1, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static long long* d_out;
template <int N0, int TS0, int D0, int A0, int BMN0, int N1 = 0, int TS1 = 0, int D1 = 0, int A1 = 0, int BMN1 = 0, int NM = 1>
static void run_ct(const char* name, double ideal) {
  const int iters = 4096;
  auto k = rate_ct_kernel<N0, TS0, D0, A0, BMN0, N1, TS1, D1, A1, BMN1, NM>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int r = 0; r < 2; ++r) k<<<148, 128, 200 * 1024>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
  printf("CT %-52s : %6.1f cyc/iter  ideal %5.1f  (%s)\n", name, (double)c / iters, ideal, cudaGetErrorString(e));
}
static double run(const Seq& q, int grid = 148) {
  const int iters = 4096;
  for (int r = 0; r < 2; ++r) rate_kernel<<<grid, 128, 200 * 1024>>>(q, iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
  long long c = 0;
  cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
  return (double)c / iters;
}

int main() {
  cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  // 0. compile-time sequences (no issue-loop overhead)
  run_ct<16, 0, 0, 0, 0>("SS N16 @0 B K", 8);
  run_ct<64, 0, 0, 0, 0>("SS N64 @0 B K", 32);
  run_ct<96, 0, 0, 0, 1>("SS N96 @0 B MN", 48);
  run_ct<128, 0, 0, 0, 0>("SS N128 @0 B K", 64);
  run_ct<160, 0, 0, 0, 0>("SS N160 @0 B K", 80);
  run_ct<160, 0, 0, 0, 1>("SS N160 @0 B MN", 80);
  run_ct<256, 0, 0, 0, 0>("SS N256 @0 B K", 128);
  run_ct<256, 0, 0, 0, 1>("SS N256 @0 B MN", 128);
  run_ct<64, 1, 0, 320, 0>("TS N64 @0 A@320 B K", 32);
  run_ct<128, 1, 0, 320, 0>("TS N128 @0 A@320 B K", 64);
  run_ct<128, 1, 304, 0, 0>("TS N128 @304 A@0 B K (fwd GEMM2)", 64);
  run_ct<160, 1, 304, 0, 0>("TS N160 @304 A@0 B K (bwd GEMM-B dQ)", 80);
  run_ct<160, 1, 304, 0, 1>("TS N160 @304 A@0 B MN (bwd GEMM-B dA)", 80);
  run_ct<256, 1, 0, 320, 0>("TS N256 @0 A@320 B K", 128);
  run_ct<160, 0, 0, 0, 1, 144, 0, 160, 0, 1, 2>("SS 160@0 + SS 144@160 B MN (GEMM1 / GEMM-A)", 152);
  run_ct<160, 0, 0, 0, 0, 144, 0, 160, 0, 0, 2>("SS 160@0 + SS 144@160 B K", 152);
  run_ct<160, 1, 0, 304, 0, 144, 1, 160, 304, 0, 2>("TS 160@0 + TS 144@160 A@304 B K (new GEMM-B)", 152);
  run_ct<160, 1, 0, 304, 1, 144, 1, 160, 304, 1, 2>("TS 160@0 + TS 144@160 A@304 B MN (new GEMM-B)", 152);
  run_ct<96, 0, 304, 0, 1>("SS N96 @304 B MN (new GEMM-A)", 48);
  run_ct<64, 0, 304, 0, 1>("SS N64 @304 B MN (new GEMM-A)", 32);
  // 1. single MMA per iteration
  for (int ts = 0; ts < 2; ++ts)
    for (int bmn = 0; bmn < 1; ++bmn)
      for (int dcol : {0}) {
        for (int N : {64, 160, 256}) {
          if (dcol + N > 512) continue;
          Seq q = {1, {N}, {ts}, {dcol}, {dcol ? 0 : 320}, {bmn}};
          const double cyc = run(q);
          printf("%s B %-2s D@%3d N %3d : %6.1f cyc/MMA  ideal %5.1f  (%.0f FLOP/clk)\n", ts ? "TS" : "SS", bmn ? "MN" : "K",
                 dcol, N, cyc, N / 2.0, 2.0 * 128 * N * 8 / cyc);
        }
      }
  // 2. the k-steps of the fused kernels
  struct Named { const char* name; Seq q; double ideal; };
  const Named seqs[] = {
      {"fwd GEMM1   SS 160@0 + SS 144@160 (B MN)", {2, {160, 144}, {0, 0}, {0, 160}, {0, 0}, {1, 1}}, 152},
      {"fwd GEMM2   TS 128@304 (A@0, B K)", {1, {128}, {1}, {304}, {0}, {0}}, 64},
      {"bwd GEMM-A  SS 160@0 + SS 144@160 (B MN)", {2, {160, 144}, {0, 0}, {0, 160}, {0, 0}, {1, 1}}, 152},
      {"bwd GEMM-B  TS 160@304 (A@0, B K)", {1, {160}, {1}, {304}, {0}, {0}}, 80},
      {"bwd GEMM-B  TS 160@304 (A@0, B MN)", {1, {160}, {1}, {304}, {0}, {1}}, 80},
      {"new GEMM-A  SS 96@304 (B MN)", {1, {96}, {0}, {304}, {0}, {1}}, 48},
      {"new GEMM-A  SS 64@304 (B MN)", {1, {64}, {0}, {304}, {0}, {1}}, 32},
      {"new GEMM-B  TS 160@0 + TS 144@160 (A@304, B K)", {2, {160, 144}, {1, 1}, {0, 160}, {304, 304}, {0, 0}}, 152},
      {"new GEMM-B  TS 160@0 + TS 144@160 (A@304, B MN)", {2, {160, 144}, {1, 1}, {0, 160}, {304, 304}, {1, 1}}, 152},
      {"new GEMM-B  TS 256@0 + TS 48@256 (A@304, B K)", {2, {256, 48}, {1, 1}, {0, 256}, {304, 304}, {0, 0}}, 152},
      {"new mix     SS 96@304 + TS 160@0 + TS 144@160 (A@400)", {3, {96, 160, 144}, {0, 1, 1}, {304, 0, 160}, {0, 400, 400}, {1, 0, 0}}, 200},
      {"new mix     SS 64@304 + TS 160@0 + TS 144@160 (A@400)", {3, {64, 160, 144}, {0, 1, 1}, {304, 0, 160}, {0, 400, 400}, {1, 0, 0}}, 184},
  };
  for (const Named& s : seqs) printf("%-58s : %6.1f cyc/iter  ideal %5.1f\n", s.name, run(s.q), s.ideal);
  return 0;
}
