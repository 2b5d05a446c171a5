// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (M = 128, K = 8) for several N, with the A operand in
// shared memory (SS) or in tensor memory (TS).  One CTA per SM, operands are whatever the memories hold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/mma_rate.cu && gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "../mms_answer_selection_b200/csrc/tc/umma.cuh"

using namespace umma;

// mode 0: one accumulator at column dcol.  mode 1: GEMM1-like pair of SS MMAs per k-step (N at column 0, N2 at column N).
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int ts, int b_mn, int iters, long long* out, int dcol,
                                                      int mode, int N2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tile = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 65536) / 4; i += blockDim.x) reinterpret_cast<float*>(tile)[i] = 0.f;
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
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_tf32(128, N, false, b_mn != 0);
    const uint32_t a_base = smem_u32(tile), b_base = a_base + 16384;
    long long t0 = clock64();
    const uint32_t idesc2 = idesc_tf32(128, N2, false, b_mn != 0);
    for (int i = 0; i < iters; ++i) {
      const int ks = i & 3;
      const uint64_t db = b_mn ? desc_mnmajor(b_base + ks * 1024, 4096) : desc_kmajor(b_base + ks * 32);
      if (mode == 1) {
        mma_tf32_ss(tmem, desc_kmajor(a_base + ks * 32), db, idesc, 1u);
        mma_tf32_ss(tmem + N, desc_kmajor(a_base + ks * 32), db, idesc2, 1u);
      } else if (ts) mma_tf32_ts(tmem + dcol, tmem + ((i * 8) % 304), db, idesc, 1u);
      else mma_tf32_ss(tmem + dcol, desc_kmajor(a_base + ks * 32), db, idesc, 1u);
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 4096;
  const int grid = 148;
  for (int ts = 0; ts < 2; ++ts) {
    for (int dcol : {256, 304, 320, 384}) {
      for (int N : {48, 128, 160, 256}) {
        if (dcol + N > 512) continue;
        for (int r = 0; r < 2; ++r) rate_kernel<<<grid, 128, 90 * 1024>>>(N, ts, 0, iters, d, dcol, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        const double cyc = (double)c / iters;
        printf("%s  D at col %3d  N %3d : %.1f cycles / MMA  -> %.2f TFLOP/s per SM at %d MHz (%s)\n",
               ts ? "TS" : "SS", dcol, N, cyc, 2.0 * 128 * N * 8 / cyc * clk_khz * 1e3 * 1e-12, clk_khz / 1000,
               cudaGetErrorString(e));
      }
    }
  }
  const int pairs[4][2] = {{256, 48}, {160, 144}, {152, 152}, {208, 96}};
  for (int b_mn = 0; b_mn < 2; ++b_mn)
    for (int p = 0; p < 4; ++p) {
      if (pairs[p][0] % 16 || pairs[p][1] % 16) continue;
      for (int r = 0; r < 2; ++r)
        rate_kernel<<<grid, 128, 90 * 1024>>>(pairs[p][0], 0, b_mn, iters, d, 0, 1, pairs[p][1]);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("SS pair N %d + %d (B %s): %.1f cycles / k-step (%s)\n", pairs[p][0], pairs[p][1], b_mn ? "MN" : "K",
             (double)c / iters, cudaGetErrorString(e));
    }
  return 0;
}
