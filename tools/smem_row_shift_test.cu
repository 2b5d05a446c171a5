// Experiment for a dedicated sentence-convolution kernel: can ONE K-major, 128-byte-swizzled tile of token rows in
// shared memory serve the kh shifted windows, i.e. can tcgen05.mma read its A operand starting at row i of the tile
// (start address + i * 128 B, not a multiple of the 1024-byte swizzle period)?  The smem descriptor has a 3-bit
// "base offset" field for start addresses inside a swizzle period; this test issues D = X[i : i+128, :] * B^T for
// several i with base offset 0 and with base offset i & 7 and compares with the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/smem_row_shift_test tools/smem_row_shift_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "../mms_answer_selection_b200/csrc/tc/umma.cuh"

using namespace umma;

constexpr int kRows = 144, kN = 32, kK = 32;

__global__ void __launch_bounds__(128, 1) shift_kernel(const float* __restrict__ X, const float* __restrict__ B,
                                                       float* __restrict__ D, int shift, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* xt = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 144 rows x 128 B
  uint8_t* bt = xt + 20480;                                                    // 32 rows x 128 B (1024-aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < kRows * 8; e += blockDim.x) {
    const int r = e >> 3, c4 = e & 7;
    *reinterpret_cast<float4*>(xt + swz128(r, c4)) = *reinterpret_cast<const float4*>(X + r * kK + c4 * 4);
  }
  for (int e = threadIdx.x; e < kN * 8; e += blockDim.x) {
    const int r = e >> 3, c4 = e & 7;
    *reinterpret_cast<float4*>(bt + swz128(r, c4)) = *reinterpret_cast<const float4*>(B + r * kK + c4 * 4);
  }
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&tmem_base, 32);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    const uint32_t idesc = idesc_tf32(128, kN, false, false);
    if (elect_one_sync()) {
      for (int ks = 0; ks < kK / 8; ++ks) {
        uint64_t da = desc_kmajor(smem_u32(xt) + shift * 128 + ks * 32);
        if (use_base_offset) da |= (uint64_t)(shift & 7) << 49;
        const uint64_t db = desc_kmajor(smem_u32(bt) + ks * 32);
        mma_tf32_ss(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
      }
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < kN; ++j) D[(warp * 32 + lane) * kN + j] = v[j];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// The same question for an MN-major operand (the reduction index k is the row of the tile): B(n, k) = Xmn[k + shift][n],
// one block of 32 n-columns x 48 k-rows, 128 B per k-row, SWIZZLE_128B_BASE32B.
__global__ void __launch_bounds__(128, 1) shift_mn_kernel(const float* __restrict__ A, const float* __restrict__ Xmn,
                                                          float* __restrict__ D, int shift) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* at = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // A: 128 rows x 128 B, K-major
  uint8_t* bt = at + 16384;                                                    // Xmn: 48 k-rows x 128 B, MN-major
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < 128 * 8; e += blockDim.x) {
    const int r = e >> 3, c4 = e & 7;
    *reinterpret_cast<float4*>(at + swz128(r, c4)) = *reinterpret_cast<const float4*>(A + r * kK + c4 * 4);
  }
  for (int e = threadIdx.x; e < 48 * 8; e += blockDim.x) {
    const int r = e >> 3, c4 = e & 7;
    *reinterpret_cast<float4*>(bt + swz128_mn(r, c4)) = *reinterpret_cast<const float4*>(Xmn + r * 32 + c4 * 4);
  }
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&tmem_base, 32);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    const uint32_t idesc = idesc_tf32(128, kN, false, true);
    if (elect_one_sync()) {
      for (int ks = 0; ks < kK / 8; ++ks) {
        const uint64_t da = desc_kmajor(smem_u32(at) + ks * 32);
        const uint64_t db = desc_mnmajor(smem_u32(bt) + shift * 128 + ks * 1024, 4096);
        mma_tf32_ss(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
      }
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < kN; ++j) D[(warp * 32 + lane) * kN + j] = v[j];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

static float tf32(float x) {
  unsigned u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x;
}

int main() {
  std::vector<float> X(kRows * kK), B(kN * kK), D(128 * kN);
  srand(1);
  for (auto& v : X) v = tf32((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = tf32((rand() % 2001 - 1000) / 1000.f);
  float *dX, *dB, *dD;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int shifts[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 16};
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int s : shifts) {
      cudaMemset(dD, 0, D.size() * 4);
      shift_kernel<<<1, 128, 64 * 1024>>>(dX, dB, dD, s, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d: %s\n", s, ubo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double err = 0, ref_max = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < kN; ++n) {
          double acc = 0;
          for (int k = 0; k < kK; ++k) acc += (double)X[(r + s) * kK + k] * B[n * kK + k];
          err = fmax(err, fabs(acc - D[r * kN + n])); ref_max = fmax(ref_max, fabs(acc));
        }
      printf("shift %2d rows  base_offset field %s : max |err| / max |ref| = %.2e  %s\n", s, ubo ? "= shift & 7" : "= 0       ",
             err / ref_max, err / ref_max < 1e-5 ? "OK" : "WRONG");
    }
  // ---- MN-major operand shifted along the reduction index
  std::vector<float> A(128 * kK), Xmn(48 * 32);
  for (auto& v : A) v = tf32((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : Xmn) v = tf32((rand() % 2001 - 1000) / 1000.f);
  float *dA, *dXmn;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dXmn, Xmn.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dXmn, Xmn.data(), Xmn.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(shift_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int s : {0, 1, 2, 3, 4, 5, 8, 9}) {
    cudaMemset(dD, 0, D.size() * 4);
    shift_mn_kernel<<<1, 128, 64 * 1024>>>(dA, dXmn, dD, s);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("MN shift %d: %s\n", s, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, ref_max = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < kN; ++n) {
        double acc = 0;
        for (int k = 0; k < kK; ++k) acc += (double)A[r * kK + k] * Xmn[(k + s) * 32 + n];
        err = fmax(err, fabs(acc - D[r * kN + n])); ref_max = fmax(ref_max, fabs(acc));
      }
    printf("MN-major operand, shift %2d k-rows (base offset 0) : max |err| / max |ref| = %.2e  %s\n", s, err / ref_max,
           err / ref_max < 1e-5 ? "OK" : "WRONG");
  }
  return 0;
}
