// Does the TMA engine's bulk reduction (cp.reduce.async.bulk.global.shared::cta.add.f32, one 1200-byte table row per
// instruction) get more fp32 adds per second out of L2 than red.global.add.v4.f32 from registers?  Both variants add the
// same 1200-byte rows to `rows` random table rows; nothing is loaded.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) bulk_kernel(const int* __restrict__ ids, int n, float* __restrict__ dW, int D, int inflight) {
  extern __shared__ __align__(128) float buf[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = buf + warp * D;
  for (int c = lane; c < D; c += 32) mine[c] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int gw = blockIdx.x * 4 + warp, nw = gridDim.x * 4;
  if (lane == 0) {
    int k = 0;
    for (int i = gw; i < n; i += nw) {
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                   ::"l"(dW + (size_t)ids[i] * D), "r"(smem_u32(mine)), "r"(D * 4) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++k >= inflight) asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void __launch_bounds__(128) red_kernel(const int* __restrict__ ids, int n, float* __restrict__ dW, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + warp, nw = gridDim.x * 4;
  const int nvec = D >> 2;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
  for (int i = gw; i < n; i += nw) {
    float4* d = reinterpret_cast<float4*>(dW + (size_t)ids[i] * D);
    for (int c = lane; c < nvec; c += 32) atomicAdd(d + c, one);
  }
}

int main() {
  const int V = 60002, D = 300;
  for (int n : {92000, 163840}) {
    std::vector<int> h(n);
    srand(1);
    for (int i = 0; i < n; ++i) h[i] = rand() % V;
    int* ids; float* dW;
    CK(cudaMalloc(&ids, 4 * n)); CK(cudaMalloc(&dW, 4ull * V * D));
    CK(cudaMemcpy(ids, h.data(), 4 * n, cudaMemcpyHostToDevice));
    CK(cudaMemset(dW, 0, 4ull * V * D));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 3; ++variant) {
      float best = 1e9f;
      for (int it = 0; it < 5; ++it) {
        CK(cudaEventRecord(e0));
        if (variant == 0) red_kernel<<<148 * 8, 128>>>(ids, n, dW, D);
        else bulk_kernel<<<148 * (variant == 1 ? 8 : 16), 128, 4 * D * 4>>>(ids, n, dW, D, 8);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it) best = fminf(best, ms);
      }
      CK(cudaGetLastError());
      printf("%6d rows x %d floats: %-28s %.1f us  (%.0f G fp32 adds/s)\n", n, D,
             variant == 0 ? "red.global.add.v4.f32" : (variant == 1 ? "cp.reduce.async.bulk (8 CTA/SM)" : "cp.reduce.async.bulk (16 CTA/SM)"),
             best * 1e3f, (double)n * D / (best * 1e-3) / 1e9);
    }
    std::vector<float> back((size_t)V * D);
    CK(cudaMemcpy(back.data(), dW, 4ull * V * D, cudaMemcpyDeviceToHost));
    double sum = 0; for (float x : back) sum += x;
    printf("   checksum %.0f (expected %.0f)\n", sum, 15.0 * n * D);
    cudaFree(ids); cudaFree(dW);
  }
  return 0;
}
