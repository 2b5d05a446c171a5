// Per-CTA timeline of the fused kernels (MMS_TC_TRACE=1 in the environment): globaltimer stamps at up to 15
// checkpoints plus the SM clock measured over the kernel.  Slot 0 = entry, slot 11 = exit, slot 15 = cycles.
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include "../mms_common.cuh"

namespace {

constexpr int kTraceSlots = 24;          // 0-14 timestamps, 15 cycles, 16-23 accumulated waits (ns)
__device__ __forceinline__ void trace(long long* tr, int slot) {
  if (!tr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  tr[(size_t)blockIdx.x * kTraceSlots + slot] = (long long)t;
}
__device__ __forceinline__ long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}
// accumulate a duration into slot 16 + i
__device__ __forceinline__ void trace_add(long long* tr, int i, long long ns) {
  if (tr) tr[(size_t)blockIdx.x * kTraceSlots + 16 + i] += ns;
}
__device__ __forceinline__ void trace_begin(long long* tr) {
  trace(tr, 0);
  if (tr) tr[(size_t)blockIdx.x * kTraceSlots + 15] = clock64();
}
__device__ __forceinline__ void trace_end(long long* tr) {
  trace(tr, 11);
  if (tr) tr[(size_t)blockIdx.x * kTraceSlots + 15] = clock64() - tr[(size_t)blockIdx.x * kTraceSlots + 15];
}

// host side: prints min/avg/max over CTAs of every checkpoint, relative to the first CTA's entry
struct TraceBuf {
  long long* dev = nullptr;
  unsigned grid = 0;
  int begin(unsigned g) {
    static const bool tracing = getenv("MMS_TC_TRACE") != nullptr;
    if (!tracing) return 0;
    grid = g;
    MMS_CUDA(cudaMalloc(&dev, sizeof(long long) * kTraceSlots * grid));
    MMS_CUDA(cudaMemset(dev, 0, sizeof(long long) * kTraceSlots * grid));
    return 0;
  }
  int end(mms_context* ctx, const char* what, const char* const* names) {
    if (!dev) return 0;
    long long* h = (long long*)malloc(sizeof(long long) * kTraceSlots * grid);
    MMS_CUDA(cudaStreamSynchronize(ctx->stream));
    MMS_CUDA(cudaMemcpy(h, dev, sizeof(long long) * kTraceSlots * grid, cudaMemcpyDeviceToHost));
    cudaFree(dev);
    long long t0 = h[0];
    for (unsigned b = 0; b < grid; ++b) t0 = mms_min(t0, h[(size_t)b * kTraceSlots]);
    double mhz = 0;
    for (unsigned b = 0; b < grid; ++b)
      mhz += 1e3 * (double)h[(size_t)b * kTraceSlots + 15] /
             (double)mms_max<long long>(1, h[(size_t)b * kTraceSlots + 11] - h[(size_t)b * kTraceSlots]);
    fprintf(stderr, "[fused trace] %s grid %u SM clock %.0f MHz | ns since first entry (min/avg/max over CTAs):", what,
            grid, mhz / grid);
    for (int s = 0; s < kTraceSlots; ++s) {
      if (s == 15 || !names[s]) continue;
      long long mn = 1LL << 62, mx = 0; double sum = 0;
      for (unsigned b = 0; b < grid; ++b) {
        const long long v = h[(size_t)b * kTraceSlots + s] - (s < 15 ? t0 : 0);
        mn = mms_min(mn, v); mx = mms_max(mx, v); sum += (double)v;
      }
      fprintf(stderr, " %s %lld/%.0f/%lld", names[s], mn, sum / grid, mx);
    }
    fprintf(stderr, "\n");
    free(h);
    return 0;
  }
};

}  // namespace
