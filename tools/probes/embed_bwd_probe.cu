// Timing probe for the Embed backward scatter-add at the C3 shape (163 840 token rows x 300 floats, ids as in
// synth.make_indices: centre-padded, real tokens uniform).  Variants separate the cost of the reads from the cost of the
// red.global.add.v4 traffic.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o embed_bwd_probe embed_bwd_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ float4 vadd(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

template <int MODE>   // 0 full, 1 no atomics (one dummy store), 2 atomics only
__global__ void __launch_bounds__(128) runs_v4(const int* __restrict__ idx, const float* __restrict__ dtop, float* __restrict__ dW,
                                               float* __restrict__ dbias, long long M, int D, int rows_per_cta) {
  __shared__ int s_idx[64];
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const int rows = (int)min((long long)rows_per_cta, M - row0);
  for (int r = threadIdx.x; r < rows; r += blockDim.x) s_idx[r] = idx[row0 + r];
  __syncthreads();
  const int nvec = D >> 2;
  for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
    float4 run = make_float4(0.f, 0.f, 0.f, 0.f), col = run;
    int cur = -1;
    const float4* src = reinterpret_cast<const float4*>(dtop + (size_t)row0 * D) + c;
    for (int rb = 0; rb < rows; rb += 8) {
      float4 g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        g[j] = (MODE != 2 && rb + j < rows) ? __ldcs(src + (size_t)(rb + j) * nvec) : make_float4(1.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (rb + j < rows) {
          const int index = s_idx[rb + j];
          col = vadd(col, g[j]);
          if (index != cur) {
            if (MODE != 1 && cur >= 0) atomicAdd(reinterpret_cast<float4*>(dW + (size_t)cur * D) + c, run);
            cur = index;
            run = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          run = vadd(run, g[j]);
        }
      }
    }
    if (MODE != 1 && cur >= 0) atomicAdd(reinterpret_cast<float4*>(dW + (size_t)cur * D) + c, run);
    if (MODE == 1 && col.x == 123.f) dW[c] = col.y;
    atomicAdd(reinterpret_cast<float4*>(dbias) + c, col);
  }
}

// warp per row, R rows in flight per warp, no run merging
template <int R, int MODE>
__global__ void __launch_bounds__(256) warp_rows(const int* __restrict__ idx, const float* __restrict__ dtop, float* __restrict__ dW,
                                                 long long M, int D) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * 256LL + threadIdx.x) >> 5, nw = (gridDim.x * 256LL) >> 5;
  const int nvec = D >> 2;
  for (long long r0 = warp * R; r0 < M; r0 += nw * R) {
    float4 v[R][3];
    int id[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      id[j] = r0 + j < M ? idx[r0 + j] : -1;
      const float4* s = reinterpret_cast<const float4*>(dtop + (size_t)(r0 + j) * D);
#pragma unroll
      for (int k = 0; k < 3; ++k) if (id[j] >= 0 && lane + 32 * k < nvec) v[j][k] = __ldcs(s + lane + 32 * k);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (id[j] < 0) continue;
      float4* d = reinterpret_cast<float4*>(dW + (size_t)id[j] * D);
#pragma unroll
      for (int k = 0; k < 3; ++k) if (lane + 32 * k < nvec) {
        if (MODE == 0) atomicAdd(d + lane + 32 * k, v[j][k]);
        else if (v[j][k].x == 123.f) d[lane] = v[j][k];
      }
    }
  }
}

// sorted form: counting sort by id, then one warp per chunk of <= C rows of one id
__global__ void hist_kernel(const int* __restrict__ idx, int* __restrict__ count, long long M) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < M; i += gridDim.x * 256LL) atomicAdd(count + idx[i], 1);
}
// single CTA scan over V bins: start[v], chunk_off[v] (chunks of C rows)
template <int C>
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ count, int* __restrict__ start, int* __restrict__ cursor,
                                                    int* __restrict__ chunk_off, int V) {
  __shared__ int s_a[1024], s_b[1024];
  const int per = (V + 1023) / 1024, lo = threadIdx.x * per, hi = min(V, lo + per);
  int a = 0, b = 0;
  for (int v = lo; v < hi; ++v) { const int c = count[v]; a += c; b += (c + C - 1) / C; }
  s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int xa = 0, xb = 0;
    if ((int)threadIdx.x >= o) { xa = s_a[threadIdx.x - o]; xb = s_b[threadIdx.x - o]; }
    __syncthreads();
    s_a[threadIdx.x] += xa; s_b[threadIdx.x] += xb;
    __syncthreads();
  }
  a = s_a[threadIdx.x] - a; b = s_b[threadIdx.x] - b;
  for (int v = lo; v < hi; ++v) {
    const int c = count[v];
    start[v] = a; cursor[v] = a; chunk_off[v] = b;
    a += c; b += (c + C - 1) / C;
  }
  if (threadIdx.x == 1023) { start[V] = a; chunk_off[V] = b; }
}
__global__ void place_kernel(const int* __restrict__ idx, int* __restrict__ cursor, int* __restrict__ svals, long long M) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < M; i += gridDim.x * 256LL) svals[atomicAdd(cursor + idx[i], 1)] = (int)i;
}
template <int C>
__global__ void __launch_bounds__(256) chunk_reduce(const float* __restrict__ dtop, float* __restrict__ dW, const int* __restrict__ start,
                                                    const int* __restrict__ chunk_off, const int* __restrict__ svals, int V, int D) {
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  const int total = chunk_off[V];
  for (int j = (blockIdx.x * 256 + threadIdx.x) >> 5; j < total; j += (gridDim.x * 256) >> 5) {
    int lo = 0, hi = V - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (chunk_off[mid] <= j) lo = mid; else hi = mid - 1; }
    const int id = lo;
    const int i0 = start[id] + (j - chunk_off[id]) * C, i1 = min(i0 + C, start[id + 1]);
    float4 acc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = i0; i < i1; i += 4) {
      float4 v[4][3];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const bool ok = i + r < i1;
        const float4* s = reinterpret_cast<const float4*>(dtop + (size_t)svals[ok ? i + r : i] * D);
#pragma unroll
        for (int k = 0; k < 3; ++k) v[r][k] = (ok && lane + 32 * k < nvec) ? __ldcs(s + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = vadd(acc[k], v[r][k]);
    }
    float4* d = reinterpret_cast<float4*>(dW + (size_t)id * D);
#pragma unroll
    for (int k = 0; k < 3; ++k) if (lane + 32 * k < nvec) atomicAdd(d + lane + 32 * k, acc[k]);
  }
}

int main() {
  const int N = 4096, L = 40, D = 300, V = 60002;
  const long long M = (long long)N * L;
  std::vector<int> h(M);
  srand(22);
  for (int n = 0; n < N; ++n) {
    const int len = 5 + rand() % 36, pad_b = (L - len) / 2;
    for (int l = 0; l < L; ++l) h[(size_t)n * L + l] = (l >= pad_b && l < pad_b + len) ? rand() % (V - 2) : V - 1;
  }
  int *idx, *count, *start, *cursor, *coff, *svals; float *dtop, *dW, *db, *flush;
  CK(cudaMalloc(&idx, 4 * M)); CK(cudaMalloc(&dtop, 4 * M * D)); CK(cudaMalloc(&dW, 4ull * V * D)); CK(cudaMalloc(&db, 4 * D));
  CK(cudaMalloc(&count, 4 * (V + 1))); CK(cudaMalloc(&start, 4 * (V + 1))); CK(cudaMalloc(&cursor, 4 * (V + 1))); CK(cudaMalloc(&coff, 4 * (V + 1)));
  CK(cudaMalloc(&svals, 4 * M)); CK(cudaMalloc(&flush, 256 << 20));
  CK(cudaMemcpy(idx, h.data(), 4 * M, cudaMemcpyHostToDevice));
  CK(cudaMemset(dtop, 0, 4 * M * D)); CK(cudaMemset(dW, 0, 4ull * V * D)); CK(cudaMemset(db, 0, 4 * D));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn) {
    float best = 1e9f, sum = 0.f;
    for (int it = 0; it < 7; ++it) {
      CK(cudaMemsetAsync(flush, it, 256 << 20));
      CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2) { best = fminf(best, ms); sum += ms; }
    }
    CK(cudaGetLastError());
    printf("%-44s best %.1f us  mean %.1f us\n", name, best * 1e3f, sum / 5 * 1e3f);
  };
  for (int rpc : {16, 32, 64}) {
    char nm[64]; snprintf(nm, 64, "runs_v4 full rows/cta %d", rpc);
    timeit(nm, [&] { runs_v4<0><<<(unsigned)((M + rpc - 1) / rpc), 96>>>(idx, dtop, dW, db, M, D, rpc); });
    snprintf(nm, 64, "runs_v4 no atomics rows/cta %d", rpc);
    timeit(nm, [&] { runs_v4<1><<<(unsigned)((M + rpc - 1) / rpc), 96>>>(idx, dtop, dW, db, M, D, rpc); });
    snprintf(nm, 64, "runs_v4 atomics only rows/cta %d", rpc);
    timeit(nm, [&] { runs_v4<2><<<(unsigned)((M + rpc - 1) / rpc), 96>>>(idx, dtop, dW, db, M, D, rpc); });
  }
  timeit("warp_rows<4> reads only", [&] { warp_rows<4, 1><<<148 * 8, 256>>>(idx, dtop, dW, M, D); });
  timeit("warp_rows<2> reads only", [&] { warp_rows<2, 1><<<148 * 8, 256>>>(idx, dtop, dW, M, D); });
  timeit("warp_rows<4> with red.v4 per row", [&] { warp_rows<4, 0><<<148 * 8, 256>>>(idx, dtop, dW, M, D); });
  auto plan = [&](auto scan) {
    cudaMemsetAsync(count, 0, 4 * (V + 1));
    hist_kernel<<<148 * 2, 256>>>(idx, count, M);
    scan();
    place_kernel<<<148 * 2, 256>>>(idx, cursor, svals, M);
  };
  timeit("plan (hist + scan + place), C=32", [&] { plan([&] { scan_kernel<32><<<1, 1024>>>(count, start, cursor, coff, V); }); });
  timeit("chunk_reduce C=32", [&] { chunk_reduce<32><<<148 * 8, 256>>>(dtop, dW, start, coff, svals, V, D); });
  timeit("plan (hist + scan + place), C=8", [&] { plan([&] { scan_kernel<8><<<1, 1024>>>(count, start, cursor, coff, V); }); });
  timeit("chunk_reduce C=8", [&] { chunk_reduce<8><<<148 * 8, 256>>>(dtop, dW, start, coff, svals, V, D); });
  timeit("plan + chunk_reduce C=8", [&] { plan([&] { scan_kernel<8><<<1, 1024>>>(count, start, cursor, coff, V); });
                                          chunk_reduce<8><<<148 * 8, 256>>>(dtop, dW, start, coff, svals, V, D); });
  return 0;
}
