// SIMT batched GEMM for fp32 / fp64 blobs: the contraction used by the double
// instantiation of every layer and by float shapes the tcgen05 kernels do not
// cover.  64x64x16 CTA tile, 256 threads, 4x4 outputs per thread, operands staged
// in shared memory with the contiguous global dimension mapped to threadIdx for
// coalescing (either operand may be "transposed" through its strides).
#include "mms_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, NT = 256;

template <typename T>
__device__ __forceinline__ void atomic_add(T* p, T v) { atomicAdd(p, v); }

template <typename T>
__global__ void __launch_bounds__(NT) simt_gemm_kernel(SimtGemmArgs<T> g) {
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int zs = blockIdx.z;
  const int z = zs / g.ksplit, split = zs % g.ksplit;
  const int z1 = z / g.nb2, z2 = z % g.nb2;
  const T* A = g.A + z1 * g.sA1 + z2 * g.sA2;
  const T* B = g.B + z1 * g.sB1 + z2 * g.sB2;
  T* C = g.C + z1 * g.sC1 + z2 * g.sC2;

  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kchunk = ((g.K + g.ksplit - 1) / g.ksplit + BK - 1) / BK * BK;
  const int kbeg = split * kchunk;
  const int kend = min(g.K, kbeg + kchunk);

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

  const bool a_k_contig = (g.sAk == 1);
  const bool b_n_contig = (g.sBn == 1);

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / NT; ++i) {
      const int e = tid + i * NT;
      int mm, kk;
      if (a_k_contig) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < g.M && gk < kend) ? A[(long long)gm * g.sAm + (long long)gk * g.sAk] : T(0);
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / NT; ++i) {
      const int e = tid + i * NT;
      int nn, kk;
      if (b_n_contig) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < g.N && gk < kend) ? B[(long long)gk * g.sBk + (long long)gn * g.sBn] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += av[i] * bv[j];
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= g.N) continue;
      T* c = C + (long long)gm * g.ldc + gn;
      const T v = g.alpha * acc[i][j];
      if (g.ksplit > 1) {
        atomic_add(c, v);
      } else {
        *c = (g.beta == T(0)) ? v : v + g.beta * (*c);
      }
    }
  }
}

template <typename T>
__global__ void fill_kernel(T* p, long long n, T v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

}  // namespace

template <typename T>
int mms_simt_gemm(mms_context* ctx, const SimtGemmArgs<T>& a) {
  if (a.M <= 0 || a.N <= 0 || a.nb1 <= 0 || a.nb2 <= 0) return 0;
  MMS_REQUIRE(a.ksplit >= 1, MMS_E_INVALID, "ksplit must be >= 1");
  MMS_REQUIRE(a.ksplit == 1 || a.beta == T(1), MMS_E_INVALID, "split-K needs beta == 1");
  const long long nz = (long long)a.nb1 * a.nb2 * a.ksplit;
  MMS_REQUIRE(nz <= 65535, MMS_E_UNSUPPORTED, "too many GEMM batches for one launch");
  dim3 grid(mms_ceil_div(a.N, BN), mms_ceil_div(a.M, BM), (unsigned)nz);
  MMS_REQUIRE(grid.y <= 65535, MMS_E_UNSUPPORTED, "GEMM M too large for one launch");
  { MmsKernelScope ks_(ctx, "simt_gemm_kernel");
    simt_gemm_kernel<T><<<grid, NT, 0, ctx->stream>>>(a); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_fill(mms_context* ctx, T* p, long long n, T v) {
  if (n <= 0) return 0;
  if (v == T(0)) {
    MMS_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * n, ctx->stream));
    return 0;
  }
  const int blocks = (int)mms_min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8);
  { MmsKernelScope ks_(ctx, "fill_kernel");
    fill_kernel<T><<<blocks, 256, 0, ctx->stream>>>(p, n, v); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template int mms_simt_gemm<float>(mms_context*, const SimtGemmArgs<float>&);
template int mms_simt_gemm<double>(mms_context*, const SimtGemmArgs<double>&);
template int mms_fill<float>(mms_context*, float*, long long, float);
template int mms_fill<double>(mms_context*, double*, long long, double);
