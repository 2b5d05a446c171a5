// Embed layer on sm_100a: index -> row gather (+bias) and gradient scatter-add.
// HBM-bound byte work; no tensor cores.  Reference: src/caffe/layers/embed_layer.cu.
//
// Forward  : one warp per token row; the float-encoded id is read once per row
//            (the reference re-reads it once per output float, embed_layer.cu:15-19),
//            the table row moves as 16-byte vectors when D and the pointers allow,
//            bias is fused (the reference runs a K=1 cuBLAS gemm, :51-55).
// Backward : each CTA owns a contiguous tile of token rows; a thread walks its column
//            down the tile and merges RUNS of equal ids before issuing one atomic per
//            run (centre-padded sentences put long runs of the pad id back to back),
//            which removes most of the same-address contention the reference's
//            one-atomic-per-float kernel has on the pad row (embed_layer.cu:29-39).
//            The bias gradient (a column sum) rides along in the same pass.
#include "mms_common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T, int VEC>
struct Vec;
template <> struct Vec<float, 4> { typedef float4 type; };
template <> struct Vec<double, 2> { typedef double2 type; };

__device__ __forceinline__ float4 vadd(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ double2 vadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ void vzero(float4& a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(double2& a) { a = make_double2(0., 0.); }

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 tf32_rn(float4 v) { return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w)); }
__device__ __forceinline__ double2 tf32_rn(double2 v) { return v; }   // staging is float only

// `staged` (MMS_OPT_STAGE_TF32, float): the same rows once more, rounded to TF32 (round to nearest, what the tensor-core
// contractions read) with a row pitch of `lds` floats -- written here, while the row is in registers, instead of by a
// rounding pass that would read the top back from HBM.
template <typename T, int VEC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
embed_forward_vec(const T* __restrict__ idx, const T* __restrict__ W, const T* __restrict__ bias,
                  T* __restrict__ top, long long M, int D, int V, int* fault, T* __restrict__ staged, int lds) {
  typedef typename Vec<T, VEC>::type VT;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
  const int nvec = D / VEC;
  for (long long row = warp; row < M; row += nwarps) {
    const int index = static_cast<int>(idx[row]);
    const bool ok = index >= 0 && index < V;
    if (!ok && lane == 0) atomicExch(fault, 1);
    const VT* src = reinterpret_cast<const VT*>(W + (size_t)(ok ? index : 0) * D);
    VT* dst = reinterpret_cast<VT*>(top + (size_t)row * D);
    for (int c = lane; c < nvec; c += 32) {
      VT v;
      if (ok) v = __ldg(src + c); else vzero(v);
      if (bias) v = vadd(v, __ldg(reinterpret_cast<const VT*>(bias) + c));
      if (top) __stcs(dst + c, v);   // streaming store: the gathered rows are consumed once (MMS_OPT_STAGE_ONLY: not at all)
      if (staged) reinterpret_cast<VT*>(staged + (size_t)row * lds)[c] = tf32_rn(v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
embed_forward_scalar(const T* __restrict__ idx, const T* __restrict__ W, const T* __restrict__ bias,
                     T* __restrict__ top, long long M, int D, int V, int* fault) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
  for (long long row = warp; row < M; row += nwarps) {
    const int index = static_cast<int>(idx[row]);
    const bool ok = index >= 0 && index < V;
    if (!ok && lane == 0) atomicExch(fault, 1);
    const T* src = W + (size_t)(ok ? index : 0) * D;
    T* dst = top + (size_t)row * D;
    for (int c = lane; c < D; c += 32) {
      T v = ok ? __ldg(src + c) : T(0);
      if (bias) v += __ldg(bias + c);
      dst[c] = v;
    }
  }
}

constexpr int kBwdMaxRows = 64;   // token rows per CTA (upper bound; small batches use fewer rows per CTA)
constexpr int kBwdThreads = 128;

// One CTA owns `rows_per_cta` consecutive token rows; thread c walks column c (and c + 128, ...)
// down the tile.  The gradient values of up to 8 rows are loaded before the first is used, so the
// loads overlap; runs of equal ids (the pad id in centre-padded sentences) are merged into one
// atomic per run.
template <typename T>
__global__ void __launch_bounds__(kBwdThreads)
embed_backward_runs(const T* __restrict__ idx, const T* __restrict__ dtop, T* __restrict__ dW,
                    T* __restrict__ dbias, long long M, int D, int V, int rows_per_cta, int* fault) {
  __shared__ int s_idx[kBwdMaxRows];
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const int rows = (int)mms_min<long long>(rows_per_cta, M - row0);
  for (int r = threadIdx.x; r < rows; r += kBwdThreads) {
    const int index = static_cast<int>(idx[row0 + r]);
    const bool ok = index >= 0 && index < V;
    if (!ok) atomicExch(fault, 1);
    s_idx[r] = ok ? index : -1;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += kBwdThreads) {
    T run = T(0), col = T(0);
    int cur = -1;
    const T* src = dtop + (size_t)row0 * D + c;
    for (int rb = 0; rb < rows; rb += 8) {
      T g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = (rb + j < rows) ? __ldg(src + (size_t)(rb + j) * D) : T(0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (rb + j < rows) {
          const int index = s_idx[rb + j];
          col += g[j];
          if (index != cur) {
            if (cur >= 0 && dW) atomicAdd(dW + (size_t)cur * D + c, run);
            cur = index;
            run = T(0);
          }
          run += g[j];
        }
      }
    }
    if (cur >= 0 && dW) atomicAdd(dW + (size_t)cur * D + c, run);
    if (dbias) atomicAdd(dbias + c, col);
  }
}

template <typename T> struct VecWidth;
template <> struct VecWidth<float> { static constexpr int value = 4; };
template <> struct VecWidth<double> { static constexpr int value = 2; };


// float, D % 4 == 0, 16-byte aligned pointers: thread c owns the 16-byte column group c of the tile (D/4 groups:
// one pass over the tile instead of ceil(D / 128) passes of 4-byte accesses, each with its own exposed load
// latency), eight rows of 16-byte loads in flight per thread, one red.global.add.v4.f32 per run of equal ids.
__global__ void __launch_bounds__(128)
embed_backward_runs_v4(const float* __restrict__ idx, const float* __restrict__ dtop, float* __restrict__ dW,
                       float* __restrict__ dbias, long long M, int D, int V, int rows_per_cta, int* fault) {
  __shared__ int s_idx[kBwdMaxRows];
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const int rows = (int)mms_min<long long>(rows_per_cta, M - row0);
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const int index = static_cast<int>(idx[row0 + r]);
    const bool ok = index >= 0 && index < V;
    if (!ok) atomicExch(fault, 1);
    s_idx[r] = ok ? index : -1;
  }
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
        g[j] = (rb + j < rows) ? __ldcs(src + (size_t)(rb + j) * nvec) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (rb + j < rows) {
          const int index = s_idx[rb + j];
          col = vadd(col, g[j]);
          if (index != cur) {
            if (cur >= 0 && dW) atomicAdd(reinterpret_cast<float4*>(dW + (size_t)cur * D) + c, run);
            cur = index;
            run = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          run = vadd(run, g[j]);
        }
      }
    }
    if (cur >= 0 && dW) atomicAdd(reinterpret_cast<float4*>(dW + (size_t)cur * D) + c, run);
    if (dbias) atomicAdd(reinterpret_cast<float4*>(dbias) + c, col);
  }
}

template <typename T>
inline bool launch_backward_v4(mms_context*, const T*, const T*, T*, T*, long long, int, int, int, long long) {
  return false;
}
template <>
inline bool launch_backward_v4<float>(mms_context* ctx, const float* idx, const float* dtop, float* dW, float* dbias,
                                      long long M, int D, int V, int rows_per_cta, long long grid) {
  const bool ok = (D % 4 == 0) && aligned16(dtop) && (!dW || aligned16(dW)) && (!dbias || aligned16(dbias));
  if (!ok) return false;
  const int threads = mms_min(128, mms_ceil_div(D / 4, 32) * 32);
  MmsKernelScope ks_(ctx, "embed_backward_runs");
  if (mms_prefer_max_shared(reinterpret_cast<const void*>(embed_backward_runs_v4)) != 0) return false;
  embed_backward_runs_v4<<<(unsigned)grid, threads, 0, ctx->stream>>>(idx, dtop, dW, dbias, M, D, V, rows_per_cta,
                                                                      ctx->fault_flag);
  return true;
}

}  // namespace

template <typename T>
int mms_embed_forward_impl(mms_context* ctx, const T* idx, const T* W, const T* bias, T* top,
                           long long M, int D, int V) {
  MMS_REQUIRE(M >= 0 && D > 0 && V > 0, MMS_E_INVALID, "bad size");
  if (M == 0) return 0;
  MMS_REQUIRE(idx && W && top, MMS_E_INVALID, "null pointer");
  constexpr int VEC = VecWidth<T>::value;
  const int grid = (int)mms_min<long long>((M + kWarpsPerCta - 1) / kWarpsPerCta, (long long)ctx->sm_count * 16);
  const bool vec_ok = (D % VEC == 0) && aligned16(W) && aligned16(top) && (!bias || aligned16(bias));
  if (vec_ok) {
    T* staged = nullptr;
    const int lds = (D + 31) & ~31;                       // rows on 128-byte lines, the pitch the contractions use
    if (ctx->stage_tf32 && sizeof(T) == 4) {
      const size_t need = sizeof(float) * (size_t)M * lds;
      if (need > ctx->stage_bytes) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        MMS_REQUIRE(cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone,
                    MMS_E_INVALID, "the staging buffer must have its size before a capture: run the step once eagerly");
        mms_stage_drop_owner(ctx);
        if (ctx->stage_buf) { MMS_CUDA(cudaStreamSynchronize(ctx->stream)); MMS_CUDA(cudaFree(ctx->stage_buf)); }
        ctx->stage_buf = nullptr; ctx->stage_bytes = 0;
        MMS_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->stage_buf), need));
        ctx->stage_bytes = need;
      }
      staged = reinterpret_cast<T*>(ctx->stage_buf);
    }
    { MmsKernelScope ks_(ctx, "embed_forward_vec");
      MMS_CARVEOUT((embed_forward_vec<T, VEC>));
      const bool only = staged && ctx->stage_only;
      // the handle owns ONE staging buffer: whatever it published for an earlier top is about to be overwritten
      if (staged) mms_stage_drop_owner(ctx);
      embed_forward_vec<T, VEC><<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(idx, W, bias, only ? nullptr : top, M, D, V,
                                                                           ctx->fault_flag, staged, lds);
      if (staged) mms_stage_publish(ctx, top, ctx->stage_buf, M, D, lds, only); }
  } else {
    { MmsKernelScope ks_(ctx, "embed_forward_scalar");
      embed_forward_scalar<T><<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(idx, W, bias, top, M, D, V,
                                                                         ctx->fault_flag); }
  }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_embed_backward_impl(mms_context* ctx, const T* idx, const T* dtop, T* dW, T* dbias,
                            long long M, int D, int V) {
  MMS_REQUIRE(M >= 0 && D > 0 && V > 0, MMS_E_INVALID, "bad size");
  if (M == 0 || (!dW && !dbias)) return 0;
  MMS_REQUIRE(idx && dtop, MMS_E_INVALID, "null pointer");
  if (ctx->embed_deterministic)      // order-independent 64-bit fixed-point sums (embed_det.cu); never the atomic path
    return mms_embed_backward_deterministic<T>(ctx, idx, dtop, dW, dbias, M, D, V);
  // enough CTAs to cover the machine about four times over, 8..64 rows each
  const long long want = mms_ceil_div(M, 4LL * ctx->sm_count);
  const int rows_per_cta = (int)mms_min<long long>(kBwdMaxRows, mms_max<long long>(8, (want + 7) / 8 * 8));
  const long long grid = (M + rows_per_cta - 1) / rows_per_cta;
  MMS_REQUIRE(grid <= 0x7fffffffLL, MMS_E_UNSUPPORTED, "too many rows");
  if (!launch_backward_v4<T>(ctx, idx, dtop, dW, dbias, M, D, V, rows_per_cta, grid)) {
    MmsKernelScope ks_(ctx, "embed_backward_runs");
    embed_backward_runs<T><<<(unsigned)grid, kBwdThreads, 0, ctx->stream>>>(idx, dtop, dW, dbias, M, D, V,
                                                                         rows_per_cta, ctx->fault_flag);
  }
  MMS_LAUNCH_CHECK();
  return 0;
}

template int mms_embed_forward_impl<float>(mms_context*, const float*, const float*, const float*, float*, long long, int, int);
template int mms_embed_forward_impl<double>(mms_context*, const double*, const double*, const double*, double*, long long, int, int);
template int mms_embed_backward_impl<float>(mms_context*, const float*, const float*, float*, float*, long long, int, int);
template int mms_embed_backward_impl<double>(mms_context*, const double*, const double*, double*, double*, long long, int, int);
