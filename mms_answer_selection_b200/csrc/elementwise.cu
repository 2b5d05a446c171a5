// PairRankLoss, FM, the loss dot-product and the 1/n gradient scaling: HBM-bound
// elementwise / reduction kernels.  Row and block reductions use warp shuffles; the
// cross-block step is a second fixed-order pass, so results are run-to-run
// reproducible (no floating-point atomics).
#include <type_traits>

#include "mms_common.cuh"
#include "adadelta.cuh"

namespace {

constexpr int kRedThreads = 256;
constexpr int kMaxPartials = 1024;

// round-to-nearest mul/add that the compiler never contracts into an FMA: the hinge
// test `ordered > 0` must see exactly the reference's two-rounding value.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ T block_sum(T v) {
  __shared__ T s[kRedThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s[w] = v;
  __syncthreads();
  v = (threadIdx.x < kRedThreads / 32) ? s[threadIdx.x] : T(0);
  if (w == 0) v = warp_sum(v);
  return v;   // valid in thread 0
}

// second pass: out = sum(partials[0..n)) / divisor
template <typename T>
__global__ void __launch_bounds__(kRedThreads) finish_sum_kernel(const T* __restrict__ partials, int n,
                                                                 T divisor, T* __restrict__ out) {
  T v = T(0);
  for (int i = threadIdx.x; i < n; i += kRedThreads) v += partials[i];
  v = block_sum(v);
  if (threadIdx.x == 0) *out = v / divisor;
}

// ---- PairRankLoss (pair_rank_loss_layer.cu:17-41 / .cpp:26-52, MKL axpby semantics) ----
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
prl_forward_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ y, T margin,
                   long long count, T* __restrict__ ordered, T* __restrict__ similar,
                   T* __restrict__ partials) {
  T acc = T(0);
  for (long long i = blockIdx.x * (long long)kRedThreads + threadIdx.x; i < count;
       i += (long long)gridDim.x * kRedThreads) {
    const T d = a[i] - b[i];
    const T yy = y[i];
    const T o = add_rn(mul_rn(T(-1), mul_rn(d, yy)), margin);   // mul, axpby(-1), add_scalar
    similar[i] = d;
    ordered[i] = o;
    acc += max(T(0), o) + abs(mul_rn(T(1) - yy, d));
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// pair_rank_loss_layer.cu:46-55 (GE) / pair_rank_loss_layer.cpp:70-80 (GT)
template <typename T>
__global__ void prl_backward_kernel(const T* __restrict__ y, const T* __restrict__ ordered,
                                    const T* __restrict__ similar, T sign_a, T sign_b, int ge,
                                    long long count, T* __restrict__ da, T* __restrict__ db) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const T yy = y[i];
    const T o = ordered[i];
    const T ot = (ge ? (o >= T(0)) : (o > T(0))) ? T(1) : T(0);
    const T st = ((T(1) - yy) * similar[i] > T(0)) ? T(1) : T(-1);
    const T core = add_rn(mul_rn(ot, yy), -mul_rn(st, T(1) - yy));
    if (da) da[i] = mul_rn(sign_a, core);
    if (db) db[i] = mul_rn(sign_b, core);
  }
}

// ---- FM (fm_layer.cpp:33-99): one warp per sample ----------------------------------------
template <typename T>
__global__ void fm_forward_kernel(const T* __restrict__ x, const T* __restrict__ bias, T* __restrict__ y,
                                  int N, int C, int Dm) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < N; i += nwarps) {
    const T* xi = x + (size_t)i * C * Dm;
    T t1 = T(0);
    for (int j = 1 + lane; j < Dm; j += 32) {
      T t2 = T(0), sq = T(0);
      for (int k = 0; k < C; ++k) { const T v = xi[k * Dm + j]; t2 += v; sq += v * v; }
      t1 += t2 * t2 - sq;
    }
    t1 = warp_sum(t1) / T(2);
    T lin = T(0);
    for (int k = lane; k < C; k += 32) lin += xi[k * Dm];
    lin = warp_sum(lin);
    if (lane == 0) y[i] = t1 + lin + (bias ? bias[0] : T(0));
  }
}

template <typename T>
__global__ void fm_backward_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                   int N, int C, int Dm) {
  const long long total = (long long)N * Dm;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % Dm);
    const int i = (int)(e / Dm);
    const T g = dy[i];
    const T* xi = x + (size_t)i * C * Dm;
    T* di = dx + (size_t)i * C * Dm;
    if (j == 0) {
      for (int k = 0; k < C; ++k) di[k * Dm] = g;
    } else {
      T tt = T(0);
      for (int k = 0; k < C; ++k) tt += xi[k * Dm + j];
      for (int k = 0; k < C; ++k) di[k * Dm + j] = g * (tt - xi[k * Dm + j]);
    }
  }
}

// out = sum_i x[i] * (yv ? yv[i] : 1) in ONE launch: every block leaves its partial sum, the block that draws the
// last ticket adds the partials in index order (the result does not depend on which block that is) and re-arms
// the ticket for the next launch on this handle.
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
sum_kernel(const T* __restrict__ x, const T* __restrict__ yv, long long n, T* __restrict__ partials,
           unsigned int* __restrict__ ticket, T* __restrict__ out, int vec) {
  T acc = T(0);
  const long long t0 = blockIdx.x * (long long)kRedThreads + threadIdx.x, stride = (long long)gridDim.x * kRedThreads;
  long long done = 0;
  if (vec) {      // 16-byte streaming loads, four per operand in flight (the scalar loop reached 0.58 of the HBM peak)
    constexpr int V = 16 / sizeof(T);
    typedef typename std::conditional<sizeof(T) == 4, float4, double2>::type VT;
    const long long nv = n / V;
    const VT* xv = reinterpret_cast<const VT*>(x);
    const VT* yw = reinterpret_cast<const VT*>(yv);
    T a4[4] = {T(0), T(0), T(0), T(0)};
    long long i = t0;
    for (; i + 3 * stride < nv; i += 4 * stride) {
      VT xs[4], ys[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { xs[j] = __ldcs(xv + i + j * stride); if (yv) ys[j] = __ldcs(yw + i + j * stride); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const T* xe = reinterpret_cast<const T*>(&xs[j]);
        const T* ye = reinterpret_cast<const T*>(&ys[j]);
#pragma unroll
        for (int e = 0; e < V; ++e) a4[j] += yv ? xe[e] * ye[e] : xe[e];
      }
    }
    for (; i < nv; i += stride) {
      const VT xs = __ldcs(xv + i);
      const T* xe = reinterpret_cast<const T*>(&xs);
      if (yv) {
        const VT ys = __ldcs(yw + i);
        const T* ye = reinterpret_cast<const T*>(&ys);
#pragma unroll
        for (int e = 0; e < V; ++e) a4[0] += xe[e] * ye[e];
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) a4[0] += xe[e];
      }
    }
    acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    done = nv * V;
  }
  for (long long i = done + t0; i < n; i += stride)
    acc += yv ? x[i] * yv[i] : x[i];
  acc = block_sum(acc);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = acc;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  T v = T(0);
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kRedThreads) v += __ldcg(partials + i);
  v = block_sum(v);
  if (threadIdx.x == 0) { *out = v; *ticket = 0u; }
}

template <typename T>
__global__ void scale_kernel(T* __restrict__ x, long long n, T alpha) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    x[i] *= alpha;
}

// One AdaDelta step of one blob, every pass the reference's solver makes over it fused into one (8 x 4 bytes per
// parameter instead of ~20 passes): gradient scale (parallel.cpp:377 / sgd_solver.cpp:131), L2 weight decay
// (sgd_solver.cpp:181-185), AdaDeltaUpdate (adadelta_solver.cu:7-16), Blob::Update (data -= diff) and, optionally,
// Net::ClearParamDiffs for the next iteration (solver.cpp:203).
template <typename T>
__global__ void __launch_bounds__(256)
adadelta_step_kernel(T* __restrict__ data, T* __restrict__ diff, T* __restrict__ hg, T* __restrict__ hu, long long n,
                     T grad_scale, T local_decay, T momentum, T delta, T local_rate, int clear, int vec) {
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool has_w = data != nullptr;
  constexpr int V = 16 / sizeof(T);
  typedef typename std::conditional<sizeof(T) == 4, float4, double2>::type VT;
  long long done = 0;
  if (vec) {
    const long long nv = n / V;
    for (long long i = t0; i < nv; i += stride) {
      VT g = reinterpret_cast<VT*>(diff)[i], h = reinterpret_cast<VT*>(hg)[i], h2 = reinterpret_cast<VT*>(hu)[i];
      VT w = g;
      if (has_w) w = reinterpret_cast<VT*>(data)[i];
#define MMS_AD1(c) adadelta_one<T>(w.c, g.c, h.c, h2.c, has_w, grad_scale, local_decay, momentum, delta, local_rate, clear != 0)
      MMS_AD1(x); MMS_AD1(y);
      if constexpr (sizeof(T) == 4) { MMS_AD1(z); MMS_AD1(w); }
#undef MMS_AD1
      if (has_w) reinterpret_cast<VT*>(data)[i] = w;
      reinterpret_cast<VT*>(diff)[i] = g;
      reinterpret_cast<VT*>(hg)[i] = h;
      reinterpret_cast<VT*>(hu)[i] = h2;
    }
    done = nv * V;
  }
  for (long long i = done + t0; i < n; i += stride) {
    T wv = has_w ? data[i] : T(0);
    adadelta_one<T>(wv, diff[i], hg[i], hu[i], has_w, grad_scale, local_decay, momentum, delta, local_rate, clear != 0);
    if (has_w) data[i] = wv;
  }
}

inline int red_grid(mms_context* ctx, long long n) {
  long long g = (n + kRedThreads - 1) / kRedThreads;
  g = mms_min<long long>(g, mms_min<long long>(kMaxPartials, (long long)ctx->sm_count * 4));
  return (int)mms_max<long long>(g, 1);
}
inline int ew_grid(mms_context* ctx, long long n) {
  return (int)mms_max<long long>(1, mms_min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16));
}

}  // namespace

template <typename T>
int mms_pairrankloss_forward_impl(mms_context* ctx, const T* a, const T* b, const T* y, T margin,
                                  long long count, T* loss, T* ordered, T* similar) {
  MMS_REQUIRE(a && b && y && loss && ordered && similar, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(count > 0, MMS_E_INVALID, "count must be positive");
  void* sp = nullptr;
  sp = ctx->partials;     // a private buffer: the scratch buffer may hold SimCross's forward cache
  T* partials = static_cast<T*>(sp);
  const int grid = red_grid(ctx, count);
  { MmsKernelScope ks_(ctx, "prl_forward_kernel");
    prl_forward_kernel<T><<<grid, kRedThreads, 0, ctx->stream>>>(a, b, y, margin, count, ordered, similar, partials); }
  MMS_LAUNCH_CHECK();
  { MmsKernelScope ks_(ctx, "finish_sum_kernel");
    finish_sum_kernel<T><<<1, kRedThreads, 0, ctx->stream>>>(partials, grid, static_cast<T>(count), loss); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_pairrankloss_backward_impl(mms_context* ctx, const T* y, const T* ordered, const T* similar,
                                   T top_diff, long long count, T* da, T* db) {
  MMS_REQUIRE(y && ordered && similar, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(count > 0, MMS_E_INVALID, "count must be positive");
  if (!da && !db) return 0;
  const T base = top_diff / static_cast<T>(count);   // sign *= top.diff / count (:67)
  { MmsKernelScope ks_(ctx, "prl_backward_kernel");
    prl_backward_kernel<T><<<ew_grid(ctx, count), 256, 0, ctx->stream>>>(y, ordered, similar, T(-1) * base,
                                                                       T(1) * base, ctx->prl_ge, count, da, db); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_fm_forward_impl(mms_context* ctx, const T* x, const T* bias, T* y, int N, int C, int Dm) {
  MMS_REQUIRE(x && y, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && C > 0 && Dm > 0, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  { MmsKernelScope ks_(ctx, "fm_forward_kernel");
    fm_forward_kernel<T><<<ew_grid(ctx, (long long)N * 32), 256, 0, ctx->stream>>>(x, bias, y, N, C, Dm); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_fm_backward_impl(mms_context* ctx, const T* x, const T* dy, T* dx, T* dbias, int N, int C, int Dm,
                         int prop0) {
  MMS_REQUIRE(x && dy, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && C > 0 && Dm > 0, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  if (dbias) MMS_TRY(mms_dot_impl<T>(ctx, dy, nullptr, N, dbias));   // db = sum dy, overwritten (:77)
  if (prop0) {
    MMS_REQUIRE(dx, MMS_E_INVALID, "null dx");
    { MmsKernelScope ks_(ctx, "fm_backward_kernel");
      fm_backward_kernel<T><<<ew_grid(ctx, (long long)N * Dm), 256, 0, ctx->stream>>>(x, dy, dx, N, C, Dm); }
    MMS_LAUNCH_CHECK();
  }
  return 0;
}

template <typename T>
int mms_dot_impl(mms_context* ctx, const T* x, const T* y, long long n, T* out) {
  MMS_REQUIRE(x && out, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(n >= 0, MMS_E_INVALID, "bad size");
  void* sp = nullptr;
  sp = ctx->partials;     // a private buffer: the scratch buffer may hold SimCross's forward cache
  T* partials = static_cast<T*>(sp);
  const int grid = red_grid(ctx, n);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(static_cast<double*>(ctx->partials) + 1024);
  { MmsKernelScope ks_(ctx, "sum_kernel");
    MMS_CARVEOUT(sum_kernel<T>);
    const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && n >= 4096;
    sum_kernel<T><<<grid, kRedThreads, 0, ctx->stream>>>(x, y, n, partials, ticket, out, vec); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_scale_impl(mms_context* ctx, T* x, long long n, T alpha) {
  MMS_REQUIRE(x || n == 0, MMS_E_INVALID, "null pointer");
  if (n <= 0) return 0;
  { MmsKernelScope ks_(ctx, "scale_kernel");
    scale_kernel<T><<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(x, n, alpha); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_adadelta_step_impl(mms_context* ctx, T* data, T* diff, T* hist_g, T* hist_u, long long n, T grad_scale,
                           T local_decay, T momentum, T delta, T local_rate, int clear_diff) {
  MMS_REQUIRE(n >= 0, MMS_E_INVALID, "bad size");
  if (n == 0) return 0;
  MMS_REQUIRE(diff && hist_g && hist_u, MMS_E_INVALID, "null pointer");
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int vec = al(diff) && al(hist_g) && al(hist_u) && (!data || al(data));
  { MmsKernelScope ks_(ctx, "adadelta_step_kernel");
    adadelta_step_kernel<T><<<ew_grid(ctx, n / (16 / sizeof(T)) + 1), 256, 0, ctx->stream>>>(
        data, diff, hist_g, hist_u, n, grad_scale, local_decay, momentum, delta, local_rate, clear_diff, vec); }
  MMS_LAUNCH_CHECK();
  return 0;
}

// ---- Dropout (dropout_layer.cu:10-45): y = x * (mask > threshold) * scale, mask = one random 32-bit word per element
template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, const unsigned* __restrict__ mask, T* __restrict__ y, long long n, unsigned threshold,
               T scale) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    y[i] = x[i] * static_cast<T>(mask[i] > threshold) * scale;
}
// counter-based generator (splitmix64 of seed and element index): the same words whatever the launch geometry
__global__ void __launch_bounds__(256) dropout_mask_kernel(unsigned* __restrict__ mask, long long n, unsigned long long seed) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    mask[i] = (unsigned)((z ^ (z >> 31)) >> 32);
  }
}
template <typename T>
int mms_dropout_impl(mms_context* ctx, const T* x, const unsigned* mask, T* y, long long n, unsigned threshold, T scale) {
  MMS_REQUIRE(n >= 0, MMS_E_INVALID, "bad size");
  if (n == 0) return 0;
  MMS_REQUIRE(x && mask && y, MMS_E_INVALID, "null pointer");
  { MmsKernelScope ks_(ctx, "dropout_kernel");
    dropout_kernel<T><<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(x, mask, y, n, threshold, scale); }
  MMS_LAUNCH_CHECK();
  return 0;
}
int mms_dropout_mask_impl(mms_context* ctx, unsigned* mask, long long n, unsigned long long seed) {
  MMS_REQUIRE(n >= 0, MMS_E_INVALID, "bad size");
  if (n == 0) return 0;
  MMS_REQUIRE(mask, MMS_E_INVALID, "null pointer");
  { MmsKernelScope ks_(ctx, "dropout_mask_kernel");
    dropout_mask_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(mask, n, seed); }
  MMS_LAUNCH_CHECK();
  return 0;
}
template int mms_dropout_impl<float>(mms_context*, const float*, const unsigned*, float*, long long, unsigned, float);
template int mms_dropout_impl<double>(mms_context*, const double*, const unsigned*, double*, long long, unsigned, double);

#define INST(T)                                                                                        \
  template int mms_pairrankloss_forward_impl<T>(mms_context*, const T*, const T*, const T*, T, long long, T*, T*, T*); \
  template int mms_pairrankloss_backward_impl<T>(mms_context*, const T*, const T*, const T*, T, long long, T*, T*);    \
  template int mms_fm_forward_impl<T>(mms_context*, const T*, const T*, T*, int, int, int);            \
  template int mms_fm_backward_impl<T>(mms_context*, const T*, const T*, T*, T*, int, int, int, int);  \
  template int mms_dot_impl<T>(mms_context*, const T*, const T*, long long, T*);                       \
  template int mms_scale_impl<T>(mms_context*, T*, long long, T);                                      \
  template int mms_adadelta_step_impl<T>(mms_context*, T*, T*, T*, T*, long long, T, T, T, T, T, int);
INST(float)
INST(double)
