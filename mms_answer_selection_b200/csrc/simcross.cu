// SimCross layer (reference: src/caffe/layers/sim_cross_layer.{cpp,cu}).
//
// mode 2 (learned bilinear Q M_k A^T + B_k) in float goes to the tcgen05 kernels in
// tc/simcross_tc.cu; this file holds (a) the dispatch, (b) the SIMT composition used for
// double blobs, MMS_MATH_FP32 and shapes the tensor-core kernels do not cover, and
// (c) modes 0/1 (cosine, 1/(1+euclid)), which are elementwise/reduction work.
#include "mms_common.cuh"

namespace {

// ---------------------------------------------------------------- modes 0 / 1 --------
// sim_cross_layer.cpp:96-139.  One thread per output (n, j, k); the D-long reductions
// run in the reference's order (dd ascending) in T.
template <typename T>
__global__ void row_norm_kernel(const T* __restrict__ x, T* __restrict__ nrm, long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    T acc = T(0);
    for (int d = lane; d < D; d += 32) { const T v = x[r * D + d]; acc += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) nrm[r] = sqrt(acc);
  }
}

template <typename T, int MODE>
__global__ void sim01_forward_kernel(const T* __restrict__ q, const T* __restrict__ a,
                                     const T* __restrict__ n0, const T* __restrict__ n1,
                                     T* __restrict__ S, int N, int Lq, int La, int D) {
  const long long total = (long long)N * Lq * La;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % La);
    const int j = (int)((e / La) % Lq);
    const int i = (int)(e / ((long long)La * Lq));
    const T* x = q + ((size_t)i * Lq + j) * D;
    const T* y = a + ((size_t)i * La + k) * D;
    T acc = T(0);
    if (MODE == 1) {
      for (int dd = 0; dd < D; ++dd) { const T diff = x[dd] - y[dd]; acc += diff * diff; }
      S[e] = T(1) / (T(1) + sqrt(acc));
    } else {
      for (int dd = 0; dd < D; ++dd) acc += x[dd] * y[dd];
      S[e] = acc / n0[(size_t)i * Lq + j] / n1[(size_t)i * La + k];
    }
  }
}

// sim_cross_layer.cpp:208-250.  SIDE 0: one thread per dq element (n, j, dd) summing over
// the La answers; SIDE 1: one thread per da element (n, m, dd) summing over the Lq
// questions -- the same per-element addition order as the reference's loops, no atomics.
template <typename T, int MODE, int SIDE>
__global__ void sim01_backward_kernel(const T* __restrict__ q, const T* __restrict__ a,
                                      const T* __restrict__ S, const T* __restrict__ dS,
                                      const T* __restrict__ n0, const T* __restrict__ n1,
                                      T* __restrict__ dout, int N, int Lq, int La, int D) {
  const int Lo = SIDE == 0 ? Lq : La;   // rows of the output side
  const int Li = SIDE == 0 ? La : Lq;   // reduced side
  const long long total = (long long)N * Lo * D;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int dd = (int)(e % D);
    const int w = (int)((e / D) % Lo);
    const int i = (int)(e / ((long long)D * Lo));
    T acc = T(0);
    for (int v = 0; v < Li; ++v) {
      const int j = SIDE == 0 ? w : v;      // question word
      const int m = SIDE == 0 ? v : w;      // answer word
      const size_t t = ((size_t)i * Lq + j) * La + m;
      const T qv = q[((size_t)i * Lq + j) * D + dd];
      const T av = a[((size_t)i * La + m) * D + dd];
      const T s = S[t], g = dS[t];
      if (MODE == 1) {
        // the reference's 1e-9 literal is a double: the division is carried out in double
        const T tt = (T)((g * s * s * s * (qv - av)) / ((double)(s - T(1)) + 1e-9));
        acc += (SIDE == 0) ? tt : -tt;
      } else {
        const T r0 = n0[(size_t)i * Lq + j], r1 = n1[(size_t)i * La + m];
        if (SIDE == 0) acc += g * (av / r0 / r1 - qv * s / (r0 * r0));
        else           acc += g * (qv / r0 / r1 - av * s / (r1 * r1));
      }
    }
    dout[e] = acc;
  }
}

// S[n, :, :, :] += B (caffe_add per sample, sim_cross_layer.cpp:155-159)
template <typename T>
__global__ void add_bias_kernel(T* __restrict__ S, const T* __restrict__ B, long long total, int per) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x)
    S[e] = B[e % per] + S[e];
}

// dB[e] += sum_n dS[n, e]  (sim_cross_layer.cpp:301-304), N split over blockIdx.y
template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ dS, T* __restrict__ dB, int N, int per) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per) return;
  const int chunk = (N + gridDim.y - 1) / gridDim.y;
  const int nb = blockIdx.y * chunk, ne = min(N, nb + chunk);
  T acc = T(0);
  for (int n = nb; n < ne; ++n) acc += dS[(size_t)n * per + e];
  if (ne > nb) atomicAdd(dB + e, acc);
}

// float blobs with 16-byte rows: one float4 column group per thread, eight samples' loads in flight per trip
__global__ void bias_grad_vec_kernel(const float* __restrict__ dS, float* __restrict__ dB, int N, int per4) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per4) return;
  const int chunk = (N + gridDim.y - 1) / gridDim.y;
  const int nb = blockIdx.y * chunk, ne = min(N, nb + chunk);
  const float4* src = reinterpret_cast<const float4*>(dS) + e;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int n = nb;
  for (; n + 8 <= ne; n += 8) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcs(src + (size_t)(n + j) * per4);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
  }
  for (; n < ne; ++n) {
    const float4 v = __ldcs(src + (size_t)n * per4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (ne > nb) {
    float* d = dB + 4 * (size_t)e;
    atomicAdd(d, acc.x); atomicAdd(d + 1, acc.y); atomicAdd(d + 2, acc.z); atomicAdd(d + 3, acc.w);
  }
}

inline int ew_grid(mms_context* ctx, long long n) {
  return (int)mms_min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
}

// ---------------------------------------------------------------- mode 2, SIMT --------
template <typename T>
int chunk_pairs(mms_context* ctx, int N, int Lmax, int D, int mc) {
  const size_t per_pair = sizeof(T) * (size_t)mc * Lmax * D;
  long long c = (long long)(ctx->scratch_cap / (per_pair ? per_pair : 1));
  c = mms_min<long long>(c, 65535 / mc);
  c = mms_min<long long>(c, N);
  return (int)mms_max<long long>(c, 1);
}

template <typename T>
int gemm(mms_context* ctx, const T* A, long long sAm, long long sAk, long long sA1, long long sA2,
         const T* B, long long sBk, long long sBn, long long sB1, long long sB2, T* C, int ldc,
         long long sC1, long long sC2, int M, int N, int K, int nb1, int nb2, T beta, int ksplit = 1) {
  SimtGemmArgs<T> g;
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K;
  g.sAm = sAm; g.sAk = sAk; g.sBk = sBk; g.sBn = sBn; g.ldc = ldc;
  g.sA1 = sA1; g.sA2 = sA2; g.sB1 = sB1; g.sB2 = sB2; g.sC1 = sC1; g.sC2 = sC2;
  g.nb1 = nb1; g.nb2 = nb2; g.alpha = T(1); g.beta = beta; g.ksplit = ksplit;
  return mms_simt_gemm<T>(ctx, g);
}

template <typename T>
int simcross2_forward_simt(mms_context* ctx, const T* q, const T* a, const T* Mw, const T* B, T* S,
                           int N, int Lq, int La, int D, int mc) {
  const int nc_max = chunk_pairs<T>(ctx, N, Lq, D, mc);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(T) * (size_t)mc * nc_max * Lq * D, &sp));
  T* Tk = static_cast<T*>(sp);
  for (int n0 = 0; n0 < N; n0 += nc_max) {
    const int nc = min(nc_max, N - n0);
    const T* qc = q + (size_t)n0 * Lq * D;
    const T* ac = a + (size_t)n0 * La * D;
    T* Sc = S + (size_t)n0 * mc * Lq * La;
    // T[k][n] = Q_n M_k                                   (sim_cross_layer.cpp:148-149)
    MMS_TRY(gemm<T>(ctx, qc, D, 1, 0, (long long)Lq * D, Mw, D, 1, (long long)D * D, 0, Tk, D,
                    (long long)nc * Lq * D, (long long)Lq * D, Lq, D, D, mc, nc, T(0)));
    // S[n][k] = T[k][n] A_n^T                             (:151-153)
    MMS_TRY(gemm<T>(ctx, Tk, D, 1, (long long)nc * Lq * D, (long long)Lq * D, ac, 1, D, 0,
                    (long long)La * D, Sc, La, (long long)Lq * La, (long long)mc * Lq * La, Lq, La, D,
                    mc, nc, T(0)));
    if (B) {
      const long long total = (long long)nc * mc * Lq * La;
      { MmsKernelScope ks_(ctx, "add_bias_kernel");
        add_bias_kernel<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(Sc, B, total, mc * Lq * La); }
      MMS_LAUNCH_CHECK();
    }
  }
  return 0;
}

template <typename T>
int simcross2_backward_simt(mms_context* ctx, const T* q, const T* a, const T* Mw, const T* dS, T* dq,
                            T* da, T* dM, int N, int Lq, int La, int D, int mc) {
  const int Lmax = max(Lq, La);
  const int nc_max = chunk_pairs<T>(ctx, N, Lmax, D, mc);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(T) * (size_t)mc * nc_max * Lmax * D, &sp));
  T* buf = static_cast<T*>(sp);
  const int tiles = mms_ceil_div(D, 64) * mms_ceil_div(D, 64) * mc;
  for (int n0 = 0; n0 < N; n0 += nc_max) {
    const int nc = min(nc_max, N - n0);
    const T* qc = q + (size_t)n0 * Lq * D;
    const T* ac = a + (size_t)n0 * La * D;
    const T* Gc = dS + (size_t)n0 * mc * Lq * La;
    T* dqc = dq + (size_t)n0 * Lq * D;
    T* dac = da + (size_t)n0 * La * D;
    const long long sU1 = (long long)nc * Lq * D, sU2 = (long long)Lq * D;
    // U[k][n] = G_nk A_n   (Lq x D, K = La)
    MMS_TRY(gemm<T>(ctx, Gc, La, 1, (long long)Lq * La, (long long)mc * Lq * La, ac, D, 1, 0,
                    (long long)La * D, buf, D, sU1, sU2, Lq, D, La, mc, nc, T(0)));
    // dM_k += Q^T U_k      (D x D, K = nc*Lq; = Q^T G A of :286-289 summed over the chunk)
    const int kdim = nc * Lq;
    int ksplit = max(1, min(mms_ceil_div(2 * ctx->sm_count, tiles), mms_ceil_div(kdim, 256)));
    ksplit = min(ksplit, 65535 / mc);
    MMS_TRY(gemm<T>(ctx, qc, 1, D, 0, 0, buf, D, 1, sU1, 0, dM, D, (long long)D * D, 0, D, D, kdim, mc,
                    1, T(1), ksplit));
    // dQ_n += U[k][n] M_k^T (Lq x D, K = D; = G (M_k A^T)^T of :291-294)
    for (int k = 0; k < mc; ++k)
      MMS_TRY(gemm<T>(ctx, buf + (size_t)k * sU1, D, 1, 0, sU2, Mw + (size_t)k * D * D, 1, D, 0, 0, dqc,
                      D, 0, (long long)Lq * D, Lq, D, D, 1, nc, T(1)));
    // T[k][n] = Q_n M_k, then dA_n += G_nk^T T[k][n]  (:296-299)
    MMS_TRY(gemm<T>(ctx, qc, D, 1, 0, (long long)Lq * D, Mw, D, 1, (long long)D * D, 0, buf, D, sU1, sU2,
                    Lq, D, D, mc, nc, T(0)));
    for (int k = 0; k < mc; ++k)
      MMS_TRY(gemm<T>(ctx, Gc + (size_t)k * Lq * La, 1, La, 0, (long long)mc * Lq * La,
                      buf + (size_t)k * sU1, D, 1, 0, sU2, dac, D, 0, (long long)La * D, La, D, Lq, 1, nc,
                      T(1)));
  }
  return 0;
}

// dB += sum_n dS[n]   (sim_cross_layer.cpp:301-304; accumulates, never zeroed by the layer)
template <typename T>
int simcross2_bias_grad(mms_context* ctx, const T* dS, T* dB, int N, int Lq, int La, int mc) {
  const int per = mc * Lq * La;
  if (sizeof(T) == 4 && per % 4 == 0 && (reinterpret_cast<uintptr_t>(dS) & 15) == 0) {
    const int per4 = per / 4, gx = mms_ceil_div(per4, 128);
    dim3 grid(gx, max(1, min(mms_ceil_div(N, 8), mms_ceil_div(8 * ctx->sm_count, gx))));
    { MmsKernelScope ks_(ctx, "bias_grad_kernel");
      MMS_CARVEOUT(bias_grad_vec_kernel);
      bias_grad_vec_kernel<<<grid, 128, 0, ctx->stream>>>(reinterpret_cast<const float*>(dS), reinterpret_cast<float*>(dB), N, per4); }
    MMS_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid(mms_ceil_div(per, 256), max(1, min(N, mms_ceil_div(4 * ctx->sm_count, mms_ceil_div(per, 256)))));
  { MmsKernelScope ks_(ctx, "bias_grad_kernel");
    bias_grad_kernel<T><<<grid, 256, 0, ctx->stream>>>(dS, dB, N, per); }
  MMS_LAUNCH_CHECK();
  return 0;
}

}  // namespace
template <typename T>
int mms_simcross2_bias_grad(mms_context* ctx, const T* dS, T* dB, int N, int Lq, int La, int mc) {
  return simcross2_bias_grad<T>(ctx, dS, dB, N, Lq, La, mc);
}
template int mms_simcross2_bias_grad<float>(mms_context*, const float*, float*, int, int, int, int);
template int mms_simcross2_bias_grad<double>(mms_context*, const double*, double*, int, int, int, int);
namespace {

template <typename T> struct IsFloat { static constexpr bool value = false; };
template <> struct IsFloat<float> { static constexpr bool value = true; };

inline int tc_fwd(mms_context* ctx, const float* q, const float* a, const float* Mw, const float* B,
                  float* S, int N, int Lq, int La, int D, int mc) {
  return mms_tc_simcross2_forward(ctx, q, a, Mw, B, S, N, Lq, La, D, mc);
}
inline int tc_fwd(mms_context*, const double*, const double*, const double*, const double*, double*,
                  int, int, int, int, int) { return MMS_E_UNSUPPORTED; }
inline int tc_bwd(mms_context* ctx, const float* q, const float* a, const float* Mw, const float* dS,
                  float* dq, float* da, float* dM, int N, int Lq, int La, int D, int mc) {
  return mms_tc_simcross2_backward(ctx, q, a, Mw, dS, dq, da, dM, N, Lq, La, D, mc);
}
inline int tc_bwd(mms_context*, const double*, const double*, const double*, const double*, double*,
                  double*, double*, int, int, int, int, int) { return MMS_E_UNSUPPORTED; }

}  // namespace

template <typename T>
int mms_simcross_forward_impl(mms_context* ctx, int mode, const T* q, const T* a, const T* Mw,
                              const T* B, T* S, T* norm0, T* norm1, int N, int Lq, int La, int D,
                              int mc) {
  MMS_REQUIRE(mode >= 0 && mode <= 2, MMS_E_INVALID, "dist_mode must be 0, 1 or 2");
  MMS_REQUIRE(q && a && S, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && Lq > 0 && La > 0 && D > 0, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  if (mode == 2) {
    MMS_REQUIRE(Mw && mc > 0, MMS_E_INVALID, "mode 2 needs M and mesure_count > 0");
    if (IsFloat<T>::value && ctx->math == MMS_MATH_TF32) {
      const int rc = tc_fwd(ctx, q, a, Mw, B, S, N, Lq, La, D, mc);
      if (rc != MMS_E_UNSUPPORTED) return rc;
    }
    MMS_TRY(mms_stage_require_real(q, false, "bottom q"));
    MMS_TRY(mms_stage_require_real(a, false, "bottom a"));
    return simcross2_forward_simt<T>(ctx, q, a, Mw, B, S, N, Lq, La, D, mc);
  }
  MMS_TRY(mms_stage_require_real(q, false, "bottom q"));
  MMS_TRY(mms_stage_require_real(a, false, "bottom a"));
  const long long total = (long long)N * Lq * La;
  if (mode == 0) {
    MMS_REQUIRE(norm0 && norm1, MMS_E_INVALID, "mode 0 needs the row-norm caches");
    { MmsKernelScope ks_(ctx, "row_norm_kernel");
      row_norm_kernel<T><<<ew_grid(ctx, (long long)N * Lq * 32), 256, 0, ctx->stream>>>(q, norm0, (long long)N * Lq, D); }
    MMS_LAUNCH_CHECK();
    { MmsKernelScope ks_(ctx, "row_norm_kernel");
      row_norm_kernel<T><<<ew_grid(ctx, (long long)N * La * 32), 256, 0, ctx->stream>>>(a, norm1, (long long)N * La, D); }
    MMS_LAUNCH_CHECK();
    { MmsKernelScope ks_(ctx, "sim01_forward_kernel");
      sim01_forward_kernel<T, 0><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(q, a, norm0, norm1, S, N, Lq, La, D); }
  } else {
    { MmsKernelScope ks_(ctx, "sim01_forward_kernel");
      sim01_forward_kernel<T, 1><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(q, a, nullptr, nullptr, S, N, Lq, La, D); }
  }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_simcross_backward_impl(mms_context* ctx, int mode, const T* q, const T* a, const T* Mw,
                               const T* S, const T* dS, const T* norm0, const T* norm1, T* dq,
                               T* da, T* dM, T* dB, int N, int Lq, int La, int D, int mc,
                               int prop0, int prop1) {
  MMS_REQUIRE(mode >= 0 && mode <= 2, MMS_E_INVALID, "dist_mode must be 0, 1 or 2");
  MMS_REQUIRE(q && a && dS && dq && da, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && Lq > 0 && La > 0 && D > 0, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  if (!(prop0 || prop1)) {
    // the reference still zeroes both bottom diffs (sim_cross_layer.cpp:176-177)
    MMS_TRY(mms_fill<T>(ctx, dq, (long long)N * Lq * D, T(0)));
    MMS_TRY(mms_fill<T>(ctx, da, (long long)N * La * D, T(0)));
    return 0;
  }
  if (mode == 2) {
    MMS_REQUIRE(Mw && dM && mc > 0, MMS_E_INVALID, "mode 2 needs M, dM and mesure_count > 0");
    int rc = MMS_E_UNSUPPORTED;
    // dB depends on dS only: with MMS_OPT_CONCURRENCY it is reduced on a private stream beside the contractions
    const bool side_bias = dB && ctx->concurrency;
    if (side_bias) {
      MMS_TRY(mms_fork(ctx, 1));
      MmsStreamSwitch sw(ctx, 1);
      MMS_TRY(simcross2_bias_grad<T>(ctx, dS, dB, N, Lq, La, mc));
    }
    struct Join {                       // every exit path joins the side stream back
      mms_context* c; bool on;
      ~Join() { if (on) mms_join(c, 1); }
    } join_bias{ctx, side_bias};
    if (IsFloat<T>::value && ctx->math == MMS_MATH_TF32)
      rc = tc_bwd(ctx, q, a, Mw, dS, dq, da, dM, N, Lq, La, D, mc);
    if (rc == MMS_E_UNSUPPORTED) {
      MMS_TRY(mms_stage_require_real(q, false, "bottom q"));
      MMS_TRY(mms_stage_require_real(a, false, "bottom a"));
      MMS_TRY(mms_fill<T>(ctx, dq, (long long)N * Lq * D, T(0)));
      MMS_TRY(mms_fill<T>(ctx, da, (long long)N * La * D, T(0)));
      MMS_TRY(mms_fill<T>(ctx, dM, (long long)mc * D * D, T(0)));     // :256
      rc = simcross2_backward_simt<T>(ctx, q, a, Mw, dS, dq, da, dM, N, Lq, La, D, mc);
    }
    MMS_TRY(rc);
    if (dB && !side_bias) MMS_TRY(simcross2_bias_grad<T>(ctx, dS, dB, N, Lq, La, mc));
    return 0;
  }
  MMS_REQUIRE(S, MMS_E_INVALID, "modes 0/1 read the forward output");
  MMS_TRY(mms_stage_require_real(q, false, "bottom q"));
  MMS_TRY(mms_stage_require_real(a, false, "bottom a"));
  const long long t0 = (long long)N * Lq * D, t1 = (long long)N * La * D;
  if (mode == 0) {
    MMS_REQUIRE(norm0 && norm1, MMS_E_INVALID, "mode 0 needs the row-norm caches");
    { MmsKernelScope ks_(ctx, "sim01_backward_kernel");
      sim01_backward_kernel<T, 0, 0><<<ew_grid(ctx, t0), 256, 0, ctx->stream>>>(q, a, S, dS, norm0, norm1, dq, N, Lq, La, D); }
    MMS_LAUNCH_CHECK();
    { MmsKernelScope ks_(ctx, "sim01_backward_kernel");
      sim01_backward_kernel<T, 0, 1><<<ew_grid(ctx, t1), 256, 0, ctx->stream>>>(q, a, S, dS, norm0, norm1, da, N, Lq, La, D); }
  } else {
    { MmsKernelScope ks_(ctx, "sim01_backward_kernel");
      sim01_backward_kernel<T, 1, 0><<<ew_grid(ctx, t0), 256, 0, ctx->stream>>>(q, a, S, dS, nullptr, nullptr, dq, N, Lq, La, D); }
    MMS_LAUNCH_CHECK();
    { MmsKernelScope ks_(ctx, "sim01_backward_kernel");
      sim01_backward_kernel<T, 1, 1><<<ew_grid(ctx, t1), 256, 0, ctx->stream>>>(q, a, S, dS, nullptr, nullptr, da, N, Lq, La, D); }
  }
  MMS_LAUNCH_CHECK();
  return 0;
}

#define INST(T)                                                                                    \
  template int mms_simcross_forward_impl<T>(mms_context*, int, const T*, const T*, const T*, const T*, \
                                            T*, T*, T*, int, int, int, int, int);                  \
  template int mms_simcross_backward_impl<T>(mms_context*, int, const T*, const T*, const T*, const T*, \
                                             const T*, const T*, const T*, T*, T*, T*, T*, int, int, \
                                             int, int, int, int, int);
INST(float)
INST(double)
