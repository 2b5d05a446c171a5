// SimMatrix layer (reference: src/caffe/layers/sim_matrix_layer.{cpp,cu}) and the
// candidate-scoring (reranking) entry point built on the same bilinear form.
//   forward : T = q W  (kept, as the reference keeps it in bottom[1].diff), s_n = <a_n, T_n>
//   backward: dW += q^T diag(ds) a ;  dq = diag(ds) a W^T ;  da = diag(ds) q W
// The reference's backward is N cblas_sger / cblas_sgemv calls on the host
// (sim_matrix_layer.cpp:73-93); here each gradient is one GEMM over the whole batch.
#include "mms_common.cuh"
#include "tc/tc_gemm.cuh"

namespace {

// s[n] = sum_c a[n,c] * T[n,c]  -- one warp per row, shuffle reduction
template <typename T>
__global__ void rowdot_kernel(const T* __restrict__ a, const T* __restrict__ Tm, T* __restrict__ s,
                              int N, int K2) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int n = warp; n < N; n += nwarps) {
    T acc = T(0);
    for (int c = lane; c < K2; c += 32) acc += a[(size_t)n * K2 + c] * Tm[(size_t)n * K2 + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s[n] = acc;
  }
}

// out[n,c] = ds[n] * in[n,c]   (out may be in: no __restrict__ on the pair)
template <typename T>
__global__ void rowscale_kernel(const T* in, const T* __restrict__ ds, T* out, long long total, int K) {
  if (sizeof(T) == 4 && (K & 3) == 0 && total < 0x7fffffffLL && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const unsigned K4 = (unsigned)K >> 2, n4 = (unsigned)(total >> 2);            // 16-byte groups, 32-bit index arithmetic
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += gridDim.x * blockDim.x) {
      const float s = (float)ds[e / K4];
      float4 v = reinterpret_cast<const float4*>(in)[e];
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      reinterpret_cast<float4*>(out)[e] = v;
    }
    return;
  }
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x)
    out[e] = ds[e / K] * in[e];
}

template <typename T>
int gemm2d(mms_context* ctx, const T* A, long long sAm, long long sAk, const T* B, long long sBk,
           long long sBn, T* C, int ldc, int M, int N, int K, T beta, int ksplit = 1) {
  SimtGemmArgs<T> g;
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K;
  g.sAm = sAm; g.sAk = sAk; g.sBk = sBk; g.sBn = sBn; g.ldc = ldc;
  g.sA1 = g.sA2 = g.sB1 = g.sB2 = g.sC1 = g.sC2 = 0;
  g.nb1 = g.nb2 = 1; g.alpha = T(1); g.beta = beta; g.ksplit = ksplit;
  // grid.y limit: tile rows of 64
  if (mms_ceil_div(M, 64) > 65535) {
    // split the M range
    const int step = 65535 * 64;
    for (int m0 = 0; m0 < M; m0 += step) {
      SimtGemmArgs<T> h = g;
      h.A = A + (long long)m0 * sAm; h.C = C + (long long)m0 * ldc; h.M = min(step, M - m0);
      MMS_TRY(mms_simt_gemm<T>(ctx, h));
    }
    return 0;
  }
  return mms_simt_gemm<T>(ctx, g);
}

inline int ew_grid(mms_context* ctx, long long n) {
  return (int)mms_min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
}

}  // namespace

// float + MMS_MATH_TF32: the three contractions run on the TMA-fed tcgen05 GEMM (tc/tc_gemm_tma.cu).
// Its operands must be TF32-exact, so q, a and W pass through one rounding launch into the scratch
// buffer (leading dimensions padded to 4 floats); diag(ds) is folded into that pass, so the scaled
// copies ds o a and ds o q cost nothing extra.
inline bool use_tc(mms_context* ctx, const float*) { return ctx->math == MMS_MATH_TF32; }
inline bool use_tc(mms_context*, const double*) { return false; }

// scratch layout shared by forward and backward: [qr | Wr | ar | as]; the forward already asks for all of it, so a
// backward on the same handle finds qr and Wr in place (MMS_OPT_REUSE_FORWARD) and rounds only the answer side
inline size_t tc_scratch_floats(int N, long long K1p, long long K2p, int K1) {
  return (size_t)N * K1p + (size_t)K1 * K2p + 2 * (size_t)N * K2p;
}

inline int tc_forward(mms_context* ctx, const float* q, const float* W, float* Tm, int N, int K1, int K2) {
  const long long K1p = tc_pad4(K1), K2p = tc_pad4(K2);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(float) * tc_scratch_floats(N, K1p, K2p, K1), &sp));
  float* qr = static_cast<float*>(sp);
  float* Wr = qr + (size_t)N * K1p;
  const RoundJob jobs[2] = {{q, qr, N, K1, K1, K1p, nullptr}, {W, Wr, K1, K2, K2, K2p, nullptr}};
  MMS_TRY(mms_tf32_round(ctx, jobs, 2));
  TcGemmArgs g = tc_gemm_args(qr, K1p, 0, Wr, K2p, 1, Tm, K2, N, K2, K1);   // B(n=c,k=t) = W[t][c]: MN-major
  g.operands_tf32 = 1;
  MMS_TRY(mms_tc_gemm(ctx, g));
  ctx->simmat_cache.valid = true; ctx->simmat_cache.generation = mms_write_clock(); ctx->simmat_cache.q = q; ctx->simmat_cache.W = W;
  ctx->simmat_cache.N = N; ctx->simmat_cache.K1 = K1; ctx->simmat_cache.K2 = K2;
  ctx->simmat_cache.T = Tm;
  return 0;
}
inline int tc_forward(mms_context*, const double*, const double*, double*, int, int, int) { return MMS_E_UNSUPPORTED; }

inline int tc_backward(mms_context* ctx, const float* q, const float* a, const float* W, const float* ds, float* dW,
                       float* dq, float* da, int N, int K1, int K2) {
  const long long K1p = tc_pad4(K1), K2p = tc_pad4(K2);
  const bool cached = ctx->reuse_forward && ctx->simmat_cache.valid &&
                      mms_unchanged_since(ctx->simmat_cache.generation, q, sizeof(float) * (size_t)N * K1) &&
                      mms_unchanged_since(ctx->simmat_cache.generation, W, sizeof(float) * (size_t)K1 * K2) &&
                      ctx->simmat_cache.q == q &&
                      ctx->simmat_cache.W == W && ctx->simmat_cache.N == N && ctx->simmat_cache.K1 == K1 &&
                      ctx->simmat_cache.K2 == K2;
  const void* before = ctx->scratch;
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(float) * tc_scratch_floats(N, K1p, K2p, K1), &sp));
  const bool have_qw = cached && sp == before;
  float* qr = static_cast<float*>(sp);          // q
  float* Wr = qr + (size_t)N * K1p;
  float* ar = Wr + (size_t)K1 * K2p;            // a
  float* as = ar + (size_t)N * K2p;             // ds o a
  // diag(ds) is applied where it is free: folded into the rounding pass for dW's operand (ds o a), and as the
  // epilogue's row scale for dq = diag(ds) (a W^T) and da = diag(ds) (q W) -- no scaled copy of q is made
  RoundJob jobs[4];
  int nj = 0;
  if (!have_qw) {
    jobs[nj++] = RoundJob{q, qr, N, K1, K1, K1p, nullptr};
    jobs[nj++] = RoundJob{W, Wr, K1, K2, K2, K2p, nullptr};
  }
  if (dq) jobs[nj++] = RoundJob{a, ar, N, K2, K2, K2p, nullptr};
  if (dW) jobs[nj++] = RoundJob{a, as, N, K2, K2, K2p, ds};
  if (nj) MMS_TRY(mms_tf32_round(ctx, jobs, nj));
  if (dW) {   // dW[r][c] += sum_n q[n][r] (ds[n] a[n][c]): both operands MN-major (rows = sample n = K index)
    TcGemmArgs g = tc_gemm_args(qr, K1p, 1, as, K2p, 1, dW, K2, K1, K2, N, TC_ATOMIC);
    const int tiles = mms_ceil_div(K1, 128) * mms_ceil_div(K2, 256);
    g.ksplit = mms_max(1, mms_min(ctx->sm_count / mms_max(tiles, 1), mms_ceil_div(N, 256)));   // one full wave
    g.operands_tf32 = 1;
    MMS_TRY(mms_tc_gemm(ctx, g));
  }
  if (dq) {   // dq[n][r] = ds[n] sum_c a[n][c] W[r][c]: A K-major, B(n=r,k=c) = W[r][c] K-major
    TcGemmArgs g = tc_gemm_args(ar, K2p, 0, Wr, K2p, 0, dq, K1, N, K1, K2);
    g.operands_tf32 = 1;
    g.out_rowscale = ds;
    MMS_TRY(mms_tc_gemm(ctx, g));
  }
  // The forward parked T = q W where the reference parks it: in bottom[1]'s diff (sim_matrix_layer.cpp:58), which is
  // the very buffer da is written to.  With MMS_OPT_REUSE_FORWARD and T untouched since, da_n = ds_n W^T q_n = ds_n T_n
  // (:88-90) is a row scaling of that buffer in place -- 2 N K2 floats of traffic instead of a fourth N x K2 x K1 GEMM.
  if (da && cached && ctx->simmat_cache.T == da &&
      mms_unchanged_since(ctx->simmat_cache.generation, da, sizeof(float) * (size_t)N * K2)) {
    { MmsKernelScope ks_(ctx, "rowscale_kernel");
      rowscale_kernel<float><<<ew_grid(ctx, (long long)N * K2), 256, 0, ctx->stream>>>(da, ds, da, (long long)N * K2, K2); }
    MMS_LAUNCH_CHECK();
    ctx->simmat_cache.T = nullptr;            // T is gone: a second backward recomputes
    return 0;
  }
  if (da) {   // da = diag(ds) (q W)
    TcGemmArgs g = tc_gemm_args(qr, K1p, 0, Wr, K2p, 1, da, K2, N, K2, K1);
    g.operands_tf32 = 1;
    g.out_rowscale = ds;
    MMS_TRY(mms_tc_gemm(ctx, g));
  }
  return 0;
}
inline int tc_backward(mms_context*, const double*, const double*, const double*, const double*, double*, double*,
                       double*, int, int, int) { return MMS_E_UNSUPPORTED; }

template <typename T>
int mms_simmatrix_forward_impl(mms_context* ctx, const T* q, const T* a, const T* W, T* s, T* Tm,
                               int N, int K1, int K2) {
  MMS_REQUIRE(q && a && W && s && Tm, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && K1 > 0 && K2 > 0, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  // T = q W   (gemm NoTrans,NoTrans M_ x K2 x K1, sim_matrix_layer.cpp:60-61)
  if (use_tc(ctx, q)) MMS_TRY(tc_forward(ctx, q, W, Tm, N, K1, K2));
  else MMS_TRY(gemm2d<T>(ctx, q, K1, 1, W, K2, 1, Tm, K2, N, K2, K1, T(0)));
  { MmsKernelScope ks_(ctx, "rowdot_kernel");
    rowdot_kernel<T><<<ew_grid(ctx, (long long)N * 32), 256, 0, ctx->stream>>>(a, Tm, s, N, K2); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_simmatrix_backward_impl(mms_context* ctx, const T* q, const T* a, const T* W, const T* ds,
                                T* dW, T* dq, T* da, int N, int K1, int K2, int prop_w, int prop0,
                                int prop1) {
  MMS_REQUIRE(q && a && W && ds, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && K1 > 0 && K2 > 0, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  if (use_tc(ctx, q))
    return tc_backward(ctx, q, a, W, ds, (prop_w && dW) ? dW : nullptr, (prop0 && dq) ? dq : nullptr,
                       (prop1 && da) ? da : nullptr, N, K1, K2);
  const bool need_as = (prop_w && dW) || (prop0 && dq);
  const bool need_qs = (prop1 && da);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(T) * ((size_t)N * K2 * need_as + (size_t)N * K1 * need_qs), &sp));
  T* As = static_cast<T*>(sp);
  T* Qs = As + (size_t)N * K2 * need_as;
  if (need_as) {
    { MmsKernelScope ks_(ctx, "rowscale_kernel");
      rowscale_kernel<T><<<ew_grid(ctx, (long long)N * K2), 256, 0, ctx->stream>>>(a, ds, As, (long long)N * K2, K2); }
    MMS_LAUNCH_CHECK();
  }
  if (need_qs) {
    { MmsKernelScope ks_(ctx, "rowscale_kernel");
      rowscale_kernel<T><<<ew_grid(ctx, (long long)N * K1), 256, 0, ctx->stream>>>(q, ds, Qs, (long long)N * K1, K1); }
    MMS_LAUNCH_CHECK();
  }
  if (prop_w && dW) {
    // dW += q^T (ds o a)      (sum of the reference's N rank-1 updates, :75-79)
    const int tiles = mms_ceil_div(K1, 64) * mms_ceil_div(K2, 64);
    const int ksplit = max(1, min(mms_ceil_div(2 * ctx->sm_count, tiles), mms_ceil_div(N, 256)));
    MMS_TRY(gemm2d<T>(ctx, q, 1, K1, As, K2, 1, dW, K2, K1, K2, N, T(1), ksplit));
  }
  if (prop0 && dq)   // dq = (ds o a) W^T     (gemv NoTrans per sample, :88-90)
    MMS_TRY(gemm2d<T>(ctx, As, K2, 1, W, 1, K2, dq, K1, N, K1, K2, T(0)));
  if (prop1 && da)   // da = (ds o q) W       (gemv Trans per sample)
    MMS_TRY(gemm2d<T>(ctx, Qs, K1, 1, W, K2, 1, da, K2, N, K2, K1, T(0)));
  return 0;
}

// scores = (Q W) C^T : the SimMatrix bilinear form for every (query, candidate) pair.
int mms_rerank_scores_impl(mms_context* ctx, const float* Q, const float* C, const float* W, float* QW,
                           float* scores, int Nq, long long Nc, int K1, int K2) {
  MMS_REQUIRE(Q && C && W && QW && scores, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(Nq > 0 && Nc > 0 && K1 > 0 && K2 > 0, MMS_E_INVALID, "bad size");
  MMS_REQUIRE(Nc <= 0x7fffffffLL, MMS_E_UNSUPPORTED, "candidate count exceeds int range");
  if (ctx->math == MMS_MATH_TF32) {
    // QW = Q W (kept un-rounded for the caller), then scores = QW C^T slab by slab: the candidates of
    // a slab are rounded to TF32 into the scratch buffer (they are read-once data, HBM-bound) and the
    // slab GEMM is TMA-fed; Q W is rounded once.
    const long long K1p = tc_pad4(K1), K2p = tc_pad4(K2);
    const size_t fixed = (size_t)Nq * K1p + (size_t)K1 * K2p + (size_t)Nq * K2p;
    long long slab = ((long long)(ctx->scratch_cap / sizeof(float)) - (long long)fixed) / K2p;
    slab = mms_max<long long>(1024, mms_min<long long>(slab, mms_min<long long>(Nc, 1 << 18)));
    // (Slabs small enough to stay dirty in L2 between the rounding pass and the GEMM -- 2 x 40 MB -- were tried: the
    // rounded copy then never reaches HBM, but 106 two-tile GEMM launches cost more than that saves: 5.0 ms vs 4.1 ms.)
    const bool pipelined = ctx->concurrency != 0 && slab < Nc;   // two slab buffers: round slab i+1 beside the GEMM of slab i
    if (pipelined) slab = mms_max<long long>(1024, mms_min<long long>(slab, (((long long)(ctx->scratch_cap / sizeof(float)) - (long long)fixed) / K2p) / 2));
    void* sp = nullptr;
    MMS_TRY(mms_scratch(ctx, sizeof(float) * (fixed + (size_t)slab * K2p * (pipelined ? 2 : 1)), &sp));
    float* Qr = static_cast<float*>(sp);
    float* Wr = Qr + (size_t)Nq * K1p;
    float* QWr = Wr + (size_t)K1 * K2p;
    float* Cr = QWr + (size_t)Nq * K2p;
    const RoundJob j0[2] = {{Q, Qr, Nq, K1, K1, K1p, nullptr}, {W, Wr, K1, K2, K2, K2p, nullptr}};
    MMS_TRY(mms_tf32_round(ctx, j0, 2));
    TcGemmArgs t = tc_gemm_args(Qr, K1p, 0, Wr, K2p, 1, QW, K2, Nq, K2, K1);
    t.operands_tf32 = 1;
    MMS_TRY(mms_tc_gemm(ctx, t));
    const RoundJob j1[1] = {{QW, QWr, Nq, K2, K2, K2p, nullptr}};
    MMS_TRY(mms_tf32_round(ctx, j1, 1));
    if (pipelined) {
      // The rounding pass is HBM-bound and uses no shared memory: it runs on a private stream on the SMs the persistent
      // GEMM occupies, one slab ahead of it (event fork/join, valid under stream capture).
      int i = 0;
      {
        const RoundJob j2[1] = {{C, Cr, mms_min<long long>(slab, Nc), K2, K2, K2p, nullptr}};
        MMS_TRY(mms_tf32_round(ctx, j2, 1));
      }
      for (long long c0 = 0; c0 < Nc; c0 += slab, ++i) {
        const long long nc = mms_min<long long>(slab, Nc - c0), next0 = c0 + slab;
        float* cur = Cr + (size_t)(i & 1) * slab * K2p;
        float* nxt = Cr + (size_t)((i + 1) & 1) * slab * K2p;
        const bool more = next0 < Nc;
        if (more) {
          MMS_TRY(mms_fork(ctx, 0));                           // after the GEMM that last read `nxt`
          MmsStreamSwitch sw_(ctx, 0);
          const RoundJob j2[1] = {{C + (size_t)next0 * K2, nxt, mms_min<long long>(slab, Nc - next0), K2, K2, K2p, nullptr}};
          MMS_TRY(mms_tf32_round(ctx, j2, 1));
        }
        TcGemmArgs g = tc_gemm_args(QWr, K2p, 0, cur, K2p, 0, scores + c0, Nc, Nq, (int)nc, K2);   // both K-major
        g.operands_tf32 = 1;
        MMS_TRY(mms_tc_gemm(ctx, g));
        if (more) MMS_TRY(mms_join(ctx, 0));
      }
      return 0;
    }
    for (long long c0 = 0; c0 < Nc; c0 += slab) {
      const long long nc = mms_min<long long>(slab, Nc - c0);
      const RoundJob j2[1] = {{C + (size_t)c0 * K2, Cr, nc, K2, K2, K2p, nullptr}};
      MMS_TRY(mms_tf32_round(ctx, j2, 1));
      TcGemmArgs g = tc_gemm_args(QWr, K2p, 0, Cr, K2p, 0, scores + c0, Nc, Nq, (int)nc, K2);   // both K-major
      g.operands_tf32 = 1;
      MMS_TRY(mms_tc_gemm(ctx, g));
    }
    return 0;
  }
  MMS_TRY(gemm2d<float>(ctx, Q, K1, 1, W, K2, 1, QW, K2, Nq, K2, K1, 0.f));
  // scores[i][j] = sum_c QW[i][c] * C[j][c]; scores row stride = Nc
  SimtGemmArgs<float> g;
  g.A = QW; g.B = C; g.C = scores; g.M = Nq; g.N = (int)Nc; g.K = K2;
  g.sAm = K2; g.sAk = 1; g.sBk = 1; g.sBn = K2; g.ldc = (int)Nc;
  g.sA1 = g.sA2 = g.sB1 = g.sB2 = g.sC1 = g.sC2 = 0;
  g.nb1 = g.nb2 = 1; g.alpha = 1.f; g.beta = 0.f; g.ksplit = 1;
  return mms_simt_gemm<float>(ctx, g);
}

// A candidate set that is scored again and again (a static index) is rounded once: mms_rerank_prepare writes the
// TF32-rounded copy the GEMM reads (rows padded to 4 floats) into a buffer the caller keeps, and
// mms_rerank_scores_prepared scores against it -- the per-call traffic drops from 16 to 8 GB at 10^6 x 1024.
int mms_rerank_prepare_impl(mms_context* ctx, const float* C, float* Cr, long long Nc, int K2) {
  MMS_REQUIRE(C && Cr && Nc > 0 && K2 > 0, MMS_E_INVALID, "bad argument");
  const RoundJob j[1] = {{C, Cr, Nc, K2, K2, tc_pad4(K2), nullptr}};
  return mms_tf32_round(ctx, j, 1);
}

int mms_rerank_scores_prepared_impl(mms_context* ctx, const float* Q, const float* Cr, const float* W, float* QW,
                                    float* scores, int Nq, long long Nc, int K1, int K2) {
  MMS_REQUIRE(Q && Cr && W && QW && scores, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(Nq > 0 && Nc > 0 && K1 > 0 && K2 > 0, MMS_E_INVALID, "bad size");
  MMS_REQUIRE(Nc <= 0x7fffffffLL, MMS_E_UNSUPPORTED, "candidate count exceeds int range");
  MMS_REQUIRE(ctx->math == MMS_MATH_TF32, MMS_E_UNSUPPORTED, "prepared candidates are a TF32 operand");
  const long long K1p = tc_pad4(K1), K2p = tc_pad4(K2);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(float) * ((size_t)Nq * K1p + (size_t)K1 * K2p + (size_t)Nq * K2p), &sp));
  float* Qr = static_cast<float*>(sp);
  float* Wr = Qr + (size_t)Nq * K1p;
  float* QWr = Wr + (size_t)K1 * K2p;
  const RoundJob j0[2] = {{Q, Qr, Nq, K1, K1, K1p, nullptr}, {W, Wr, K1, K2, K2, K2p, nullptr}};
  MMS_TRY(mms_tf32_round(ctx, j0, 2));
  TcGemmArgs t = tc_gemm_args(Qr, K1p, 0, Wr, K2p, 1, QW, K2, Nq, K2, K1);
  t.operands_tf32 = 1;
  MMS_TRY(mms_tc_gemm(ctx, t));
  const RoundJob j1[1] = {{QW, QWr, Nq, K2, K2, K2p, nullptr}};
  MMS_TRY(mms_tf32_round(ctx, j1, 1));
  TcGemmArgs g = tc_gemm_args(QWr, K2p, 0, Cr, K2p, 0, scores, Nc, Nq, (int)Nc, K2);   // both K-major
  g.operands_tf32 = 1;
  return mms_tc_gemm(ctx, g);
}

template int mms_simmatrix_forward_impl<float>(mms_context*, const float*, const float*, const float*, float*, float*, int, int, int);
template int mms_simmatrix_forward_impl<double>(mms_context*, const double*, const double*, const double*, double*, double*, int, int, int);
template int mms_simmatrix_backward_impl<float>(mms_context*, const float*, const float*, const float*, const float*, float*, float*, float*, int, int, int, int, int, int);
template int mms_simmatrix_backward_impl<double>(mms_context*, const double*, const double*, const double*, const double*, double*, double*, double*, int, int, int, int, int, int);
