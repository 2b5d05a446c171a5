// Host-side interface of the generic tcgen05 TF32 GEMM (tc/tc_gemm.cu).
#pragma once
#include "../mms_common.cuh"

enum { TC_STORE = 0, TC_ACCUM = 1, TC_ATOMIC = 2 };

struct TcGemmArgs {
  const float* A; long long lda; int a_mn;   // a_mn = 0: A(m,k) = A[m*lda + k];  1: A[k*lda + m]
  const float* B; long long ldb; int b_mn;   // b_mn = 0: B(n,k) = B[n*ldb + k];  1: B[k*ldb + n]
  float* C; long long ldc;                   // C(m,n) = C[m*ldc + n]
  int M, N, K;
  long long sA, sB, sC;                      // batch strides in elements
  int batch;
  int ksplit;                                // > 1: K split over CTAs, needs mode == TC_ATOMIC
  int mode;                                  // TC_STORE: C = acc, TC_ACCUM: C += acc, TC_ATOMIC: atomicAdd
  const float* a_rowscale;                   // optional: scales A's staged rows (K-major: per m, MN-major: per k)
  const float* b_rowscale;                   // optional: scales B's staged rows (K-major: per n, MN-major: per k)
  const float* out_rowscale;                 // optional: acc[m][n] *= out_rowscale[m]
};

inline TcGemmArgs tc_gemm_args(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn,
                               float* C, long long ldc, int M, int N, int K, int mode = TC_STORE) {
  TcGemmArgs g;
  g.A = A; g.lda = lda; g.a_mn = a_mn; g.B = B; g.ldb = ldb; g.b_mn = b_mn; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.sA = g.sB = g.sC = 0; g.batch = 1; g.ksplit = 1; g.mode = mode;
  g.a_rowscale = g.b_rowscale = g.out_rowscale = nullptr;
  return g;
}

int mms_tc_gemm(mms_context* ctx, const TcGemmArgs& args);
