// Host-side interface of the generic tcgen05 TF32 GEMM (tc/tc_gemm.cu).
#pragma once
#include "../mms_common.cuh"

enum { TC_STORE = 0, TC_ACCUM = 1, TC_ATOMIC = 2 };

// C[z1][z2] (+)= sum over K segments s of  op(A[z1][z2][s]) * op(B[z1][z2][s])   (+ c_add)
//   A operand base for batch (z1, z2), segment s: A + z1*sA1 + z2*sA2 + s*segA   (likewise B)
//   C base: C + z1*sC1 + z2*sC2;   c_add base: c_add + z1*s_add1 + z2*s_add2
struct TcGemmArgs {
  const float* A; long long lda; int a_mn;   // a_mn = 0: A(m,k) = A[m*lda + k];  1: A[k*lda + m]
  const float* B; long long ldb; int b_mn;   // b_mn = 0: B(n,k) = B[n*ldb + k];  1: B[k*ldb + n]
  float* C; long long ldc;                   // C(m,n) = C[m*ldc + n]
  int M, N, K;                               // K = extent of ONE segment
  int nb1, nb2;                              // batch grid (z = z1*nb2 + z2)
  long long sA1, sA2, sB1, sB2, sC1, sC2;    // batch strides in elements
  int nseg; long long segA, segB;            // the reduction runs over nseg segments of K each
  int ksplit;                                // > 1: reduction split over CTAs, needs mode == TC_ATOMIC
  int mode;                                  // TC_STORE: C = acc, TC_ACCUM: C += acc, TC_ATOMIC: atomicAdd
  const float* a_rowscale;                   // optional: scales A's staged rows (K-major: per m, MN-major: per k)
  const float* b_rowscale;                   // optional: scales B's staged rows (K-major: per n, MN-major: per k)
  const float* out_rowscale;                 // optional: acc[m][n] *= out_rowscale[m]
  const float* c_add; long long ld_add, s_add1, s_add2;   // optional: acc[m][n] += c_add[m*ld_add + n]
  int operands_tf32;                         // 1: A and B already hold TF32-exact values (rounded by a previous
                                             //    kernel) -> they may be fetched by TMA without a register pass
  int round_out;                             // 1: round the stored result to TF32 (it feeds another contraction)
  // Threshold epilogue (per-query top-k of candidate scores, tc/../topk.cu): when flt_s != null C is NOT written;
  // acc[m][n] is compared with row m's current k-th list entry (flt_s / flt_i at [m*flt_k + flt_k-1]; order: value
  // descending, index ascending, index = flt_base + n) and survivors are appended to row m's candidate buffer
  // (cand_s / cand_i [m*cand_cap + pos], pos = atomicAdd(cand_n + m, 1); counts beyond cand_cap mean "overflowed").
  const float* flt_s; const long long* flt_i; int flt_k; long long flt_base;
  float* cand_s; long long* cand_i; int* cand_n; int cand_cap;
};

inline TcGemmArgs tc_gemm_args(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn,
                               float* C, long long ldc, int M, int N, int K, int mode = TC_STORE) {
  TcGemmArgs g;
  g.A = A; g.lda = lda; g.a_mn = a_mn; g.B = B; g.ldb = ldb; g.b_mn = b_mn; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K;
  g.nb1 = g.nb2 = 1; g.sA1 = g.sA2 = g.sB1 = g.sB2 = g.sC1 = g.sC2 = 0;
  g.nseg = 1; g.segA = g.segB = 0;
  g.ksplit = 1; g.mode = mode;
  g.a_rowscale = g.b_rowscale = g.out_rowscale = nullptr;
  g.c_add = nullptr; g.ld_add = g.s_add1 = g.s_add2 = 0;
  g.operands_tf32 = 0; g.round_out = 0;
  g.flt_s = nullptr; g.flt_i = nullptr; g.flt_k = 0; g.flt_base = 0;
  g.cand_s = nullptr; g.cand_i = nullptr; g.cand_n = nullptr; g.cand_cap = 0;
  return g;
}

// Dispatch: the TMA-fed kernel (tc_gemm_tma.cu) when the operands are TF32-exact and 16-byte
// aligned in every stride, else the register-staged kernel (tc_gemm.cu), which rounds on the fly.
int mms_tc_gemm(mms_context* ctx, const TcGemmArgs& args);
int mms_tc_gemm_staged(mms_context* ctx, const TcGemmArgs& args);
int mms_tc_gemm_tma(mms_context* ctx, const TcGemmArgs& args);   // MMS_E_UNSUPPORTED if not eligible

// Tensor map (cached per handle) over operand X(mn, k): K-major X[mn*ld + k] (box = rows_box x 32 k) or
// MN-major X[k*ld + mn] (box = k_box k x 32 mn); batch strides s1 (extent nb1), s2 (nb2), segment stride sseg
// (nseg) in elements, 0 = broadcast.  `out` points at a CUtensorMap.  TMA coordinates: K-major
// (k, mn, z2, z1, seg), MN-major (mn, k, z2, z1, seg); out-of-range elements read as zero.
int mms_tc_make_map(mms_context* ctx, void* out, const float* ptr, long long ld, bool mn_major, long long MN,
                    long long K, int rows_box, long long s1, long long s2, long long sseg, int nb1, int nb2,
                    int nseg, int k_box = 32);

int mms_tc_make_map_raw(mms_context* ctx, void* out, const float* ptr, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box, bool atom32b);

// Sentence convolution forward as a dedicated kernel (tc/sentconv_fwd.cu): Yt[c][r] = sum_i sum_d Wr[c][i*D + d] *
// xr[(r + i)*D + d] for r < rows; MMS_E_UNSUPPORTED when the shape does not fit it (C > 128, kh > 8, D % 4).
int mms_tc_sentconv_forward(mms_context* ctx, const float* xr, long long rows_total, const float* Wr, float* Yt,
                            long long rows, int D, int C, int kh, long long ldyt);
// dW[c][i*D + d] += sum_r G[r*ldg + c] * xr[(r + i)*D + d] on a dedicated kernel (MMS_E_UNSUPPORTED: C > 128, kh > 5, ...)
int mms_tc_sentconv_dw(mms_context* ctx, const float* G, long long rows, int ldg, const float* xr, long long rows_total,
                       float* dW, int D, int C, int kh);
// The general form (also the backward's dx): out(m, r) = sum_{i<kh} sum_{k<Kd} F[m*ldf + i*Kdf + k] * X[(r+i)*ldx + k].
int mms_tc_sentconv_shifted(mms_context* ctx, const float* X, long long rows_total, long long ldx, int Kd, const float* F,
                            long long ldf, int Kdf, int Mtot, int kh, float* out, long long ld_out, long long rows,
                            int row_major_out, int dry_run);

// dst[r*ldd + c] = tf32_rna(src[r*lds + c] * (scale ? scale[r] : 1)) for up to 4 matrices in one launch.
struct RoundJob {
  const float* src; float* dst; long long rows; int cols; long long lds, ldd; const float* scale;
};
int mms_tf32_round(mms_context* ctx, const RoundJob* jobs, int njobs);
inline long long tc_pad4(long long n) { return (n + 3) & ~3LL; }
