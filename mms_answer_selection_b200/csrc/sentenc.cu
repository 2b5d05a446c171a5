// Sentence encoder of the sentence-vector variant of the net (reference examples/trec_qa_w2v_mms/do_trec_qa_clean.py:
// 352-375, 412-422): Convolution(kernel 5 x D over the (N,1,L,D) embedded sentence) -> BN -> max-over-time Pooling ->
// TanH, whose (N,K) outputs feed SimMatrix.  Reference layers: src/caffe/layers/{base_conv,conv,bn,pooling,tanh}_layer.cpp
// (the stock Caffe classes plus the fork's own "BN").
//
// Sentence convolution = three contractions on the GEMM engines, with NO im2col buffer.  A window of kh token rows of a
// sentence is kh*D consecutive floats of x, so the im2col matrix the reference materialises per sample
// (base_conv_layer.cpp:257-321, util/im2col.cpp) is an overlapping view of x itself: row r = n*L + t starts at x + r*D.
//   forward   Yt[c][r] = sum_k W[c][k] x[r*D + k],  k < kh*D          rows of the x operand overlap (ld D < kh*D)
//   dW        dW[c][k] += sum_r G[r][c] x[r*D + k]                     the same view as the MN-major operand, split over r
//   dx        dx[r][d] = sum_k Gpad[r*ldg + k] Wf[k][d],  k < kh*ldg    kh consecutive rows of the padded gradient
// TMA tensor maps take overlapping rows as they are (tools/tma_overlap_test.py), so each product is ONE plain GEMM.
// Rows r whose window crosses a sentence boundary (t > L-kh) are junk in Y and are dropped by the kernel that
// transposes Y into Caffe's (N,C,T,1) top and adds the bias; G (the top gradient, transposed back to rows, TF32
// rounded) holds zeros there, which is exactly the zero padding dx needs.  float blobs run on the TMA-fed tcgen05
// engine (tc/tc_gemm_tma.cu), double blobs and MMS_MATH_FP32 on the SIMT GEMM with the same views.
//
// BN, pooling and TanH are HBM-bound passes.  BN statistics: CTAs walk whole samples with coalesced reads and keep
// per-thread double accumulators (the channel of a thread's elements is the same in every sample); the elementwise
// passes and the max over time take one warp per (n, c) plane; general pooling is one thread per output, and gather
// (no atomics) for its gradients.
#include <cfloat>

#include "mms_common.cuh"
#include "tc/tc_gemm.cuh"

namespace {

__device__ __forceinline__ float round_operand(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ double round_operand(double v) { return v; }

inline int ew_grid(mms_context* ctx, long long n) {
  return (int)mms_max<long long>(1, mms_min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16));
}

// top[n][c][t] = Y[(n*L + t)*ldy + c] + bias[c], t < T.  One sample per CTA pass: the T x C tile of Y goes through
// shared memory so that both the reads (along c) and the writes (along t) are coalesced.
template <typename T>
__global__ void sentconv_unpack_kernel(const T* __restrict__ Y, const T* __restrict__ bias, T* __restrict__ top, int N,
                                       int L, int Tn, int C, int ldy) {
  extern __shared__ unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);                 // [Tn][C + 1]
  const int ldt = C + 1;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const T* src = Y + (size_t)n * L * ldy;
    for (int e = threadIdx.x; e < Tn * C; e += blockDim.x) {
      const int t = e / C, c = e - t * C;
      tile[t * ldt + c] = src[(size_t)t * ldy + c];
    }
    __syncthreads();
    T* dst = top + (size_t)n * C * Tn;
    for (int e = threadIdx.x; e < C * Tn; e += blockDim.x) {
      const int c = e / Tn, t = e - c * Tn;
      dst[e] = tile[t * ldt + c] + (bias ? bias[c] : T(0));
    }
    __syncthreads();
  }
}

// Channel-major variant for the tensor-core forward, which computes Yt[c][r] (the weights on the 128-row side of the
// MMA, 256 sentence rows per tile): top[n][c][t] = Yt[c*ldyt + n*L + t] + bias[c] is a copy of T-element runs.
template <typename T, int VL>
__global__ void sentconv_unpack_t_kernel(const T* __restrict__ Yt, const T* __restrict__ bias, T* __restrict__ top,
                                         long long nvec, int L, int Tn, int C, long long ldyt) {
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const long long e = v * VL, plane = e / Tn;              // plane = n*C + c; a vector never straddles two planes
    const int t = (int)(e - plane * Tn), c = (int)(plane % C);
    const long long n = plane / C;
    const T b = bias ? bias[c] : T(0);
    const T* src = Yt + (size_t)c * ldyt + n * L + t;
    if (VL == 4) {
      float4 q = *reinterpret_cast<const float4*>(src);
      q.x += b; q.y += b; q.z += b; q.w += b;
      *reinterpret_cast<float4*>(top + e) = q;
    } else {
      top[e] = src[0] + b;
    }
  }
}

// G[(n*L + t)*ldg + c] = round(dtop[n][c][t]) for t < T, 0 for the other rows and the pad columns; and
// dbias[c] += sum_{n,t} dtop[n][c][t] (conv_layer.cpp:47-52, accumulating) with one atomic per channel and CTA.
template <typename T>
__global__ void sentconv_pack_kernel(const T* __restrict__ dtop, T* __restrict__ G, T* __restrict__ dbias, int N, int L,
                                     int Tn, int C, int ldg, int do_round) {
  extern __shared__ unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);                 // [C][Tn + 1]
  T* bsum = tile + (size_t)C * (Tn + 1);                     // [C]
  const int ldt = Tn + 1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) bsum[c] = T(0);
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    __syncthreads();
    const T* src = dtop + (size_t)n * C * Tn;
    for (int e = threadIdx.x; e < C * Tn; e += blockDim.x) {
      const int c = e / Tn, t = e - c * Tn;
      tile[c * ldt + t] = src[e];
    }
    __syncthreads();
    if (dbias)
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        T s = T(0);
        for (int t = 0; t < Tn; ++t) s += tile[c * ldt + t];
        bsum[c] += s;
      }
    if (G) {
      T* dst = G + (size_t)n * L * ldg;
      for (int e = threadIdx.x; e < L * ldg; e += blockDim.x) {
        const int t = e / ldg, c = e - t * ldg;
        T v = (t < Tn && c < C) ? tile[c * ldt + t] : T(0);
        dst[e] = do_round ? round_operand(v) : v;
      }
    }
  }
  __syncthreads();
  if (dbias)
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dbias + c, bsum[c]);
}

// The same with no block-wide barrier: one WARP per (sample, block of 32 channels).  A warp keeps its channel block for
// all the samples it visits, so the bias-gradient partial sums live in registers (lane = channel) and cost one atomic
// per lane at the end; the 32 x T tile is transposed through a private slice of shared memory under __syncwarp.
template <typename T>
__global__ void sentconv_pack_warp_kernel(const T* __restrict__ dtop, T* __restrict__ G, T* __restrict__ dbias, int N, int L,
                                          int Tn, int C, int ldg, int do_round) {
  extern __shared__ unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, ldt = Tn + 1;
  T* tile = reinterpret_cast<T*>(smem_raw) + (size_t)wib * 32 * ldt;      // [32][Tn + 1]
  const int cblocks = (ldg + 31) / 32;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int stride = nwarps / cblocks;                       // samples advance by the number of complete warp groups
  if (warp >= stride * cblocks) return;                      // leftover warps (no block-wide barrier below)
  const int cb = warp % cblocks, c0 = cb * 32, nc = max(0, min(32, C - c0));
  const int col = c0 + lane;
  T bsum = T(0);
  for (int n = warp / cblocks; n < N; n += stride) {
    const T* src = dtop + ((size_t)n * C + c0) * Tn;          // channels c0 .. c0+nc of sample n: nc*Tn contiguous values
    if (sizeof(T) == 4 && (Tn & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const float4* src4 = reinterpret_cast<const float4*>(src);   // 16-byte loads; a vector never straddles two channels
      const int tn4 = Tn >> 2;
      for (int e = lane; e < nc * tn4; e += 32) {
        const int c = e / tn4, t = (e - c * tn4) * 4;
        const float4 v = src4[e];
        T* d = tile + c * ldt + t;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      for (int e = lane; e < nc * Tn; e += 32) {
        const int c = e / Tn;
        tile[c * ldt + (e - c * Tn)] = src[e];
      }
    }
    __syncwarp();
    if (dbias && lane < nc) {
      T s_ = T(0);
      for (int t = 0; t < Tn; ++t) s_ += tile[lane * ldt + t];
      bsum += s_;
    }
    if (G && col < ldg) {
      T* dst = G + (size_t)n * L * ldg + col;
      for (int t = 0; t < L; ++t) {
        const T v = (t < Tn && lane < nc) ? tile[lane * ldt + t] : T(0);
        dst[(size_t)t * ldg] = do_round ? round_operand(v) : v;
      }
    }
    __syncwarp();
  }
  if (dbias && lane < nc) atomicAdd(dbias + col, bsum);
}

// Wf[(i'*Cp + c)*D + d] = round(W[c][kh-1-i'][d]): the kernel rows in reverse order, channel-minor, for dx
template <typename T>
__global__ void sentconv_flip_weights_kernel(const T* __restrict__ W, T* __restrict__ Wf, int C, int Cp, int kh, int D,
                                             int do_round) {
  const long long total = (long long)kh * Cp * D;            // Cp >= C rows per kernel row, the extra ones zero
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(e % D);
    const int c = (int)((e / D) % Cp);
    const int ip = (int)(e / ((long long)D * Cp));
    const T v = c < C ? W[((size_t)c * kh + (kh - 1 - ip)) * D + d] : T(0);
    Wf[e] = do_round ? round_operand(v) : v;
  }
}

// Wt[(d*kh + i')*Cp + c] = round(W[c][kh-1-i'][d]): the same filters with the OUTPUT index d outermost -- K-major rows for
// the dedicated shifted-window kernel (tc/sentconv_fwd.cu), which then computes dx exactly like a forward pass over
// the padded gradient rows
template <typename T>
__global__ void sentconv_flip_weights_t_kernel(const T* __restrict__ W, T* __restrict__ Wt, int C, int Cp, int kh, int D,
                                               int do_round) {
  const long long total = (long long)D * kh * Cp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cp);
    const int ip = (int)((e / Cp) % kh);
    const int d = (int)(e / ((long long)Cp * kh));
    const T v = c < C ? W[((size_t)c * kh + (kh - 1 - ip)) * D + d] : T(0);
    Wt[e] = do_round ? round_operand(v) : v;
  }
}

template <typename T>
int simt_gemm(mms_context* ctx, const T* A, long long sAm, long long sAk, const T* B, long long sBk, long long sBn, T* C,
              int ldc, long long M, int N, int K, T beta, int ksplit) {
  SimtGemmArgs<T> g;
  g.B = B; g.N = N; g.K = K; g.sAm = sAm; g.sAk = sAk; g.sBk = sBk; g.sBn = sBn; g.ldc = ldc;
  g.sA1 = g.sA2 = g.sB1 = g.sB2 = g.sC1 = g.sC2 = 0;
  g.nb1 = g.nb2 = 1; g.alpha = T(1); g.beta = beta; g.ksplit = ksplit;
  const long long step = 65535LL * 64;                      // grid.y limit of the SIMT kernel (row tiles of 64)
  for (long long m0 = 0; m0 < M; m0 += step) {
    g.A = A + m0 * sAm; g.C = C + m0 * ldc; g.M = (int)mms_min<long long>(step, M - m0);
    MMS_TRY(mms_simt_gemm<T>(ctx, g));
  }
  return 0;
}

inline bool tensor_path(mms_context* ctx, const float*, int D) { return ctx->math == MMS_MATH_TF32 && D % 4 == 0; }
inline bool tensor_path(mms_context*, const double*, int) { return false; }

// ---- tensor-core legs (float only; the double overloads exist to keep the templates well-formed)
int tc_conv_forward(mms_context* ctx, const float* xr, const float* Wr, float* Y, long long rows, int D, int C, int kh,
                    int ldy) {
  TcGemmArgs g = tc_gemm_args(xr, D, 0, Wr, (long long)kh * D, 0, Y, ldy, (int)rows, C, D);
  g.nseg = kh; g.segA = D; g.segB = D;                       // segment i: window row i of x against W[:, i, :]
  g.operands_tf32 = 1;
  return mms_tc_gemm(ctx, g);
}
int tc_conv_forward(mms_context*, const double*, const double*, double*, long long, int, int, int, int) {
  return MMS_E_UNSUPPORTED;
}
// The same product with the operands' roles swapped: Yt[c][r], the C filters on the MMA's 128-row side and 256
// sentence rows per tile -- 48 KB of operands per 512 tensor-pipe cycles instead of 30 KB per 224 (the kernel is
// bound by what an SM ingests, not by the tensor pipe), and the top then needs no transpose.
int tc_conv_forward_t(mms_context* ctx, const float* xr, const float* Wr, float* Yt, long long rows, int D, int C, int kh,
                      long long ldyt) {
  {  // the dedicated kernel that stages every slab of token rows once for all kh kernel rows (tc/sentconv_fwd.cu)
    const int rc = mms_tc_sentconv_forward(ctx, xr, rows + kh - 1, Wr, Yt, rows, D, C, kh, ldyt);
    if (rc != MMS_E_UNSUPPORTED) return rc;
  }
  // B(n = r, k) = xr[r*D + k], k < kh*D: rows of the operand overlap (leading dimension D, row length kh*D) -- the
  // tensor map takes that as it is (tools/tma_overlap_test.py), so the kh window rows are one contiguous reduction
  TcGemmArgs g = tc_gemm_args(Wr, (long long)kh * D, 0, xr, D, 0, Yt, ldyt, C, (int)rows, kh * D);
  g.operands_tf32 = 1;
  return mms_tc_gemm(ctx, g);
}
int tc_conv_forward_t(mms_context*, const double*, const double*, double*, long long, int, int, int, long long) {
  return MMS_E_UNSUPPORTED;
}

int tc_conv_dw(mms_context* ctx, const float* G, int ldg, const float* xr, float* dW, long long rows, int D, int C, int kh) {
  // dW[c][i*D + d] += sum_r G[r][c] xr[(r+i)*D + d]: both operands MN-major (the reduction index r is the row), the
  // reduction is split over CTAs
  // B(n, k = r) = xr[r*D + n], n < kh*D: the overlapping view again, so all kh kernel rows are columns of ONE product
  TcGemmArgs g = tc_gemm_args(G, ldg, 1, xr, D, 1, dW, (long long)kh * D, C, kh * D, (int)rows, TC_ATOMIC);
  const int tiles = mms_ceil_div(C, 128) * mms_ceil_div(kh * D, 256);
  // one full wave and no more: 150 tiles on 148 persistent CTAs take as long as 296 (measured: 0.64 ms with 15 splits)
  g.ksplit = (int)mms_max<long long>(1, mms_min<long long>(ctx->sm_count / tiles, (rows + 511) / 512));
  g.operands_tf32 = 1;
  return mms_tc_gemm(ctx, g);
}
int tc_conv_dw(mms_context*, const double*, int, const double*, double*, long long, int, int, int) { return MMS_E_UNSUPPORTED; }
// the dedicated kernel (tc/sentconv_fwd.cu) when the shape fits it, else the generic engine
int tc_conv_dw_any(mms_context* ctx, const float* G, int ldg, const float* xr, float* dW, long long mrows, long long rows,
                   int D, int C, int kh) {
  const int rc = mms_tc_sentconv_dw(ctx, G, mrows, ldg, xr, rows, dW, D, C, kh);
  if (rc != MMS_E_UNSUPPORTED) return rc;
  return tc_conv_dw(ctx, G, ldg, xr, dW, mrows, D, C, kh);
}
int tc_conv_dw_any(mms_context*, const double*, int, const double*, double*, long long, long long, int, int, int) {
  return MMS_E_UNSUPPORTED;
}

int tc_conv_dx(mms_context* ctx, const float* Gpad, int ldg, const float* Wf, float* dx, long long rows, int D, int C,
               int kh) {
  // A(m = r, k) = Gpad[r*ldg + k], k < kh*ldg: kh consecutive gradient rows are one contiguous reduction (overlapping
  // rows again); Wf holds ldg rows per kernel row, the pad rows zero like G's pad columns
  TcGemmArgs g = tc_gemm_args(Gpad, ldg, 0, Wf, D, 1, dx, D, (int)rows, D, kh * ldg);
  g.operands_tf32 = 1;
  return mms_tc_gemm(ctx, g);
}
int tc_conv_dx(mms_context*, const double*, int, const double*, double*, long long, int, int, int) { return MMS_E_UNSUPPORTED; }

// dx on the dedicated shifted-window kernel: out(m = d, r) = sum_i' sum_c Wt[d][i'*ldg + c] Gpad[(r + i')*ldg + c]
int tc_conv_dx_shifted(mms_context* ctx, const float* Gpad, int ldg, const float* Wt, float* dx, long long rows, int D, int C,
                       int kh, int dry_run) {
  return mms_tc_sentconv_shifted(ctx, Gpad, rows + 2 * (kh - 1), ldg, C, Wt, (long long)kh * ldg, ldg, D, kh, dx, D, rows, 1,
                                 dry_run);
}
int tc_conv_dx_shifted(mms_context*, const double*, int, const double*, double*, long long, int, int, int, int) {
  return MMS_E_UNSUPPORTED;
}

int round_copies(mms_context* ctx, const float* x, float* xr, long long rows, int D, const float* W, float* Wr, int C,
                 int kh) {
  const RoundJob jobs[2] = {{x, xr, rows, D, D, D, nullptr}, {W, Wr, C, kh * D, (long long)kh * D, (long long)kh * D, nullptr}};
  return mms_tf32_round(ctx, jobs, Wr ? 2 : 1);
}
int round_copies(mms_context*, const double*, double*, long long, int, const double*, double*, int, int) {
  return MMS_E_UNSUPPORTED;
}

}  // namespace

// ------------------------------------------------------------------------------------------ sentence convolution
template <typename T>
int mms_sentconv_forward_impl(mms_context* ctx, const T* x, const T* W, const T* bias, T* top, int N, int L, int D, int C,
                              int kh) {
  MMS_REQUIRE(x && W && top, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && L > 0 && D > 0 && C > 0 && kh > 0 && kh <= L, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  const int Tn = L - kh + 1;
  const long long rows = (long long)N * L;
  const long long mrows = rows - (kh - 1);                   // windows that stay inside x
  MMS_REQUIRE(rows <= 0x7fffffffLL, MMS_E_UNSUPPORTED, "more than 2^31 token rows");
  const bool tc = tensor_path(ctx, x, D);
  const int ldy = tc ? (int)tc_pad4(C) : C;
  const long long ldyt = tc_pad4(rows);                      // tensor path: Yt[c][r], C rows of ldyt
  // scratch: [xr | Y | Wr].  xr sits at the head and the request already covers what the backward needs
  // ([xr | Gpad | Wf]), so that a backward on this handle finds the rounded x where the forward left it.
  void* sp = nullptr;
  const size_t n_y = tc ? (size_t)C * ldyt : (size_t)rows * ldy, n_xr = tc ? (size_t)rows * D : 0,
               n_wr = tc ? (size_t)C * kh * D : 0;
  const size_t n_bwd = n_xr + (size_t)(rows + 2 * (kh - 1)) * ldy + (size_t)kh * ldy * D + 4;
  MMS_TRY(mms_scratch(ctx, sizeof(T) * mms_max(n_xr + n_y + n_wr, n_bwd), &sp));
  T* xr = static_cast<T*>(sp);
  T* Y = xr + n_xr;
  if (tc) {
    T* Wr = Y + n_y;
    MMS_TRY(round_copies(ctx, x, xr, rows, D, W, Wr, C, kh));
    MMS_TRY(tc_conv_forward_t(ctx, xr, Wr, Y, mrows, D, C, kh, ldyt));
    ctx->sent_cache.valid = true; ctx->sent_cache.generation = mms_write_clock(); ctx->sent_cache.x = x; ctx->sent_cache.rows = rows; ctx->sent_cache.D = D;
    const long long total = (long long)N * C * Tn;
    const bool vec = sizeof(T) == 4 && Tn % 4 == 0 && L % 4 == 0 && (reinterpret_cast<uintptr_t>(top) & 15) == 0;
    { MmsKernelScope ks_(ctx, "sentconv_unpack_t_kernel");
      if (vec) sentconv_unpack_t_kernel<T, 4><<<ew_grid(ctx, total / 4), 256, 0, ctx->stream>>>(Y, bias, top, total / 4, L, Tn, C, ldyt);
      else sentconv_unpack_t_kernel<T, 1><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(Y, bias, top, total, L, Tn, C, ldyt); }
    MMS_LAUNCH_CHECK();
    return 0;
  } else {
    // Y[r][c] = sum_k x[r*D + k] W[c*kh*D + k], k < kh*D
    MMS_TRY(simt_gemm<T>(ctx, x, D, 1, W, 1, (long long)kh * D, Y, ldy, mrows, C, kh * D, T(0), 1));
  }
  const size_t smem = sizeof(T) * (size_t)Tn * (C + 1);
  MMS_REQUIRE(smem <= 200 * 1024, MMS_E_UNSUPPORTED, "output tile of one sentence exceeds shared memory");
  static bool configured[2] = {false, false};
  if (!configured[sizeof(T) == 8]) {
    MMS_MAX_SMEM(sentconv_unpack_kernel<T>, 200 * 1024);
    configured[sizeof(T) == 8] = true;
  }
  { MmsKernelScope ks_(ctx, "sentconv_unpack_kernel");
    sentconv_unpack_kernel<T><<<mms_min(N, ctx->sm_count * 8), 256, smem, ctx->stream>>>(Y, bias, top, N, L, Tn, C, ldy); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_sentconv_backward_impl(mms_context* ctx, const T* x, const T* W, const T* dtop, T* dW, T* dbias, T* dx, int N,
                               int L, int D, int C, int kh) {
  MMS_REQUIRE(x && W && dtop, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && L > 0 && D > 0 && C > 0 && kh > 0 && kh <= L, MMS_E_INVALID, "bad size");
  if (N == 0 || !(dW || dbias || dx)) return 0;
  const int Tn = L - kh + 1;
  const long long rows = (long long)N * L;
  MMS_REQUIRE(rows <= 0x7fffffffLL, MMS_E_UNSUPPORTED, "more than 2^31 token rows");
  const bool tc = tensor_path(ctx, x, D);
  const int ldg = tc ? (int)tc_pad4(C) : C;
  const bool need_g = dW || dx;
  // Gpad: kh-1 zero rows, the N*L gradient rows (row n*L + t holds dtop[n][:, t], zero for t >= T), kh-1 zero rows.
  // G = Gpad + (kh-1) rows is the row-aligned view dW uses; Gpad itself is the shifted, zero-padded view dx uses.
  const size_t n_g = need_g ? (size_t)(rows + 2 * (kh - 1)) * ldg : 0;
  const size_t n_xr = (tc && dW) ? (size_t)rows * D : 0, n_wf = dx ? (size_t)kh * ldg * D : 0;
  const bool cached = ctx->reuse_forward && ctx->sent_cache.valid &&
                      mms_unchanged_since(ctx->sent_cache.generation, x, sizeof(T) * (size_t)rows * D) &&
                      ctx->sent_cache.x == x &&
                      ctx->sent_cache.rows == rows && ctx->sent_cache.D == D;
  const void* before = ctx->scratch;
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(T) * (n_xr + n_g + n_wf + 4), &sp));
  const bool have_xr = cached && sp == before && n_xr > 0;   // the forward's rounded x is still at the head
  T* xr = static_cast<T*>(sp);
  T* Gpad = xr + n_xr;
  T* G = Gpad + (size_t)(kh - 1) * ldg;
  T* Wf = Gpad + n_g;
  if (need_g && kh > 1) {
    MMS_CUDA(cudaMemsetAsync(Gpad, 0, sizeof(T) * (size_t)(kh - 1) * ldg, ctx->stream));
    MMS_CUDA(cudaMemsetAsync(G + (size_t)rows * ldg, 0, sizeof(T) * (size_t)(kh - 1) * ldg, ctx->stream));
  }
  const int cblocks = (ldg + 31) / 32;
  const size_t wsmem = sizeof(T) * 8 * 32 * (size_t)(Tn + 1);           // 8 warps, a 32 x (T+1) tile each
  if (wsmem <= 48 * 1024 && (long long)ctx->sm_count * 4 * 8 >= 2 * cblocks) {
    // every warp keeps one channel block: the warp count must be a multiple of the number of channel blocks
    long long warps = mms_min<long long>((long long)N * cblocks, (long long)ctx->sm_count * 4 * 8);
    warps = mms_max<long long>(cblocks, warps / cblocks * cblocks);
    const int blocks = (int)((warps + 7) / 8);
    { MmsKernelScope ks_(ctx, "sentconv_pack_warp_kernel");
      sentconv_pack_warp_kernel<T><<<blocks, 256, wsmem, ctx->stream>>>(dtop, need_g ? G : nullptr, dbias, N, L, Tn, C, ldg,
                                                                        tc ? 1 : 0); }
    MMS_LAUNCH_CHECK();
  } else {
  const size_t smem = sizeof(T) * ((size_t)C * (Tn + 1) + C);
  MMS_REQUIRE(smem <= 200 * 1024, MMS_E_UNSUPPORTED, "gradient tile of one sentence exceeds shared memory");
  static bool configured[2] = {false, false};
  if (!configured[sizeof(T) == 8]) {
    MMS_MAX_SMEM(sentconv_pack_kernel<T>, 200 * 1024);
    configured[sizeof(T) == 8] = true;
  }
  { MmsKernelScope ks_(ctx, "sentconv_pack_kernel");
    sentconv_pack_kernel<T><<<mms_min(N, ctx->sm_count * 4), 256, smem, ctx->stream>>>(
        dtop, need_g ? G : nullptr, dbias, N, L, Tn, C, ldg, tc ? 1 : 0); }
  MMS_LAUNCH_CHECK();
  }
  const long long mrows = rows - (kh - 1);
  if (dW) {                                                  // accumulates (weight_cpu_gemm, beta = 1; conv_layer.cpp:57-60)
    if (tc) {
      if (!have_xr) MMS_TRY(round_copies(ctx, x, xr, rows, D, nullptr, nullptr, C, kh));
      MMS_TRY(tc_conv_dw_any(ctx, G, ldg, xr, dW, mrows, rows, D, C, kh));
    } else {
      const int tiles = mms_ceil_div(C, 64) * mms_ceil_div(kh * D, 64);
      const int ksplit = (int)mms_max<long long>(1, mms_min<long long>(mms_ceil_div(2 * ctx->sm_count, tiles), (mrows + 255) / 256));
      MMS_TRY(simt_gemm<T>(ctx, G, 1, ldg, x, D, 1, dW, kh * D, C, kh * D, (int)mrows, T(1), ksplit));
    }
  }
  if (dx) {                                                  // overwrites (backward_cpu_gemm + col2im; conv_layer.cpp:62-65)
    const bool shifted = tc && tc_conv_dx_shifted(ctx, Gpad, ldg, Wf, dx, rows, D, C, kh, /*dry_run=*/1) == 0;
    { MmsKernelScope ks_(ctx, "sentconv_flip_weights_kernel");
      if (shifted)
        sentconv_flip_weights_t_kernel<T><<<ew_grid(ctx, (long long)kh * ldg * D), 256, 0, ctx->stream>>>(W, Wf, C, ldg, kh, D, 1);
      else
        sentconv_flip_weights_kernel<T><<<ew_grid(ctx, (long long)kh * ldg * D), 256, 0, ctx->stream>>>(W, Wf, C, ldg, kh, D,
                                                                                                   tc ? 1 : 0); }
    MMS_LAUNCH_CHECK();
    if (shifted) MMS_TRY(tc_conv_dx_shifted(ctx, Gpad, ldg, Wf, dx, rows, D, C, kh, 0));
    else if (tc) MMS_TRY(tc_conv_dx(ctx, Gpad, ldg, Wf, dx, rows, D, C, kh));
    else MMS_TRY(simt_gemm<T>(ctx, Gpad, ldg, 1, Wf, D, 1, dx, D, rows, D, kh * C, T(0), 1));   // ldg == C here
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ pooling (pooling_layer.cpp)
namespace {

template <typename T>
__global__ void pool_forward_kernel(const T* __restrict__ x, T* __restrict__ top, int* __restrict__ mask, long long total,
                                    int H, int W, int PH, int PW, int kh, int kw, int sh, int sw, int ph_, int pw_,
                                    int method) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(e % PW), ph = (int)((e / PW) % PH);
    const long long nc = e / ((long long)PW * PH);
    const T* src = x + nc * H * W;
    int hs = ph * sh - ph_, ws = pw * sw - pw_;
    if (method == 0) {                                       // MAX: first maximum in scan order wins (:150-163)
      const int he = min(hs + kh, H), we = min(ws + kw, W);
      hs = max(hs, 0); ws = max(ws, 0);
      T best = -FLT_MAX;
      int arg = -1;
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w)
          if (src[h * W + w] > best) { best = src[h * W + w]; arg = h * W + w; }
      top[e] = best;
      if (mask) mask[e] = arg;
    } else {                                                 // AVE: the divisor counts the padding (:186-203)
      int he = min(hs + kh, H + ph_), we = min(ws + kw, W + pw_);
      const int pool_size = (he - hs) * (we - ws);
      hs = max(hs, 0); ws = max(ws, 0); he = min(he, H); we = min(we, W);
      T s = T(0);
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) s += src[h * W + w];
      top[e] = s / pool_size;
    }
  }
}

// AVE pooling whose windows tile the plane exactly (kernel == stride, no padding, H % kh == 0, W % kw == 0 -- the
// 4 x 4 / stride 4 pooling of the CNN over the similarity tensor): every input element belongs to one window, so the
// gradient is a broadcast of dtop / (kh kw); 32-bit index arithmetic, one pass at HBM speed.
template <typename T>
__global__ void __launch_bounds__(256)
pool_ave_tiled_backward_kernel(const T* __restrict__ dtop, T* __restrict__ dx, unsigned total, unsigned H, unsigned W,
                               unsigned PW, unsigned kh, unsigned kw, T inv) {
  const unsigned PHW = (H / kh) * PW;
  if ((W & 3u) == 0 && sizeof(T) == 4 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0) {   // four outputs of one row per thread
    const unsigned W4 = W >> 2, total4 = total >> 2;
    for (unsigned e = blockIdx.x * 256u + threadIdx.x; e < total4; e += gridDim.x * 256u) {
      const unsigned w0 = (e % W4) << 2, r = e / W4, h = r % H, nc = r / H;
      const T* g = dtop + nc * PHW + (h / kh) * PW;
      float4 o;
      o.x = (float)(g[w0 / kw] * inv); o.y = (float)(g[(w0 + 1) / kw] * inv);
      o.z = (float)(g[(w0 + 2) / kw] * inv); o.w = (float)(g[(w0 + 3) / kw] * inv);
      __stcs(reinterpret_cast<float4*>(dx) + e, o);
    }
    return;
  }
  for (unsigned e = blockIdx.x * 256u + threadIdx.x; e < total; e += gridDim.x * 256u) {
    const unsigned w = e % W, r = e / W, h = r % H, nc = r / H;
    dx[e] = dtop[nc * PHW + (h / kh) * PW + w / kw] * inv;
  }
}

template <typename T>
__global__ void pool_backward_kernel(const T* __restrict__ dtop, const int* __restrict__ mask, T* __restrict__ dx,
                                     long long total, int H, int W, int PH, int PW, int kh, int kw, int sh, int sw, int ph_,
                                     int pw_, int method) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(e % W), h = (int)((e / W) % H);
    const long long nc = e / ((long long)W * H);
    const int phs = (h + ph_ < kh) ? 0 : (h + ph_ - kh) / sh + 1, phe = min((h + ph_) / sh + 1, PH);
    const int pws = (w + pw_ < kw) ? 0 : (w + pw_ - kw) / sw + 1, pwe = min((w + pw_) / sw + 1, PW);
    const T* g = dtop + nc * PH * PW;
    T s = T(0);
    if (method == 0) {
      const int* m = mask + nc * PH * PW;
      for (int ph = phs; ph < phe; ++ph)
        for (int pw = pws; pw < pwe; ++pw)
          if (m[ph * PW + pw] == h * W + w) s += g[ph * PW + pw];
    } else {
      for (int ph = phs; ph < phe; ++ph)
        for (int pw = pws; pw < pwe; ++pw) {
          const int hs = ph * sh - ph_, ws = pw * sw - pw_;
          const int he = min(hs + kh, H + ph_), we = min(ws + kw, W + pw_);
          s += g[ph * PW + pw] / ((he - hs) * (we - ws));
        }
    }
    dx[e] = s;
  }
}

// One window covering the whole (H, W) plane (max over time: kernel (L-kh+1) x 1): one WARP per plane, coalesced
// reads, shuffle arg-max in which the smaller index wins a tie (= the first maximum of the reference's scan).
template <typename T>
__global__ void pool_plane_max_kernel(const T* __restrict__ x, T* __restrict__ top, int* __restrict__ mask, long long NC,
                                      int HW) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < NC; p += nwarps) {
    const T* src = x + p * HW;
    T best = -FLT_MAX;
    int arg = 0x7fffffff;
    for (int i = lane; i < HW; i += 32) {
      const T v = src[i];
      if (v > best) { best = v; arg = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) { top[p] = best; mask[p] = arg == 0x7fffffff ? -1 : arg; }
  }
}
// The same for planes whose size is a multiple of the 16-byte vector: one THREAD per plane, all of its loads in flight
// at once (a warp-per-plane pass over 36 elements is latency-bound: two dependent loads and a shuffle chain per plane).
template <typename T, typename V, int VL>
__global__ void pool_plane_max_vec_kernel(const T* __restrict__ x, T* __restrict__ top, int* __restrict__ mask, long long NC,
                                          int HW) {
  const int nv = HW / VL;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < NC; p += (long long)gridDim.x * blockDim.x) {
    const V* src = reinterpret_cast<const V*>(x + p * HW);
    T best = -FLT_MAX;
    int arg = -1;
#pragma unroll 4
    for (int v = 0; v < nv; ++v) {
      const V q = src[v];
      const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int j = 0; j < VL; ++j)
        if (e[j] > best) { best = e[j]; arg = v * VL + j; }
    }
    top[p] = best;
    mask[p] = arg;
  }
}

template <typename T>
__global__ void pool_plane_max_backward_kernel(const T* __restrict__ dtop, const int* __restrict__ mask, T* __restrict__ dx,
                                               long long total, int HW) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / HW;
    dx[e] = mask[p] == (int)(e - p * HW) ? dtop[p] : T(0);
  }
}

template <typename T, typename V, int VL>
__global__ void pool_plane_max_backward_vec_kernel(const T* __restrict__ dtop, const int* __restrict__ mask,
                                                   T* __restrict__ dx, long long nvec, int HW) {
  V* dd = reinterpret_cast<V*>(dx);
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const long long p = (v * VL) / HW;                       // a vector never straddles two planes (HW % VL == 0)
    const int at = mask[p] - (int)(v * VL - p * HW);
    const T g = dtop[p];
    V o;
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int j = 0; j < VL; ++j) oe[j] = at == j ? g : T(0);
    dd[v] = o;
  }
}

}  // namespace

template <typename T>
int mms_pool_forward_impl(mms_context* ctx, const T* x, T* top, int* mask, long long NC, int H, int W, int PH, int PW,
                          int kh, int kw, int sh, int sw, int pad_h, int pad_w, int method) {
  MMS_REQUIRE(x && top && (method == 1 || mask), MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(NC >= 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && pad_h >= 0 &&
              pad_w >= 0 && (method == 0 || method == 1), MMS_E_INVALID, "bad argument");
  const long long total = NC * PH * PW;
  if (total == 0) return 0;
  if (method == 0 && PH == 1 && PW == 1 && pad_h == 0 && pad_w == 0 && kh >= H && kw >= W) {
    const int HW = H * W, VL = 16 / (int)sizeof(T);
    if (HW % VL == 0 && HW <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
      MmsKernelScope ks_(ctx, "pool_plane_max_vec_kernel");
      if (sizeof(T) == 4)
        pool_plane_max_vec_kernel<T, float4, 16 / sizeof(T)><<<ew_grid(ctx, NC), 128, 0, ctx->stream>>>(x, top, mask, NC, HW);
      else
        pool_plane_max_vec_kernel<T, double2, 16 / sizeof(T)><<<ew_grid(ctx, NC), 128, 0, ctx->stream>>>(x, top, mask, NC, HW);
    } else {
      MmsKernelScope ks_(ctx, "pool_plane_max_kernel");
      pool_plane_max_kernel<T><<<ew_grid(ctx, NC * 32), 256, 0, ctx->stream>>>(x, top, mask, NC, HW);
    }
    MMS_LAUNCH_CHECK();
    return 0;
  }
  { MmsKernelScope ks_(ctx, "pool_forward_kernel");
    pool_forward_kernel<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(x, top, mask, total, H, W, PH, PW, kh, kw, sh, sw,
                                                                       pad_h, pad_w, method); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_pool_backward_impl(mms_context* ctx, const T* dtop, const int* mask, T* dx, long long NC, int H, int W, int PH,
                           int PW, int kh, int kw, int sh, int sw, int pad_h, int pad_w, int method) {
  MMS_REQUIRE(dtop && dx && (method == 1 || mask), MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(NC >= 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && pad_h >= 0 &&
              pad_w >= 0 && (method == 0 || method == 1), MMS_E_INVALID, "bad argument");
  const long long total = NC * H * W;
  if (total == 0) return 0;
  if (method == 0 && PH == 1 && PW == 1 && pad_h == 0 && pad_w == 0 && kh >= H && kw >= W) {
    const int HW = H * W, VL = 16 / (int)sizeof(T);
    if (HW % VL == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0) {
      MmsKernelScope ks_(ctx, "pool_plane_max_backward_vec_kernel");
      if (sizeof(T) == 4)
        pool_plane_max_backward_vec_kernel<T, float4, 16 / sizeof(T)><<<ew_grid(ctx, total / VL), 256, 0, ctx->stream>>>(
            dtop, mask, dx, total / VL, HW);
      else
        pool_plane_max_backward_vec_kernel<T, double2, 16 / sizeof(T)><<<ew_grid(ctx, total / VL), 256, 0, ctx->stream>>>(
            dtop, mask, dx, total / VL, HW);
    } else {
      MmsKernelScope ks_(ctx, "pool_plane_max_backward_kernel");
      pool_plane_max_backward_kernel<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(dtop, mask, dx, total, HW);
    }
    MMS_LAUNCH_CHECK();
    return 0;
  }
  if (method == 1 && pad_h == 0 && pad_w == 0 && kh == sh && kw == sw && H % kh == 0 && W % kw == 0 && PH == H / kh &&
      PW == W / kw && total < 0xffffffffLL) {
    MmsKernelScope ks_(ctx, "pool_ave_tiled_backward_kernel");
    pool_ave_tiled_backward_kernel<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(
        dtop, dx, (unsigned)total, (unsigned)H, (unsigned)W, (unsigned)PW, (unsigned)kh, (unsigned)kw, T(1) / T(kh * kw));
    MMS_LAUNCH_CHECK();
    return 0;
  }
  { MmsKernelScope ks_(ctx, "pool_backward_kernel");
    pool_backward_kernel<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(dtop, mask, dx, total, H, W, PH, PW, kh, kw, sh,
                                                                        sw, pad_h, pad_w, method); }
  MMS_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------ TanH (tanh_layer.cpp)
namespace {
template <typename T>
__global__ void tanh_forward_kernel(const T* __restrict__ x, T* __restrict__ y, long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    y[e] = tanh(x[e]);
}
template <typename T>
__global__ void tanh_backward_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const T t = y[e];
    dx[e] = dy[e] * (T(1) - t * t);                          // tanh_layer.cpp:27-33
  }
}
}  // namespace

template <typename T>
int mms_tanh_forward_impl(mms_context* ctx, const T* x, T* y, long long n) {
  MMS_REQUIRE(n >= 0 && (n == 0 || (x && y)), MMS_E_INVALID, "bad argument");
  if (n == 0) return 0;
  { MmsKernelScope ks_(ctx, "tanh_forward_kernel");
    tanh_forward_kernel<T><<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(x, y, n); }
  MMS_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int mms_tanh_backward_impl(mms_context* ctx, const T* y, const T* dy, T* dx, long long n) {
  MMS_REQUIRE(n >= 0 && (n == 0 || (y && dy && dx)), MMS_E_INVALID, "bad argument");
  if (n == 0) return 0;
  { MmsKernelScope ks_(ctx, "tanh_backward_kernel");
    tanh_backward_kernel<T><<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(y, dy, dx, n); }
  MMS_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------ BN (the fork's bn_layer.cpp)
namespace {

// acc[c] += sum a, acc[C + c] += sum b over a slice of the N*HW elements of channel c; a/b chosen by MODE:
//   0: (x, x^2)   forward statistics        1: (g, g * xn)   backward sums
template <typename T, int MODE>
__global__ void bn_channel_sums_kernel(const T* __restrict__ p, const T* __restrict__ q, double* __restrict__ acc, int N,
                                       int C, int HW, int slices) {
  const int c = blockIdx.x / slices, slice = blockIdx.x - c * slices;
  const long long per = (long long)N * HW;
  double s0 = 0, s1 = 0;
  for (long long e = slice * (long long)blockDim.x + threadIdx.x; e < per; e += (long long)slices * blockDim.x) {
    const long long n = e / HW, i = e - n * HW;
    const size_t at = ((size_t)n * C + c) * HW + i;
    const double a = (double)p[at];
    s0 += a;
    s1 += MODE == 0 ? a * a : a * (double)q[at];
  }
  __shared__ double red[2][32];
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s0 = threadIdx.x < nw ? red[0][threadIdx.x] : 0.0;
    s1 = threadIdx.x < nw ? red[1][threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if (threadIdx.x == 0) { atomicAdd(acc + c, s0); atomicAdd(acc + C + c, s1); }
  }
}

// Large planes (HW >= 256: the 36 x 36 planes of the CNN over the similarity tensor): one CTA per (channel, slice of the
// samples) streams whole planes -- HW contiguous elements -- with no index division in the loop; partial sums in the
// blob's type per thread (a few hundred values), double across threads and CTAs.
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
bn_channel_sums_planes_kernel(const T* __restrict__ p, const T* __restrict__ q, double* __restrict__ acc, int N, int C, int HW,
                              int slices) {
  const int c = blockIdx.x / slices, slice = blockIdx.x - c * slices;
  T t0 = T(0), t1 = T(0);
  for (int n = slice; n < N; n += slices) {
    const size_t base = ((size_t)n * C + c) * HW;
    for (int i = threadIdx.x; i < HW; i += 256) {
      const T a = p[base + i];
      t0 += a;
      t1 += MODE == 0 ? a * a : a * q[base + i];
    }
  }
  double s0 = (double)t0, s1 = (double)t1;
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s0 = threadIdx.x < 8 ? red[0][threadIdx.x] : 0.0;
    s1 = threadIdx.x < 8 ? red[1][threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if (threadIdx.x == 0) { atomicAdd(acc + c, s0); atomicAdd(acc + C + c, s1); }
  }
}

// The same sums with coalesced reads, for C*HW <= 16 * 1024: a CTA walks whole samples (C*HW contiguous elements each);
// element j + k*blockDim of every sample belongs to the same channel, so thread j keeps one pair of register
// accumulators per k and the channel reduction happens once, at the end, through shared memory.
template <typename T, int MODE>
__global__ void bn_channel_sums_rows_kernel(const T* __restrict__ p, const T* __restrict__ q, double* __restrict__ acc, int N,
                                            int C, int HW) {
  extern __shared__ unsigned char bn_smem_raw[];
  T* part = reinterpret_cast<T*>(bn_smem_raw);              // [C*HW]: the per-thread partial sums, one pass per sum
  const int per = C * HW;
  // per-thread partial sums over this CTA's share of the samples (a few dozen values each) stay in the blob's own
  // type -- no FP64 or conversions in the streaming loop; everything across threads and CTAs is summed in double
  T d0[16], d1[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) d0[k] = d1[k] = T(0);
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const size_t base = (size_t)n * per;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int e = threadIdx.x + k * blockDim.x;
      if (e < per) {
        const T a = p[base + e];
        d0[k] += a;
        d1[k] += MODE == 0 ? a * a : a * q[base + e];
      }
    }
  }
  // channel reduction without atomics on shared memory (a 36-way same-address CAS loop per element is slower than
  // the whole streaming pass): park the partials, then thread c adds up the HW entries of channel c
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int e = threadIdx.x + k * blockDim.x;
      if (e < per) part[e] = pass == 0 ? d0[k] : d1[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double s = 0.0;
      for (int i = 0; i < HW; ++i) s += (double)part[c * HW + i];
      atomicAdd(acc + pass * C + c, s);
    }
  }
}

// TRAIN: mean / variance of the batch (var = E[x^2] - E[x]^2, :131-165), running statistics blended with bn_memory
// (:168-172);  TEST: the running statistics (:177-182).  stat[c] = mean, stat[C + c] = sqrt(var + eps) (:206-210).
template <typename T>
__global__ void bn_finalize_stats_kernel(const double* __restrict__ acc, T* __restrict__ run_mean, T* __restrict__ run_var,
                                         T* __restrict__ mean_out, T* __restrict__ std_out, int C, double count, int train,
                                         T memory, T eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  T mean, var;
  if (train) {
    mean = static_cast<T>(acc[c] / count);
    const T ex2 = static_cast<T>(acc[C + c] / count);
    var = ex2 - mean * mean;
    run_mean[c] = (T(1) - memory) * mean + memory * run_mean[c];
    run_var[c] = (T(1) - memory) * var + memory * run_var[c];
  } else {
    mean = run_mean[c];
    var = run_var[c];
  }
  mean_out[c] = mean;
  std_out[c] = static_cast<T>(pow(var + eps, T(0.5)));
}

// Elementwise passes walk whole (n, c) planes so that the channel -- and with it mean, std, scale, shift and the two
// backward sums -- is fixed per plane instead of being re-derived per element: one warp per plane, lanes across HW.
template <typename T>
__global__ void bn_normalize_planes_kernel(const T* __restrict__ x, const T* __restrict__ mean, const T* __restrict__ stdv,
                                           const T* __restrict__ scale, const T* __restrict__ shift, T* __restrict__ xn,
                                           T* __restrict__ top, long long planes, int C, int HW) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < planes; p += nwarps) {
    const int c = (int)(p % C);
    const T m = mean[c], sd = stdv[c], sc = scale[c], sh = shift[c];
    const size_t base = (size_t)p * HW;
    for (int i = lane; i < HW; i += 32) {
      const T v = (x[base + i] - m) / sd;
      xn[base + i] = v;
      top[base + i] = v * sc + sh;
    }
  }
}
template <typename T>
__global__ void bn_backward_planes_kernel(const T* __restrict__ g, const T* __restrict__ xn, const T* __restrict__ scale,
                                          const T* __restrict__ stdv, const double* __restrict__ acc, T* __restrict__ dx,
                                          long long planes, int C, int HW, double count) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < planes; p += nwarps) {
    const int c = (int)(p % C);
    const T sc = scale[c], sd = stdv[c];
    const T sum_t = static_cast<T>(acc[c]) * sc, sum_xt = static_cast<T>(acc[C + c]) * sc, cnt = static_cast<T>(count);
    const size_t base = (size_t)p * HW;
    for (int i = lane; i < HW; i += 32) {
      const T t = g[base + i] * sc;
      dx[base + i] = (t - (xn[base + i] * sum_xt + sum_t) / cnt) / sd;
    }
  }
}

// Planes whose size is a multiple of the 16-byte vector: a flat, fully coalesced pass of 16-byte accesses; a vector never
// straddles two planes, so the channel is derived once per vector.  (One thread per plane was tried: its two strided
// 16-byte store streams halve the bandwidth of the normalising pass.)
template <typename T, typename V, int VL>
__global__ void bn_normalize_vec_kernel(const T* __restrict__ x, const T* __restrict__ mean, const T* __restrict__ stdv,
                                        const T* __restrict__ scale, const T* __restrict__ shift, T* __restrict__ xn,
                                        T* __restrict__ top, long long nvec, int C, int HW) {
  const V* src = reinterpret_cast<const V*>(x);
  V* dn = reinterpret_cast<V*>(xn);
  V* dt = reinterpret_cast<V*>(top);
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(((v * VL) / HW) % C);
    const T m = mean[c], sd = stdv[c], sc = scale[c], sh = shift[c];
    V a = src[v], n_, t_;
    const T* ae = reinterpret_cast<const T*>(&a);
    T* ne = reinterpret_cast<T*>(&n_);
    T* te = reinterpret_cast<T*>(&t_);
#pragma unroll
    for (int j = 0; j < VL; ++j) { ne[j] = (ae[j] - m) / sd; te[j] = ne[j] * sc + sh; }
    dn[v] = n_;
    dt[v] = t_;
  }
}
template <typename T, typename V, int VL>
__global__ void bn_backward_vec_kernel(const T* __restrict__ g, const T* __restrict__ xn, const T* __restrict__ scale,
                                       const T* __restrict__ stdv, const double* __restrict__ acc, T* __restrict__ dx,
                                       long long nvec, int C, int HW, double count) {
  const V* gs = reinterpret_cast<const V*>(g);
  const V* ns = reinterpret_cast<const V*>(xn);
  V* dd = reinterpret_cast<V*>(dx);
  const T cnt = static_cast<T>(count);
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(((v * VL) / HW) % C);
    const T sc = scale[c], sd = stdv[c];
    const T sum_t = static_cast<T>(acc[c]) * sc, sum_xt = static_cast<T>(acc[C + c]) * sc;
    const V gv = gs[v], nvv = ns[v];
    V o;
    const T* ge = reinterpret_cast<const T*>(&gv);
    const T* ne = reinterpret_cast<const T*>(&nvv);
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int j = 0; j < VL; ++j) {
      const T t = ge[j] * sc;
      oe[j] = (t - (ne[j] * sum_xt + sum_t) / cnt) / sd;
    }
    dd[v] = o;
  }
}
inline bool vec_ok(const void* a, const void* b, const void* c, int HW, int VL) {
  return HW % VL == 0 &&
         ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

template <typename T>
__global__ void bn_normalize_kernel(const T* __restrict__ x, const T* __restrict__ mean, const T* __restrict__ stdv,
                                    const T* __restrict__ scale, const T* __restrict__ shift, T* __restrict__ xn,
                                    T* __restrict__ top, long long total, int C, int HW) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((e / HW) % C);
    const T v = (x[e] - mean[c]) / stdv[c];
    xn[e] = v;                                               // "Saving x_norm" :225-227
    top[e] = v * scale[c] + shift[c];
  }
}

// dscale = sum g xn, dshift = sum g (both OVERWRITTEN, gemv beta 0, :271-292)
template <typename T>
__global__ void bn_param_grads_kernel(const double* __restrict__ acc, T* __restrict__ dscale, T* __restrict__ dshift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dshift) dshift[c] = static_cast<T>(acc[c]);
  if (dscale) dscale[c] = static_cast<T>(acc[C + c]);
}

// dx = (t - (xn * sum(xn t) + sum(t)) / m) / std with t = g * scale  (:296-384)
template <typename T>
__global__ void bn_backward_kernel(const T* __restrict__ g, const T* __restrict__ xn, const T* __restrict__ scale,
                                   const T* __restrict__ stdv, const double* __restrict__ acc, T* __restrict__ dx,
                                   long long total, int C, int HW, double count) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((e / HW) % C);
    const T sc = scale[c];
    const T sum_t = static_cast<T>(acc[c]) * sc, sum_xt = static_cast<T>(acc[C + c]) * sc;
    const T t = g[e] * sc;
    dx[e] = (t - (xn[e] * sum_xt + sum_t) / static_cast<T>(count)) / stdv[c];
  }
}

template <typename T, int MODE>
int bn_channel_sums(mms_context* ctx, const T* p, const T* q, double* acc, int N, int C, int HW) {
  const long long per = (long long)C * HW;
  if (per * sizeof(T) <= 48 * 1024) {
    const int threads = (int)mms_min<long long>(1024, mms_max<long long>(128, ((per + 15) / 16 + 31) / 32 * 32));
    { MmsKernelScope ks_(ctx, "bn_channel_sums_rows_kernel");
      bn_channel_sums_rows_kernel<T, MODE><<<mms_min(N, ctx->sm_count * 4), threads, per * sizeof(T), ctx->stream>>>(
          p, q, acc, N, C, HW); }
    MMS_LAUNCH_CHECK();
    return 0;
  }
  if (HW >= 256) {
    const int slices = (int)mms_max<long long>(1, mms_min<long long>(N, (8LL * ctx->sm_count + C - 1) / C));
    { MmsKernelScope ks_(ctx, "bn_channel_sums_planes_kernel");
      bn_channel_sums_planes_kernel<T, MODE><<<C * slices, 256, 0, ctx->stream>>>(p, q, acc, N, C, HW, slices); }
    MMS_LAUNCH_CHECK();
    return 0;
  }
  const long long want = mms_max<long long>(1, (4LL * ctx->sm_count + C - 1) / C);
  const int slices = (int)mms_max<long long>(1, mms_min<long long>(want, ((long long)N * HW + 1023) / 1024));
  { MmsKernelScope ks_(ctx, "bn_channel_sums_kernel");
    bn_channel_sums_kernel<T, MODE><<<C * slices, 256, 0, ctx->stream>>>(p, q, acc, N, C, HW, slices); }
  MMS_LAUNCH_CHECK();
  return 0;
}

}  // namespace

template <typename T>
int mms_bn_forward_impl(mms_context* ctx, const T* x, const T* scale, const T* shift, T* run_mean, T* run_var, T* top,
                        T* x_norm, T* batch_mean, T* batch_std, int N, int C, int HW, int train, T memory, T eps) {
  MMS_REQUIRE(x && scale && shift && run_mean && run_var && top && x_norm && batch_mean && batch_std, MMS_E_INVALID,
              "null pointer");
  MMS_REQUIRE(N > 0 && C > 0 && HW > 0, MMS_E_INVALID, "bad size");
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(double) * 2 * C, &sp));
  double* acc = static_cast<double*>(sp);
  if (train) {
    MMS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, ctx->stream));
    MMS_TRY((bn_channel_sums<T, 0>(ctx, x, nullptr, acc, N, C, HW)));
  }
  { MmsKernelScope ks_(ctx, "bn_finalize_stats_kernel");
    bn_finalize_stats_kernel<T><<<mms_ceil_div(C, 128), 128, 0, ctx->stream>>>(acc, run_mean, run_var, batch_mean, batch_std,
                                                                             C, (double)N * HW, train, memory, eps); }
  MMS_LAUNCH_CHECK();
  const long long total = (long long)N * C * HW;
  if (vec_ok(x, x_norm, top, HW, 16 / (int)sizeof(T))) {
    MmsKernelScope ks_(ctx, "bn_normalize_vec_kernel");
    if (sizeof(T) == 4)
      bn_normalize_vec_kernel<T, float4, 16 / sizeof(T)><<<ew_grid(ctx, total / 4), 256, 0, ctx->stream>>>(
          x, batch_mean, batch_std, scale, shift, x_norm, top, total / 4, C, HW);
    else
      bn_normalize_vec_kernel<T, double2, 16 / sizeof(T)><<<ew_grid(ctx, total / 2), 256, 0, ctx->stream>>>(
          x, batch_mean, batch_std, scale, shift, x_norm, top, total / 2, C, HW);
  } else if (HW >= 16) {
    { MmsKernelScope ks_(ctx, "bn_normalize_planes_kernel");
      bn_normalize_planes_kernel<T><<<ew_grid(ctx, (long long)N * C * 32), 256, 0, ctx->stream>>>(
          x, batch_mean, batch_std, scale, shift, x_norm, top, (long long)N * C, C, HW); }
  } else {
    { MmsKernelScope ks_(ctx, "bn_normalize_kernel");
      bn_normalize_kernel<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(x, batch_mean, batch_std, scale, shift, x_norm,
                                                                         top, total, C, HW); }
  }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_bn_backward_impl(mms_context* ctx, const T* dtop, const T* x_norm, const T* scale, const T* batch_std, T* dscale,
                         T* dshift, T* dx, int N, int C, int HW) {
  MMS_REQUIRE(dtop && x_norm && scale && batch_std, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N > 0 && C > 0 && HW > 0, MMS_E_INVALID, "bad size");
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(double) * 2 * C, &sp));
  double* acc = static_cast<double*>(sp);
  MMS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, ctx->stream));
  MMS_TRY((bn_channel_sums<T, 1>(ctx, dtop, x_norm, acc, N, C, HW)));
  if (dscale || dshift) {
    { MmsKernelScope ks_(ctx, "bn_param_grads_kernel");
      bn_param_grads_kernel<T><<<mms_ceil_div(C, 128), 128, 0, ctx->stream>>>(acc, dscale, dshift, C); }
    MMS_LAUNCH_CHECK();
  }
  if (dx) {
    const long long total = (long long)N * C * HW;
    if (vec_ok(dtop, x_norm, dx, HW, 16 / (int)sizeof(T))) {
      MmsKernelScope ks_(ctx, "bn_backward_vec_kernel");
      if (sizeof(T) == 4)
        bn_backward_vec_kernel<T, float4, 16 / sizeof(T)><<<ew_grid(ctx, total / 4), 256, 0, ctx->stream>>>(
            dtop, x_norm, scale, batch_std, acc, dx, total / 4, C, HW, (double)N * HW);
      else
        bn_backward_vec_kernel<T, double2, 16 / sizeof(T)><<<ew_grid(ctx, total / 2), 256, 0, ctx->stream>>>(
            dtop, x_norm, scale, batch_std, acc, dx, total / 2, C, HW, (double)N * HW);
    } else if (HW >= 16) {
      { MmsKernelScope ks_(ctx, "bn_backward_planes_kernel");
        bn_backward_planes_kernel<T><<<ew_grid(ctx, (long long)N * C * 32), 256, 0, ctx->stream>>>(
            dtop, x_norm, scale, batch_std, acc, dx, (long long)N * C, C, HW, (double)N * HW); }
    } else {
      { MmsKernelScope ks_(ctx, "bn_backward_kernel");
        bn_backward_kernel<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(dtop, x_norm, scale, batch_std, acc, dx, total,
                                                                          C, HW, (double)N * HW); }
    }
    MMS_LAUNCH_CHECK();
  }
  return 0;
}

#define INST(T)                                                                                                       \
  template int mms_sentconv_forward_impl<T>(mms_context*, const T*, const T*, const T*, T*, int, int, int, int, int); \
  template int mms_sentconv_backward_impl<T>(mms_context*, const T*, const T*, const T*, T*, T*, T*, int, int, int,   \
                                             int, int);                                                              \
  template int mms_pool_forward_impl<T>(mms_context*, const T*, T*, int*, long long, int, int, int, int, int, int,    \
                                        int, int, int, int, int);                                                    \
  template int mms_pool_backward_impl<T>(mms_context*, const T*, const int*, T*, long long, int, int, int, int, int,  \
                                         int, int, int, int, int, int);                                              \
  template int mms_tanh_forward_impl<T>(mms_context*, const T*, T*, long long);                                       \
  template int mms_tanh_backward_impl<T>(mms_context*, const T*, const T*, T*, long long);                            \
  template int mms_bn_forward_impl<T>(mms_context*, const T*, const T*, const T*, T*, T*, T*, T*, T*, T*, int, int,   \
                                      int, int, T, T);                                                               \
  template int mms_bn_backward_impl<T>(mms_context*, const T*, const T*, const T*, const T*, T*, T*, T*, int, int, int);
INST(float)
INST(double)
