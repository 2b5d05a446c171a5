// Per-query top-k over candidate scores: the consumer of the reranking scores in the reference is a ranking by score
// per query (examples/trec_qa_w2v_mms/do_trec_qa_clean.py:617-650 writes trec_eval run files from the per-query score
// lists; map_layer.cpp:41-100 / mrr_layer.cpp:38-79 sort each query's candidates by score).  SURVEY.md 8(e): candidates
// are sharded over the GPUs, every GPU keeps a per-query local top-k, the lists are all-gathered and merged, ties broken
// by candidate index so that the ranking does not depend on the sharding.
//
// Order: score descending, then candidate index ascending (deterministic; NaN scores never enter a list).
//
// topk_update_kernel: one CTA per query.  The running list (k entries, sorted) sits at the head of a shared-memory array;
// the row of new scores is scanned against the list's current threshold (its k-th entry), survivors are appended, and
// when the array is about to overflow -- or at the end of the row -- the array is bitonic-sorted and cut back to k.
// After the first slab almost nothing survives the threshold, so a slab row costs one read of its scores (which are
// still in L2: mms_rerank_topk sizes the score slab to stay there) and at most one small sort.
#include <math_constants.h>
#include <stdint.h>

#include "mms_common.cuh"
#include "tc/tc_gemm.cuh"

namespace {

constexpr int kTopkThreads = 256;

struct Entry { float s; long long i; };

__device__ __forceinline__ bool better(float s, long long i, float ts, long long ti) {
  return s > ts || (s == ts && i < ti);
}

// sorts (sc, ix)[0..CAP) by (score desc, index asc); CAP a power of two
template <int CAP>
__device__ void bitonic_sort(float* sc, long long* ix) {
  for (int size = 2; size <= CAP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < CAP / 2; t += kTopkThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;             // "descending" halves keep the better entry first
        const float a = sc[lo], b = sc[hi];
        const long long ia = ix[lo], ib = ix[hi];
        const bool swap = desc ? better(b, ib, a, ia) : better(a, ia, b, ib);
        if (swap) { sc[lo] = b; sc[hi] = a; ix[lo] = ib; ix[hi] = ia; }
      }
    }
  }
  __syncthreads();
}

// row: `n` scores of query blockIdx.x at scores + q * ld; their candidate indices are idx[q * ld + c] when idx != null,
// else idx_base + c.  run_s / run_i: the query's running list (k entries, sorted; unused slots = (-inf, LLONG_MAX)).
// cand_n (optional): per-query count written by topk_filter_kernel.  mode 1: the row holds that many filtered candidates
// (skip the query when the count is 0 or exceeded the buffer `cand_cap`); mode 2: the row is the full score slab, scanned
// only for the queries whose candidate buffer overflowed; the count is cleared here for the next slab.
template <int CAP>
__global__ void __launch_bounds__(kTopkThreads)
topk_update_kernel(const float* __restrict__ scores, const long long* __restrict__ idx, long long ld, int n,
                   long long idx_base, float* __restrict__ run_s, long long* __restrict__ run_i, int k,
                   int* __restrict__ cand_n, int cand_cap, int mode) {
  __shared__ float sc[CAP];
  __shared__ long long ix[CAP];
  __shared__ int cnt;
  const int q = blockIdx.x;
  if (cand_n) {
    const int have = cand_n[q];
    if (mode == 1) {
      if (have == 0 || have > cand_cap) return;
      n = have;
    } else if (mode == 3) {                              // candidates of the threshold epilogue: merge, clear, flag overflow
      __syncthreads();
      if (threadIdx.x == 0) {
        cand_n[q] = 0;
        if (have > cand_cap) atomicExch(cand_n + gridDim.x, 1);    // the word behind the counts: "some buffer overflowed"
      }
      if (have == 0 || have > cand_cap) return;
      n = have;
    } else {
      __syncthreads();                                   // every thread has read the count before it is cleared
      if (threadIdx.x == 0) cand_n[q] = 0;
      if (have <= cand_cap) return;
    }
  }
  const float* row = scores + (size_t)q * ld;
  const long long* irow = idx ? idx + (size_t)q * ld : nullptr;
  for (int t = threadIdx.x; t < CAP; t += kTopkThreads) {
    sc[t] = t < k ? run_s[(size_t)q * k + t] : -CUDART_INF_F;
    ix[t] = t < k ? run_i[(size_t)q * k + t] : 0x7fffffffffffffffLL;
  }
  if (threadIdx.x == 0) cnt = k;
  __syncthreads();
  float ts = sc[k - 1];
  long long ti = ix[k - 1];
  __shared__ int overflow[2];                            // alternating: chunk i resets / reads flag i & 1
  // cut: sort everything appended so far and keep the k best (called by all threads)
  auto cut = [&](int have) {
    for (int t = have + threadIdx.x; t < CAP; t += kTopkThreads) { sc[t] = -CUDART_INF_F; ix[t] = 0x7fffffffffffffffLL; }
    bitonic_sort<CAP>(sc, ix);
    if (threadIdx.x == 0) cnt = k;
    __syncthreads();
    ts = sc[k - 1]; ti = ix[k - 1];
  };
  constexpr int U = 8;                                   // scores per thread and chunk: U loads in flight (L2 latency)
  int chunk = 0;
  for (int c0 = 0; c0 < n; c0 += U * kTopkThreads, ++chunk) {
    int* ovf = &overflow[chunk & 1];
    float v[U]; long long gi[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int c = c0 + j * kTopkThreads + threadIdx.x;
      v[j] = c < n ? __ldcg(row + c) : CUDART_NAN_F;     // read through L2: the GEMM of this slab has just written it
      gi[j] = c < n ? (irow ? irow[c] : idx_base + c) : -1;
    }
    const int have0 = cnt;                               // (all threads passed a barrier since cnt last changed)
    if (threadIdx.x == 0) *ovf = 0;
    __syncthreads();
    int mine = 0;
#pragma unroll
    for (int j = 0; j < U; ++j) mine += (v[j] == v[j] && gi[j] >= 0 && better(v[j], gi[j], ts, ti)) ? 1 : 0;
    if (mine) {
      int pos = atomicAdd(&cnt, mine);
      if (pos + mine <= CAP) {
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (v[j] == v[j] && gi[j] >= 0 && better(v[j], gi[j], ts, ti)) { sc[pos] = v[j]; ix[pos] = gi[j]; ++pos; }
      } else {
        *ovf = 1;
      }
    }
    __syncthreads();
    if (*ovf) {
      // too many survivors for the array (an unfilled list, or scores that keep rising): discard this chunk's appends,
      // cut, and feed the chunk 256 scores at a time -- at most 256 appends between cuts always fit (CAP - k >= 256)
      __syncthreads();
      if (threadIdx.x == 0) cnt = have0;
      __syncthreads();
      if (have0 > k) cut(have0);
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (v[j] == v[j] && gi[j] >= 0 && better(v[j], gi[j], ts, ti)) {
          const int pos = atomicAdd(&cnt, 1);
          sc[pos] = v[j]; ix[pos] = gi[j];
        }
        __syncthreads();
        const int have = cnt;
        if (have + kTopkThreads > CAP) cut(have);
        else __syncthreads();
      }
    }
  }
  __syncthreads();
  if (cnt > k) cut(cnt);
  for (int t = threadIdx.x; t < k; t += kTopkThreads) {
    run_s[(size_t)q * k + t] = sc[t];
    run_i[(size_t)q * k + t] = ix[t];
  }
}

// Memory-speed pre-filter of one score slab: scores that beat the query's current k-th entry go to the query's candidate
// buffer (global atomics: after the first slab a handful per query), everything else is only read.  A query whose list is
// not full yet, or whose buffer overflows, is handled by the full scan (mode 2 above).
__global__ void __launch_bounds__(256)
topk_filter_kernel(const float* __restrict__ scores, long long ld, int n, long long idx_base, const float* __restrict__ run_s,
                   const long long* __restrict__ run_i, int k, float* __restrict__ cand_s, long long* __restrict__ cand_i,
                   int* __restrict__ cand_n, int cand_cap, int chunks) {
  const int q = blockIdx.x / chunks, ch = blockIdx.x - q * chunks;
  const float ts = run_s[(size_t)q * k + k - 1];
  const long long ti = run_i[(size_t)q * k + k - 1];
  const float* row = scores + (size_t)q * ld;
  const int per = (n + chunks - 1) / chunks, c_beg = ch * per, c_end = min(n, c_beg + per);
  for (int c0 = c_beg + threadIdx.x; c0 < c_end; c0 += 256 * 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j * 256 < c_end) ? __ldcg(row + c0 + j * 256) : CUDART_NAN_F;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long gi = idx_base + c0 + j * 256;
      if (v[j] == v[j] && better(v[j], gi, ts, ti)) {
        const int pos = atomicAdd(cand_n + q, 1);
        if (pos < cand_cap) { cand_s[(size_t)q * cand_cap + pos] = v[j]; cand_i[(size_t)q * cand_cap + pos] = gi; }
      }
    }
  }
}

__global__ void topk_init_kernel(float* s, long long* i, long long n) {
  for (long long t = blockIdx.x * 256LL + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
    s[t] = -CUDART_INF_F;
    i[t] = 0x7fffffffffffffffLL;
  }
}

int launch_update(mms_context* ctx, const float* scores, const long long* idx, long long ld, int n, long long idx_base,
                  float* run_s, long long* run_i, int Nq, int k, int* cand_n = nullptr, int cand_cap = 0, int mode = 0) {
  MmsKernelScope ks_(ctx, "topk_update_kernel");
  if (k <= 128)
    topk_update_kernel<512><<<Nq, kTopkThreads, 0, ctx->stream>>>(scores, idx, ld, n, idx_base, run_s, run_i, k, cand_n, cand_cap, mode);
  else
    topk_update_kernel<2048><<<Nq, kTopkThreads, 0, ctx->stream>>>(scores, idx, ld, n, idx_base, run_s, run_i, k, cand_n, cand_cap, mode);
  MMS_LAUNCH_CHECK();
  return 0;
}

constexpr int kCandCap = 1024;          // filtered candidates per query and slab before the full scan takes over

// one slab of scores into the running lists: filter -> merge of the filtered candidates -> full scan where that overflowed
int fold_slab(mms_context* ctx, const float* S, long long ld, int nc, long long idx_base, float* top_s, long long* top_i,
              int Nq, int k, float* cand_s, long long* cand_i, int* cand_n) {
  const int chunks = mms_max(1, mms_min(64, mms_ceil_div(nc, 4096)));
  { MmsKernelScope ks_(ctx, "topk_filter_kernel");
    topk_filter_kernel<<<Nq * chunks, 256, 0, ctx->stream>>>(S, ld, nc, idx_base, top_s, top_i, k, cand_s, cand_i, cand_n,
                                                             kCandCap, chunks); }
  MMS_LAUNCH_CHECK();
  MMS_TRY(launch_update(ctx, cand_s, cand_i, kCandCap, kCandCap, 0, top_s, top_i, Nq, k, cand_n, kCandCap, 1));
  return launch_update(ctx, S, nullptr, ld, nc, idx_base, top_s, top_i, Nq, k, cand_n, kCandCap, 2);
}

}  // namespace

int mms_topk_init(mms_context* ctx, float* run_s, long long* run_i, int Nq, int k) {
  MmsKernelScope ks_(ctx, "topk_init_kernel");
  topk_init_kernel<<<mms_ceil_div((long long)Nq * k, 256), 256, 0, ctx->stream>>>(run_s, run_i, (long long)Nq * k);
  MMS_LAUNCH_CHECK();
  return 0;
}

int mms_topk_update(mms_context* ctx, const float* scores, const long long* idx, long long ld, long long n,
                    long long idx_base, float* run_s, long long* run_i, int Nq, int k) {
  MMS_REQUIRE(scores && run_s && run_i, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(Nq > 0 && k > 0 && k <= 1024 && n >= 0 && n <= 0x7fffffffLL, MMS_E_INVALID, "bad size (k <= 1024)");
  if (n == 0) return 0;
  return launch_update(ctx, scores, idx, ld, (int)n, idx_base, run_s, run_i, Nq, k);
}

// scores = (Q W) C^T slab by slab, each slab folded into the per-query top-k while it is still in L2; the full score
// matrix is never written.
namespace {
// Form 1 (always correct, capturable): score slabs are written (sized to stay in L2) and folded into the lists.
int rerank_topk_stored(mms_context* ctx, const float* Q, const float* C, const float* W, float* QW, float* top_s,
                       long long* top_i, int Nq, long long Nc, int K1, int K2, int k, long long idx_base, int prepared) {
  const long long K1p = tc_pad4(K1), K2p = tc_pad4(K2);
  // score slab: ~48 MB so that it stays in the 126 MB L2 between the GEMM that writes it and the scan that reads it
  // ... and so that the slab's GEMM is a whole number of waves of the 74 CTA pairs (256 x 256 tiles): 1000 queries
  // -> 4 row tiles x 74 column tiles = 4 waves per slab of 18 944 candidates (74 MB of scores)
  const long long m_tiles = mms_ceil_div(Nq, 256), pairs = mms_max(1, ctx->sm_count / 2);
  long long slab = (72LL << 20) / (4LL * Nq) / 256;                     // column tiles that fit ~72 MB
  const long long per_wave = mms_max<long long>(1, pairs / m_tiles);    // column tiles per wave set
  slab = mms_max<long long>(per_wave, slab / per_wave * per_wave) * 256;
  slab = mms_max<long long>(2048, mms_min<long long>(slab, Nc));
  // + per-query candidate buffers of the pre-filter: scores (4 B), indices (8 B), counts
  const size_t cand_floats = ((size_t)Nq * kCandCap * 3 + (size_t)Nq + 16 + 3) & ~(size_t)3;
  const size_t fixed = (size_t)Nq * K1p + (size_t)K1 * K2p + (size_t)Nq * K2p + (size_t)Nq * slab + cand_floats;
  const bool pipelined = !prepared && ctx->concurrency != 0 && slab < Nc;
  const size_t cbuf = prepared ? 0 : (size_t)slab * K2p * (pipelined ? 2 : 1);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(float) * (fixed + cbuf), &sp));
  float* Qr = static_cast<float*>(sp);
  float* Wr = Qr + (size_t)Nq * K1p;
  float* QWr = Wr + (size_t)K1 * K2p;
  float* S = QWr + (size_t)Nq * K2p;
  long long* cand_i = reinterpret_cast<long long*>(((reinterpret_cast<uintptr_t>(S + (size_t)Nq * slab) + 15) & ~(uintptr_t)15));
  float* cand_s = reinterpret_cast<float*>(cand_i + (size_t)Nq * kCandCap);
  int* cand_n = reinterpret_cast<int*>(cand_s + (size_t)Nq * kCandCap);
  float* Cr = S + (size_t)Nq * slab + cand_floats;
  MMS_CUDA(cudaMemsetAsync(cand_n, 0, sizeof(int) * (size_t)Nq, ctx->stream));
  const RoundJob j0[2] = {{Q, Qr, Nq, K1, K1, K1p, nullptr}, {W, Wr, K1, K2, K2, K2p, nullptr}};
  MMS_TRY(mms_tf32_round(ctx, j0, 2));
  TcGemmArgs t = tc_gemm_args(Qr, K1p, 0, Wr, K2p, 1, QW, K2, Nq, K2, K1);
  t.operands_tf32 = 1;
  MMS_TRY(mms_tc_gemm(ctx, t));
  const RoundJob j1[1] = {{QW, QWr, Nq, K2, K2, K2p, nullptr}};
  MMS_TRY(mms_tf32_round(ctx, j1, 1));
  MMS_TRY(mms_topk_init(ctx, top_s, top_i, Nq, k));
  if (!prepared) {
    const RoundJob j2[1] = {{C, Cr, mms_min<long long>(slab, Nc), K2, K2, K2p, nullptr}};
    MMS_TRY(mms_tf32_round(ctx, j2, 1));
  }
  int i = 0;
  for (long long c0 = 0; c0 < Nc; c0 += slab, ++i) {
    const long long nc = mms_min<long long>(slab, Nc - c0), next0 = c0 + slab;
    const float* cur = prepared ? C + (size_t)c0 * K2p : Cr + (pipelined ? (size_t)(i & 1) * slab * K2p : 0);
    const bool more = next0 < Nc;
    if (!prepared && more && pipelined) {                 // round slab i+1 beside the GEMM of slab i
      float* nxt = Cr + (size_t)((i + 1) & 1) * slab * K2p;
      MMS_TRY(mms_fork(ctx, 0));
      MmsStreamSwitch sw_(ctx, 0);
      const RoundJob j2[1] = {{C + (size_t)next0 * K2, nxt, mms_min<long long>(slab, Nc - next0), K2, K2, K2p, nullptr}};
      MMS_TRY(mms_tf32_round(ctx, j2, 1));
    }
    TcGemmArgs g = tc_gemm_args(QWr, K2p, 0, cur, K2p, 0, S, slab, Nq, (int)nc, K2);   // both K-major
    g.operands_tf32 = 1;
    MMS_TRY(mms_tc_gemm(ctx, g));
    MMS_TRY(fold_slab(ctx, S, slab, (int)nc, idx_base + c0, top_s, top_i, Nq, k, cand_s, cand_i, cand_n));
    if (!prepared && more) {
      if (pipelined) MMS_TRY(mms_join(ctx, 0));
      else {
        const RoundJob j2[1] = {{C + (size_t)next0 * K2, Cr, mms_min<long long>(slab, Nc - next0), K2, K2, K2p, nullptr}};
        MMS_TRY(mms_tf32_round(ctx, j2, 1));
      }
    }
  }
  return 0;
}
}  // namespace

// Form 2 (default): the scores never leave the tensor memory.  A short prefix of the candidates is scored the stored way
// to fill the lists; every later slab runs the GEMM with the THRESHOLD EPILOGUE (tc_gemm.cuh flt_*): each accumulator
// value is compared with its query's current k-th entry and only survivors are appended to per-query buffers, which one
// small merge per slab folds into the lists.  Slabs grow geometrically (expected survivors = slab * k / seen stays a
// quarter of the buffer).  Scores that keep rising can still overflow a buffer: that is detected (one flag, one host
// read at the end) and the call is redone the stored way, so the result is always exact.
int mms_rerank_topk_impl(mms_context* ctx, const float* Q, const float* C, const float* W, float* QW, float* top_s,
                         long long* top_i, int Nq, long long Nc, int K1, int K2, int k, long long idx_base, int prepared) {
  MMS_REQUIRE(Q && C && W && QW && top_s && top_i, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(Nq > 0 && Nc > 0 && K1 > 0 && K2 > 0 && k > 0 && k <= 1024, MMS_E_INVALID, "bad size (k <= 1024)");
  MMS_REQUIRE(ctx->math == MMS_MATH_TF32, MMS_E_UNSUPPORTED, "top-k reranking runs on the tensor-core path");
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool capturing = cudaStreamIsCapturing(ctx->stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone;
  const long long seed = mms_min<long long>(Nc, 8192);
  if (capturing || Nc <= 4 * seed)
    return rerank_topk_stored(ctx, Q, C, W, QW, top_s, top_i, Nq, Nc, K1, K2, k, idx_base, prepared);
  const long long K1p = tc_pad4(K1), K2p = tc_pad4(K2);
  int cand_cap = 4096;
  while (cand_cap > 1024 && (size_t)Nq * cand_cap * 12 > ((size_t)256 << 20)) cand_cap >>= 1;
  const long long grow = mms_max<long long>(1, cand_cap / (4LL * k));
  const size_t cand_floats = ((size_t)Nq * cand_cap * 3 + (size_t)Nq + 32 + 3) & ~(size_t)3;
  const size_t fixed = (size_t)Nq * K1p + (size_t)K1 * K2p + (size_t)Nq * K2p + (size_t)Nq * seed + cand_floats;
  long long slab_max = ((long long)(ctx->scratch_cap / sizeof(float)) - (long long)fixed) / K2p / (prepared ? 1 : 2);
  slab_max = mms_max<long long>(seed, mms_min<long long>(slab_max, 1 << 18)) / 256 * 256;
  const size_t cbuf = prepared ? 0 : (size_t)slab_max * K2p * 2;
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(float) * (fixed + cbuf), &sp));
  float* Qr = static_cast<float*>(sp);
  float* Wr = Qr + (size_t)Nq * K1p;
  float* QWr = Wr + (size_t)K1 * K2p;
  float* S = QWr + (size_t)Nq * K2p;
  long long* cand_i = reinterpret_cast<long long*>(((reinterpret_cast<uintptr_t>(S + (size_t)Nq * seed) + 15) & ~(uintptr_t)15));
  float* cand_s = reinterpret_cast<float*>(cand_i + (size_t)Nq * cand_cap);
  int* cand_n = reinterpret_cast<int*>(cand_s + (size_t)Nq * cand_cap);       // Nq counts + the overflow word
  float* Cr = S + (size_t)Nq * seed + cand_floats;
  MMS_CUDA(cudaMemsetAsync(cand_n, 0, sizeof(int) * ((size_t)Nq + 1), ctx->stream));
  const RoundJob j0[2] = {{Q, Qr, Nq, K1, K1, K1p, nullptr}, {W, Wr, K1, K2, K2, K2p, nullptr}};
  MMS_TRY(mms_tf32_round(ctx, j0, 2));
  TcGemmArgs t = tc_gemm_args(Qr, K1p, 0, Wr, K2p, 1, QW, K2, Nq, K2, K1);
  t.operands_tf32 = 1;
  MMS_TRY(mms_tc_gemm(ctx, t));
  const RoundJob j1[1] = {{QW, QWr, Nq, K2, K2, K2p, nullptr}};
  MMS_TRY(mms_tf32_round(ctx, j1, 1));
  MMS_TRY(mms_topk_init(ctx, top_s, top_i, Nq, k));
  const bool pipelined = !prepared && ctx->concurrency != 0;
  auto rounded = [&](long long c0, long long nc, float* dst) -> int {
    const RoundJob j2[1] = {{C + (size_t)c0 * K2, dst, nc, K2, K2, K2p, nullptr}};
    return mms_tf32_round(ctx, j2, 1);
  };
  if (!prepared) MMS_TRY(rounded(0, seed, Cr));
  int i = 0;
  long long c0 = 0, seen = 0;
  while (c0 < Nc) {
    const long long nc = seen == 0 ? seed : mms_min<long long>(mms_min<long long>(slab_max, grow * seen / 256 * 256), Nc - c0);
    const long long next0 = c0 + nc;
    const long long nnext = next0 < Nc ? mms_min<long long>(mms_min<long long>(slab_max, grow * (seen + nc) / 256 * 256), Nc - next0) : 0;
    const float* cur = prepared ? C + (size_t)c0 * K2p : Cr + (size_t)(i & 1) * slab_max * K2p;
    if (!prepared && nnext > 0 && pipelined) {            // round the next slab beside this slab's GEMM
      MMS_TRY(mms_fork(ctx, 0));
      MmsStreamSwitch sw_(ctx, 0);
      MMS_TRY(rounded(next0, nnext, Cr + (size_t)((i + 1) & 1) * slab_max * K2p));
    }
    TcGemmArgs g = tc_gemm_args(QWr, K2p, 0, cur, K2p, 0, S, seed, Nq, (int)nc, K2);   // both K-major; C only written by the seed slab
    g.operands_tf32 = 1;
    if (seen == 0) {
      MMS_TRY(mms_tc_gemm(ctx, g));
      MMS_TRY(launch_update(ctx, S, nullptr, seed, (int)nc, idx_base + c0, top_s, top_i, Nq, k));
    } else {
      g.flt_s = top_s; g.flt_i = top_i; g.flt_k = k; g.flt_base = idx_base + c0;
      g.cand_s = cand_s; g.cand_i = cand_i; g.cand_n = cand_n; g.cand_cap = cand_cap;
      const int rc = mms_tc_gemm_tma(ctx, g);
      if (rc == MMS_E_UNSUPPORTED) return rerank_topk_stored(ctx, Q, C, W, QW, top_s, top_i, Nq, Nc, K1, K2, k, idx_base, prepared);
      MMS_TRY(rc);
      MMS_TRY(launch_update(ctx, cand_s, cand_i, cand_cap, cand_cap, 0, top_s, top_i, Nq, k, cand_n, cand_cap, 3));
    }
    if (!prepared && nnext > 0) {
      if (pipelined) MMS_TRY(mms_join(ctx, 0));
      else MMS_TRY(rounded(next0, nnext, Cr + (size_t)((i + 1) & 1) * slab_max * K2p));
    }
    seen += nc; c0 = next0; ++i;
  }
  int overflowed = 0;
  MMS_CUDA(cudaMemcpyAsync(&overflowed, cand_n + Nq, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MMS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (overflowed)        // a candidate buffer overflowed (scores that keep rising): exact answer the stored way
    return rerank_topk_stored(ctx, Q, C, W, QW, top_s, top_i, Nq, Nc, K1, K2, k, idx_base, prepared);
  return 0;
}
