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
template <int CAP>
__global__ void __launch_bounds__(kTopkThreads)
topk_update_kernel(const float* __restrict__ scores, const long long* __restrict__ idx, long long ld, int n,
                   long long idx_base, float* __restrict__ run_s, long long* __restrict__ run_i, int k) {
  __shared__ float sc[CAP];
  __shared__ long long ix[CAP];
  __shared__ int cnt;
  const int q = blockIdx.x;
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

__global__ void topk_init_kernel(float* s, long long* i, long long n) {
  for (long long t = blockIdx.x * 256LL + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
    s[t] = -CUDART_INF_F;
    i[t] = 0x7fffffffffffffffLL;
  }
}

int launch_update(mms_context* ctx, const float* scores, const long long* idx, long long ld, int n, long long idx_base,
                  float* run_s, long long* run_i, int Nq, int k) {
  MmsKernelScope ks_(ctx, "topk_update_kernel");
  if (k <= 128) topk_update_kernel<512><<<Nq, kTopkThreads, 0, ctx->stream>>>(scores, idx, ld, n, idx_base, run_s, run_i, k);
  else topk_update_kernel<2048><<<Nq, kTopkThreads, 0, ctx->stream>>>(scores, idx, ld, n, idx_base, run_s, run_i, k);
  MMS_LAUNCH_CHECK();
  return 0;
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
int mms_rerank_topk_impl(mms_context* ctx, const float* Q, const float* C, const float* W, float* QW, float* top_s,
                         long long* top_i, int Nq, long long Nc, int K1, int K2, int k, long long idx_base, int prepared) {
  MMS_REQUIRE(Q && C && W && QW && top_s && top_i, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(Nq > 0 && Nc > 0 && K1 > 0 && K2 > 0 && k > 0 && k <= 1024, MMS_E_INVALID, "bad size (k <= 1024)");
  MMS_REQUIRE(ctx->math == MMS_MATH_TF32, MMS_E_UNSUPPORTED, "top-k reranking runs on the tensor-core path");
  const long long K1p = tc_pad4(K1), K2p = tc_pad4(K2);
  // score slab: ~48 MB so that it stays in the 126 MB L2 between the GEMM that writes it and the scan that reads it
  // ... and so that the slab's GEMM is a whole number of waves of the 74 CTA pairs (256 x 256 tiles): 1000 queries
  // -> 4 row tiles x 74 column tiles = 4 waves per slab of 18 944 candidates (74 MB of scores)
  const long long m_tiles = mms_ceil_div(Nq, 256), pairs = mms_max(1, ctx->sm_count / 2);
  long long slab = (72LL << 20) / (4LL * Nq) / 256;                     // column tiles that fit ~72 MB
  const long long per_wave = mms_max<long long>(1, pairs / m_tiles);    // column tiles per wave set
  slab = mms_max<long long>(per_wave, slab / per_wave * per_wave) * 256;
  slab = mms_max<long long>(2048, mms_min<long long>(slab, Nc));
  const size_t fixed = (size_t)Nq * K1p + (size_t)K1 * K2p + (size_t)Nq * K2p + (size_t)Nq * slab;
  const bool pipelined = !prepared && ctx->concurrency != 0 && slab < Nc;
  const size_t cbuf = prepared ? 0 : (size_t)slab * K2p * (pipelined ? 2 : 1);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(float) * (fixed + cbuf), &sp));
  float* Qr = static_cast<float*>(sp);
  float* Wr = Qr + (size_t)Nq * K1p;
  float* QWr = Wr + (size_t)K1 * K2p;
  float* S = QWr + (size_t)Nq * K2p;
  float* Cr = S + (size_t)Nq * slab;
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
    MMS_TRY(launch_update(ctx, S, nullptr, slab, (int)nc, idx_base + c0, top_s, top_i, Nq, k));
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
