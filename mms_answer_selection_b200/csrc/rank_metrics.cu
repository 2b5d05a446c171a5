// Ranking metrics on the device: MAP, MRR, AUC, RankAccuracy (reference: the CPU-only evaluation layers
// src/caffe/layers/map_layer.cpp:41-100, mrr_layer.cpp:38-79, auc_layer.cpp:47-136, rank_accuracy_layer.cpp:36-50,
// which pull every score to the host, bucket them in a std::map and std::sort each bucket).
//
// MAP / MRR:  one stable radix sort of 64-bit keys (group id ascending | score descending) brings every group
//             together in rank order; two segmented scans (cub::DeviceScan::InclusiveScanByKey) give every sample its
//             rank in the group, the number of positives up to it and the running sum of precision@rank; the LAST
//             sample of each group then holds the group's average precision / reciprocal rank and adds it to two
//             double accumulators.  Everything is O(n) parallel work whatever the group sizes (30 candidates per
//             question in TREC-QA, 10^6 per query in candidate reranking).
// AUC:        one sort by score (descending), an inclusive sum of the labels, one reduction.
// Equal scores keep their input order (radix sort is stable); the reference's std::sort leaves it unspecified.
// The sort and the scans are CUB device primitives (part of the CUDA toolkit); the key construction, the
// per-sample terms and the reductions are kernels of this file.  Temporaries are carved out of the context's workspace.
#include <cub/cub.cuh>

#include "mms_common.cuh"

namespace {

__device__ __forceinline__ uint32_t score_desc_key(float s) {
  const uint32_t b = __float_as_uint(s);
  const uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);   // unsigned order == float order
  return ~asc;
}

template <typename T>
__global__ void grouped_keys_kernel(const T* __restrict__ data, long long stride, long long offset,
                                    const T* __restrict__ label, const T* __restrict__ group, long long n,
                                    unsigned long long* __restrict__ keys, int* __restrict__ labels) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t g = (uint32_t)static_cast<int>(group[i]) ^ 0x80000000u;          // map<int, ...>: signed order
    keys[i] = ((unsigned long long)g << 32) | score_desc_key(static_cast<float>(data[i * stride + offset]));
    labels[i] = static_cast<int>(label[i]);
  }
}

struct GroupOf {
  __host__ __device__ __forceinline__ uint32_t operator()(unsigned long long k) const { return (uint32_t)(k >> 32); }
};

struct Counts { int cnt, pos, zero; };                      // rank in group, positives so far, label-0 samples so far
struct CountsSum {
  __host__ __device__ __forceinline__ Counts operator()(const Counts& a, const Counts& b) const {
    return Counts{a.cnt + b.cnt, a.pos + b.pos, a.zero + b.zero};
  }
};
struct CountsOf {
  __host__ __device__ __forceinline__ Counts operator()(int label) const {
    return Counts{1, label == 1 ? 1 : 0, label == 0 ? 1 : 0};
  }
};

// precision at the rank of every positive: (positives so far) / (rank)            map_layer.cpp:78-80
__global__ void precision_terms_kernel(const int* __restrict__ labels, const Counts* __restrict__ c, long long n,
                                       double* __restrict__ term) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    term[i] = labels[i] == 1 ? (double)c[i].pos / (double)c[i].cnt : 0.0;
}

// acc[0] += AP of every effective group, acc[1] += RR; eff[0], eff[1] count them.  The last sample of a group holds the
// group totals; the first positive of a group (pos == 1 at a label-1 sample) holds its reciprocal rank.
__global__ void group_totals_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ labels,
                                    const Counts* __restrict__ c, const double* __restrict__ apsum, long long n,
                                    double* __restrict__ acc, int* __restrict__ eff, double* __restrict__ first_rr) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool tail = (i == n - 1) || ((keys[i] >> 32) != (keys[i + 1] >> 32));
    if (!tail) continue;
    const Counts t = c[i];
    if (t.pos >= 1 && t.cnt - t.pos > 0) {                  // a positive and something that is not one   :87-89
      atomicAdd(acc, apsum[i] / (double)t.pos);
      atomicAdd(eff, 1);
    }
    if (t.pos >= 1 && t.zero > 0) {                          // a positive and a label-0 sample           mrr :70-72
      // rank of the group's first positive = samples before it + 1: found by the thread that owns it (below)
      atomicAdd(acc + 1, first_rr[i - (t.cnt - 1)]);         // stored at the group's head by first_positive_kernel
      atomicAdd(eff + 1, 1);
    }
  }
}

// first_rr[head of the group] = 1 / rank of the group's first positive
__global__ void first_positive_kernel(const int* __restrict__ labels, const Counts* __restrict__ c, long long n,
                                      double* __restrict__ first_rr) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (labels[i] == 1 && c[i].pos == 1) first_rr[i - (c[i].cnt - 1)] = 1.0 / (double)c[i].cnt;
}

template <typename T>
__global__ void finalize_map_mrr_kernel(const double* acc, const int* eff, T* map_out, T* mrr_out) {
  if (map_out) *map_out = static_cast<T>(acc[0] / (double)eff[0]);      // 0/0 when no group qualifies, as the reference
  if (mrr_out) *mrr_out = static_cast<T>(acc[1] / (double)eff[1]);
}

// ---- AUC
template <typename T>
__global__ void auc_keys_kernel(const T* __restrict__ data, long long stride, long long offset,
                                const T* __restrict__ label, long long n, int has_ignore, int ignore_label,
                                uint32_t* __restrict__ keys, int* __restrict__ labels) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int lv = static_cast<int>(label[i]);
    const bool skip = has_ignore && lv == ignore_label;      // auc_layer.cpp:69-71
    keys[i] = skip ? 0xFFFFFFFFu : score_desc_key(static_cast<float>(data[i * stride + offset]));
    labels[i] = skip ? 0x40000000 : lv;                      // marker: contributes nothing (filtered below)
  }
}
struct AucLabel {                                            // ignored samples count as 0 in the running sum
  __host__ __device__ __forceinline__ int operator()(int l) const { return l == 0x40000000 ? 0 : l; }
};
// acc[0] += high_i * (1 - label_i) over the counted samples; cnt[0] = counted samples, cnt[1] = total of the labels
__global__ void auc_terms_kernel(const int* __restrict__ labels, const int* __restrict__ high, long long n,
                                 double* __restrict__ acc, unsigned long long* __restrict__ cnt) {
  double s = 0;
  unsigned long long c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = labels[i];
    if (l == 0x40000000) continue;
    s += (double)high[i] * (double)(1 - l);
    ++c;
  }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(acc, s); atomicAdd(cnt, c); }
}
template <typename T>
__global__ void finalize_auc_kernel(const double* acc, const unsigned long long* cnt, const int* high, long long n, T* out) {
  const double h = n > 0 ? (double)high[n - 1] : 0.0;       // ignored samples sort last and add 0
  *out = h > 0 ? static_cast<T>(acc[0] / h / ((double)cnt[0] - h)) : T(0);   // :126-133
}

template <typename T>
__global__ void rank_accuracy_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ label,
                                     long long n, unsigned long long* __restrict__ hits) {
  unsigned long long c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += (label[i] * (a[i] - b[i])) > T(0) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(hits, c);
}
template <typename T>
__global__ void finalize_accuracy_kernel(const unsigned long long* hits, long long n, T* out) {
  *out = static_cast<T>(static_cast<T>(*hits) / static_cast<T>(n));
}

inline int grid_for(mms_context* ctx, long long n) {
  return (int)mms_max<long long>(1, mms_min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8));
}

// scratch: one carve-up of the context's growable workspace (allocated once, kept between calls: a stream-ordered
// allocation of the 2.3 GB a 32 M-score call needs costs more than the sort itself when the pool hands it back)
struct Arena {
  mms_context* ctx; char* base = nullptr; size_t used = 0, cap = 0;
  explicit Arena(mms_context* c) : ctx(c) {}
  int reserve(size_t bytes) {
    cap = bytes;
    void* p = nullptr;
    MMS_TRY(mms_scratch(ctx, bytes, &p));
    base = static_cast<char*>(p);
    return 0;
  }
  template <typename U> U* take(size_t count) {
    used = (used + 255) & ~(size_t)255;
    U* p = reinterpret_cast<U*>(base + used);
    used += count * sizeof(U);
    return p;
  }
};
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

template <typename T>
int mms_rank_map_mrr_impl(mms_context* ctx, const T* data, long long stride, long long offset, const T* label,
                          const T* group, long long n, T* map_out, T* mrr_out) {
  MMS_REQUIRE(n >= 0 && n < 0x7fffffffLL && stride >= 1 && offset >= 0 && offset < stride, MMS_E_INVALID, "bad size");
  MMS_REQUIRE((map_out || mrr_out) && (n == 0 || (data && label && group)), MMS_E_INVALID, "null pointer");
  cudaStream_t st = ctx->stream;
  const int N = (int)n;
  typedef unsigned long long u64;
  cub::TransformInputIterator<uint32_t, GroupOf, const u64*> gkeys(static_cast<const u64*>(nullptr), GroupOf());
  cub::TransformInputIterator<Counts, CountsOf, const int*> cin(static_cast<const int*>(nullptr), CountsOf());
  size_t t_sort = 0, t_scan1 = 0, t_scan2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (const u64*)nullptr, (u64*)nullptr, (const int*)nullptr, (int*)nullptr, N,
                                  0, 64, st);
  cub::DeviceScan::InclusiveScanByKey(nullptr, t_scan1, gkeys, cin, (Counts*)nullptr, CountsSum(), N, cub::Equality(), st);
  cub::DeviceScan::InclusiveSumByKey(nullptr, t_scan2, gkeys, (const double*)nullptr, (double*)nullptr, N, cub::Equality(), st);
  const size_t t_tmp = mms_max(t_sort, mms_max(t_scan1, t_scan2));
  const size_t cnt = (size_t)mms_max<long long>(n, 1);
  Arena ar(ctx);
  MMS_TRY(ar.reserve(pad256(t_tmp) + 2 * pad256(cnt * 8) + 2 * pad256(cnt * 4) + pad256(cnt * sizeof(Counts)) +
                     3 * pad256(cnt * 8) + 4096));
  void* tmp = ar.take<char>(t_tmp);
  u64* k_in = ar.take<u64>(cnt); u64* k_out = ar.take<u64>(cnt);
  int* l_in = ar.take<int>(cnt); int* l_out = ar.take<int>(cnt);
  Counts* c = ar.take<Counts>(cnt);
  double* term = ar.take<double>(cnt); double* apsum = ar.take<double>(cnt); double* first_rr = ar.take<double>(cnt);
  double* acc = ar.take<double>(2); int* eff = ar.take<int>(2);
  MMS_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  MMS_CUDA(cudaMemsetAsync(eff, 0, 2 * sizeof(int), st));
  if (n > 0) {
    const int grid = grid_for(ctx, n);
    { MmsKernelScope ks_(ctx, "grouped_keys_kernel");
      grouped_keys_kernel<T><<<grid, 256, 0, st>>>(data, stride, offset, label, group, n, k_in, l_in); }
    MMS_LAUNCH_CHECK();
    size_t tb = t_tmp;
    MMS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const u64*)k_in, k_out, (const int*)l_in, l_out, N, 0, 64, st));
    cub::TransformInputIterator<uint32_t, GroupOf, const u64*> gk(k_out, GroupOf());
    cub::TransformInputIterator<Counts, CountsOf, const int*> ci(l_out, CountsOf());
    tb = t_tmp;
    MMS_CUDA(cub::DeviceScan::InclusiveScanByKey(tmp, tb, gk, ci, c, CountsSum(), N, cub::Equality(), st));
    { MmsKernelScope ks_(ctx, "precision_terms_kernel");
      precision_terms_kernel<<<grid, 256, 0, st>>>(l_out, c, n, term); }
    MMS_LAUNCH_CHECK();
    tb = t_tmp;
    MMS_CUDA(cub::DeviceScan::InclusiveSumByKey(tmp, tb, gk, (const double*)term, apsum, N, cub::Equality(), st));
    { MmsKernelScope ks_(ctx, "first_positive_kernel");
      first_positive_kernel<<<grid, 256, 0, st>>>(l_out, c, n, first_rr); }
    MMS_LAUNCH_CHECK();
    { MmsKernelScope ks_(ctx, "group_totals_kernel");
      group_totals_kernel<<<grid, 256, 0, st>>>(k_out, l_out, c, apsum, n, acc, eff, first_rr); }
    MMS_LAUNCH_CHECK();
  }
  { MmsKernelScope ks_(ctx, "finalize_map_mrr_kernel");
    finalize_map_mrr_kernel<T><<<1, 1, 0, st>>>(acc, eff, map_out, mrr_out); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_rank_auc_impl(mms_context* ctx, const T* data, long long stride, long long offset, const T* label,
                      long long n, int has_ignore, int ignore_label, T* out) {
  MMS_REQUIRE(n >= 0 && n < 0x7fffffffLL && stride >= 1 && offset >= 0 && offset < stride, MMS_E_INVALID, "bad size");
  MMS_REQUIRE(out && (n == 0 || (data && label)), MMS_E_INVALID, "null pointer");
  cudaStream_t st = ctx->stream;
  const int N = (int)n;
  size_t t_sort = 0, t_scan = 0;
  cub::TransformInputIterator<int, AucLabel, const int*> lin(static_cast<const int*>(nullptr), AucLabel());
  cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int*)nullptr,
                                  (int*)nullptr, N, 0, 32, st);
  cub::DeviceScan::InclusiveSum(nullptr, t_scan, lin, (int*)nullptr, N, st);
  const size_t t_tmp = mms_max(t_sort, t_scan);
  const size_t cnt = (size_t)mms_max<long long>(n, 1);
  Arena ar(ctx);
  MMS_TRY(ar.reserve(pad256(t_tmp) + 5 * pad256(cnt * 4) + 4096));
  void* tmp = ar.take<char>(t_tmp);
  uint32_t* k_in = ar.take<uint32_t>(cnt); uint32_t* k_out = ar.take<uint32_t>(cnt);
  int* l_in = ar.take<int>(cnt); int* l_out = ar.take<int>(cnt); int* high = ar.take<int>(cnt);
  double* acc = ar.take<double>(1); unsigned long long* counted = ar.take<unsigned long long>(1);
  MMS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  MMS_CUDA(cudaMemsetAsync(counted, 0, sizeof(unsigned long long), st));
  if (n > 0) {
    const int grid = grid_for(ctx, n);
    { MmsKernelScope ks_(ctx, "auc_keys_kernel");
      auc_keys_kernel<T><<<grid, 256, 0, st>>>(data, stride, offset, label, n, has_ignore, ignore_label, k_in, l_in); }
    MMS_LAUNCH_CHECK();
    size_t tb = t_tmp;
    MMS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const uint32_t*)k_in, k_out, (const int*)l_in, l_out, N, 0, 32, st));
    cub::TransformInputIterator<int, AucLabel, const int*> li(l_out, AucLabel());
    tb = t_tmp;
    MMS_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, li, high, N, st));
    { MmsKernelScope ks_(ctx, "auc_terms_kernel");
      auc_terms_kernel<<<grid, 256, 0, st>>>(l_out, high, n, acc, counted); }
    MMS_LAUNCH_CHECK();
  }
  { MmsKernelScope ks_(ctx, "finalize_auc_kernel");
    finalize_auc_kernel<T><<<1, 1, 0, st>>>(acc, counted, high, n, out); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_rank_accuracy_impl(mms_context* ctx, const T* a, const T* b, const T* label, long long n, T* out) {
  MMS_REQUIRE(n > 0 && a && b && label && out, MMS_E_INVALID, "bad argument");
  cudaStream_t st = ctx->stream;
  Arena ar(ctx);
  MMS_TRY(ar.reserve(256));
  unsigned long long* hits = ar.take<unsigned long long>(1);
  MMS_CUDA(cudaMemsetAsync(hits, 0, sizeof(unsigned long long), st));
  { MmsKernelScope ks_(ctx, "rank_accuracy_kernel");
    rank_accuracy_kernel<T><<<grid_for(ctx, n), 256, 0, st>>>(a, b, label, n, hits); }
  MMS_LAUNCH_CHECK();
  { MmsKernelScope ks_(ctx, "finalize_accuracy_kernel");
    finalize_accuracy_kernel<T><<<1, 1, 0, st>>>(hits, n, out); }
  MMS_LAUNCH_CHECK();
  return 0;
}

#define INST(T)                                                                                                  \
  template int mms_rank_map_mrr_impl<T>(mms_context*, const T*, long long, long long, const T*, const T*, long long, T*, T*); \
  template int mms_rank_auc_impl<T>(mms_context*, const T*, long long, long long, const T*, long long, int, int, T*);         \
  template int mms_rank_accuracy_impl<T>(mms_context*, const T*, const T*, const T*, long long, T*);
INST(float)
INST(double)
