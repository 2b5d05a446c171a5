// Embed backward for the two Embed layers of a QA net (they share one table: do_trec_qa_clean.py:461-468, param sharing by
// name) as ONE scatter-add with the token rows grouped by id:  dW[id, :] += sum of the gradient rows of every token with
// that id, dbias += sum of all rows.  Reference op: src/caffe/layers/embed_layer.cu:29-39 (one atomicAdd per float).
//
// Why: every atomic add on a table row that is not in L2 -- after 1.5 GB of other traffic none is -- makes L2 fetch the
// row from HBM and write it back later: the per-layer kernels of embed.cu move the touched rows of dW once per OCCURRENCE
// run (tools/probes/embed_bwd_probe.cu: 96 us of red.global.add.v4.f32 against a cold table; the same adds against an
// L2-resident one take 25 us, tools/probes/bulk_reduce_probe.cu, and the TMA engine's cp.reduce.async.bulk is no faster).
// Grouped by id, both layers' occurrences of a row cost one round trip, the padding id costs a few hundred atomics
// instead of thousands, and what remains is reading the 393 MB of dq / da once.
//
//   plan    (needs the ids only -- MMSNet runs it on a side stream while the forward contractions run)
//     1  count[id]++            warp-aggregated (match.any): a warp of padding tokens issues one atomic
//     2  scan of the V bins: start[id]; the list of SHORT runs (<= 128 rows); LONG runs cut into chunks of 256 rows
//     3  rows[cursor[id]++] = row   warp-aggregated the same way
//   reduce  (after SimCross backward)
//     4  one warp per short run: rows summed in registers, one red.global.add.v4.f32 per 16 bytes of the table row
//     5  one CTA per long chunk: eight warps sum 32 rows each, shared-memory reduction, one set of atomics per chunk
//   dbias rides along: per-warp column sums -> shared memory -> one set of atomics per CTA.
//
// Float only (the double nets use the per-layer kernels); D % 4 == 0, D <= 512, 16-byte aligned blobs, at least 32 768
// token rows (smaller batches are launch-bound: two launches beat six) -- anything else runs the per-layer kernels of
// embed.cu, as does a handle with MMS_OPT_EMBED_DETERMINISTIC (embed_det.cu).
#include <cub/cub.cuh>

#include "mms_common.cuh"

namespace {

constexpr int kLongRun = 128;      // runs up to this many rows are summed by one warp
constexpr int kChunkRows = 256;    // rows of a long run per CTA (kWarps x 32: a lane holds one row number of its warp)
constexpr int kWarps = 8;
constexpr int kBatch = 8;          // short runs a warp describes at once
constexpr long long kMinRows = 32768;   // below this the step is launch-bound: two per-layer launches beat plan + reduce

inline bool grouped_shape(const mms_context* ctx, long long Mt, int D) {
  return !ctx->embed_deterministic && Mt >= kMinRows && D % 4 == 0 && D <= 512;
}

struct SortedPlan {
  const void* idx0 = nullptr; const void* idx1 = nullptr;
  long long M0 = 0, M1 = 0;
  int V = 0;
  bool valid = false;
  int* buf = nullptr;              // one allocation: count | start | cursor | run_id | chunks | rows | counters
  size_t ints = 0;
  int *count = nullptr, *start = nullptr, *cursor = nullptr, *run_id = nullptr, *chunks = nullptr, *rows = nullptr,
      *counters = nullptr,         // counters[0] = short runs, [1] = long chunks
      *totals = nullptr;           // per scan CTA: rows, short runs, long chunks
  int max_chunks = 0;
};

__device__ __forceinline__ int load_id(const float* __restrict__ idx0, const float* __restrict__ idx1, long long M0,
                                       long long r) {
  return static_cast<int>(r < M0 ? idx0[r] : idx1[r - M0]);
}

// Both passes over the ids share this shape: every lane of a warp takes one row; lanes holding the same id elect a leader.
__global__ void __launch_bounds__(256)
plan_count_kernel(const float* __restrict__ idx0, const float* __restrict__ idx1, long long M0, long long Mt, int V,
                  int* __restrict__ count, int* fault) {
  const int lane = threadIdx.x & 31;
  for (long long base = (blockIdx.x * 256LL + threadIdx.x) - lane; base < Mt; base += gridDim.x * 256LL) {
    const long long r = base + lane;
    int id = -1;
    if (r < Mt) {
      id = load_id(idx0, idx1, M0, r);
      if (id < 0 || id >= V) { id = -1; atomicExch(fault, 1); }
    }
    const unsigned same = __match_any_sync(0xffffffffu, id);
    if (id >= 0 && lane == __ffs(same) - 1) atomicAdd(count + id, __popc(same));
  }
}

__global__ void __launch_bounds__(256)
plan_place_kernel(const float* __restrict__ idx0, const float* __restrict__ idx1, long long M0, long long Mt, int V,
                  int* __restrict__ cursor, int* __restrict__ rows) {
  const int lane = threadIdx.x & 31;
  for (long long base = (blockIdx.x * 256LL + threadIdx.x) - lane; base < Mt; base += gridDim.x * 256LL) {
    const long long r = base + lane;
    int id = -1;
    if (r < Mt) {
      id = load_id(idx0, idx1, M0, r);
      if (id < 0 || id >= V) id = -1;
    }
    const unsigned same = __match_any_sync(0xffffffffu, id);
    const int leader = __ffs(same) - 1;
    int pos = 0;
    if (id >= 0 && lane == leader) pos = atomicAdd(cursor + id, __popc(same));
    pos = __shfl_sync(0xffffffffu, pos, leader);
    if (id >= 0) rows[pos + __popc(same & ((1u << lane) - 1u))] = (int)r;
  }
}

// Exclusive scan of count[] -> start[] / cursor[], and compaction of the non-empty bins into the short-run list and the
// long-chunk list (id, first row slot, rows).  Two launches of V / 1024 CTAs: per-CTA totals, then every CTA adds up the
// totals in front of it and scans its own 1024 bins (four consecutive bins per thread, one 16-byte load).
constexpr int kScanThreads = 256, kScanBins = 4 * kScanThreads;

__device__ __forceinline__ void bin_sums(int n, int& a, int& b, int& c) {
  a += n;
  if (n > kLongRun) c += (n + kChunkRows - 1) / kChunkRows; else if (n > 0) ++b;
}

__device__ __forceinline__ int4 load_counts(const int* __restrict__ count, int v0, int V) {     // count[] is padded by 8
  int4 n = *reinterpret_cast<const int4*>(count + v0);
  if (v0 + 1 >= V) n.y = 0;
  if (v0 + 2 >= V) n.z = 0;
  if (v0 + 3 >= V) n.w = 0;
  if (v0 >= V) n.x = 0;
  return n;
}

__global__ void __launch_bounds__(kScanThreads)
plan_totals_kernel(const int* __restrict__ count, int* __restrict__ totals, int V) {
  typedef cub::BlockReduce<int, kScanThreads> Reduce;
  __shared__ typename Reduce::TempStorage tmp[3];
  const int4 n = load_counts(count, blockIdx.x * kScanBins + 4 * threadIdx.x, V);
  int a = 0, b = 0, c = 0;
  bin_sums(n.x, a, b, c); bin_sums(n.y, a, b, c); bin_sums(n.z, a, b, c); bin_sums(n.w, a, b, c);
  a = Reduce(tmp[0]).Sum(a); b = Reduce(tmp[1]).Sum(b); c = Reduce(tmp[2]).Sum(c);
  if (threadIdx.x == 0) { totals[3 * blockIdx.x] = a; totals[3 * blockIdx.x + 1] = b; totals[3 * blockIdx.x + 2] = c; }
}

__global__ void __launch_bounds__(kScanThreads)
plan_scan_kernel(const int* __restrict__ count, const int* __restrict__ totals, int* __restrict__ start,
                 int* __restrict__ cursor, int* __restrict__ run_id, int* __restrict__ chunks, int* __restrict__ counters,
                 int V) {
  typedef cub::BlockScan<int, kScanThreads> Scan;
  typedef cub::BlockReduce<int, kScanThreads> Reduce;
  __shared__ typename Scan::TempStorage tmp[3];
  __shared__ typename Reduce::TempStorage rtmp[3];
  __shared__ int s_base[3];
  // totals of the CTAs in front of this one (the last CTA also publishes the grand totals)
  int pa = 0, pb = 0, pc = 0;
  for (int j = threadIdx.x; j < (int)blockIdx.x; j += kScanThreads) { pa += totals[3 * j]; pb += totals[3 * j + 1]; pc += totals[3 * j + 2]; }
  pa = Reduce(rtmp[0]).Sum(pa); pb = Reduce(rtmp[1]).Sum(pb); pc = Reduce(rtmp[2]).Sum(pc);
  if (threadIdx.x == 0) { s_base[0] = pa; s_base[1] = pb; s_base[2] = pc; }
  __syncthreads();
  const int v0 = blockIdx.x * kScanBins + 4 * threadIdx.x;
  const int4 n4 = load_counts(count, v0, V);
  const int n[4] = {n4.x, n4.y, n4.z, n4.w};
  int a = 0, b = 0, c = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) bin_sums(n[i], a, b, c);
  int ea, eb, ec;
  Scan(tmp[0]).ExclusiveSum(a, ea); Scan(tmp[1]).ExclusiveSum(b, eb); Scan(tmp[2]).ExclusiveSum(c, ec);
  a = ea + s_base[0]; b = eb + s_base[1]; c = ec + s_base[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = v0 + i;
    if (v >= V) break;
    start[v] = a; cursor[v] = a;
    if (n[i] > kLongRun) {
      for (int o = 0; o < n[i]; o += kChunkRows, ++c) {
        chunks[3 * c] = v; chunks[3 * c + 1] = a + o; chunks[3 * c + 2] = min(kChunkRows, n[i] - o);
      }
    } else if (n[i] > 0) {
      run_id[b++] = v;
    }
    a += n[i];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) { counters[0] = b; counters[1] = c; }
}

__device__ __forceinline__ float4 vadd4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct ReduceArgs {
  const float* dtop0; const float* dtop1; long long M0;
  const int* count; const int* start; const int* run_id; const int* chunks; const int* rows; const int* counters;
  float* dW; float* dbias; int D;
};

__device__ __forceinline__ const float4* row_ptr(const ReduceArgs& a, int r) {
  return reinterpret_cast<const float4*>(r < a.M0 ? a.dtop0 + (size_t)r * a.D : a.dtop1 + (size_t)(r - a.M0) * a.D);
}

template <int VPL>
__device__ __forceinline__ void load_row(const ReduceArgs& a, int r, bool ok, int lane, int nvec, float4 (&v)[VPL]) {
  const float4* s = row_ptr(a, r);
#pragma unroll
  for (int k = 0; k < VPL; ++k)
    v[k] = (ok && lane + 32 * k < nvec) ? __ldcs(s + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Short runs, kBatch at a time per warp: lane t fetches the description of run t (id, first row slot, length) and its first
// two row numbers, so the chain  run list -> start / count -> row numbers  is walked once per batch instead of once per
// run (a batch of 32 left too few warps with work: 54 k runs are 1 700 batches, 11 per SM); then the warp sums run after
// run, each one exposed HBM latency long -- hidden by 32 resident warps per SM.
template <int VPL>
__global__ void __launch_bounds__(kWarps * 32, VPL <= 3 ? 4 : 2)
short_runs_kernel(const ReduceArgs a) {
  extern __shared__ float4 s_col[];                    // kWarps x nvec column sums (dbias): lane-owned slots, no atomics
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nvec = a.D >> 2;
  const int nruns = a.counters[0];
  float4* my_col = s_col + wid * nvec;
#pragma unroll
  for (int k = 0; k < VPL; ++k)
    if (lane + 32 * k < nvec) my_col[lane + 32 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = (blockIdx.x * kWarps + wid) * kBatch; base < nruns; base += gridDim.x * kWarps * kBatch) {
    int id = 0, s = 0, n = 0, r0 = 0, r1 = 0;
    if (lane < kBatch && base + lane < nruns) {
      id = a.run_id[base + lane];
      s = a.start[id]; n = a.count[id];
      r0 = a.rows[s];
      r1 = n > 1 ? a.rows[s + 1] : r0;
    }
    const int cnt = min(kBatch, nruns - base);
    for (int t = 0; t < cnt; ++t) {
      const int id_t = __shfl_sync(0xffffffffu, id, t), n_t = __shfl_sync(0xffffffffu, n, t), s_t = __shfl_sync(0xffffffffu, s, t);
      float4 va[VPL], vb[VPL];
      load_row<VPL>(a, __shfl_sync(0xffffffffu, r0, t), true, lane, nvec, va);
      load_row<VPL>(a, __shfl_sync(0xffffffffu, r1, t), n_t > 1, lane, nvec, vb);
      float4 acc[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc[k] = vadd4(va[k], vb[k]);
      for (int i = 2; i < n_t; i += 2) {                // longer runs: two more rows per trip
        const bool two = i + 1 < n_t;
        load_row<VPL>(a, a.rows[s_t + i], true, lane, nvec, va);
        load_row<VPL>(a, a.rows[s_t + (two ? i + 1 : i)], two, lane, nvec, vb);
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc[k] = vadd4(acc[k], vadd4(va[k], vb[k]));
      }
      if (a.dW) {        // (a plain read-modify-write -- this warp is the row's only writer -- measured no faster: the
                         // table row comes from HBM either way, tools/probes/bulk_reduce_probe.cu)
        float4* d = reinterpret_cast<float4*>(a.dW + (size_t)id_t * a.D);
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (lane + 32 * k < nvec) atomicAdd(d + lane + 32 * k, acc[k]);
      }
      if (a.dbias) {
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (lane + 32 * k < nvec) my_col[lane + 32 * k] = vadd4(my_col[lane + 32 * k], acc[k]);
      }
    }
  }
  if (a.dbias) {
    __syncthreads();
    for (int c = threadIdx.x; c < nvec; c += kWarps * 32) {
      float4 t = s_col[c];
      for (int w = 1; w < kWarps; ++w) t = vadd4(t, s_col[w * nvec + c]);
      atomicAdd(reinterpret_cast<float4*>(a.dbias) + c, t);
    }
  }
}

// Long runs, one CTA per chunk of 256 rows: warp w takes rows w, w + 8, ...; lane j holds the row number of the warp's
// j-th row (one coalesced load for all 32), eight gradient rows in flight per warp.
template <int VPL>
__global__ void __launch_bounds__(kWarps * 32, 2)
long_chunks_kernel(const ReduceArgs a) {
  extern __shared__ float4 s_col[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nvec = a.D >> 2;
  const int nchunks = a.counters[1];
  for (int j = blockIdx.x; j < nchunks; j += gridDim.x) {
    const int id = a.chunks[3 * j], s = a.chunks[3 * j + 1], n = a.chunks[3 * j + 2];
    const int mine = (n - wid + kWarps - 1) / kWarps;             // rows of this warp (<= 32)
    const int my_row = lane < mine ? a.rows[s + wid + kWarps * lane] : 0;
    float4 acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < mine; i += 8) {
      float4 v[8][VPL];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        load_row<VPL>(a, __shfl_sync(0xffffffffu, my_row, (i + u) & 31), i + u < mine, lane, nvec, v[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc[k] = vadd4(acc[k], v[u][k]);
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (lane + 32 * k < nvec) s_col[wid * nvec + lane + 32 * k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < nvec; c += kWarps * 32) {
      float4 t = s_col[c];
      for (int w = 1; w < kWarps; ++w) t = vadd4(t, s_col[w * nvec + c]);
      if (a.dW) atomicAdd(reinterpret_cast<float4*>(a.dW + (size_t)id * a.D) + c, t);
      if (a.dbias) atomicAdd(reinterpret_cast<float4*>(a.dbias) + c, t);
    }
    __syncthreads();
  }
}

SortedPlan* plan_of(mms_context* ctx) {
  if (!ctx->embed_plan) ctx->embed_plan = new SortedPlan();
  return static_cast<SortedPlan*>(ctx->embed_plan);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int build_plan(mms_context* ctx, const float* idx0, long long M0, const float* idx1, long long M1, int V) {
  SortedPlan* p = plan_of(ctx);
  p->valid = false;
  const long long Mt = M0 + M1;
  MMS_REQUIRE(Mt < 0x7fffffffLL, MMS_E_UNSUPPORTED, "too many rows");
  const int max_chunks = (int)(Mt / kLongRun + 1);     // a long run has > 128 rows: ceil(n / 256) <= n / 128
  const int scan_ctas = mms_ceil_div(V, kScanBins);
  const size_t need = 3 * (size_t)(V + 8) + (size_t)(V + 8) + 3 * (size_t)max_chunks + (size_t)Mt + 8 + 3 * (size_t)scan_ctas + 8;
  if (need > p->ints) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    MMS_REQUIRE(cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone,
                MMS_E_INVALID, "the id plan must have its size before a capture: run the step once eagerly");
    if (p->buf) { MMS_CUDA(cudaDeviceSynchronize()); MMS_CUDA(cudaFree(p->buf)); p->buf = nullptr; p->ints = 0; }
    MMS_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->buf), sizeof(int) * need));
    p->ints = need;
  }
  int* w = p->buf;
  p->count = w; w += V + 8;
  p->start = w; w += V + 8;
  p->cursor = w; w += V + 8;
  p->run_id = w; w += V + 8;
  p->chunks = w; w += 3 * (size_t)max_chunks;
  p->counters = w; w += 8;
  p->totals = w; w += 3 * (size_t)scan_ctas + 5;
  p->rows = w;
  p->max_chunks = max_chunks;
  MMS_CUDA(cudaMemsetAsync(p->count, 0, sizeof(int) * (size_t)(V + 8), ctx->stream));
  const int grid = (int)mms_min<long long>((Mt + 255) / 256, (long long)ctx->sm_count * 4);
  { MmsKernelScope ks_(ctx, "embed_plan_count");
    plan_count_kernel<<<grid, 256, 0, ctx->stream>>>(idx0, idx1, M0, Mt, V, p->count, ctx->fault_flag); }
  MMS_LAUNCH_CHECK();
  const int gs = mms_ceil_div(V, kScanBins);
  { MmsKernelScope ks_(ctx, "embed_plan_totals");
    plan_totals_kernel<<<gs, kScanThreads, 0, ctx->stream>>>(p->count, p->totals, V); }
  MMS_LAUNCH_CHECK();
  { MmsKernelScope ks_(ctx, "embed_plan_scan");
    plan_scan_kernel<<<gs, kScanThreads, 0, ctx->stream>>>(p->count, p->totals, p->start, p->cursor, p->run_id, p->chunks,
                                                           p->counters, V); }
  MMS_LAUNCH_CHECK();
  { MmsKernelScope ks_(ctx, "embed_plan_place");
    plan_place_kernel<<<grid, 256, 0, ctx->stream>>>(idx0, idx1, M0, Mt, V, p->cursor, p->rows); }
  MMS_LAUNCH_CHECK();
  p->idx0 = idx0; p->idx1 = idx1; p->M0 = M0; p->M1 = M1; p->V = V;
  p->valid = true;
  return 0;
}

template <int VPL>
int launch_reduce(mms_context* ctx, const ReduceArgs& a, long long Mt) {
  const size_t smem = sizeof(float4) * (size_t)kWarps * (a.D >> 2);
  // short runs: the machine four times over (a warp's run is a few KB: latency is hidden by the number of warps)
  const int g_short = ctx->sm_count * 4;
  const int g_long = (int)mms_min<long long>(Mt / kLongRun + 1, (long long)ctx->sm_count * 4);
  { MmsKernelScope ks_(ctx, "embed_backward_long_chunks");
    MMS_CARVEOUT(long_chunks_kernel<VPL>);
    long_chunks_kernel<VPL><<<g_long, kWarps * 32, smem, ctx->stream>>>(a); }
  MMS_LAUNCH_CHECK();
  { MmsKernelScope ks_(ctx, "embed_backward_short_runs");
    MMS_CARVEOUT(short_runs_kernel<VPL>);
    short_runs_kernel<VPL><<<g_short, kWarps * 32, smem, ctx->stream>>>(a); }
  MMS_LAUNCH_CHECK();
  return 0;
}

}  // namespace

void mms_embed_plan_destroy(mms_context* ctx) {
  SortedPlan* p = static_cast<SortedPlan*>(ctx->embed_plan);
  if (!p) return;
  if (p->buf) cudaFree(p->buf);
  delete p;
  ctx->embed_plan = nullptr;
}

int mms_embed_plan_pair_impl(mms_context* ctx, const float* idx0, long long M0, const float* idx1, long long M1, int V) {
  MMS_REQUIRE(M0 >= 0 && M1 >= 0 && V > 0, MMS_E_INVALID, "bad size");
  MMS_REQUIRE((idx0 || M0 == 0) && (idx1 || M1 == 0), MMS_E_INVALID, "null pointer");
  if (M0 + M1 == 0) return 0;
  if (ctx->embed_deterministic || M0 + M1 < kMinRows) return 0;      // the backward will take the per-layer kernels
  return build_plan(ctx, idx0, M0, idx1, M1, V);
}

int mms_embed_backward_pair_impl(mms_context* ctx, const float* idx0, const float* dtop0, long long M0, const float* idx1,
                                 const float* dtop1, long long M1, float* dW, float* dbias, int D, int V) {
  MMS_REQUIRE(M0 >= 0 && M1 >= 0 && D > 0 && V > 0, MMS_E_INVALID, "bad size");
  if (M0 + M1 == 0 || (!dW && !dbias)) return 0;
  MMS_REQUIRE((M0 == 0 || (idx0 && dtop0)) && (M1 == 0 || (idx1 && dtop1)), MMS_E_INVALID, "null pointer");
  const bool grouped = grouped_shape(ctx, M0 + M1, D) && (M0 == 0 || aligned16(dtop0)) &&
                       (M1 == 0 || aligned16(dtop1)) && (!dW || aligned16(dW)) && (!dbias || aligned16(dbias));
  if (!grouped) {      // the per-layer kernels, one blob after the other (the deterministic form needs that order)
    if (M0) MMS_TRY(mms_embed_backward_impl<float>(ctx, idx0, dtop0, dW, dbias, M0, D, V));
    if (M1) MMS_TRY(mms_embed_backward_impl<float>(ctx, idx1, dtop1, dW, dbias, M1, D, V));
    return 0;
  }
  SortedPlan* p = plan_of(ctx);
  // a plan made earlier for these very blobs (mms_embed_plan_pair: the caller vouches that the ids have not changed since)
  if (!(p->valid && p->idx0 == idx0 && p->idx1 == idx1 && p->M0 == M0 && p->M1 == M1 && p->V == V))
    MMS_TRY(build_plan(ctx, idx0, M0, idx1, M1, V));
  p->valid = false;                                       // one plan, one backward
  ReduceArgs a;
  a.dtop0 = dtop0; a.dtop1 = dtop1; a.M0 = M0;
  a.count = p->count; a.start = p->start; a.run_id = p->run_id; a.chunks = p->chunks; a.rows = p->rows; a.counters = p->counters;
  a.dW = dW; a.dbias = dbias; a.D = D;
  const int vpl = ((D >> 2) + 31) / 32;
  switch (vpl) {
    case 1: return launch_reduce<1>(ctx, a, M0 + M1);
    case 2: return launch_reduce<2>(ctx, a, M0 + M1);
    case 3: return launch_reduce<3>(ctx, a, M0 + M1);
    default: return launch_reduce<4>(ctx, a, M0 + M1);
  }
}
