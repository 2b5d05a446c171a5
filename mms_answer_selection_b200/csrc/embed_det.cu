// Embed backward, deterministic form (MMS_OPT_EMBED_DETERMINISTIC): dW[idx[n], :] += dtop[n, :], dbias += sum_n dtop[n, :]
// with results that do not depend on the order in which rows are visited -- bit-identical from run to run AND under
// any permutation of the (id, gradient row) pairs.  Reference op: src/caffe/layers/embed_layer.cu:29-39 (one atomicAdd
// per float, i.e. an arrival-order float sum); SURVEY.md 7 "hard parts" asks for a reproducible variant.
//
// Float addition is not associative, integer addition is.  Every table row that receives gradient ("run": the token rows
// sharing one id) is summed in 64-bit FIXED POINT with a per-run power-of-two scale chosen from the run's largest
// magnitude and its length, so that no sum can overflow and the quantum is max|x| * len * 2^-61 -- far below a float ulp
// of any partial sum -- and converted back once: dW[id] += float(sum).  One writer per table row, no float atomics.
//
//   1  keys = (int) idx, vals = row number;  radix sort by key (cub)              -> rows of one id are adjacent
//   2  run heads, run ids (scan), run starts (select)                             -> R runs
//   3  max|dtop[row, :]| per sorted row -> atomicMax into its run's slot (order-free) and into a global slot (for dbias)
//   4  runs are cut into chunks of 128 rows (a centre-padded batch puts ~half of all rows on the pad id); offsets by scan
//   5  one WARP per chunk: 64-bit accumulators in registers over the chunk's rows; single-chunk runs write dW; chunks of
//      longer runs park their integer partials and the last one to finish (ticket) adds them up -- any order gives the
//      same integer -- and writes.  dbias rides along with the global scale: per-CTA shared-memory integer accumulators,
//      one 64-bit atomicAdd per column and CTA.
//   6  dbias += float(bias sum)
//
// Two passes over dtop (3 and 5) instead of one: the price of order independence (measured beside the atomic path
// in bench.py extra.embed_backward_modes).
#include <cub/cub.cuh>

#include "mms_common.cuh"

namespace {

constexpr int kChunkRows = 128;
constexpr int kWarps = 8;                 // warps per CTA of the reduction kernel
constexpr int kMaxVecPerLane = 8;         // 16-byte column groups a lane may own: D <= 8 * 32 * (16 / sizeof(T))

typedef unsigned long long u64;

template <typename T>
__global__ void __launch_bounds__(256)
det_keys_kernel(const T* __restrict__ idx, int* __restrict__ keys, int* __restrict__ vals, long long M, int V, int* fault) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < M; i += (long long)gridDim.x * 256) {
    const int id = static_cast<int>(idx[i]);
    const bool ok = id >= 0 && id < V;
    if (!ok) atomicExch(fault, 1);
    keys[i] = ok ? id : V;                 // out-of-range ids sort behind every table row and are skipped
    vals[i] = (int)i;
  }
}

__global__ void __launch_bounds__(256) det_heads_kernel(const int* __restrict__ skeys, int* __restrict__ head, long long M) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < M; i += (long long)gridDim.x * 256)
    head[i] = (i == 0 || skeys[i] != skeys[i - 1]) ? 1 : 0;
}

// one warp per sorted row: largest magnitude of the gradient row -> its run's slot and the global slot
template <typename T>
__global__ void __launch_bounds__(256)
det_rowmax_kernel(const T* __restrict__ dtop, const int* __restrict__ svals, const int* __restrict__ rid,
                  unsigned* __restrict__ runmax, unsigned* __restrict__ gmax, long long M, int D) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * 256LL + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * 256) >> 5;
  float cta_max = 0.f;
  for (long long i = warp; i < M; i += nwarps) {
    const T* row = dtop + (size_t)svals[i] * D;
    float m = 0.f;
    for (int c = lane; c < D; c += 32) m = fmaxf(m, fabsf((float)row[c]));      // |x| rounded to float: monotone, enough for a bound
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // a double rounded to float may round DOWN: bump by one ulp so that the bound stays a bound
    if (sizeof(T) == 8 && m > 0.f) m = __uint_as_float(__float_as_uint(m) + 1u);
    if (lane == 0 && m > 0.f) atomicMax(runmax + (rid[i] - 1), __float_as_uint(m));   // non-negative floats order like their bits
    cta_max = fmaxf(cta_max, m);
  }
  if (lane == 0 && cta_max > 0.f) atomicMax(gmax, __float_as_uint(cta_max));
}

__global__ void __launch_bounds__(256)
det_nchunks_kernel(const int* __restrict__ run_start, const int* __restrict__ num_runs, int* __restrict__ nch,
                   int* __restrict__ nch_multi, long long M) {
  const int R = *num_runs;
  for (long long r = blockIdx.x * 256LL + threadIdx.x; r <= M; r += (long long)gridDim.x * 256) {
    int n = 0;
    if (r < R) {
      const long long end = (r + 1 < R) ? run_start[r + 1] : M;
      n = (int)((end - run_start[r] + kChunkRows - 1) / kChunkRows);
    }
    nch[r] = n;
    nch_multi[r] = n > 1 ? n : 0;
  }
}

// power-of-two scale 2^e such that |x| <= maxabs and `len` terms cannot overflow 61 bits
__device__ __forceinline__ int scale_exp(float maxabs, long long len) {
  int lg = 0;
  while ((1LL << lg) < len) ++lg;
  return 61 - lg - (ilogbf(maxabs) + 1);
}
__device__ __forceinline__ long long to_fixed(double x, double scale) { return __double2ll_rn(x * scale); }   // scale = 2^e: exact
__device__ __forceinline__ double from_fixed(long long v, int e) { return scalbn((double)v, -e); }

template <typename T> struct Ld16;
template <> struct Ld16<float> { typedef float4 type; static constexpr int n = 4; };
template <> struct Ld16<double> { typedef double2 type; static constexpr int n = 2; };

struct DetArgs {
  const int* skeys; const int* svals; const int* run_start; const int* num_runs;
  const int* chunk_off; const int* multi_off;       // exclusive scans over runs (M + 1 entries; [M] = totals)
  const unsigned* runmax; const unsigned* gmax;
  unsigned* ticket; long long* partials; u64* bias_acc;
  long long M; int D, V;
};

template <typename T, int VPL>      // VPL: 16-byte column groups per lane
__global__ void __launch_bounds__(kWarps * 32)
det_reduce_kernel(const T* __restrict__ dtop, T* __restrict__ dW, const DetArgs a, int want_bias) {
  typedef typename Ld16<T>::type VT;
  constexpr int VN = Ld16<T>::n;
  extern __shared__ u64 s_bias[];                      // D integer accumulators of this CTA (dbias, global scale)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nvec = a.D / VN;
  for (int c = threadIdx.x; c < a.D; c += blockDim.x) s_bias[c] = 0;
  __syncthreads();
  const int R = *a.num_runs;
  const int total = a.chunk_off[a.M];
  const float gm = __uint_as_float(*a.gmax);
  const int eg = gm > 0.f ? scale_exp(gm, a.M) : 0;
  for (int j = blockIdx.x * kWarps + wid; j < total; j += gridDim.x * kWarps) {
    int lo = 0, hi = R - 1;                            // the run this chunk belongs to: last r with chunk_off[r] <= j
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (a.chunk_off[mid] <= j) lo = mid; else hi = mid - 1;
    }
    const int r = lo;
    const long long rs = a.run_start[r], re = (r + 1 < R) ? a.run_start[r + 1] : a.M;
    const int nch = (int)((re - rs + kChunkRows - 1) / kChunkRows), cj = j - a.chunk_off[r];
    const long long i0 = rs + (long long)cj * kChunkRows, i1 = mms_min<long long>(i0 + kChunkRows, re);
    const int id = a.skeys[rs];
    const float rm = __uint_as_float(a.runmax[r]);
    const int er = rm > 0.f ? scale_exp(rm, re - rs) : 0;
    const double sr = scalbn(1.0, er), sg = scalbn(1.0, eg);      // the two power-of-two scales, once per chunk
    long long acc[VPL][VN], bacc[VPL][VN];
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int c = 0; c < VN; ++c) { acc[v][c] = 0; bacc[v][c] = 0; }
    for (long long i = i0; i < i1; i += 2) {             // two rows in flight per lane and column group
      VT x0[VPL], x1[VPL];
      const VT* p0 = reinterpret_cast<const VT*>(dtop + (size_t)a.svals[i] * a.D);
      const bool two = i + 1 < i1;
      const VT* p1 = reinterpret_cast<const VT*>(dtop + (size_t)a.svals[two ? i + 1 : i] * a.D);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int cv = lane + 32 * v;
        if (cv < nvec) { x0[v] = __ldcs(p0 + cv); x1[v] = __ldcs(p1 + cv); }
      }
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        if (lane + 32 * v >= nvec) continue;
        const T* e0 = reinterpret_cast<const T*>(&x0[v]);
        const T* e1 = reinterpret_cast<const T*>(&x1[v]);
#pragma unroll
        for (int c = 0; c < VN; ++c) {
          acc[v][c] += to_fixed((double)e0[c], sr);
          if (want_bias) bacc[v][c] += to_fixed((double)e0[c], sg);
          if (two) {
            acc[v][c] += to_fixed((double)e1[c], sr);
            if (want_bias) bacc[v][c] += to_fixed((double)e1[c], sg);
          }
        }
      }
    }
    if (want_bias) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int cv = lane + 32 * v;
        if (cv >= nvec) continue;
#pragma unroll
        for (int c = 0; c < VN; ++c) atomicAdd(&s_bias[cv * VN + c], (u64)bacc[v][c]);
      }
    }
    if (id >= a.V || dW == nullptr) continue;           // out-of-range ids were flagged; nothing to write
    bool writer = nch == 1;
    if (nch > 1) {                                       // park the integer partial; the last chunk of the run adds them up
      long long* part = a.partials + ((size_t)a.multi_off[r] + cj) * a.D;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int cv = lane + 32 * v;
        if (cv >= nvec) continue;
#pragma unroll
        for (int c = 0; c < VN; ++c) part[cv * VN + c] = acc[v][c];
      }
      __threadfence();
      __syncwarp();
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(a.ticket + r, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      writer = t == (unsigned)(nch - 1);
      if (writer) {
        __threadfence();
        const long long* base = a.partials + (size_t)a.multi_off[r] * a.D;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int c = 0; c < VN; ++c) acc[v][c] = 0;
        for (int k = 0; k < nch; ++k) {
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int cv = lane + 32 * v;
            if (cv >= nvec) continue;
#pragma unroll
            for (int c = 0; c < VN; ++c) acc[v][c] += __ldcg(base + (size_t)k * a.D + cv * VN + c);
          }
        }
      }
    }
    if (writer) {
      VT* dst = reinterpret_cast<VT*>(dW + (size_t)id * a.D);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int cv = lane + 32 * v;
        if (cv >= nvec) continue;
        VT w = dst[cv];
        T* we = reinterpret_cast<T*>(&w);
#pragma unroll
        for (int c = 0; c < VN; ++c) we[c] += (T)from_fixed(acc[v][c], er);
        dst[cv] = w;
      }
    }
  }
  if (want_bias) {
    __syncthreads();
    for (int c = threadIdx.x; c < a.D; c += blockDim.x)
      if (s_bias[c]) atomicAdd(a.bias_acc + c, s_bias[c]);
  }
}

template <typename T>
__global__ void det_bias_finish_kernel(const u64* __restrict__ bias_acc, const unsigned* __restrict__ gmax, T* __restrict__ dbias,
                                       long long M, int D) {
  const float gm = __uint_as_float(*gmax);
  if (!(gm > 0.f)) return;
  const int eg = scale_exp(gm, M);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x)
    dbias[c] += (T)from_fixed((long long)bias_acc[c], eg);
}

size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

// Returns MMS_E_UNSUPPORTED for shapes the vector kernel does not take (the caller then uses the atomic path only if
// determinism was not requested; see mms_embed_backward_impl).
template <typename T>
int mms_embed_backward_deterministic(mms_context* ctx, const T* idx, const T* dtop, T* dW, T* dbias, long long M, int D,
                                     int V) {
  constexpr int VN = 16 / (int)sizeof(T);
  MMS_REQUIRE(M <= 0x7fffffffLL - 1, MMS_E_UNSUPPORTED, "too many rows");
  MMS_REQUIRE(D % VN == 0 && D / VN <= 32 * kMaxVecPerLane, MMS_E_UNSUPPORTED,
              "deterministic Embed backward needs D a multiple of 16 bytes and <= 4096 bytes per row");
  MMS_REQUIRE((reinterpret_cast<uintptr_t>(dtop) & 15) == 0 && (!dW || (reinterpret_cast<uintptr_t>(dW) & 15) == 0),
              MMS_E_UNSUPPORTED, "deterministic Embed backward needs 16-byte aligned blobs");
  cudaStream_t st = ctx->stream;
  const int n = (int)M;
  // ---- workspace carve-up
  size_t t_sort = 0, t_scan = 0, t_sel = 0;
  int bits = 1;
  while ((1LL << bits) <= V) ++bits;                       // keys are 0..V
  cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, n, 0, bits, st);
  cub::DeviceScan::InclusiveSum(nullptr, t_scan, (const int*)nullptr, (int*)nullptr, n + 1, st);
  cub::DeviceSelect::Flagged(nullptr, t_sel, cub::CountingInputIterator<int>(0), (const int*)nullptr, (int*)nullptr,
                             (int*)nullptr, n, st);
  const size_t t_cub = al256(mms_max(t_sort, mms_max(t_scan, t_sel)));
  const size_t n_part = (size_t)(2 * (M / kChunkRows) + 2);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += al256(bytes); return o; };
  const size_t o_keys = take(4 * M), o_vals = take(4 * M), o_skeys = take(4 * M), o_svals = take(4 * M);
  const size_t o_head = take(4 * M), o_rid = take(4 * M), o_rstart = take(4 * (M + 1));
  const size_t o_nch = take(4 * (M + 1)), o_nchm = take(4 * (M + 1)), o_coff = take(4 * (M + 2)), o_moff = take(4 * (M + 2));
  const size_t o_zero = off;                               // everything from here to o_zero_end is zeroed per call
  const size_t o_runmax = take(4 * M), o_ticket = take(4 * M), o_bias = take(8 * (size_t)D), o_scal = take(64);
  const size_t o_zero_end = off;
  const size_t o_part = take(8 * n_part * D), o_cub = take(t_cub);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, off, &sp));
  char* w = static_cast<char*>(sp);
  int* keys = (int*)(w + o_keys); int* vals = (int*)(w + o_vals); int* skeys = (int*)(w + o_skeys); int* svals = (int*)(w + o_svals);
  int* head = (int*)(w + o_head); int* rid = (int*)(w + o_rid); int* rstart = (int*)(w + o_rstart);
  int* nch = (int*)(w + o_nch); int* nchm = (int*)(w + o_nchm); int* coff = (int*)(w + o_coff); int* moff = (int*)(w + o_moff);
  unsigned* runmax = (unsigned*)(w + o_runmax); unsigned* ticket = (unsigned*)(w + o_ticket);
  u64* bias_acc = (u64*)(w + o_bias);
  int* num_runs = (int*)(w + o_scal); unsigned* gmax = (unsigned*)(w + o_scal + 8);
  void* cub_tmp = w + o_cub;
  size_t tb = t_cub;
  const int g256 = (int)mms_min<long long>((M + 255) / 256, (long long)ctx->sm_count * 8);

  MMS_CUDA(cudaMemsetAsync(w + o_zero, 0, o_zero_end - o_zero, st));
  { MmsKernelScope ks_(ctx, "embed_det_keys");
    det_keys_kernel<T><<<g256, 256, 0, st>>>(idx, keys, vals, M, V, ctx->fault_flag); }
  MMS_LAUNCH_CHECK();
  { MmsKernelScope ks_(ctx, "embed_det_sort");
    MMS_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, tb, (const int*)keys, skeys, (const int*)vals, svals, n, 0, bits, st)); }
  { MmsKernelScope ks_(ctx, "embed_det_runs");
    det_heads_kernel<<<g256, 256, 0, st>>>(skeys, head, M);
    MMS_LAUNCH_CHECK();
    tb = t_cub;
    MMS_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, tb, (const int*)head, rid, n, st));
    tb = t_cub;
    MMS_CUDA(cub::DeviceSelect::Flagged(cub_tmp, tb, cub::CountingInputIterator<int>(0), (const int*)head, rstart, num_runs, n, st)); }
  { MmsKernelScope ks_(ctx, "embed_det_rowmax");
    const int g = (int)mms_min<long long>((M + 7) / 8, (long long)ctx->sm_count * 16);
    det_rowmax_kernel<T><<<g, 256, 0, st>>>(dtop, svals, rid, runmax, gmax, M, D); }
  MMS_LAUNCH_CHECK();
  { MmsKernelScope ks_(ctx, "embed_det_runs");
    det_nchunks_kernel<<<g256, 256, 0, st>>>(rstart, num_runs, nch, nchm, M);
    MMS_LAUNCH_CHECK();
    tb = t_cub;
    MMS_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, tb, (const int*)nch, coff, n + 1, st));
    tb = t_cub;
    MMS_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, tb, (const int*)nchm, moff, n + 1, st)); }
  DetArgs a;
  a.skeys = skeys; a.svals = svals; a.run_start = rstart; a.num_runs = num_runs; a.chunk_off = coff; a.multi_off = moff;
  a.runmax = runmax; a.gmax = gmax; a.ticket = ticket; a.partials = (long long*)(w + o_part); a.bias_acc = bias_acc;
  a.M = M; a.D = D; a.V = V;
  const int nvec = D / VN;
  const int vpl = (nvec + 31) / 32;
  const int grid = (int)mms_min<long long>((M + kWarps - 1) / kWarps, (long long)ctx->sm_count * 8);
  const size_t smem = sizeof(u64) * (size_t)D;
  const int want_bias = dbias != nullptr;
  { MmsKernelScope ks_(ctx, "embed_det_reduce");
    switch (vpl) {
      case 1: det_reduce_kernel<T, 1><<<grid, kWarps * 32, smem, st>>>(dtop, dW, a, want_bias); break;
      case 2: det_reduce_kernel<T, 2><<<grid, kWarps * 32, smem, st>>>(dtop, dW, a, want_bias); break;
      case 3: det_reduce_kernel<T, 3><<<grid, kWarps * 32, smem, st>>>(dtop, dW, a, want_bias); break;
      case 4: det_reduce_kernel<T, 4><<<grid, kWarps * 32, smem, st>>>(dtop, dW, a, want_bias); break;
      default: det_reduce_kernel<T, kMaxVecPerLane><<<grid, kWarps * 32, smem, st>>>(dtop, dW, a, want_bias); break;
    } }
  MMS_LAUNCH_CHECK();
  if (want_bias) {
    MmsKernelScope ks_(ctx, "embed_det_bias");
    det_bias_finish_kernel<T><<<mms_ceil_div(D, 256), 256, 0, st>>>(bias_acc, gmax, dbias, M, D);
    MMS_LAUNCH_CHECK();
  }
  return 0;
}

template int mms_embed_backward_deterministic<float>(mms_context*, const float*, const float*, float*, float*, long long, int, int);
template int mms_embed_backward_deterministic<double>(mms_context*, const double*, const double*, double*, double*, long long, int, int);
