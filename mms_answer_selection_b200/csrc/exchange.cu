// Data-parallel gradient exchange over NVLink peer memory: the B200-native replacement of the reference's
// P2PSync (src/caffe/parallel.cpp).
//
// The reference keeps all learnable blobs in two flat device buffers per solver (Params / GPUParams,
// parallel.cpp:60-115), and every iteration (a) the root broadcasts the weights down a binary tree of peer
// cudaMemcpyAsync calls (on_start, :287-322), (b) the children's gradient buffers are copied up the tree and
// added level by level (on_gradients_ready, :325-380, caffe_gpu_add), (c) the root scales the sum by
// 1/solver_count (:377) and alone applies the solver update.  Every level is a stream sync on the host.
//
// Here every rank (one process -- or thread -- per GPU) owns one allocation  [flags | data | diff]  that all
// peers map (cudaIpc handles across processes, plain peer pointers inside one process, or a symmetric-memory
// allocation handed in by the host together with its NVSwitch multicast address).  One kernel per rank does the
// whole exchange of a range [begin, end) of the flat buffer; rank r owns slice r of the range:
//
//   allreduce   diff_p[i] <- scale * sum_q diff_q[i]   for every rank p       (= (b) + (c), result on every rank)
//   adadelta    the same sum, then the owner applies the AdaDelta solver step to ITS slice of the weights
//               (history only exists on the owner: optimizer state is sharded 1/world) and stores the NEW WEIGHTS
//               into every rank's data buffer, then clears the gradient range locally             (= (b)+(c)+update+(a))
//   broadcast   data_p <- data_root                                                               (= (a), used once)
//
// Slice r is read from all peers with 16-byte loads in the fixed order q = 0..world-1 (bit-identical results on
// every rank and from run to run) and the result is stored to all peers with 16-byte stores -- a reduce-scatter and
// an all-gather in one pass, every byte crossing NVLink once in each direction.  With a multicast address the loop
// body becomes one `multimem.ld_reduce.add.v4.f32` (the NVSwitch adds the eight replicas in flight) and one
// `multimem.st.v4.f32` (the switch replicates the store), which divides the bytes the SMs have to issue by `world`.
//
// Cross-rank ordering is two flag rounds per call in the peers' flag blocks (system-scope release/acquire):
// READY (my gradient range is final -- sent when the kernel starts, i.e. after everything queued before it on the
// stream) and DONE (all my stores into your buffers are performed -- sent by my last CTA).  A kernel exits only
// after it has seen DONE from every peer, so whatever follows on the stream sees the complete result.  The epoch
// lives in device memory, which makes a call replayable inside a CUDA graph.  Waits are bounded (default 10 s):
// a missing peer raises a fault flag that mms_exchange_check reports instead of hanging the GPU.
#include <new>

#include "adadelta.cuh"
#include "mms_common.cuh"

namespace {

constexpr int kMaxWorld = MMS_EXCHANGE_MAX_WORLD;
constexpr int kChannels = MMS_EXCHANGE_CHANNELS;
constexpr size_t kHeaderBytes = 4096;
constexpr int kThreads = 256;   // 256 x ~96 registers: a CTA leaves room for a contraction CTA on the same SM
constexpr int kMaxSegs = 8;

struct XChan {                       // one per channel, 256 bytes
  unsigned ready[kMaxWorld];         // ready[q] = last epoch rank q announced "my range is final"
  unsigned done[kMaxWorld];          // done[q]  = last epoch rank q announced "my stores into you are performed"
  unsigned epoch;                    // last completed epoch of this rank on this channel
  unsigned cta_done;                 // CTAs of the running kernel that have finished their slice
  unsigned fault;                    // 1: a wait timed out
  unsigned pad[64 - 2 * kMaxWorld - 3];
};
static_assert(sizeof(XChan) == 256, "channel block");
static_assert(sizeof(XChan) * kChannels <= kHeaderBytes, "header");

struct XDev {
  char* base[kMaxWorld];
  char* mc;                          // multicast mapping of all bases (or null)
  size_t data_off, diff_off;
  int rank, world;
  unsigned long long timeout_ns;
};

struct XSegs {                       // per-blob solver multipliers over the flat buffer (element offsets)
  int n;
  long long end[kMaxSegs];
  double rate[kMaxSegs], decay[kMaxSegs];
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ XChan* chan_of(char* base, int channel) {
  return reinterpret_cast<XChan*>(base) + channel;
}
// spin until *flag >= epoch (epochs are monotonic; the difference test survives wrap-around)
__device__ __forceinline__ bool wait_flag(const unsigned* flag, unsigned epoch, unsigned long long timeout_ns) {
  const unsigned long long t0 = now_ns();
  unsigned spins = 0;
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
    if ((++spins & 63u) == 0) {
      __nanosleep(64);
      if (now_ns() - t0 > timeout_ns) return false;
    }
  }
  return true;
}

template <typename T> struct V16;
template <> struct V16<float> { typedef float4 type; static constexpr int n = 4; };
template <> struct V16<double> { typedef double2 type; static constexpr int n = 2; };

__device__ __forceinline__ float4 ld16(const float4* p) { return __ldcg(p); }      // L2 only: never a stale L1 line
__device__ __forceinline__ double2 ld16(const double2* p) { return __ldcg(p); }
__device__ __forceinline__ void st16(float4* p, float4 v) { __stcg(p, v); }
__device__ __forceinline__ void st16(double2* p, double2 v) { __stcg(p, v); }
__device__ __forceinline__ void vacc(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void vacc(double2& a, const double2 b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void vscale(float4& a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }
__device__ __forceinline__ void vscale(double2& a, double s) { a.x *= s; a.y *= s; }

// NVSwitch in-fabric reduction / replication through the multicast mapping (float only)
__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float4* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double2 mc_ld_reduce(const double2* p) { return *p; }   // never instantiated with MC
__device__ __forceinline__ void mc_st(double2*, double2) {}

enum { kAllreduce = 0, kAdadelta = 1, kBroadcast = 2 };

template <typename T>
struct XArgs {
  long long begin, end;              // element range of the flat buffers, multiples of 16 bytes
  T scale;                           // allreduce: result scale; adadelta: gradient scale (1/world, 1/iter_size)
  T momentum, delta;
  T* hist_g; T* hist_u;              // adadelta: full-length history arrays of this rank (only its slices are used)
  int clear_diff;                    // adadelta: leave this rank's gradient range zeroed
  int root;                          // broadcast
  int channel;
};

// READY round: CTA 0 announces, every CTA waits for every peer before it touches peer memory.
__device__ __forceinline__ bool ready_round(const XDev& d, XChan* my, int channel, unsigned epoch) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < d.world) {
    if (blockIdx.x == 0) st_release_sys(&chan_of(d.base[threadIdx.x], channel)->ready[d.rank], epoch);
    if (!wait_flag(&my->ready[threadIdx.x], epoch, d.timeout_ns)) { s_ok = 0; atomicExch(&my->fault, 1u); }
  }
  __syncthreads();
  return s_ok != 0;
}

// DONE round: the CTA that finishes last tells every peer and waits for every peer, then closes the epoch.
__device__ __forceinline__ void done_round(const XDev& d, XChan* my, int channel, unsigned epoch) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                            // this CTA's peer stores, ordered before the ticket
    s_last = atomicAdd(&my->cta_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < d.world) {
    __threadfence_system();
    st_release_sys(&chan_of(d.base[threadIdx.x], channel)->done[d.rank], epoch);
    if (!wait_flag(&my->done[threadIdx.x], epoch, d.timeout_ns)) atomicExch(&my->fault, 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    my->cta_done = 0;
    __threadfence();
    *reinterpret_cast<volatile unsigned*>(&my->epoch) = epoch;
  }
}

// U = 16-byte units per thread and trip: U x world peer loads are in flight per thread (latency, not bandwidth, is what a
// thread waits for on NVLink), so U is chosen as 16 / world by the host
template <typename T, int MODE, bool MC, int U>
__global__ void __launch_bounds__(kThreads)
exchange_kernel(const XDev d, const XArgs<T> a, const XSegs segs) {
  typedef typename V16<T>::type VT;
  constexpr int VN = V16<T>::n;
  constexpr int W = MC ? 1 : 16 / U;          // peers this instantiation walks (world <= W)
  XChan* my = chan_of(d.base[d.rank], a.channel);
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(&my->epoch) + 1u;
  const bool ok = ready_round(d, my, a.channel, epoch);
  const long long n16 = (a.end - a.begin) / VN, b16 = a.begin / VN;
  if (ok) {
    if (MODE == kBroadcast) {
      if (d.rank != a.root) {
        const VT* src = reinterpret_cast<const VT*>(d.base[a.root] + d.data_off) + b16;
        VT* dst = reinterpret_cast<VT*>(d.base[d.rank] + d.data_off) + b16;
        for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n16; i += (long long)gridDim.x * kThreads)
          st16(dst + i, ld16(src + i));
      }
    } else {
      // slice of this rank, in 16-byte units of the range
      const long long lo = n16 * d.rank / d.world, hi = n16 * (d.rank + 1) / d.world;
      const VT* mc_diff = reinterpret_cast<const VT*>(d.mc + d.diff_off) + b16;
      for (long long i0 = lo + blockIdx.x * (long long)kThreads * U; i0 < hi; i0 += (long long)gridDim.x * kThreads * U) {
        VT acc[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {                    // U independent 16-byte columns per thread
          const long long i = i0 + u * kThreads + threadIdx.x;
          live[u] = i < hi;
          if (!live[u]) continue;
          if (MC) {
            acc[u] = mc_ld_reduce(mc_diff + i);
          } else {
            VT part[W];
#pragma unroll
            for (int q = 0; q < W; ++q)
              if (q < d.world) part[q] = ld16(reinterpret_cast<const VT*>(d.base[q] + d.diff_off) + b16 + i);
            acc[u] = part[0];
#pragma unroll
            for (int q = 1; q < W; ++q)
              if (q < d.world) vacc(acc[u], part[q]);      // fixed order 0..world-1: the same bits on every rank
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!live[u]) continue;
          const long long i = i0 + u * kThreads + threadIdx.x;
          if (MODE == kAllreduce) {
            vscale(acc[u], a.scale);
            if (MC) {
              mc_st(reinterpret_cast<VT*>(d.mc + d.diff_off) + b16 + i, acc[u]);
            } else {
#pragma unroll
              for (int q = 0; q < W; ++q)
                if (q < d.world) st16(reinterpret_cast<VT*>(d.base[q] + d.diff_off) + b16 + i, acc[u]);
            }
          } else {   // kAdadelta: the owner updates its slice of the weights and publishes the new weights
            const long long e0 = a.begin + i * VN;         // first element of this 16-byte unit
            int s = 0;
            while (s + 1 < segs.n && e0 >= segs.end[s]) ++s;
            const T rate = (T)segs.rate[s], decay = (T)segs.decay[s];
            VT* wl = reinterpret_cast<VT*>(d.base[d.rank] + d.data_off) + b16 + i;
            VT w = *wl;
            VT hg = reinterpret_cast<VT*>(a.hist_g)[b16 + i], hu = reinterpret_cast<VT*>(a.hist_u)[b16 + i];
            T* wp = reinterpret_cast<T*>(&w); T* gp = reinterpret_cast<T*>(&acc[u]);
            T* hgp = reinterpret_cast<T*>(&hg); T* hup = reinterpret_cast<T*>(&hu);
#pragma unroll
            for (int c = 0; c < VN; ++c)
              adadelta_one<T>(wp[c], gp[c], hgp[c], hup[c], true, a.scale, decay, a.momentum, a.delta, rate, false);
            reinterpret_cast<VT*>(a.hist_g)[b16 + i] = hg;
            reinterpret_cast<VT*>(a.hist_u)[b16 + i] = hu;
            if (MC) {
              mc_st(reinterpret_cast<VT*>(d.mc + d.data_off) + b16 + i, w);
            } else {
#pragma unroll
              for (int q = 0; q < W; ++q)
                if (q < d.world) st16(reinterpret_cast<VT*>(d.base[q] + d.data_off) + b16 + i, w);
            }
          }
        }
      }
    }
  }
  done_round(d, my, a.channel, epoch);
}

template <typename T>
__global__ void __launch_bounds__(256) scale_range_kernel(T* x, long long n, T alpha) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) x[i] *= alpha;
}

size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

struct mms_exchange {
  int rank = 0, world = 1, device = 0, sm_count = 148;
  long long count = 0;
  int elem = 4;
  char* base = nullptr;              // this rank's allocation
  bool owns_base = false;
  size_t bytes = 0;
  char* peer[kMaxWorld] = {};
  bool opened[kMaxWorld] = {};       // peer mapped through cudaIpcOpenMemHandle (to be closed)
  char* mc = nullptr;
  bool attached = false;
  void* hist_g = nullptr; void* hist_u = nullptr;
  int ctas = 0;                      // 0: one per SM
  long long timeout_ms = 10000;
  unsigned long long launches = 0;
};

namespace {

XDev dev_of(const mms_exchange* x) {
  XDev d;
  for (int q = 0; q < kMaxWorld; ++q) d.base[q] = q < x->world ? x->peer[q] : nullptr;
  d.mc = x->mc;
  d.data_off = kHeaderBytes;
  d.diff_off = kHeaderBytes + pad256((size_t)x->count * x->elem);
  d.rank = x->rank; d.world = x->world;
  d.timeout_ns = (unsigned long long)x->timeout_ms * 1000000ULL;
  return d;
}

int check_range(const mms_exchange* x, long long begin, long long end, int channel) {
  MMS_REQUIRE(x, MMS_E_INVALID, "null exchange");
  MMS_REQUIRE(x->attached, MMS_E_INVALID, "exchange not attached to its peers yet");
  MMS_REQUIRE(0 <= begin && begin <= end && end <= x->count, MMS_E_INVALID, "bad range");
  const int vn = 16 / x->elem;
  MMS_REQUIRE(begin % vn == 0 && (end % vn == 0 || end == x->count), MMS_E_INVALID,
              "range bounds must be multiples of 16 bytes (pad blob offsets)");
  MMS_REQUIRE(channel >= 0 && channel < kChannels, MMS_E_INVALID, "bad channel");
  return 0;
}

int grid_for(const mms_exchange* x, long long units, int U) {
  const long long want = (units + kThreads * U - 1) / (kThreads * U);
  const int cap = x->ctas > 0 ? x->ctas : x->sm_count;
  return (int)mms_max<long long>(1, mms_min<long long>(want, cap));
}

template <typename T>
int launch(mms_exchange* x, cudaStream_t st, int mode, XArgs<T> a, const XSegs& segs) {
  const XDev d = dev_of(x);
  const int vn = 16 / (int)sizeof(T);
  a.end = (a.end + vn - 1) / vn * vn;           // the buffers are padded to 256 bytes: a ragged tail rounds up
  const long long n16 = (a.end - a.begin) / vn;
  if (n16 <= 0) return 0;
  const long long units = mode == kBroadcast ? n16 : (n16 + x->world - 1) / x->world;
  const bool mc = x->mc != nullptr && sizeof(T) == 4;
  const int U = (mc || x->world <= 2) ? 8 : (x->world <= 4 ? 4 : 2);      // world <= 16 / U
  const int grid = grid_for(x, units, mode == kBroadcast ? 2 : U);
  x->launches++;
#define MMS_X_LAUNCH(MODE, MCF, UU)                                                   \
  do {                                                                                \
    MMS_CARVEOUT((exchange_kernel<T, MODE, MCF, UU>));                                 \
    exchange_kernel<T, MODE, MCF, UU><<<grid, kThreads, 0, st>>>(d, a, segs);          \
  } while (0)
#define MMS_X_BY_U(MODE)                                                             \
  do {                                                                               \
    if (mc) MMS_X_LAUNCH(MODE, true, 8);                                             \
    else if (U == 8) MMS_X_LAUNCH(MODE, false, 8);                                   \
    else if (U == 4) MMS_X_LAUNCH(MODE, false, 4);                                   \
    else MMS_X_LAUNCH(MODE, false, 2);                                               \
  } while (0)
  if (mode == kAllreduce) MMS_X_BY_U(kAllreduce);
  else if (mode == kAdadelta) MMS_X_BY_U(kAdadelta);
  else MMS_X_LAUNCH(kBroadcast, false, 2);
#undef MMS_X_BY_U
#undef MMS_X_LAUNCH
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int allreduce_impl(mms_exchange* x, cudaStream_t st, int channel, long long begin, long long end, double scale) {
  MMS_TRY(check_range(x, begin, end, channel));
  MMS_REQUIRE(x->elem == (int)sizeof(T), MMS_E_INVALID, "element size does not match the call's type");
  if (end == begin) return 0;
  if (x->world == 1) {                            // nothing to exchange: the 1/n scale alone (parallel.cpp:377)
    if (scale != 1.0) {
      T* diff = reinterpret_cast<T*>(x->base + kHeaderBytes + pad256((size_t)x->count * x->elem)) + begin;
      x->launches++;
      scale_range_kernel<T><<<mms_min<long long>(4 * x->sm_count, (end - begin + 255) / 256), 256, 0, st>>>(diff, end - begin, (T)scale);
      MMS_LAUNCH_CHECK();
    }
    return 0;
  }
  XArgs<T> a = {};
  a.begin = begin; a.end = end; a.scale = (T)scale; a.channel = channel;
  XSegs segs = {};
  return launch<T>(x, st, kAllreduce, a, segs);
}

template <typename T>
int adadelta_impl(mms_exchange* x, cudaStream_t st, int channel, long long begin, long long end, double grad_scale,
                  const long long* seg_end, const double* seg_rate, const double* seg_decay, int nseg,
                  double momentum, double delta, int clear_diff) {
  MMS_TRY(check_range(x, begin, end, channel));
  MMS_REQUIRE(x->elem == (int)sizeof(T), MMS_E_INVALID, "element size does not match the call's type");
  MMS_REQUIRE(nseg >= 1 && nseg <= kMaxSegs && seg_end && seg_rate && seg_decay, MMS_E_INVALID, "1..8 segments");
  if (end == begin) return 0;
  if (!x->hist_g) {                               // AdaDeltaPreSolve (adadelta_solver.cpp:12-22): zero history
    const size_t hb = pad256((size_t)x->count * x->elem);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    MMS_REQUIRE(cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone, MMS_E_INVALID,
                "the first fused solver step allocates the history: run it once before capturing a graph");
    MMS_CUDA(cudaMalloc(&x->hist_g, hb)); MMS_CUDA(cudaMalloc(&x->hist_u, hb));
    MMS_CUDA(cudaMemset(x->hist_g, 0, hb)); MMS_CUDA(cudaMemset(x->hist_u, 0, hb));
  }
  XArgs<T> a = {};
  a.begin = begin; a.end = end; a.scale = (T)grad_scale; a.momentum = (T)momentum; a.delta = (T)delta;
  a.hist_g = static_cast<T*>(x->hist_g); a.hist_u = static_cast<T*>(x->hist_u);
  a.clear_diff = clear_diff; a.channel = channel;
  XSegs segs = {};
  segs.n = nseg;
  for (int s = 0; s < nseg; ++s) { segs.end[s] = seg_end[s]; segs.rate[s] = seg_rate[s]; segs.decay[s] = seg_decay[s]; }
  mms_note_write(x->base + kHeaderBytes + (size_t)begin * x->elem, (size_t)(end - begin) * x->elem);   // weights change
  MMS_TRY(launch<T>(x, st, kAdadelta, a, segs));
  if (clear_diff) {   // Net::ClearParamDiffs of the next iteration (solver.cpp:203): every peer has read this range
    char* diff = x->base + kHeaderBytes + pad256((size_t)x->count * x->elem);
    MMS_CUDA(cudaMemsetAsync(diff + (size_t)begin * x->elem, 0, (size_t)(end - begin) * x->elem, st));
  }
  return 0;
}

}  // namespace

extern "C" {

long long mms_exchange_bytes(long long count, int elem_bytes) {
  if (count < 0 || (elem_bytes != 4 && elem_bytes != 8)) return -1;
  return (long long)(kHeaderBytes + 2 * pad256((size_t)count * elem_bytes));
}

int mms_exchange_create(mms_exchange_t* out, int rank, int world, long long count, int elem_bytes, void* external_base) {
  MMS_REQUIRE(out, MMS_E_INVALID, "null pointer");
  *out = nullptr;
  MMS_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, MMS_E_INVALID, "bad rank / world (<= 16)");
  MMS_REQUIRE(count > 0 && (elem_bytes == 4 || elem_bytes == 8), MMS_E_INVALID, "bad size");
  int dev = 0, major = 0, sms = 0;
  MMS_CUDA(cudaGetDevice(&dev));
  MMS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MMS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MMS_REQUIRE(major == 10, MMS_E_UNSUPPORTED, "libmms_b200 is built for sm_100a only; no CPU or other-GPU fallback exists");
  mms_exchange* x = new (std::nothrow) mms_exchange();
  MMS_REQUIRE(x, MMS_E_NOMEM, "out of host memory");
  x->rank = rank; x->world = world; x->device = dev; x->sm_count = sms; x->count = count; x->elem = elem_bytes;
  x->bytes = (size_t)mms_exchange_bytes(count, elem_bytes);
  if (external_base) {
    x->base = static_cast<char*>(external_base);
  } else {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->base), x->bytes);
    if (e != cudaSuccess) {
      mms_set_error("exchange allocation of %zu bytes failed: %s", x->bytes, cudaGetErrorString(e));
      delete x;
      return MMS_E_NOMEM;
    }
    x->owns_base = true;
  }
  cudaError_t e = cudaMemset(x->base, 0, x->bytes);
  if (e != cudaSuccess) {
    mms_set_error("exchange allocation could not be cleared: %s", cudaGetErrorString(e));
    if (x->owns_base) cudaFree(x->base);
    delete x;
    return (int)e;
  }
  x->peer[rank] = x->base;
  if (world == 1) x->attached = true;
  *out = x;
  return 0;
}

int mms_exchange_destroy(mms_exchange_t x) {
  if (!x) return 0;
  cudaDeviceSynchronize();
  for (int q = 0; q < x->world; ++q)
    if (x->opened[q]) cudaIpcCloseMemHandle(x->peer[q]);
  if (x->hist_g) cudaFree(x->hist_g);
  if (x->hist_u) cudaFree(x->hist_u);
  if (x->owns_base) cudaFree(x->base);
  delete x;
  return 0;
}

int mms_exchange_buffers(mms_exchange_t x, void** data, void** diff) {
  MMS_REQUIRE(x, MMS_E_INVALID, "null exchange");
  if (data) *data = x->base + kHeaderBytes;
  if (diff) *diff = x->base + kHeaderBytes + pad256((size_t)x->count * x->elem);
  return 0;
}

int mms_exchange_base(mms_exchange_t x, void** base) {
  MMS_REQUIRE(x && base, MMS_E_INVALID, "null argument");
  *base = x->base;
  return 0;
}

int mms_exchange_export_ipc(mms_exchange_t x, void* handle64) {
  MMS_REQUIRE(x && handle64, MMS_E_INVALID, "null argument");
  MMS_REQUIRE(x->owns_base, MMS_E_INVALID, "only library-allocated buffers can be exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == MMS_EXCHANGE_IPC_BYTES, "ipc handle size");
  MMS_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), x->base));
  return 0;
}

int mms_exchange_attach_ipc(mms_exchange_t x, const void* handles) {
  MMS_REQUIRE(x && handles, MMS_E_INVALID, "null argument");
  MMS_REQUIRE(!x->attached || x->world == 1, MMS_E_INVALID, "already attached");
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
  for (int q = 0; q < x->world; ++q) {
    if (q == x->rank) continue;
    void* p = nullptr;
    MMS_CUDA(cudaIpcOpenMemHandle(&p, h[q], cudaIpcMemLazyEnablePeerAccess));
    x->peer[q] = static_cast<char*>(p);
    x->opened[q] = true;
  }
  x->attached = true;
  return 0;
}

int mms_exchange_attach_ptrs(mms_exchange_t x, void* const* peer_bases, void* multicast_base) {
  MMS_REQUIRE(x && peer_bases, MMS_E_INVALID, "null argument");
  MMS_REQUIRE(peer_bases[x->rank] == x->base, MMS_E_INVALID, "peer_bases[rank] must be this rank's own allocation");
  for (int q = 0; q < x->world; ++q) {
    MMS_REQUIRE(peer_bases[q], MMS_E_INVALID, "null peer pointer");
    x->peer[q] = static_cast<char*>(peer_bases[q]);
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, peer_bases[q]) == cudaSuccess && at.type == cudaMemoryTypeDevice &&
        at.device != x->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);     // same process, another GPU (P2PSync's threads)
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        mms_set_error("peer access from device %d to %d: %s", x->device, at.device, cudaGetErrorString(e));
        return (int)e;
      }
    }
    cudaGetLastError();
  }
  x->mc = static_cast<char*>(multicast_base);
  x->attached = true;
  return 0;
}

int mms_exchange_set_option(mms_exchange_t x, int option, long long value) {
  MMS_REQUIRE(x, MMS_E_INVALID, "null exchange");
  switch (option) {
    case MMS_EXCHANGE_OPT_CTAS: MMS_REQUIRE(value >= 0 && value <= 4096, MMS_E_INVALID, "bad CTA count"); x->ctas = (int)value; return 0;
    case MMS_EXCHANGE_OPT_TIMEOUT_MS: MMS_REQUIRE(value > 0, MMS_E_INVALID, "bad timeout"); x->timeout_ms = value; return 0;
    case MMS_EXCHANGE_OPT_MULTICAST: if (!value) x->mc = nullptr; return 0;
    default: mms_set_error("unknown exchange option %d", option); return MMS_E_INVALID;
  }
}

unsigned long long mms_exchange_launch_count(mms_exchange_t x) { return x ? x->launches : 0; }

int mms_exchange_check(mms_exchange_t x, void* stream) {
  MMS_REQUIRE(x, MMS_E_INVALID, "null exchange");
  MMS_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  XChan h[kChannels];
  MMS_CUDA(cudaMemcpy(h, x->base, sizeof(h), cudaMemcpyDeviceToHost));
  for (int c = 0; c < kChannels; ++c)
    if (h[c].fault) {
      MMS_CUDA(cudaMemset(x->base + c * sizeof(XChan) + offsetof(XChan, fault), 0, sizeof(unsigned)));
      mms_set_error("gradient exchange, channel %d: a peer did not arrive within %lld ms", c, x->timeout_ms);
      return MMS_E_FAULT;
    }
  return 0;
}

int mms_exchange_allreduce_f32(mms_exchange_t x, void* stream, int channel, long long begin, long long end, float scale) {
  return allreduce_impl<float>(x, static_cast<cudaStream_t>(stream), channel, begin, end, scale);
}
int mms_exchange_allreduce_f64(mms_exchange_t x, void* stream, int channel, long long begin, long long end, double scale) {
  return allreduce_impl<double>(x, static_cast<cudaStream_t>(stream), channel, begin, end, scale);
}

int mms_exchange_adadelta_f32(mms_exchange_t x, void* stream, int channel, long long begin, long long end, float grad_scale,
                              const long long* seg_end, const double* seg_rate, const double* seg_decay, int nseg,
                              float momentum, float delta, int clear_diff) {
  return adadelta_impl<float>(x, static_cast<cudaStream_t>(stream), channel, begin, end, grad_scale, seg_end, seg_rate,
                              seg_decay, nseg, momentum, delta, clear_diff);
}
int mms_exchange_adadelta_f64(mms_exchange_t x, void* stream, int channel, long long begin, long long end, double grad_scale,
                              const long long* seg_end, const double* seg_rate, const double* seg_decay, int nseg,
                              double momentum, double delta, int clear_diff) {
  return adadelta_impl<double>(x, static_cast<cudaStream_t>(stream), channel, begin, end, grad_scale, seg_end, seg_rate,
                               seg_decay, nseg, momentum, delta, clear_diff);
}

int mms_exchange_broadcast(mms_exchange_t x, void* stream, int channel, int root) {
  MMS_TRY(check_range(x, 0, x ? x->count : 0, channel));
  MMS_REQUIRE(root >= 0 && root < x->world, MMS_E_INVALID, "bad root");
  if (x->world == 1) return 0;
  mms_note_write(x->base + kHeaderBytes, (size_t)x->count * x->elem);
  XSegs segs = {};
  if (x->elem == 4) {
    XArgs<float> a = {};
    a.begin = 0; a.end = x->count; a.root = root; a.channel = channel;
    return launch<float>(x, static_cast<cudaStream_t>(stream), kBroadcast, a, segs);
  }
  XArgs<double> a = {};
  a.begin = 0; a.end = x->count; a.root = root; a.channel = channel;
  return launch<double>(x, static_cast<cudaStream_t>(stream), kBroadcast, a, segs);
}

int mms_exchange_history(mms_exchange_t x, void** hist_g, void** hist_u) {
  MMS_REQUIRE(x, MMS_E_INVALID, "null exchange");
  if (hist_g) *hist_g = x->hist_g;
  if (hist_u) *hist_u = x->hist_u;
  return 0;
}

}  // extern "C"
