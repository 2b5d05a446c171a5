// C-ABI of libmms_b200.so (declared in include/mms_b200.h): context management and thin
// extern "C" wrappers over the templated layer implementations.
#include <stdarg.h>
#include <string.h>

#include <new>

#include "mms_common.cuh"
#include "tc/tc_gemm.cuh"

static thread_local char g_err[512] = "";

void mms_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#include <mutex>
namespace {
struct WriteNote { uintptr_t lo, hi; unsigned long long clock; };
constexpr int kWriteLog = 256;
WriteNote g_writes[kWriteLog];
unsigned long long g_write_clock = 1;      // clock of the next note; notes [clock - kWriteLog, clock) are retained
std::mutex g_write_mu;
}  // namespace
unsigned long long mms_write_clock() {
  std::lock_guard<std::mutex> lk(g_write_mu);
  return g_write_clock;
}
void mms_note_write(const void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_write_mu);
  WriteNote& w = g_writes[g_write_clock % kWriteLog];
  w.lo = p ? reinterpret_cast<uintptr_t>(p) : 0;
  w.hi = p ? w.lo + bytes : ~(uintptr_t)0;
  w.clock = g_write_clock++;
}
bool mms_unchanged_since(unsigned long long clock, const void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_write_mu);
  if (g_write_clock - clock > (unsigned long long)kWriteLog) return false;     // older notes are gone: assume changed
  const uintptr_t lo = reinterpret_cast<uintptr_t>(p), hi = lo + bytes;
  for (unsigned long long c = clock; c < g_write_clock; ++c) {
    const WriteNote& w = g_writes[c % kWriteLog];
    if (w.lo < hi && lo < w.hi) return false;
  }
  return true;
}

// Has a write to exactly this range (not a blanket "everything may have changed") been noted since `clock`?
static bool written_since(unsigned long long clock, const void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_write_mu);
  const uintptr_t lo = reinterpret_cast<uintptr_t>(p), hi = lo + bytes;
  const unsigned long long first = g_write_clock - clock > (unsigned long long)kWriteLog ? g_write_clock - kWriteLog : clock;
  for (unsigned long long c = first; c < g_write_clock; ++c) {
    const WriteNote& w = g_writes[c % kWriteLog];
    if (w.lo != 0 && w.lo < hi && lo < w.hi) return true;
  }
  return false;
}

int mms_scratch(mms_context* ctx, size_t bytes, void** out) {
  ctx->fwd_cache.valid = false;        // whoever asks for the scratch buffer is about to overwrite it
  ctx->m_prepared.valid = false;       // (mms_tc_simcross2_forward reads this flag BEFORE it asks)
  ctx->sent_cache.valid = false;
  ctx->simmat_cache.valid = false;
  ctx->dm_pending.valid = false;
  if (bytes > ctx->scratch_bytes) {
    // growing means cudaStreamSynchronize + cudaFree + cudaMalloc, none of which is legal while the stream is being
    // captured into a CUDA graph: the workspace must have its size before the capture starts (run the step once
    // eagerly, or call mms_reserve_scratch)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
      mms_set_error("workspace of %zu bytes needed but only %zu reserved while the stream is being captured: "
                    "pre-size it with mms_reserve_scratch or one eager call before the capture", bytes, ctx->scratch_bytes);
      return MMS_E_INVALID;
    }
    if (ctx->scratch) {
      // earlier launches on the stream and on the private streams may still read the old buffer
      MMS_CUDA(cudaStreamSynchronize(ctx->stream));
      for (int i = 0; i < 2; ++i)
        if (ctx->side[i]) MMS_CUDA(cudaStreamSynchronize(ctx->side[i]));
      MMS_CUDA(cudaFree(ctx->scratch));
      ctx->scratch = nullptr;
      ctx->scratch_bytes = 0;
    }
    const size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    cudaError_t e = cudaMalloc(&ctx->scratch, want);
    if (e != cudaSuccess) {
      mms_set_error("scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
      return MMS_E_NOMEM;
    }
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return 0;
}

#include <unordered_map>
namespace {
struct StageEntry { mms_context* owner; const float* staged; long long rows; int cols, ld; unsigned long long clock; bool virt; };
std::unordered_map<const void*, StageEntry> g_stage;
std::mutex g_stage_mu;
}  // namespace
void mms_stage_publish(mms_context* owner, const void* src, const float* staged, long long rows, int cols, int ld, bool virt) {
  const unsigned long long clock = mms_write_clock();
  std::lock_guard<std::mutex> lk(g_stage_mu);
  g_stage[src] = StageEntry{owner, staged, rows, cols, ld, clock, virt};
}
const float* mms_stage_lookup(const void* src, long long rows, int cols, int ld) {
  StageEntry e;
  {
    std::lock_guard<std::mutex> lk(g_stage_mu);
    auto it = g_stage.find(src);
    if (it == g_stage.end()) return nullptr;
    e = it->second;
  }
  if (e.rows != rows || e.cols != cols || e.ld != ld) return nullptr;
  return mms_unchanged_since(e.clock, src, sizeof(float) * (size_t)rows * cols) ? e.staged : nullptr;
}
bool mms_stage_virtual(const void* src) {
  StageEntry e;
  {
    std::lock_guard<std::mutex> lk(g_stage_mu);
    auto it = g_stage.find(src);
    if (it == g_stage.end() || !it->second.virt) return false;
    e = it->second;
  }
  // only a later write INTO this top through the library puts real data there; mms_invalidate_caches() ("everything may
  // have changed") makes the staged copy stale but does not materialise the top -- the consumer then fails loudly
  return !written_since(e.clock, src, sizeof(float) * (size_t)e.rows * e.cols);
}
int mms_stage_require_real(const void* src, bool have_staged, const char* what) {
  if (have_staged || !mms_stage_virtual(src)) return 0;
  mms_set_error("%s was produced with MMS_OPT_STAGE_ONLY (its fp32 values were never written) and this call cannot use the "
                "staged copy: only mode 2 on the tensor-core path with the whole batch in one chunk can", what);
  return MMS_E_INVALID;
}
void mms_stage_drop_owner(mms_context* owner) {
  std::lock_guard<std::mutex> lk(g_stage_mu);
  for (auto it = g_stage.begin(); it != g_stage.end();) it = it->second.owner == owner ? g_stage.erase(it) : ++it;
}

#include <set>
int mms_prefer_max_shared(const void* func) {
  static std::mutex mu;
  static std::set<const void*> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(func)) return 0;
  MMS_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  done.insert(func);
  return 0;
}

void mms_tc_destroy_state(mms_context* ctx);

int mms_fork(mms_context* ctx, int i) {
  if (!ctx->side[i]) {
    MMS_CUDA(cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking));
    MMS_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork[i], cudaEventDisableTiming));
    MMS_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
  }
  MMS_CUDA(cudaEventRecord(ctx->ev_fork[i], ctx->stream));
  MMS_CUDA(cudaStreamWaitEvent(ctx->side[i], ctx->ev_fork[i], 0));
  return 0;
}

int mms_join(mms_context* ctx, int i) {
  MMS_CUDA(cudaEventRecord(ctx->ev_join[i], ctx->side[i]));
  MMS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
  return 0;
}

// ---- per-launch CUDA-event profiler -----------------------------------------------------
#include <map>
#include <string>
#include <vector>
namespace {
struct ProfRecord { const char* name; cudaEvent_t e0, e1; };
std::vector<ProfRecord>* prof_of(mms_context* ctx) {
  if (!ctx->prof) ctx->prof = new std::vector<ProfRecord>();
  return static_cast<std::vector<ProfRecord>*>(ctx->prof);
}
void prof_clear(mms_context* ctx) {
  if (!ctx->prof) return;
  for (auto& r : *prof_of(ctx)) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  prof_of(ctx)->clear();
}
}  // namespace

MmsKernelScope::MmsKernelScope(mms_context* c, const char* name) : ctx(c), slot(-1) {
  ctx->launches++;
  if (!ctx->profile) return;
  ProfRecord r; r.name = name;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, ctx->stream);
  prof_of(ctx)->push_back(r);
  slot = (int)prof_of(ctx)->size() - 1;
}
MmsKernelScope::~MmsKernelScope() {
  if (slot >= 0) cudaEventRecord((*prof_of(ctx))[slot].e1, ctx->stream);
}

extern "C" {

const char* mms_last_error(void) { return g_err; }
int mms_version(void) { return MMS_B200_VERSION; }

int mms_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

int mms_create(mms_handle_t* out) {
  MMS_REQUIRE(out, MMS_E_INVALID, "null handle pointer");
  *out = nullptr;
  int dev = 0;
  MMS_CUDA(cudaGetDevice(&dev));
  int major = 0, sms = 0;
  MMS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MMS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MMS_REQUIRE(major == 10, MMS_E_UNSUPPORTED,
              "libmms_b200 is built for sm_100a only; no CPU or other-GPU fallback exists");
  mms_context* ctx = new (std::nothrow) mms_context();
  MMS_REQUIRE(ctx, MMS_E_NOMEM, "out of host memory");
  ctx->device = dev;
  ctx->sm_count = sms;
  cudaError_t e = cudaMalloc(&ctx->fault_flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(ctx->fault_flag, 0, sizeof(int));
  // 1024 block partials + one ticket word (zero between launches) for the single-launch reductions
  if (e == cudaSuccess) e = cudaMalloc(&ctx->partials, 1025 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(ctx->partials, 0, 1025 * sizeof(double));
  if (e != cudaSuccess) {
    mms_set_error("context allocation failed: %s", cudaGetErrorString(e));
    delete ctx;
    return (int)e;
  }
  *out = ctx;
  return 0;
}

int mms_destroy(mms_handle_t h) {
  if (!h) return 0;
  cudaStreamSynchronize(h->stream);
  mms_tc_destroy_state(h);
  mms_embed_plan_destroy(h);
  prof_clear(h);
  delete static_cast<std::vector<ProfRecord>*>(h->prof);
  mms_stage_drop_owner(h);
  if (h->stage_buf) cudaFree(h->stage_buf);
  if (h->scratch) cudaFree(h->scratch);
  if (h->fault_flag) cudaFree(h->fault_flag);
  if (h->partials) cudaFree(h->partials);
  for (int i = 0; i < 2; ++i) {
    if (h->side[i]) { cudaStreamSynchronize(h->side[i]); cudaStreamDestroy(h->side[i]); }
    if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  delete h;
  return 0;
}

int mms_set_stream(mms_handle_t h, void* s) {
  MMS_REQUIRE(h, MMS_E_INVALID, "null handle");
  h->stream = static_cast<cudaStream_t>(s);
  return 0;
}

int mms_set_option(mms_handle_t h, int option, long long value) {
  MMS_REQUIRE(h, MMS_E_INVALID, "null handle");
  switch (option) {
    case MMS_OPT_MATH:
      MMS_REQUIRE(value == MMS_MATH_TF32 || value == MMS_MATH_FP32, MMS_E_INVALID, "bad math mode");
      h->math = (int)value; return 0;
    case MMS_OPT_PRL_GE: h->prl_ge = value != 0; return 0;
    case MMS_OPT_SCRATCH_BYTES:
      MMS_REQUIRE(value >= (1 << 20), MMS_E_INVALID, "scratch cap below 1 MiB");
      h->scratch_cap = (size_t)value; return 0;
    case MMS_OPT_EMBED_DETERMINISTIC: h->embed_deterministic = value != 0; return 0;
    case MMS_OPT_REUSE_FORWARD: h->reuse_forward = value != 0; h->fwd_cache.valid = false; return 0;
    case MMS_OPT_CONCURRENCY: h->concurrency = value != 0; return 0;
    case MMS_OPT_STAGE_TF32:
      h->stage_tf32 = value != 0;
      if (!h->stage_tf32) mms_stage_drop_owner(h);
      return 0;
    case MMS_OPT_STAGE_ONLY: h->stage_only = value != 0; return 0;
    default: mms_set_error("unknown option %d", option); return MMS_E_INVALID;
  }
}

int mms_reserve_scratch(mms_handle_t h, long long bytes) {
  MMS_REQUIRE(h, MMS_E_INVALID, "null handle");
  MMS_REQUIRE(bytes >= 0, MMS_E_INVALID, "negative size");
  if ((size_t)bytes <= h->scratch_bytes) return 0;
  void* p = nullptr;
  return mms_scratch(h, (size_t)bytes, &p);
}

int mms_invalidate_caches(void) {
  mms_note_write(nullptr, 0);      // "everything may have changed"
  return 0;
}

int mms_get_option(mms_handle_t h, int option, long long* value) {
  MMS_REQUIRE(h && value, MMS_E_INVALID, "null argument");
  switch (option) {
    case MMS_OPT_MATH: *value = h->math; return 0;
    case MMS_OPT_PRL_GE: *value = h->prl_ge; return 0;
    case MMS_OPT_SCRATCH_BYTES: *value = (long long)h->scratch_cap; return 0;
    case MMS_OPT_EMBED_DETERMINISTIC: *value = h->embed_deterministic; return 0;
    case MMS_OPT_REUSE_FORWARD: *value = h->reuse_forward; return 0;
    case MMS_OPT_CONCURRENCY: *value = h->concurrency; return 0;
    case MMS_OPT_STAGE_TF32: *value = h->stage_tf32; return 0;
    case MMS_OPT_STAGE_ONLY: *value = h->stage_only; return 0;
    default: mms_set_error("unknown option %d", option); return MMS_E_INVALID;
  }
}

unsigned long long mms_launch_count(mms_handle_t h) { return h ? h->launches : 0; }

int mms_profile_enable(mms_handle_t h, int on) {
  MMS_REQUIRE(h, MMS_E_INVALID, "null handle");
  MMS_CUDA(cudaStreamSynchronize(h->stream));
  prof_clear(h);
  h->profile = on != 0;
  return 0;
}

// Writes one line per kernel name: "<name> <launches> <total_ms>\n"; clears the records.
int mms_profile_report(mms_handle_t h, char* buf, size_t size) {
  MMS_REQUIRE(h && buf && size > 0, MMS_E_INVALID, "null argument");
  MMS_CUDA(cudaStreamSynchronize(h->stream));
  std::map<std::string, std::pair<long long, double> > agg;
  std::vector<std::string> order;
  if (h->prof) {
    for (auto& r : *prof_of(h)) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
      if (!agg.count(r.name)) order.push_back(r.name);
      agg[r.name].first += 1;
      agg[r.name].second += ms;
    }
  }
  std::string out;
  char line[256];
  for (auto& n : order) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", n.c_str(), agg[n].first, agg[n].second);
    out += line;
  }
  snprintf(buf, size, "%s", out.c_str());
  prof_clear(h);
  return 0;
}

int mms_check_faults(mms_handle_t h) {
  MMS_REQUIRE(h, MMS_E_INVALID, "null handle");
  int flag = 0;
  MMS_CUDA(cudaMemcpyAsync(&flag, h->fault_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  MMS_CUDA(cudaStreamSynchronize(h->stream));
  if (flag) {
    MMS_CUDA(cudaMemsetAsync(h->fault_flag, 0, sizeof(int), h->stream));
    mms_set_error("Embed: an input index was outside [0, input_dim)");
    return MMS_E_FAULT;
  }
  return 0;
}

#define H MMS_REQUIRE(h, MMS_E_INVALID, "null handle")

int mms_embed_plan_pair_f32(mms_handle_t h, const float* idx0, long long M0, const float* idx1, long long M1, int V) {
  H; return mms_embed_plan_pair_impl(h, idx0, M0, idx1, M1, V);
}
int mms_embed_backward_pair_f32(mms_handle_t h, const float* idx0, const float* dtop0, long long M0, const float* idx1,
                                const float* dtop1, long long M1, float* dW, float* dbias, int D, int V) {
  H; return mms_embed_backward_pair_impl(h, idx0, dtop0, M0, idx1, dtop1, M1, dW, dbias, D, V);
}

#define MMS_DEFINE_TYPED(T, SUF)                                                                   \
  int mms_embed_forward_##SUF(mms_handle_t h, const T* idx, const T* W, const T* bias, T* top,     \
                              long long M, int D, int V) {                                         \
    H; mms_note_write(top, sizeof(T) * (size_t)(M > 0 ? M : 0) * (size_t)(D > 0 ? D : 0));         \
    return mms_embed_forward_impl<T>(h, idx, W, bias, top, M, D, V);                               \
  }                                                                                                \
  int mms_embed_backward_##SUF(mms_handle_t h, const T* idx, const T* dtop, T* dW, T* dbias,       \
                               long long M, int D, int V) {                                        \
    H; return mms_embed_backward_impl<T>(h, idx, dtop, dW, dbias, M, D, V);                        \
  }                                                                                                \
  int mms_simcross_forward_##SUF(mms_handle_t h, int mode, const T* q, const T* a, const T* Mw,    \
                                 const T* B, T* S, T* n0, T* n1, int N, int Lq, int La, int D,     \
                                 int mc) {                                                         \
    H; return mms_simcross_forward_impl<T>(h, mode, q, a, Mw, B, S, n0, n1, N, Lq, La, D, mc);     \
  }                                                                                                \
  int mms_simcross_backward_##SUF(mms_handle_t h, int mode, const T* q, const T* a, const T* Mw,   \
                                  const T* S, const T* dS, const T* n0, const T* n1, T* dq, T* da, \
                                  T* dM, T* dB, int N, int Lq, int La, int D, int mc, int p0,      \
                                  int p1) {                                                        \
    H; return mms_simcross_backward_impl<T>(h, mode, q, a, Mw, S, dS, n0, n1, dq, da, dM, dB, N,   \
                                            Lq, La, D, mc, p0, p1);                                \
  }                                                                                                \
  int mms_simmatrix_forward_##SUF(mms_handle_t h, const T* q, const T* a, const T* W, T* s, T* Tm, \
                                  int N, int K1, int K2) {                                         \
    H; return mms_simmatrix_forward_impl<T>(h, q, a, W, s, Tm, N, K1, K2);                         \
  }                                                                                                \
  int mms_simmatrix_backward_##SUF(mms_handle_t h, const T* q, const T* a, const T* W,             \
                                   const T* ds, T* dW, T* dq, T* da, int N, int K1, int K2,        \
                                   int pw, int p0, int p1) {                                       \
    H; return mms_simmatrix_backward_impl<T>(h, q, a, W, ds, dW, dq, da, N, K1, K2, pw, p0, p1);   \
  }                                                                                                \
  int mms_pairrankloss_forward_##SUF(mms_handle_t h, const T* a, const T* b, const T* y, T margin, \
                                     long long count, T* loss, T* ordered, T* similar) {           \
    H; return mms_pairrankloss_forward_impl<T>(h, a, b, y, margin, count, loss, ordered, similar); \
  }                                                                                                \
  int mms_pairrankloss_backward_##SUF(mms_handle_t h, const T* y, const T* ordered,                \
                                      const T* similar, T top_diff, long long count, T* da,        \
                                      T* db) {                                                     \
    H; return mms_pairrankloss_backward_impl<T>(h, y, ordered, similar, top_diff, count, da, db);  \
  }                                                                                                \
  int mms_fm_forward_##SUF(mms_handle_t h, const T* x, const T* bias, T* y, int N, int C, int Dm) {\
    H; return mms_fm_forward_impl<T>(h, x, bias, y, N, C, Dm);                                     \
  }                                                                                                \
  int mms_fm_backward_##SUF(mms_handle_t h, const T* x, const T* dy, T* dx, T* dbias, int N,       \
                            int C, int Dm, int prop0) {                                            \
    H; return mms_fm_backward_impl<T>(h, x, dy, dx, dbias, N, C, Dm, prop0);                       \
  }                                                                                                \
  int mms_dot_##SUF(mms_handle_t h, const T* data, const T* diff, long long count, T* out) {       \
    H; MMS_REQUIRE(diff, MMS_E_INVALID, "null pointer");                                           \
    return mms_dot_impl<T>(h, data, diff, count, out);                                             \
  }                                                                                                \
  int mms_scale_##SUF(mms_handle_t h, T* x, long long count, T alpha) {                            \
    H; return mms_scale_impl<T>(h, x, count, alpha);                                               \
  }                                                                                                \
  int mms_adadelta_step_##SUF(mms_handle_t h, T* data, T* diff, T* hist_g, T* hist_u,              \
                              long long count, T grad_scale, T local_decay, T momentum, T delta,   \
                              T local_rate, int clear_diff) {                                      \
    H; if (data) mms_note_write(data, sizeof(T) * (size_t)(count > 0 ? count : 0));                \
    return mms_adadelta_step_impl<T>(h, data, diff, hist_g, hist_u, count, grad_scale,             \
                                        local_decay, momentum, delta, local_rate, clear_diff);     \
  }                                                                                                \
  int mms_rank_map_mrr_##SUF(mms_handle_t h, const T* data, long long stride, long long offset,    \
                             const T* label, const T* group, long long count, T* map_out,          \
                             T* mrr_out) {                                                         \
    H; return mms_rank_map_mrr_impl<T>(h, data, stride, offset, label, group, count, map_out,      \
                                       mrr_out);                                                   \
  }                                                                                                \
  int mms_rank_auc_##SUF(mms_handle_t h, const T* data, long long stride, long long offset,        \
                         const T* label, long long count, int has_ignore_label, int ignore_label,  \
                         T* out) {                                                                 \
    H; return mms_rank_auc_impl<T>(h, data, stride, offset, label, count, has_ignore_label,        \
                                   ignore_label, out);                                             \
  }                                                                                                \
  int mms_rank_accuracy_##SUF(mms_handle_t h, const T* a, const T* b, const T* label,              \
                              long long count, T* out) {                                           \
    H; return mms_rank_accuracy_impl<T>(h, a, b, label, count, out);                               \
  }                                                                                                \
  int mms_sentconv_forward_##SUF(mms_handle_t h, const T* x, const T* W, const T* bias, T* top,    \
                                 int N, int L, int D, int C, int kh) {                             \
    H; return mms_sentconv_forward_impl<T>(h, x, W, bias, top, N, L, D, C, kh);                    \
  }                                                                                                \
  int mms_sentconv_backward_##SUF(mms_handle_t h, const T* x, const T* W, const T* dtop, T* dW,    \
                                  T* dbias, T* dx, int N, int L, int D, int C, int kh) {           \
    H; return mms_sentconv_backward_impl<T>(h, x, W, dtop, dW, dbias, dx, N, L, D, C, kh);         \
  }                                                                                                \
  int mms_conv2d_forward_##SUF(mms_handle_t h, const T* x, const T* W, const T* bias, T* top, int N,  \
                               int C, int H_, int W_, int Co, int kh, int kw) {                     \
    H; mms_note_write(top, sizeof(T) * (size_t)(N > 0 ? N : 0) * Co * (size_t)((H_ - kh + 1) > 0 ? (H_ - kh + 1) : 0) * \
                               (size_t)((W_ - kw + 1) > 0 ? (W_ - kw + 1) : 0));                   \
    return mms_conv2d_forward_impl<T>(h, x, W, bias, top, N, C, H_, W_, Co, kh, kw);                \
  }                                                                                                \
  int mms_conv2d_backward_##SUF(mms_handle_t h, const T* x, const T* W, const T* dtop, T* dW,       \
                                T* dbias, T* dx, int N, int C, int H_, int W_, int Co, int kh,      \
                                int kw) {                                                          \
    H; return mms_conv2d_backward_impl<T>(h, x, W, dtop, dW, dbias, dx, N, C, H_, W_, Co, kh, kw);  \
  }                                                                                                \
  int mms_dropout_##SUF(mms_handle_t h, const T* x, const unsigned* mask, T* y, long long count,    \
                        unsigned threshold, T scale) {                                             \
    H; mms_note_write(y, sizeof(T) * (size_t)(count > 0 ? count : 0));                             \
    return mms_dropout_impl<T>(h, x, mask, y, count, threshold, scale);                            \
  }                                                                                                \
  int mms_pool_forward_##SUF(mms_handle_t h, const T* x, T* top, int* mask, long long NC, int H_,  \
                             int W_, int PH, int PW, int kh, int kw, int sh, int sw, int pad_h,    \
                             int pad_w, int method) {                                              \
    H; mms_note_write(top, sizeof(T) * (size_t)(NC > 0 ? NC : 0) * (size_t)(PH > 0 ? PH : 0) * (size_t)(PW > 0 ? PW : 0)); \
    return mms_pool_forward_impl<T>(h, x, top, mask, NC, H_, W_, PH, PW, kh, kw, sh, sw, pad_h,    \
                                    pad_w, method);                                             \
  }                                                                                                \
  int mms_pool_backward_##SUF(mms_handle_t h, const T* dtop, const int* mask, T* dx, long long NC, \
                              int H_, int W_, int PH, int PW, int kh, int kw, int sh, int sw,      \
                              int pad_h, int pad_w, int method) {                                  \
    H; return mms_pool_backward_impl<T>(h, dtop, mask, dx, NC, H_, W_, PH, PW, kh, kw, sh, sw,     \
                                        pad_h, pad_w, method);                                     \
  }                                                                                                \
  int mms_tanh_forward_##SUF(mms_handle_t h, const T* x, T* y, long long count) {                  \
    H; mms_note_write(y, sizeof(T) * (size_t)(count > 0 ? count : 0));                             \
    return mms_tanh_forward_impl<T>(h, x, y, count);                                            \
  }                                                                                                \
  int mms_tanh_backward_##SUF(mms_handle_t h, const T* y, const T* dy, T* dx, long long count) {   \
    H; return mms_tanh_backward_impl<T>(h, y, dy, dx, count);                                      \
  }                                                                                                \
  int mms_bn_forward_##SUF(mms_handle_t h, const T* x, const T* scale, const T* shift,             \
                           T* run_mean, T* run_var, T* top, T* x_norm, T* batch_mean,              \
                           T* batch_std, int N, int C, int HW, int train, T bn_memory, T var_eps) {\
    H; mms_note_write(top, sizeof(T) * (size_t)(N > 0 ? N : 0) * (size_t)(C > 0 ? C : 0) * (size_t)(HW > 0 ? HW : 0)); \
    return mms_bn_forward_impl<T>(h, x, scale, shift, run_mean, run_var, top, x_norm,              \
                                     batch_mean, batch_std, N, C, HW, train, bn_memory, var_eps);  \
  }                                                                                                \
  int mms_bn_backward_##SUF(mms_handle_t h, const T* dtop, const T* x_norm, const T* scale,        \
                            const T* batch_std, T* dscale, T* dshift, T* dx, int N, int C,         \
                            int HW) {                                                              \
    H; return mms_bn_backward_impl<T>(h, dtop, x_norm, scale, batch_std, dscale, dshift, dx, N, C, \
                                      HW);                                                         \
  }                                                                                                \
  int mms_adadelta_update_##SUF(mms_handle_t h, T* g, T* hist_g, T* hist_u, long long count,       \
                                T momentum, T delta, T local_rate) {                               \
    H; return mms_adadelta_step_impl<T>(h, nullptr, g, hist_g, hist_u, count, T(1), T(0), momentum,\
                                        delta, local_rate, 0);                                     \
  }

MMS_DEFINE_TYPED(float, f32)
MMS_DEFINE_TYPED(double, f64)

int mms_simcross_prepare_f32(mms_handle_t h, const float* Mw, int D, int mc) {
  H;
  MMS_REQUIRE(Mw && D > 0 && mc > 0, MMS_E_INVALID, "bad argument");
  return mms_tc_simcross2_prepare(h, Mw, D, mc);
}
int mms_simcross_backward_bottoms_f32(mms_handle_t h, const float* q, const float* a, const float* Mw, const float* dS,
                                      float* dq, float* da, int N, int Lq, int La, int D, int mc) {
  H;
  MMS_REQUIRE(q && a && Mw && dS && dq && da, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N > 0 && Lq > 0 && La > 0 && D > 0 && mc > 0, MMS_E_INVALID, "bad size");
  if (h->math != MMS_MATH_TF32) return MMS_E_UNSUPPORTED;
  return mms_tc_simcross2_backward_bottoms(h, q, a, Mw, dS, dq, da, N, Lq, La, D, mc);
}
int mms_simcross_backward_params_f32(mms_handle_t h, const float* dS, float* dM, float* dB, int N, int Lq, int La, int D,
                                     int mc) {
  H;
  MMS_REQUIRE(dS && dM, MMS_E_INVALID, "null pointer");
  // dB depends on dS only: with MMS_OPT_CONCURRENCY it is reduced on a private stream beside the dM contraction
  const bool side_bias = dB && h->concurrency;
  if (side_bias) {
    MMS_TRY(mms_fork(h, 1));
    MmsStreamSwitch sw(h, 1);
    MMS_TRY(mms_simcross2_bias_grad<float>(h, dS, dB, N, Lq, La, mc));
  }
  const int rc = mms_tc_simcross2_backward_params(h, dM, N, Lq, La, D, mc);
  if (side_bias) MMS_TRY(mms_join(h, 1));
  MMS_TRY(rc);
  if (dB && !side_bias) MMS_TRY(mms_simcross2_bias_grad<float>(h, dS, dB, N, Lq, La, mc));
  return 0;
}

int mms_rerank_scores_f32(mms_handle_t h, const float* Q, const float* C, const float* W, float* QW,
                          float* scores, int Nq, long long Nc, int K1, int K2) {
  H; return mms_rerank_scores_impl(h, Q, C, W, QW, scores, Nq, Nc, K1, K2);
}

int mms_rerank_prepare_f32(mms_handle_t h, const float* C, float* C_tf32, long long Nc, int K2) {
  H; return mms_rerank_prepare_impl(h, C, C_tf32, Nc, K2);
}
int mms_rerank_scores_prepared_f32(mms_handle_t h, const float* Q, const float* C_tf32, const float* W, float* QW,
                                   float* scores, int Nq, long long Nc, int K1, int K2) {
  H; return mms_rerank_scores_prepared_impl(h, Q, C_tf32, W, QW, scores, Nq, Nc, K1, K2);
}

int mms_dropout_mask(mms_handle_t h, unsigned* mask, long long count, unsigned long long seed) {
  H; return mms_dropout_mask_impl(h, mask, count, seed);
}

int mms_rerank_topk_f32(mms_handle_t h, const float* Q, const float* C, const float* W, float* QW, float* top_scores,
                        long long* top_idx, int Nq, long long Nc, int K1, int K2, int k, long long idx_base) {
  H; return mms_rerank_topk_impl(h, Q, C, W, QW, top_scores, top_idx, Nq, Nc, K1, K2, k, idx_base, 0);
}
int mms_rerank_topk_prepared_f32(mms_handle_t h, const float* Q, const float* C_tf32, const float* W, float* QW,
                                 float* top_scores, long long* top_idx, int Nq, long long Nc, int K1, int K2, int k,
                                 long long idx_base) {
  H; return mms_rerank_topk_impl(h, Q, C_tf32, W, QW, top_scores, top_idx, Nq, Nc, K1, K2, k, idx_base, 1);
}
int mms_topk_merge_f32(mms_handle_t h, const float* scores, const long long* idx, long long ld, long long n,
                       float* out_scores, long long* out_idx, int Nq, int k) {
  H;
  MMS_REQUIRE(scores && idx && out_scores && out_idx, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(Nq > 0 && k > 0 && k <= 1024, MMS_E_INVALID, "bad size (k <= 1024)");
  MMS_TRY(mms_topk_init(h, out_scores, out_idx, Nq, k));
  return mms_topk_update(h, scores, idx, ld, n, 0, out_scores, out_idx, Nq, k);
}

// Test/diagnostic entry: C (+)= op(A) op(B) on the tcgen05 TF32 GEMM (see tc/tc_gemm.cuh).
int mms_tc_gemm_f32(mms_handle_t h, const float* A, long long lda, int a_mn, const float* B, long long ldb,
                    int b_mn, float* C, long long ldc, int M, int N, int K, int ksplit, int mode) {
  H;
  TcGemmArgs g = tc_gemm_args(A, lda, a_mn, B, ldb, b_mn, C, ldc, M, N, K, mode & 0xff);
  g.ksplit = ksplit;
  g.operands_tf32 = (mode & 0x100) != 0;
  return mms_tc_gemm(h, g);
}

}  // extern "C"
