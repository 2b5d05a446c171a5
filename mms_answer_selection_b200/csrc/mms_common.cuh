// Shared host/device plumbing for libmms_b200.so (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mms_b200.h"

struct mms_context {
  cudaStream_t stream = 0;
  int math = MMS_MATH_TF32;
  int prl_ge = 0;
  int embed_deterministic = 0;
  struct MPrepared {             // mms_simcross_prepare: M already rounded at the head of the scratch buffer
    bool valid = false; const void* M = nullptr; const void* at = nullptr; int D = 0, mc = 0;
    unsigned long long clock = 0;
  } m_prepared;
  void* embed_plan = nullptr;    // embed_sorted.cu: token rows grouped by id (mms_embed_plan_pair)
  size_t scratch_cap = (size_t)4 << 30;
  void* scratch = nullptr;       // grows on demand, reused across calls
  size_t scratch_bytes = 0;
  int* fault_flag = nullptr;     // device word set by kernels on data faults
  int sm_count = 148;
  int device = 0;
  void* tc_state = nullptr;      // tensor-core path private state (tensor maps etc.)
  unsigned long long launches = 0;   // kernels launched through this handle
  int profile = 0;               // 1: bracket every launch with CUDA events (mms_profile_*)
  void* prof = nullptr;          // std::vector<ProfRecord>*
  void* partials = nullptr;      // 1024 doubles: block partial sums of the loss reductions
  // fork/join inside one call: independent contractions of SimCross backward run on two private
  // streams, ordered against the caller's stream with events (works under stream capture too)
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork[2] = {nullptr, nullptr};
  cudaEvent_t ev_join[2] = {nullptr, nullptr};
  int concurrency = 1;           // MMS_OPT_CONCURRENCY
  int stage_only = 0;            // MMS_OPT_STAGE_ONLY: ... and not the fp32 top itself
  int stage_tf32 = 0;            // MMS_OPT_STAGE_TF32: Embed forward also writes the TF32-rounded, row-padded copy
  float* stage_buf = nullptr;    // ... into this buffer (owned by the handle, published in the staging registry)
  size_t stage_bytes = 0;
  // SimCross forward leaves its rounded operands and T = Q M_k in the scratch buffer; backward reuses
  // them when MMS_OPT_REUSE_FORWARD is set and nothing has touched the scratch buffer in between.
  int reuse_forward = 0;
  struct FwdCache {
    bool valid = false;
    const void *q = nullptr, *a = nullptr, *M = nullptr;
    int N = 0, Lq = 0, La = 0, D = 0, mc = 0;
    unsigned long long generation = 0;   // mms_write_clock() when the forward ran
    const float *qr = nullptr, *ar = nullptr;   // the rounded operands it used (workspace or a staged copy)
  } fwd_cache;
  // mms_simcross_backward_bottoms left U = dS A (and the rounded q) in the workspace for mms_simcross_backward_params
  struct DmPending {
    bool valid = false;
    int N = 0, Lq = 0, La = 0, D = 0, mc = 0;
    const float* qr = nullptr;                  // the rounded questions _bottoms used
  } dm_pending;
  // Sentence convolution: the forward leaves the TF32-rounded copy of x at the head of the scratch buffer; with
  // MMS_OPT_REUSE_FORWARD the backward on the same handle reads it instead of rounding x again.
  struct SentCache {
    bool valid = false;
    const void* x = nullptr;
    long long rows = 0;
    int D = 0;
    unsigned long long generation = 0;
  } sent_cache;
  // SimMatrix: likewise the rounded q and W ([qr | Wr] at the head of the scratch buffer)
  struct SimMatCache {
    bool valid = false;
    const void *q = nullptr, *W = nullptr;
    const void* T = nullptr;                   // where the forward left T = q W (the reference: bottom[1]'s diff)
    int N = 0, K1 = 0, K2 = 0;
    unsigned long long generation = 0;
  } simmat_cache;
};

// Process-wide write log behind the MMS_OPT_REUSE_FORWARD caches.  Every entry point that rewrites blob CONTENTS
// behind unchanged pointers -- the tops written by the forwards, the weights written by the optimizer step and the
// gradient exchange -- notes the byte range it writes (mms_note_write); mms_invalidate_caches() notes "everything".
// A cache records mms_write_clock() when its forward ran and is honoured by a later backward only if no range noted
// since then overlaps the operands it was built from (mms_unchanged_since): a weight update, an in-place layer or a
// refilled bottom between a Forward and a later Backward can therefore never be paired with stale rounded copies.
// Staging registry (MMS_OPT_STAGE_TF32): producer kernels that know their top will be a tensor-core operand write the
// rounded, padded copy themselves and publish it under the top's address; consumers look the address up.  An entry is
// honoured only while no logged write overlaps the top.
void mms_stage_publish(struct mms_context* owner, const void* src, const float* staged, long long rows, int cols, int ld, bool virt = false);
bool mms_stage_virtual(const void* src);   // src was published by a producer that did not write src itself (MMS_OPT_STAGE_ONLY)
int mms_stage_require_real(const void* src, bool have_staged, const char* what);
const float* mms_stage_lookup(const void* src, long long rows, int cols, int ld);
void mms_stage_drop_owner(struct mms_context* owner);
unsigned long long mms_write_clock();
void mms_note_write(const void* p, size_t bytes);
bool mms_unchanged_since(unsigned long long clock, const void* p, size_t bytes);

// Runs the launches issued between fork(i) and join(i) on private stream i, after everything already
// queued on the caller's stream; join(i) makes the caller's stream wait for them.
int mms_fork(mms_context* ctx, int i);
int mms_join(mms_context* ctx, int i);
struct MmsStreamSwitch {          // RAII: ctx->stream = side stream i for the scope
  mms_context* ctx; cudaStream_t saved;
  MmsStreamSwitch(mms_context* c, int i) : ctx(c), saved(c->stream) { c->stream = c->side[i]; }
  ~MmsStreamSwitch() { ctx->stream = saved; }
};

// Placed in front of every kernel launch: counts it and, when profiling is on, records a
// CUDA event on the handle's stream before and after the launch.
struct MmsKernelScope {
  mms_context* ctx;
  int slot;
  MmsKernelScope(mms_context* c, const char* name);
  ~MmsKernelScope();
};

void mms_set_error(const char* fmt, ...);

#define MMS_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      mms_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

#define MMS_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      mms_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

// Opt-in to `bytes` of dynamic shared memory AND pin the kernel's shared-memory carveout to the maximum.  Kernels whose
// carveouts differ cannot be resident on one SM at the same time (the SM drains before it is re-partitioned), which
// silently serialises the branches of the step graph -- the contraction kernels, the HBM-bound kernels beside them
// and the gradient exchange all ask for the same (largest) carveout, so they do co-reside (measured: the table
// exchange did not overlap the dM kernel at all before this).
#define MMS_MAX_SMEM(func, bytes)                                                                                  \
  do {                                                                                                             \
    MMS_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));                      \
    MMS_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
  } while (0)
// the same pin for kernels that need no opt-in (small static shared memory), once per kernel
int mms_prefer_max_shared(const void* func);
#define MMS_CARVEOUT(func) MMS_TRY(mms_prefer_max_shared(reinterpret_cast<const void*>(func)))

#define MMS_REQUIRE(cond, code, msg)                                                    \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      mms_set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);                  \
      return (code);                                                                    \
    }                                                                                   \
  } while (0)

#define MMS_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// Returns a scratch buffer of at least `bytes` (owned by the context).
int mms_scratch(mms_context* ctx, size_t bytes, void** out);

static inline int mms_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
// Developer kill switches (MMS_NO_FUSED, MMS_NO_2CTA, ...: force the generic engine instead of a dedicated kernel) are
// read from the environment only in builds with -DMMS_DEV_KNOBS (MMS_NVCC_EXTRA); the shipped library ignores them.
inline bool mms_dev_knob(const char* name) {
#ifdef MMS_DEV_KNOBS
  return getenv(name) != nullptr;
#else
  (void)name;
  return false;
#endif
}
template <typename T> __host__ __device__ inline T mms_min(T a, T b) { return a < b ? a : b; }
template <typename T> __host__ __device__ inline T mms_max(T a, T b) { return a > b ? a : b; }

// ---- SIMT batched GEMM (gemm_simt.cu): the fp32/fp64 fallback contraction ------------
// C[z][m][n] = alpha * sum_k A(z,m,k) * B(z,k,n) + beta * C[z][m][n]
// A(z,m,k) = A[z1*sA1 + z2*sA2 + m*sAm + k*sAk],  z = z1*nb2 + z2  (likewise B, C).
// ksplit > 1 splits K across CTAs and combines with atomicAdd (beta must be 1 then).
template <typename T>
struct SimtGemmArgs {
  const T* A; const T* B; T* C;
  int M, N, K;
  long long sAm, sAk, sBk, sBn;
  int ldc;
  long long sA1, sA2, sB1, sB2, sC1, sC2;
  int nb1, nb2;
  T alpha, beta;
  int ksplit;
};
template <typename T>
int mms_simt_gemm(mms_context* ctx, const SimtGemmArgs<T>& args);

template <typename T>
int mms_fill(mms_context* ctx, T* p, long long n, T v);

// ---- layer implementations (templated on the blob dtype) ------------------------------
template <typename T>
int mms_embed_forward_impl(mms_context*, const T* idx, const T* W, const T* bias, T* top,
                           long long M, int D, int V);
template <typename T>
int mms_embed_backward_impl(mms_context*, const T* idx, const T* dtop, T* dW, T* dbias,
                            long long M, int D, int V);
int mms_tc_simcross2_prepare(mms_context* ctx, const float* Mw, int D, int mc);
void mms_embed_plan_destroy(mms_context* ctx);
int mms_embed_plan_pair_impl(mms_context* ctx, const float* idx0, long long M0, const float* idx1, long long M1, int V);
int mms_embed_backward_pair_impl(mms_context* ctx, const float* idx0, const float* dtop0, long long M0, const float* idx1,
                                 const float* dtop1, long long M1, float* dW, float* dbias, int D, int V);
template <typename T>
int mms_embed_backward_deterministic(mms_context*, const T* idx, const T* dtop, T* dW, T* dbias, long long M, int D, int V);
template <typename T>
int mms_simcross_forward_impl(mms_context*, int mode, const T* q, const T* a, const T* Mw,
                              const T* B, T* S, T* norm0, T* norm1, int N, int Lq, int La, int D,
                              int mc);
template <typename T>
int mms_simcross_backward_impl(mms_context*, int mode, const T* q, const T* a, const T* Mw,
                               const T* S, const T* dS, const T* norm0, const T* norm1, T* dq,
                               T* da, T* dM, T* dB, int N, int Lq, int La, int D, int mc,
                               int prop0, int prop1);
template <typename T>
int mms_simmatrix_forward_impl(mms_context*, const T* q, const T* a, const T* W, T* s, T* Tm,
                               int N, int K1, int K2);
template <typename T>
int mms_simmatrix_backward_impl(mms_context*, const T* q, const T* a, const T* W, const T* ds,
                                T* dW, T* dq, T* da, int N, int K1, int K2, int prop_w,
                                int prop0, int prop1);
template <typename T>
int mms_pairrankloss_forward_impl(mms_context*, const T* a, const T* b, const T* y, T margin,
                                  long long count, T* loss, T* ordered, T* similar);
template <typename T>
int mms_pairrankloss_backward_impl(mms_context*, const T* y, const T* ordered, const T* similar,
                                   T top_diff, long long count, T* da, T* db);
template <typename T>
int mms_fm_forward_impl(mms_context*, const T* x, const T* bias, T* y, int N, int C, int Dm);
template <typename T>
int mms_fm_backward_impl(mms_context*, const T* x, const T* dy, T* dx, T* dbias, int N, int C,
                         int Dm, int prop0);
template <typename T>
int mms_dot_impl(mms_context*, const T* x, const T* y, long long n, T* out);
template <typename T>
int mms_scale_impl(mms_context*, T* x, long long n, T alpha);
template <typename T>
int mms_rank_map_mrr_impl(mms_context*, const T* data, long long stride, long long offset, const T* label, const T* group,
                          long long n, T* map_out, T* mrr_out);
template <typename T>
int mms_rank_auc_impl(mms_context*, const T* data, long long stride, long long offset, const T* label, long long n,
                      int has_ignore, int ignore_label, T* out);
template <typename T>
int mms_rank_accuracy_impl(mms_context*, const T* a, const T* b, const T* label, long long n, T* out);
int mms_rerank_topk_impl(mms_context*, const float* Q, const float* C, const float* W, float* QW, float* top_s,
                         long long* top_i, int Nq, long long Nc, int K1, int K2, int k, long long idx_base, int prepared);
int mms_topk_init(mms_context*, float* run_s, long long* run_i, int Nq, int k);
int mms_topk_update(mms_context*, const float* scores, const long long* idx, long long ld, long long n, long long idx_base,
                    float* run_s, long long* run_i, int Nq, int k);
int mms_rerank_prepare_impl(mms_context*, const float* C, float* Cr, long long Nc, int K2);
int mms_rerank_scores_prepared_impl(mms_context*, const float* Q, const float* Cr, const float* W, float* QW,
                                    float* scores, int Nq, long long Nc, int K1, int K2);
template <typename T>
int mms_sentconv_forward_impl(mms_context*, const T* x, const T* W, const T* bias, T* top, int N, int L, int D, int C, int kh);
template <typename T>
int mms_sentconv_backward_impl(mms_context*, const T* x, const T* W, const T* dtop, T* dW, T* dbias, T* dx, int N, int L,
                               int D, int C, int kh);
template <typename T>
int mms_conv2d_forward_impl(mms_context*, const T* x, const T* W, const T* bias, T* top, int N, int C, int H, int Wd, int Co,
                            int kh, int kw);
template <typename T>
int mms_conv2d_backward_impl(mms_context*, const T* x, const T* W, const T* dtop, T* dW, T* dbias, T* dx, int N, int C, int H,
                             int Wd, int Co, int kh, int kw);
template <typename T>
int mms_dropout_impl(mms_context*, const T* x, const unsigned* mask, T* y, long long n, unsigned threshold, T scale);
int mms_dropout_mask_impl(mms_context*, unsigned* mask, long long n, unsigned long long seed);
template <typename T>
int mms_pool_forward_impl(mms_context*, const T* x, T* top, int* mask, long long NC, int H, int W, int PH, int PW, int kh,
                          int kw, int sh, int sw, int pad_h, int pad_w, int method);
template <typename T>
int mms_pool_backward_impl(mms_context*, const T* dtop, const int* mask, T* dx, long long NC, int H, int W, int PH, int PW,
                           int kh, int kw, int sh, int sw, int pad_h, int pad_w, int method);
template <typename T>
int mms_tanh_forward_impl(mms_context*, const T* x, T* y, long long n);
template <typename T>
int mms_tanh_backward_impl(mms_context*, const T* y, const T* dy, T* dx, long long n);
template <typename T>
int mms_bn_forward_impl(mms_context*, const T* x, const T* scale, const T* shift, T* run_mean, T* run_var, T* top,
                        T* x_norm, T* batch_mean, T* batch_std, int N, int C, int HW, int train, T memory, T eps);
template <typename T>
int mms_bn_backward_impl(mms_context*, const T* dtop, const T* x_norm, const T* scale, const T* batch_std, T* dscale,
                         T* dshift, T* dx, int N, int C, int HW);
template <typename T>
int mms_adadelta_step_impl(mms_context*, T* data, T* diff, T* hist_g, T* hist_u, long long n, T grad_scale, T local_decay,
                           T momentum, T delta, T local_rate, int clear_diff);

int mms_rerank_scores_impl(mms_context*, const float* Q, const float* C, const float* W, float* QW,
                           float* scores, int Nq, long long Nc, int K1, int K2);

// ---- tensor-core (tcgen05, TF32) path, float only (tc/*.cu) ---------------------------
// Each returns MMS_E_UNSUPPORTED when the shape is outside what the tcgen05 kernels
// handle; the caller then uses the SIMT path (still on the GPU).
int mms_tc_simcross2_forward(mms_context*, const float* q, const float* a, const float* Mw,
                             const float* B, float* S, int N, int Lq, int La, int D, int mc);
// fused forward over TF32-rounded, row-padded copies (tc/simcross_fused.cu)
int mms_tc_simcross2_forward_fused(mms_context*, const float* qr, const float* ar, const float* Mr,
                                   const float* B, float* S, int N, int Lq, int La, int D, int mc, int Dp);
// fused bottom gradients (tc/simcross_fused_bwd.cu): which = 0 dq (xr = rounded answers; exports U = G A for the
// dM contraction), which = 1 da (xr = rounded questions).  _plan tells whether the shape is covered and how many
// of the `ctas` CTAs it may use share one tile's measures (> 1: `out` must be zeroed by the caller).
int mms_tc_simcross2_backward_fused_plan(int which, int N, int Lq, int La, int D, int mc, int ctas, int allow_split,
                                         int* ksplit, int* split_from);
int mms_tc_simcross2_backward_fused(mms_context*, int which, const float* xr, const float* Mr, const float* dS,
                                    float* out, float* Uexp, int N, int Lq, int La, int D, int mc, int Dp, int ksplit,
                                    int split_from, int u_blocked = 0);
// weight gradient dM_k += Qall^T U_k from the blocked U export (tc/simcross_dm.cu); _plan: shape covered?
int mms_tc_simcross2_dm_plan(int D);
int mms_tc_simcross2_dm(mms_context*, const float* qr, const float* Ub, float* dM, long long rows, int D, int Dp, int mc,
                        int blocked = 1);
// (dq, da, dM overwritten; dB is accumulated by the caller)
int mms_tc_simcross2_backward(mms_context*, const float* q, const float* a, const float* Mw,
                              const float* dS, float* dq, float* da, float* dM, int N, int Lq, int La,
                              int D, int mc);
// the same in two calls: bottoms (dq, da; U stays in the workspace) and, later, the weight gradient from it
int mms_tc_simcross2_backward_bottoms(mms_context*, const float* q, const float* a, const float* Mw, const float* dS,
                                      float* dq, float* da, int N, int Lq, int La, int D, int mc);
int mms_tc_simcross2_backward_params(mms_context*, float* dM, int N, int Lq, int La, int D, int mc);
template <typename T>
int mms_simcross2_bias_grad(mms_context*, const T* dS, T* dB, int N, int Lq, int La, int mc);
