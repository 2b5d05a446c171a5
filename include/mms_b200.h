/*
 * mms_b200.h -- C-ABI of the B200-native MMS hot path (libmms_b200.so).
 *
 * Drop-in boundary: every entry point below is what a Caffe Layer subclass's
 * Forward_gpu / Backward_gpu for this path binds to (see INTEGRATION.md and
 * mms_answer_selection_b200/caffe_layers/).  Each declaration cites the reference
 * routine it replaces (paths relative to the reference tree).
 *
 * Conventions
 *  - plain C, plain pointers and sizes; no C++/torch types cross this boundary;
 *  - all data pointers are DEVICE pointers (cudaMalloc'ed, e.g. Blob::gpu_data())
 *    unless the parameter name ends in _host;
 *  - tensors are dense row-major, exactly Caffe's Blob layout;
 *  - <name>_f32 computes on float blobs, <name>_f64 on double blobs
 *    (Caffe instantiates every layer for both, common.hpp:41-66);
 *  - work is enqueued on the handle's stream (default: the legacy default stream 0,
 *    which is what Caffe uses); calls do not synchronise unless
 *    stated;
 *  - every function returns 0 on success.  A positive value is a cudaError_t, a
 *    negative value one of MMS_E_* below; mms_last_error() gives a message for the
 *    calling thread.  There is NO CPU fallback: without a CUDA device every compute
 *    entry point fails with a CUDA error.
 *  - "ACCUMULATES" / "OVERWRITES" state the diff semantics, which follow the
 *    reference layer by layer (SURVEY.md 8(a)/8(b) "Diff conventions").
 */
#ifndef MMS_B200_H_
#define MMS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMS_B200_VERSION 100 /* 0.1.0 */

enum {
  MMS_E_INVALID = -1,     /* bad argument (null pointer, non-positive size, bad mode) */
  MMS_E_UNSUPPORTED = -2, /* shape outside what the kernels implement */
  MMS_E_NOMEM = -3,       /* workspace allocation failed */
  MMS_E_FAULT = -4        /* a kernel flagged a data fault (e.g. Embed index out of range) */
};

/* Arithmetic used by the float contractions (SimCross mode 2, SimMatrix, rerank). */
enum {
  MMS_MATH_TF32 = 0, /* default: tcgen05 kind::tf32, fp32 accumulate in TMEM */
  MMS_MATH_FP32 = 1  /* SIMT FFMA path, fp32 throughout (also the shape fallback) */
};

enum {
  MMS_OPT_MATH = 1,          /* MMS_MATH_* */
  MMS_OPT_PRL_GE = 2,        /* PairRankLoss backward hinge test: 0 `ordered > 0` (reference CPU,
                                pair_rank_loss_layer.cpp:76), 1 `ordered >= 0` (reference GPU,
                                pair_rank_loss_layer.cu:51).  Default 0. */
  MMS_OPT_SCRATCH_BYTES = 3, /* cap for the per-call scratch chunk (default 4 GiB; grows on demand) */
  MMS_OPT_EMBED_DETERMINISTIC = 4, /* Embed backward: 1 = order-independent reduction: rows are sorted by id and every
                                table row is summed in 64-bit fixed point by one writer, so dW / dbias are bit-identical
                                from run to run and under any permutation of the token rows (needs D a multiple of
                                16 bytes and 16-byte aligned blobs, else MMS_E_UNSUPPORTED); 0 = run-merged float
                                atomics (arrival order, like the reference's kernel).  On a handle used for
                                mms_simcross_backward (mode 2): 1 = every dq / da row has one writer (the measures of a
                                pair group are never spread over CTAs that add with float atomics), so the bottom
                                gradients are bit-identical from run to run too.  Default 0. */
  MMS_OPT_REUSE_FORWARD = 5, /* SimCross mode 2, float: 1 = mms_simcross_backward may reuse the TF32-rounded
                                operands and T = Q M_k that the LAST mms_simcross_forward on this handle left
                                in the workspace, provided it was called with the same q / a / M pointers and
                                sizes and no other call used the workspace since.  The caller promises that
                                q, a and M are unchanged between that forward and this backward -- which is
                                how Net::ForwardBackward and GradientChecker drive a layer; writes made through
                                this library (optimizer step, gradient exchange, forwards refilling a bottom)
                                are detected and drop the cache, other in-place writes need
                                mms_invalidate_caches().  Default 0
                                (stateless: backward recomputes, like the reference, sim_cross_layer.cpp:296).
                                The same promise lets mms_simmatrix_backward reuse the rounded q and W, and
                                mms_sentconv_backward the rounded x, of the last forward on the handle. */
  MMS_OPT_CONCURRENCY = 6,   /* 1 (default): independent contractions inside one call may run on private
                                streams, joined back before the call's last launch on the handle's stream. */
  MMS_OPT_STAGE_TF32 = 7,    /* Embed forward, float: 1 = the gather also writes the operand copy the tensor-core
                                contractions read (TF32 round-to-nearest, rows padded to 128-byte lines) into a
                                buffer owned by this handle (ONE buffer per handle: a second Embed forward on the same
                                handle replaces the copy of the first -- give every Embed layer its own handle), and
                                publishes it under the top's address.  A later
                                mms_simcross_forward / _backward whose q or a IS that top (same pointer, shape) reads
                                the staged copy instead of re-reading and rounding the top in a pass of its own
                                (q and a then cross HBM twice per forward instead of three times).  The copy is
                                dropped when the library itself rewrites the top; a caller that lets a foreign
                                in-place layer modify the top between this Embed and its consumer must not set the
                                option (or call mms_invalidate_caches).  Default 0. */
  MMS_OPT_STAGE_ONLY = 8     /* Embed forward with MMS_OPT_STAGE_TF32: 1 = the staged copy is the ONLY thing written; the
                                fp32 top keeps its address (it is the key the copy is published under) but its memory is
                                left untouched, which takes a third off the gather's HBM traffic.  For tops whose only
                                consumers are mms_simcross_forward / _backward mode 2 on the tensor-core path: those
                                read the staged copy, and fail with MMS_E_INVALID -- rather than read memory nobody
                                wrote -- when they cannot (other modes, double, MMS_MATH_FP32, a batch cut into
                                chunks, a stale copy).  Anything else that reads the top sees whatever the buffer
                                held before.  Default 0. */
};

typedef struct mms_context* mms_handle_t;

/* Opaque per-layer workspace: stream, scratch, options.  Create in LayerSetUp,
 * destroy in the layer destructor; one handle is used by one host thread at a time
 * (Caffe enters a layer instance from one thread, layer.hpp:451-487). */
int mms_create(mms_handle_t* out);
int mms_destroy(mms_handle_t h);
int mms_set_stream(mms_handle_t h, void* cuda_stream /* cudaStream_t */);
int mms_set_option(mms_handle_t h, int option, long long value);
int mms_get_option(mms_handle_t h, int option, long long* value);
/* Grows the handle's workspace to at least `bytes` now.  The workspace otherwise grows on demand with
 * cudaStreamSynchronize + cudaFree + cudaMalloc, which is illegal while the handle's stream is being captured into a
 * CUDA graph: a call that would have to grow it during a capture fails with MMS_E_INVALID instead of invalidating
 * the capture.  Run the step once eagerly, or reserve here, before capturing. */
int mms_reserve_scratch(mms_handle_t h, long long bytes);
/* MMS_OPT_REUSE_FORWARD bookkeeping.  The library logs the byte ranges its own entry points rewrite behind
 * unchanged pointers (tops of the forwards, weights in the optimizer step and the gradient exchange) and a backward
 * only reuses what its forward left in the workspace if none of them overlaps the forward's operands.  A caller that
 * changes q / a / weights IN PLACE by other means between a Forward and a later Backward (pycaffe assigning
 * net.params[...], a cudaMemcpy into a blob, a foreign in-place layer) calls this to drop every such cache. */
int mms_invalidate_caches(void);
/* Synchronises the handle's stream and reports (then clears) kernel-flagged data
 * faults: returns MMS_E_FAULT if any Embed index was outside [0, V). */
int mms_check_faults(mms_handle_t h);
/* Number of kernels launched through this handle so far (bench.py's gpu_launches). */
unsigned long long mms_launch_count(mms_handle_t h);
/* Per-launch device timing with CUDA events on the handle's stream (used by bench.py for
 * the roofline line; never on by default).  mms_profile_report writes one line per kernel
 * name, "<name> <launches> <total_ms>", synchronises the stream and clears the records. */
int mms_profile_enable(mms_handle_t h, int on);
int mms_profile_report(mms_handle_t h, char* buf, size_t size);
const char* mms_last_error(void);
int mms_version(void);
/* 1 if the library can run here (CUDA device of compute capability 10.x present). */
int mms_device_ok(void);

/* ------------------------------------------------------------------ Embed --
 * Replaces EmbedLayer::Forward_gpu (src/caffe/layers/embed_layer.cu:42-56):
 * top[n,:] = W[(int)idx[n],:] (+ bias).  idx holds M float-encoded row ids.
 * Bit-exact.  bias may be NULL (bias_term: false). */
int mms_embed_forward_f32(mms_handle_t h, const float* idx, const float* W, const float* bias,
                          float* top, long long M, int D, int V);
int mms_embed_forward_f64(mms_handle_t h, const double* idx, const double* W, const double* bias,
                          double* top, long long M, int D, int V);
/* Replaces EmbedLayer::Backward_gpu (embed_layer.cu:59-77): dW[idx[n],:] += dtop[n,:],
 * dbias += sum_n dtop[n,:].  ACCUMULATES into dW/dbias (the solver zeroes them).
 * dW or dbias may be NULL (param_propagate_down_ false / no bias). */
int mms_embed_backward_f32(mms_handle_t h, const float* idx, const float* dtop, float* dW,
                           float* dbias, long long M, int D, int V);
int mms_embed_backward_f64(mms_handle_t h, const double* idx, const double* dtop, double* dW,
                           double* dbias, long long M, int D, int V);
/* The Backward_gpu of TWO Embed layers that share their table (param sharing by name, do_trec_qa_clean.py:461-468) as one
 * scatter-add: dW[idx0[n],:] += dtop0[n,:], dW[idx1[n],:] += dtop1[n,:], dbias += both column sums (embed_layer.cu:59-77
 * twice).  The token rows of both blobs are grouped by id first, so a table row that occurs k times receives one atomic
 * add per float instead of k (and its 1200 bytes of dW cross HBM once instead of up to k times: the table gradient is
 * never L2-resident when the scatter runs).  Same accumulate semantics and the same float-atomic (arrival-order) rounding as
 * mms_embed_backward; idx1 / dtop1 may be NULL with M1 = 0.
 * mms_embed_plan_pair does the grouping alone -- it needs the ids only, so a net can run it beside the forward pass; a
 * following mms_embed_backward_pair on the same handle with the same id blobs and sizes uses it (the caller vouches that
 * the ids did not change in between), any other call groups the rows itself.  Shapes the grouped kernels do not take
 * (D % 4 != 0, D > 512, unaligned blobs), batches below 32 768 token rows (launch-bound: two launches beat six) and
 * handles with MMS_OPT_EMBED_DETERMINISTIC run mms_embed_backward per blob. */
/* Optional, SimCross mode 2, float, MMS_MATH_TF32: rounds M to TF32 into the handle's workspace AHEAD of the next
 * mms_simcross_forward on this handle (which then skips that pass when M is the same pointer and unchanged): M does not
 * depend on the step's inputs, so a net runs this on a side stream beside the Embed gathers.  No reference counterpart
 * (sim_cross_layer.cpp:146-160 reads M in place).  A no-op until the workspace exists (first forward). */
int mms_simcross_prepare_f32(mms_handle_t h, const float* M, int D, int mc);
int mms_embed_plan_pair_f32(mms_handle_t h, const float* idx0, long long M0, const float* idx1, long long M1, int V);
int mms_embed_backward_pair_f32(mms_handle_t h, const float* idx0, const float* dtop0, long long M0,
                                const float* idx1, const float* dtop1, long long M1, float* dW, float* dbias,
                                int D, int V);

/* --------------------------------------------------------------- SimCross --
 * Replaces SimCrossLayer::Forward_gpu (src/caffe/layers/sim_cross_layer.cu:127-190,
 * whose mode 2 falls back to Forward_cpu, sim_cross_layer.cpp:140-161).
 *   q (N,Lq,D), a (N,La,D); mode 0 cosine, 1 1/(1+||q-a||), 2 bilinear.
 *   mode 2: Mw (mc,D,D), B (mc,Lq,La) or NULL; S (N,mc,Lq,La) = Q M_k A^T + B_k.
 *   modes 0/1: S (N,1,Lq,La); Mw/B ignored; norm0 (N,Lq) / norm1 (N,La) are the
 *   mode-0 row-norm caches (data0_norm_/data1_norm_), NULL otherwise. */
int mms_simcross_forward_f32(mms_handle_t h, int mode, const float* q, const float* a,
                             const float* Mw, const float* B, float* S, float* norm0, float* norm1,
                             int N, int Lq, int La, int D, int mc);
int mms_simcross_forward_f64(mms_handle_t h, int mode, const double* q, const double* a,
                             const double* Mw, const double* B, double* S, double* norm0,
                             double* norm1, int N, int Lq, int La, int D, int mc);
/* Replaces SimCrossLayer::Backward_gpu (sim_cross_layer.cu:193-243 -> Backward_cpu,
 * sim_cross_layer.cpp:166-307).  OVERWRITES dq and da (zeroed first, :176-177, even when
 * nothing propagates); the rest runs only if prop0 || prop1 (:201).  Mode 2: OVERWRITES dM
 * (zeroed, :256) and ACCUMULATES into dB (:301-304; NULL when bias_term is false).
 * S is the forward output (read by modes 0/1 only). */
int mms_simcross_backward_f32(mms_handle_t h, int mode, const float* q, const float* a,
                              const float* Mw, const float* S, const float* dS, const float* norm0,
                              const float* norm1, float* dq, float* da, float* dM, float* dB,
                              int N, int Lq, int La, int D, int mc, int prop0, int prop1);
int mms_simcross_backward_f64(mms_handle_t h, int mode, const double* q, const double* a,
                              const double* Mw, const double* S, const double* dS,
                              const double* norm0, const double* norm1, double* dq, double* da,
                              double* dM, double* dB, int N, int Lq, int La, int D, int mc,
                              int prop0, int prop1);

/* The same mode-2 backward in two calls, for a host that orders the step by data dependencies rather than layer by
 * layer (data-parallel training: the embedding scatter-add and the exchange of the table gradient only need dq / da,
 * so they can run while the weight gradient dM is still being computed):
 *   _bottoms  dq, da OVERWRITTEN (sim_cross_layer.cpp:291-299); U = dS A stays in the handle's workspace.  Float /
 *             TF32 only; MMS_E_UNSUPPORTED when the shape is outside the fused kernels or the batch does not fit
 *             the workspace in one piece -- call mms_simcross_backward_f32 then.
 *   _params   dM OVERWRITTEN (:256, :286-289) from that workspace -- it must be the next call on the handle after
 *             the matching _bottoms -- and dB ACCUMULATES (:301-304; NULL without bias_term). */
int mms_simcross_backward_bottoms_f32(mms_handle_t h, const float* q, const float* a, const float* Mw,
                                      const float* dS, float* dq, float* da, int N, int Lq, int La, int D,
                                      int mc);
int mms_simcross_backward_params_f32(mms_handle_t h, const float* dS, float* dM, float* dB, int N, int Lq,
                                     int La, int D, int mc);

/* -------------------------------------------------------------- SimMatrix --
 * Replaces SimMatrixLayer::Forward_gpu (src/caffe/layers/sim_matrix_layer.cu:20-40):
 * T = q W (N,K2), written to `T` -- the reference keeps it in bottom[1]'s diff buffer
 * (sim_matrix_layer.cpp:58) -- and s_n = <a_n, T_n>.  q (N,K1), a (N,K2), W (K1,K2), s (N). */
int mms_simmatrix_forward_f32(mms_handle_t h, const float* q, const float* a, const float* W,
                              float* s, float* T, int N, int K1, int K2);
int mms_simmatrix_forward_f64(mms_handle_t h, const double* q, const double* a, const double* W,
                              double* s, double* T, int N, int K1, int K2);
/* Replaces SimMatrixLayer::Backward_gpu (= Backward_cpu, sim_matrix_layer.cpp:68-95):
 * dW += sum_n ds_n q_n a_n^T (ACCUMULATES, if prop_w); dq_n = ds_n W a_n (OVERWRITES, if
 * prop0); da_n = ds_n W^T q_n (OVERWRITES, if prop1). */
int mms_simmatrix_backward_f32(mms_handle_t h, const float* q, const float* a, const float* W,
                               const float* ds, float* dW, float* dq, float* da, int N, int K1,
                               int K2, int prop_w, int prop0, int prop1);
int mms_simmatrix_backward_f64(mms_handle_t h, const double* q, const double* a, const double* W,
                               const double* ds, double* dW, double* dq, double* da, int N, int K1,
                               int K2, int prop_w, int prop0, int prop1);

/* ----------------------------------------------------------- PairRankLoss --
 * Replaces PairRankLossLayer::Forward_gpu (src/caffe/layers/pair_rank_loss_layer.cu:11-43):
 * ordered = margin - y (a-b), similar = a-b (both cached, `count` elements each),
 * *loss = (1/count) sum[max(0,ordered) + |(1-y) similar|].  loss is a DEVICE scalar. */
int mms_pairrankloss_forward_f32(mms_handle_t h, const float* a, const float* b, const float* y,
                                 float margin, long long count, float* loss, float* ordered,
                                 float* similar);
int mms_pairrankloss_forward_f64(mms_handle_t h, const double* a, const double* b, const double* y,
                                 double margin, long long count, double* loss, double* ordered,
                                 double* similar);
/* Replaces PairRankLossLayer::Backward_gpu (pair_rank_loss_layer.cu:46-82):
 * d{a,b}[i] = sign * (top_diff/count) * ([ordered>0] y - sgn((1-y) similar) (1-y)), sign -1 for a,
 * +1 for b, sgn(0) = -1.  OVERWRITES; da / db may be NULL (propagate_down false).
 * top_diff_host is top[0]->cpu_diff()[0] (the loss weight), a host scalar. */
int mms_pairrankloss_backward_f32(mms_handle_t h, const float* y, const float* ordered,
                                  const float* similar, float top_diff_host, long long count,
                                  float* da, float* db);
int mms_pairrankloss_backward_f64(mms_handle_t h, const double* y, const double* ordered,
                                  const double* similar, double top_diff_host, long long count,
                                  double* da, double* db);

/* --------------------------------------------------------------------- FM --
 * Replaces FMLayer::Forward_gpu (src/caffe/layers/fm_layer.cu:12-15 -> Forward_cpu,
 * fm_layer.cpp:33-62): x (N,C,Dm), y_n = 1/2 sum_{j>=1}[(sum_k x_kj)^2 - sum_k x_kj^2]
 * + sum_k x_k0 + bias.  bias: device scalar or NULL. */
int mms_fm_forward_f32(mms_handle_t h, const float* x, const float* bias, float* y, int N, int C,
                       int Dm);
int mms_fm_forward_f64(mms_handle_t h, const double* x, const double* bias, double* y, int N, int C,
                       int Dm);
/* Replaces FMLayer::Backward_gpu (fm_layer.cu:18-21 -> Backward_cpu, fm_layer.cpp:65-99):
 * dbias = sum dy (OVERWRITES, :77; NULL to skip); dx OVERWRITTEN if prop0. */
int mms_fm_backward_f32(mms_handle_t h, const float* x, const float* dy, float* dx, float* dbias,
                        int N, int C, int Dm, int prop0);
int mms_fm_backward_f64(mms_handle_t h, const double* x, const double* dy, double* dx,
                        double* dbias, int N, int C, int Dm, int prop0);

/* ------------------------------------------------------- loss plumbing ------
 * Replaces the loss reduction in Layer::Forward (include/caffe/layer.hpp:471-479,
 * caffe_gpu_dot(count, top.data, top.diff)): *out = sum_i data[i]*diff[i], device scalar. */
int mms_dot_f32(mms_handle_t h, const float* data, const float* diff, long long count, float* out);
int mms_dot_f64(mms_handle_t h, const double* data, const double* diff, long long count,
                double* out);

/* ------------------------------------------------- data-parallel exchange ---
 * Replaces the root's 1/solver_count scaling of the summed gradient buffer in
 * P2PSync::on_gradients_ready (src/caffe/parallel.cpp:377, caffe_gpu_scal): x *= alpha. */
int mms_scale_f32(mms_handle_t h, float* x, long long count, float alpha);
int mms_scale_f64(mms_handle_t h, double* x, long long count, double alpha);

/* The exchange itself: replaces Params / GPUParams (parallel.cpp:60-115: all learnable blobs re-bound to one flat
 * data and one flat diff buffer per solver) and P2PSync::on_start / on_gradients_ready (parallel.cpp:287-322 weight
 * broadcast down the tree, :325-380 gradient sum up the tree + 1/solver_count scale on the root), one rank per GPU
 * -- a process (torchrun) or a thread (P2PSync::InternalThreadEntry, :271-284).  Every rank owns ONE allocation
 * [flags | data | diff] that all peers map over NVLink; one kernel launch per rank exchanges a range of the flat
 * buffer by reading the rank's 1/world slice from every peer (fixed order: identical bits on every rank), and
 * writing the result into every peer -- with an NVSwitch multicast mapping as multimem.ld_reduce / multimem.st.
 * Calls are stream-ordered and capturable into CUDA graphs; cross-rank waits are bounded
 * (MMS_EXCHANGE_OPT_TIMEOUT_MS) and reported by mms_exchange_check, never a hang.
 *
 *   mms_exchange_bytes        size of one rank's allocation for `count` elements of 4 / 8 bytes
 *   mms_exchange_create       external_base NULL: the library cudaMalloc's the allocation (exportable by IPC);
 *                             otherwise the caller's allocation of mms_exchange_bytes() (e.g. symmetric memory)
 *   mms_exchange_buffers      this rank's flat data / diff buffers (count elements each) to re-bind the blobs to;
 *                             blob offsets inside them must be multiples of 16 bytes
 *   mms_exchange_export_ipc / _attach_ipc   one process per GPU: every rank exports 64 bytes, the host gathers the
 *                             world x 64 bytes by any means (torch.distributed, MPI, a file) and attaches
 *   mms_exchange_attach_ptrs  peers' allocations as plain pointers (threads of one process -- peer access is
 *                             enabled here -- or symmetric memory), plus the multicast address of the same
 *                             allocations or NULL
 *   mms_exchange_allreduce    diff[begin, end) <- scale * sum over ranks, on every rank          (:325-380, :377)
 *   mms_exchange_adadelta     the same sum, then the owner of each slice applies SGDSolver::ApplyUpdate with the
 *                             AdaDelta rule (see mms_adadelta_step; history sharded over the ranks) and stores the
 *                             new weights into every rank's data: all-reduce + replicated optimizer + the next
 *                             on_start broadcast in one pass.  seg_end[s] (element offsets, ascending, last = end of
 *                             the buffer) delimits blobs with local_rate seg_rate[s] and local_decay seg_decay[s].
 *                             clear_diff: this rank's gradient range is left zeroed (Net::ClearParamDiffs).
 *   mms_exchange_broadcast    data <- root's data on every rank                                   (:287-322)
 * `channel` (0..MMS_EXCHANGE_CHANNELS-1) selects an independent set of flags: calls that may overlap in time (two
 * buckets on two streams) use different channels; every rank issues the same calls in the same order per channel. */
#define MMS_EXCHANGE_MAX_WORLD 8
#define MMS_EXCHANGE_CHANNELS 4
#define MMS_EXCHANGE_IPC_BYTES 64
enum {
  MMS_EXCHANGE_OPT_CTAS = 1,       /* CTAs per exchange kernel (0 = one per SM) */
  MMS_EXCHANGE_OPT_TIMEOUT_MS = 2, /* bound of every cross-rank wait (default 10000) */
  MMS_EXCHANGE_OPT_MULTICAST = 3   /* 0: ignore the multicast mapping (plain peer loads / stores) */
};
typedef struct mms_exchange* mms_exchange_t;
long long mms_exchange_bytes(long long count, int elem_bytes);
int mms_exchange_create(mms_exchange_t* out, int rank, int world, long long count, int elem_bytes,
                        void* external_base);
int mms_exchange_destroy(mms_exchange_t x);
int mms_exchange_buffers(mms_exchange_t x, void** data, void** diff);
int mms_exchange_base(mms_exchange_t x, void** base);
int mms_exchange_export_ipc(mms_exchange_t x, void* handle64);
int mms_exchange_attach_ipc(mms_exchange_t x, const void* handles /* world x 64 bytes, rank order */);
int mms_exchange_attach_ptrs(mms_exchange_t x, void* const* peer_bases, void* multicast_base);
int mms_exchange_set_option(mms_exchange_t x, int option, long long value);
int mms_exchange_allreduce_f32(mms_exchange_t x, void* cuda_stream, int channel, long long begin, long long end,
                               float scale);
int mms_exchange_allreduce_f64(mms_exchange_t x, void* cuda_stream, int channel, long long begin, long long end,
                               double scale);
int mms_exchange_adadelta_f32(mms_exchange_t x, void* cuda_stream, int channel, long long begin, long long end,
                              float grad_scale, const long long* seg_end, const double* seg_rate,
                              const double* seg_decay, int nseg, float momentum, float delta, int clear_diff);
int mms_exchange_adadelta_f64(mms_exchange_t x, void* cuda_stream, int channel, long long begin, long long end,
                              double grad_scale, const long long* seg_end, const double* seg_rate,
                              const double* seg_decay, int nseg, double momentum, double delta, int clear_diff);
int mms_exchange_broadcast(mms_exchange_t x, void* cuda_stream, int channel, int root);
/* this rank's history arrays of the fused solver step (count elements each; NULL before the first step) */
int mms_exchange_history(mms_exchange_t x, void** hist_g, void** hist_u);
/* synchronises the stream; MMS_E_FAULT if a cross-rank wait timed out since the last check */
int mms_exchange_check(mms_exchange_t x, void* cuda_stream);
unsigned long long mms_exchange_launch_count(mms_exchange_t x);

/* ------------------------------------------------------------ optimizer step ---
 * mms_adadelta_update_* has the argument meaning of the reference's
 * adadelta_update_gpu(N, g, h, h2, momentum, delta, local_rate) (src/caffe/solvers/adadelta_solver.cu:6-26,
 * called from AdaDeltaSolver::ComputeUpdateValue, adadelta_solver.cpp:96-101): g (the blob's diff) becomes the
 * update, hist_g / hist_u are the two history blobs of the parameter.
 * mms_adadelta_step_* is the whole SGDSolver::ApplyUpdate of one learnable blob (sgd_solver.cpp:102-116) in ONE pass:
 *   diff *= grad_scale              Normalize (:118-141) and the 1/solver_count of P2PSync (parallel.cpp:377)
 *   diff += local_decay * data      Regularize, L2 (:181-185); local_decay = weight_decay * decay_mult
 *   AdaDelta update of diff         as above; local_rate = base_lr * lr_mult
 *   data -= diff                    Net::Update -> Blob::Update
 *   diff = 0 if clear_diff          Net::ClearParamDiffs of the next iteration (solver.cpp:203)
 * data may be NULL (then only the first three). */
int mms_adadelta_update_f32(mms_handle_t h, float* g, float* hist_g, float* hist_u, long long count,
                            float momentum, float delta, float local_rate);
int mms_adadelta_update_f64(mms_handle_t h, double* g, double* hist_g, double* hist_u, long long count,
                            double momentum, double delta, double local_rate);
int mms_adadelta_step_f32(mms_handle_t h, float* data, float* diff, float* hist_g, float* hist_u,
                          long long count, float grad_scale, float local_decay, float momentum,
                          float delta, float local_rate, int clear_diff);
int mms_adadelta_step_f64(mms_handle_t h, double* data, double* diff, double* hist_g, double* hist_u,
                          long long count, double grad_scale, double local_decay, double momentum,
                          double delta, double local_rate, int clear_diff);

/* ------------------------------------------------------------ ranking metrics ---
 * The reference's evaluation layers are CPU-only (Forward_cpu; the scores are pulled to the host, bucketed in a
 * std::map and std::sort-ed per bucket).  Here: device in, device scalar out (a pointer to one T in device memory).
 * The score of sample i is data[i * stride + offset] -- the reference's bottom_data[i * (fixed_axis + 1) + fixed_axis]
 * (map_layer.cpp:50, mrr_layer.cpp:49) / bottom_data[i * dim + fixed_axis] (auc_layer.cpp:76): stride = number of
 * classes, offset = fixed_axis.  labels and group ids are float-encoded integers as in the reference's blobs.
 *   mms_rank_map_mrr  MAPLayer::Forward_cpu (map_layer.cpp:41-100) and MRRLayer::Forward_cpu (mrr_layer.cpp:38-79) in
 *                     one pass (either output may be NULL): groups without a positive (label 1) or without a negative
 *                     (MAP: any other label; MRR: label 0) are skipped, mean over the rest
 *   mms_rank_auc      AUCLayer::Forward_cpu (auc_layer.cpp:47-136) for (N, C) predictions
 *   mms_rank_accuracy RankAccuracyLayer::Forward_cpu (rank_accuracy_layer.cpp:36-50)
 * Samples with EQUAL scores keep their input order (the reference's std::sort leaves their order unspecified). */
int mms_rank_map_mrr_f32(mms_handle_t h, const float* data, long long stride, long long offset,
                         const float* label, const float* group, long long count, float* map_out,
                         float* mrr_out);
int mms_rank_map_mrr_f64(mms_handle_t h, const double* data, long long stride, long long offset,
                         const double* label, const double* group, long long count, double* map_out,
                         double* mrr_out);
int mms_rank_auc_f32(mms_handle_t h, const float* data, long long stride, long long offset,
                     const float* label, long long count, int has_ignore_label, int ignore_label,
                     float* out);
int mms_rank_auc_f64(mms_handle_t h, const double* data, long long stride, long long offset,
                     const double* label, long long count, int has_ignore_label, int ignore_label,
                     double* out);
int mms_rank_accuracy_f32(mms_handle_t h, const float* a, const float* b, const float* label,
                          long long count, float* out);
int mms_rank_accuracy_f64(mms_handle_t h, const double* a, const double* b, const double* label,
                          long long count, double* out);

/* ------------------------------------------------------ candidate scoring ---
 * Reranking with the SimMatrix bilinear form (BASELINE config "1k queries x 1M candidates"):
 * scores[i,j] = q_i^T W c_j.  Q (Nq,K1), C (Nc,K2), W (K1,K2), scores (Nq,Nc) row-major.
 * Same arithmetic as SimMatrixLayer::Forward (sim_matrix_layer.cpp:53-65) applied to every
 * (query, candidate) pair; QW (Nq,K2) is scratch supplied by the caller. */
int mms_rerank_scores_f32(mms_handle_t h, const float* Q, const float* C, const float* W,
                          float* QW, float* scores, int Nq, long long Nc, int K1, int K2);

/* The same against a candidate set that is scored repeatedly (a static index): mms_rerank_prepare writes the
 * TF32-rounded copy the contraction reads -- (Nc, pad4(K2)) floats, rows padded to a multiple of 4 -- into C_tf32
 * once; mms_rerank_scores_prepared then scores any number of query batches against it without re-reading and
 * re-writing the candidates (16 -> 8 GB per call at 10^6 x 1024).  Results are identical to mms_rerank_scores. */
int mms_rerank_prepare_f32(mms_handle_t h, const float* C, float* C_tf32, long long Nc, int K2);
int mms_rerank_scores_prepared_f32(mms_handle_t h, const float* Q, const float* C_tf32, const float* W, float* QW,
                                   float* scores, int Nq, long long Nc, int K1, int K2);

/* Per-query top-k instead of the full score matrix (the consumer ranks each query's candidates by
 * score: do_trec_qa_clean.py:617-650, map_layer.cpp:41-100).  top_scores / top_idx (Nq, k): the k best candidates of
 * every query, score descending, ties by candidate index ascending (NaN scores are never listed; unused slots hold
 * -inf / INT64_MAX); indices are idx_base + the candidate's row in C, so that a shard of a larger candidate set
 * reports global indices.  The scores are folded into the lists slab by slab while they are in L2 -- the Nq x Nc
 * matrix is never written.  k <= 1024.  _prepared takes the mms_rerank_prepare copy of the candidates.
 * mms_topk_merge_f32: merges lists of several shards: row q of `scores` / `idx` holds n candidate (score, index)
 * pairs of query q (the all-gathered per-shard lists, ld elements apart); out_* as above.  The result does not depend
 * on how the candidates were sharded. */
int mms_rerank_topk_f32(mms_handle_t h, const float* Q, const float* C, const float* W, float* QW, float* top_scores,
                        long long* top_idx, int Nq, long long Nc, int K1, int K2, int k, long long idx_base);
int mms_rerank_topk_prepared_f32(mms_handle_t h, const float* Q, const float* C_tf32, const float* W, float* QW,
                                 float* top_scores, long long* top_idx, int Nq, long long Nc, int K1, int K2, int k,
                                 long long idx_base);
int mms_topk_merge_f32(mms_handle_t h, const float* scores, const long long* idx, long long ld, long long n,
                       float* out_scores, long long* out_idx, int Nq, int k);

/* ------------------------------------------------------ sentence encoder ---
 * The sentence-vector variant of the net (examples/trec_qa_w2v_mms/do_trec_qa_clean.py:352-375, 412-422):
 * Convolution(kernel kh x D over the (N,1,L,D) embedded sentence) -> BN -> Pooling(MAX over time) -> TanH -> SimMatrix.
 *
 * mms_sentconv_*: ConvolutionLayer (conv_layer.cpp:25-73, base_conv_layer.cpp:257-321) for ONE input channel and a
 *   kernel as wide as the input (kernel_w = D, stride 1, pad 0, group 1): x (N,1,L,D), W (C,1,kh,D), bias (C) or
 *   NULL, top (N,C,L-kh+1,1).  No im2col buffer: the windows are overlapping views of x.  backward: dW and dbias
 *   ACCUMULATE (gemm/gemv beta 1), dx is OVERWRITTEN; any of the three may be NULL (param_propagate_down /
 *   propagate_down false).
 * mms_pool_*: PoolingLayer (pooling_layer.cpp:80-227), method 0 MAX / 1 AVE over (NC, H, W) planes, NC = num*channels;
 *   PH, PW are the pooled sizes the caller computed in Reshape (:80-110).  mask (NC,PH,PW) int32 is the reference's
 *   max_idx_ (argmax as h*W+w; first maximum in scan order), required for MAX.  backward OVERWRITES dx.
 * mms_tanh_*: TanHLayer (tanh_layer.cpp:11-37); backward takes the forward OUTPUT y.  In place is allowed.
 * mms_bn_*: the fork's BNLayer (type "BN", bn_layer.cpp:121-257 forward, :261-384 backward) over (N,C,HW):
 *   train != 0: batch statistics (var = E[x^2] - E[x]^2), running mean/var <- (1-bn_memory)*batch + bn_memory*running;
 *   train == 0: the running statistics.  x_norm (N,C,HW), batch_mean (C) and batch_std (C) = sqrt(var + var_eps) are
 *   outputs the backward reads (the reference keeps them in buffer_blob_.diff / batch_variance_; var_eps is 1e-9
 *   there).  backward OVERWRITES dscale, dshift (may be NULL) and dx (may be NULL). */
int mms_sentconv_forward_f32(mms_handle_t h, const float* x, const float* W, const float* bias, float* top, int N, int L,
                             int D, int C, int kh);
int mms_sentconv_forward_f64(mms_handle_t h, const double* x, const double* W, const double* bias, double* top, int N,
                             int L, int D, int C, int kh);
int mms_sentconv_backward_f32(mms_handle_t h, const float* x, const float* W, const float* dtop, float* dW, float* dbias,
                              float* dx, int N, int L, int D, int C, int kh);
int mms_sentconv_backward_f64(mms_handle_t h, const double* x, const double* W, const double* dtop, double* dW,
                              double* dbias, double* dx, int N, int L, int D, int C, int kh);
int mms_pool_forward_f32(mms_handle_t h, const float* x, float* top, int* mask, long long NC, int H, int W, int PH, int PW,
                         int kh, int kw, int sh, int sw, int pad_h, int pad_w, int method);
int mms_pool_forward_f64(mms_handle_t h, const double* x, double* top, int* mask, long long NC, int H, int W, int PH,
                         int PW, int kh, int kw, int sh, int sw, int pad_h, int pad_w, int method);
int mms_pool_backward_f32(mms_handle_t h, const float* dtop, const int* mask, float* dx, long long NC, int H, int W, int PH,
                          int PW, int kh, int kw, int sh, int sw, int pad_h, int pad_w, int method);
int mms_pool_backward_f64(mms_handle_t h, const double* dtop, const int* mask, double* dx, long long NC, int H, int W,
                          int PH, int PW, int kh, int kw, int sh, int sw, int pad_h, int pad_w, int method);
int mms_tanh_forward_f32(mms_handle_t h, const float* x, float* y, long long count);
int mms_tanh_forward_f64(mms_handle_t h, const double* x, double* y, long long count);
int mms_tanh_backward_f32(mms_handle_t h, const float* y, const float* dy, float* dx, long long count);
int mms_tanh_backward_f64(mms_handle_t h, const double* y, const double* dy, double* dx, long long count);
int mms_bn_forward_f32(mms_handle_t h, const float* x, const float* scale, const float* shift, float* run_mean,
                       float* run_var, float* top, float* x_norm, float* batch_mean, float* batch_std, int N, int C, int HW,
                       int train, float bn_memory, float var_eps);
int mms_bn_forward_f64(mms_handle_t h, const double* x, const double* scale, const double* shift, double* run_mean,
                       double* run_var, double* top, double* x_norm, double* batch_mean, double* batch_std, int N, int C,
                       int HW, int train, double bn_memory, double var_eps);
int mms_bn_backward_f32(mms_handle_t h, const float* dtop, const float* x_norm, const float* scale, const float* batch_std,
                        float* dscale, float* dshift, float* dx, int N, int C, int HW);
int mms_bn_backward_f64(mms_handle_t h, const double* dtop, const double* x_norm, const double* scale,
                        const double* batch_std, double* dscale, double* dshift, double* dx, int N, int C, int HW);

/* ------------------------------------------------ the CNN over the similarity tensor ---
 * network_v4 (examples/trec_qa_w2v_mms/do_trec_qa_clean.py:470-477): Dropout(0.1) -> Convolution(5x5, 32) + BN ->
 * Pooling(AVE 4, stride 4) -> TanH -> Convolution(5x5, 64) + BN -> Pooling(AVE 5) -> TanH over S (N, mc, Lq, La).  BN,
 * Pooling and TanH are the entry points above (BN takes (N, C, H*W)).
 * mms_conv2d_*: ConvolutionLayer (conv_layer.cpp:25-73, base_conv_layer.cpp:257-321) for stride 1, pad 0, group 1:
 *   x (N,C,H,W), W (Co,C,kh,kw), bias (Co) or NULL, top (N,Co,H-kh+1,W-kw+1).  Implicit GEMMs on the tensor cores for
 *   float (no im2col buffer); direct sums for double / MMS_MATH_FP32.  backward: dW and dbias ACCUMULATE, dx is
 *   OVERWRITTEN; any of the three may be NULL.
 * mms_dropout_*: DropoutLayer::Forward_gpu / Backward_gpu (dropout_layer.cu:10-45): y = x * (mask > threshold) * scale
 *   with one random 32-bit word per element, threshold = UINT_MAX * dropout_ratio, scale = 1 / (1 - dropout_ratio);
 *   the backward is the same call on the top gradient.  mms_dropout_mask fills `mask` from a counter-based generator
 *   (the reference draws it with cuRAND, dropout_layer.cu:27; any source of uniform words is equivalent). */
int mms_conv2d_forward_f32(mms_handle_t h, const float* x, const float* W, const float* bias, float* top, int N, int C,
                           int H, int Wd, int Co, int kh, int kw);
int mms_conv2d_forward_f64(mms_handle_t h, const double* x, const double* W, const double* bias, double* top, int N,
                           int C, int H, int Wd, int Co, int kh, int kw);
int mms_conv2d_backward_f32(mms_handle_t h, const float* x, const float* W, const float* dtop, float* dW, float* dbias,
                            float* dx, int N, int C, int H, int Wd, int Co, int kh, int kw);
int mms_conv2d_backward_f64(mms_handle_t h, const double* x, const double* W, const double* dtop, double* dW,
                            double* dbias, double* dx, int N, int C, int H, int Wd, int Co, int kh, int kw);
int mms_dropout_f32(mms_handle_t h, const float* x, const unsigned* mask, float* y, long long count,
                    unsigned threshold, float scale);
int mms_dropout_f64(mms_handle_t h, const double* x, const unsigned* mask, double* y, long long count,
                    unsigned threshold, double scale);
int mms_dropout_mask(mms_handle_t h, unsigned* mask, long long count, unsigned long long seed);

/* ------------------------------------------------- input formats (host) ---
 * embed_param.weight_source: the pre-trained word-vector file EmbedLayer::LayerSetUp reads into blobs_[0]
 * (embed_layer.cpp:46-113).  HOST memory: table_host is the (input_dim, num_output) table as the weight filler left
 * it (blobs_[0]->mutable_cpu_data()); records overwrite rows 0.. in file order, the other rows keep their values.
 * Format by file-name suffix as in the reference: "txt" GloVe text, "all" the fork's id/vector/word dump (header
 * must say input_dim-1, num_output-1), anything else word2vec binary (its dimension must equal num_output).
 * Values are stored as 32-bit floats into the first four bytes of each element -- for _f64 that is the reference's
 * `(float*)(weight_data + w_index)` behaviour, kept for identical results.  Unlike the reference, a missing file, a
 * malformed record or more records than rows is an error (MMS_E_INVALID), not undefined behaviour.
 * rows_loaded (may be NULL) receives the number of records read.  No GPU is touched. */
int mms_load_weight_source_f32(const char* path, float* table_host, long long input_dim, long long num_output,
                               long long* rows_loaded);
int mms_load_weight_source_f64(const char* path, double* table_host, long long input_dim, long long num_output,
                               long long* rows_loaded);

/* --------------------------------------------------------- diagnostics ------
 * The tcgen05 TF32 GEMM building block, exposed for tests and profiling:
 * C (+)= op(A) op(B), M x N x K.  a_mn = 0: A(m,k) = A[m*lda + k] (K-major), 1: A[k*lda + m]
 * (MN-major); b_mn likewise with B(n,k).  mode 0 store, 1 +=, 2 atomicAdd (required when
 * ksplit > 1); mode | 0x100 declares that A and B already hold TF32-exact values, which lets the
 * TMA-fed kernel fetch them (otherwise operands are rounded to TF32 in registers on the way in). */
int mms_tc_gemm_f32(mms_handle_t h, const float* A, long long lda, int a_mn, const float* B,
                    long long ldb, int b_mn, float* C, long long ldc, int M, int N, int K, int ksplit,
                    int mode);

#ifdef __cplusplus
}
#endif
#endif /* MMS_B200_H_ */
