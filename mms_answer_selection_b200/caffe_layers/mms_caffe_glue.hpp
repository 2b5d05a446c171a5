// Glue between Caffe's Layer API and the C-ABI of libmms_b200.so (include/mms_b200.h).
//
// The TUs in this directory REPLACE src/caffe/layers/{embed,sim_cross,sim_matrix,
// pair_rank_loss,fm,bn}_layer.{cpp,cu} and {map,mrr,auc,rank_accuracy}_layer.cpp of the reference in the link (a layer type can be registered
// once, include/caffe/layer_factory.hpp:69-70).  They are compiled against the reference's own,
// unmodified headers, so the class declarations -- members included -- are the reference's; the
// per-thread workspace handle therefore lives here, not in the classes.
//
// Host code only: no CUDA kernels, no cuBLAS.  Forward_gpu / Backward_gpu marshal Blob device
// pointers and sizes into the C-ABI; Forward_cpu / Backward_cpu abort -- there is no CPU path.
#ifndef MMS_CAFFE_GLUE_HPP_
#define MMS_CAFFE_GLUE_HPP_

#include "caffe/common.hpp"
#include "mms_b200.h"

namespace caffe {
namespace mms {

// One workspace per host thread (Caffe runs one solver thread per GPU, parallel.cpp:271-284, and its
// Caffe singleton is thread-local, common.cpp:15-20).  Work is enqueued on the legacy default
// stream, which is what every other Caffe layer uses (SURVEY.md 8(b) "Stream / threading").
mms_handle_t handle();

// Caffe's error convention is abort-with-log (CUDA_CHECK -> LOG(FATAL), device_alternate.hpp:48-53).
#define MMS_CAFFE_CHECK(expr)                                                   \
  do {                                                                          \
    const int mms_rc_ = (expr);                                                 \
    CHECK_EQ(mms_rc_, 0) << "libmms_b200: " << mms_last_error();                \
  } while (0)

#define MMS_NO_CPU_PATH(Layer)                                                  \
  LOG(FATAL) << #Layer " (mms_b200) runs on the GPU only: set Caffe::set_mode(Caffe::GPU)"

// Overloads that pick the _f32 / _f64 entry point from the blob type (Caffe instantiates every
// layer for float and double, common.hpp:41-66).
#define MMS_OVERLOAD(name, PARAMS_F, PARAMS_D, ARGS)                            \
  inline int name PARAMS_F { return mms_##name##_f32 ARGS; }                    \
  inline int name PARAMS_D { return mms_##name##_f64 ARGS; }

MMS_OVERLOAD(embed_forward,
             (mms_handle_t h, const float* i, const float* W, const float* b, float* t, long long M, int D, int V),
             (mms_handle_t h, const double* i, const double* W, const double* b, double* t, long long M, int D, int V),
             (h, i, W, b, t, M, D, V))
MMS_OVERLOAD(embed_backward,
             (mms_handle_t h, const float* i, const float* dt, float* dW, float* db, long long M, int D, int V),
             (mms_handle_t h, const double* i, const double* dt, double* dW, double* db, long long M, int D, int V),
             (h, i, dt, dW, db, M, D, V))
MMS_OVERLOAD(simcross_forward,
             (mms_handle_t h, int mode, const float* q, const float* a, const float* M, const float* B, float* S,
              float* n0, float* n1, int N, int Lq, int La, int D, int mc),
             (mms_handle_t h, int mode, const double* q, const double* a, const double* M, const double* B, double* S,
              double* n0, double* n1, int N, int Lq, int La, int D, int mc),
             (h, mode, q, a, M, B, S, n0, n1, N, Lq, La, D, mc))
MMS_OVERLOAD(simcross_backward,
             (mms_handle_t h, int mode, const float* q, const float* a, const float* M, const float* S,
              const float* dS, const float* n0, const float* n1, float* dq, float* da, float* dM, float* dB, int N,
              int Lq, int La, int D, int mc, int p0, int p1),
             (mms_handle_t h, int mode, const double* q, const double* a, const double* M, const double* S,
              const double* dS, const double* n0, const double* n1, double* dq, double* da, double* dM, double* dB,
              int N, int Lq, int La, int D, int mc, int p0, int p1),
             (h, mode, q, a, M, S, dS, n0, n1, dq, da, dM, dB, N, Lq, La, D, mc, p0, p1))
MMS_OVERLOAD(simmatrix_forward,
             (mms_handle_t h, const float* q, const float* a, const float* W, float* s, float* T, int N, int K1, int K2),
             (mms_handle_t h, const double* q, const double* a, const double* W, double* s, double* T, int N, int K1,
              int K2),
             (h, q, a, W, s, T, N, K1, K2))
MMS_OVERLOAD(simmatrix_backward,
             (mms_handle_t h, const float* q, const float* a, const float* W, const float* ds, float* dW, float* dq,
              float* da, int N, int K1, int K2, int pw, int p0, int p1),
             (mms_handle_t h, const double* q, const double* a, const double* W, const double* ds, double* dW,
              double* dq, double* da, int N, int K1, int K2, int pw, int p0, int p1),
             (h, q, a, W, ds, dW, dq, da, N, K1, K2, pw, p0, p1))
MMS_OVERLOAD(pairrankloss_forward,
             (mms_handle_t h, const float* a, const float* b, const float* y, float m, long long n, float* loss,
              float* o, float* s),
             (mms_handle_t h, const double* a, const double* b, const double* y, double m, long long n, double* loss,
              double* o, double* s),
             (h, a, b, y, m, n, loss, o, s))
MMS_OVERLOAD(pairrankloss_backward,
             (mms_handle_t h, const float* y, const float* o, const float* s, float td, long long n, float* da,
              float* db),
             (mms_handle_t h, const double* y, const double* o, const double* s, double td, long long n, double* da,
              double* db),
             (h, y, o, s, td, n, da, db))
MMS_OVERLOAD(fm_forward,
             (mms_handle_t h, const float* x, const float* b, float* y, int N, int C, int Dm),
             (mms_handle_t h, const double* x, const double* b, double* y, int N, int C, int Dm),
             (h, x, b, y, N, C, Dm))
MMS_OVERLOAD(fm_backward,
             (mms_handle_t h, const float* x, const float* dy, float* dx, float* db, int N, int C, int Dm, int p0),
             (mms_handle_t h, const double* x, const double* dy, double* dx, double* db, int N, int C, int Dm, int p0),
             (h, x, dy, dx, db, N, C, Dm, p0))
MMS_OVERLOAD(rank_map_mrr,
             (mms_handle_t h, const float* d, long long st, long long off, const float* l, const float* g, long long n,
              float* map, float* mrr),
             (mms_handle_t h, const double* d, long long st, long long off, const double* l, const double* g,
              long long n, double* map, double* mrr),
             (h, d, st, off, l, g, n, map, mrr))
MMS_OVERLOAD(rank_auc,
             (mms_handle_t h, const float* d, long long st, long long off, const float* l, long long n, int hi, int il,
              float* out),
             (mms_handle_t h, const double* d, long long st, long long off, const double* l, long long n, int hi,
              int il, double* out),
             (h, d, st, off, l, n, hi, il, out))
MMS_OVERLOAD(rank_accuracy,
             (mms_handle_t h, const float* a, const float* b, const float* y, long long n, float* out),
             (mms_handle_t h, const double* a, const double* b, const double* y, long long n, double* out),
             (h, a, b, y, n, out))
MMS_OVERLOAD(bn_forward,
             (mms_handle_t h, const float* x, const float* sc, const float* sh, float* rm, float* rv, float* top, float* xn,
              float* bm, float* bs, int N, int C, int HW, int train, float mem, float eps),
             (mms_handle_t h, const double* x, const double* sc, const double* sh, double* rm, double* rv, double* top,
              double* xn, double* bm, double* bs, int N, int C, int HW, int train, double mem, double eps),
             (h, x, sc, sh, rm, rv, top, xn, bm, bs, N, C, HW, train, mem, eps))
MMS_OVERLOAD(bn_backward,
             (mms_handle_t h, const float* g, const float* xn, const float* sc, const float* bs, float* dsc, float* dsh,
              float* dx, int N, int C, int HW),
             (mms_handle_t h, const double* g, const double* xn, const double* sc, const double* bs, double* dsc,
              double* dsh, double* dx, int N, int C, int HW),
             (h, g, xn, sc, bs, dsc, dsh, dx, N, C, HW))
MMS_OVERLOAD(load_weight_source,
             (const char* path, float* t, long long K, long long N, long long* n),
             (const char* path, double* t, long long K, long long N, long long* n),
             (path, t, K, N, n))
#undef MMS_OVERLOAD

}  // namespace mms
}  // namespace caffe

#endif  // MMS_CAFFE_GLUE_HPP_
