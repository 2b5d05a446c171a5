// TEST INFRASTRUCTURE ONLY -- GPU flavour of ref_runtime.cpp for the drop-in harness.
//
// oracle/_ref/libmms_dropin.so links the reference's OWN blob.cpp / syncedmem.cpp / layer.cpp /
// loss_layer.cpp (compiled in place, GPU build = no CPU_ONLY) with the PRODUCT's drop-in layer
// TUs (mms_answer_selection_b200/caffe_layers/*.cpp) and libmms_b200.so, so that tests can drive
// the new layers through the reference's real Layer API, Blob and SyncedMemory in Caffe::GPU mode.
// This file stands in for the two reference TUs that cannot be compiled here (common.cpp needs
// boost/glog/cuRAND setup, math_functions.{cpp,cu} need boost.random / cuBLAS): it provides the
// Caffe singleton and exactly the math wrappers the linked reference TUs call.
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

#include "caffe/common.hpp"
#include "caffe/util/math_functions.hpp"
#include "mms_b200.h"

namespace caffe {

static thread_local Caffe* tls_caffe = nullptr;

Caffe& Caffe::Get() {
  if (!tls_caffe) tls_caffe = new Caffe();
  return *tls_caffe;
}

// common.cpp:105-122 creates a cuBLAS handle and a cuRAND generator here; none of the linked TUs
// uses either, so both stay NULL.
Caffe::Caffe()
    : cublas_handle_(NULL), curand_generator_(NULL), random_generator_(), mode_(Caffe::GPU), solver_count_(1),
      root_solver_(true) {}
Caffe::~Caffe() {}

static std::mt19937& harness_rng() {
  static thread_local std::mt19937 g(1701);
  return g;
}
void Caffe::set_random_seed(const unsigned int seed) { harness_rng().seed(seed); }
void Caffe::SetDevice(const int device_id) { CUDA_CHECK(cudaSetDevice(device_id)); }
void Caffe::DeviceQuery() {}
bool Caffe::CheckDevice(const int) { return true; }
int Caffe::FindDevice(const int start_id) { return start_id; }

const char* cublasGetErrorString(cublasStatus_t) { return "cuBLAS is not linked into the drop-in harness"; }
const char* curandGetErrorString(curandStatus_t) { return "cuRAND is not linked into the drop-in harness"; }

void caffe_gpu_memcpy(const size_t N, const void* X, void* Y) {
  if (X != Y) CUDA_CHECK(cudaMemcpy(Y, X, N, cudaMemcpyDefault));   // math_functions.cu:91-95
}

namespace {
mms_handle_t harness_handle() {
  static thread_local mms_handle_t h = nullptr;
  if (!h) CHECK_EQ(mms_create(&h), 0) << mms_last_error();
  return h;
}
template <typename T> struct Dev {
  T* p;
  explicit Dev(size_t n) { CUDA_CHECK(cudaMalloc(&p, n * sizeof(T))); }
  ~Dev() { cudaFree(p); }
};
int dot_abi(const float* x, const float* y, int n, float* out) { return mms_dot_f32(harness_handle(), x, y, n, out); }
int dot_abi(const double* x, const double* y, int n, double* out) { return mms_dot_f64(harness_handle(), x, y, n, out); }
int scale_abi(float* x, int n, float a) { return mms_scale_f32(harness_handle(), x, n, a); }
int scale_abi(double* x, int n, double a) { return mms_scale_f64(harness_handle(), x, n, a); }
}  // namespace

// the loss reduction of Layer::Forward (layer.hpp:471-479) goes through the product's own kernel
template <typename Dtype>
void caffe_gpu_dot(const int n, const Dtype* x, const Dtype* y, Dtype* out) {
  Dev<Dtype> d(1);
  CHECK_EQ(dot_abi(x, y, n, d.p), 0) << mms_last_error();
  CUDA_CHECK(cudaMemcpy(out, d.p, sizeof(Dtype), cudaMemcpyDeviceToHost));
}
template void caffe_gpu_dot<float>(const int, const float*, const float*, float*);
template void caffe_gpu_dot<double>(const int, const double*, const double*, double*);

template <typename Dtype>
void caffe_gpu_scal(const int N, const Dtype alpha, Dtype* X) {
  CHECK_EQ(scale_abi(X, N, alpha), 0) << mms_last_error();
}
template void caffe_gpu_scal<float>(const int, const float, float*);
template void caffe_gpu_scal<double>(const int, const double, double*);

// Blob::Update / asum_* in GPU mode: not on the tested path; host round trips keep them correct.
template <typename Dtype>
void caffe_gpu_axpy(const int N, const Dtype alpha, const Dtype* X, Dtype* Y) {
  std::vector<Dtype> x(N), y(N);
  CUDA_CHECK(cudaMemcpy(x.data(), X, N * sizeof(Dtype), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(y.data(), Y, N * sizeof(Dtype), cudaMemcpyDeviceToHost));
  for (int i = 0; i < N; ++i) y[i] += alpha * x[i];
  CUDA_CHECK(cudaMemcpy(Y, y.data(), N * sizeof(Dtype), cudaMemcpyHostToDevice));
}
template void caffe_gpu_axpy<float>(const int, const float, const float*, float*);
template void caffe_gpu_axpy<double>(const int, const double, const double*, double*);

template <typename Dtype>
void caffe_gpu_asum(const int n, const Dtype* x, Dtype* y) {
  std::vector<Dtype> h(n);
  CUDA_CHECK(cudaMemcpy(h.data(), x, n * sizeof(Dtype), cudaMemcpyDeviceToHost));
  Dtype s = 0;
  for (int i = 0; i < n; ++i) s += std::fabs(h[i]);
  *y = s;
}
template void caffe_gpu_asum<float>(const int, const float*, float*);
template void caffe_gpu_asum<double>(const int, const double*, double*);

// ---- host-side wrappers used by blob.cpp and filler.hpp -----------------------------------
#define MMS_HOST_MATH(T)                                                                        \
  template <> void caffe_axpy<T>(const int N, const T alpha, const T* X, T* Y) {                \
    for (int i = 0; i < N; ++i) Y[i] += alpha * X[i];                                           \
  }                                                                                             \
  template <> void caffe_scal<T>(const int N, const T alpha, T* X) {                            \
    for (int i = 0; i < N; ++i) X[i] *= alpha;                                                  \
  }                                                                                             \
  template <> T caffe_cpu_asum<T>(const int n, const T* x) {                                    \
    T s = 0; for (int i = 0; i < n; ++i) s += std::fabs(x[i]); return s;                        \
  }                                                                                             \
  template <> void caffe_rng_uniform<T>(const int n, const T a, const T b, T* r) {              \
    std::uniform_real_distribution<T> dist(a, b);                                               \
    for (int i = 0; i < n; ++i) r[i] = dist(harness_rng());                                     \
  }                                                                                             \
  template <> void caffe_rng_gaussian<T>(const int n, const T mu, const T sigma, T* r) {        \
    std::normal_distribution<T> dist(mu, sigma);                                                \
    for (int i = 0; i < n; ++i) r[i] = dist(harness_rng());                                     \
  }                                                                                             \
  template <> void caffe_rng_bernoulli<T>(const int n, const T p, int* r) {                     \
    std::bernoulli_distribution dist(p);                                                        \
    for (int i = 0; i < n; ++i) r[i] = dist(harness_rng());                                     \
  }
MMS_HOST_MATH(float)
MMS_HOST_MATH(double)

template <typename Dtype>
Dtype caffe_cpu_dot(const int n, const Dtype* x, const Dtype* y) {
  Dtype s = 0;
  for (int i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
template float caffe_cpu_dot<float>(const int, const float*, const float*);
template double caffe_cpu_dot<double>(const int, const double*, const double*);

template <typename Dtype>
void caffe_set(const int N, const Dtype alpha, Dtype* Y) {
  for (int i = 0; i < N; ++i) Y[i] = alpha;
}
template void caffe_set<int>(const int, const int, int*);
template void caffe_set<float>(const int, const float, float*);
template void caffe_set<double>(const int, const double, double*);

// math_functions.cpp:96-118: cudaMemcpyDefault in GPU mode (pointers may be host or device)
template <typename Dtype>
void caffe_copy(const int N, const Dtype* X, Dtype* Y) {
  if (X == Y) return;
  if (Caffe::mode() == Caffe::GPU) CUDA_CHECK(cudaMemcpy(Y, X, sizeof(Dtype) * N, cudaMemcpyDefault));
  else memcpy(Y, X, sizeof(Dtype) * N);
}
template void caffe_copy<int>(const int, const int*, int*);
template void caffe_copy<unsigned int>(const int, const unsigned int*, unsigned int*);
template void caffe_copy<float>(const int, const float*, float*);
template void caffe_copy<double>(const int, const double*, double*);

}  // namespace caffe
