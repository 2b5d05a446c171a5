// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle, never of the product.
//
// Runtime glue that lets the reference's hot-path layer sources link without
// boost / cuBLAS / cuRAND.  It stands in for two reference TUs that cannot be
// compiled here because they include boost.random / boost.math:
//   * src/caffe/common.cpp            (the Caffe singleton, CPU_ONLY flavour)
//   * src/caffe/util/math_functions.cpp (only the wrappers the five MMS layers,
//                                        blob.cpp and filler.hpp call)
// The BLAS calling conventions (row-major, lda/ldb/ldc choice) are the ones
// fixed by the reference at src/caffe/util/math_functions.cpp:13-58; the
// arithmetic itself lives in the un-vendored BLAS (OpenBLAS 0.3.15 here).
#include <cmath>
#include <cstring>
#include <random>

#include "caffe/common.hpp"
#include "caffe/util/math_functions.hpp"

namespace caffe {

// ---- Caffe singleton (common.cpp:15-20,55-79 CPU_ONLY branch) ---------------
static thread_local Caffe* tls_caffe = nullptr;

Caffe& Caffe::Get() {
  if (!tls_caffe) tls_caffe = new Caffe();
  return *tls_caffe;
}

Caffe::Caffe()
    : random_generator_(), mode_(Caffe::CPU), solver_count_(1), root_solver_(true) {}
Caffe::~Caffe() {}

static std::mt19937& oracle_rng() {
  static thread_local std::mt19937 g(1701);
  return g;
}

void Caffe::set_random_seed(const unsigned int seed) { oracle_rng().seed(seed); }
void Caffe::SetDevice(const int) { NO_GPU; }
void Caffe::DeviceQuery() { NO_GPU; }
bool Caffe::CheckDevice(const int) { NO_GPU; return false; }
int Caffe::FindDevice(const int) { NO_GPU; return -1; }

// ---- BLAS level 1/2/3 wrappers (math_functions.cpp:13-58, 112-160, 340-387) --
#define MMS_REF_BLAS(T, pfx)                                                          \
  template <>                                                                       \
  void caffe_cpu_gemm<T>(const CBLAS_TRANSPOSE TransA, const CBLAS_TRANSPOSE TransB, \
                         const int M, const int N, const int K, const T alpha,      \
                         const T* A, const T* B, const T beta, T* C) {              \
    const int lda = (TransA == CblasNoTrans) ? K : M;                               \
    const int ldb = (TransB == CblasNoTrans) ? N : K;                               \
    cblas_##pfx##gemm(CblasRowMajor, TransA, TransB, M, N, K, alpha, A, lda, B, ldb,  \
                    beta, C, N);                                                    \
  }                                                                                 \
  template <>                                                                       \
  void caffe_cpu_gemv<T>(const CBLAS_TRANSPOSE TransA, const int M, const int N,    \
                         const T alpha, const T* A, const T* x, const T beta,       \
                         T* y) {                                                    \
    cblas_##pfx##gemv(CblasRowMajor, TransA, M, N, alpha, A, N, x, 1, beta, y, 1);    \
  }                                                                                 \
  template <>                                                                       \
  void caffe_cpu_ger<T>(const int M, const int N, const T alpha, const T* x,        \
                        const T* y, T* A) {                                         \
    cblas_##pfx##ger(CblasRowMajor, M, N, alpha, x, 1, y, 1, A, N);                   \
  }                                                                                 \
  template <>                                                                       \
  void caffe_axpy<T>(const int N, const T alpha, const T* X, T* Y) {                \
    cblas_##pfx##axpy(N, alpha, X, 1, Y, 1);                                          \
  }                                                                                 \
  template <>                                                                       \
  void caffe_scal<T>(const int N, const T alpha, T* X) {                            \
    cblas_##pfx##scal(N, alpha, X, 1);                                                \
  }                                                                                 \
  template <>                                                                       \
  void caffe_cpu_axpby<T>(const int N, const T alpha, const T* X, const T beta,     \
                          T* Y) {                                                   \
    cblas_##pfx##axpby(N, alpha, X, 1, beta, Y, 1);                                   \
  }                                                                                 \
  template <>                                                                       \
  T caffe_cpu_strided_dot<T>(const int n, const T* x, const int incx, const T* y,   \
                             const int incy) {                                      \
    return cblas_##pfx##dot(n, x, incx, y, incy);                                     \
  }                                                                                 \
  template <>                                                                       \
  T caffe_cpu_asum<T>(const int n, const T* x) {                                    \
    return cblas_##pfx##asum(n, x, 1);                                                \
  }                                                                                 \
  template <>                                                                       \
  void caffe_cpu_scale<T>(const int n, const T alpha, const T* x, T* y) {           \
    cblas_##pfx##copy(n, x, 1, y, 1);                                                 \
    cblas_##pfx##scal(n, alpha, y, 1);                                                \
  }                                                                                 \
  template <>                                                                       \
  void caffe_add_scalar(const int N, const T alpha, T* Y) {                         \
    for (int i = 0; i < N; ++i) Y[i] += alpha;                                      \
  }                                                                                 \
  template <>                                                                       \
  void caffe_add<T>(const int n, const T* a, const T* b, T* y) {                    \
    for (int i = 0; i < n; ++i) y[i] = a[i] + b[i];                                 \
  }                                                                                 \
  template <>                                                                       \
  void caffe_sub<T>(const int n, const T* a, const T* b, T* y) {                    \
    for (int i = 0; i < n; ++i) y[i] = a[i] - b[i];                                 \
  }                                                                                 \
  template <>                                                                       \
  void caffe_mul<T>(const int n, const T* a, const T* b, T* y) {                    \
    for (int i = 0; i < n; ++i) y[i] = a[i] * b[i];                                 \
  }                                                                                 \
  template <>                                                                       \
  void caffe_div<T>(const int n, const T* a, const T* b, T* y) {                    \
    for (int i = 0; i < n; ++i) y[i] = a[i] / b[i];                                 \
  }                                                                                 \
  template <>                                                                       \
  void caffe_sqr<T>(const int n, const T* a, T* y) {                                \
    for (int i = 0; i < n; ++i) y[i] = a[i] * a[i];                                 \
  }                                                                                 \
  template <>                                                                       \
  void caffe_abs<T>(const int n, const T* a, T* y) {                                \
    for (int i = 0; i < n; ++i) y[i] = std::fabs(a[i]);                             \
  }                                                                                 \
  template <>                                                                       \
  void caffe_powx<T>(const int n, const T* a, const T b, T* y) {                    \
    for (int i = 0; i < n; ++i) y[i] = std::pow(a[i], b);                           \
  }                                                                                 \
  template <>                                                                       \
  void caffe_rng_uniform<T>(const int n, const T a, const T b, T* r) {              \
    std::uniform_real_distribution<T> dist(a, b);                                      \
    for (int i = 0; i < n; ++i) r[i] = dist(oracle_rng());                             \
  }                                                                                 \
  template <>                                                                       \
  void caffe_rng_gaussian<T>(const int n, const T mu, const T sigma, T* r) {        \
    std::normal_distribution<T> dist(mu, sigma);                                       \
    for (int i = 0; i < n; ++i) r[i] = dist(oracle_rng());                             \
  }                                                                                 \
  template <>                                                                       \
  void caffe_rng_bernoulli<T>(const int n, const T p, int* r) {                     \
    std::bernoulli_distribution dist(p);                                               \
    for (int i = 0; i < n; ++i) r[i] = dist(oracle_rng());                             \
  }                                                                                 \
  template <>                                                                       \
  void caffe_rng_bernoulli<T>(const int n, const T p, unsigned int* r) {            \
    std::bernoulli_distribution dist(p);                                               \
    for (int i = 0; i < n; ++i) r[i] = dist(oracle_rng());                             \
  }

MMS_REF_BLAS(float, s)
MMS_REF_BLAS(double, d)

unsigned int caffe_rng_rand() { return oracle_rng()(); }

template <typename Dtype>
Dtype caffe_cpu_dot(const int n, const Dtype* x, const Dtype* y) {
  return caffe_cpu_strided_dot(n, x, 1, y, 1);
}
template float caffe_cpu_dot<float>(const int, const float*, const float*);
template double caffe_cpu_dot<double>(const int, const double*, const double*);

// math_functions.cpp:66-79
template <typename Dtype>
void caffe_set(const int N, const Dtype alpha, Dtype* Y) {
  if (alpha == 0) {
    memset(Y, 0, sizeof(Dtype) * N);
    return;
  }
  for (int i = 0; i < N; ++i) Y[i] = alpha;
}
template void caffe_set<int>(const int, const int, int*);
template void caffe_set<float>(const int, const float, float*);
template void caffe_set<double>(const int, const double, double*);

// math_functions.cpp:96-118 (CPU branch)
template <typename Dtype>
void caffe_copy(const int N, const Dtype* X, Dtype* Y) {
  if (X != Y) memcpy(Y, X, sizeof(Dtype) * N);
}
template void caffe_copy<int>(const int, const int*, int*);
template void caffe_copy<unsigned int>(const int, const unsigned int*, unsigned int*);
template void caffe_copy<float>(const int, const float*, float*);
template void caffe_copy<double>(const int, const double*, double*);

}  // namespace caffe
