// TEST INFRASTRUCTURE ONLY -- never part of the product.
//
// oracle/_ref/libmms_refcuda.so = the reference's OWN CUDA sources for the solver step, compiled verbatim
// where they lie (oracle/Makefile `refcuda`):
//     src/caffe/solvers/adadelta_solver.cu   AdaDeltaUpdate / adadelta_update_gpu        (:5-26)
//     src/caffe/util/math_functions.cu       caffe_gpu_scal / caffe_gpu_axpy (cuBLAS)    (:96-120, :98-110)
// plus this file: the Caffe singleton those TUs need (a cuBLAS handle and a cuRAND generator, as
// common.cpp:105-122 creates them) and a C entry per checked routine.  It pins the product's
// mms_adadelta_update_* / mms_adadelta_step_* BY EXECUTION on the GPU box (tests/test_gpu_parity.py).
//
// mmsrefcu_apply_update runs, for ONE learnable blob, exactly the calls SGDSolver::ApplyUpdate makes in
// GPU mode, each through the reference's own routine:
//     Normalize   caffe_gpu_scal(count, 1/iter_size, diff)          sgd_solver.cpp:118-141 (skipped when the factor is 1)
//     Regularize  caffe_gpu_axpy(count, local_decay, data, diff)    sgd_solver.cpp:143-204 (L2; skipped when local_decay == 0)
//     ComputeUpdateValue  adadelta_update_gpu(count, diff, h, h2, momentum, delta, local_rate)   adadelta_solver.cpp:96-101
//     Net::Update -> Blob::Update  caffe_gpu_axpy(count, -1, diff, data)                          blob.cpp:160-183
// SGDSolver / Net themselves cannot be compiled here: sgd_solver.cpp pulls in solver.hpp -> net.hpp ->
// the protobuf-generated NetParameter/SolverParameter/SolverState classes and caffe/util/hdf5.hpp (hdf5.h), neither
// of which exists in this image; the four calls above are everything those two classes contribute to one blob's step.
#include <cublas_v2.h>
#include <curand.h>

#include "caffe/common.hpp"
#include "caffe/util/math_functions.hpp"

namespace caffe {

template <typename Dtype>
void adadelta_update_gpu(int N, Dtype* g, Dtype* h, Dtype* h2, Dtype momentum, Dtype delta, Dtype local_rate);

static thread_local Caffe* tls_caffe = nullptr;
Caffe& Caffe::Get() {
  if (!tls_caffe) tls_caffe = new Caffe();
  return *tls_caffe;
}
Caffe::Caffe()
    : cublas_handle_(NULL), curand_generator_(NULL), random_generator_(), mode_(Caffe::GPU), solver_count_(1),
      root_solver_(true) {
  if (cublasCreate(&cublas_handle_) != CUBLAS_STATUS_SUCCESS) cublas_handle_ = NULL;
  if (curandCreateGenerator(&curand_generator_, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS)
    curand_generator_ = NULL;
}
Caffe::~Caffe() {
  if (cublas_handle_) cublasDestroy(cublas_handle_);
  if (curand_generator_) curandDestroyGenerator(curand_generator_);
}
void Caffe::set_random_seed(const unsigned int) {}
void Caffe::SetDevice(const int device_id) { cudaSetDevice(device_id); }
void Caffe::DeviceQuery() {}
bool Caffe::CheckDevice(const int) { return true; }
int Caffe::FindDevice(const int start_id) { return start_id; }
const char* cublasGetErrorString(cublasStatus_t) { return "cuBLAS error"; }
const char* curandGetErrorString(curandStatus_t) { return "cuRAND error"; }

}  // namespace caffe

namespace {
template <typename T>
int apply_update(long long n, T* data, T* diff, T* h, T* h2, double accum_normalization, double local_decay,
                 double momentum, double delta, double local_rate) {
  if (!caffe::Caffe::cublas_handle()) return -1;
  const int N = (int)n;
  if (accum_normalization != 1.0) caffe::caffe_gpu_scal<T>(N, (T)accum_normalization, diff);
  if (data && local_decay != 0.0) caffe::caffe_gpu_axpy<T>(N, (T)local_decay, data, diff);
  caffe::adadelta_update_gpu<T>(N, diff, h, h2, (T)momentum, (T)delta, (T)local_rate);
  if (data) caffe::caffe_gpu_axpy<T>(N, T(-1), diff, data);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}
}  // namespace

extern "C" {
// dtype 0 float / 1 double; all pointers are DEVICE pointers; data may be NULL (bare adadelta_update_gpu)
int mmsrefcu_apply_update(int dtype, long long n, void* data, void* diff, void* h, void* h2, double accum_normalization,
                          double local_decay, double momentum, double delta, double local_rate) {
  try {
    if (dtype == 0)
      return apply_update<float>(n, (float*)data, (float*)diff, (float*)h, (float*)h2, accum_normalization, local_decay,
                                 momentum, delta, local_rate);
    return apply_update<double>(n, (double*)data, (double*)diff, (double*)h, (double*)h2, accum_normalization,
                                local_decay, momentum, delta, local_rate);
  } catch (...) {
    return -3;
  }
}
}
