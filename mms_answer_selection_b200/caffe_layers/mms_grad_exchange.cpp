// See mms_grad_exchange.hpp.
#include "mms_grad_exchange.hpp"

#include "caffe/syncedmem.hpp"

namespace caffe {
namespace mms {

namespace {
inline int allreduce(mms_exchange_t x, cudaStream_t s, long long n, float scale) { return mms_exchange_allreduce_f32(x, s, 0, 0, n, scale); }
inline int allreduce(mms_exchange_t x, cudaStream_t s, long long n, double scale) { return mms_exchange_allreduce_f64(x, s, 0, 0, n, scale); }
inline int adadelta(mms_exchange_t x, cudaStream_t s, long long n, float gs, const long long* e, const double* r, const double* d,
                    int ns, float mom, float delta) {
  return mms_exchange_adadelta_f32(x, s, 0, 0, n, gs, e, r, d, ns, mom, delta, 1);
}
inline int adadelta(mms_exchange_t x, cudaStream_t s, long long n, double gs, const long long* e, const double* r,
                    const double* d, int ns, double mom, double delta) {
  return mms_exchange_adadelta_f64(x, s, 0, 0, n, gs, e, r, d, ns, mom, delta, 1);
}
}  // namespace

template <typename Dtype>
GradExchange<Dtype>::GradExchange(const std::vector<Blob<Dtype>*>& params, int rank, int world)
    : x_(NULL), rank_(rank), world_(world), count_(0), data_(NULL), diff_(NULL) {
  // every blob on a 16-byte boundary, so that any run of blobs is a vector-aligned range of the flat buffers
  const long long vn = 16 / sizeof(Dtype);
  for (size_t i = 0; i < params.size(); ++i) {
    offsets_.push_back(count_);
    counts_.push_back(params[i]->count());
    count_ += (params[i]->count() + vn - 1) / vn * vn;
  }
  CHECK_GT(count_, 0) << "a net without learnable parameters has nothing to exchange";
  MMS_CAFFE_CHECK(mms_exchange_create(&x_, rank, world, count_, sizeof(Dtype), NULL));
  void *d = NULL, *g = NULL;
  MMS_CAFFE_CHECK(mms_exchange_buffers(x_, &d, &g));
  data_ = static_cast<Dtype*>(d);
  diff_ = static_cast<Dtype*>(g);
  for (size_t i = 0; i < params.size(); ++i) {
    // apply_buffers(net, data_, size_, copy) then replace_gpu / replace_gpu_diff (parallel.cpp:36-55, :110-115)
    CUDA_CHECK(cudaMemcpy(data_ + offsets_[i], params[i]->gpu_data(), sizeof(Dtype) * counts_[i], cudaMemcpyDefault));
    params[i]->data()->set_gpu_data(data_ + offsets_[i]);
    params[i]->diff()->set_gpu_data(diff_ + offsets_[i]);
  }
}

template <typename Dtype>
GradExchange<Dtype>::~GradExchange() {
  if (x_) mms_exchange_destroy(x_);
}

template <typename Dtype>
void GradExchange<Dtype>::Attach(const std::vector<GradExchange<Dtype>*>& all) {
  std::vector<void*> bases(all.size());
  for (size_t q = 0; q < all.size(); ++q) MMS_CAFFE_CHECK(mms_exchange_base(all[q]->x_, &bases[q]));
  int initial = 0;
  CUDA_CHECK(cudaGetDevice(&initial));
  for (size_t q = 0; q < all.size(); ++q) {
    cudaPointerAttributes at;
    CUDA_CHECK(cudaPointerGetAttributes(&at, bases[q]));
    CUDA_CHECK(cudaSetDevice(at.device));                    // peer access is enabled from the owner's device
    MMS_CAFFE_CHECK(mms_exchange_attach_ptrs(all[q]->x_, bases.data(), NULL));
  }
  CUDA_CHECK(cudaSetDevice(initial));
}

template <typename Dtype>
void GradExchange<Dtype>::on_start(cudaStream_t stream, bool sync) {
  MMS_CAFFE_CHECK(mms_exchange_broadcast(x_, stream, 0, 0));
  if (sync) MMS_CAFFE_CHECK(mms_exchange_check(x_, stream));
}

template <typename Dtype>
void GradExchange<Dtype>::on_gradients_ready(cudaStream_t stream, bool sync) {
  MMS_CAFFE_CHECK(allreduce(x_, stream, count_, Dtype(1) / Dtype(world_)));
  if (sync) MMS_CAFFE_CHECK(mms_exchange_check(x_, stream));
}

template <typename Dtype>
void GradExchange<Dtype>::on_gradients_ready_adadelta(const std::vector<float>& lr_mult, const std::vector<float>& decay_mult,
                                                      Dtype base_lr, Dtype momentum, Dtype delta, Dtype weight_decay,
                                                      int iter_size, cudaStream_t stream, bool sync) {
  CHECK_EQ(lr_mult.size(), offsets_.size());
  CHECK_EQ(decay_mult.size(), offsets_.size());
  CHECK_LE(offsets_.size(), 8u) << "at most 8 blobs per fused call: split the parameters into buckets";
  std::vector<long long> ends(offsets_.size());
  std::vector<double> rate(offsets_.size()), decay(offsets_.size());
  for (size_t i = 0; i < offsets_.size(); ++i) {
    ends[i] = i + 1 < offsets_.size() ? offsets_[i + 1] : count_;
    rate[i] = static_cast<double>(base_lr) * lr_mult[i];                 // sgd_solver.cpp:27-30, adadelta_solver.cpp:30
    decay[i] = static_cast<double>(weight_decay) * decay_mult[i];        // sgd_solver.cpp:148
  }
  MMS_CAFFE_CHECK(adadelta(x_, stream, count_, Dtype(1) / Dtype(world_ * iter_size), ends.data(), rate.data(), decay.data(),
                           static_cast<int>(ends.size()), momentum, delta));
  if (sync) MMS_CAFFE_CHECK(mms_exchange_check(x_, stream));
}

template class GradExchange<float>;
template class GradExchange<double>;

}  // namespace mms
}  // namespace caffe
