// Drop-in replacement of src/caffe/layers/fm_layer.{cpp,cu}: factorization-machine combination of the
// per-modality vectors, y_n = 1/2 sum_{j>=1}[(sum_k x_kj)^2 - sum_k x_kj^2] + sum_k x_k0 + b.  The
// reference's "GPU" methods call the CPU code (fm_layer.cu:14,20); here both passes are kernels.
#include <vector>

#include "caffe/layers/fm_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void FMLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  bias_term_ = this->layer_param_.fm_param().bias_term();
  if (bias_term_) {
    this->blobs_.resize(1);
    this->blobs_[0].reset(new Blob<Dtype>(vector<int>(1, 1)));
    this->blobs_[0]->mutable_cpu_data()[0] = Dtype(0);
  }
  this->param_propagate_down_.resize(this->blobs_.size(), true);
}

template <typename Dtype>
void FMLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  vector<int> top_shape(2);
  top_shape[0] = bottom[0]->num();
  top_shape[1] = 1;
  top[0]->Reshape(top_shape);
}

template <typename Dtype>
void FMLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  MMS_CAFFE_CHECK(mms::fm_forward(mms::handle(), bottom[0]->gpu_data(), bias_term_ ? this->blobs_[0]->gpu_data() : NULL,
                                  top[0]->mutable_gpu_data(), bottom[0]->num(), bottom[0]->channels(),
                                  bottom[0]->height()));
}

template <typename Dtype>
void FMLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                  const vector<Blob<Dtype>*>& bottom) {
  Dtype* db = (bias_term_ && this->param_propagate_down_[0]) ? this->blobs_[0]->mutable_gpu_diff() : NULL;
  MMS_CAFFE_CHECK(mms::fm_backward(mms::handle(), bottom[0]->gpu_data(), top[0]->gpu_diff(),
                                   bottom[0]->mutable_gpu_diff(), db, bottom[0]->num(), bottom[0]->channels(),
                                   bottom[0]->height(), propagate_down[0] ? 1 : 0));
}

template <typename Dtype>
void FMLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(FMLayer);
}
template <typename Dtype>
void FMLayer<Dtype>::Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(FMLayer);
}

INSTANTIATE_CLASS(FMLayer);
REGISTER_LAYER_CLASS(FM);

}  // namespace caffe
