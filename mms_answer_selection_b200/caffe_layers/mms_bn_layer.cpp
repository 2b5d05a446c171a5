// Drop-in replacement of src/caffe/layers/bn_layer.{cpp,cu}: the fork's own batch normalisation (type "BN",
// include/caffe/layers/batch_norm_v0_layer.hpp): blobs scale, shift, running mean, running variance, each (1,C,1,1).
// The reference composes each pass from ~20 BLAS calls over N*C*H*W buffers (bn_layer.cpp:121-257, :261-384); here a
// pass is two or three kernels behind mms_bn_forward / mms_bn_backward.  The members keep the roles they have in the
// reference: buffer_blob_'s diff holds x_norm (:225-227), batch_variance_ the standard deviation (:206-210),
// batch_mean_ the mean that was subtracted.
#include <vector>

#include "caffe/filler.hpp"
#include "caffe/layers/batch_norm_v0_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void BNLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_NE(top[0], bottom[0]) << this->type() << " Layer does not allow in-place computation.";
  var_eps_ = 1e-9;
  bn_memory_ = this->layer_param_.bn_param().bn_memory();
  C_ = bottom[0]->channels();
  if (this->blobs_.size() > 0) {
    LOG(INFO) << "Skipping parameter initialization";
  } else {
    this->blobs_.resize(4);
    for (int i = 0; i < 4; ++i) this->blobs_[i].reset(new Blob<Dtype>(1, C_, 1, 1));
    shared_ptr<Filler<Dtype> > scale_filler(GetFiller<Dtype>(this->layer_param_.bn_param().scale_filler()));
    scale_filler->Fill(this->blobs_[0].get());
    shared_ptr<Filler<Dtype> > shift_filler(GetFiller<Dtype>(this->layer_param_.bn_param().shift_filler()));
    shift_filler->Fill(this->blobs_[1].get());
    caffe_set(C_, Dtype(0), this->blobs_[2]->mutable_cpu_data());       // running mean and variance start at zero
    caffe_set(C_, Dtype(0), this->blobs_[3]->mutable_cpu_data());
  }
  this->param_propagate_down_.resize(this->blobs_.size(), true);
}

template <typename Dtype>
void BNLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  N_ = bottom[0]->num();
  C_ = bottom[0]->channels();
  H_ = bottom[0]->height();
  W_ = bottom[0]->width();
  top[0]->Reshape(N_, C_, H_, W_);
  buffer_blob_.Reshape(N_, C_, H_, W_);
  batch_mean_.Reshape(1, C_, 1, 1);
  batch_variance_.Reshape(1, C_, 1, 1);
  this->param_propagate_down_.resize(this->blobs_.size(), true);
}

template <typename Dtype>
void BNLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  MMS_CAFFE_CHECK(mms::bn_forward(mms::handle(), bottom[0]->gpu_data(), this->blobs_[0]->gpu_data(),
                                  this->blobs_[1]->gpu_data(), this->blobs_[2]->mutable_gpu_data(),
                                  this->blobs_[3]->mutable_gpu_data(), top[0]->mutable_gpu_data(),
                                  buffer_blob_.mutable_gpu_diff(), batch_mean_.mutable_gpu_data(),
                                  batch_variance_.mutable_gpu_data(), N_, C_, H_ * W_, this->phase_ == TRAIN ? 1 : 0,
                                  bn_memory_, var_eps_));
}

template <typename Dtype>
void BNLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                  const vector<Blob<Dtype>*>& bottom) {
  // the reference computes all three gradients whatever propagate_down says, and OVERWRITES the two param diffs
  MMS_CAFFE_CHECK(mms::bn_backward(mms::handle(), top[0]->gpu_diff(), buffer_blob_.gpu_diff(), this->blobs_[0]->gpu_data(),
                                   batch_variance_.gpu_data(), this->blobs_[0]->mutable_gpu_diff(),
                                   this->blobs_[1]->mutable_gpu_diff(), bottom[0]->mutable_gpu_diff(), N_, C_, H_ * W_));
}

template <typename Dtype>
void BNLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(BNLayer);
}
template <typename Dtype>
void BNLayer<Dtype>::Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(BNLayer);
}

INSTANTIATE_CLASS(BNLayer);
REGISTER_LAYER_CLASS(BN);

}  // namespace caffe
