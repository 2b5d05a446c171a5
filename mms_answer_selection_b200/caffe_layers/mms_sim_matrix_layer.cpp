// Drop-in replacement of src/caffe/layers/sim_matrix_layer.{cpp,cu}: sentence-level bilinear score
// s_n = q_n^T W a_n with W [K1, K2].  As in the reference, forward leaves T = q W in bottom[1]'s DIFF
// buffer (sim_matrix_layer.cpp:58, .cu:25); backward is three batched GEMMs instead of the
// reference's N host-side sger / gemv calls (sim_matrix_layer.cpp:73-93).
#include <vector>

#include "caffe/filler.hpp"
#include "caffe/layers/sim_matrix_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void SimMatrixLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(bottom[0]->num(), bottom[1]->num());
  K1_ = bottom[0]->count(1);
  K2_ = bottom[1]->count(1);
  if (this->blobs_.size() > 0) {
    LOG(INFO) << "Skipping parameter initialization";
  } else {
    this->blobs_.resize(1);
    vector<int> shape(2);
    shape[0] = K1_; shape[1] = K2_;
    this->blobs_[0].reset(new Blob<Dtype>(shape));
    shared_ptr<Filler<Dtype> > wf(GetFiller<Dtype>(this->layer_param_.sim_matrix_param().weight_filler()));
    wf->Fill(this->blobs_[0].get());
  }
  this->param_propagate_down_.resize(this->blobs_.size(), true);
}

template <typename Dtype>
void SimMatrixLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(K1_, bottom[0]->count(1)) << "Input size incompatible with inner product parameters.";
  CHECK_EQ(K2_, bottom[1]->count(1)) << "Input size incompatible with inner product parameters.";
  M_ = bottom[0]->count(0, 1);
  vector<int> top_shape(2);
  top_shape[0] = bottom[0]->shape(0);
  top_shape[1] = 1;
  top[0]->Reshape(top_shape);
}

template <typename Dtype>
void SimMatrixLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  MMS_CAFFE_CHECK(mms::simmatrix_forward(mms::handle(), bottom[0]->gpu_data(), bottom[1]->gpu_data(),
                                         this->blobs_[0]->gpu_data(), top[0]->mutable_gpu_data(),
                                         bottom[1]->mutable_gpu_diff(), M_, K1_, K2_));
}

template <typename Dtype>
void SimMatrixLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                         const vector<Blob<Dtype>*>& bottom) {
  const bool pw = this->param_propagate_down_[0];
  MMS_CAFFE_CHECK(mms::simmatrix_backward(
      mms::handle(), bottom[0]->gpu_data(), bottom[1]->gpu_data(), this->blobs_[0]->gpu_data(), top[0]->gpu_diff(),
      pw ? this->blobs_[0]->mutable_gpu_diff() : NULL, propagate_down[0] ? bottom[0]->mutable_gpu_diff() : NULL,
      propagate_down[1] ? bottom[1]->mutable_gpu_diff() : NULL, M_, K1_, K2_, pw ? 1 : 0, propagate_down[0] ? 1 : 0,
      propagate_down[1] ? 1 : 0));
}

template <typename Dtype>
void SimMatrixLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(SimMatrixLayer);
}
template <typename Dtype>
void SimMatrixLayer<Dtype>::Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(SimMatrixLayer);
}

INSTANTIATE_CLASS(SimMatrixLayer);
REGISTER_LAYER_CLASS(SimMatrix);

}  // namespace caffe
