// Drop-in replacement of src/caffe/layers/embed_layer.{cpp,cu} for the B200: same class (declared in
// the reference's include/caffe/layers/embed_layer.hpp), same blobs ([input_dim, num_output] table and
// optional [num_output] bias), same prototxt fields; the gather and the gradient scatter-add run in
// libmms_b200.so.  The bias multiplier of the reference (a vector of ones for its K=1 gemm,
// embed_layer.cpp:126-131) is not needed: the bias add is fused into the gather kernel.
#include <vector>

#include "caffe/filler.hpp"
#include "caffe/layers/embed_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void EmbedLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const EmbedParameter& ep = this->layer_param_.embed_param();
  N_ = ep.num_output();
  CHECK_GT(N_, 0) << "EmbedLayer num_output must be positive.";
  K_ = ep.input_dim();
  CHECK_GT(K_, 0) << "EmbedLayer input_dim must be positive.";
  bias_term_ = ep.bias_term();
  if (this->blobs_.size() > 0) {
    LOG(INFO) << "Skipping parameter initialization";
  } else {
    this->blobs_.resize(bias_term_ ? 2 : 1);
    vector<int> table_shape(2);
    table_shape[0] = K_;
    table_shape[1] = N_;
    this->blobs_[0].reset(new Blob<Dtype>(table_shape));
    shared_ptr<Filler<Dtype> > wf(GetFiller<Dtype>(ep.weight_filler()));
    wf->Fill(this->blobs_[0].get());
    if (bias_term_) {
      this->blobs_[1].reset(new Blob<Dtype>(vector<int>(1, N_)));
      shared_ptr<Filler<Dtype> > bf(GetFiller<Dtype>(ep.bias_filler()));
      bf->Fill(this->blobs_[1].get());
    }
    // embed_param.weight_source: pre-trained vectors over the filler's values (embed_layer.cpp:46-113); the loader is
    // host code of the library and writes the blob's CPU copy, the table reaches the GPU through the blob's own sync
    if (ep.has_weight_source() && !ep.weight_source().empty()) {
      long long loaded = 0;
      MMS_CAFFE_CHECK(mms::load_weight_source(ep.weight_source().c_str(), this->blobs_[0]->mutable_cpu_data(), K_, N_,
                                              &loaded));
      LOG(INFO) << "loaded " << loaded << " words from " << ep.weight_source();
    } else {
      LOG(INFO) << "Not loading word embeddings";
    }
  }
  this->param_propagate_down_.resize(this->blobs_.size(), true);
}

template <typename Dtype>
void EmbedLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  M_ = bottom[0]->count();                       // one table row per input index
  vector<int> top_shape = bottom[0]->shape();
  top_shape.push_back(N_);
  top[0]->Reshape(top_shape);
}

template <typename Dtype>
void EmbedLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const Dtype* bias = bias_term_ ? this->blobs_[1]->gpu_data() : NULL;
  MMS_CAFFE_CHECK(mms::embed_forward(mms::handle(), bottom[0]->gpu_data(), this->blobs_[0]->gpu_data(), bias,
                                     top[0]->mutable_gpu_data(), M_, N_, K_));
#ifdef DEBUG
  // the reference DCHECKs every index (embed_layer.cpp:142-146); the kernel flags them on the device
  MMS_CAFFE_CHECK(mms_check_faults(mms::handle()));
#endif
}

template <typename Dtype>
void EmbedLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                     const vector<Blob<Dtype>*>& bottom) {
  CHECK(!propagate_down[0]) << "Can't backpropagate to EmbedLayer input.";
  Dtype* dW = this->param_propagate_down_[0] ? this->blobs_[0]->mutable_gpu_diff() : NULL;
  Dtype* db = (bias_term_ && this->param_propagate_down_[1]) ? this->blobs_[1]->mutable_gpu_diff() : NULL;
  MMS_CAFFE_CHECK(mms::embed_backward(mms::handle(), bottom[0]->gpu_data(), top[0]->gpu_diff(), dW, db, M_, N_, K_));
}

template <typename Dtype>
void EmbedLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(EmbedLayer);
}
template <typename Dtype>
void EmbedLayer<Dtype>::Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(EmbedLayer);
}

INSTANTIATE_CLASS(EmbedLayer);
REGISTER_LAYER_CLASS(Embed);

}  // namespace caffe
