// Drop-in replacement of src/caffe/layers/pair_rank_loss_layer.{cpp,cu}: pairwise margin ranking loss
// plus the "similar pair" L1 term.  The reference's GPU forward runs four small kernels and then a
// host double loop over a D2H copy (pair_rank_loss_layer.cu:17-41); here forward is one fused kernel
// with a warp-shuffle reduction that leaves the scalar on the device.
#include <vector>

#include "caffe/layers/pair_rank_loss_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void PairRankLossLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  LossLayer<Dtype>::LayerSetUp(bottom, top);           // default loss_weight 1
  CHECK_EQ(bottom[0]->num(), bottom[1]->num());
  CHECK_EQ(bottom[0]->num(), bottom[2]->num());
  CHECK_EQ(bottom[0]->count(1), bottom[2]->count(1));
  CHECK_EQ(bottom[0]->count(1), bottom[1]->count(1));
  margin_ = static_cast<Dtype>(this->layer_param_.pair_rank_loss_param().margin());
  // the caches are sized here, once, exactly like the reference (pair_rank_loss_layer.cpp:21-22)
  ordered_diff_.Reshape(bottom[0]->num(), bottom[0]->channels(), 1, 1);
  similar_diff_.Reshape(bottom[0]->num(), bottom[0]->channels(), 1, 1);
}

template <typename Dtype>
void PairRankLossLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  MMS_CAFFE_CHECK(mms::pairrankloss_forward(mms::handle(), bottom[0]->gpu_data(), bottom[1]->gpu_data(),
                                            bottom[2]->gpu_data(), margin_, bottom[0]->count(),
                                            top[0]->mutable_gpu_data(), ordered_diff_.mutable_gpu_data(),
                                            similar_diff_.mutable_gpu_data()));
}

template <typename Dtype>
void PairRankLossLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                            const vector<Blob<Dtype>*>& bottom) {
  if (propagate_down[2]) LOG(FATAL) << this->type() << " Layer cannot backpropagate to label inputs.";
  const Dtype top_diff = top[0]->cpu_diff()[0];        // the loss weight (layer.hpp:414-428)
  MMS_CAFFE_CHECK(mms::pairrankloss_backward(mms::handle(), bottom[2]->gpu_data(), ordered_diff_.gpu_data(),
                                             similar_diff_.gpu_data(), top_diff, bottom[0]->count(),
                                             propagate_down[0] ? bottom[0]->mutable_gpu_diff() : NULL,
                                             propagate_down[1] ? bottom[1]->mutable_gpu_diff() : NULL));
}

template <typename Dtype>
void PairRankLossLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(PairRankLossLayer);
}
template <typename Dtype>
void PairRankLossLayer<Dtype>::Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(PairRankLossLayer);
}

INSTANTIATE_CLASS(PairRankLossLayer);
REGISTER_LAYER_CLASS(PairRankLoss);

}  // namespace caffe
