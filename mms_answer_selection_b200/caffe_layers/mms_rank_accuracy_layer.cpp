// Drop-in replacement of src/caffe/layers/rank_accuracy_layer.cpp.  The reference's class declares Forward_cpu only
// (include/caffe/layers/rank_accuracy_layer.hpp), so in Caffe::GPU mode Layer::Forward_gpu lands here (layer.hpp:344-348): the
// scores stay on the device -- bottom[i]->gpu_data() in, top[0]->mutable_gpu_data() out -- instead of being pulled
// to the host, bucketed in a std::map and std::sort-ed (rank_accuracy_layer.cpp:36-50).  No CPU path.
#include <vector>

#include "caffe/layers/rank_accuracy_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void RankAccuracyLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {}

template <typename Dtype>
void RankAccuracyLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(bottom[0]->count(), bottom[1]->count()) << "two pairs have the same dimension!.";
  CHECK_EQ(bottom[0]->count(), bottom[2]->count()) << "pair should have the same dimension with the label!.";
  top[0]->Reshape(vector<int>(0));                          // a scalar: 0 axes
}

template <typename Dtype>
void RankAccuracyLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  if (Caffe::mode() != Caffe::GPU) MMS_NO_CPU_PATH(RankAccuracyLayer);
  MMS_CAFFE_CHECK(mms::rank_accuracy(mms::handle(), bottom[0]->gpu_data(), bottom[1]->gpu_data(),
                                     bottom[2]->gpu_data(), bottom[0]->count(), top[0]->mutable_gpu_data()));
}

INSTANTIATE_CLASS(RankAccuracyLayer);
REGISTER_LAYER_CLASS(RankAccuracy);

}  // namespace caffe
