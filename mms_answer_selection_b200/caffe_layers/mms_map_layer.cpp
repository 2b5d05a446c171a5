// Drop-in replacement of src/caffe/layers/map_layer.cpp.  The reference's class declares Forward_cpu only
// (include/caffe/layers/map_layer.hpp), so in Caffe::GPU mode Layer::Forward_gpu lands here (layer.hpp:344-348): the
// scores stay on the device -- bottom[i]->gpu_data() in, top[0]->mutable_gpu_data() out -- instead of being pulled
// to the host, bucketed in a std::map and std::sort-ed (map_layer.cpp:41-100).  No CPU path.
#include <vector>

#include "caffe/layers/map_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void MAPLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  fixed_axis_ = this->layer_param_.map_param().fixed_axis();
}

template <typename Dtype>
void MAPLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_LE(fixed_axis_, bottom[0]->count() / bottom[1]->count())
      << "top_k must be less than or equal to the number of classes.";
  const int samples = bottom[0]->count(0, 1) * bottom[0]->count(2);
  CHECK_EQ(samples, bottom[1]->count()) << "Number of labels must match number of predictions; ";
  CHECK_EQ(samples, bottom[2]->count());
  top[0]->Reshape(vector<int>(0));                          // a scalar: 0 axes
}

template <typename Dtype>
void MAPLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  if (Caffe::mode() != Caffe::GPU) MMS_NO_CPU_PATH(MAPLayer);
  // score of sample i = bottom_data[i * (fixed_axis_ + 1) + fixed_axis_]  (map_layer.cpp:50)
  Dtype* out = top[0]->mutable_gpu_data();
  MMS_CAFFE_CHECK(mms::rank_map_mrr(mms::handle(), bottom[0]->gpu_data(), fixed_axis_ + 1, fixed_axis_,
                                    bottom[1]->gpu_data(), bottom[2]->gpu_data(), bottom[1]->count(),
                                    out, static_cast<Dtype*>(NULL)));
}

INSTANTIATE_CLASS(MAPLayer);
REGISTER_LAYER_CLASS(MAP);

}  // namespace caffe
