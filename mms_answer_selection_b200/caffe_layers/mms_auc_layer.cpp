// Drop-in replacement of src/caffe/layers/auc_layer.cpp.  The reference's class declares Forward_cpu only
// (include/caffe/layers/auc_layer.hpp), so in Caffe::GPU mode Layer::Forward_gpu lands here (layer.hpp:344-348): the
// scores stay on the device -- bottom[i]->gpu_data() in, top[0]->mutable_gpu_data() out -- instead of being pulled
// to the host, bucketed in a std::map and std::sort-ed (auc_layer.cpp:47-136).  No CPU path.
#include <vector>

#include "caffe/layers/auc_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void AUCLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  fixed_axis_ = this->layer_param_.auc_param().fixed_axis();
  has_ignore_label_ = this->layer_param_.auc_param().has_ignore_label();
  if (has_ignore_label_) ignore_label_ = this->layer_param_.auc_param().ignore_label();
}

template <typename Dtype>
void AUCLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_LE(fixed_axis_, bottom[0]->count() / bottom[1]->count())
      << "top_k must be less than or equal to the number of classes.";
  label_axis_ = bottom[0]->CanonicalAxisIndex(this->layer_param_.auc_param().axis());
  outer_num_ = bottom[0]->count(0, label_axis_);
  inner_num_ = bottom[0]->count(label_axis_ + 1);
  CHECK_EQ(outer_num_ * inner_num_, bottom[1]->count()) << "Number of labels must match number of predictions; ";
  // the device path takes the (N, C) predictions the reference nets feed it (softmax output, label axis 1)
  CHECK_EQ(inner_num_, 1) << "AUCLayer (mms_b200): predictions must be (N, C)";
  top[0]->Reshape(vector<int>(0));                          // a scalar: 0 axes
}

template <typename Dtype>
void AUCLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  if (Caffe::mode() != Caffe::GPU) MMS_NO_CPU_PATH(AUCLayer);
  // score of sample i = bottom_data[i * dim + fixed_axis_]  (auc_layer.cpp:76 with inner_num_ == 1)
  MMS_CAFFE_CHECK(mms::rank_auc(mms::handle(), bottom[0]->gpu_data(), bottom[0]->count() / outer_num_, fixed_axis_,
                                bottom[1]->gpu_data(), outer_num_, has_ignore_label_ ? 1 : 0,
                                has_ignore_label_ ? ignore_label_ : 0, top[0]->mutable_gpu_data()));
}

INSTANTIATE_CLASS(AUCLayer);
REGISTER_LAYER_CLASS(AUC);

}  // namespace caffe
