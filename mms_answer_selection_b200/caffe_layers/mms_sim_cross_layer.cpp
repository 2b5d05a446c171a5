// Drop-in replacement of src/caffe/layers/sim_cross_layer.{cpp,cu}: word-by-word similarity tensor
// between two embedded sentences, dist_mode 0 cosine / 1 1/(1+euclid) / 2 learned bilinear
// Q M_k A^T + B_k for mesure_count matrices.  Blobs as in the reference: M [mesure_count, D, D],
// B [mesure_count, Lq, La] (bias_term).  In the reference mode 2 has no GPU code at all
// (sim_cross_layer.cu:187-189 calls Forward_cpu); here every mode runs in libmms_b200.so.
#include <vector>

#include "caffe/filler.hpp"
#include "caffe/layers/sim_cross_layer.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {

template <typename Dtype>
void SimCrossLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(bottom.size(), 2);
  CHECK_EQ(bottom[0]->num(), bottom[1]->num());
  CHECK_EQ(bottom[0]->height(), bottom[1]->height());
  const SimCrossParameter& sp = this->layer_param_.sim_cross_param();
  dist_mode_ = sp.dist_mode();
  if (dist_mode_ != 2) return;
  // like the reference (sim_cross_layer.cpp:17-46) the parameters are (re)created unconditionally
  const int mc = sp.mesure_count();
  this->blobs_.resize(sp.bias_term() ? 2 : 1);
  vector<int> shape(3);
  shape[0] = mc; shape[1] = bottom[0]->height(); shape[2] = bottom[1]->height();
  this->blobs_[0].reset(new Blob<Dtype>(shape));
  shared_ptr<Filler<Dtype> > wf(GetFiller<Dtype>(sp.weight_filler()));
  wf->Fill(this->blobs_[0].get());
  if (sp.bias_term()) {
    shape[1] = bottom[0]->channels(); shape[2] = bottom[1]->channels();
    this->blobs_[1].reset(new Blob<Dtype>(shape));
    shared_ptr<Filler<Dtype> > bf(GetFiller<Dtype>(sp.bias_filler()));
    bf->Fill(this->blobs_[1].get());
  }
}

template <typename Dtype>
void SimCrossLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const int N = bottom[0]->num(), Lq = bottom[0]->channels(), La = bottom[1]->channels();
  vector<int> top_shape(4);
  top_shape[0] = N;
  top_shape[1] = dist_mode_ == 2 ? this->layer_param_.sim_cross_param().mesure_count() : 1;
  top_shape[2] = Lq;
  top_shape[3] = La;
  top[0]->Reshape(top_shape);
  if (dist_mode_ == 0) {           // cached row norms of both sentences (data0_norm_/data1_norm_)
    vector<int> sz(2);
    sz[0] = N; sz[1] = Lq;
    data0_norm_.Reshape(sz);
    sz[1] = La;
    data1_norm_.Reshape(sz);
  }
  // measure_temp0_/measure_temp1_ (the reference's per-pair gemm scratch) stay empty: the
  // contraction intermediates live in the workspace of libmms_b200.so.
}

namespace {
template <typename Dtype>
struct SimCrossDims {
  int N, Lq, La, D, mc;
  SimCrossDims(const vector<Blob<Dtype>*>& bottom, int mode, int mesure_count)
      : N(bottom[0]->num()), Lq(bottom[0]->channels()), La(bottom[1]->channels()), D(bottom[0]->height()),
        mc(mode == 2 ? mesure_count : 1) {}
};
}  // namespace

template <typename Dtype>
void SimCrossLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const SimCrossParameter& sp = this->layer_param_.sim_cross_param();
  const SimCrossDims<Dtype> d(bottom, dist_mode_, sp.mesure_count());
  const bool bilinear = dist_mode_ == 2;
  const Dtype* M = bilinear ? this->blobs_[0]->gpu_data() : NULL;
  const Dtype* B = (bilinear && sp.bias_term()) ? this->blobs_[1]->gpu_data() : NULL;
  Dtype* n0 = dist_mode_ == 0 ? data0_norm_.mutable_gpu_data() : NULL;
  Dtype* n1 = dist_mode_ == 0 ? data1_norm_.mutable_gpu_data() : NULL;
  MMS_CAFFE_CHECK(mms::simcross_forward(mms::handle(), dist_mode_, bottom[0]->gpu_data(), bottom[1]->gpu_data(), M, B,
                                        top[0]->mutable_gpu_data(), n0, n1, d.N, d.Lq, d.La, d.D, d.mc));
}

template <typename Dtype>
void SimCrossLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                        const vector<Blob<Dtype>*>& bottom) {
  const SimCrossParameter& sp = this->layer_param_.sim_cross_param();
  const SimCrossDims<Dtype> d(bottom, dist_mode_, sp.mesure_count());
  const bool bilinear = dist_mode_ == 2;
  const Dtype* M = bilinear ? this->blobs_[0]->gpu_data() : NULL;
  Dtype* dM = bilinear ? this->blobs_[0]->mutable_gpu_diff() : NULL;
  Dtype* dB = (bilinear && sp.bias_term()) ? this->blobs_[1]->mutable_gpu_diff() : NULL;
  const Dtype* n0 = dist_mode_ == 0 ? data0_norm_.gpu_data() : NULL;
  const Dtype* n1 = dist_mode_ == 0 ? data1_norm_.gpu_data() : NULL;
  // diff semantics (bottom diffs overwritten, dM overwritten, dB accumulated) are those of
  // sim_cross_layer.cpp:166-307 and are implemented behind the C-ABI.
  MMS_CAFFE_CHECK(mms::simcross_backward(mms::handle(), dist_mode_, bottom[0]->gpu_data(), bottom[1]->gpu_data(), M,
                                         top[0]->gpu_data(), top[0]->gpu_diff(), n0, n1,
                                         bottom[0]->mutable_gpu_diff(), bottom[1]->mutable_gpu_diff(), dM, dB,
                                         d.N, d.Lq, d.La, d.D, d.mc, propagate_down[0] ? 1 : 0,
                                         propagate_down[1] ? 1 : 0));
}

template <typename Dtype>
void SimCrossLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(SimCrossLayer);
}
template <typename Dtype>
void SimCrossLayer<Dtype>::Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {
  MMS_NO_CPU_PATH(SimCrossLayer);
}

INSTANTIATE_CLASS(SimCrossLayer);
REGISTER_LAYER_CLASS(SimCross);

}  // namespace caffe
