// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle, never of the product.
//
// C-ABI around the reference's OWN layer code (compiled in place from
// /root/reference by oracle/Makefile, CPU_ONLY).  Layers are created through the
// reference's LayerRegistry::CreateLayer (include/caffe/layer_factory.hpp:73-81)
// and driven through Layer::SetUp / Forward / Backward (include/caffe/layer.hpp:
// 67-77, 451-500) -- the same entry points caffe::Net uses (src/caffe/net.cpp:
// 139, 541, 586).  Loaded from Python with ctypes (oracle/refbind.py) by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "caffe/blob.hpp"
#include "caffe/layer.hpp"
#include "caffe/layer_factory.hpp"
#include "caffe/layers/pair_rank_loss_layer.hpp"

#ifndef MMS_DROPIN
#include "caffe/layers/conv_layer.hpp"
#include "caffe/layers/pooling_layer.hpp"
#include "caffe/layers/tanh_layer.hpp"
namespace caffe {
// Convolution / Pooling / TanH are registered by creator functions in layer_factory.cpp (engine selection,
// :38-110, :220-239), which is not compiled here (cuDNN, Python layers): register the CAFFE-engine classes directly.
template <typename Dtype> shared_ptr<Layer<Dtype> > MmsRefConv(const LayerParameter& p) {
  return shared_ptr<Layer<Dtype> >(new ConvolutionLayer<Dtype>(p));
}
template <typename Dtype> shared_ptr<Layer<Dtype> > MmsRefPool(const LayerParameter& p) {
  return shared_ptr<Layer<Dtype> >(new PoolingLayer<Dtype>(p));
}
template <typename Dtype> shared_ptr<Layer<Dtype> > MmsRefTanH(const LayerParameter& p) {
  return shared_ptr<Layer<Dtype> >(new TanHLayer<Dtype>(p));
}
REGISTER_LAYER_CREATOR(Convolution, MmsRefConv);
REGISTER_LAYER_CREATOR(Pooling, MmsRefPool);
REGISTER_LAYER_CREATOR(TanH, MmsRefTanH);
// pair_rank_loss_layer.cpp:86-87 has no STUB_GPU, so a CPU_ONLY link lacks the
// Forward_gpu/Backward_gpu bodies its header declares; supply the stubs here.
STUB_GPU(PairRankLossLayer);
template class PairRankLossLayer<float>;
template class PairRankLossLayer<double>;
}  // namespace caffe
#else
// Built a second time with -DMMS_DROPIN (no CPU_ONLY) into oracle/_ref/libmms_dropin.so: the same
// harness then drives the PRODUCT's drop-in layers (mms_answer_selection_b200/caffe_layers/) through
// the reference's Layer API in Caffe::GPU mode -- see dropin_runtime.cpp.
#include <cuda_runtime.h>
#include "mms_grad_exchange.hpp"
#endif

namespace {

using caffe::Blob;
using caffe::Layer;
using caffe::LayerParameter;
using caffe::shared_ptr;
using std::vector;

struct SessionBase {
  virtual ~SessionBase() {}
  LayerParameter param;
  int dtype = 0;
  std::string error;
};

template <typename Dtype>
struct Session : public SessionBase {
  vector<shared_ptr<Blob<Dtype> > > bottom_own, top_own;
  vector<Blob<Dtype>*> bottom, top;
  shared_ptr<Layer<Dtype> > layer;
};

thread_local std::string g_last_error;

template <typename F>
int guarded(SessionBase* s, F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    if (s) s->error = e.what();
    return 1;
  }
}

caffe::FillerParameter* filler_of(LayerParameter* p, const std::string& which) {
  const std::string& t = p->type();
  if (t == "SimCross") {
    return which == "weight_filler" ? p->mutable_sim_cross_param()->mutable_weight_filler()
                                    : p->mutable_sim_cross_param()->mutable_bias_filler();
  } else if (t == "SimMatrix") {
    return p->mutable_sim_matrix_param()->mutable_weight_filler();
  } else if (t == "Convolution") {
    return which == "weight_filler" ? p->mutable_convolution_param()->mutable_weight_filler()
                                    : p->mutable_convolution_param()->mutable_bias_filler();
  } else if (t == "BN") {
    return which == "scale_filler" ? p->mutable_bn_param()->mutable_scale_filler()
                                   : p->mutable_bn_param()->mutable_shift_filler();
  } else if (t == "Embed") {
    return which == "weight_filler" ? p->mutable_embed_param()->mutable_weight_filler()
                                    : p->mutable_embed_param()->mutable_bias_filler();
  }
  return nullptr;
}

#define DISPATCH(s, ...)                                         \
  do {                                                           \
    if ((s)->dtype == 0) {                                       \
      auto* S = static_cast<Session<float>*>(s);                 \
      typedef float Dtype;                                       \
      (void)sizeof(Dtype);                                       \
      __VA_ARGS__;                                                   \
    } else {                                                     \
      auto* S = static_cast<Session<double>*>(s);                \
      typedef double Dtype;                                      \
      (void)sizeof(Dtype);                                       \
      __VA_ARGS__;                                                   \
    }                                                            \
  } while (0)

}  // namespace

extern "C" {

const char* mmsref_last_error() { return g_last_error.c_str(); }

#ifndef MMS_DROPIN
void mmsref_set_blas_threads(int n) { openblas_set_num_threads(n); }
int mmsref_get_blas_threads() { return openblas_get_num_threads(); }
int mmsref_is_dropin() { return 0; }
#else
void mmsref_set_blas_threads(int) {}
int mmsref_get_blas_threads() { return 0; }
int mmsref_is_dropin() { return 1; }
// 0 = Caffe::CPU (every drop-in layer then aborts: there is no CPU path), 1 = Caffe::GPU
void mmsref_set_mode(int gpu) { caffe::Caffe::set_mode(gpu ? caffe::Caffe::GPU : caffe::Caffe::CPU); }
#endif

void* mmsref_create(const char* type, int dtype) {
  SessionBase* s = dtype == 0 ? static_cast<SessionBase*>(new Session<float>())
                              : static_cast<SessionBase*>(new Session<double>());
  s->dtype = dtype;
  s->param.set_type(type);
  s->param.set_name(std::string("ref_") + type);
  return s;
}

void mmsref_destroy(void* h) { delete static_cast<SessionBase*>(h); }

// Integer/boolean prototxt fields.  Returns 0 on success, 2 for an unknown key.
int mmsref_set_i(void* h, const char* key, long long v) {
  SessionBase* s = static_cast<SessionBase*>(h);
  const std::string k(key);
  if (k == "dist_mode") s->param.mutable_sim_cross_param()->set_dist_mode(static_cast<int>(v));
  else if (k == "mesure_count") s->param.mutable_sim_cross_param()->set_mesure_count(static_cast<int>(v));
  else if (k == "sim_cross.bias_term") s->param.mutable_sim_cross_param()->set_bias_term(v != 0);
  else if (k == "num_output") s->param.mutable_embed_param()->set_num_output(static_cast<unsigned>(v));
  else if (k == "input_dim") s->param.mutable_embed_param()->set_input_dim(static_cast<unsigned>(v));
  else if (k == "embed.bias_term") s->param.mutable_embed_param()->set_bias_term(v != 0);
  else if (k == "fm.bias_term") s->param.mutable_fm_param()->set_bias_term(v != 0);
  else if (k == "phase") s->param.set_phase(v ? caffe::TEST : caffe::TRAIN);
  else if (k == "conv.num_output") s->param.mutable_convolution_param()->set_num_output(static_cast<unsigned>(v));
  else if (k == "conv.kernel_h") s->param.mutable_convolution_param()->set_kernel_h(static_cast<unsigned>(v));
  else if (k == "conv.kernel_w") s->param.mutable_convolution_param()->set_kernel_w(static_cast<unsigned>(v));
  else if (k == "conv.stride") s->param.mutable_convolution_param()->add_stride(static_cast<unsigned>(v));
  else if (k == "conv.pad") s->param.mutable_convolution_param()->add_pad(static_cast<unsigned>(v));
  else if (k == "conv.group") s->param.mutable_convolution_param()->set_group(static_cast<unsigned>(v));
  else if (k == "conv.bias_term") s->param.mutable_convolution_param()->set_bias_term(v != 0);
  else if (k == "pool.method") s->param.mutable_pooling_param()->set_pool(static_cast<caffe::PoolingParameter_PoolMethod>(v));
  else if (k == "pool.kernel_h") s->param.mutable_pooling_param()->set_kernel_h(static_cast<unsigned>(v));
  else if (k == "pool.kernel_w") s->param.mutable_pooling_param()->set_kernel_w(static_cast<unsigned>(v));
  else if (k == "pool.stride_h") s->param.mutable_pooling_param()->set_stride_h(static_cast<unsigned>(v));
  else if (k == "pool.stride_w") s->param.mutable_pooling_param()->set_stride_w(static_cast<unsigned>(v));
  else if (k == "map.fixed_axis") s->param.mutable_map_param()->set_fixed_axis(static_cast<int>(v));
  else if (k == "mrr.fixed_axis") s->param.mutable_mrr_param()->set_fixed_axis(static_cast<int>(v));
  else if (k == "auc.fixed_axis") s->param.mutable_auc_param()->set_fixed_axis(static_cast<int>(v));
  else if (k == "auc.axis") s->param.mutable_auc_param()->set_axis(static_cast<int>(v));
  else if (k == "auc.ignore_label") s->param.mutable_auc_param()->set_ignore_label(static_cast<int>(v));
  else return 2;
  return 0;
}

// Float prototxt fields, incl. "<weight_filler|bias_filler>.<value|min|max|mean|std>".
int mmsref_set_f(void* h, const char* key, double v) {
  SessionBase* s = static_cast<SessionBase*>(h);
  const std::string k(key);
  if (k == "margin") { s->param.mutable_pair_rank_loss_param()->set_margin(static_cast<float>(v)); return 0; }
  if (k == "loss_weight") { s->param.add_loss_weight(static_cast<float>(v)); return 0; }
  if (k == "bn_memory") { s->param.mutable_bn_param()->set_bn_memory(static_cast<float>(v)); return 0; }
  const size_t dot = k.find('.');
  if (dot == std::string::npos) return 2;
  caffe::FillerParameter* f = filler_of(&s->param, k.substr(0, dot));
  if (!f) return 2;
  const std::string field = k.substr(dot + 1);
  if (field == "value") f->set_value(static_cast<float>(v));
  else if (field == "min") f->set_min(static_cast<float>(v));
  else if (field == "max") f->set_max(static_cast<float>(v));
  else if (field == "mean") f->set_mean(static_cast<float>(v));
  else if (field == "std") f->set_std(static_cast<float>(v));
  else return 2;
  return 0;
}

// String prototxt fields: "<weight_filler|bias_filler>.type", "weight_source".
int mmsref_set_s(void* h, const char* key, const char* v) {
  SessionBase* s = static_cast<SessionBase*>(h);
  const std::string k(key);
  if (k == "weight_source") { s->param.mutable_embed_param()->set_weight_source(v); return 0; }
  const size_t dot = k.find('.');
  if (dot == std::string::npos) return 2;
  caffe::FillerParameter* f = filler_of(&s->param, k.substr(0, dot));
  if (!f || k.substr(dot + 1) != "type") return 2;
  f->set_type(v);
  return 0;
}

void mmsref_seed(unsigned int seed) { caffe::Caffe::set_random_seed(seed); }

int mmsref_add_bottom(void* h, int ndim, const int* shape) {
  SessionBase* s = static_cast<SessionBase*>(h);
  int idx = -1;
  int rc = guarded(s, [&] {
    vector<int> sh(shape, shape + ndim);
    DISPATCH(s, {
      S->bottom_own.push_back(shared_ptr<Blob<Dtype> >(new Blob<Dtype>(sh)));
      S->bottom.push_back(S->bottom_own.back().get());
      idx = static_cast<int>(S->bottom.size()) - 1;
    });
  });
  return rc ? -1 : idx;
}

int mmsref_reshape_bottom(void* h, int i, int ndim, const int* shape) {
  SessionBase* s = static_cast<SessionBase*>(h);
  return guarded(s, [&] {
    vector<int> sh(shape, shape + ndim);
    DISPATCH(s, S->bottom[i]->Reshape(sh));
  });
}

int mmsref_setup(void* h, int num_top) {
  SessionBase* s = static_cast<SessionBase*>(h);
  return guarded(s, [&] {
    DISPATCH(s, {
      for (int t = 0; t < num_top; ++t) {
        S->top_own.push_back(shared_ptr<Blob<Dtype> >(new Blob<Dtype>()));
        S->top.push_back(S->top_own.back().get());
      }
      S->layer = caffe::LayerRegistry<Dtype>::CreateLayer(S->param);
      S->layer->SetUp(S->bottom, S->top);
    });
  });
}

int mmsref_num_blobs(void* h) {
  SessionBase* s = static_cast<SessionBase*>(h);
  int n = 0;
  DISPATCH(s, n = static_cast<int>(S->layer->blobs().size()));
  return n;
}

// kind: 0 bottom, 1 top, 2 param blob.  Writes up to 8 dims; returns ndim.
int mmsref_shape(void* h, int kind, int i, int* shape_out) {
  SessionBase* s = static_cast<SessionBase*>(h);
  int nd = 0;
  DISPATCH(s, {
    const Blob<Dtype>* b = kind == 0 ? S->bottom[i] : kind == 1 ? S->top[i] : S->layer->blobs()[i].get();
    nd = b->num_axes();
    for (int a = 0; a < nd && a < 8; ++a) shape_out[a] = b->shape(a);
  });
  return nd;
}

// what: 0 data, 1 diff.  Buffers hold count() elements of the session dtype.
int mmsref_write(void* h, int kind, int i, int what, const void* src) {
  SessionBase* s = static_cast<SessionBase*>(h);
  return guarded(s, [&] {
    DISPATCH(s, {
      Blob<Dtype>* b = kind == 0 ? S->bottom[i] : kind == 1 ? S->top[i] : S->layer->blobs()[i].get();
      Dtype* dst = what == 0 ? b->mutable_cpu_data() : b->mutable_cpu_diff();
      std::memcpy(dst, src, sizeof(Dtype) * b->count());
    });
  });
}

int mmsref_read(void* h, int kind, int i, int what, void* dst) {
  SessionBase* s = static_cast<SessionBase*>(h);
  return guarded(s, [&] {
    DISPATCH(s, {
      const Blob<Dtype>* b = kind == 0 ? S->bottom[i] : kind == 1 ? S->top[i] : S->layer->blobs()[i].get();
      const Dtype* src = what == 0 ? b->cpu_data() : b->cpu_diff();
      std::memcpy(dst, src, sizeof(Dtype) * b->count());
    });
  });
}

int mmsref_forward(void* h, double* loss_out) {
  SessionBase* s = static_cast<SessionBase*>(h);
  return guarded(s, [&] {
    DISPATCH(s, {
      Dtype loss = S->layer->Forward(S->bottom, S->top);
      if (loss_out) *loss_out = static_cast<double>(loss);
    });
  });
}

int mmsref_backward(void* h, const int* propagate_down) {
  SessionBase* s = static_cast<SessionBase*>(h);
  return guarded(s, [&] {
    DISPATCH(s, {
      vector<bool> pd(S->bottom.size());
      for (size_t i = 0; i < pd.size(); ++i) pd[i] = propagate_down[i] != 0;
      S->layer->Backward(S->top, pd, S->bottom);
    });
  });
}

// `caffe time` protocol (tools/caffe.cpp:345-365): the caller does the warm-up;
// this runs `iters` forward(+backward) passes and returns mean wall ms per pass.
int mmsref_time(void* h, int iters, int do_backward, const int* propagate_down, double* ms_out) {
  SessionBase* s = static_cast<SessionBase*>(h);
  return guarded(s, [&] {
    DISPATCH(s, {
      vector<bool> pd(S->bottom.size());
      for (size_t i = 0; i < pd.size(); ++i) pd[i] = propagate_down ? propagate_down[i] != 0 : false;
      auto t0 = std::chrono::steady_clock::now();
      for (int it = 0; it < iters; ++it) {
        S->layer->Forward(S->bottom, S->top);
        if (do_backward) S->layer->Backward(S->top, pd, S->bottom);
      }
#ifdef MMS_DROPIN
      cudaDeviceSynchronize();
#endif
      auto t1 = std::chrono::steady_clock::now();
      *ms_out = std::chrono::duration<double, std::milli>(t1 - t0).count() / iters;
    });
  });
}

#ifdef MMS_DROPIN
// The C++ gradient-exchange glue (caffe_layers/mms_grad_exchange.cpp) with `world` solver replicas on the CURRENT device
// (virtual ranks): every replica owns `nblobs` reference Blobs with the given counts; their data / diff are filled from
// data_in / diff_in ([world][sum of counts], host), the blobs are re-bound by GradExchange, on_start and
// on_gradients_ready (or the fused AdaDelta form when fused != 0) run on one stream per replica, and the replicas'
// blob contents -- read back through the reference's own Blob::cpu_data() / cpu_diff() -- land in data_out / diff_out.
int mmsref_grad_exchange_run(int world, int nblobs, const int* counts, const float* data_in, const float* diff_in, int fused,
                             float* data_out, float* diff_out) {
  return guarded(nullptr, [&]() {
    caffe::Caffe::set_mode(caffe::Caffe::GPU);
    long long total = 0;
    for (int i = 0; i < nblobs; ++i) total += counts[i];
    std::vector<std::vector<shared_ptr<Blob<float> > > > own(world);
    std::vector<caffe::mms::GradExchange<float>*> ex(world);
    std::vector<cudaStream_t> streams(world);
    for (int r = 0; r < world; ++r) {
      std::vector<Blob<float>*> params;
      long long off = 0;
      for (int i = 0; i < nblobs; ++i) {
        shared_ptr<Blob<float> > b(new Blob<float>(std::vector<int>(1, counts[i])));
        memcpy(b->mutable_cpu_data(), data_in + r * total + off, sizeof(float) * counts[i]);
        memcpy(b->mutable_cpu_diff(), diff_in + r * total + off, sizeof(float) * counts[i]);
        b->gpu_data(); b->gpu_diff();                       // device copies exist before the re-binding
        own[r].push_back(b); params.push_back(b.get());
        off += counts[i];
      }
      ex[r] = new caffe::mms::GradExchange<float>(params, r, world);
      // the diffs were on the host: write them into the re-bound device buffers
      off = 0;
      for (int i = 0; i < nblobs; ++i) {
        CUDA_CHECK(cudaMemcpy(own[r][i]->mutable_gpu_diff(), diff_in + r * total + off, sizeof(float) * counts[i], cudaMemcpyHostToDevice));
        off += counts[i];
      }
      CUDA_CHECK(cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking));
      MMS_CAFFE_CHECK(mms_exchange_set_option(ex[r]->handle(), MMS_EXCHANGE_OPT_CTAS, 8));
    }
    caffe::mms::GradExchange<float>::Attach(ex);
    CUDA_CHECK(cudaDeviceSynchronize());
    for (int r = 0; r < world; ++r) ex[r]->on_start(streams[r], false);
    for (int r = 0; r < world; ++r) MMS_CAFFE_CHECK(mms_exchange_check(ex[r]->handle(), streams[r]));
    std::vector<float> lr(nblobs, 1.f), dec(nblobs, 1.f);
    for (int r = 0; r < world; ++r) {
      if (fused) ex[r]->on_gradients_ready_adadelta(lr, dec, 1.f, 0.95f, 5e-7f, 5e-4f, 1, streams[r], false);
      else ex[r]->on_gradients_ready(streams[r], false);
    }
    for (int r = 0; r < world; ++r) MMS_CAFFE_CHECK(mms_exchange_check(ex[r]->handle(), streams[r]));
    for (int r = 0; r < world; ++r) {
      long long off = 0;
      for (int i = 0; i < nblobs; ++i) {
        memcpy(data_out + r * total + off, own[r][i]->cpu_data(), sizeof(float) * counts[i]);
        memcpy(diff_out + r * total + off, own[r][i]->cpu_diff(), sizeof(float) * counts[i]);
        off += counts[i];
      }
    }
    for (int r = 0; r < world; ++r) { own[r].clear(); delete ex[r]; cudaStreamDestroy(streams[r]); }
  });
}
#endif

}  // extern "C"
