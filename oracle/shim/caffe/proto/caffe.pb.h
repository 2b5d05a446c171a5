// Hand-written stand-in for the protoc output of the reference's
// src/caffe/proto/caffe.proto (protoc / libprotobuf are not installed here).
// Only the messages and accessors that the hot-path sources name are provided:
// blob.cpp, syncedmem.cpp, layer.cpp, filler.hpp, layer_factory.hpp and
// layers/{embed,sim_cross,sim_matrix,pair_rank_loss,fm,loss}_layer.cpp.
// Field names, types and DEFAULTS follow caffe.proto (cited per message); the
// accessor naming follows protobuf's generated-code convention.
// Test infrastructure only -- never linked into the product library.
#pragma once
#include <string>
#include <vector>

namespace caffe {

#define MMS_PB_OPTIONAL(T, name, dflt)                          \
 private:                                                       \
  T name##_ = dflt;                                             \
  bool has_##name##_ = false;                                   \
 public:                                                        \
  T name() const { return name##_; }                            \
  bool has_##name() const { return has_##name##_; }             \
  void set_##name(T v) { name##_ = v; has_##name##_ = true; }   \
  void clear_##name() { name##_ = dflt; has_##name##_ = false; }

#define MMS_PB_STRING(name, dflt)                                         \
 private:                                                                 \
  std::string name##_ = dflt;                                             \
  bool has_##name##_ = false;                                             \
 public:                                                                  \
  const std::string& name() const { return name##_; }                     \
  bool has_##name() const { return has_##name##_; }                       \
  void set_##name(const std::string& v) { name##_ = v; has_##name##_ = true; }

#define MMS_PB_REPEATED(T, name)                                \
 private:                                                       \
  std::vector<T> name##_;                                       \
 public:                                                        \
  int name##_size() const { return static_cast<int>(name##_.size()); } \
  T name(int i) const { return name##_[i]; }                    \
  void add_##name(T v) { name##_.push_back(v); }                \
  void clear_##name() { name##_.clear(); }                      \
  const std::vector<T>& name() const { return name##_; }

#define MMS_PB_MESSAGE(T, name)                                 \
 private:                                                       \
  T name##_;                                                    \
  bool has_##name##_ = false;                                   \
 public:                                                        \
  const T& name() const { return name##_; }                     \
  T* mutable_##name() { has_##name##_ = true; return &name##_; } \
  bool has_##name() const { return has_##name##_; }             \
  void clear_##name() { name##_ = T(); has_##name##_ = false; }

// caffe.proto:287-292
enum Phase { TRAIN = 0, TEST = 1 };

// caffe.proto:6-8
class BlobShape {
  MMS_PB_REPEATED(long long, dim)
};

// caffe.proto:11-24
class BlobProto {
  MMS_PB_MESSAGE(BlobShape, shape)
  MMS_PB_REPEATED(float, data)
  MMS_PB_REPEATED(float, diff)
  MMS_PB_REPEATED(double, double_data)
  MMS_PB_REPEATED(double, double_diff)
  MMS_PB_OPTIONAL(int, num, 0)
  MMS_PB_OPTIONAL(int, channels, 0)
  MMS_PB_OPTIONAL(int, height, 0)
  MMS_PB_OPTIONAL(int, width, 0)
};

// caffe.proto:41-62  (default filler: type "constant", value 0)
enum FillerParameter_VarianceNorm {
  FillerParameter_VarianceNorm_FAN_IN = 0,
  FillerParameter_VarianceNorm_FAN_OUT = 1,
  FillerParameter_VarianceNorm_AVERAGE = 2
};
class FillerParameter {
 public:
  typedef FillerParameter_VarianceNorm VarianceNorm;
  static const VarianceNorm FAN_IN = FillerParameter_VarianceNorm_FAN_IN;
  static const VarianceNorm FAN_OUT = FillerParameter_VarianceNorm_FAN_OUT;
  static const VarianceNorm AVERAGE = FillerParameter_VarianceNorm_AVERAGE;
  MMS_PB_STRING(type, "constant")
  MMS_PB_OPTIONAL(float, value, 0.f)
  MMS_PB_OPTIONAL(float, min, 0.f)
  MMS_PB_OPTIONAL(float, max, 1.f)
  MMS_PB_OPTIONAL(float, mean, 0.f)
  MMS_PB_OPTIONAL(float, std, 1.f)
  MMS_PB_OPTIONAL(int, sparse, -1)
  MMS_PB_OPTIONAL(FillerParameter_VarianceNorm, variance_norm,
                  FillerParameter_VarianceNorm_FAN_IN)
};

// caffe.proto:471-477 (note the reference's spelling: mesure_count)
class SimCrossParameter {
  MMS_PB_OPTIONAL(int, dist_mode, 1)
  MMS_PB_OPTIONAL(int, mesure_count, 1)
  MMS_PB_MESSAGE(FillerParameter, weight_filler)
  MMS_PB_OPTIONAL(bool, bias_term, true)
  MMS_PB_MESSAGE(FillerParameter, bias_filler)
};

// caffe.proto:430-432
class SimMatrixParameter {
  MMS_PB_MESSAGE(FillerParameter, weight_filler)
};

// caffe.proto:479-481
class PairRankLossParameter {
  MMS_PB_OPTIONAL(float, margin, 1.0f)
};

// caffe.proto:790-803
class EmbedParameter {
  MMS_PB_OPTIONAL(unsigned, num_output, 0u)
  MMS_PB_OPTIONAL(unsigned, input_dim, 0u)
  MMS_PB_OPTIONAL(bool, bias_term, true)
  MMS_PB_MESSAGE(FillerParameter, weight_filler)
  MMS_PB_MESSAGE(FillerParameter, bias_filler)
  MMS_PB_STRING(weight_source, "")
};

// caffe.proto:418-420
class FMParameter {
  MMS_PB_OPTIONAL(bool, bias_term, true)
};

// caffe.proto:422-428
class MAPParameter {
  MMS_PB_OPTIONAL(int, fixed_axis, 1)
};
class MRRParameter {
  MMS_PB_OPTIONAL(int, fixed_axis, 1)
};

// caffe.proto:465-469
class AUCParameter {
  MMS_PB_OPTIONAL(int, fixed_axis, 1)
  MMS_PB_OPTIONAL(int, axis, 1)
  MMS_PB_OPTIONAL(int, ignore_label, 0)
};

// caffe.proto (ConvolutionParameter): the fields base_conv_layer.cpp reads
enum ConvolutionParameter_Engine { ConvolutionParameter_Engine_DEFAULT = 0, ConvolutionParameter_Engine_CAFFE = 1,
                                   ConvolutionParameter_Engine_CUDNN = 2 };
class ConvolutionParameter {
  MMS_PB_OPTIONAL(unsigned, num_output, 0u)
  MMS_PB_OPTIONAL(bool, bias_term, true)
  MMS_PB_REPEATED(unsigned, pad)
  MMS_PB_REPEATED(unsigned, kernel_size)
  MMS_PB_REPEATED(unsigned, stride)
  MMS_PB_REPEATED(unsigned, dilation)
  MMS_PB_OPTIONAL(unsigned, pad_h, 0u)
  MMS_PB_OPTIONAL(unsigned, pad_w, 0u)
  MMS_PB_OPTIONAL(unsigned, kernel_h, 0u)
  MMS_PB_OPTIONAL(unsigned, kernel_w, 0u)
  MMS_PB_OPTIONAL(unsigned, stride_h, 0u)
  MMS_PB_OPTIONAL(unsigned, stride_w, 0u)
  MMS_PB_OPTIONAL(unsigned, group, 1u)
  MMS_PB_MESSAGE(FillerParameter, weight_filler)
  MMS_PB_MESSAGE(FillerParameter, bias_filler)
  MMS_PB_OPTIONAL(ConvolutionParameter_Engine, engine, ConvolutionParameter_Engine_DEFAULT)
  MMS_PB_OPTIONAL(int, axis, 1)
  MMS_PB_OPTIONAL(bool, force_nd_im2col, false)
};

// caffe.proto (PoolingParameter)
enum PoolingParameter_PoolMethod { PoolingParameter_PoolMethod_MAX = 0, PoolingParameter_PoolMethod_AVE = 1,
                                   PoolingParameter_PoolMethod_STOCHASTIC = 2 };
class PoolingParameter {
 public:
  typedef PoolingParameter_PoolMethod PoolMethod;
  MMS_PB_OPTIONAL(PoolingParameter_PoolMethod, pool, PoolingParameter_PoolMethod_MAX)
  MMS_PB_OPTIONAL(unsigned, pad, 0u)
  MMS_PB_OPTIONAL(unsigned, pad_h, 0u)
  MMS_PB_OPTIONAL(unsigned, pad_w, 0u)
  MMS_PB_OPTIONAL(unsigned, kernel_size, 0u)
  MMS_PB_OPTIONAL(unsigned, kernel_h, 0u)
  MMS_PB_OPTIONAL(unsigned, kernel_w, 0u)
  MMS_PB_OPTIONAL(unsigned, stride, 1u)
  MMS_PB_OPTIONAL(unsigned, stride_h, 0u)
  MMS_PB_OPTIONAL(unsigned, stride_w, 0u)
  MMS_PB_OPTIONAL(bool, global_pooling, false)
};

// caffe.proto:484-488
class BNParameter {
  MMS_PB_OPTIONAL(float, bn_memory, 0.9f)
  MMS_PB_MESSAGE(FillerParameter, scale_filler)
  MMS_PB_MESSAGE(FillerParameter, shift_filler)
};

// caffe.proto:1216-1223
class TanHParameter {
  MMS_PB_OPTIONAL(int, engine, 0)
};

// caffe.proto (LossParameter): ignore_label, normalize
class LossParameter {
  MMS_PB_OPTIONAL(int, ignore_label, 0)
  MMS_PB_OPTIONAL(bool, normalize, true)
};

// caffe.proto:310-416, restricted to what the hot-path layers read.
class LayerParameter {
  MMS_PB_STRING(name, "")
  MMS_PB_STRING(type, "")
  MMS_PB_OPTIONAL(Phase, phase, TRAIN)
  MMS_PB_REPEATED(float, loss_weight)
  MMS_PB_REPEATED(bool, propagate_down)
  MMS_PB_MESSAGE(SimCrossParameter, sim_cross_param)
  MMS_PB_MESSAGE(SimMatrixParameter, sim_matrix_param)
  MMS_PB_MESSAGE(PairRankLossParameter, pair_rank_loss_param)
  MMS_PB_MESSAGE(EmbedParameter, embed_param)
  MMS_PB_MESSAGE(FMParameter, fm_param)
  MMS_PB_MESSAGE(LossParameter, loss_param)
  MMS_PB_MESSAGE(MAPParameter, map_param)
  MMS_PB_MESSAGE(MRRParameter, mrr_param)
  MMS_PB_MESSAGE(AUCParameter, auc_param)
  MMS_PB_MESSAGE(ConvolutionParameter, convolution_param)
  MMS_PB_MESSAGE(PoolingParameter, pooling_param)
  MMS_PB_MESSAGE(BNParameter, bn_param)
  MMS_PB_MESSAGE(TanHParameter, tanh_param)
 private:
  std::vector<BlobProto> blobs_;
 public:
  int blobs_size() const { return static_cast<int>(blobs_.size()); }
  const BlobProto& blobs(int i) const { return blobs_[i]; }
  BlobProto* add_blobs() { blobs_.emplace_back(); return &blobs_.back(); }
  void clear_blobs() { blobs_.clear(); }
  void Clear() { *this = LayerParameter(); }
  void CopyFrom(const LayerParameter& other) { *this = other; }
};

}  // namespace caffe
