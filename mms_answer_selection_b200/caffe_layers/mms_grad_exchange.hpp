// Data-parallel gradient exchange for Caffe solvers: the C++ host side of mms_exchange_* (include/mms_b200.h).
//
// Replaces, in the reference's src/caffe/parallel.cpp:
//   Params / GPUParams (:60-115)              one flat data and one flat diff buffer per solver, the net's learnable
//                                             blobs re-pointed into them (apply_buffers + replace_gpu / replace_gpu_diff,
//                                             :36-55, :110-115)                                   -> GradExchange ctor
//   P2PSync::on_start (:287-322)              weights down the tree of peer copies                -> on_start()
//   P2PSync::on_gradients_ready (:325-380)    gradients up the tree, caffe_gpu_add per level, 1/solver_count on the
//                                             root (:377)                                          -> on_gradients_ready()
// One GradExchange per solver thread / GPU (P2PSync::InternalThreadEntry, :271-284); after every thread has built its
// object, any one thread calls GradExchange::Attach(all) once (the reference wires parent / children pointers and
// enables peer access in the P2PSync constructor, :208-262).  With processes instead of threads use
// mms_exchange_export_ipc / _attach_ipc.  Compiled against the reference's own headers; host code only.
#ifndef MMS_GRAD_EXCHANGE_HPP_
#define MMS_GRAD_EXCHANGE_HPP_

#include <vector>

#include "caffe/blob.hpp"
#include "mms_caffe_glue.hpp"

namespace caffe {
namespace mms {

template <typename Dtype>
class GradExchange {
 public:
  // `params`: Net::learnable_params() of this thread's solver (shared blobs once, net order)
  GradExchange(const std::vector<Blob<Dtype>*>& params, int rank, int world);
  ~GradExchange();
  // peers' allocations as plain pointers (same process); enables peer access
  static void Attach(const std::vector<GradExchange<Dtype>*>& all);

  // Solver callbacks (solver.hpp:88-95).  `stream`: the stream the solver's layers run on (Caffe: the legacy default
  // stream, 0).  on_start copies the root's weights into every replica; on_gradients_ready leaves the MEAN gradient
  // over the solvers in every replica's diff (sum and the 1/solver_count of :377), so that every solver can apply the
  // identical update -- or, with on_gradients_ready_adadelta, has it applied by the owners of the slices and receives
  // the new weights.  `sync`: block until the exchange has finished, as the reference's callbacks do.
  void on_start(cudaStream_t stream = 0, bool sync = true);
  void on_gradients_ready(cudaStream_t stream = 0, bool sync = true);
  void on_gradients_ready_adadelta(const std::vector<float>& lr_mult, const std::vector<float>& decay_mult, Dtype base_lr,
                                   Dtype momentum, Dtype delta, Dtype weight_decay, int iter_size, cudaStream_t stream = 0,
                                   bool sync = true);

  long long count() const { return count_; }
  long long offset(int i) const { return offsets_[i]; }
  Dtype* data() const { return data_; }
  Dtype* diff() const { return diff_; }
  mms_exchange_t handle() const { return x_; }

 private:
  mms_exchange_t x_;
  int rank_, world_;
  long long count_;
  std::vector<long long> offsets_, counts_;
  Dtype* data_;
  Dtype* diff_;
};

}  // namespace mms
}  // namespace caffe

#endif  // MMS_GRAD_EXCHANGE_HPP_
