// Per-thread workspace handle of the drop-in layers (see mms_caffe_glue.hpp).
#include "mms_caffe_glue.hpp"

namespace caffe {
namespace mms {

namespace {
struct ThreadHandle {
  mms_handle_t h = nullptr;
  int device = -1;
  ~ThreadHandle() { if (h) mms_destroy(h); }
};
}  // namespace

mms_handle_t handle() {
  static thread_local ThreadHandle th;
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  if (th.h && th.device != dev) {     // Caffe::SetDevice moved this thread to another GPU
    mms_destroy(th.h);
    th.h = nullptr;
  }
  if (!th.h) {
    MMS_CAFFE_CHECK(mms_create(&th.h));
    th.device = dev;
    // legacy default stream 0: implicit ordering with every other Caffe layer
    MMS_CAFFE_CHECK(mms_set_stream(th.h, nullptr));
    // Net::ForwardBackward / GradientChecker run a layer's Backward right after its Forward on unchanged bottoms
    // and weights (net.cpp:535-591), so SimCross backward may reuse the TF32-rounded operands the forward left
    // in the workspace.  The library checks pointers and sizes and re-rounds whenever another call touched the
    // workspace in between (e.g. a second SimCross layer on this thread).
    MMS_CAFFE_CHECK(mms_set_option(th.h, MMS_OPT_REUSE_FORWARD, 1));
  }
  return th.h;
}

}  // namespace mms
}  // namespace caffe
