// placeholder until the tcgen05 kernels land: every shape reports "unsupported", so the
// dispatcher in simcross.cu uses the SIMT composition.
#include "../mms_common.cuh"
int mms_tc_simcross2_forward(mms_context*, const float*, const float*, const float*, const float*,
                             float*, int, int, int, int, int) { return MMS_E_UNSUPPORTED; }
int mms_tc_simcross2_backward(mms_context*, const float*, const float*, const float*, const float*,
                              float*, float*, float*, float*, int, int, int, int, int) { return MMS_E_UNSUPPORTED; }
void mms_tc_destroy_state(mms_context*) {}
