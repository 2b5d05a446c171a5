// One element of the AdaDelta solver step, shared by the single-GPU optimizer kernel (elementwise.cu) and the
// fused gradient-exchange tail (exchange.cu).  Reference: src/caffe/solvers/adadelta_solver.cu:7-16 (the update),
// sgd_solver.cpp:118-204 (Normalize / Regularize), blob.cpp:160-183 (Blob::Update).
#pragma once

// The update itself is written the way the reference's kernel is (adadelta_solver.cu:9-13), INCLUDING its quirk: gi
// and hi are `float` locals there whatever Dtype is, so for double blobs the gradient and the refreshed gradient
// history are narrowed to float before they enter the square root, and the update that reaches g / h2 carries float
// precision.  Results are pinned by execution against that kernel compiled verbatim (oracle/_ref/libmms_refcuda.so).
template <typename T>
__device__ __forceinline__ void adadelta_one(T& w, T& g, T& h, T& h2, bool has_w, T grad_scale, T local_decay,
                                             T momentum, T delta, T local_rate, bool clear) {
  T gd = g;
  if (grad_scale != T(1)) gd = gd * grad_scale;      // caffe_gpu_scal (Normalize / P2PSync's 1/n)
  if (has_w && local_decay != T(0)) gd = local_decay * w + gd;   // caffe_gpu_axpy (Regularize, L2)
  float gi = gd;
  float hi = h = momentum * h + (1 - momentum) * gi * gi;
  gi = gi * sqrt((h2 + delta) / (hi + delta));
  h2 = momentum * h2 + (1 - momentum) * gi * gi;
  const T upd = local_rate * gi;
  if (has_w) w = T(-1) * upd + w;                    // Blob::Update: caffe_gpu_axpy(count, -1, diff, data)
  g = clear ? T(0) : upd;
}

