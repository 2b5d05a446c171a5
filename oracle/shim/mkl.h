// Stand-in for Intel MKL's umbrella header.  The reference author built with
// BLAS := mkl (Makefile.config:33), which makes caffe_cpu_axpby resolve to the
// library's element-wise cblas_?axpby instead of Caffe's scal-then-axpy fallback
// (include/caffe/util/mkl_alternate.hpp:83-94).  -DUSE_MKL + this header keeps
// that behaviour with OpenBLAS.  Test infrastructure only.
#pragma once
#include "cblas.h"
