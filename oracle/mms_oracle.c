/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's MMS hot path (the CPU oracle that can
 * travel to a box without /root/reference).  PARITY PINNED: checked in
 * tests/test_oracle.py against (1) the reference's only known-answer tests for
 * this path, src/caffe/test/test_embed_layer.cpp:54-176, (2) outputs of the
 * reference's own layer code run here (oracle/_ref, built by oracle/Makefile) and
 * (3) the committed fixtures tests/golden/ (generated from oracle/_ref by
 * tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or load this file.  The product library
 * (mms_answer_selection_b200/csrc) never does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mms_oracle.h"

#define MMS_CAT_(a, b) a##b
#define MMS_CAT(a, b) MMS_CAT_(a, b)

#define REAL float
#define FN(name) MMS_CAT(name, _f32)
#define MMSO_POW(x, y) powf(x, y)
#include "mms_oracle_impl.h"
#undef REAL
#undef FN
#undef MMSO_POW

#define REAL double
#define FN(name) MMS_CAT(name, _f64)
#define MMSO_POW(x, y) pow(x, y)
#include "mms_oracle_impl.h"
#undef REAL
#undef FN
#undef MMSO_POW
