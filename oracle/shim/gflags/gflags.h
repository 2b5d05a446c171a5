// Stand-in for <gflags/gflags.h>.  Defining GFLAGS_GFLAGS_H_ keeps the reference's
// common.hpp from aliasing `namespace gflags = google`.  Test infrastructure only.
#pragma once
#define GFLAGS_GFLAGS_H_
namespace gflags {}
