// Stand-in for <boost/shared_ptr.hpp> (boost is not installed in this image).
// Test infrastructure only: lets the reference's unmodified headers compile.
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; }
