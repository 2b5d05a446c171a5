// Stand-in for <boost/thread.hpp>: only boost::mutex is named by the reference's
// layer.hpp / layer.cpp (forward_mutex_).  Test infrastructure only.
#pragma once
#include <mutex>
namespace boost {
class mutex {
 public:
  void lock() { m_.lock(); }
  void unlock() { m_.unlock(); }
 private:
  std::mutex m_;
};
}  // namespace boost
