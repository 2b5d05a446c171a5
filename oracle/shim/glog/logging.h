// Stand-in for <glog/logging.h>: CHECK*/LOG*/DCHECK* with glog's abort-on-FATAL
// behaviour.  Also pulls in the libc headers real glog drags in transitively
// (the reference relies on them for memset/strlen/fscanf/FLT_MAX).
// Test infrastructure only.
#pragma once
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
namespace mms_shim_log {
// glog aborts the process on LOG(FATAL)/CHECK failure.  The oracle harness wants
// to observe that outcome from a test, so the stand-in throws instead; the
// harness (oracle/ref_harness.cpp) turns it into an error code + message.
struct FatalError : public std::runtime_error {
  explicit FatalError(const std::string& m) : std::runtime_error(m) {}
};
class Message {
 public:
  Message(bool fatal, const char* file, int line) : fatal_(fatal) {
    stream_ << file << ":" << line << "] ";
  }
  ~Message() noexcept(false) {
    if (fatal_) {
      throw FatalError(stream_.str());
    } else if (std::getenv("MMS_REF_VERBOSE")) {
      std::cerr << "I " << stream_.str() << std::endl;
    }
  }
  std::ostream& stream() { return stream_; }
 private:
  bool fatal_;
  std::ostringstream stream_;
};
struct Voidify { void operator&(std::ostream&) {} };
}  // namespace mms_shim_log
#define MMS_SHIM_LOG(fatal) mms_shim_log::Message(fatal, __FILE__, __LINE__).stream()
#define MMS_SHIM_LOG_INFO MMS_SHIM_LOG(false)
#define MMS_SHIM_LOG_WARNING MMS_SHIM_LOG(false)
#define MMS_SHIM_LOG_ERROR MMS_SHIM_LOG(false)
#define MMS_SHIM_LOG_FATAL MMS_SHIM_LOG(true)
#define LOG(severity) MMS_SHIM_LOG_##severity
#define DLOG(severity) LOG(severity)
#define LOG_IF(severity, cond) !(cond) ? (void)0 : mms_shim_log::Voidify() & LOG(severity)
#define LOG_EVERY_N(severity, n) LOG(severity)
#define LOG_FIRST_N(severity, n) LOG(severity)
#define CHECK(cond) \
  (cond) ? (void)0 : mms_shim_log::Voidify() & MMS_SHIM_LOG(true) << "Check failed: " #cond " "
#define MMS_SHIM_CHECK_OP(a, b, op) \
  ((a)op(b)) ? (void)0 : mms_shim_log::Voidify() & MMS_SHIM_LOG(true) \
      << "Check failed: " #a " " #op " " #b " "
#define CHECK_EQ(a, b) MMS_SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) MMS_SHIM_CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) MMS_SHIM_CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) MMS_SHIM_CHECK_OP(a, b, <)
#define CHECK_GE(a, b) MMS_SHIM_CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) MMS_SHIM_CHECK_OP(a, b, >)
#define CHECK_NOTNULL(p) (p)
// Release-build glog compiles DCHECKs away; the reference is built -DNDEBUG
// (Makefile: DEBUG := 0), so do the same.
#define DCHECK(cond) while (false) CHECK(cond)
#define DCHECK_EQ(a, b) while (false) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) while (false) CHECK_NE(a, b)
#define DCHECK_LE(a, b) while (false) CHECK_LE(a, b)
#define DCHECK_LT(a, b) while (false) CHECK_LT(a, b)
#define DCHECK_GE(a, b) while (false) CHECK_GE(a, b)
#define DCHECK_GT(a, b) while (false) CHECK_GT(a, b)
