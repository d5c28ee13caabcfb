// TEST INFRASTRUCTURE -- stand-in for glog: CHECK_* abort like glog's fatal checks, VLOG / LOG swallow their stream.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <ostream>

namespace glogstub {
struct Null {
  template <class T> Null& operator<<(const T&) { return *this; }
  Null& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }   // std::endl
};
struct Fatal {
  const char* what;
  explicit Fatal(const char* w) : what(w) {}
  ~Fatal() { std::fprintf(stderr, "glog stub: CHECK failed: %s\n", what); std::abort(); }
  template <class T> Fatal& operator<<(const T&) { return *this; }
  Fatal& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
}  // namespace glogstub
#define GLOGSTUB_CHECK_OP(a, b, op) if ((a) op (b)) ; else ::glogstub::Fatal(#a " " #op " " #b)
#define CHECK(c) if (c) ; else ::glogstub::Fatal(#c)
#define CHECK_EQ(a, b) GLOGSTUB_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) GLOGSTUB_CHECK_OP(a, b, !=)
#define CHECK_GE(a, b) GLOGSTUB_CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) GLOGSTUB_CHECK_OP(a, b, >)
#define CHECK_LE(a, b) GLOGSTUB_CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) GLOGSTUB_CHECK_OP(a, b, <)
#define CHECK_NOTNULL(p) (p)
#define VLOG(n) if (true) ; else ::glogstub::Null()
#define VLOG_IF(n, c) if (true) ; else ::glogstub::Null()
#define LOG(x) if (true) ; else ::glogstub::Null()
