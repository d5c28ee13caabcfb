// TEST INFRASTRUCTURE -- stand-in for roscpp's ros::Time / ros::Duration (rostime is not installed here), so that the
// reference's src/backend/trajectory.cpp compiles (oracle/ref_traj_shim.cpp).  The arithmetic restates rostime
// (time.h / duration.h / impl): normalised (sec, nsec) pairs, toSec() = sec + 1e-9 nsec, fromSec() = floor + round to ns,
// Duration * double -> Duration(toSec() * scale).  It is the same restatement the oracle and the library use; the tests
// built on this stub pin the reference's TRAJECTORY code, not ROS time arithmetic (which stays "parity unpinned").
#pragma once
#include <cmath>
#include <cstdint>

namespace ros {

struct Duration {
  int32_t sec = 0, nsec = 0;
  Duration() {}
  Duration(int32_t s, int32_t ns) : sec(s), nsec(ns) { normalize(); }
  explicit Duration(double d) { fromSec(d); }
  Duration& fromSec(double d) {
    int64_t s = (int64_t)std::floor(d);
    int64_t ns = (int64_t)std::round((d - (double)s) * 1e9);
    s += ns / 1000000000ll;
    ns %= 1000000000ll;
    sec = (int32_t)s; nsec = (int32_t)ns;
    return *this;
  }
  void normalize() {
    int64_t s = sec, ns = nsec;
    while (ns >= 1000000000ll) { ns -= 1000000000ll; ++s; }
    while (ns < 0) { ns += 1000000000ll; --s; }
    sec = (int32_t)s; nsec = (int32_t)ns;
  }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  Duration operator*(double scale) const { return Duration(toSec() * scale); }
};

struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
  explicit Time(double t) { fromSec(t); }
  Time& fromSec(double t) {
    int64_t s = (int64_t)std::floor(t);
    int64_t ns = (int64_t)std::round((t - (double)s) * 1e9);
    s += ns / 1000000000ll;
    ns %= 1000000000ll;
    sec = (uint32_t)s; nsec = (uint32_t)ns;
    return *this;
  }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  uint64_t toNSec() const { return (uint64_t)sec * 1000000000ull + (uint64_t)nsec; }
  bool operator<(const Time& o) const { return sec < o.sec || (sec == o.sec && nsec < o.nsec); }
  bool operator>(const Time& o) const { return o < *this; }
  bool operator<=(const Time& o) const { return !(o < *this); }
  bool operator>=(const Time& o) const { return !(*this < o); }
  bool operator==(const Time& o) const { return sec == o.sec && nsec == o.nsec; }
  bool operator!=(const Time& o) const { return !(*this == o); }
  Time operator+(const Duration& d) const {
    int64_t s = (int64_t)sec + d.sec, ns = (int64_t)nsec + d.nsec;
    while (ns >= 1000000000ll) { ns -= 1000000000ll; ++s; }
    while (ns < 0) { ns += 1000000000ll; --s; }
    return Time((uint32_t)s, (uint32_t)ns);
  }
  Time operator-(const Duration& d) const { return *this + Duration(-d.sec, -d.nsec); }
  Duration operator-(const Time& o) const { return Duration((int32_t)((int64_t)sec - (int64_t)o.sec), (int32_t)((int64_t)nsec - (int64_t)o.nsec)); }
  Time& operator+=(const Duration& d) { *this = *this + d; return *this; }
};

}  // namespace ros
