// TEST INFRASTRUCTURE stub (inert publishers)
#pragma once
#include <ros/ros.h>
#include <string>
namespace image_transport {
struct Publisher {
  void shutdown() {}
  int getNumSubscribers() const { return 0; }
  template <class M> void publish(const M&) const {}
};
struct ImageTransport {
  explicit ImageTransport(const ros::NodeHandle&) {}
  Publisher advertise(const std::string&, int) { return Publisher(); }
};
}  // namespace image_transport
