// TEST INFRASTRUCTURE stub: ros::Time / ros::Duration (time.h) plus inert node plumbing so that the reference's node classes
// compile and run without ROS: publishers publish nothing and have no subscribers.
#pragma once
#include <string>

#include "time.h"
namespace ros {
struct Publisher {
  void shutdown() {}
  int getNumSubscribers() const { return 0; }
  template <class M> void publish(const M&) const {}
};
struct NodeHandle {
  template <class M> Publisher advertise(const std::string&, int) { return Publisher(); }
};
}  // namespace ros
namespace std_msgs { struct Header { ros::Time stamp; std::string frame_id; }; }
