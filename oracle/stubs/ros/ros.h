// TEST INFRASTRUCTURE stub: ros::Time / ros::Duration (time.h) plus empty node plumbing types so that the reference's class
// headers compile; no ROS communication exists in the code that is compiled.
#pragma once
#include "time.h"
namespace ros {
struct NodeHandle {};
struct Publisher {};
}  // namespace ros
