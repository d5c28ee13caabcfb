// TEST INFRASTRUCTURE stub
#pragma once
#include <ros/ros.h>
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Twist { Vector3 linear, angular; };
struct TwistStamped { std_msgs::Header header; Twist twist; };
}  // namespace geometry_msgs
