// TEST INFRASTRUCTURE stub
#pragma once
#include <ros/ros.h>
namespace image_transport {
struct Publisher {};
struct ImageTransport { explicit ImageTransport(const ros::NodeHandle&) {} };
}
