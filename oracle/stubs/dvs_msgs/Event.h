// TEST INFRASTRUCTURE stub: dvs_msgs::Event as the ROS C++ message generator lays it out (uint16 x, uint16 y, ros::Time ts,
// bool polarity): 16 bytes, identical to cmaxb_event / orc_event, so test arrays can be reinterpreted.
#pragma once
#include <cstdint>

#include "../ros/time.h"
namespace dvs_msgs {
struct Event {
  uint16_t x = 0, y = 0;
  ros::Time ts;
  uint8_t polarity = 0;
};
static_assert(sizeof(Event) == 16, "dvs_msgs::Event layout");
}  // namespace dvs_msgs
