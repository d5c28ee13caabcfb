// TEST INFRASTRUCTURE stub
#pragma once
#include <memory>
#include <ros/ros.h>
namespace sensor_msgs { struct Image { std_msgs::Header header; }; typedef std::shared_ptr<Image> ImagePtr; }
