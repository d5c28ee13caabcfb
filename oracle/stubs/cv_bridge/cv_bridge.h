// TEST INFRASTRUCTURE stub
#pragma once
#include <memory>
#include <string>
#include <opencv2/core.hpp>
#include <sensor_msgs/Image.h>
namespace cv_bridge {
struct CvImage {
  std_msgs::Header header;
  std::string encoding;
  cv::Mat image;
  sensor_msgs::ImagePtr toImageMsg() const { return std::make_shared<sensor_msgs::Image>(); }
};
}  // namespace cv_bridge
