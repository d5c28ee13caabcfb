// TEST INFRASTRUCTURE stub: the two accessors AngVelEstimator::initialize reads
#pragma once
#include <opencv2/core.hpp>
namespace image_geometry {
class PinholeCameraModel {
 public:
  cv::Size res;
  cv::Matx33d K;
  cv::Size fullResolution() const { return res; }
  cv::Matx33d fullIntrinsicMatrix() const { return K; }
};
}  // namespace image_geometry
