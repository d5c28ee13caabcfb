/*
 * ref_fe_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" wrapper (our code) around the REAL front-end image builder of the
 * reference, compiled from where it lies (never copied):
 *   src/frontend/local_image_warped_events.cpp   AngVelEstimator::computeImageOfWarpedEvents / warpAndAccumulateEvents  (A1, A2)
 *   src/utils/image_geom_util.cpp                canonicalProjection / applyIntrinsics / cross2Matrix                    (A3)
 * ROS, OpenCV and glog are absent from this image: oracle/stubs/ supplies ros::Time, dvs_msgs::Event, empty node plumbing types, a
 * cv::Mat stand-in (GaussianBlur = the oracle's cv2-pinned blur) and CHECK macros.  AngVelEstimator's constructor / destructor live in
 * src/frontend/ang_vel_estimator.cpp (ROS publishers); they are defined here as empty bodies and the members the image builder reads
 * are set directly.  Built by oracle/Makefile into oracle/_ref/libref_fe.so (tests/test_oracle_fe_firstparty.py).
 */
#include <cmath>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <vector>

#include "backend/trajectory.h"
#include "backend/equirectangular_camera.h"
#include "utils/image_geom_util.h"
#include "utils/image_utils.h"
#include "utils/parameters.h"
#include <cv_bridge/cv_bridge.h>
#include <dvs_msgs/Event.h>
#include <dvs_msgs/EventArray.h>
#include <image_transport/image_transport.h>
#include <opencv2/imgproc.hpp>
#define private public        /* test access to the members the image builder reads */
#include "backend/event_pano_warper.h"
#include "frontend/ang_vel_estimator.h"
#undef private

namespace cmax_slam {
AngVelEstimator::AngVelEstimator(ros::NodeHandle* nh) : nh_(nh), it_(*nh) {}
AngVelEstimator::~AngVelEstimator() {}
}  // namespace cmax_slam

/* One image build exactly as local_contrast_fdf does it.  iwe: H*W floats; deriv: H*W*3 floats (CV_32FC3) or null. */
extern "C" int ref1p_fe_images(const dvs_msgs::Event* events, long long n, const unsigned t_ref[2], const double* lut_xyz, int W, int H,
                               const double K4[4], double blur_sigma, int batch_size, const double omega[3], float* iwe, float* deriv) {
  static ros::NodeHandle nh;
  cmax_slam::AngVelEstimator est(&nh);
  est.params.warp_opt.blur_sigma = blur_sigma;
  est.params.warp_opt.event_batch_size = batch_size;
  est.params.warp_opt.event_sample_rate = 1;
  est.cam_width_ = W; est.cam_height_ = H;
  est.camera_matrix_ = cv::Matx33d(K4[0], 0., K4[2], 0., K4[1], K4[3], 0., 0., 1.);
  est.precomputed_bearing_vectors_.resize((size_t)W * H);
  for (size_t i = 0; i < (size_t)W * H; ++i) est.precomputed_bearing_vectors_[i] = cv::Point3d(lut_xyz[3 * i], lut_xyz[3 * i + 1], lut_xyz[3 * i + 2]);
  est.event_subset_.assign(events, events + n);
  est.time_packet_ = ros::Time(t_ref[0], t_ref[1]);
  cv::Mat img, der;
  est.computeImageOfWarpedEvents(cv::Point3d(omega[0], omega[1], omega[2]), &img, deriv ? &der : nullptr);
  std::memcpy(iwe, img.ptr<float>(), sizeof(float) * (size_t)W * H);
  if (deriv) std::memcpy(deriv, der.ptr<float>(), sizeof(float) * (size_t)W * H * 3);
  return 0;
}
