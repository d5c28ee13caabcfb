/*
 * ref_warper_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" wrapper (our code) around the REAL back-end warp code of the
 * reference, compiled from where it lies (never copied):
 *   src/backend/event_pano_warper.cpp   EventWarper::computeImageOfWarpedEvents / warpAndAccumulateEvents / updateAlpha /
 *                                       updateIGp / updateIG / setUpdateTimesIG / warpEventToMap        
 *   src/backend/trajectory.cpp, include/backend/equirectangular_camera.h, real basalt / Sophus / Eigen
 * ROS, OpenCV and glog are absent from this image: oracle/stubs/ supplies ros::Time, dvs_msgs::Event, a cv::Mat stand-in whose
 * element-wise helpers carry the oracle's restatement of the OpenCV arithmetic (GaussianBlur = the cv2-pinned blur), and CHECK
 * macros.  What this pins is the reference's own batching, indexing, bounds, votes, old/new split, band indices and
 * expression structure.  Built by oracle/Makefile into oracle/_ref/libref_warper.so (tests/test_oracle_be_firstparty.py).
 */
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>

#include "backend/trajectory.h"
#include "backend/equirectangular_camera.h"
#include "utils/image_geom_util.h"
#include "utils/image_utils.h"
#include "utils/parameters.h"
#include <dvs_msgs/Event.h>
#include <dvs_msgs/EventArray.h>
#define private public        /* test access to EventWarper's images (IG_, IL_old_, IL_new_, alpha_) */
#include "backend/event_pano_warper.h"
#undef private

using namespace cmax_slam;

struct RefWarper {
  std::unique_ptr<EventWarper> w;
  std::vector<cv::Point3d> lut;
  int PW = 0, PH = 0, order = 2;
};

extern "C" RefWarper* ref1p_warper_create(const double* lut_xyz, int SW, int SH, int PW, int PH, double blur_sigma, int batch_size,
                                          int sample_rate, int max_update_times, int order) {
  OptionsWarp wo; wo.blur_sigma = blur_sigma; wo.event_batch_size = batch_size; wo.event_sample_rate = sample_rate;
  OptionPanoMap mo; mo.pano_height = PH; mo.pano_width = PW; mo.Y_angle = 0.0; mo.max_update_times = max_update_times; mo.backend_min_ev_rate = 0;
  RefWarper* r = new RefWarper();
  r->PW = PW; r->PH = PH; r->order = order;
  r->lut.resize((size_t)SW * SH);
  for (size_t i = 0; i < r->lut.size(); ++i) r->lut[i] = cv::Point3d(lut_xyz[3 * i], lut_xyz[3 * i + 1], lut_xyz[3 * i + 2]);
  r->w.reset(new EventWarper(wo, mo));
  r->w->initialize(SW, SH, &r->lut);
  return r;
}
extern "C" void ref1p_warper_destroy(RefWarper* r) { delete r; }

/* IG_ <- given image (what updateIG accumulated over earlier windows) */
extern "C" void ref1p_warper_set_ig(RefWarper* r, const float* IG) {
  std::memcpy(r->w->IG_.ptr<float>(), IG, sizeof(float) * (size_t)r->PW * r->PH);
}
extern "C" void ref1p_warper_get_map(RefWarper* r, float* IG, unsigned char* times) {
  if (IG) std::memcpy(IG, r->w->IG_.ptr<float>(), sizeof(float) * (size_t)r->PW * r->PH);
  if (times) std::memcpy(times, r->w->IG_update_times_map_.ptr<unsigned char>(), (size_t)r->PW * r->PH);
}

/* One cost-function image build exactly as global_contrast_fdf does it: temporary trajectory = (t_beg, dt_knots, knots)
 * [the constructor CopyAndIncrementalUpdate uses], computeImageOfWarpedEvents(traj, events, &iwe, want_grad ? &bands : 0).
 * first_iter != 0: setFirstIter(true) (IGp <- IG, alpha from updateAlpha).  Outputs may be null. */
extern "C" int ref1p_warper_eval(RefWarper* r, const dvs_msgs::Event* events, long long n, double t_beg, double dt_knots,
                                 const double* knots_xyzw, int K, int n_fixed, const unsigned* tnext, int first_iter, int want_grad,
                                 float* iwe_out, float* bands_out, float* il_old, float* il_new, double* alpha) {
  std::vector<Sophus::SO3d> cps;
  for (int i = 0; i < K; ++i) cps.push_back(Sophus::SO3d(Eigen::Quaterniond(knots_xyzw[4 * i + 3], knots_xyzw[4 * i], knots_xyzw[4 * i + 1], knots_xyzw[4 * i + 2])));
  std::unique_ptr<Trajectory> traj;
  if (r->order == 4) traj.reset(new CubicTrajectory(t_beg, dt_knots, cps)); else traj.reset(new LinearTrajectory(t_beg, dt_knots, cps));
  r->w->setNumFixedCtrlPoses(n_fixed);
  r->w->setNextWinBegTime(ros::Time(tnext[0], tnext[1]));
  if (first_iter) r->w->setFirstIter(true);
  std::vector<dvs_msgs::Event> ev(events, events + n);
  cv::Mat iwe;
  std::vector<cv::Mat> bands;
  r->w->computeImageOfWarpedEvents(traj.get(), &ev, &iwe, want_grad ? &bands : nullptr);
  const size_t A = (size_t)r->PW * r->PH;
  if (iwe_out) std::memcpy(iwe_out, iwe.ptr<float>(), sizeof(float) * A);
  if (bands_out && want_grad) for (size_t p = 0; p < bands.size(); ++p) std::memcpy(bands_out + p * A, bands[p].ptr<float>(), sizeof(float) * A);
  if (il_old) std::memcpy(il_old, r->w->IL_old_.ptr<float>(), sizeof(float) * A);
  if (il_new) std::memcpy(il_new, r->w->IL_new_.ptr<float>(), sizeof(float) * A);
  if (alpha) *alpha = r->w->alpha_;
  return (int)bands.size();
}
extern "C" void ref1p_warper_update_ig(RefWarper* r) { r->w->updateIG(); }
extern "C" void ref1p_warper_mark_fov(RefWarper* r, const double q[4], int radius) {
  r->w->setUpdateTimesIG(Sophus::SO3d(Eigen::Quaterniond(q[3], q[0], q[1], q[2])), radius);
}
