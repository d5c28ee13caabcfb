/*
 * ref_node_shim.cpp -- TEST INFRASTRUCTURE.  The reference's OWN node classes driven without ROS, compiled from where they lie
 * (never copied):
 *   src/frontend/ang_vel_estimator.cpp     AngVelEstimator: pushEvent, getEventSubset, deleteOldEvents, slideWindow        (8f-3)
 *   src/backend/pose_graph_optimizer.cpp   PoseGraphOptimizer: pushAngVel, isReadyFrontendPoses, getEventSubset,
 *                                          getAngVelSubset, integrateAngVel, processTimeWindow, setUpdateTimesIG, slideWindow (8f-4)
 *   + trajectory.cpp, event_pano_warper.cpp, local_image_warped_events.cpp, image_geom_util.cpp, real basalt / Sophus / Eigen
 * with the stand-in headers of oracle/stubs/ (inert publishers, ros::Time, cv::Mat stand-in, glog CHECKs).  The two GSL solves
 * (local_optim_contrast_gsl.cpp, global_optim_contrast_gsl.cpp need GSL) are replaced here: the front-end "solve" returns the next
 * angular velocity of a caller-supplied table, the back-end solve leaves the control poses unchanged -- what this pins is the
 * reference's event bookkeeping, packet / window cutting, trajectory initialisation and control-pose index arithmetic.
 * The reference polls PoseGraphOptimizer::Run() from a second thread; here its loop body runs after every pushEvent.
 * Built by oracle/Makefile into oracle/_ref/libref_node.so (tests/test_node_firstparty.py).
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
#include <image_geometry/pinhole_camera_model.h>
#include <image_transport/image_transport.h>
#include <opencv2/imgproc.hpp>
#define private public        /* the Run() loop body and the state that is compared are private */
#include "backend/event_pano_warper.h"
#include "backend/pose_graph_optimizer.h"
#include "frontend/ang_vel_estimator.h"
#undef private

using namespace cmax_slam;

struct PacketRec { uint32_t sec, nsec; long long n; uint32_t f_sec, f_nsec, l_sec, l_nsec; unsigned long long hash; double w[3]; };
struct WindowRec {
  uint32_t beg[2], end[2]; long long n_events; uint32_t f[2], l[2]; unsigned long long hash;
  int n_ctrl, idx_traj, idx_opt, num_opt; uint32_t latest[2]; double latest_q[4];
  std::vector<double> knots;
};
static unsigned long long hash_events(const std::vector<dvs_msgs::Event>& ev) {
  unsigned long long h = 1469598103934665603ull;
  for (const auto& e : ev) {
    const unsigned long long v[4] = {e.x, e.y, e.ts.sec, e.ts.nsec};
    for (unsigned long long x : v) { h ^= x; h *= 1099511628211ull; }
  }
  return h;
}

struct RefNode {
  ros::NodeHandle nh;
  image_geometry::PinholeCameraModel cam;
  std::vector<cv::Point3d> lut;
  std::unique_ptr<AngVelEstimator> fe;
  std::unique_ptr<PoseGraphOptimizer> be;
  std::vector<double> omegas; size_t next_omega = 0;
  int fe_rate = 1;
  std::vector<PacketRec> packets;
  std::vector<WindowRec> windows;
};
static RefNode* g_node = nullptr;     /* the stand-in solves find their node here */

#ifndef REF_FULL
namespace cmax_slam {
/* stands in for src/frontend/local_optim_contrast_gsl.cpp:74-233 (needs GSL): records the packet, returns the next table entry */
double AngVelEstimator::setupProblemAndOptimize_gsl(cv::Point3d& ang_vel) {
  RefNode* n = g_node;
  PacketRec r{};
  r.sec = time_packet_.sec; r.nsec = time_packet_.nsec; r.n = (long long)event_subset_.size();
  r.f_sec = event_subset_.front().ts.sec; r.f_nsec = event_subset_.front().ts.nsec;
  r.l_sec = event_subset_.back().ts.sec; r.l_nsec = event_subset_.back().ts.nsec;
  r.hash = hash_events(event_subset_);
  n->packets.push_back(r);
  const size_t k = n->next_omega++ % (n->omegas.size() / 3);
  ang_vel = cv::Point3d(n->omegas[3 * k], n->omegas[3 * k + 1], n->omegas[3 * k + 2]);
  return 0.0;
}
/* stands in for src/backend/global_optim_contrast_gsl.cpp:15-145 (needs GSL): no solve, control poses unchanged */
void PoseGraphOptimizer::setupProblemAndOptimize_gsl() {}
}  // namespace cmax_slam
#endif  /* !REF_FULL: with REF_FULL the reference's own *_optim_contrast_gsl*.cpp are linked (over the GSL stand-in of oracle/stubs/gsl) */

extern "C" RefNode* ref1p_node_create(int W, int H, const double K4[4], double dt_ang_vel, int num_events_per_packet, int fe_sample_rate,
                                      double win_size, double win_stride, double dt_knots, int spline_degree, int pano_height,
                                      double y_angle, int min_ev_rate, int max_update_times, const double* omegas, int n_omegas) {
  RefNode* n = new RefNode();
  g_node = n;
  n->cam.res = cv::Size(W, H);
  n->cam.K = cv::Matx33d(K4[0], 0., K4[2], 0., K4[1], K4[3], 0., 0., 1.);
  n->lut.resize((size_t)W * H);
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) n->lut[(size_t)y * W + x] = cv::Point3d((x - K4[2]) / K4[0], (y - K4[3]) / K4[1], 1.0);
  n->omegas.assign(omegas, omegas + 3 * n_omegas);
  n->fe_rate = fe_sample_rate;
  /* wiring of CMaxSLAM::CMaxSLAM (src/cmax_slam.cpp:85-96) and cameraInfoCallback (:122-145) */
  n->fe.reset(new AngVelEstimator(&n->nh));
  n->be.reset(new PoseGraphOptimizer(&n->nh));
  n->fe->setBackend(n->be.get());
  n->be->setFrontend(n->fe.get());
  AngVelEstParams fp;
  fp.process_opt.contrast_measure = 0;
  fp.num_events_per_packet = num_events_per_packet; fp.dt_ang_vel = dt_ang_vel;
  fp.warp_opt.blur_sigma = 1.0; fp.warp_opt.event_batch_size = 100; fp.warp_opt.event_sample_rate = fe_sample_rate;
  fp.data_opt.show_iwe = false;
  PoseGraphParams bp;
  bp.process_opt.contrast_measure = 0;
  bp.sliding_window_opt.time_window_size = win_size; bp.sliding_window_opt.sliding_window_stride = win_stride;
  bp.warp_opt.blur_sigma = 1.0; bp.warp_opt.event_batch_size = 100; bp.warp_opt.event_sample_rate = 1;
  bp.traj_opt.dt_knots = dt_knots; bp.traj_opt.spline_degree = spline_degree;
  bp.data_opt.show_iwe = false;
  bp.map_opt.pano_height = pano_height; bp.map_opt.pano_width = 2 * pano_height; bp.map_opt.Y_angle = y_angle;
  bp.map_opt.backend_min_ev_rate = min_ev_rate; bp.map_opt.max_update_times = max_update_times;
  bp.draw_FOV = false; bp.gamma = 0.75;
  n->fe->initialize(&n->cam, fp, n->lut);
  n->be->initialize(W, H, bp, &(n->fe->events_), &n->lut);
  return n;
}
extern "C" void ref1p_node_destroy(RefNode* n) { if (g_node == n) g_node = nullptr; delete n; }

/* CMaxSLAM::eventsCallback (src/cmax_slam.cpp:147-161) for one message, with the body of PoseGraphOptimizer::Run() (:357-377)
 * executed after every pushEvent */
extern "C" void ref1p_node_events(RefNode* n, const dvs_msgs::Event* ev, long long count) {
  g_node = n;
  for (long long i = 0; i < count; i += n->fe_rate) {
    const bool was_init = n->fe->sliding_window_initialized_;
    const ros::Time tp_before = n->fe->time_packet_;
    const size_t solved_before = n->packets.size();
    n->fe->pushEvent(ev[i]);
    if (was_init && n->fe->time_packet_ != tp_before && n->packets.size() == solved_before) {
      /* a packet was consumed that the stand-in solve did not see: REF_FULL (the real solve ran: n = -2), or its time span
         exceeded 10 dt_ang_vel and the reference set the angular velocity to zero (ang_vel_estimator.cpp:109-114: n = -1) */
      PacketRec r{};
      r.sec = tp_before.sec; r.nsec = tp_before.nsec;
#ifdef REF_FULL
      r.n = -2;
#else
      r.n = -1;
#endif
      n->packets.push_back(r);
    }
    if (n->packets.size() > solved_before) {
      PacketRec& r = n->packets.back();
      r.w[0] = n->fe->ang_vel_.x; r.w[1] = n->fe->ang_vel_.y; r.w[2] = n->fe->ang_vel_.z;
    }
    PoseGraphOptimizer& b = *n->be;
    while (b.isReadyFrontendPoses()) {
      WindowRec w{};
      w.beg[0] = b.t_win_beg_.sec; w.beg[1] = b.t_win_beg_.nsec; w.end[0] = b.t_win_end_.sec; w.end[1] = b.t_win_end_.nsec;
      b.getEventSubset(b.t_win_beg_, b.t_win_end_);
      w.n_events = (long long)b.event_subset_.size();
      if (w.n_events) {
        w.f[0] = b.event_subset_.front().ts.sec; w.f[1] = b.event_subset_.front().ts.nsec;
        w.l[0] = b.event_subset_.back().ts.sec; w.l[1] = b.event_subset_.back().ts.nsec;
      }
      w.hash = hash_events(b.event_subset_);
      AngVelMap sub = b.getAngVelSubset(b.t_ang_vel_beg_, b.t_ang_vel_end_);
      b.processTimeWindow(sub);
      w.n_ctrl = (int)b.traj_->size(); w.idx_traj = b.idx_cp_traj_beg_; w.idx_opt = b.idx_cp_opt_beg_; w.num_opt = b.num_cp_opt_;
      w.latest[0] = b.pose_latest_.first.sec; w.latest[1] = b.pose_latest_.first.nsec;
      const auto& q = b.pose_latest_.second.unit_quaternion();
      w.latest_q[0] = q.x(); w.latest_q[1] = q.y(); w.latest_q[2] = q.z(); w.latest_q[3] = q.w();
      for (int k = 0; k < w.n_ctrl; ++k) {
        const auto& c = b.traj_->getControlPose(k).unit_quaternion();
        w.knots.insert(w.knots.end(), {c.x(), c.y(), c.z(), c.w()});
      }
      n->windows.push_back(w);
      b.slideWindow();
    }
  }
}

#ifdef REF_FULL
/* the reference's own front-end solve (src/frontend/local_optim_contrast_gsl.cpp:74-233) on one packet */
extern "C" double ref1p_fe_solve(const dvs_msgs::Event* events, long long n, const unsigned t_ref[2], int W, int H, const double K4[4],
                                 double blur_sigma, int batch_size, int measure, const double omega_in[3], double omega_out[3]) {
  static ros::NodeHandle nh;
  AngVelEstimator est(&nh);
  est.params.warp_opt.blur_sigma = blur_sigma; est.params.warp_opt.event_batch_size = batch_size; est.params.warp_opt.event_sample_rate = 1;
  est.params.process_opt.contrast_measure = measure;
  est.cam_width_ = W; est.cam_height_ = H;
  est.camera_matrix_ = cv::Matx33d(K4[0], 0., K4[2], 0., K4[1], K4[3], 0., 0., 1.);
  est.precomputed_bearing_vectors_.resize((size_t)W * H);
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) est.precomputed_bearing_vectors_[(size_t)y * W + x] = cv::Point3d((x - K4[2]) / K4[0], (y - K4[3]) / K4[1], 1.0);
  est.event_subset_.assign(events, events + n);
  est.time_packet_ = ros::Time(t_ref[0], t_ref[1]);
  cv::Point3d w(omega_in[0], omega_in[1], omega_in[2]);
  const double cost = est.setupProblemAndOptimize_gsl(w);
  omega_out[0] = w.x; omega_out[1] = w.y; omega_out[2] = w.z;
  return cost;
}
#endif

extern "C" void ref1p_node_packet_omega(RefNode* n, int i, double w[3]) { std::memcpy(w, n->packets[(size_t)i].w, sizeof(double) * 3); }
extern "C" void ref1p_node_get_map(RefNode* n, float* IG, unsigned char* times) {
  EventWarper* w = n->be->event_warper_;
  const size_t A = (size_t)w->IG_.rows * w->IG_.cols;
  if (IG) std::memcpy(IG, w->IG_.ptr<float>(), sizeof(float) * A);
  if (times) std::memcpy(times, w->IG_update_times_map_.ptr<unsigned char>(), A);
}
extern "C" int ref1p_node_counts(RefNode* n, int* n_packets, int* n_windows, long long* n_stored) {
  *n_packets = (int)n->packets.size(); *n_windows = (int)n->windows.size(); *n_stored = (long long)n->fe->events_.size();
  return 0;
}
/* packet i: [sec, nsec, n, first sec, first nsec, last sec, last nsec] + hash */
extern "C" void ref1p_node_packet(RefNode* n, int i, long long out7[7], unsigned long long* hash) {
  const PacketRec& r = n->packets[(size_t)i];
  const long long v[7] = {r.sec, r.nsec, r.n, r.f_sec, r.f_nsec, r.l_sec, r.l_nsec};
  std::memcpy(out7, v, sizeof(v));
  *hash = r.hash;
}
/* window i: ints [beg s, beg ns, end s, end ns, n_events, first s, first ns, last s, last ns, n_ctrl, idx_traj, idx_opt, num_opt, latest s, latest ns] */
extern "C" void ref1p_node_window(RefNode* n, int i, long long out15[15], unsigned long long* hash, double latest_q[4], double* knots, int cap) {
  const WindowRec& w = n->windows[(size_t)i];
  const long long v[15] = {w.beg[0], w.beg[1], w.end[0], w.end[1], w.n_events, w.f[0], w.f[1], w.l[0], w.l[1], w.n_ctrl, w.idx_traj, w.idx_opt,
                           w.num_opt, w.latest[0], w.latest[1]};
  std::memcpy(out15, v, sizeof(v));
  *hash = w.hash;
  std::memcpy(latest_q, w.latest_q, sizeof(w.latest_q));
  if (knots && (int)w.knots.size() <= cap) std::memcpy(knots, w.knots.data(), sizeof(double) * w.knots.size());
}
