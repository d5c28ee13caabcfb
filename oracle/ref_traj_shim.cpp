/*
 * ref_traj_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" wrappers (our code) around the REAL first-party trajectory code of
 * the reference, compiled from where it lies (never copied):
 *   src/backend/trajectory.cpp + include/backend/trajectory.h   Linear/CubicTrajectory: constructors, pushbackCtrlPoses,
 *       evaluate (+ the f32 repacking of the knot Jacobians), generateCtrlPoses / fitCtrlPoses (Eigen fullPivHouseholderQr),
 *       incrementalUpdate, CopyAndIncrementalUpdate (window time origin int64(1e9 * double))        SURVEY rows A7, 8f-4
 * on top of the real basalt / Sophus / Eigen headers the reference vendors.  ROS, OpenCV and glog are absent from this image:
 * oracle/stubs/ supplies ros::Time / ros::Duration (restated rostime arithmetic), a container-only cv::Mat and CHECK macros.
 * Built by oracle/Makefile into oracle/_ref/libref_traj.so; pins csrc/traj_init.cu, csrc/pgo.cu's window origin and the
 * oracle's spline wrapper (tests/test_traj_firstparty.py), and generates tests/golden/traj_firstparty.npz.
 */
#include "backend/trajectory.h"

#include <memory>

using cmax_slam::CubicTrajectory;
using cmax_slam::LinearTrajectory;
using cmax_slam::Trajectory;
using cmax_slam::TrajectorySettings;

static Sophus::SO3d so3(const double* q) { return Sophus::SO3d(Eigen::Quaterniond(q[3], q[0], q[1], q[2])); }
static void put(const Sophus::SO3d& r, double* q) {
  q[0] = r.unit_quaternion().x(); q[1] = r.unit_quaternion().y(); q[2] = r.unit_quaternion().z(); q[3] = r.unit_quaternion().w();
}
static std::unique_ptr<Trajectory> make(int order, double t_beg, double dt_knots, const double* knots, int K) {
  std::vector<Sophus::SO3d> cps;
  for (int i = 0; i < K; ++i) cps.push_back(so3(knots + 4 * i));
  if (order == 4) return std::unique_ptr<Trajectory>(new CubicTrajectory(t_beg, dt_knots, cps));
  return std::unique_ptr<Trajectory>(new LinearTrajectory(t_beg, dt_knots, cps));
}

/* Trajectory::generateCtrlPoses(poses, t_beg, t_end) on an empty trajectory created as PoseGraphOptimizer::pushAngVel does
 * (TrajectorySettings{t_traj_beg, ., dt_knots}).  Returns the number of control poses (written to ctrl_xyzw, capacity cap). */
extern "C" int ref1p_generate_ctrl_poses(int order, double dt_knots, const uint32_t t_traj_beg[2], const uint32_t t_beg[2],
                                         const uint32_t t_end[2], const uint32_t* stamps, const double* poses_xyzw, int n,
                                         double* ctrl_xyzw, int cap) {
  TrajectorySettings cfg;
  cfg.t_beg = ros::Time(t_traj_beg[0], t_traj_beg[1]);
  cfg.t_end = ros::Time(t_end[0], t_end[1]);
  cfg.dt_knots = dt_knots;
  std::unique_ptr<Trajectory> traj;
  if (order == 4) traj.reset(new CubicTrajectory(cfg)); else traj.reset(new LinearTrajectory(cfg));
  PoseMap poses;
  for (int i = 0; i < n; ++i) poses.insert(PoseEntry(ros::Time(stamps[2 * i], stamps[2 * i + 1]), so3(poses_xyzw + 4 * i)));
  std::vector<Sophus::SO3d> cps = traj->generateCtrlPoses(poses, ros::Time(t_beg[0], t_beg[1]), ros::Time(t_end[0], t_end[1]));
  if ((int)cps.size() > cap) return -1;
  for (size_t i = 0; i < cps.size(); ++i) put(cps[i], ctrl_xyzw + 4 * i);
  return (int)cps.size();
}

/* evaluate(t, &idx, &jacobian) of a trajectory built with the (double t_beg, dt_knots, ctrl_poses) constructor -- the one
 * CopyAndIncrementalUpdate uses for the window's temporary trajectory.  J: 3 x 3N floats row-major. */
extern "C" int ref1p_evaluate(int order, double t_beg, double dt_knots, const double* knots_xyzw, int K, const uint32_t t[2],
                              double q_xyzw[4], int32_t* idx_beg, float* J) {
  std::unique_ptr<Trajectory> traj = make(order, t_beg, dt_knots, knots_xyzw, K);
  cv::Mat jac;
  int idx = 0;
  Sophus::SO3d r = J ? traj->evaluate(ros::Time(t[0], t[1]), &idx, &jac) : traj->evaluate(ros::Time(t[0], t[1]));
  put(r, q_xyzw);
  if (J) {
    *idx_beg = idx;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3 * order; ++b) J[a * 3 * order + b] = jac.at<float>(a, b);
  }
  return 0;
}

/* CopyAndIncrementalUpdate(drotv, idx_traj_beg, idx_opt_beg) on a trajectory whose origin is a ros::Time (as traj_ in
 * PoseGraphOptimizer), then evaluate the temporary trajectory at t: value, first knot index and f32 Jacobian -- exactly what
 * EventWarper::warpAndAccumulateEvents receives per batch (event_pano_warper.cpp:250). */
extern "C" int ref1p_window_evaluate(int order, const uint32_t t_traj_beg[2], double dt_knots, const double* knots_xyzw, int K,
                                     int idx_traj_beg, int idx_opt_beg, const double* drotv, const uint32_t t[2],
                                     double q_xyzw[4], int32_t* idx_beg, float* J, double* knots_after_xyzw) {
  TrajectorySettings cfg;
  cfg.t_beg = ros::Time(t_traj_beg[0], t_traj_beg[1]);
  cfg.t_end = cfg.t_beg;
  cfg.dt_knots = dt_knots;
  std::unique_ptr<Trajectory> traj;
  if (order == 4) traj.reset(new CubicTrajectory(cfg)); else traj.reset(new LinearTrajectory(cfg));
  std::vector<Sophus::SO3d> cps;
  for (int i = 0; i < K; ++i) cps.push_back(so3(knots_xyzw + 4 * i));
  traj->pushbackCtrlPoses(cps);
  std::vector<Eigen::Vector3d> d;
  for (int i = 0; i < K - idx_opt_beg; ++i) d.emplace_back(drotv[3 * i], drotv[3 * i + 1], drotv[3 * i + 2]);
  std::unique_ptr<Trajectory> tmp(traj->CopyAndIncrementalUpdate(d, idx_traj_beg, idx_opt_beg));
  cv::Mat jac;
  int idx = 0;
  Sophus::SO3d r = tmp->evaluate(ros::Time(t[0], t[1]), &idx, &jac);
  put(r, q_xyzw);
  *idx_beg = idx;
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3 * order; ++b) J[a * 3 * order + b] = jac.at<float>(a, b);
  if (knots_after_xyzw) {      // traj_->incrementalUpdate(drotv, idx_opt_beg) on the full trajectory
    traj->incrementalUpdate(d, idx_opt_beg);
    for (int i = 0; i < K; ++i) put(traj->getControlPose(i), knots_after_xyzw + 4 * i);
  }
  return 0;
}
