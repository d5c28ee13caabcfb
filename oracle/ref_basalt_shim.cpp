/*
 * ref_basalt_shim.cpp -- TEST INFRASTRUCTURE.  A thin extern "C" wrapper (our code) around the
 * REAL basalt::So3Spline / Sophus::SO3d / sophus_utils of the reference, included read-only
 * from /root/reference/thirdparty/basalt-headers (never copied into this repo).  Built by
 * oracle/Makefile into oracle/_ref/libref_basalt.so; used only to pin the restated spline math
 * of cmax_oracle.cpp and to generate tests/golden/spline_*.npz.
 *
 * Wrapped: So3Spline<N>::evaluate (so3_spline.h:218-274), SO3d::exp/log (so3.hpp:247-290,583-619),
 * leftJacobianSO3 / leftJacobianInvSO3 (sophus_utils.hpp:332-414), and the f32 repacking done by
 * Linear/CubicTrajectory::evaluate (src/backend/trajectory.cpp:86-110,329-355).
 */
#include <basalt/spline/so3_spline.h>
#include <basalt/utils/sophus_utils.hpp>
#include <cstdint>
#include <cstring>

template <int N>
static int eval_impl(const double* knots_xyzw, int K, int64_t t0_ns, int64_t dt_ns, int64_t t_ns,
                     double* q_xyzw, double* R, int32_t* start_idx, double* J) {
  basalt::So3Spline<N> spline(dt_ns, t0_ns);
  for (int i = 0; i < K; ++i) {
    Eigen::Quaterniond q(knots_xyzw[4 * i + 3], knots_xyzw[4 * i], knots_xyzw[4 * i + 1], knots_xyzw[4 * i + 2]);
    spline.knotsPushBack(Sophus::SO3d(q));
  }
  if (t_ns < spline.minTimeNs() || t_ns > spline.maxTimeNs()) return -1;
  typename basalt::So3Spline<N>::JacobianStruct Js;
  Sophus::SO3d r = spline.evaluate(t_ns, J ? &Js : nullptr);
  const auto& uq = r.unit_quaternion();
  if (q_xyzw) { q_xyzw[0] = uq.x(); q_xyzw[1] = uq.y(); q_xyzw[2] = uq.z(); q_xyzw[3] = uq.w(); }
  if (R) { Eigen::Matrix3d m = r.matrix(); for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) R[a * 3 + b] = m(a, b); }
  if (J) {
    *start_idx = (int32_t)Js.start_idx;
    for (int k = 0; k < N; ++k)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) J[9 * k + a * 3 + b] = Js.d_val_d_knot[k](a, b);
  }
  return 0;
}

extern "C" int ref_so3_spline_eval(int order, const double* knots_xyzw, int K, int64_t t0_ns, int64_t dt_ns,
                                   int64_t t_ns, double* q_xyzw, double* R, int32_t* start_idx, double* J) {
  switch (order) {
    case 2: return eval_impl<2>(knots_xyzw, K, t0_ns, dt_ns, t_ns, q_xyzw, R, start_idx, J);
    case 3: return eval_impl<3>(knots_xyzw, K, t0_ns, dt_ns, t_ns, q_xyzw, R, start_idx, J);
    case 4: return eval_impl<4>(knots_xyzw, K, t0_ns, dt_ns, t_ns, q_xyzw, R, start_idx, J);
    default: return -2;
  }
}
extern "C" void ref_so3_exp(const double w[3], double q[4]) {
  Sophus::SO3d r = Sophus::SO3d::exp(Eigen::Vector3d(w[0], w[1], w[2]));
  q[0] = r.unit_quaternion().x(); q[1] = r.unit_quaternion().y(); q[2] = r.unit_quaternion().z(); q[3] = r.unit_quaternion().w();
}
extern "C" void ref_so3_log(const double q[4], double w[3]) {
  Sophus::SO3d r(Eigen::Quaterniond(q[3], q[0], q[1], q[2]));
  Eigen::Vector3d v = r.log();
  w[0] = v[0]; w[1] = v[1]; w[2] = v[2];
}
/* exp(x) * knot : the left-multiplicative knot update of trajectory.cpp:236,497 */
extern "C" void ref_knot_update(const double x[3], const double knot_xyzw[4], double out_xyzw[4]) {
  Sophus::SO3d k(Eigen::Quaterniond(knot_xyzw[3], knot_xyzw[0], knot_xyzw[1], knot_xyzw[2]));
  Sophus::SO3d r = Sophus::SO3d::exp(Eigen::Vector3d(x[0], x[1], x[2])) * k;
  out_xyzw[0] = r.unit_quaternion().x(); out_xyzw[1] = r.unit_quaternion().y(); out_xyzw[2] = r.unit_quaternion().z(); out_xyzw[3] = r.unit_quaternion().w();
}
extern "C" void ref_left_jacobians(const double phi[3], double Jl[9], double Jlinv[9]) {
  Eigen::Matrix3d a, b;
  Eigen::Vector3d p(phi[0], phi[1], phi[2]);
  Sophus::leftJacobianSO3(p, a);
  Sophus::leftJacobianInvSO3(p, b);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { Jl[r * 3 + c] = a(r, c); Jlinv[r * 3 + c] = b(r, c); }
}

/* ---- trajectory initialisation ---------------------------------------------------
 * The first-party functions (src/backend/trajectory.cpp, pose_graph_optimizer.cpp) need ROS/OpenCV/glog and do
 * not compile here; the functions below restate their control flow line by line around the REAL third-party
 * calls they make -- Sophus::SO3d exp / log / inverse / operator* and Eigen::FullPivHouseholderQR::solve
 * (Eigen 3.3.9, vendored by the reference) -- so the product's own QR and SO(3) code is pinned against them. */
#include <Eigen/Dense>
#include <cmath>
#include <vector>

static double ref_stamp_diff(uint32_t as, uint32_t an, uint32_t bs, uint32_t bn) {   /* (a - b).toSec(), ros::Duration normalisation */
  long long s = (long long)as - (long long)bs, ns = (long long)an - (long long)bn;
  if (ns < 0) { ns += 1000000000ll; --s; }
  return (double)(int)s + 1e-9 * (double)(int)ns;
}

/* Linear/CubicTrajectory::fitCtrlPoses, trajectory.cpp:112-186 / 357-463.  stamps: n x (sec, nsec). */
extern "C" int ref_fit_ctrl_poses(int order, double dt_knots, double t_beg, int num_cps, const uint32_t* stamps,
                                  const double* poses_xyzw, int n, double* ctrl_xyzw) {
  if (n < num_cps) return -1;
  auto so3 = [](const double* q) { return Sophus::SO3d(Eigen::Quaterniond(q[3], q[0], q[1], q[2])); };
  Sophus::SO3d offset = so3(poses_xyzw);
  Sophus::SO3d offset_inv = offset.inverse();
  Eigen::MatrixXd N = Eigen::MatrixXd::Zero(n, num_cps);
  Eigen::VectorXd Dx(n), Dy(n), Dz(n);
  Eigen::Matrix2d M2 = (Eigen::MatrixXd(2, 2) << 1.0, 0.0, -1.0, 1.0).finished();
  Eigen::Matrix4d M4 = (Eigen::MatrixXd(4, 4) << 1. / 6, 2. / 3, 1. / 6, 0.0, -0.5, 0.0, 0.5, 0.0, 0.5, -1.0, 0.5, 0.0, -1. / 6, 0.5, -0.5, 1. / 6).finished();
  for (int idx = 0; idx < n; ++idx) {
    Sophus::SO3d drot = offset_inv * so3(poses_xyzw + 4 * idx);
    double t = (double)stamps[2 * idx] + 1e-9 * (double)stamps[2 * idx + 1];
    int t_i = std::floor((t - t_beg) / dt_knots);
    double u = (t - (t_i * dt_knots + t_beg)) / dt_knots;
    if (t_i < 0 || t_i + order > num_cps) return -2;
    if (order == 2) {
      Eigen::Matrix<double, 1, 2> U;
      for (int i = 0; i < 2; i++) U(i) = std::pow(u, i);
      Eigen::Matrix<double, 1, 2> N_idx = U * M2;
      for (int j = 0; j < 2; j++) N(idx, t_i + j) = N_idx(j);
    } else {
      Eigen::Matrix<double, 1, 4> U;
      for (int i = 0; i < 4; i++) U(i) = std::pow(u, i);
      Eigen::Matrix<double, 1, 4> N_idx = U * M4;
      for (int j = 0; j < 4; j++) N(idx, t_i + j) = N_idx(j);
    }
    Eigen::Vector3d rv = drot.log();
    Dx(idx) = rv(0); Dy(idx) = rv(1); Dz(idx) = rv(2);
  }
  Eigen::VectorXd Px = N.fullPivHouseholderQr().solve(Dx);
  Eigen::VectorXd Py = N.fullPivHouseholderQr().solve(Dy);
  Eigen::VectorXd Pz = N.fullPivHouseholderQr().solve(Dz);
  for (int i = 0; i < num_cps; ++i) {
    Sophus::SO3d cp = offset * Sophus::SO3d::exp(Eigen::Vector3d(Px(i), Py(i), Pz(i)));
    ctrl_xyzw[4 * i] = cp.unit_quaternion().x(); ctrl_xyzw[4 * i + 1] = cp.unit_quaternion().y();
    ctrl_xyzw[4 * i + 2] = cp.unit_quaternion().z(); ctrl_xyzw[4 * i + 3] = cp.unit_quaternion().w();
  }
  return 0;
}

/* PoseGraphOptimizer::integrateAngVel, pose_graph_optimizer.cpp:191-222.  state = (prev sec, prev nsec) + prev w. */
extern "C" int ref_integrate_ang_vel(const uint32_t latest_stamp[2], const double latest_xyzw[4], uint32_t prev_stamp[2],
                                     double prev_w[3], int first_time_window, const uint32_t* stamps, const double* w, int m,
                                     uint32_t* out_stamps, double* out_xyzw) {
  Sophus::SO3d cur(Eigen::Quaterniond(latest_xyzw[3], latest_xyzw[0], latest_xyzw[1], latest_xyzw[2]));
  uint32_t cs = latest_stamp[0], cn = latest_stamp[1];
  int n = 0;
  for (int i = 0; i < m; ++i) {
    const uint32_t s = stamps[2 * i], ns = stamps[2 * i + 1];
    const bool newer = s > prev_stamp[0] || (s == prev_stamp[0] && ns > prev_stamp[1]);
    if (!newer && !first_time_window) continue;
    const double dt = ref_stamp_diff(s, ns, cs, cn);
    Eigen::Vector3d drotv = dt * ((Eigen::Vector3d(prev_w[0], prev_w[1], prev_w[2]) + Eigen::Vector3d(w[3 * i], w[3 * i + 1], w[3 * i + 2])) / 2.0);
    cs = s; cn = ns;
    cur = cur * Sophus::SO3d::exp(drotv);
    out_stamps[2 * n] = s; out_stamps[2 * n + 1] = ns;
    out_xyzw[4 * n] = cur.unit_quaternion().x(); out_xyzw[4 * n + 1] = cur.unit_quaternion().y();
    out_xyzw[4 * n + 2] = cur.unit_quaternion().z(); out_xyzw[4 * n + 3] = cur.unit_quaternion().w();
    ++n;
    prev_stamp[0] = s; prev_stamp[1] = ns;
    prev_w[0] = w[3 * i]; prev_w[1] = w[3 * i + 1]; prev_w[2] = w[3 * i + 2];
  }
  return n;
}
