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
