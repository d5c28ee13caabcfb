// so3_math.cuh -- SO(3) exp/log, left Jacobians and the cumulative uniform SO(3) B-spline
// (value + left-perturbation knot Jacobians) in f64, host + device.
//
// Computes what the reference obtains from
//   Sophus::SO3d::exp / log / matrix / inverse / operator*   thirdparty/basalt-headers/thirdparty/Sophus/sophus/so3.hpp:229-339,583-619
//   Sophus::leftJacobianSO3 / leftJacobianInvSO3              thirdparty/basalt-headers/include/basalt/utils/sophus_utils.hpp:332-414
//   basalt::So3Spline<N>::evaluate(time_ns, &J)               thirdparty/basalt-headers/include/basalt/spline/so3_spline.h:218-274
//   computeBlendingMatrix<N,double,true>                      thirdparty/basalt-headers/include/basalt/spline/spline_common.h:69-100
// written from the published formulas (arXiv:1911.08860) with the same branch thresholds, so the
// results agree with the reference to rounding.  Orders N = 2 (linear) and N = 4 (cubic) are the
// two the reference instantiates (include/backend/trajectory.h:83,142).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define CMAXB_HD __host__ __device__ __forceinline__
#else
#define CMAXB_HD inline
#endif

namespace cmaxb {

struct Vec3 { double x, y, z; };
struct Mat3 { double m[9]; };  // row-major
struct Quat { double x, y, z, w; };

constexpr double kSophusEps = 1e-10;  // Sophus::Constants<double>::epsilon()
constexpr double kPi = 3.14159265358979323846;

CMAXB_HD Mat3 mat_identity() { Mat3 r; for (int i = 0; i < 9; ++i) r.m[i] = 0.0; r.m[0] = r.m[4] = r.m[8] = 1.0; return r; }
CMAXB_HD Mat3 mat_mul(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = a.m[i * 3] * b.m[j];
      s += a.m[i * 3 + 1] * b.m[3 + j];
      s += a.m[i * 3 + 2] * b.m[6 + j];
      r.m[i * 3 + j] = s;
    }
  return r;
}
CMAXB_HD Mat3 mat_hat(const Vec3& p) {
  Mat3 r;
  r.m[0] = 0.0;  r.m[1] = -p.z; r.m[2] = p.y;
  r.m[3] = p.z;  r.m[4] = 0.0;  r.m[5] = -p.x;
  r.m[6] = -p.y; r.m[7] = p.x;  r.m[8] = 0.0;
  return r;
}
CMAXB_HD Quat quat_normalized(Quat q) {
  const double len = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= len; q.y /= len; q.z /= len; q.w /= len;
  return q;
}
// group product; the SO3(quaternion) constructor re-normalises every product
CMAXB_HD Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return quat_normalized(r);
}
CMAXB_HD Quat quat_inv(const Quat& a) { Quat c; c.x = -a.x; c.y = -a.y; c.z = -a.z; c.w = a.w; return quat_normalized(c); }
// unit quaternion -> rotation matrix (Eigen::QuaternionBase::toRotationMatrix operation order)
CMAXB_HD Mat3 quat_to_mat(const Quat& q) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 r;
  r.m[0] = 1.0 - (tyy + tzz); r.m[1] = txy - twz;         r.m[2] = txz + twy;
  r.m[3] = txy + twz;         r.m[4] = 1.0 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy;         r.m[7] = tyz + twx;         r.m[8] = 1.0 - (txx + tyy);
  return r;
}
CMAXB_HD Quat so3_exp(const Vec3& o) {
  const double theta_sq = o.x * o.x + o.y * o.y + o.z * o.z;
  double imag, real;
  if (theta_sq < kSophusEps * kSophusEps) {
    const double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    const double theta = sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  Quat q; q.x = imag * o.x; q.y = imag * o.y; q.z = imag * o.z; q.w = real;
  return q;
}
CMAXB_HD Vec3 so3_log(const Quat& q) {
  const double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  const double w = q.w;
  double f;
  if (squared_n < kSophusEps * kSophusEps) {
    const double squared_w = w * w;
    f = 2.0 / w - (2.0 / 3.0) * squared_n / (w * squared_w);
  } else {
    const double n = sqrt(squared_n);
    if (fabs(w) < kSophusEps) f = (w > 0.0) ? kPi / n : -kPi / n;
    else f = 2.0 * atan(n / w) / n;
  }
  Vec3 r; r.x = f * q.x; r.y = f * q.y; r.z = f * q.z;
  return r;
}
CMAXB_HD Mat3 so3_left_jacobian(const Vec3& phi) {
  const double n2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z;
  const Mat3 ph = mat_hat(phi);
  const Mat3 ph2 = mat_mul(ph, ph);
  Mat3 J = mat_identity();
  if (n2 > kSophusEps) {
    const double n = sqrt(n2);
    const double n3 = n2 * n;
    const double a = 1.0 - cos(n);
    const double b = n - sin(n);
    for (int i = 0; i < 9; ++i) J.m[i] += ph.m[i] * a / n2;
    for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] * b / n3;
  } else {
    for (int i = 0; i < 9; ++i) J.m[i] += ph.m[i] / 2.0;
    for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] / 6.0;
  }
  return J;
}
CMAXB_HD Mat3 so3_left_jacobian_inv(const Vec3& phi) {
  const double n2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z;
  const Mat3 ph = mat_hat(phi);
  const Mat3 ph2 = mat_mul(ph, ph);
  Mat3 J = mat_identity();
  for (int i = 0; i < 9; ++i) J.m[i] -= ph.m[i] / 2.0;
  if (n2 > kSophusEps) {
    const double n = sqrt(n2);
    if (n < kPi - sqrt(kSophusEps)) {
      const double c = 1.0 / n2 - (1.0 + cos(n)) / (2.0 * n * sin(n));
      for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] * c;
    } else {
      for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] / (kPi * kPi);
    }
  } else {
    for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] / 12.0;
  }
  return J;
}

// Cumulative blending coefficients coeff[0..N) of a uniform B-spline of order N at u in [0,1):
// coeff = M_cumulative * [1 u u^2 u^3]^T.  The matrices below are the values of
// computeBlendingMatrix<N,double,true>() for N = 2 and N = 4 (checked against the real
// basalt code in tests/test_so3_math.py).
template <int N>
CMAXB_HD void spline_cum_coeffs(double u, double* coeff);
template <>
CMAXB_HD void spline_cum_coeffs<2>(double u, double* coeff) {
  coeff[0] = 1.0;
  coeff[1] = u;
}
template <>
CMAXB_HD void spline_cum_coeffs<4>(double u, double* coeff) {
  const double u2 = u * u, u3 = u2 * u;
  // rows of M_c4 = 1/6 * [6 0 0 0; 5 3 -3 1; 1 3 3 -2; 0 0 0 1]
  coeff[0] = 1.0;
  coeff[1] = (5.0 / 6.0) + (3.0 / 6.0) * u + (-3.0 / 6.0) * u2 + (1.0 / 6.0) * u3;
  coeff[2] = (1.0 / 6.0) + (3.0 / 6.0) * u + (3.0 / 6.0) * u2 + (-2.0 / 6.0) * u3;
  coeff[3] = (1.0 / 6.0) * u3;
}

// So3Spline<N>::evaluate.  knots: the whole (already updated) knot array; s = first knot of the
// segment; u = fractional position.  J (N blocks) may be null.
template <int N>
CMAXB_HD Quat so3_spline_eval(const Quat* knots, int s, double u, Mat3* J) {
  double coeff[N];
  spline_cum_coeffs<N>(u, coeff);
  Quat res = knots[s];
  Mat3 J_helper = mat_identity();
  for (int i = 0; i < N - 1; ++i) {
    const Quat p0 = knots[s + i];
    const Quat p1 = knots[s + i + 1];
    const Quat p0inv = quat_inv(p0);
    const Quat r01 = quat_mul(p0inv, p1);
    const Vec3 delta = so3_log(r01);
    Vec3 kdelta; kdelta.x = delta.x * coeff[i + 1]; kdelta.y = delta.y * coeff[i + 1]; kdelta.z = delta.z * coeff[i + 1];
    if (J) {
      const Mat3 Jl_inv_delta = so3_left_jacobian_inv(delta);
      const Mat3 Jl_k_delta = so3_left_jacobian(kdelta);
      J[i] = J_helper;
      Mat3 Rs = quat_to_mat(res);
      for (int e = 0; e < 9; ++e) Rs.m[e] *= coeff[i + 1];
      J_helper = mat_mul(mat_mul(mat_mul(Rs, Jl_k_delta), Jl_inv_delta), quat_to_mat(p0inv));
      for (int e = 0; e < 9; ++e) J[i].m[e] -= J_helper.m[e];
    }
    res = quat_mul(res, so3_exp(kdelta));
  }
  if (J) J[N - 1] = J_helper;
  return res;
}

}  // namespace cmaxb
