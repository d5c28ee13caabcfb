/*
 * cmax_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Single-threaded, line-faithful C++17 restatement of the CMax-SLAM contrast-maximisation
 * inner loop of tub-rip/cmax_slam @ 12342de.  Every function cites the reference file:line it
 * follows (paths relative to the reference root).  Precision choices mirror the reference:
 * f64 geometry, f32 bilinear weights / accumulators / Jacobian rows, f32 separable blur,
 * f64 reductions, sequential event order.  Compile with -ffp-contract=off (the reference is
 * built for baseline x86-64, i.e. without FMA contraction).
 *
 * Self-contained on purpose (no Eigen / Sophus / OpenCV / ROS): those are restated here and the
 * restatement is pinned against the real thing where it exists -- see cmax_oracle.h.
 */
#include "cmax_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <atomic>
#include <thread>

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }

// ----------------------------------------------------------------------------------------------
// ros::Time / ros::Duration semantics (roscpp Noetic, rostime/{time.h,duration.h,impl/*.h};
// un-vendored third party, call sites local_image_warped_events.cpp:68-75,
// event_pano_warper.cpp:239-242,298, trajectory.cpp:89,332).
// ----------------------------------------------------------------------------------------------
struct RosTime { uint32_t sec, nsec; };
struct RosDur { int32_t sec, nsec; };

inline double toSec(RosTime t) { return (double)t.sec + 1e-9 * (double)t.nsec; }
inline double toSec(RosDur d) { return (double)d.sec + 1e-9 * (double)d.nsec; }
inline int64_t toNSec(RosTime t) { return (int64_t)((uint64_t)t.sec * 1000000000ull + (uint64_t)t.nsec); }

// TimeBase::operator-(T) -> Duration((int32)sec - (int32)rhs.sec, (int32)nsec - (int32)rhs.nsec),
// normalised so that 0 <= nsec < 1e9.
inline RosDur sub(RosTime a, RosTime b) {
  int64_t s = (int64_t)a.sec - (int64_t)b.sec;
  int64_t ns = (int64_t)a.nsec - (int64_t)b.nsec;
  while (ns >= 1000000000ll) { ns -= 1000000000ll; ++s; }
  while (ns < 0) { ns += 1000000000ll; --s; }
  return RosDur{(int32_t)s, (int32_t)ns};
}
// DurationBase::fromSec: sec = floor(d); nsec = boost::math::round((d-sec)*1e9); rollover.
inline RosDur durFromSec(double d) {
  int64_t sec64 = (int64_t)std::floor(d);
  int32_t sec = (int32_t)sec64;
  int32_t nsec = (int32_t)std::round((d - (double)sec) * 1e9);  // half away from zero
  int32_t rollover = nsec / 1000000000;
  sec += rollover;
  nsec %= 1000000000;
  return RosDur{sec, nsec};
}
inline RosDur mul(RosDur d, double scale) { return durFromSec(toSec(d) * scale); }
inline RosTime add(RosTime t, RosDur d) {
  int64_t s = (int64_t)t.sec + d.sec;
  int64_t ns = (int64_t)t.nsec + d.nsec;
  while (ns >= 1000000000ll) { ns -= 1000000000ll; ++s; }
  while (ns < 0) { ns += 1000000000ll; --s; }
  return RosTime{(uint32_t)s, (uint32_t)ns};
}
inline bool lessThan(RosTime a, RosTime b) { return a.sec < b.sec || (a.sec == b.sec && a.nsec < b.nsec); }
inline RosTime evTime(const orc_event& e) { return RosTime{e.sec, e.nsec}; }

// batch mid-time: time_first + (time_last - time_first) * 0.5
// (local_image_warped_events.cpp:68-73, event_pano_warper.cpp:239-242)
inline RosTime batchMid(RosTime first, RosTime last) { return add(first, mul(sub(last, first), 0.5)); }

// ----------------------------------------------------------------------------------------------
// Small fixed-size math (restating cv::Matx / Eigen 3x3 / Sophus::SO3d arithmetic order).
// ----------------------------------------------------------------------------------------------
struct V3 { double x, y, z; };
struct M3 { double m[9]; double& operator()(int r, int c) { return m[r * 3 + c]; } double operator()(int r, int c) const { return m[r * 3 + c]; } };
struct Quat { double x, y, z, w; };

inline M3 matmul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = a(i, 0) * b(0, j);
      s += a(i, 1) * b(1, j);
      s += a(i, 2) * b(2, j);
      r(i, j) = s;
    }
  return r;
}
inline M3 identity() { M3 r{}; r(0, 0) = r(1, 1) = r(2, 2) = 1.0; return r; }
inline M3 scale(const M3& a, double s) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * s; return r; }
inline M3 hat(const V3& p) {  // Sophus::SO3::hat, so3.hpp:669-690
  M3 r{};
  r(0, 1) = -p.z; r(0, 2) = p.y;
  r(1, 0) = p.z;  r(1, 2) = -p.x;
  r(2, 0) = -p.y; r(2, 1) = p.x;
  return r;
}

// Sophus::SO3 class invariant: every construction from a quaternion normalises (so3.hpp:297-305,481-487).
inline Quat normalized(Quat q) {
  double len = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return Quat{q.x / len, q.y / len, q.z / len, q.w / len};
}
// so3.hpp:325-339 group multiplication (+ normalisation by the SO3(quat) constructor)
inline Quat qmul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return normalized(r);
}
// so3.hpp:229-231 inverse = SO3(conjugate)
inline Quat qinv(const Quat& a) { return normalized(Quat{-a.x, -a.y, -a.z, a.w}); }
// so3.hpp:310-312 -> Eigen::QuaternionBase::toRotationMatrix (Eigen/src/Geometry/Quaternion.h)
inline M3 qmat(const Quat& q) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r(0, 0) = 1.0 - (tyy + tzz); r(0, 1) = txy - twz;         r(0, 2) = txz + twy;
  r(1, 0) = txy + twz;         r(1, 1) = 1.0 - (txx + tzz); r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy;         r(2, 1) = tyz + twx;         r(2, 2) = 1.0 - (txx + tyy);
  return r;
}
constexpr double kEps = 1e-10;  // Sophus::Constants<double>::epsilon(), sophus/common.hpp:94

// so3.hpp:583-619 expAndTheta
inline Quat so3exp(const V3& o) {
  const double theta_sq = o.x * o.x + o.y * o.y + o.z * o.z;
  double imag, real;
  if (theta_sq < kEps * kEps) {
    const double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    const double theta = std::sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  // assigned directly to unit_quaternion_nonconst(): no normalisation here
  return Quat{imag * o.x, imag * o.y, imag * o.z, real};
}
// so3.hpp:247-290 logAndTheta
inline V3 so3log(const Quat& q) {
  const double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  const double w = q.w;
  double f;
  if (squared_n < kEps * kEps) {
    const double squared_w = w * w;
    f = 2.0 / w - (2.0 / 3.0) * squared_n / (w * squared_w);
  } else {
    const double n = std::sqrt(squared_n);
    if (std::fabs(w) < kEps) {
      f = (w > 0.0) ? M_PI / n : -M_PI / n;
    } else {
      f = 2.0 * std::atan(n / w) / n;
    }
  }
  return V3{f * q.x, f * q.y, f * q.z};
}
// basalt/utils/sophus_utils.hpp:332-362
inline M3 leftJacobianSO3(const V3& phi) {
  const double n2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z;
  const M3 ph = hat(phi);
  const M3 ph2 = matmul(ph, ph);
  M3 J = identity();
  if (n2 > kEps) {
    const double n = std::sqrt(n2);
    const double n3 = n2 * n;
    const double a = (1.0 - std::cos(n));
    const double b = (n - std::sin(n));
    for (int i = 0; i < 9; ++i) J.m[i] += ph.m[i] * a / n2;
    for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] * b / n3;
  } else {
    for (int i = 0; i < 9; ++i) J.m[i] += ph.m[i] / 2.0;
    for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] / 6.0;
  }
  return J;
}
// basalt/utils/sophus_utils.hpp:372-414
inline M3 leftJacobianInvSO3(const V3& phi) {
  const double n2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z;
  const M3 ph = hat(phi);
  const M3 ph2 = matmul(ph, ph);
  M3 J = identity();
  for (int i = 0; i < 9; ++i) J.m[i] -= ph.m[i] / 2.0;
  if (n2 > kEps) {
    const double n = std::sqrt(n2);
    if (n < M_PI - std::sqrt(kEps)) {
      const double c = (1.0 / n2 - (1.0 + std::cos(n)) / (2.0 * n * std::sin(n)));
      for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] * c;
    } else {
      for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] / (M_PI * M_PI);
    }
  } else {
    for (int i = 0; i < 9; ++i) J.m[i] += ph2.m[i] / 12.0;
  }
  return J;
}

// basalt/spline/spline_common.h:69-100 computeBlendingMatrix<N,double,true> (cumulative)
inline double binom(int n, int k) {
  if (k > n) return 0;
  double r = 1;
  for (int d = 1; d <= k; ++d) { r *= n--; r /= d; }
  return r;
}
inline void blendingMatrixCumulative(int N, double* M /* N*N row-major */) {
  std::vector<double> m(N * N, 0.0);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double sum = 0;
      for (int s = j; s < N; ++s)
        sum += std::pow(-1.0, s - j) * binom(N, s - j) * std::pow(N - s - 1.0, N - 1.0 - i);
      m[j * N + i] = binom(N - 1, N - 1 - i) * sum;
    }
  for (int i = 0; i < N; ++i)
    for (int j = i + 1; j < N; ++j)
      for (int c = 0; c < N; ++c) m[i * N + c] += m[j * N + c];
  uint64_t fact = 1;
  for (int i = 2; i < N; ++i) fact *= i;
  for (int i = 0; i < N * N; ++i) M[i] = m[i] / (double)fact;
}

// basalt/spline/so3_spline.h:218-274 So3Spline<N>::evaluate(time_ns, &J)
// J: N blocks of 3x3 (d_val_d_knot[i]).  Returns 0, or -1 when the time is outside the spline
// (BASALT_ASSERT aborts in the reference, so3_spline.h:221-230).
int splineEvaluate(int N, const Quat* knots, int K, int64_t t0_ns, int64_t dt_ns, int64_t t_ns,
                   Quat* res_out, int* start_idx, M3* J) {
  const int DEG = N - 1;
  const int64_t st_ns = t_ns - t0_ns;
  if (st_ns < 0) return -1;
  const int64_t s = st_ns / dt_ns;
  const double u = double(st_ns % dt_ns) / double(dt_ns);
  if (s < 0 || (int64_t)(s + N) > (int64_t)K) return -1;

  double B[16];
  blendingMatrixCumulative(N, B);
  // baseCoeffsWithTime<0>: p = [1, u, u^2, ...]   (so3_spline.h:754-772)
  double p[4] = {0, 0, 0, 0};
  p[0] = 1.0;
  double ti = u;
  for (int j = 1; j < N; ++j) { p[j] = 1.0 * ti; ti = ti * u; }
  double coeff[4];
  for (int i = 0; i < N; ++i) {
    double acc = 0;
    for (int j = 0; j < N; ++j) acc += B[i * N + j] * p[j];
    coeff[i] = acc;
  }

  Quat res = knots[s];
  M3 J_helper = identity();
  if (J) *start_idx = (int)s;
  for (int i = 0; i < DEG; ++i) {
    const Quat& p0 = knots[s + i];
    const Quat& p1 = knots[s + i + 1];
    const Quat r01 = qmul(qinv(p0), p1);
    const V3 delta = so3log(r01);
    const V3 kdelta{delta.x * coeff[i + 1], delta.y * coeff[i + 1], delta.z * coeff[i + 1]};
    if (J) {
      const M3 Jl_inv_delta = leftJacobianInvSO3(delta);
      const M3 Jl_k_delta = leftJacobianSO3(kdelta);
      J[i] = J_helper;
      J_helper = matmul(matmul(matmul(scale(qmat(res), coeff[i + 1]), Jl_k_delta), Jl_inv_delta),
                        qmat(qinv(p0)));
      for (int e = 0; e < 9; ++e) J[i].m[e] -= J_helper.m[e];
    }
    res = qmul(res, so3exp(kdelta));
  }
  if (J) J[DEG] = J_helper;
  *res_out = res;
  if (!J && start_idx) *start_idx = (int)s;
  return 0;
}

// ----------------------------------------------------------------------------------------------
// OpenCV image primitives (un-vendored third party; restated, pinned vs cv2 4.13 by our tests).
// ----------------------------------------------------------------------------------------------
// cv::GaussianBlur(src,dst,Size(0,0),sigma) for CV_32F: ksize = cvRound(sigma*4*2+1)|1,
// kernel = getGaussianKernel(ksize, sigma, CV_32F), separable, BORDER_REFLECT_101.
int gaussianKernel(double sigma, std::vector<float>& taps) {
  int ksize = (int)std::lrint(sigma * 4 * 2 + 1) | 1;
  std::vector<double> k(ksize);
  const double scale2X = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < ksize; ++i) {
    const double x = i - (ksize - 1) * 0.5;
    const double t = std::exp(scale2X * x * x);
    k[i] = t;
    sum += t;
  }
  sum = 1.0 / sum;
  taps.resize(ksize);
  for (int i = 0; i < ksize; ++i) taps[i] = (float)(k[i] * sum);
  return ksize;
}
inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}
// Row pass: s = k[0]*src[x-r]; s = fma(k[j], src[x-r+j], s) left to right (RowVec_32f, AVX2/FMA
// dispatch); column pass: symmetric form s = k[r]*c; s = fma(k[r+j], down_j + up_j, s)
// (SymmColumnVec_32f).  This op order is BIT-EXACT against cv2 4.13.0 on x86-64 with FMA
// (tests/test_oracle_opencv.py, tests/golden/blur_*.npz).
void gaussianBlur(const float* src, float* dst, int W, int H, int C, double sigma) {
  std::vector<float> taps;
  const int ksize = gaussianKernel(sigma, taps);
  const int r = ksize / 2;
  std::vector<float> tmp((size_t)W * H * C);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) {
        float s = 0.f;
        for (int k = 0; k < ksize; ++k) {
          const int xs = reflect101(x + k - r, W);
          const float v = src[((size_t)y * W + xs) * C + c];
          s = (k == 0) ? taps[0] * v : std::fmaf(taps[k], v, s);
        }
        tmp[((size_t)y * W + x) * C + c] = s;
      }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) {
        float s = taps[r] * tmp[((size_t)y * W + x) * C + c];
        for (int j = 1; j <= r; ++j) {
          const int y0 = reflect101(y - j, H), y1 = reflect101(y + j, H);
          s = std::fmaf(taps[r + j], tmp[((size_t)y1 * W + x) * C + c] + tmp[((size_t)y0 * W + x) * C + c], s);
        }
        dst[((size_t)y * W + x) * C + c] = s;
      }
}
// cv::meanStdDev for CV_32FC1: f64 sum and sum of squares; var = max(sq/N - mean^2, 0).
void meanStdDev(const float* img, int64_t n, double* mean, double* stddev) {
  double s = 0, sq = 0;
  for (int64_t i = 0; i < n; ++i) { const double v = img[i]; s += v; sq += v * v; }
  const double m = s / (double)n;
  double var = sq / (double)n - m * m;
  if (var < 0) var = 0;
  *mean = m;
  *stddev = std::sqrt(var);
}
inline double meanOf(const float* img, int64_t n, int stride = 1, int off = 0) {
  double s = 0;
  for (int64_t i = 0; i < n; ++i) s += (double)img[i * stride + off];
  return s / (double)n;
}

// contrast_Variance / contrast_MeanSquare over one image and P derivative planes accessed as
// plane p -> deriv[i*stride + p*pstride...]: a generic accessor keeps FE (interleaved HxWx3,
// after cv::split) and BE (planar vector<Mat>) on one code path.
// FE: local_focus_funcs.cpp:9-44,82-120; BE: global_focus_funcs.cpp:11-47,52-80.
struct PlaneView { const float* base; int64_t elem_stride; };
double computeContrast(const float* img, int64_t n, const std::vector<PlaneView>& planes,
                       double* grad, int measure) {
  if (measure == 1) {
    // MEAN SQUARE: cv::norm(img, NORM_L2SQR)/N ; g_i = 2*mean(img.mul(ch_i))
    double sq = 0;
    for (int64_t i = 0; i < n; ++i) { const double v = img[i]; sq += v * v; }
    const double contrast = sq / (double)n;
    if (grad)
      for (size_t p = 0; p < planes.size(); ++p) {
        double s = 0;
        for (int64_t i = 0; i < n; ++i) {
          const float prod = img[i] * planes[p].base[i * planes[p].elem_stride];
          s += (double)prod;
        }
        grad[p] = 2.0 * (s / (double)n);
      }
    return contrast;
  }
  double mean, stddev;
  meanStdDev(img, n, &mean, &stddev);
  const double contrast = stddev * stddev;
  if (grad) {
    // img_zeromean = 2.*(img - mean): MatExpr folds to convertTo(alpha=2, beta=-2*mean), f32 arithmetic
    std::vector<float> zm((size_t)n);
    const float a = 2.0f, b = (float)(-2.0 * mean);
    for (int64_t i = 0; i < n; ++i) zm[i] = img[i] * a + b;
    for (size_t p = 0; p < planes.size(); ++p) {
      const float* ch = planes[p].base;
      const int64_t st = planes[p].elem_stride;
      const double mean_ch = meanOf(ch, n, (int)st, 0);
      const float neg = (float)(-mean_ch);  // channels - mean: cv::add with the scalar cast to f32
      double s = 0;
      for (int64_t i = 0; i < n; ++i) {
        const float d = ch[i * st] + neg;
        const float prod = zm[i] * d;
        s += (double)prod;
      }
      grad[p] = s / (double)n;
    }
  }
  return contrast;
}

}  // namespace

// ================================================================================================
// Front-end
// ================================================================================================
// ---- first-party geometry of the reference, restated; pinned against the reference's OWN sources compiled with
// container stubs (oracle/ref_geom_shim.cpp -> oracle/_ref/libref_geom.so, tests/test_oracle_geom.py) -----------------
// canonicalProjection, src/utils/image_geom_util.cpp:24-41
inline void canonicalProjectionO(const V3& p, double* u, double* v, double Jp[2][3]) {
  const double inv = 1.0 / p.z;
  *u = p.x * inv;
  *v = p.y * inv;
  Jp[0][0] = inv; Jp[0][1] = 0.0; Jp[0][2] = -*u * inv;
  Jp[1][0] = 0.0; Jp[1][1] = inv; Jp[1][2] = -*v * inv;
}
// applyIntrinsics, src/utils/image_geom_util.cpp:7-22
inline void applyIntrinsicsO(double u, double v, double fx, double fy, double cx, double cy, double* px, double* py) {
  *px = fx * u + cx;
  *py = fy * v + cy;
}
// cross2Matrix, include/utils/image_geom_util.h:5-8
inline void cross2MatrixO(const V3& m, double M[3][3]) {
  M[0][0] = 0;    M[0][1] = -m.z; M[0][2] = m.y;
  M[1][0] = m.z;  M[1][1] = 0;    M[1][2] = -m.x;
  M[2][0] = -m.y; M[2][1] = m.x;  M[2][2] = 0;
}
// dvs::EquirectangularCamera::projectToImage, include/backend/equirectangular_camera.h:18-45
inline void equirectProjectO(double wx, double wy, double wz, double fx, double fy, double cxp, double cyp, double* px, double* py,
                             float dpm_drb[2][3]) {
  const double phi = std::atan2(wx, wz);
  const double theta = std::asin(wy / std::sqrt(wx * wx + wy * wy + wz * wz));
  const double rho = std::sqrt(wx * wx + wy * wy + wz * wz);
  const double Ydivrho = wy / rho;
  const double XdivZ = wx / wz;
  const double tmp1 = fx / ((1 + XdivZ * XdivZ) * wz);
  const double tmp2 = -fy / std::sqrt(1 - Ydivrho * Ydivrho);
  const double tmp3 = Ydivrho / (rho * rho);
  dpm_drb[0][0] = (float)tmp1;
  dpm_drb[0][1] = 0.f;
  dpm_drb[0][2] = (float)(-tmp1 * XdivZ);
  dpm_drb[1][0] = (float)(tmp2 * tmp3 * wx);
  dpm_drb[1][1] = (float)(tmp2 * (tmp3 * wy - 1 / rho));
  dpm_drb[1][2] = (float)(tmp2 * tmp3 * wz);
  *px = cxp + phi * fx;
  *py = cyp + theta * fy;
}
// hooks for the pinning tests
extern "C" void orc_geom_pinhole(const double p[3], const double K4[4], double uv[2], double px[2], double Jproj[6]) {
  double Jp[2][3];
  canonicalProjectionO(V3{p[0], p[1], p[2]}, &uv[0], &uv[1], Jp);
  applyIntrinsicsO(uv[0], uv[1], K4[0], K4[1], K4[2], K4[3], &px[0], &px[1]);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) Jproj[i * 3 + j] = Jp[i][j];
}
extern "C" void orc_geom_cross2matrix(const double v[3], double M[9]) {
  double m[3][3];
  cross2MatrixO(V3{v[0], v[1], v[2]}, m);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i * 3 + j] = m[i][j];
}
extern "C" void orc_geom_equirect(const double w[3], int PW, int PH, double px[2], float J[6]) {
  // EquirectangularCamera(pano_size, 360, 180): centre and focal lengths, equirectangular_camera.h:11-16,64-67
  const double cxp = (double)PW / 2.0, cyp = (double)PH / 2.0;
  const double fx = double((PW / 360.0) * 180.0 / M_PI), fy = double((PH / 180.0) * 180.0 / M_PI);
  float d[2][3];
  equirectProjectO(w[0], w[1], w[2], fx, fy, cxp, cyp, &px[0], &px[1], d);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) J[i * 3 + j] = d[i][j];
}

extern "C" int orc_fe_eval(const orc_fe_args* a, const double omega[3], int want_grad, orc_fe_out* out) {
  const int W = a->width, H = a->height;
  const int64_t A = (int64_t)W * H;
  const int64_t n = a->n_events;
  if (a->batch_size <= 0) return fail("batch_size must be > 0");
  // computeImageOfWarpedEvents, local_image_warped_events.cpp:10-39
  std::vector<float> iwe((size_t)A, 0.f);
  std::vector<float> deriv(want_grad ? (size_t)A * 3 : 0, 0.f);
  int64_t n_in = 0;

  for (int64_t beg = 0; beg < n; beg += a->batch_size) {
    const int64_t end = std::min<int64_t>(beg + a->batch_size, n);
    // warpAndAccumulateEvents, local_image_warped_events.cpp:59-170
    const RosTime t_first = evTime(a->events[beg]);
    const RosTime t_last = evTime(a->events[end - 1]);
    const RosDur time_dt = sub(t_last, t_first);
    if (!(toSec(time_dt) >= 0.)) return fail("Events must span a non-negative time interval");  // :72
    const RosTime time_batch = add(t_first, mul(time_dt, 0.5));                                   // :73
    const double dt = toSec(time_batch) - a->t_ref_sec;                                           // :75
    const V3 delta{omega[0] * dt, omega[1] * dt, omega[2] * dt};                                  // :76

    for (int64_t i = beg; i < end; ++i) {
      const orc_event& e = a->events[i];
      if (e.x >= W || e.y >= H) return fail("event outside the sensor (vector::at would throw)");
      const double* bp = a->lut_xyz + 3 * ((int64_t)e.y * W + e.x);                               // :100
      const V3 b{bp[0], bp[1], bp[2]};
      // point_3D + delta_rot.cross(point_3D)  (cv::Point3_::cross)                               // :101
      const V3 cr{delta.y * b.z - delta.z * b.y, delta.z * b.x - delta.x * b.z, delta.x * b.y - delta.y * b.x};
      const V3 pr{b.x + cr.x, b.y + cr.y, b.z + cr.z};

      // canonicalProjection + applyIntrinsics (pinned against the reference's own source: tests/test_oracle_geom.py)
      double u, v, px, py, Jp[2][3];
      canonicalProjectionO(pr, &u, &v, Jp);
      applyIntrinsicsO(u, v, a->fx, a->fy, a->cx, a->cy, &px, &py);

      double J[2][3] = {{0, 0, 0}, {0, 0, 0}};
      if (want_grad) {
        // cross2Matrix((-dt)*point_3D), image_geom_util.h:5-8, local_image_warped_events.cpp:110
        const V3 mv{(-dt) * b.x, (-dt) * b.y, (-dt) * b.z};
        double Mx[3][3];
        cross2MatrixO(mv, Mx);
        double Jc[2][3];
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3; ++c) {  // cv::Matx product: s = 0; s += a(i,k)*b(k,j)
            double s = 0;
            for (int k = 0; k < 3; ++k) s += Jp[r][k] * Mx[k][c];
            Jc[r][c] = s;
          }
        const double Ji[2][2] = {{a->fx, 0.}, {0., a->fy}};
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 2; ++k) s += Ji[r][k] * Jc[k][c];
            J[r][c] = s;
          }
      }

      // :139-142  truncation, bounds
      int32_t cell = -1;
      if (std::fabs(px) < 2e9 && std::fabs(py) < 2e9) {
        const int xx = (int)px, yy = (int)py;
        if (1 <= xx && xx < W - 2 && 1 <= yy && yy < H - 2) {
          cell = yy * W + xx;
          ++n_in;
          const float dx = (float)(px - xx), dy = (float)(py - yy);
          iwe[(size_t)yy * W + xx] += (1.f - dx) * (1.f - dy);
          iwe[(size_t)yy * W + xx + 1] += dx * (1.f - dy);
          iwe[(size_t)(yy + 1) * W + xx] += (1.f - dx) * dy;
          iwe[(size_t)(yy + 1) * W + xx + 1] += dx * dy;
          if (want_grad) {
            const float r0[3] = {(float)J[0][0], (float)J[0][1], (float)J[0][2]};
            const float r1[3] = {(float)J[1][0], (float)J[1][1], (float)J[1][2]};
            float* d00 = &deriv[((size_t)yy * W + xx) * 3];
            float* d01 = &deriv[((size_t)yy * W + xx + 1) * 3];
            float* d10 = &deriv[((size_t)(yy + 1) * W + xx) * 3];
            float* d11 = &deriv[((size_t)(yy + 1) * W + xx + 1) * 3];
            for (int c = 0; c < 3; ++c) {  // :163-166
              d00[c] += r0[c] * (-(1.f - dy)) + r1[c] * (-(1.f - dx));
              d01[c] += r0[c] * (1.f - dy) + r1[c] * (-dx);
              d10[c] += r0[c] * (-dy) + r1[c] * (1.f - dx);
              d11[c] += r0[c] * dy + r1[c] * dx;
            }
          }
        }
      }
      if (out->cells) out->cells[i] = cell;
    }
  }
  out->n_inbounds = n_in;
  if (out->iwe_raw) std::memcpy(out->iwe_raw, iwe.data(), sizeof(float) * A);
  if (out->deriv_raw && want_grad) std::memcpy(out->deriv_raw, deriv.data(), sizeof(float) * A * 3);

  if (a->blur_sigma > 0) {  // :32-38
    std::vector<float> t((size_t)A);
    gaussianBlur(iwe.data(), t.data(), W, H, 1, a->blur_sigma);
    iwe.swap(t);
    if (want_grad) {
      std::vector<float> t3((size_t)A * 3);
      gaussianBlur(deriv.data(), t3.data(), W, H, 3, a->blur_sigma);
      deriv.swap(t3);
    }
  }
  if (out->iwe) std::memcpy(out->iwe, iwe.data(), sizeof(float) * A);
  if (out->deriv && want_grad) std::memcpy(out->deriv, deriv.data(), sizeof(float) * A * 3);

  std::vector<PlaneView> planes;
  if (want_grad)
    for (int c = 0; c < 3; ++c) planes.push_back(PlaneView{deriv.data() + c, 3});
  out->contrast = computeContrast(iwe.data(), A, planes, want_grad ? out->grad : nullptr, a->contrast_measure);
  return 0;
}

extern "C" int orc_fe_eval_batch(const orc_fe_args* a, const double* omegas, int k, int want_grad,
                                 double* contrasts, double* grads, int n_threads) {
  std::atomic<int> rc{0}, next{0};
  auto work = [&]() {
    for (int h = next.fetch_add(1); h < k; h = next.fetch_add(1)) {
      orc_fe_out o;
      std::memset(&o, 0, sizeof(o));
      const int r = orc_fe_eval(a, omegas + 3 * h, want_grad, &o);
      if (r != 0) rc = r;
      contrasts[h] = o.contrast;
      if (want_grad && grads) { grads[3 * h] = o.grad[0]; grads[3 * h + 1] = o.grad[1]; grads[3 * h + 2] = o.grad[2]; }
    }
  };
  const int nt = std::max(1, std::min(n_threads, k));
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return rc.load();
}

// ================================================================================================
// Back-end
// ================================================================================================
extern "C" double orc_update_alpha(const float* IGp, const float* IL, int64_t n) {
  // EventWarper::updateAlpha, event_pano_warper.cpp:134-165
  int64_t nz = 0;
  for (int64_t i = 0; i < n; ++i) nz += (IGp[i] != 0.f);
  if (nz < 1) return 0.0;
  auto density = [n](const float* img) {
    double area = 0, num = 0;
    for (int64_t i = 0; i < n; ++i) {
      const float e = std::exp(-1.0f * img[i]);  // cv::exp on (-(1.0/lambda0))*img, f32
      const float integrand = 1.f - e;
      area += (double)integrand;
      num += (double)img[i];
    }
    return num / area;
  };
  const double d_IGp = density(IGp);
  const double d_IL = density(IL);
  return d_IL / d_IGp;
}

extern "C" int orc_be_eval(const orc_be_args* a, const double* x, int want_grad, orc_be_out* out) {
  const int W = a->pano_width, H = a->pano_height;
  const int SW = a->sensor_width, SH = a->sensor_height;
  const int64_t A = (int64_t)W * H;
  const int64_t n = a->n_events;
  const int N = a->spline_order;
  if (N != 2 && N != 4) return fail("spline_order must be 2 or 4");
  const int K = a->n_knots;
  const int n_opt = K - a->n_fixed;
  if (n_opt < 0) return fail("n_fixed > n_knots");
  const int P = 3 * n_opt;
  const int Nk = N;  // NumInvolvedControlPoses(): 2 linear / 4 cubic, trajectory.h:100,158

  // copyAndUpdateTraj -> CopyAndIncrementalUpdate -> incrementalUpdate: K_i <- exp(x_i) * K_i
  // for the optimised knots (trajectory.cpp:221-263,491-522; ...analytical.cpp:22-33).
  std::vector<Quat> knots(K);
  for (int i = 0; i < K; ++i) knots[i] = Quat{a->knots_xyzw[4 * i], a->knots_xyzw[4 * i + 1], a->knots_xyzw[4 * i + 2], a->knots_xyzw[4 * i + 3]};
  for (int i = a->n_fixed; i < K; ++i) {
    const int j = i - a->n_fixed;
    const V3 d = x ? V3{x[3 * j], x[3 * j + 1], x[3 * j + 2]} : V3{0, 0, 0};
    knots[i] = qmul(so3exp(d), knots[i]);
  }

  // EquirectangularCamera(pano_size, 360, 180): equirectangular_camera.h:11-16,64-67
  const double cxp = (double)W / 2.0, cyp = (double)H / 2.0;
  const double fx = double((W / 360.0) * 180.0 / M_PI);
  const double fy = double((H / 180.0) * 180.0 / M_PI);
  const RosTime t_next{a->tnext_sec, a->tnext_nsec};

  // computeImageOfWarpedEvents, event_pano_warper.cpp:167-231
  std::vector<float> il_old((size_t)A, 0.f), il_new((size_t)A, 0.f);
  std::vector<float> bands(want_grad ? (size_t)A * P : 0, 0.f);
  if (out->cells) for (int64_t i = 0; i < n; ++i) out->cells[i] = -2;
  int64_t n_in = 0;
  const int bs = a->batch_size, sr = a->event_sample_rate;
  if (bs <= 0 || sr <= 0) return fail("batch_size and event_sample_rate must be > 0");

  for (int64_t beg = 0; beg < n - 1; beg += bs) {  // :188-196 (note: < end()-1)
    const int64_t left = n - beg;
    const int64_t end = (left > bs) ? beg + bs : n;
    // warpAndAccumulateEvents, :233-336
    const RosTime t_first = evTime(a->events[beg]);
    const RosTime t_last = evTime(a->events[end - 1]);
    const RosTime time_batch = add(t_first, mul(sub(t_last, t_first), 0.5));
    const int64_t t_ns = toNSec(time_batch);  // trajectory.cpp:89,332

    Quat q;
    int idx_cp_beg = 0;
    M3 Jd[4];
    if (splineEvaluate(N, knots.data(), K, a->t0_ns, a->dt_ns, t_ns, &q, &idx_cp_beg, want_grad ? Jd : nullptr) != 0)
      return fail("batch time outside the spline (BASALT_ASSERT would abort)");
    const M3 R = qmat(q);
    // trajectory.cpp:93-106 / :336-351: 3 x 3Nk f32, jacobian(j, i+3k) = d_val_d_knot[k](j,i)
    float Jk[3][12];
    if (want_grad)
      for (int k = 0; k < Nk; ++k)
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) Jk[r][3 * k + c] = (float)Jd[k](r, c);

    for (int64_t i = beg; i < end; i += sr) {  // :262
      const orc_event& e = a->events[i];
      if (e.x >= SW || e.y >= SH) return fail("event outside the sensor (vector::at would throw)");
      const double* bp = a->lut_xyz + 3 * ((int64_t)e.y * SW + e.x);
      const V3 b{bp[0], bp[1], bp[2]};
      // e_ray_w = R * e_ray_cam                                                     :269
      const double wx = R(0, 0) * b.x + R(0, 1) * b.y + R(0, 2) * b.z;
      const double wy = R(1, 0) * b.x + R(1, 1) * b.y + R(1, 2) * b.z;
      const double wz = R(2, 0) * b.x + R(2, 1) * b.y + R(2, 2) * b.z;
      // projectToImage, equirectangular_camera.h:18-45
      float dpm_drb[2][3];
      double px, py;
      equirectProjectO(wx, wy, wz, fx, fy, cxp, cyp, &px, &py, dpm_drb);

      float jac[2][12];
      if (want_grad) {
        // rb = R_cv * bvec (cv::Matx33d * Point3d: s = 0; s += ...)                :280
        const double rbx = 0 + R(0, 0) * b.x + R(0, 1) * b.y + R(0, 2) * b.z;
        const double rby = 0 + R(1, 0) * b.x + R(1, 1) * b.y + R(1, 2) * b.z;
        const double rbz = 0 + R(2, 0) * b.x + R(2, 1) * b.y + R(2, 2) * b.z;
        const float D[3][3] = {{0.f, (float)rbz, (float)(-rby)},                    // :281
                               {(float)(-rbz), 0.f, (float)rbx},
                               {(float)rby, (float)(-rbx), 0.f}};
        float dpm_ddrot[2][3];                                                       // :282 Matx23f*Matx33f
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3; ++c) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s += dpm_drb[r][k] * D[k][c];
            dpm_ddrot[r][c] = s;
          }
        // :285 cv::Mat product -> cv::gemm CV_32F: products accumulated in double, cast to f32
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3 * Nk; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)dpm_ddrot[r][k] * (double)Jk[k][c];
            jac[r][c] = (float)s;
          }
      }

      int32_t cell = -1;
      if (std::fabs(px) < 2e9 && std::fabs(py) < 2e9) {
        const int xx = (int)px, yy = (int)py;                                        // :290-293
        const float dx = (float)(px - xx), dy = (float)(py - yy);
        if (1 <= xx && xx < W - 2 && 1 <= yy && yy < H - 2) {                        // :296
          cell = yy * W + xx;
          ++n_in;
          float* il = lessThan(evTime(e), t_next) ? il_old.data() : il_new.data();   // :298
          il[(size_t)yy * W + xx] += (1.f - dx) * (1.f - dy);
          il[(size_t)yy * W + xx + 1] += dx * (1.f - dy);
          il[(size_t)(yy + 1) * W + xx] += (1.f - dx) * dy;
          il[(size_t)(yy + 1) * W + xx + 1] += dx * dy;
          if (want_grad) {
            for (int c = 0; c < 3 * Nk; ++c) {                                       // :316-332
              const float r0 = jac[0][c], r1 = jac[1][c];
              const int j = 3 * (idx_cp_beg - a->n_fixed) + c;
              if (j >= 0) {
                if (j >= P) return fail("band index out of range (vector::at would throw)");
                float* bd = &bands[(size_t)j * A];
                bd[(size_t)yy * W + xx] += r0 * (-(1.f - dy)) + r1 * (-(1.f - dx));
                bd[(size_t)yy * W + xx + 1] += r0 * (1.f - dy) + r1 * (-dx);
                bd[(size_t)(yy + 1) * W + xx] += r0 * (-dy) + r1 * (1.f - dx);
                bd[(size_t)(yy + 1) * W + xx + 1] += r0 * dy + r1 * dx;
              }
            }
          }
        }
      }
      if (out->cells) out->cells[i] = cell;
    }
  }
  out->n_inbounds = n_in;
  if (out->il_old) std::memcpy(out->il_old, il_old.data(), sizeof(float) * A);
  if (out->il_new) std::memcpy(out->il_new, il_new.data(), sizeof(float) * A);
  if (out->bands_raw && want_grad) std::memcpy(out->bands_raw, bands.data(), sizeof(float) * A * P);

  // IL = old + new (:199); I = IL + alpha*IGp (:213, cv::scaleAdd: f32, alpha cast to f32)
  std::vector<float> I((size_t)A);
  const float alpha_f = (float)a->alpha;
  for (int64_t i = 0; i < A; ++i) {
    const float il = il_old[i] + il_new[i];
    I[i] = a->IGp ? a->IGp[i] * alpha_f + il : il;
  }
  if (a->blur_sigma > 0) {  // :217-230
    std::vector<float> t((size_t)A);
    gaussianBlur(I.data(), t.data(), W, H, 1, a->blur_sigma);
    I.swap(t);
    if (want_grad)
      for (int p = 0; p < P; ++p) {
        gaussianBlur(&bands[(size_t)p * A], t.data(), W, H, 1, a->blur_sigma);
        std::memcpy(&bands[(size_t)p * A], t.data(), sizeof(float) * A);
      }
  }
  if (out->iwe) std::memcpy(out->iwe, I.data(), sizeof(float) * A);
  if (out->bands && want_grad) std::memcpy(out->bands, bands.data(), sizeof(float) * A * P);

  std::vector<PlaneView> planes;
  if (want_grad)
    for (int p = 0; p < P; ++p) planes.push_back(PlaneView{&bands[(size_t)p * A], 1});
  out->contrast = computeContrast(I.data(), A, planes, (want_grad ? out->grad : nullptr), a->contrast_measure);
  return 0;
}

// EventWarper::setUpdateTimesIG(rot, radius) (event_pano_warper.cpp:81-107) with warpEventToMap (:37-54)
extern "C" void orc_set_update_times(const double* lut_xyz, int SW, int SH, int PW, int PH, const double rot_xyzw[4],
                                     int radius, uint8_t* times) {
  std::vector<uint8_t> mask((size_t)PW * PH, 0);
  const M3 R = qmat(Quat{rot_xyzw[0], rot_xyzw[1], rot_xyzw[2], rot_xyzw[3]});
  const double cxp = (double)PW / 2.0, cyp = (double)PH / 2.0;
  const double fx = double((PW / 360.0) * 180.0 / M_PI), fy = double((PH / 180.0) * 180.0 / M_PI);
  for (int x = 0; x < SW; ++x)
    for (int y = 0; y < SH; ++y) {
      const double* b = lut_xyz + 3 * ((size_t)y * SW + x);
      const double wx = R(0, 0) * b[0] + R(0, 1) * b[1] + R(0, 2) * b[2];
      const double wy = R(1, 0) * b[0] + R(1, 1) * b[1] + R(1, 2) * b[2];
      const double wz = R(2, 0) * b[0] + R(2, 1) * b[1] + R(2, 2) * b[2];
      const double phi = std::atan2(wx, wz);
      const double theta = std::asin(wy / std::sqrt(wx * wx + wy * wy + wz * wz));
      const double px = cxp + phi * fx, py = cyp + theta * fy;
      const int ic = (int)px, ir = (int)py;                                          // :91
      for (int i = -radius; i <= radius; i++)
        for (int j = -radius; j <= radius; j++) {
          const int x_mask = ic + i, y_mask = ir + j;
          // :97 as written (`0 <= y_mask+j`); y_mask >= 0 added: Mat::at with a negative row is out of bounds
          if (0 <= y_mask + j && y_mask >= 0 && y_mask < PH && 0 <= x_mask && x_mask < PW) mask[(size_t)y_mask * PW + x_mask] = 1;
        }
    }
  for (size_t i = 0; i < mask.size(); ++i) {                                          // cv::add on CV_8U saturates (:106)
    const int v = (int)times[i] + (int)mask[i];
    times[i] = (uint8_t)(v > 255 ? 255 : v);
  }
}
// EventWarper::updateIG (event_pano_warper.cpp:109-126)
extern "C" void orc_update_ig(float* IG, const float* il_old, const uint8_t* times, int max_update_times, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    if ((int)times[i] <= max_update_times) IG[i] += il_old[i];
}

// ================================================================================================
// Building blocks for the pinning tests
// ================================================================================================
extern "C" int orc_gaussian_kernel(double sigma, float* taps) {
  std::vector<float> t;
  const int k = gaussianKernel(sigma, t);
  for (int i = 0; i < k && i < 64; ++i) taps[i] = t[i];
  return k;
}
extern "C" void orc_gaussian_blur(const float* src, float* dst, int W, int H, int C, double sigma) {
  gaussianBlur(src, dst, W, H, C, sigma);
}
extern "C" void orc_mean_stddev(const float* img, int64_t n, double* mean, double* stddev) {
  meanStdDev(img, n, mean, stddev);
}
extern "C" int orc_so3_spline_eval(int order, const double* knots_xyzw, int K, int64_t t0_ns, int64_t dt_ns,
                                   int64_t t_ns, double* q_xyzw, double* R, int32_t* start_idx, double* J) {
  if (order < 2 || order > 4) return fail("order must be 2..4");
  std::vector<Quat> knots(K);
  for (int i = 0; i < K; ++i) knots[i] = Quat{knots_xyzw[4 * i], knots_xyzw[4 * i + 1], knots_xyzw[4 * i + 2], knots_xyzw[4 * i + 3]};
  Quat q;
  int idx = 0;
  M3 Jd[4];
  if (splineEvaluate(order, knots.data(), K, t0_ns, dt_ns, t_ns, &q, &idx, J ? Jd : nullptr) != 0)
    return fail("time outside the spline");
  if (q_xyzw) { q_xyzw[0] = q.x; q_xyzw[1] = q.y; q_xyzw[2] = q.z; q_xyzw[3] = q.w; }
  if (R) { const M3 m = qmat(q); std::memcpy(R, m.m, sizeof(double) * 9); }
  if (start_idx) *start_idx = idx;
  if (J) for (int k = 0; k < order; ++k) std::memcpy(J + 9 * k, Jd[k].m, sizeof(double) * 9);
  return 0;
}
extern "C" void orc_so3_exp(const double w[3], double q[4]) {
  const Quat r = so3exp(V3{w[0], w[1], w[2]});
  q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
}
extern "C" void orc_so3_log(const double q[4], double w[3]) {
  const V3 r = so3log(Quat{q[0], q[1], q[2], q[3]});
  w[0] = r.x; w[1] = r.y; w[2] = r.z;
}
extern "C" void orc_batch_mid_time(uint32_t s0, uint32_t ns0, uint32_t s1, uint32_t ns1, uint32_t* s, uint32_t* ns) {
  const RosTime m = batchMid(RosTime{s0, ns0}, RosTime{s1, ns1});
  *s = m.sec; *ns = m.nsec;
}
extern "C" const char* orc_last_error(void) { return g_err.c_str(); }
