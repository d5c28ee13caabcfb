// TEST INFRASTRUCTURE -- the OpenCV free functions / operators the reference's warp code calls, for the cv::Mat stand-in of
// core.hpp.  Arithmetic = the oracle's restatement of OpenCV (see the note in core.hpp); GaussianBlur IS the oracle's
// cv2-pinned blur (orc_gaussian_blur).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "core.hpp"

extern "C" void orc_gaussian_blur(const float* src, float* dst, int W, int H, int C, double sigma);
extern "C" void orc_mean_stddev(const float* img, int64_t n, double* mean, double* stddev);

namespace cv {

inline void add(const Mat& a, const Mat& b, Mat& dst) {
  Mat out(a.rows, a.cols, a.type());
  if (a.type() == CV_32FC1) for (size_t i = 0; i < a.total(); ++i) out.ptr<float>()[i] = a.ptr<float>()[i] + b.ptr<float>()[i];
  else for (size_t i = 0; i < a.total(); ++i) { const int v = (int)a.ptr<uchar>()[i] + (int)b.ptr<uchar>()[i]; out.ptr<uchar>()[i] = (uchar)(v > 255 ? 255 : v); }   // saturate_cast
  dst = out;
}
// dst = src1 * alpha + src2, f32, alpha cast to f32
inline void scaleAdd(const Mat& src1, double alpha, const Mat& src2, Mat& dst) {
  Mat out(src1.rows, src1.cols, CV_32FC1);
  const float af = (float)alpha;
  for (size_t i = 0; i < src1.total(); ++i) out.ptr<float>()[i] = src1.ptr<float>()[i] * af + src2.ptr<float>()[i];
  dst = out;
}
inline void exp(const Mat& src, Mat& dst) {
  Mat out(src.rows, src.cols, CV_32FC1);
  for (size_t i = 0; i < src.total(); ++i) out.ptr<float>()[i] = std::exp(src.ptr<float>()[i]);
  dst = out;
}
inline Scalar sum(const Mat& a) {
  double s = 0;
  for (size_t i = 0; i < a.total(); ++i) s += (double)a.ptr<float>()[i];
  return Scalar(s, 0.0, 0.0, 0.0);
}
inline int countNonZero(const Mat& a) {
  int n = 0;
  for (size_t i = 0; i < a.total(); ++i) n += a.ptr<float>()[i] != 0.f;
  return n;
}
inline Mat operator*(double s, const Mat& a) {
  Mat out(a.rows, a.cols, CV_32FC1);
  const float sf = (float)s;
  for (size_t i = 0; i < a.total(); ++i) out.ptr<float>()[i] = sf * a.ptr<float>()[i];
  return out;
}
inline Mat operator-(const Mat& a) {
  Mat out(a.rows, a.cols, CV_32FC1);
  for (size_t i = 0; i < a.total(); ++i) out.ptr<float>()[i] = -a.ptr<float>()[i];
  return out;
}
inline Mat operator-(float s, const Mat& a) {
  Mat out(a.rows, a.cols, CV_32FC1);
  for (size_t i = 0; i < a.total(); ++i) out.ptr<float>()[i] = s - a.ptr<float>()[i];
  return out;
}
// Matx (f32) * Mat (f32): cv::gemm on small CV_32F matrices accumulates the products in f64 and casts the sum to f32
template <int M, int K> inline Mat operator*(const Matx<float, M, K>& a, const Mat& b) {
  Mat out(M, b.cols, CV_32FC1);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < b.cols; ++j) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)a(i, k) * (double)b.at<float>(k, j);
      out.at<float>(i, j) = (float)s;
    }
  return out;
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size /*ksize = (0,0)*/, double sigma) {
  Mat out(src.rows, src.cols, src.type());
  orc_gaussian_blur(src.ptr<float>(), out.ptr<float>(), src.cols, src.rows, src.channels(), sigma);
  dst = out;
}

// ---- reductions used by the focus functions (f64 accumulation, as cv::mean / meanStdDev / norm do for CV_32F) ---------
enum { NORM_L2SQR = 5 };
inline Scalar mean(const Mat& a) {
  double s = 0;
  for (size_t i = 0; i < a.total(); ++i) s += (double)a.ptr<float>()[i];
  return Scalar(s / (double)a.total(), 0.0, 0.0, 0.0);
}
inline void meanStdDev(const Mat& a, Vec4d& mean_out, Vec4d& stddev_out) {
  double m = 0, sd = 0;
  orc_mean_stddev(a.ptr<float>(), (int64_t)a.total(), &m, &sd);
  mean_out = Vec4d(m, 0.0, 0.0, 0.0);
  stddev_out = Vec4d(sd, 0.0, 0.0, 0.0);
}
inline double norm(const Mat& a, int /*NORM_L2SQR*/) {
  double s = 0;
  for (size_t i = 0; i < a.total(); ++i) { const double v = a.ptr<float>()[i]; s += v * v; }
  return s;
}
inline void split(const Mat& src, std::vector<Mat>& channels) {
  const int cn = src.channels();
  channels.clear();
  for (int c = 0; c < cn; ++c) {
    Mat m(src.rows, src.cols, CV_32FC1);
    for (size_t i = 0; i < m.total(); ++i) m.ptr<float>()[i] = src.ptr<float>()[i * cn + c];
    channels.push_back(m);
  }
}
inline Mat operator+(const Mat& a, const Mat& b) { Mat r; add(a, b, r); return r; }
// display-only helpers of publishEventImage (never executed: the stand-in publishers have no subscribers): compile only
enum { NORM_MINMAX = 32, COLOR_GRAY2BGR = 8 };
#define CV_GRAY2BGR 8
inline void cv_stub_unreachable(const char* what) { std::fprintf(stderr, "cv::%s is not part of the stand-in\n", what); std::abort(); }
inline void hconcat(const Mat&, const Mat&, Mat&) { cv_stub_unreachable("hconcat"); }
inline void normalize(const Mat&, Mat&, double, double, int, int) { cv_stub_unreachable("normalize"); }
inline void pow(const Mat&, double, Mat&) { cv_stub_unreachable("pow"); }
inline void cvtColor(const Mat&, Mat&, int) { cv_stub_unreachable("cvtColor"); }
// IMAGE_GRADIENT_MAGNITUDE_CONTRAST is unreachable from the reference's launch files and is not restated: compile only
inline void Sobel(const Mat&, Mat&, int, int, int) { std::fprintf(stderr, "cv::Sobel is not part of the stand-in\n"); std::abort(); }

}  // namespace cv
