// TEST INFRASTRUCTURE -- container-only stand-in for the few OpenCV core types the reference's geometry utilities use
// (cv::Point_, cv::Point3_, cv::Matx, cv::Size, CV_PI), so that the REAL reference sources
//   src/utils/image_geom_util.cpp, include/utils/image_geom_util.h, include/backend/equirectangular_camera.h
// compile here (OpenCV's C++ headers are not installed) and pin the oracle's restatement of them (oracle/ref_geom_shim.cpp).
// Every arithmetic expression that the pinning tests check is the reference's own; the only arithmetic defined here is the
// element-wise Point algebra and the textbook Matx product (s = 0; s += a(i,k) * b(k,j)), neither of which is on the hot path
// of the functions that are pinned (canonicalProjection, applyIntrinsics, cross2Matrix, projectToImage).
#pragma once
#include <cmath>
#include <initializer_list>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795

namespace cv {

template <class T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<double> Point2d;
typedef Point_<float> Point2f;

template <class T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
  Point3_ cross(const Point3_& p) const { return Point3_(y * p.z - z * p.y, z * p.x - x * p.z, x * p.y - y * p.x); }
};
typedef Point3_<double> Point3d;
template <class T> inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> inline Point3_<T> operator*(double s, const Point3_<T>& a) { return Point3_<T>((T)(s * a.x), (T)(s * a.y), (T)(s * a.z)); }
template <class T> inline Point3_<T> operator*(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x * s), (T)(a.y * s), (T)(a.z * s)); }

template <class T, int M, int N> struct Matx {
  T val[M * N];
  Matx() { for (int i = 0; i < M * N; ++i) val[i] = T(0); }
  template <class... A> Matx(A... a) : val{T(a)...} { static_assert(sizeof...(A) == M * N, "Matx initialiser count"); }
  T& operator()(int i, int j) { return val[i * N + j]; }
  const T& operator()(int i, int j) const { return val[i * N + j]; }
};
template <class T, int M, int K, int N> inline Matx<T, M, N> operator*(const Matx<T, M, K>& a, const Matx<T, K, N>& b) {
  Matx<T, M, N> r;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      T s = 0;
      for (int k = 0; k < K; ++k) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}
typedef Matx<double, 2, 2> Matx22d;
typedef Matx<double, 2, 3> Matx23d;
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<float, 2, 3> Matx23f;

// cv::Mat as the trajectory code uses it: a small zero-initialised CV_32FC1 matrix addressed with at<float>(r, c)
#define CV_32FC1 5
struct Mat {
  int rows = 0, cols = 0;
  std::vector<float> data;
  static Mat zeros(int r, int c, int /*type*/) { Mat m; m.rows = r; m.cols = c; m.data.assign((size_t)r * c, 0.f); return m; }
  template <class T> T& at(int r, int c) { return reinterpret_cast<T&>(data[(size_t)r * cols + c]); }
  template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T&>(data[(size_t)r * cols + c]); }
};

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

}  // namespace cv
