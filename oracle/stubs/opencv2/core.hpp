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
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;   // OpenCV defines uchar at global scope (cvdef.h)

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
  explicit Point3_(T v0) : x(v0), y(0), z(0) {}   // OpenCV reaches this through Vec<T,3>(v0): first element set, rest zero
  Point3_ cross(const Point3_& p) const { return Point3_(y * p.z - z * p.y, z * p.x - x * p.z, x * p.y - y * p.x); }
};
typedef Point3_<double> Point3d;
typedef Point3_<float> Point3f;
template <class T> inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> inline Point3_<T> operator*(double s, const Point3_<T>& a) { return Point3_<T>((T)(s * a.x), (T)(s * a.y), (T)(s * a.z)); }
template <class T> inline Point3_<T> operator*(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x * s), (T)(a.y * s), (T)(a.z * s)); }
// Point3f * float: OpenCV's operator*(Point3_<T>, float) computes saturate_cast<T>(a.x * b) -- a float product for T = float
inline Point3_<float> operator*(const Point3_<float>& a, float s) { return Point3_<float>(a.x * s, a.y * s, a.z * s); }
template <class T> inline Point3_<T>& operator+=(Point3_<T>& a, const Point3_<T>& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }

template <class T, int M, int N> struct Matx {
  T val[M * N];
  Matx() { for (int i = 0; i < M * N; ++i) val[i] = T(0); }
  template <class... A> Matx(A... a) : val{T(a)...} { static_assert(sizeof...(A) == M * N, "Matx initialiser count"); }
  T& operator()(int i, int j) { return val[i * N + j]; }
  const T& operator()(int i, int j) const { return val[i * N + j]; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  Matx<T, 1, N> row(int i) const { Matx<T, 1, N> r; for (int j = 0; j < N; ++j) r.val[j] = val[i * N + j]; return r; }
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
typedef Matx<double, 1, 3> Matx13d;
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<float, 2, 3> Matx23f;
typedef Matx<float, 3, 3> Matx33f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

// ---- Vec / Scalar -------------------------------------------------------------------------------------------------
template <class T, int N> struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
  template <class... A> Vec(A... a) : val{T(a)...} {}
  template <class U> Vec(const Vec<U, N>& o) { for (int i = 0; i < N; ++i) val[i] = (T)o.val[i]; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<int, 3> Vec3i;
typedef Vec<double, 4> Vec4d;
typedef Vec<double, 4> Scalar;
typedef Point_<int> Point2i;
typedef Point_<int> Point;

// Matx33 * Point3 (OpenCV: Matx product with the point as a 3x1 Matx: s = 0; s += a(i,k) * b(k))
template <class T> inline Point3_<T> operator*(const Matx<T, 3, 3>& a, const Point3_<T>& b) {
  const T v[3] = {b.x, b.y, b.z};
  T r[3];
  for (int i = 0; i < 3; ++i) { T s = 0; for (int k = 0; k < 3; ++k) s += a(i, k) * v[k]; r[i] = s; }
  return Point3_<T>(r[0], r[1], r[2]);
}

// ---- cv::Mat: a reference-counted dense single-channel matrix (copies are shallow, as in OpenCV) ---------------------
// The element-wise helpers below (add, scaleAdd, exp, sum, countNonZero, gemm, GaussianBlur) carry the ORACLE's restatement
// of the OpenCV arithmetic (f32 element ops; sums and small gemm accumulated in f64; blur = the cv2-pinned restatement).
// What the shims built on this header pin is the reference's own control flow, indexing and expression structure around them.
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_8UC3 16
#define CV_32FC3 21
struct Mat {
  int rows = 0, cols = 0, type_ = CV_32FC1;
  std::shared_ptr<std::vector<unsigned char>> buf;
  static int esz(int t) { return t == CV_8UC1 ? 1 : t == CV_32FC1 ? 4 : t == CV_64FC1 ? 8 : t == CV_32FC3 ? 12 : 3; }
  int channels() const { return (type_ == CV_32FC3 || type_ == CV_8UC3) ? 3 : 1; }
  Mat() {}
  Mat(int r, int c, int t) { create(r, c, t); }
  Mat(Size sz, int t);
  void create(int r, int c, int t) { rows = r; cols = c; type_ = t; buf = std::make_shared<std::vector<unsigned char>>((size_t)r * c * esz(t), (unsigned char)0); }
  static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
  static Mat zeros(Size sz, int t);
  bool empty() const { return !buf || rows * cols == 0; }
  int type() const { return type_; }
  size_t total() const { return (size_t)rows * cols; }
  template <class T> T* ptr() { return reinterpret_cast<T*>(buf->data()); }
  template <class T> const T* ptr() const { return reinterpret_cast<const T*>(buf->data()); }
  template <class T> T& at(int r, int c) { return ptr<T>()[(size_t)r * cols + c]; }
  template <class T> const T& at(int r, int c) const { return ptr<T>()[(size_t)r * cols + c]; }
  template <class T, class P> T& at(const Point_<P>& p) { return at<T>((int)p.y, (int)p.x); }
  Mat& setTo(double v) {
    if (type_ == CV_32FC1) for (size_t i = 0; i < total(); ++i) ptr<float>()[i] = (float)v;
    else if (type_ == CV_64FC1) for (size_t i = 0; i < total(); ++i) ptr<double>()[i] = v;
    else std::memset(buf->data(), (int)v, buf->size());
    return *this;
  }
  void copyTo(Mat& dst) const { dst.create(rows, cols, type_); std::memcpy(dst.buf->data(), buf->data(), buf->size()); }
  Mat clone() const { Mat m; copyTo(m); return m; }
  void convertTo(Mat&, int) const { std::abort(); }   // display only, never executed
  Mat mul(const Mat& o) const { Mat r(rows, cols, CV_32FC1); for (size_t i = 0; i < total(); ++i) r.ptr<float>()[i] = ptr<float>()[i] * o.ptr<float>()[i]; return r; }
  inline Mat mul(const struct MatExpr& e) const;
};

// The two lazy expressions the focus functions build: (A - s) and k * (A - s).  OpenCV folds them into ONE pass
// dst = A * (float)alpha + (float)beta in f32 (MatOp_AddEx -> convertTo / add with a scalar); evaluated on conversion to Mat.
struct MatExpr {
  Mat a; double alpha = 1.0, beta = 0.0;
  operator Mat() const {
    Mat r(a.rows, a.cols, CV_32FC1);
    const float af = (float)alpha, bf = (float)beta;
    for (size_t i = 0; i < a.total(); ++i) r.ptr<float>()[i] = a.ptr<float>()[i] * af + bf;
    return r;
  }
};
inline MatExpr operator-(const Mat& a, double s) { MatExpr e; e.a = a; e.alpha = 1.0; e.beta = -s; return e; }
inline MatExpr operator*(double k, const MatExpr& e) { MatExpr r = e; r.alpha *= k; r.beta *= k; return r; }
inline Mat Mat::mul(const MatExpr& e) const { return mul((Mat)e); }

inline Mat::Mat(Size sz, int t) { create(sz.height, sz.width, t); }
inline Mat Mat::zeros(Size sz, int t) { return Mat(sz.height, sz.width, t); }

}  // namespace cv
