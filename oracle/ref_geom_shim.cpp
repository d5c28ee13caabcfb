/*
 * ref_geom_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" wrappers (our code) around the REAL first-party geometry of the
 * reference, compiled from where it lies under /root/reference (never copied):
 *   canonicalProjection, applyIntrinsics        src/utils/image_geom_util.cpp:7-41      
 *   cross2Matrix                                include/utils/image_geom_util.h:5-8     (A3)
 *   dvs::EquirectangularCamera::projectToImage  include/backend/equirectangular_camera.h:11-45,64-67  (A8)
 * OpenCV's C++ headers and glog are absent from this image: oracle/stubs/ supplies container-only stand-ins for
 * cv::Point_/Point3_/Matx/Size (see oracle/stubs/opencv2/core.hpp).  Built by oracle/Makefile into
 * oracle/_ref/libref_geom.so; pins the restated geometry of cmax_oracle.cpp (tests/test_oracle_geom.py) and generates
 * tests/golden/geom_ref.npz.
 */
#include "utils/image_geom_util.h"
#include "backend/equirectangular_camera.h"

/* pinhole: p (3) -> calibrated uv (2), pixel (2), Jproj (2x3 row-major), Jintr (2x2) */
extern "C" void ref_pinhole(const double p[3], const double K4[4], double uv[2], double px[2], double Jproj[6], double Jintr[4]) {
  cv::Point2d c, out;
  cv::Matx23d jp;
  cv::Matx22d ji;
  canonicalProjection(cv::Point3d(p[0], p[1], p[2]), &c, &jp);
  const cv::Matx33d K(K4[0], 0., K4[2], 0., K4[1], K4[3], 0., 0., 1.);
  applyIntrinsics(c, K, &out, &ji);
  uv[0] = c.x; uv[1] = c.y; px[0] = out.x; px[1] = out.y;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) Jproj[i * 3 + j] = jp(i, j);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) Jintr[i * 2 + j] = ji(i, j);
}
extern "C" void ref_cross2matrix(const double v[3], double M[9]) {
  cv::Matx33d m;
  cross2Matrix(cv::Point3d(v[0], v[1], v[2]), &m);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i * 3 + j] = m(i, j);
}
/* equirectangular panorama PW x PH (360 x 180 degrees, as EventWarper constructs it): world ray -> pixel, 2x3 f32 Jacobian */
extern "C" void ref_equirect(const double w[3], int PW, int PH, double px[2], float J[6]) {
  dvs::EquirectangularCamera cam(cv::Size(PW, PH), 360.0, 180.0);
  cv::Matx23f j;
  const Eigen::Vector2d p = cam.projectToImage(Eigen::Vector3d(w[0], w[1], w[2]), &j);
  px[0] = p[0]; px[1] = p[1];
  for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) J[a * 3 + b] = j(a, b);
}
