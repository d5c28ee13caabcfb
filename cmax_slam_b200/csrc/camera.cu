// camera.cu -- bearing-vector look-up table from the camera calibration ("LUT precompute"):
// one thread per sensor pixel.
//
// Computes what CMaxSLAM::precomputeBearingVectors does (src/cmax_slam.cpp:106-120):
//     rectified = cam.rectifyPoint(cv::Point2d(x, y));  bearing = cam.projectPixelTo3dRay(rectified)
// with image_geometry::PinholeCameraModel (ROS, un-vendored) on top of cv::undistortPoints (OpenCV, un-vendored):
//   rectifyPoint        : no distortion -> the raw pixel; otherwise the pixel goes through cv::undistortPoints as a
//                         CV_32FC2 point (float in, float out) with K, D, R, P
//   undistortPoints     : x = (u-cx)/fx, y = (v-cy)/fy; 5 fixed-point iterations of the inverse plumb-bob / rational
//                         model; then [x y 1] -> (P[:3,:3] R) [x y 1], perspective divide; result rounded to float
//   projectPixelTo3dRay : ((u - P02 - P03) / P00, (v - P12 - P13) / P11, 1)
// Pinned by us against cv2 4.13 (tests/golden/lut_cv2.npz); PARITY UNPINNED against the reference's own OpenCV / ROS.
#include "capi_common.cuh"

using namespace cmaxb;

namespace {

struct CamDev {
  int W, H, distorted;
  double fx, fy, cx, cy, ifx, ify;
  double k[12];
  double RR[9];         // P[:3,:3] * R
  double pfx, pfy, pcx, pcy, ptx, pty;
};

__global__ void lut_precompute_kernel(CamDev c, double* __restrict__ lut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.W * c.H) return;
  const int px = i % c.W, py = i / c.W;
  double ru = (double)px, rv = (double)py;                 // rectified pixel
  if (c.distorted) {
    const double u = (double)(float)px, v = (double)(float)py;   // cv::Point2f raw32 = uv_raw
    double x = (u - c.cx) * c.ifx, y = (v - c.cy) * c.ify;
    const double x0 = x, y0 = y;
    const double* k = c.k;
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) { x = (u - c.cx) * c.ifx; y = (v - c.cy) * c.ify; break; }
      const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
      const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
      x = (x0 - deltaX) * icdist;
      y = (y0 - deltaY) * icdist;
    }
    const double xx = c.RR[0] * x + c.RR[1] * y + c.RR[2];
    const double yy = c.RR[3] * x + c.RR[4] * y + c.RR[5];
    const double ww = 1. / (c.RR[6] * x + c.RR[7] * y + c.RR[8]);
    ru = (double)(float)(xx * ww);                         // rect32 is a cv::Point2f
    rv = (double)(float)(yy * ww);
  }
  lut[3 * (long long)i] = (ru - c.pcx - c.ptx) / c.pfx;
  lut[3 * (long long)i + 1] = (rv - c.pcy - c.pty) / c.pfy;
  lut[3 * (long long)i + 2] = 1.0;
}

}  // namespace

extern "C" int cmaxb_precompute_bearing_vectors(const cmaxb_camera_info* info, int device, double* lut_xyz) {
  if (!info || !lut_xyz) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (info->width < 1 || info->height < 1 || info->n_D < 0 || info->n_D > 12) return set_error(CMAXB_ERR_INVALID, "bad camera info");
  if (info->K[0] == 0.0 || info->K[4] == 0.0 || info->P[0] == 0.0 || info->P[5] == 0.0) return set_error(CMAXB_ERR_INVALID, "uncalibrated camera (zero focal length)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return set_error(CMAXB_ERR_CUDA, "no CUDA device: libcmax_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(CMAXB_ERR_INVALID, "bad device ordinal");
  CMAXB_CUDA_TRY(cudaSetDevice(device));
  CamDev c{};
  c.W = info->width; c.H = info->height;
  c.fx = info->K[0]; c.fy = info->K[4]; c.cx = info->K[2]; c.cy = info->K[5];
  c.ifx = 1. / c.fx; c.ify = 1. / c.fy;
  c.distorted = 0;
  for (int i = 0; i < 12; ++i) c.k[i] = 0.0;
  for (int i = 0; i < info->n_D; ++i) { c.k[i] = info->D[i]; if (info->D[i] != 0.0) c.distorted = 1; }
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q) {
      double s = 0;
      for (int m = 0; m < 3; ++m) s += info->P[4 * r + m] * info->R[3 * m + q];
      c.RR[3 * r + q] = s;
    }
  c.pfx = info->P[0]; c.pfy = info->P[5]; c.pcx = info->P[2]; c.pcy = info->P[6]; c.ptx = info->P[3]; c.pty = info->P[7];
  const long long A = (long long)c.W * c.H;
  double* d = nullptr;
  CMAXB_TRY(dev_alloc(&d, (size_t)(3 * A)));
  lut_precompute_kernel<<<(unsigned)((A + 255) / 256), 256>>>(c, d);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) { g_launch_count.fetch_add(1, std::memory_order_relaxed); e = cudaMemcpy(lut_xyz, d, sizeof(double) * 3 * A, cudaMemcpyDeviceToHost); }
  cudaFree(d);
  if (e != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("bearing-vector kernel: ") + cudaGetErrorString(e));
  return CMAXB_OK;
}
