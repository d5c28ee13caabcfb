/*
 * ref_focus_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" wrappers (our code) around the REAL focus functions of the reference,
 * compiled from where they lie (never copied):
 *   src/frontend/local_focus_funcs.cpp   computeContrast / contrast_Variance / contrast_MeanSquare (cv::Matx13d gradient)     (A4)
 *   src/backend/global_focus_funcs.cpp   computeContrast / contrast_Variance / contrast_MeanSquare (1 x P cv::Mat gradient)   (A9)
 * OpenCV is absent from this image: oracle/stubs/opencv2 supplies a cv::Mat stand-in whose reductions and lazy expressions carry
 * the oracle's restatement of the OpenCV arithmetic (pinned against cv2 4.13 separately, tests/test_oracle_opencv.py); what this
 * pins is the reference's own formulae and operation structure.  Built into oracle/_ref/libref_focus.so.
 */
#include <cstring>
#include <vector>

#include "frontend/local_focus_funcs.h"
namespace fe_focus = cmax_slam;
#include <opencv2/imgproc.hpp>

/* the two headers declare different overloads of cmax_slam::computeContrast and clashing enums; the back-end one is declared by hand */
namespace cmax_slam {
double computeContrast(const cv::Mat& image, std::vector<cv::Mat>* image_deriv, cv::Mat* gradient, const int contrast_measure);
}

static cv::Mat wrap(const float* p, int W, int H, int type) {
  cv::Mat m(H, W, type);
  std::memcpy(m.ptr<float>(), p, sizeof(float) * (size_t)W * H * m.channels());
  return m;
}

extern "C" double ref1p_fe_contrast(const float* iwe, const float* deriv3, int W, int H, int measure, double* grad3) {
  cv::Mat img = wrap(iwe, W, H, CV_32FC1);
  if (!grad3) return cmax_slam::computeContrast(img, (cv::Mat*)nullptr, (cv::Matx13d*)nullptr, measure);
  cv::Mat der = wrap(deriv3, W, H, CV_32FC3);
  cv::Matx13d g;
  const double c = cmax_slam::computeContrast(img, &der, &g, measure);
  for (int i = 0; i < 3; ++i) grad3[i] = g(i);
  return c;
}

extern "C" double ref1p_be_contrast(const float* iwe, const float* bands, int P, int W, int H, int measure, double* grad) {
  cv::Mat img = wrap(iwe, W, H, CV_32FC1);
  std::vector<cv::Mat> ch;
  for (int p = 0; p < P; ++p) ch.push_back(wrap(bands + (size_t)p * W * H, W, H, CV_32FC1));
  cv::Mat g;
  const double c = cmax_slam::computeContrast(img, &ch, grad ? &g : nullptr, measure);
  if (grad) for (int p = 0; p < P; ++p) grad[p] = g.at<double>(0, p);
  return c;
}
