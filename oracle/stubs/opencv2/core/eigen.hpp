// TEST INFRASTRUCTURE stub: cv::eigen2cv for fixed-size matrices (element copy)
#pragma once
#include <Eigen/Core>

#include "../core.hpp"
#include "../ops_stub.hpp"
namespace cv {
template <class T, int R, int C> inline void eigen2cv(const Eigen::Matrix<T, R, C>& src, Matx<T, R, C>& dst) {
  for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) dst(i, j) = src(i, j);
}
}  // namespace cv
