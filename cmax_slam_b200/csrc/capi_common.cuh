// capi_common.cuh -- error plumbing, launch accounting and per-kernel CUDA-event timing shared by
// the C-ABI translation units.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace cmaxb {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launch_count;

inline int set_error(int code, const std::string& msg) { g_last_error = msg; return code; }

#define CMAXB_CUDA_TRY(expr)                                                                   \
  do {                                                                                         \
    cudaError_t err__ = (expr);                                                                \
    if (err__ != cudaSuccess) {                                                                \
      return ::cmaxb::set_error(CMAXB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    }                                                                                          \
  } while (0)

#define CMAXB_TRY(expr)          \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != CMAXB_OK) return rc__; \
  } while (0)

// Per-kernel timing: when enabled every launch is bracketed by two events on the launching stream
// and waited for (serialising the stream -- a measurement mode, not the production mode).
struct KernelProfiler {
  bool enabled = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double ms[CMAXB_K_COUNT] = {};
  uint64_t launches[CMAXB_K_COUNT] = {};

  int init() {
    CMAXB_CUDA_TRY(cudaEventCreate(&e0));
    CMAXB_CUDA_TRY(cudaEventCreate(&e1));
    return CMAXB_OK;
  }
  void destroy() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    e0 = e1 = nullptr;
  }
  void reset() {
    for (int i = 0; i < CMAXB_K_COUNT; ++i) { ms[i] = 0; launches[i] = 0; }
  }
  template <class F>
  int run(int kind, cudaStream_t s, bool is_kernel, F&& f) {
    if (enabled) CMAXB_CUDA_TRY(cudaEventRecord(e0, s));
    (void)cudaGetLastError();   // do not inherit a stale error of another library in this process
    f();
    {
      cudaError_t err = cudaGetLastError();
      if (err != cudaSuccess)
        return set_error(CMAXB_ERR_CUDA, std::string("launch of kernel kind ") + std::to_string(kind) + " failed: " + cudaGetErrorString(err));
    }
    if (is_kernel) g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (enabled) {
      CMAXB_CUDA_TRY(cudaEventRecord(e1, s));
      CMAXB_CUDA_TRY(cudaEventSynchronize(e1));
      float t = 0.f;
      CMAXB_CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
      ms[kind] += t;
      launches[kind] += 1;
    }
    return CMAXB_OK;
  }
};

// cv::GaussianBlur(.., Size(0,0), sigma) for CV_32F: ksize = cvRound(sigma*8+1)|1,
// getGaussianKernel(ksize, sigma, CV_32F): exp(-x^2/(2 sigma^2)) normalised in f64, cast to f32.
inline int make_taps(double sigma, Taps* t) {
  std::memset(t, 0, sizeof(*t));
  if (!(sigma > 0)) { t->r = 0; t->w[0] = 1.0f; return CMAXB_OK; }
  const int ksize = (int)std::lrint(sigma * 4 * 2 + 1) | 1;
  const int r = ksize / 2;
  if (r > kMaxRadius) return set_error(CMAXB_ERR_INVALID, "blur_sigma too large (kernel radius > 16)");
  double k[kMaxTaps];
  const double scale2X = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < ksize; ++i) {
    const double x = i - (ksize - 1) * 0.5;
    k[i] = std::exp(scale2X * x * x);
    sum += k[i];
  }
  sum = 1.0 / sum;
  for (int i = 0; i < ksize; ++i) t->w[i] = (float)(k[i] * sum);
  t->r = r;
  return CMAXB_OK;
}

template <class T>
inline int dev_alloc(T** p, size_t count) {
  CMAXB_CUDA_TRY(cudaMalloc((void**)p, count * sizeof(T)));
  return CMAXB_OK;
}

}  // namespace cmaxb
