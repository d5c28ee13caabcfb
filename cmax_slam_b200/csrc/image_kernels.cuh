// image_kernels.cuh -- fused Gaussian blur + contrast/gradient reduction, and the adjoint
// (transposed) blur used by the gather formulation of the gradient.
//
// Replaces, per cost evaluation, the reference's full-image passes
//   cv::GaussianBlur(iwe) / cv::GaussianBlur(deriv)      local_image_warped_events.cpp:32-38,
//                                                         event_pano_warper.cpp:217-230
//   cv::split, cv::meanStdDev, cv::mean, Mat::mul ...     local_focus_funcs.cpp:9-44,82-120,
//                                                         global_focus_funcs.cpp:11-47
// with ONE pass: each CTA stages a (TH+2r)x(TW+2r) tile (BORDER_REFLECT_101 resolved at load
// time) in shared memory, runs the separable filter with OpenCV's operation order (row: left to
// right FMA chain; column: symmetric pairs, FMA chain -- bit-exact vs cv2 4.13 for equal input),
// and reduces sum(I), sum(I^2), sum(D_c), sum(I*D_c) in f64.  The last CTA to finish (atomic
// ticket) turns the sums into contrast / gradient and re-arms the accumulators, so no extra
// launch and no host round trip is needed.
#pragma once
#include "common.cuh"

namespace cmaxb {

constexpr int kTW = 32, kTH = 32, kImgThreads = 256;
constexpr int kNAcc = 8;  // S1, S2, SD[3], SID[3]

__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * (len - 1) - p;
  return p;
}

// ---- pixel sources ------------------------------------------------------------------------------
struct SrcPlane {            // one float plane per hypothesis
  const float* base; long long stride_h;
  __device__ __forceinline__ float load(int h, int x, int y, int W) const {
    return base[h * stride_h + (long long)y * W + x];
  }
};
struct SrcPlane4 {           // interleaved (I, D0, D1, D2)
  const float4* base; long long stride_h;
  __device__ __forceinline__ float4 load(int h, int x, int y, int W) const {
    return base[h * stride_h + (long long)y * W + x];
  }
};
struct SrcBeI {              // I = IL_old + IL_new + alpha * IGp     (event_pano_warper.cpp:199,213)
  const float* il_old; const float* il_new; const float* igp; float alpha;
  __device__ __forceinline__ float load(int, int x, int y, int W) const {
    const long long i = (long long)y * W + x;
    const float il = il_old[i] + il_new[i];
    return igp ? igp[i] * alpha + il : il;
  }
};

template <int C> struct PixT;
template <> struct PixT<1> { using type = float; };
template <> struct PixT<4> { using type = float4; };

__device__ __forceinline__ float pmul(float w, float a) { return w * a; }
__device__ __forceinline__ float pfma(float w, float a, float s) { return fmaf(w, a, s); }
__device__ __forceinline__ float padd(float a, float b) { return a + b; }
__device__ __forceinline__ float4 pmul(float w, float4 a) { return make_float4(w * a.x, w * a.y, w * a.z, w * a.w); }
__device__ __forceinline__ float4 pfma(float w, float4 a, float4 s) {
  return make_float4(fmaf(w, a.x, s.x), fmaf(w, a.y, s.y), fmaf(w, a.z, s.z), fmaf(w, a.w, s.w));
}
__device__ __forceinline__ float4 padd(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct ReduceOut {
  double* acc;          // [n_hyp][kNAcc]
  unsigned int* ticket; // [n_hyp]
  double* result;       // [n_hyp][4]  contrast, g0, g1, g2
  double* mean;         // [n_hyp]     mean of the blurred image (for the adjoint pass)
};

// Dynamic shared memory: in[(TH+2r)*(TW+2r)] + tmp[(TH+2r)*TW] pixels + reduction scratch.
template <int C, class Src, bool WRITE_OUT>
__global__ void __launch_bounds__(kImgThreads)
blur_reduce_kernel(Src src, int W, int H, Taps taps, typename PixT<C>::type* out, long long out_stride_h,
                   ReduceOut ro, int measure) {
  using Pix = typename PixT<C>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int r = taps.r;
  const int IW = kTW + 2 * r, IH = kTH + 2 * r;
  Pix* s_in = reinterpret_cast<Pix*>(smem_raw);
  Pix* s_tmp = s_in + IW * IH;
  double* s_red = reinterpret_cast<double*>(s_tmp + IH * kTW);

  const int h = blockIdx.z;
  const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
  const int tid = threadIdx.x;

  for (int i = tid; i < IW * IH; i += kImgThreads) {
    const int ly = i / IW, lx = i - ly * IW;
    const int gx = reflect101(min(tx0 + lx - r, W + r), W);
    const int gy = reflect101(min(ty0 + ly - r, H + r), H);
    s_in[i] = src.load(h, gx, gy, W);
  }
  __syncthreads();
  // row pass: s = w0*x0; s = fma(w_j, x_j, s)
  for (int i = tid; i < IH * kTW; i += kImgThreads) {
    const int ly = i / kTW, lx = i - ly * kTW;
    const Pix* p = s_in + ly * IW + lx;
    Pix s = pmul(taps.w[0], p[0]);
    for (int j = 1; j <= 2 * r; ++j) s = pfma(taps.w[j], p[j], s);
    s_tmp[i] = s;
  }
  __syncthreads();
  // column pass (symmetric form) + reduction
  double a[kNAcc];
#pragma unroll
  for (int i = 0; i < kNAcc; ++i) a[i] = 0.0;
  const int lx = tid & (kTW - 1);
  for (int ly = tid / kTW; ly < kTH; ly += kImgThreads / kTW) {
    const int gx = tx0 + lx, gy = ty0 + ly;
    if (gx < W && gy < H) {
      const Pix* c = s_tmp + (ly + r) * kTW + lx;
      Pix s = pmul(taps.w[r], c[0]);
      for (int j = 1; j <= r; ++j) s = pfma(taps.w[r + j], padd(c[j * kTW], c[-j * kTW]), s);
      if (WRITE_OUT) out[h * out_stride_h + (long long)gy * W + gx] = s;
      if constexpr (C == 1) {
        const double v = (double)s;
        a[0] += v; a[1] += v * v;
      } else {
        const double v = (double)s.x;
        a[0] += v; a[1] += v * v;
        a[2] += (double)s.y; a[3] += (double)s.z; a[4] += (double)s.w;
        a[5] += v * (double)s.y; a[6] += v * (double)s.z; a[7] += v * (double)s.w;
      }
    }
  }
  constexpr int NV = (C == 1) ? 2 : kNAcc;
  double* acc = ro.acc + (long long)h * kNAcc;
  {
    double v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = a[i];
    block_atomic_add<NV>(v, acc, s_red);
  }
  // last CTA of this hypothesis finalises
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    is_last = (atomicAdd(ro.ticket + h, 1u) == total - 1);
  }
  __syncthreads();
  if (is_last && tid == 0) {
    __threadfence();
    volatile double* va = acc;
    const double Np = (double)W * (double)H;
    const double S1 = va[0], S2 = va[1];
    const double mean = S1 / Np;
    double contrast;
    double* res = ro.result + (long long)h * 4;
    if (measure == CMAXB_CONTRAST_MEAN_SQUARE) {
      contrast = S2 / Np;                                    // cv::norm(L2SQR)/N
      if (C == 4) for (int c = 0; c < 3; ++c) res[1 + c] = 2.0 * (va[5 + c] / Np);
    } else {
      double var = S2 / Np - mean * mean;                    // cv::meanStdDev
      if (var < 0.0) var = 0.0;
      const double sd = sqrt(var);
      contrast = sd * sd;                                    // stddev[0]*stddev[0]
      if (C == 4) for (int c = 0; c < 3; ++c) res[1 + c] = 2.0 * (va[5 + c] / Np - mean * (va[2 + c] / Np));
    }
    res[0] = contrast;
    ro.mean[h] = mean;
    for (int i = 0; i < kNAcc; ++i) va[i] = 0.0;
    ro.ticket[h] = 0u;
    __threadfence();
  }
}

template <int C>
inline size_t blur_smem_bytes(int r) {
  const size_t pix = (C == 1) ? sizeof(float) : sizeof(float4);
  const int IW = kTW + 2 * r, IH = kTH + 2 * r;
  return pix * ((size_t)IW * IH + (size_t)IH * kTW) + sizeof(double) * (kImgThreads / 32) * kNAcc;
}

// ---- adjoint blur -------------------------------------------------------------------------------
// G = B^T z with z = 2*(I - mean) (variance) or 2*I (mean square), B = separable Gaussian with
// BORDER_REFLECT_101.  For one axis of length n and zero-extended z0:
//   (B^T z)(q) = conv(q) + [1<=q<=r] conv(-q) + [n-1-r<=q<=n-2] conv(2(n-1)-q),
//   conv(j) = sum_d w[r+d] z0(j+d).
// Then  g_j = (1/Np) sum_events sum_corners dw_c^(j) * G(corner)  reproduces
// mean( 2(I-mu) .* (blur(D_j) - mean(blur(D_j))) ) of local_focus_funcs.cpp:36-41 /
// global_focus_funcs.cpp:39-43 (the mean(blur(D_j)) term multiplies sum(2(I-mu)) == 0).
static __global__ void __launch_bounds__(kImgThreads)
adjoint_blur_kernel(const float* __restrict__ blurred, long long stride_h, int W, int H, Taps taps,
                    const double* __restrict__ mean, int measure, float* __restrict__ G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int r = taps.r;
  const int IW = kTW + 2 * r, IH = kTH + 2 * r;
  float* s_in = reinterpret_cast<float*>(smem_raw);
  float* s_tmp = s_in + IW * IH;
  const int h = blockIdx.z;
  const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
  const int tid = threadIdx.x;
  const float a2 = 2.0f;
  const float b2 = (measure == CMAXB_CONTRAST_MEAN_SQUARE) ? 0.0f : (float)(-2.0 * mean[h]);
  const float* img = blurred + h * stride_h;
  for (int i = tid; i < IW * IH; i += kImgThreads) {
    const int ly = i / IW, lx = i - ly * IW;
    const int gx = tx0 + lx - r, gy = ty0 + ly - r;
    float z = 0.f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H) z = img[(long long)gy * W + gx] * a2 + b2;  // img_zeromean (f32)
    s_in[i] = z;
  }
  __syncthreads();
  // x-adjoint for every staged row
  for (int i = tid; i < IH * kTW; i += kImgThreads) {
    const int ly = i / kTW, lx = i - ly * kTW;
    const int q = tx0 + lx;
    const float* row = s_in + ly * IW;     // row[j] holds z0 at x = tx0 - r + j
    float s = 0.f;
    if (q < W) {
      for (int d = -r; d <= r; ++d) s = fmaf(taps.w[r + d], row[lx + r + d], s);
      if (q >= 1 && q <= r)
        for (int d = q; d <= r; ++d) s = fmaf(taps.w[r + d], row[(-q + d) - tx0 + r], s);
      if (q <= W - 2 && q >= W - 1 - r)
        for (int d = -r; d <= q - (W - 1); ++d) s = fmaf(taps.w[r + d], row[(2 * (W - 1) - q + d) - tx0 + r], s);
    }
    s_tmp[i] = s;
  }
  __syncthreads();
  const int lx = tid & (kTW - 1);
  for (int ly = tid / kTW; ly < kTH; ly += kImgThreads / kTW) {
    const int gx = tx0 + lx, q = ty0 + ly;
    if (gx < W && q < H) {
      const float* col = s_tmp + lx;       // col[j*kTW] holds the row at y = ty0 - r + j
      float s = 0.f;
      for (int d = -r; d <= r; ++d) s = fmaf(taps.w[r + d], col[(ly + r + d) * kTW], s);
      if (q >= 1 && q <= r)
        for (int d = q; d <= r; ++d) s = fmaf(taps.w[r + d], col[((-q + d) - ty0 + r) * kTW], s);
      if (q <= H - 2 && q >= H - 1 - r)
        for (int d = -r; d <= q - (H - 1); ++d) s = fmaf(taps.w[r + d], col[((2 * (H - 1) - q + d) - ty0 + r) * kTW], s);
      G[h * stride_h + (long long)q * W + gx] = s;
    }
  }
}
inline size_t adjoint_smem_bytes(int r) {
  const int IW = kTW + 2 * r, IH = kTH + 2 * r;
  return sizeof(float) * ((size_t)IW * IH + (size_t)IH * kTW);
}

}  // namespace cmaxb
