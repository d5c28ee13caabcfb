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

constexpr int kTW = 32, kTH = 32, kImgThreads = 256;   // 32 x 32 tiles: the halo rows cost 25 % (16-row tiles: 50 %)
constexpr int kNAcc = 8;  // S1, S2, SD[3], SID[3]

__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * (len - 1) - p;
  return p;
}

// ---- pixel sources ------------------------------------------------------------------------------
struct SrcPlane {            // one float plane per hypothesis
  static constexpr bool kQuad = false;
  const float* base; long long stride_h;
  __device__ __forceinline__ float load(int h, int x, int y, int W) const {
    return base[h * stride_h + (long long)y * W + x];
  }
};
struct SrcPlane4 {           // interleaved (I, D0, D1, D2)
  static constexpr bool kQuad = false;
  const float4* base; long long stride_h;
  __device__ __forceinline__ float4 load(int h, int x, int y, int W) const {
    return base[h * stride_h + (long long)y * W + x];
  }
};
// Corner-split accumulator ("quad" image): cell (y,x) holds the four bilinear votes of the events
// whose truncated position is (y,x) -- (.x -> pixel (y,x), .y -> (y,x+1), .z -> (y+1,x),
// .w -> (y+1,x+1)) -- so the scatter issues ONE 16-byte vector reduction per event instead of four
// scalar ones (the scatter is bound by L2 reduction operations, not bytes).  The image proper is
// re-assembled here, at load time, in a fixed order.
struct SrcQuad {
  static constexpr bool kQuad = true;
  const float4* base; long long stride_h;
  static constexpr bool kExtra = false;
  __device__ __forceinline__ float4 cell(int h, int x, int y, int W) const { return __ldcg(base + h * stride_h + (long long)y * W + x); }
  __device__ __forceinline__ float extra(int, int, int) const { return 0.f; }
  __device__ __forceinline__ float combine(float, float v) const { return v; }
  // __ldcg (L2-coherent) rather than the read-only path: in the fused evaluation kernel the image
  // was written by other SMs earlier in the SAME launch.
  __device__ __forceinline__ float load(int h, int x, int y, int W) const {
    const float4* q = base + h * stride_h + (long long)y * W + x;
    float v = __ldcg(q).x;
    if (x > 0) v += __ldcg(q - 1).y;
    if (y > 0) {
      v += __ldcg(q - W).z;
      if (x > 0) v += __ldcg(q - W - 1).w;
    }
    return v;
  }
};
// back-end: I = IL + alpha * IGp with IL held as a quad image
struct SrcBeQuad {
  static constexpr bool kQuad = true;
  SrcQuad il; const float* igp; float alpha;
  static constexpr bool kExtra = true;   // + alpha * IGp: fetched with the cells (batched), combined at assembly
  __device__ __forceinline__ float4 cell(int h, int x, int y, int W) const { return il.cell(h, x, y, W); }
  __device__ __forceinline__ float extra(int x, int y, int W) const { return igp ? __ldg(igp + (long long)y * W + x) : 0.f; }
  __device__ __forceinline__ float combine(float e, float v) const { return igp ? e * alpha + v : v; }
  __device__ __forceinline__ float load(int h, int x, int y, int W) const {
    const float v = il.load(h, x, y, W);
    return igp ? igp[(long long)y * W + x] * alpha + v : v;
  }
};
struct SrcBeI {              // I = IL_old + IL_new + alpha * IGp     (event_pano_warper.cpp:199,213)
  static constexpr bool kQuad = false;
  const float* il_old; const float* il_new; const float* igp; float alpha;
  __device__ __forceinline__ float load(int, int x, int y, int W) const {
    const long long i = (long long)y * W + x;
    const float il = il_old[i] + il_new[i];
    return igp ? igp[i] * alpha + il : il;
  }
};

struct SrcBePlane {          // I = IL + alpha * IGp with IL already assembled (and summed across ranks) as a float plane
  static constexpr bool kQuad = false;
  const float* il; const float* igp; float alpha;
  __device__ __forceinline__ float load(int, int x, int y, int W) const {
    const long long i = (long long)y * W + x;
    const float v = il[i];
    return igp ? igp[i] * alpha + v : v;
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

constexpr int kMaxImgCtas = 148 * 16;  // CTAs per image plane (larger images: each CTA strides over several tiles)

struct ReduceOut {
  double* partials;     // [n_planes][kMaxImgCtas][kNAcc]  per-CTA sums (plain stores, no atomics)
  unsigned int* ticket; // [n_planes]
  double* result;       // [n_planes][4]  contrast, g0, g1, g2
  double* mean;         // [n_planes]     mean of the blurred image (for the adjoint pass)
  double* host_result = nullptr;   // optional mapped host copy of result[0] of plane 0 (no D2H copy on the stream)
  double* raw = nullptr;        // optional [n_planes][2]: the plain sums S1, S2 (row-band evaluation: summed across ranks by the caller)
  int sum_y0 = 0, sum_y1 = 0x7fffffff;   // rows whose pixels enter the sums (row-band evaluation: the band's own rows)
};

// Deterministic block-wide sum of NV doubles per thread; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[wid * NV + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0;
      for (int w = 0; w < nw; ++w) s += red[w * NV + i];
      v[i] = s;
    }
  }
}

// Dynamic shared memory: in[(TH+2r)*(TW+2r)] + tmp[(TH+2r)*TW] pixels + reduction scratch.
// Persistent CTAs: blockIdx.x strides over the tiles of plane blockIdx.z; every CTA keeps its f64
// sums in registers, stores ONE partial record, and the last CTA to arrive (atomic ticket) adds the
// records in a fixed order -- the reduction is deterministic for a given image and free of
// same-address atomics.
// R >= 0: compile-time radius (unrolled filter, constant index math); R < 0: runtime taps.r.
// zero_ptr != nullptr: the CTA also clears its tiles of ANOTHER accumulator image (the one the next
// evaluation scatters into), which removes the separate memset from the evaluation.
template <int C, class Src, bool WRITE_OUT, int R = -1>
__global__ void __launch_bounds__(kImgThreads)
blur_reduce_kernel(Src src, int W, int H, Taps taps, typename PixT<C>::type* out, long long out_stride_h,
                   ReduceOut ro, int measure, float4* zero_ptr = nullptr, long long zero_stride_h = 0) {
  using Pix = typename PixT<C>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int r = (R >= 0) ? R : taps.r;
  const int IW = kTW + 2 * r, IH = kTH + 2 * r;
  // quad sources: the cells of the tile (+halo+1) are staged once as float4 (one 16-byte request per
  // cell, cells outside the image as zero) and the pixels are assembled from shared memory
  const int QW = IW + 1, QH = IH + 1;
  float4* s_q = reinterpret_cast<float4*>(smem_raw);
  Pix* s_in = reinterpret_cast<Pix*>(smem_raw + (Src::kQuad ? sizeof(float4) * QW * QH : 0));
  Pix* s_tmp = s_in + IW * IH;
  double* s_red = reinterpret_cast<double*>(s_tmp + IH * kTW);

  const int h = blockIdx.z;
  const int tid = threadIdx.x;
  const int ntx = (W + kTW - 1) / kTW, nty = (H + kTH - 1) / kTH;
  constexpr int NV = (C == 1) ? 2 : kNAcc;
  double a[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = 0.0;

  for (int tile = blockIdx.x; tile < ntx * nty; tile += gridDim.x) {
    const int tx0 = (tile % ntx) * kTW, ty0 = (tile / ntx) * kTH;
    __syncthreads();   // previous tile's column pass is done with s_tmp / s_in
    if constexpr (Src::kQuad) {
      const int qx0 = tx0 - r - 1, qy0 = ty0 - r - 1;
      // all of a thread's requests (cells, and the IGp pixels of the back-end) are issued back to back and
      // only then stored: one L2 round trip per tile instead of one per element
      constexpr int kPerThread = 6;
      for (int base = 0; base < QW * QH; base += kPerThread * kImgThreads) {
        float4 v[kPerThread];
#pragma unroll
        for (int u = 0; u < kPerThread; ++u) {
          const int i = base + u * kImgThreads + tid;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < QW * QH) {
            const int ly = i / QW, lx = i - ly * QW;
            const int gx = qx0 + lx, gy = qy0 + ly;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) v[u] = src.cell(h, gx, gy, W);
          }
        }
#pragma unroll
        for (int u = 0; u < kPerThread; ++u) {
          const int i = base + u * kImgThreads + tid;
          if (i < QW * QH) s_q[i] = v[u];
        }
      }
      if constexpr (Src::kExtra) {
        for (int base = 0; base < IW * IH; base += kPerThread * kImgThreads) {
          float e[kPerThread];
#pragma unroll
          for (int u = 0; u < kPerThread; ++u) {
            const int i = base + u * kImgThreads + tid;
            e[u] = 0.f;
            if (i < IW * IH) {
              const int ly = i / IW, lx = i - ly * IW;
              e[u] = src.extra(reflect101(min(tx0 + lx - r, W + r), W), reflect101(min(ty0 + ly - r, H + r), H), W);
            }
          }
#pragma unroll
          for (int u = 0; u < kPerThread; ++u) {
            const int i = base + u * kImgThreads + tid;
            if (i < IW * IH) s_in[i] = e[u];
          }
        }
      }
      __syncthreads();
      for (int i = tid; i < IW * IH; i += kImgThreads) {
        const int ly = i / IW, lx = i - ly * IW;
        const int gx = reflect101(min(tx0 + lx - r, W + r), W);
        const int gy = reflect101(min(ty0 + ly - r, H + r), H);
        const int cx = gx - qx0, cy = gy - qy0;
        float v = 0.f;
        if (cx >= 1 && cy >= 1 && cx < QW && cy < QH) {
          const float4* c = s_q + cy * QW + cx;
          v = c[0].x;
          v += c[-1].y;
          v += c[-QW].z;
          v += c[-QW - 1].w;
          if constexpr (Src::kExtra) v = src.combine(s_in[i], v);
        }
        s_in[i] = v;
      }
    } else {
      for (int i = tid; i < IW * IH; i += kImgThreads) {
        const int ly = i / IW, lx = i - ly * IW;
        const int gx = reflect101(min(tx0 + lx - r, W + r), W);
        const int gy = reflect101(min(ty0 + ly - r, H + r), H);
        s_in[i] = src.load(h, gx, gy, W);
      }
    }
    __syncthreads();
    // row pass: s = w0*x0; s = fma(w_j, x_j, s)
    for (int i = tid; i < IH * kTW; i += kImgThreads) {
      const int ly = i / kTW, lx = i - ly * kTW;
      const Pix* p = s_in + ly * IW + lx;
      Pix s = pmul(taps.w[0], p[0]);
#pragma unroll
      for (int j = 1; j <= 2 * r; ++j) s = pfma(taps.w[j], p[j], s);
      s_tmp[i] = s;
    }
    __syncthreads();
    // column pass (symmetric form) + reduction
    const int lx = tid & (kTW - 1);
    for (int ly = tid / kTW; ly < kTH; ly += kImgThreads / kTW) {
      const int gx = tx0 + lx, gy = ty0 + ly;
      if (gx < W && gy < H) {
        const Pix* c = s_tmp + (ly + r) * kTW + lx;
        Pix s = pmul(taps.w[r], c[0]);
#pragma unroll
        for (int j = 1; j <= r; ++j) s = pfma(taps.w[r + j], padd(c[j * kTW], c[-j * kTW]), s);
        if (WRITE_OUT) out[h * out_stride_h + (long long)gy * W + gx] = s;
        if (gy < ro.sum_y0 || gy >= ro.sum_y1) {
          // outside the rows this launch is responsible for (halo rows of a row band)
        } else if constexpr (C == 1) {
          const double v = (double)s;
          a[0] += v; a[1] += v * v;
        } else {
          const double v = (double)s.x;
          a[0] += v; a[1] += v * v;
          a[2] += (double)s.y; a[3] += (double)s.z; a[4] += (double)s.w;
          a[5] += v * (double)s.y; a[6] += v * (double)s.z; a[7] += v * (double)s.w;
        }
        if (zero_ptr) zero_ptr[h * zero_stride_h + (long long)gy * W + gx] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  // one partial record per CTA
  block_sum<NV>(a, s_red);
  double* part = ro.partials + ((long long)h * kMaxImgCtas + blockIdx.x) * kNAcc;
  __shared__ bool is_last;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) part[i] = a[i];
    __threadfence();
    is_last = (atomicAdd(ro.ticket + h, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // last CTA: fixed-order sum of the records
  __threadfence();
  double t[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) t[i] = 0.0;
  const volatile double* all = ro.partials + (long long)h * kMaxImgCtas * kNAcc;
  for (int c = tid; c < (int)gridDim.x; c += kImgThreads) {
#pragma unroll
    for (int i = 0; i < NV; ++i) t[i] += all[(long long)c * kNAcc + i];
  }
  block_sum<NV>(t, s_red);
  if (tid == 0) {
    const double Np = (double)W * (double)H;
    const double S1 = t[0], S2 = t[1];
    const double mean = S1 / Np;
    double contrast;
    double* res = ro.result + (long long)h * 4;
    if (measure == CMAXB_CONTRAST_MEAN_SQUARE) {
      contrast = S2 / Np;                                    // cv::norm(L2SQR)/N
      if constexpr (C == 4) for (int c = 0; c < 3; ++c) res[1 + c] = 2.0 * (t[5 + c] / Np);
    } else {
      double var = S2 / Np - mean * mean;                    // cv::meanStdDev
      if (var < 0.0) var = 0.0;
      const double sd = sqrt(var);
      contrast = sd * sd;                                    // stddev[0]*stddev[0]
      if constexpr (C == 4) for (int c = 0; c < 3; ++c) res[1 + c] = 2.0 * (t[5 + c] / Np - mean * (t[2 + c] / Np));
    }
    res[0] = contrast;
    if (ro.host_result && h == 0) ro.host_result[0] = contrast;
    if (ro.raw) { ro.raw[2 * h] = S1; ro.raw[2 * h + 1] = S2; }
    ro.mean[h] = mean;
    ro.ticket[h] = 0u;
    __threadfence();
  }
}

template <int C>
inline size_t blur_smem_bytes(int r, bool quad = false) {
  const size_t pix = (C == 1) ? sizeof(float) : sizeof(float4);
  const int IW = kTW + 2 * r, IH = kTH + 2 * r;
  return pix * ((size_t)IW * IH + (size_t)IH * kTW) + sizeof(double) * (kImgThreads / 32) * kNAcc +
         (quad ? sizeof(float4) * (size_t)(IW + 1) * (IH + 1) : 0);
}

inline dim3 image_grid(int W, int H, int planes) { return dim3((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, planes); }
inline dim3 image_grid_persistent(int W, int H, int planes) {
  const int tiles = ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH);
  return dim3(tiles < kMaxImgCtas ? tiles : kMaxImgCtas, 1, planes);
}

// Host launcher: picks the unrolled radius-4 instantiation (sigma = 1, the value every launch file
// of the reference uses) or the runtime-radius one, and opts in to the dynamic shared memory once.
template <int C, class Src, bool WRITE_OUT, int R>
inline cudaError_t launch_blur_reduce_r(cudaStream_t s, int planes, const Src& src, int W, int H, const Taps& taps,
                                        typename PixT<C>::type* out, long long out_stride_h, const ReduceOut& ro, int measure,
                                        float4* zero_ptr, long long zero_stride_h) {
  auto kern = blur_reduce_kernel<C, Src, WRITE_OUT, R>;
  const size_t smem = blur_smem_bytes<C>(taps.r, Src::kQuad);
  static size_t configured[64] = {};   // per device: the attribute belongs to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = smem;
  }
  kern<<<image_grid_persistent(W, H, planes), kImgThreads, smem, s>>>(src, W, H, taps, out, out_stride_h, ro, measure, zero_ptr, zero_stride_h);
  return cudaGetLastError();
}
template <int C, class Src, bool WRITE_OUT>
inline cudaError_t launch_blur_reduce(cudaStream_t s, int planes, const Src& src, int W, int H, const Taps& taps,
                                      typename PixT<C>::type* out, long long out_stride_h, const ReduceOut& ro, int measure,
                                      float4* zero_ptr = nullptr, long long zero_stride_h = 0) {
  if (taps.r == 4) return launch_blur_reduce_r<C, Src, WRITE_OUT, 4>(s, planes, src, W, H, taps, out, out_stride_h, ro, measure, zero_ptr, zero_stride_h);
  if (taps.r == 0) return launch_blur_reduce_r<C, Src, WRITE_OUT, 0>(s, planes, src, W, H, taps, out, out_stride_h, ro, measure, zero_ptr, zero_stride_h);
  return launch_blur_reduce_r<C, Src, WRITE_OUT, -1>(s, planes, src, W, H, taps, out, out_stride_h, ro, measure, zero_ptr, zero_stride_h);
}

// ---- adjoint blur -------------------------------------------------------------------------------
// G = B^T z with z = 2*(I - mean) (variance) or 2*I (mean square), B = separable Gaussian with
// BORDER_REFLECT_101.  For one axis of length n and zero-extended z0:
//   (B^T z)(q) = conv(q) + [1<=q<=r] conv(-q) + [n-1-r<=q<=n-2] conv(2(n-1)-q),
//   conv(j) = sum_d w[r+d] z0(j+d).
// Then  g_j = (1/Np) sum_events sum_corners dw_c^(j) * G(corner)  reproduces
// mean( 2(I-mu) .* (blur(D_j) - mean(blur(D_j))) ) of local_focus_funcs.cpp:36-41 /
// global_focus_funcs.cpp:39-43 (the mean(blur(D_j)) term multiplies sum(2(I-mu)) == 0).
// QUAD_OUT: G is written as one float4 per CELL, (G(y,x), G(y,x+1), G(y+1,x), G(y+1,x+1)), so the
// gather needs a single 16-byte load per event instead of two sector requests.
template <bool QUAD_OUT, int R = -1>
__global__ void __launch_bounds__(kImgThreads)
adjoint_blur_kernel(const float* __restrict__ blurred, long long stride_h, int W, int H, Taps taps,
                    const double* __restrict__ mean, int measure, float* __restrict__ G, float4* __restrict__ GQ) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int r = (R >= 0) ? R : taps.r;
  const int IW = kTW + 1 + 2 * r, IH = kTH + 1 + 2 * r;   // one extra row/column for the quad packing
  constexpr int OW = kTW + 1, OH = kTH + 1;
  float* s_in = reinterpret_cast<float*>(smem_raw);
  float* s_tmp = s_in + IW * IH;       // [IH][OW]
  float* s_g = s_tmp + IH * OW;        // [OH][OW]
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
  for (int i = tid; i < IH * OW; i += kImgThreads) {
    const int ly = i / OW, lx = i - ly * OW;
    const int q = tx0 + lx;
    const float* row = s_in + ly * IW;     // row[j] holds z0 at x = tx0 - r + j
    float s = 0.f;
    if (q < W) {
#pragma unroll
      for (int d = -r; d <= r; ++d) s = fmaf(taps.w[r + d], row[lx + r + d], s);
      if (q >= 1 && q <= r)
        for (int d = q; d <= r; ++d) s = fmaf(taps.w[r + d], row[(-q + d) - tx0 + r], s);
      if (q <= W - 2 && q >= W - 1 - r)
        for (int d = -r; d <= q - (W - 1); ++d) s = fmaf(taps.w[r + d], row[(2 * (W - 1) - q + d) - tx0 + r], s);
    }
    s_tmp[i] = s;
  }
  __syncthreads();
  for (int i = tid; i < OH * OW; i += kImgThreads) {
    const int ly = i / OW, lx = i - ly * OW;
    const int gx = tx0 + lx, q = ty0 + ly;
    float s = 0.f;
    if (gx < W && q < H) {
      const float* col = s_tmp + lx;       // col[j*OW] holds the row at y = ty0 - r + j
#pragma unroll
      for (int d = -r; d <= r; ++d) s = fmaf(taps.w[r + d], col[(ly + r + d) * OW], s);
      if (q >= 1 && q <= r)
        for (int d = q; d <= r; ++d) s = fmaf(taps.w[r + d], col[((-q + d) - ty0 + r) * OW], s);
      if (q <= H - 2 && q >= H - 1 - r)
        for (int d = -r; d <= q - (H - 1); ++d) s = fmaf(taps.w[r + d], col[((2 * (H - 1) - q + d) - ty0 + r) * OW], s);
    }
    s_g[i] = s;
  }
  __syncthreads();
  for (int i = tid; i < kTW * kTH; i += kImgThreads) {
    const int ly = i / kTW, lx = i & (kTW - 1);
    const int gx = tx0 + lx, gy = ty0 + ly;
    if (gx < W && gy < H) {
      const float* p = s_g + ly * OW + lx;
      if (QUAD_OUT) GQ[h * stride_h + (long long)gy * W + gx] = make_float4(p[0], p[1], p[OW], p[OW + 1]);
      else G[h * stride_h + (long long)gy * W + gx] = p[0];
    }
  }
}
inline size_t adjoint_smem_bytes(int r) {
  const int IW = kTW + 1 + 2 * r, IH = kTH + 1 + 2 * r;
  return sizeof(float) * ((size_t)IW * IH + (size_t)IH * (kTW + 1) + (size_t)(kTH + 1) * (kTW + 1));
}
template <bool QUAD_OUT, int R>
inline cudaError_t launch_adjoint_blur_r(cudaStream_t s, int planes, const float* blurred, long long stride_h, int W, int H,
                                         const Taps& taps, const double* mean, int measure, float* G, float4* GQ);
template <bool QUAD_OUT>
inline cudaError_t launch_adjoint_blur(cudaStream_t s, int planes, const float* blurred, long long stride_h, int W, int H,
                                       const Taps& taps, const double* mean, int measure, float* G, float4* GQ) {
  if (taps.r == 4) return launch_adjoint_blur_r<QUAD_OUT, 4>(s, planes, blurred, stride_h, W, H, taps, mean, measure, G, GQ);
  return launch_adjoint_blur_r<QUAD_OUT, -1>(s, planes, blurred, stride_h, W, H, taps, mean, measure, G, GQ);
}
template <bool QUAD_OUT, int R>
inline cudaError_t launch_adjoint_blur_r(cudaStream_t s, int planes, const float* blurred, long long stride_h, int W, int H,
                                         const Taps& taps, const double* mean, int measure, float* G, float4* GQ) {
  auto kern = adjoint_blur_kernel<QUAD_OUT, R>;
  const size_t smem = adjoint_smem_bytes(taps.r);
  static size_t configured[64] = {};   // per device: the attribute belongs to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = smem;
  }
  kern<<<image_grid(W, H, planes), kImgThreads, smem, s>>>(blurred, stride_h, W, H, taps, mean, measure, G, GQ);
  return cudaGetLastError();
}

}  // namespace cmaxb
