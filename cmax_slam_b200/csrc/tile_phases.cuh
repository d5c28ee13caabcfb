// tile_phases.cuh -- image phases of the fused (single cooperative launch) evaluation kernels, shared by
// the front-end (fe_mega.cuh) and the back-end (be_mega.cuh).
//
// Every CTA strides over 32 x th tiles of one image (th chosen on the host so that an image has at most
// gridDim.x tiles: one tile per CTA per phase).  Arithmetic and operation order are those of the
// stand-alone kernels in image_kernels.cuh (OpenCV's separable-filter order; adjoint with REFLECT_101
// folding).
#pragma once
#include "image_kernels.cuh"

namespace cmaxb {

constexpr int kMegaThreads = 256;
constexpr int kMegaMaxCtas = 148 * 8;
constexpr int kMegaMaxTH = 32;

struct TileCtx {
  int W, H, th;
  Taps taps;
};

// ---- pixel sources of the blur phase ---------------------------------------------------------------
struct MQuad {               // corner-split accumulator (+ alpha * IGp for the back-end)
  static constexpr bool kQuad = true;
  const float4* quad; const float* igp; float alpha;
  __device__ __forceinline__ float4 cell(int x, int y, int W) const { return __ldcg(quad + (long long)y * W + x); }
  __device__ __forceinline__ float finish(float v, int x, int y, int W) const {
    return igp ? __ldg(igp + (long long)y * W + x) * alpha + v : v;
  }
};
struct MPlanes {             // IL_old (+ IL_new) (+ alpha * IGp) float planes
  static constexpr bool kQuad = false;
  const float* il_old; const float* il_new; const float* igp; float alpha;
  __device__ __forceinline__ float load(int x, int y, int W) const {
    const long long i = (long long)y * W + x;
    float il = __ldcg(il_old + i);
    if (il_new) il += __ldcg(il_new + i);
    return igp ? __ldg(igp + i) * alpha + il : il;
  }
};

inline size_t tile_smem_bytes(int r, int th = kMegaMaxTH) {
  const int IW = kTW + 2 * r, IH = th + 2 * r;
  const size_t a = sizeof(float4) * (size_t)(IW + 1) * (IH + 1) + sizeof(float) * ((size_t)IW * IH + (size_t)IH * kTW) +
                   sizeof(double) * (kMegaThreads / 32) * kNAcc;
  const int JW = kTW + 1 + 2 * r, JH = th + 1 + 2 * r;
  const size_t b = sizeof(float) * ((size_t)JW * JH + (size_t)JH * (kTW + 1) + (size_t)(th + 1) * (kTW + 1));
  return a > b ? a : b;
}
// tile height such that one image has at most `grid` tiles (each CTA: one tile per phase)
inline int tile_height_for_grid(int W, int H, int grid) {
  const int ntx = (W + kTW - 1) / kTW;
  int rows_of_tiles = grid / ntx;
  if (rows_of_tiles < 1) rows_of_tiles = 1;
  int th = (H + rows_of_tiles - 1) / rows_of_tiles;
  if (th < 8) th = 8;
  if (th > kMegaMaxTH) th = kMegaMaxTH;
  return th;
}

// Blur + S1, S2.  out: blurred image (or null); zero_ptr: a quad image whose tiles are cleared here (or
// null); part2: this image's per-CTA records [kMegaMaxCtas][2].
template <int R, class Src>
__device__ __forceinline__ void tile_blur_phase(const TileCtx& c, const Src& src, float* __restrict__ out,
                                                float4* __restrict__ zero_ptr, double* __restrict__ part2, unsigned char* smem_raw) {
  const int W = c.W, H = c.H;
  const int r = (R >= 0) ? R : c.taps.r;
  const int TH = c.th;
  const int IW = kTW + 2 * r, IH = TH + 2 * r;
  const int QW = IW + 1, QH = IH + 1;
  float4* s_q = reinterpret_cast<float4*>(smem_raw);          // [QH][QW] cells at image coords (tx0-r-1.., ty0-r-1..)
  float* s_in = reinterpret_cast<float*>(s_q + QW * QH);      // [IH][IW]
  float* s_tmp = s_in + IW * IH;                              // [IH][kTW]
  double* s_red = reinterpret_cast<double*>(s_tmp + IH * kTW);
  const int tid = threadIdx.x;
  const int ntx = (W + kTW - 1) / kTW, nty = (H + TH - 1) / TH;
  double a[2] = {0.0, 0.0};
  for (int tile = blockIdx.x; tile < ntx * nty; tile += gridDim.x) {
    const int tx0 = (tile % ntx) * kTW, ty0 = (tile / ntx) * TH;
    const int qx0 = tx0 - r - 1, qy0 = ty0 - r - 1;
    __syncthreads();
    if constexpr (Src::kQuad) {
      // the quad cells of the tile (+halo+1) are staged ONCE as float4, cells outside the image as zero;
      // all of a thread's requests are issued back to back, then stored (one L2 round trip per tile)
      constexpr int kCellsPerThread = 6;
      for (int base = 0; base < QW * QH; base += kCellsPerThread * kMegaThreads) {
        float4 v[kCellsPerThread];
#pragma unroll
        for (int u = 0; u < kCellsPerThread; ++u) {
          const int i = base + u * kMegaThreads + tid;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < QW * QH) {
            const int ly = i / QW, lx = i - ly * QW;
            const int gx = qx0 + lx, gy = qy0 + ly;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) v[u] = src.cell(gx, gy, W);
          }
        }
#pragma unroll
        for (int u = 0; u < kCellsPerThread; ++u) {
          const int i = base + u * kMegaThreads + tid;
          if (i < QW * QH) s_q[i] = v[u];
        }
      }
      __syncthreads();
      // assemble the image (BORDER_REFLECT_101 halo included) from the staged cells
      for (int i = tid; i < IW * IH; i += kMegaThreads) {
        const int ly = i / IW, lx = i - ly * IW;
        const int gx = reflect101(min(tx0 + lx - r, W + r), W);
        const int gy = reflect101(min(ty0 + ly - r, H + r), H);
        const int cx = gx - qx0, cy = gy - qy0;
        float v = 0.f;
        if (cx >= 1 && cy >= 1 && cx < QW && cy < QH) {        // always true for pixels that feed a valid output
          const float4* q = s_q + cy * QW + cx;
          v = q[0].x;
          v += q[-1].y;
          v += q[-QW].z;
          v += q[-QW - 1].w;
          v = src.finish(v, gx, gy, W);
        }
        s_in[i] = v;
      }
    } else {
      constexpr int kPixPerThread = 6;
      for (int base = 0; base < IW * IH; base += kPixPerThread * kMegaThreads) {
        float v[kPixPerThread];
#pragma unroll
        for (int u = 0; u < kPixPerThread; ++u) {
          const int i = base + u * kMegaThreads + tid;
          v[u] = 0.f;
          if (i < IW * IH) {
            const int ly = i / IW, lx = i - ly * IW;
            const int gx = reflect101(min(tx0 + lx - r, W + r), W);
            const int gy = reflect101(min(ty0 + ly - r, H + r), H);
            v[u] = src.load(gx, gy, W);
          }
        }
#pragma unroll
        for (int u = 0; u < kPixPerThread; ++u) {
          const int i = base + u * kMegaThreads + tid;
          if (i < IW * IH) s_in[i] = v[u];
        }
      }
    }
    __syncthreads();
    // row pass: s = w0*x0; s = fma(w_j, x_j, s)
    for (int i = tid; i < IH * kTW; i += kMegaThreads) {
      const int ly = i / kTW, lx = i - ly * kTW;
      const float* q = s_in + ly * IW + lx;
      float s = c.taps.w[0] * q[0];
#pragma unroll
      for (int j = 1; j <= 2 * r; ++j) s = fmaf(c.taps.w[j], q[j], s);
      s_tmp[i] = s;
    }
    __syncthreads();
    // column pass (symmetric form) + sums
    const int lx = tid & (kTW - 1);
    for (int ly = tid / kTW; ly < TH; ly += kMegaThreads / kTW) {
      const int gx = tx0 + lx, gy = ty0 + ly;
      if (gx < W && gy < H) {
        const float* q = s_tmp + (ly + r) * kTW + lx;
        float s = c.taps.w[r] * q[0];
#pragma unroll
        for (int j = 1; j <= r; ++j) s = fmaf(c.taps.w[r + j], q[j * kTW] + q[-j * kTW], s);
        if (out) out[(long long)gy * W + gx] = s;
        const double v = (double)s;
        a[0] += v; a[1] += v * v;
        if (zero_ptr) zero_ptr[(long long)gy * W + gx] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  block_sum<2>(a, s_red);
  if (tid == 0) {
    double* part = part2 + (long long)blockIdx.x * 2;
    part[0] = a[0]; part[1] = a[1];
  }
}

// every CTA adds the per-CTA records in the same fixed order -> identical S1, S2 everywhere
__device__ __forceinline__ void tile_sum_partials(const double* __restrict__ part2, double* s_red, double* S1, double* S2) {
  double t[2] = {0.0, 0.0};
  for (int c = threadIdx.x; c < (int)gridDim.x; c += kMegaThreads) {
    t[0] += __ldcg(part2 + 2 * c); t[1] += __ldcg(part2 + 2 * c + 1);
  }
  block_sum<2>(t, s_red);
  __shared__ double s_bc[2];
  if (threadIdx.x == 0) { s_bc[0] = t[0]; s_bc[1] = t[1]; }
  __syncthreads();
  *S1 = s_bc[0]; *S2 = s_bc[1];
  __syncthreads();
}

// G = B^T z, z = a2 * blurred + b2 (zero outside the image); written as a plane (G) or as one float4 of the
// four corner values per cell (GQ).  See adjoint_blur_kernel in image_kernels.cuh for the folding.
template <int R>
__device__ __forceinline__ void tile_adjoint_phase(const TileCtx& c, const float* __restrict__ img, float a2, float b2,
                                                   float* __restrict__ G, float4* __restrict__ GQ, unsigned char* smem_raw) {
  const int W = c.W, H = c.H;
  const int r = (R >= 0) ? R : c.taps.r;
  const int TH = c.th;
  const int IW = kTW + 1 + 2 * r, IH = TH + 1 + 2 * r;
  constexpr int OW = kTW + 1;
  const int OH = TH + 1;
  float* s_in = reinterpret_cast<float*>(smem_raw);
  float* s_tmp = s_in + IW * IH;
  float* s_g = s_tmp + IH * OW;
  const int tid = threadIdx.x;
  const int ntx = (W + kTW - 1) / kTW, nty = (H + TH - 1) / TH;
  for (int tile = blockIdx.x; tile < ntx * nty; tile += gridDim.x) {
    const int tx0 = (tile % ntx) * kTW, ty0 = (tile / ntx) * TH;
    __syncthreads();
    constexpr int kPixPerThread = 6;
    for (int base = 0; base < IW * IH; base += kPixPerThread * kMegaThreads) {
      float z[kPixPerThread];
      bool inside[kPixPerThread];
#pragma unroll
      for (int u = 0; u < kPixPerThread; ++u) {
        const int i = base + u * kMegaThreads + tid;
        z[u] = 0.f; inside[u] = false;
        if (i < IW * IH) {
          const int ly = i / IW, lx = i - ly * IW;
          const int gx = tx0 + lx - r, gy = ty0 + ly - r;
          inside[u] = gx >= 0 && gx < W && gy >= 0 && gy < H;
          if (inside[u]) z[u] = __ldcg(img + (long long)gy * W + gx);
        }
      }
#pragma unroll
      for (int u = 0; u < kPixPerThread; ++u) {
        const int i = base + u * kMegaThreads + tid;
        if (i < IW * IH) s_in[i] = inside[u] ? z[u] * a2 + b2 : 0.f;     // img_zeromean (f32), zero outside the image
      }
    }
    __syncthreads();
    for (int i = tid; i < IH * OW; i += kMegaThreads) {
      const int ly = i / OW, lx = i - ly * OW;
      const int q = tx0 + lx;
      const float* row = s_in + ly * IW;
      float s = 0.f;
      if (q < W) {
#pragma unroll
        for (int d = -r; d <= r; ++d) s = fmaf(c.taps.w[r + d], row[lx + r + d], s);
        if (q >= 1 && q <= r)
          for (int d = q; d <= r; ++d) s = fmaf(c.taps.w[r + d], row[(-q + d) - tx0 + r], s);
        if (q <= W - 2 && q >= W - 1 - r)
          for (int d = -r; d <= q - (W - 1); ++d) s = fmaf(c.taps.w[r + d], row[(2 * (W - 1) - q + d) - tx0 + r], s);
      }
      s_tmp[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < OH * OW; i += kMegaThreads) {
      const int ly = i / OW, lx = i - ly * OW;
      const int gx = tx0 + lx, q = ty0 + ly;
      float s = 0.f;
      if (gx < W && q < H) {
        const float* col = s_tmp + lx;
#pragma unroll
        for (int d = -r; d <= r; ++d) s = fmaf(c.taps.w[r + d], col[(ly + r + d) * OW], s);
        if (q >= 1 && q <= r)
          for (int d = q; d <= r; ++d) s = fmaf(c.taps.w[r + d], col[((-q + d) - ty0 + r) * OW], s);
        if (q <= H - 2 && q >= H - 1 - r)
          for (int d = -r; d <= q - (H - 1); ++d) s = fmaf(c.taps.w[r + d], col[((2 * (H - 1) - q + d) - ty0 + r) * OW], s);
      }
      s_g[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < kTW * TH; i += kMegaThreads) {
      const int ly = i / kTW, lx = i & (kTW - 1);
      const int gx = tx0 + lx, gy = ty0 + ly;
      if (gx < W && gy < H) {
        const float* q = s_g + ly * OW + lx;
        if (GQ) GQ[(long long)gy * W + gx] = make_float4(q[0], q[1], q[OW], q[OW + 1]);
        else G[(long long)gy * W + gx] = q[0];
      }
    }
  }
}

// contrast from S1, S2 (cv::meanStdDev / cv::norm semantics, local_focus_funcs.cpp:9-44)
__device__ __forceinline__ double contrast_from_sums(double S1, double S2, double Np, int measure) {
  const double mean = S1 / Np;
  if (measure == CMAXB_CONTRAST_MEAN_SQUARE) return S2 / Np;
  double var = S2 / Np - mean * mean;
  if (var < 0.0) var = 0.0;
  const double sd = sqrt(var);
  return sd * sd;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

}  // namespace cmaxb
