// fe_mega.cuh -- the front-end cost evaluation as ONE persistent cooperative kernel.
//
// ncu on the multi-kernel pipeline (profiles/r01b_*) showed that each of its four small kernels
// keeps the SMs busy for only 5-10 us while its launch / ramp / tail costs another 5-8 us, and the
// gaps between them add ~24 us per evaluation.  Here the whole evaluation
//     scatter -> | -> blur + sums (+ clear next accumulator) -> | -> adjoint blur -> | -> gather -> | -> final
// runs in one launch of co-resident CTAs (cooperative launch, one grid barrier per '|'), the
// hypotheses come in as kernel parameters (no H2D copy) and the result is stored straight into
// mapped pinned host memory (no D2H copy).  Work decomposition, arithmetic and operation order are
// exactly those of the stand-alone kernels (fe_kernels.cuh, image_kernels.cuh).
#pragma once
#include <cooperative_groups.h>

#include "fe_kernels.cuh"
#include "image_kernels.cuh"

namespace cmaxb {

namespace cg = cooperative_groups;

constexpr int kMegaThreads = 256;
constexpr int kMegaMaxHyp = 32;       // hypotheses per launch (kernel-parameter space)
constexpr int kMegaMaxCtas = 148 * 8;

struct FeMegaParams {
  FeGeom g;
  int k;                      // hypotheses in this launch
  int th;                     // image tile height (rows), chosen so that #tiles <= #CTAs: one tile per CTA per phase
  int want_grad;
  int measure;
  Taps taps;
  double omegas[3 * kMegaMaxHyp];
  float4* quad;               // [k][A]  accumulator being filled and consumed (clean on entry)
  float4* quad_next;          // [k][A]  accumulator of the next evaluation: cleared here (or null)
  float* blurred;             // [k][A]
  float4* GQ;                 // [k][A]
  long long A;
  double* part_img;           // [k][kMegaMaxCtas][2]
  double* part_ev;            // [k][kMegaMaxCtas][3]
  unsigned int* ticket;       // arrival counter for the final reduction (re-armed by the kernel)
  double* contrast_dev;       // [k] device scratch: contrast of a gradient evaluation until the last CTA publishes it
  double* result;             // [k][4] mapped pinned host memory (device pointer)
  double* mirror;             // optional [k][4] DEVICE copy of the results (feeds an NCCL collective without a host hop)
  unsigned long long* done_flag; // mapped host word: receives `seq` after the results are visible to the host
  unsigned long long seq;
  unsigned long long* phase_ns; // optional [8]: %globaltimer of CTA 0 at every phase boundary (mapped host memory)
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define CMAXB_PHASE_MARK(idx) do { if (p.phase_ns && blockIdx.x == 0 && threadIdx.x == 0) p.phase_ns[idx] = global_timer_ns(); } while (0)
#define CMAXB_PHASE_MARK_ANY(idx) do { if (p.phase_ns && threadIdx.x == 0) p.phase_ns[idx] = global_timer_ns(); } while (0)

constexpr int kEvUnroll = 4;

__device__ __forceinline__ void mega_scatter(const FeMegaParams& p) {
  const FeGeom& g = p.g;
  // every CTA owns one contiguous run of the (tile-binned) packet, so that the LUT / accumulator lines of
  // a source tile stay in its SM's L1; kEvUnroll events per thread-iteration, all event records, dt
  // entries and LUT sectors requested before the first dependent use
  const long long chunk = (g.n + gridDim.x - 1) / gridDim.x;
  const long long c_beg = blockIdx.x * chunk;
  const long long c_end = (c_beg + chunk < g.n) ? c_beg + chunk : g.n;
  constexpr long long stride = kMegaThreads;
  for (long long i = c_beg + threadIdx.x; i < c_end; i += kEvUnroll * stride) {
    uint4 e[kEvUnroll];
    double dt[kEvUnroll];
    double2 bxy[kEvUnroll];
    double bz[kEvUnroll];
    bool ok[kEvUnroll];
    unsigned int bidx[kEvUnroll];
#pragma unroll
    for (int u = 0; u < kEvUnroll; ++u) {
      const long long j = i + u * stride;
      ok[u] = j < c_end;
      const long long jj = ok[u] ? j : i;
      if (g.bev) {
        const uint2 r = __ldg(g.bev + jj);
        e[u].x = r.x; bidx[u] = r.y;
      } else {
        e[u] = load_event(g.ev, jj);
        bidx[u] = (unsigned)jj / (unsigned)g.batch_size;
      }
    }
#pragma unroll
    for (int u = 0; u < kEvUnroll; ++u) {
      dt[u] = __ldg(g.dt_tab + bidx[u]);
      const int ex = e[u].x & 0xffff, ey = e[u].x >> 16;
      const double2* lp = reinterpret_cast<const double2*>(g.lut + (ey * g.W + ex));
      bxy[u] = __ldg(lp);
      bz[u] = __ldg(reinterpret_cast<const double*>(lp + 1));
    }
    for (int h = 0; h < p.k; ++h) {
      const double ox = p.omegas[3 * h], oy = p.omegas[3 * h + 1], oz = p.omegas[3 * h + 2];
      float4* q = p.quad + h * p.A;
#pragma unroll
      for (int u = 0; u < kEvUnroll; ++u) {
        const FeWarp w = fe_warp_b<false>(g, bxy[u].x, bxy[u].y, bz[u], dt[u], ox, oy, oz);
        if (ok[u] && w.in) {
          const float dx = w.dx, dy = w.dy;
          atomicAdd(q + (long long)w.yy * g.W + w.xx,
                    make_float4((1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy));
        }
      }
    }
  }
}

// blur + S1,S2 partial sums of hypothesis h; clears the same tiles of the next accumulator.
// The quad cells of the tile (+halo+1) are staged ONCE in shared memory as float4 (one 16-byte L2
// request per cell, all requests of a thread issued back to back), cells outside the image staged as
// zero, and the image pixels -- including the BORDER_REFLECT_101 halo -- are assembled from there.
template <int R>
__device__ __forceinline__ void mega_blur(const FeMegaParams& p, int h, unsigned char* smem_raw, bool write_out) {
  const int W = p.g.W, H = p.g.H;
  const int r = (R >= 0) ? R : p.taps.r;
  const int TH = p.th;
  const int IW = kTW + 2 * r, IH = TH + 2 * r;
  const int QW = IW + 1, QH = IH + 1;
  float4* s_q = reinterpret_cast<float4*>(smem_raw);          // [QH][QW] cells at image coords (tx0-r-1.., ty0-r-1..)
  float* s_in = reinterpret_cast<float*>(s_q + QW * QH);      // [IH][IW]
  float* s_tmp = s_in + IW * IH;                              // [IH][kTW]
  double* s_red = reinterpret_cast<double*>(s_tmp + IH * kTW);
  const int tid = threadIdx.x;
  const int ntx = (W + kTW - 1) / kTW, nty = (H + TH - 1) / TH;
  const float4* quad = p.quad + h * p.A;
  float* out = p.blurred + h * p.A;
  float4* zero_ptr = p.quad_next ? p.quad_next + h * p.A : nullptr;
  double a[2] = {0.0, 0.0};
  for (int tile = blockIdx.x; tile < ntx * nty; tile += gridDim.x) {
    const int tx0 = (tile % ntx) * kTW, ty0 = (tile / ntx) * TH;
    const int qx0 = tx0 - r - 1, qy0 = ty0 - r - 1;
    __syncthreads();
    CMAXB_PHASE_MARK(10);
    // all of a thread's cell requests are issued back to back (registers), then stored: one L2 round trip
    // per tile instead of one per cell
    constexpr int kCellsPerThread = 6;   // >= ceil((kTW+2*16+1)*(kMegaMaxTH+... )) is not needed: loop below handles the rest
    for (int base = 0; base < QW * QH; base += kCellsPerThread * kMegaThreads) {
      float4 v[kCellsPerThread];
#pragma unroll
      for (int u = 0; u < kCellsPerThread; ++u) {
        const int i = base + u * kMegaThreads + tid;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < QW * QH) {
          const int ly = i / QW, lx = i - ly * QW;
          const int gx = qx0 + lx, gy = qy0 + ly;
          if (gx >= 0 && gx < W && gy >= 0 && gy < H) v[u] = __ldcg(quad + (long long)gy * W + gx);
        }
      }
#pragma unroll
      for (int u = 0; u < kCellsPerThread; ++u) {
        const int i = base + u * kMegaThreads + tid;
        if (i < QW * QH) s_q[i] = v[u];
      }
    }
    __syncthreads();
    CMAXB_PHASE_MARK(11);
    for (int i = tid; i < IW * IH; i += kMegaThreads) {
      const int ly = i / IW, lx = i - ly * IW;
      const int gx = reflect101(min(tx0 + lx - r, W + r), W);
      const int gy = reflect101(min(ty0 + ly - r, H + r), H);
      const int cx = gx - qx0, cy = gy - qy0;                 // >= 1 by construction
      float v = 0.f;
      if (cx >= 1 && cy >= 1 && cx < QW && cy < QH) {          // always true for pixels that feed a valid output
        const float4* c = s_q + cy * QW + cx;
        v = c[0].x;
        v += c[-1].y;
        v += c[-QW].z;
        v += c[-QW - 1].w;
      }
      s_in[i] = v;
    }
    __syncthreads();
    CMAXB_PHASE_MARK(12);
    for (int i = tid; i < IH * kTW; i += kMegaThreads) {
      const int ly = i / kTW, lx = i - ly * kTW;
      const float* q = s_in + ly * IW + lx;
      float s = p.taps.w[0] * q[0];
#pragma unroll
      for (int j = 1; j <= 2 * r; ++j) s = fmaf(p.taps.w[j], q[j], s);
      s_tmp[i] = s;
    }
    __syncthreads();
    CMAXB_PHASE_MARK(13);
    const int lx = tid & (kTW - 1);
    for (int ly = tid / kTW; ly < TH; ly += kMegaThreads / kTW) {
      const int gx = tx0 + lx, gy = ty0 + ly;
      if (gx < W && gy < H) {
        const float* c = s_tmp + (ly + r) * kTW + lx;
        float s = p.taps.w[r] * c[0];
#pragma unroll
        for (int j = 1; j <= r; ++j) s = fmaf(p.taps.w[r + j], c[j * kTW] + c[-j * kTW], s);
        if (write_out) out[(long long)gy * W + gx] = s;
        const double v = (double)s;
        a[0] += v; a[1] += v * v;
        if (zero_ptr) zero_ptr[(long long)gy * W + gx] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  CMAXB_PHASE_MARK(14);
  block_sum<2>(a, s_red);
  if (tid == 0) {
    double* part = p.part_img + ((long long)h * kMegaMaxCtas + blockIdx.x) * 2;
    part[0] = a[0]; part[1] = a[1];
  }
  CMAXB_PHASE_MARK(15);
}

// every CTA adds the per-CTA records of hypothesis h in the same fixed order -> identical S1, S2
__device__ __forceinline__ void mega_image_sums(const FeMegaParams& p, int h, double* s_red, double* S1, double* S2) {
  const double* all = p.part_img + (long long)h * kMegaMaxCtas * 2;
  double t[2] = {0.0, 0.0};
  for (int c = threadIdx.x; c < (int)gridDim.x; c += kMegaThreads) {
    t[0] += __ldcg(all + 2 * c); t[1] += __ldcg(all + 2 * c + 1);
  }
  block_sum<2>(t, s_red);
  __shared__ double s_bc[2];
  if (threadIdx.x == 0) { s_bc[0] = t[0]; s_bc[1] = t[1]; }
  __syncthreads();
  *S1 = s_bc[0]; *S2 = s_bc[1];
  __syncthreads();
}

template <int R>
__device__ __forceinline__ void mega_adjoint(const FeMegaParams& p, int h, double mean, unsigned char* smem_raw) {
  const int W = p.g.W, H = p.g.H;
  const int r = (R >= 0) ? R : p.taps.r;
  const int TH = p.th;
  const int IW = kTW + 1 + 2 * r, IH = TH + 1 + 2 * r;
  constexpr int OW = kTW + 1;
  const int OH = TH + 1;
  float* s_in = reinterpret_cast<float*>(smem_raw);
  float* s_tmp = s_in + IW * IH;
  float* s_g = s_tmp + IH * OW;
  const int tid = threadIdx.x;
  const float a2 = 2.0f;
  const float b2 = (p.measure == CMAXB_CONTRAST_MEAN_SQUARE) ? 0.0f : (float)(-2.0 * mean);
  const float* img = p.blurred + h * p.A;
  float4* GQ = p.GQ + h * p.A;
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
        for (int d = -r; d <= r; ++d) s = fmaf(p.taps.w[r + d], row[lx + r + d], s);
        if (q >= 1 && q <= r)
          for (int d = q; d <= r; ++d) s = fmaf(p.taps.w[r + d], row[(-q + d) - tx0 + r], s);
        if (q <= W - 2 && q >= W - 1 - r)
          for (int d = -r; d <= q - (W - 1); ++d) s = fmaf(p.taps.w[r + d], row[(2 * (W - 1) - q + d) - tx0 + r], s);
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
        for (int d = -r; d <= r; ++d) s = fmaf(p.taps.w[r + d], col[(ly + r + d) * OW], s);
        if (q >= 1 && q <= r)
          for (int d = q; d <= r; ++d) s = fmaf(p.taps.w[r + d], col[((-q + d) - ty0 + r) * OW], s);
        if (q <= H - 2 && q >= H - 1 - r)
          for (int d = -r; d <= q - (H - 1); ++d) s = fmaf(p.taps.w[r + d], col[((2 * (H - 1) - q + d) - ty0 + r) * OW], s);
      }
      s_g[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < kTW * TH; i += kMegaThreads) {
      const int ly = i / kTW, lx = i & (kTW - 1);
      const int gx = tx0 + lx, gy = ty0 + ly;
      if (gx < W && gy < H) {
        const float* q = s_g + ly * OW + lx;
        GQ[(long long)gy * W + gx] = make_float4(q[0], q[1], q[OW], q[OW + 1]);
      }
    }
  }
}

__device__ __forceinline__ void mega_gather(const FeMegaParams& p, int h, double* s_red) {
  const FeGeom& g = p.g;
  const double ox = p.omegas[3 * h], oy = p.omegas[3 * h + 1], oz = p.omegas[3 * h + 2];
  const float4* GQh = p.GQ + h * p.A;
  double acc[3] = {0.0, 0.0, 0.0};
  constexpr int U = 2;
  const long long chunk = (g.n + gridDim.x - 1) / gridDim.x;
  const long long c_beg = blockIdx.x * chunk;
  const long long c_end = (c_beg + chunk < g.n) ? c_beg + chunk : g.n;
  constexpr long long stride = kMegaThreads;
  for (long long i = c_beg + threadIdx.x; i < c_end; i += U * stride) {
    uint4 e[U]; double dt[U]; double2 bxy[U]; double bz[U]; bool ok[U]; unsigned int bidx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long j = i + u * stride;
      ok[u] = j < c_end;
      const long long jj = ok[u] ? j : i;
      if (g.bev) {
        const uint2 r = __ldg(g.bev + jj);
        e[u].x = r.x; bidx[u] = r.y;
      } else {
        e[u] = load_event(g.ev, jj);
        bidx[u] = (unsigned)jj / (unsigned)g.batch_size;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      dt[u] = __ldg(g.dt_tab + bidx[u]);
      const int ex = e[u].x & 0xffff, ey = e[u].x >> 16;
      const double2* lp = reinterpret_cast<const double2*>(g.lut + (ey * g.W + ex));
      bxy[u] = __ldg(lp);
      bz[u] = __ldg(reinterpret_cast<const double*>(lp + 1));
    }
    FeWarp w[U];
    float4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      w[u] = fe_warp_b<true>(g, bxy[u].x, bxy[u].y, bz[u], dt[u], ox, oy, oz);
      q[u] = (ok[u] && w[u].in) ? __ldcg(GQh + (long long)w[u].yy * g.W + w[u].xx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!(ok[u] && w[u].in)) continue;
      const double g00 = q[u].x, g01 = q[u].y, g10 = q[u].z, g11 = q[u].w;
      const double dx = w[u].dx, dy = w[u].dy;
      const double a = (1.0 - dy) * (g01 - g00) + dy * (g11 - g10);
      const double b = (1.0 - dx) * (g10 - g00) + dx * (g11 - g01);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += (double)w[u].r0[c] * a + (double)w[u].r1[c] * b;
    }
  }
  block_sum<3>(acc, s_red);
  if (threadIdx.x == 0) {
    double* part = p.part_ev + ((long long)h * kMegaMaxCtas + blockIdx.x) * 3;
    part[0] = acc[0]; part[1] = acc[1]; part[2] = acc[2];
  }
}

template <int R>
__global__ void __launch_bounds__(kMegaThreads)
fe_eval_megakernel(const __grid_constant__ FeMegaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[(kMegaThreads / 32) * 3];
  __shared__ double s_mean[kMegaMaxHyp];
  cg::grid_group grid = cg::this_grid();
  const double Np = (double)p.g.W * (double)p.g.H;

  CMAXB_PHASE_MARK(0);
  if (p.g.n > 0) mega_scatter(p);
  CMAXB_PHASE_MARK(1);
  grid.sync();
  CMAXB_PHASE_MARK(2);
  for (int h = 0; h < p.k; ++h) mega_blur<R>(p, h, smem_raw, p.want_grad != 0);
  CMAXB_PHASE_MARK(3);
  grid.sync();
  CMAXB_PHASE_MARK(4);
  for (int h = 0; h < p.k; ++h) {
    double S1, S2;
    mega_image_sums(p, h, s_red, &S1, &S2);
    const double mean = S1 / Np;
    if (threadIdx.x == 0) s_mean[h] = mean;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      double contrast;
      if (p.measure == CMAXB_CONTRAST_MEAN_SQUARE) contrast = S2 / Np;
      else {
        double var = S2 / Np - mean * mean;
        if (var < 0.0) var = 0.0;
        const double sd = sqrt(var);
        contrast = sd * sd;
      }
      if (p.want_grad) {
        p.contrast_dev[h] = contrast;   // published to the host by the last CTA, together with the gradient
      } else {
        p.result[4 * h] = contrast;
        if (p.mirror) { p.mirror[4 * h] = contrast; p.mirror[4 * h + 1] = 0.0; p.mirror[4 * h + 2] = 0.0; p.mirror[4 * h + 3] = 0.0; }
      }
    }
  }
  __syncthreads();
  if (!p.want_grad) {
    CMAXB_PHASE_MARK(5);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(p.done_flag) = p.seq;
    }
    return;
  }
  for (int h = 0; h < p.k; ++h) mega_adjoint<R>(p, h, s_mean[h], smem_raw);
  CMAXB_PHASE_MARK(5);
  grid.sync();
  CMAXB_PHASE_MARK(6);
  for (int h = 0; h < p.k; ++h) {
    __syncthreads();
    mega_gather(p, h, s_red);
  }
  CMAXB_PHASE_MARK(7);
  // no fourth grid barrier: the last CTA to publish its gather record (atomic ticket) does the final sum
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  CMAXB_PHASE_MARK(8);
  if (s_last) {
    __threadfence();
    if (threadIdx.x == 0) *p.ticket = 0u;
    // one WARP per hypothesis: lanes add the per-CTA records in a fixed order, lane 0 publishes
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int h = wid; h < p.k; h += kMegaThreads / 32) {
      const double* all = p.part_ev + (long long)h * kMegaMaxCtas * 3;
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      constexpr int kRecUnroll = 8;    // 24 independent loads in flight per lane (the loop is one L2 round trip per step)
      for (int c0 = lane; c0 < (int)gridDim.x; c0 += 32 * kRecUnroll) {
        double v[kRecUnroll][3];
#pragma unroll
        for (int u = 0; u < kRecUnroll; ++u) {
          const int c = c0 + 32 * u;
          const bool ok = c < (int)gridDim.x;
          v[u][0] = ok ? __ldcg(all + 3 * c) : 0.0; v[u][1] = ok ? __ldcg(all + 3 * c + 1) : 0.0; v[u][2] = ok ? __ldcg(all + 3 * c + 2) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kRecUnroll; ++u) { t0 += v[u][0]; t1 += v[u][1]; t2 += v[u][2]; }
      }
      t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
      if (lane == 0) {
        const double c = __ldcg(p.contrast_dev + h);
        p.result[4 * h] = c;
        p.result[4 * h + 1] = t0 / Np; p.result[4 * h + 2] = t1 / Np; p.result[4 * h + 3] = t2 / Np;
        if (p.mirror) { p.mirror[4 * h] = c; p.mirror[4 * h + 1] = t0 / Np; p.mirror[4 * h + 2] = t1 / Np; p.mirror[4 * h + 3] = t2 / Np; }
      }
    }
    __syncthreads();
    CMAXB_PHASE_MARK_ANY(9);
    if (threadIdx.x == 0) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(p.done_flag) = p.seq;
    }
  }
}

constexpr int kMegaMaxTH = 32;
inline size_t mega_smem_bytes(int r, int th = kMegaMaxTH) {
  const int IW = kTW + 2 * r, IH = th + 2 * r;
  const size_t a = sizeof(float4) * (size_t)(IW + 1) * (IH + 1) + sizeof(float) * ((size_t)IW * IH + (size_t)IH * kTW) +
                   sizeof(double) * (kMegaThreads / 32) * kNAcc;
  const int JW = kTW + 1 + 2 * r, JH = th + 1 + 2 * r;
  const size_t b = sizeof(float) * ((size_t)JW * JH + (size_t)JH * (kTW + 1) + (size_t)(th + 1) * (kTW + 1));
  return a > b ? a : b;
}
// tile height such that one hypothesis plane has at most `grid` tiles (each CTA: one tile per phase)
inline int mega_tile_height(int W, int H, int grid) {
  const int ntx = (W + kTW - 1) / kTW;
  int rows_of_tiles = grid / ntx;
  if (rows_of_tiles < 1) rows_of_tiles = 1;
  int th = (H + rows_of_tiles - 1) / rows_of_tiles;
  if (th < 8) th = 8;
  if (th > kMegaMaxTH) th = kMegaMaxTH;
  return th;
}

}  // namespace cmaxb
