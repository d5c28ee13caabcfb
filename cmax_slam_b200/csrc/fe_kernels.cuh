// fe_kernels.cuh -- front-end event kernels: per-batch dt table, first-order rotational warp +
// pinhole projection in f64 registers, bilinear scatter (value / dense derivative images) and the
// adjoint gather of the gradient.
//
// Computes what AngVelEstimator::warpAndAccumulateEvents does per event
// (src/frontend/local_image_warped_events.cpp:59-170) with canonicalProjection / applyIntrinsics /
// cross2Matrix (src/utils/image_geom_util.cpp:7-41, include/utils/image_geom_util.h:5-8).
#pragma once
#include "common.cuh"

namespace cmaxb {

struct FeGeom {
  const uint4* ev;          // raw 16-byte dvs_msgs::Event records
  long long n;
  int batch_size;
  const double* dt_tab;     // per batch: t_mid.toSec() - t_ref            (:68-75)
  const double4* lut;       // bearing vectors padded to 32 B              (:100)
  int W, H;
  double fx, fy, cx, cy;
};

struct FeWarp {
  bool in;                  // passes the bounds test                      (:142)
  int xx, yy;
  float dx, dy;
  float r0[3], r1[3];       // rows of d(pixel)/d(omega) as float          (:157-160)
};

// One event, one hypothesis.  No FMA contraction (see common.cuh).
// (bx,by,bz) = bearing vector of the event's pixel, already loaded.
// GRAD: 0 = cell + bilinear fractions only; 1 = Jacobian rows evaluated in f64 and rounded to f32 exactly as
// the reference does (:110-135,157-160); 2 = the same chain evaluated in f32 from the f64 projection (the
// rows differ from GRAD 1 by <= a few f32 ulps; used by the adjoint gather, whose sum is accumulated in f64
// and checked to 1e-5 of the gradient -- it halves the pass's f64 instruction count).
template <int GRAD>
__device__ __forceinline__ FeWarp fe_warp_b(const FeGeom& g, double bx, double by, double bz, double dt, double ox, double oy, double oz) {
  FeWarp o;
  // delta_rot = ang_vel * dt ; p' = p + delta_rot x p                     (:76,101)
  const double dlx = ox * dt, dly = oy * dt, dlz = oz * dt;
  const double px3 = bx + (dly * bz - dlz * by);
  const double py3 = by + (dlz * bx - dlx * bz);
  const double pz3 = bz + (dlx * by - dly * bx);
  // canonicalProjection                                                   (image_geom_util.cpp:29-33)
  const double inv = 1.0 / pz3;
  const double u = px3 * inv, v = py3 * inv;
  // applyIntrinsics                                                       (image_geom_util.cpp:15-16)
  const double px = g.fx * u + g.cx;
  const double py = g.fy * v + g.cy;
  o.in = false;
  o.xx = o.yy = 0;
  o.dx = o.dy = 0.f;
  if (fabs(px) < 2e9 && fabs(py) < 2e9) {
    const int xx = (int)px, yy = (int)py;                                  // truncation (:139)
    if (1 <= xx && xx < g.W - 2 && 1 <= yy && yy < g.H - 2) {
      o.in = true;
      o.xx = xx; o.yy = yy;
      o.dx = (float)(px - (double)xx);
      o.dy = (float)(py - (double)yy);
    }
  }
  if (GRAD == 2) {
    const float fu = (float)u, fv = (float)v, finv = (float)inv;
    const float ndt = -(float)dt;
    const float mx = ndt * (float)bx, my = ndt * (float)by, mz = ndt * (float)bz;
    const float a02 = -fu * finv, a12 = -fv * finv;
    const float ffx = (float)g.fx, ffy = (float)g.fy;
    o.r0[0] = ffx * (a02 * (-my));
    o.r0[1] = ffx * (finv * (-mz) + a02 * mx);
    o.r0[2] = ffx * (finv * my);
    o.r1[0] = ffy * (finv * mz + a12 * (-my));
    o.r1[1] = ffy * (a12 * mx);
    o.r1[2] = ffy * (finv * (-mx));
  }
  if (GRAD == 1) {
    // M = cross2Matrix((-dt)*p)                                           (:110)
    const double ndt = -dt;
    const double mx = ndt * bx, my = ndt * by, mz = ndt * bz;
    const double a02 = -u * inv, a12 = -v * inv;                           // Jp(0,2), Jp(1,2)
    // Jc = Jp * M (terms multiplied by structural zeros dropped: x + 0 == x)
    const double c00 = a02 * (-my);
    const double c01 = inv * (-mz) + a02 * mx;
    const double c02 = inv * my;
    const double c10 = inv * mz + a12 * (-my);
    const double c11 = a12 * mx;
    const double c12 = inv * (-mx);
    // J = diag(fx, fy) * Jc                                               (:135)
    o.r0[0] = (float)(g.fx * c00); o.r0[1] = (float)(g.fx * c01); o.r0[2] = (float)(g.fx * c02);
    o.r1[0] = (float)(g.fy * c10); o.r1[1] = (float)(g.fy * c11); o.r1[2] = (float)(g.fy * c12);
  }
  return o;
}

template <int GRAD>
__device__ __forceinline__ FeWarp fe_warp(const FeGeom& g, uint4 e, double dt, double ox, double oy, double oz) {
  const int ex = e.x & 0xffff, ey = e.x >> 16;
  const double2* lp = reinterpret_cast<const double2*>(g.lut + (ey * g.W + ex));
  const double2 bxy = __ldg(lp);
  const double bz = __ldg(reinterpret_cast<const double*>(lp + 1));
  return fe_warp_b<GRAD>(g, bxy.x, bxy.y, bz, dt, ox, oy, oz);
}

// Packet verdict words live in MAPPED host memory (no memset / copy on the stream): kernels raise them with plain stores.
// flags[0] = 1: a batch spans a negative time interval (:72); flags[1] = 1: an event lies outside the sensor (:100)
__device__ __forceinline__ void fe_flag_raise(int* flags, int which) { reinterpret_cast<volatile int*>(flags)[which] = 1; }

__global__ void fe_validate_kernel(const uint4* __restrict__ ev, long long n, int W, int H, int* flags) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 e = ev[i];
  if ((int)(e.x & 0xffff) >= W || (int)(e.x >> 16) >= H) fe_flag_raise(flags, 1);
}

// per-batch reference time offsets; flags[0] = 1 on a negative batch span (:72)
__global__ void fe_batch_dt_kernel(const uint4* __restrict__ ev, long long n, int bs, double t_ref,
                                   double* __restrict__ dt_tab, long long nb, int* flags) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const long long beg = b * bs;
  long long end = beg + bs; if (end > n) end = n;
  const uint4 e0 = ev[beg], e1 = ev[end - 1];
  RosTime mid;
  const bool ok = ros_batch_mid(RosTime{e0.y, e0.z}, RosTime{e1.y, e1.z}, &mid);
  if (!ok) fe_flag_raise(flags, 0);
  dt_tab[b] = ros_to_sec(mid.sec, mid.nsec) - t_ref;                       // (:75)
}


constexpr int kFeThreads = 256;

// MODE 2: value only  -> corner-split float4 accumulator quad[h] (one vector reduction per event;
//                        image_kernels.cuh SrcQuad re-assembles the image)
// MODE 0: value only  -> float plane img1[h]
// MODE 1: dense       -> float4 plane img4[h] = (I, dI/dwx, dI/dwy, dI/dwz), one 16-byte vector
//                        reduction per bilinear corner (red.global.add.v4.f32, sm_90+)
template <int MODE>
__global__ void __launch_bounds__(kFeThreads)
fe_scatter_kernel(FeGeom g, const double* __restrict__ omegas, float* __restrict__ img1,
                  float4* __restrict__ img4, long long stride_h) {
  const int h = blockIdx.y;
  const double ox = omegas[3 * h], oy = omegas[3 * h + 1], oz = omegas[3 * h + 2];
  const long long stride = (long long)gridDim.x * kFeThreads;
  for (long long i = blockIdx.x * (long long)kFeThreads + threadIdx.x; i < g.n; i += stride) {
    const uint4 e = load_event(g.ev, i);
    const double dt = __ldg(g.dt_tab + i / g.batch_size);
    const FeWarp w = fe_warp<(MODE == 1) ? 1 : 0>(g, e, dt, ox, oy, oz);
    if (!w.in) continue;
    const float dx = w.dx, dy = w.dy;
    const float w00 = (1.f - dx) * (1.f - dy), w01 = dx * (1.f - dy);
    const float w10 = (1.f - dx) * dy, w11 = dx * dy;
    const long long p = h * stride_h + (long long)w.yy * g.W + w.xx;
    if (MODE == 2) {
      atomicAdd(img4 + p, make_float4(w00, w01, w10, w11));
    } else if (MODE == 0) {
      atomicAdd(img1 + p, w00);
      atomicAdd(img1 + p + 1, w01);
      atomicAdd(img1 + p + g.W, w10);
      atomicAdd(img1 + p + g.W + 1, w11);
    } else {
      // derivative votes (:163-166)
      const float s00 = -(1.f - dy), t00 = -(1.f - dx);
      const float s01 = (1.f - dy), t01 = -dx;
      const float s10 = -dy, t10 = (1.f - dx);
      const float s11 = dy, t11 = dx;
      float4 v;
      v = make_float4(w00, w.r0[0] * s00 + w.r1[0] * t00, w.r0[1] * s00 + w.r1[1] * t00, w.r0[2] * s00 + w.r1[2] * t00);
      atomicAdd(img4 + p, v);
      v = make_float4(w01, w.r0[0] * s01 + w.r1[0] * t01, w.r0[1] * s01 + w.r1[1] * t01, w.r0[2] * s01 + w.r1[2] * t01);
      atomicAdd(img4 + p + 1, v);
      v = make_float4(w10, w.r0[0] * s10 + w.r1[0] * t10, w.r0[1] * s10 + w.r1[1] * t10, w.r0[2] * s10 + w.r1[2] * t10);
      atomicAdd(img4 + p + g.W, v);
      v = make_float4(w11, w.r0[0] * s11 + w.r1[0] * t11, w.r0[1] * s11 + w.r1[1] * t11, w.r0[2] * s11 + w.r1[2] * t11);
      atomicAdd(img4 + p + g.W + 1, v);
    }
  }
}

// debug / parity: per-event cell index
__global__ void fe_cells_kernel(FeGeom g, const double* __restrict__ omegas, int* __restrict__ cells) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const uint4 e = load_event(g.ev, i);
  const double dt = g.dt_tab[i / g.batch_size];
  const FeWarp w = fe_warp<0>(g, e, dt, omegas[0], omegas[1], omegas[2]);
  cells[i] = w.in ? w.yy * g.W + w.xx : -1;
}

// Adjoint gather: g_c = (1/Np) * sum_events [ r0_c * a + r1_c * b ],
//   a = sum_corners s_corner * G(corner), b = sum_corners t_corner * G(corner)
// (s, t = the derivative-vote weights of :163-166).  Last CTA finalises into result[h][1..3].
constexpr int kMaxEventCtas = 148 * 8;

template <bool QUAD>
__global__ void __launch_bounds__(kFeThreads)
fe_gather_kernel(FeGeom g, const double* __restrict__ omegas, const float* __restrict__ G, const float4* __restrict__ GQ,
                 long long stride_h, double* partials /*[n_hyp][kMaxEventCtas][3]*/, unsigned int* ticket, double* result) {
  __shared__ double s_red[(kFeThreads / 32) * 3];
  __shared__ bool is_last;
  const int h = blockIdx.y;
  const double ox = omegas[3 * h], oy = omegas[3 * h + 1], oz = omegas[3 * h + 2];
  const float* Gh = QUAD ? nullptr : G + h * stride_h;
  const float4* GQh = QUAD ? GQ + h * stride_h : nullptr;
  double acc[3] = {0.0, 0.0, 0.0};
  const long long stride = (long long)gridDim.x * kFeThreads;
  for (long long i = blockIdx.x * (long long)kFeThreads + threadIdx.x; i < g.n; i += stride) {
    const uint4 e = load_event(g.ev, i);
    const double dt = __ldg(g.dt_tab + i / g.batch_size);
    const FeWarp w = fe_warp<1>(g, e, dt, ox, oy, oz);
    if (!w.in) continue;
    double g00, g01, g10, g11;
    if (QUAD) {
      const float4 q = __ldg(GQh + (long long)w.yy * g.W + w.xx);
      g00 = q.x; g01 = q.y; g10 = q.z; g11 = q.w;
    } else {
      const float* p = Gh + (long long)w.yy * g.W + w.xx;
      g00 = __ldg(p); g01 = __ldg(p + 1); g10 = __ldg(p + g.W); g11 = __ldg(p + g.W + 1);
    }
    const double dx = w.dx, dy = w.dy;
    const double a = (1.0 - dy) * (g01 - g00) + dy * (g11 - g10);
    const double b = (1.0 - dx) * (g10 - g00) + dx * (g11 - g01);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += (double)w.r0[c] * a + (double)w.r1[c] * b;
  }
  // one partial record per CTA, fixed-order final sum by the last CTA (no same-address atomics)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) acc[c] = warp_sum(acc[c]);
  if (lane == 0) { s_red[wid * 3] = acc[0]; s_red[wid * 3 + 1] = acc[1]; s_red[wid * 3 + 2] = acc[2]; }
  __syncthreads();
  double* part = partials + ((long long)h * kMaxEventCtas + blockIdx.x) * 3;
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (int w = 0; w < kFeThreads / 32; ++w) s += s_red[w * 3 + c];
      part[c] = s;
    }
    __threadfence();
    is_last = (atomicAdd(ticket + h, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const volatile double* all = partials + (long long)h * kMaxEventCtas * 3;
  double t[3] = {0.0, 0.0, 0.0};
  for (int c = threadIdx.x; c < (int)gridDim.x; c += kFeThreads) {
    t[0] += all[3 * c]; t[1] += all[3 * c + 1]; t[2] += all[3 * c + 2];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) t[c] = warp_sum(t[c]);
  __syncthreads();
  if (lane == 0) { s_red[wid * 3] = t[0]; s_red[wid * 3 + 1] = t[1]; s_red[wid * 3 + 2] = t[2]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double Np = (double)g.W * (double)g.H;
    for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (int w = 0; w < kFeThreads / 32; ++w) s += s_red[w * 3 + c];
      result[4 * h + 1 + c] = s / Np;
    }
    ticket[h] = 0u;
    __threadfence();
  }
}

}  // namespace cmaxb
