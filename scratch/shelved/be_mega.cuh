// be_mega.cuh -- the back-end cost evaluation as ONE persistent cooperative kernel.
//
// The stand-alone pipeline needs 7-8 launches per evaluation (poses, clear, scatter, blur, adjoint, gather,
// per-knot reduce) plus two copies and a stream synchronise; at the window sizes of the reference's launch
// files (1e5-1e6 events) that is launch-bound (~50 us of gaps per evaluation).  Here
//     poses | scatter (+cache) | blur + S1,S2 (+ clear next accumulator) | mu, contrast; adjoint | gather | per-knot sums
// run in one cooperative launch with grid barriers at '|'.  The increments x are copied to the device
// once (3*K_opt doubles), results are stored into mapped pinned host memory followed by a sequence number
// the host spins on.  Phase bodies are the device functions of be_kernels.cuh / tile_phases.cuh, so the
// arithmetic is identical to the stand-alone kernels.
#pragma once
#include <cooperative_groups.h>

#include "be_kernels.cuh"
#include "tile_phases.cuh"

namespace cmaxb {

namespace cgb = cooperative_groups;

struct BeMegaParams {
  BeGeom g;
  int N;                      // spline order
  int th;                     // image tile height
  int want_grad, measure, use_quad, n_opt;
  Taps taps;
  const Quat* knots0;
  const double* x;            // device, 3*n_opt
  const BeBatchTime* bt;
  BePose* poses;
  int* idx;
  float4* ilq;                // corner-split IL accumulator (clean on entry) -- use_quad
  float4* ilq_next;           // cleared here for the next evaluation (or null)
  float* il_old; float* il_new;   // float-plane accumulators (clean on entry) -- !use_quad
  const float* igp; float alpha;
  float* blurred; float* G; float4* GQ;
  BeCache cache;
  double* wgrad;
  const int* seg_lo; const int* seg_hi;
  double* part_img;           // [kMegaMaxCtas][2]
  double* grad_dev;           // [3*n_opt]
  double* contrast_dev;       // [1]
  unsigned int* ticket;
  double* result;             // mapped host: [0] contrast, [1..] gradient
  unsigned long long* done_flag; unsigned long long seq;
  unsigned long long* phase_ns;
};

#define CMAXB_BE_MARK(idx) do { if (p.phase_ns && blockIdx.x == 0 && threadIdx.x == 0) p.phase_ns[idx] = global_timer_ns(); } while (0)

// 3 CTAs per SM: the spline evaluation of the (tiny) pose phase would otherwise set the register budget of
// the whole kernel (126 regs -> 2 CTAs/SM); it is allowed to spill instead.
template <int N, int R>
__global__ void __launch_bounds__(kMegaThreads, 3)
be_eval_megakernel(const __grid_constant__ BeMegaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[(kMegaThreads / 32) * 3];
  cgb::grid_group grid = cgb::this_grid();
  const BeGeom& g = p.g;
  const double Np = (double)g.W * (double)g.H;
  const long long gtid = blockIdx.x * (long long)kMegaThreads + threadIdx.x;
  const long long gthreads = (long long)gridDim.x * kMegaThreads;

  // ---- poses: one thread per batch --------------------------------------------------------------
  CMAXB_BE_MARK(0);
  for (long long b = gtid; b < g.nb; b += gthreads)
    be_pose_one<N>(p.knots0, p.x, g.n_fixed, p.bt[b], p.want_grad, p.poses + b, p.idx + b);
  CMAXB_BE_MARK(1);
  grid.sync();
  // ---- scatter: one thread per visited event ------------------------------------------------------
  if (g.nb > 0) {
    if (p.use_quad) {
      if (p.want_grad) be_scatter_range<2, true>(g, p.poses, nullptr, nullptr, p.ilq, p.cache, gtid, gthreads);
      else be_scatter_range<2, false>(g, p.poses, nullptr, nullptr, p.ilq, p.cache, gtid, gthreads);
    } else {
      if (p.want_grad) be_scatter_range<0, true>(g, p.poses, p.il_old, p.il_new, nullptr, p.cache, gtid, gthreads);
      else be_scatter_range<0, false>(g, p.poses, p.il_old, p.il_new, nullptr, p.cache, gtid, gthreads);
    }
  }
  CMAXB_BE_MARK(2);
  grid.sync();
  // ---- blur(IL + alpha*IGp) + S1, S2 ----------------------------------------------------------------
  const TileCtx tc{g.W, g.H, p.th, p.taps};
  if (p.use_quad) {
    const MQuad src{p.ilq, p.igp, p.alpha};
    tile_blur_phase<R>(tc, src, p.blurred, p.ilq_next, p.part_img, smem_raw);
  } else {
    const MPlanes src{p.il_old, p.il_new, p.igp, p.alpha};
    tile_blur_phase<R>(tc, src, p.blurred, nullptr, p.part_img, smem_raw);
  }
  CMAXB_BE_MARK(3);
  grid.sync();
  double S1, S2;
  tile_sum_partials(p.part_img, s_red, &S1, &S2);
  const double mean = S1 / Np;
  const double contrast = contrast_from_sums(S1, S2, Np, p.measure);
  if (!p.want_grad || p.n_opt == 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      p.result[0] = contrast;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(p.done_flag) = p.seq;
    }
    CMAXB_BE_MARK(4);
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) p.contrast_dev[0] = contrast;
  // ---- adjoint image ---------------------------------------------------------------------------------
  {
    const float b2 = (p.measure == CMAXB_CONTRAST_MEAN_SQUARE) ? 0.0f : (float)(-2.0 * mean);
    tile_adjoint_phase<R>(tc, p.blurred, 2.0f, b2, p.use_quad ? nullptr : p.G, p.use_quad ? p.GQ : nullptr, smem_raw);
  }
  CMAXB_BE_MARK(4);
  grid.sync();
  // ---- gather: one warp per batch ------------------------------------------------------------------
  {
    const long long w0 = blockIdx.x * (long long)(kMegaThreads / 32) + (threadIdx.x >> 5);
    const long long ws = (long long)gridDim.x * (kMegaThreads / 32);
    if (p.use_quad) be_gather_range<N, true>(g, p.poses, nullptr, p.GQ, p.cache, p.wgrad, w0, ws);
    else be_gather_range<N, false>(g, p.poses, p.G, nullptr, p.cache, p.wgrad, w0, ws);
  }
  CMAXB_BE_MARK(5);
  grid.sync();
  // ---- per-knot sums (fixed order) ------------------------------------------------------------------
  for (int kk = blockIdx.x; kk < p.n_opt; kk += gridDim.x) {
    __syncthreads();
    be_grad_reduce_knot<N>(kk, p.idx, p.seg_lo, p.seg_hi, p.wgrad, g.n_fixed, 1.0 / Np, p.grad_dev, s_red);
  }
  CMAXB_BE_MARK(6);
  // last CTA to arrive publishes contrast + gradient to the host
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x == 0) { *p.ticket = 0u; p.result[0] = __ldcg(p.contrast_dev); }
    for (int i = threadIdx.x; i < 3 * p.n_opt; i += kMegaThreads) p.result[1 + i] = __ldcg(p.grad_dev + i);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (p.phase_ns) p.phase_ns[7] = global_timer_ns();
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(p.done_flag) = p.seq;
    }
  }
}

}  // namespace cmaxb
