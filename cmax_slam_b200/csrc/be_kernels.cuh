// be_kernels.cuh -- back-end (panoramic bundle adjustment) event kernels.
//
// Per cost evaluation the reference runs EventWarper::computeImageOfWarpedEvents ->
// warpAndAccumulateEvents (src/backend/event_pano_warper.cpp:167-336): one spline pose and knot
// Jacobian per batch of `event_batch_size` events (Trajectory::evaluate, src/backend/
// trajectory.cpp:86-110,329-355 -> basalt So3Spline<N>::evaluate), then per event a rotation of
// the bearing vector, the equirectangular projection with its 2x3 Jacobian
// (include/backend/equirectangular_camera.h:18-45), bilinear votes into IL_old_/IL_new_ and into
// 3*Nk derivative bands.
//
// Mapping here: one WARP per batch (the pose is loaded once and broadcast; no division by the
// batch size; the 16-byte event records of a batch are one coalesced run), lanes stride over the
// batch's events.
#pragma once
#include "common.cuh"
#include "so3_math.cuh"

namespace cmaxb {

struct BePose {            // per batch, 232 bytes
  double R[9];             // so3.matrix()                                   (:254)
  float Jk[36];            // 3 x 3Nk, row-major with row stride 3Nk         (trajectory.cpp:99-106,342-351)
  int idx_cp_beg;          // J.start_idx
  int valid;               // 0: batch not processed
};

struct BeBatchTime {       // per batch, fixed for a window
  long long s;             // segment index  (so3_spline.h:224)
  double u;                // fractional position (:225)
};

struct BeGeom {
  const uint4* ev;
  long long n_eff;         // events actually visited by the reference loop (:188-196)
  long long nb;
  int batch_size, sample_rate;
  const double4* lut;
  int SW, SH;              // sensor
  int W, H;                // panorama
  double fx, fy, cx, cy;   // equirectangular focal lengths / centre      (equirectangular_camera.h:11-16,64-67)
  uint32_t tnext_sec, tnext_nsec;
  int n_fixed, Nk;
  // enumeration of the events the reference loop visits (stride sample_rate inside each batch, :262):
  // visited event j lives in batch b = min(j / m, nb-1) at offset (j - b*m) * sample_rate
  int m;                   // visited events per full batch = ceil(batch_size / sample_rate)
  long long n_visit;       // total visited events
};

// per visited event, written by the gradient scatter pass and consumed by the gather pass: 24 bytes
// (round 1 stored the 2x3 Jacobian factor as well, 36 bytes: 2.6x the algorithmic traffic at C4; the factor is a
// closed form of the rotated ray, rebuilt in f32 by the gather)
struct BeCache {
  float4* a;               // (cell = yy*W+xx as int bits or -1, dx, dy, ray.x)
  float2* b;               // (ray.y, ray.z)   ray = R * bearing (world frame), f32
};

// t_mid of each batch -> (s, u); flags |= 4 when outside the spline (BASALT_ASSERT, so3_spline.h:221-230)
__global__ void be_batch_time_kernel(const uint4* __restrict__ ev, long long n, long long n_eff, int bs, long long nb,
                                     long long t0_ns, long long dt_ns, int n_knots, int order,
                                     BeBatchTime* __restrict__ out, int* flags) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const long long beg = b * bs;
  long long end = beg + bs;
  if (n - beg <= bs) end = n;                                             // (:192-194)
  const uint4 e0 = ev[beg], e1 = ev[end - 1];
  RosTime mid;
  ros_batch_mid(RosTime{e0.y, e0.z}, RosTime{e1.y, e1.z}, &mid);
  const long long t_ns = (long long)((unsigned long long)mid.sec * 1000000000ull + (unsigned long long)mid.nsec);  // toNSec()
  const long long st = t_ns - t0_ns;
  BeBatchTime bt;
  bt.s = 0; bt.u = 0.0;
  if (st < 0) { atomicOr(flags, 4); }
  else {
    bt.s = st / dt_ns;
    bt.u = (double)(st % dt_ns) / (double)dt_ns;
    if (bt.s + order > (long long)n_knots) { atomicOr(flags, 4); bt.s = 0; }
  }
  out[b] = bt;
}

// one batch: K_i <- exp(x_i) K_i for the N knots of its segment (incrementalUpdate, trajectory.cpp:221-238,
// 491-499), then So3Spline<N>::evaluate (+ f32 knot Jacobians).  knots0: the window's knots; x: 3*n_opt increments.
template <int N>
__device__ __forceinline__ void be_pose_one(const Quat* __restrict__ knots0, const double* __restrict__ x, int n_fixed,
                                            BeBatchTime t, int want_grad, BePose* __restrict__ p, int* __restrict__ idx_out) {
  Quat loc[N];
  const int s = (int)t.s;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    Quat q = knots0[s + i];
    if (s + i >= n_fixed) {
      const int j = s + i - n_fixed;
      Vec3 d; d.x = x[3 * j]; d.y = x[3 * j + 1]; d.z = x[3 * j + 2];
      q = quat_mul(so3_exp(d), q);
    }
    loc[i] = q;
  }
  Mat3 J[N];
  const Quat q = so3_spline_eval<N>(loc, 0, t.u, want_grad ? J : nullptr);
  const Mat3 R = quat_to_mat(q);
#pragma unroll
  for (int i = 0; i < 9; ++i) p->R[i] = R.m[i];
  if (want_grad) {
#pragma unroll
    for (int k = 0; k < N; ++k)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) p->Jk[r * (3 * N) + 3 * k + c] = (float)J[k].m[r * 3 + c];
  }
  p->idx_cp_beg = s;
  p->valid = 1;
  *idx_out = s;     // packed copy for the per-knot reduction (coalesced scan)
}

template <int N>
__global__ void be_pose_kernel(const Quat* __restrict__ knots0, const double* __restrict__ x, int n_fixed,
                               const BeBatchTime* __restrict__ bt, long long nb, int want_grad, BePose* __restrict__ poses,
                               int* __restrict__ idx_out, float4* __restrict__ zero_a = nullptr, long long zero_na = 0,
                               float4* __restrict__ zero_b = nullptr, long long zero_nb = 0) {
  // the evaluation's accumulators are cleared here as well (one launch instead of a memset node per buffer)
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = tid; i < zero_na; i += nth) zero_a[i] = z4;
  for (long long i = tid; i < zero_nb; i += nth) zero_b[i] = z4;
  const long long b = tid;
  if (b >= nb) return;
  be_pose_one<N>(knots0, x, n_fixed, bt[b], want_grad, poses + b, idx_out + b);
}

struct BeWarp {
  bool in;
  int xx, yy;
  float dx, dy;
  bool is_old;
  float rx, ry, rz;  // R * bearing as f32 (drb_ddrot is built from exactly these casts, :280-281)
  float dd[2][3];   // dpm_ddrot = dpm_drb * drb_ddrot (2x3, f32)            (:273-282)
};

template <bool GRAD>
__device__ __forceinline__ BeWarp be_warp(const BeGeom& g, const double* R, uint4 e) {
  BeWarp o;
  const int ex = e.x & 0xffff, ey = e.x >> 16;
  const double2* lp = reinterpret_cast<const double2*>(g.lut + (ey * g.SW + ex));
  const double2 bxy = __ldg(lp);
  const double bz = __ldg(reinterpret_cast<const double*>(lp + 1));
  const double bx = bxy.x, by = bxy.y;
  // e_ray_w = R * e_ray_cam                                               (:269)
  const double wx = R[0] * bx + R[1] * by + R[2] * bz;
  const double wy = R[3] * bx + R[4] * by + R[5] * bz;
  const double wz = R[6] * bx + R[7] * by + R[8] * bz;
  // projectToImage                                                        (equirectangular_camera.h:18-45)
  const double phi = atan2(wx, wz);
  const double n2 = wx * wx + wy * wy + wz * wz;
  const double rho = sqrt(n2);
  const double theta = asin(wy / rho);
  const double px = g.cx + phi * g.fx;
  const double py = g.cy + theta * g.fy;
  o.in = false;
  o.xx = o.yy = 0;
  o.dx = o.dy = 0.f;
  if (fabs(px) < 2e9 && fabs(py) < 2e9) {
    const int xx = (int)px, yy = (int)py;                                  // (:290-291)
    if (1 <= xx && xx < g.W - 2 && 1 <= yy && yy < g.H - 2) {              // (:296)
      o.in = true;
      o.xx = xx; o.yy = yy;
      o.dx = (float)(px - (double)xx);
      o.dy = (float)(py - (double)yy);
    }
  }
  o.is_old = (e.y < g.tnext_sec) || (e.y == g.tnext_sec && e.z < g.tnext_nsec);   // ev->ts < t_next_win_beg_ (:298)
  o.rx = (float)wx; o.ry = (float)wy; o.rz = (float)wz;
  if (GRAD) {
    const double Ydivrho = wy / rho;
    const double XdivZ = wx / wz;
    const double tmp1 = g.fx / ((1 + XdivZ * XdivZ) * wz);
    const double tmp2 = -g.fy / sqrt(1 - Ydivrho * Ydivrho);
    const double tmp3 = Ydivrho / (rho * rho);
    const float j00 = (float)tmp1, j02 = (float)(-tmp1 * XdivZ);
    const float j10 = (float)(tmp2 * tmp3 * wx), j11 = (float)(tmp2 * (tmp3 * wy - 1 / rho)), j12 = (float)(tmp2 * tmp3 * wz);
    // drb_ddrot = [0 rb.z -rb.y; -rb.z 0 rb.x; rb.y -rb.x 0] as f32          (:280-281)
    const float rx = (float)wx, ry = (float)wy, rz = (float)wz;
    const float nrx = (float)(-wx), nry = (float)(-wy), nrz = (float)(-wz);
    // dpm_ddrot = dpm_drb * drb_ddrot, f32, s = 0; s += a*b  (j01 == 0)      (:282)
    o.dd[0][0] = j00 * 0.f + 0.f * nrz + j02 * ry;
    o.dd[0][1] = j00 * rz + 0.f * 0.f + j02 * nrx;
    o.dd[0][2] = j00 * nry + 0.f * rx + j02 * 0.f;
    o.dd[1][0] = j10 * 0.f + j11 * nrz + j12 * ry;
    o.dd[1][1] = j10 * rz + j11 * 0.f + j12 * nrx;
    o.dd[1][2] = j10 * nry + j11 * rx + j12 * 0.f;
  }
  return o;
}

constexpr int kBeThreads = 256;
constexpr int kBeDirtyWords = 1024;     // dirty-tile bitmap of the peer exchange: up to 32768 tiles of 32 x 32 (8192 x 4096 panorama)
constexpr int kBeWarps = kBeThreads / 32;

// value scatter, one THREAD per visited event (full lane efficiency; the batch pose is read through
// L1, where the ~100 consecutive threads of a batch hit the same 72 bytes).
// MODE 0: IL_old_ / IL_new_ as two float planes (what updateIG needs).
// MODE 2: IL = old + new as ONE corner-split float4 image, one vector reduction per event (the cost
// evaluation only ever uses the sum, event_pano_warper.cpp:199).
// CACHE: also compute the 2x3 Jacobian factor and store (cell, dx, dy, dd) for the gather pass.
template <int MODE, bool CACHE>
__device__ __forceinline__ void be_scatter_range(const BeGeom& g, const BePose* __restrict__ poses, float* __restrict__ il_old,
                                                 float* __restrict__ il_new, float4* __restrict__ il_quad, const BeCache& cache,
                                                 long long j0, long long stride, int dirty_ntx = 0, unsigned int* s_dirty = nullptr) {
  for (long long j = j0; j < g.n_visit; j += stride) {
    long long b = j / g.m;
    if (b > g.nb - 1) b = g.nb - 1;
    const long long i = b * g.batch_size + (j - b * g.m) * g.sample_rate;
    if (i >= g.n_eff) continue;   // never true by construction of n_visit; kept as a guard
    double R[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = __ldcg(&poses[b].R[q]);
    const uint4 e = load_event(g.ev, i);
    const BeWarp w = be_warp<false>(g, R, e);
    if (CACHE) {
      __stcg(cache.a + j, make_float4(__int_as_float(w.in ? w.yy * g.W + w.xx : -1), w.dx, w.dy, w.rx));
      if (w.in) __stcg(cache.b + j, make_float2(w.ry, w.rz));
    }
    if (!w.in) continue;
    const float dx = w.dx, dy = w.dy;
    const long long p = (long long)w.yy * g.W + w.xx;
    if (MODE == 2) {
      atomicAdd(il_quad + p, make_float4((1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy));
      if (s_dirty) {                        // 32x32 panorama tile touched (peer exchange, be_xchg.cuh): CTA-local bitmap, flushed once
        const int t = (w.yy >> 5) * dirty_ntx + (w.xx >> 5);     // (millions of stores to a few global sectors would serialise in L2)
        const unsigned int bit = 1u << (t & 31);
        if (!(s_dirty[t >> 5] & bit)) atomicOr(&s_dirty[t >> 5], bit);
      }
    } else {
      float* il = w.is_old ? il_old : il_new;
      atomicAdd(il + p, (1.f - dx) * (1.f - dy));
      atomicAdd(il + p + 1, dx * (1.f - dy));
      atomicAdd(il + p + g.W, (1.f - dx) * dy);
      atomicAdd(il + p + g.W + 1, dx * dy);
    }
  }
}

template <int MODE, bool CACHE>
__global__ void __launch_bounds__(kBeThreads)
be_scatter_kernel(BeGeom g, const BePose* __restrict__ poses, float* __restrict__ il_old, float* __restrict__ il_new,
                  float4* __restrict__ il_quad, BeCache cache, unsigned int* __restrict__ dirty = nullptr, int dirty_ntx = 0,
                  int dirty_ntiles = 0) {
  __shared__ unsigned int s_dirty[kBeDirtyWords];
  const bool local = MODE == 2 && dirty != nullptr;           // dirty_ntiles <= 32 kBeDirtyWords (checked by cmaxb_be_exchange_init)
  const int nwords = (dirty_ntiles + 31) >> 5;
  if (local) {
    for (int i = threadIdx.x; i < nwords; i += kBeThreads) s_dirty[i] = 0u;
    __syncthreads();
  }
  be_scatter_range<MODE, CACHE>(g, poses, il_old, il_new, il_quad, cache, blockIdx.x * (long long)kBeThreads + threadIdx.x,
                                (long long)gridDim.x * kBeThreads, dirty_ntx, local ? s_dirty : nullptr);
  if (local) {
    __syncthreads();
    for (int i = threadIdx.x; i < nwords; i += kBeThreads) {
      const unsigned int m = s_dirty[i];
      if (m && (m & ~__ldcg(dirty + i))) atomicOr(dirty + i, m);
    }
  }
}

// dense derivative bands (reference-faithful DENSE mode / parity): planar bands[P][A]
template <int N>
__global__ void __launch_bounds__(kBeThreads)
be_scatter_bands_kernel(BeGeom g, const BePose* __restrict__ poses, float* __restrict__ bands, long long A, int P) {
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * kBeWarps;
  for (long long b = blockIdx.x * (long long)kBeWarps + (threadIdx.x >> 5); b < g.nb; b += wstride) {
    const long long beg = b * g.batch_size;
    long long end = beg + g.batch_size;
    if (end > g.n_eff || g.n_eff - beg <= g.batch_size) end = g.n_eff;
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = __ldg(&poses[b].R[i]);
    const int idx = poses[b].idx_cp_beg;
    const float* Jk = poses[b].Jk;
    for (long long i = beg + (long long)lane * g.sample_rate; i < end; i += 32LL * g.sample_rate) {
      const uint4 e = load_event(g.ev, i);
      const BeWarp w = be_warp<true>(g, R, e);
      if (!w.in) continue;
      const float dx = w.dx, dy = w.dy;
      const long long p = (long long)w.yy * g.W + w.xx;
      for (int c = 0; c < 3 * N; ++c) {
        const int j = 3 * (idx - g.n_fixed) + c;                            // (:322)
        if (j < 0 || j >= P) continue;
        // jac = dpm_ddrot * ddrot_ddrot_cp: cv::gemm CV_32F accumulates in double   (:285)
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double jk = (double)__ldg(Jk + k * (3 * N) + c);
          s0 += (double)w.dd[0][k] * jk;
          s1 += (double)w.dd[1][k] * jk;
        }
        const float r0 = (float)s0, r1 = (float)s1;
        float* bd = bands + (long long)j * A + p;
        atomicAdd(bd, r0 * (-(1.f - dy)) + r1 * (-(1.f - dx)));                // (:327-330)
        atomicAdd(bd + 1, r0 * (1.f - dy) + r1 * (-dx));
        atomicAdd(bd + g.W, r0 * (-dy) + r1 * (1.f - dx));
        atomicAdd(bd + g.W + 1, r0 * dy + r1 * dx);
      }
    }
  }
}

// debug / parity: per-event cell (-1 rejected, -2 never visited)
__global__ void be_cells_kernel(BeGeom g, const BePose* __restrict__ poses, long long n_total, int* __restrict__ cells) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  int out = -2;
  if (i < g.n_eff) {
    const long long b = i / g.batch_size;
    if ((i - b * g.batch_size) % g.sample_rate == 0) {
      double R[9];
      for (int k = 0; k < 9; ++k) R[k] = poses[b].R[k];
      const BeWarp w = be_warp<false>(g, R, load_event(g.ev, i));
      out = w.in ? w.yy * g.W + w.xx : -1;
    }
  }
  cells[i] = out;
}

// Adjoint gather.  For an event with derivative-vote weights (s_c, t_c) (:327-330) and
// jac = dd * Jk, the contribution to g_j is  (a*dd[0,:] + b*dd[1,:]) . Jk[:, j]  with
// a = sum_c s_c G(c), b = sum_c t_c G(c).  The bracket is summed over the batch first (3 numbers),
// then multiplied by the batch's Jk once: wgrad[b][c] = V_b . Jk[:, c].
// One WARP per batch; the f64 geometry (atan2 / asin) is NOT recomputed: (cell, dx, dy, rotated ray) come from the
// cache the gradient scatter pass wrote (24 bytes per event, streamed once), and the 2x3 factor
// dpm_ddrot = dpm_drb * drb_ddrot (event_pano_warper.cpp:273-282, equirectangular_camera.h:30-42) is rebuilt from the
// ray in f32 through its closed form (s2 = x^2 + z^2, rho2 = s2 + y^2):
//   dpm_drb = [ fx z / s2, 0, -fx x / s2 ;  -fy x y / (rho2 s),  fy s / rho2,  -fy y z / (rho2 s) ]
template <int N, bool QUAD>
__device__ __forceinline__ void be_gather_range(const BeGeom& g, const BePose* __restrict__ poses, const float* __restrict__ G,
                                                const float4* __restrict__ GQ, const BeCache& cache, double* __restrict__ wgrad,
                                                long long b0, long long wstride) {
  const int lane = threadIdx.x & 31;
  for (long long b = b0; b < g.nb; b += wstride) {
    const long long j0 = b * g.m;
    long long j1 = j0 + g.m;
    if (b == g.nb - 1 || j1 > g.n_visit) j1 = g.n_visit;
    double v0 = 0, v1 = 0, v2 = 0;
    // kGatherIl x 32 events per pass: the cache records of both first, then both adjoint-image cells, then the arithmetic
    // (the kernel waits on loads: 12 stalled warps per issue in ncu; 4 x 32 cost too many registers, profiles/r02k_*)
    constexpr int kGatherIl = 2;
    for (long long jb = j0; jb < j1; jb += 32 * kGatherIl) {
      float4 ca4[kGatherIl]; float2 cb2[kGatherIl]; float4 q4[kGatherIl]; int cell[kGatherIl];
#pragma unroll
      for (int k = 0; k < kGatherIl; ++k) {
        const long long j = jb + k * 32 + lane;
        ca4[k] = make_float4(__int_as_float(-1), 0.f, 0.f, 0.f);
        cb2[k] = make_float2(0.f, 0.f);
        if (j < j1) { ca4[k] = __ldcg(cache.a + j); cb2[k] = __ldcg(cache.b + j); }   // (b is only meaningful when the cell is valid)
      }
#pragma unroll
      for (int k = 0; k < kGatherIl; ++k) {
        cell[k] = __float_as_int(ca4[k].x);
        q4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cell[k] >= 0) {
          if (QUAD) q4[k] = __ldcg(GQ + cell[k]);
          else {
            const float* p = G + cell[k];
            q4[k] = make_float4(__ldcg(p), __ldcg(p + 1), __ldcg(p + g.W), __ldcg(p + g.W + 1));
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kGatherIl; ++k) {
        if (cell[k] < 0) continue;
        float dd[2][3];
        {
          const float rx = ca4[k].w, ry = cb2[k].x, rz = cb2[k].y;
          const float s2 = rx * rx + rz * rz, rho2 = s2 + ry * ry;
          const float sr = sqrtf(s2);
          const float i_s2 = __fdividef(1.0f, s2), i_r2s = __fdividef(1.0f, rho2 * sr);
          const float ffx = (float)g.fx, ffy = (float)g.fy;
          const float j00 = ffx * rz * i_s2, j02 = -ffx * rx * i_s2;
          const float j10 = -ffy * rx * ry * i_r2s, j11 = ffy * sr * __fdividef(1.0f, rho2), j12 = -ffy * ry * rz * i_r2s;
          // dpm_ddrot = dpm_drb * (-[ray]x)
          dd[0][0] = j02 * ry;               dd[0][1] = j00 * rz - j02 * rx;   dd[0][2] = -j00 * ry;
          dd[1][0] = -j11 * rz + j12 * ry;   dd[1][1] = j10 * rz - j12 * rx;   dd[1][2] = -j10 * ry + j11 * rx;
        }
        const double g00 = q4[k].x, g01 = q4[k].y, g10 = q4[k].z, g11 = q4[k].w;
        const double dx = ca4[k].y, dy = ca4[k].z;
        const double a = (1.0 - dy) * (g01 - g00) + dy * (g11 - g10);
        const double bb = (1.0 - dx) * (g10 - g00) + dx * (g11 - g01);
        v0 += a * (double)dd[0][0] + bb * (double)dd[1][0];
        v1 += a * (double)dd[0][1] + bb * (double)dd[1][1];
        v2 += a * (double)dd[0][2] + bb * (double)dd[1][2];
      }
    }
    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
    if (lane < 3 * N) {
      const float* Jk = poses[b].Jk;
      wgrad[b * (3 * N) + lane] = v0 * (double)__ldcg(Jk + lane) + v1 * (double)__ldcg(Jk + 3 * N + lane) + v2 * (double)__ldcg(Jk + 6 * N + lane);
    }
  }
}

template <int N, bool QUAD>
__global__ void __launch_bounds__(kBeThreads)
be_gather_kernel(BeGeom g, const BePose* __restrict__ poses, const float* __restrict__ G, const float4* __restrict__ GQ,
                 BeCache cache, double* __restrict__ wgrad) {
  be_gather_range<N, QUAD>(g, poses, G, GQ, cache, wgrad, blockIdx.x * (long long)kBeWarps + (threadIdx.x >> 5),
                           (long long)gridDim.x * kBeWarps);
}

// Batches of spline segment s form a contiguous run when the events are time-sorted; record the
// bounding run [lo, hi) of every segment once per window (correct for unsorted input too: the
// reduction re-checks idx inside the run).
__global__ void be_segment_ranges_kernel(const BeBatchTime* __restrict__ bt, long long nb, int* __restrict__ seg_lo,
                                         int* __restrict__ seg_hi) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int s = (int)bt[b].s;
  atomicMin(seg_lo + s, (int)b);
  atomicMax(seg_hi + s, (int)b + 1);
}

// 1024 threads per knot: the loop below is a chain of dependent L2 round trips (index, then three partial sums per
// batch), so its length in iterations is what the kernel costs (2 000 batches per knot in C5: 8 iterations at 256 threads)
constexpr int kBeReduceThreads = 1024;

// g[3*kk + c] = (1/Np) * sum over batches touching knot (kk + n_fixed) of wgrad[b][3*(knot-idx_b)+c].
// One CTA per optimised knot; fixed summation order (deterministic).
template <int N>
__device__ __forceinline__ void be_grad_reduce_knot(int kk, const int* __restrict__ idx, const int* __restrict__ seg_lo,
                                                    const int* __restrict__ seg_hi, const double* __restrict__ wgrad, int n_fixed,
                                                    double inv_np, double* __restrict__ grad, double* s_red /*[warps*3]*/,
                                                    double* __restrict__ host_grad = nullptr /* mapped host copy of the result */) {
  const int knot = kk + n_fixed;
  double a[3] = {0.0, 0.0, 0.0};
  for (int rel = 0; rel < N; ++rel) {
    const int s = knot - rel;               // batches of segment s touch knots s .. s+N-1
    if (s < 0) continue;
    const int lo = seg_lo[s], hi = seg_hi[s];
    for (long long b = lo + (long long)threadIdx.x; b < hi; b += blockDim.x) {
      if (__ldcg(idx + b) != s) continue;
      const double* w = wgrad + b * (3 * N) + 3 * rel;
      a[0] += __ldcg(w); a[1] += __ldcg(w + 1); a[2] += __ldcg(w + 2);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) a[c] = warp_sum(a[c]);
  __syncthreads();
  if (lane == 0) { s_red[wid * 3] = a[0]; s_red[wid * 3 + 1] = a[1]; s_red[wid * 3 + 2] = a[2]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_red[w * 3 + threadIdx.x];
    grad[3 * kk + threadIdx.x] = s * inv_np;
    if (host_grad) host_grad[3 * kk + threadIdx.x] = s * inv_np;
  }
}

template <int N>
__global__ void __launch_bounds__(kBeReduceThreads)
be_grad_reduce_kernel(const int* __restrict__ idx, const int* __restrict__ seg_lo, const int* __restrict__ seg_hi,
                      const double* __restrict__ wgrad, long long nb, int n_fixed, double inv_np, double* __restrict__ grad,
                      double* __restrict__ host_grad = nullptr) {
  __shared__ double s_red[(kBeReduceThreads / 32) * 3];
  be_grad_reduce_knot<N>(blockIdx.x, idx, seg_lo, seg_hi, wgrad, n_fixed, inv_np, grad, s_red, host_grad);
}

// DENSE mode: reduce blurred bands against the blurred image.
// g_j = mean( 2(I-mu) .* (D_j - mean(D_j)) ) = 2*(mean(I*D_j) - mu*mean(D_j))   (global_focus_funcs.cpp:39-43)
// or 2*mean(I*D_j) (mean square, :22).  One CTA per band.
__global__ void __launch_bounds__(256)
be_band_reduce_kernel(const float* __restrict__ I, const float* __restrict__ bands, long long A, const double* mean,
                      int measure, double* __restrict__ grad) {
  __shared__ double s_red[8 * 2];
  const float* D = bands + (long long)blockIdx.x * A;
  double sd = 0, sid = 0;
  for (long long i = threadIdx.x; i < A; i += blockDim.x) {
    const double d = D[i];
    sd += d; sid += (double)I[i] * d;
  }
  sd = warp_sum(sd); sid = warp_sum(sid);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_red[wid * 2] = sd; s_red[wid * 2 + 1] = sid; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += s_red[w * 2]; b += s_red[w * 2 + 1]; }
    const double Np = (double)A;
    grad[blockIdx.x] = (measure == CMAXB_CONTRAST_MEAN_SQUARE) ? 2.0 * (b / Np) : 2.0 * (b / Np - mean[0] * (a / Np));
  }
}

// ---- global-map upkeep (event_pano_warper.cpp:81-132) ---------------------------
// EventWarper::setUpdateTimesIG(rot, radius): rasterise the sensor's field of view at pose `rot` into a 0/1 mask
// (every sensor pixel warped with warpEventToMap, dilated by `radius`, including the reference's
// `0 <= y_mask + j` bound), one thread per sensor pixel.  Stores of 1 are idempotent: no atomics.
__global__ void __launch_bounds__(256)
be_fov_mask_kernel(BeGeom g, Quat rot, int radius, unsigned char* __restrict__ mask) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)g.SW * g.SH) return;
  const Mat3 Rm = quat_to_mat(rot);
  const double2* lp = reinterpret_cast<const double2*>(g.lut + i);
  const double2 bxy = __ldg(lp);
  const double bz = __ldg(reinterpret_cast<const double*>(lp + 1));
  const double wx = Rm.m[0] * bxy.x + Rm.m[1] * bxy.y + Rm.m[2] * bz;
  const double wy = Rm.m[3] * bxy.x + Rm.m[4] * bxy.y + Rm.m[5] * bz;
  const double wz = Rm.m[6] * bxy.x + Rm.m[7] * bxy.y + Rm.m[8] * bz;
  const double phi = atan2(wx, wz);
  const double theta = asin(wy / sqrt(wx * wx + wy * wy + wz * wz));
  const double px = g.cx + phi * g.fx, py = g.cy + theta * g.fy;
  if (!(fabs(px) < 2e9 && fabs(py) < 2e9)) return;
  const int ic = (int)px, ir = (int)py;                                   // (:91)
  for (int a = -radius; a <= radius; ++a)
    for (int j = -radius; j <= radius; ++j) {
      const int x_mask = ic + a, y_mask = ir + j;
      if (0 <= y_mask + j && y_mask >= 0 && y_mask < g.H && 0 <= x_mask && x_mask < g.W)   // (:97; y_mask >= 0 guards the store)
        mask[(long long)y_mask * g.W + x_mask] = 1;
    }
}
// cv::add(IG_update_times_map_, mask, IG_update_times_map_) on CV_8U saturates            (:106)
__global__ void be_times_add_kernel(unsigned char* __restrict__ times, const unsigned char* __restrict__ mask, long long A) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)times[i] + (int)mask[i];
    times[i] = (unsigned char)(v > 255 ? 255 : v);
  }
}
// EventWarper::updateIG: IG += IL_old where the pixel has been visited at most max_update_times times   (:109-126)
__global__ void be_update_ig_kernel(float* __restrict__ IG, const float* __restrict__ il_old, const unsigned char* __restrict__ times,
                                    int max_update_times, long long A) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A; i += (long long)gridDim.x * blockDim.x)
    if ((int)times[i] <= max_update_times) IG[i] += il_old[i];
}

// IL as one float plane (from the corner-split accumulator or from IL_old + IL_new): the buffer that is
// summed across GPUs when one window is sharded by time.
__global__ void __launch_bounds__(256)
be_assemble_il_kernel(const float* __restrict__ il_old, const float* __restrict__ il_new, const float4* __restrict__ il_quad,
                      int W, long long A, float* __restrict__ plane) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A; i += (long long)gridDim.x * blockDim.x) {
    float l;
    if (il_quad) {
      const int y = (int)(i / W), x = (int)(i - (long long)y * W);
      l = il_quad[i].x;
      if (x > 0) l += il_quad[i - 1].y;
      if (y > 0) { l += il_quad[i - W].z; if (x > 0) l += il_quad[i - W - 1].w; }
    } else {
      l = il_old[i] + il_new[i];
    }
    plane[i] = l;
  }
}

// Row-band sharding of the image phases (one window sharded by time over several GPUs): the IL of THIS
// rank's events is packed as `world` chunks of ce = hb + 2 hl rows -- chunk c = rows [c hb - hl, (c+1) hb + hl) of the
// panorama, zero outside it -- so that ONE reduce-scatter hands every rank the summed rows of its own band INCLUDING the
// halo the blur and the adjoint blur need (hl = 2 r + 1).
__global__ void __launch_bounds__(256)
be_pack_bands_kernel(const float* __restrict__ il_old, const float* __restrict__ il_new, const float4* __restrict__ il_quad,
                     int W, int H, int world, int hb, int hl, float* __restrict__ send) {
  const int ce = hb + 2 * hl;
  const long long total = (long long)world * ce * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long row = i / W;
    const int c = (int)(row / ce), j = (int)(row - (long long)c * ce);
    const int y = c * hb - hl + j;
    float l = 0.f;
    if (y >= 0 && y < H) {
      const long long p = (long long)y * W + x;
      if (il_quad) {
        l = il_quad[p].x;
        if (x > 0) l += il_quad[p - 1].y;
        if (y > 0) { l += il_quad[p - W].z; if (x > 0) l += il_quad[p - W - 1].w; }
      } else {
        l = il_old[p] + il_new[p];
      }
    }
    send[i] = l;
  }
}

// contrast and mean from the (all-reduced) sums of the whole panorama
__global__ void be_band_finalize_kernel(const double* __restrict__ sums2, double Np, int measure, double* __restrict__ result,
                                        double* __restrict__ mean) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double S1 = sums2[0], S2 = sums2[1];
  const double m = S1 / Np;
  double contrast;
  if (measure == CMAXB_CONTRAST_MEAN_SQUARE) contrast = S2 / Np;
  else {
    double var = S2 / Np - m * m;
    if (var < 0.0) var = 0.0;
    const double sd = sqrt(var);
    contrast = sd * sd;
  }
  result[0] = contrast;
  mean[0] = m;
}

// updateAlpha sums (event_pano_warper.cpp:134-165): out[0..4] = sum(1-exp(-IGp)), sum(IGp),
// sum(1-exp(-IL)), sum(IL), countNonZero(IGp); IL = IL_old + IL_new.
// il_quad != nullptr: IL is re-assembled from the corner-split accumulator.
__global__ void __launch_bounds__(256)
be_alpha_sums_kernel(const float* __restrict__ igp, const float* __restrict__ il_old, const float* __restrict__ il_new,
                     const float4* __restrict__ il_quad, int W, long long A, double* __restrict__ out) {
  __shared__ double s_red[8 * 5];
  double v[5] = {0, 0, 0, 0, 0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A; i += (long long)gridDim.x * blockDim.x) {
    const float a = igp[i];
    float l;
    if (il_quad) {
      const int y = (int)(i / W), x = (int)(i - (long long)y * W);
      l = il_quad[i].x;
      if (x > 0) l += il_quad[i - 1].y;
      if (y > 0) { l += il_quad[i - W].z; if (x > 0) l += il_quad[i - W - 1].w; }
    } else {
      l = il_new ? il_old[i] + il_new[i] : il_old[i];
    }
    v[0] += (double)(1.f - expf(-1.0f * a));
    v[1] += (double)a;
    v[2] += (double)(1.f - expf(-1.0f * l));
    v[3] += (double)l;
    v[4] += (a != 0.f) ? 1.0 : 0.0;
  }
  block_atomic_add<5>(v, out, s_red);
}

}  // namespace cmaxb
