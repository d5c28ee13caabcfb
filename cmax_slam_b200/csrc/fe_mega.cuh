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
#include "tile_phases.cuh"

namespace cmaxb {

namespace cg = cooperative_groups;

constexpr int kMegaMaxHyp = 32;       // hypotheses per launch (kernel-parameter space)

// ---- fused result exchange ---------------------------------------------------------------------------------
// Multi-GPU hypothesis sharding: every rank evaluates its own hypotheses of the replicated
// packet and all ranks need all (contrast, g) rows.  Instead of a separate NCCL all-gather after the kernel, the
// CTA that publishes the result also stores its rows straight into every peer's exchange buffer (peer-to-peer
// stores over NVLink / NVSwitch, buffers opened with CUDA IPC), waits for the peers' rows and copies the
// gathered rows to mapped host memory -- compute + collective in ONE launch.
// Wire format ("low latency": data and flag travel together, no fence, no separate flag round trip): every
// double is sent as ONE 16-byte store {value bits, sequence number}; the receiver spins on the slot until the
// tag equals the sequence number of the launch.  Exchange buffer of a rank: slots[2][world][kmax][4] x 16 B,
// double-buffered by the parity of the sequence number (a rank can run at most one evaluation ahead of a peer
// that has not yet read its rows, because its own next kernel waits for that peer's rows of the same number).
constexpr int kXMaxWorld = 8;
constexpr unsigned long long kXTimeoutNs = 10ull * 1000ull * 1000ull * 1000ull;   // a dead peer must not hang the GPU

struct FeXchgParams {
  int world, rank, kmax;
  unsigned long long seq;           // exchange sequence number of this launch (same on all ranks), >= 1
  ulonglong2* peer[kXMaxWorld];     // exchange buffer of every rank as mapped into THIS process (peer[rank] = own)
  double* all_host;                 // mapped host memory [world][k][4]: gathered rows of this launch
  double* all_dev;                  // optional device copy [world][k][4] (caller owned)
  unsigned int* err;                // mapped host word, set to 1 when a peer's rows do not arrive in time
};

__device__ __forceinline__ ulonglong2* xchg_slots(const FeXchgParams& x, ulonglong2* base, int par, int r) {
  return base + ((long long)(par * x.world + r) * x.kmax) * 4;
}

// Called by ALL threads of ONE CTA.  s_rows[k*4] (shared memory) = this rank's rows of the launch.
__device__ __forceinline__ void mega_exchange(const FeXchgParams& x, int k, const double* s_rows) {
  const int par = (int)(x.seq & 1ull);
  const int nv = k * 4;
  __syncthreads();
  // 1. own rows -> every rank's buffer (own copy included), one tagged 16-byte store per value
  for (int i = threadIdx.x; i < x.world * nv; i += blockDim.x) {
    const int r = i / nv, j = i - r * nv;
    ulonglong2* dst = xchg_slots(x, x.peer[r], par, x.rank) + j;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(s_rows[j]);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(bits), "l"(x.seq) : "memory");
  }
  // 2. every rank's rows out of OUR buffer -> mapped host memory (+ device copy); spin until the tag arrives
  for (int i = threadIdx.x; i < x.world * nv; i += blockDim.x) {
    const int r = i / nv, j = i - r * nv;
    const ulonglong2* src = xchg_slots(x, x.peer[x.rank], par, r) + j;
    const unsigned long long t0 = global_timer_ns();
    unsigned long long bits = 0, tag = 0;
    unsigned int spins = 0;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tag) : "l"(src) : "memory");
      if (tag == x.seq) break;
      if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > kXTimeoutNs) { *x.err = 1u; break; }
    }
    const double v = __longlong_as_double((long long)bits);
    x.all_host[i] = v;
    if (x.all_dev) x.all_dev[i] = v;
  }
  __syncthreads();
}

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
  int gather_f32;             // 1: Jacobian rows and bilinear adjoint of the gather pass in f32 (f64 accumulation)
  FeXchgParams x;             // in-kernel all-gather of the result rows over peer memory (x.world <= 1: off)
};

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
        const FeWarp w = fe_warp_b<0>(g, bxy[u].x, bxy[u].y, bz[u], dt[u], ox, oy, oz);
        if (ok[u] && w.in) {
          const float dx = w.dx, dy = w.dy;
          atomicAdd(q + (long long)w.yy * g.W + w.xx,
                    make_float4((1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy));
        }
      }
    }
  }
}

// image phases: thin wrappers over tile_phases.cuh
template <int R>
__device__ __forceinline__ void mega_blur(const FeMegaParams& p, int h, unsigned char* smem_raw, bool write_out) {
  const TileCtx c{p.g.W, p.g.H, p.th, p.taps};
  const MQuad src{p.quad + h * p.A, nullptr, 0.f};
  tile_blur_phase<R>(c, src, write_out ? p.blurred + h * p.A : nullptr, p.quad_next ? p.quad_next + h * p.A : nullptr,
                     p.part_img + (long long)h * kMegaMaxCtas * 2, smem_raw);
}
__device__ __forceinline__ void mega_image_sums(const FeMegaParams& p, int h, double* s_red, double* S1, double* S2) {
  tile_sum_partials(p.part_img + (long long)h * kMegaMaxCtas * 2, s_red, S1, S2);
}
template <int R>
__device__ __forceinline__ void mega_adjoint(const FeMegaParams& p, int h, double mean, unsigned char* smem_raw) {
  const TileCtx c{p.g.W, p.g.H, p.th, p.taps};
  const float b2 = (p.measure == CMAXB_CONTRAST_MEAN_SQUARE) ? 0.0f : (float)(-2.0 * mean);
  tile_adjoint_phase<R>(c, p.blurred + h * p.A, 2.0f, b2, nullptr, p.GQ + h * p.A, smem_raw);
}

template <bool F32>
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
      w[u] = fe_warp_b<F32 ? 2 : 1>(g, bxy[u].x, bxy[u].y, bz[u], dt[u], ox, oy, oz);
      q[u] = (ok[u] && w[u].in) ? __ldcg(GQh + (long long)w[u].yy * g.W + w[u].xx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!(ok[u] && w[u].in)) continue;
      if (F32) {
        const float dx = w[u].dx, dy = w[u].dy;
        const float a = (1.f - dy) * (q[u].y - q[u].x) + dy * (q[u].w - q[u].z);
        const float b = (1.f - dx) * (q[u].z - q[u].x) + dx * (q[u].w - q[u].y);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += (double)(w[u].r0[c] * a + w[u].r1[c] * b);
        continue;
      }
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

// rows of this launch (shared memory, [k][4]) -> mapped host result (+ device mirror) (+ exchange with the
// peers), then the completion word the host spins on.  Called by all threads of ONE CTA.
__device__ __forceinline__ void mega_publish(const FeMegaParams& p, const double* s_rows) {
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * p.k; i += kMegaThreads) {
    const double v = s_rows[i];
    p.result[i] = v;
    if (p.mirror) p.mirror[i] = v;
  }
  if (p.x.world > 1) mega_exchange(p.x, p.k, s_rows);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(p.done_flag) = p.seq;
  }
}

template <int R>
__global__ void __launch_bounds__(kMegaThreads)
fe_eval_megakernel(const __grid_constant__ FeMegaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[(kMegaThreads / 32) * 3];
  __shared__ double s_mean[kMegaMaxHyp];
  __shared__ double s_rows[kMegaMaxHyp * 4];
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
    if (threadIdx.x == 0) {
      s_mean[h] = mean;
      double contrast;
      if (p.measure == CMAXB_CONTRAST_MEAN_SQUARE) contrast = S2 / Np;
      else {
        double var = S2 / Np - mean * mean;
        if (var < 0.0) var = 0.0;
        const double sd = sqrt(var);
        contrast = sd * sd;
      }
      // every CTA holds the identical value (fixed-order sums); the publishing CTA uses its own copy
      s_rows[4 * h] = contrast; s_rows[4 * h + 1] = 0.0; s_rows[4 * h + 2] = 0.0; s_rows[4 * h + 3] = 0.0;
    }
  }
  __syncthreads();
  if (!p.want_grad) {
    CMAXB_PHASE_MARK(5);
    if (blockIdx.x == 0) mega_publish(p, s_rows);
    return;
  }
  for (int h = 0; h < p.k; ++h) mega_adjoint<R>(p, h, s_mean[h], smem_raw);
  CMAXB_PHASE_MARK(5);
  grid.sync();
  CMAXB_PHASE_MARK(6);
  for (int h = 0; h < p.k; ++h) {
    __syncthreads();
    if (p.gather_f32) mega_gather<true>(p, h, s_red);
    else mega_gather<false>(p, h, s_red);
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
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (p.k <= 4) {
      // few hypotheses: the whole CTA adds the per-CTA records of one hypothesis (all loads of a thread in
      // flight at once: one L2 round trip), fixed order
      constexpr int kRec = (kMegaMaxCtas + kMegaThreads - 1) / kMegaThreads;
      for (int h = 0; h < p.k; ++h) {
        const double* all = p.part_ev + (long long)h * kMegaMaxCtas * 3;
        double v[kRec][3];
#pragma unroll
        for (int u = 0; u < kRec; ++u) {
          const int c = threadIdx.x + u * kMegaThreads;
          const bool ok = c < (int)gridDim.x;
          v[u][0] = ok ? __ldcg(all + 3 * c) : 0.0; v[u][1] = ok ? __ldcg(all + 3 * c + 1) : 0.0; v[u][2] = ok ? __ldcg(all + 3 * c + 2) : 0.0;
        }
        double t[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int u = 0; u < kRec; ++u) { t[0] += v[u][0]; t[1] += v[u][1]; t[2] += v[u][2]; }
        __syncthreads();
        block_sum<3>(t, s_red);
        if (threadIdx.x == 0) { s_rows[4 * h + 1] = t[0] / Np; s_rows[4 * h + 2] = t[1] / Np; s_rows[4 * h + 3] = t[2] / Np; }
      }
    } else {
      // one WARP per hypothesis: lanes add the per-CTA records in a fixed order
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
        if (lane == 0) { s_rows[4 * h + 1] = t0 / Np; s_rows[4 * h + 2] = t1 / Np; s_rows[4 * h + 3] = t2 / Np; }
      }
    }
    __syncthreads();
    CMAXB_PHASE_MARK_ANY(9);
    mega_publish(p, s_rows);
  }
}

inline size_t mega_smem_bytes(int r, int th = kMegaMaxTH) { return tile_smem_bytes(r, th); }
inline int mega_tile_height(int W, int H, int grid) { return tile_height_for_grid(W, H, grid); }

}  // namespace cmaxb
