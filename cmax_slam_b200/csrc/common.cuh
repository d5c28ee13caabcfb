// common.cuh -- shared device/host helpers of libcmax_b200 (sm_100a).
//
// Arithmetic contract: every geometry expression that decides an integer pixel cell is evaluated
// in IEEE f64 WITHOUT fused multiply-add (this translation unit set is compiled with
// -fmad=false), operation for operation as the reference does on baseline x86-64, so that the
// truncated cell index (local_image_warped_events.cpp:139, event_pano_warper.cpp:290) is
// bit-identical to the CPU.  Explicit fmaf() is used only where OpenCV's AVX2 filter uses FMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cmax_b200.h"

namespace cmaxb {

constexpr int kMaxRadius = 16;        // Gaussian kernel radius supported (sigma <= 4)
constexpr int kMaxTaps = 2 * kMaxRadius + 1;

struct Taps {
  int r;                // radius; 0 = no blur
  float w[kMaxTaps];    // w[0..2r]
};

// ---- ros::Time / ros::Duration arithmetic (roscpp rostime; see oracle/cmax_oracle.cpp) ----------
struct RosTime { uint32_t sec, nsec; };

__host__ __device__ inline double ros_to_sec(uint32_t sec, uint32_t nsec) {
  return (double)sec + 1e-9 * (double)nsec;
}
// time_first + (time_last - time_first) * 0.5   (local_image_warped_events.cpp:68-73,
// event_pano_warper.cpp:239-242).  Returns false when the span is negative.
__host__ __device__ inline bool ros_batch_mid(RosTime first, RosTime last, RosTime* mid) {
  long long s = (long long)last.sec - (long long)first.sec;
  long long ns = (long long)last.nsec - (long long)first.nsec;
  if (ns < 0) { ns += 1000000000ll; --s; }
  // Duration::toSec()
  const double dsec = (double)(int)s + 1e-9 * (double)(int)ns;
  const bool ok = dsec >= 0.0;
  // Duration * 0.5 -> Duration(toSec()*0.5) -> fromSec: floor / round-half-away
  const double d = dsec * 0.5;
  const double fl = floor(d);
  int hs = (int)(long long)fl;
  int hns = (int)round((d - (double)hs) * 1e9);
  hs += hns / 1000000000;
  hns %= 1000000000;
  long long rs = (long long)first.sec + hs;
  long long rns = (long long)first.nsec + hns;
  if (rns >= 1000000000ll) { rns -= 1000000000ll; ++rs; }
  if (rns < 0) { rns += 1000000000ll; --rs; }
  mid->sec = (uint32_t)rs;
  mid->nsec = (uint32_t)rns;
  return ok;
}

// ---- event record: one 16-byte vector load ------------------------------------------------------
// uint4 = { x | y<<16, sec, nsec, polarity | pad }
__device__ __forceinline__ uint4 load_event(const uint4* __restrict__ ev, long long i) {
  return __ldg(ev + i);
}

// ---- warp / block reductions of doubles ---------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduce NV doubles per thread across the block and atomically add them to dst[0..NV).
// red: shared scratch of at least (blockDim.x/32)*NV doubles.
template <int NV>
__device__ __forceinline__ void block_atomic_add(double (&v)[NV], double* dst, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[wid * NV + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int w = 0; w < nw; ++w) s += red[w * NV + threadIdx.x];
    atomicAdd(dst + threadIdx.x, s);
  }
}

// flags[0] |= 2 when an event lies outside the sensor (the reference's
// precomputed_bearing_vectors_.at(...) would throw, local_image_warped_events.cpp:100)
static __global__ void validate_events_kernel(const uint4* __restrict__ ev, long long n, int W, int H, int* flags) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 e = ev[i];
  if ((int)(e.x & 0xffff) >= W || (int)(e.x >> 16) >= H) atomicOr(flags, 2);
}

}  // namespace cmaxb
