// fe_binning.cuh -- one-time per-packet spatial binning of the events.
//
// The evaluation kernels are bound by the L2 rate of RANDOM 32-byte sector requests: the events of a
// packet arrive in time order, so consecutive threads touch unrelated pixels -- every bearing-vector
// read, every vote and every adjoint-image read is its own L2 request.  A packet is evaluated ~100-300
// times by the optimiser, so it pays to reorder it ONCE by 32x32 source tile: a warp's events then share
// a few LUT / accumulator / adjoint-image lines, and the fused kernel can stage the tile's bearing vectors in
// shared memory.  The binned record is 16 bytes {x | y<<16, batch index, dt (f64)}: dt is the reference's ONE time
// offset per 100 consecutive events (local_image_warped_events.cpp:67-76) of the event's batch in ARRIVAL order,
// carried along so that the evaluation kernels have no dependent table lookup left.
// Order inside a tile is arbitrary (atomic cursors); sums are reordered anyway by the f32 atomics.
// After pass 3, tile_cursor[t] = end offset of tile t's run (= start of tile t+1's).
#pragma once
#include "common.cuh"

namespace cmaxb {

constexpr int kBinTile = 32;          // source tile edge in pixels
constexpr int kBinThreads = 256;
constexpr int kBinChunk = 2048;       // events per CTA (8 per thread: ~490 CTAs for a 1M-event packet)
constexpr int kBinMaxTiles = 6144;    // shared-memory histogram capacity: the scatter pass needs 2 x 4 B per tile within the 48 KB default limit

__device__ __forceinline__ int bin_tile_of(uint4 e, int W, int H, int ntx) {
  int x = e.x & 0xffff, y = e.x >> 16;
  x = min(x, W - 1); y = min(y, H - 1);   // out-of-sensor events are rejected by validate_events_kernel
  return (y / kBinTile) * ntx + (x / kBinTile);
}

// pass 1: tile histogram (per-CTA shared histogram, one global atomic per non-empty bin) + validation of the pixel range
// (flags[0] |= 2: the reference's precomputed_bearing_vectors_.at(...) would throw, local_image_warped_events.cpp:100)
__global__ void __launch_bounds__(kBinThreads)
fe_bin_count_kernel(const uint4* __restrict__ ev, long long n, int W, int H, int ntx, int ntiles, unsigned int* __restrict__ tile_count,
                    int* __restrict__ flags) {
  extern __shared__ unsigned int s_hist[];
  for (int i = threadIdx.x; i < ntiles; i += kBinThreads) s_hist[i] = 0u;
  __syncthreads();
  const long long beg = blockIdx.x * (long long)kBinChunk;
  const long long end = min(beg + (long long)kBinChunk, n);
  bool bad = false;
  for (long long i = beg + threadIdx.x; i < end; i += kBinThreads) {
    const uint4 e = __ldg(ev + i);
    bad = bad || (int)(e.x & 0xffff) >= W || (int)(e.x >> 16) >= H;
    atomicAdd(&s_hist[bin_tile_of(e, W, H, ntx)], 1u);
  }
  if (bad) atomicOr(flags, 2);
  __syncthreads();
  for (int i = threadIdx.x; i < ntiles; i += kBinThreads)
    if (s_hist[i]) atomicAdd(&tile_count[i], s_hist[i]);
}

// pass 2: exclusive scan of the tile counts (one CTA; ntiles <= kBinMaxTiles) -> cursors
__global__ void __launch_bounds__(1024)
fe_bin_scan_kernel(const unsigned int* __restrict__ tile_count, int ntiles, unsigned int* __restrict__ tile_cursor) {
  __shared__ unsigned int s_part[1024];
  const int per = (ntiles + 1023) / 1024;
  const int b = threadIdx.x * per;
  unsigned int sum = 0;
  for (int i = b; i < min(b + per, ntiles); ++i) sum += tile_count[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan of the 1024 partials
  for (int off = 1; off < 1024; off <<= 1) {
    unsigned int v = (threadIdx.x >= off) ? s_part[threadIdx.x - off] : 0u;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned int run = s_part[threadIdx.x] - sum;   // exclusive prefix of this thread's range
  for (int i = b; i < min(b + per, ntiles); ++i) { tile_cursor[i] = run; run += tile_count[i]; }
}

// pass 3: write the 8-byte records tile by tile
__global__ void __launch_bounds__(kBinThreads)
fe_bin_scatter_kernel(const uint4* __restrict__ ev, long long n, int W, int H, int ntx, int ntiles, int batch_size,
                      const double* __restrict__ dt_tab, unsigned int* __restrict__ tile_cursor, uint4* __restrict__ binned) {
  extern __shared__ unsigned int s_mem[];
  unsigned int* s_hist = s_mem;            // counts, then running local cursor
  unsigned int* s_base = s_mem + ntiles;   // global base of this CTA's run inside each tile
  for (int i = threadIdx.x; i < ntiles; i += kBinThreads) s_hist[i] = 0u;
  __syncthreads();
  const long long beg = blockIdx.x * (long long)kBinChunk;
  const long long end = min(beg + (long long)kBinChunk, n);
  for (long long i = beg + threadIdx.x; i < end; i += kBinThreads) atomicAdd(&s_hist[bin_tile_of(__ldg(ev + i), W, H, ntx)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < ntiles; i += kBinThreads) {
    const unsigned int c = s_hist[i];
    s_base[i] = c ? atomicAdd(&tile_cursor[i], c) : 0u;
    s_hist[i] = 0u;
  }
  __syncthreads();
  for (long long i = beg + threadIdx.x; i < end; i += kBinThreads) {
    const uint4 e = __ldg(ev + i);
    const int t = bin_tile_of(e, W, H, ntx);
    const unsigned int pos = s_base[t] + atomicAdd(&s_hist[t], 1u);
    const unsigned int b = (unsigned int)(i / batch_size);
    const double dt = __ldg(dt_tab + b);
    binned[pos] = make_uint4(e.x, b, (unsigned int)__double2loint(dt), (unsigned int)__double2hiint(dt));
  }
}

}  // namespace cmaxb
