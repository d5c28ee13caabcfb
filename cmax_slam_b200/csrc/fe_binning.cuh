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
// Two launches per packet: count (+ validation + batch time offsets), scatter (+ the scan of the counts, in every CTA).
#pragma once
#include "common.cuh"
#include "fe_kernels.cuh"

namespace cmaxb {

constexpr int kBinTile = 32;          // source tile edge in pixels
constexpr int kBinThreads = 256;
constexpr int kBinChunk = 2048;       // events per CTA at most (8 per thread: ~490 CTAs for a 1M-event packet); rounded down to a multiple of the batch size
constexpr int kBinMaxTiles = 6016;    // shared-memory histogram capacity: the scatter pass needs 2 x 4 B per tile + 40 B static within the 48 KB default limit

__device__ __forceinline__ int bin_tile_of(uint4 e, int W, int H, int ntx) {
  int x = e.x & 0xffff, y = e.x >> 16;
  x = min(x, W - 1); y = min(y, H - 1);   // out-of-sensor events are rejected by validate_events_kernel
  return (y / kBinTile) * ntx + (x / kBinTile);
}

constexpr int kBinPer = kBinChunk / kBinThreads;   // events per thread, all loaded before the first use (one memory latency per pass)

// events [beg, end) of a chunk into registers: kBinPer independent coalesced 16-byte loads per thread
__device__ __forceinline__ void bin_load_chunk(const uint4* __restrict__ ev, long long beg, long long end, uint4 (&e)[kBinPer]) {
#pragma unroll
  for (int k = 0; k < kBinPer; ++k) {
    const long long i = beg + threadIdx.x + (long long)k * kBinThreads;
    e[k] = i < end ? __ldg(ev + i) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// pass 1: tile histogram (per-CTA shared histogram, one global atomic per non-empty bin) + validation of the pixel range
// (flags[1] = 1: the reference's precomputed_bearing_vectors_.at(...) would throw, local_image_warped_events.cpp:100)
// + the per-batch time offsets of the chunk's batches when `chunk` is a multiple of the batch size (dt_tab != null;
// flags[0] = 1 on a negative batch span, local_image_warped_events.cpp:72)
__global__ void __launch_bounds__(kBinThreads)
fe_bin_count_kernel(const uint4* __restrict__ ev, long long n, int chunk, int W, int H, int ntx, int ntiles,
                    unsigned int* __restrict__ tile_count, int* __restrict__ flags, int bs, double t_ref, double* __restrict__ dt_tab,
                    long long nb) {
  extern __shared__ unsigned int s_hist[];
  const long long beg = blockIdx.x * (long long)chunk;
  const long long end = min(beg + (long long)chunk, n);
  uint4 e[kBinPer];
  bin_load_chunk(ev, beg, end, e);
  for (int i = threadIdx.x; i < ntiles; i += kBinThreads) s_hist[i] = 0u;
  __syncthreads();
  bool bad = false;
#pragma unroll
  for (int k = 0; k < kBinPer; ++k) {
    if (beg + threadIdx.x + (long long)k * kBinThreads >= end) break;
    bad = bad || (int)(e[k].x & 0xffff) >= W || (int)(e[k].x >> 16) >= H;
    atomicAdd(&s_hist[bin_tile_of(e[k], W, H, ntx)], 1u);
  }
  if (bad) fe_flag_raise(flags, 1);
  if (dt_tab) {
    const int per = chunk / bs;
    for (int lb = threadIdx.x; lb < per; lb += kBinThreads) {     // (more than 256 batches per chunk when the batch size is < 8)
      const long long b = blockIdx.x * (long long)per + lb;
      if (b < nb) {
        const long long b0 = b * bs;
        long long b1 = b0 + bs; if (b1 > n) b1 = n;
        const uint4 e0 = __ldg(ev + b0), e1 = __ldg(ev + b1 - 1);
        RosTime mid;
        if (!ros_batch_mid(RosTime{e0.y, e0.z}, RosTime{e1.y, e1.z}, &mid)) fe_flag_raise(flags, 0);
        dt_tab[b] = ros_to_sec(mid.sec, mid.nsec) - t_ref;                   // (:75)
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntiles; i += kBinThreads)
    if (s_hist[i]) atomicAdd(&tile_count[i], s_hist[i]);
}

// pass 2: every CTA forms the exclusive prefix of the tile counts itself (a few hundred numbers: cheaper than a launch of
// its own), claims its run inside every tile with one global atomic per non-empty bin, writes the 16-byte records.
// tile_count / tile_cursor are zero before the count pass; the LAST CTA to finish (ticket) zeroes them again for the next
// packet, so a packet costs two launches and nothing else on the stream.  CTA 0 writes tile_end[t] = end offset of tile t's run.
__global__ void __launch_bounds__(kBinThreads)
fe_bin_scatter_kernel(const uint4* __restrict__ ev, long long n, int chunk, int W, int H, int ntx, int ntiles, int batch_size,
                      const double* __restrict__ dt_tab, unsigned int* __restrict__ tile_count,
                      unsigned int* __restrict__ tile_cursor, unsigned int* __restrict__ tile_end, uint4* __restrict__ binned,
                      unsigned int* __restrict__ ticket) {
  extern __shared__ unsigned int s_mem[];
  unsigned int* s_hist = s_mem;            // counts, then running local cursor
  unsigned int* s_base = s_mem + ntiles;   // exclusive prefix of the tile counts, then global base of this CTA's run inside each tile
  __shared__ unsigned int s_warp[kBinThreads / 32];
  const long long beg = blockIdx.x * (long long)chunk;
  const long long end = min(beg + (long long)chunk, n);
  uint4 e[kBinPer];
  bin_load_chunk(ev, beg, end, e);
  // exclusive scan of tile_count: thread t owns tiles [t per, (t + 1) per)
  const int per = (ntiles + kBinThreads - 1) / kBinThreads;
  const int t0 = threadIdx.x * per;
  unsigned int sum = 0;
  for (int i = t0; i < min(t0 + per, ntiles); ++i) { s_hist[i] = 0u; sum += __ldcg(tile_count + i); }
  unsigned int inc = sum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += v;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  unsigned int run = inc - sum;
  for (int w = 0; w < wid; ++w) run += s_warp[w];
  for (int i = t0; i < min(t0 + per, ntiles); ++i) {
    const unsigned int c = __ldcg(tile_count + i);
    s_base[i] = run;
    run += c;
    if (blockIdx.x == 0) tile_end[i] = run;
  }
  __syncthreads();
  int tile[kBinPer];
#pragma unroll
  for (int k = 0; k < kBinPer; ++k) {
    tile[k] = -1;
    if (beg + threadIdx.x + (long long)k * kBinThreads < end) {
      tile[k] = bin_tile_of(e[k], W, H, ntx);
      atomicAdd(&s_hist[tile[k]], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntiles; i += kBinThreads) {
    const unsigned int c = s_hist[i];
    if (c) s_base[i] += atomicAdd(&tile_cursor[i], c);
    s_hist[i] = 0u;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kBinPer; ++k) {
    if (tile[k] < 0) continue;
    const long long i = beg + threadIdx.x + (long long)k * kBinThreads;
    const unsigned int pos = s_base[tile[k]] + atomicAdd(&s_hist[tile[k]], 1u);
    const unsigned int b = (unsigned int)(i / batch_size);
    const double dt = __ldcg(dt_tab + b);
    binned[pos] = make_uint4(e[k].x, b, (unsigned int)__double2loint(dt), (unsigned int)__double2hiint(dt));
  }
  // every CTA has read the counts and claimed its runs by the time it arrives here
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < ntiles; i += kBinThreads) { tile_count[i] = 0u; tile_cursor[i] = 0u; }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

}  // namespace cmaxb
