// be_xchg.cuh -- time-sharded back-end window over several GPUs with the exchange done by the kernels themselves over
// peer memory (NVLink / NVSwitch, buffers opened with CUDA IPC): SURVEY section 8e, BASELINE config C5.
//
// What is exchanged, and why it is SPARSE: rank r scatters the events of its time slab.  A slab covers a fraction of the
// trajectory, so its votes land in the part of the panorama the camera saw during the slab (a few percent of a 4096x2048
// panorama) -- the whole-plane all-reduce (and the reduce-scatter of the row-band variant) moves 33.5 MB per rank of which
// >90 % are zeros.  Here every rank keeps a byte per 32x32 panorama tile ("dirty": some vote of this evaluation landed
// there) and only dirty tiles travel:
//   push   : pixels of the dirty tiles are ADDED (red.global.add.v4.f32 over NVLink) into the accumulator of the rank
//            that owns the row band (and into the neighbour's when the row lies in its halo); then one flag per peer
//   blur   : the owner blurs its band (+ halo) once all peers' flags arrived; S1, S2 of its own rows go to every peer as
//            tagged words (data and flag in one 8-byte store); every rank adds them up in rank order -> identical mean
//   adjoint: the owner forms G on its band and raises a flag at every peer
//   pull   : every rank reads G for ITS dirty tiles out of the owners' bands (coalesced peer loads) into its local
//            corner-packed adjoint image, so the gather pass is the single-GPU one
//   grad   : partial gradients (3 K_opt doubles) go to every peer as tagged words and are summed in rank order
// Every wait is bounded (kBeXTimeoutNs): a peer that never arrives raises the fault word instead of hanging the GPU.
#pragma once
#include "common.cuh"
#include "be_kernels.cuh"

namespace cmaxb {

constexpr int kBeXMaxWorld = 8;
constexpr int kBeXTile = 32;
constexpr int kBeXSumSlots = 4;
constexpr int kBeXGradMax = 3 * 1024;          // doubles per rank in the gradient exchange
constexpr unsigned long long kBeXTimeoutNs = 5ull * 1000ull * 1000ull * 1000ull;

// One rank's IPC-exported block (byte offsets; every rank computes the same layout)
struct BeXLayout {
  size_t acc[2];        // float [ce * W]     extended-band accumulators (parity of the evaluation number)
  size_t G;             // float [ce * W]     adjoint image of the band (row j = panorama row first + j)
  size_t push_flag;     // u64 [world]        evaluation number up to which rank q's pushes are complete
  size_t g_flag;        // u64 [world]        evaluation number of owner q's G band
  size_t sums;          // u64 [slots][world][4]   tagged words of (S1, S2)
  size_t grad;          // u64 [2][world][2 * kBeXGradMax]
  size_t total;
};

inline BeXLayout be_x_layout(int W, int ce, int world) {
  BeXLayout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  L.acc[0] = take(sizeof(float) * (size_t)ce * W);
  L.acc[1] = take(sizeof(float) * (size_t)ce * W);
  L.G = take(sizeof(float) * (size_t)ce * W);
  L.push_flag = take(8 * (size_t)world);
  L.g_flag = take(8 * (size_t)world);
  L.sums = take(8 * (size_t)kBeXSumSlots * world * 4);
  L.grad = take(8 * (size_t)2 * world * 2 * kBeXGradMax);
  L.total = o;
  return L;
}

struct BeXPeers {
  char* base[kBeXMaxWorld];
  int world, rank;
};

struct BeXGeom {
  int W, H, ntx, nty;     // panorama, tiles
  int hb, hl, ce;         // band height, halo, rows of an extended band
};

__device__ __forceinline__ unsigned long long be_x_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned long long be_x_ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// wait until flags[0 .. world) >= seq (thread 0 of the CTA; the others wait at the barrier)
__device__ __forceinline__ void be_x_wait_flags(const unsigned long long* flags, int world, unsigned long long seq,
                                                unsigned long long* fault, unsigned long long code) {
  if (threadIdx.x == 0) {
    const unsigned long long t0 = be_x_timer();
    for (int r = 0; r < world; ++r) {
      unsigned int spins = 0;
      while (be_x_ld_flag(flags + r) < seq) {
        if ((++spins & 0xffu) == 0) {
          if (*(volatile unsigned long long*)fault) break;
          if (be_x_timer() - t0 > kBeXTimeoutNs) { *(volatile unsigned long long*)fault = code | ((unsigned long long)r << 8); break; }
        }
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}

__device__ __forceinline__ void be_x_ll_store(unsigned long long* dst, double v, unsigned long long tag_hi) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"((bits & 0xffffffffull) | tag_hi) : "memory");
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 1), "l"((bits >> 32) | tag_hi) : "memory");
}

// spin until both words carry the tag; returns the value (0 and the fault word set on time-out)
__device__ __forceinline__ double be_x_ll_load(const unsigned long long* src, unsigned long long tag_hi, unsigned long long* fault,
                                               unsigned long long code) {
  const unsigned long long t0 = be_x_timer();
  unsigned long long lo = 0, hi = 0;
  unsigned int spins = 0;
  for (;;) {
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(lo) : "l"(src) : "memory");
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(hi) : "l"(src + 1) : "memory");
    if ((lo & 0xffffffff00000000ull) == tag_hi && (hi & 0xffffffff00000000ull) == tag_hi) break;
    if ((++spins & 0xffu) == 0) {
      if (*(volatile unsigned long long*)fault) return 0.0;
      if (be_x_timer() - t0 > kBeXTimeoutNs) { *(volatile unsigned long long*)fault = code; return 0.0; }
    }
  }
  return __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
}

// dirty-tile bitmap (one bit per 32x32 panorama tile) staged in shared memory: the tile loops below then cost no global
// round trip per tile
__device__ __forceinline__ void be_x_stage_bits(const unsigned int* __restrict__ dirty, int ntiles, unsigned int* s_bits) {
  const int nwords = (ntiles + 31) >> 5;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) s_bits[i] = __ldcg(dirty + i);
  __syncthreads();
}
__device__ __forceinline__ bool be_x_bit(const unsigned int* s_bits, int t) { return (s_bits[t >> 5] >> (t & 31)) & 1u; }

// ---- clean: zero the cells of the tiles the PREVIOUS evaluation dirtied, clear their flags -----------------------------
__global__ void __launch_bounds__(256)
be_x_clean_kernel(float4* __restrict__ quad, unsigned int* __restrict__ dirty, BeXGeom g) {
  __shared__ unsigned int s_bits[kBeDirtyWords];
  const int ntiles = g.ntx * g.nty;
  be_x_stage_bits(dirty, ntiles, s_bits);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (!be_x_bit(s_bits, t)) continue;            // uniform per CTA
    const int tx = t % g.ntx, ty = t / g.ntx;
    for (int i = threadIdx.x; i < kBeXTile * kBeXTile; i += blockDim.x) {
      const int x = tx * kBeXTile + (i & 31), y = ty * kBeXTile + (i >> 5);
      if (x < g.W && y < g.H) quad[(long long)y * g.W + x] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x == 0) atomicAnd(dirty + (t >> 5), ~(1u << (t & 31)));
  }
}

// ---- push: pixels of the dirty tiles (and of the first row / column of their right / lower neighbours, which receive
// the other three corners of the votes) are added into the owners' extended-band accumulators ------------------------------
__global__ void __launch_bounds__(256)
be_x_push_kernel(const float4* __restrict__ quad, const unsigned int* __restrict__ dirty, BeXGeom g, BeXPeers peers, size_t acc_off,
                 size_t flag_off, unsigned long long seq, unsigned int* __restrict__ ticket) {
  __shared__ unsigned int s_bits[kBeDirtyWords];
  const int ntiles = g.ntx * g.nty;
  be_x_stage_bits(dirty, ntiles, s_bits);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tx = t % g.ntx, ty = t / g.ntx;
    const bool d_here = be_x_bit(s_bits, t);
    const bool d_left = tx > 0 && be_x_bit(s_bits, t - 1);
    const bool d_up = ty > 0 && be_x_bit(s_bits, t - g.ntx);
    const bool d_ul = tx > 0 && ty > 0 && be_x_bit(s_bits, t - g.ntx - 1);
    if (!(d_here || d_left || d_up || d_ul)) continue;       // uniform per CTA
    // four consecutive pixels of one row per thread: 8 threads per row, 32 rows
    const int row = threadIdx.x >> 3, x0 = tx * kBeXTile + (threadIdx.x & 7) * 4, y = ty * kBeXTile + row;
    // a tile that is only a neighbour of a dirty one holds votes in its first row / column only
    if (!d_here && !(row == 0 && (d_up || d_ul)) && !((threadIdx.x & 7) == 0 && (d_left || d_ul))) continue;
    if (y >= g.H || x0 >= g.W) continue;
    float l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      float v = 0.f;
      if (x < g.W) {
        const long long p = (long long)y * g.W + x;
        v = __ldcg(&quad[p]).x;
        if (x > 0) v += __ldcg(&quad[p - 1]).y;
        if (y > 0) { v += __ldcg(&quad[p - g.W]).z; if (x > 0) v += __ldcg(&quad[p - g.W - 1]).w; }
      }
      l[k] = v;
    }
    if (l[0] == 0.f && l[1] == 0.f && l[2] == 0.f && l[3] == 0.f) continue;
    // owners: every band c with c hb - hl <= y < (c + 1) hb + hl
    const int c_lo = y >= g.hl ? (y - g.hl) / g.hb : 0;
    const int c_hi = min(peers.world - 1, (y + g.hl) / g.hb);
    for (int c = c_lo; c <= c_hi; ++c) {
      const int j = y - c * g.hb + g.hl;
      if (j < 0 || j >= g.ce) continue;
      float* dst = reinterpret_cast<float*>(peers.base[c] + acc_off) + (long long)j * g.W + x0;
      if (x0 + 3 < g.W && ((g.W & 3) == 0)) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(l[0]), "f"(l[1]), "f"(l[2]), "f"(l[3]) : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (x0 + k < g.W && l[k] != 0.f) atomicAdd(dst + k, l[k]);
      }
    }
  }
  // all pushes of this CTA performed before the ticket; the last CTA raises this rank's flag at every peer
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int n = atomicAdd(ticket, 1u);
    if (n == gridDim.x - 1) {
      *ticket = 0u;
      __threadfence_system();
      for (int r = 0; r < peers.world; ++r) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.base[r] + flag_off) + peers.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
      }
    }
  }
}

// ---- wait for every rank's pushes (stream-ordered before the band blur) ------------------------------------------------
__global__ void be_x_wait_kernel(const unsigned long long* flags, int world, unsigned long long seq, unsigned long long* fault,
                                 unsigned long long code) {
  be_x_wait_flags(flags, world, seq, fault, code);
}

// ---- S1, S2 of the own rows -> every peer (tagged words); all ranks' sums, added in rank order -> contrast, mean --------
__global__ void be_x_sums_kernel(const double* __restrict__ sums2, BeXPeers peers, size_t sums_off, unsigned long long seq, double Np,
                                 int measure, double* __restrict__ result, double* __restrict__ mean, unsigned long long* fault) {
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  {
    const int r = threadIdx.x >> 1, k = threadIdx.x & 1;        // thread = (peer, which sum)
    if (r < peers.world) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(peers.base[r] + sums_off) +
                                ((seq % kBeXSumSlots) * peers.world + peers.rank) * 4 + 2 * k;
      be_x_ll_store(dst, sums2[k], tag);
    }
  }
  if (threadIdx.x != 0) return;
  const unsigned long long* slot = reinterpret_cast<const unsigned long long*>(peers.base[peers.rank] + sums_off) + (seq % kBeXSumSlots) * peers.world * 4;
  double S1 = 0.0, S2 = 0.0;
  for (int r = 0; r < peers.world; ++r) {
    S1 += be_x_ll_load(slot + r * 4, tag, fault, 0x30 | ((unsigned long long)r << 8));
    S2 += be_x_ll_load(slot + r * 4 + 2, tag, fault, 0x30 | ((unsigned long long)r << 8));
  }
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

// ---- G band ready: flag at every peer -----------------------------------------------------------------------------------
__global__ void be_x_flag_kernel(BeXPeers peers, size_t flag_off, unsigned long long seq) {
  if (threadIdx.x >= peers.world) return;
  __threadfence_system();
  unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.base[threadIdx.x] + flag_off) + peers.rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
}

// ---- pull: G of the dirty tiles out of the owners' bands -> local corner-packed adjoint image ---------------------------
__device__ __forceinline__ float be_x_peer_G(const BeXPeers& peers, const BeXGeom& g, size_t g_off, int x, int y) {
  const int c = min(y / g.hb, peers.world - 1);
  const int first = max(0, c * g.hb - g.hl);
  const float* src = reinterpret_cast<const float*>(peers.base[c] + g_off) + (long long)(y - first) * g.W + x;
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(src) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
be_x_pull_kernel(float4* __restrict__ GQ, const unsigned int* __restrict__ dirty, BeXGeom g, BeXPeers peers, size_t g_off,
                 const unsigned long long* __restrict__ g_flags, unsigned long long seq, unsigned long long* fault) {
  __shared__ float s_g[kBeXTile + 1][kBeXTile + 1];
  __shared__ unsigned int s_bits[kBeDirtyWords];
  const int ntiles = g.ntx * g.nty;
  be_x_stage_bits(dirty, ntiles, s_bits);
  // a CTA without a dirty tile has nothing to wait for
  bool any = false;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) any = any || be_x_bit(s_bits, t);
  if (!any) return;
  be_x_wait_flags(g_flags, peers.world, seq, fault, 0x40);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (!be_x_bit(s_bits, t)) continue;            // uniform per CTA
    const int tx = t % g.ntx, ty = t / g.ntx;
    __syncthreads();
    // 33 x 33 values, up to 5 per thread: all peer loads issued before the first shared-memory store (NVLink latency once)
    constexpr int kPer = ((kBeXTile + 1) * (kBeXTile + 1) + 255) / 256;
    float v[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = threadIdx.x + q * 256;
      const int r = i / (kBeXTile + 1), c = i - r * (kBeXTile + 1);
      const int x = tx * kBeXTile + c, y = ty * kBeXTile + r;
      v[q] = (i < (kBeXTile + 1) * (kBeXTile + 1) && x < g.W && y < g.H) ? be_x_peer_G(peers, g, g_off, x, y) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = threadIdx.x + q * 256;
      if (i < (kBeXTile + 1) * (kBeXTile + 1)) s_g[i / (kBeXTile + 1)][i % (kBeXTile + 1)] = v[q];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBeXTile * kBeXTile; i += blockDim.x) {
      const int r = i >> 5, c = i & 31;
      const int x = tx * kBeXTile + c, y = ty * kBeXTile + r;
      if (x < g.W && y < g.H) GQ[(long long)y * g.W + x] = make_float4(s_g[r][c], s_g[r][c + 1], s_g[r + 1][c], s_g[r + 1][c + 1]);
    }
  }
}

// ---- gradient: partial sums to every peer, total in rank order ----------------------------------------------------------
__global__ void __launch_bounds__(256)
be_x_grad_kernel(double* __restrict__ grad, int P, BeXPeers peers, size_t grad_off, unsigned long long seq, unsigned long long* fault) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P) return;
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  const size_t slot = (size_t)(seq & 1) * peers.world * 2 * kBeXGradMax;
  const double mine = grad[j];
  for (int r = 0; r < peers.world; ++r) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(peers.base[r] + grad_off) + slot + (size_t)peers.rank * 2 * kBeXGradMax + 2 * j;
    be_x_ll_store(dst, mine, tag);
  }
  const unsigned long long* own = reinterpret_cast<const unsigned long long*>(peers.base[peers.rank] + grad_off) + slot;
  double s = 0.0;
  for (int r = 0; r < peers.world; ++r) s += be_x_ll_load(own + (size_t)r * 2 * kBeXGradMax + 2 * j, tag, fault, 0x50 | ((unsigned long long)r << 8));
  grad[j] = s;
}

}  // namespace cmaxb
