// fe_capi.cu -- C ABI of the front-end path (see include/cmax_b200.h).
#include "capi_common.cuh"
#include "fe_kernels.cuh"
#include "image_kernels.cuh"
#include "fe_mega.cuh"
#include "fe_binning.cuh"

#include <deque>

using namespace cmaxb;

// evaluations that may be queued on the stream before the oldest is fetched (results live in a ring of
// mapped host slots, so the host never has to drain the stream between launches)
constexpr int kFeRing = 4;

struct FeInflight {
  int k; bool grad; bool mega; int slot; unsigned long long seq; bool xchg;
};

namespace cmaxb {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};
}  // namespace cmaxb

struct cmaxb_fe {
  cmaxb_fe_cfg cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long long A = 0;          // pixels
  int kmax = 1;
  Taps taps{};
  double4* d_lut = nullptr;
  uint4* d_ev = nullptr; size_t ev_cap = 0;
  double* d_dt = nullptr; size_t dt_cap = 0;
  uint2* d_bev = nullptr; size_t bev_cap = 0; bool have_bins = false; bool use_bins = true;   // spatially binned copy of the packet
  unsigned int* d_tile_count = nullptr; unsigned int* d_tile_cursor = nullptr; int ntiles = 0, ntx = 0;
  long long n = 0, nb = 0;
  bool have_packet = false; bool flags_pending = false;
  int* d_flags = nullptr; int* h_flags = nullptr;
  // value accumulators: two corner-split ("quad") images used alternately; the blur kernel of one
  // evaluation clears the image the next evaluation scatters into (no memset in steady state)
  float4* d_quad[2] = {nullptr, nullptr}; int quad_cur = 0; int quad_dirty[2] = {0, 0};
  float* d_blur1 = nullptr;        // blurred IWE (adjoint mode, get_iwe)
  float4* d_GQ = nullptr;          // adjoint image, one float4 per cell
  float4* d_img4 = nullptr; float4* d_blur4 = nullptr;   // DENSE mode: (I, dI/dw) accumulators
  int* d_cells = nullptr; size_t cells_cap = 0;
  double* d_omegas = nullptr; double* h_omegas = nullptr;
  double* d_acc = nullptr; unsigned int* d_ticket = nullptr; unsigned int* d_ticket2 = nullptr;
  double* d_gacc = nullptr; double* d_result = nullptr; double* d_mean = nullptr; double* h_result = nullptr;
  // fused (single cooperative kernel) evaluation
  int mega_grid = 0; int mega_th = 16; bool mega_ok = false;
  double* d_part_img = nullptr; double* d_part_ev = nullptr;
  double* h_mega_result = nullptr; double* d_mega_result = nullptr;   // mapped pinned memory
  bool last_mega = false;
  unsigned long long* h_phase = nullptr; unsigned long long* d_phase = nullptr;   // mapped: phase boundary timestamps
  unsigned long long* h_done = nullptr; unsigned long long* d_done = nullptr;     // mapped: completion sequence number
  unsigned long long seq = 0;
  double* d_mirror = nullptr;   // caller-owned device buffer [kmax][4] (cmaxb_fe_set_result_mirror)
  bool force_multi_kernel = false;   // CMAXB_FE_MULTI_KERNEL=1: stand-alone kernels (profiling / A-B comparison)
  int last_k = 0; bool last_grad = false; bool pending = false;
  std::deque<FeInflight> inflight;   // launched, not yet fetched (FIFO)
  int ring_next = 0;
  bool gather_f32 = false;           // CMAXB_FE_GATHER_F32=1: Jacobian chain of the gather in f32 (measured 1.5 % faster; default = the reference's f64 chain)
  // fused result exchange over peer memory (cmaxb_fe_exchange_*)
  int x_world = 0, x_rank = 0; bool x_on = false;
  ulonglong2* x_local = nullptr; size_t x_bytes = 0;
  ulonglong2* x_peer[kXMaxWorld] = {};
  double* x_all_dev = nullptr;
  double* h_xall = nullptr; double* d_xall = nullptr;        // mapped: [kFeRing][world][kmax][4]
  unsigned int* h_xerr = nullptr; unsigned int* d_xerr = nullptr;
  unsigned long long x_seq = 0;
  KernelProfiler prof;
};

static FeGeom fe_geom(const cmaxb_fe* fe) {
  FeGeom g;
  g.ev = fe->d_ev; g.bev = nullptr; g.n = fe->n; g.batch_size = fe->cfg.batch_size; g.dt_tab = fe->d_dt; g.lut = fe->d_lut;
  g.W = fe->cfg.width; g.H = fe->cfg.height;
  g.fx = fe->cfg.fx; g.fy = fe->cfg.fy; g.cx = fe->cfg.cx; g.cy = fe->cfg.cy;
  return g;
}

extern "C" int cmaxb_fe_create(const cmaxb_fe_cfg* cfg, cmaxb_fe** out) {
  if (!cfg || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->width < 4 || cfg->height < 4 || !cfg->lut_xyz || cfg->batch_size <= 0)
    return set_error(CMAXB_ERR_INVALID, "bad front-end configuration");
  if (cfg->grad_mode != CMAXB_GRAD_DENSE && cfg->grad_mode != CMAXB_GRAD_ADJOINT)
    return set_error(CMAXB_ERR_INVALID, "bad grad_mode");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return set_error(CMAXB_ERR_CUDA, "no CUDA device: libcmax_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return set_error(CMAXB_ERR_INVALID, "bad device ordinal");
  CMAXB_CUDA_TRY(cudaSetDevice(cfg->device));
  cmaxb_fe* fe = new cmaxb_fe();
  fe->cfg = *cfg;
  fe->cfg.lut_xyz = nullptr;
  fe->device = cfg->device;
  fe->A = (long long)cfg->width * cfg->height;
  fe->kmax = cfg->max_hypotheses > 0 ? cfg->max_hypotheses : 1;
  {
    const char* mk = getenv("CMAXB_FE_MULTI_KERNEL");
    fe->force_multi_kernel = mk && mk[0] == '1';
    const char* nb = getenv("CMAXB_FE_NO_BINNING");   // A/B switch: evaluate the packet in arrival (time) order
    fe->use_bins = !(nb && nb[0] == '1');
    const char* g32 = getenv("CMAXB_FE_GATHER_F32");  // A/B switch: Jacobian chain of the gather pass in f32
    fe->gather_f32 = g32 && g32[0] == '1';
  }
  int rc = make_taps(cfg->blur_sigma, &fe->taps);
  if (rc != CMAXB_OK) { delete fe; return rc; }
  if (fe->taps.r + 2 > cfg->width || fe->taps.r + 2 > cfg->height) { delete fe; return set_error(CMAXB_ERR_INVALID, "image smaller than the blur kernel"); }
  auto fail = [&](int code) { cmaxb_fe_destroy(fe); return code; };
  if (cfg->stream) fe->stream = (cudaStream_t)cfg->stream;
  else {
    if (cudaStreamCreateWithFlags(&fe->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(set_error(CMAXB_ERR_CUDA, "cudaStreamCreate failed"));
    fe->own_stream = true;
  }
  if (fe->prof.init() != CMAXB_OK) return fail(CMAXB_ERR_CUDA);
  // LUT padded to 32-byte records
  {
    std::vector<double4> lut((size_t)fe->A);
    for (long long i = 0; i < fe->A; ++i) lut[i] = make_double4(cfg->lut_xyz[3 * i], cfg->lut_xyz[3 * i + 1], cfg->lut_xyz[3 * i + 2], 0.0);
    if (dev_alloc(&fe->d_lut, (size_t)fe->A) != CMAXB_OK) return fail(CMAXB_ERR_CUDA);
    if (cudaMemcpy(fe->d_lut, lut.data(), sizeof(double4) * fe->A, cudaMemcpyHostToDevice) != cudaSuccess) return fail(set_error(CMAXB_ERR_CUDA, "LUT upload failed"));
  }
  const size_t k = (size_t)fe->kmax, A = (size_t)fe->A;
  bool ok = true;
  ok = ok && dev_alloc(&fe->d_flags, 1) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_quad[0], k * A) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_quad[1], k * A) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_omegas, k * 3) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_acc, k * kNAcc * kMaxImgCtas) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_ticket, k) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_ticket2, k) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_gacc, k * 3 * kMaxEventCtas) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_result, k * 4) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_mean, k) == CMAXB_OK;
  if (cfg->grad_mode == CMAXB_GRAD_ADJOINT) {
    ok = ok && dev_alloc(&fe->d_blur1, k * A) == CMAXB_OK;
    ok = ok && dev_alloc(&fe->d_GQ, k * A) == CMAXB_OK;
  } else {
    ok = ok && dev_alloc(&fe->d_img4, k * A) == CMAXB_OK;
  }
  if (!ok) return fail(CMAXB_ERR_CUDA);
  ok = ok && cudaMallocHost((void**)&fe->h_flags, sizeof(int)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&fe->h_omegas, sizeof(double) * 3 * k) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&fe->h_result, sizeof(double) * 4 * k) == cudaSuccess;
  ok = ok && cudaMemset(fe->d_ticket, 0, sizeof(unsigned) * k) == cudaSuccess;
  ok = ok && cudaMemset(fe->d_ticket2, 0, sizeof(unsigned) * k) == cudaSuccess;
  ok = ok && cudaMemset(fe->d_result, 0, sizeof(double) * k * 4) == cudaSuccess;
  ok = ok && cudaMemset(fe->d_quad[0], 0, sizeof(float4) * k * A) == cudaSuccess;
  ok = ok && cudaMemset(fe->d_quad[1], 0, sizeof(float4) * k * A) == cudaSuccess;
  if (!ok) return fail(set_error(CMAXB_ERR_CUDA, "front-end buffer allocation failed"));
  // fused evaluation kernel: co-resident grid size from the occupancy API
  {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, fe->device);
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, fe->device);
    const size_t smem = mega_smem_bytes(fe->taps.r);   // sized for the tallest tile
    int occ = 0;
    cudaError_t e1 = cudaFuncSetAttribute(fe_eval_megakernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e2 = cudaFuncSetAttribute(fe_eval_megakernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e3 = (fe->taps.r == 4)
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fe_eval_megakernel<4>, kMegaThreads, smem)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fe_eval_megakernel<-1>, kMegaThreads, smem);
    if (coop && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess && occ > 0) {
      int grid = occ * nsm;
      if (grid > kMegaMaxCtas) grid = kMegaMaxCtas;
      {
        // tuning aid: CMAXB_FE_GRID_FRACTION=0.5 launches the fused kernel on half of the co-resident CTAs, so that two
        // handles on two streams can run their (latency-bound) evaluations side by side; default 1 = whole GPU
        const char* gf = getenv("CMAXB_FE_GRID_FRACTION");
        const double f = gf ? atof(gf) : 1.0;
        if (f > 0.0 && f < 1.0) { grid = (int)(grid * f); if (grid < nsm / 4) grid = nsm / 4; if (grid < 1) grid = 1; }
      }
      fe->mega_grid = grid;
      fe->mega_th = mega_tile_height(cfg->width, cfg->height, grid);
      const size_t kk = (size_t)fe->kmax;
      bool okm = dev_alloc(&fe->d_part_img, kk * kMegaMaxCtas * 2) == CMAXB_OK && dev_alloc(&fe->d_part_ev, kk * kMegaMaxCtas * 3) == CMAXB_OK;
      okm = okm && cudaHostAlloc((void**)&fe->h_mega_result, sizeof(double) * 4 * kk * kFeRing, cudaHostAllocMapped) == cudaSuccess;
      okm = okm && cudaHostGetDevicePointer((void**)&fe->d_mega_result, fe->h_mega_result, 0) == cudaSuccess;
      okm = okm && cudaHostAlloc((void**)&fe->h_done, sizeof(unsigned long long) * 8, cudaHostAllocMapped) == cudaSuccess;
      okm = okm && cudaHostGetDevicePointer((void**)&fe->d_done, fe->h_done, 0) == cudaSuccess;
      if (okm) fe->h_done[0] = 0;
      okm = okm && cudaHostAlloc((void**)&fe->h_phase, sizeof(unsigned long long) * 16, cudaHostAllocMapped) == cudaSuccess;
      okm = okm && cudaHostGetDevicePointer((void**)&fe->d_phase, fe->h_phase, 0) == cudaSuccess;
      if (okm && !fe->d_blur1) okm = dev_alloc(&fe->d_blur1, kk * A) == CMAXB_OK;
      if (okm && !fe->d_GQ) okm = dev_alloc(&fe->d_GQ, kk * A) == CMAXB_OK;
      fe->mega_ok = okm;
    }
    (void)cudaGetLastError();
    if (!fe->mega_ok) return fail(set_error(CMAXB_ERR_CUDA, "cooperative launch unavailable: the fused evaluation kernel cannot run on this device"));
  }
  *out = fe;
  return CMAXB_OK;
}

extern "C" void cmaxb_fe_destroy(cmaxb_fe* fe) {
  if (!fe) return;
  cudaSetDevice(fe->device);
  if (fe->stream) cudaStreamSynchronize(fe->stream);
  cudaFree(fe->d_lut); cudaFree(fe->d_ev); cudaFree(fe->d_dt); cudaFree(fe->d_flags);
  cudaFree(fe->d_bev); cudaFree(fe->d_tile_count); cudaFree(fe->d_tile_cursor);
  cudaFree(fe->d_quad[0]); cudaFree(fe->d_quad[1]); cudaFree(fe->d_blur1); cudaFree(fe->d_GQ); cudaFree(fe->d_img4); cudaFree(fe->d_blur4);
  cudaFree(fe->d_cells); cudaFree(fe->d_omegas); cudaFree(fe->d_acc); cudaFree(fe->d_ticket); cudaFree(fe->d_ticket2);
  cudaFree(fe->d_gacc); cudaFree(fe->d_result); cudaFree(fe->d_mean);
  if (fe->h_flags) cudaFreeHost(fe->h_flags);
  if (fe->h_omegas) cudaFreeHost(fe->h_omegas);
  if (fe->h_result) cudaFreeHost(fe->h_result);
  if (fe->h_mega_result) cudaFreeHost(fe->h_mega_result);
  if (fe->h_phase) cudaFreeHost(fe->h_phase);
  if (fe->h_done) cudaFreeHost(fe->h_done);
  cudaFree(fe->d_part_img); cudaFree(fe->d_part_ev);
  for (int r = 0; r < fe->x_world; ++r)
    if (r != fe->x_rank && fe->x_peer[r]) cudaIpcCloseMemHandle(fe->x_peer[r]);
  cudaFree(fe->x_local);
  if (fe->h_xall) cudaFreeHost(fe->h_xall);
  if (fe->h_xerr) cudaFreeHost(fe->h_xerr);
  fe->prof.destroy();
  if (fe->own_stream && fe->stream) cudaStreamDestroy(fe->stream);
  delete fe;
}

static int fe_check_packet_flags(cmaxb_fe* fe) {
  // validation result of an asynchronous set_packet (its D2H copy precedes every later operation on the stream)
  if (!fe->flags_pending) return CMAXB_OK;
  fe->flags_pending = false;
  if (*fe->h_flags & 2) { fe->have_packet = false; return set_error(CMAXB_ERR_EVENT_RANGE, "event pixel outside the sensor"); }
  if (*fe->h_flags & 1) { fe->have_packet = false; return set_error(CMAXB_ERR_TIME_ORDER, "Events must span a non-negative time interval"); }
  return CMAXB_OK;
}

static int fe_set_packet_impl(cmaxb_fe* fe, const cmaxb_event* events, size_t n, double t_ref_sec, bool wait) {
  if (!fe || (!events && n > 0)) return set_error(CMAXB_ERR_INVALID, "null argument");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  if (fe->pending) { CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->stream)); fe->pending = false; }
  fe->have_packet = false;
  const long long bs = fe->cfg.batch_size;
  const long long nb = ((long long)n + bs - 1) / bs;
  if (n > fe->ev_cap) {
    cudaFree(fe->d_ev); fe->d_ev = nullptr; fe->ev_cap = 0;
    CMAXB_TRY(dev_alloc(&fe->d_ev, n));
    fe->ev_cap = n;
  }
  if ((size_t)nb > fe->dt_cap) {
    cudaFree(fe->d_dt); fe->d_dt = nullptr; fe->dt_cap = 0;
    CMAXB_TRY(dev_alloc(&fe->d_dt, (size_t)nb));
    fe->dt_cap = (size_t)nb;
  }
  fe->n = (long long)n; fe->nb = nb;
  if (n > 0) {
    cudaStream_t s = fe->stream;
    CMAXB_CUDA_TRY(cudaMemcpyAsync(fe->d_ev, events, sizeof(cmaxb_event) * n, cudaMemcpyDefault, s));   // host (pinned: DMA) or device memory (UVA)
    CMAXB_CUDA_TRY(cudaMemsetAsync(fe->d_flags, 0, sizeof(int), s));
    const uint4* ev = fe->d_ev; const long long nn = fe->n; int* flags = fe->d_flags;
    const int W = fe->cfg.width, H = fe->cfg.height; double* dt = fe->d_dt; const int ibs = (int)bs;
    CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
      validate_events_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(ev, nn, W, H, flags);
    }));
    CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
      fe_batch_dt_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(ev, nn, ibs, t_ref_sec, dt, nb, flags);
    }));
    // one-time spatial binning of the packet (reused by every evaluation until the next set_packet)
    fe->have_bins = false;
    fe->ntx = (W + kBinTile - 1) / kBinTile;
    fe->ntiles = fe->ntx * ((H + kBinTile - 1) / kBinTile);
    if (fe->use_bins && fe->ntiles <= kBinMaxTiles && nn < (1LL << 32)) {
      if (n > fe->bev_cap) {
        cudaFree(fe->d_bev); fe->d_bev = nullptr; fe->bev_cap = 0;
        CMAXB_TRY(dev_alloc(&fe->d_bev, n));
        fe->bev_cap = n;
      }
      if (!fe->d_tile_count) {
        CMAXB_TRY(dev_alloc(&fe->d_tile_count, (size_t)kBinMaxTiles));
        CMAXB_TRY(dev_alloc(&fe->d_tile_cursor, (size_t)kBinMaxTiles));
      }
      const int ntiles = fe->ntiles, ntx = fe->ntx;
      const unsigned nchunks = (unsigned)((nn + kBinChunk - 1) / kBinChunk);
      CMAXB_CUDA_TRY(cudaMemsetAsync(fe->d_tile_count, 0, sizeof(unsigned int) * ntiles, s));
      CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
        fe_bin_count_kernel<<<nchunks, kBinThreads, sizeof(unsigned int) * ntiles, s>>>(ev, nn, W, H, ntx, ntiles, fe->d_tile_count);
      }));
      CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
        fe_bin_scan_kernel<<<1, 1024, 0, s>>>(fe->d_tile_count, ntiles, fe->d_tile_cursor);
      }));
      CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
        fe_bin_scatter_kernel<<<nchunks, kBinThreads, 2 * sizeof(unsigned int) * ntiles, s>>>(ev, nn, W, H, ntx, ntiles, ibs, fe->d_tile_cursor, fe->d_bev);
      }));
      fe->have_bins = true;
    }
    CMAXB_CUDA_TRY(cudaMemcpyAsync(fe->h_flags, fe->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s));
    fe->flags_pending = true;
    fe->have_packet = true;
    if (wait) {
      CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
      return fe_check_packet_flags(fe);
    }
    return CMAXB_OK;
  }
  fe->have_packet = true;
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_set_packet(cmaxb_fe* fe, const cmaxb_event* events, size_t n, double t_ref_sec) {
  return fe_set_packet_impl(fe, events, n, t_ref_sec, true);
}
extern "C" int cmaxb_fe_set_packet_async(cmaxb_fe* fe, const cmaxb_event* events, size_t n, double t_ref_sec) {
  return fe_set_packet_impl(fe, events, n, t_ref_sec, false);
}

static int fe_upload_omegas(cmaxb_fe* fe, const double* omegas, int k) {
  if (fe->pending) { CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->stream)); fe->pending = false; }
  std::memcpy(fe->h_omegas, omegas, sizeof(double) * 3 * k);
  CMAXB_CUDA_TRY(cudaMemcpyAsync(fe->d_omegas, fe->h_omegas, sizeof(double) * 3 * k, cudaMemcpyHostToDevice, fe->stream));
  return CMAXB_OK;
}

static dim3 fe_event_grid(const cmaxb_fe* fe, int k) {
  long long blocks = (fe->n + kFeThreads - 1) / kFeThreads;
  const long long cap = kMaxEventCtas;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return dim3((unsigned)blocks, (unsigned)k, 1);
}

// scatter the value votes of k hypotheses into the current quad accumulator
static int fe_run_scatter_value(cmaxb_fe* fe, int k) {
  cudaStream_t s = fe->stream;
  const int cur = fe->quad_cur;
  if (fe->quad_dirty[cur] > 0) {   // only after start-up / a change of k / a debug call: steady state skips this
    const int planes = fe->quad_dirty[cur];
    CMAXB_TRY(fe->prof.run(CMAXB_K_ZERO, s, false, [&] { cudaMemsetAsync(fe->d_quad[cur], 0, sizeof(float4) * fe->A * planes, s); }));
    fe->quad_dirty[cur] = 0;
  }
  if (fe->n > 0) {
    const FeGeom g = fe_geom(fe);
    CMAXB_TRY(fe->prof.run(CMAXB_K_FE_SCATTER, s, true, [&] {
      fe_scatter_kernel<2><<<fe_event_grid(fe, k), kFeThreads, 0, s>>>(g, fe->d_omegas, nullptr, fe->d_quad[cur], fe->A);
    }));
    fe->quad_dirty[cur] = k;
  }
  return CMAXB_OK;
}
// blur + reduce the current quad accumulator (k planes); optionally keep the blurred image; clears the
// OTHER quad accumulator and makes it current.
static int fe_run_value_image(cmaxb_fe* fe, int k, const Taps& taps, bool write_out) {
  cudaStream_t s = fe->stream;
  const int cur = fe->quad_cur, oth = cur ^ 1;
  const SrcQuad src{fe->d_quad[cur], fe->A};
  const ReduceOut ro{fe->d_acc, fe->d_ticket, fe->d_result, fe->d_mean};
  const int W = fe->cfg.width, H = fe->cfg.height, measure = fe->cfg.contrast_measure;
  // the other image can be cleared by this kernel if its dirty planes are covered by our k planes
  float4* zero_ptr = (fe->quad_dirty[oth] > 0 && fe->quad_dirty[oth] <= k) ? fe->d_quad[oth] : nullptr;
  if (write_out && !fe->d_blur1) CMAXB_TRY(dev_alloc(&fe->d_blur1, (size_t)fe->kmax * fe->A));
  cudaError_t le = cudaSuccess;
  CMAXB_TRY(fe->prof.run(CMAXB_K_BLUR_REDUCE, s, true, [&] {
    le = write_out ? launch_blur_reduce<1, SrcQuad, true>(s, k, src, W, H, taps, fe->d_blur1, fe->A, ro, measure, zero_ptr, fe->A)
                   : launch_blur_reduce<1, SrcQuad, false>(s, k, src, W, H, taps, nullptr, 0, ro, measure, zero_ptr, fe->A);
  }));
  if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("blur_reduce launch: ") + cudaGetErrorString(le));
  if (zero_ptr) fe->quad_dirty[oth] = 0;
  fe->quad_cur = oth;
  return CMAXB_OK;
}
static int fe_run_scatter_dense(cmaxb_fe* fe, int k) {
  cudaStream_t s = fe->stream;
  if (!fe->d_img4) CMAXB_TRY(dev_alloc(&fe->d_img4, (size_t)fe->kmax * fe->A));
  CMAXB_TRY(fe->prof.run(CMAXB_K_ZERO, s, false, [&] { cudaMemsetAsync(fe->d_img4, 0, sizeof(float4) * fe->A * k, s); }));
  if (fe->n > 0) {
    const FeGeom g = fe_geom(fe);
    CMAXB_TRY(fe->prof.run(CMAXB_K_FE_SCATTER, s, true, [&] {
      fe_scatter_kernel<1><<<fe_event_grid(fe, k), kFeThreads, 0, s>>>(g, fe->d_omegas, nullptr, fe->d_img4, fe->A);
    }));
  }
  return CMAXB_OK;
}

// Wait until the stream is idle.  Evaluations already launched stay fetchable (their rows sit in the mapped ring).
static int fe_drain_stream(cmaxb_fe* fe) {
  if (fe->pending) { CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->stream)); fe->pending = false; }
  return CMAXB_OK;
}

// One cooperative launch per <= kMegaMaxHyp hypotheses; omegas travel as kernel parameters and the
// results land in a ring slot of mapped pinned memory (up to kFeRing launches may be outstanding).
static int fe_eval_launch_fused(cmaxb_fe* fe, const double* omegas, int k, int want_grad) {
  cudaStream_t s = fe->stream;
  if ((int)fe->inflight.size() >= kFeRing)
    return set_error(CMAXB_ERR_STATE, "too many outstanding evaluations: call cmaxb_fe_eval_fetch first");
  if (!fe->inflight.empty() && !fe->inflight.back().mega) CMAXB_TRY(fe_drain_stream(fe));
  if (fe->x_on && k > kMegaMaxHyp)
    return set_error(CMAXB_ERR_INVALID, "result exchange supports at most 32 hypotheses per launch");
  const int slot = fe->ring_next;
  fe->ring_next = (fe->ring_next + 1) % kFeRing;
  const int cur = fe->quad_cur, oth = cur ^ 1;
  if (fe->quad_dirty[cur] > 0) {
    const int planes = fe->quad_dirty[cur];
    CMAXB_TRY(fe->prof.run(CMAXB_K_ZERO, s, false, [&] { cudaMemsetAsync(fe->d_quad[cur], 0, sizeof(float4) * fe->A * planes, s); }));
    fe->quad_dirty[cur] = 0;
  }
  const bool clear_next = fe->quad_dirty[oth] > 0 && fe->quad_dirty[oth] <= k;
  for (int c0 = 0; c0 < k; c0 += kMegaMaxHyp) {
    const int kc = (k - c0 < kMegaMaxHyp) ? k - c0 : kMegaMaxHyp;
    FeMegaParams p;
    p.g = fe_geom(fe);
    p.g.bev = fe->have_bins ? fe->d_bev : nullptr;   // the fused kernel walks the tile-binned copy of the packet
    p.k = kc; p.th = fe->mega_th; p.want_grad = want_grad; p.measure = fe->cfg.contrast_measure; p.taps = fe->taps;
    for (int i = 0; i < 3 * kc; ++i) p.omegas[i] = omegas[3 * c0 + i];
    p.quad = fe->d_quad[cur] + (long long)c0 * fe->A;
    p.quad_next = clear_next ? fe->d_quad[oth] + (long long)c0 * fe->A : nullptr;
    p.blurred = fe->d_blur1 + (long long)c0 * fe->A;
    p.GQ = fe->d_GQ + (long long)c0 * fe->A;
    p.A = fe->A;
    p.part_img = fe->d_part_img + (long long)c0 * kMegaMaxCtas * 2;
    p.part_ev = fe->d_part_ev + (long long)c0 * kMegaMaxCtas * 3;
    p.ticket = fe->d_ticket;
    p.contrast_dev = fe->d_mean + c0;
    p.result = fe->d_mega_result + ((long long)slot * fe->kmax + c0) * 4;
    p.mirror = fe->d_mirror ? fe->d_mirror + 4 * c0 : nullptr;
    p.done_flag = fe->d_done;
    p.seq = ++fe->seq;
    p.phase_ns = fe->prof.enabled ? fe->d_phase : nullptr;
    p.gather_f32 = fe->gather_f32 ? 1 : 0;
    std::memset(&p.x, 0, sizeof(p.x));
    if (fe->x_on) {
      p.x.world = fe->x_world; p.x.rank = fe->x_rank; p.x.kmax = fe->kmax;
      p.x.seq = ++fe->x_seq;
      for (int r = 0; r < fe->x_world; ++r) p.x.peer[r] = fe->x_peer[r];
      p.x.all_host = fe->d_xall + (long long)slot * fe->x_world * fe->kmax * 4;
      p.x.all_dev = fe->x_all_dev;
      p.x.err = fe->d_xerr;
    }
    if (fe->prof.enabled) for (int i = 0; i < 16; ++i) fe->h_phase[i] = 0;
    void* args[] = {&p};
    const size_t smem = mega_smem_bytes(fe->taps.r);
    cudaError_t le = cudaSuccess;
    CMAXB_TRY(fe->prof.run(CMAXB_K_FE_EVAL_FUSED, s, true, [&] {
      le = (fe->taps.r == 4)
          ? cudaLaunchCooperativeKernel((void*)fe_eval_megakernel<4>, dim3(fe->mega_grid), dim3(kMegaThreads), args, smem, s)
          : cudaLaunchCooperativeKernel((void*)fe_eval_megakernel<-1>, dim3(fe->mega_grid), dim3(kMegaThreads), args, smem, s);
    }));
    if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("fused evaluation launch: ") + cudaGetErrorString(le));
  }
  if (clear_next) fe->quad_dirty[oth] = 0;
  if (fe->n > 0) fe->quad_dirty[cur] = k;
  fe->quad_cur = oth;
  fe->inflight.push_back(FeInflight{k, want_grad != 0, true, slot, fe->seq, fe->x_on});
  fe->last_k = k; fe->last_grad = want_grad != 0; fe->pending = true; fe->last_mega = true;
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_eval_launch(cmaxb_fe* fe, const double* omegas, int k, int want_grad) {
  if (!fe || !omegas) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!fe->have_packet) return set_error(CMAXB_ERR_STATE, "no event packet: call cmaxb_fe_set_packet first");
  if (k < 1 || k > fe->kmax) return set_error(CMAXB_ERR_INVALID, "k exceeds cfg.max_hypotheses");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  if (!(want_grad && fe->cfg.grad_mode == CMAXB_GRAD_DENSE) && !fe->force_multi_kernel)
    return fe_eval_launch_fused(fe, omegas, k, want_grad);
  // multi-kernel pipeline (DENSE gradients, A/B runs): one evaluation outstanding at a time
  if (!fe->inflight.empty())
    return set_error(CMAXB_ERR_STATE, "multi-kernel evaluation: fetch the outstanding evaluation first");
  CMAXB_TRY(fe_upload_omegas(fe, omegas, k));
  fe->last_mega = false;
  cudaStream_t s = fe->stream;
  const int W = fe->cfg.width, H = fe->cfg.height, measure = fe->cfg.contrast_measure;
  if (want_grad && fe->cfg.grad_mode == CMAXB_GRAD_DENSE) {
    CMAXB_TRY(fe_run_scatter_dense(fe, k));
    const SrcPlane4 src{fe->d_img4, fe->A};
    const ReduceOut ro{fe->d_acc, fe->d_ticket, fe->d_result, fe->d_mean};
    cudaError_t le = cudaSuccess;
    CMAXB_TRY(fe->prof.run(CMAXB_K_BLUR_REDUCE, s, true, [&] {
      le = launch_blur_reduce<4, SrcPlane4, false>(s, k, src, W, H, fe->taps, nullptr, 0, ro, measure);
    }));
    if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("blur_reduce launch: ") + cudaGetErrorString(le));
  } else {
    CMAXB_TRY(fe_run_scatter_value(fe, k));
    CMAXB_TRY(fe_run_value_image(fe, k, fe->taps, want_grad != 0));
    if (want_grad) {
      if (!fe->d_GQ) CMAXB_TRY(dev_alloc(&fe->d_GQ, (size_t)fe->kmax * fe->A));
      cudaError_t le = cudaSuccess;
      CMAXB_TRY(fe->prof.run(CMAXB_K_ADJOINT_BLUR, s, true, [&] {
        le = launch_adjoint_blur<true>(s, k, fe->d_blur1, fe->A, W, H, fe->taps, fe->d_mean, measure, nullptr, fe->d_GQ);
      }));
      if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("adjoint_blur launch: ") + cudaGetErrorString(le));
      const FeGeom g = fe_geom(fe);
      CMAXB_TRY(fe->prof.run(CMAXB_K_FE_GATHER, s, true, [&] {
        fe_gather_kernel<true><<<fe_event_grid(fe, k), kFeThreads, 0, s>>>(g, fe->d_omegas, nullptr, fe->d_GQ, fe->A, fe->d_gacc, fe->d_ticket2, fe->d_result);
      }));
    }
  }
  CMAXB_CUDA_TRY(cudaMemcpyAsync(fe->h_result, fe->d_result, sizeof(double) * 4 * k, cudaMemcpyDeviceToHost, s));
  fe->inflight.push_back(FeInflight{k, want_grad != 0, false, 0, 0ull, false});
  fe->last_k = k; fe->last_grad = want_grad != 0; fe->pending = true;
  return CMAXB_OK;
}

// waits for the OLDEST outstanding launch and pops it; *out = its record
static int fe_wait_oldest(cmaxb_fe* fe, FeInflight* out) {
  if (fe->inflight.empty()) return set_error(CMAXB_ERR_STATE, "no evaluation launched");
  const FeInflight f = fe->inflight.front();
  if (f.mega) {
    // The fused kernel publishes its results in mapped pinned memory and then stores its sequence
    // number (monotonic): spin on that word (~1 us) instead of paying the driver's stream-synchronise
    // latency; check the stream now and then so that a faulted kernel cannot hang the caller.
    volatile unsigned long long* done = fe->h_done;
    unsigned long long spins = 0;
    while (*done < f.seq) {
      if ((++spins & 0x3fff) == 0) {
        cudaError_t q = cudaStreamQuery(fe->stream);
        if (q == cudaSuccess) break;                       // finished (flag write raced the query) or faulted
        if (q != cudaErrorNotReady) { fe->inflight.clear(); return set_error(CMAXB_ERR_CUDA, std::string("fused evaluation kernel: ") + cudaGetErrorString(q)); }
      }
    }
    if (*done < f.seq) CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->stream));
    if (*done < f.seq) { fe->inflight.clear(); return set_error(CMAXB_ERR_CUDA, "fused evaluation kernel finished without publishing its result"); }
  } else {
    CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->stream));
  }
  fe->inflight.pop_front();
  if (fe->inflight.empty() && (!f.mega || *fe->h_done >= fe->seq)) fe->pending = false;   // nothing of ours is left on the stream
  *out = f;
  return fe_check_packet_flags(fe);
}

extern "C" int cmaxb_fe_eval_fetch(cmaxb_fe* fe, double* contrasts, double* grads3k) {
  if (!fe || !contrasts) return set_error(CMAXB_ERR_INVALID, "null argument");
  FeInflight f;
  CMAXB_TRY(fe_wait_oldest(fe, &f));
  if (f.xchg && *fe->h_xerr) return set_error(CMAXB_ERR_CUDA, "result exchange: a peer's rows did not arrive (timeout)");
  const double* res = f.mega ? fe->h_mega_result + (long long)f.slot * fe->kmax * 4 : fe->h_result;
  for (int h = 0; h < f.k; ++h) {
    contrasts[h] = res[4 * h];
    if (grads3k && f.grad)
      for (int c = 0; c < 3; ++c) grads3k[3 * h + c] = res[4 * h + 1 + c];
  }
  return CMAXB_OK;
}

// ---- fused result exchange over peer memory -----------------------------------------------------------
extern "C" int cmaxb_fe_exchange_init(cmaxb_fe* fe, int world, int rank, void* handle64_out) {
  if (!fe || !handle64_out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (world < 1 || world > kXMaxWorld || rank < 0 || rank >= world) return set_error(CMAXB_ERR_INVALID, "bad world / rank (at most 8 ranks)");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  CMAXB_TRY(fe_drain_stream(fe));
  if (fe->x_local) return set_error(CMAXB_ERR_STATE, "exchange already initialised");
  const size_t need = sizeof(ulonglong2) * 2 * (size_t)world * fe->kmax * 4;
  size_t bytes = (size_t)2 << 20;            // a whole 2 MiB block: the IPC handle exports nothing else
  while (bytes < need) bytes <<= 1;
  CMAXB_CUDA_TRY(cudaMalloc((void**)&fe->x_local, bytes));
  CMAXB_CUDA_TRY(cudaMemset(fe->x_local, 0, bytes));
  CMAXB_CUDA_TRY(cudaDeviceSynchronize());
  fe->x_bytes = bytes;
  cudaIpcMemHandle_t h;
  CMAXB_CUDA_TRY(cudaIpcGetMemHandle(&h, fe->x_local));
  std::memcpy(handle64_out, &h, 64);
  fe->x_world = world; fe->x_rank = rank;
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_exchange_connect(cmaxb_fe* fe, const void* handles, double* gathered_dev) {
  if (!fe || !handles) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!fe->x_local) return set_error(CMAXB_ERR_STATE, "call cmaxb_fe_exchange_init first");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  CMAXB_TRY(fe_drain_stream(fe));
  for (int r = 0; r < fe->x_world; ++r) {
    if (r == fe->x_rank) { fe->x_peer[r] = fe->x_local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + 64 * r, 64);
    void* ptr = nullptr;
    CMAXB_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    fe->x_peer[r] = (ulonglong2*)ptr;
  }
  if (!fe->h_xall) {
    CMAXB_CUDA_TRY(cudaHostAlloc((void**)&fe->h_xall, sizeof(double) * 4 * (size_t)fe->kmax * fe->x_world * kFeRing, cudaHostAllocMapped));
    CMAXB_CUDA_TRY(cudaHostGetDevicePointer((void**)&fe->d_xall, fe->h_xall, 0));
    CMAXB_CUDA_TRY(cudaHostAlloc((void**)&fe->h_xerr, sizeof(unsigned int) * 4, cudaHostAllocMapped));
    CMAXB_CUDA_TRY(cudaHostGetDevicePointer((void**)&fe->d_xerr, fe->h_xerr, 0));
    fe->h_xerr[0] = 0;
  }
  fe->x_all_dev = gathered_dev;
  fe->x_on = fe->x_world > 1;
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_exchange_close(cmaxb_fe* fe) {
  if (!fe) return set_error(CMAXB_ERR_INVALID, "null argument");
  cudaSetDevice(fe->device);
  if (fe->stream) cudaStreamSynchronize(fe->stream);
  fe->pending = false;
  fe->x_on = false;
  for (int r = 0; r < fe->x_world; ++r) {
    if (r != fe->x_rank && fe->x_peer[r]) cudaIpcCloseMemHandle(fe->x_peer[r]);
    fe->x_peer[r] = nullptr;
  }
  // the local buffer stays allocated until destroy: a peer may still have it mapped
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_eval_fetch_all(cmaxb_fe* fe, double* rows) {
  if (!fe || !rows) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (fe->inflight.empty()) return set_error(CMAXB_ERR_STATE, "no evaluation launched");
  if (!fe->inflight.front().xchg) return set_error(CMAXB_ERR_STATE, "the oldest evaluation was launched without a result exchange");
  FeInflight f;
  CMAXB_TRY(fe_wait_oldest(fe, &f));
  if (*fe->h_xerr) return set_error(CMAXB_ERR_CUDA, "result exchange: a peer's rows did not arrive (timeout)");
  const double* all = fe->h_xall + (long long)f.slot * fe->x_world * fe->kmax * 4;
  std::memcpy(rows, all, sizeof(double) * 4 * (size_t)f.k * fe->x_world);
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_eval_batch(cmaxb_fe* fe, const double* omegas, int k, double* contrasts, double* grads3k) {
  if (fe && !fe->inflight.empty()) return set_error(CMAXB_ERR_STATE, "outstanding cmaxb_fe_eval_launch calls: fetch them first");
  CMAXB_TRY(cmaxb_fe_eval_launch(fe, omegas, k, grads3k != nullptr));
  return cmaxb_fe_eval_fetch(fe, contrasts, grads3k);
}

extern "C" int cmaxb_fe_eval(cmaxb_fe* fe, const double omega[3], double* contrast, double* grad3) {
  return cmaxb_fe_eval_batch(fe, omega, 1, contrast, grad3);
}

extern "C" int cmaxb_fe_get_iwe(cmaxb_fe* fe, const double omega[3], int blurred, float* out) {
  if (!fe || !omega || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!fe->have_packet) return set_error(CMAXB_ERR_STATE, "no event packet");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  CMAXB_TRY(fe_upload_omegas(fe, omega, 1));
  CMAXB_TRY(fe_run_scatter_value(fe, 1));
  Taps t0{}; t0.r = 0; t0.w[0] = 1.0f;   // raw image: the same kernel with a radius-0 filter
  CMAXB_TRY(fe_run_value_image(fe, 1, (blurred && fe->taps.r > 0) ? fe->taps : t0, true));
  cudaStream_t s = fe->stream;
  CMAXB_CUDA_TRY(cudaMemcpyAsync(out, fe->d_blur1, sizeof(float) * fe->A, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_get_deriv(cmaxb_fe* fe, const double omega[3], int blurred, float* out) {
  if (!fe || !omega || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!fe->have_packet) return set_error(CMAXB_ERR_STATE, "no event packet");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  CMAXB_TRY(fe_upload_omegas(fe, omega, 1));
  CMAXB_TRY(fe_run_scatter_dense(fe, 1));
  cudaStream_t s = fe->stream;
  const float4* src_ptr = fe->d_img4;
  if (blurred && fe->taps.r > 0) {
    if (!fe->d_blur4) CMAXB_TRY(dev_alloc(&fe->d_blur4, (size_t)fe->A));
    const SrcPlane4 src{fe->d_img4, fe->A};
    const ReduceOut ro{fe->d_acc, fe->d_ticket, fe->d_result, fe->d_mean};
    const int W = fe->cfg.width, H = fe->cfg.height;
    cudaError_t le = cudaSuccess;
    CMAXB_TRY(fe->prof.run(CMAXB_K_BLUR_REDUCE, s, true, [&] {
      le = launch_blur_reduce<4, SrcPlane4, true>(s, 1, src, W, H, fe->taps, fe->d_blur4, fe->A, ro, fe->cfg.contrast_measure);
    }));
    if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("blur_reduce launch: ") + cudaGetErrorString(le));
    src_ptr = fe->d_blur4;
  }
  std::vector<float4> tmp((size_t)fe->A);
  CMAXB_CUDA_TRY(cudaMemcpyAsync(tmp.data(), src_ptr, sizeof(float4) * fe->A, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  for (long long i = 0; i < fe->A; ++i) { out[3 * i] = tmp[i].y; out[3 * i + 1] = tmp[i].z; out[3 * i + 2] = tmp[i].w; }
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_get_cells(cmaxb_fe* fe, const double omega[3], int32_t* out) {
  if (!fe || !omega || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!fe->have_packet) return set_error(CMAXB_ERR_STATE, "no event packet");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  if (fe->n == 0) return CMAXB_OK;
  CMAXB_TRY(fe_upload_omegas(fe, omega, 1));
  if ((size_t)fe->n > fe->cells_cap) {
    cudaFree(fe->d_cells); fe->d_cells = nullptr; fe->cells_cap = 0;
    CMAXB_TRY(dev_alloc(&fe->d_cells, (size_t)fe->n));
    fe->cells_cap = (size_t)fe->n;
  }
  cudaStream_t s = fe->stream;
  const FeGeom g = fe_geom(fe);
  CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
    fe_cells_kernel<<<(unsigned)((fe->n + 255) / 256), 256, 0, s>>>(g, fe->d_omegas, fe->d_cells);
  }));
  CMAXB_CUDA_TRY(cudaMemcpyAsync(out, fe->d_cells, sizeof(int) * fe->n, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_profile(cmaxb_fe* fe, int enable) {
  if (!fe) return set_error(CMAXB_ERR_INVALID, "null argument");
  fe->prof.enabled = enable != 0;
  fe->prof.reset();
  return CMAXB_OK;
}
extern "C" int cmaxb_fe_set_result_mirror(cmaxb_fe* fe, double* device_ptr) {
  if (!fe) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (fe->pending) { CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->stream)); fe->pending = false; }
  fe->d_mirror = device_ptr;
  return CMAXB_OK;
}
extern "C" int cmaxb_fe_phase_times(cmaxb_fe* fe, double* us10) {
  if (!fe || !us10) return set_error(CMAXB_ERR_INVALID, "null argument");
  CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->stream));
  for (int i = 0; i < 10; ++i)
    us10[i] = (fe->h_phase && fe->h_phase[i] && fe->h_phase[0]) ? (double)(fe->h_phase[i] - fe->h_phase[0]) * 1e-3 : -1.0;
  return CMAXB_OK;
}
extern "C" int cmaxb_fe_kernel_times(cmaxb_fe* fe, double* ms, uint64_t* launches) {
  if (!fe || !ms || !launches) return set_error(CMAXB_ERR_INVALID, "null argument");
  for (int i = 0; i < CMAXB_K_COUNT; ++i) { ms[i] = fe->prof.ms[i]; launches[i] = fe->prof.launches[i]; }
  return CMAXB_OK;
}

// ---- diagnostics ----------------------------------------------------------------------------------
extern "C" const char* cmaxb_last_error(void) { return g_last_error.c_str(); }
extern "C" int cmaxb_version(void) { return CMAXB_VERSION; }
extern "C" int cmaxb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" uint64_t cmaxb_launch_count(void) { return g_launch_count.load(); }
extern "C" const char* cmaxb_kernel_name(int kind) {
  static const char* names[CMAXB_K_COUNT] = {"zero(memset)", "fe_scatter", "fe_gather", "blur_reduce", "adjoint_blur",
                                             "be_poses", "be_scatter", "be_gather", "be_grad_reduce", "misc",
                                             "fe_eval_fused", "be_eval_fused"};
  return (kind >= 0 && kind < CMAXB_K_COUNT) ? names[kind] : "?";
}
