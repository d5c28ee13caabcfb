// be_capi.cu -- C ABI of the back-end (panoramic CMax bundle adjustment) path, include/cmax_b200.h.
#include "capi_common.cuh"
#include "be_kernels.cuh"
#include "be_xchg.cuh"
#include "image_kernels.cuh"

using namespace cmaxb;

struct cmaxb_be {
  cmaxb_be_cfg cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long long A = 0, SA = 0;
  int N = 2;                 // spline order
  Taps taps{};
  double4* d_lut = nullptr;
  uint4* d_ev = nullptr; size_t ev_cap = 0;
  long long n = 0, n_eff = 0, nb = 0;
  BeBatchTime* d_bt = nullptr; BePose* d_poses = nullptr; double* d_wgrad = nullptr; int* d_idx = nullptr; size_t nb_cap = 0;
  Quat* d_knots0 = nullptr; Quat* d_knots = nullptr; double* d_x = nullptr; double* d_grad = nullptr; size_t knots_cap = 0;
  int* d_seg_lo = nullptr; int* d_seg_hi = nullptr;
  double* h_x = nullptr; double* h_grad = nullptr; size_t hx_cap = 0;
  int n_knots = 0, n_fixed = 0, n_opt = 0;
  long long t0_ns = 0, dt_ns = 0;
  uint32_t tnext_sec = 0, tnext_nsec = 0;
  float* d_igp = nullptr; bool have_igp = false;
  double alpha = 0.0; bool alpha_pending = false;
  float* d_il_old = nullptr; float* d_il_new = nullptr; float* d_blur = nullptr; float* d_G = nullptr;
  float4* d_ilq = nullptr; float4* d_GQ = nullptr;   // corner-split accumulator / adjoint image (event-dense windows)
  bool use_quad = false; bool il_is_quad = false;
  float* d_il_plane = nullptr; bool il_is_plane = false;   // assembled IL (event-sharded evaluation)
  // row-band sharding of the image phases (cmaxb_be_shard_*)
  int sh_world = 0, sh_rank = 0, sh_hb = 0, sh_hl = 0, sh_ce = 0, sh_y0 = 0, sh_y1 = 0, sh_first = 0, sh_rows = 0;
  float* d_sh_send = nullptr; float* d_sh_recv = nullptr; float* d_sh_blur = nullptr; float* d_sh_gband = nullptr;
  float* d_sh_gfull = nullptr; double* d_sh_sums = nullptr; size_t sh_cap = 0;
  int sh_stage = 0; bool sh_grad = false;
  // exchange by the kernels themselves over peer memory (cmaxb_be_exchange_* / cmaxb_be_xeval, be_xchg.cuh)
  char* x_local = nullptr; BeXLayout x_lay{}; BeXPeers x_peers{}; bool x_on = false; unsigned long long x_seq = 0;
  unsigned int* d_dirty = nullptr; unsigned int* d_xticket = nullptr; bool x_invariant = false;
  unsigned long long* h_xfault = nullptr; unsigned long long* d_xfault = nullptr;
  bool split_pending = false; bool split_grad = false; bool end_launched = false; bool end_grad = false;
  // device-resident global map (IG_, IG_update_times_map_)
  float* d_IG = nullptr; unsigned char* d_times = nullptr; unsigned char* d_mask = nullptr;
  float4* d_ca = nullptr; float2* d_cb = nullptr; size_t cache_cap = 0;   // per-event gather cache (24 B)
  long long n_visit = 0; int m_visit = 1;
  float* d_bands = nullptr; float* d_bands_blur = nullptr; size_t bands_cap = 0;
  double* d_acc = nullptr; unsigned int* d_ticket = nullptr; double* d_result = nullptr; double* d_mean = nullptr;
  double* d_bacc = nullptr; unsigned int* d_bticket = nullptr; double* d_bresult = nullptr; double* d_bmean = nullptr; size_t bacc_cap = 0;
  double* d_alpha_sums = nullptr;
  double* h_result = nullptr; double* h_alpha_sums = nullptr;
  double* d_hresult = nullptr; double* d_hgrad = nullptr;   // device aliases of the MAPPED h_result / h_grad (plain evaluations write them directly)
  bool zeroed_by_poses = false;                             // the pose kernel of this evaluation cleared the accumulators
  bool direct_out = false;                                  // this evaluation's kernels write h_result / h_grad themselves
  int* d_flags = nullptr; int* h_flags = nullptr;
  int* d_cells = nullptr; size_t cells_cap = 0;
  bool have_window = false;
  // CUDA-graph replay of a whole evaluation (x upload, poses, zeroing, scatter, blur, adjoint, gather, per-knot reduction,
  // result download): small windows are launch-bound (ten launches + a synchronisation, ~90 us at 9k events); the first
  // evaluation of each kind (value / value + gradient) after set_window runs launch by launch (it allocates and fixes
  // alpha), the second is captured, later ones are one cudaGraphLaunch.  CMAXB_BE_GRAPH=0 turns it off.
  bool graph_on = true;
  cudaGraphExec_t gexec[2] = {nullptr, nullptr};
  int evals_in_window[2] = {0, 0};
  std::vector<double> last_x;   // parameter vector of the most recent cmaxb_be_eval (the reference's IL_old_ / IL_new_ members hold THAT evaluation's images)
  KernelProfiler prof;
};

static BeGeom be_geom(const cmaxb_be* be) {
  BeGeom g;
  g.ev = be->d_ev; g.n_eff = be->n_eff; g.nb = be->nb;
  g.batch_size = be->cfg.batch_size; g.sample_rate = be->cfg.event_sample_rate;
  g.lut = be->d_lut; g.SW = be->cfg.sensor_width; g.SH = be->cfg.sensor_height;
  g.W = be->cfg.pano_width; g.H = be->cfg.pano_height;
  // EquirectangularCamera(pano_size, 360, 180)            equirectangular_camera.h:11-16,64-67
  g.cx = (double)g.W / 2.0; g.cy = (double)g.H / 2.0;
  g.fx = (double)((g.W / 360.0) * 180.0 / 3.1415926535897932384626433832795);
  g.fy = (double)((g.H / 180.0) * 180.0 / 3.1415926535897932384626433832795);
  g.tnext_sec = be->tnext_sec; g.tnext_nsec = be->tnext_nsec;
  g.n_fixed = be->n_fixed; g.Nk = be->N;
  g.m = be->m_visit; g.n_visit = be->n_visit;
  return g;
}

extern "C" int cmaxb_be_create(const cmaxb_be_cfg* cfg, cmaxb_be** out) {
  if (!cfg || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->sensor_width < 1 || cfg->sensor_height < 1 || cfg->pano_width < 4 || cfg->pano_height < 4 || !cfg->lut_xyz ||
      cfg->batch_size <= 0 || cfg->event_sample_rate <= 0)
    return set_error(CMAXB_ERR_INVALID, "bad back-end configuration");
  if (cfg->spline_order != 2 && cfg->spline_order != 4) return set_error(CMAXB_ERR_INVALID, "spline_order must be 2 (linear) or 4 (cubic)");
  if (cfg->grad_mode != CMAXB_GRAD_DENSE && cfg->grad_mode != CMAXB_GRAD_ADJOINT) return set_error(CMAXB_ERR_INVALID, "bad grad_mode");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return set_error(CMAXB_ERR_CUDA, "no CUDA device: libcmax_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return set_error(CMAXB_ERR_INVALID, "bad device ordinal");
  CMAXB_CUDA_TRY(cudaSetDevice(cfg->device));
  cmaxb_be* be = new cmaxb_be();
  be->cfg = *cfg;
  be->cfg.lut_xyz = nullptr;
  be->device = cfg->device;
  be->N = cfg->spline_order;
  be->A = (long long)cfg->pano_width * cfg->pano_height;
  { const char* gr = getenv("CMAXB_BE_GRAPH"); be->graph_on = !(gr && gr[0] == '0'); }
  be->SA = (long long)cfg->sensor_width * cfg->sensor_height;
  int rc = make_taps(cfg->blur_sigma, &be->taps);
  if (rc != CMAXB_OK) { delete be; return rc; }
  auto fail = [&](int code) { cmaxb_be_destroy(be); return code; };
  if (cfg->stream) be->stream = (cudaStream_t)cfg->stream;
  else {
    if (cudaStreamCreateWithFlags(&be->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(set_error(CMAXB_ERR_CUDA, "cudaStreamCreate failed"));
    be->own_stream = true;
  }
  if (be->prof.init() != CMAXB_OK) return fail(CMAXB_ERR_CUDA);
  {
    std::vector<double4> lut((size_t)be->SA);
    for (long long i = 0; i < be->SA; ++i) lut[i] = make_double4(cfg->lut_xyz[3 * i], cfg->lut_xyz[3 * i + 1], cfg->lut_xyz[3 * i + 2], 0.0);
    if (dev_alloc(&be->d_lut, (size_t)be->SA) != CMAXB_OK) return fail(CMAXB_ERR_CUDA);
    if (cudaMemcpy(be->d_lut, lut.data(), sizeof(double4) * be->SA, cudaMemcpyHostToDevice) != cudaSuccess) return fail(set_error(CMAXB_ERR_CUDA, "LUT upload failed"));
  }
  const size_t A = (size_t)be->A;
  bool ok = true;
  ok = ok && dev_alloc(&be->d_il_old, A) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_il_new, A) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_blur, A) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_G, A) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_igp, A) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_acc, (size_t)kNAcc * kMaxImgCtas) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_ticket, 1) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_result, 4) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_mean, 1) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_alpha_sums, 8) == CMAXB_OK;
  ok = ok && dev_alloc(&be->d_flags, 1) == CMAXB_OK;
  if (!ok) return fail(CMAXB_ERR_CUDA);
  ok = ok && cudaHostAlloc((void**)&be->h_result, sizeof(double) * 4, cudaHostAllocMapped) == cudaSuccess;
  ok = ok && cudaHostGetDevicePointer((void**)&be->d_hresult, be->h_result, 0) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&be->h_alpha_sums, sizeof(double) * 8) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&be->h_flags, sizeof(int)) == cudaSuccess;
  ok = ok && cudaMemset(be->d_ticket, 0, sizeof(unsigned)) == cudaSuccess;
  ok = ok && cudaMemset(be->d_result, 0, sizeof(double) * 4) == cudaSuccess;
  if (!ok) return fail(set_error(CMAXB_ERR_CUDA, "back-end buffer allocation failed"));
  *out = be;
  return CMAXB_OK;
}

extern "C" void cmaxb_be_destroy(cmaxb_be* be) {
  if (!be) return;
  cudaSetDevice(be->device);
  if (be->stream) cudaStreamSynchronize(be->stream);
  for (int k = 0; k < 2; ++k) if (be->gexec[k]) cudaGraphExecDestroy(be->gexec[k]);
  cudaFree(be->d_lut); cudaFree(be->d_ev); cudaFree(be->d_bt); cudaFree(be->d_poses); cudaFree(be->d_wgrad); cudaFree(be->d_idx);
  cudaFree(be->d_knots0); cudaFree(be->d_knots); cudaFree(be->d_x); cudaFree(be->d_grad);
  cudaFree(be->d_seg_lo); cudaFree(be->d_seg_hi);
  cudaFree(be->d_igp); cudaFree(be->d_il_old); cudaFree(be->d_il_new); cudaFree(be->d_blur); cudaFree(be->d_G);
  cudaFree(be->d_bands); cudaFree(be->d_bands_blur); cudaFree(be->d_ilq); cudaFree(be->d_GQ);
  cudaFree(be->d_ca); cudaFree(be->d_cb); cudaFree(be->d_il_plane);
  cudaFree(be->x_local); cudaFree(be->d_dirty); cudaFree(be->d_xticket);
  if (be->h_xfault) cudaFreeHost(be->h_xfault);
  cudaFree(be->d_sh_send); cudaFree(be->d_sh_recv); cudaFree(be->d_sh_blur); cudaFree(be->d_sh_gband); cudaFree(be->d_sh_gfull); cudaFree(be->d_sh_sums);
  cudaFree(be->d_IG); cudaFree(be->d_times); cudaFree(be->d_mask);

  cudaFree(be->d_acc); cudaFree(be->d_ticket); cudaFree(be->d_result); cudaFree(be->d_mean);
  cudaFree(be->d_bacc); cudaFree(be->d_bticket); cudaFree(be->d_bresult); cudaFree(be->d_bmean);
  cudaFree(be->d_alpha_sums); cudaFree(be->d_flags); cudaFree(be->d_cells);
  if (be->h_x) cudaFreeHost(be->h_x);
  if (be->h_grad) cudaFreeHost(be->h_grad);
  if (be->h_result) cudaFreeHost(be->h_result);
  if (be->h_alpha_sums) cudaFreeHost(be->h_alpha_sums);
  if (be->h_flags) cudaFreeHost(be->h_flags);
  be->prof.destroy();
  if (be->own_stream && be->stream) cudaStreamDestroy(be->stream);
  delete be;
}

extern "C" int cmaxb_be_set_window(cmaxb_be* be, const cmaxb_be_window* w) {
  if (!be || !w || (!w->events && w->n_events > 0) || !w->knots_xyzw) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (w->n_knots < be->N || w->n_fixed < 0 || w->n_fixed > w->n_knots || w->dt_ns <= 0)
    return set_error(CMAXB_ERR_INVALID, "bad window description");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  cudaStream_t s = be->stream;
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  be->have_window = false;
  be->il_is_plane = false; be->split_pending = false; be->end_launched = false; be->sh_stage = 0;
  for (int k = 0; k < 2; ++k) {        // buffers may move and the window geometry is baked into the captured launches
    if (be->gexec[k]) { cudaGraphExecDestroy(be->gexec[k]); be->gexec[k] = nullptr; }
    be->evals_in_window[k] = 0;
  }
  be->last_x.clear();
  const long long n = (long long)w->n_events, bs = be->cfg.batch_size;
  // the reference loop `for (beg = begin; beg < end-1; beg += bs)` never visits a trailing batch
  // of exactly one event (event_pano_warper.cpp:188-196)
  const long long n_eff = (n >= 1 && (n - 1) % bs == 0) ? n - 1 : n;
  const long long nb = (n_eff + bs - 1) / bs;
  be->n = n; be->n_eff = n_eff; be->nb = nb;
  {
    const long long sr = be->cfg.event_sample_rate;
    be->m_visit = (int)((bs + sr - 1) / sr);
    const long long len_last = nb > 0 ? n_eff - (nb - 1) * bs : 0;
    be->n_visit = nb > 0 ? (nb - 1) * be->m_visit + (len_last + sr - 1) / sr : 0;
  }
  // Event-dense windows accumulate IL in a corner-split float4 image (1 vector reduction per event,
  // 4x the image bytes); sparse windows on a big panorama keep float planes (4 reductions per event).
  be->use_quad = (2 * n_eff >= be->A);
  if (be->use_quad && !be->d_ilq) {
    CMAXB_TRY(dev_alloc(&be->d_ilq, (size_t)be->A));
    CMAXB_TRY(dev_alloc(&be->d_GQ, (size_t)be->A));
  }
  if ((size_t)n > be->ev_cap) {
    cudaFree(be->d_ev); be->d_ev = nullptr; be->ev_cap = 0;
    CMAXB_TRY(dev_alloc(&be->d_ev, (size_t)n));
    be->ev_cap = (size_t)n;
  }
  if ((size_t)nb > be->nb_cap) {
    cudaFree(be->d_bt); cudaFree(be->d_poses); cudaFree(be->d_wgrad); cudaFree(be->d_idx);
    be->d_bt = nullptr; be->d_poses = nullptr; be->d_wgrad = nullptr; be->d_idx = nullptr; be->nb_cap = 0;
    CMAXB_TRY(dev_alloc(&be->d_idx, (size_t)nb));
    CMAXB_TRY(dev_alloc(&be->d_bt, (size_t)nb));
    CMAXB_TRY(dev_alloc(&be->d_poses, (size_t)nb));
    CMAXB_TRY(dev_alloc(&be->d_wgrad, (size_t)nb * 12));
    be->nb_cap = (size_t)nb;
  }
  if ((size_t)w->n_knots > be->knots_cap) {
    cudaFree(be->d_knots0); cudaFree(be->d_knots); cudaFree(be->d_x); cudaFree(be->d_grad);
    cudaFree(be->d_seg_lo); cudaFree(be->d_seg_hi); be->d_seg_lo = be->d_seg_hi = nullptr;
    if (be->h_x) cudaFreeHost(be->h_x);
    if (be->h_grad) cudaFreeHost(be->h_grad);
    be->h_x = be->h_grad = nullptr; be->knots_cap = 0;
    const size_t K = (size_t)w->n_knots;
    CMAXB_TRY(dev_alloc(&be->d_knots0, K));
    CMAXB_TRY(dev_alloc(&be->d_knots, K));
    CMAXB_TRY(dev_alloc(&be->d_x, 3 * K));
    CMAXB_TRY(dev_alloc(&be->d_grad, 3 * K));
    CMAXB_TRY(dev_alloc(&be->d_seg_lo, K));
    CMAXB_TRY(dev_alloc(&be->d_seg_hi, K));
    CMAXB_CUDA_TRY(cudaMallocHost((void**)&be->h_x, sizeof(double) * 3 * K));
    CMAXB_CUDA_TRY(cudaHostAlloc((void**)&be->h_grad, sizeof(double) * 3 * K, cudaHostAllocMapped));
    CMAXB_CUDA_TRY(cudaHostGetDevicePointer((void**)&be->d_hgrad, be->h_grad, 0));
    be->knots_cap = K;
  }
  be->n_knots = w->n_knots; be->n_fixed = w->n_fixed; be->n_opt = w->n_knots - w->n_fixed;
  be->t0_ns = w->t0_ns; be->dt_ns = w->dt_ns;
  be->tnext_sec = w->tnext_sec; be->tnext_nsec = w->tnext_nsec;
  static_assert(sizeof(Quat) == 4 * sizeof(double), "Quat must be 4 packed doubles (x,y,z,w)");
  CMAXB_CUDA_TRY(cudaMemcpyAsync(be->d_knots0, w->knots_xyzw, sizeof(Quat) * w->n_knots, cudaMemcpyHostToDevice, s));
  if (w->IGp) {
    CMAXB_CUDA_TRY(cudaMemcpyAsync(be->d_igp, w->IGp, sizeof(float) * be->A, cudaMemcpyHostToDevice, s));
    be->have_igp = true;
  } else {
    CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_igp, 0, sizeof(float) * be->A, s));
    be->have_igp = false;
  }
  if (std::isnan(w->alpha)) { be->alpha = 0.0; be->alpha_pending = true; }
  else { be->alpha = w->alpha; be->alpha_pending = false; }
  if (n > 0) {
    CMAXB_CUDA_TRY(cudaMemcpyAsync(be->d_ev, w->events, sizeof(cmaxb_event) * n, cudaMemcpyHostToDevice, s));
    CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_flags, 0, sizeof(int), s));
    const uint4* ev = be->d_ev; int* flags = be->d_flags;
    const int SW = be->cfg.sensor_width, SH = be->cfg.sensor_height;
    CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
      validate_events_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ev, n, SW, SH, flags);
    }));
    if (nb > 0) {
      CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
        be_batch_time_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(ev, n, n_eff, (int)bs, nb, be->t0_ns, be->dt_ns,
                                                                             be->n_knots, be->N, be->d_bt, flags);
      }));
    }
    CMAXB_CUDA_TRY(cudaMemcpyAsync(be->h_flags, be->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s));
    CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
    if (!(*be->h_flags & 4) && nb > 0) {   // segment indices are valid: bounding batch run of every segment
      CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_seg_lo, 0x7f, sizeof(int) * be->n_knots, s));
      CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_seg_hi, 0, sizeof(int) * be->n_knots, s));
      CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
        be_segment_ranges_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(be->d_bt, nb, be->d_seg_lo, be->d_seg_hi);
      }));
      CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
    }
    if (*be->h_flags & 2) return set_error(CMAXB_ERR_EVENT_RANGE, "event pixel outside the sensor");
    if (*be->h_flags & 4) return set_error(CMAXB_ERR_SPLINE_RANGE, "batch time outside the spline's valid range");
  } else {
    CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  }
  be->have_window = true;
  return CMAXB_OK;
}

static unsigned be_warp_grid(const cmaxb_be* be) {
  long long blocks = (be->nb + kBeWarps - 1) / kBeWarps;
  const long long cap = 148LL * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// x -> updated knots -> per-batch pose table
static int be_fill_x(cmaxb_be* be, const double* x, int n) {
  if (x && n != 3 * be->n_opt) return set_error(CMAXB_ERR_INVALID, "x must have 3*(n_knots-n_fixed) entries");
  for (int i = 0; i < 3 * be->n_opt; ++i) be->h_x[i] = x ? x[i] : 0.0;
  return CMAXB_OK;
}
// zero_acc: the pose kernel also clears the accumulators of the scatter that follows (be_run_scatter(..., zeroed = true))
static int be_enqueue_poses(cmaxb_be* be, bool want_grad, bool zero_acc = false) {
  cudaStream_t s = be->stream;
  if (be->n_opt > 0) CMAXB_CUDA_TRY(cudaMemcpyAsync(be->d_x, be->h_x, sizeof(double) * 3 * be->n_opt, cudaMemcpyHostToDevice, s));
  be->zeroed_by_poses = false;
  if (be->nb > 0) {
    unsigned grid = (unsigned)((be->nb + 127) / 128);
    float4* za = nullptr; float4* zb = nullptr; long long na = 0, nbz = 0;
    if (zero_acc && (be->A & 3) == 0) {
      if (be->use_quad) { za = be->d_ilq; na = be->A; }
      else { za = reinterpret_cast<float4*>(be->d_il_old); na = be->A / 4; zb = reinterpret_cast<float4*>(be->d_il_new); nbz = be->A / 4; }
      grid = std::max(grid, 148u * 8u);
      be->zeroed_by_poses = true;
    }
    CMAXB_TRY(be->prof.run(CMAXB_K_BE_POSES, s, true, [&] {
      if (be->N == 2) be_pose_kernel<2><<<grid, 128, 0, s>>>(be->d_knots0, be->d_x, be->n_fixed, be->d_bt, be->nb, want_grad, be->d_poses, be->d_idx, za, na, zb, nbz);
      else be_pose_kernel<4><<<grid, 128, 0, s>>>(be->d_knots0, be->d_x, be->n_fixed, be->d_bt, be->nb, want_grad, be->d_poses, be->d_idx, za, na, zb, nbz);
    }));
  }
  return CMAXB_OK;
}
static int be_run_poses(cmaxb_be* be, const double* x, int n, bool want_grad) {
  CMAXB_TRY(be_fill_x(be, x, n));
  return be_enqueue_poses(be, want_grad);
}

// updateAlpha (event_pano_warper.cpp:134-165) from the current IL; il_old may be an assembled plane (il_new = null)
static int be_run_alpha(cmaxb_be* be, const float* il_old, const float* il_new, const float4* il_quad) {
  cudaStream_t s = be->stream;
  CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_alpha_sums, 0, sizeof(double) * 8, s));
  CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
    be_alpha_sums_kernel<<<148 * 4, 256, 0, s>>>(be->d_igp, il_old, il_new, il_quad, be->cfg.pano_width, be->A, be->d_alpha_sums);
  }));
  CMAXB_CUDA_TRY(cudaMemcpyAsync(be->h_alpha_sums, be->d_alpha_sums, sizeof(double) * 5, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  const double* v = be->h_alpha_sums;
  if (v[4] < 1.0) be->alpha = 0.0;                                              // countNonZero(IGp_) < 1
  else be->alpha = (v[3] / v[2]) / (v[1] / v[0]);
  be->alpha_pending = false;
  return CMAXB_OK;
}

static unsigned be_event_grid(const cmaxb_be* be) {
  long long blocks = (be->n_visit + kBeThreads - 1) / kBeThreads;
  const long long cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

static int be_run_scatter(cmaxb_be* be, bool allow_quad, bool want_cache = false, bool defer_alpha = false, unsigned int* dirty = nullptr) {
  cudaStream_t s = be->stream;
  const bool quad = dirty ? true : (allow_quad && be->use_quad);
  const int dirty_ntx = (be->cfg.pano_width + kBeXTile - 1) / kBeXTile;
  const int dirty_ntiles = dirty_ntx * ((be->cfg.pano_height + kBeXTile - 1) / kBeXTile);
  if (!dirty) be->x_invariant = false;      // the accumulator is about to hold votes the peer exchange's tile flags do not know of
  if (want_cache && (size_t)be->n_visit > be->cache_cap) {
    cudaFree(be->d_ca); cudaFree(be->d_cb);
    be->d_ca = nullptr; be->d_cb = nullptr; be->cache_cap = 0;
    CMAXB_TRY(dev_alloc(&be->d_ca, (size_t)be->n_visit));
    CMAXB_TRY(dev_alloc(&be->d_cb, (size_t)be->n_visit));
    be->cache_cap = (size_t)be->n_visit;
  }
  const BeCache cache{be->d_ca, be->d_cb};
  be->il_is_quad = quad;
  be->il_is_plane = false;   // a fresh scatter supersedes the assembled plane of an earlier sharded evaluation (eval_begin re-sets it)
  const bool zeroed = be->zeroed_by_poses && allow_quad;      // (the pose kernel cleared what allow_quad = true selects)
  be->zeroed_by_poses = false;
  CMAXB_TRY(be->prof.run(CMAXB_K_ZERO, s, false, [&] {
    if (dirty || zeroed) return;             // peer exchange: the caller cleaned the dirty tiles; plain evaluation: the pose kernel did
    if (quad) cudaMemsetAsync(be->d_ilq, 0, sizeof(float4) * be->A, s);
    else {
      cudaMemsetAsync(be->d_il_old, 0, sizeof(float) * be->A, s);
      cudaMemsetAsync(be->d_il_new, 0, sizeof(float) * be->A, s);
    }
  }));
  if (be->nb > 0) {
    const BeGeom g = be_geom(be);
    CMAXB_TRY(be->prof.run(CMAXB_K_BE_SCATTER, s, true, [&] {
      const unsigned grid = be_event_grid(be);
      if (quad) {
        if (want_cache) be_scatter_kernel<2, true><<<grid, kBeThreads, 0, s>>>(g, be->d_poses, nullptr, nullptr, be->d_ilq, cache, dirty, dirty_ntx, dirty_ntiles);
        else be_scatter_kernel<2, false><<<grid, kBeThreads, 0, s>>>(g, be->d_poses, nullptr, nullptr, be->d_ilq, cache, dirty, dirty_ntx, dirty_ntiles);
      } else {
        if (want_cache) be_scatter_kernel<0, true><<<grid, kBeThreads, 0, s>>>(g, be->d_poses, be->d_il_old, be->d_il_new, nullptr, cache);
        else be_scatter_kernel<0, false><<<grid, kBeThreads, 0, s>>>(g, be->d_poses, be->d_il_old, be->d_il_new, nullptr, cache);
      }
    }));
  }
  // first evaluation of a window with alpha unspecified: updateAlpha            (:201-210)
  if (be->alpha_pending && !defer_alpha) CMAXB_TRY(be_run_alpha(be, be->d_il_old, be->d_il_new, quad ? be->d_ilq : nullptr));
  return CMAXB_OK;
}


// blur(I) + contrast; leaves the blurred image in d_blur and its mean in d_mean
static int be_run_image(cmaxb_be* be, const Taps& taps) {
  cudaStream_t s = be->stream;
  ReduceOut ro{be->d_acc, be->d_ticket, be->d_result, be->d_mean};
  if (be->direct_out) ro.host_result = be->d_hresult;
  const int W = be->cfg.pano_width, H = be->cfg.pano_height;
  const float* igp = be->have_igp ? be->d_igp : nullptr;
  cudaError_t le = cudaSuccess;
  CMAXB_TRY(be->prof.run(CMAXB_K_BLUR_REDUCE, s, true, [&] {
    if (be->il_is_plane) {
      const SrcBePlane src{be->d_il_plane, igp, (float)be->alpha};
      le = launch_blur_reduce<1, SrcBePlane, true>(s, 1, src, W, H, taps, be->d_blur, 0, ro, be->cfg.contrast_measure);
    } else if (be->il_is_quad) {
      const SrcBeQuad src{SrcQuad{be->d_ilq, 0}, igp, (float)be->alpha};
      le = launch_blur_reduce<1, SrcBeQuad, true>(s, 1, src, W, H, taps, be->d_blur, 0, ro, be->cfg.contrast_measure);
    } else {
      const SrcBeI src{be->d_il_old, be->d_il_new, igp, (float)be->alpha};
      le = launch_blur_reduce<1, SrcBeI, true>(s, 1, src, W, H, taps, be->d_blur, 0, ro, be->cfg.contrast_measure);
    }
  }));
  if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("blur_reduce launch: ") + cudaGetErrorString(le));
  return CMAXB_OK;
}

static int be_ensure_bands(cmaxb_be* be) {
  const size_t need = (size_t)3 * be->n_opt * (size_t)be->A;
  if (need > be->bands_cap) {
    cudaFree(be->d_bands); cudaFree(be->d_bands_blur);
    be->d_bands = be->d_bands_blur = nullptr; be->bands_cap = 0;
    CMAXB_TRY(dev_alloc(&be->d_bands, need));
    CMAXB_TRY(dev_alloc(&be->d_bands_blur, need));
    be->bands_cap = need;
  }
  const size_t P = (size_t)3 * be->n_opt;
  if (P > be->bacc_cap) {
    cudaFree(be->d_bacc); cudaFree(be->d_bticket); cudaFree(be->d_bresult); cudaFree(be->d_bmean);
    CMAXB_TRY(dev_alloc(&be->d_bacc, P * kNAcc * kMaxImgCtas));
    CMAXB_TRY(dev_alloc(&be->d_bticket, P));
    CMAXB_TRY(dev_alloc(&be->d_bresult, P * 4));
    CMAXB_TRY(dev_alloc(&be->d_bmean, P));
    CMAXB_CUDA_TRY(cudaMemset(be->d_bticket, 0, sizeof(unsigned) * P));
    be->bacc_cap = P;
  }
  return CMAXB_OK;
}

// dense bands at the current pose table: scatter, then blur into d_bands_blur
static int be_run_bands(cmaxb_be* be, bool blur) {
  CMAXB_TRY(be_ensure_bands(be));
  cudaStream_t s = be->stream;
  const int P = 3 * be->n_opt;
  if (P == 0) return CMAXB_OK;
  CMAXB_TRY(be->prof.run(CMAXB_K_ZERO, s, false, [&] { cudaMemsetAsync(be->d_bands, 0, sizeof(float) * (size_t)P * be->A, s); }));
  if (be->nb > 0) {
    const BeGeom g = be_geom(be);
    CMAXB_TRY(be->prof.run(CMAXB_K_BE_SCATTER, s, true, [&] {
      if (be->N == 2) be_scatter_bands_kernel<2><<<be_warp_grid(be), kBeThreads, 0, s>>>(g, be->d_poses, be->d_bands, be->A, P);
      else be_scatter_bands_kernel<4><<<be_warp_grid(be), kBeThreads, 0, s>>>(g, be->d_poses, be->d_bands, be->A, P);
    }));
  }
  if (blur) {
    const SrcPlane src{be->d_bands, be->A};
    const ReduceOut ro{be->d_bacc, be->d_bticket, be->d_bresult, be->d_bmean};
    const int W = be->cfg.pano_width, H = be->cfg.pano_height;
    // blockIdx.z is limited to 65535 planes
    cudaError_t le = cudaSuccess;
    CMAXB_TRY(be->prof.run(CMAXB_K_BLUR_REDUCE, s, true, [&] {
      le = launch_blur_reduce<1, SrcPlane, true>(s, P, src, W, H, be->taps, be->d_bands_blur, be->A, ro, be->cfg.contrast_measure);
    }));
    if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("blur_reduce launch: ") + cudaGetErrorString(le));
  }
  return CMAXB_OK;
}

// adjoint gather over this handle's events + per-knot reduction -> d_grad (G: float plane, or GQ: corner-packed cells)
static int be_gather_launch(cmaxb_be* be, const float* G, const float4* GQ) {
  cudaStream_t s = be->stream;
  const int W = be->cfg.pano_width, H = be->cfg.pano_height;
  const bool quad = GQ != nullptr;
  if (be->nb > 0) {
    const BeGeom g = be_geom(be);
    const BeCache cache{be->d_ca, be->d_cb};
    CMAXB_TRY(be->prof.run(CMAXB_K_BE_GATHER, s, true, [&] {
      if (be->N == 2) {
        if (quad) be_gather_kernel<2, true><<<be_warp_grid(be), kBeThreads, 0, s>>>(g, be->d_poses, nullptr, GQ, cache, be->d_wgrad);
        else be_gather_kernel<2, false><<<be_warp_grid(be), kBeThreads, 0, s>>>(g, be->d_poses, G, nullptr, cache, be->d_wgrad);
      } else {
        if (quad) be_gather_kernel<4, true><<<be_warp_grid(be), kBeThreads, 0, s>>>(g, be->d_poses, nullptr, GQ, cache, be->d_wgrad);
        else be_gather_kernel<4, false><<<be_warp_grid(be), kBeThreads, 0, s>>>(g, be->d_poses, G, nullptr, cache, be->d_wgrad);
      }
    }));
  }
  const double inv_np = 1.0 / ((double)W * (double)H);
  CMAXB_TRY(be->prof.run(CMAXB_K_BE_GRAD_REDUCE, s, true, [&] {
    double* hg = be->direct_out ? be->d_hgrad : nullptr;
    if (be->N == 2) be_grad_reduce_kernel<2><<<be->n_opt, kBeReduceThreads, 0, s>>>(be->d_idx, be->d_seg_lo, be->d_seg_hi, be->d_wgrad, be->nb, be->n_fixed, inv_np, be->d_grad, hg);
    else be_grad_reduce_kernel<4><<<be->n_opt, kBeReduceThreads, 0, s>>>(be->d_idx, be->d_seg_lo, be->d_seg_hi, be->d_wgrad, be->nb, be->n_fixed, inv_np, be->d_grad, hg);
  }));
  return CMAXB_OK;
}

// image -> contrast (+ gradient) on the current IL (quad / planes / assembled plane); results to the host
// queue blur + contrast (+ adjoint image, gather, per-knot reduction): results stay on the device
static int be_finish_launch(cmaxb_be* be, bool want_grad) {
  cudaStream_t s = be->stream;
  const int P = 3 * be->n_opt;
  CMAXB_TRY(be_run_image(be, be->taps));
  if (want_grad && P > 0) {
    const int W = be->cfg.pano_width, H = be->cfg.pano_height;
    if (be->cfg.grad_mode == CMAXB_GRAD_ADJOINT) {
      const bool quad = be->use_quad;
      cudaError_t le = cudaSuccess;
      CMAXB_TRY(be->prof.run(CMAXB_K_ADJOINT_BLUR, s, true, [&] {
        le = quad ? launch_adjoint_blur<true>(s, 1, be->d_blur, 0, W, H, be->taps, be->d_mean, be->cfg.contrast_measure, nullptr, be->d_GQ)
                  : launch_adjoint_blur<false>(s, 1, be->d_blur, 0, W, H, be->taps, be->d_mean, be->cfg.contrast_measure, be->d_G, nullptr);
      }));
      if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("adjoint_blur launch: ") + cudaGetErrorString(le));
      CMAXB_TRY(be_gather_launch(be, quad ? nullptr : be->d_G, quad ? be->d_GQ : nullptr));
    } else {
      CMAXB_TRY(be_run_bands(be, true));
      CMAXB_TRY(be->prof.run(CMAXB_K_BE_GRAD_REDUCE, s, true, [&] {
        be_band_reduce_kernel<<<P, 256, 0, s>>>(be->d_blur, be->d_bands_blur, be->A, be->d_mean, be->cfg.contrast_measure, be->d_grad);
      }));
    }
  }
  return CMAXB_OK;
}
// copy contrast (+ gradient) to the host (queued) / wait and hand over
static int be_enqueue_fetch(cmaxb_be* be, bool want_grad) {
  cudaStream_t s = be->stream;
  const int P = 3 * be->n_opt;
  if (be->direct_out) return CMAXB_OK;      // the kernels wrote h_result / h_grad (mapped) themselves
  if (want_grad && P > 0) CMAXB_CUDA_TRY(cudaMemcpyAsync(be->h_grad, be->d_grad, sizeof(double) * P, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaMemcpyAsync(be->h_result, be->d_result, sizeof(double) * 4, cudaMemcpyDeviceToHost, s));
  return CMAXB_OK;
}
static int be_wait_fetch(cmaxb_be* be, bool want_grad, double* contrast, double* grad) {
  const int P = 3 * be->n_opt;
  CMAXB_CUDA_TRY(cudaStreamSynchronize(be->stream));
  *contrast = be->h_result[0];
  if (want_grad && grad) for (int i = 0; i < P; ++i) grad[i] = be->h_grad[i];
  return CMAXB_OK;
}
static int be_finish_fetch(cmaxb_be* be, bool want_grad, double* contrast, double* grad) {
  CMAXB_TRY(be_enqueue_fetch(be, want_grad));
  return be_wait_fetch(be, want_grad, contrast, grad);
}
static int be_finish_eval(cmaxb_be* be, bool want_grad, double* contrast, double* grad) {
  CMAXB_TRY(be_finish_launch(be, want_grad));
  return be_finish_fetch(be, want_grad, contrast, grad);
}

extern "C" int cmaxb_be_eval(cmaxb_be* be, const double* x, int n, double* contrast, double* grad) {
  if (!be || !contrast) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window: call cmaxb_be_set_window first");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  const bool want_grad = grad != nullptr;
  const int P = 3 * be->n_opt;
  const bool adjoint_grad = want_grad && P > 0 && be->cfg.grad_mode == CMAXB_GRAD_ADJOINT;
  be->split_pending = false;
  const int kind = want_grad ? 1 : 0;
  // results straight into mapped host memory (ADJOINT gradient or value only; the dense-band reduction keeps its copies)
  be->direct_out = !(want_grad && P > 0 && be->cfg.grad_mode == CMAXB_GRAD_DENSE);
  struct Reset { cmaxb_be* b; ~Reset() { b->direct_out = false; } } reset_direct{be};
  CMAXB_TRY(be_fill_x(be, x, n));
  if (x && n > 0) be->last_x.assign(x, x + n); else be->last_x.assign((size_t)(n > 0 ? n : 0), 0.0);
  const bool replay = be->graph_on && !be->prof.enabled && !be->alpha_pending && be->evals_in_window[kind] >= 1 &&
                      !(want_grad && be->cfg.grad_mode == CMAXB_GRAD_DENSE);
  be->evals_in_window[kind] += 1;
  if (replay) {
    cudaStream_t s = be->stream;
    if (!be->gexec[kind]) {
      cudaGraph_t graph = nullptr;
      CMAXB_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      int rc = be_enqueue_poses(be, want_grad, /*zero_acc=*/true);
      if (rc == CMAXB_OK) rc = be_run_scatter(be, true, adjoint_grad);
      if (rc == CMAXB_OK) rc = be_finish_launch(be, want_grad);
      if (rc == CMAXB_OK) rc = be_enqueue_fetch(be, want_grad);
      cudaError_t ce = cudaStreamEndCapture(s, &graph);
      if (rc != CMAXB_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        be->graph_on = false;                 // fall back to launch-by-launch evaluation for good
        if (rc != CMAXB_OK) return rc;
        return set_error(CMAXB_ERR_CUDA, std::string("graph capture of the back-end evaluation failed: ") + cudaGetErrorString(ce));
      }
      ce = cudaGraphInstantiate(&be->gexec[kind], graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { be->gexec[kind] = nullptr; be->graph_on = false; return set_error(CMAXB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
    }
    CMAXB_CUDA_TRY(cudaGraphLaunch(be->gexec[kind], s));
    g_launch_count.fetch_add(want_grad ? 6 : 3, std::memory_order_relaxed);    // kernels replayed by the graph
    return be_wait_fetch(be, want_grad, contrast, grad);
  }
  CMAXB_TRY(be_enqueue_poses(be, want_grad, /*zero_acc=*/true));
  CMAXB_TRY(be_run_scatter(be, true, adjoint_grad));
  return be_finish_eval(be, want_grad, contrast, grad);
}

extern "C" int cmaxb_be_last_eval_x(cmaxb_be* be, double* x, int n) {
  if (!be || (!x && n > 0)) return set_error(CMAXB_ERR_INVALID, "null argument");
  if ((int)be->last_x.size() != n) return set_error(CMAXB_ERR_STATE, "no evaluation with that many parameters since set_window");
  for (int i = 0; i < n; ++i) x[i] = be->last_x[(size_t)i];
  return CMAXB_OK;
}

// ---- event-sharded evaluation: one window split by TIME across GPUs ----------------
// begin : poses + scatter of THIS rank's events, IL assembled into one float plane
// (caller): all-reduce (SUM) of that plane across ranks, on the handle's stream
// end   : blur + variance on the summed plane (identical on every rank), adjoint image, gather over this
//         rank's events -> contrast and this rank's PARTIAL gradient (caller sums the partials)
extern "C" int cmaxb_be_eval_begin(cmaxb_be* be, const double* x, int n, int want_grad) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window: call cmaxb_be_set_window first");
  if (want_grad && be->cfg.grad_mode != CMAXB_GRAD_ADJOINT) return set_error(CMAXB_ERR_INVALID, "event-sharded evaluation needs CMAXB_GRAD_ADJOINT");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  const int P = 3 * be->n_opt;
  const bool g = want_grad && P > 0;
  if (!be->d_il_plane) CMAXB_TRY(dev_alloc(&be->d_il_plane, (size_t)be->A));
  CMAXB_TRY(be_run_poses(be, x, n, want_grad != 0));
  CMAXB_TRY(be_run_scatter(be, true, g, /*defer_alpha=*/true));
  cudaStream_t s = be->stream;
  CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
    be_assemble_il_kernel<<<148 * 8, 256, 0, s>>>(be->d_il_old, be->d_il_new, be->il_is_quad ? be->d_ilq : nullptr,
                                                 be->cfg.pano_width, be->A, be->d_il_plane);
  }));
  be->il_is_plane = true;
  be->split_pending = true; be->split_grad = want_grad != 0;
  return CMAXB_OK;
}

extern "C" int cmaxb_be_il_plane(cmaxb_be* be, float** device_ptr, size_t* count) {
  if (!be || !device_ptr || !count) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->d_il_plane) {
    CMAXB_CUDA_TRY(cudaSetDevice(be->device));
    CMAXB_TRY(dev_alloc(&be->d_il_plane, (size_t)be->A));
  }
  *device_ptr = be->d_il_plane;
  *count = (size_t)be->A;
  return CMAXB_OK;
}

extern "C" int cmaxb_be_eval_end(cmaxb_be* be, double* contrast, double* grad_partial) {
  if (!be || !contrast) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->split_pending) return set_error(CMAXB_ERR_STATE, "cmaxb_be_eval_end without cmaxb_be_eval_begin");
  if (grad_partial && !be->split_grad) return set_error(CMAXB_ERR_STATE, "gradient requested but eval_begin ran without want_grad");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  be->split_pending = false;
  be->il_is_plane = true;
  if (be->alpha_pending) CMAXB_TRY(be_run_alpha(be, be->d_il_plane, nullptr, nullptr));
  return be_finish_eval(be, grad_partial != nullptr, contrast, grad_partial);
}

// eval_end in two halves, so that the partial gradients can be summed across ranks ON THE DEVICE (NCCL all-reduce of
// cmaxb_be_grad_device() on the handle's stream) between them -- no host hop inside a sharded evaluation
extern "C" int cmaxb_be_eval_end_launch(cmaxb_be* be, int want_grad) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->split_pending) return set_error(CMAXB_ERR_STATE, "cmaxb_be_eval_end_launch without cmaxb_be_eval_begin");
  if (want_grad && !be->split_grad) return set_error(CMAXB_ERR_STATE, "gradient requested but eval_begin ran without want_grad");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  be->split_pending = false;
  be->il_is_plane = true;
  if (be->alpha_pending) CMAXB_TRY(be_run_alpha(be, be->d_il_plane, nullptr, nullptr));
  be->end_launched = true; be->end_grad = want_grad != 0;
  return be_finish_launch(be, want_grad != 0);
}
extern "C" int cmaxb_be_grad_device(cmaxb_be* be, double** device_ptr, size_t* count) {
  if (!be || !device_ptr || !count) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window");
  *device_ptr = be->d_grad;
  *count = (size_t)(3 * be->n_opt);
  return CMAXB_OK;
}
extern "C" int cmaxb_be_eval_end_fetch(cmaxb_be* be, double* contrast, double* grad) {
  if (!be || !contrast) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->end_launched) return set_error(CMAXB_ERR_STATE, "cmaxb_be_eval_end_fetch without cmaxb_be_eval_end_launch");
  if (grad && !be->end_grad) return set_error(CMAXB_ERR_STATE, "gradient requested but eval_end_launch ran without want_grad");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  be->end_launched = false;
  return be_finish_fetch(be, grad != nullptr, contrast, grad);
}

// row band of `rank` among `world`: own rows [y0, y1), extended by the halo hl = 2 r + 1 the blur and its adjoint need
static void be_band_geometry(cmaxb_be* be, int world, int rank) {
  const int H = be->cfg.pano_height;
  const int hl = 2 * be->taps.r + 1;
  const int hb = (H + world - 1) / world;
  be->sh_world = world; be->sh_rank = rank; be->sh_hb = hb; be->sh_hl = hl; be->sh_ce = hb + 2 * hl;
  be->sh_y0 = rank * hb; be->sh_y1 = std::min(H, (rank + 1) * hb);
  be->sh_first = std::max(0, be->sh_y0 - hl);                  // first REAL panorama row of the extended band
  be->sh_rows = std::min(H, be->sh_y1 + hl) - be->sh_first;     // real rows in it
}

static bool be_bands_ok(const cmaxb_be* be, int world) {
  const int H = be->cfg.pano_height;
  const int hb = (H + world - 1) / world;
  return world >= 1 && hb >= 2 * be->taps.r + 1 && (long long)(world - 1) * hb < H;
}

// ---- row-band sharding of the image phases (time-sharded window over several GPUs) ----------------------------------
// begin:   poses + scatter of THIS rank's events; IL packed as `world` extended bands (send buffer)
// (caller) reduce_scatter(recv <- send, SUM): every rank now holds the summed rows of its band + halo
// image:   blur of the band (+ alpha IGp), S1 / S2 over the band's OWN rows -> sums (2 doubles)
// (caller) all_reduce(sums, SUM)
// adjoint: contrast + mean from the sums; adjoint blur of the band -> own rows of G
// (caller) all_gather(g_full <- g_own)
// gather:  adjoint gather over this rank's events with the full G plane -> partial gradient (cmaxb_be_grad_device)
// (caller) all_reduce(gradient, SUM); cmaxb_be_eval_end_fetch
// Compared with the whole-plane exchange (cmaxb_be_eval_begin / _end) the same bytes cross NVLink (reduce-scatter + all-gather
// = one all-reduce) but blur and adjoint blur run on 1/world of the panorama instead of being replicated on every rank.
extern "C" int cmaxb_be_shard_begin(cmaxb_be* be, const double* x, int n, int want_grad, int world, int rank, float** send_dev,
                                    float** recv_dev, size_t* chunk_floats) {
  if (!be || !send_dev || !recv_dev || !chunk_floats) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window: call cmaxb_be_set_window first");
  if (be->alpha_pending) return set_error(CMAXB_ERR_STATE, "alpha is not fixed yet: run the window's first evaluation through cmaxb_be_eval_begin / _end");
  if (want_grad && be->cfg.grad_mode != CMAXB_GRAD_ADJOINT) return set_error(CMAXB_ERR_INVALID, "event-sharded evaluation needs CMAXB_GRAD_ADJOINT");
  const int W = be->cfg.pano_width, H = be->cfg.pano_height;
  const int hl = 2 * be->taps.r + 1;
  const int hb = (H + world - 1) / world;
  if (world < 1 || rank < 0 || rank >= world || hb < hl || (long long)(world - 1) * hb >= H)
    return set_error(CMAXB_ERR_INVALID, "bad world / rank, or bands thinner than the halo (use the whole-plane exchange)");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  const int ce = hb + 2 * hl;
  const size_t need = (size_t)world * ce * W;
  if (need > be->sh_cap || be->sh_world != world) {
    cudaFree(be->d_sh_send); cudaFree(be->d_sh_recv); cudaFree(be->d_sh_blur); cudaFree(be->d_sh_gband); cudaFree(be->d_sh_gfull); cudaFree(be->d_sh_sums);
    be->d_sh_send = be->d_sh_recv = be->d_sh_blur = be->d_sh_gband = be->d_sh_gfull = nullptr; be->d_sh_sums = nullptr; be->sh_cap = 0;
    CMAXB_TRY(dev_alloc(&be->d_sh_send, need));
    CMAXB_TRY(dev_alloc(&be->d_sh_recv, (size_t)ce * W));
    CMAXB_TRY(dev_alloc(&be->d_sh_blur, (size_t)ce * W));
    CMAXB_TRY(dev_alloc(&be->d_sh_gband, (size_t)ce * W));
    CMAXB_TRY(dev_alloc(&be->d_sh_gfull, (size_t)world * hb * W + W + 1));
    CMAXB_TRY(dev_alloc(&be->d_sh_sums, 2));
    CMAXB_CUDA_TRY(cudaMemset(be->d_sh_gfull, 0, sizeof(float) * ((size_t)world * hb * W + W + 1)));
    be->sh_cap = need;
  }
  be_band_geometry(be, world, rank);
  const int P = 3 * be->n_opt;
  const bool g = want_grad && P > 0;
  CMAXB_TRY(be_run_poses(be, x, n, want_grad != 0));
  CMAXB_TRY(be_run_scatter(be, true, g, /*defer_alpha=*/true));
  cudaStream_t s = be->stream;
  CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
    be_pack_bands_kernel<<<148 * 8, 256, 0, s>>>(be->d_il_old, be->d_il_new, be->il_is_quad ? be->d_ilq : nullptr, W, H, world, hb, hl, be->d_sh_send);
  }));
  if (x && n > 0) be->last_x.assign(x, x + n); else be->last_x.assign((size_t)(n > 0 ? n : 0), 0.0);
  be->sh_stage = 1; be->sh_grad = g;
  *send_dev = be->d_sh_send; *recv_dev = be->d_sh_recv; *chunk_floats = (size_t)ce * W;
  return CMAXB_OK;
}

extern "C" int cmaxb_be_shard_image(cmaxb_be* be, double** sums_dev) {
  if (!be || !sums_dev) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (be->sh_stage != 1) return set_error(CMAXB_ERR_STATE, "cmaxb_be_shard_image without cmaxb_be_shard_begin");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  cudaStream_t s = be->stream;
  const int W = be->cfg.pano_width;
  // the band as a stand-alone image of its REAL rows: a true panorama edge gets BORDER_REFLECT_101 as it should, an
  // artificial one spoils r rows of the blur and r more of the adjoint -- inside the halo, never used
  const int skip = be->sh_first - (be->sh_y0 - be->sh_hl);      // zero rows above the panorama (first rank only)
  const float* il = be->d_sh_recv + (size_t)skip * W;
  const float* igp = be->have_igp ? be->d_igp + (size_t)be->sh_first * W : nullptr;
  ReduceOut ro{be->d_acc, be->d_ticket, be->d_result, be->d_mean};
  ro.raw = be->d_sh_sums;
  ro.sum_y0 = be->sh_y0 - be->sh_first; ro.sum_y1 = be->sh_y1 - be->sh_first;
  const SrcBePlane src{il, igp, (float)be->alpha};
  cudaError_t le = cudaSuccess;
  CMAXB_TRY(be->prof.run(CMAXB_K_BLUR_REDUCE, s, true, [&] {
    le = launch_blur_reduce<1, SrcBePlane, true>(s, 1, src, W, be->sh_rows, be->taps, be->d_sh_blur, 0, ro, be->cfg.contrast_measure);
  }));
  if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("blur_reduce launch: ") + cudaGetErrorString(le));
  be->sh_stage = 2;
  *sums_dev = be->d_sh_sums;
  return CMAXB_OK;
}

extern "C" int cmaxb_be_shard_adjoint(cmaxb_be* be, float** g_own_dev, float** g_full_dev, size_t* own_floats) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (be->sh_stage != 2) return set_error(CMAXB_ERR_STATE, "cmaxb_be_shard_adjoint without cmaxb_be_shard_image");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  cudaStream_t s = be->stream;
  const int W = be->cfg.pano_width, H = be->cfg.pano_height;
  CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
    be_band_finalize_kernel<<<1, 32, 0, s>>>(be->d_sh_sums, (double)W * (double)H, be->cfg.contrast_measure, be->d_result, be->d_mean);
  }));
  if (g_own_dev) *g_own_dev = nullptr;
  if (g_full_dev) *g_full_dev = nullptr;
  if (own_floats) *own_floats = (size_t)be->sh_hb * W;
  if (be->sh_grad) {
    cudaError_t le = cudaSuccess;
    CMAXB_TRY(be->prof.run(CMAXB_K_ADJOINT_BLUR, s, true, [&] {
      le = launch_adjoint_blur<false>(s, 1, be->d_sh_blur, 0, W, be->sh_rows, be->taps, be->d_mean, be->cfg.contrast_measure, be->d_sh_gband, nullptr);
    }));
    if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("adjoint_blur launch: ") + cudaGetErrorString(le));
    // the band's own rows of G (hb rows; a short last band is zero padded by the allocation)
    if (g_own_dev) *g_own_dev = be->d_sh_gband + (size_t)(be->sh_y0 - be->sh_first) * W;
    if (g_full_dev) *g_full_dev = be->d_sh_gfull;
    be->sh_stage = 3;
  } else {
    be->sh_stage = 0;
    be->end_launched = true; be->end_grad = false;
  }
  return CMAXB_OK;
}

extern "C" int cmaxb_be_shard_gather(cmaxb_be* be) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (be->sh_stage != 3) return set_error(CMAXB_ERR_STATE, "cmaxb_be_shard_gather without cmaxb_be_shard_adjoint (gradient evaluation)");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_gather_launch(be, be->d_sh_gfull, nullptr));
  be->sh_stage = 0;
  be->end_launched = true; be->end_grad = true;
  return CMAXB_OK;
}

// ---- the same evaluation with the exchange done by the kernels over peer memory (be_xchg.cuh) ---------------------------
extern "C" int cmaxb_be_exchange_init(cmaxb_be* be, int world, int rank, void* handle64_out) {
  if (!be || !handle64_out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (world < 1 || world > kBeXMaxWorld || rank < 0 || rank >= world) return set_error(CMAXB_ERR_INVALID, "bad world / rank (at most 8 ranks)");
  if (!be_bands_ok(be, world)) return set_error(CMAXB_ERR_INVALID, "row bands thinner than the blur halo: use the whole-plane exchange");
  if (be->x_local) return set_error(CMAXB_ERR_STATE, "exchange already initialised");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(be->stream));
  const int W = be->cfg.pano_width, H = be->cfg.pano_height;
  be_band_geometry(be, world, rank);
  be->x_lay = be_x_layout(W, be->sh_ce, world);
  size_t bytes = (size_t)2 << 20;            // whole 2 MiB blocks: the IPC handle exports nothing else
  while (bytes < be->x_lay.total) bytes += (size_t)2 << 20;
  CMAXB_CUDA_TRY(cudaMalloc((void**)&be->x_local, bytes));
  CMAXB_CUDA_TRY(cudaMemset(be->x_local, 0, bytes));
  const int ntiles = ((W + kBeXTile - 1) / kBeXTile) * ((H + kBeXTile - 1) / kBeXTile);
  if (ntiles > 32 * kBeDirtyWords) return set_error(CMAXB_ERR_INVALID, "panorama too large for the dirty-tile bitmap (32768 tiles of 32 x 32)");
  CMAXB_TRY(dev_alloc(&be->d_dirty, (size_t)kBeDirtyWords));
  CMAXB_CUDA_TRY(cudaMemset(be->d_dirty, 0, sizeof(unsigned int) * kBeDirtyWords));
  CMAXB_TRY(dev_alloc(&be->d_xticket, 1));
  CMAXB_CUDA_TRY(cudaMemset(be->d_xticket, 0, sizeof(unsigned int)));
  CMAXB_CUDA_TRY(cudaHostAlloc((void**)&be->h_xfault, sizeof(unsigned long long), cudaHostAllocMapped));
  CMAXB_CUDA_TRY(cudaHostGetDevicePointer((void**)&be->d_xfault, be->h_xfault, 0));
  be->h_xfault[0] = 0;
  if (!be->d_sh_blur) {
    CMAXB_TRY(dev_alloc(&be->d_sh_blur, (size_t)be->sh_ce * W));
    CMAXB_TRY(dev_alloc(&be->d_sh_sums, 2));
  }
  CMAXB_CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CMAXB_CUDA_TRY(cudaIpcGetMemHandle(&h, be->x_local));
  std::memcpy(handle64_out, &h, 64);
  be->x_peers.world = world; be->x_peers.rank = rank;
  be->x_invariant = false; be->x_seq = 0;
  return CMAXB_OK;
}

extern "C" int cmaxb_be_exchange_connect(cmaxb_be* be, const void* handles) {
  if (!be || !handles) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->x_local) return set_error(CMAXB_ERR_STATE, "call cmaxb_be_exchange_init first");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  for (int r = 0; r < be->x_peers.world; ++r) {
    if (r == be->x_peers.rank) { be->x_peers.base[r] = be->x_local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + 64 * r, 64);
    void* ptr = nullptr;
    CMAXB_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    be->x_peers.base[r] = (char*)ptr;
  }
  be->x_on = true;
  return CMAXB_OK;
}

extern "C" int cmaxb_be_exchange_close(cmaxb_be* be) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  cudaSetDevice(be->device);
  cudaStreamSynchronize(be->stream);
  be->x_on = false;
  for (int r = 0; r < be->x_peers.world; ++r) {
    if (r != be->x_peers.rank && be->x_peers.base[r]) cudaIpcCloseMemHandle(be->x_peers.base[r]);
    be->x_peers.base[r] = nullptr;
  }
  // the local block stays allocated until destroy: a peer may still have it mapped
  return CMAXB_OK;
}

// tiles[0] = panorama tiles this rank's events touched in the last cmaxb_be_xeval, tiles[1] = tiles of the panorama
extern "C" int cmaxb_be_exchange_stats(cmaxb_be* be, int64_t* tiles2) {
  if (!be || !tiles2) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->d_dirty) return set_error(CMAXB_ERR_STATE, "no peer exchange");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  std::vector<unsigned int> bits(kBeDirtyWords);
  CMAXB_CUDA_TRY(cudaMemcpyAsync(bits.data(), be->d_dirty, sizeof(unsigned int) * kBeDirtyWords, cudaMemcpyDeviceToHost, be->stream));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(be->stream));
  int64_t n = 0;
  for (unsigned int w : bits) n += __builtin_popcount(w);
  tiles2[0] = n;
  tiles2[1] = (int64_t)((be->cfg.pano_width + kBeXTile - 1) / kBeXTile) * ((be->cfg.pano_height + kBeXTile - 1) / kBeXTile);
  return CMAXB_OK;
}

// Collective: every rank of the exchange calls it with the same x, in the same order.  contrast and gradient of the WHOLE
// window come back on every rank (bitwise identical: all sums across ranks are formed in rank order).
extern "C" int cmaxb_be_xeval(cmaxb_be* be, const double* x, int n, double* contrast, double* grad) {
  if (!be || !contrast) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->x_on) return set_error(CMAXB_ERR_STATE, "no peer exchange: call cmaxb_be_exchange_init / _connect first");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window: call cmaxb_be_set_window first");
  if (be->alpha_pending) return set_error(CMAXB_ERR_STATE, "alpha is not fixed yet: run the window's first evaluation through cmaxb_be_eval_begin / _end");
  if (grad && be->cfg.grad_mode != CMAXB_GRAD_ADJOINT) return set_error(CMAXB_ERR_INVALID, "event-sharded evaluation needs CMAXB_GRAD_ADJOINT");
  const int P = 3 * be->n_opt;
  if (P > kBeXGradMax) return set_error(CMAXB_ERR_INVALID, "too many free control poses for the gradient exchange (1024)");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  cudaStream_t s = be->stream;
  const int W = be->cfg.pano_width, H = be->cfg.pano_height;
  const bool g = grad != nullptr && P > 0;
  if (!be->d_ilq) {
    CMAXB_TRY(dev_alloc(&be->d_ilq, (size_t)be->A));
    CMAXB_TRY(dev_alloc(&be->d_GQ, (size_t)be->A));
  }
  BeXGeom xg{W, H, (W + kBeXTile - 1) / kBeXTile, (H + kBeXTile - 1) / kBeXTile, be->sh_hb, be->sh_hl, be->sh_ce};
  const int ntiles = xg.ntx * xg.nty;
  const BeXLayout& L = be->x_lay;
  const BeXPeers& peers = be->x_peers;
  const unsigned long long seq = ++be->x_seq;
  const int par = (int)(seq & 1);
  const unsigned tile_grid = (unsigned)std::min(ntiles, 148 * 8);
  // 1. accumulator: only the tiles the previous evaluation touched are non-zero
  CMAXB_TRY(be->prof.run(CMAXB_K_ZERO, s, true, [&] {
    if (!be->x_invariant) {
      cudaMemsetAsync(be->d_ilq, 0, sizeof(float4) * be->A, s);
      cudaMemsetAsync(be->d_dirty, 0, sizeof(unsigned int) * kBeDirtyWords, s);
    } else {
      be_x_clean_kernel<<<tile_grid, 256, 0, s>>>(be->d_ilq, be->d_dirty, xg);
    }
  }));
  // 2. poses + scatter of the own slab (tiles flagged)
  CMAXB_TRY(be_run_poses(be, x, n, g));
  CMAXB_TRY(be_run_scatter(be, true, g, /*defer_alpha=*/true, be->d_dirty));
  be->x_invariant = true;
  // 3. dirty tiles -> the band owners' accumulators; flag at every peer
  CMAXB_TRY(be->prof.run(CMAXB_K_BE_X_PUSH, s, true, [&] {
    be_x_push_kernel<<<tile_grid, 256, 0, s>>>(be->d_ilq, be->d_dirty, xg, peers, L.acc[par], L.push_flag, seq, be->d_xticket);
    be_x_wait_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const unsigned long long*>(be->x_local + L.push_flag), peers.world, seq, be->d_xfault, 0x20);
  }));
  // 4. blur of the own band (+ alpha IGp), S1 / S2 over its own rows
  {
    const int skip = be->sh_first - (be->sh_y0 - be->sh_hl);
    const float* il = reinterpret_cast<const float*>(be->x_local + L.acc[par]) + (size_t)skip * W;
    const float* igp = be->have_igp ? be->d_igp + (size_t)be->sh_first * W : nullptr;
    ReduceOut ro{be->d_acc, be->d_ticket, be->d_result, be->d_mean};
    ro.raw = be->d_sh_sums;
    ro.sum_y0 = be->sh_y0 - be->sh_first; ro.sum_y1 = be->sh_y1 - be->sh_first;
    const SrcBePlane src{il, igp, (float)be->alpha};
    cudaError_t le = cudaSuccess;
    CMAXB_TRY(be->prof.run(CMAXB_K_BLUR_REDUCE, s, true, [&] {
      le = launch_blur_reduce<1, SrcBePlane, true>(s, 1, src, W, be->sh_rows, be->taps, be->d_sh_blur, 0, ro, be->cfg.contrast_measure);
    }));
    if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("blur_reduce launch: ") + cudaGetErrorString(le));
  }
  CMAXB_TRY(be->prof.run(CMAXB_K_BE_X_SUMS, s, true, [&] {
    cudaMemsetAsync(be->x_local + L.acc[par], 0, sizeof(float) * (size_t)be->sh_ce * W, s);       // ready for evaluation seq + 2
    be_x_sums_kernel<<<1, 32, 0, s>>>(be->d_sh_sums, peers, L.sums, seq, (double)W * (double)H, be->cfg.contrast_measure, be->d_result,
                                      be->d_mean, be->d_xfault);
  }));
  if (g) {
    // 5. adjoint image of the band; every rank pulls G for its dirty tiles; gather; gradient exchange
    cudaError_t le = cudaSuccess;
    CMAXB_TRY(be->prof.run(CMAXB_K_ADJOINT_BLUR, s, true, [&] {
      le = launch_adjoint_blur<false>(s, 1, be->d_sh_blur, 0, W, be->sh_rows, be->taps, be->d_mean, be->cfg.contrast_measure,
                                      reinterpret_cast<float*>(be->x_local + L.G), nullptr);
    }));
    if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("adjoint_blur launch: ") + cudaGetErrorString(le));
    CMAXB_TRY(be->prof.run(CMAXB_K_BE_X_PULL, s, true, [&] {
      be_x_flag_kernel<<<1, 32, 0, s>>>(peers, L.g_flag, seq);
      be_x_pull_kernel<<<tile_grid, 256, 0, s>>>(be->d_GQ, be->d_dirty, xg, peers, L.G,
                                                 reinterpret_cast<const unsigned long long*>(be->x_local + L.g_flag), seq, be->d_xfault);
    }));
    CMAXB_TRY(be_gather_launch(be, nullptr, be->d_GQ));
    CMAXB_TRY(be->prof.run(CMAXB_K_BE_X_GRAD, s, true, [&] {
      be_x_grad_kernel<<<(P + 255) / 256, 256, 0, s>>>(be->d_grad, P, peers, L.grad, seq, be->d_xfault);
    }));
  }
  if (x && n > 0) be->last_x.assign(x, x + n); else be->last_x.assign((size_t)(n > 0 ? n : 0), 0.0);
  be->end_launched = true; be->end_grad = g;
  CMAXB_TRY(cmaxb_be_eval_end_fetch(be, contrast, g ? grad : nullptr));
  if (be->h_xfault[0]) {
    const unsigned long long why = be->h_xfault[0];
    be->h_xfault[0] = 0;
    return set_error(CMAXB_ERR_CUDA, "peer exchange: a rank did not arrive (time-out), stage 0x" + std::to_string(why & 0xff) + " rank " + std::to_string((why >> 8) & 0xff));
  }
  if (grad && !g) for (int i = 0; i < P; ++i) grad[i] = 0.0;
  return CMAXB_OK;
}

extern "C" int cmaxb_be_get_alpha(cmaxb_be* be, double* alpha) {
  if (!be || !alpha) return set_error(CMAXB_ERR_INVALID, "null argument");
  *alpha = be->alpha_pending ? NAN : be->alpha;
  return CMAXB_OK;
}

extern "C" int cmaxb_be_get_il(cmaxb_be* be, const double* x, int n, float* il_old, float* il_new) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_run_poses(be, x, n, false));
  CMAXB_TRY(be_run_scatter(be, false));
  cudaStream_t s = be->stream;
  if (il_old) CMAXB_CUDA_TRY(cudaMemcpyAsync(il_old, be->d_il_old, sizeof(float) * be->A, cudaMemcpyDeviceToHost, s));
  if (il_new) CMAXB_CUDA_TRY(cudaMemcpyAsync(il_new, be->d_il_new, sizeof(float) * be->A, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_get_iwe(cmaxb_be* be, const double* x, int n, int blurred, float* out) {
  if (!be || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_run_poses(be, x, n, false));
  CMAXB_TRY(be_run_scatter(be, true));
  cudaStream_t s = be->stream;
  Taps t0{}; t0.r = 0; t0.w[0] = 1.0f;   // un-blurred I = IL + alpha*IGp: the same kernel with a radius-0 filter
  CMAXB_TRY(be_run_image(be, blurred ? be->taps : t0));
  CMAXB_CUDA_TRY(cudaMemcpyAsync(out, be->d_blur, sizeof(float) * be->A, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_get_bands(cmaxb_be* be, const double* x, int n, int blurred, float* out) {
  if (!be || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_run_poses(be, x, n, true));
  const bool do_blur = blurred && be->taps.r > 0;
  CMAXB_TRY(be_run_bands(be, do_blur));
  cudaStream_t s = be->stream;
  const size_t bytes = sizeof(float) * (size_t)3 * be->n_opt * be->A;
  CMAXB_CUDA_TRY(cudaMemcpyAsync(out, do_blur ? be->d_bands_blur : be->d_bands, bytes, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_get_cells(cmaxb_be* be, const double* x, int n, int32_t* out) {
  if (!be || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  if (be->n == 0) return CMAXB_OK;
  CMAXB_TRY(be_run_poses(be, x, n, false));
  if ((size_t)be->n > be->cells_cap) {
    cudaFree(be->d_cells); be->d_cells = nullptr; be->cells_cap = 0;
    CMAXB_TRY(dev_alloc(&be->d_cells, (size_t)be->n));
    be->cells_cap = (size_t)be->n;
  }
  cudaStream_t s = be->stream;
  const BeGeom g = be_geom(be);
  CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
    be_cells_kernel<<<(unsigned)((be->n + 255) / 256), 256, 0, s>>>(g, be->d_poses, be->n, be->d_cells);
  }));
  CMAXB_CUDA_TRY(cudaMemcpyAsync(out, be->d_cells, sizeof(int) * be->n, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_get_poses(cmaxb_be* be, const double* x, int n, int64_t* n_batches, double* R9, float* Jk,
                                  int32_t* idx_cp_beg, int64_t capacity) {
  if (!be || !n_batches) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  *n_batches = be->nb;
  if (capacity < be->nb || be->nb == 0) return CMAXB_OK;
  CMAXB_TRY(be_run_poses(be, x, n, true));
  std::vector<BePose> h((size_t)be->nb);
  CMAXB_CUDA_TRY(cudaMemcpyAsync(h.data(), be->d_poses, sizeof(BePose) * be->nb, cudaMemcpyDeviceToHost, be->stream));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(be->stream));
  const int nj = 9 * be->N;
  for (long long b = 0; b < be->nb; ++b) {
    if (R9) for (int i = 0; i < 9; ++i) R9[9 * b + i] = h[b].R[i];
    if (Jk) for (int i = 0; i < nj; ++i) Jk[nj * b + i] = h[b].Jk[i];
    if (idx_cp_beg) idx_cp_beg[b] = h[b].idx_cp_beg;
  }
  return CMAXB_OK;
}

// ---- device-resident global map: IG_ and IG_update_times_map_ stay in HBM between windows -------------------
static int be_map_ensure(cmaxb_be* be) {
  if (be->d_IG) return CMAXB_OK;
  CMAXB_TRY(dev_alloc(&be->d_IG, (size_t)be->A));
  CMAXB_TRY(dev_alloc(&be->d_times, (size_t)be->A));
  CMAXB_TRY(dev_alloc(&be->d_mask, (size_t)be->A));
  CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_IG, 0, sizeof(float) * be->A, be->stream));
  CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_times, 0, (size_t)be->A, be->stream));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_map_reset(cmaxb_be* be) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_map_ensure(be));
  CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_IG, 0, sizeof(float) * be->A, be->stream));
  CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_times, 0, (size_t)be->A, be->stream));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(be->stream));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_map_set(cmaxb_be* be, const float* IG, const uint8_t* times) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_map_ensure(be));
  if (IG) CMAXB_CUDA_TRY(cudaMemcpyAsync(be->d_IG, IG, sizeof(float) * be->A, cudaMemcpyHostToDevice, be->stream));
  if (times) CMAXB_CUDA_TRY(cudaMemcpyAsync(be->d_times, times, (size_t)be->A, cudaMemcpyHostToDevice, be->stream));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(be->stream));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_map_get(cmaxb_be* be, float* IG, uint8_t* times) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_map_ensure(be));
  if (IG) CMAXB_CUDA_TRY(cudaMemcpyAsync(IG, be->d_IG, sizeof(float) * be->A, cudaMemcpyDeviceToHost, be->stream));
  if (times) CMAXB_CUDA_TRY(cudaMemcpyAsync(times, be->d_times, (size_t)be->A, cudaMemcpyDeviceToHost, be->stream));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(be->stream));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_map_use_as_igp(cmaxb_be* be, double alpha) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window: call cmaxb_be_set_window first");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_map_ensure(be));
  // updateIGp: IGp <- IG (event_pano_warper.cpp:128-132), device to device
  CMAXB_CUDA_TRY(cudaMemcpyAsync(be->d_igp, be->d_IG, sizeof(float) * be->A, cudaMemcpyDeviceToDevice, be->stream));
  be->have_igp = true;
  if (std::isnan(alpha)) { be->alpha = 0.0; be->alpha_pending = true; }
  else { be->alpha = alpha; be->alpha_pending = false; }
  return CMAXB_OK;
}

extern "C" int cmaxb_be_map_update(cmaxb_be* be, const double* x, int n, int max_update_times) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!be->have_window) return set_error(CMAXB_ERR_STATE, "no window: call cmaxb_be_set_window first");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_map_ensure(be));
  // IL_old_ at the optimum, then updateIG (event_pano_warper.cpp:109-126)
  CMAXB_TRY(be_run_poses(be, x, n, false));
  CMAXB_TRY(be_run_scatter(be, false, false, /*defer_alpha=*/true));
  cudaStream_t s = be->stream;
  CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
    be_update_ig_kernel<<<148 * 8, 256, 0, s>>>(be->d_IG, be->d_il_old, be->d_times, max_update_times, be->A);
  }));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_map_mark_fov(cmaxb_be* be, const double* rot_xyzw, int m, int radius) {
  if (!be || (!rot_xyzw && m > 0) || radius < 0) return set_error(CMAXB_ERR_INVALID, "bad argument");
  CMAXB_CUDA_TRY(cudaSetDevice(be->device));
  CMAXB_TRY(be_map_ensure(be));
  cudaStream_t s = be->stream;
  const BeGeom g = be_geom(be);
  for (int i = 0; i < m; ++i) {   // one setUpdateTimesIG(rot_check, radius) per pose (pose_graph_optimizer.cpp:325-337)
    Quat q; q.x = rot_xyzw[4 * i]; q.y = rot_xyzw[4 * i + 1]; q.z = rot_xyzw[4 * i + 2]; q.w = rot_xyzw[4 * i + 3];
    CMAXB_CUDA_TRY(cudaMemsetAsync(be->d_mask, 0, (size_t)be->A, s));
    CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
      be_fov_mask_kernel<<<(unsigned)((be->SA + 255) / 256), 256, 0, s>>>(g, q, radius, be->d_mask);
    }));
    CMAXB_TRY(be->prof.run(CMAXB_K_MISC, s, true, [&] {
      be_times_add_kernel<<<148 * 8, 256, 0, s>>>(be->d_times, be->d_mask, be->A);
    }));
  }
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

extern "C" int cmaxb_be_profile(cmaxb_be* be, int enable) {
  if (!be) return set_error(CMAXB_ERR_INVALID, "null argument");
  be->prof.enabled = enable != 0;
  be->prof.reset();
  return CMAXB_OK;
}
extern "C" int cmaxb_be_kernel_times(cmaxb_be* be, double* ms, uint64_t* launches) {
  if (!be || !ms || !launches) return set_error(CMAXB_ERR_INVALID, "null argument");
  for (int i = 0; i < CMAXB_K_COUNT; ++i) { ms[i] = be->prof.ms[i]; launches[i] = be->prof.launches[i]; }
  return CMAXB_OK;
}
