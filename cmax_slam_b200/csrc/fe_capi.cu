// fe_capi.cu -- C ABI of the front-end path (see include/cmax_b200.h).
//
// Handle layout: `packets` = resident event packets (slots; the current one is selected with
// cmaxb_fe_select_packet), `lanes` = independent evaluation pipelines (stream + accumulators + gather records +
// reduction state).  Lane 0 is the handle's main stream: packet uploads / preparation, the synchronous
// cmaxb_fe_eval / _eval_batch (the GSL-callback pattern, whole co-resident grid) and the getters run there.
// With cfg.lanes >= 2, cmaxb_fe_eval_launch round-robins over lanes 1..cfg.lanes (own streams) with a partial grid,
// so that the latency-bound phases of several evaluations overlap on the device and with the next packet's upload
// (profiles/r01f_two_stream_probe.txt).
#include "capi_common.cuh"
#include "fe_kernels.cuh"
#include "image_kernels.cuh"
#include "fe_fused.cuh"
#include "fe_binning.cuh"

#include <deque>

using namespace cmaxb;

constexpr int kFeRing = CMAXB_FE_MAX_OUTSTANDING;   // launches that may be outstanding (mapped result slots)
constexpr int kFeMaxLanes = 1 + 4;         // lane 0 (main stream) + up to 4 throughput lanes
constexpr int kFeMaxPackets = 16;
static_assert(kXSlots >= 2 * kFeRing, "exchange slots must cover twice the launch ring");

struct FeInflight {
  int k; bool grad; bool fused; int slot; int lane; unsigned long long tag; bool xchg; int pkt;
};

namespace cmaxb {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};
}  // namespace cmaxb

struct FePacket {
  uint4* d_ev = nullptr; size_t ev_cap = 0;      // owned copy of the raw events
  const uint4* ev = nullptr;                     // raw events in use: d_ev, or a caller-owned device buffer (view)
  double* d_dt = nullptr; size_t dt_cap = 0;
  uint4* d_bev = nullptr; size_t bev_cap = 0; bool have_bins = false;   // tile-binned records {x|y<<16, batch, dt}
  unsigned int* d_tile_end = nullptr;            // [ntiles] end offset of every source tile's run (cursor after the binning scatter)
  long long n = 0, nb = 0;
  bool have = false; bool flags_pending = false;
  int* d_flags = nullptr; int* h_flags = nullptr;
  cudaEvent_t ready = nullptr;                   // recorded on lane 0's stream after the preparation kernels
  unsigned long long gen = 0;                    // bumped by every set_packet
  int users = 0;                                 // evaluations launched on this packet and not yet fetched
};

struct FeLane {
  cudaStream_t stream = nullptr; bool own_stream = false; bool ready = false;
  // value accumulators: two corner-split ("quad") images used alternately; the image phase of one evaluation clears
  // the image the next evaluation scatters into (no memset in steady state)
  float4* d_quad[2] = {nullptr, nullptr}; int quad_cur = 0; int quad_dirty[2] = {0, 0};
  float4* d_GQ = nullptr;
  double* d_sums = nullptr;                      // [kmax][8] accumulators, kSumStride doubles apart
  unsigned int* d_ticket = nullptr;
  unsigned long long* d_bar = nullptr; unsigned long long bar_count = 0;
  unsigned long long* h_fault = nullptr; unsigned long long* d_fault = nullptr;   // mapped: barrier / tile-copy timeout
  unsigned long long* d_phase = nullptr;         // [16] device words: phase stamps of the last profiled launch
  CUtensorMap tmap[2][2];                        // [grid mode][quad buffer]
  unsigned long long seen_gen[kFeMaxPackets] = {};   // packet generation this lane's stream is ordered after
  int inflight = 0;
};

struct cmaxb_fe {
  cmaxb_fe_cfg cfg{};
  int device = 0;
  long long A = 0;          // pixels
  int kmax = 1;
  Taps taps{};
  float cxl[kMaxRadius + 1], cxr[kMaxRadius + 1], cyl[kMaxRadius + 1], cyr[kMaxRadius + 1];
  float acorr[4 * kMaxRadius + 1];
  float* d_adj_tab = nullptr;       // [xl | xr | yl | yr][2r][4r+1]
  double4* d_lut = nullptr;
  unsigned int* d_tile_count = nullptr; int ntiles = 0, ntx = 0;
  CUtensorMap tmap_lut;             // bearing-vector LUT as a 2-D tensor, box = one 32 x 32 source tile
  bool use_bins = true;
  FePacket packets[kFeMaxPackets]; int npackets = 1; int cur = 0;
  FeLane lanes[kFeMaxLanes]; int nlanes = 1; int lane_next = 0;   // nlanes = throughput lanes (lanes[1..nlanes]); <= 1: everything on lane 0
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  cudaStream_t stream = nullptr;    // == lanes[0].stream
  // stand-alone kernels (DENSE gradients, getters, A/B): lane 0 only
  float* d_blur1 = nullptr; float4* d_img4 = nullptr; float4* d_blur4 = nullptr;
  int* d_cells = nullptr; size_t cells_cap = 0;
  double* d_omegas = nullptr; double* h_omegas = nullptr;
  double* d_acc = nullptr; unsigned int* d_ticket = nullptr; unsigned int* d_ticket2 = nullptr;
  double* d_gacc = nullptr; double* d_result = nullptr; double* d_mean = nullptr; double* h_result = nullptr;
  // fused evaluation
  int grid[2] = {0, 0}; int th[2] = {16, 16};   // [0] whole co-resident grid (latency), [1] partial grid (throughput lanes)
  bool use_tma = false; bool tma_ok = false;
  unsigned long long* h_ring = nullptr; unsigned long long* d_ring = nullptr;   // mapped pinned memory, tagged words: [kFeRing][kmax][4][2]
  unsigned long long launch_no = 0; // tags the result words of a launch
  bool force_multi_kernel = false;  // CMAXB_FE_MULTI_KERNEL=1: stand-alone kernels (profiling / A-B comparison)
  bool pending = false;             // work of ours may still be running on some lane
  std::deque<FeInflight> inflight;  // launched, not yet fetched (FIFO)
  int ring_next = 0;
  double* d_mirror = nullptr;       // caller-owned device buffer [kmax][4] (cmaxb_fe_set_result_mirror)
  // fused result exchange over peer memory (cmaxb_fe_exchange_*)
  int x_world = 0, x_rank = 0; bool x_on = false;
  unsigned long long* x_local = nullptr; size_t x_bytes = 0;
  unsigned long long* x_peer[kXMaxWorld] = {};
  double* x_all_dev = nullptr;
  unsigned long long* h_xall = nullptr; unsigned long long* d_xall = nullptr;   // mapped, tagged words: [kFeRing][world][kmax][4][2]
  unsigned int* h_xerr = nullptr; unsigned int* d_xerr = nullptr;
  unsigned long long x_seq = 0;
  KernelProfiler prof;
};

static FeGeom fe_geom(const cmaxb_fe* fe, const FePacket& pk) {
  FeGeom g;
  g.ev = pk.ev; g.n = pk.n; g.batch_size = fe->cfg.batch_size; g.dt_tab = pk.d_dt; g.lut = fe->d_lut;
  g.W = fe->cfg.width; g.H = fe->cfg.height;
  g.fx = fe->cfg.fx; g.fy = fe->cfg.fy; g.cx = fe->cfg.cx; g.cy = fe->cfg.cy;
  return g;
}

// C = B^T 1 along one axis of length n (see fe_fused.cuh): the adjoint's own formula applied to an all-ones signal
static void border_table(const Taps& t, int n, float* lo, float* hi) {
  const int r = t.r;
  auto conv1 = [&](int j) {
    float s = 0.f;
    for (int d = -r; d <= r; ++d) if (j + d >= 0 && j + d < n) s = fmaf(t.w[r + d], 1.0f, s);
    return s;
  };
  auto full = [&](int q) {
    if (q < 0 || q >= n) return 1.0f;
    float s = conv1(q);
    if (q >= 1 && q <= r) for (int d = q; d <= r; ++d) if (-q + d >= 0 && -q + d < n) s = fmaf(t.w[r + d], 1.0f, s);
    if (q <= n - 2 && q >= n - 1 - r) for (int d = -r; d <= q - (n - 1); ++d) { const int j = 2 * (n - 1) - q + d; if (j >= 0 && j < n) s = fmaf(t.w[r + d], 1.0f, s); }
    return s;
  };
  for (int i = 0; i <= kMaxRadius; ++i) { lo[i] = (i <= r) ? full(i) : 1.0f; hi[i] = (i <= r) ? full(n - 1 - r + i) : 1.0f; }
}

// B^T B along one axis of length n (B = the Gaussian row filter with BORDER_REFLECT_101): away from the borders it is the
// autocorrelation of the taps, acorr[s + 2r] = sum_d w[d] w[d + s]; rows q < 2r (lo) and q >= n - 2r (hi) differ and are
// tabulated: lo[q][s + 2r] = (B^T B)[q][q + s], hi[m][s + 2r] = (B^T B)[n - 2r + m][n - 2r + m + s] (0 outside [0, n)).
static void adjoint_tables(const Taps& t, int n, float* acorr, float* lo, float* hi) {
  const int r = t.r, r2 = 2 * r, nt = 4 * r + 1;
  for (int s = -r2; s <= r2; ++s) {
    double v = 0.0;
    for (int d = -r; d <= r; ++d) if (d + s >= -r && d + s <= r) v += (double)t.w[r + d] * (double)t.w[r + d + s];
    acorr[s + r2] = (float)v;
  }
  auto refl = [&](int p) { if (n == 1) return 0; while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p; return p; };
  auto bval = [&](int i, int col) {        // B[i][col]
    double v = 0.0;
    for (int d = -r; d <= r; ++d) if (refl(i + d) == col) v += (double)t.w[r + d];
    return v;
  };
  auto ata = [&](int q, int j) {           // (B^T B)[q][j] = sum_i B[i][q] B[i][j]; B[i][q] != 0 needs |i - q| <= r or a reflection: i <= r - q, i >= 2(n-1) - q - r
    if (q < 0 || q >= n || j < 0 || j >= n) return 0.0;
    double v = 0.0;
    for (int i = 0; i < n; ++i) {
      if (!(std::abs(i - q) <= r || i <= r - q || i >= 2 * (n - 1) - q - r)) continue;
      const double a = bval(i, q);
      if (a != 0.0) v += a * bval(i, j);
    }
    return v;
  };
  for (int q = 0; q < r2; ++q)
    for (int s = -r2; s <= r2; ++s) {
      lo[q * nt + s + r2] = (float)ata(q, q + s);
      const int qh = n - r2 + q;
      hi[q * nt + s + r2] = (float)ata(qh, qh + s);
    }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
    (void)cudaGetLastError();
  }
  return fn;
}
// quad accumulator [planes][H][W] float4 seen as a 3-D f32 tensor {4W, H, planes}; box = the staged cell region of a tile
static bool make_quad_tmap(CUtensorMap* out, float4* base, int W, int H, int planes, int r, int th) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)4 * W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t gstr[2] = {(cuuint64_t)16 * W, (cuuint64_t)16 * W * H};
  const cuuint32_t box[3] = {(cuuint32_t)(4 * fused_qw(r)), (cuuint32_t)fused_qh(r, th), 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  if (box[0] > 256u || box[1] > 256u) return false;
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bearing-vector LUT [H][W] double4 seen as a 2-D tensor of 8-byte words {4W, H}; box = one 32 x 32 source tile (32 KB)
static bool make_lut_tmap(CUtensorMap* out, double4* base, int W, int H) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)4 * W, (cuuint64_t)H};
  const cuuint64_t gstr[1] = {(cuuint64_t)32 * W};
  const cuuint32_t box[2] = {(cuuint32_t)(4 * kBinTile), (cuuint32_t)kBinTile};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static void* fused_kernel(const cmaxb_fe* fe) {
  if (fe->use_tma) return fe->taps.r == 4 ? (void*)fe_eval_fused_kernel<4, true> : (void*)fe_eval_fused_kernel<-1, true>;
  return fe->taps.r == 4 ? (void*)fe_eval_fused_kernel<4, false> : (void*)fe_eval_fused_kernel<-1, false>;
}

// allocate the buffers of lane `li` on first use
static int fe_lane_prepare(cmaxb_fe* fe, int li) {
  FeLane& L = fe->lanes[li];
  if (L.ready) return CMAXB_OK;
  const size_t k = (size_t)fe->kmax, A = (size_t)fe->A;
  if (!L.stream) {
    CMAXB_CUDA_TRY(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    L.own_stream = true;
  }
  CMAXB_TRY(dev_alloc(&L.d_quad[0], k * A));
  CMAXB_TRY(dev_alloc(&L.d_quad[1], k * A));
  CMAXB_TRY(dev_alloc(&L.d_GQ, k * A));
  CMAXB_TRY(dev_alloc(&L.d_sums, k * 8 * kSumStride));
  CMAXB_CUDA_TRY(cudaMemset(L.d_sums, 0, sizeof(double) * k * 8 * kSumStride));
  CMAXB_TRY(dev_alloc(&L.d_ticket, 1));
  CMAXB_TRY(dev_alloc(&L.d_bar, 1));
  CMAXB_CUDA_TRY(cudaMemset(L.d_quad[0], 0, sizeof(float4) * k * A));
  CMAXB_CUDA_TRY(cudaMemset(L.d_quad[1], 0, sizeof(float4) * k * A));
  CMAXB_CUDA_TRY(cudaMemset(L.d_ticket, 0, sizeof(unsigned int)));
  CMAXB_CUDA_TRY(cudaMemset(L.d_bar, 0, sizeof(unsigned long long)));
  CMAXB_CUDA_TRY(cudaHostAlloc((void**)&L.h_fault, sizeof(unsigned long long) * 8, cudaHostAllocMapped));
  CMAXB_CUDA_TRY(cudaHostGetDevicePointer((void**)&L.d_fault, L.h_fault, 0));
  L.h_fault[0] = 0;
  CMAXB_TRY(dev_alloc(&L.d_phase, 16 + 4 * kCtaTraceMax));
  CMAXB_CUDA_TRY(cudaMemset(L.d_phase, 0, sizeof(unsigned long long) * (16 + 4 * kCtaTraceMax)));
  if (fe->use_tma) {
    for (int m = 0; m < 2; ++m)
      for (int b = 0; b < 2; ++b)
        if (!make_quad_tmap(&L.tmap[m][b], L.d_quad[b], fe->cfg.width, fe->cfg.height, fe->kmax, fe->taps.r, fe->th[m]))
          return set_error(CMAXB_ERR_CUDA, "cuTensorMapEncodeTiled failed for the accumulator image");
  }
  CMAXB_CUDA_TRY(cudaDeviceSynchronize());
  L.ready = true;
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_create(const cmaxb_fe_cfg* cfg, cmaxb_fe** out) {
  if (!cfg || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->width < 4 || cfg->height < 4 || cfg->width > 65535 || cfg->height > 65535 || !cfg->lut_xyz || cfg->batch_size <= 0)
    return set_error(CMAXB_ERR_INVALID, "bad front-end configuration");
  if (cfg->grad_mode != CMAXB_GRAD_DENSE && cfg->grad_mode != CMAXB_GRAD_ADJOINT)
    return set_error(CMAXB_ERR_INVALID, "bad grad_mode");
  if (cfg->lanes < 0 || cfg->lanes > kFeMaxLanes - 1 || cfg->packet_slots < 0 || cfg->packet_slots > kFeMaxPackets)
    return set_error(CMAXB_ERR_INVALID, "lanes must be 0..4 and packet_slots 0..16");
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
  fe->nlanes = cfg->lanes > 0 ? cfg->lanes : 3;
  fe->npackets = cfg->packet_slots > 0 ? cfg->packet_slots : 1;
  double grid_fraction = 0.5;
  bool want_tma = true;
  {
    const char* mk = getenv("CMAXB_FE_MULTI_KERNEL");
    fe->force_multi_kernel = mk && mk[0] == '1';
    const char* nb = getenv("CMAXB_FE_NO_BINNING");   // A/B switch: evaluate the packet in arrival (time) order
    fe->use_bins = !(nb && nb[0] == '1');
    const char* tm = getenv("CMAXB_FE_TMA");          // A/B switch: 0 = tiles staged with per-thread loads
    if (tm && tm[0] == '0') want_tma = false;
    const char* ln = getenv("CMAXB_FE_LANES");
    if (ln && atoi(ln) >= 1 && atoi(ln) <= kFeMaxLanes - 1) fe->nlanes = atoi(ln);
    const char* gf = getenv("CMAXB_FE_GRID_FRACTION");   // share of the co-resident CTAs one throughput-lane launch uses
    if (gf && atof(gf) > 0.0 && atof(gf) <= 1.0) grid_fraction = atof(gf);
  }
  int rc = make_taps(cfg->blur_sigma, &fe->taps);
  if (rc != CMAXB_OK) { delete fe; return rc; }
  if (fe->taps.r + 2 > cfg->width || fe->taps.r + 2 > cfg->height) { delete fe; return set_error(CMAXB_ERR_INVALID, "image smaller than the blur kernel"); }
  border_table(fe->taps, cfg->width, fe->cxl, fe->cxr);
  border_table(fe->taps, cfg->height, fe->cyl, fe->cyr);
  {
    const int r2 = 2 * fe->taps.r, nt = 4 * fe->taps.r + 1;
    std::vector<float> tab((size_t)4 * (r2 > 0 ? r2 : 1) * nt, 0.f);
    std::memset(fe->acorr, 0, sizeof(fe->acorr));
    adjoint_tables(fe->taps, cfg->width, fe->acorr, tab.data(), tab.data() + (size_t)r2 * nt);
    adjoint_tables(fe->taps, cfg->height, fe->acorr, tab.data() + (size_t)2 * r2 * nt, tab.data() + (size_t)3 * r2 * nt);
    if (dev_alloc(&fe->d_adj_tab, tab.size()) != CMAXB_OK) { delete fe; return CMAXB_ERR_CUDA; }
    if (cudaMemcpy(fe->d_adj_tab, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(fe->d_adj_tab); delete fe; return set_error(CMAXB_ERR_CUDA, "adjoint table upload failed");
    }
  }
  auto fail = [&](int code) { cmaxb_fe_destroy(fe); return code; };
  if (cfg->stream) fe->lanes[0].stream = (cudaStream_t)cfg->stream;
  else {
    if (cudaStreamCreateWithFlags(&fe->lanes[0].stream, cudaStreamNonBlocking) != cudaSuccess) return fail(set_error(CMAXB_ERR_CUDA, "cudaStreamCreate failed"));
    fe->lanes[0].own_stream = true;
  }
  fe->stream = fe->lanes[0].stream;
  if (fe->prof.init() != CMAXB_OK) return fail(CMAXB_ERR_CUDA);
  if (cudaEventCreateWithFlags(&fe->fork_ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&fe->join_ev, cudaEventDisableTiming) != cudaSuccess) return fail(set_error(CMAXB_ERR_CUDA, "cudaEventCreate failed"));
  // LUT padded to 32-byte records
  {
    std::vector<double4> lut((size_t)fe->A);
    for (long long i = 0; i < fe->A; ++i) lut[i] = make_double4(cfg->lut_xyz[3 * i], cfg->lut_xyz[3 * i + 1], cfg->lut_xyz[3 * i + 2], 0.0);
    if (dev_alloc(&fe->d_lut, (size_t)fe->A) != CMAXB_OK) return fail(CMAXB_ERR_CUDA);
    if (cudaMemcpy(fe->d_lut, lut.data(), sizeof(double4) * fe->A, cudaMemcpyHostToDevice) != cudaSuccess) return fail(set_error(CMAXB_ERR_CUDA, "LUT upload failed"));
  }
  const size_t k = (size_t)fe->kmax;
  bool ok = true;
  for (int s = 0; s < fe->npackets; ++s) {
    // verdict words of the packet in mapped host memory: [0] time order, [1] pixel range (written by the preparation kernels)
    ok = ok && cudaHostAlloc((void**)&fe->packets[s].h_flags, 2 * sizeof(int), cudaHostAllocMapped) == cudaSuccess;
    ok = ok && cudaHostGetDevicePointer((void**)&fe->packets[s].d_flags, fe->packets[s].h_flags, 0) == cudaSuccess;
    if (ok) fe->packets[s].h_flags[0] = fe->packets[s].h_flags[1] = 0;
    ok = ok && cudaEventCreateWithFlags(&fe->packets[s].ready, cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && dev_alloc(&fe->d_omegas, k * 3) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_acc, k * kNAcc * kMaxImgCtas) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_ticket, k) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_ticket2, k) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_gacc, k * 3 * kMaxEventCtas) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_result, k * 4) == CMAXB_OK;
  ok = ok && dev_alloc(&fe->d_mean, k) == CMAXB_OK;
  if (!ok) return fail(set_error(CMAXB_ERR_CUDA, "front-end buffer allocation failed"));
  ok = ok && cudaMallocHost((void**)&fe->h_omegas, sizeof(double) * 3 * k) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&fe->h_result, sizeof(double) * 4 * k) == cudaSuccess;
  ok = ok && cudaMemset(fe->d_ticket, 0, sizeof(unsigned) * k) == cudaSuccess;
  ok = ok && cudaMemset(fe->d_ticket2, 0, sizeof(unsigned) * k) == cudaSuccess;
  ok = ok && cudaMemset(fe->d_result, 0, sizeof(double) * k * 4) == cudaSuccess;
  ok = ok && cudaHostAlloc((void**)&fe->h_ring, sizeof(unsigned long long) * 8 * k * kFeRing, cudaHostAllocMapped) == cudaSuccess;
  ok = ok && cudaHostGetDevicePointer((void**)&fe->d_ring, fe->h_ring, 0) == cudaSuccess;
  if (ok) std::memset(fe->h_ring, 0, sizeof(unsigned long long) * 8 * k * kFeRing);
  if (!ok) return fail(set_error(CMAXB_ERR_CUDA, "front-end buffer allocation failed"));
  // fused evaluation kernel: co-resident grid size from the occupancy API; TMA staging when the tile box fits a tensor map
  {
    int coop = 0, nsm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, fe->device);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, fe->device);
    fe->use_tma = want_tma && tmap_encoder() != nullptr && 4 * fused_qw(fe->taps.r) <= 256 && fused_qh(fe->taps.r, kFusedMaxTH) <= 256 &&
                  make_lut_tmap(&fe->tmap_lut, fe->d_lut, cfg->width, cfg->height);
    const size_t smem_max = fused_smem_bytes(fe->taps.r, kFusedMaxTH);
    int occ = 0;
    void* kern = fused_kernel(fe);
    cudaError_t e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    cudaError_t e3 = cudaErrorUnknown;
    if (e1 == cudaSuccess) {
      if (fe->use_tma) e3 = (fe->taps.r == 4) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fe_eval_fused_kernel<4, true>, kFusedThreads, smem_max)
                                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fe_eval_fused_kernel<-1, true>, kFusedThreads, smem_max);
      else e3 = (fe->taps.r == 4) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fe_eval_fused_kernel<4, false>, kFusedThreads, smem_max)
                                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fe_eval_fused_kernel<-1, false>, kFusedThreads, smem_max);
    }
    (void)cudaGetLastError();
    if (!(coop && e1 == cudaSuccess && e3 == cudaSuccess && occ > 0))
      return fail(set_error(CMAXB_ERR_CUDA, "cooperative launch unavailable: the fused evaluation kernel cannot run on this device"));
    int grid = occ * nsm;
    if (grid > kFusedMaxCtas) grid = kFusedMaxCtas;
    fe->grid[0] = grid;
    int part = (int)(grid * grid_fraction);
    if (part < nsm / 4) part = nsm / 4;
    if (part < 1) part = 1;
    if (part > grid) part = grid;
    fe->grid[1] = part;
    for (int m = 0; m < 2; ++m) {
      fe->th[m] = fused_tile_height(cfg->width, cfg->height, fe->grid[m]);
    }
  }
  rc = fe_lane_prepare(fe, 0);
  if (rc != CMAXB_OK) return fail(rc);
  *out = fe;
  return CMAXB_OK;
}

extern "C" void cmaxb_fe_destroy(cmaxb_fe* fe) {
  if (!fe) return;
  cudaSetDevice(fe->device);
  for (int l = 0; l < kFeMaxLanes; ++l) if (fe->lanes[l].stream) cudaStreamSynchronize(fe->lanes[l].stream);
  cudaFree(fe->d_lut); cudaFree(fe->d_tile_count); cudaFree(fe->d_adj_tab);
  for (int s = 0; s < kFeMaxPackets; ++s) {
    FePacket& pk = fe->packets[s];
    cudaFree(pk.d_ev); cudaFree(pk.d_dt); cudaFree(pk.d_bev); cudaFree(pk.d_tile_end);
    if (pk.h_flags) cudaFreeHost(pk.h_flags);
    if (pk.ready) cudaEventDestroy(pk.ready);
  }
  for (int l = 0; l < kFeMaxLanes; ++l) {
    FeLane& L = fe->lanes[l];
    cudaFree(L.d_quad[0]); cudaFree(L.d_quad[1]); cudaFree(L.d_GQ);
    cudaFree(L.d_sums); cudaFree(L.d_ticket); cudaFree(L.d_bar); cudaFree(L.d_phase);
    if (L.h_fault) cudaFreeHost(L.h_fault);
    if (L.own_stream && L.stream) cudaStreamDestroy(L.stream);
  }
  cudaFree(fe->d_blur1); cudaFree(fe->d_img4); cudaFree(fe->d_blur4);
  cudaFree(fe->d_cells); cudaFree(fe->d_omegas); cudaFree(fe->d_acc); cudaFree(fe->d_ticket); cudaFree(fe->d_ticket2);
  cudaFree(fe->d_gacc); cudaFree(fe->d_result); cudaFree(fe->d_mean);
  if (fe->h_omegas) cudaFreeHost(fe->h_omegas);
  if (fe->h_result) cudaFreeHost(fe->h_result);
  if (fe->h_ring) cudaFreeHost(fe->h_ring);
  for (int r = 0; r < fe->x_world; ++r)
    if (r != fe->x_rank && fe->x_peer[r]) cudaIpcCloseMemHandle(fe->x_peer[r]);
  cudaFree(fe->x_local);
  if (fe->h_xall) cudaFreeHost(fe->h_xall);
  if (fe->h_xerr) cudaFreeHost(fe->h_xerr);
  fe->prof.destroy();
  if (fe->fork_ev) cudaEventDestroy(fe->fork_ev);
  if (fe->join_ev) cudaEventDestroy(fe->join_ev);
  delete fe;
}

// Wait until every lane is idle.  Evaluations already launched stay fetchable (their rows sit in the mapped ring).
static int fe_drain(cmaxb_fe* fe) {
  if (fe->pending) {
    for (int l = 0; l < kFeMaxLanes; ++l)
      if (fe->lanes[l].stream && (l == 0 || fe->lanes[l].ready)) CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->lanes[l].stream));
    fe->pending = false;
  }
  return CMAXB_OK;
}

static int fe_check_packet_flags(cmaxb_fe* fe, FePacket& pk) {
  // validation result of an asynchronous set_packet (its D2H copy precedes every later operation on the stream)
  if (!pk.flags_pending) return CMAXB_OK;
  pk.flags_pending = false;
  if (pk.h_flags[1]) { pk.have = false; return set_error(CMAXB_ERR_EVENT_RANGE, "event pixel outside the sensor"); }
  if (pk.h_flags[0]) { pk.have = false; return set_error(CMAXB_ERR_TIME_ORDER, "Events must span a non-negative time interval"); }
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_select_packet(cmaxb_fe* fe, int slot) {
  if (!fe) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (slot < 0 || slot >= fe->npackets) return set_error(CMAXB_ERR_INVALID, "packet slot out of range (cfg.packet_slots)");
  fe->cur = slot;
  return CMAXB_OK;
}

// view: use the caller's DEVICE buffer in place (no copy; it must stay valid and unchanged until the slot's next set_packet)
static int fe_set_packet_impl(cmaxb_fe* fe, const cmaxb_event* events, size_t n, double t_ref_sec, bool wait, bool view) {
  if (!fe || (!events && n > 0)) return set_error(CMAXB_ERR_INVALID, "null argument");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  FePacket& pk = fe->packets[fe->cur];
  // evaluations of THIS slot's previous packet must have finished before its buffers are rewritten: lane 0 is ordered
  // by its stream; a fetched evaluation has finished; the ones still outstanding on a throughput lane are waited for
  if (pk.users > 0)
    for (const FeInflight& f : fe->inflight)
      if (f.pkt == fe->cur && f.lane > 0) CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->lanes[f.lane].stream));
  pk.have = false;
  const long long bs = fe->cfg.batch_size;
  const long long nb = ((long long)n + bs - 1) / bs;
  if (view) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, events) != cudaSuccess || at.type != cudaMemoryTypeDevice || at.device != fe->device) {
      (void)cudaGetLastError();
      return set_error(CMAXB_ERR_INVALID, "cmaxb_fe_set_packet_view needs device memory of the handle's GPU");
    }
  } else if (n > pk.ev_cap) {
    cudaStreamSynchronize(fe->stream);
    cudaFree(pk.d_ev); pk.d_ev = nullptr; pk.ev_cap = 0;
    CMAXB_TRY(dev_alloc(&pk.d_ev, n));
    pk.ev_cap = n;
  }
  if ((size_t)nb > pk.dt_cap) {
    cudaStreamSynchronize(fe->stream);
    cudaFree(pk.d_dt); pk.d_dt = nullptr; pk.dt_cap = 0;
    CMAXB_TRY(dev_alloc(&pk.d_dt, (size_t)nb));
    pk.dt_cap = (size_t)nb;
  }
  pk.n = (long long)n; pk.nb = nb;
  pk.gen += 1;
  if (n > 0) {
    cudaStream_t s = fe->stream;
    if (view) pk.ev = reinterpret_cast<const uint4*>(events);
    else {
      CMAXB_CUDA_TRY(cudaMemcpyAsync(pk.d_ev, events, sizeof(cmaxb_event) * n, cudaMemcpyDefault, s));   // host (pinned: DMA) or device memory (UVA)
      pk.ev = pk.d_ev;
    }
    // verdict words: the previous packet of this slot must have delivered its own before they are cleared
    if (pk.flags_pending) CMAXB_CUDA_TRY(cudaEventSynchronize(pk.ready));
    pk.h_flags[0] = pk.h_flags[1] = 0;
    const uint4* ev = pk.ev; const long long nn = pk.n; int* flags = pk.d_flags;
    const int W = fe->cfg.width, H = fe->cfg.height; double* dt = pk.d_dt; const int ibs = (int)bs;
    // one-time spatial binning of the packet (reused by every evaluation until the next set_packet); its counting pass
    // also validates the pixel range (without bins a stand-alone validation kernel does)
    fe->ntx = (W + kBinTile - 1) / kBinTile;
    fe->ntiles = fe->ntx * ((H + kBinTile - 1) / kBinTile);
    const bool do_bins = fe->use_bins && fe->ntiles <= kBinMaxTiles && nn < (1LL << 32);
    if (!do_bins) {
      CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
        fe_validate_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(ev, nn, W, H, flags);
      }));
    }
    // chunk of the binning kernels: a multiple of the batch size, so that the counting pass also forms the batch time offsets
    const int chunk = (do_bins && ibs <= kBinChunk) ? (kBinChunk / ibs) * ibs : kBinChunk;
    const bool dt_in_count = do_bins && chunk % ibs == 0;
    if (!dt_in_count) {
      CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
        fe_batch_dt_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(ev, nn, ibs, t_ref_sec, dt, nb, flags);
      }));
    }
    pk.have_bins = false;
    if (do_bins) {
      if (n > pk.bev_cap) {
        cudaStreamSynchronize(s);
        cudaFree(pk.d_bev); pk.d_bev = nullptr; pk.bev_cap = 0;
        CMAXB_TRY(dev_alloc(&pk.d_bev, n));
        pk.bev_cap = n;
      }
      if (!fe->d_tile_count) {              // counts, cursors, ticket: zero here, left zero by every binning scatter pass
        CMAXB_TRY(dev_alloc(&fe->d_tile_count, (size_t)2 * kBinMaxTiles + 1));
        CMAXB_CUDA_TRY(cudaMemsetAsync(fe->d_tile_count, 0, sizeof(unsigned int) * (2 * kBinMaxTiles + 1), s));
      }
      if (!pk.d_tile_end) CMAXB_TRY(dev_alloc(&pk.d_tile_end, (size_t)kBinMaxTiles));
      const int ntiles = fe->ntiles, ntx = fe->ntx;
      const unsigned nchunks = (unsigned)((nn + chunk - 1) / chunk);
      uint4* bev = pk.d_bev;
      unsigned int* count = fe->d_tile_count;
      unsigned int* cursor = fe->d_tile_count + kBinMaxTiles;
      CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
        fe_bin_count_kernel<<<nchunks, kBinThreads, sizeof(unsigned int) * ntiles, s>>>(ev, nn, chunk, W, H, ntx, ntiles, count, flags, ibs,
                                                                                       t_ref_sec, dt_in_count ? dt : nullptr, nb);
      }));
      CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
        fe_bin_scatter_kernel<<<nchunks, kBinThreads, 2 * sizeof(unsigned int) * ntiles, s>>>(ev, nn, chunk, W, H, ntx, ntiles, ibs, dt, count,
                                                                                              cursor, pk.d_tile_end, bev, fe->d_tile_count + 2 * kBinMaxTiles);
      }));
      pk.have_bins = true;
    }
    CMAXB_CUDA_TRY(cudaEventRecord(pk.ready, s));
    pk.flags_pending = true;
    pk.have = true;
    fe->pending = true;
    if (wait) {
      CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
      return fe_check_packet_flags(fe, pk);
    }
    return CMAXB_OK;
  }
  pk.ev = pk.d_ev;
  CMAXB_CUDA_TRY(cudaEventRecord(pk.ready, fe->stream));
  pk.have = true;
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_set_packet(cmaxb_fe* fe, const cmaxb_event* events, size_t n, double t_ref_sec) {
  return fe_set_packet_impl(fe, events, n, t_ref_sec, true, false);
}
extern "C" int cmaxb_fe_set_packet_async(cmaxb_fe* fe, const cmaxb_event* events, size_t n, double t_ref_sec) {
  return fe_set_packet_impl(fe, events, n, t_ref_sec, false, false);
}
extern "C" int cmaxb_fe_set_packet_view(cmaxb_fe* fe, const cmaxb_event* device_events, size_t n, double t_ref_sec) {
  return fe_set_packet_impl(fe, device_events, n, t_ref_sec, false, true);
}

static int fe_upload_omegas(cmaxb_fe* fe, const double* omegas, int k) {
  CMAXB_TRY(fe_drain(fe));
  std::memcpy(fe->h_omegas, omegas, sizeof(double) * 3 * k);
  CMAXB_CUDA_TRY(cudaMemcpyAsync(fe->d_omegas, fe->h_omegas, sizeof(double) * 3 * k, cudaMemcpyHostToDevice, fe->stream));
  return CMAXB_OK;
}

static dim3 fe_event_grid(long long n, int k) {
  long long blocks = (n + kFeThreads - 1) / kFeThreads;
  const long long cap = kMaxEventCtas;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return dim3((unsigned)blocks, (unsigned)k, 1);
}

// ---- stand-alone kernels on lane 0 (DENSE gradients, getters, A/B) ------------------------------------------
// scatter the value votes of k hypotheses into lane 0's current quad accumulator
static int fe_run_scatter_value(cmaxb_fe* fe, int k) {
  FeLane& L = fe->lanes[0];
  const FePacket& pk = fe->packets[fe->cur];
  cudaStream_t s = fe->stream;
  const int cur = L.quad_cur;
  if (L.quad_dirty[cur] > 0) {   // only after start-up / a change of k / a debug call: steady state skips this
    const int planes = L.quad_dirty[cur];
    CMAXB_TRY(fe->prof.run(CMAXB_K_ZERO, s, false, [&] { cudaMemsetAsync(L.d_quad[cur], 0, sizeof(float4) * fe->A * planes, s); }));
    L.quad_dirty[cur] = 0;
  }
  if (pk.n > 0) {
    const FeGeom g = fe_geom(fe, pk);
    CMAXB_TRY(fe->prof.run(CMAXB_K_FE_SCATTER, s, true, [&] {
      fe_scatter_kernel<2><<<fe_event_grid(pk.n, k), kFeThreads, 0, s>>>(g, fe->d_omegas, nullptr, L.d_quad[cur], fe->A);
    }));
    L.quad_dirty[cur] = k;
  }
  return CMAXB_OK;
}
// blur + reduce lane 0's current quad accumulator (k planes); optionally keep the blurred image; clears the
// OTHER quad accumulator and makes it current.
static int fe_run_value_image(cmaxb_fe* fe, int k, const Taps& taps, bool write_out) {
  FeLane& L = fe->lanes[0];
  cudaStream_t s = fe->stream;
  const int cur = L.quad_cur, oth = cur ^ 1;
  const SrcQuad src{L.d_quad[cur], fe->A};
  const ReduceOut ro{fe->d_acc, fe->d_ticket, fe->d_result, fe->d_mean};
  const int W = fe->cfg.width, H = fe->cfg.height, measure = fe->cfg.contrast_measure;
  // the other image can be cleared by this kernel if its dirty planes are covered by our k planes
  float4* zero_ptr = (L.quad_dirty[oth] > 0 && L.quad_dirty[oth] <= k) ? L.d_quad[oth] : nullptr;
  if (write_out && !fe->d_blur1) CMAXB_TRY(dev_alloc(&fe->d_blur1, (size_t)fe->kmax * fe->A));
  cudaError_t le = cudaSuccess;
  CMAXB_TRY(fe->prof.run(CMAXB_K_BLUR_REDUCE, s, true, [&] {
    le = write_out ? launch_blur_reduce<1, SrcQuad, true>(s, k, src, W, H, taps, fe->d_blur1, fe->A, ro, measure, zero_ptr, fe->A)
                   : launch_blur_reduce<1, SrcQuad, false>(s, k, src, W, H, taps, nullptr, 0, ro, measure, zero_ptr, fe->A);
  }));
  if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("blur_reduce launch: ") + cudaGetErrorString(le));
  if (zero_ptr) L.quad_dirty[oth] = 0;
  L.quad_cur = oth;
  return CMAXB_OK;
}
static int fe_run_scatter_dense(cmaxb_fe* fe, int k) {
  const FePacket& pk = fe->packets[fe->cur];
  cudaStream_t s = fe->stream;
  if (!fe->d_img4) CMAXB_TRY(dev_alloc(&fe->d_img4, (size_t)fe->kmax * fe->A));
  CMAXB_TRY(fe->prof.run(CMAXB_K_ZERO, s, false, [&] { cudaMemsetAsync(fe->d_img4, 0, sizeof(float4) * fe->A * k, s); }));
  if (pk.n > 0) {
    const FeGeom g = fe_geom(fe, pk);
    CMAXB_TRY(fe->prof.run(CMAXB_K_FE_SCATTER, s, true, [&] {
      fe_scatter_kernel<1><<<fe_event_grid(pk.n, k), kFeThreads, 0, s>>>(g, fe->d_omegas, nullptr, fe->d_img4, fe->A);
    }));
  }
  return CMAXB_OK;
}

// One cooperative launch per <= kFusedMaxHyp hypotheses; omegas travel as kernel parameters and the results land
// in a ring slot of mapped pinned memory (up to kFeRing launches may be outstanding).  mode 0: lane 0, whole grid;
// mode 1: next throughput lane, partial grid.
static int fe_eval_launch_fused(cmaxb_fe* fe, const double* omegas, int k, int want_grad, int mode) {
  if ((int)fe->inflight.size() >= kFeRing)
    return set_error(CMAXB_ERR_STATE, "too many outstanding evaluations: call cmaxb_fe_eval_fetch first");
  if (!fe->inflight.empty() && !fe->inflight.back().fused) CMAXB_TRY(fe_drain(fe));
  if (fe->x_on && k > kFusedMaxHyp)
    return set_error(CMAXB_ERR_INVALID, "result exchange supports at most 32 hypotheses per launch");
  int li = 0;
  if (mode == 1 && fe->nlanes > 1) { li = 1 + fe->lane_next; fe->lane_next = (fe->lane_next + 1) % fe->nlanes; }
  CMAXB_TRY(fe_lane_prepare(fe, li));
  FeLane& L = fe->lanes[li];
  FePacket& pk = fe->packets[fe->cur];
  cudaStream_t s = L.stream;
  if (li > 0 && L.seen_gen[fe->cur] != pk.gen) {      // order this lane's stream after the packet's preparation kernels
    CMAXB_CUDA_TRY(cudaStreamWaitEvent(s, pk.ready, 0));
    L.seen_gen[fe->cur] = pk.gen;
  }
  const int gm = li > 0 ? 1 : 0;
  const int grid = fe->grid[gm], th = fe->th[gm];
  const int slot = fe->ring_next;
  fe->ring_next = (fe->ring_next + 1) % kFeRing;
  fe->launch_no += 1;
  if ((fe->launch_no & 0xffffffffull) == 0) fe->launch_no += 1;       // tag 0 = "never written"
  const unsigned long long tag = (fe->launch_no & 0xffffffffull) << 32;
  const int cur = L.quad_cur, oth = cur ^ 1;
  if (L.quad_dirty[cur] > 0) {
    const int planes = L.quad_dirty[cur];
    CMAXB_TRY(fe->prof.run(CMAXB_K_ZERO, s, false, [&] { cudaMemsetAsync(L.d_quad[cur], 0, sizeof(float4) * fe->A * planes, s); }));
    L.quad_dirty[cur] = 0;
  }
  const bool clear_next = L.quad_dirty[oth] > 0 && L.quad_dirty[oth] <= k;
  const int W = fe->cfg.width, H = fe->cfg.height;
  for (int c0 = 0; c0 < k; c0 += kFusedMaxHyp) {
    const int kc = (k - c0 < kFusedMaxHyp) ? k - c0 : kFusedMaxHyp;
    FeFusedParams p;
    p.g = fe_geom(fe, pk);
    p.bev = pk.have_bins ? pk.d_bev : nullptr;     // the fused kernel walks the tile-binned copy of the packet
    p.tile_end = pk.d_tile_end; p.bin_ntx = fe->ntx; p.bin_ntiles = fe->ntiles;
    p.k = kc; p.th = th; p.ntx = (W + kTW - 1) / kTW; p.nty = (H + th - 1) / th;
    p.want_grad = want_grad; p.measure = fe->cfg.contrast_measure; 
    p.quad_plane0 = c0;
    p.taps = fe->taps;
    std::memcpy(p.cxl, fe->cxl, sizeof(p.cxl)); std::memcpy(p.cxr, fe->cxr, sizeof(p.cxr));
    std::memcpy(p.cyl, fe->cyl, sizeof(p.cyl)); std::memcpy(p.cyr, fe->cyr, sizeof(p.cyr));
    std::memcpy(p.acorr, fe->acorr, sizeof(p.acorr)); p.adj_tab = fe->d_adj_tab;
    for (int i = 0; i < 3 * kc; ++i) p.omegas[i] = omegas[3 * c0 + i];
    p.quad = L.d_quad[cur] + (long long)c0 * fe->A;
    p.quad_next = clear_next ? L.d_quad[oth] + (long long)c0 * fe->A : nullptr;
    p.GQ = L.d_GQ + (long long)c0 * fe->A;
    p.A = fe->A;
    p.sums = L.d_sums + (long long)c0 * 8 * kSumStride;
    p.ticket = L.d_ticket;
    p.bar = L.d_bar; p.bar_base = L.bar_count;
    L.bar_count += (unsigned long long)grid * (want_grad ? 2ull : 1ull);
    p.result = fe->d_ring + ((long long)slot * fe->kmax + c0) * 8;
    p.tag = tag;
    p.mirror = fe->d_mirror ? fe->d_mirror + 4 * c0 : nullptr;
    p.fault_flag = L.d_fault;
    p.phase_ns = fe->prof.enabled ? L.d_phase : nullptr;
    std::memset(&p.x, 0, sizeof(p.x));
    if (fe->x_on) {
      p.x.world = fe->x_world; p.x.rank = fe->x_rank; p.x.kmax = fe->kmax;
      p.x.seq = ++fe->x_seq;
      for (int r = 0; r < fe->x_world; ++r) p.x.peer[r] = fe->x_peer[r];
      p.x.all_host = fe->d_xall + (long long)slot * fe->x_world * fe->kmax * 8;
      p.x.all_dev = fe->x_all_dev;
      p.x.err = fe->d_xerr;
    }
    if (fe->prof.enabled) CMAXB_CUDA_TRY(cudaMemsetAsync(L.d_phase, 0, sizeof(unsigned long long) * 16, s));
    void* args[] = {&p, &L.tmap[gm][cur], &fe->tmap_lut};
    const size_t smem = fused_smem_bytes(fe->taps.r, th);
    cudaError_t le = cudaSuccess;
    CMAXB_TRY(fe->prof.run(CMAXB_K_FE_EVAL_FUSED, s, true, [&] {
      le = cudaLaunchCooperativeKernel(fused_kernel(fe), dim3(grid), dim3(kFusedThreads), args, smem, s);
    }));
    if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("fused evaluation launch: ") + cudaGetErrorString(le));
  }
  if (clear_next) L.quad_dirty[oth] = 0;
  if (pk.n > 0) L.quad_dirty[cur] = k;
  L.quad_cur = oth;
  L.inflight += 1;
  pk.users += 1;
  fe->inflight.push_back(FeInflight{k, want_grad != 0, true, slot, li, tag, fe->x_on, fe->cur});
  fe->pending = true;
  return CMAXB_OK;
}

static int fe_eval_launch_mode(cmaxb_fe* fe, const double* omegas, int k, int want_grad, int mode) {
  if (!fe || !omegas) return set_error(CMAXB_ERR_INVALID, "null argument");
  FePacket& pk = fe->packets[fe->cur];
  if (!pk.have) return set_error(CMAXB_ERR_STATE, "no event packet: call cmaxb_fe_set_packet first");
  if (k < 1 || k > fe->kmax) return set_error(CMAXB_ERR_INVALID, "k exceeds cfg.max_hypotheses");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  if (!(want_grad && fe->cfg.grad_mode == CMAXB_GRAD_DENSE) && !fe->force_multi_kernel)
    return fe_eval_launch_fused(fe, omegas, k, want_grad, mode);
  // multi-kernel pipeline (DENSE gradients, A/B runs): one evaluation outstanding at a time
  if (!fe->inflight.empty())
    return set_error(CMAXB_ERR_STATE, "multi-kernel evaluation: fetch the outstanding evaluation first");
  CMAXB_TRY(fe_upload_omegas(fe, omegas, k));
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
      cudaError_t le = cudaSuccess;
      CMAXB_TRY(fe->prof.run(CMAXB_K_ADJOINT_BLUR, s, true, [&] {
        le = launch_adjoint_blur<true>(s, k, fe->d_blur1, fe->A, W, H, fe->taps, fe->d_mean, measure, nullptr, fe->lanes[0].d_GQ);
      }));
      if (le != cudaSuccess) return set_error(CMAXB_ERR_CUDA, std::string("adjoint_blur launch: ") + cudaGetErrorString(le));
      const FeGeom g = fe_geom(fe, pk);
      CMAXB_TRY(fe->prof.run(CMAXB_K_FE_GATHER, s, true, [&] {
        fe_gather_kernel<true><<<fe_event_grid(pk.n, k), kFeThreads, 0, s>>>(g, fe->d_omegas, nullptr, fe->lanes[0].d_GQ, fe->A, fe->d_gacc, fe->d_ticket2, fe->d_result);
      }));
    }
  }
  CMAXB_CUDA_TRY(cudaMemcpyAsync(fe->h_result, fe->d_result, sizeof(double) * 4 * k, cudaMemcpyDeviceToHost, s));
  pk.users += 1;
  fe->inflight.push_back(FeInflight{k, want_grad != 0, false, 0, 0, 0ull, false, fe->cur});
  fe->pending = true;
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_eval_launch(cmaxb_fe* fe, const double* omegas, int k, int want_grad) {
  return fe_eval_launch_mode(fe, omegas, k, want_grad, 1);
}

// All `count` doubles of a tagged-word block carry `tag`?  (see ll_store in fe_fused.cuh)
static bool ll_ready(const volatile unsigned long long* words, int count, unsigned long long tag) {
  for (int i = 2 * count - 1; i >= 0; --i)
    if ((words[i] & 0xffffffff00000000ull) != tag) return false;
  return true;
}
static double ll_value(const volatile unsigned long long* words, int i) {
  const unsigned long long bits = (words[2 * i] & 0xffffffffull) | (words[2 * i + 1] << 32);
  double v;
  std::memcpy(&v, &bits, sizeof(v));
  return v;
}

// waits for the OLDEST outstanding launch and pops it; *out = its record
static int fe_wait_oldest(cmaxb_fe* fe, FeInflight* out) {
  if (fe->inflight.empty()) return set_error(CMAXB_ERR_STATE, "no evaluation launched");
  const FeInflight f = fe->inflight.front();
  auto abandon = [&](const std::string& msg) {
    fe->inflight.clear();
    for (auto& pp : fe->packets) pp.users = 0;
    for (auto& ll : fe->lanes) ll.inflight = 0;
    return set_error(CMAXB_ERR_CUDA, msg);
  };
  if (f.fused) {
    // The fused kernel publishes its rows as tagged words in mapped pinned memory: poll the words themselves (no
    // completion flag, no stream synchronisation); check the stream now and then so that a faulted kernel cannot
    // hang the caller.
    FeLane& L = fe->lanes[f.lane];
    const volatile unsigned long long* words = f.xchg ? fe->h_xall + (long long)f.slot * fe->x_world * fe->kmax * 8
                                                      : fe->h_ring + (long long)f.slot * fe->kmax * 8;
    const int count = f.xchg ? 4 * f.k * fe->x_world : 4 * f.k;
    unsigned long long spins = 0;
    bool ready = false;
    while (!(ready = ll_ready(words, count, f.tag))) {
      if ((++spins & 0x3fff) == 0) {
        cudaError_t q = cudaStreamQuery(L.stream);
        if (q == cudaSuccess) break;                       // finished (the words raced the query) or faulted
        if (q != cudaErrorNotReady) return abandon(std::string("fused evaluation kernel: ") + cudaGetErrorString(q));
      }
    }
    if (!ready) {
      CMAXB_CUDA_TRY(cudaStreamSynchronize(L.stream));
      if (!ll_ready(words, count, f.tag)) return abandon("fused evaluation kernel finished without publishing its result");
    }
    L.inflight -= 1;
    if (L.h_fault[0]) {
      const unsigned long long why = L.h_fault[0];
      L.h_fault[0] = 0;
      return abandon(why == 2 ? "fused evaluation kernel: tile copy (TMA) timed out" : "fused evaluation kernel: grid barrier timed out");
    }
  } else {
    CMAXB_CUDA_TRY(cudaStreamSynchronize(fe->stream));
  }
  fe->inflight.pop_front();
  fe->packets[f.pkt].users -= 1;
  *out = f;
  return fe_check_packet_flags(fe, fe->packets[f.pkt]);
}

extern "C" int cmaxb_fe_eval_fetch(cmaxb_fe* fe, double* contrasts, double* grads3k) {
  if (!fe || !contrasts) return set_error(CMAXB_ERR_INVALID, "null argument");
  FeInflight f;
  CMAXB_TRY(fe_wait_oldest(fe, &f));
  if (f.xchg && *fe->h_xerr) return set_error(CMAXB_ERR_CUDA, "result exchange: a peer's rows did not arrive (timeout)");
  if (f.fused) {
    // own rows: the result ring (without exchange) or this rank's block of the gathered rows
    const volatile unsigned long long* words = f.xchg
        ? fe->h_xall + (long long)f.slot * fe->x_world * fe->kmax * 8 + (long long)fe->x_rank * f.k * 8     // rows are packed [world][k][4]
        : fe->h_ring + (long long)f.slot * fe->kmax * 8;
    for (int h = 0; h < f.k; ++h) {
      contrasts[h] = ll_value(words, 4 * h);
      if (grads3k && f.grad)
        for (int c = 0; c < 3; ++c) grads3k[3 * h + c] = ll_value(words, 4 * h + 1 + c);
    }
    return CMAXB_OK;
  }
  const double* res = fe->h_result;
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
  CMAXB_TRY(fe_drain(fe));
  if (fe->x_local) return set_error(CMAXB_ERR_STATE, "exchange already initialised");
  const size_t need = sizeof(unsigned long long) * 8 * (size_t)kXSlots * world * fe->kmax;
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
  CMAXB_TRY(fe_drain(fe));
  for (int r = 0; r < fe->x_world; ++r) {
    if (r == fe->x_rank) { fe->x_peer[r] = fe->x_local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + 64 * r, 64);
    void* ptr = nullptr;
    CMAXB_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    fe->x_peer[r] = (unsigned long long*)ptr;
  }
  if (!fe->h_xall) {
    const size_t xbytes = sizeof(unsigned long long) * 8 * (size_t)fe->kmax * fe->x_world * kFeRing;
    CMAXB_CUDA_TRY(cudaHostAlloc((void**)&fe->h_xall, xbytes, cudaHostAllocMapped));
    std::memset(fe->h_xall, 0, xbytes);
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
  fe->pending = true;
  fe_drain(fe);
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
  const volatile unsigned long long* all = fe->h_xall + (long long)f.slot * fe->x_world * fe->kmax * 8;
  for (int i = 0; i < 4 * f.k * fe->x_world; ++i) rows[i] = ll_value(all, i);
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_eval_batch(cmaxb_fe* fe, const double* omegas, int k, double* contrasts, double* grads3k) {
  if (fe && !fe->inflight.empty()) return set_error(CMAXB_ERR_STATE, "outstanding cmaxb_fe_eval_launch calls: fetch them first");
  CMAXB_TRY(fe_eval_launch_mode(fe, omegas, k, grads3k != nullptr, 0));
  return cmaxb_fe_eval_fetch(fe, contrasts, grads3k);
}

extern "C" int cmaxb_fe_eval(cmaxb_fe* fe, const double omega[3], double* contrast, double* grad3) {
  return cmaxb_fe_eval_batch(fe, omega, 1, contrast, grad3);
}

extern "C" int cmaxb_fe_get_iwe(cmaxb_fe* fe, const double omega[3], int blurred, float* out) {
  if (!fe || !omega || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!fe->packets[fe->cur].have) return set_error(CMAXB_ERR_STATE, "no event packet");
  if (!fe->inflight.empty()) return set_error(CMAXB_ERR_STATE, "outstanding cmaxb_fe_eval_launch calls: fetch them first");
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
  if (!fe->packets[fe->cur].have) return set_error(CMAXB_ERR_STATE, "no event packet");
  if (!fe->inflight.empty()) return set_error(CMAXB_ERR_STATE, "outstanding cmaxb_fe_eval_launch calls: fetch them first");
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
  const FePacket& pk = fe->packets[fe->cur];
  if (!pk.have) return set_error(CMAXB_ERR_STATE, "no event packet");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  if (pk.n == 0) return CMAXB_OK;
  CMAXB_TRY(fe_upload_omegas(fe, omega, 1));
  if ((size_t)pk.n > fe->cells_cap) {
    cudaFree(fe->d_cells); fe->d_cells = nullptr; fe->cells_cap = 0;
    CMAXB_TRY(dev_alloc(&fe->d_cells, (size_t)pk.n));
    fe->cells_cap = (size_t)pk.n;
  }
  cudaStream_t s = fe->stream;
  const FeGeom g = fe_geom(fe, pk);
  CMAXB_TRY(fe->prof.run(CMAXB_K_MISC, s, true, [&] {
    fe_cells_kernel<<<(unsigned)((pk.n + 255) / 256), 256, 0, s>>>(g, fe->d_omegas, fe->d_cells);
  }));
  CMAXB_CUDA_TRY(cudaMemcpyAsync(out, fe->d_cells, sizeof(int) * pk.n, cudaMemcpyDeviceToHost, s));
  CMAXB_CUDA_TRY(cudaStreamSynchronize(s));
  return CMAXB_OK;
}

// lanes_fork: every throughput lane's stream waits for the work queued so far on the main stream (e.g. a timing event);
// lanes_join: the main stream waits for everything queued so far on the throughput lanes.
extern "C" int cmaxb_fe_lanes_fork(cmaxb_fe* fe) {
  if (!fe) return set_error(CMAXB_ERR_INVALID, "null argument");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  if (fe->nlanes <= 1) return CMAXB_OK;
  CMAXB_CUDA_TRY(cudaEventRecord(fe->fork_ev, fe->stream));
  for (int l = 1; l <= fe->nlanes; ++l) {
    CMAXB_TRY(fe_lane_prepare(fe, l));
    CMAXB_CUDA_TRY(cudaStreamWaitEvent(fe->lanes[l].stream, fe->fork_ev, 0));
  }
  return CMAXB_OK;
}
extern "C" int cmaxb_fe_lanes_join(cmaxb_fe* fe) {
  if (!fe) return set_error(CMAXB_ERR_INVALID, "null argument");
  CMAXB_CUDA_TRY(cudaSetDevice(fe->device));
  if (fe->nlanes <= 1) return CMAXB_OK;
  for (int l = 1; l <= fe->nlanes; ++l) {
    if (!fe->lanes[l].ready) continue;
    CMAXB_CUDA_TRY(cudaEventRecord(fe->join_ev, fe->lanes[l].stream));
    CMAXB_CUDA_TRY(cudaStreamWaitEvent(fe->stream, fe->join_ev, 0));
  }
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
  CMAXB_TRY(fe_drain(fe));
  fe->d_mirror = device_ptr;
  return CMAXB_OK;
}
extern "C" int cmaxb_fe_phase_times(cmaxb_fe* fe, double* us10) {
  if (!fe || !us10) return set_error(CMAXB_ERR_INVALID, "null argument");
  fe->pending = true;
  CMAXB_TRY(fe_drain(fe));
  unsigned long long ph[16];
  CMAXB_CUDA_TRY(cudaMemcpy(ph, fe->lanes[0].d_phase, sizeof(ph), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 10; ++i) us10[i] = (ph[i] && ph[0]) ? (double)(ph[i] - ph[0]) * 1e-3 : -1.0;
  return CMAXB_OK;
}
// profiling aid: per-CTA stamps of the last profiled whole-grid launch, us since kernel entry: out[cta][4] = scatter end, image
// phase end, gather start, gather end (0 where not reached); *n_ctas = CTAs of the launch (at most 1024 are traced)
extern "C" int cmaxb_fe_cta_times(cmaxb_fe* fe, double* out, int max_ctas, int* n_ctas) {
  if (!fe || !out || !n_ctas) return set_error(CMAXB_ERR_INVALID, "null argument");
  fe->pending = true;
  CMAXB_TRY(fe_drain(fe));
  std::vector<unsigned long long> ph(16 + 4 * kCtaTraceMax);
  CMAXB_CUDA_TRY(cudaMemcpy(ph.data(), fe->lanes[0].d_phase, sizeof(unsigned long long) * ph.size(), cudaMemcpyDeviceToHost));
  const int n = std::min(std::min(fe->grid[0], kCtaTraceMax), max_ctas);
  for (int c = 0; c < n; ++c)
    for (int w = 0; w < 4; ++w) {
      const unsigned long long v = ph[16 + 4 * c + w];
      out[4 * c + w] = (v && ph[0]) ? (double)(v - ph[0]) * 1e-3 : 0.0;
    }
  *n_ctas = n;
  return CMAXB_OK;
}
extern "C" int cmaxb_fe_kernel_times(cmaxb_fe* fe, double* ms, uint64_t* launches) {
  if (!fe || !ms || !launches) return set_error(CMAXB_ERR_INVALID, "null argument");
  for (int i = 0; i < CMAXB_K_COUNT; ++i) { ms[i] = fe->prof.ms[i]; launches[i] = fe->prof.launches[i]; }
  return CMAXB_OK;
}
/* geometry of the fused launches: [0] whole-grid CTAs, [1] throughput-lane CTAs, [2] lanes, [3] TMA staging on/off,
 * [4] tile height (whole grid), [5] tile height (lane), [6] gather records in use by the last policy (-1 auto, 0, 1) */
extern "C" int cmaxb_fe_launch_info(cmaxb_fe* fe, int32_t* info7) {
  if (!fe || !info7) return set_error(CMAXB_ERR_INVALID, "null argument");
  info7[0] = fe->grid[0]; info7[1] = fe->grid[1]; info7[2] = fe->nlanes; info7[3] = fe->use_tma ? 1 : 0;
  info7[4] = fe->th[0]; info7[5] = fe->th[1]; info7[6] = 0;
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
                                             "fe_eval_fused", "be_eval_fused", "be_x_push", "be_x_sums", "be_x_pull", "be_x_grad"};
  return (kind >= 0 && kind < CMAXB_K_COUNT) ? names[kind] : "?";
}
