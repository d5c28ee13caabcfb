/*
 * cmax_b200.h -- C ABI of libcmax_b200.so: the B200-native (sm_100a) contrast-maximisation
 * inner loop of CMax-SLAM (reference: tub-rip/cmax_slam @ 12342de).
 *
 * The library replaces, beneath the reference's GSL f/df/fdf cost callbacks, the pair of calls
 *   computeImageOfWarpedEvents(...) + computeContrast(...)
 * made by
 *   front-end: local_contrast_fdf            src/frontend/local_optim_contrast_gsl.cpp:20-56
 *   back-end : global_contrast_fdf           src/backend/global_optim_contrast_gsl_analytical.cpp:17-68
 * Plain C: pointers and sizes only, no C++ / OpenCV / ROS / GSL / torch types.  Every function
 * returns 0 on success or a negative cmaxb_status; the message is in cmaxb_last_error()
 * (thread local).  The library returns +contrast / +gradient; the GSL adapter negates, as the
 * reference does (local_optim_contrast_gsl.cpp:48-54).  There is NO CPU fallback: without a CUDA
 * device every create() fails with CMAXB_ERR_CUDA.
 *
 * Threading: one handle = one CUDA stream + private device buffers.  Calls on different handles
 * may run concurrently (FE thread / BE thread, src/cmax_slam.cpp:92); one handle is not
 * re-entrant (neither are the reference functions: function-local statics).
 */
#ifndef CMAX_B200_H_
#define CMAX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMAXB_VERSION 200

typedef enum cmaxb_status {
  CMAXB_OK = 0,
  CMAXB_ERR_INVALID = -1,      /* bad argument */
  CMAXB_ERR_CUDA = -2,         /* CUDA runtime error / no device */
  CMAXB_ERR_EVENT_RANGE = -3,  /* event pixel outside the sensor (reference: vector::at throws) */
  CMAXB_ERR_TIME_ORDER = -4,   /* a batch spans a negative time (reference: CHECK_GE aborts,
                                  local_image_warped_events.cpp:72) */
  CMAXB_ERR_SPLINE_RANGE = -5, /* batch time outside the spline (reference: BASALT_ASSERT aborts,
                                  so3_spline.h:221-230) */
  CMAXB_ERR_STATE = -6         /* call order (e.g. eval before set_packet) */
} cmaxb_status;

/* dvs_msgs::Event exactly as the ROS C++ generator lays it out (uint16 x, uint16 y,
 * ros::Time ts{uint32 sec, uint32 nsec}, bool polarity, 3 pad bytes): 16 bytes, so a
 * std::vector<dvs_msgs::Event>::data() can be passed as is (event_subset_,
 * include/frontend/ang_vel_estimator.h, include/backend/pose_graph_optimizer.h). */
typedef struct cmaxb_event {
  uint16_t x, y;
  uint32_t sec, nsec;
  uint8_t polarity;
  uint8_t pad_[3];
} cmaxb_event;

enum { CMAXB_CONTRAST_VARIANCE = 0, CMAXB_CONTRAST_MEAN_SQUARE = 1 }; /* local_focus_funcs.h:7-11 */

/* How the analytic gradient is evaluated (identical mathematics, different data movement):
 *   DENSE   : accumulate the derivative images, blur them, reduce -- what the reference does
 *             (local_image_warped_events.cpp:153-167, local_focus_funcs.cpp:34-42).
 *   ADJOINT : g = sum_events sum_corners dw_c * [blur^T(2(I-mu))/Np](corner): one extra image and a
 *             gather pass over the events; no derivative images exist.                       */
enum { CMAXB_GRAD_DENSE = 0, CMAXB_GRAD_ADJOINT = 1 };

/* ------------------------------------------------------------------ front-end ------------ */
typedef struct cmaxb_fe cmaxb_fe;

typedef struct cmaxb_fe_cfg {
  int32_t width, height;     /* cam_width_, cam_height_ */
  double fx, fy, cx, cy;     /* camera_matrix_ (0,0),(1,1),(0,2),(1,2); ang_vel_estimator.cpp:42 */
  const double* lut_xyz;     /* precomputed_bearing_vectors_: width*height*3 doubles, copied */
  double blur_sigma;         /* params.warp_opt.blur_sigma (<=0: no blur) */
  int32_t batch_size;        /* params.warp_opt.event_batch_size */
  int32_t contrast_measure;  /* params.process_opt.contrast_measure */
  int32_t grad_mode;         /* CMAXB_GRAD_* */
  int32_t device;            /* CUDA device ordinal */
  void* stream;              /* optional cudaStream_t to run on; NULL = library-owned stream */
  int32_t max_hypotheses;    /* capacity of eval_batch (<=0: 1) */
  int32_t lanes;             /* throughput lanes of cmaxb_fe_eval_launch: 0 = default (3); 1 = none, every evaluation runs on
                                the handle's stream with the whole GPU; 2..4 = that many library-owned streams, each launch on
                                a partial grid so that consecutive evaluations overlap on the device */
  int32_t packet_slots;      /* resident packets (cmaxb_fe_select_packet); 0 = 1 */
} cmaxb_fe_cfg;

int cmaxb_fe_create(const cmaxb_fe_cfg* cfg, cmaxb_fe** out);
void cmaxb_fe_destroy(cmaxb_fe* fe);

/* Upload one event packet (replaces the copy into event_subset_, ang_vel_estimator.cpp:137-147)
 * and its reference time time_packet_.toSec().  `events` is host memory (pinned memory makes
 * the copy asynchronous) or device memory of the handle's GPU (e.g. the output of an NVLink all-gather of packet
 * shards uploaded by several ranks: the copy is then device to device).  Validates pixel range and batch time
 * order. */
int cmaxb_fe_set_packet(cmaxb_fe* fe, const cmaxb_event* events, size_t n, double t_ref_sec);
/* Same, but returns as soon as the copy and the preparation kernels are queued on the handle's stream (events
 * must stay valid -- pinned -- until the next call that waits).  The validation verdict is delivered by the next
 * cmaxb_fe_eval / cmaxb_fe_eval_fetch.  Lets the upload of packet i+1 (on another handle / stream) overlap the
 * evaluations of packet i. */
int cmaxb_fe_set_packet_async(cmaxb_fe* fe, const cmaxb_event* events, size_t n, double t_ref_sec);

/* Same as _async for events that ALREADY live in device memory of the handle's GPU (e.g. a device-resident event store,
 * cmaxb_stream_*_device): nothing is copied, the library reads the caller's buffer in place, which must stay valid and
 * unchanged until the slot's next set_packet.  Overlapping packets of one store so cross PCIe once. */
int cmaxb_fe_set_packet_view(cmaxb_fe* fe, const cmaxb_event* device_events, size_t n, double t_ref_sec);
/* A handle holds cfg.packet_slots resident packets; set_packet*, the evaluations and the getters act on the selected
 * slot (default 0).  An evaluation keeps the slot it was launched on, so packet i+1 can be uploaded into another slot
 * while packet i is still being evaluated. */
int cmaxb_fe_select_packet(cmaxb_fe* fe, int slot);

/* One cost evaluation = computeImageOfWarpedEvents + computeContrast.  grad3 == NULL => value
 * only (the local_contrast_f path, local_optim_contrast_gsl.cpp:58-63). */
int cmaxb_fe_eval(cmaxb_fe* fe, const double omega[3], double* contrast, double* grad3);

/* k hypotheses on the resident packet in one pass over the events (BASELINE config C3). */
int cmaxb_fe_eval_batch(cmaxb_fe* fe, const double* omegas, int k, double* contrasts, double* grads3k);

/* Asynchronous pair used by the multi-GPU driver, batched line searches and the benchmark: launch returns as
 * soon as the work is queued on the stream; fetch waits for the OLDEST outstanding launch and returns its
 * results (FIFO).  Up to CMAXB_FE_MAX_OUTSTANDING launches may be queued before the first fetch (each has its
 * own result slot in mapped host memory), so consecutive evaluations run back to back on the device with no
 * host round trip between them; one more launch returns CMAXB_ERR_STATE. */
#define CMAXB_FE_MAX_OUTSTANDING 8
int cmaxb_fe_eval_launch(cmaxb_fe* fe, const double* omegas, int k, int want_grad);
int cmaxb_fe_eval_fetch(cmaxb_fe* fe, double* contrasts, double* grads3k);
/* With cfg.lanes >= 2 the launches above run on library-owned streams.  _fork makes those streams wait for the work
 * queued so far on the handle's stream (cfg.stream), _join makes the handle's stream wait for the launches queued so
 * far -- e.g. to bracket a batch of launches with events recorded on the handle's stream. */
int cmaxb_fe_lanes_fork(cmaxb_fe* fe);
int cmaxb_fe_lanes_join(cmaxb_fe* fe);

/* Fused result exchange for hypothesis sharding across GPUs (BASELINE config C3): one process
 * per GPU, every rank holds the packet and evaluates its own hypotheses.  After connect, the evaluation kernel
 * itself stores its (contrast, g0, g1, g2) rows into every peer's exchange buffer (peer-to-peer stores over
 * NVLink / NVSwitch through CUDA IPC mappings), signals a per-rank sequence flag, waits for the peers' flags and
 * hands the rows of ALL ranks to the host -- the all-gather the reference-side driver would otherwise issue as
 * a separate NCCL collective happens inside the one launch.
 *   cmaxb_fe_exchange_init    allocates this rank's exchange buffer; handle64_out receives its 64-byte
 *                             cudaIpcMemHandle_t, which the caller distributes to all ranks (any transport)
 *   cmaxb_fe_exchange_connect handles = world x 64 bytes in rank order; gathered_dev (optional, caller-owned
 *                             device buffer of world*k*4 doubles) also receives the gathered rows
 *   cmaxb_fe_eval_fetch_all   like cmaxb_fe_eval_fetch, but returns rows[world][k][4] of all ranks
 * Contract: after connect every rank must issue the same sequence of cmaxb_fe_eval_launch calls (same k,
 * k <= 32); a peer that does not show up within 10 s makes the fetch fail with CMAXB_ERR_CUDA instead of hanging. */
int cmaxb_fe_exchange_init(cmaxb_fe* fe, int world, int rank, void* handle64_out);
int cmaxb_fe_exchange_connect(cmaxb_fe* fe, const void* handles, double* gathered_dev);
int cmaxb_fe_exchange_close(cmaxb_fe* fe);
int cmaxb_fe_eval_fetch_all(cmaxb_fe* fe, double* rows);

/* Optional: a caller-owned DEVICE buffer of max_hypotheses*4 doubles that also receives (contrast, g0, g1, g2)
 * of every evaluation, so that a collective (NCCL all-gather / all-reduce of the per-hypothesis rows,
 * SURVEY section 8e) can consume the results on the handle's stream without a host round trip.  NULL = off. */
int cmaxb_fe_set_result_mirror(cmaxb_fe* fe, double* device_ptr);

/* IWE for display / parity (publishEventImage, ang_vel_estimator.cpp:203-233).  blurred=0: raw
 * accumulator; 1: after the Gaussian blur.  out: height*width floats. */
int cmaxb_fe_get_iwe(cmaxb_fe* fe, const double omega[3], int blurred, float* out);
/* Derivative images (height*width*3 floats, interleaved like CV_32FC3).  Parity / debugging. */
int cmaxb_fe_get_deriv(cmaxb_fe* fe, const double omega[3], int blurred, float* out);
/* Per-event integer cell yy*width+xx (-1 = rejected by the bounds test,
 * local_image_warped_events.cpp:139-142).  n ints.  Parity / debugging. */
int cmaxb_fe_get_cells(cmaxb_fe* fe, const double omega[3], int32_t* out);

/* ------------------------------------------------------------------ back-end ------------- */
typedef struct cmaxb_be cmaxb_be;

typedef struct cmaxb_be_cfg {
  int32_t sensor_width, sensor_height;
  const double* lut_xyz;        /* sensor_width*sensor_height*3, copied */
  int32_t pano_width, pano_height; /* map_opt_.pano_width / pano_height */
  double blur_sigma;
  int32_t batch_size;           /* warp_opt_.event_batch_size */
  int32_t event_sample_rate;    /* warp_opt_.event_sample_rate */
  int32_t spline_order;         /* 2 = LinearTrajectory (So3Spline<2>), 4 = CubicTrajectory (So3Spline<4>) */
  int32_t contrast_measure;
  int32_t grad_mode;            /* CMAXB_GRAD_* (DENSE keeps 3*K_opt band images: small problems only) */
  int32_t device;
  void* stream;
} cmaxb_be_cfg;

typedef struct cmaxb_be_window {
  const cmaxb_event* events;    /* event_subset_ of the window (host) */
  size_t n_events;
  const double* knots_xyzw;     /* n_knots unit quaternions (x,y,z,w): control poses
                                   idx_cp_traj_beg_ .. end (trajectory.cpp:249-254) */
  int32_t n_knots;
  int64_t t0_ns;                /* int64_t(1e9 * t_traj_temp_beg)   trajectory.cpp:60-61,255 */
  int64_t dt_ns;                /* int64_t(1e9 * dt_knots)          trajectory.cpp:60 */
  int32_t n_fixed;              /* num_cps_fixed_ = idx_cp_opt_beg_ - idx_cp_traj_beg_ */
  uint32_t tnext_sec, tnext_nsec; /* t_next_win_beg_ (pose_graph_optimizer.cpp:290) */
  const float* IGp;             /* pano floats or NULL (= zeros) */
  double alpha;                 /* alpha_; NaN => computed on the first eval from IGp and that
                                   eval's IL exactly as updateAlpha (event_pano_warper.cpp:134-165,
                                   201-210) and then frozen for the window */
} cmaxb_be_window;

int cmaxb_be_create(const cmaxb_be_cfg* cfg, cmaxb_be** out);
void cmaxb_be_destroy(cmaxb_be* be);
int cmaxb_be_set_window(cmaxb_be* be, const cmaxb_be_window* w);
/* x: 3*(n_knots-n_fixed) incremental rotation vectors (the gsl_vector of global_contrast_fdf);
 * grad: same length or NULL (value only). */
int cmaxb_be_eval(cmaxb_be* be, const double* x, int n, double* contrast, double* grad);
int cmaxb_be_get_alpha(cmaxb_be* be, double* alpha);
/* parameter vector of the most recent cmaxb_be_eval since set_window.  The reference's updateIG consumes the
 * IL_old_ image left behind by the LAST cost evaluation of the solve (event_pano_warper.cpp:109-126), which is
 * not necessarily the optimum the solver returns; this is how a caller reproduces that. */
int cmaxb_be_last_eval_x(cmaxb_be* be, double* x, int n);

/* Event-sharded evaluation: ONE window split by time across GPUs (BASELINE config C5).  Every
 * rank holds a batch-aligned time slab of the window's events and ALL knots.  Variance is non-linear in the
 * image, so the partial images must be summed before mean / variance:
 *   cmaxb_be_eval_begin : spline poses + scatter of this rank's events; IL assembled into one float plane
 *   (caller)            : all-reduce SUM of cmaxb_be_il_plane() across ranks (NCCL) on the handle's stream
 *   cmaxb_be_eval_end   : blur + contrast on the summed plane (identical on all ranks), adjoint image, gather
 *                         over this rank's events -> contrast and this rank's PARTIAL gradient; the caller
 *                         sums the partial gradients (all-reduce of 3*K_opt doubles).
 * With one rank (no exchange) begin + end == cmaxb_be_eval.  Needs CMAXB_GRAD_ADJOINT for gradients. */
int cmaxb_be_eval_begin(cmaxb_be* be, const double* x, int n, int want_grad);
int cmaxb_be_il_plane(cmaxb_be* be, float** device_ptr, size_t* count);
int cmaxb_be_eval_end(cmaxb_be* be, double* contrast, double* grad_partial);
/* cmaxb_be_eval_end in two halves: _launch queues blur, contrast, adjoint image, gather and per-knot reduction and
 * leaves this rank's partial gradient in the device buffer cmaxb_be_grad_device() reports (3*K_opt doubles); the
 * caller all-reduces that buffer on the handle's stream (NCCL); _fetch then copies contrast and the SUMMED gradient
 * to the host.  A sharded evaluation so has no host round trip between its two collectives. */
int cmaxb_be_eval_end_launch(cmaxb_be* be, int want_grad);
int cmaxb_be_grad_device(cmaxb_be* be, double** device_ptr, size_t* count);
int cmaxb_be_eval_end_fetch(cmaxb_be* be, double* contrast, double* grad);
/* The same time-sharded evaluation with the IMAGE PHASES sharded by row band (each rank blurs / adjoint-blurs 1/world of the
 * panorama instead of all of it; the bytes over NVLink are those of the whole-plane exchange: reduce-scatter + all-gather):
 *   cmaxb_be_shard_begin   poses + scatter of this rank's events; IL packed as `world` bands of chunk_floats floats each
 *                          (band + halo of 2 r + 1 rows, zero outside the panorama) in *send_dev
 *   (caller)               reduce_scatter: *recv_dev <- SUM over ranks of chunk `rank` of *send_dev
 *   cmaxb_be_shard_image   blur of the band, S1 / S2 over the band's own rows -> *sums_dev (2 doubles)
 *   (caller)               all_reduce SUM of *sums_dev
 *   cmaxb_be_shard_adjoint contrast and mean; gradient evaluations: adjoint blur of the band, *g_own_dev = the band's own rows
 *                          of G (own_floats floats), *g_full_dev = the buffer to all-gather them into (world * own_floats)
 *   (caller)               all_gather: *g_full_dev <- *g_own_dev of every rank
 *   cmaxb_be_shard_gather  gather over this rank's events -> partial gradient in cmaxb_be_grad_device()
 *   (caller)               all_reduce SUM of the gradient; cmaxb_be_eval_end_fetch returns contrast (+ gradient)
 * alpha must be fixed (run the first evaluation of a NaN-alpha window through cmaxb_be_eval_begin / _end); bands must be at
 * least 2 r + 1 rows high. */
int cmaxb_be_shard_begin(cmaxb_be* be, const double* x, int n, int want_grad, int world, int rank, float** send_dev,
                         float** recv_dev, size_t* chunk_floats);
int cmaxb_be_shard_image(cmaxb_be* be, double** sums_dev);
int cmaxb_be_shard_adjoint(cmaxb_be* be, float** g_own_dev, float** g_full_dev, size_t* own_floats);
int cmaxb_be_shard_gather(cmaxb_be* be);
/* The same time-sharded evaluation with the exchange done BY THE KERNELS over peer memory (NVLink / NVSwitch, CUDA IPC),
 * and only for the panorama tiles a rank's events touched (csrc/be_xchg.cuh) -- no NCCL call on the data path:
 *   cmaxb_be_exchange_init     allocates this rank's exchange block, returns its 64-byte IPC handle
 *   cmaxb_be_exchange_connect  handles = world x 64 bytes in rank order (gathered by the caller through any channel)
 *   cmaxb_be_xeval             collective: every rank calls it with the same x in the same order; contrast (and gradient,
 *                              when grad != NULL) of the WHOLE window come back on every rank, bitwise identical
 *   cmaxb_be_exchange_close    unmaps the peers' blocks
 * Requirements: at most 8 ranks on one node with peer access; alpha fixed (first evaluation of a NaN-alpha window through
 * cmaxb_be_eval_begin / _end); row bands at least 2 r + 1 rows; at most 1024 free control poses.  A rank that does not
 * arrive within 5 s makes cmaxb_be_xeval fail with CMAXB_ERR_CUDA instead of hanging the GPU. */
int cmaxb_be_exchange_init(cmaxb_be* be, int world, int rank, void* handle64_out);
int cmaxb_be_exchange_connect(cmaxb_be* be, const void* handles);
int cmaxb_be_exchange_close(cmaxb_be* be);
int cmaxb_be_xeval(cmaxb_be* be, const double* x, int n, double* contrast, double* grad);
/* tiles2[0] = 32x32 panorama tiles this rank's events touched in the last cmaxb_be_xeval, tiles2[1] = tiles of the panorama */
int cmaxb_be_exchange_stats(cmaxb_be* be, int64_t* tiles2);
/* IL_old_ / IL_new_ at x (needed by updateIG, event_pano_warper.cpp:109-126); either may be NULL */
int cmaxb_be_get_il(cmaxb_be* be, const double* x, int n, float* il_old, float* il_new);
/* final image I = blur(IL + alpha*IGp) at x */
int cmaxb_be_get_iwe(cmaxb_be* be, const double* x, int n, int blurred, float* out);
/* dense derivative bands (3*(n_knots-n_fixed) planes of pano floats, planar). Parity / debugging;
 * allocates the band images on first use. */
int cmaxb_be_get_bands(cmaxb_be* be, const double* x, int n, int blurred, float* out);
int cmaxb_be_get_cells(cmaxb_be* be, const double* x, int n, int32_t* out);
/* per-batch pose table at x: R (9 doubles, row-major), Jk (3 x 3*order floats, row-major),
 * idx_cp_beg.  Parity / debugging of the device So3Spline. Arrays sized n_batches. */
int cmaxb_be_get_poses(cmaxb_be* be, const double* x, int n, int64_t* n_batches, double* R9,
                       float* Jk, int32_t* idx_cp_beg, int64_t capacity);

/* ------------------------------------------------------------------ global map upkeep ---- */
/* IG_ and IG_update_times_map_ (event_pano_warper.h) kept in HBM between windows, so a
 * window neither uploads a panorama (IGp) nor downloads IL_old:
 *   cmaxb_be_map_reset      resetIG + zero visit counts                      event_pano_warper.h:56, .cpp:21-28
 *   cmaxb_be_map_set / _get host <-> device copy of IG (pano floats) and the visit counts (pano bytes); NULL = skip
 *   cmaxb_be_map_use_as_igp after set_window: IGp <- IG on the device (updateIGp, .cpp:128-132); alpha as in
 *                           cmaxb_be_window.alpha (NaN => updateAlpha on the first evaluation)
 *   cmaxb_be_map_update     IG += IL_old(x) where visits <= max_update_times (updateIG, .cpp:109-126)
 *   cmaxb_be_map_mark_fov   setUpdateTimesIG(rot, radius) for m poses (x,y,z,w): sensor FOV raster dilated by
 *                           radius, added (saturating) to the visit counts   (.cpp:81-107, pose_graph_optimizer.cpp:325-337) */
int cmaxb_be_map_reset(cmaxb_be* be);
int cmaxb_be_map_set(cmaxb_be* be, const float* IG, const uint8_t* visit_counts);
int cmaxb_be_map_get(cmaxb_be* be, float* IG, uint8_t* visit_counts);
int cmaxb_be_map_use_as_igp(cmaxb_be* be, double alpha);
int cmaxb_be_map_update(cmaxb_be* be, const double* x, int n, int max_update_times);
int cmaxb_be_map_mark_fov(cmaxb_be* be, const double* rot_xyzw, int m, int radius);

/* ------------------------------------------------------------------ optimiser loop ------- */
/* the reference's solve loops (GSL Fletcher-Reeves conjugate gradient with the
 * reference's constants and stopping rules; local_optim_contrast_gsl.cpp:74-233,
 * global_optim_contrast_gsl.cpp:15-145) on top of the evaluation entry points, so a whole packet / window
 * solve needs neither GSL nor per-evaluation glue.  GSL's algorithm is restated (conjugate_fr.c,
 * directional_minimize.c); cost = -contrast.  params == NULL selects the reference's constants. */
typedef struct cmaxb_opt_params {
  double initial_step;   /* 0.1 */
  double line_tol;       /* FE 0.05, BE 0.1 */
  int32_t max_iterations;/* 50 */
  double epsabs_grad;    /* FE 1e-3, BE 1e-4 */
  double tolfun;         /* 1e-4 */
  int32_t fused_trials;  /* 0: the reference's call pattern (value-only trials, then df at the accepted point);
                            1: every line-search trial is evaluated WITH its gradient and memoised, so the df request GSL
                            issues at the accepted trial point costs no second evaluation -- same requests, same iterates
                            (up to the ~1e-7 by which the value of a value+gradient evaluation differs from a value-only
                            one), fewer launches */
} cmaxb_opt_params;
typedef struct cmaxb_opt_result {
  double cost_initial, cost_final;   /* -contrast */
  int32_t iterations, f_evals, g_evals;
  int32_t stop_reason;               /* 0 iteration limit, 1 cost stagnation, 2 gradient norm, 3 no progress */
  int32_t cost_launches;             /* cost evaluations actually run (f_evals / g_evals count the minimiser's requests) */
} cmaxb_opt_result;
int cmaxb_fe_optimize(cmaxb_fe* fe, const double omega0[3], const cmaxb_opt_params* params, double omega_out[3],
                      cmaxb_opt_result* result);
/* The same Fletcher-Reeves loop over a caller-supplied cost to be MINIMISED (callbacks return 0 on success); `params` is
 * required (reference constants: FE {0.1, 0.05, 50, 1e-3, 1e-4}, BE {0.1, 0.1, 50, 1e-4, 1e-4}).  Host code only. */
typedef int (*cmaxb_cost_f)(const double* x, int n, void* user, double* f);
typedef int (*cmaxb_cost_fdf)(const double* x, int n, void* user, double* f, double* grad);
int cmaxb_optimize_callback(int n, const double* x0, cmaxb_cost_f f, cmaxb_cost_fdf fdf, void* user,
                            const cmaxb_opt_params* params, double* x_out, cmaxb_opt_result* result);
/* x0 == NULL: start from zero increments (global_optim_contrast_gsl.cpp:36-37) */
int cmaxb_be_optimize(cmaxb_be* be, const double* x0, int n, const cmaxb_opt_params* params, double* x_out,
                      cmaxb_opt_result* result);

/* ------------------------------------------------------------------ trajectory initialisation ---- */
/* the step between the front-end's angular velocities and the back-end window solve.
 * Host C++ inside the library (a few dozen poses, one small dense least-squares system) so that a whole window
 * -- initialise, solve, update the map -- runs behind this ABI.  Rotations are unit quaternions (x, y, z, w). */
typedef struct cmaxb_stamp { uint32_t sec, nsec; } cmaxb_stamp;   /* ros::Time */

/* PoseGraphOptimizer::integrateAngVel (pose_graph_optimizer.cpp:191-222): trapezoid integration of the m angular
 * velocities (strictly increasing stamps) from pose_latest_; entries not newer than ang_vel_prev_ are skipped
 * unless first_time_window.  ang_vel_prev_t / ang_vel_prev are updated in place (state carried across windows).
 * pose_*_out: capacity m; *n_out = poses produced. */
int cmaxb_traj_integrate_ang_vel(cmaxb_stamp pose_latest_t, const double pose_latest_xyzw[4],
                                 cmaxb_stamp* ang_vel_prev_t, double ang_vel_prev[3], int first_time_window,
                                 const cmaxb_stamp* t, const double* ang_vel, int m,
                                 cmaxb_stamp* pose_t_out, double* pose_xyzw_out, int* n_out);
/* number of control poses generateCtrlPoses fits for [t_beg, t_end]: round(span / dt_knots) + 1 (order 2) or + 3
 * (order 4)   (trajectory.cpp:203-212, 479-489).  Negative = error. */
int cmaxb_traj_num_ctrl_poses(int spline_order, cmaxb_stamp t_beg, cmaxb_stamp t_end, double dt_knots);
/* Linear/CubicTrajectory::fitCtrlPoses (trajectory.cpp:112-186, 357-463): lift the poses to the tangent space at
 * the first pose, solve the B-spline collocation system N P = D in the least-squares sense (full-pivoting
 * Householder QR, as Eigen::FullPivHouseholderQR), retract.  ctrl_xyzw_out: num_cps quaternions. */
int cmaxb_traj_fit_ctrl_poses(int spline_order, double dt_knots, double t_beg_sec, int num_cps,
                              const cmaxb_stamp* pose_t, const double* pose_xyzw, int n_poses, double* ctrl_xyzw_out);
/* Trajectory::evaluate(t), value only (pose_latest_ update pose_graph_optimizer.cpp:316-317, setUpdateTimesIG
 * :325-337).  t0_ns / dt_ns as in cmaxb_be_window. */
int cmaxb_traj_evaluate(int spline_order, const double* knots_xyzw, int n_knots, int64_t t0_ns, int64_t dt_ns,
                        cmaxb_stamp t, double out_xyzw[4]);
/* Trajectory::incrementalUpdate: knots[i] <- exp(x[i - idx_beg]) * knots[i], i >= idx_beg (trajectory.cpp:221-238,
 * 491-499) -- applies the optimiser's result to the trajectory. */
int cmaxb_traj_incremental_update(double* knots_xyzw, int n_knots, int idx_beg, const double* x);

/* ------------------------------------------------------------------ bearing-vector LUT ---------- */
/* CMaxSLAM::precomputeBearingVectors (src/cmax_slam.cpp:106-120): for every sensor pixel,
 * cam.projectPixelTo3dRay(cam.rectifyPoint(pixel)) with image_geometry::PinholeCameraModel semantics (float32 round
 * trip through cv::undistortPoints with 5 iterations when any distortion coefficient is non-zero; ray from the
 * projection matrix P).  One thread per pixel on the device; lut_xyz receives width*height*3 doubles, the input of
 * cmaxb_fe_cfg.lut_xyz / cmaxb_be_cfg.lut_xyz.  Fields = sensor_msgs::CameraInfo (no binning / ROI). */
typedef struct cmaxb_camera_info {
  int32_t width, height;
  double K[9];      /* row-major 3x3 */
  double D[12];     /* k1 k2 p1 p2 k3 [k4 k5 k6 [s1 s2 s3 s4]] (plumb_bob: 5, rational_polynomial: 8) */
  int32_t n_D;
  double R[9];      /* rectification rotation */
  double P[12];     /* row-major 3x4 projection matrix */
} cmaxb_camera_info;
int cmaxb_precompute_bearing_vectors(const cmaxb_camera_info* info, int device, double* lut_xyz);

/* ------------------------------------------------------------------ event ingestion / staging ---- */
/* the event store both ends share, the front-end's packet cutter and the back-end's
 * window cutter (host C++).  Packets / windows are handed out in PINNED host buffers (when a CUDA device is present),
 * so cmaxb_fe_set_packet_async / cmaxb_be_set_window move them by DMA while the previous packet is evaluated; the
 * per-packet spatial binning and validation happen on the device inside set_packet.  Replaces
 * CMaxSLAM::eventsCallback (src/cmax_slam.cpp:147-161), AngVelEstimator::pushEvent / getEventSubset /
 * deleteOldEvents / slideWindow (src/frontend/ang_vel_estimator.cpp:68-183) and PoseGraphOptimizer::getEventSubset
 * (src/backend/pose_graph_optimizer.cpp:133-166). */
typedef struct cmaxb_stream cmaxb_stream;
typedef struct cmaxb_stream_cfg {
  double dt_ang_vel;              /* params.dt_ang_vel: one packet (one angular velocity) per dt */
  int32_t num_events_per_packet;  /* params.num_events_per_packet */
  int32_t event_sample_rate;      /* front_end_params_.warp_opt.event_sample_rate: stride over each incoming message */
} cmaxb_stream_cfg;
int cmaxb_stream_create(const cmaxb_stream_cfg* cfg, cmaxb_stream** out);
void cmaxb_stream_destroy(cmaxb_stream* s);
/* one dvs_msgs::EventArray; *packets_ready = packets whose trailing half has arrived */
int cmaxb_stream_push(cmaxb_stream* s, const cmaxb_event* msg_events, size_t n, int* packets_ready);
/* oldest complete packet: event_subset_ (valid until the call after next), time_packet_, and whether its time
 * span exceeds 10 dt_ang_vel (the reference then assumes zero angular velocity).  Returns 1 when none is ready. */
int cmaxb_stream_next_packet(cmaxb_stream* s, const cmaxb_event** events, size_t* n, cmaxb_stamp* time_packet,
                             int* span_too_long);
/* events of the back-end window [t_beg, t_end) by the reference's coarse-to-fine search (packet-level look-up table,
 * then 100-event strides back from the end); consumes the look-up entries and deletes the events no end needs any
 * more.  Returns 1 (not an error, nothing consumed) when the store does not cover the window yet; CMAXB_ERR_STATE is
 * reserved for an inconsistent store. */
int cmaxb_stream_window_events(cmaxb_stream* s, cmaxb_stamp t_beg, cmaxb_stamp t_end, const cmaxb_event** events, size_t* n);
int cmaxb_stream_state(cmaxb_stream* s, int64_t* n_stored, int64_t* n_subsets_pending, int64_t* n_ts_map, cmaxb_stamp* time_packet);

/* Device-resident event store.  The reference cuts OVERLAPPING packets out of one host vector (a packet = the half
 * packet before and after every dt_ang_vel tick, ang_vel_estimator.cpp:84-92,137-147), so handing each packet to the
 * device separately moves every event over PCIe several times.  With a device store attached, every pushed event is
 * copied to a ring in device memory ONCE (asynchronously, on `cuda_stream`) and packets are handed out as views of that
 * ring for cmaxb_fe_set_packet_view (same stream => ordered after the copy).
 *   cmaxb_stream_attach_device   before the first push; ring_events = ring capacity in events (0: 8 packets)
 *   cmaxb_stream_push_ex         flags: CMAXB_PUSH_BORROW -- msg_events is page-locked host memory that stays valid and
 *                                unchanged until cmaxb_stream_released() has passed it: the message is referenced in place
 *                                (no host copy) and DMA-ed from where it lies (needs event_sample_rate 1);
 *                                CMAXB_PUSH_SORTED -- the message is sorted by time stamp (as DVS drivers deliver it): the
 *                                packet ticks inside it are found by bisection instead of an event-by-event scan
 *   cmaxb_stream_next_packet_device   like cmaxb_stream_next_packet, but *device_events points into the device ring (valid
 *                                until ring_events - packet more events have been pushed)
 *   cmaxb_stream_released        number of events, counted from the first push, that the store has dropped */
#define CMAXB_PUSH_BORROW 1
#define CMAXB_PUSH_SORTED 2
int cmaxb_stream_attach_device(cmaxb_stream* s, int device, void* cuda_stream, size_t ring_events);
int cmaxb_stream_push_ex(cmaxb_stream* s, const cmaxb_event* msg_events, size_t n, int flags, int* packets_ready);
int cmaxb_stream_next_packet_device(cmaxb_stream* s, const cmaxb_event** device_events, size_t* n, cmaxb_stamp* time_packet,
                                    int* span_too_long);
int cmaxb_stream_released(cmaxb_stream* s, int64_t* n_released);
/* makes consumer_stream wait for the device copies of everything pushed so far: needed when the store copies on its own
 * stream (so that uploads overlap the evaluations) and the packets are prepared on another one */
int cmaxb_stream_wait_copied(cmaxb_stream* s, void* consumer_stream);

/* ------------------------------------------------------------------ back-end window pipeline ---- */
/* Everything PoseGraphOptimizer does for one sliding time window (pose_graph_optimizer.cpp:72-354), as host C++
 * over a cmaxb_be handle: the caller pushes the front-end's angular velocities and hands over the events of the
 * current window; the library integrates them, fits and appends the new control poses, freezes / slides the
 * control-pose indices, solves the window (Fletcher-Reeves over cmaxb_be_eval, x0 = 0), applies the increments,
 * updates the panoramic map and its visit counts ON THE DEVICE and prepares the next window.  Replaces, in
 * src/backend/pose_graph_optimizer.cpp: pushAngVel, getAngVelSubset, integrateAngVel, processTimeWindow,
 * setUpdateTimesIG, slideWindow (+ setupProblemAndOptimize_gsl).  The event cut for [t_win_beg, t_win_end) is
 * the caller's (cmaxb_stream_*, or the reference's own getEventSubset). */
typedef struct cmaxb_pgo cmaxb_pgo;
typedef struct cmaxb_pgo_cfg {
  int32_t spline_order;            /* 2 (traj_opt.spline_degree 1) or 4 (degree 3); must match the cmaxb_be handle */
  double dt_knots;                 /* traj_opt.dt_knots */
  double time_window_size;         /* sliding_window_opt.time_window_size (s) */
  double sliding_window_stride;    /* sliding_window_opt.sliding_window_stride (s) */
  double y_angle_deg;              /* map_opt.Y_angle: initial pose = rotation about Y (pose_graph_optimizer.cpp:99-103) */
  int32_t max_update_times;        /* map_opt.max_update_times (updateIG gate) */
  double min_num_ev_per_win;       /* min_num_ev_per_win_ (pose_graph_optimizer.cpp:64-67): fewer events => no solve */
  int32_t use_opt_params;          /* 0: the reference's constants (step 0.1, tol 0.1, 50 iterations, 1e-4, 1e-4) */
  cmaxb_opt_params opt_params;
} cmaxb_pgo_cfg;
typedef struct cmaxb_pgo_report {
  int32_t window;                  /* count_window_ of the processed window */
  cmaxb_stamp t_win_beg, t_win_end;
  int32_t n_ang_vel, n_frontend_poses;
  int32_t n_ctrl_poses, idx_cp_traj_beg, idx_cp_opt_beg, num_cp_opt;
  int32_t optimized;               /* 0: too few events, camera assumed still */
  cmaxb_opt_result opt;
  double alpha;
  int32_t n_fov_marks;
  cmaxb_stamp pose_latest_t; double pose_latest_xyzw[4];
} cmaxb_pgo_report;
int cmaxb_pgo_create(const cmaxb_pgo_cfg* cfg, cmaxb_be* be /* borrowed; NULL = trajectory bookkeeping only, no solve */, cmaxb_pgo** out);
void cmaxb_pgo_destroy(cmaxb_pgo* pgo);
int cmaxb_pgo_push_ang_vel(cmaxb_pgo* pgo, cmaxb_stamp ts, const double ang_vel[3]);
/* current window cursors; *ang_vel_ready = the latest angular velocity lies beyond t_win_end (isReadyFrontendPoses) */
int cmaxb_pgo_window(cmaxb_pgo* pgo, cmaxb_stamp* t_win_beg, cmaxb_stamp* t_win_end, int* ang_vel_ready);
/* getAngVelSubset + processTimeWindow + slideWindow on the given event subset (host memory).  On an error (e.g. fewer
 * front-end poses than control poses to fit -- the reference aborts on that CHECK_GE, trajectory.cpp:116) the window
 * is NOT slid but the angular velocities it consumed are gone: treat the handle as dead, as the reference's abort does. */
int cmaxb_pgo_process_window(cmaxb_pgo* pgo, const cmaxb_event* events, size_t n_events, cmaxb_pgo_report* report);
/* all control poses so far (x,y,z,w) and the spline's time origin / knot spacing; xyzw may be NULL to query *n */
int cmaxb_pgo_get_ctrl_poses(cmaxb_pgo* pgo, double* xyzw, int capacity, int* n, int64_t* t0_ns, int64_t* dt_ns);

/* ------------------------------------------------------------------ diagnostics ---------- */
const char* cmaxb_last_error(void);
int cmaxb_version(void);
int cmaxb_device_count(void);
/* number of kernels THIS LIBRARY launched since load (all handles) */
uint64_t cmaxb_launch_count(void);

/* Per-kernel device timing with CUDA events on the handle's stream.  Enable, run evaluations,
 * then read accumulated milliseconds / launch counts per kernel kind. */
enum {
  CMAXB_K_ZERO = 0, CMAXB_K_FE_SCATTER, CMAXB_K_FE_GATHER, CMAXB_K_BLUR_REDUCE, CMAXB_K_ADJOINT_BLUR,
  CMAXB_K_BE_POSES, CMAXB_K_BE_SCATTER, CMAXB_K_BE_GATHER, CMAXB_K_BE_GRAD_REDUCE, CMAXB_K_MISC,
  CMAXB_K_FE_EVAL_FUSED, CMAXB_K_BE_EVAL_FUSED,
  CMAXB_K_BE_X_PUSH, CMAXB_K_BE_X_SUMS, CMAXB_K_BE_X_PULL, CMAXB_K_BE_X_GRAD,   /* peer exchange of cmaxb_be_xeval, waits for the peers included */
  CMAXB_K_COUNT
};
int cmaxb_fe_profile(cmaxb_fe* fe, int enable);
int cmaxb_fe_kernel_times(cmaxb_fe* fe, double* ms /*CMAXB_K_COUNT*/, uint64_t* launches /*CMAXB_K_COUNT*/);
/* profiling aid: per-CTA %globaltimer stamps of the last profiled whole-grid fused launch, us since kernel entry:
 * out[cta][4] = scatter end, image phase end, gather start, gather end; *n_ctas = CTAs traced (<= max_ctas, <= 1024) */
int cmaxb_fe_cta_times(cmaxb_fe* fe, double* out, int max_ctas, int* n_ctas);
/* Fused evaluation kernel, profiling enabled: microseconds from kernel entry (CTA 0) to the phase boundaries of the LAST
 * synchronous evaluation: [1] scatter end, [2] grid barrier, [3] image phase end, [4] grid barrier, [5] gather end,
 * [6] last CTA starts the final sums, [7] result rows stored (all CTA 0 except [6], [7]); [8] / [9]: the SLOWEST CTA's
 * scatter end / image phase end; -1 = not reached. */
int cmaxb_fe_phase_times(cmaxb_fe* fe, double* us10);
/* fused launches: [0] CTAs of a whole-GPU launch, [1] CTAs of a throughput-lane launch, [2] throughput lanes,
 * [3] TMA tile staging on, [4] / [5] image tile height of the two launch shapes, [6] reserved (0) */
int cmaxb_fe_launch_info(cmaxb_fe* fe, int32_t* info7);
int cmaxb_be_profile(cmaxb_be* be, int enable);
int cmaxb_be_kernel_times(cmaxb_be* be, double* ms, uint64_t* launches);
const char* cmaxb_kernel_name(int kind);

#ifdef __cplusplus
}
#endif
#endif /* CMAX_B200_H_ */
