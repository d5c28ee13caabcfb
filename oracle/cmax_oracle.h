/*
 * cmax_oracle.h -- C interface of the CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * A line-faithful, single-threaded C++17 restatement of the CMax-SLAM contrast-maximisation
 * inner loop (reference: tub-rip/cmax_slam @ 12342de).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * path (cmax_slam_b200/) never does.
 *
 * Parity pinning status (see oracle/README.md):
 *   - SO(3) spline / Sophus / Jacobians: PINNED against the real basalt-headers code of the
 *     reference, compiled from /root/reference into oracle/_ref (tests/test_oracle_spline.py,
 *     tests/golden/spline_*.npz).
 *   - Gaussian blur / meanStdDev / mean (OpenCV, un-vendored, unpinned by the reference):
 *     pinned by us against cv2 4.13 (tests/golden/blur_*.npz).
 *   - first-party geometry (canonicalProjection, applyIntrinsics, cross2Matrix, EquirectangularCamera::
 *     projectToImage): PINNED bit-exact against the reference's own sources compiled with container stubs
 *     (oracle/ref_geom_shim.cpp -> oracle/_ref/libref_geom.so, tests/test_oracle_geom.py, tests/golden/geom_ref.npz).
 *   - spline wrapper (time origin, f32 Jacobian repacking): PINNED against the reference's own
 *     src/backend/trajectory.cpp compiled with ROS / OpenCV / glog stand-ins (oracle/ref_traj_shim.cpp ->
 *     oracle/_ref/libref_traj.so, tests/test_traj_firstparty.py, tests/golden/traj_firstparty.npz).
 *   - the hot path as a whole (FE / BE image builders, map upkeep, both computeContrast families): PINNED BIT-EXACT
 *     against the reference's own translation units compiled unmodified with stand-in ROS / OpenCV / glog headers
 *     (oracle/ref_{fe,focus,warper}_shim.cpp -> oracle/_ref, tests/test_oracle_firstparty.py,
 *     tests/golden/hotpath_firstparty.npz).  Underneath those translation units the OpenCV image primitives are this
 *     file's restatements (pinned against cv2 4.13) and ros::Time is restated.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dvs_msgs::Event as laid out by the ROS C++ message generator: uint16 x, uint16 y,
 * ros::Time ts {uint32 sec, uint32 nsec}, bool polarity (+3 pad) = 16 bytes. */
typedef struct orc_event {
  uint16_t x, y;
  uint32_t sec, nsec;
  uint8_t polarity;
  uint8_t pad[3];
} orc_event;

typedef struct orc_fe_args {
  const orc_event* events;
  int64_t n_events;
  double t_ref_sec;        /* time_packet_.toSec() */
  const double* lut_xyz;   /* W*H*3 bearing vectors */
  int32_t width, height;
  double fx, fy, cx, cy;   /* camera_matrix_ entries (0,0) (1,1) (0,2) (1,2) */
  int32_t batch_size;
  double blur_sigma;
  int32_t contrast_measure; /* 0 variance, 1 mean square */
} orc_fe_args;

typedef struct orc_fe_out {
  double contrast;
  double grad[3];
  float* iwe;        /* H*W      final image fed to computeContrast (blurred if sigma>0), or NULL */
  float* deriv;      /* H*W*3    interleaved, final, or NULL */
  float* iwe_raw;    /* H*W      before blur, or NULL */
  float* deriv_raw;  /* H*W*3    before blur, or NULL */
  int32_t* cells;    /* n_events: yy*W+xx of each event, -1 if rejected by the bounds test, or NULL */
  int64_t n_inbounds;
} orc_fe_out;

int orc_fe_eval(const orc_fe_args* a, const double omega[3], int want_grad, orc_fe_out* out);
/* k independent hypotheses, parallel over hypotheses with n_threads std::threads (each one single-threaded like the
 * reference).  contrasts[k], grads[3k] (or NULL). */
int orc_fe_eval_batch(const orc_fe_args* a, const double* omegas, int k, int want_grad,
                      double* contrasts, double* grads, int n_threads);

typedef struct orc_be_args {
  const orc_event* events;
  int64_t n_events;
  const double* lut_xyz;
  int32_t sensor_width, sensor_height;
  int32_t pano_width, pano_height;
  const double* knots_xyzw;  /* K*4 unit quaternions (x,y,z,w) of the temp trajectory */
  int32_t n_knots;
  int64_t t0_ns, dt_ns;
  int32_t spline_order;      /* 2 linear, 4 cubic (basalt So3Spline<N>) */
  int32_t n_fixed;           /* num_cps_fixed_ */
  uint32_t tnext_sec, tnext_nsec; /* t_next_win_beg_ */
  const float* IGp;          /* pano or NULL (treated as zeros) */
  double alpha;
  int32_t batch_size, event_sample_rate;
  double blur_sigma;
  int32_t contrast_measure;
} orc_be_args;

typedef struct orc_be_out {
  double contrast;
  double* grad;      /* 3*(K-n_fixed) or NULL */
  float* iwe;        /* final I (blurred) or NULL */
  float* bands;      /* P planes, final (blurred), or NULL */
  float* bands_raw;  /* P planes before blur or NULL */
  float* il_old;     /* or NULL */
  float* il_new;     /* or NULL */
  int32_t* cells;    /* n_events; -1 rejected, -2 skipped (sampling / trailing batch) */
  int64_t n_inbounds;
} orc_be_out;

/* x: 3*(K-n_fixed) incremental rotation vectors (NULL = zeros) */
int orc_be_eval(const orc_be_args* a, const double* x, int want_grad, orc_be_out* out);

/* EventWarper::updateAlpha (event_pano_warper.cpp:134-165) */
double orc_update_alpha(const float* IGp, const float* IL, int64_t n_px);

/* map upkeep: EventWarper::setUpdateTimesIG (event_pano_warper.cpp:81-107), updateIG (:109-126) */
void orc_set_update_times(const double* lut_xyz, int SW, int SH, int PW, int PH, const double rot_xyzw[4], int radius,
                          uint8_t* times);
void orc_update_ig(float* IG, const float* il_old, const uint8_t* times, int max_update_times, int64_t n);

/* building blocks exposed for pinning tests */
int orc_gaussian_kernel(double sigma, float* taps /* >= 64 */);                 /* returns ksize */
void orc_gaussian_blur(const float* src, float* dst, int W, int H, int C, double sigma);
void orc_mean_stddev(const float* img, int64_t n, double* mean, double* stddev);
/* basalt So3Spline<order>::evaluate restated. knots K*4 xyzw. Outputs: q_xyzw[4], R[9] row-major,
 * start_idx, J (order blocks of 3x3 row-major, f64). returns 0 or <0 if time out of range. */
int orc_so3_spline_eval(int order, const double* knots_xyzw, int K, int64_t t0_ns, int64_t dt_ns,
                        int64_t t_ns, double* q_xyzw, double* R, int32_t* start_idx, double* J);
void orc_so3_exp(const double w[3], double q_xyzw[4]);
void orc_so3_log(const double q_xyzw[4], double w[3]);
/* ros::Time helpers: batch mid-time (time_first + (time_last-time_first)*0.5) */
void orc_batch_mid_time(uint32_t s0, uint32_t ns0, uint32_t s1, uint32_t ns1, uint32_t* s, uint32_t* ns);
const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
