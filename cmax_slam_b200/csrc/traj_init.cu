// traj_init.cu -- host side of the back-end trajectory initialisation: the step
// between the front-end (angular velocities) and the back-end window solve (control poses of the SO(3) B-spline).
// Plain host C++ (no device code: the problem is a few dozen poses and a dense <= 100 x 30 least-squares system);
// it lives in libcmax_b200.so so that a whole window -- initialise, solve, update the map -- runs behind the C ABI.
//
// Computes what the reference does in
//   PoseGraphOptimizer::integrateAngVel            src/backend/pose_graph_optimizer.cpp:191-222
//   Linear/CubicTrajectory::generateCtrlPoses      src/backend/trajectory.cpp:203-212, 479-489
//   Linear/CubicTrajectory::fitCtrlPoses           src/backend/trajectory.cpp:112-186, 357-463
//     (Eigen::FullPivHouseholderQR::solve -- Eigen 3.3.9 vendored under thirdparty/basalt-headers/thirdparty/eigen;
//      restated here: full-pivoting Householder QR with Eigen's pivot / rank rules)
//   Linear/CubicTrajectory::evaluate (value)       src/backend/trajectory.cpp:86-110, 329-355
//   Linear/CubicTrajectory::incrementalUpdate      src/backend/trajectory.cpp:221-238, 491-499
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "capi_common.cuh"
#include "so3_math.cuh"

using namespace cmaxb;

namespace {

inline Quat q_load(const double* p) { Quat q; q.x = p[0]; q.y = p[1]; q.z = p[2]; q.w = p[3]; return q; }
inline void q_store(const Quat& q, double* p) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }

// (a - b).toSec() with ros::Duration normalisation (sec int32, nsec in [0, 1e9))
inline double stamp_diff_sec(cmaxb_stamp a, cmaxb_stamp b) {
  long long s = (long long)a.sec - (long long)b.sec;
  long long ns = (long long)a.nsec - (long long)b.nsec;
  if (ns < 0) { ns += 1000000000ll; --s; }
  return (double)(int)s + 1e-9 * (double)(int)ns;
}
inline bool stamp_gt(cmaxb_stamp a, cmaxb_stamp b) { return a.sec > b.sec || (a.sec == b.sec && a.nsec > b.nsec); }

// Least squares min |A x - b| for nrhs right-hand sides by Householder QR with FULL pivoting, as
// Eigen::FullPivHouseholderQR (computeInPlace / _solve_impl): pivot = largest |a_ij| of the remaining corner,
// stop when it is negligible against the largest pivot so far (epsilon * min(rows, cols)), rank counted with
// the same threshold against the largest diagonal; unknowns beyond the rank are set to zero.
// A: rows x cols row-major (destroyed), B: rows x nrhs row-major (destroyed), X: cols x nrhs.
void lstsq_fullpiv_qr(std::vector<double>& A, int rows, int cols, std::vector<double>& B, int nrhs, std::vector<double>& X) {
  const int size = std::min(rows, cols);
  std::vector<int> colperm(cols);
  for (int j = 0; j < cols; ++j) colperm[j] = j;
  std::vector<double> diag((size_t)size, 0.0);
  const double precision = std::numeric_limits<double>::epsilon() * (double)size;
  double biggest = 0.0, maxpivot = 0.0;
  int nonzero_pivots = size;
  auto a = [&](int i, int j) -> double& { return A[(size_t)i * cols + j]; };
  for (int k = 0; k < size; ++k) {
    int pr = k, pc = k;
    double big = 0.0;
    for (int j = k; j < cols; ++j)          // column-major visit, first maximum wins (Eigen's maxCoeff visitor)
      for (int i = k; i < rows; ++i) {
        const double v = std::fabs(a(i, j));
        if (v > big) { big = v; pr = i; pc = j; }
      }
    if (k == 0) biggest = big;
    if (big <= biggest * precision) { nonzero_pivots = k; break; }      // isMuchSmallerThan
    if (pr != k) {
      for (int j = 0; j < cols; ++j) std::swap(a(k, j), a(pr, j));
      for (int j = 0; j < nrhs; ++j) std::swap(B[(size_t)k * nrhs + j], B[(size_t)pr * nrhs + j]);
    }
    if (pc != k) {
      for (int i = 0; i < rows; ++i) std::swap(a(i, k), a(i, pc));
      std::swap(colperm[k], colperm[pc]);
    }
    // Householder vector of column k below the diagonal (Eigen makeHouseholderInPlace)
    double tail2 = 0.0;
    for (int i = k + 1; i < rows; ++i) tail2 += a(i, k) * a(i, k);
    const double c0 = a(k, k);
    double beta, tau;
    if (tail2 <= std::numeric_limits<double>::min()) { tau = 0.0; beta = c0; for (int i = k + 1; i < rows; ++i) a(i, k) = 0.0; }
    else {
      beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
      for (int i = k + 1; i < rows; ++i) a(i, k) /= (c0 - beta);
      tau = (beta - c0) / beta;
    }
    a(k, k) = beta;
    diag[(size_t)k] = beta;
    if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
    // apply H = I - tau v v^T (v = [1; a(k+1.., k)]) to the remaining columns and to B
    if (tau != 0.0) {
      for (int j = k + 1; j < cols; ++j) {
        double s = a(k, j);
        for (int i = k + 1; i < rows; ++i) s += a(i, k) * a(i, j);
        s *= tau;
        a(k, j) -= s;
        for (int i = k + 1; i < rows; ++i) a(i, j) -= s * a(i, k);
      }
      for (int j = 0; j < nrhs; ++j) {
        double s = B[(size_t)k * nrhs + j];
        for (int i = k + 1; i < rows; ++i) s += a(i, k) * B[(size_t)i * nrhs + j];
        s *= tau;
        B[(size_t)k * nrhs + j] -= s;
        for (int i = k + 1; i < rows; ++i) B[(size_t)i * nrhs + j] -= s * a(i, k);
      }
    }
  }
  const double thresh = std::fabs(maxpivot) * std::numeric_limits<double>::epsilon() * (double)size;
  int rank = 0;
  for (int i = 0; i < nonzero_pivots; ++i) rank += (std::fabs(diag[(size_t)i]) > thresh) ? 1 : 0;
  X.assign((size_t)cols * nrhs, 0.0);
  // back substitution on the leading rank x rank triangle
  for (int j = 0; j < nrhs; ++j) {
    std::vector<double> y((size_t)rank);
    for (int i = rank - 1; i >= 0; --i) {
      double s = B[(size_t)i * nrhs + j];
      for (int l = i + 1; l < rank; ++l) s -= a(i, l) * y[(size_t)l];
      y[(size_t)i] = s / a(i, i);
    }
    for (int i = 0; i < rank; ++i) X[(size_t)colperm[i] * nrhs + j] = y[(size_t)i];
  }
}

}  // namespace

extern "C" int cmaxb_traj_integrate_ang_vel(cmaxb_stamp pose_latest_t, const double pose_latest_xyzw[4],
                                            cmaxb_stamp* ang_vel_prev_t, double ang_vel_prev[3], int first_time_window,
                                            const cmaxb_stamp* t, const double* ang_vel, int m,
                                            cmaxb_stamp* pose_t_out, double* pose_xyzw_out, int* n_out) {
  if (!pose_latest_xyzw || !ang_vel_prev_t || !ang_vel_prev || (m > 0 && (!t || !ang_vel || !pose_t_out || !pose_xyzw_out)) || !n_out)
    return set_error(CMAXB_ERR_INVALID, "null argument");
  for (int i = 1; i < m; ++i)
    if (!stamp_gt(t[i], t[i - 1])) return set_error(CMAXB_ERR_TIME_ORDER, "angular velocities must have strictly increasing stamps (they are keys of a std::map)");
  cmaxb_stamp cur_t = pose_latest_t;
  Quat cur = q_load(pose_latest_xyzw);
  int n = 0;
  for (int i = 0; i < m; ++i) {
    if (!stamp_gt(t[i], *ang_vel_prev_t) && !first_time_window) continue;              // (:199-203)
    const double dt = stamp_diff_sec(t[i], cur_t);                                      // (:205)
    Vec3 d;
    d.x = dt * ((ang_vel_prev[0] + ang_vel[3 * i]) / 2.0);                              // (:206)
    d.y = dt * ((ang_vel_prev[1] + ang_vel[3 * i + 1]) / 2.0);
    d.z = dt * ((ang_vel_prev[2] + ang_vel[3 * i + 2]) / 2.0);
    cur_t = t[i];
    cur = quat_mul(cur, so3_exp(d));                                                    // post-multiplication (:210)
    pose_t_out[n] = cur_t;
    q_store(cur, pose_xyzw_out + 4 * n);
    ++n;
    *ang_vel_prev_t = t[i];
    ang_vel_prev[0] = ang_vel[3 * i]; ang_vel_prev[1] = ang_vel[3 * i + 1]; ang_vel_prev[2] = ang_vel[3 * i + 2];
  }
  *n_out = n;
  return CMAXB_OK;
}

extern "C" int cmaxb_traj_num_ctrl_poses(int spline_order, cmaxb_stamp t_beg, cmaxb_stamp t_end, double dt_knots) {
  if ((spline_order != 2 && spline_order != 4) || !(dt_knots > 0)) return set_error(CMAXB_ERR_INVALID, "bad spline order / knot spacing");
  // std::round((t_end - t_beg).toSec() / dt_knots_) + 1 (linear) or + 3 (cubic)
  return (int)std::round(stamp_diff_sec(t_end, t_beg) / dt_knots) + (spline_order == 4 ? 3 : 1);
}

extern "C" int cmaxb_traj_fit_ctrl_poses(int spline_order, double dt_knots, double t_beg_sec, int num_cps,
                                         const cmaxb_stamp* pose_t, const double* pose_xyzw, int n_poses, double* ctrl_xyzw_out) {
  if (!pose_t || !pose_xyzw || !ctrl_xyzw_out) return set_error(CMAXB_ERR_INVALID, "null argument");
  if ((spline_order != 2 && spline_order != 4) || !(dt_knots > 0) || num_cps < 1) return set_error(CMAXB_ERR_INVALID, "bad spline order / knot spacing / count");
  if (n_poses < num_cps) return set_error(CMAXB_ERR_INVALID, "fewer poses than control poses (reference: CHECK_GE(poses.size(), num_cps))");
  const int N = spline_order;
  // basis matrices of the uniform linear / cubic B-spline (trajectory.cpp:146-147, 417-420)
  static const double M2[2][2] = {{1.0, 0.0}, {-1.0, 1.0}};
  static const double M4[4][4] = {{1. / 6, 2. / 3, 1. / 6, 0.0}, {-0.5, 0.0, 0.5, 0.0}, {0.5, -1.0, 0.5, 0.0}, {-1. / 6, 0.5, -0.5, 1. / 6}};
  // 1. lift: increments w.r.t. the first pose
  const Quat offset = q_load(pose_xyzw);
  const Quat offset_inv = quat_inv(offset);
  std::vector<double> A((size_t)n_poses * num_cps, 0.0), D((size_t)n_poses * 3, 0.0), P;
  for (int i = 0; i < n_poses; ++i) {
    const Quat drot = quat_mul(offset_inv, q_load(pose_xyzw + 4 * i));
    const double t = ros_to_sec(pose_t[i].sec, pose_t[i].nsec);
    const int t_i = (int)std::floor((t - t_beg_sec) / dt_knots);                        // first control pose affecting p(t)
    const double u = (t - (t_i * dt_knots + t_beg_sec)) / dt_knots;
    if (t_i < 0 || t_i + N > num_cps)
      return set_error(CMAXB_ERR_SPLINE_RANGE, "pose stamp outside the span of the control poses being fitted");
    double U[4];
    for (int k = 0; k < N; ++k) U[k] = std::pow(u, k);
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int k = 0; k < N; ++k) s += U[k] * (N == 2 ? M2[k][j] : M4[k][j]);
      A[(size_t)i * num_cps + t_i + j] = s;
    }
    const Vec3 rv = so3_log(drot);
    D[(size_t)i * 3] = rv.x; D[(size_t)i * 3 + 1] = rv.y; D[(size_t)i * 3 + 2] = rv.z;
  }
  // 2. solve N P = D in the tangent space
  lstsq_fullpiv_qr(A, n_poses, num_cps, D, 3, P);
  // 3. retract
  for (int i = 0; i < num_cps; ++i) {
    Vec3 d; d.x = P[(size_t)i * 3]; d.y = P[(size_t)i * 3 + 1]; d.z = P[(size_t)i * 3 + 2];
    q_store(quat_mul(offset, so3_exp(d)), ctrl_xyzw_out + 4 * i);
  }
  return CMAXB_OK;
}

extern "C" int cmaxb_traj_evaluate(int spline_order, const double* knots_xyzw, int n_knots, int64_t t0_ns, int64_t dt_ns,
                                   cmaxb_stamp t, double out_xyzw[4]) {
  if (!knots_xyzw || !out_xyzw) return set_error(CMAXB_ERR_INVALID, "null argument");
  if ((spline_order != 2 && spline_order != 4) || dt_ns <= 0) return set_error(CMAXB_ERR_INVALID, "bad spline order / knot spacing");
  const long long t_ns = (long long)((unsigned long long)t.sec * 1000000000ull + (unsigned long long)t.nsec);   // toNSec()
  const long long st = t_ns - t0_ns;
  if (st < 0) return set_error(CMAXB_ERR_SPLINE_RANGE, "time before the start of the spline");
  const long long s = st / dt_ns;
  const double u = (double)(st % dt_ns) / (double)dt_ns;
  if (s + spline_order > (long long)n_knots) return set_error(CMAXB_ERR_SPLINE_RANGE, "time beyond the last spline segment");
  std::vector<Quat> k((size_t)spline_order);
  for (int i = 0; i < spline_order; ++i) k[(size_t)i] = q_load(knots_xyzw + 4 * (s + i));
  const Quat r = (spline_order == 2) ? so3_spline_eval<2>(k.data(), 0, u, nullptr) : so3_spline_eval<4>(k.data(), 0, u, nullptr);
  q_store(r, out_xyzw);
  return CMAXB_OK;
}

extern "C" int cmaxb_traj_incremental_update(double* knots_xyzw, int n_knots, int idx_beg, const double* x) {
  if (!knots_xyzw || (!x && idx_beg < n_knots)) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (idx_beg < 0 || idx_beg > n_knots) return set_error(CMAXB_ERR_INVALID, "bad first optimised control pose");
  for (int i = idx_beg; i < n_knots; ++i) {
    Vec3 d; d.x = x[3 * (i - idx_beg)]; d.y = x[3 * (i - idx_beg) + 1]; d.z = x[3 * (i - idx_beg) + 2];
    q_store(quat_mul(so3_exp(d), q_load(knots_xyzw + 4 * i)), knots_xyzw + 4 * i);       // left perturbation (:236, :497)
  }
  return CMAXB_OK;
}
