// pgo.cu -- the back-end window pipeline on top of the evaluation shim: everything PoseGraphOptimizer does for one
// sliding time window, as host C++ inside the library, so that a caller only feeds angular velocities and the
// events of the window.  No ROS, no GSL, no OpenCV; the panoramic map never leaves HBM.
//
// Mirrors (file:line of the reference)
//   PoseGraphOptimizer::initialize (cursors, cp_stride_)          src/backend/pose_graph_optimizer.cpp:35-69
//   PoseGraphOptimizer::pushAngVel                                :72-110
//   PoseGraphOptimizer::isReadyFrontendPoses                      :112-131
//   PoseGraphOptimizer::getAngVelSubset                           :168-189
//   PoseGraphOptimizer::processTimeWindow                         :244-323
//   PoseGraphOptimizer::setUpdateTimesIG                          :325-337
//   PoseGraphOptimizer::slideWindow                               :339-354
//   setupProblemAndOptimize_gsl (x0 = 0, incrementalUpdate)       src/backend/global_optim_contrast_gsl.cpp:15-145
//   Trajectory::CopyAndIncrementalUpdate (window time origin)     src/backend/trajectory.cpp:240-263
//   ros::Time / ros::Duration arithmetic                          roscpp (un-vendored; restated, see oracle/cmax_oracle.cpp)
#include <algorithm>
#include <cmath>
#include <vector>

#include "capi_common.cuh"
#include "so3_math.cuh"

using namespace cmaxb;

namespace {

struct RDur { int sec, nsec; };   // ros::Duration, normalised: nsec in [0, 1e9)

inline RDur dur_from_sec(double d) {            // ros::Duration(double) -> fromSec
  const double fl = std::floor(d);
  long long s = (long long)fl;
  long long ns = (long long)std::round((d - (double)s) * 1e9);
  s += ns / 1000000000ll;
  ns %= 1000000000ll;
  return RDur{(int)s, (int)ns};
}
inline cmaxb_stamp stamp_add(cmaxb_stamp t, RDur d) {     // ros::Time + ros::Duration
  long long s = (long long)t.sec + d.sec, ns = (long long)t.nsec + d.nsec;
  while (ns >= 1000000000ll) { ns -= 1000000000ll; ++s; }
  while (ns < 0) { ns += 1000000000ll; --s; }
  return cmaxb_stamp{(uint32_t)s, (uint32_t)ns};
}
inline cmaxb_stamp stamp_sub(cmaxb_stamp t, RDur d) { return stamp_add(t, RDur{-d.sec, -d.nsec}); }
inline bool st_lt(cmaxb_stamp a, cmaxb_stamp b) { return a.sec < b.sec || (a.sec == b.sec && a.nsec < b.nsec); }
inline bool st_gt(cmaxb_stamp a, cmaxb_stamp b) { return st_lt(b, a); }
inline double st_sec(cmaxb_stamp t) { return ros_to_sec(t.sec, t.nsec); }
inline long long st_nsec(cmaxb_stamp t) { return (long long)((unsigned long long)t.sec * 1000000000ull + (unsigned long long)t.nsec); }

struct AngVel { cmaxb_stamp t; double w[3]; };

}  // namespace

struct cmaxb_pgo {
  cmaxb_pgo_cfg cfg{};
  cmaxb_be* be = nullptr;
  bool time_window_initialized = false, first_time_window = true;
  int count_window = 0, cp_stride = 1;
  RDur win_size{}, win_stride{};
  cmaxb_stamp t_win_beg{}, t_win_end{}, t_av_beg{}, t_av_end{};
  // trajectory (Linear/CubicTrajectory): t_beg_, dt_knots_, spline_(dt_ns, t_beg_ns)
  double traj_t_beg = 0.0; long long traj_t_beg_ns = 0, traj_dt_ns = 0;
  std::vector<double> knots;                 // xyzw per control pose
  std::vector<AngVel> ang_vel;               // frontend_ang_vel_ (std::map: sorted, unique stamps)
  AngVel ang_vel_prev{};
  cmaxb_stamp pose_latest_t{}; double pose_latest_q[4] = {0, 0, 0, 1};
  int idx_cp_traj_beg = 0, idx_cp_opt_beg = 0, num_cp_opt = 0;
};

extern "C" int cmaxb_pgo_create(const cmaxb_pgo_cfg* cfg, cmaxb_be* be, cmaxb_pgo** out) {
  if (!cfg || !out) return set_error(CMAXB_ERR_INVALID, "null argument");   // be == NULL: trajectory bookkeeping only, no window is ever solved
  *out = nullptr;
  if ((cfg->spline_order != 2 && cfg->spline_order != 4) || !(cfg->dt_knots > 0) || !(cfg->time_window_size > 0) ||
      !(cfg->sliding_window_stride > 0))
    return set_error(CMAXB_ERR_INVALID, "bad back-end pipeline configuration");
  cmaxb_pgo* p = new cmaxb_pgo();
  p->cfg = *cfg;
  p->be = be;
  p->win_size = dur_from_sec(cfg->time_window_size);
  p->win_stride = dur_from_sec(cfg->sliding_window_stride);
  p->cp_stride = (int)std::round(cfg->sliding_window_stride / cfg->dt_knots);                 // (:58)
  *out = p;
  return CMAXB_OK;
}

extern "C" void cmaxb_pgo_destroy(cmaxb_pgo* p) { delete p; }

extern "C" int cmaxb_pgo_push_ang_vel(cmaxb_pgo* p, cmaxb_stamp ts, const double w[3]) {
  if (!p || !w) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!p->time_window_initialized) {
    p->t_win_beg = ts;                                                                          // (:79-82)
    p->t_win_end = stamp_add(ts, p->win_size);
    p->t_av_beg = p->t_win_beg;
    p->t_av_end = p->t_win_end;
    p->traj_t_beg = st_sec(ts);                                                                 // TrajectorySettings (:85-88), trajectory.cpp:25-33
    p->traj_t_beg_ns = st_nsec(ts);
    p->traj_dt_ns = (long long)(1e9 * p->cfg.dt_knots);
    p->knots.clear();
    p->ang_vel_prev.t = ts;                                                                     // (:95)
    for (int i = 0; i < 3; ++i) p->ang_vel_prev.w[i] = w[i];
    p->time_window_initialized = true;
    // first pose: rotation about Y by map_opt.Y_angle degrees (:99-103); Sophus::SO3d(R0) = that quaternion
    const double theta = p->cfg.y_angle_deg * 3.14159265358979323846 / 180.0;
    p->pose_latest_t = ts;
    p->pose_latest_q[0] = 0.0; p->pose_latest_q[1] = std::sin(0.5 * theta); p->pose_latest_q[2] = 0.0; p->pose_latest_q[3] = std::cos(0.5 * theta);
  }
  // std::map::insert: sorted by stamp, an existing key is kept
  auto it = std::lower_bound(p->ang_vel.begin(), p->ang_vel.end(), ts, [](const AngVel& a, cmaxb_stamp t) { return st_lt(a.t, t); });
  if (it != p->ang_vel.end() && it->t.sec == ts.sec && it->t.nsec == ts.nsec) return CMAXB_OK;
  AngVel a; a.t = ts; a.w[0] = w[0]; a.w[1] = w[1]; a.w[2] = w[2];
  p->ang_vel.insert(it, a);
  return CMAXB_OK;
}

extern "C" int cmaxb_pgo_window(cmaxb_pgo* p, cmaxb_stamp* t_win_beg, cmaxb_stamp* t_win_end, int* ang_vel_ready) {
  if (!p) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!p->time_window_initialized) return set_error(CMAXB_ERR_STATE, "no angular velocity pushed yet");
  if (t_win_beg) *t_win_beg = p->t_win_beg;
  if (t_win_end) *t_win_end = p->t_win_end;
  // isReadyFrontendPoses (:112-131), angular-velocity half: the latest stamp is beyond the window
  if (ang_vel_ready) *ang_vel_ready = (!p->ang_vel.empty() && st_gt(p->ang_vel.back().t, p->t_win_end)) ? 1 : 0;
  return CMAXB_OK;
}

extern "C" int cmaxb_pgo_process_window(cmaxb_pgo* p, const cmaxb_event* events, size_t n_events, cmaxb_pgo_report* rep) {
  if (!p || (!events && n_events > 0)) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!p->time_window_initialized) return set_error(CMAXB_ERR_STATE, "no angular velocity pushed yet");
  const int N = p->cfg.spline_order;
  cmaxb_pgo_report r{};
  r.window = p->count_window;
  r.t_win_beg = p->t_win_beg; r.t_win_end = p->t_win_end;

  // ---- getAngVelSubset(t_ang_vel_beg_, t_ang_vel_end_) (:168-189): upper_bound(beg) .. lower_bound(end), erase [begin, end)
  auto ib = std::upper_bound(p->ang_vel.begin(), p->ang_vel.end(), p->t_av_beg, [](cmaxb_stamp t, const AngVel& a) { return st_lt(t, a.t); });
  auto ie = std::lower_bound(p->ang_vel.begin(), p->ang_vel.end(), p->t_av_end, [](const AngVel& a, cmaxb_stamp t) { return st_lt(a.t, t); });
  std::vector<AngVel> subset;
  if (ib < ie) subset.assign(ib, ie);
  const size_t n_erase = (size_t)(ie - p->ang_vel.begin());     // committed below, once nothing can fail any more
  r.n_ang_vel = (int)subset.size();

  // ---- processTimeWindow (:244-323)
  std::vector<cmaxb_stamp> st(subset.size()), pst(subset.size() + 1);
  std::vector<double> ws(3 * subset.size() + 3), pq(4 * subset.size() + 4);
  for (size_t i = 0; i < subset.size(); ++i) { st[i] = subset[i].t; for (int c = 0; c < 3; ++c) ws[3 * i + c] = subset[i].w[c]; }
  // everything that can reject the window runs on copies first: a failed call leaves the optimiser's state untouched
  const int num_new = cmaxb_traj_num_ctrl_poses(N, p->t_av_beg, p->t_av_end, p->cfg.dt_knots);
  if (num_new < 0) return num_new;
  int n_poses = 0;
  AngVel prev = p->ang_vel_prev;
  CMAXB_TRY(cmaxb_traj_integrate_ang_vel(p->pose_latest_t, p->pose_latest_q, &prev.t, prev.w,
                                         p->first_time_window ? 1 : 0, st.data(), ws.data(), (int)subset.size(), pst.data(), pq.data(), &n_poses));
  r.n_frontend_poses = n_poses;
  std::vector<double> ctrl((size_t)4 * num_new);
  CMAXB_TRY(cmaxb_traj_fit_ctrl_poses(N, p->cfg.dt_knots, st_sec(p->t_av_beg), num_new, pst.data(), pq.data(), n_poses, ctrl.data()));
  p->ang_vel_prev = prev;
  p->ang_vel.erase(p->ang_vel.begin(), p->ang_vel.begin() + (long)n_erase);
  int first_new = 0;
  if (p->first_time_window) {
    p->idx_cp_opt_beg = (N == 4) ? 3 : 1;                                                       // (:259-263)
    p->first_time_window = false;
  } else {
    first_new = (N == 4) ? 3 : 1;                                                               // num_cps_erase (:268-273)
    if (first_new > num_new) first_new = num_new;
  }
  p->knots.insert(p->knots.end(), ctrl.begin() + 4 * first_new, ctrl.end());                    // pushbackCtrlPoses (:277)
  const int size = (int)(p->knots.size() / 4);
  p->idx_cp_traj_beg = p->count_window * p->cp_stride;                                          // (:283-285)
  p->idx_cp_opt_beg = std::max(p->idx_cp_traj_beg, p->idx_cp_opt_beg);
  p->num_cp_opt = size - p->idx_cp_opt_beg;
  r.n_ctrl_poses = size; r.idx_cp_traj_beg = p->idx_cp_traj_beg; r.idx_cp_opt_beg = p->idx_cp_opt_beg; r.num_cp_opt = p->num_cp_opt;
  if (p->idx_cp_traj_beg + N > size || p->num_cp_opt < 0)
    return set_error(CMAXB_ERR_STATE, "window holds fewer control poses than the spline order");
  const cmaxb_stamp tnext = stamp_add(p->t_win_beg, p->win_stride);                             // setNextWinBegTime (:290)

  r.optimized = 0;
  if (p->be && (double)n_events > p->cfg.min_num_ev_per_win && p->num_cp_opt > 0) {             // (:297)
    // window trajectory = control poses idx_cp_traj_beg.. with origin t_beg_ + idx*dt_knots_ (CopyAndIncrementalUpdate)
    cmaxb_be_window w{};
    w.events = events; w.n_events = n_events;
    w.knots_xyzw = p->knots.data() + 4 * p->idx_cp_traj_beg;
    w.n_knots = size - p->idx_cp_traj_beg;
    const double t_traj_temp_beg = p->traj_t_beg + p->idx_cp_traj_beg * p->cfg.dt_knots;        // trajectory.cpp:255
    w.t0_ns = (int64_t)(1e9 * t_traj_temp_beg);                                                 // trajectory.cpp:60-63
    w.dt_ns = (int64_t)(1e9 * p->cfg.dt_knots);
    w.n_fixed = p->idx_cp_opt_beg - p->idx_cp_traj_beg;                                         // setNumFixedCtrlPoses (:288)
    w.tnext_sec = tnext.sec; w.tnext_nsec = tnext.nsec;
    w.IGp = nullptr;
    w.alpha = std::nan("");
    CMAXB_TRY(cmaxb_be_set_window(p->be, &w));
    CMAXB_TRY(cmaxb_be_map_use_as_igp(p->be, std::nan("")));                                    // setFirstIter(true): IGp <- IG, alpha on the first evaluation (:293)
    const int np = 3 * p->num_cp_opt;
    std::vector<double> x((size_t)np, 0.0), x_last((size_t)np, 0.0);
    cmaxb_opt_result res{};
    CMAXB_TRY(cmaxb_be_optimize(p->be, nullptr, np, p->cfg.use_opt_params ? &p->cfg.opt_params : nullptr, x.data(), &res));
    r.opt = res;
    CMAXB_TRY(cmaxb_be_get_alpha(p->be, &r.alpha));
    // traj_->incrementalUpdate(optimal_drotv, idx_cp_opt_beg_)                                 global_optim_contrast_gsl.cpp:127-129
    CMAXB_TRY(cmaxb_traj_incremental_update(p->knots.data(), size, p->idx_cp_opt_beg, x.data()));
    // updateIG(): IG += IL_old_, the member image left by the LAST cost evaluation of the solve (not necessarily the
    // optimum: a rejected line-search trial may have come last)                                event_pano_warper.cpp:109-126
    CMAXB_TRY(cmaxb_be_last_eval_x(p->be, x_last.data(), np));
    CMAXB_TRY(cmaxb_be_map_update(p->be, x_last.data(), np, p->cfg.max_update_times));
    // setUpdateTimesIG (:325-337): poses every 0.05 s over [t_win_beg, t_win_beg + stride) on the UPDATED trajectory
    std::vector<double> rots;
    const RDur dchk = dur_from_sec(0.05);
    for (cmaxb_stamp t = p->t_win_beg; st_lt(t, tnext); t = stamp_add(t, dchk)) {
      double q[4];
      CMAXB_TRY(cmaxb_traj_evaluate(N, p->knots.data(), size, p->traj_t_beg_ns, p->traj_dt_ns, t, q));
      rots.insert(rots.end(), q, q + 4);
    }
    CMAXB_TRY(cmaxb_be_map_mark_fov(p->be, rots.data(), (int)(rots.size() / 4), 3));
    r.n_fov_marks = (int)(rots.size() / 4);
    r.optimized = 1;
  }
  // latest pose for the next window (:316-318)
  p->pose_latest_t = stamp_sub(p->t_win_end, dur_from_sec(1e-6));
  CMAXB_TRY(cmaxb_traj_evaluate(N, p->knots.data(), size, p->traj_t_beg_ns, p->traj_dt_ns, p->pose_latest_t, p->pose_latest_q));
  r.pose_latest_t = p->pose_latest_t;
  for (int i = 0; i < 4; ++i) r.pose_latest_xyzw[i] = p->pose_latest_q[i];

  // ---- slideWindow (:339-354)
  p->t_win_beg = stamp_add(p->t_win_beg, p->win_stride);
  p->t_av_beg = p->t_win_end;
  p->t_win_end = stamp_add(p->t_win_end, p->win_stride);
  p->t_av_end = p->t_win_end;
  p->count_window += 1;
  if (rep) *rep = r;
  return CMAXB_OK;
}

extern "C" int cmaxb_pgo_get_ctrl_poses(cmaxb_pgo* p, double* xyzw, int capacity, int* n, int64_t* t0_ns, int64_t* dt_ns) {
  if (!p || !n) return set_error(CMAXB_ERR_INVALID, "null argument");
  const int size = (int)(p->knots.size() / 4);
  *n = size;
  if (t0_ns) *t0_ns = p->traj_t_beg_ns;
  if (dt_ns) *dt_ns = p->traj_dt_ns;
  if (xyzw) {
    if (capacity < size) return set_error(CMAXB_ERR_INVALID, "capacity too small");
    std::copy(p->knots.begin(), p->knots.end(), xyzw);
  }
  return CMAXB_OK;
}
