// cmax_b200_pipeline.hpp -- reference-side binding of the entry points AROUND the cost function (SURVEY section 8f
// ranks 2-4): thin C++17 RAII wrappers a maintainer drops into the reference's node in place of the event store,
// the packet / window cutters and PoseGraphOptimizer's per-window work.  Header-only; links against
// libcmax_b200.so directly (or resolve the same symbols with dlsym, as cmax_b200_gsl.hpp does).
// INTEGRATION.md section 3c lists the edit sites.
//
// Replaces, in the reference:
//   EventStore   : CMaxSLAM::eventsCallback (src/cmax_slam.cpp:147-161); AngVelEstimator::pushEvent /
//                  getEventSubset / deleteOldEvents / slideWindow (src/frontend/ang_vel_estimator.cpp:68-183);
//                  PoseGraphOptimizer::getEventSubset (src/backend/pose_graph_optimizer.cpp:133-166)
//   WindowSolver : PoseGraphOptimizer::pushAngVel / isReadyFrontendPoses / getAngVelSubset / processTimeWindow /
//                  setUpdateTimesIG / slideWindow (src/backend/pose_graph_optimizer.cpp:72-354)
//   bearing_vectors : CMaxSLAM::precomputeBearingVectors (src/cmax_slam.cpp:106-120)
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "cmax_b200.h"

namespace cmaxb_pipeline {

struct Error : std::runtime_error {
  int code;
  Error(int c, const char* where) : std::runtime_error(std::string(where) + ": " + cmaxb_last_error()), code(c) {}
};
inline void check(int rc, const char* where) { if (rc != CMAXB_OK) throw Error(rc, where); }

// ros::Time <-> cmaxb_stamp: same two uint32 fields
template <class RosTime> inline cmaxb_stamp stamp(const RosTime& t) { return cmaxb_stamp{t.sec, t.nsec}; }

struct Span { const cmaxb_event* data = nullptr; size_t size = 0; };

struct Packet {
  Span events;               // event_subset_ (pinned host memory inside the library; valid until the call after next)
  cmaxb_stamp time_packet{}; // time_packet_
  bool span_too_long = false;// timespan > 10 dt_ang_vel: the reference sets ang_vel_ = 0 instead of solving
};

class EventStore {
 public:
  EventStore(double dt_ang_vel, int num_events_per_packet, int event_sample_rate = 1) {
    cmaxb_stream_cfg c{dt_ang_vel, num_events_per_packet, event_sample_rate};
    check(cmaxb_stream_create(&c, &s_), "cmaxb_stream_create");
  }
  ~EventStore() { cmaxb_stream_destroy(s_); }
  EventStore(const EventStore&) = delete;
  EventStore& operator=(const EventStore&) = delete;
  // eventsCallback: `events` = msg->events.data() (dvs_msgs::Event is layout-compatible with cmaxb_event)
  int push(const void* events, size_t n) {
    int ready = 0;
    check(cmaxb_stream_push(s_, static_cast<const cmaxb_event*>(events), n, &ready), "cmaxb_stream_push");
    return ready;
  }
  bool next_packet(Packet* out) {
    int flag = 0;
    const int rc = cmaxb_stream_next_packet(s_, &out->events.data, &out->events.size, &out->time_packet, &flag);
    if (rc == 1) return false;
    check(rc, "cmaxb_stream_next_packet");
    out->span_too_long = flag != 0;
    return true;
  }
  // false when the store does not reach the end of the window yet
  bool window_events(cmaxb_stamp t_beg, cmaxb_stamp t_end, Span* out) {
    const int rc = cmaxb_stream_window_events(s_, t_beg, t_end, &out->data, &out->size);
    if (rc == 1) return false;                  // not an error: nothing consumed (CMAXB_ERR_STATE = inconsistent store, thrown below)
    check(rc, "cmaxb_stream_window_events");
    return true;
  }
  // ---- device-resident store: every event crosses PCIe once, packets come out as views of a device ring --------------
  // before the first push; cuda_stream = the stream the copies run on (its own, not the evaluator's: they overlap the kernels)
  void attach_device(int device, void* cuda_stream, size_t ring_events = 0) {
    check(cmaxb_stream_attach_device(s_, device, cuda_stream, ring_events), "cmaxb_stream_attach_device");
  }
  // flags: CMAXB_PUSH_BORROW (page-locked message, referenced in place until released() passes it), CMAXB_PUSH_SORTED
  int push(const void* events, size_t n, int flags) {
    int ready = 0;
    check(cmaxb_stream_push_ex(s_, static_cast<const cmaxb_event*>(events), n, flags, &ready), "cmaxb_stream_push_ex");
    return ready;
  }
  // device pointer + count for cmaxb_fe_set_packet_view; call wait_copied(evaluator stream) before using it
  bool next_packet_device(const cmaxb_event** dev_events, size_t* n, cmaxb_stamp* time_packet, bool* span_too_long = nullptr) {
    int flag = 0;
    const int rc = cmaxb_stream_next_packet_device(s_, dev_events, n, time_packet, &flag);
    if (rc == 1) return false;
    check(rc, "cmaxb_stream_next_packet_device");
    if (span_too_long) *span_too_long = flag != 0;
    return true;
  }
  void wait_copied(void* consumer_stream) { check(cmaxb_stream_wait_copied(s_, consumer_stream), "cmaxb_stream_wait_copied"); }
  cmaxb_stream* handle() const { return s_; }

 private:
  cmaxb_stream* s_ = nullptr;
};

class WindowSolver {
 public:
  // `be` may be null: trajectory bookkeeping only (no window is solved)
  WindowSolver(const cmaxb_pgo_cfg& cfg, cmaxb_be* be) { check(cmaxb_pgo_create(&cfg, be, &p_), "cmaxb_pgo_create"); }
  ~WindowSolver() { cmaxb_pgo_destroy(p_); }
  WindowSolver(const WindowSolver&) = delete;
  WindowSolver& operator=(const WindowSolver&) = delete;
  void push_ang_vel(cmaxb_stamp ts, const double w[3]) { check(cmaxb_pgo_push_ang_vel(p_, ts, w), "cmaxb_pgo_push_ang_vel"); }
  // isReadyFrontendPoses + the window cursors
  bool ready(cmaxb_stamp* t_beg, cmaxb_stamp* t_end) {
    int r = 0;
    check(cmaxb_pgo_window(p_, t_beg, t_end, &r), "cmaxb_pgo_window");
    return r != 0;
  }
  cmaxb_pgo_report process(Span events) {
    cmaxb_pgo_report rep{};
    check(cmaxb_pgo_process_window(p_, events.data, events.size, &rep), "cmaxb_pgo_process_window");
    return rep;
  }
  std::vector<double> ctrl_poses_xyzw(int64_t* t0_ns = nullptr, int64_t* dt_ns = nullptr) {
    int n = 0;
    check(cmaxb_pgo_get_ctrl_poses(p_, nullptr, 0, &n, t0_ns, dt_ns), "cmaxb_pgo_get_ctrl_poses");
    std::vector<double> q(static_cast<size_t>(4) * n);
    if (n) check(cmaxb_pgo_get_ctrl_poses(p_, q.data(), n, &n, t0_ns, dt_ns), "cmaxb_pgo_get_ctrl_poses");
    return q;
  }
  // The reference's PoseGraphOptimizer::Run loop body (src/backend/pose_graph_optimizer.cpp:357-377): process every
  // window that both the angular velocities and the event store cover.  Returns the number of windows processed.
  template <class OnWindow>
  int run_ready_windows(EventStore& store, OnWindow&& on_window) {
    int n = 0;
    cmaxb_stamp tb{}, te{};
    Span ev;
    while (ready(&tb, &te) && store.window_events(tb, te, &ev)) {
      on_window(process(ev));
      ++n;
    }
    return n;
  }

 private:
  cmaxb_pgo* p_ = nullptr;
};

// cam.projectPixelTo3dRay(cam.rectifyPoint(pixel)) for every pixel; K (9), D (n_D <= 12), R (9), P (12) straight from
// sensor_msgs::CameraInfo.  Result: width*height*3 doubles = std::vector<cv::Point3d> layout.
inline std::vector<double> bearing_vectors(int width, int height, const double* K, const double* D, int n_D, const double* R,
                                           const double* P, int device = 0) {
  cmaxb_camera_info info{};
  info.width = width; info.height = height; info.n_D = n_D;
  for (int i = 0; i < 9; ++i) { info.K[i] = K[i]; info.R[i] = R[i]; }
  for (int i = 0; i < n_D && i < 12; ++i) info.D[i] = D[i];
  for (int i = 0; i < 12; ++i) info.P[i] = P[i];
  std::vector<double> lut(static_cast<size_t>(3) * width * height);
  check(cmaxb_precompute_bearing_vectors(&info, device, lut.data()), "cmaxb_precompute_bearing_vectors");
  return lut;
}

}  // namespace cmaxb_pipeline
