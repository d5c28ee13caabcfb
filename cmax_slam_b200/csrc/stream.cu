// stream.cu -- event ingestion / staging: the event store both ends share, the
// front-end's packet cutter and the back-end's window cutter, as host C++ inside the library.  Packets and
// windows are handed out in PINNED host buffers, so cmaxb_fe_set_packet_async / cmaxb_be_set_window copy them to
// the device by DMA while the previous packet is still being evaluated.
//
// Mirrors (file:line of the reference)
//   CMaxSLAM::eventsCallback (front-end subsampling stride)          src/cmax_slam.cpp:147-161
//   AngVelEstimator::pushEvent (cursors, subset bookkeeping)         src/frontend/ang_vel_estimator.cpp:68-136
//   AngVelEstimator::getEventSubset / slideWindow                    :138-147, 176-183
//   AngVelEstimator::deleteOldEvents                                 :149-174
//   PoseGraphOptimizer::getEventSubset (coarse-to-fine window cut)   src/backend/pose_graph_optimizer.cpp:133-166
// The reference handles one event at a time and solves a packet synchronously inside pushEvent; here push() only
// does the bookkeeping and next_packet() hands out the completed packets in the same order with the same cursors
// (time_packet_ advances by dt_ang_vel per packet handed out, exactly as slideWindow does).
#include <algorithm>
#include <cmath>
#include <deque>
#include <vector>

#include "capi_common.cuh"

using namespace cmaxb;

namespace {

struct SDur { int sec, nsec; };
inline SDur sdur_from_sec(double d) {
  const double fl = std::floor(d);
  long long s = (long long)fl;
  long long ns = (long long)std::round((d - (double)s) * 1e9);
  s += ns / 1000000000ll;
  ns %= 1000000000ll;
  return SDur{(int)s, (int)ns};
}
inline cmaxb_stamp sadd(cmaxb_stamp t, SDur d) {
  long long s = (long long)t.sec + d.sec, ns = (long long)t.nsec + d.nsec;
  while (ns >= 1000000000ll) { ns -= 1000000000ll; ++s; }
  while (ns < 0) { ns += 1000000000ll; --s; }
  return cmaxb_stamp{(uint32_t)s, (uint32_t)ns};
}
inline bool slt(cmaxb_stamp a, cmaxb_stamp b) { return a.sec < b.sec || (a.sec == b.sec && a.nsec < b.nsec); }
inline cmaxb_stamp ev_ts(const cmaxb_event& e) { return cmaxb_stamp{e.sec, e.nsec}; }

// a host buffer that is page-locked when a CUDA device is present (asynchronous H2D), pageable otherwise
struct StageBuf {
  cmaxb_event* p = nullptr; size_t cap = 0; bool pinned = false;
  int reserve(size_t n) {
    if (n <= cap) return CMAXB_OK;
    release();
    size_t want = std::max<size_t>(n, 1024);
    want += want / 2;
    void* q = nullptr;
    if (cudaHostAlloc(&q, want * sizeof(cmaxb_event), cudaHostAllocDefault) == cudaSuccess) pinned = true;
    else {
      (void)cudaGetLastError();
      q = std::malloc(want * sizeof(cmaxb_event));
      pinned = false;
      if (!q) return set_error(CMAXB_ERR_INVALID, "out of host memory");
    }
    p = (cmaxb_event*)q; cap = want;
    return CMAXB_OK;
  }
  void release() {
    if (!p) return;
    if (pinned) cudaFreeHost(p); else std::free(p);
    p = nullptr; cap = 0;
  }
};

}  // namespace

struct cmaxb_stream {
  cmaxb_stream_cfg cfg{};
  SDur dt_av{};
  std::vector<cmaxb_event> events;                 // events_
  long long num_event_total = 0;                   // num_event_total_
  bool sliding_window_initialized = false;
  cmaxb_stamp time_packet{}, time_get_subset{};
  int num_ev_half_packet = 0;
  std::deque<std::pair<long long, long long>> subsets_info;       // event_subsets_info_
  std::vector<std::pair<cmaxb_stamp, long long>> ts_map;          // ev_subset_ts_map_ (sorted by stamp, unique)
  long long ev_beg_idx = 0, ev_end_idx = 0;
  StageBuf packet[2]; int packet_cur = 0;          // double-buffered: packet i stays valid while packet i+1 is cut
  StageBuf window;
};

extern "C" int cmaxb_stream_create(const cmaxb_stream_cfg* cfg, cmaxb_stream** out) {
  if (!cfg || !out) return set_error(CMAXB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (!(cfg->dt_ang_vel > 0) || cfg->num_events_per_packet < 2 || cfg->event_sample_rate < 1)
    return set_error(CMAXB_ERR_INVALID, "bad stream configuration");
  cmaxb_stream* s = new cmaxb_stream();
  s->cfg = *cfg;
  s->dt_av = sdur_from_sec(cfg->dt_ang_vel);                          // ang_vel_estimator.cpp:60
  s->num_ev_half_packet = cfg->num_events_per_packet / 2;             // :64
  *out = s;
  return CMAXB_OK;
}

extern "C" void cmaxb_stream_destroy(cmaxb_stream* s) {
  if (!s) return;
  s->packet[0].release(); s->packet[1].release(); s->window.release();
  delete s;
}

// pushEvent without the solve (ang_vel_estimator.cpp:68-100)
static void stream_push_one(cmaxb_stream* s, const cmaxb_event& e) {
  const cmaxb_stamp ts = ev_ts(e);
  if (!s->sliding_window_initialized) {
    const SDur half = sdur_from_sec(((double)s->dt_av.sec + 1e-9 * (double)s->dt_av.nsec) * 0.5);   // dt_av_ * 0.5
    s->time_packet = sadd(ts, half);
    s->time_get_subset = s->time_packet;
    s->sliding_window_initialized = true;
  }
  s->events.push_back(e);
  s->num_event_total += 1;
  if (slt(s->time_get_subset, ts)) {                                  // event.ts > time_get_subset_
    const long long beg = std::max(s->num_event_total - (long long)s->num_ev_half_packet, 0ll);
    const long long end = s->num_event_total + (long long)s->num_ev_half_packet;
    s->subsets_info.emplace_back(beg, end);
    // std::map::insert keeps an existing key
    auto it = std::lower_bound(s->ts_map.begin(), s->ts_map.end(), ts,
                               [](const std::pair<cmaxb_stamp, long long>& a, cmaxb_stamp t) { return slt(a.first, t); });
    if (!(it != s->ts_map.end() && it->first.sec == ts.sec && it->first.nsec == ts.nsec)) s->ts_map.insert(it, {ts, s->num_event_total - 1});
    s->time_get_subset = sadd(s->time_get_subset, s->dt_av);
  }
}

extern "C" int cmaxb_stream_push(cmaxb_stream* s, const cmaxb_event* msg_events, size_t n, int* packets_ready) {
  if (!s || (!msg_events && n > 0)) return set_error(CMAXB_ERR_INVALID, "null argument");
  // eventsCallback: for (ev = begin; ev < end; ev += event_sample_rate) pushEvent(*ev)
  for (size_t i = 0; i < n; i += (size_t)s->cfg.event_sample_rate) stream_push_one(s, msg_events[i]);
  if (packets_ready) {
    int k = 0;
    for (const auto& si : s->subsets_info) { if (s->num_event_total > si.second) ++k; else break; }
    *packets_ready = k;
  }
  return CMAXB_OK;
}

extern "C" int cmaxb_stream_next_packet(cmaxb_stream* s, const cmaxb_event** events, size_t* n, cmaxb_stamp* time_packet,
                                        int* span_too_long) {
  if (!s || !events || !n) return set_error(CMAXB_ERR_INVALID, "null argument");
  *events = nullptr; *n = 0;
  // "once the whole event packet is received" (:103)
  if (s->subsets_info.empty() || !(s->num_event_total > s->subsets_info.front().second)) return 1;   // nothing ready
  // getEventSubset (:138-147)
  s->ev_beg_idx = s->subsets_info.front().first;
  s->ev_end_idx = s->subsets_info.front().second;
  s->subsets_info.pop_front();
  if (s->ev_beg_idx < 0 || s->ev_end_idx > (long long)s->events.size() || s->ev_beg_idx >= s->ev_end_idx)
    return set_error(CMAXB_ERR_STATE, "packet indices outside the event store");
  s->packet_cur ^= 1;
  StageBuf& b = s->packet[s->packet_cur];
  const size_t cnt = (size_t)(s->ev_end_idx - s->ev_beg_idx);
  CMAXB_TRY(b.reserve(cnt));
  std::copy(s->events.begin() + s->ev_beg_idx, s->events.begin() + s->ev_end_idx, b.p);
  *events = b.p; *n = cnt;
  if (time_packet) *time_packet = s->time_packet;
  if (span_too_long) {
    // timespan_packet > 10 * dt_ang_vel => the reference assumes zero angular velocity (:109-114)
    const cmaxb_event& f = b.p[0]; const cmaxb_event& l = b.p[cnt - 1];
    long long ds = (long long)l.sec - (long long)f.sec, dn = (long long)l.nsec - (long long)f.nsec;
    if (dn < 0) { dn += 1000000000ll; --ds; }
    const double span = (double)(int)ds + 1e-9 * (double)(int)dn;
    *span_too_long = span > 10 * s->cfg.dt_ang_vel ? 1 : 0;
  }
  s->time_packet = sadd(s->time_packet, s->dt_av);                    // slideWindow (:176-183)
  return CMAXB_OK;
}

// AngVelEstimator::deleteOldEvents (:149-174)
static void stream_delete_old(cmaxb_stream* s, long long idx_backend) {
  const long long del = std::min(idx_backend, s->ev_beg_idx);
  if (del <= 0) return;
  s->events.erase(s->events.begin(), s->events.begin() + del);
  s->num_event_total -= del;
  s->ev_beg_idx -= del; s->ev_end_idx -= del;
  for (auto& si : s->subsets_info) { si.first -= del; si.second -= del; }
  for (auto& m : s->ts_map) m.second -= del;
}

extern "C" int cmaxb_stream_window_events(cmaxb_stream* s, cmaxb_stamp t_beg, cmaxb_stamp t_end, const cmaxb_event** events, size_t* n) {
  if (!s || !events || !n) return set_error(CMAXB_ERR_INVALID, "null argument");
  *events = nullptr; *n = 0;
  // PoseGraphOptimizer::getEventSubset (:133-166)
  auto ib = std::upper_bound(s->ts_map.begin(), s->ts_map.end(), t_beg,
                             [](cmaxb_stamp t, const std::pair<cmaxb_stamp, long long>& a) { return slt(t, a.first); });
  auto ie = std::lower_bound(s->ts_map.begin(), s->ts_map.end(), t_end,
                             [](const std::pair<cmaxb_stamp, long long>& a, cmaxb_stamp t) { return slt(a.first, t); });
  if (ib == s->ts_map.end() || ie == s->ts_map.end())
    return set_error(CMAXB_ERR_STATE, "the event store does not cover the window yet (the reference would dereference map.end())");
  const long long beg = ib->second;
  long long end = ie->second;
  const cmaxb_stamp t_end_mod = sadd(t_end, SDur{0, -1000});          // t_end - ros::Duration(1e-6)
  if (beg < 0 || end >= (long long)s->events.size()) return set_error(CMAXB_ERR_STATE, "window indices outside the event store");
  while (slt(t_end_mod, ev_ts(s->events[(size_t)end]))) {             // events.at(end).ts > t_end_mod
    end -= 100;
    if (end <= beg) { end = beg + 1; break; }
  }
  const size_t cnt = end > beg ? (size_t)(end - beg) : 0;
  CMAXB_TRY(s->window.reserve(cnt));
  std::copy(s->events.begin() + beg, s->events.begin() + end, s->window.p);
  *events = s->window.p; *n = cnt;
  s->ts_map.erase(s->ts_map.begin(), ib + 1);                         // erase(begin, std::next(ev_beg_iter))
  stream_delete_old(s, beg);
  return CMAXB_OK;
}

extern "C" int cmaxb_stream_state(cmaxb_stream* s, int64_t* n_stored, int64_t* n_subsets_pending, int64_t* n_ts_map, cmaxb_stamp* time_packet) {
  if (!s) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (n_stored) *n_stored = s->num_event_total;
  if (n_subsets_pending) *n_subsets_pending = (int64_t)s->subsets_info.size();
  if (n_ts_map) *n_ts_map = (int64_t)s->ts_map.size();
  if (time_packet) *time_packet = s->time_packet;
  return CMAXB_OK;
}
