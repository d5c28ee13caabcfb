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

// The host-side event store (events_ of the reference): a queue of contiguous segments, each either owned (a copy of
// the pushed message, subsampled) or BORROWED (the caller's page-locked buffer, referenced in place).  Indices are
// relative to the oldest stored event, exactly like the indices of the reference's std::vector after its erase().
struct EventStore {
  struct Seg { const cmaxb_event* p; size_t n; size_t off; std::vector<cmaxb_event>* own; };   // events p[off .. n)
  std::deque<Seg> segs;
  std::deque<long long> starts;      // index of every segment's first live event
  long long count = 0;
  ~EventStore() { for (auto& g : segs) delete g.own; }
  long long size() const { return count; }
  void push_borrow(const cmaxb_event* p, size_t n) {
    if (!n) return;
    starts.push_back(count); segs.push_back(Seg{p, n, 0, nullptr}); count += (long long)n;
  }
  // copies msg[0], msg[stride], ...; returns the owned chunk
  const std::vector<cmaxb_event>* push_copy(const cmaxb_event* msg, size_t n, size_t stride) {
    auto* v = new std::vector<cmaxb_event>();
    v->reserve((n + stride - 1) / stride);
    for (size_t i = 0; i < n; i += stride) v->push_back(msg[i]);
    if (v->empty()) { delete v; return nullptr; }
    starts.push_back(count); segs.push_back(Seg{v->data(), v->size(), 0, v}); count += (long long)v->size();
    return v;
  }
  size_t seg_of(long long i) const {
    return (size_t)(std::upper_bound(starts.begin(), starts.end(), i) - starts.begin()) - 1;
  }
  const cmaxb_event& at(long long i) const {
    const size_t k = seg_of(i);
    return segs[k].p[segs[k].off + (size_t)(i - starts[k])];
  }
  void copy_range(long long beg, long long end, cmaxb_event* dst) const {
    if (end <= beg) return;
    size_t k = seg_of(beg);
    long long i = beg;
    while (i < end) {
      const Seg& g = segs[k];
      const long long lo = i - starts[k], hi = std::min<long long>((long long)(g.n - g.off), end - starts[k]);
      std::copy(g.p + g.off + lo, g.p + g.off + hi, dst);
      dst += hi - lo; i += hi - lo; ++k;
    }
  }
  void erase_front(long long del) {
    count -= del;
    while (del > 0 && !segs.empty()) {
      Seg& g = segs.front();
      const long long live = (long long)(g.n - g.off);
      if (del >= live) { del -= live; delete g.own; segs.pop_front(); starts.pop_front(); }
      else { g.off += (size_t)del; del = 0; }
    }
    long long run = 0;
    for (size_t k = 0; k < segs.size(); ++k) { starts[k] = run; run += (long long)(segs[k].n - segs[k].off); }
  }
};

struct cmaxb_stream {
  cmaxb_stream_cfg cfg{};
  SDur dt_av{};
  EventStore events;                               // events_
  long long num_event_total = 0;                   // num_event_total_ (== events.size())
  long long abs_front = 0;                         // absolute number (since creation) of the oldest stored event
  bool sliding_window_initialized = false;
  cmaxb_stamp time_packet{}, time_get_subset{};
  int num_ev_half_packet = 0;
  std::deque<std::pair<long long, long long>> subsets_info;       // event_subsets_info_
  std::vector<std::pair<cmaxb_stamp, long long>> ts_map;          // ev_subset_ts_map_ (sorted by stamp, unique)
  long long ev_beg_idx = 0, ev_end_idx = 0;
  StageBuf packet[2]; int packet_cur = 0;          // double-buffered: packet i stays valid while packet i+1 is cut
  StageBuf window;
  // device-resident mirror of the store (cmaxb_stream_attach_device): a ring of `cap` events whose first `mirror`
  // events are duplicated behind its end, so that every packet is one contiguous device range
  int device = -1; cudaStream_t cu_stream = nullptr;
  cmaxb_event* d_ring = nullptr; size_t cap = 0, mirror = 0;
  StageBuf bounce[2]; int bounce_cur = 0; cudaEvent_t bounce_done[2] = {nullptr, nullptr};   // owned copies travel through pinned memory
  cudaEvent_t copied = nullptr;                    // recorded on cu_stream after the copies of every push
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
  if (s->d_ring) {
    cudaSetDevice(s->device);
    if (s->cu_stream) cudaStreamSynchronize(s->cu_stream);
    cudaFree(s->d_ring);
    for (int i = 0; i < 2; ++i) if (s->bounce_done[i]) cudaEventDestroy(s->bounce_done[i]);
    if (s->copied) cudaEventDestroy(s->copied);
  }
  s->packet[0].release(); s->packet[1].release(); s->window.release();
  s->bounce[0].release(); s->bounce[1].release();
  delete s;
}

extern "C" int cmaxb_stream_attach_device(cmaxb_stream* s, int device, void* cuda_stream, size_t ring_events) {
  if (!s) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (s->d_ring) return set_error(CMAXB_ERR_STATE, "device store already attached");
  if (s->num_event_total > 0) return set_error(CMAXB_ERR_STATE, "attach the device store before the first push");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { (void)cudaGetLastError(); return set_error(CMAXB_ERR_CUDA, "no such CUDA device"); }
  CMAXB_CUDA_TRY(cudaSetDevice(device));
  const size_t L = (size_t)2 * s->num_ev_half_packet;                 // a packet is [total - half, total + half)
  size_t C = ring_events ? ring_events : 8 * L;
  if (C < 2 * L) C = 2 * L;
  CMAXB_CUDA_TRY(cudaMalloc((void**)&s->d_ring, sizeof(cmaxb_event) * (C + L)));
  for (int i = 0; i < 2; ++i) CMAXB_CUDA_TRY(cudaEventCreateWithFlags(&s->bounce_done[i], cudaEventDisableTiming));
  CMAXB_CUDA_TRY(cudaEventCreateWithFlags(&s->copied, cudaEventDisableTiming));
  s->device = device; s->cu_stream = (cudaStream_t)cuda_stream; s->cap = C; s->mirror = L;
  return CMAXB_OK;
}

// events [abs0, abs0 + m) from src (host memory) into the device ring (+ the mirror of the ring's first events)
static int stream_to_device(cmaxb_stream* s, long long abs0, const cmaxb_event* src, size_t m) {
  const size_t C = s->cap, L = s->mirror;
  size_t done = 0;
  while (done < m) {
    const size_t pos = (size_t)((abs0 + (long long)done) % (long long)C);
    const size_t cnt = std::min(m - done, C - pos);
    CMAXB_CUDA_TRY(cudaMemcpyAsync(s->d_ring + pos, src + done, sizeof(cmaxb_event) * cnt, cudaMemcpyHostToDevice, s->cu_stream));
    if (pos < L) {
      const size_t mc = std::min(cnt, L - pos);
      // the mirror of the ring's first events (packets stay contiguous across the wrap): device to device, not over PCIe again
      CMAXB_CUDA_TRY(cudaMemcpyAsync(s->d_ring + C + pos, s->d_ring + pos, sizeof(cmaxb_event) * mc, cudaMemcpyDeviceToDevice, s->cu_stream));
    }
    done += cnt;
  }
  return CMAXB_OK;
}

// pushEvent without the solve (ang_vel_estimator.cpp:68-100) for the event that becomes number `total` (1-based) of the store
static void stream_note_event(cmaxb_stream* s, cmaxb_stamp ts, long long total) {
  const long long beg = std::max(total - (long long)s->num_ev_half_packet, 0ll);
  const long long end = total + (long long)s->num_ev_half_packet;
  s->subsets_info.emplace_back(beg, end);
  // std::map::insert keeps an existing key
  auto it = std::lower_bound(s->ts_map.begin(), s->ts_map.end(), ts,
                             [](const std::pair<cmaxb_stamp, long long>& a, cmaxb_stamp t) { return slt(a.first, t); });
  if (!(it != s->ts_map.end() && it->first.sec == ts.sec && it->first.nsec == ts.nsec)) s->ts_map.insert(it, {ts, total - 1});
  s->time_get_subset = sadd(s->time_get_subset, s->dt_av);
}

extern "C" int cmaxb_stream_push_ex(cmaxb_stream* s, const cmaxb_event* msg_events, size_t n, int flags, int* packets_ready) {
  if (!s || (!msg_events && n > 0)) return set_error(CMAXB_ERR_INVALID, "null argument");
  const size_t stride = (size_t)s->cfg.event_sample_rate;
  const bool borrow = (flags & CMAXB_PUSH_BORROW) != 0;
  if (borrow && stride != 1) return set_error(CMAXB_ERR_INVALID, "CMAXB_PUSH_BORROW needs event_sample_rate 1 (the message is referenced in place)");
  if (n > 0) {
    // eventsCallback: for (ev = begin; ev < end; ev += event_sample_rate) pushEvent(*ev)
    const long long total0 = s->num_event_total;
    const cmaxb_event* src;          // the m events that enter the store, contiguous
    size_t m;
    if (borrow) { s->events.push_borrow(msg_events, n); src = msg_events; m = n; }
    else {
      const std::vector<cmaxb_event>* v = s->events.push_copy(msg_events, n, stride);
      src = v ? v->data() : nullptr; m = v ? v->size() : 0;
    }
    if (m > 0) {
      if (!s->sliding_window_initialized) {
        const SDur half = sdur_from_sec(((double)s->dt_av.sec + 1e-9 * (double)s->dt_av.nsec) * 0.5);   // dt_av_ * 0.5
        s->time_packet = sadd(ev_ts(src[0]), half);
        s->time_get_subset = s->time_packet;
        s->sliding_window_initialized = true;
      }
      if (flags & CMAXB_PUSH_SORTED) {
        // time-sorted message: the first event beyond time_get_subset_ is found by bisection, then the search resumes
        // behind it with the advanced cursor -- the same events as the event-by-event scan selects
        size_t from = 0;
        while (from < m) {
          const cmaxb_stamp tg = s->time_get_subset;
          const cmaxb_event* it = std::upper_bound(src + from, src + m, tg, [](cmaxb_stamp t, const cmaxb_event& e) { return slt(t, ev_ts(e)); });
          if (it == src + m) break;
          const size_t k = (size_t)(it - src);
          stream_note_event(s, ev_ts(*it), total0 + (long long)k + 1);
          from = k + 1;
        }
      } else {
        for (size_t k = 0; k < m; ++k)
          if (slt(s->time_get_subset, ev_ts(src[k]))) stream_note_event(s, ev_ts(src[k]), total0 + (long long)k + 1);   // event.ts > time_get_subset_
      }
      s->num_event_total = total0 + (long long)m;
      if (s->d_ring) {
        CMAXB_CUDA_TRY(cudaSetDevice(s->device));
        const long long abs0 = s->abs_front + total0;
        if (borrow) CMAXB_TRY(stream_to_device(s, abs0, src, m));
        else {
          // owned copies live in pageable memory: stage them through a pinned bounce buffer so that the copy is a DMA
          s->bounce_cur ^= 1;
          StageBuf& bb = s->bounce[s->bounce_cur];
          if (bb.p) CMAXB_CUDA_TRY(cudaEventSynchronize(s->bounce_done[s->bounce_cur]));
          CMAXB_TRY(bb.reserve(m));
          std::copy(src, src + m, bb.p);
          CMAXB_TRY(stream_to_device(s, abs0, bb.p, m));
          CMAXB_CUDA_TRY(cudaEventRecord(s->bounce_done[s->bounce_cur], s->cu_stream));
        }
        CMAXB_CUDA_TRY(cudaEventRecord(s->copied, s->cu_stream));
      }
    }
  }
  if (packets_ready) {
    int k = 0;
    for (const auto& si : s->subsets_info) { if (s->num_event_total > si.second) ++k; else break; }
    *packets_ready = k;
  }
  return CMAXB_OK;
}

extern "C" int cmaxb_stream_push(cmaxb_stream* s, const cmaxb_event* msg_events, size_t n, int* packets_ready) {
  return cmaxb_stream_push_ex(s, msg_events, n, 0, packets_ready);
}

// getEventSubset (:138-147) + the span test (:109-114) + slideWindow (:176-183); fills ev_beg_idx / ev_end_idx
static int stream_take_packet(cmaxb_stream* s, cmaxb_stamp* time_packet, int* span_too_long) {
  // "once the whole event packet is received" (:103)
  if (s->subsets_info.empty() || !(s->num_event_total > s->subsets_info.front().second)) return 1;   // nothing ready
  s->ev_beg_idx = s->subsets_info.front().first;
  s->ev_end_idx = s->subsets_info.front().second;
  s->subsets_info.pop_front();
  if (s->ev_beg_idx < 0 || s->ev_end_idx > s->events.size() || s->ev_beg_idx >= s->ev_end_idx)
    return set_error(CMAXB_ERR_STATE, "packet indices outside the event store");
  if (time_packet) *time_packet = s->time_packet;
  if (span_too_long) {
    // timespan_packet > 10 * dt_ang_vel => the reference assumes zero angular velocity (:109-114)
    const cmaxb_event& f = s->events.at(s->ev_beg_idx); const cmaxb_event& l = s->events.at(s->ev_end_idx - 1);
    long long ds = (long long)l.sec - (long long)f.sec, dn = (long long)l.nsec - (long long)f.nsec;
    if (dn < 0) { dn += 1000000000ll; --ds; }
    const double span = (double)(int)ds + 1e-9 * (double)(int)dn;
    *span_too_long = span > 10 * s->cfg.dt_ang_vel ? 1 : 0;
  }
  s->time_packet = sadd(s->time_packet, s->dt_av);                    // slideWindow (:176-183)
  return CMAXB_OK;
}

extern "C" int cmaxb_stream_next_packet(cmaxb_stream* s, const cmaxb_event** events, size_t* n, cmaxb_stamp* time_packet,
                                        int* span_too_long) {
  if (!s || !events || !n) return set_error(CMAXB_ERR_INVALID, "null argument");
  *events = nullptr; *n = 0;
  const int rc = stream_take_packet(s, time_packet, span_too_long);
  if (rc != CMAXB_OK) return rc;
  s->packet_cur ^= 1;
  StageBuf& b = s->packet[s->packet_cur];
  const size_t cnt = (size_t)(s->ev_end_idx - s->ev_beg_idx);
  CMAXB_TRY(b.reserve(cnt));
  s->events.copy_range(s->ev_beg_idx, s->ev_end_idx, b.p);
  *events = b.p; *n = cnt;
  return CMAXB_OK;
}

extern "C" int cmaxb_stream_next_packet_device(cmaxb_stream* s, const cmaxb_event** device_events, size_t* n, cmaxb_stamp* time_packet,
                                               int* span_too_long) {
  if (!s || !device_events || !n) return set_error(CMAXB_ERR_INVALID, "null argument");
  *device_events = nullptr; *n = 0;
  if (!s->d_ring) return set_error(CMAXB_ERR_STATE, "no device store: call cmaxb_stream_attach_device first");
  // the packet must still be in the ring: checked BEFORE it is consumed
  if (!s->subsets_info.empty() && s->num_event_total > s->subsets_info.front().second &&
      s->num_event_total - s->subsets_info.front().first > (long long)s->cap)
    return set_error(CMAXB_ERR_STATE, "the device ring no longer holds the packet (consumer too far behind): use cmaxb_stream_next_packet");
  const int rc = stream_take_packet(s, time_packet, span_too_long);
  if (rc != CMAXB_OK) return rc;
  const size_t cnt = (size_t)(s->ev_end_idx - s->ev_beg_idx);
  if (cnt > s->mirror) return set_error(CMAXB_ERR_STATE, "packet longer than the ring's mirror");
  const size_t pos = (size_t)((s->abs_front + s->ev_beg_idx) % (long long)s->cap);
  *device_events = s->d_ring + pos;      // contiguous thanks to the mirror behind the ring's end
  *n = cnt;
  return CMAXB_OK;
}

// AngVelEstimator::deleteOldEvents (:149-174)
static void stream_delete_old(cmaxb_stream* s, long long idx_backend) {
  const long long del = std::min(idx_backend, s->ev_beg_idx);
  if (del <= 0) return;
  s->events.erase_front(del);
  s->abs_front += del;
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
    return 1;   // the event store does not cover the window yet (the reference would dereference map.end()): not an error, try later
  const long long beg = ib->second;
  long long end = ie->second;
  const cmaxb_stamp t_end_mod = sadd(t_end, SDur{0, -1000});          // t_end - ros::Duration(1e-6)
  if (beg < 0 || end >= s->events.size()) return set_error(CMAXB_ERR_STATE, "window indices outside the event store");
  while (slt(t_end_mod, ev_ts(s->events.at(end)))) {                  // events.at(end).ts > t_end_mod
    end -= 100;
    if (end <= beg) { end = beg + 1; break; }
  }
  const size_t cnt = end > beg ? (size_t)(end - beg) : 0;
  CMAXB_TRY(s->window.reserve(cnt));
  s->events.copy_range(beg, end, s->window.p);
  *events = s->window.p; *n = cnt;
  s->ts_map.erase(s->ts_map.begin(), ib + 1);                         // erase(begin, std::next(ev_beg_iter))
  stream_delete_old(s, beg);
  return CMAXB_OK;
}

// makes `consumer_stream` wait for the device copies of everything pushed so far (needed when the packets are consumed
// on another stream than the one the store copies on, e.g. a dedicated copy stream that overlaps the evaluations)
extern "C" int cmaxb_stream_wait_copied(cmaxb_stream* s, void* consumer_stream) {
  if (!s) return set_error(CMAXB_ERR_INVALID, "null argument");
  if (!s->d_ring) return set_error(CMAXB_ERR_STATE, "no device store");
  CMAXB_CUDA_TRY(cudaSetDevice(s->device));
  if ((cudaStream_t)consumer_stream != s->cu_stream && s->num_event_total + s->abs_front > 0)
    CMAXB_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)consumer_stream, s->copied, 0));
  return CMAXB_OK;
}

// number of events (counted from the first push) the store has dropped: borrowed message buffers that lie entirely
// below this mark may be reused by the caller
extern "C" int cmaxb_stream_released(cmaxb_stream* s, int64_t* n_released) {
  if (!s || !n_released) return set_error(CMAXB_ERR_INVALID, "null argument");
  *n_released = s->abs_front;
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
