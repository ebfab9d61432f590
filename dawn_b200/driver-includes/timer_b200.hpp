// Event timer base class of generated stencils.  API of the reference's timer_cuda
// (driver-includes/timer_cuda.hpp:39-84: start/pause/reset/total_time in seconds) without its
// per-run cudaEventSynchronize: event pairs are recorded on the stencil's stream and only resolved
// when total_time() is asked for, so back-to-back run() calls never stall the host.
#pragma once
#include "b200_runtime.hpp"

#include <string>
#include <utility>
#include <vector>

namespace dawn_b200 {

class timer_b200 {
  std::string name_;
  void* stream_ = nullptr;
  std::vector<std::pair<void*, void*>> pending_, pool_;
  std::pair<void*, void*> cur_{nullptr, nullptr};
  double total_s_ = 0;
  bool enabled_ = true;

public:
  explicit timer_b200(std::string name) : name_(std::move(name)) {}
  timer_b200(const timer_b200&) = delete;
  ~timer_b200() {
    for(auto& p : pending_)
      b200rt_event_destroy(p.first), b200rt_event_destroy(p.second);
    for(auto& p : pool_)
      b200rt_event_destroy(p.first), b200rt_event_destroy(p.second);
  }
  const std::string& name() const { return name_; }
  void* stream() const { return stream_; }
  void set_stream(void* s) { stream_ = s; }
  void enable_meters(bool on) { enabled_ = on; }

  void start() {
    if(!enabled_)
      return;
    if(pool_.empty()) {
      check(b200rt_event_create(&cur_.first));
      check(b200rt_event_create(&cur_.second));
    } else {
      cur_ = pool_.back();
      pool_.pop_back();
    }
    check(b200rt_event_record(cur_.first, stream_));
  }
  void pause() {
    if(!enabled_ || !cur_.first)
      return;
    check(b200rt_event_record(cur_.second, stream_));
    pending_.push_back(cur_);
    cur_ = {nullptr, nullptr};
    if(pending_.size() >= 512)
      resolve();
  }
  void resolve() {
    for(auto& p : pending_) {
      float ms = 0;
      check(b200rt_event_synchronize(p.second));
      check(b200rt_event_elapsed_ms(p.first, p.second, &ms));
      total_s_ += ms * 1e-3;
      pool_.push_back(p);
    }
    pending_.clear();
  }
  void reset() {
    resolve();
    total_s_ = 0;
  }
  double total_time() {
    resolve();
    return total_s_;
  }
};

} // namespace dawn_b200
