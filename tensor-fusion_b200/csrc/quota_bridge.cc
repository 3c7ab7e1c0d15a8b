// quota_bridge.cc -- the limiter bridge thread.
//
// The unchanged Go hypervisor keeps running its PID controller
// (pkg/hypervisor/worker/computing/quota_controller.go:378-458) against the
// quota file: it stores rate/capacity and FetchAdd()s refills into
// erl_current_tokens at 2 Hz.  The reference limiter consumes those tokens
// with FetchSubERLTokens from the CPU, once per kernel launch.  Here the
// consumer is a GPU kernel, so a small prepaid window of tokens is moved
//   quota file --take_up_to()--> device bucket --gate kernel--> launches
// every `period`.  Token conservation: tokens only ever move, the sum
// (file + device) changes exactly by what the hypervisor adds and what the
// gate kernels consume; the window is sized so that the hypervisor's drain
// logic (rebalanceTokenBucket, :349-376) still sees most of the balance.
#include "quota_bridge.h"

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <string>
#include <thread>

#include "bridge_pacing.h"
#include "shm_quota.h"

namespace tfw {

struct QuotaBridge {
  tfw_gate* gate = nullptr;
  tfq::QuotaFile* file = nullptr;
  uint32_t idx = 0;
  std::thread thr;
  std::atomic<bool> stop{false};
  std::atomic<double> max_cost{1.0};
  std::atomic<double> rate{0.0};
  std::atomic<double> file_cap{0.0};
  std::atomic<uint64_t> moved_milli{0};
  double window = 0.0;
  unsigned period_us = 2000;
  double prepaid_s = 0.02;
  Pacer pace;  // TFW_BRIDGE_PACED=0 / TFW_BRIDGE_QUANTUM_MS (bridge_pacing.h)
};

static void bridge_loop(QuotaBridge* b) {
  cudaSetDevice(gate_device(b->gate));
  auto last = std::chrono::steady_clock::now();
  while (!b->stop.load(std::memory_order_acquire)) {
    const auto now = std::chrono::steady_clock::now();
    const double dt = std::chrono::duration<double>(now - last).count();
    last = now;
    if (b->file->has_device(b->idx)) {
      const double rate = b->file->rate(b->idx);
      const double cap = b->file->capacity(b->idx);
      b->rate.store(rate, std::memory_order_relaxed);
      b->file_cap.store(cap, std::memory_order_relaxed);
      const double window = pace_window(b->pace, rate, cap, b->prepaid_s, b->max_cost.load(std::memory_order_relaxed));
      if (window != b->window) {
        tfw_gate_set_capacity(b->gate, window);
        b->window = window;
      }
      const uint64_t unix_now = (uint64_t)time(nullptr);
      double take = 0.0;
      double headroom = window - gate_mirror_tokens(b->gate);
      const bool alive = b->file->is_healthy(10, unix_now);
      headroom = pace_headroom(b->pace, rate, cap, window, headroom, dt, b->file->tokens(b->idx) > 0.0, alive);  // bridge_pacing.h
      if (headroom > 0.0) {
        if (alive) {
          take = b->file->take_up_to(b->idx, headroom);
          pace_spent(b->pace, take);
        } else {
          // hypervisor gone (heartbeat stale > 10 s): keep enforcing the last
          // rate it set instead of starving or un-limiting the vGPU.
          take = rate * dt;
          if (take > headroom) take = headroom;
        }
      }
      if (take > 0.0) {
        tfw_gate_refill(b->gate, take, nullptr);
        b->moved_milli.fetch_add((uint64_t)(take * 1000.0), std::memory_order_relaxed);
      }
    }
    std::this_thread::sleep_for(std::chrono::microseconds(b->period_us));
  }
}

tfw_status quota_bridge_start(tfw_gate* g, const char* shm_file, uint32_t device_index, QuotaBridge** out) {
  if (!g || !shm_file || !out || device_index >= TF_SHM_MAX_DEVICES) return TFW_ERR_INVALID;
  tfq::QuotaFile* f = nullptr;
  std::string err;
  tfq::Status s = tfq::QuotaFile::open_file(shm_file, &f, &err);
  if (s != tfq::kOk) return s == tfq::kNotFound ? TFW_ERR_NOT_FOUND : TFW_ERR_INVALID;
  QuotaBridge* b = new QuotaBridge();
  b->gate = g;
  b->file = f;
  b->idx = device_index;
  if (const char* e = getenv("TFW_BRIDGE_PERIOD_US")) { int v = atoi(e); if (v >= 100) b->period_us = (unsigned)v; }
  if (const char* e = getenv("TFW_BRIDGE_PREPAID_MS")) { double v = atof(e); if (v > 0) b->prepaid_s = v / 1000.0; }
  if (const char* e = getenv("TFW_BRIDGE_PACED")) b->pace.paced = !(e[0] == '0');
  if (const char* e = getenv("TFW_BRIDGE_QUANTUM_MS")) { double v = atof(e); if (v > 0) b->pace.quantum_s = v / 1000.0; }
  if (f->has_device(device_index)) b->pace.carry = f->capacity(device_index);  // a fresh vGPU may burst like a full bucket
  // the device bucket starts empty: every token it ever holds came out of the file
  tfw_gate_set_tokens(g, 0.0);
  if (f->has_device(device_index)) b->file_cap.store(f->capacity(device_index), std::memory_order_relaxed);
  b->thr = std::thread(bridge_loop, b);
  *out = b;
  return TFW_OK;
}

void quota_bridge_stop(QuotaBridge* b) {
  if (!b) return;
  b->stop.store(true, std::memory_order_release);
  if (b->thr.joinable()) b->thr.join();
  // hand unspent prepaid tokens back to the file
  tfw_gate_state st{};
  if (tfw_gate_get_state(b->gate, &st) == TFW_OK && st.tokens > 0.0 && b->file->has_device(b->idx)) {
    tfw_gate_set_tokens(b->gate, 0.0);
    b->file->give_back(b->idx, st.tokens);
  }
  delete b->file;
  delete b;
}

void quota_bridge_note_cost(QuotaBridge* b, double cost) {
  double cur = b->max_cost.load(std::memory_order_relaxed);
  while (cost > cur && !b->max_cost.compare_exchange_weak(cur, cost)) {}
}
double quota_bridge_rate(QuotaBridge* b) { return b->rate.load(std::memory_order_relaxed); }
double quota_bridge_capacity(QuotaBridge* b) { return b->file_cap.load(std::memory_order_relaxed); }
uint64_t quota_bridge_moved_milli(QuotaBridge* b) { return b->moved_milli.load(std::memory_order_relaxed); }

}  // namespace tfw
