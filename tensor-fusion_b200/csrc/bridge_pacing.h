// bridge_pacing.h -- when may the limiter bridge move tokens from the quota file to the device bucket?
//
// Plain arithmetic, no CUDA: quota_bridge.cc calls it every period, tools/bridge_pacing_sim.cc replays it against a
// model of the hypervisor's 2 Hz refill (tests/test_cpu_host.py).  See DESIGN 5.3 "Pacing".
#pragma once

namespace tfw {

struct Pacer {
  double carry = 0.0;       // credit: tokens the bridge may still move at the controller's rate
  bool paced = true;        // false: move whatever fits as soon as the file has it (round-1 behaviour)
  double quantum_s = 0.05;  // tokens are handed over in bursts worth this much time at the controller's rate
};

// One bridge period of `dt` seconds.  `headroom` = free room in the device bucket (window - tokens it holds).
// Returns how many tokens may move now (0 .. headroom).
//
// The hypervisor refills the file in one lump per 500 ms tick (rate * dt, quota_controller.go:349-376).  Handing a
// saturating vGPU the whole lump at once makes it run flat out for a fraction of the tick and starve until the next.
// Metering one launch at a time is no answer either: tenants are separate processes, the GPU time-slices between their
// contexts, and evenly interleaved 200 us kernels pay a context switch each.  So the file's tokens are metered out at the
// controller's own rate IN BURSTS worth `quantum` of that rate, and unused credit accumulates up to the file's capacity,
// so a burst after idle time gets its burst.
//
// Credit accrues only while the file holds tokens to spend it on.  An empty file means the tenant is ahead of the
// controller already: credit saved up while starving would let it swallow the next lump in one go, starve for the rest
// of that tick, save up again ... (measured on B200: a 60 ms burst and a 440 ms stall in every tick).  With the hypervisor
// gone (stale heartbeat) nobody fills the file, the bridge mints at the last rate itself and the credit runs with the clock.
inline double pace_headroom(Pacer& p, double rate, double file_capacity, double window, double headroom, double dt,
                            bool file_has_tokens, bool hypervisor_alive) {
  if (!p.paced) return headroom;
  if (file_has_tokens || !hypervisor_alive) p.carry += rate * dt;
  const double carry_cap = file_capacity > window ? file_capacity : window;
  if (p.carry > carry_cap) p.carry = carry_cap;
  const double quantum = rate * p.quantum_s;
  if (p.carry < quantum && p.carry < carry_cap) return 0.0;  // not a burst's worth yet
  return headroom > p.carry ? p.carry : headroom;
}

// Tokens actually taken out of the file are paid for with credit.
inline void pace_spent(Pacer& p, double taken) {
  if (p.paced) p.carry -= taken;
}

// The device bucket holds one burst (at least two of the largest launch, at most what the file may hold).
inline double pace_window(const Pacer& p, double rate, double file_capacity, double prepaid_s, double max_cost) {
  double window = rate * (p.paced && p.quantum_s > prepaid_s ? p.quantum_s : prepaid_s);
  const double floor_ = 2.0 * max_cost;
  if (window < floor_) window = floor_;
  if (file_capacity > 0.0 && window > file_capacity) window = file_capacity;
  return window;
}

}  // namespace tfw
