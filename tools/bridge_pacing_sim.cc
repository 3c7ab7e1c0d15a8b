// bridge_pacing_sim.cc -- the limiter bridge's pacing rule (csrc/bridge_pacing.h) against a model of what it sits between:
// a hypervisor that refills the quota file in one lump per 500 ms tick (quota_controller.go:349-376) and a saturating
// tenant that launches a 200 us kernel whenever the device bucket holds a token.  Prints one JSON object per scenario:
// how evenly the launches were admitted.  No GPU, no CUDA: the very function quota_bridge.cc calls.
//
//   bridge_pacing_sim <rate tokens/s> <seconds> <mode>
//     mode: paced        the product rule
//           credit-while-starving   the rule before the fix: credit accrues whatever the file holds
//           unpaced      TFW_BRIDGE_PACED=0
//           idle-then-burst   the tenant sleeps for 2 s, then saturates: how much may it burst?
//           hypervisor-dies   the ticks stop after 1 s (heartbeat goes stale): the bridge mints at the last rate
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bridge_pacing.h"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const double rate = atof(argv[1]), seconds = atof(argv[2]);
  const char* mode = argv[3];
  const bool always_credit = !strcmp(mode, "credit-while-starving");
  const bool idle_first = !strcmp(mode, "idle-then-burst");
  const bool dies = !strcmp(mode, "hypervisor-dies");
  tfw::Pacer p;
  p.paced = strcmp(mode, "unpaced") != 0;
  const double step = 0.002, tick = 0.5, kernel_s = 200e-6, cost = 1.0;
  const double capacity = std::max(rate * 0.5, 200.0);  // quota_controller.go:425-433
  double file = capacity, bucket = 0.0;
  p.carry = capacity;                                    // quota_bridge_start: a fresh vGPU may burst like a full bucket
  double next_tick = tick, busy_until = 0.0, last_launch = -1.0, max_gap = 0.0, first_burst = 0.0;
  long launches = 0, launches_after_idle_100ms = 0;
  std::vector<double> gaps;
  for (double t = 0.0; t < seconds; t += step) {
    const bool alive = !(dies && t > 1.0 + 10.0);        // heartbeat stale after 10 s without a tick
    if (t >= next_tick) {
      if (!(dies && t > 1.0)) file = std::min(capacity, file + rate * tick);
      next_tick += tick;
    }
    const double window = tfw::pace_window(p, rate, capacity, 0.02, cost);
    double room = window - bucket;
    room = tfw::pace_headroom(p, rate, capacity, window, room, step, always_credit || file > 0.0, alive);
    if (room > 0.0) {
      double take = alive ? std::min(room, file) : std::min(room, rate * step);
      if (alive) { file -= take; tfw::pace_spent(p, take); }
      bucket += take;
    }
    // the tenant: back-to-back 200 us kernels while tokens last
    const bool idle = idle_first && t < 2.0;
    for (double u = std::max(t, busy_until); !idle && u < t + step && bucket >= cost; u += kernel_s) {
      bucket -= cost;
      busy_until = u + kernel_s;
      if (last_launch >= 0.0 && t > (idle_first ? 2.2 : 2.0)) { gaps.push_back(u - last_launch); max_gap = std::max(max_gap, u - last_launch); }
      last_launch = u;
      ++launches;
      if (idle_first && u < 2.1) ++launches_after_idle_100ms;
    }
    (void)first_burst;
  }
  std::sort(gaps.begin(), gaps.end());
  const double p99 = gaps.empty() ? 0.0 : gaps[(size_t)(gaps.size() * 0.99)];
  printf("{\"mode\": \"%s\", \"rate\": %.1f, \"seconds\": %.1f, \"launches\": %ld, \"launches_per_s\": %.1f, \"max_gap_ms\": %.2f, \"p99_gap_ms\": %.2f, "
         "\"launches_in_first_100ms_after_idle\": %ld, \"capacity\": %.1f}\n",
         mode, rate, seconds, launches, launches / seconds, max_gap * 1e3, p99 * 1e3, launches_after_idle_100ms, capacity);
  return 0;
}
