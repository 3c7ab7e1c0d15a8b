/*
 * erl_oracle.c -- CPU restatement of the elastic-rate-limit controller.
 * TEST INFRASTRUCTURE.  Follows
 *   pkg/hypervisor/worker/computing/quota_controller.go
 *     constants :20-47, defaultERLConfig :118-131, loadERLConfigFromEnv :143-178,
 *     getOrCreateERLState :253-268, slewRate :314-319, computeDesiredRate :321-347,
 *     rebalanceTokenBucket :349-376, updateERLControllers :378-458 (loop body)
 *   pkg/hypervisor/worker/controller.go computeUpLimit :307-325
 * Go on amd64 does not fuse multiply-add; build with -ffp-contract=off so the
 * float64 results are bit-identical.  Pinned by quota_controller_test.go:11-73.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tfo_shm.h"

static const double kUpdateInterval = 0.5;
static const double kUtilDeadband = 0.03, kIntegralClamp = 1.5;
static const double kMaxRateIncreaseRatio = 0.35, kMaxRateDecreaseRatio = 0.25;
static const double kTokenReserveRatio = 0.35, kTokenDrainRatio = 0.80, kTokenDrainMin = 25.0;

static double clampf(double v, double lo, double hi) { return tfo_go_min(tfo_go_max(v, lo), hi); }

void tfo_erl_default_cfg(tfo_erl_cfg* c) {
  c->burst_window = 0.5; c->rate_min = 10.0; c->rate_max = 200000.0;
  c->capacity_min = 200.0; c->capacity_max = 200000.0;
  c->util_alpha = 0.25; c->kp = 0.9; c->ki = 0.35; c->kd = 0.10; c->integral_decay = 0.85;
}
void tfo_erl_new_state(tfo_erl_state* s) { memset(s, 0, sizeof *s); s->current_rate = 100.0; }

/* parsePositiveFloat :133-141 applied to "key":"value" of the flat JSON object */
static double json_pos(const char* json, const char* key, double fallback) {
  char pat[96];
  snprintf(pat, sizeof pat, "\"%s\"", key);
  const char* k = strstr(json, pat);
  if (!k) return fallback;
  k += strlen(pat);
  while (*k && (isspace((unsigned char)*k) || *k == ':')) ++k;
  if (*k != '"') return fallback;
  ++k;
  char buf[64];
  size_t n = 0;
  while (*k && *k != '"' && n + 1 < sizeof buf) buf[n++] = *k++;
  buf[n] = 0;
  char* s = buf;
  while (isspace((unsigned char)*s)) ++s;
  if (!*s) return fallback;
  char* end;
  double v = strtod(s, &end);
  while (isspace((unsigned char)*end)) ++end;
  if (*end || !(v > 0)) return fallback;
  return v;
}
int tfo_erl_cfg_from_json(const char* json, tfo_erl_cfg* c) {
  tfo_erl_default_cfg(c);
  if (!json || !*json) return 0;
  if (!strstr(json, "elasticRateLimitParameters")) return 0;
  c->rate_max = json_pos(json, "maxRefillRate", c->rate_max);
  c->rate_min = json_pos(json, "minRefillRate", c->rate_min);
  c->util_alpha = json_pos(json, "filterAlpha", c->util_alpha);
  c->kp = json_pos(json, "kp", c->kp);
  c->ki = json_pos(json, "ki", c->ki);
  c->kd = json_pos(json, "kd", c->kd);
  c->burst_window = json_pos(json, "burstWindow", c->burst_window);
  c->capacity_min = json_pos(json, "capacityMin", c->capacity_min);
  c->capacity_max = json_pos(json, "capacityMax", c->capacity_max);
  c->integral_decay = json_pos(json, "integralDecayFactor", c->integral_decay);
  if (c->rate_min > c->rate_max) c->rate_min = c->rate_max;
  if (c->capacity_min > c->capacity_max) c->capacity_min = c->capacity_max;
  c->util_alpha = clampf(c->util_alpha, 0.01, 0.95);
  c->integral_decay = clampf(c->integral_decay, 0.01, 0.999);
  return 0;
}

double tfo_erl_slew(double current, double target, double up, double down) {
  if (target > current) return tfo_go_min(target, current * (1.0 + up));
  return tfo_go_max(target, current * (1.0 - down));
}

double tfo_erl_compute_desired_rate(double current_rate, double target_util, double smoothed_util, double dt,
                                    tfo_erl_state* es, const tfo_erl_cfg* cfg) {
  if (smoothed_util <= 0.01) return tfo_go_min(current_rate * (1.0 + kMaxRateIncreaseRatio), cfg->rate_max);
  double error = target_util - smoothed_util;
  if (fabs(error) < kUtilDeadband) {
    es->integral_err *= cfg->integral_decay;
    return current_rate;
  }
  es->integral_err = clampf(es->integral_err * cfg->integral_decay + error * dt, -kIntegralClamp, kIntegralClamp);
  double derivative = 0.0;
  if (dt > 0) derivative = (error - es->last_error) / dt;
  es->last_error = error;
  double feed_forward = current_rate * (target_util / tfo_go_max(smoothed_util, 0.05));
  double control = 1.0 + cfg->kp * error + cfg->ki * es->integral_err + cfg->kd * derivative;
  control = clampf(control, 0.5, 1.5);
  double desired = clampf(feed_forward * control, cfg->rate_min, cfg->rate_max);
  return tfo_erl_slew(current_rate, desired, kMaxRateIncreaseRatio, kMaxRateDecreaseRatio);
}

double tfo_erl_rebalance(uint8_t* f, uint32_t idx, double now_secs, double refill_rate, double capacity,
                         double target_util, double smoothed_util) {
  double current = tfo_shm_get(f, idx, 2), last = tfo_shm_get(f, idx, 3);
  if (last > 0) {
    double elapsed = now_secs - last;
    if (elapsed > 0 && elapsed < 5.0) {
      tfo_shm_fetch_add(f, idx, refill_rate * elapsed);
      current = tfo_shm_get(f, idx, 2);
    }
  }
  double reserve = clampf(capacity * kTokenReserveRatio, 0.0, capacity);
  if (current > capacity) {
    double drain = tfo_go_max(kTokenDrainMin, capacity * kTokenDrainRatio) * kUpdateInterval;
    current = tfo_go_max(capacity, current - drain);
    tfo_shm_set(f, idx, 2, current);
  } else if (smoothed_util > target_util + kUtilDeadband && current > reserve) {
    double drain = tfo_go_max(kTokenDrainMin, capacity * kTokenDrainRatio) * kUpdateInterval;
    current = tfo_go_max(reserve, current - drain);
    tfo_shm_set(f, idx, 2, current);
  }
  tfo_shm_set(f, idx, 3, now_secs);
  return current;
}

double tfo_erl_tick(uint8_t* f, uint32_t idx, tfo_erl_state* es, const tfo_erl_cfg* cfg, uint32_t up_limit,
                    double nvml_util_percent, double now_secs) {
  const double dt = kUpdateInterval;
  double target = (double)up_limit / 100.0;
  double util = nvml_util_percent / 100.0;
  if (!es->initialized) { es->smoothed_util = util; es->initialized = 1; }
  else es->smoothed_util = cfg->util_alpha * util + (1 - cfg->util_alpha) * es->smoothed_util;
  es->current_rate = tfo_erl_compute_desired_rate(es->current_rate, target, es->smoothed_util, dt, es, cfg);
  double cap = clampf(es->current_rate * cfg->burst_window, cfg->capacity_min, cfg->capacity_max);
  tfo_shm_set(f, idx, 0, es->current_rate);
  tfo_shm_set(f, idx, 1, cap);
  return tfo_erl_rebalance(f, idx, now_secs, es->current_rate, cap, target, es->smoothed_util);
}

uint32_t tfo_compute_up_limit(int64_t compute_percent, double tflops_limit, double max_tflops) {
  if (compute_percent > 0) return (uint32_t)compute_percent;
  if (tflops_limit > 0 && max_tflops > 0) {
    double percent = ceil(tflops_limit / max_tflops * 100.0);
    if (percent < 1) return 1;
    if (percent > 100) return 100;
    return (uint32_t)percent;
  }
  return 100;
}
