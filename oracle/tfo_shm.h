/* tfo_shm.h -- oracle API for the quota file + ERL controller.  TEST INFRASTRUCTURE. */
#ifndef TFO_SHM_H
#define TFO_SHM_H
#include <stddef.h>
#include <stdint.h>
#include <time.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint32_t device_idx;
  char uuid[128];
  uint32_t up_limit;
  uint64_t mem_limit;
  uint32_t total_cuda_cores;
} tfo_dev_cfg;

typedef struct tfo_shm tfo_shm;

size_t tfo_shm_file_bytes(void);
size_t tfo_shm_legacy_bytes(void);
size_t tfo_shm_offset(const char* what);
int tfo_shm_init_image(uint8_t* f, const tfo_dev_cfg* cfgs, size_t n, uint64_t now, uint64_t creator_pid);
int tfo_shm_create(const char* base, const char* ns, const char* pod, const tfo_dev_cfg* cfgs, size_t n, tfo_shm** out);
int tfo_shm_open(const char* base, const char* ns, const char* pod, tfo_shm** out);
uint8_t* tfo_shm_data(tfo_shm* h);
void tfo_shm_close(tfo_shm* h);

/* field: 0 refill rate, 1 capacity, 2 current tokens, 3 last update */
double tfo_shm_get(uint8_t* f, uint32_t idx, int field);
void tfo_shm_set(uint8_t* f, uint32_t idx, int field, double v);
double tfo_shm_fetch_sub(uint8_t* f, uint32_t idx, double cost);
double tfo_shm_fetch_add(uint8_t* f, uint32_t idx, double amount);
int tfo_shm_has_device(uint8_t* f, uint32_t idx);
int tfo_shm_set_pod_memory_used(uint8_t* f, uint32_t idx, uint64_t v);
uint64_t tfo_shm_pod_memory_used(uint8_t* f, uint32_t idx);
int tfo_shm_is_healthy(uint8_t* f, uint64_t timeout_secs, uint64_t now);
int tfo_shm_pid_insert(uint8_t* f, uint64_t pid);
int tfo_shm_pid_remove(uint8_t* f, uint64_t pid);
size_t tfo_shm_pid_values(uint8_t* f, uint64_t* out, size_t cap);
int tfo_valid_component(const char* s);
int tfo_from_shm_path(const char* path, char* ns, char* name, size_t cap);
double tfo_go_max(double x, double y);
double tfo_go_min(double x, double y);

/* ---- ERL controller (erl_oracle.c) ---- */
typedef struct {
  double burst_window, rate_min, rate_max, capacity_min, capacity_max, util_alpha, kp, ki, kd, integral_decay;
} tfo_erl_cfg;
typedef struct {
  double current_rate, smoothed_util, integral_err, last_error;
  int initialized;
} tfo_erl_state;
void tfo_erl_default_cfg(tfo_erl_cfg* c);
int tfo_erl_cfg_from_json(const char* json, tfo_erl_cfg* c);
void tfo_erl_new_state(tfo_erl_state* s);
double tfo_erl_slew(double current, double target, double up, double down);
double tfo_erl_compute_desired_rate(double current_rate, double target_util, double smoothed_util, double dt,
                                    tfo_erl_state* es, const tfo_erl_cfg* cfg);
double tfo_erl_rebalance(uint8_t* f, uint32_t idx, double now_secs, double refill_rate, double capacity,
                         double target_util, double smoothed_util);
/* per-(worker,device) body of updateERLControllers; returns current tokens */
double tfo_erl_tick(uint8_t* f, uint32_t idx, tfo_erl_state* es, const tfo_erl_cfg* cfg, uint32_t up_limit,
                    double nvml_util_percent, double now_secs);
uint32_t tfo_compute_up_limit(int64_t compute_percent, double tflops_limit, double max_tflops);
double tfo_gate_bench(int nthreads, uint64_t ops_per_thread, double cost, double* deny_ratio);
#ifdef __cplusplus
}
#endif
#endif
