/*
 * tfw_gate.h -- C-ABI of the device-resident ERL token bucket (north_star (b)).
 *
 * Reference semantics restated on the GPU (bit-exact, float64 carried in a
 * 64-bit word and updated with compare-and-swap):
 *   FetchSubERLTokens  pkg/hypervisor/worker/state/soft_limiter_shm.go:715-731
 *   FetchAddERLTokens  pkg/hypervisor/worker/state/soft_limiter_shm.go:734-748
 * The reference evaluates the gate on the CPU inside an LD_PRELOAD hook
 * (provider/limiter.h:71-75, CheckAndRecordComputeOps); here the bucket lives
 * in HBM and a one-thread gate kernel enqueued in front of every client launch
 * takes the tokens in stream order, so the host thread that deserializes the
 * command stream never blocks on the limiter.
 *
 * The hypervisor's PID controller (quota_controller.go:378-458, unchanged Go)
 * keeps writing rate / capacity / refills into the shared-memory quota file; a
 * bridge thread inside this library moves tokens from that file into the
 * device bucket (see DESIGN.md "limiter bridge").
 */
#ifndef TFW_GATE_H
#define TFW_GATE_H

#include <stddef.h>
#include <stdint.h>

#include "tfw_worker.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfw_gate tfw_gate;

typedef struct {
  double tokens;        /* device bucket */
  double capacity;      /* device bucket capacity */
  double refill_rate;   /* tokens/s as last seen in the quota file (0 if standalone) */
  uint64_t admitted;    /* successful FetchSub operations */
  uint64_t denied;      /* FetchSub attempts that found tokens < cost (try: 1 per call; blocking: 1 per poll) */
  uint64_t blocked_gates; /* blocking gates that found too few tokens when enqueued (compute_throttled_cnt) */
  uint64_t wait_ns;     /* device-side wait time; only measured by the spin fallback (TFW_GATE_SPIN=1) */
  uint64_t bridged_tokens_milli; /* tokens moved quota file -> device, x1000 */
  uint64_t timeouts;    /* blocking gates released by the fail-open timer (5 s) */
} tfw_gate_state;

typedef struct {
  uint32_t kind;   /* 0 = FetchSub(amount) non-blocking, 1 = FetchAdd(amount), 2 = SetCapacity(amount), 3 = SetTokens(amount) */
  uint32_t pad;
  double amount;
} tfw_gate_op;

/* shm_path == NULL: stand-alone bucket (tokens 100, capacity 100: the quota
 * file's initial values, soft_limiter_shm.go:186-189). */
TFW_API tfw_status tfw_gate_create(int device, const char* shm_path, uint32_t device_index, tfw_gate** out);
TFW_API tfw_status tfw_gate_destroy(tfw_gate* g);
/* What happens to a gate whose tokens never come (dead hypervisor, stalled controller):
 * fail_closed = 0 (default; env TFW_GATE_FAIL_POLICY=open): the watchdog releases it after max_wait_ms
 * (availability over isolation; counted in `timeouts`); fail_closed = 1 (TFW_GATE_FAIL_POLICY=closed,
 * tfw_config flag TFW_F_GATE_FAIL_CLOSED): it waits for its refill, however long.  max_wait_ms <= 0
 * keeps the current value (default 5000, env TFW_GATE_MAX_WAIT_MS). */
TFW_API tfw_status tfw_gate_set_policy(tfw_gate* g, int fail_closed, double max_wait_ms);

/* One non-blocking FetchSubERLTokens executed by a kernel. *before = value
 * found; *admitted = 1 iff tokens were taken. */
TFW_API tfw_status tfw_gate_try(tfw_gate* g, double cost, double* before, int* admitted);
/* Stream-ordered blocking gate: kernels enqueued on `cuda_stream` after this call start
 * only after `cost` tokens were taken from the bucket (a cost above the bucket's capacity is charged
 * as one full bucket: FetchSub can never admit more than the capacity).  The wait is a stream memory
 * operation (cuStreamWaitValue64 on the token word, GEQ bits(cost)) followed by a one-thread
 * take kernel, so a throttled vGPU keeps no kernel running while it waits; a watchdog
 * releases a gate that waited longer than 5 s (fail-open, counted in `timeouts`). */
TFW_API tfw_status tfw_gate_enqueue(tfw_gate* g, double cost, void* cuda_stream);
/* FetchAddERLTokens on the device bucket (capped at capacity). */
TFW_API tfw_status tfw_gate_refill(tfw_gate* g, double amount, double* before);
TFW_API tfw_status tfw_gate_set_capacity(tfw_gate* g, double capacity);
TFW_API tfw_status tfw_gate_set_tokens(tfw_gate* g, double tokens);
TFW_API tfw_status tfw_gate_get_state(tfw_gate* g, tfw_gate_state* out);
/* the same for the gate a worker created from its tfw_config.shm_path (TFW_ERR_NOT_FOUND: the vGPU has no limiter) */
TFW_API tfw_status tfw_worker_gate_state(tfw_worker* w, tfw_gate_state* out);
/* Run a recorded sequence of bucket operations in ONE single-thread kernel and
 * return the value found before each op (parity test vs the oracle). */
TFW_API tfw_status tfw_gate_run_sequence(tfw_gate* g, const tfw_gate_op* ops, uint32_t n, double* before);
/* Contended variant: `nthreads` device threads (one per CTA) each issue
 * `per_thread` FetchSub(cost); returns how many were admitted.  Conservation
 * (admitted*cost + tokens_left == tokens_initial, exactly, for integer costs)
 * is the size-independent property checked at scale. */
TFW_API tfw_status tfw_gate_contend(tfw_gate* g, uint32_t nthreads, uint32_t per_thread, double cost,
                                    uint64_t* admitted);

#ifdef __cplusplus
}
#endif
#endif /* TFW_GATE_H */
