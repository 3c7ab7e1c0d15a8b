/*
 * tfw_stats_file.h -- the per-worker counters record a vGPU worker publishes next to its
 * quota file (<dir of TF_SHM_PATH>/tfw_stats) and the provider folds into
 * AccelGetDeviceMetrics().extraMetrics (provider/accelerator.h:195-229).  The hypervisor
 * writes every extra metric as a field of its `tf_gpu_usage` line
 * (pkg/hypervisor/metrics/metrics.go:135-139), so staged-GB/s, swap and throttle counters reach
 * the existing metrics pipeline without a Go change (the schema already has
 * compute_throttled_cnt, internal/metrics/types.go:181-183).
 */
#ifndef TFW_STATS_FILE_H
#define TFW_STATS_FILE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFW_STATS_MAGIC 0x53574654u /* 'TFWS' */
#define TFW_STATS_VERSION 2u
#define TFW_STATS_FILE_NAME "tfw_stats"
#define TFW_STATS_STALE_SECS 15u /* records older than this are ignored (dead worker) */

typedef struct {
  uint32_t magic, version;
  uint64_t seq;               /* odd while the writer is mid-update (seqlock) */
  uint64_t pid;
  uint64_t updated_unix_secs;
  char device_uuid[64];       /* "GPU-xxxxxxxx-...." as NVML prints it */
  uint64_t frames, payload_bytes, h2d_dma_bytes, d2h_bytes, d2d_bytes, fill_bytes;
  uint64_t mover_launches, client_launches, gate_launches;
  uint64_t vram_bytes, vram_peak_bytes, live_buffers;
  uint64_t gate_admitted, gate_blocked, gate_timeouts;
  /* control channel, provider -> worker ("send snapshot command to worker via shared memory",
   * pkg/hypervisor/server/handlers/worker.go:94-129): AccelSnapshot / AccelResume write ctl_request,
   * the worker executes it at its next poll (tfw_worker_poll_control) and answers in ctl_ack. */
  uint64_t ctl_request;     /* (sequence << 8) | TFW_CTL_* ; 0 = none yet */
  uint64_t ctl_ack;         /* the request the worker completed last */
  uint64_t ctl_status;      /* tfw_status of that request */
  uint64_t ctl_frozen;      /* 1 while the vGPU is frozen */
  uint64_t ctl_moved_bytes; /* bytes the last freeze moved out of HBM */
  uint64_t parked_bytes;    /* bytes currently held in host memory for a frozen vGPU */
  uint64_t ctl_arg;         /* argument of ctl_request (TFW_CTL_SM_LIMIT: percent, TFW_CTL_MEM_LIMIT: bytes); written before it */
  uint64_t reserved;
  /* version 2: who this worker is to the hypervisor (FreezeWorker / ResumeWorker / AutoFreeze / AutoResume of
   * provider/limiter.h:77-81 name a worker, not a process) and when it froze */
  char worker_id[64];       /* $TF_WORKER_ID, else $POD_UID, else "<POD_NAMESPACE>/<POD_NAME>", else "" */
  uint64_t frozen_unix_ms;  /* when the current freeze began; 0 while running */
  uint64_t frozen_auto;     /* 1 if the worker froze itself (idle longer than auto_freeze.freeze_to_mem_ttl) */
  uint64_t auto_freezes, auto_resumes;
  uint64_t sm_limit_percent;  /* hard compute limit in force (0 = whole GPU) and the SMs it translates to */
  uint64_t sm_count;
  uint64_t vram_limit_bytes;  /* hard memory limit in force (0 = none) */
} tfw_stats_record;

#define TFW_CTL_FREEZE 1u
#define TFW_CTL_RESUME 2u
#define TFW_CTL_SM_LIMIT 3u  /* AccelSetComputeUnitHardLimit: ctl_arg = percent of the SMs */
#define TFW_CTL_MEM_LIMIT 4u /* AccelSetMemHardLimit: ctl_arg = bytes */

#ifdef __cplusplus
}
#endif
#endif
