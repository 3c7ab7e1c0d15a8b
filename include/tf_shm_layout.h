/*
 * tf_shm_layout.h -- byte layout of the per-pod soft-limiter quota file.
 *
 * Boundary #2 of the drop-in: the Go hypervisor creates and
 * updates this file (pkg/hypervisor/worker/state/soft_limiter_shm.go), the
 * limiter inside the worker maps it.  The file is Rust's
 *   #[repr(C)] enum SharedDeviceState { V1(..), V2(SharedDeviceStateV2) }
 * i.e. a u32 discriminant, 4 bytes of padding, then the V2 payload
 * (soft_limiter_shm.go:345-352).  Offsets below are the ones the reference's
 * own tests pin (soft_limiter_shm_test.go:217-232: payload @8, LastHeartbeat
 * @0x890, PIDs @0x898) and the sizes Go's unsafe.Sizeof yields on amd64.
 *
 * All multi-byte fields are little-endian; the four erl_* words hold IEEE-754
 * float64 bit patterns and are only ever touched with 64-bit atomics.
 */
#ifndef TF_SHM_LAYOUT_H
#define TF_SHM_LAYOUT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TF_SHM_MAX_DEVICES 16      /* soft_limiter_shm.go:21 */
#define TF_SHM_MAX_PROCESSES 2048  /* soft_limiter_shm.go:20 */
#define TF_SHM_UUID_LEN 64         /* soft_limiter_shm.go:22 */
#define TF_SHM_DISCRIMINANT_V2 1u  /* soft_limiter_shm.go:26 */
#define TF_SHM_FILE_NAME "shm"     /* soft_limiter_shm.go:23 */

typedef struct {
  char uuid[TF_SHM_UUID_LEN];      /* +0   NUL-terminated, at most 63 chars (:262-275) */
  uint32_t up_limit;               /* +64  percent 0..100 */
  uint32_t pad0_;
  uint64_t mem_limit;              /* +72  bytes */
  uint32_t total_cuda_cores;       /* +80 */
  uint32_t pad1_;
  uint64_t pod_memory_used;        /* +88  bytes, atomic store by the hypervisor @2 Hz */
  uint64_t erl_token_refill_rate;  /* +96  f64 bits, tokens/s   (init 10.0)  */
  uint64_t erl_token_capacity;     /* +104 f64 bits             (init 100.0) */
  uint64_t erl_current_tokens;     /* +112 f64 bits, CAS target (init 100.0) */
  uint64_t erl_last_token_update;  /* +120 f64 bits, unix seconds */
  uint32_t is_active;              /* +128 0/1 */
  uint32_t pad2_;
} tf_shm_device_entry;             /* 136 bytes */

typedef struct {
  uint64_t lock;                              /* holder PID or 0 (:842-878) */
  uint64_t values[TF_SHM_MAX_PROCESSES];      /* :779-783 */
  uint64_t bitmap[TF_SHM_MAX_PROCESSES];      /* oversized on purpose; bit i = word i/64, mask (1<<63)>>(i&63) (:751-776) */
  uint64_t len;
  uint64_t creator_pid;
} tf_shm_pid_registry;                        /* 32 792 bytes */

typedef struct {
  uint32_t discriminant;                      /* 0x0000 = 1 */
  uint32_t pad_;
  tf_shm_device_entry devices[TF_SHM_MAX_DEVICES]; /* 0x0008, indexed by device index (not packed) */
  uint32_t device_count;                      /* 0x0888 */
  uint32_t pad1_;
  uint64_t last_heartbeat;                    /* 0x0890 unix seconds */
  tf_shm_pid_registry pids;                   /* 0x0898 */
  uint8_t reserved[512];                      /* 0x88B0 */
} tf_shm_file;                                /* 35 504 bytes */

#define TF_SHM_FILE_BYTES 35504u
#define TF_SHM_LEGACY_BYTES 35496u /* the V2 payload without the enum header: rejected (:981-987) */

#if defined(__cplusplus)
#define TF_SHM_ASSERT(c, m) static_assert(c, m)
#else
#define TF_SHM_ASSERT(c, m) _Static_assert(c, m)
#endif
TF_SHM_ASSERT(sizeof(tf_shm_device_entry) == 136, "DeviceEntryV2 is 136 bytes");
TF_SHM_ASSERT(offsetof(tf_shm_device_entry, erl_current_tokens) == 112, "token word @112");
TF_SHM_ASSERT(offsetof(tf_shm_device_entry, is_active) == 128, "is_active @128");
TF_SHM_ASSERT(offsetof(tf_shm_file, devices) == 8, "V2 payload starts at 8");
TF_SHM_ASSERT(offsetof(tf_shm_file, device_count) == 0x888, "device_count @0x888");
TF_SHM_ASSERT(offsetof(tf_shm_file, last_heartbeat) == 0x890, "LastHeartbeat @0x890");
TF_SHM_ASSERT(offsetof(tf_shm_file, pids) == 0x898, "PIDs @0x898");
TF_SHM_ASSERT(offsetof(tf_shm_pid_registry, bitmap) == 16392, "bitmap @+16392");
TF_SHM_ASSERT(offsetof(tf_shm_pid_registry, len) == 32776, "len @+32776");
TF_SHM_ASSERT(offsetof(tf_shm_pid_registry, creator_pid) == 32784, "pid @+32784");
TF_SHM_ASSERT(offsetof(tf_shm_file, reserved) == 0x88B0, "padding @0x88B0");
TF_SHM_ASSERT(sizeof(tf_shm_file) == TF_SHM_FILE_BYTES, "quota file is 35 504 bytes");

#ifdef __cplusplus
}
#endif
#endif /* TF_SHM_LAYOUT_H */
