/*
 * tfw_worker.h -- C-ABI of the B200 vGPU worker data path (libtfw_b200.so).
 *
 * This is the boundary the Go host (pkg/hypervisor / the worker process the
 * operator starts with `./tensor-fusion-worker -p 8000`, reference:
 * internal/utils/compose.go:1304-1325, pkg/constants/env.go:155-156) binds
 * through cgo.  The reference implements this side in a closed binary
 * (README.md:131); the entry points below are what a Go `net.Conn` read loop
 * needs to hand a forwarded-CUDA byte stream to the GPU:
 *
 *   conn.Read(ring) -> tfw_submit() -> [deserialize | H2D DMA | unpack kernel]
 *   tfw_poll_responses() -> conn.Write()
 *
 * Plain pointers and sizes only; no C++/torch types.  All functions are
 * thread-compatible per worker (one submitting thread per worker, any number
 * of workers per process) and return tfw_status; values 0..6 coincide with the
 * reference's AccelResult (provider/accelerator.h:47-55) so a Go caller can
 * map both through one table.
 *
 * There is NO CPU fallback: without a usable CUDA device every entry point
 * that touches the data path returns TFW_ERR_NO_DEVICE.
 */
#ifndef TFW_WORKER_H
#define TFW_WORKER_H

#include <stddef.h>
#include <stdint.h>

#include "tfw_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TFW_API __attribute__((visibility("default")))

typedef enum {
  TFW_OK = 0,
  TFW_ERR_INVALID = 1,       /* == ACCEL_ERROR_INVALID_PARAM */
  TFW_ERR_NOT_FOUND = 2,     /* unknown handle */
  TFW_ERR_NOT_SUPPORTED = 3,
  TFW_ERR_EXHAUSTED = 4,     /* VRAM quota / ring space */
  TFW_ERR_FAILED = 5,        /* CUDA call failed; see tfw_last_error */
  TFW_ERR_INTERNAL = 6,
  TFW_ERR_PROTOCOL = 7,      /* malformed frame */
  TFW_ERR_NO_DEVICE = 8      /* CUDA unavailable: the worker refuses to run */
} tfw_status;

typedef struct tfw_worker tfw_worker;
typedef struct tfw_trace tfw_trace;

/* Flags for tfw_config.flags */
#define TFW_F_MOVER_TMA 0x1u      /* table-driven batches whose copies are congruent modulo 16 go to the TMA bulk mover
                                     (one cp.async.bulk load + store per 32 KiB tile); env TFW_MOVER=tma|ldg overrides */
#define TFW_F_MOVER_LDG 0x2u      /* force the 16-B vector ld/st mover */
#define TFW_F_NO_ZERO_FILL 0x4u   /* do not scrub fresh allocations (native-CUDA semantics) */
#define TFW_F_NO_LIMITER 0x8u     /* DISABLE_GPU_LIMITER (pkg/constants/env.go:140-146) */
#define TFW_F_GATE_FAIL_CLOSED 0x10u /* a launch whose tokens never come waits for ever instead of being released after 5 s */

typedef struct {
  uint32_t struct_size;      /* sizeof(tfw_config), for forward compatibility */
  int32_t device;            /* CUDA ordinal of the home GPU */
  uint64_t chunk_bytes;      /* staging slot size; 0 = default (32 MiB) */
  uint32_t num_slots;        /* staging slots (ring depth); 0 = default (4) */
  uint32_t flags;            /* TFW_F_* */
  uint64_t vram_limit_bytes; /* hard VRAM quota (TF_CUDA_MEMORY_LIMIT); 0 = unlimited */
  const char* shm_path;      /* quota file (TF_SHM_PATH); NULL = no soft limiter */
  uint32_t shm_device_index; /* device entry inside the quota file */
  uint32_t mover_ctas_per_sm; /* 0 = default */
  /* Optional (struct_size >= 56): allocate client buffers inside a tiered vGPU address space
   * (include/tfw_vram.h).  Points at a tfw_vspace_config; home_device is overridden by `device`.
   * MALLOC then hands out region-aligned ranges of that space, regions are made resident on
   * first touch, cold ones move to peer HBM / host DRAM, and client handles keep working. */
  const void* tiering;
  /* Optional (struct_size >= 64): hard compute limit in percent of the GPU's SMs (TF_CUDA_SM_PERCENT_LIMIT,
   * internal/utils/compose.go:1287-1295; AccelSetComputeUnitHardLimit).  1..99: the vGPU's stream lives in a
   * green context that owns that share of the SMs -- every kernel of the tenant runs there and nowhere else.
   * 0 or 100: the whole GPU. */
  uint32_t sm_percent_limit;
  uint32_t reserved0;
} tfw_config;
#define TFW_CONFIG_SIZE_V1 48u /* the layout without `tiering` is still accepted */
#define TFW_CONFIG_SIZE_V2 56u /* ... and the one without `sm_percent_limit` */

typedef struct {
  uint64_t frames;            /* frames executed */
  uint64_t payload_bytes;     /* H2D payload bytes staged into HBM */
  uint64_t d2h_bytes;         /* bytes returned to the client */
  uint64_t d2d_bytes;
  uint64_t fill_bytes;
  uint64_t h2d_dma_bytes;     /* bytes moved host->device by the copy engine */
  uint64_t mover_launches;    /* unpack/scatter kernel launches */
  uint64_t gate_launches;     /* limiter gate kernels enqueued */
  uint64_t client_launches;   /* TFCS_OP_LAUNCH kernels launched */
  uint64_t batches_hazard;    /* batches cut by a RAW/WAW/WAR hazard */
  uint64_t vram_bytes;        /* bytes currently allocated to this vGPU */
  uint64_t vram_peak_bytes;
  uint64_t live_buffers;
  uint64_t other_launches;    /* digest etc. */
  /* ABI version 2 */
  uint64_t h2d_ref_bytes;     /* H2D bytes DMA'd straight from client arenas (no staging, no unpack kernel) */
  uint64_t d2h_ref_bytes;     /* D2H bytes DMA'd straight into client arenas */
  uint64_t user_launches;     /* TFCS_OP_LAUNCH_USER kernels (subset of client_launches) */
} tfw_stats;

/* One unit of work of the byte-mover kernel (device addresses). */
typedef struct {
  uint64_t dst;
  uint64_t src;      /* 0 => fill */
  uint64_t len;
  uint32_t tile0;    /* exclusive prefix sum of tiles; filled by the library */
  uint32_t fill;     /* fill byte replicated x4 when src == 0 */
} tfw_move_desc;

/* ---- lifecycle ---------------------------------------------------------- */
TFW_API tfw_status tfw_worker_create(const tfw_config* cfg, tfw_worker** out);
TFW_API tfw_status tfw_worker_destroy(tfw_worker* w);
TFW_API const char* tfw_last_error(const tfw_worker* w); /* never NULL */
TFW_API uint32_t tfw_abi_version(void);

/* ---- pinned host memory (the receive ring lives in it) ------------------ */
TFW_API tfw_status tfw_host_alloc(size_t bytes, void** out);
TFW_API tfw_status tfw_host_free(void* p);
TFW_API tfw_status tfw_host_register(void* p, size_t bytes);
TFW_API tfw_status tfw_host_unregister(void* p);

/* ---- streaming data path ------------------------------------------------ */
/* Feed wire bytes.  `stream` must start at a frame boundary plus whatever
 * the previous call left unconsumed (the library keeps no partial frames:
 * `*consumed` tells the caller how many bytes were whole frames; re-submit
 * the remainder together with the next bytes).  If `stream` lies in memory
 * obtained from tfw_host_alloc / tfw_host_register the DMA engine reads it in
 * place; otherwise it is copied through the worker's internal pinned ring.
 * Work is enqueued asynchronously; the memory must stay valid until
 * tfw_flush() returns. */
TFW_API tfw_status tfw_submit(tfw_worker* w, const void* stream, size_t nbytes, size_t* consumed);
/* Fences for double-buffered receive rings: *ticket names everything submitted so far;
 * tfw_fence_wait returns once the GPU (DMA reads of the host memory included) is done with
 * it, i.e. the memory handed to those tfw_submit calls may be overwritten. */
TFW_API tfw_status tfw_fence(tfw_worker* w, uint64_t* ticket);
TFW_API tfw_status tfw_fence_wait(tfw_worker* w, uint64_t ticket);
/* Non-blocking form: *done = 1 when everything submitted before the ticket has left the host buffers. */
TFW_API tfw_status tfw_fence_query(tfw_worker* w, uint64_t ticket, int* done);
/* Freeze / resume the vGPU ("freeze to mem", api/v1/schedulingconfigtemplate_types.go:221-231;
 * provider/limiter.h:77-81 FreezeWorker/ResumeWorker; handlers/legacy.go:111-139 HandleTrap):
 * freeze drains the vGPU stream and releases the vGPU's HBM to other tenants: on a tiered worker every
 * resident region moves to the host tier (pointers stay valid, regions come back on first touch after
 * resume); on a plain worker every buffer is copied to host memory and freed, and resume allocates and
 * refills them (clients hold handles, not pointers, so the move is invisible to them).  While frozen
 * tfw_submit answers TFW_ERR_NOT_SUPPORTED.  *moved_bytes (optional) = bytes moved out by this call.
 * Resume answers TFW_ERR_EXHAUSTED, and the vGPU stays frozen, if the HBM is not available yet. */
TFW_API tfw_status tfw_worker_freeze(tfw_worker* w, uint64_t* moved_bytes);
TFW_API tfw_status tfw_worker_resume(tfw_worker* w);
/* The worker's own idle policy ("auto_freeze": {"freeze_to_mem_ttl": ...} of the hypervisor's RemotePodInfo,
 * pkg/hypervisor/api/http_types.go:82-100; AutoFreeze / AutoResume of provider/limiter.h:80-81): the same freeze,
 * remembered as self-inflicted -- tfw_worker_auto_resume undoes only such a freeze, one ordered through
 * AccelSnapshot / FreezeWorker stays until AccelResume / ResumeWorker. */
TFW_API tfw_status tfw_worker_auto_freeze(tfw_worker* w, uint64_t* moved_bytes);
TFW_API tfw_status tfw_worker_auto_resume(tfw_worker* w);
/* Change the hard limits of a running vGPU (AccelSetComputeUnitHardLimit / AccelSetMemHardLimit reach the worker
 * through the control words of its stats record).  The compute limit drains the vGPU once and moves its stream
 * into a new SM partition; the memory limit applies to further MALLOCs. */
TFW_API tfw_status tfw_worker_set_sm_limit(tfw_worker* w, uint32_t percent);
TFW_API tfw_status tfw_worker_set_vram_limit(tfw_worker* w, uint64_t bytes);
/* Execute a pending freeze / resume request of the provider (AccelSnapshot / AccelResume write it into
 * the worker's stats record, include/tfw_stats_file.h).  Call it from the thread that owns the worker,
 * between submits; costs one memory read when nothing is pending.  *frozen (optional) = state after the call:
 * 0 running, 1 frozen by the provider, 2 frozen by the worker's own idle policy. */
TFW_API tfw_status tfw_worker_poll_control(tfw_worker* w, int* frozen);
/* Same-node transports: where the client's page-locked host memory lives.  Arena k of a session is
 * the file "<prefix>.a<k>" (the shm transport passes its ring file's path); the worker maps and
 * page-locks it on TFCS_OP_HOST_REGISTER, and MEMCPY_*_REF frames make the copy engine move bytes
 * between those pages and HBM directly.  Without a prefix HOST_REGISTER answers NOT_SUPPORTED. */
TFW_API tfw_status tfw_set_arena_prefix(tfw_worker* w, const char* prefix);
/* Produce responses straight into a page-locked byte ring shared with the client (the worker ->
 * client ring of include/tfw_shm_ring.h): headers are written by the CPU, D2H payloads by the copy
 * engine, and *head is advanced (release) over each piece once the GPU has completed it; *tail is
 * the client's consumer cursor.  With a sink tfw_poll_responses copies nothing: it only publishes
 * (*nbytes stays 0), and a D2H that finds the ring full makes tfw_submit answer TFW_ERR_EXHAUSTED
 * with the frame kept: call tfw_submit again (zero bytes are fine) once the client has consumed. */
typedef struct {
  uint32_t struct_size;
  uint32_t reserved;
  void* ring;
  uint64_t ring_bytes;       /* multiple of 64 */
  uint64_t* head;            /* producer cursor, monotonic byte count (written by the library) */
  const uint64_t* tail;      /* consumer cursor (read by the library) */
} tfw_response_sink;
TFW_API tfw_status tfw_set_response_sink(tfw_worker* w, const tfw_response_sink* sink);
/* Block until every submitted frame has executed on the GPU. */
TFW_API tfw_status tfw_flush(tfw_worker* w);
/* Drain response frames (RESP_D2H / RESP_SYNC / RESP_ERROR) produced so far, in order, as a byte
 * stream: a response larger than `cap` is handed out in pieces over consecutive calls.  *nbytes == 0
 * means nothing is ready yet. */
TFW_API tfw_status tfw_poll_responses(tfw_worker* w, void* out, size_t cap, size_t* nbytes);

/* ---- recorded-trace replay with the trace resident in HBM --------------- */
/* Parse a whole trace, copy its bytes into HBM once, pre-build every batch's
 * descriptor table on the device.  tfw_trace_replay() then only launches
 * kernels: this is the "inputs already resident in HBM" leg of the bench. */
TFW_API tfw_status tfw_trace_load(tfw_worker* w, const void* stream, size_t nbytes, tfw_trace** out);
TFW_API tfw_status tfw_trace_replay(tfw_worker* w, tfw_trace* t);
/* Device time of the last replay's mover kernels and their count/bytes. */
TFW_API tfw_status tfw_trace_info(const tfw_trace* t, uint64_t* payload_bytes, uint64_t* mover_launches,
                                  uint64_t* algorithmic_bytes);
TFW_API tfw_status tfw_trace_free(tfw_worker* w, tfw_trace* t);
/* Final handle table of a loaded trace (buffers are owned by the trace). */
TFW_API tfw_status tfw_trace_buffer_info(const tfw_trace* t, uint32_t handle, uint64_t* size, uint64_t* dev_ptr);

/* ---- introspection (tests, metrics) ------------------------------------- */
TFW_API tfw_status tfw_buffer_info(tfw_worker* w, uint32_t handle, uint64_t* size, uint64_t* dev_ptr);
/* Synchronous device->host read of a client-visible buffer range. */
TFW_API tfw_status tfw_buffer_read(tfw_worker* w, uint32_t handle, uint64_t off, void* dst, uint64_t n);
/* 64-bit order-sensitive digest of a whole buffer, computed on the GPU
 * (definition: tfw_digest64 in DESIGN.md; the oracle restates it on the CPU). */
TFW_API tfw_status tfw_buffer_digest(tfw_worker* w, uint32_t handle, uint64_t* digest);
TFW_API tfw_status tfw_get_stats(tfw_worker* w, tfw_stats* out);
/* Raw CUDA stream handle (cudaStream_t) of the vGPU execution stream. */
TFW_API void* tfw_exec_stream(tfw_worker* w);

/* ---- the mover kernel on its own (kernel-level tests and roofline) ------ */
/* Launch the byte mover over `n` host-side descriptors of device addresses on
 * the worker's exec stream; if ms != NULL the call synchronises and returns the
 * kernel's CUDA-event time. */
TFW_API tfw_status tfw_move_batch(tfw_worker* w, tfw_move_desc* descs, uint32_t n, float* ms);
/* Device scratch memory for kernel-level tests (not client-visible). */
TFW_API tfw_status tfw_dev_alloc(tfw_worker* w, uint64_t bytes, uint64_t* dev_ptr);
TFW_API tfw_status tfw_dev_free(tfw_worker* w, uint64_t dev_ptr);
TFW_API tfw_status tfw_dev_write(tfw_worker* w, uint64_t dev_ptr, const void* src, uint64_t n);
TFW_API tfw_status tfw_dev_read(tfw_worker* w, uint64_t dev_ptr, void* dst, uint64_t n);
/* tfw_buffer_digest over a raw device range (8-byte aligned base). */
TFW_API tfw_status tfw_dev_digest(tfw_worker* w, uint64_t dev_ptr, uint64_t bytes, uint64_t* digest);

#ifdef __cplusplus
}
#endif
#endif /* TFW_WORKER_H */
