/*
 * tfw_shm_ring.h -- shared-memory transport between a client process and the vGPU worker on the
 * same node: `./tensor-fusion-worker -n shmem -m tf_shm -M 1024` (internal/utils/compose.go:1311-1317)
 * and the client's connection URL "shmem+tf_shm+1024+1" (internal/webhook/v1/pod_webhook.go:584); the
 * file lives in the pod-shared /dev/shm (pkg/constants/constants.go:291, env.go:76-77).
 *
 * The reference's ring layout is unpublished (both halves are closed), so this is this repo's own:
 * one file = a 4 KiB header + two byte rings carrying the TFCS stream (include/tfw_wire.h),
 * client -> worker (3/4 of the space) and worker -> client (1/4).  Cursors are monotonic byte
 * counts; position = cursor % ring size; a ring is empty when head == tail.  The worker page-locks the
 * whole mapping (cudaHostRegister), so the GPU's copy engine reads H2D payloads straight out of
 * the ring the client wrote them into: one CPU copy end to end, no system call per frame.
 * The consumer cursor of the client -> worker ring (`c2w_tail`) therefore only advances once the
 * DMA that reads a span has completed.
 *
 * Session protocol: the worker creates / sizes the file, fills the header, sets `worker_ready`.
 * A client attaches by CAS `client_pid` 0 -> its pid and notes `session`; it streams frames and stores
 * its session number into `client_closed` when it is done writing; the worker drains, answers, stores
 * the number into `worker_closed`; the client detaches; the worker bumps `session` and clears
 * `client_pid` for the next client (cursors keep counting).
 */
#ifndef TFW_SHM_RING_H
#define TFW_SHM_RING_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFSR_MAGIC 0x52534654u /* 'TFSR' */
#define TFSR_VERSION 1u /* the "initVersion" field of the connection URL (pod_webhook.go:583) */
#define TFSR_HDR_BYTES 4096u
#define TFSR_MIN_BYTES (1u << 20)

typedef struct {
  uint32_t magic, version;
  uint64_t total_bytes;
  uint64_t c2w_off, c2w_size; /* client -> worker ring: offset from the start of the file, bytes (multiple of 64) */
  uint64_t w2c_off, w2c_size; /* worker -> client ring */
  uint32_t worker_pid;
  uint32_t worker_ready;      /* 1 once the mapping is sized, initialised and page-locked */
  uint32_t client_pid;        /* 0 = free; CAS to attach */
  uint32_t client_closed;     /* = session once the client of that session will write no more */
  uint32_t worker_closed;     /* = session once the worker has answered everything of that session (or failed) */
  uint32_t session;           /* number of the current (or next) client, from 1; bumped by the worker */
  uint32_t client_lock_session; /* = session when the attached client holds the liveness lock (see tfsr_client_lock) */
  uint32_t reserved0;
  uint8_t pad0[128 - 80];
  /* cursors, one cache line each: written by one side, read by the other */
  uint64_t c2w_head; uint8_t pad1[56]; /* client: bytes produced */
  uint64_t c2w_tail; uint8_t pad2[56]; /* worker: bytes released (their DMA has completed) */
  uint64_t w2c_head; uint8_t pad3[56]; /* worker: bytes produced */
  uint64_t w2c_tail; uint8_t pad4[56]; /* client: bytes consumed */
} tfsr_header;

#ifdef __cplusplus
static_assert(sizeof(tfsr_header) == 128 + 4 * 64, "tfsr_header layout");
static_assert(sizeof(tfsr_header) <= TFSR_HDR_BYTES, "tfsr_header must fit its page");
#endif

/* split of a file of `total` bytes; both rings are multiples of 64 bytes */
static inline void tfsr_layout(uint64_t total, uint64_t* c2w_off, uint64_t* c2w_size, uint64_t* w2c_off, uint64_t* w2c_size) {
  const uint64_t space = (total - TFSR_HDR_BYTES) & ~(uint64_t)4095;
  const uint64_t up = (space / 4 * 3) & ~(uint64_t)4095;
  *c2w_off = TFSR_HDR_BYTES;
  *c2w_size = up;
  *w2c_off = TFSR_HDR_BYTES + up;
  *w2c_size = space - up;
}

/* Liveness of the attached client.  PIDs mean nothing across the containers of a pod, a file lock does: the client
 * keeps an open-file-description write lock on byte 0 of the ring file for as long as it is attached -- the kernel
 * drops it when the process dies, however it dies -- and says so in client_lock_session.  The worker probes the
 * lock when a session has been silent for a while.  Any error reads as "alive". */
#if defined(__linux__)
#ifndef _GNU_SOURCE
#define _GNU_SOURCE 1
#endif
#include <fcntl.h>
#include <string.h>
#ifndef F_OFD_GETLK /* <fcntl.h> was included earlier without _GNU_SOURCE: asm-generic/fcntl.h values */
#define F_OFD_GETLK 36
#define F_OFD_SETLK 37
#endif
#ifdef F_OFD_SETLK
static inline int tfsr_client_lock(int fd, int lock) { /* 0 = done */
  struct flock fl;
  memset(&fl, 0, sizeof fl);
  fl.l_type = lock ? F_WRLCK : F_UNLCK;
  fl.l_whence = SEEK_SET;
  fl.l_start = 0;
  fl.l_len = 1;
  return fcntl(fd, F_OFD_SETLK, &fl) == 0 ? 0 : -1;
}
static inline int tfsr_client_alive(int fd) { /* 1 = a client holds the lock (or we cannot tell), 0 = nobody does */
  struct flock fl;
  memset(&fl, 0, sizeof fl);
  fl.l_type = F_WRLCK;
  fl.l_whence = SEEK_SET;
  fl.l_start = 0;
  fl.l_len = 1;
  if (fcntl(fd, F_OFD_GETLK, &fl) != 0) return 1;
  return fl.l_type != F_UNLCK;
}
#define TFSR_HAVE_LIVENESS 1
#endif
#endif

#ifdef __cplusplus
}
#endif
#endif /* TFW_SHM_RING_H */
