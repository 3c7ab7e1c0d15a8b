/*
 * tfw_wire.h -- "TFCS" forwarded-CUDA command-stream wire format, version 1.
 *
 * The reference ships the remote-vGPU worker as a closed binary
 * (reference: README.md:131, internal/utils/compose.go:1304-1325 only launch
 * `./tensor-fusion-worker -p 8000`); it publishes no wire format.  This header
 * is therefore the format *this* worker defines ("parity unpinned" at the
 * reference boundary).  Precedent followed: the mock
 * driver's handle<->pointer indirection
 * (provider/example/device_mock/driver_mock.c:320) -- clients name buffers by
 * opaque handles, never by device addresses.
 *
 * Stream layout: a sequence of frames.  Every frame starts with a 64-byte
 * little-endian header; MEMCPY_H2D frames are followed by `length` payload
 * bytes, zero-padded up to the next 16-byte boundary.  Because the header is
 * 64 bytes and every payload is padded to 16, *every payload starts 16-byte
 * aligned relative to the start of the stream* -- the property the staging
 * kernel's 16-byte vector loads rely on.
 *
 * Plain C; shared by the CUDA worker, the C oracle and the host bindings.
 */
#ifndef TFW_WIRE_H
#define TFW_WIRE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFCS_MAGIC 0x53434654u /* bytes 'T','F','C','S' on the wire */
#define TFCS_VERSION 1u
#define TFCS_HDR_BYTES 64u
#define TFCS_PAYLOAD_ALIGN 16u

/* Limits of one vGPU session (SURVEY.md 8d, C1 trace definition). */
#define TFCS_MAX_HANDLES 65536u
#define TFCS_MAX_BUFFER_BYTES (1ull << 40)

typedef enum {
  TFCS_OP_NOP = 0,
  TFCS_OP_MALLOC = 1,     /* h0 = new handle, length = bytes; contents are zero */
  TFCS_OP_FREE = 2,       /* h0 */
  TFCS_OP_MEMCPY_H2D = 3, /* h0 @ off0, length bytes of payload follow */
  TFCS_OP_MEMCPY_D2H = 4, /* h0 @ off0, length; answered by a RESP_D2H frame */
  TFCS_OP_MEMCPY_D2D = 5, /* dst h0 @ off0  <-  src h1 @ off1, length */
  TFCS_OP_MEMSET = 6,     /* h0 @ off0, length, arg0 & 0xff = fill byte */
  TFCS_OP_LAUNCH = 7,     /* arg0 = built-in kernel id, arg1 = grid, arg2 = block,
                             arg3 = the client's cost estimate (advisory: the worker charges
                             the limiter with a cost it computes itself from grid and block),
                             h0 @ off0 .. +length = the buffer range it works on,
                             off1 = kernel-specific scalar */
  TFCS_OP_SYNC = 8,       /* drain the vGPU stream; answered by RESP_SYNC */
  /* -- pinned client memory shared with the worker (same-node transports only).  A client that
   *    allocates page-locked host memory (cuMemAllocHost) gets it from an "arena": a tmpfs file
   *    next to the ring file, named <ring file>.a<id>, which the worker maps and page-locks too.
   *    Copies from / to arena memory carry no payload: the GPU's copy engine moves the bytes
   *    between the client's own pages and HBM, exactly as native CUDA does for pinned memory. */
  TFCS_OP_HOST_REGISTER = 9,    /* h0 = arena id (1..TFCS_MAX_ARENAS), length = bytes of the file */
  TFCS_OP_HOST_UNREGISTER = 10, /* h0 = arena id; drains the vGPU stream first */
  TFCS_OP_MEMCPY_H2D_REF = 11,  /* dst h0 @ off0  <-  arena h1 @ off1, length; no payload */
  TFCS_OP_MEMCPY_D2H_REF = 12,  /* arena h1 @ off1  <-  src h0 @ off0, length; answered by RESP_ACK
                                   only if flags & TFCS_F_ACK (otherwise the next SYNC covers it) */
  /* -- user modules: the client ships the code image, the worker loads it with the driver --- */
  TFCS_OP_MODULE_LOAD = 13,     /* h0 = module id (client-chosen, 1..TFCS_MAX_MODULES), length bytes of
                                   cubin / PTX / fatbin follow as payload */
  TFCS_OP_MODULE_UNLOAD = 14,   /* h0 = module id */
  TFCS_OP_MODULE_GET_FUNCTION = 15, /* h0 = module id, h1 = function id (client-chosen), payload = the
                                   kernel's name (length bytes, no NUL); answered by RESP_FUNCTION */
  TFCS_OP_LAUNCH_USER = 16,     /* h1 = function id; payload = tfcs_launch_params + parameter block.
                                   Every 8-byte aligned word of the block that carries the client
                                   stub's device-pointer tag (TFCS_PTR_TAG) and names a live buffer is
                                   replaced by that buffer's real address on the worker. */
  /* -- a TCP client on the worker's own node (loopback) proposes to move the session onto shared-memory rings:
   *    first frame of the connection; payload = name of a ring file the client created in the shared-memory
   *    directory (no '/'), off0 = its size in bytes.  RESP_ACK: the worker has mapped, page-locked and initialised
   *    it -- both sides continue on the rings (include/tfw_shm_ring.h), the socket stays open as the session's
   *    lifeline.  RESP_ERROR (no shared /dev/shm, e.g. another pod): the session stays on the socket. */
  TFCS_OP_UPGRADE_SHM = 17,
  /* worker -> client */
  TFCS_OP_RESP_D2H = 0x84,  /* call_id echoes the request, payload follows */
  TFCS_OP_RESP_SYNC = 0x88, /* arg0 = status (0 ok) */
  TFCS_OP_RESP_ACK = 0x8C,  /* completion of a D2H_REF that asked for it */
  TFCS_OP_RESP_FUNCTION = 0x8F, /* payload = parameter layout: arg0 = count, then count x {u32 offset, u32 size};
                                   arg1 = bytes of the parameter block */
  TFCS_OP_RESP_ERROR = 0xFF /* arg0 = tfw_status, call_id = offending call */
} tfcs_opcode;

#define TFCS_F_ACK 0x1u       /* tfcs_frame_hdr.flags: answer with RESP_ACK when the copy has completed */
#define TFCS_MAX_ARENAS 15u
#define TFCS_MAX_MODULES 4096u
#define TFCS_MAX_FUNCTIONS 65536u
#define TFCS_MAX_PARAM_BYTES 4096u /* CUDA's own limit for a kernel parameter block (32764 on sm_100 with
                                      large-parameter kernels; the stub accepts the classic 4 KiB) */

/* Device pointers handed to a remote-mode application: bit 62 set, bits 40..61 the buffer handle,
 * bits 0..39 the byte offset inside the buffer (so pointer arithmetic within a buffer works on the
 * client and the worker keeps addressing by handle). */
#define TFCS_PTR_TAG (1ull << 62)
#define TFCS_PTR_HANDLE(p) ((uint32_t)(((p) >> 40) & 0x3FFFFFu))
#define TFCS_PTR_OFFSET(p) ((p) & ((1ull << 40) - 1))
#define TFCS_PTR_IS_TAGGED(p) ((((p) >> 62) & 3u) == 1u)

/* First bytes of a LAUNCH_USER payload; `param_bytes` bytes of parameter block follow. */
typedef struct {
  uint32_t grid[3];
  uint32_t block[3];
  uint32_t shared_bytes;
  uint32_t param_bytes;
} tfcs_launch_params;

/* Built-in kernel registry for TFCS_OP_LAUNCH.  A real client ships cubins;
 * the synthetic traces of SURVEY.md 8d only need these. */
typedef enum {
  TFCS_KERNEL_NOOP = 0,    /* empty kernel, grid x block as requested */
  TFCS_KERNEL_SPIN = 1,    /* busy-wait off1 nanoseconds per CTA (limiter load) */
  TFCS_KERNEL_ADD_U8 = 2,  /* buf[i] += (off1 & 0xff) (mod 256) over the range */
  TFCS_KERNEL_XOR_IDX = 3  /* buf[i] ^= (uint8)(i*off1 >> 3), i relative to range */
} tfcs_kernel_id;

typedef struct {
  uint32_t magic;   /* TFCS_MAGIC */
  uint16_t version; /* TFCS_VERSION */
  uint16_t opcode;  /* tfcs_opcode */
  uint32_t call_id; /* monotonically increasing per connection */
  uint32_t flags;   /* reserved, 0 */
  uint32_t h0;      /* destination / subject buffer handle */
  uint32_t h1;      /* source buffer handle (D2D) */
  uint64_t off0;    /* byte offset inside h0 */
  uint64_t off1;    /* byte offset inside h1, or kernel scalar */
  uint64_t length;  /* bytes */
  uint32_t arg0, arg1, arg2, arg3;
} tfcs_frame_hdr;

#if defined(__cplusplus)
static_assert(sizeof(tfcs_frame_hdr) == TFCS_HDR_BYTES, "TFCS header must be 64 bytes");
#else
_Static_assert(sizeof(tfcs_frame_hdr) == TFCS_HDR_BYTES, "TFCS header must be 64 bytes");
#endif

static inline uint64_t tfcs_pad16(uint64_t n) { return (n + 15u) & ~(uint64_t)15u; }

/* Bytes a frame occupies on the wire (header + padded payload). */
/* Opcodes whose header is followed by `length` payload bytes (zero-padded to 16). */
static inline int tfcs_has_payload(uint32_t opcode) {
  return opcode == TFCS_OP_MEMCPY_H2D || opcode == TFCS_OP_RESP_D2H || opcode == TFCS_OP_MODULE_LOAD ||
         opcode == TFCS_OP_MODULE_GET_FUNCTION || opcode == TFCS_OP_LAUNCH_USER || opcode == TFCS_OP_RESP_FUNCTION ||
         opcode == TFCS_OP_UPGRADE_SHM;
}

static inline uint64_t tfcs_frame_bytes(const tfcs_frame_hdr* h) {
  if (tfcs_has_payload(h->opcode)) return TFCS_HDR_BYTES + tfcs_pad16(h->length);
  return TFCS_HDR_BYTES;
}

#ifdef __cplusplus
}
#endif
#endif /* TFW_WIRE_H */
