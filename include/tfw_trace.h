/*
 * tfw_trace.h -- C-ABI of the synthetic TFCS trace generators (host only).
 *
 * The reference has no recorded command traces ("parity
 * unpinned" for the data path); these generators define the synthetic traces
 * of SURVEY.md 8d so that the GPU worker, the CPU oracle and the native-CUDA
 * comparator all consume byte-identical input.
 *
 * Payload bytes of call `i` are the little-endian output words of
 * xoshiro256** seeded with four splitmix64 outputs of (seed + i).
 */
#ifndef TFW_TRACE_H
#define TFW_TRACE_H

#include <stddef.h>
#include <stdint.h>

#include "tfw_worker.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TFW_TRACE_SEED_C1 0x7F5EEDull /* SURVEY.md 8d */

typedef struct {
  uint64_t seed;
  uint32_t ncalls;            /* frames to emit (a trailing SYNC is added) */
  uint32_t max_live;          /* live-buffer cap (64) */
  uint64_t max_buffer_bytes;  /* largest MALLOC (<= 64 MiB) */
  uint64_t max_payload_bytes; /* largest copy (4 MiB) */
  uint32_t unaligned_percent; /* copies forced off 16-byte alignment (25) */
  uint32_t error_permille;    /* deliberately invalid frames (bad handle / out of range) */
  uint32_t launch_cost;       /* tokens charged per LAUNCH frame (arg3); 0 = limiter not exercised */
  uint32_t reserved;
} tfw_trace_c1_params;

TFW_API void tfw_trace_c1_defaults(tfw_trace_c1_params* p);
/* Generate the mixed trace (40% H2D, 10% D2H, 10% D2D, 5% memset, 5% malloc,
 * 5% free, 25% launch).  out == NULL or cap too small: only *nbytes is set. */
TFW_API tfw_status tfw_trace_gen_c1(const tfw_trace_c1_params* p, void* out, size_t cap, size_t* nbytes);
/* Bulk stream of SURVEY.md 8d C2: `nbuf` MALLOCs of `bytes_each`, then `ncopies`
 * H2D frames of `bytes_each` round-robin into them, one noop LAUNCH, one SYNC. */
TFW_API tfw_status tfw_trace_gen_bulk(uint64_t seed, uint32_t nbuf, uint32_t ncopies, uint64_t bytes_each,
                                      uint32_t nthreads, void* out, size_t cap, size_t* nbytes);
/* Latency leg: `ncalls` x { H2D of `bytes_each` ; noop LAUNCH } into one buffer. */
TFW_API tfw_status tfw_trace_gen_small(uint64_t seed, uint32_t ncalls, uint64_t bytes_each, void* out, size_t cap,
                                       size_t* nbytes);
/* Native-CUDA comparator: the same call stream issued directly to
 * the CUDA runtime from one host thread.  One warm-up pass, then `passes` timed
 * passes (host wall-clock, stream synchronised on both sides). */
TFW_API tfw_status tfw_native_replay(int device, const void* stream, size_t nbytes, uint32_t passes,
                                     double* seconds_per_pass, uint64_t* payload_bytes, uint64_t* calls);
/* Bulk-copy comparator for the process-boundary legs of the bench: `ncopies` x cudaMemcpyAsync of `each`
 * bytes between `nsrc` host buffers (pinned != 0: page-locked, else pageable; touched beforehand) and `nbuf`
 * device buffers on one stream + one synchronize.  direction 0 = host -> device, 1 = device -> host. */
TFW_API tfw_status tfw_native_copy(int device, int direction, int pinned, uint64_t each, uint32_t nsrc, uint32_t nbuf,
                                   uint32_t ncopies, uint32_t passes, double* seconds_per_pass);
/* The payload generator on its own. */
TFW_API void tfw_trace_payload(uint64_t seed, uint32_t call_id, void* dst, uint64_t nbytes);

#ifdef __cplusplus
}
#endif
#endif /* TFW_TRACE_H */
