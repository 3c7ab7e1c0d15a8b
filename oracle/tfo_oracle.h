/*
 * tfo_oracle.h -- CPU oracle of the vGPU worker hot path.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (libtfw_b200.so,
 * libaccelerator_b200.so) never links, loads or calls anything under oracle/.
 *
 * Parity status
 *   - quota file + ERL arithmetic (shm_oracle.c, erl_oracle.c): PINNED by the
 *     reference's own golden vectors (pkg/hypervisor/worker/state/
 *     soft_limiter_shm_test.go, computing/quota_controller_test.go), see
 *     tests/test_oracle_golden.py.
 *   - command-stream replay (replay_oracle.c): "PARITY UNPINNED" at the
 *     reference boundary -- the reference worker is closed source and publishes
 *     neither a wire format nor golden buffers.  The invariant
 *     restated here is the north_star's: the path moves bytes, so a sequential
 *     malloc/memcpy/memset replay on the CPU defines the result.
 */
#ifndef TFO_ORACLE_H
#define TFO_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfo_session tfo_session;

/* Sequential replay of a TFCS stream on host memory.  Returns 0, or 7
 * (protocol error) for a malformed stream.  flags bit 2 (0x4) = do not zero
 * fresh allocations. */
int tfo_replay(const void* stream, size_t nbytes, uint64_t vram_limit, uint32_t flags, tfo_session** out);
size_t tfo_responses(const tfo_session* s, const uint8_t** p);
int tfo_buffer(const tfo_session* s, uint32_t handle, const uint8_t** p, uint64_t* size);
uint64_t tfo_stat(const tfo_session* s, int which); /* 0 frames, 1 payload bytes, 2 live buffers, 3 vram bytes, 4 errors */
void tfo_free(tfo_session* s);
void tfo_set_threads(int n);          /* persistent pool for copies / fills of 4 MiB and more (1 = off) */
void tfo_set_buffer_cache(int on);   /* freed buffers keep their pages for the next MALLOC of that size (baseline timing only) */

/* digest of a byte range: restatement of tfw_digest64 */
uint64_t tfo_digest(const void* p, uint64_t n);
/* xoshiro256** / splitmix64 payload stream */
void tfo_payload(uint64_t seed, uint32_t call_id, void* dst, uint64_t n);
uint64_t tfo_splitmix64_nth(uint64_t seed, uint32_t n);
void tfo_pattern(uint64_t seed, void* dst, uint64_t nbytes); /* tfw_vspace_fill_pattern restated */ /* n-th output (0-based) of splitmix64 seeded with `seed` */

#ifdef __cplusplus
}
#endif
#endif
