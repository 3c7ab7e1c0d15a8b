/*
 * tfc_client.h -- minimal client of the TFCS transport.
 *
 * The reference's client is a closed libcuda shim that dials the worker found in
 * TensorFusionConnection.status.connectionURL =
 *     "native+<workerPodIP>+<port>+<podName>-<resourceVersion>"
 * (internal/controller/tensorfusionconnection_controller.go:136-138).  This library speaks the
 * same URL and this repo's wire format (include/tfw_wire.h); it is what a CUDA-interposing shim
 * would sit on.  Host only -- no CUDA on the client side.  Calls are stream-ordered on the
 * worker; tfc_memcpy_d2h and tfc_sync wait for their response, everything else is fire-and-forget
 * and errors surface through tfc_sync / tfc_last_error_code.
 */
#ifndef TFC_CLIENT_H
#define TFC_CLIENT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#define TFC_API __attribute__((visibility("default")))

typedef struct tfc_conn tfc_conn;

/* url: "native+<ip>+<port>+<anything>", "<ip>:<port>", or "shmem+<name>+<MiB>+<initVersion>" (same-node worker started
 * with `-n shmem -m <name> -M <MiB>`, rings in /dev/shm/<name>, include/tfw_shm_ring.h).  Returns 0 on success. */
TFC_API int tfc_connect(const char* url, tfc_conn** out);
TFC_API void tfc_close(tfc_conn* c);
TFC_API int tfc_malloc(tfc_conn* c, uint64_t bytes, uint32_t* handle);
TFC_API int tfc_free(tfc_conn* c, uint32_t handle);
TFC_API int tfc_memcpy_h2d(tfc_conn* c, uint32_t dst, uint64_t off, const void* src, uint64_t n);
TFC_API int tfc_memcpy_d2h(tfc_conn* c, void* dst, uint32_t src, uint64_t off, uint64_t n); /* blocking */
TFC_API int tfc_memcpy_d2d(tfc_conn* c, uint32_t dst, uint64_t doff, uint32_t src, uint64_t soff, uint64_t n);
TFC_API int tfc_memset(tfc_conn* c, uint32_t dst, uint64_t off, int value, uint64_t n);
TFC_API int tfc_launch(tfc_conn* c, uint32_t kernel_id, uint32_t grid, uint32_t block, uint32_t handle, uint64_t off,
                       uint64_t n, uint64_t scalar, uint32_t cost_tokens);
TFC_API int tfc_sync(tfc_conn* c);                 /* blocking; returns the first error code seen since the last sync (0 = none) */
TFC_API int tfc_last_error_code(const tfc_conn* c);  /* tfw_status of the most recent RESP_ERROR */
TFC_API uint32_t tfc_last_error_call(const tfc_conn* c);

#ifdef __cplusplus
}
#endif
#endif
