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
/* Page-locked host memory shared with the worker ("arena", TFCS_OP_HOST_REGISTER): what a CUDA application gets
 * from cuMemAllocHost in remote mode.  Copies whose host side lies in such memory carry no payload -- the GPU's
 * copy engine moves the bytes between these very pages and HBM, as native CUDA does for pinned memory.  Same-node
 * (shmem) connections only: 3 (not supported) over TCP, 4 when /dev/shm or the arena table (15) is exhausted. */
TFC_API int tfc_host_alloc(tfc_conn* c, uint64_t bytes, void** out);
TFC_API int tfc_host_free(tfc_conn* c, void* p);
/* Asynchronous D2H into arena memory: complete after the next tfc_sync (1 if dst is not arena memory). */
TFC_API int tfc_memcpy_d2h_async(tfc_conn* c, void* dst, uint32_t src, uint64_t off, uint64_t n);
/* User modules (TFCS_OP_MODULE_LOAD / MODULE_GET_FUNCTION / LAUNCH_USER): the worker loads the image with the
 * driver; get_function blocks for the kernel's parameter layout (offsets/sizes, up to `cap` entries are stored). */
TFC_API int tfc_module_load(tfc_conn* c, const void* image, uint64_t bytes, uint32_t* module);
TFC_API int tfc_module_unload(tfc_conn* c, uint32_t module);
TFC_API int tfc_module_get_function(tfc_conn* c, uint32_t module, const char* name, uint32_t* function, uint32_t* nparams,
                                    uint32_t* offsets, uint32_t* sizes, uint32_t cap, uint32_t* param_bytes);
/* params: the kernel's parameter block (param_bytes as reported by get_function); device pointers inside it are
 * the tagged values of include/tfw_wire.h (TFCS_PTR_TAG | handle << 40 | offset). */
TFC_API int tfc_launch_user(tfc_conn* c, uint32_t function, const uint32_t grid[3], const uint32_t block[3],
                            uint32_t shared_bytes, const void* params, uint32_t param_bytes, uint32_t cost_tokens);
TFC_API int tfc_sync(tfc_conn* c);                 /* blocking; returns the first error code seen since the last sync (0 = none) */
TFC_API int tfc_last_error_code(const tfc_conn* c);  /* tfw_status of the most recent RESP_ERROR */
TFC_API uint32_t tfc_last_error_call(const tfc_conn* c);

#ifdef __cplusplus
}
#endif
#endif
