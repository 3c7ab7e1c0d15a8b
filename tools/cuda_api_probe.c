/* cuda_api_probe.c -- the parts of the driver API that libcuda_remote.so builds on the client out of wire operations
 * the worker already has: 16- and 32-bit pattern memsets (one seed block + doubling D2D copies), unified-addressing
 * cuMemcpy, events on the vGPU's one ordered stream, pointer attributes, host functions.  Linked against "libcuda.so.1"
 * like any application; prints one JSON line, exit code 0 = everything as expected. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
extern CUresult cuInit(unsigned);
extern CUresult cuDeviceGet(int*, int);
extern CUresult cuDeviceComputeCapability(int*, int*, int);
extern CUresult cuCtxCreate_v2(void**, unsigned, int);
extern CUresult cuCtxGetApiVersion(void*, unsigned*);
extern CUresult cuCtxSynchronize(void);
extern CUresult cuMemAlloc_v2(CUdeviceptr*, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemAllocHost_v2(void**, size_t);
extern CUresult cuMemFreeHost(void*);
extern CUresult cuMemsetD8_v2(CUdeviceptr, unsigned char, size_t);
extern CUresult cuMemsetD16_v2(CUdeviceptr, unsigned short, size_t);
extern CUresult cuMemsetD32_v2(CUdeviceptr, unsigned, size_t);
extern CUresult cuMemcpy(CUdeviceptr, CUdeviceptr, size_t);
extern CUresult cuMemcpyAsync(CUdeviceptr, CUdeviceptr, size_t, void*);
extern CUresult cuMemcpyDtoH_v2(void*, CUdeviceptr, size_t);
extern CUresult cuStreamCreate(void**, unsigned);
extern CUresult cuStreamWaitEvent(void*, void*, unsigned);
extern CUresult cuEventCreate(void**, unsigned);
extern CUresult cuEventRecord(void*, void*);
extern CUresult cuEventQuery(void*);
extern CUresult cuEventSynchronize(void*);
extern CUresult cuEventElapsedTime(float*, void*, void*);
extern CUresult cuEventDestroy_v2(void*);
extern CUresult cuPointerGetAttribute(void*, int, CUdeviceptr);
extern CUresult cuLaunchHostFunc(void*, void (*)(void*), void*);
extern CUresult cuGetErrorName(CUresult, const char**);

#define CK(call)                                                        \
  do {                                                                  \
    CUresult r_ = (call);                                               \
    if (r_ != 0) {                                                      \
      const char* n_ = "?";                                             \
      cuGetErrorName(r_, &n_);                                          \
      fprintf(stderr, "%s -> %d (%s)\n", #call, r_, n_);                \
      return 10;                                                        \
    }                                                                   \
  } while (0)

static void host_fn(void* p) { *(int*)p += 1; }

int main(int argc, char** argv) {
  const size_t count = argc > 1 ? (size_t)strtoull(argv[1], NULL, 10) : 300007;  /* 32-bit words: 1.2 MB, not a power of two */
  int dev = -1, major = 0, minor = 0, called = 0;
  unsigned api = 0;
  void *ctx, *stream, *ev0, *ev1;
  CK(cuInit(0));
  CK(cuDeviceGet(&dev, 0));
  CK(cuDeviceComputeCapability(&major, &minor, dev));
  CK(cuCtxCreate_v2(&ctx, 0, dev));
  CK(cuCtxGetApiVersion(ctx, &api));
  CK(cuStreamCreate(&stream, 0));
  CUdeviceptr a = 0, b = 0;
  const size_t bytes = count * 4;
  CK(cuMemAlloc_v2(&a, bytes));
  CK(cuMemAlloc_v2(&b, bytes));
  uint32_t* host = malloc(bytes);
  uint8_t* hb = (uint8_t*)host;
  if (!host) return 11;

  /* 32-bit pattern whose bytes differ (1.0f) */
  CK(cuMemsetD32_v2(a, 0x3f800000u, count));
  CK(cuMemcpyDtoH_v2(host, a, bytes));
  int ok_d32 = 1;
  for (size_t i = 0; i < count; ++i) ok_d32 &= host[i] == 0x3f800000u;
  /* byte-replicated value: the plain byte fill */
  CK(cuMemsetD32_v2(a, 0x7f7f7f7fu, count));
  CK(cuMemcpyDtoH_v2(host, a, bytes));
  int ok_d32_bytes = 1;
  for (size_t i = 0; i < count; ++i) ok_d32_bytes &= host[i] == 0x7f7f7f7fu;
  /* 16-bit pattern in the middle of a buffer: the bytes around it stay */
  const size_t n16 = 70001;
  CK(cuMemsetD8_v2(b, 0x11, bytes));
  CK(cuMemsetD16_v2(b + 6, 0xBEEF, n16));
  CK(cuMemcpyDtoH_v2(host, b, bytes));
  int ok_d16 = hb[0] == 0x11 && hb[5] == 0x11 && hb[6 + 2 * n16] == 0x11 && hb[bytes - 1] == 0x11;
  for (size_t i = 0; i < n16; ++i) ok_d16 &= hb[6 + 2 * i] == 0xEF && hb[7 + 2 * i] == 0xBE;
  const CUresult misaligned = cuMemsetD32_v2(a + 2, 0x01020304u, 4);   /* INVALID_VALUE like the driver */
  const CUresult past_end = cuMemsetD16_v2(b, 0x0102, bytes);           /* 2 x bytes: beyond the allocation */

  /* unified addressing: one entry point, the pointers say which way */
  for (size_t i = 0; i < count; ++i) host[i] = (uint32_t)(i * 2654435761u);
  CK(cuMemcpy(a, (CUdeviceptr)(uintptr_t)host, bytes));                  /* host -> device */
  CK(cuMemcpy(b, a, bytes));                                             /* device -> device */
  void* pinned = NULL;
  CK(cuMemAllocHost_v2(&pinned, bytes));
  memset(pinned, 0, bytes);
  CK(cuEventCreate(&ev0, 0));
  CK(cuEventCreate(&ev1, 0));
  CK(cuEventRecord(ev0, stream));
  CK(cuMemcpyAsync((CUdeviceptr)(uintptr_t)pinned, b, bytes, stream));   /* device -> page-locked host, asynchronous */
  CK(cuEventRecord(ev1, stream));
  CK(cuStreamWaitEvent(stream, ev1, 0));
  CK(cuEventSynchronize(ev1));                                           /* covers the copy */
  const int ok_memcpy = memcmp(pinned, host, bytes) == 0;
  const CUresult query_done = cuEventQuery(ev0);
  float ms = -1.0f;
  const CUresult elapsed = cuEventElapsedTime(&ms, ev0, ev1);            /* NOT_SUPPORTED: no GPU timestamps on the wire */
  CK(cuEventDestroy_v2(ev0));
  const CUresult stale = cuEventSynchronize(ev0);                        /* INVALID_HANDLE */
  CK(cuEventDestroy_v2(ev1));

  unsigned mtype = 0;
  CUdeviceptr base = 0;
  size_t range = 0;
  CK(cuPointerGetAttribute(&mtype, 2, a + 100));
  CK(cuPointerGetAttribute(&base, 11, a + 100));
  CK(cuPointerGetAttribute(&range, 12, a + 100));
  const CUresult host_ptr = cuPointerGetAttribute(&mtype, 2, (CUdeviceptr)(uintptr_t)host);
  CK(cuLaunchHostFunc(stream, host_fn, &called));
  CK(cuMemFreeHost(pinned));
  CK(cuMemFree_v2(a));
  CK(cuMemFree_v2(b));
  CK(cuCtxSynchronize());
  printf("{\"cc\": [%d, %d], \"api\": %u, \"ok_d32\": %d, \"ok_d32_bytes\": %d, \"ok_d16\": %d, \"misaligned\": %d, \"past_end\": %d, \"ok_memcpy\": %d, "
         "\"query_done\": %d, \"elapsed\": %d, \"stale\": %d, \"mtype\": %u, \"base_ok\": %d, \"range\": %zu, \"host_ptr\": %d, \"called\": %d}\n",
         major, minor, api, ok_d32, ok_d32_bytes, ok_d16, misaligned, past_end, ok_memcpy, query_done, elapsed, stale, mtype, base == a, range, host_ptr, called);
  free(host);
  return (ok_d32 && ok_d32_bytes && ok_d16 && ok_memcpy) ? 0 : 12;
}
