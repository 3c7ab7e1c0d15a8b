/* cuda_remote_probe.c -- a small CUDA driver-API application, linked against "libcuda.so.1" like any other.
 * With the stub directory first in LD_LIBRARY_PATH that name resolves to libcuda_remote.so and every call
 * below travels to the vGPU worker.  Prints one JSON line; exit code 0 = all results as expected. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
extern CUresult cuInit(unsigned);
extern CUresult cuDriverGetVersion(int*);
extern CUresult cuDeviceGetCount(int*);
extern CUresult cuDeviceGet(int*, int);
extern CUresult cuDeviceGetName(char*, int, int);
extern CUresult cuDeviceGetAttribute(int*, int, int);
extern CUresult cuDeviceTotalMem_v2(size_t*, int);
extern CUresult cuCtxCreate_v2(void**, unsigned, int);
extern CUresult cuCtxSynchronize(void);
extern CUresult cuMemAlloc_v2(CUdeviceptr*, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemGetInfo_v2(size_t*, size_t*);
extern CUresult cuMemcpyHtoD_v2(CUdeviceptr, const void*, size_t);
extern CUresult cuMemcpyDtoH_v2(void*, CUdeviceptr, size_t);
extern CUresult cuMemcpyDtoD_v2(CUdeviceptr, CUdeviceptr, size_t);
extern CUresult cuMemsetD8_v2(CUdeviceptr, unsigned char, size_t);
extern CUresult cuModuleLoadData(void**, const void*);
extern CUresult cuModuleGetFunction(void**, void*, const char*);
extern CUresult cuLaunchKernel(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**);
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

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? (size_t)strtoull(argv[1], NULL, 10) : 5000003;
  int ver = 0, count = 0, dev = -1, sms = 0, cc = 0;
  char name[128];
  size_t total = 0, free_b = 0, total2 = 0;
  void *ctx, *mod, *add, *xorf, *missing = NULL;
  CK(cuInit(0));
  CK(cuDriverGetVersion(&ver));
  CK(cuDeviceGetCount(&count));
  CK(cuDeviceGet(&dev, 0));
  CK(cuDeviceGetName(name, sizeof name, dev));
  CK(cuDeviceGetAttribute(&sms, 16, dev));
  CK(cuDeviceGetAttribute(&cc, 75, dev));
  CK(cuDeviceTotalMem_v2(&total, dev));
  CK(cuCtxCreate_v2(&ctx, 0, dev));
  CUdeviceptr a = 0, b = 0, huge = 0;
  CK(cuMemAlloc_v2(&a, n));
  CK(cuMemAlloc_v2(&b, n));
  const CUresult oom = cuMemAlloc_v2(&huge, (size_t)1 << 39);  /* over any quota of the test */
  CK(cuMemGetInfo_v2(&free_b, &total2));
  uint8_t* src = malloc(n);
  uint8_t* want_a = malloc(n);
  uint8_t* want_b = calloc(n, 1);
  uint8_t* got = malloc(n);
  uint64_t s = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; src[i] = (uint8_t)(s >> 32); }
  CK(cuMemcpyHtoD_v2(a, src, n));
  CK(cuMemcpyHtoD_v2(b + 7, src + 100, 1000));                 /* pointer arithmetic on a device pointer */
  memcpy(want_a, src, n);
  memcpy(want_b + 7, src + 100, 1000);
  CK(cuModuleLoadData(&mod, NULL));
  CK(cuModuleGetFunction(&add, mod, "tfw_add_u8"));
  CK(cuModuleGetFunction(&xorf, mod, "tfw_xor_idx"));
  const CUresult nf = cuModuleGetFunction(&missing, mod, "my_own_kernel");
  CUdeviceptr p = a + 5;
  uint64_t cnt = n - 5, scalar = 9;
  void* params[3] = {&p, &cnt, &scalar};
  CK(cuLaunchKernel(add, 64, 1, 1, 256, 1, 1, 0, NULL, params, NULL));
  for (size_t i = 5; i < n; ++i) want_a[i] = (uint8_t)(want_a[i] + 9);
  p = b + 2000; cnt = 4096; scalar = 24;
  CK(cuLaunchKernel(xorf, 16, 1, 1, 128, 1, 1, 0, NULL, params, NULL));
  for (size_t i = 0; i < 4096; ++i) want_b[2000 + i] ^= (uint8_t)((i * 24) >> 3);
  CK(cuMemcpyDtoD_v2(b + 100000, a + 1, 50000));
  memcpy(want_b + 100000, want_a + 1, 50000);
  CK(cuMemsetD8_v2(b + 300000, 0xEE, 333));
  memset(want_b + 300000, 0xEE, 333);
  CK(cuCtxSynchronize());
  CK(cuMemcpyDtoH_v2(got, a, n));
  const int ok_a = memcmp(got, want_a, n) == 0;
  CK(cuMemcpyDtoH_v2(got, b, n));
  const int ok_b = memcmp(got, want_b, n) == 0;
  const CUresult bad_ptr = cuMemcpyDtoH_v2(got, 0x1234, 16);    /* not a pointer this driver handed out */
  const CUresult past_end = cuMemcpyHtoD_v2(a + n - 4, src, 8);
  CK(cuMemFree_v2(a));
  CK(cuMemFree_v2(b));
  const CUresult dbl = cuMemFree_v2(a);
  CK(cuCtxSynchronize());
  printf("{\"version\": %d, \"count\": %d, \"name\": \"%s\", \"sms\": %d, \"cc_major\": %d, \"total\": %zu, \"free_after_alloc\": %zu, "
         "\"oom\": %d, \"not_found\": %d, \"ok_a\": %d, \"ok_b\": %d, \"bad_ptr\": %d, \"past_end\": %d, \"double_free\": %d}\n",
         ver, count, name, sms, cc, total, free_b, oom, nf, ok_a, ok_b, bad_ptr, past_end, dbl);
  return ok_a && ok_b ? 0 : 11;
}
