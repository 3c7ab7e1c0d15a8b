/* cuda_user_probe.c -- a CUDA driver-API application that ships its OWN kernels: loads a code image
 * (cubin, PTX or fatbin file given on the command line) with cuModuleLoadData, allocates page-locked host
 * memory with cuMemAllocHost, launches `saxpy_u32` and `vec_add_struct` (tools/user_kernels.cu) and checks
 * the results against the CPU.  Linked against "libcuda.so.1": with the stub directory first in
 * LD_LIBRARY_PATH every call travels to the vGPU worker (MODULE_LOAD / LAUNCH_USER / *_REF frames); with the
 * real driver it runs natively.  Prints one JSON line with a digest of both outputs: the two runs must agree. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
extern CUresult cuInit(unsigned);
extern CUresult cuDeviceGet(int*, int);
extern CUresult cuCtxCreate_v2(void**, unsigned, int);
extern CUresult cuCtxSynchronize(void);
extern CUresult cuMemAlloc_v2(CUdeviceptr*, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemAllocHost_v2(void**, size_t);
extern CUresult cuMemFreeHost(void*);
extern CUresult cuMemcpyHtoDAsync_v2(CUdeviceptr, const void*, size_t, void*);
extern CUresult cuMemcpyDtoHAsync_v2(void*, CUdeviceptr, size_t, void*);
extern CUresult cuMemcpyDtoH_v2(void*, CUdeviceptr, size_t);
extern CUresult cuModuleLoadData(void**, const void*);
extern CUresult cuModuleUnload(void*);
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

typedef struct { uint32_t n, bias; CUdeviceptr a, b, out; } VecArgs;

static uint64_t fnv(const uint32_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: cuda_user_probe <image file> [n]\n"); return 2; }
  const uint32_t n = argc > 2 ? (uint32_t)strtoul(argv[2], NULL, 10) : 1000003u;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  fseek(f, 0, SEEK_END);
  const long len = ftell(f);
  fseek(f, 0, SEEK_SET);
  char* image = malloc((size_t)len + 1);
  if (fread(image, 1, (size_t)len, f) != (size_t)len) return 2;
  image[len] = 0;
  fclose(f);

  int dev = -1;
  void *ctx, *mod, *saxpy, *vadd, *missing = NULL;
  CK(cuInit(0));
  CK(cuDeviceGet(&dev, 0));
  CK(cuCtxCreate_v2(&ctx, 0, dev));
  CK(cuModuleLoadData(&mod, image));
  CK(cuModuleGetFunction(&saxpy, mod, "saxpy_u32"));
  CK(cuModuleGetFunction(&vadd, mod, "vec_add_struct"));
  const CUresult nf = cuModuleGetFunction(&missing, mod, "no_such_kernel");
  void* bad = NULL;
  const CUresult bad_image = cuModuleLoadData(&bad, "this is not a code image at all");

  uint32_t *hx, *hy, *hout;  /* page-locked: copies are asynchronous, zero-copy through the stub */
  const size_t bytes = (size_t)n * 4;
  CK(cuMemAllocHost_v2((void**)&hx, bytes));
  CK(cuMemAllocHost_v2((void**)&hy, bytes));
  CK(cuMemAllocHost_v2((void**)&hout, bytes));
  uint32_t* want = malloc(bytes);
  for (uint32_t i = 0; i < n; ++i) { hx[i] = i * 2654435761u + 12345u; hy[i] = (i ^ 0x9E3779B9u) * 40503u; want[i] = 7u * hx[i] + hy[i]; }
  CUdeviceptr dx = 0, dy = 0, dout = 0;
  CK(cuMemAlloc_v2(&dx, bytes));
  CK(cuMemAlloc_v2(&dy, bytes));
  CK(cuMemAlloc_v2(&dout, bytes + 64));
  CK(cuMemcpyHtoDAsync_v2(dx, hx, bytes, NULL));
  CK(cuMemcpyHtoDAsync_v2(dy, hy, bytes, NULL));
  uint32_t a = 7u, nn = n;
  void* p1[] = {&dx, &dy, &a, &nn};
  CK(cuLaunchKernel(saxpy, 592, 1, 1, 256, 1, 1, 0, NULL, p1, NULL));
  VecArgs va = {n, 5u, dx, dy, dout + 64};  /* an interior pointer: offset arithmetic on the client */
  void* p2[] = {&va};
  CK(cuLaunchKernel(vadd, 296, 1, 1, 128, 1, 1, 0, NULL, p2, NULL));
  CK(cuMemcpyDtoHAsync_v2(hout, dout + 64, bytes, NULL));
  uint32_t* got_y = malloc(bytes);           /* pageable destination: a payload-carrying response */
  CK(cuMemcpyDtoH_v2(got_y, dy, bytes));
  CK(cuCtxSynchronize());
  int ok_saxpy = memcmp(got_y, want, bytes) == 0, ok_vadd = 1;
  for (uint32_t i = 0; i < n; ++i) if (hout[i] != hx[i] + want[i] + 5u) { ok_vadd = 0; break; }
  printf("{\"n\": %u, \"ok_saxpy\": %d, \"ok_vec_add_struct\": %d, \"digest_y\": \"%016llx\", \"digest_out\": \"%016llx\", \"not_found\": %d, \"bad_image\": %d}\n",
         n, ok_saxpy, ok_vadd, (unsigned long long)fnv(got_y, n), (unsigned long long)fnv(hout, n), nf, bad_image);
  CK(cuMemFree_v2(dx)); CK(cuMemFree_v2(dy)); CK(cuMemFree_v2(dout));
  CK(cuMemFreeHost(hx)); CK(cuMemFreeHost(hy)); CK(cuMemFreeHost(hout));
  CK(cuModuleUnload(mod));
  return ok_saxpy && ok_vadd ? 0 : 1;
}
