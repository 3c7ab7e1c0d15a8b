/* mock_cuda.c -- a stand-in libcuda.so.1 for the CPU tests of libcuda_limiter.so.
 * Exports the handful of driver entry points the hook interposes or calls, with the driver's
 * signatures; launches and allocations only count.  Test infrastructure, never shipped. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))
typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef struct { unsigned char bytes[16]; } CUuuid;

static uint64_t g_launches, g_launches_ptsz, g_allocs, g_frees, g_live_bytes;
static unsigned long long g_next = 0x7f0000000000ull;
static uint64_t g_sizes[4096];

EXPORT void mock_cuda_counts(uint64_t out[5]) {
  out[0] = g_launches; out[1] = g_launches_ptsz; out[2] = g_allocs; out[3] = g_frees; out[4] = g_live_bytes;
}
EXPORT CUresult cuInit(unsigned flags) { (void)flags; return 0; }
EXPORT CUresult cuCtxGetDevice(int* d) { *d = 0; return 0; }
EXPORT CUresult cuDeviceGetUuid_v2(CUuuid* u, int d) {
  (void)d;
  for (int i = 0; i < 16; ++i) u->bytes[i] = (unsigned char)(0x10 + i);  /* GPU-10111213-1415-1617-1819-1a1b1c1d1e1f */
  return 0;
}
EXPORT CUresult cuLaunchKernel(void* f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned smem,
                               void* st, void** params, void** extra) {
  (void)f; (void)gx; (void)gy; (void)gz; (void)bx; (void)by; (void)bz; (void)smem; (void)st; (void)params; (void)extra;
  __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
  return 0;
}
EXPORT CUresult cuLaunchKernel_ptsz(void* f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                    unsigned smem, void* st, void** params, void** extra) {
  (void)f; (void)gx; (void)gy; (void)gz; (void)bx; (void)by; (void)bz; (void)smem; (void)st; (void)params; (void)extra;
  __atomic_add_fetch(&g_launches_ptsz, 1, __ATOMIC_RELAXED);
  return 0;
}
EXPORT CUresult cuMemAlloc_v2(CUdeviceptr* p, size_t n) {
  if (g_allocs >= 4096) return 2;
  g_sizes[g_allocs] = n;
  *p = g_next + ((unsigned long long)g_allocs << 32);
  g_allocs++;
  g_live_bytes += n;
  return 0;
}
EXPORT CUresult cuMemFree_v2(CUdeviceptr p) {
  const unsigned long long i = (p - g_next) >> 32;
  if (i >= g_allocs) return 1;
  g_live_bytes -= g_sizes[i];
  g_sizes[i] = 0;
  g_frees++;
  return 0;
}
EXPORT CUresult cuMemGetInfo_v2(size_t* f, size_t* t) {
  *t = 180ull << 30;
  *f = (180ull << 30) - g_live_bytes;
  return 0;
}
EXPORT CUresult cuDeviceTotalMem_v2(size_t* b, int d) { (void)d; *b = 180ull << 30; return 0; }

/* ---- graphs: one fixed graph = 3 kernel nodes (4x128, 2x64, 1x32 threads), a memcpy node and a child graph with
 * one kernel node (8x256): blocks x warps = 16 + 4 + 1 + 64 = 85 ---- */
typedef struct { int type; unsigned grid, block; void* child; } mock_node;
typedef struct { mock_node* nodes; size_t n; } mock_graph;
static mock_node g_child_nodes[] = {{0, 8, 256, NULL}};
static mock_graph g_child = {g_child_nodes, 1};
static mock_node g_nodes[] = {{0, 4, 128, NULL}, {0, 2, 64, NULL}, {1, 0, 0, NULL}, {0, 1, 32, NULL}, {4, 0, 0, &g_child}};
static mock_graph g_graph = {g_nodes, 5};
static uint64_t g_graph_launches, g_graph_launches_ptsz, g_graph_destroys;
EXPORT void* mock_cuda_graph(void) { return &g_graph; }
EXPORT void mock_cuda_graph_counts(uint64_t out[3]) { out[0] = g_graph_launches; out[1] = g_graph_launches_ptsz; out[2] = g_graph_destroys; }
EXPORT CUresult cuGraphGetNodes(void* g, void** nodes, size_t* n) {
  mock_graph* mg = (mock_graph*)g;
  if (!nodes) { *n = mg->n; return 0; }
  size_t k = *n < mg->n ? *n : mg->n;
  for (size_t i = 0; i < k; ++i) nodes[i] = &mg->nodes[i];
  *n = k;
  return 0;
}
EXPORT CUresult cuGraphNodeGetType(void* node, int* type) { *type = ((mock_node*)node)->type; return 0; }
typedef struct { void* func; unsigned gx, gy, gz, bx, by, bz, smem; void** kp; void** extra; void* kern; void* ctx; } mock_kparams;
EXPORT CUresult cuGraphKernelNodeGetParams_v2(void* node, mock_kparams* p) {
  mock_node* mn = (mock_node*)node;
  if (mn->type != 0) return 1;
  memset(p, 0, sizeof *p);
  p->gx = mn->grid; p->gy = p->gz = 1; p->bx = mn->block; p->by = p->bz = 1;
  return 0;
}
EXPORT CUresult cuGraphChildGraphNodeGetGraph(void* node, void** g) {
  mock_node* mn = (mock_node*)node;
  if (mn->type != 4) return 1;
  *g = mn->child;
  return 0;
}
EXPORT CUresult cuGraphInstantiateWithFlags(void** exec, void* g, unsigned long long flags) { (void)flags; *exec = (char*)g + 1; return 0; }
EXPORT CUresult cuGraphLaunch(void* exec, void* st) { (void)exec; (void)st; __atomic_add_fetch(&g_graph_launches, 1, __ATOMIC_RELAXED); return 0; }
EXPORT CUresult cuGraphLaunch_ptsz(void* exec, void* st) { (void)exec; (void)st; __atomic_add_fetch(&g_graph_launches_ptsz, 1, __ATOMIC_RELAXED); return 0; }
EXPORT CUresult cuGraphExecDestroy(void* exec) { (void)exec; g_graph_destroys++; return 0; }

EXPORT CUresult cuGetProcAddress_v2(const char* name, void** pfn, int version, uint64_t flags, void* status);
EXPORT CUresult cuGetProcAddress(const char* name, void** pfn, int version, uint64_t flags) {
  return cuGetProcAddress_v2(name, pfn, version, flags, NULL);
}
CUresult cuGetProcAddress_v2(const char* name, void** pfn, int version, uint64_t flags, void* status) {
  (void)status;
  const int ptsz = (flags & 2) != 0;
  void* p = NULL;
  if (!strcmp(name, "cuLaunchKernel")) p = ptsz ? (void*)cuLaunchKernel_ptsz : (void*)cuLaunchKernel;
  else if (!strcmp(name, "cuMemAlloc")) p = (void*)cuMemAlloc_v2;
  else if (!strcmp(name, "cuMemFree")) p = (void*)cuMemFree_v2;
  else if (!strcmp(name, "cuMemGetInfo")) p = (void*)cuMemGetInfo_v2;
  else if (!strcmp(name, "cuDeviceTotalMem")) p = (void*)cuDeviceTotalMem_v2;
  else if (!strcmp(name, "cuCtxGetDevice")) p = (void*)cuCtxGetDevice;
  else if (!strcmp(name, "cuDeviceGetUuid")) p = (void*)cuDeviceGetUuid_v2;
  else if (!strcmp(name, "cuInit")) p = (void*)cuInit;
  else if (!strcmp(name, "cuGraphInstantiateWithFlags")) p = (void*)cuGraphInstantiateWithFlags;
  else if (!strcmp(name, "cuGraphLaunch")) p = ptsz ? (void*)cuGraphLaunch_ptsz : (void*)cuGraphLaunch;
  else if (!strcmp(name, "cuGraphExecDestroy")) p = (void*)cuGraphExecDestroy;
  else if (!strcmp(name, "cuGetProcAddress")) p = version >= 12000 ? (void*)cuGetProcAddress_v2 : (void*)cuGetProcAddress;
  *pfn = p;
  return p ? 0 : 500;
}
