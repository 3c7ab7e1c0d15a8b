/* hook_probe.c -- a CUDA "application" for the CPU tests of libcuda_limiter.so: reaches the driver
 * the way libcudart does (dlopen + dlsym(cuGetProcAddress_v2) + cuGetProcAddress for the rest) or by
 * plain dlsym, launches and allocates, and prints what it saw as one JSON line.
 *   hook_probe <procaddr|dlsym> <launches> <grid> <block> [alloc_bytes ...] */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int (*launch_fn)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**);
typedef int (*alloc_fn)(unsigned long long*, size_t);
typedef int (*free_fn)(unsigned long long);
typedef int (*info_fn)(size_t*, size_t*);
typedef int (*gpa_fn)(const char*, void**, int, uint64_t, void*);

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const int by_proc = !strcmp(argv[1], "procaddr");
  const int per_thread = !strcmp(argv[1], "procaddr_ptsz");
  const long n = atol(argv[2]);
  const unsigned grid = (unsigned)atol(argv[3]), block = (unsigned)atol(argv[4]);
  void* h = dlopen("libcuda.so.1", RTLD_NOW);
  if (!h) { fprintf(stderr, "no libcuda.so.1: %s\n", dlerror()); return 3; }
  launch_fn launch = NULL; alloc_fn alloc = NULL; free_fn mfree = NULL; info_fn info = NULL;
  if (by_proc || per_thread) {
    gpa_fn gpa = (gpa_fn)dlsym(h, "cuGetProcAddress_v2");
    if (!gpa) return 4;
    gpa_fn gpa2 = NULL;  /* libcudart re-resolves cuGetProcAddress through itself */
    if (gpa("cuGetProcAddress", (void**)&gpa2, 12080, 0, NULL) || !gpa2) return 5;
    gpa2("cuLaunchKernel", (void**)&launch, 12080, per_thread ? 2 : 0, NULL);
    gpa2("cuMemAlloc", (void**)&alloc, 12080, 0, NULL);
    gpa2("cuMemFree", (void**)&mfree, 12080, 0, NULL);
    gpa2("cuMemGetInfo", (void**)&info, 12080, 0, NULL);
  } else {
    launch = (launch_fn)dlsym(h, "cuLaunchKernel");
    alloc = (alloc_fn)dlsym(h, "cuMemAlloc_v2");
    mfree = (free_fn)dlsym(h, "cuMemFree_v2");
    info = (info_fn)dlsym(h, "cuMemGetInfo_v2");
  }
  if (!launch || !alloc || !mfree || !info) return 6;
  /* "graph" as the 6th argument: instantiate the mock's graph and replay it <launches> times instead of launching kernels */
  if (argc > 5 && !strcmp(argv[5], "graph")) {
    typedef int (*inst_fn)(void**, void*, unsigned long long);
    typedef int (*glaunch_fn)(void*, void*);
    typedef int (*gdestroy_fn)(void*);
    inst_fn inst = NULL; glaunch_fn gl = NULL; gdestroy_fn gd = NULL;
    if (by_proc || per_thread) {
      gpa_fn gpa = (gpa_fn)dlsym(h, "cuGetProcAddress_v2");
      gpa("cuGraphInstantiateWithFlags", (void**)&inst, 12080, 0, NULL);
      gpa("cuGraphLaunch", (void**)&gl, 12080, per_thread ? 2 : 0, NULL);
      gpa("cuGraphExecDestroy", (void**)&gd, 12080, 0, NULL);
    } else {
      inst = (inst_fn)dlsym(h, "cuGraphInstantiateWithFlags");
      gl = (glaunch_fn)dlsym(h, "cuGraphLaunch");
      gd = (gdestroy_fn)dlsym(h, "cuGraphExecDestroy");
    }
    void* (*mg)(void) = (void* (*)(void))dlsym(h, "mock_cuda_graph");
    if (!inst || !gl || !gd || !mg) return 7;
    void* exec = NULL;
    if (inst(&exec, mg(), 0)) return 8;
    const double g0 = now_ms();
    int grc = 0;
    for (long i = 0; i < n && !grc; ++i) grc = gl(exec, NULL);
    const double gms = now_ms() - g0;
    gd(exec);
    uint64_t gc[3] = {0};
    void (*gcounts)(uint64_t*) = (void (*)(uint64_t*))dlsym(h, "mock_cuda_graph_counts");
    if (gcounts) gcounts(gc);
    struct { uint64_t launches, blocked, timeouts, wait_ns, tokens, denied, active; } gs = {0};
    void (*ghs)(void*) = (void (*)(void*))dlsym(RTLD_DEFAULT, "tf_hook_get_stats");
    if (ghs) ghs(&gs);
    printf("{\"graph_rc\": %d, \"graph_ms\": %.3f, \"driver_graph_launches\": %llu, \"driver_graph_launches_ptsz\": %llu, \"driver_graph_destroys\": %llu, "
           "\"hook_launches\": %llu, \"hook_tokens\": %llu, \"hook_blocked\": %llu}\n", grc, gms, (unsigned long long)gc[0], (unsigned long long)gc[1],
           (unsigned long long)gc[2], (unsigned long long)gs.launches, (unsigned long long)gs.tokens, (unsigned long long)gs.blocked);
    return 0;
  }
  /* an unrelated lookup must be untouched by the interposed dlsym */
  const int libc_ok = dlsym(RTLD_DEFAULT, "printf") != NULL && dlsym(RTLD_NEXT, "malloc") != NULL;

  const double t0 = now_ms();
  int rc = 0;
  for (long i = 0; i < n && !rc; ++i) rc = launch(NULL, grid, 1, 1, block, 1, 1, 0, NULL, NULL, NULL);
  const double launch_ms = now_ms() - t0;

  printf("{\"launch_rc\": %d, \"launch_ms\": %.3f, \"libc_ok\": %d, \"allocs\": [", rc, launch_ms, libc_ok);
  unsigned long long ptrs[64];
  int na = 0;
  for (int i = 5; i < argc && na < 64; ++i, ++na) {
    ptrs[na] = 0;
    const int r = alloc(&ptrs[na], (size_t)strtoull(argv[i], NULL, 10));
    printf("%s%d", na ? ", " : "", r);
    if (r) ptrs[na] = 0;
  }
  size_t f = 0, t = 0;
  info(&f, &t);
  printf("], \"free\": %zu, \"total\": %zu", f, t);
  for (int i = 0; i < na; ++i)
    if (ptrs[i]) mfree(ptrs[i]);
  info(&f, &t);
  printf(", \"free_after\": %zu", f);
  uint64_t c[5] = {0};
  void (*counts)(uint64_t*) = (void (*)(uint64_t*))dlsym(h, "mock_cuda_counts");
  if (counts) counts(c);
  printf(", \"driver_launches\": %llu, \"driver_launches_ptsz\": %llu, \"driver_allocs\": %llu, \"driver_frees\": %llu", (unsigned long long)c[0],
         (unsigned long long)c[1], (unsigned long long)c[2], (unsigned long long)c[3]);
  struct { uint64_t launches, blocked, timeouts, wait_ns, tokens, denied, active; } hs = {0};
  void (*hstats)(void*) = (void (*)(void*))dlsym(RTLD_DEFAULT, "tf_hook_get_stats");
  if (hstats) hstats(&hs);
  printf(", \"hooked\": %d, \"hook_active\": %llu, \"hook_launches\": %llu, \"hook_blocked\": %llu, \"hook_timeouts\": %llu, \"hook_tokens\": %llu, "
         "\"hook_denied\": %llu}\n", hstats != NULL, (unsigned long long)hs.active, (unsigned long long)hs.launches, (unsigned long long)hs.blocked,
         (unsigned long long)hs.timeouts, (unsigned long long)hs.tokens, (unsigned long long)hs.denied);
  return 0;
}
