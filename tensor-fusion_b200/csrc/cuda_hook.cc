// cuda_hook.cc -- libcuda_limiter.so: the LD_PRELOAD limiter of *local* soft isolation mode.
//
// The operator mounts a vendor limiter at /home/app/libcuda_limiter.so (or
// /tensor-fusion-limiter/libcuda_limiter.so), sets LD_PRELOAD to it together with
// TF_ISOLATION_MODE=soft and TF_SHM_PATH (pkg/constants/env.go:123-138,
// internal/utils/compose.go:1415-1458); the library is closed source in the reference
// ("cuda_hook", provider/limiter.h:67-83 documents the functions it calls).  This is that
// library for CUDA 12 / B200:
//
//   * every kernel launch of the process (cuLaunchKernel, cuLaunchKernelEx,
//     cuLaunchCooperativeKernel, stream-legacy and per-thread variants) is charged to the
//     pod's ERL token bucket in the quota file with CheckAndRecordComputeOps; a launch that
//     finds the bucket short waits (sleeping, not spinning) until the hypervisor's refill
//     covers it;
//   * device memory (cuMemAlloc/Managed/Pitch/Async, cuMemCreate) is charged with
//     CheckAndRecordMemoryOps; an allocation over the pod's limit fails with
//     CUDA_ERROR_OUT_OF_MEMORY before reaching the driver; cuMemGetInfo / cuDeviceTotalMem
//     report the pod's view.
//
// Interposition covers the three ways a CUDA program reaches the driver: direct linking
// (exported symbols), dlopen + dlsym (the interposed dlsym below) and cuGetProcAddress (what
// libcudart >= 11.3 uses for everything).  The real driver is $TENSOR_FUSION_NGPU_PATH
// (pkg/constants/env.go:119-120) when set, else libcuda.so.1.
//
// Inside this repo's own worker the gate lives on the GPU (gate.cu); the hook recognises the
// worker library in the process and only forwards.
//
// Cost of a launch in tokens: thread blocks x warps per block.
#include <dlfcn.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "hv_handshake.h"
#include "provider_log.h"
#include "tf_provider_abi.h"

#define HOOK_EXPORT extern "C" __attribute__((visibility("default")))

// ---- the slice of the driver API this file needs (cuda.h renames symbols with macros, so it is not included)
typedef int CUresult;
typedef int CUdevice;
typedef unsigned long long CUdeviceptr;
typedef unsigned long long CUmemGenericAllocationHandle;
typedef struct CUstream_st* CUstream;
typedef struct CUfunc_st* CUfunction;
typedef struct CUmemPool_st* CUmemoryPool;
typedef uint64_t cuuint64_t;
struct CUuuid { unsigned char bytes[16]; };
struct CUlaunchConfigHead {  // leading members of CUlaunchConfig (cuda.h: CUlaunchConfig_st)
  unsigned gridDimX, gridDimY, gridDimZ, blockDimX, blockDimY, blockDimZ, sharedMemBytes;
};
enum : CUresult { CUDA_SUCCESS_ = 0, CUDA_ERROR_OUT_OF_MEMORY_ = 2, CUDA_ERROR_NOT_FOUND_ = 500 };
constexpr cuuint64_t kProcPerThreadStream = 1u << 1;  // CU_GET_PROC_ADDRESS_PER_THREAD_DEFAULT_STREAM

namespace {

// ---------------------------------------------------------------- configuration and log
struct Config {
  bool active = false;       // charge launches / memory
  bool log = false;
  bool graphs = true;        // cuGraphLaunch is charged with the sum of the graph's kernel nodes (TF_LIMITER_CHARGE_GRAPHS=0 turns it off)
  long max_wait_ms = 5000;   // fail-open bound of one blocked launch (same as the GPU gate's watchdog)
};
Config g_cfg;
std::atomic<uint64_t> g_launches{0}, g_blocked{0}, g_timeouts{0}, g_wait_ns{0}, g_tokens{0}, g_denied_allocs{0};

void hlog(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void hlog(const char* fmt, ...) {
  if (!g_cfg.log) return;
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[libcuda_limiter] ");
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  va_end(ap);
}

// ---------------------------------------------------------------- the real dlsym and the real driver
using dlsym_fn = void* (*)(void*, const char*);
dlsym_fn real_dlsym() {
  static dlsym_fn fn = [] {
    void* p = dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.2.5");  // x86-64 baseline version
    if (!p) p = dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.34");   // libdl merged into libc
    if (!p) p = dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.17");   // aarch64 baseline
    return reinterpret_cast<dlsym_fn>(p);
  }();
  return fn;
}

void* driver_handle() {
  static void* h = [] {
    const char* ngpu = getenv("TENSOR_FUSION_NGPU_PATH");
    void* p = nullptr;
    if (ngpu && *ngpu) p = dlopen(ngpu, RTLD_NOW | RTLD_GLOBAL);
    if (!p) p = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!p) p = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    return p;
  }();
  return h;
}

// true when `p` points into this library (a driver that was not linked -Bsymbolic can hand our own
// interposed exports back to us; treating those as "real" would recurse)
bool is_ours(void* p) {
  static const void* base = [] {
    Dl_info me{};
    return dladdr(reinterpret_cast<void*>(&is_ours), &me) ? me.dli_fbase : nullptr;
  }();
  Dl_info d{};
  return p && dladdr(p, &d) && d.dli_fbase == base;
}

void* driver_sym(const char* name) {
  void* h = driver_handle();
  dlsym_fn ds = real_dlsym();
  return h && ds ? ds(h, name) : nullptr;
}

// ---------------------------------------------------------------- hook table
// One slot per (entry point, stream flavour).  `real` is whatever the driver handed out for that
// request last (by export, dlsym or cuGetProcAddress); all of them share one ABI per slot.
enum Slot {
  kLaunch, kLaunchPtsz, kLaunchEx, kLaunchExPtsz, kLaunchCoop, kLaunchCoopPtsz,
  kMemAlloc, kMemAllocManaged, kMemAllocPitch, kMemFree,
  kMemAllocAsync, kMemAllocAsyncPtsz, kMemAllocFromPoolAsync, kMemAllocFromPoolAsyncPtsz, kMemFreeAsync, kMemFreeAsyncPtsz,
  kMemCreate, kMemRelease, kMemGetInfo, kDeviceTotalMem,
  kGetProcAddress, kGetProcAddressV2,
  kGraphInstantiateWithFlags, kGraphInstantiateWithParams, kGraphInstantiateWithParamsPtsz, kGraphLaunch, kGraphLaunchPtsz, kGraphExecUpdate, kGraphExecDestroy,
  kSlotCount
};
std::atomic<void*> g_real[kSlotCount];

struct HookName {
  const char* exported;  // symbol in libcuda.so
  const char* proc;      // name used with cuGetProcAddress (no suffix); nullptr = same as previous row
  bool ptsz;
  Slot slot;
  void* hook;
};
extern const HookName kHooks[];
extern const size_t kHookCount;

const HookName* hook_by_export(const char* name) {
  for (size_t i = 0; i < kHookCount; ++i)
    if (std::strcmp(kHooks[i].exported, name) == 0) return &kHooks[i];
  return nullptr;
}
template <typename F>
F real_of(Slot s) {
  void* p = g_real[s].load(std::memory_order_acquire);
  if (!p) {
    for (size_t i = 0; i < kHookCount; ++i)
      if (kHooks[i].slot == s) {
        p = driver_sym(kHooks[i].exported);
        break;
      }
    if (p) g_real[s].store(p, std::memory_order_release);
  }
  return reinterpret_cast<F>(p);
}

// ---------------------------------------------------------------- device identity
struct DevId {
  std::once_flag once;
  char uuid[64] = {0};
  bool ok = false;
  void* bucket = nullptr;  // the device's entry in the pod's quota file, resolved once
  int shm_idx = -1;
};
DevId g_dev[64];
char g_pid[32];

DevId* current_dev();
const char* current_uuid() {
  DevId* d = current_dev();
  return d ? d->uuid : nullptr;
}

DevId* current_dev() {
  using ctx_get_dev = CUresult (*)(CUdevice*);
  using dev_uuid = CUresult (*)(CUuuid*, CUdevice);
  static ctx_get_dev get_dev = reinterpret_cast<ctx_get_dev>(driver_sym("cuCtxGetDevice"));
  static dev_uuid get_uuid = [] {
    void* p = driver_sym("cuDeviceGetUuid_v2");
    if (!p) p = driver_sym("cuDeviceGetUuid");
    return reinterpret_cast<dev_uuid>(p);
  }();
  if (!get_dev || !get_uuid) return nullptr;
  CUdevice d = -1;
  if (get_dev(&d) != CUDA_SUCCESS_ || d < 0 || d >= 64) return nullptr;
  DevId& id = g_dev[d];
  std::call_once(id.once, [&] {
    CUuuid u{};
    if (get_uuid(&u, d) != CUDA_SUCCESS_) return;
    const unsigned char* b = u.bytes;
    snprintf(id.uuid, sizeof id.uuid, "GPU-%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", b[0], b[1], b[2],
             b[3], b[4], b[5], b[6], b[7], b[8], b[9], b[10], b[11], b[12], b[13], b[14], b[15]);
    uint64_t lim = 0, used = 0;
    uint32_t up = 0;
    id.ok = tfprov::self_limits(id.uuid, &lim, &used, &up);
    if (id.ok) id.bucket = tfprov::self_bucket(id.uuid, &id.shm_idx);
    id.ok = id.ok && id.bucket;
    hlog("device %d = %s: %s (up_limit %u%%, mem_limit %llu)", d, id.uuid, id.ok ? "limited" : "not in this pod's quota file", up,
         (unsigned long long)lim);
  });
  return id.ok ? &id : nullptr;
}

// ---------------------------------------------------------------- gates
uint64_t mono_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

inline uint64_t launch_tokens(uint64_t blocks, uint64_t threads_per_block) {
  const uint64_t t = blocks * ((threads_per_block + 31) / 32);
  return t ? t : 1;
}

void gate_tokens(uint64_t tokens);
void gate_compute(uint64_t blocks, uint64_t threads_per_block) {
  if (g_cfg.active) gate_tokens(launch_tokens(blocks, threads_per_block));
}

void gate_tokens(uint64_t tokens) {
  if (!g_cfg.active || !tokens) return;
  DevId* dev = current_dev();
  if (!dev) return;
  const char* uuid = dev->uuid;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  // fast path: one lock-free FetchSub on the bucket resolved at first use -- the same arithmetic as
  // CheckAndRecordComputeOps (which the slow path below keeps calling while it waits)
  const double cost = tfprov::clamp_cost(dev->bucket, dev->shm_idx, (double)tokens);  // never more than the bucket can hold
  tokens = (uint64_t)cost;
  const bool short_of = tfprov::self_charge(dev->bucket, dev->shm_idx, cost) < cost;
  ComputeOpRecord rec;
  rec.shouldBlock = short_of;
  if (rec.shouldBlock) {
    g_blocked.fetch_add(1, std::memory_order_relaxed);
    const uint64_t t0 = mono_ns();
    long sleep_us = 50;
    bool admitted = false;
    while ((long)((mono_ns() - t0) / 1000000ull) < g_cfg.max_wait_ms) {
      timespec ts{0, sleep_us * 1000};
      nanosleep(&ts, nullptr);
      if (sleep_us < 2000) sleep_us *= 2;
      if (CheckAndRecordComputeOps(g_pid, uuid, tokens, &rec) != ACCEL_SUCCESS || !rec.shouldBlock) {
        admitted = true;
        break;
      }
    }
    // a launch larger than the bucket can ever hold, or a dead hypervisor, must not hang the tenant
    if (!admitted) g_timeouts.fetch_add(1, std::memory_order_relaxed);
    g_wait_ns.fetch_add(mono_ns() - t0, std::memory_order_relaxed);
  }
  g_tokens.fetch_add(tokens, std::memory_order_relaxed);
}

// true = the allocation may proceed (and has been recorded)
bool charge_memory(int64_t bytes) {
  if (!g_cfg.active || bytes == 0) return true;
  const char* uuid = current_uuid();
  if (!uuid) return true;
  MemoryOpRecord rec;
  if (CheckAndRecordMemoryOps(g_pid, uuid, bytes, &rec) != ACCEL_SUCCESS) return true;
  if (rec.shouldBlock) {
    g_denied_allocs.fetch_add(1, std::memory_order_relaxed);
    hlog("allocation of %lld bytes denied: %llu bytes left of the pod's limit", (long long)bytes, (unsigned long long)rec.availableBytes);
    return false;
  }
  return true;
}

std::mutex g_mem_mu;
std::unordered_map<unsigned long long, uint64_t>& allocs() {
  static auto* m = new std::unordered_map<unsigned long long, uint64_t>();  // leaked on purpose: frees may run during exit
  return *m;
}
void remember(unsigned long long key, uint64_t bytes) {
  std::lock_guard<std::mutex> lk(g_mem_mu);
  allocs()[key] = bytes;
}
uint64_t forget(unsigned long long key) {
  std::lock_guard<std::mutex> lk(g_mem_mu);
  auto it = allocs().find(key);
  if (it == allocs().end()) return 0;
  const uint64_t b = it->second;
  allocs().erase(it);
  return b;
}

// ---------------------------------------------------------------- start-up
std::atomic<int> g_handshake{0};  // 0 none, 1 in flight, 2 finished

void print_summary() {
  // a short-lived process still gets registered: give an in-flight handshake up to a second
  for (int i = 0; i < 100 && g_handshake.load() == 1; ++i) usleep(10000);
  if (!g_cfg.log) return;
  hlog("launches %llu (blocked %llu, timed out %llu, waited %.3f ms), tokens %llu, denied allocations %llu",
       (unsigned long long)g_launches.load(), (unsigned long long)g_blocked.load(), (unsigned long long)g_timeouts.load(),
       (double)g_wait_ns.load() / 1e6, (unsigned long long)g_tokens.load(), (unsigned long long)g_denied_allocs.load());
}

__attribute__((constructor)) void hook_init() {
  snprintf(g_pid, sizeof g_pid, "%ld", (long)getpid());
  const char* lg = getenv("TF_LIMITER_LOG");
  g_cfg.log = lg && *lg && std::strcmp(lg, "0") != 0;
  const char* gr = getenv("TF_LIMITER_CHARGE_GRAPHS");
  g_cfg.graphs = !(gr && std::strcmp(gr, "0") == 0);  // on unless switched off: replayed graphs would otherwise run for free
  if (const char* w = getenv("TF_LIMITER_MAX_WAIT_MS")) g_cfg.max_wait_ms = atol(w) > 0 ? atol(w) : g_cfg.max_wait_ms;
  const char* shm = getenv("TF_SHM_PATH");                 // pkg/constants/env.go:133-136
  const char* off = getenv("DISABLE_GPU_LIMITER");         // env.go:140-141
  const char* mode = getenv("TF_ISOLATION_MODE");          // "soft" is the mode this library implements
  const bool mode_ok = !mode || !*mode || strcasecmp(mode, "soft") == 0;
  // this repo's worker charges launches on the GPU: do not charge them twice
  const bool in_worker = real_dlsym() && real_dlsym()(RTLD_DEFAULT, "tfw_worker_create") != nullptr;
  g_cfg.active = shm && *shm && !(off && *off) && mode_ok && !in_worker && access(shm, R_OK | W_OK) == 0;
  hlog("pid %s: %s (TF_SHM_PATH=%s, TF_ISOLATION_MODE=%s%s)", g_pid, g_cfg.active ? "limiting" : "forwarding only", shm ? shm : "",
       mode ? mode : "", in_worker ? ", inside tensor-fusion-worker" : "");
  if (g_cfg.active && getenv("HYPERVISOR_IP")) {
    // registers this PID in the quota file through the hypervisor (which knows the host PID)
    g_handshake.store(1);
    std::thread([] {
      const tfhv::Result r = tfhv::handshake("main");
      g_handshake.store(2);
      hlog("hypervisor handshake: %s", r.registered ? "registered" : r.reached ? "pod found, process not registered" : "not reachable");
    }).detach();
  }
  atexit(print_summary);
}

// ---------------------------------------------------------------- hooks (signatures: cuda.h 12.x)
#define LAUNCH_ARGS                                                                                                            \
  CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned smem, CUstream st, \
      void **params, void **extra
using launch_fn = CUresult (*)(LAUNCH_ARGS);
template <Slot S>
CUresult launch_hook(LAUNCH_ARGS) {
  launch_fn real = real_of<launch_fn>(S);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  gate_compute((uint64_t)gx * gy * gz, (uint64_t)bx * by * bz);
  return real(f, gx, gy, gz, bx, by, bz, smem, st, params, extra);
}

using launch_ex_fn = CUresult (*)(const void* cfg, CUfunction f, void** params, void** extra);
template <Slot S>
CUresult launch_ex_hook(const void* cfg, CUfunction f, void** params, void** extra) {
  launch_ex_fn real = real_of<launch_ex_fn>(S);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  if (cfg) {
    const CUlaunchConfigHead* c = static_cast<const CUlaunchConfigHead*>(cfg);
    gate_compute((uint64_t)c->gridDimX * c->gridDimY * c->gridDimZ, (uint64_t)c->blockDimX * c->blockDimY * c->blockDimZ);
  }
  return real(cfg, f, params, extra);
}

using launch_coop_fn = CUresult (*)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**);
template <Slot S>
CUresult launch_coop_hook(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned smem,
                          CUstream st, void** params) {
  launch_coop_fn real = real_of<launch_coop_fn>(S);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  gate_compute((uint64_t)gx * gy * gz, (uint64_t)bx * by * bz);
  return real(f, gx, gy, gz, bx, by, bz, smem, st, params);
}

CUresult mem_alloc_hook(CUdeviceptr* p, size_t bytes) {
  auto real = real_of<CUresult (*)(CUdeviceptr*, size_t)>(kMemAlloc);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  if (!charge_memory((int64_t)bytes)) return CUDA_ERROR_OUT_OF_MEMORY_;
  const CUresult r = real(p, bytes);
  if (r == CUDA_SUCCESS_ && p) remember(*p, bytes);
  else charge_memory(-(int64_t)bytes);
  return r;
}

CUresult mem_alloc_managed_hook(CUdeviceptr* p, size_t bytes, unsigned flags) {
  auto real = real_of<CUresult (*)(CUdeviceptr*, size_t, unsigned)>(kMemAllocManaged);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  if (!charge_memory((int64_t)bytes)) return CUDA_ERROR_OUT_OF_MEMORY_;
  const CUresult r = real(p, bytes, flags);
  if (r == CUDA_SUCCESS_ && p) remember(*p, bytes);
  else charge_memory(-(int64_t)bytes);
  return r;
}

CUresult mem_alloc_pitch_hook(CUdeviceptr* p, size_t* pitch, size_t width, size_t height, unsigned elem) {
  auto real = real_of<CUresult (*)(CUdeviceptr*, size_t*, size_t, size_t, unsigned)>(kMemAllocPitch);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const uint64_t est = (uint64_t)((width + 511) & ~(size_t)511) * height;  // pitch is at least the 512-B rounded width
  if (!charge_memory((int64_t)est)) return CUDA_ERROR_OUT_OF_MEMORY_;
  const CUresult r = real(p, pitch, width, height, elem);
  if (r == CUDA_SUCCESS_ && p && pitch) {
    const uint64_t actual = (uint64_t)*pitch * height;
    if (actual != est) charge_memory((int64_t)actual - (int64_t)est);
    remember(*p, actual);
  } else {
    charge_memory(-(int64_t)est);
  }
  return r;
}

CUresult mem_free_hook(CUdeviceptr p) {
  auto real = real_of<CUresult (*)(CUdeviceptr)>(kMemFree);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(p);
  if (r == CUDA_SUCCESS_) {
    const uint64_t b = forget(p);
    if (b) charge_memory(-(int64_t)b);
  }
  return r;
}

template <Slot S>
CUresult mem_alloc_async_hook(CUdeviceptr* p, size_t bytes, CUstream st) {
  auto real = real_of<CUresult (*)(CUdeviceptr*, size_t, CUstream)>(S);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  if (!charge_memory((int64_t)bytes)) return CUDA_ERROR_OUT_OF_MEMORY_;
  const CUresult r = real(p, bytes, st);
  if (r == CUDA_SUCCESS_ && p) remember(*p, bytes);
  else charge_memory(-(int64_t)bytes);
  return r;
}

template <Slot S>
CUresult mem_alloc_from_pool_async_hook(CUdeviceptr* p, size_t bytes, CUmemoryPool pool, CUstream st) {
  auto real = real_of<CUresult (*)(CUdeviceptr*, size_t, CUmemoryPool, CUstream)>(S);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  if (!charge_memory((int64_t)bytes)) return CUDA_ERROR_OUT_OF_MEMORY_;
  const CUresult r = real(p, bytes, pool, st);
  if (r == CUDA_SUCCESS_ && p) remember(*p, bytes);
  else charge_memory(-(int64_t)bytes);
  return r;
}

template <Slot S>
CUresult mem_free_async_hook(CUdeviceptr p, CUstream st) {
  auto real = real_of<CUresult (*)(CUdeviceptr, CUstream)>(S);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(p, st);
  if (r == CUDA_SUCCESS_) {
    const uint64_t b = forget(p);
    if (b) charge_memory(-(int64_t)b);
  }
  return r;
}

// physical allocations of the virtual-memory API (what PyTorch's expandable segments and this repo's
// own tiering use); the handle is the key
CUresult mem_create_hook(CUmemGenericAllocationHandle* h, size_t bytes, const void* prop, unsigned long long flags) {
  auto real = real_of<CUresult (*)(CUmemGenericAllocationHandle*, size_t, const void*, unsigned long long)>(kMemCreate);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  // CUmemAllocationProp.location.type == CU_MEM_LOCATION_TYPE_DEVICE(1) is the only kind that uses HBM
  const int* pr = static_cast<const int*>(prop);
  const bool device = !pr || pr[2] == 1;  // {type, requestedHandleTypes, location.type, location.id, ...}
  if (device && !charge_memory((int64_t)bytes)) return CUDA_ERROR_OUT_OF_MEMORY_;
  const CUresult r = real(h, bytes, prop, flags);
  if (device) {
    if (r == CUDA_SUCCESS_ && h) remember(*h | (1ull << 63), bytes);
    else charge_memory(-(int64_t)bytes);
  }
  return r;
}

CUresult mem_release_hook(CUmemGenericAllocationHandle h) {
  auto real = real_of<CUresult (*)(CUmemGenericAllocationHandle)>(kMemRelease);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(h);
  if (r == CUDA_SUCCESS_) {
    const uint64_t b = forget(h | (1ull << 63));
    if (b) charge_memory(-(int64_t)b);
  }
  return r;
}

CUresult mem_get_info_hook(size_t* free_b, size_t* total_b) {
  auto real = real_of<CUresult (*)(size_t*, size_t*)>(kMemGetInfo);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(free_b, total_b);
  if (r != CUDA_SUCCESS_ || !g_cfg.active) return r;
  const char* uuid = current_uuid();
  uint64_t lim = 0, used = 0;
  if (uuid && tfprov::self_limits(uuid, &lim, &used, nullptr) && lim) {
    const uint64_t avail = lim > used ? lim - used : 0;
    if (total_b && *total_b > lim) *total_b = lim;
    if (free_b && *free_b > avail) *free_b = avail;
  }
  return r;
}

CUresult device_total_mem_hook(size_t* bytes, CUdevice dev) {
  auto real = real_of<CUresult (*)(size_t*, CUdevice)>(kDeviceTotalMem);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(bytes, dev);
  if (r != CUDA_SUCCESS_ || !g_cfg.active || !bytes) return r;
  const char* uuid = current_uuid();  // the pod's devices share one limit per device entry; current device is the best key
  uint64_t lim = 0;
  if (uuid && tfprov::self_limits(uuid, &lim, nullptr, nullptr) && lim && *bytes > lim) *bytes = lim;
  return r;
}

// ---- CUDA graphs (TF_LIMITER_CHARGE_GRAPHS=0 turns the accounting off) --------------------------------------
// A replayed graph launches its kernels without passing cuLaunchKernel.  At instantiation the cost of the graph
// is computed once -- blocks x warps summed over its kernel nodes, child graphs included -- and every
// cuGraphLaunch of that executable graph is charged with it (clamped to the bucket's capacity like any launch).
// cuGraphExecUpdate re-computes the cost from the graph it was updated with.
typedef struct CUgraph_st* CUgraph;
typedef struct CUgraphNode_st* CUgraphNode;
typedef struct CUgraphExec_st* CUgraphExec;
struct KernelNodeParams {  // CUDA_KERNEL_NODE_PARAMS_v2 (cuda.h 12.x); the v1 struct is a prefix of it
  void* func;
  unsigned gridDimX, gridDimY, gridDimZ, blockDimX, blockDimY, blockDimZ, sharedMemBytes;
  void** kernelParams;
  void** extra;
  void* kern;
  void* ctx;
};
std::mutex g_graph_mu;
std::unordered_map<CUgraphExec, uint64_t>& graph_costs() {
  static auto* m = new std::unordered_map<CUgraphExec, uint64_t>();
  return *m;
}

uint64_t graph_cost(CUgraph g, int depth) {
  using get_nodes_fn = CUresult (*)(CUgraph, CUgraphNode*, size_t*);
  using node_type_fn = CUresult (*)(CUgraphNode, int*);
  using kparams_fn = CUresult (*)(CUgraphNode, KernelNodeParams*);
  using child_fn = CUresult (*)(CUgraphNode, CUgraph*);
  static get_nodes_fn get_nodes = reinterpret_cast<get_nodes_fn>(driver_sym("cuGraphGetNodes"));
  static node_type_fn node_type = reinterpret_cast<node_type_fn>(driver_sym("cuGraphNodeGetType"));
  static kparams_fn kparams = [] {
    void* p = driver_sym("cuGraphKernelNodeGetParams_v2");
    if (!p) p = driver_sym("cuGraphKernelNodeGetParams");
    return reinterpret_cast<kparams_fn>(p);
  }();
  static child_fn child = reinterpret_cast<child_fn>(driver_sym("cuGraphChildGraphNodeGetGraph"));
  if (!g || !get_nodes || !node_type || !kparams || depth > 8) return 0;
  size_t n = 0;
  if (get_nodes(g, nullptr, &n) != CUDA_SUCCESS_ || !n) return 0;
  std::vector<CUgraphNode> nodes(n);
  if (get_nodes(g, nodes.data(), &n) != CUDA_SUCCESS_) return 0;
  uint64_t cost = 0;
  for (size_t i = 0; i < n && i < nodes.size(); ++i) {
    int type = -1;
    if (node_type(nodes[i], &type) != CUDA_SUCCESS_) continue;
    if (type == 0) {  // CU_GRAPH_NODE_TYPE_KERNEL
      KernelNodeParams p{};
      if (kparams(nodes[i], &p) == CUDA_SUCCESS_)
        cost += launch_tokens((uint64_t)p.gridDimX * p.gridDimY * p.gridDimZ, (uint64_t)p.blockDimX * p.blockDimY * p.blockDimZ);
    } else if (type == 4 && child) {  // CU_GRAPH_NODE_TYPE_GRAPH
      CUgraph sub = nullptr;
      if (child(nodes[i], &sub) == CUDA_SUCCESS_) cost += graph_cost(sub, depth + 1);
    }
  }
  return cost;
}

void remember_graph(CUgraphExec e, CUgraph g) {
  if (!g_cfg.active || !g_cfg.graphs || !e) return;
  const uint64_t cost = graph_cost(g, 0);
  std::lock_guard<std::mutex> lk(g_graph_mu);
  graph_costs()[e] = cost;
}

CUresult graph_instantiate_flags_hook(CUgraphExec* e, CUgraph g, unsigned long long flags) {
  auto real = real_of<CUresult (*)(CUgraphExec*, CUgraph, unsigned long long)>(kGraphInstantiateWithFlags);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(e, g, flags);
  if (r == CUDA_SUCCESS_ && e) remember_graph(*e, g);
  return r;
}
template <Slot S>
CUresult graph_instantiate_params_hook(CUgraphExec* e, CUgraph g, void* params) {
  auto real = real_of<CUresult (*)(CUgraphExec*, CUgraph, void*)>(S);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(e, g, params);
  if (r == CUDA_SUCCESS_ && e) remember_graph(*e, g);
  return r;
}
template <Slot S>
CUresult graph_launch_hook(CUgraphExec e, CUstream st) {
  auto real = real_of<CUresult (*)(CUgraphExec, CUstream)>(S);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  if (g_cfg.active && g_cfg.graphs) {
    uint64_t cost = 0;
    {
      std::lock_guard<std::mutex> lk(g_graph_mu);
      const auto it = graph_costs().find(e);
      if (it != graph_costs().end()) cost = it->second;
    }
    gate_tokens(cost);
  }
  return real(e, st);
}
CUresult graph_exec_update_hook(CUgraphExec e, CUgraph g, void* result_info) {
  auto real = real_of<CUresult (*)(CUgraphExec, CUgraph, void*)>(kGraphExecUpdate);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(e, g, result_info);
  if (r == CUDA_SUCCESS_) remember_graph(e, g);  // the executable graph now runs g's kernels: charge those
  return r;
}
CUresult graph_exec_destroy_hook(CUgraphExec e) {
  auto real = real_of<CUresult (*)(CUgraphExec)>(kGraphExecDestroy);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  {
    std::lock_guard<std::mutex> lk(g_graph_mu);
    graph_costs().erase(e);
  }
  return real(e);
}
inline bool graph_slot(Slot s) { return s >= kGraphInstantiateWithFlags && s <= kGraphExecDestroy; }

// cuGetProcAddress: let the driver resolve the (name, version, flags) request, then substitute the hook
void* substitute(const char* symbol, int version, cuuint64_t flags, void* real) {
  if (!symbol || !real) return real;
  const bool ptsz = (flags & kProcPerThreadStream) != 0;
  if (std::strcmp(symbol, "cuGetProcAddress") == 0) {  // the 5-argument form is what a >= 12.0 request gets
    const HookName* h = hook_by_export(version >= 12000 ? "cuGetProcAddress_v2" : "cuGetProcAddress");
    if (!is_ours(real)) g_real[h->slot].store(real, std::memory_order_release);
    return h->hook;
  }
  for (size_t i = 0; i < kHookCount; ++i) {
    const HookName& h = kHooks[i];
    if (std::strcmp(h.proc, symbol) != 0) continue;
    // stream-flavoured entry points have two rows; the others one
    bool has_ptsz_row = false;
    for (size_t j = 0; j < kHookCount; ++j)
      if (j != i && std::strcmp(kHooks[j].proc, symbol) == 0) has_ptsz_row = true;
    if (has_ptsz_row && h.ptsz != ptsz) continue;
    if (graph_slot(h.slot) && !g_cfg.graphs) return real;  // opt-in feature off: the application talks to the driver directly
    if (is_ours(real)) return h.hook;  // the slot falls back to the driver's export
    g_real[h.slot].store(real, std::memory_order_release);
    return h.hook;
  }
  return real;
}

CUresult get_proc_address_hook(const char* symbol, void** pfn, int version, cuuint64_t flags) {
  auto real = real_of<CUresult (*)(const char*, void**, int, cuuint64_t)>(kGetProcAddress);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(symbol, pfn, version, flags);
  if (r == CUDA_SUCCESS_ && pfn) *pfn = substitute(symbol, version, flags, *pfn);
  return r;
}

CUresult get_proc_address_v2_hook(const char* symbol, void** pfn, int version, cuuint64_t flags, void* status) {
  auto real = real_of<CUresult (*)(const char*, void**, int, cuuint64_t, void*)>(kGetProcAddressV2);
  if (!real) return CUDA_ERROR_NOT_FOUND_;
  const CUresult r = real(symbol, pfn, version, flags, status);
  if (r == CUDA_SUCCESS_ && pfn) *pfn = substitute(symbol, version, flags, *pfn);
  return r;
}

#define H(fn) reinterpret_cast<void*>(fn)
const HookName kHooks[] = {
    {"cuLaunchKernel", "cuLaunchKernel", false, kLaunch, H(launch_hook<kLaunch>)},
    {"cuLaunchKernel_ptsz", "cuLaunchKernel", true, kLaunchPtsz, H(launch_hook<kLaunchPtsz>)},
    {"cuLaunchKernelEx", "cuLaunchKernelEx", false, kLaunchEx, H(launch_ex_hook<kLaunchEx>)},
    {"cuLaunchKernelEx_ptsz", "cuLaunchKernelEx", true, kLaunchExPtsz, H(launch_ex_hook<kLaunchExPtsz>)},
    {"cuLaunchCooperativeKernel", "cuLaunchCooperativeKernel", false, kLaunchCoop, H(launch_coop_hook<kLaunchCoop>)},
    {"cuLaunchCooperativeKernel_ptsz", "cuLaunchCooperativeKernel", true, kLaunchCoopPtsz, H(launch_coop_hook<kLaunchCoopPtsz>)},
    {"cuMemAlloc_v2", "cuMemAlloc", false, kMemAlloc, H(mem_alloc_hook)},
    {"cuMemAllocManaged", "cuMemAllocManaged", false, kMemAllocManaged, H(mem_alloc_managed_hook)},
    {"cuMemAllocPitch_v2", "cuMemAllocPitch", false, kMemAllocPitch, H(mem_alloc_pitch_hook)},
    {"cuMemFree_v2", "cuMemFree", false, kMemFree, H(mem_free_hook)},
    {"cuMemAllocAsync", "cuMemAllocAsync", false, kMemAllocAsync, H(mem_alloc_async_hook<kMemAllocAsync>)},
    {"cuMemAllocAsync_ptsz", "cuMemAllocAsync", true, kMemAllocAsyncPtsz, H(mem_alloc_async_hook<kMemAllocAsyncPtsz>)},
    {"cuMemAllocFromPoolAsync", "cuMemAllocFromPoolAsync", false, kMemAllocFromPoolAsync,
     H(mem_alloc_from_pool_async_hook<kMemAllocFromPoolAsync>)},
    {"cuMemAllocFromPoolAsync_ptsz", "cuMemAllocFromPoolAsync", true, kMemAllocFromPoolAsyncPtsz,
     H(mem_alloc_from_pool_async_hook<kMemAllocFromPoolAsyncPtsz>)},
    {"cuMemFreeAsync", "cuMemFreeAsync", false, kMemFreeAsync, H(mem_free_async_hook<kMemFreeAsync>)},
    {"cuMemFreeAsync_ptsz", "cuMemFreeAsync", true, kMemFreeAsyncPtsz, H(mem_free_async_hook<kMemFreeAsyncPtsz>)},
    {"cuMemCreate", "cuMemCreate", false, kMemCreate, H(mem_create_hook)},
    {"cuMemRelease", "cuMemRelease", false, kMemRelease, H(mem_release_hook)},
    {"cuMemGetInfo_v2", "cuMemGetInfo", false, kMemGetInfo, H(mem_get_info_hook)},
    {"cuDeviceTotalMem_v2", "cuDeviceTotalMem", false, kDeviceTotalMem, H(device_total_mem_hook)},
    {"cuGraphInstantiateWithFlags", "cuGraphInstantiateWithFlags", false, kGraphInstantiateWithFlags, H(graph_instantiate_flags_hook)},
    {"cuGraphInstantiateWithParams", "cuGraphInstantiateWithParams", false, kGraphInstantiateWithParams,
     H(graph_instantiate_params_hook<kGraphInstantiateWithParams>)},
    {"cuGraphInstantiateWithParams_ptsz", "cuGraphInstantiateWithParams", true, kGraphInstantiateWithParamsPtsz,
     H(graph_instantiate_params_hook<kGraphInstantiateWithParamsPtsz>)},
    {"cuGraphLaunch", "cuGraphLaunch", false, kGraphLaunch, H(graph_launch_hook<kGraphLaunch>)},
    {"cuGraphLaunch_ptsz", "cuGraphLaunch", true, kGraphLaunchPtsz, H(graph_launch_hook<kGraphLaunchPtsz>)},
    {"cuGraphExecUpdate_v2", "cuGraphExecUpdate", false, kGraphExecUpdate, H(graph_exec_update_hook)},
    {"cuGraphExecDestroy", "cuGraphExecDestroy", false, kGraphExecDestroy, H(graph_exec_destroy_hook)},
    {"cuGetProcAddress", "cuGetProcAddress", false, kGetProcAddress, H(get_proc_address_hook)},
    {"cuGetProcAddress_v2", "cuGetProcAddress", false, kGetProcAddressV2, H(get_proc_address_v2_hook)},
};
const size_t kHookCount = sizeof(kHooks) / sizeof(kHooks[0]);

}  // namespace

// the provider's log sink, for the limiter objects linked into this library
namespace tfprov {
void log(const char* level, const char* msg) { hlog("%s %s", level, msg); }
}  // namespace tfprov

// ---------------------------------------------------------------- exported interposers
// dlopen()+dlsym() users (libcudart looks up cuGetProcAddress_v2 this way).  Everything that is not a
// hooked driver symbol leaves through a sibling call (a jmp, checked in the build), so glibc still sees the
// original caller's return address and RTLD_NEXT keeps its meaning for other interposers.
static void* hooked_dlsym(const HookName* h, void* handle, const char* name) {
  dlsym_fn ds = real_dlsym();
  void* real = (handle == RTLD_DEFAULT || handle == RTLD_NEXT) ? nullptr : ds(handle, name);
  if (!real || is_ours(real)) real = driver_sym(name);
  if (!real) return nullptr;
  if (!is_ours(real)) g_real[h->slot].store(real, std::memory_order_release);
  return h->hook;
}

HOOK_EXPORT void* dlsym(void* handle, const char* name) {
  if (name[0] == 'c' && name[1] == 'u') {
    if (const HookName* h = hook_by_export(name))
      if (!graph_slot(h->slot) || g_cfg.graphs) return hooked_dlsym(h, handle, name);
  }
  dlsym_fn ds = real_dlsym();
  if (!ds) return nullptr;
  return ds(handle, name);
}

// directly linked users
HOOK_EXPORT CUresult cuLaunchKernel(LAUNCH_ARGS) { return launch_hook<kLaunch>(f, gx, gy, gz, bx, by, bz, smem, st, params, extra); }
HOOK_EXPORT CUresult cuLaunchKernel_ptsz(LAUNCH_ARGS) { return launch_hook<kLaunchPtsz>(f, gx, gy, gz, bx, by, bz, smem, st, params, extra); }
HOOK_EXPORT CUresult cuLaunchKernelEx(const void* c, CUfunction f, void** p, void** e) { return launch_ex_hook<kLaunchEx>(c, f, p, e); }
HOOK_EXPORT CUresult cuLaunchKernelEx_ptsz(const void* c, CUfunction f, void** p, void** e) { return launch_ex_hook<kLaunchExPtsz>(c, f, p, e); }
HOOK_EXPORT CUresult cuLaunchCooperativeKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                               unsigned smem, CUstream st, void** params) {
  return launch_coop_hook<kLaunchCoop>(f, gx, gy, gz, bx, by, bz, smem, st, params);
}
HOOK_EXPORT CUresult cuLaunchCooperativeKernel_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                                    unsigned bz, unsigned smem, CUstream st, void** params) {
  return launch_coop_hook<kLaunchCoopPtsz>(f, gx, gy, gz, bx, by, bz, smem, st, params);
}
HOOK_EXPORT CUresult cuMemAlloc_v2(CUdeviceptr* p, size_t n) { return mem_alloc_hook(p, n); }
HOOK_EXPORT CUresult cuMemAllocManaged(CUdeviceptr* p, size_t n, unsigned fl) { return mem_alloc_managed_hook(p, n, fl); }
HOOK_EXPORT CUresult cuMemAllocPitch_v2(CUdeviceptr* p, size_t* pitch, size_t w, size_t h, unsigned e) { return mem_alloc_pitch_hook(p, pitch, w, h, e); }
HOOK_EXPORT CUresult cuMemFree_v2(CUdeviceptr p) { return mem_free_hook(p); }
HOOK_EXPORT CUresult cuMemAllocAsync(CUdeviceptr* p, size_t n, CUstream s) { return mem_alloc_async_hook<kMemAllocAsync>(p, n, s); }
HOOK_EXPORT CUresult cuMemAllocAsync_ptsz(CUdeviceptr* p, size_t n, CUstream s) { return mem_alloc_async_hook<kMemAllocAsyncPtsz>(p, n, s); }
HOOK_EXPORT CUresult cuMemAllocFromPoolAsync(CUdeviceptr* p, size_t n, CUmemoryPool pool, CUstream s) {
  return mem_alloc_from_pool_async_hook<kMemAllocFromPoolAsync>(p, n, pool, s);
}
HOOK_EXPORT CUresult cuMemAllocFromPoolAsync_ptsz(CUdeviceptr* p, size_t n, CUmemoryPool pool, CUstream s) {
  return mem_alloc_from_pool_async_hook<kMemAllocFromPoolAsyncPtsz>(p, n, pool, s);
}
HOOK_EXPORT CUresult cuMemFreeAsync(CUdeviceptr p, CUstream s) { return mem_free_async_hook<kMemFreeAsync>(p, s); }
HOOK_EXPORT CUresult cuMemFreeAsync_ptsz(CUdeviceptr p, CUstream s) { return mem_free_async_hook<kMemFreeAsyncPtsz>(p, s); }
HOOK_EXPORT CUresult cuMemCreate(CUmemGenericAllocationHandle* h, size_t n, const void* prop, unsigned long long fl) { return mem_create_hook(h, n, prop, fl); }
HOOK_EXPORT CUresult cuMemRelease(CUmemGenericAllocationHandle h) { return mem_release_hook(h); }
HOOK_EXPORT CUresult cuMemGetInfo_v2(size_t* f, size_t* t) { return mem_get_info_hook(f, t); }
HOOK_EXPORT CUresult cuDeviceTotalMem_v2(size_t* b, CUdevice d) { return device_total_mem_hook(b, d); }
HOOK_EXPORT CUresult cuGraphInstantiateWithFlags(CUgraphExec* e, CUgraph g, unsigned long long fl) { return graph_instantiate_flags_hook(e, g, fl); }
HOOK_EXPORT CUresult cuGraphInstantiateWithParams(CUgraphExec* e, CUgraph g, void* p) { return graph_instantiate_params_hook<kGraphInstantiateWithParams>(e, g, p); }
HOOK_EXPORT CUresult cuGraphInstantiateWithParams_ptsz(CUgraphExec* e, CUgraph g, void* p) { return graph_instantiate_params_hook<kGraphInstantiateWithParamsPtsz>(e, g, p); }
HOOK_EXPORT CUresult cuGraphLaunch(CUgraphExec e, CUstream s) { return graph_launch_hook<kGraphLaunch>(e, s); }
HOOK_EXPORT CUresult cuGraphLaunch_ptsz(CUgraphExec e, CUstream s) { return graph_launch_hook<kGraphLaunchPtsz>(e, s); }
HOOK_EXPORT CUresult cuGraphExecUpdate_v2(CUgraphExec e, CUgraph g, void* info) { return graph_exec_update_hook(e, g, info); }
HOOK_EXPORT CUresult cuGraphExecDestroy(CUgraphExec e) { return graph_exec_destroy_hook(e); }
HOOK_EXPORT CUresult cuGetProcAddress(const char* s, void** pfn, int v, cuuint64_t fl) { return get_proc_address_hook(s, pfn, v, fl); }
HOOK_EXPORT CUresult cuGetProcAddress_v2(const char* s, void** pfn, int v, cuuint64_t fl, void* st) { return get_proc_address_v2_hook(s, pfn, v, fl, st); }

// counters for tests and for the worker's stats file
struct tf_hook_stats {
  uint64_t launches, blocked, timeouts, wait_ns, tokens, denied_allocs, active;
};
HOOK_EXPORT void tf_hook_get_stats(tf_hook_stats* out) {
  if (!out) return;
  out->launches = g_launches.load();
  out->blocked = g_blocked.load();
  out->timeouts = g_timeouts.load();
  out->wait_ns = g_wait_ns.load();
  out->tokens = g_tokens.load();
  out->denied_allocs = g_denied_allocs.load();
  out->active = g_cfg.active ? 1 : 0;
}
