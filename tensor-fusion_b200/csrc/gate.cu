// gate.cu -- device-resident ERL token bucket: gate / refill kernels + C-ABI.
//
// Arithmetic restated from pkg/hypervisor/worker/state/soft_limiter_shm.go:
//   FetchSubERLTokens :715-731   deny (no mutation) if current < cost, else
//                                CAS(current -> max(0, current - cost))
//   FetchAddERLTokens :734-748   CAS(current -> max(0, min(capacity, current + amount)))
// float64 values travel as u64 bit patterns; Go's math.Max/Min NaN rules are
// reproduced by go_max/go_min (CUDA's fmax/fmin drop NaNs, Go propagates them).
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <deque>
#include <mutex>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <thread>

#include "gate.h"
#include "kernels.h"
#include "quota_bridge.h"

namespace tfw {

__device__ __forceinline__ double go_max(double a, double b) {
  if (a != a || b != b) return __longlong_as_double(0x7FF8000000000001ll);
  if (a == 0.0 && b == 0.0) return (__double_as_longlong(a) < 0) ? b : a;  // Max(-0,+0) = +0
  return a > b ? a : b;
}
__device__ __forceinline__ double go_min(double a, double b) {
  if (a != a || b != b) return __longlong_as_double(0x7FF8000000000001ll);
  if (a == 0.0 && b == 0.0) return (__double_as_longlong(a) < 0) ? a : b;  // Min(-0,+0) = -0
  return a < b ? a : b;
}

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// returns the value found; *ok tells whether tokens were taken
__device__ double bucket_fetch_sub(DevBucket* b, double cost, bool* ok) {
  for (;;) {
    const unsigned long long cur_bits = ld_acquire(&b->tokens);
    const double cur = __longlong_as_double((long long)cur_bits);
    if (cur < cost) { *ok = false; return cur; }
    const double nv = go_max(0.0, __dsub_rn(cur, cost));
    if (atomicCAS(&b->tokens, cur_bits, (unsigned long long)__double_as_longlong(nv)) == cur_bits) { *ok = true; return cur; }
  }
}

__device__ double bucket_fetch_add(DevBucket* b, double amount) {
  const double cap = __longlong_as_double((long long)ld_acquire(&b->capacity));
  for (;;) {
    const unsigned long long cur_bits = ld_acquire(&b->tokens);
    const double cur = __longlong_as_double((long long)cur_bits);
    const double nv = go_max(0.0, go_min(cap, __dadd_rn(cur, amount)));
    if (atomicCAS(&b->tokens, cur_bits, (unsigned long long)__double_as_longlong(nv)) == cur_bits) return cur;
  }
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Blocking gate: one thread; spins (with back-off) until the bucket admits `cost`.
__global__ void tfw_gate_block(DevBucket* b, double cost, unsigned long long* mirror) {
  const unsigned long long t0 = gtimer();
  bool ok = false, waited = false;
  unsigned ns = 100;
  for (;;) {
    const double before = bucket_fetch_sub(b, cost, &ok);
    (void)before;
    if (ok) break;
    atomicAdd(&b->denied, 1ull);
    waited = true;
    if (gtimer() - t0 > b->max_wait_ns) { atomicAdd(&b->timeouts, 1ull); break; }  // fail open, never hang the GPU
    __nanosleep(ns);
    if (ns < 2000) ns *= 2;
  }
  if (ok) atomicAdd(&b->admitted, 1ull);
  if (waited) {
    atomicAdd(&b->blocked, 1ull);
    atomicAdd(&b->wait_ns, gtimer() - t0);
  }
  if (mirror) *mirror = ld_acquire(&b->tokens);
}

// Take kernel of the stream-wait gate: the stream has already waited (cuStreamWaitValue64)
// until tokens >= cost, so the first FetchSub normally succeeds.  If another consumer raced
// us to the tokens it degrades to the bounded spin above.  mirror[1] counts completed gates.
__global__ void tfw_gate_take(DevBucket* b, double cost, unsigned long long* mirror) {
  const unsigned long long t0 = gtimer();
  bool ok = false;
  unsigned ns = 100;
  for (;;) {
    bucket_fetch_sub(b, cost, &ok);
    if (ok) break;
    atomicAdd(&b->denied, 1ull);
    if (gtimer() - t0 > b->max_wait_ns) { atomicAdd(&b->timeouts, 1ull); break; }
    __nanosleep(ns);
    if (ns < 2000) ns *= 2;
  }
  if (ok) atomicAdd(&b->admitted, 1ull);
  if (mirror) {
    mirror[0] = ld_acquire(&b->tokens);
    mirror[1] = ld_acquire(&b->admitted) + ld_acquire(&b->timeouts);
  }
}

// fail-open of a gate whose refill never came: make exactly its cost available
__global__ void tfw_gate_force_k(DevBucket* b, double cost, unsigned long long* mirror) {
  for (;;) {
    const unsigned long long cur_bits = ld_acquire(&b->tokens);
    const double cur = __longlong_as_double((long long)cur_bits);
    if (cur >= cost) break;
    if (atomicCAS(&b->tokens, cur_bits, (unsigned long long)__double_as_longlong(cost)) == cur_bits) break;
  }
  // counted apart from `timeouts`: the released take kernel bumps `admitted`, and mirror[1] = admitted +
  // timeouts must count every gate exactly once (the watchdog tracks the head of the stream with it)
  atomicAdd(&b->forced, 1ull);
  if (mirror) mirror[0] = ld_acquire(&b->tokens);
}

__global__ void tfw_gate_try_k(DevBucket* b, double cost, double* before, int* admitted, unsigned long long* mirror) {
  bool ok = false;
  *before = bucket_fetch_sub(b, cost, &ok);
  *admitted = ok ? 1 : 0;
  atomicAdd(ok ? &b->admitted : &b->denied, 1ull);
  if (mirror) *mirror = ld_acquire(&b->tokens);
}

__global__ void tfw_gate_refill_k(DevBucket* b, double amount, double* before, unsigned long long* mirror) {
  const double v = bucket_fetch_add(b, amount);
  if (before) *before = v;
  if (mirror) *mirror = ld_acquire(&b->tokens);
}

__global__ void tfw_gate_set_k(DevBucket* b, int what, double v, unsigned long long* mirror) {
  if (what == 0) atomicExch(&b->capacity, (unsigned long long)__double_as_longlong(v));
  else atomicExch(&b->tokens, (unsigned long long)__double_as_longlong(v));
  if (mirror) *mirror = ld_acquire(&b->tokens);
}

__global__ void tfw_gate_seq_k(DevBucket* b, const tfw_gate_op* ops, uint32_t n, double* before) {
  for (uint32_t i = 0; i < n; ++i) {
    bool ok;
    switch (ops[i].kind) {
      case 0: before[i] = bucket_fetch_sub(b, ops[i].amount, &ok); break;
      case 1: before[i] = bucket_fetch_add(b, ops[i].amount); break;
      case 2: before[i] = __longlong_as_double((long long)atomicExch(&b->capacity, (unsigned long long)__double_as_longlong(ops[i].amount))); break;
      default: before[i] = __longlong_as_double((long long)atomicExch(&b->tokens, (unsigned long long)__double_as_longlong(ops[i].amount))); break;
    }
  }
}

__global__ void tfw_gate_contend_k(DevBucket* b, uint32_t per_thread, double cost, unsigned long long* admitted) {
  unsigned long long mine = 0;
  for (uint32_t i = 0; i < per_thread; ++i) {
    bool ok;
    bucket_fetch_sub(b, cost, &ok);
    mine += ok ? 1 : 0;
  }
  atomicAdd(admitted, mine);
}

cudaError_t preload_gate_kernels() {
  cudaFuncAttributes a;
  const void* fns[] = {(const void*)tfw_gate_block, (const void*)tfw_gate_take, (const void*)tfw_gate_force_k, (const void*)tfw_gate_try_k, (const void*)tfw_gate_refill_k,
                       (const void*)tfw_gate_set_k, (const void*)tfw_gate_seq_k, (const void*)tfw_gate_contend_k};
  for (const void* f : fns) {
    cudaError_t e = cudaFuncGetAttributes(&a, f);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace tfw

struct tfw_gate {
  int device = 0;
  tfw::DevBucket* bucket = nullptr;
  unsigned long long* mirror_host = nullptr;  // pinned + mapped: last token value the GPU published
  unsigned long long* mirror_dev = nullptr;
  cudaStream_t side = nullptr;                // refill / query stream (never the vGPU exec stream)
  double* d_scratch = nullptr;                // before / admitted outputs
  tfw::QuotaBridge* bridge = nullptr;
  double refill_rate = 0.0;
  std::string err;
  // stream-wait gate: the wait itself is a stream memory operation, not a kernel, so a
  // throttled vGPU does not show up as GPU utilisation (NVML counts a spinning gate kernel as
  // 100 % busy, which drives the hypervisor's PID loop to the minimum rate).
  CUresult (*wait_value64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;
  std::mutex mu;
  std::deque<std::pair<double, std::chrono::steady_clock::time_point>> pending;  // cost, enqueue time
  uint64_t enqueued = 0, forced = 0;
  std::atomic<uint64_t> host_blocked{0};
  std::thread watchdog;
  std::atomic<bool> stop{false};
  double max_wait_s = 5.0;
  bool fail_closed = false;                   // TFW_GATE_FAIL_POLICY=closed: a starved gate waits for its refill, however long
  std::atomic<double> capacity_host{100.0};   // capacity of the device bucket as last set from the host
};

namespace {
// Fail-open watchdog: a gate that has been at the head of the queue for max_wait_s without
// its refill arriving is released (counted in `timeouts`), so a dead hypervisor or a cost
// above the bucket capacity can never wedge the vGPU stream.
void gate_watchdog(tfw_gate* g) {
  cudaSetDevice(g->device);
  while (!g->stop.load(std::memory_order_acquire)) {
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    const uint64_t done = const_cast<volatile unsigned long long*>(g->mirror_host)[1];
    std::lock_guard<std::mutex> lk(g->mu);
    while (!g->pending.empty() && g->enqueued - g->pending.size() < done) g->pending.pop_front();
    if (g->pending.empty()) continue;
    const auto age = std::chrono::duration<double>(std::chrono::steady_clock::now() - g->pending.front().second).count();
    if (age > g->max_wait_s && !g->fail_closed) {
      tfw::tfw_gate_force_k<<<1, 1, 0, g->side>>>(g->bucket, g->pending.front().first, g->mirror_dev);
      cudaGetLastError();
      g->pending.front().second = std::chrono::steady_clock::now();  // give the released gate time to run
      g->forced++;
    }
  }
}
}  // namespace

extern "C" {

#define G_OK(call) do { if ((call) != cudaSuccess) { cudaGetLastError(); return TFW_ERR_FAILED; } } while (0)

tfw_status tfw_gate_create(int device, const char* shm_path, uint32_t device_index, tfw_gate** out) {
  if (!out) return TFW_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return TFW_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return TFW_ERR_INVALID;
  tfw_gate* g = new (std::nothrow) tfw_gate();
  if (!g) return TFW_ERR_EXHAUSTED;
  g->device = device;
  auto bail = [&](tfw_status s) { tfw_gate_destroy(g); return s; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (tfw::preload_gate_kernels() != cudaSuccess || tfw::preload_kernels() != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (cudaMalloc(reinterpret_cast<void**>(&g->bucket), sizeof(tfw::DevBucket)) != cudaSuccess) return bail(TFW_ERR_EXHAUSTED);
  tfw::DevBucket init{};
  const double hundred = 100.0;  // quota-file defaults, soft_limiter_shm.go:186-189
  std::memcpy(&init.tokens, &hundred, 8);
  std::memcpy(&init.capacity, &hundred, 8);
  init.max_wait_ns = 5ull * 1000 * 1000 * 1000;
  if (cudaMemcpy(g->bucket, &init, sizeof(init), cudaMemcpyHostToDevice) != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (cudaHostAlloc(reinterpret_cast<void**>(&g->mirror_host), 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return bail(TFW_ERR_EXHAUSTED);
  std::memset(g->mirror_host, 0, 64);
  std::memcpy(g->mirror_host, &hundred, 8);
  if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&g->mirror_dev), g->mirror_host, 0) != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking) != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (cudaMalloc(reinterpret_cast<void**>(&g->d_scratch), 64) != cudaSuccess) return bail(TFW_ERR_EXHAUSTED);
  {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    int can = 0;
    // CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS (122): the runtime enum has no name for it
    if (cudaDeviceGetAttribute(&can, static_cast<cudaDeviceAttr>(122), device) != cudaSuccess) { cudaGetLastError(); can = 1; }
    if (can && !getenv("TFW_GATE_SPIN") && cudaGetDriverEntryPoint("cuStreamWaitValue64", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g->wait_value64 = reinterpret_cast<decltype(g->wait_value64)>(fn);
    else
      cudaGetLastError();
    if (const char* e = getenv("TFW_GATE_MAX_WAIT_MS")) { double v = atof(e); if (v > 0) g->max_wait_s = v / 1000.0; }
    // What happens to a launch whose tokens never come (dead hypervisor, stalled controller):
    //   open   (default) the watchdog releases it after max_wait -- availability over isolation;
    //   closed the gate waits for its refill, however long -- isolation over availability.
    if (const char* e = getenv("TFW_GATE_FAIL_POLICY")) g->fail_closed = !strcmp(e, "closed");
    g->watchdog = std::thread(gate_watchdog, g);
  }
  if (shm_path) {
    tfw_status s = tfw::quota_bridge_start(g, shm_path, device_index, &g->bridge);
    if (s != TFW_OK) return bail(s);
  }
  *out = g;
  return TFW_OK;
}

tfw_status tfw_gate_destroy(tfw_gate* g) {
  if (!g) return TFW_ERR_INVALID;
  cudaSetDevice(g->device);
  g->stop.store(true, std::memory_order_release);
  if (g->watchdog.joinable()) g->watchdog.join();
  if (g->bridge) tfw::quota_bridge_stop(g->bridge);
  if (g->side) { cudaStreamSynchronize(g->side); cudaStreamDestroy(g->side); }
  if (g->bucket) cudaFree(g->bucket);
  if (g->d_scratch) cudaFree(g->d_scratch);
  if (g->mirror_host) cudaFreeHost(g->mirror_host);
  delete g;
  return TFW_OK;
}

tfw_status tfw_gate_set_policy(tfw_gate* g, int fail_closed, double max_wait_ms) {
  if (!g) return TFW_ERR_INVALID;
  std::lock_guard<std::mutex> lk(g->mu);
  g->fail_closed = fail_closed != 0;
  if (max_wait_ms > 0) g->max_wait_s = max_wait_ms / 1000.0;
  return TFW_OK;
}

tfw_status tfw_gate_try(tfw_gate* g, double cost, double* before, int* admitted) {
  if (!g || !before || !admitted) return TFW_ERR_INVALID;
  cudaSetDevice(g->device);
  tfw::tfw_gate_try_k<<<1, 1, 0, g->side>>>(g->bucket, cost, g->d_scratch, reinterpret_cast<int*>(g->d_scratch + 1), g->mirror_dev);
  G_OK(cudaGetLastError());
  double host[2];
  G_OK(cudaMemcpyAsync(host, g->d_scratch, sizeof(host), cudaMemcpyDeviceToHost, g->side));
  G_OK(cudaStreamSynchronize(g->side));
  *before = host[0];
  int a;
  std::memcpy(&a, &host[1], sizeof(int));
  *admitted = a;
  return TFW_OK;
}

tfw_status tfw_gate_enqueue(tfw_gate* g, double cost, void* cuda_stream) {
  if (!g || !(cost >= 0.0)) return TFW_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  // A launch can cost more than the bucket will ever hold (the reference controller keeps capacity =
  // clamp(rate * 0.5 s, 200, 200000), quota_controller.go:425-433, while blocks x warps of one large
  // grid exceeds 10^5): FetchSub never admits cost > capacity, so such a launch would sit at the gate
  // until the watchdog.  It is charged one full bucket instead -- the most the bucket can express.
  if (g->bridge) {
    const double file_cap = tfw::quota_bridge_capacity(g->bridge);
    if (file_cap > 0.0 && cost > file_cap) cost = file_cap;
    tfw::quota_bridge_note_cost(g->bridge, cost);  // the prepaid window grows to hold it
  } else {
    const double cap = g->capacity_host.load(std::memory_order_relaxed);
    if (cap > 0.0 && cost > cap) cost = cap;
  }
  if (!g->wait_value64) {  // no stream memory operations on this device/driver: bounded spin gate
    tfw::tfw_gate_block<<<1, 1, 0, st>>>(g->bucket, cost, g->mirror_dev);
    G_OK(cudaGetLastError());
    return TFW_OK;
  }
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->pending.emplace_back(cost, std::chrono::steady_clock::now());
    g->enqueued++;
  }
  double seen;
  std::memcpy(&seen, const_cast<const unsigned long long*>(g->mirror_host), 8);
  if (seen < cost) g->host_blocked.fetch_add(1, std::memory_order_relaxed);  // estimate: would have to wait
  uint64_t bits;
  std::memcpy(&bits, &cost, 8);  // non-negative float64 bit patterns order like unsigned integers
  // GEQ waits until (int64)(*addr - value) >= 0: exact for two non-negative float64 bit patterns
  if (g->wait_value64(reinterpret_cast<CUstream>(st), reinterpret_cast<CUdeviceptr>(&g->bucket->tokens), bits, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
    g->wait_value64 = nullptr;  // driver refuses stream memory operations: the take kernel's bounded spin still gates
  tfw::tfw_gate_take<<<1, 1, 0, st>>>(g->bucket, cost, g->mirror_dev);
  G_OK(cudaGetLastError());
  return TFW_OK;
}

tfw_status tfw_gate_refill(tfw_gate* g, double amount, double* before) {
  if (!g) return TFW_ERR_INVALID;
  cudaSetDevice(g->device);
  tfw::tfw_gate_refill_k<<<1, 1, 0, g->side>>>(g->bucket, amount, before ? g->d_scratch : nullptr, g->mirror_dev);
  G_OK(cudaGetLastError());
  if (before) {
    G_OK(cudaMemcpyAsync(before, g->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, g->side));
    G_OK(cudaStreamSynchronize(g->side));
  }
  return TFW_OK;
}

static tfw_status gate_set(tfw_gate* g, int what, double v) {
  if (!g) return TFW_ERR_INVALID;
  cudaSetDevice(g->device);
  tfw::tfw_gate_set_k<<<1, 1, 0, g->side>>>(g->bucket, what, v, g->mirror_dev);
  G_OK(cudaGetLastError());
  G_OK(cudaStreamSynchronize(g->side));
  return TFW_OK;
}
tfw_status tfw_gate_set_capacity(tfw_gate* g, double capacity) {
  tfw_status s = gate_set(g, 0, capacity);
  if (s == TFW_OK) g->capacity_host.store(capacity, std::memory_order_relaxed);
  return s;
}
tfw_status tfw_gate_set_tokens(tfw_gate* g, double tokens) { return gate_set(g, 1, tokens); }

tfw_status tfw_gate_get_state(tfw_gate* g, tfw_gate_state* out) {
  if (!g || !out) return TFW_ERR_INVALID;
  cudaSetDevice(g->device);
  tfw::DevBucket b{};
  G_OK(cudaMemcpyAsync(&b, g->bucket, sizeof(b), cudaMemcpyDeviceToHost, g->side));
  G_OK(cudaStreamSynchronize(g->side));
  std::memcpy(&out->tokens, &b.tokens, 8);
  std::memcpy(&out->capacity, &b.capacity, 8);
  out->refill_rate = g->bridge ? tfw::quota_bridge_rate(g->bridge) : 0.0;
  out->admitted = b.admitted;
  out->denied = b.denied;
  out->blocked_gates = b.blocked + g->host_blocked.load(std::memory_order_relaxed);
  out->wait_ns = b.wait_ns;
  out->bridged_tokens_milli = g->bridge ? tfw::quota_bridge_moved_milli(g->bridge) : 0;
  out->timeouts = b.timeouts + b.forced;
  return TFW_OK;
}

tfw_status tfw_gate_run_sequence(tfw_gate* g, const tfw_gate_op* ops, uint32_t n, double* before) {
  if (!g || !ops || !before || !n) return TFW_ERR_INVALID;
  cudaSetDevice(g->device);
  tfw_gate_op* d_ops = nullptr;
  double* d_before = nullptr;
  G_OK(cudaMalloc(reinterpret_cast<void**>(&d_ops), sizeof(tfw_gate_op) * n));
  if (cudaMalloc(reinterpret_cast<void**>(&d_before), sizeof(double) * n) != cudaSuccess) { cudaFree(d_ops); cudaGetLastError(); return TFW_ERR_EXHAUSTED; }
  tfw_status rc = TFW_OK;
  if (cudaMemcpyAsync(d_ops, ops, sizeof(tfw_gate_op) * n, cudaMemcpyHostToDevice, g->side) != cudaSuccess) rc = TFW_ERR_FAILED;
  if (rc == TFW_OK) {
    tfw::tfw_gate_seq_k<<<1, 1, 0, g->side>>>(g->bucket, d_ops, n, d_before);
    if (cudaGetLastError() != cudaSuccess) rc = TFW_ERR_FAILED;
  }
  if (rc == TFW_OK && cudaMemcpyAsync(before, d_before, sizeof(double) * n, cudaMemcpyDeviceToHost, g->side) != cudaSuccess) rc = TFW_ERR_FAILED;
  if (cudaStreamSynchronize(g->side) != cudaSuccess) rc = TFW_ERR_FAILED;
  cudaFree(d_ops);
  cudaFree(d_before);
  if (rc != TFW_OK) cudaGetLastError();
  return rc;
}

tfw_status tfw_gate_contend(tfw_gate* g, uint32_t nthreads, uint32_t per_thread, double cost, uint64_t* admitted) {
  if (!g || !admitted || !nthreads || !per_thread) return TFW_ERR_INVALID;
  cudaSetDevice(g->device);
  unsigned long long* d = reinterpret_cast<unsigned long long*>(g->d_scratch + 4);
  G_OK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), g->side));
  tfw::tfw_gate_contend_k<<<nthreads, 1, 0, g->side>>>(g->bucket, per_thread, cost, d);
  G_OK(cudaGetLastError());
  unsigned long long h = 0;
  G_OK(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, g->side));
  G_OK(cudaStreamSynchronize(g->side));
  *admitted = h;
  return TFW_OK;
}

}  // extern "C"

// accessors used by the bridge (quota_bridge.cc is plain C++ and cannot see the struct)
namespace tfw {
int gate_device(tfw_gate* g) { return g->device; }
double gate_mirror_tokens(tfw_gate* g) { double v; std::memcpy(&v, const_cast<const unsigned long long*>(g->mirror_host), 8); return v; }
}  // namespace tfw
