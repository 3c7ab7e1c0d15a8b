// copy_lab.cu -- tuning lab for the aligned-copy inner loop of the byte mover.
// Prints GB/s (read+write bytes / CUDA-event time) for a matrix of kernel shapes
// on a 4 GiB device-to-device copy.  Build: see tools/Makefile target copy_lab.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <functional>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum { LD_NC_NA = 0, LD_PLAIN = 1, LD_EF = 2 };
enum { ST_NA = 0, ST_PLAIN = 1, ST_EF = 2, ST_CS = 3 };

template <int MODE> __device__ __forceinline__ int4 ld16(const void* p, uint64_t pol) {
  int4 r;
  if (MODE == LD_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if (MODE == LD_PLAIN) asm volatile("ld.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
template <int MODE> __device__ __forceinline__ void st16(void* p, const int4& v, uint64_t pol) {
  if (MODE == ST_NA) asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else if (MODE == ST_PLAIN) asm volatile("st.global.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else if (MODE == ST_CS) asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// persistent (grid-stride over tiles) or one-shot (grid == ntiles): same code
template <int THREADS, int U, int MINB, int LD, int ST>
__global__ void __launch_bounds__(THREADS, MINB) k_copy(const uint8_t* __restrict__ s, uint8_t* __restrict__ d, uint64_t ntiles) {
  const uint64_t pol = (LD == LD_EF || ST == ST_EF) ? evict_first_policy() : 0;
  constexpr uint64_t TILE = (uint64_t)THREADS * U * 16;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint8_t* sp = s + t * TILE + threadIdx.x * 16;
    uint8_t* dp = d + t * TILE + threadIdx.x * 16;
    int4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld16<LD>(sp + (size_t)u * THREADS * 16, pol);
#pragma unroll
    for (int u = 0; u < U; ++u) st16<ST>(dp + (size_t)u * THREADS * 16, v[u], pol);
  }
}

// ---- TMA bulk pipeline ------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGE_BYTES, int STAGES, int LOOK>
__global__ void __launch_bounds__(32) k_tma(const uint8_t* __restrict__ s, uint8_t* __restrict__ d, uint64_t nchunks) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[STAGES];
  if (threadIdx.x != 0) return;
  const uint32_t sb = s32(smem), bb = s32(bars);
  for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb + 8 * i));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  uint32_t nl = 0, ns = 0;
  auto store_one = [&](uint64_t chunk) {
    const uint32_t st = ns % STAGES, par = (ns / STAGES) & 1;
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bb + 8 * st), "r"(par) : "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + chunk * STAGE_BYTES), "r"(sb + st * STAGE_BYTES), "r"(STAGE_BYTES) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    ++ns;
  };
  // chunks handled by this CTA: c = blockIdx.x + k*gridDim.x
  uint64_t pending[STAGES];
  for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const uint32_t st = nl % STAGES;
    if (nl >= STAGES) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1 - LOOK) : "memory");
    pending[st] = c;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb + 8 * st), "r"(STAGE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + st * STAGE_BYTES), "l"(s + c * STAGE_BYTES), "r"(STAGE_BYTES), "r"(bb + 8 * st) : "memory");
    ++nl;
    if (nl - ns > LOOK) store_one(pending[ns % STAGES]);
  }
  while (ns < nl) store_one(pending[ns % STAGES]);
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- TMA one-shot: grid == tiles, each CTA moves ONE tile with one bulk load and one bulk store.
// Like the one-shot LDG kernel, the hardware CTA scheduler hands tiles out in address order, so the set
// of tiles in flight is one contiguous DRAM window; unlike it, no register file is involved at all.
template <int TILE>
__global__ void __launch_bounds__(32) k_tma_oneshot(const uint8_t* __restrict__ s, uint8_t* __restrict__ d) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x != 0) return;
  const uint32_t sb = s32(smem), bb = s32(&bar);
  const uint64_t off = (uint64_t)blockIdx.x * TILE;
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(TILE) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb), "l"(s + off), "r"(TILE), "r"(bb) : "memory");
  asm volatile("{\n.reg .pred p;\nW1: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(bb) : "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + off), "r"(sb), "r"(TILE) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the CTA's shared memory may go once it has been read
}

#include <functional>
#include <string>
static float timed(cudaStream_t st, const std::function<void()>& f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  std::vector<float> ts;
  for (int i = 0; i < 7; ++i) {
    CK(cudaEventRecord(a, st)); f(); CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ts.push_back(ms);
  }
  CK(cudaGetLastError());
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

template <int THREADS, int U, int MINB, int LD, int ST>
void run_copy(const char* name, const uint8_t* s, uint8_t* d, uint64_t bytes, int sms, int ctas /*0 = one-shot*/, cudaStream_t st) {
  const uint64_t TILE = (uint64_t)THREADS * U * 16;
  const uint64_t ntiles = bytes / TILE;
  const unsigned grid = ctas ? (unsigned)(sms * ctas) : (unsigned)ntiles;
  float ms = timed(st, [&] { k_copy<THREADS, U, MINB, LD, ST><<<grid, THREADS, 0, st>>>(s, d, ntiles); });
  printf("{\"kernel\":\"%s\",\"threads\":%d,\"unroll\":%d,\"minb\":%d,\"ld\":%d,\"st\":%d,\"ctas_per_sm\":%d,\"ms\":%.4f,\"GBps\":%.1f}\n",
         name, THREADS, U, MINB, LD, ST, ctas, ms, 2.0 * bytes / (ms * 1e-3) / 1e9);
  fflush(stdout);
}
template <int SB, int STAGES, int LOOK>
void run_tma(const uint8_t* s, uint8_t* d, uint64_t bytes, int sms, int ctas, cudaStream_t st) {
  const int smem = SB * STAGES;
  CK(cudaFuncSetAttribute(k_tma<SB, STAGES, LOOK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const uint64_t nchunks = bytes / SB;
  float ms = timed(st, [&] { k_tma<SB, STAGES, LOOK><<<sms * ctas, 32, smem, st>>>(s, d, nchunks); });
  printf("{\"kernel\":\"tma\",\"stage_bytes\":%d,\"stages\":%d,\"look\":%d,\"ctas_per_sm\":%d,\"ms\":%.4f,\"GBps\":%.1f}\n", SB, STAGES, LOOK, ctas, ms,
         2.0 * bytes / (ms * 1e-3) / 1e9);
  fflush(stdout);
}

template <int TILE>
void run_tma_oneshot(const uint8_t* s, uint8_t* d, uint64_t bytes, cudaStream_t st) {
  CK(cudaFuncSetAttribute(k_tma_oneshot<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE));
  const unsigned grid = (unsigned)(bytes / TILE);
  float ms = timed(st, [&] { k_tma_oneshot<TILE><<<grid, 32, TILE, st>>>(s, d); });
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tma_oneshot<TILE>, 32, TILE));
  printf("{\"kernel\":\"tma_oneshot\",\"tile_bytes\":%d,\"ctas_per_sm_resident\":%d,\"ms\":%.4f,\"GBps\":%.1f}\n", TILE, occ, ms, 2.0 * bytes / (ms * 1e-3) / 1e9);
  fflush(stdout);
}

static void peer_lab(const char* tag, const uint8_t* s, uint8_t* d, uint64_t bytes, int sms, cudaStream_t st) {
  printf("{\"section\":\"%s\"}\n", tag);
  { float ms = timed(st, [&] { CK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDefault, st)); });
    printf("{\"kernel\":\"cudaMemcpyAsync\",\"ms\":%.4f,\"GBps_one_way\":%.1f}\n", ms, (double)bytes / (ms * 1e-3) / 1e9); }
  // note: GBps printed by run_copy/run_tma counts read+write (2x); one-way NVLink rate is half of it
  run_copy<256, 8, 3, LD_NC_NA, ST_NA>("oneshot", s, d, bytes, sms, 0, st);
  run_copy<256, 4, 6, LD_NC_NA, ST_NA>("oneshot", s, d, bytes, sms, 0, st);
  run_copy<256, 2, 8, LD_NC_NA, ST_NA>("oneshot", s, d, bytes, sms, 0, st);
  run_copy<256, 8, 3, LD_PLAIN, ST_PLAIN>("oneshot", s, d, bytes, sms, 0, st);
  run_copy<256, 8, 3, LD_NC_NA, ST_CS>("oneshot", s, d, bytes, sms, 0, st);
  for (int c : {1, 2, 3, 4}) run_copy<256, 8, 3, LD_NC_NA, ST_NA>("persist", s, d, bytes, sms, c, st);
  for (int c : {2, 4}) run_copy<256, 16, 2, LD_NC_NA, ST_NA>("persist", s, d, bytes, sms, c / 2, st);
  run_tma_oneshot<16384>(s, d, bytes, st);
  run_tma_oneshot<32768>(s, d, bytes, st);
  run_tma<16384, 6, 4>(s, d, bytes, sms, 1, st);
  run_tma<16384, 6, 4>(s, d, bytes, sms, 2, st);
  run_tma<32768, 3, 1>(s, d, bytes, sms, 1, st);
  run_tma<32768, 6, 4>(s, d, bytes, sms, 1, st);
  run_tma<65536, 3, 1>(s, d, bytes, sms, 1, st);
  run_tma<8192, 8, 6>(s, d, bytes, sms, 4, st);
}

int main(int argc, char** argv) {
  const uint64_t bytes = 4ull << 30;
  uint8_t *s, *d;
  CK(cudaMalloc(&s, bytes)); CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(s, 1, bytes)); CK(cudaMemset(d, 2, bytes));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  if (argc > 1 && std::string(argv[1]) == "peer") {
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("{\"error\":\"peer mode needs 2 GPUs\"}\n"); return 0; }
    uint8_t* r;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&r, bytes)); CK(cudaMemset(r, 3, bytes)); CK(cudaDeviceSynchronize());
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    peer_lab("push: local HBM -> peer HBM, kernel on the source GPU", s, r, bytes, sms, st);
    peer_lab("pull: peer HBM -> local HBM, kernel on the destination GPU", r, d, bytes, sms, st);
    return 0;
  }
  { float ms = timed(st, [&] { CK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, st)); });
    printf("{\"kernel\":\"cudaMemcpyAsync_D2D\",\"ms\":%.4f,\"GBps\":%.1f}\n", ms, 2.0 * bytes / (ms * 1e-3) / 1e9); }
  // persistent, 256 threads
  for (int c : {2, 3, 4}) run_copy<256, 8, 3, LD_NC_NA, ST_NA>("persist", s, d, bytes, sms, c, st);
  for (int c : {2, 4, 6}) run_copy<256, 4, 6, LD_NC_NA, ST_NA>("persist", s, d, bytes, sms, c, st);
  for (int c : {4, 8}) run_copy<256, 2, 8, LD_NC_NA, ST_NA>("persist", s, d, bytes, sms, c, st);
  for (int c : {1, 2}) run_copy<256, 16, 2, LD_NC_NA, ST_NA>("persist", s, d, bytes, sms, c, st);
  for (int c : {1, 2}) run_copy<512, 8, 2, LD_NC_NA, ST_NA>("persist", s, d, bytes, sms, c, st);
  for (int c : {2, 4}) run_copy<128, 8, 8, LD_NC_NA, ST_NA>("persist", s, d, bytes, sms, c * 2, st);
  // one-shot
  run_copy<256, 8, 3, LD_NC_NA, ST_NA>("oneshot", s, d, bytes, sms, 0, st);
  run_copy<256, 4, 6, LD_NC_NA, ST_NA>("oneshot", s, d, bytes, sms, 0, st);
  run_copy<128, 4, 12, LD_NC_NA, ST_NA>("oneshot", s, d, bytes, sms, 0, st);
  run_copy<128, 8, 6, LD_NC_NA, ST_NA>("oneshot", s, d, bytes, sms, 0, st);
  run_copy<256, 2, 8, LD_NC_NA, ST_NA>("oneshot", s, d, bytes, sms, 0, st);
  // cache-hint variants on the best persistent shape
  run_copy<256, 8, 3, LD_PLAIN, ST_PLAIN>("persist", s, d, bytes, sms, 3, st);
  run_copy<256, 8, 3, LD_NC_NA, ST_PLAIN>("persist", s, d, bytes, sms, 3, st);
  run_copy<256, 8, 3, LD_EF, ST_NA>("persist", s, d, bytes, sms, 3, st);
  run_copy<256, 8, 3, LD_EF, ST_EF>("persist", s, d, bytes, sms, 3, st);
  run_copy<256, 8, 3, LD_NC_NA, ST_CS>("persist", s, d, bytes, sms, 3, st);
  run_copy<256, 8, 3, LD_EF, ST_CS>("persist", s, d, bytes, sms, 3, st);
  run_copy<256, 4, 6, LD_EF, ST_EF>("oneshot", s, d, bytes, sms, 0, st);
  // TMA one-shot (grid == tiles)
  run_tma_oneshot<8192>(s, d, bytes, st);
  run_tma_oneshot<16384>(s, d, bytes, st);
  run_tma_oneshot<32768>(s, d, bytes, st);
  run_tma_oneshot<65536>(s, d, bytes, st);
  // TMA shapes
  run_tma<16384, 6, 4>(s, d, bytes, sms, 1, st);
  run_tma<16384, 6, 4>(s, d, bytes, sms, 2, st);
  run_tma<16384, 4, 2>(s, d, bytes, sms, 3, st);
  run_tma<8192, 8, 6>(s, d, bytes, sms, 3, st);
  run_tma<8192, 6, 4>(s, d, bytes, sms, 4, st);
  run_tma<32768, 3, 1>(s, d, bytes, sms, 2, st);
  run_tma<32768, 6, 4>(s, d, bytes, sms, 1, st);
  run_tma<32768, 3, 1>(s, d, bytes, sms, 1, st);
  run_tma<4096, 12, 10>(s, d, bytes, sms, 4, st);
  run_tma<65536, 3, 1>(s, d, bytes, sms, 1, st);
  return 0;
}
