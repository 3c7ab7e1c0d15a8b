// vmm_lab.cu -- what re-pointing a 1 GiB region costs: cuMemMap / cuMemSetAccess / cuMemUnmap on a GPU that is idle, that
// runs a kernel, that runs a copy-engine copy; access granted to 1 GPU or to all; one call per region or one per 4 regions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/vmm_lab tools/vmm_lab.cu -lcuda      Run: tools/vmm_lab [ngpus]
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* m; cuGetErrorString(r_, &m); printf("{\"error\":\"%s -> %s\"}\n", #x, m); exit(1);} } while (0)
#define RT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\":\"%s -> %s\"}\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void spin(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}
int main(int argc, char** argv) {
  int ndev = 0;
  RT(cudaGetDeviceCount(&ndev));
  if (argc > 1) ndev = std::min(ndev, atoi(argv[1]));
  for (int d = 0; d < ndev; ++d) { RT(cudaSetDevice(d)); RT(cudaFree(0)); }
  RT(cudaSetDevice(0));
  for (int d = 1; d < ndev; ++d) cudaDeviceEnablePeerAccess(d, 0);
  const size_t R = 1ull << 30;
  const int K = 4;
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = 0;
  std::vector<CUmemGenericAllocationHandle> h(K);
  double t = now_ms();
  for (auto& x : h) CK(cuMemCreate(&x, R, &prop, 0));
  printf("{\"op\":\"cuMemCreate 1 GiB\",\"ms_each\":%.3f}\n", (now_ms() - t) / K);
  CUdeviceptr va = 0;
  CK(cuMemAddressReserve(&va, (size_t)K * R * 2, R, 0, 0));
  std::vector<CUmemAccessDesc> one(1), all(ndev);
  one[0].location.type = CU_MEM_LOCATION_TYPE_DEVICE; one[0].location.id = 0; one[0].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (int d = 0; d < ndev; ++d) { all[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE; all[d].location.id = d; all[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; }
  cudaStream_t st;
  RT(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  void *src, *dst;
  RT(cudaMalloc(&src, 8ull << 30)); RT(cudaMalloc(&dst, 8ull << 30));
  auto cycle = [&](const char* what, const std::vector<CUmemAccessDesc>& acc, int background /*0 idle, 1 kernel, 2 copy engine*/, bool batched) {
    double tm = 0, ta = 0, tu = 0;
    const int reps = 3;
    for (int rep = 0; rep < reps; ++rep) {
      if (background == 1) spin<<<148, 64, 0, st>>>(60ull * 1000 * 1000);
      if (background == 2) RT(cudaMemcpyAsync(dst, src, 8ull << 30, cudaMemcpyDeviceToDevice, st));
      double t0 = now_ms();
      for (int i = 0; i < K; ++i) CK(cuMemMap(va + (size_t)i * R, R, 0, h[i], 0));
      double t1 = now_ms();
      if (batched) CK(cuMemSetAccess(va, (size_t)K * R, acc.data(), acc.size()));
      else for (int i = 0; i < K; ++i) CK(cuMemSetAccess(va + (size_t)i * R, R, acc.data(), acc.size()));
      double t2 = now_ms();
      for (int i = 0; i < K; ++i) CK(cuMemUnmap(va + (size_t)i * R, R));
      double t3 = now_ms();
      tm += t1 - t0; ta += t2 - t1; tu += t3 - t2;
      RT(cudaStreamSynchronize(st));
    }
    printf("{\"case\":\"%s\",\"gpus_granted\":%zu,\"background\":\"%s\",\"setaccess_calls\":\"%s\",\"map_ms_per_GiB\":%.3f,\"setaccess_ms_per_GiB\":%.3f,\"unmap_ms_per_GiB\":%.3f}\n", what,
           acc.size(), background == 0 ? "idle" : background == 1 ? "kernel running" : "copy engine running", batched ? "one per 4 GiB" : "one per GiB", tm / reps / K, ta / reps / K,
           tu / reps / K);
    fflush(stdout);
  };
  printf("{\"gpus\":%d}\n", ndev);
  if (ndev > 1) {  // memory that lives on GPU 1, mapped for GPU 0 only: what re-pointing an evicted region's VA at its peer backing costs
    std::vector<CUmemGenericAllocationHandle> hp(K);
    CUmemAllocationProp pp = prop;
    pp.location.id = 1;
    for (auto& x : hp) CK(cuMemCreate(&x, R, &pp, 0));
    std::swap(h, hp);
    cycle("peer memory, home granted", one, 0, false);
    cycle("peer memory, home granted", one, 1, false);
    std::swap(h, hp);
    for (auto& x : hp) CK(cuMemRelease(x));
  }
  for (int bg = 0; bg < 3; ++bg) {
    cycle("home only", one, bg, false);
    cycle("home only", one, bg, true);
    if (ndev > 1) cycle("all gpus", all, bg, false);
  }
  return 0;
}
