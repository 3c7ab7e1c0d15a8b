// vram.cu -- vGPU VRAM tiering: VA reservation + per-region physical backing on
// home HBM / peer HBM / pinned host, with evict & prefetch (see include/tfw_vram.h).
//
// Driver-API VMM calls (cuMemAddressReserve / cuMemCreate / cuMemMap /
// cuMemSetAccess) are resolved through cudaGetDriverEntryPoint so the library
// keeps loading on hosts without libcuda (it then answers TFW_ERR_NO_DEVICE).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <list>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"
#include "tfw_vram.h"

namespace {

struct Drv {
  CUresult (*cuMemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*cuMemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*cuMemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*cuMemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*cuMemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*cuMemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*cuGetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
  bool load() {
    if (ok) return true;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    ok = get("cuMemAddressReserve", (void**)&cuMemAddressReserve) && get("cuMemAddressFree", (void**)&cuMemAddressFree) &&
         get("cuMemCreate", (void**)&cuMemCreate) && get("cuMemRelease", (void**)&cuMemRelease) &&
         get("cuMemMap", (void**)&cuMemMap) && get("cuMemUnmap", (void**)&cuMemUnmap) &&
         get("cuMemSetAccess", (void**)&cuMemSetAccess) &&
         get("cuMemGetAllocationGranularity", (void**)&cuMemGetAllocationGranularity) &&
         get("cuGetErrorString", (void**)&cuGetErrorString);
    if (!ok) cudaGetLastError();
    return ok;
  }
};
Drv g_drv;

// One physical backing of region size.  It keeps a permanent alias mapping in the pool VA
// range, so copies into fresh backing need no VMM call and a migration costs a single
// unmap + map + setAccess of the region's own VA; freed backing is pooled per GPU (within
// the tier budget) instead of being released and re-created (cuMemCreate of 1 GiB ~ms).
struct Phys {
  CUmemGenericAllocationHandle h = 0;
  int device = -1;
  CUdeviceptr alias = 0;
};

struct Region {
  uint32_t tier = TFW_TIER_NONE;
  int32_t peer_slot = -1;
  Phys* phys = nullptr;
  int host_slot = -1;
  std::list<uint32_t>::iterator lru;  // valid when tier == HOME
  bool pinned = false;
};

constexpr uint32_t kWindowSlots = tfw::kInlineDescs;  // regions moved by one mover launch

}  // namespace

struct tfw_vspace {
  tfw_vspace_config cfg{};
  int sm_count = 148;
  CUdeviceptr base = 0, window = 0;   // window = the alias (pool) VA range
  uint64_t alias_slots = 0, alias_next = 0;
  std::vector<uint64_t> alias_free;                 // recycled alias slot indices
  std::vector<std::vector<Phys*>> pool;             // free backing per CUDA device ordinal
  uint64_t R = 0;
  uint32_t n = 0;
  std::vector<Region> regions;
  std::list<uint32_t> lru;  // front = most recently used HOME region
  uint64_t home_used = 0, host_used = 0;
  std::vector<uint64_t> peer_used;
  uint8_t* host_pool = nullptr;
  std::vector<int> host_free;
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: host->device DMAs, so evictions and prefetches of
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;  // the host tier use both PCIe directions at once
  unsigned long long* d_digest = nullptr;
  tfw_move_desc* d_descs = nullptr;  // device copy of a batch's descriptors (TMA mover)
  // Receiver-driven P2P: a copy INTO GPU d is launched on GPU d (one-sided get).  SM-initiated
  // NVLink writes top out at ~718 GB/s on B200 while reads reach ~790 (profiles/r01_peer_lab.jsonl),
  // so evictions are pulled by the peer and prefetches by the home GPU.
  struct DevCtx { cudaStream_t stream = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr; bool used = false; };
  std::vector<DevCtx> dev;           // indexed by CUDA ordinal; [home] aliases `stream`
  tfw_vspace_stats st{};
  std::string err;
};

namespace {

tfw_status vfail(tfw_vspace* vs, tfw_status s, const std::string& m) { vs->err = m; return s; }

#define DRV(vs, call)                                                                       \
  do {                                                                                      \
    CUresult r__ = (call);                                                                  \
    if (r__ != CUDA_SUCCESS) {                                                              \
      const char* m__ = nullptr;                                                            \
      g_drv.cuGetErrorString(r__, &m__);                                                    \
      return vfail(vs, r__ == CUDA_ERROR_OUT_OF_MEMORY ? TFW_ERR_EXHAUSTED : TFW_ERR_FAILED, \
                   std::string(#call) + ": " + (m__ ? m__ : "?"));                          \
    }                                                                                       \
  } while (0)
#define RT(vs, call)                                                                        \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) return vfail(vs, TFW_ERR_FAILED, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

CUmemAllocationProp prop_for(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}

int device_of(const tfw_vspace* vs, uint32_t tier, int32_t peer_slot) {
  return tier == TFW_TIER_PEER ? vs->cfg.peer_devices[peer_slot] : vs->cfg.home_device;
}

// Region VAs are the vGPU's own pointers: only the home GPU uses them.  Alias mappings are
// what the copies read and write, from whichever GPU drives the copy: every GPU of the
// vspace gets access (a one-time cost per backing, the aliases are permanent).
tfw_status set_access(tfw_vspace* vs, CUdeviceptr va, bool all_devices = false) {
  CUmemAccessDesc a[TFW_VRAM_MAX_PEERS + 1];
  size_t n = 0;
  a[n].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  a[n].location.id = vs->cfg.home_device;
  a[n].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  ++n;
  if (all_devices)
    for (uint32_t i = 0; i < vs->cfg.n_peers; ++i) {
      a[n].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      a[n].location.id = vs->cfg.peer_devices[i];
      a[n].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      ++n;
    }
  DRV(vs, g_drv.cuMemSetAccess(va, vs->R, a, n));
  return TFW_OK;
}

// Backing on `device`: recycled from the pool, else created and given its permanent alias mapping.
tfw_status acquire_phys(tfw_vspace* vs, int device, Phys** out) {
  auto& pl = vs->pool[device];
  if (!pl.empty()) { *out = pl.back(); pl.pop_back(); return TFW_OK; }
  uint64_t slot;
  if (!vs->alias_free.empty()) { slot = vs->alias_free.back(); vs->alias_free.pop_back(); }
  else if (vs->alias_next < vs->alias_slots) slot = vs->alias_next++;
  else return vfail(vs, TFW_ERR_EXHAUSTED, "alias VA range exhausted");
  Phys* ph = new (std::nothrow) Phys();
  if (!ph) return TFW_ERR_EXHAUSTED;
  ph->device = device;
  ph->alias = vs->window + slot * vs->R;
  CUmemAllocationProp p = prop_for(device);
  CUresult r = g_drv.cuMemCreate(&ph->h, vs->R, &p, 0);
  if (r != CUDA_SUCCESS) { delete ph; vs->alias_free.push_back(slot); return vfail(vs, r == CUDA_ERROR_OUT_OF_MEMORY ? TFW_ERR_EXHAUSTED : TFW_ERR_FAILED, "cuMemCreate failed"); }
  r = g_drv.cuMemMap(ph->alias, vs->R, 0, ph->h, 0);
  if (r == CUDA_SUCCESS && set_access(vs, ph->alias, true) != TFW_OK) { g_drv.cuMemUnmap(ph->alias, vs->R); r = CUDA_ERROR_UNKNOWN; }
  if (r != CUDA_SUCCESS) {
    g_drv.cuMemRelease(ph->h);
    delete ph;
    vs->alias_free.push_back(slot);
    return vfail(vs, TFW_ERR_FAILED, "mapping new backing failed (no P2P path between home and peer GPU?)");
  }
  *out = ph;
  return TFW_OK;
}

void destroy_phys(tfw_vspace* vs, Phys* ph) {
  g_drv.cuMemUnmap(ph->alias, vs->R);
  g_drv.cuMemRelease(ph->h);
  vs->alias_free.push_back((uint64_t)(ph->alias - vs->window) / vs->R);
  delete ph;
}

// Backing no longer referenced by a region: keep it for reuse while the tier budget allows.
void release_phys(tfw_vspace* vs, Phys* ph, uint64_t used_bytes, uint64_t budget_bytes) {
  auto& pl = vs->pool[ph->device];
  if (used_bytes + (pl.size() + 1) * vs->R <= budget_bytes) pl.push_back(ph);
  else destroy_phys(vs, ph);
}

// Point the region's own VA at `ph` (the alias mapping stays).
tfw_status point_region(tfw_vspace* vs, uint32_t region, Phys* ph) {
  const CUdeviceptr va = vs->base + (uint64_t)region * vs->R;
  DRV(vs, g_drv.cuMemMap(va, vs->R, 0, ph->h, 0));
  return set_access(vs, va);
}

bool budget_ok(const tfw_vspace* vs, uint32_t tier, int32_t slot) {
  if (tier == TFW_TIER_HOME) return vs->home_used + vs->R <= vs->cfg.home_budget_bytes;
  if (tier == TFW_TIER_PEER) return slot >= 0 && (uint32_t)slot < vs->cfg.n_peers && vs->peer_used[slot] + vs->R <= vs->cfg.peer_budget_bytes;
  if (tier == TFW_TIER_HOST) return !vs->host_free.empty();
  return false;
}

void account(tfw_vspace* vs, uint32_t region, uint32_t tier, int32_t slot, int sign) {
  Region& r = vs->regions[region];
  if (tier == TFW_TIER_HOME) {
    vs->home_used += sign * (int64_t)vs->R;
    vs->st.regions_home += sign;
    if (sign > 0) { vs->lru.push_front(region); r.lru = vs->lru.begin(); } else { vs->lru.erase(r.lru); }
  } else if (tier == TFW_TIER_PEER) {
    vs->peer_used[slot] += sign * (int64_t)vs->R;
    vs->st.regions_peer += sign;
  } else if (tier == TFW_TIER_HOST) {
    vs->host_used += sign * (int64_t)vs->R;
    vs->st.regions_host += sign;
  }
}

CUdeviceptr va_of(const tfw_vspace* vs, uint32_t region) { return vs->base + (uint64_t)region * vs->R; }

}  // namespace

extern "C" {

const char* tfw_vspace_last_error(const tfw_vspace* vs) { return vs ? vs->err.c_str() : "null vspace"; }

tfw_status tfw_vspace_create(const tfw_vspace_config* cfg, tfw_vspace** out) {
  if (!cfg || !out || cfg->struct_size != sizeof(tfw_vspace_config)) return TFW_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return TFW_ERR_NO_DEVICE; }
  if (cfg->home_device < 0 || cfg->home_device >= ndev || cfg->n_peers > TFW_VRAM_MAX_PEERS) return TFW_ERR_INVALID;
  if (!cfg->region_bytes || (cfg->region_bytes & ((2u << 20) - 1)) || !cfg->va_bytes || cfg->va_bytes % cfg->region_bytes) return TFW_ERR_INVALID;
  for (uint32_t i = 0; i < cfg->n_peers; ++i)
    if (cfg->peer_devices[i] < 0 || cfg->peer_devices[i] >= ndev || cfg->peer_devices[i] == cfg->home_device) return TFW_ERR_INVALID;
  // primary contexts everywhere we will place memory
  for (uint32_t i = 0; i < cfg->n_peers; ++i) {
    if (cudaSetDevice(cfg->peer_devices[i]) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { cudaGetLastError(); return TFW_ERR_FAILED; }
    int can = 0;
    cudaDeviceCanAccessPeer(&can, cfg->home_device, cfg->peer_devices[i]);
    if (!can) return TFW_ERR_NOT_SUPPORTED;
  }
  if (cudaSetDevice(cfg->home_device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { cudaGetLastError(); return TFW_ERR_FAILED; }
  if (!g_drv.load()) return TFW_ERR_NO_DEVICE;
  tfw_vspace* vs = new (std::nothrow) tfw_vspace();
  if (!vs) return TFW_ERR_EXHAUSTED;
  vs->cfg = *cfg;
  vs->R = cfg->region_bytes;
  vs->n = (uint32_t)(cfg->va_bytes / cfg->region_bytes);
  vs->regions.resize(vs->n);
  vs->peer_used.assign(cfg->n_peers, 0);
  auto bail = [&](tfw_status s) { tfw_vspace_destroy(vs); return s; };
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, cfg->home_device) != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (prop.major < 10) return bail(TFW_ERR_NOT_SUPPORTED);
  vs->sm_count = prop.multiProcessorCount;
  if (tfw::preload_kernels() != cudaSuccess) return bail(TFW_ERR_FAILED);
  CUmemAllocationProp p = prop_for(cfg->home_device);
  size_t gran = 0;
  if (g_drv.cuMemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || vs->R % gran) return bail(TFW_ERR_INVALID);
  if (g_drv.cuMemAddressReserve(&vs->base, cfg->va_bytes, vs->R > (1ull << 30) ? (1ull << 30) : vs->R, 0, 0) != CUDA_SUCCESS) return bail(TFW_ERR_EXHAUSTED);
  // alias range: one slot per backing that can exist at once (every region + one batch in flight)
  vs->alias_slots = (uint64_t)vs->n + kWindowSlots;
  vs->pool.assign((size_t)ndev, {});
  if (g_drv.cuMemAddressReserve(&vs->window, vs->alias_slots * vs->R, vs->R > (1ull << 30) ? (1ull << 30) : vs->R, 0, 0) != CUDA_SUCCESS) return bail(TFW_ERR_EXHAUSTED);
  if (cudaStreamCreateWithFlags(&vs->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (cudaStreamCreateWithFlags(&vs->stream2, cudaStreamNonBlocking) != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (cudaEventCreate(&vs->e0) != cudaSuccess || cudaEventCreate(&vs->e1) != cudaSuccess || cudaEventCreate(&vs->e2) != cudaSuccess) return bail(TFW_ERR_FAILED);
  if (cudaMalloc(reinterpret_cast<void**>(&vs->d_digest), 8) != cudaSuccess) return bail(TFW_ERR_EXHAUSTED);
  if (cudaMalloc(reinterpret_cast<void**>(&vs->d_descs), sizeof(tfw_move_desc) * kWindowSlots) != cudaSuccess) return bail(TFW_ERR_EXHAUSTED);
  vs->dev.resize((size_t)ndev);
  vs->dev[cfg->home_device].stream = vs->stream;
  vs->dev[cfg->home_device].e0 = vs->e0;
  vs->dev[cfg->home_device].e1 = vs->e1;
  for (uint32_t i = 0; i < cfg->n_peers; ++i) {
    const int d = cfg->peer_devices[i];
    int back = 0;
    cudaDeviceCanAccessPeer(&back, d, cfg->home_device);
    if (!back) return bail(TFW_ERR_NOT_SUPPORTED);
    if (cudaSetDevice(d) != cudaSuccess || tfw::preload_kernels() != cudaSuccess ||
        cudaStreamCreateWithFlags(&vs->dev[d].stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&vs->dev[d].e0) != cudaSuccess || cudaEventCreate(&vs->dev[d].e1) != cudaSuccess) {
      cudaGetLastError();
      cudaSetDevice(cfg->home_device);
      return bail(TFW_ERR_FAILED);
    }
  }
  cudaSetDevice(cfg->home_device);
  const uint64_t host_slots = cfg->host_budget_bytes / vs->R;
  if (host_slots) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&vs->host_pool), host_slots * vs->R, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return bail(TFW_ERR_EXHAUSTED); }
    for (int i = (int)host_slots - 1; i >= 0; --i) vs->host_free.push_back(i);
  }
  *out = vs;
  return TFW_OK;
}

tfw_status tfw_vspace_destroy(tfw_vspace* vs) {
  if (!vs) return TFW_ERR_INVALID;
  cudaSetDevice(vs->cfg.home_device);
  if (vs->stream) cudaStreamSynchronize(vs->stream);
  for (uint32_t i = 0; i < vs->n && i < vs->regions.size(); ++i) {
    Region& r = vs->regions[i];
    if ((r.tier == TFW_TIER_HOME || r.tier == TFW_TIER_PEER) && r.phys) {
      g_drv.cuMemUnmap(va_of(vs, i), vs->R);
      destroy_phys(vs, r.phys);
    }
  }
  for (auto& pl : vs->pool) for (Phys* ph : pl) destroy_phys(vs, ph);
  for (size_t d = 0; d < vs->dev.size(); ++d) {
    if ((int)d == vs->cfg.home_device || !vs->dev[d].stream) continue;
    cudaSetDevice((int)d);
    cudaStreamSynchronize(vs->dev[d].stream);
    cudaEventDestroy(vs->dev[d].e0);
    cudaEventDestroy(vs->dev[d].e1);
    cudaStreamDestroy(vs->dev[d].stream);
  }
  cudaSetDevice(vs->cfg.home_device);
  if (vs->base) g_drv.cuMemAddressFree(vs->base, vs->cfg.va_bytes);
  if (vs->window) g_drv.cuMemAddressFree(vs->window, vs->alias_slots * vs->R);
  if (vs->host_pool) cudaFreeHost(vs->host_pool);
  if (vs->d_digest) cudaFree(vs->d_digest);
  if (vs->d_descs) cudaFree(vs->d_descs);
  if (vs->e0) cudaEventDestroy(vs->e0);
  if (vs->e1) cudaEventDestroy(vs->e1);
  if (vs->e2) cudaEventDestroy(vs->e2);
  if (vs->stream2) cudaStreamDestroy(vs->stream2);
  if (vs->stream) cudaStreamDestroy(vs->stream);
  delete vs;
  return TFW_OK;
}

tfw_status tfw_vspace_info(tfw_vspace* vs, uint64_t* base, uint64_t* region_bytes, uint32_t* n_regions) {
  if (!vs) return TFW_ERR_INVALID;
  if (base) *base = (uint64_t)vs->base;
  if (region_bytes) *region_bytes = vs->R;
  if (n_regions) *n_regions = vs->n;
  return TFW_OK;
}

tfw_status tfw_vspace_residency(tfw_vspace* vs, uint32_t region, uint32_t* tier, int32_t* device) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  const Region& r = vs->regions[region];
  if (tier) *tier = r.tier;
  if (device) *device = r.tier == TFW_TIER_HOME ? vs->cfg.home_device : r.tier == TFW_TIER_PEER ? vs->cfg.peer_devices[r.peer_slot] : -1;
  return TFW_OK;
}

tfw_status tfw_vspace_populate(tfw_vspace* vs, uint32_t region, uint32_t tier, int32_t peer_slot) {
  if (!vs || region >= vs->n || tier == TFW_TIER_NONE || tier > TFW_TIER_HOST) return TFW_ERR_INVALID;
  Region& r = vs->regions[region];
  if (r.tier != TFW_TIER_NONE) return vfail(vs, TFW_ERR_INVALID, "region already backed");
  if (!budget_ok(vs, tier, peer_slot)) return vfail(vs, TFW_ERR_EXHAUSTED, "tier budget exhausted");
  cudaSetDevice(vs->cfg.home_device);
  if (tier == TFW_TIER_HOST) {
    r.host_slot = vs->host_free.back();
    vs->host_free.pop_back();
    std::memset(vs->host_pool + (uint64_t)r.host_slot * vs->R, 0, vs->R);
  } else {
    tfw_status s = acquire_phys(vs, device_of(vs, tier, peer_slot), &r.phys);
    if (s != TFW_OK) return s;
    s = point_region(vs, region, r.phys);
    if (s != TFW_OK) return s;
    tfw_move_desc d{};
    d.dst = (uint64_t)va_of(vs, region);
    d.len = vs->R;
    d.tile0 = 0;
    RT(vs, tfw::launch_mover_inline(&d, 1, tfw::mover_tiles(d.dst, d.len), vs->sm_count, 0, vs->stream));  // scrub
    RT(vs, cudaStreamSynchronize(vs->stream));
    vs->st.mover_launches++;
  }
  r.tier = tier;
  r.peer_slot = tier == TFW_TIER_PEER ? peer_slot : -1;
  account(vs, region, tier, peer_slot, +1);
  return TFW_OK;
}

tfw_status tfw_vspace_migrate(tfw_vspace* vs, const uint32_t* regions, const uint8_t* tiers, const int32_t* peer_slots,
                              uint32_t n, tfw_migrate_result* res) {
  if (!vs || !regions || !tiers || !n) return TFW_ERR_INVALID;
  cudaSetDevice(vs->cfg.home_device);
  const auto t_begin = std::chrono::steady_clock::now();
  tfw_migrate_result acc{};
  for (uint32_t off = 0; off < n; off += kWindowSlots) {
    const uint32_t m = std::min(kWindowSlots, n - off);
    struct Move { uint32_t region; uint32_t to; int32_t slot; Phys* nphys = nullptr; int nhost = -1; bool noop = false; };
    std::vector<Move> mv(m);
    // ---- validate -------------------------------------------------------------------
    for (uint32_t k = 0; k < m; ++k) {
      Move& x = mv[k];
      x.region = regions[off + k];
      x.to = tiers[off + k];
      x.slot = peer_slots ? peer_slots[off + k] : -1;
      if (x.region >= vs->n || x.to == TFW_TIER_NONE || x.to > TFW_TIER_HOST) return vfail(vs, TFW_ERR_INVALID, "bad migrate request");
      for (uint32_t j = 0; j < k; ++j) if (mv[j].region == x.region) return vfail(vs, TFW_ERR_INVALID, "region listed twice in one batch");
      const Region& r = vs->regions[x.region];
      if (r.tier == TFW_TIER_NONE) return vfail(vs, TFW_ERR_INVALID, "region not populated");
      x.noop = r.tier == x.to && (x.to != TFW_TIER_PEER || r.peer_slot == x.slot);
    }
    // copies grouped by the GPU that drives them (the destination GPU; the home GPU for host moves)
    struct P2P { int exec; uint64_t dst, src; };
    std::vector<P2P> p2p;
    tfw_status rc = TFW_OK;
    struct Dma { void* dst; const void* src; cudaMemcpyKind kind; };
    std::vector<Dma> dma;
    // budgets are checked against the state after the moves already planned in this window
    uint64_t home_plan = vs->home_used, host_plan = vs->host_free.size();
    std::vector<uint64_t> peer_plan = vs->peer_used;
    for (uint32_t k = 0; k < m; ++k) {  // regions leaving a tier in this window make room in it
      if (mv[k].noop) continue;
      const Region& r = vs->regions[mv[k].region];
      if (r.tier == TFW_TIER_HOME) home_plan -= vs->R;
      else if (r.tier == TFW_TIER_PEER) peer_plan[r.peer_slot] -= vs->R;
      else if (r.tier == TFW_TIER_HOST) ++host_plan;
    }
    // ---- new backing (pooled: no VMM call in the steady state) and the copy list --------
    for (uint32_t k = 0; k < m && rc == TFW_OK; ++k) {
      Move& x = mv[k];
      if (x.noop) continue;
      Region& r = vs->regions[x.region];
      if (x.to == TFW_TIER_HOME) { if (home_plan + vs->R > vs->cfg.home_budget_bytes) rc = TFW_ERR_EXHAUSTED; else home_plan += vs->R; }
      else if (x.to == TFW_TIER_PEER) { if (x.slot < 0 || (uint32_t)x.slot >= vs->cfg.n_peers || peer_plan[x.slot] + vs->R > vs->cfg.peer_budget_bytes) rc = TFW_ERR_EXHAUSTED; else peer_plan[x.slot] += vs->R; }
      else { if (host_plan == 0) rc = TFW_ERR_EXHAUSTED; else --host_plan; }
      if (rc != TFW_OK) { vfail(vs, rc, "target tier budget exhausted"); break; }
      if (x.to == TFW_TIER_HOST) {
        if (vs->host_free.empty()) { rc = vfail(vs, TFW_ERR_EXHAUSTED, "no free host slot until this batch completes; split the batch"); break; }
        x.nhost = vs->host_free.back();
        vs->host_free.pop_back();
        dma.push_back({vs->host_pool + (uint64_t)x.nhost * vs->R, reinterpret_cast<const void*>(r.phys->alias), cudaMemcpyDeviceToHost});
      } else {
        rc = acquire_phys(vs, device_of(vs, x.to, x.slot), &x.nphys);
        if (rc != TFW_OK) break;
        if (r.tier == TFW_TIER_HOST) dma.push_back({reinterpret_cast<void*>(x.nphys->alias), vs->host_pool + (uint64_t)r.host_slot * vs->R, cudaMemcpyHostToDevice});
        else p2p.push_back({(vs->cfg.flags & TFW_VS_PUSH_EVICT) ? vs->cfg.home_device : x.nphys->device, (uint64_t)x.nphys->alias, (uint64_t)r.phys->alias});
      }
    }
    if (rc != TFW_OK) {  // undo this window's reservations
      for (auto& x : mv) {
        if (x.nphys) vs->pool[x.nphys->device].push_back(x.nphys);
        if (x.nhost >= 0) vs->host_free.push_back(x.nhost);
      }
      return rc;
    }
    // ---- copy: per destination GPU ONE mover launch (or one DMA per region), all GPUs at once;
    //      the copy engine of the home GPU for the host tier ------------------------------------
    for (auto& dc : vs->dev) dc.used = false;
    for (const P2P& c : p2p) vs->dev[c.exec].used = true;
    if (!dma.empty()) vs->dev[vs->cfg.home_device].used = true;
    for (size_t d = 0; d < vs->dev.size(); ++d) {
      tfw_vspace::DevCtx& dc = vs->dev[d];
      if (!dc.used) continue;
      RT(vs, cudaSetDevice((int)d));
      RT(vs, cudaEventRecord(dc.e0, dc.stream));
      tfw_move_desc descs[kWindowSlots];
      uint32_t nd = 0;
      for (const P2P& c : p2p) {
        if (c.exec != (int)d) continue;
        if (vs->cfg.flags & TFW_VS_COPY_ENGINE) { RT(vs, cudaMemcpyAsync(reinterpret_cast<void*>(c.dst), reinterpret_cast<const void*>(c.src), vs->R, cudaMemcpyDeviceToDevice, dc.stream)); continue; }
        descs[nd].dst = c.dst; descs[nd].src = c.src; descs[nd].len = vs->R; descs[nd].fill = 0;
        ++nd;
      }
      if (nd) {
        uint64_t t = 0;
        for (uint32_t i = 0; i < nd; ++i) { descs[i].tile0 = (uint32_t)t; t += tfw::mover_tiles(descs[i].dst, descs[i].len); }
        RT(vs, tfw::launch_mover_inline(descs, nd, (uint32_t)t, vs->sm_count, 0, dc.stream));
        vs->st.mover_launches++;
        acc.launches++;
      }
      if ((int)d == vs->cfg.home_device && !dma.empty()) {
        bool used2 = false;
        for (const Dma& c : dma) {
          cudaStream_t st = c.kind == cudaMemcpyHostToDevice ? vs->stream2 : vs->stream;
          if (st == vs->stream2 && !used2) { RT(vs, cudaStreamWaitEvent(vs->stream2, vs->e0, 0)); used2 = true; }
          RT(vs, cudaMemcpyAsync(c.dst, c.src, vs->R, c.kind, st));
        }
        if (used2) { RT(vs, cudaEventRecord(vs->e2, vs->stream2)); RT(vs, cudaStreamWaitEvent(vs->stream, vs->e2, 0)); }
      }
      RT(vs, cudaEventRecord(dc.e1, dc.stream));
    }
    float batch_ms = 0;
    for (size_t d = 0; d < vs->dev.size(); ++d) {
      tfw_vspace::DevCtx& dc = vs->dev[d];
      if (!dc.used) continue;
      RT(vs, cudaSetDevice((int)d));
      RT(vs, cudaEventSynchronize(dc.e1));
      float ms = 0;
      RT(vs, cudaEventElapsedTime(&ms, dc.e0, dc.e1));
      batch_ms = std::max(batch_ms, ms);  // the GPUs copy concurrently: the batch takes as long as the slowest
    }
    RT(vs, cudaSetDevice(vs->cfg.home_device));
    acc.copy_ms += batch_ms;
    // ---- re-point: the region's VA now names the new backing ------------------------------
    for (uint32_t k = 0; k < m; ++k) {
      Move& x = mv[k];
      if (x.noop) continue;
      Region& r = vs->regions[x.region];
      const uint32_t from = r.tier;
      const int32_t from_slot = r.peer_slot;
      if (from == TFW_TIER_PEER && x.to == TFW_TIER_HOME) vs->st.prefetch_bytes_peer += vs->R;
      if (from == TFW_TIER_HOME && x.to == TFW_TIER_PEER) vs->st.evict_bytes_peer += vs->R;
      if (from == TFW_TIER_PEER && x.to == TFW_TIER_PEER) { vs->st.evict_bytes_peer += vs->R; vs->st.prefetch_bytes_peer += vs->R; }
      if (x.to == TFW_TIER_HOST) vs->st.evict_bytes_host += vs->R;
      if (from == TFW_TIER_HOST) vs->st.prefetch_bytes_host += vs->R;
      Phys* old = r.phys;
      if (from != TFW_TIER_HOST) DRV(vs, g_drv.cuMemUnmap(va_of(vs, x.region), vs->R));
      else { vs->host_free.push_back(r.host_slot); r.host_slot = -1; }
      account(vs, x.region, from, from_slot, -1);
      if (old) {
        const uint64_t used = from == TFW_TIER_HOME ? vs->home_used : vs->peer_used[from_slot];
        const uint64_t budget = from == TFW_TIER_HOME ? vs->cfg.home_budget_bytes : vs->cfg.peer_budget_bytes;
        release_phys(vs, old, used, budget);
      }
      r.phys = nullptr;
      if (x.to == TFW_TIER_HOST) {
        r.host_slot = x.nhost;
      } else {
        tfw_status s = point_region(vs, x.region, x.nphys);
        if (s != TFW_OK) return s;
        r.phys = x.nphys;
      }
      r.tier = x.to;
      r.peer_slot = x.to == TFW_TIER_PEER ? x.slot : -1;
      account(vs, x.region, x.to, x.slot, +1);
      vs->st.remaps++;
      acc.bytes += vs->R;
    }
  }
  acc.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (res) *res = acc;
  return TFW_OK;
}

tfw_status tfw_vspace_access(tfw_vspace* vs, uint32_t region) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  Region& r = vs->regions[region];
  if (r.tier == TFW_TIER_HOME) {
    vs->lru.erase(r.lru);
    vs->lru.push_front(region);
    r.lru = vs->lru.begin();
    vs->st.policy_hits++;
    return TFW_OK;
  }
  // One batch: the least-recently-used HOME regions leave (emptiest peer first, else host) while
  // the wanted region comes in -- the copies share the launch / run on both PCIe directions.
  uint32_t regs[kWindowSlots];
  uint8_t tiers[kWindowSlots];
  int32_t slots[kWindowSlots];
  uint32_t nmv = 0;
  uint64_t home_after = vs->home_used;
  std::vector<uint64_t> peer_after = vs->peer_used;
  size_t host_after = vs->host_free.size() + (r.tier == TFW_TIER_HOST ? 0 : 0);
  auto victim = vs->lru.rbegin();
  while (home_after + vs->R > vs->cfg.home_budget_bytes) {
    while (victim != vs->lru.rend() && vs->regions[*victim].pinned) ++victim;
    if (victim == vs->lru.rend() || nmv + 1 >= kWindowSlots) return vfail(vs, TFW_ERR_EXHAUSTED, "home budget exhausted by pinned regions");
    int best = -1;
    for (uint32_t p = 0; p < vs->cfg.n_peers; ++p)
      if (peer_after[p] + vs->R <= vs->cfg.peer_budget_bytes && (best < 0 || peer_after[p] < peer_after[best])) best = (int)p;
    if (best < 0 && host_after == 0) return vfail(vs, TFW_ERR_EXHAUSTED, "no tier has room for an evicted region");
    regs[nmv] = *victim;
    tiers[nmv] = best >= 0 ? TFW_TIER_PEER : TFW_TIER_HOST;
    slots[nmv] = best;
    if (best >= 0) peer_after[best] += vs->R; else --host_after;
    home_after -= vs->R;
    ++nmv;
    ++victim;
  }
  const uint32_t evictions = nmv;
  if (r.tier == TFW_TIER_NONE) {  // first touch: fresh zero-filled HOME backing once the victims are out
    if (nmv) {
      tfw_status s = tfw_vspace_migrate(vs, regs, tiers, slots, nmv, nullptr);
      if (s != TFW_OK) return s;
      vs->st.policy_evictions += evictions;
    }
    return tfw_vspace_populate(vs, region, TFW_TIER_HOME, -1);
  }
  regs[nmv] = region;
  tiers[nmv] = TFW_TIER_HOME;
  slots[nmv] = -1;
  ++nmv;
  tfw_status s = tfw_vspace_migrate(vs, regs, tiers, slots, nmv, nullptr);
  if (s != TFW_OK) return s;
  vs->st.policy_evictions += evictions;
  vs->st.policy_prefetches++;
  return TFW_OK;
}

tfw_status tfw_vspace_unpopulate(tfw_vspace* vs, uint32_t region) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  Region& r = vs->regions[region];
  if (r.tier == TFW_TIER_NONE) return TFW_OK;
  cudaSetDevice(vs->cfg.home_device);
  const uint32_t from = r.tier;
  const int32_t from_slot = r.peer_slot;
  if (from == TFW_TIER_HOST) {
    vs->host_free.push_back(r.host_slot);
    r.host_slot = -1;
  } else {
    DRV(vs, g_drv.cuMemUnmap(va_of(vs, region), vs->R));
  }
  account(vs, region, from, from_slot, -1);
  if (r.phys) {
    const uint64_t used = from == TFW_TIER_HOME ? vs->home_used : vs->peer_used[from_slot];
    const uint64_t budget = from == TFW_TIER_HOME ? vs->cfg.home_budget_bytes : vs->cfg.peer_budget_bytes;
    release_phys(vs, r.phys, used, budget);
    r.phys = nullptr;
  }
  r.tier = TFW_TIER_NONE;
  r.peer_slot = -1;
  r.pinned = false;
  return TFW_OK;
}

tfw_status tfw_vspace_pin(tfw_vspace* vs, uint32_t region, int pinned) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  vs->regions[region].pinned = pinned != 0;
  return TFW_OK;
}

tfw_status tfw_vspace_get_stats(tfw_vspace* vs, tfw_vspace_stats* out) {
  if (!vs || !out) return TFW_ERR_INVALID;
  *out = vs->st;
  return TFW_OK;
}

tfw_status tfw_vspace_fill_pattern(tfw_vspace* vs, uint32_t region, uint64_t seed) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  const Region& r = vs->regions[region];
  if (r.tier != TFW_TIER_HOME && r.tier != TFW_TIER_PEER) return TFW_ERR_NOT_SUPPORTED;
  cudaSetDevice(vs->cfg.home_device);
  RT(vs, tfw::launch_pattern(reinterpret_cast<void*>(va_of(vs, region)), vs->R, seed, vs->sm_count, vs->stream));
  RT(vs, cudaStreamSynchronize(vs->stream));
  return TFW_OK;
}

tfw_status tfw_vspace_digest(tfw_vspace* vs, uint32_t region, uint64_t* digest) {
  if (!vs || region >= vs->n || !digest) return TFW_ERR_INVALID;
  const Region& r = vs->regions[region];
  if (r.tier != TFW_TIER_HOME && r.tier != TFW_TIER_PEER) return TFW_ERR_NOT_SUPPORTED;
  cudaSetDevice(vs->cfg.home_device);
  RT(vs, cudaMemsetAsync(vs->d_digest, 0, 8, vs->stream));
  RT(vs, tfw::launch_digest(reinterpret_cast<void*>(va_of(vs, region)), vs->R, vs->d_digest, vs->sm_count, vs->stream));
  unsigned long long sum = 0;
  RT(vs, cudaMemcpyAsync(&sum, vs->d_digest, 8, cudaMemcpyDeviceToHost, vs->stream));
  RT(vs, cudaStreamSynchronize(vs->stream));
  *digest = tfw::digest_mix((uint64_t)sum ^ (vs->R * tfw::kDigestK1));
  return TFW_OK;
}

tfw_status tfw_vspace_read(tfw_vspace* vs, uint32_t region, uint64_t off, void* dst, uint64_t nbytes) {
  if (!vs || region >= vs->n || !dst || off > vs->R || nbytes > vs->R - off) return TFW_ERR_INVALID;
  const Region& r = vs->regions[region];
  cudaSetDevice(vs->cfg.home_device);
  if (r.tier == TFW_TIER_HOST) { std::memcpy(dst, vs->host_pool + (uint64_t)r.host_slot * vs->R + off, nbytes); return TFW_OK; }
  if (r.tier == TFW_TIER_NONE) return TFW_ERR_NOT_SUPPORTED;
  RT(vs, cudaMemcpyAsync(dst, reinterpret_cast<void*>(va_of(vs, region) + off), nbytes, cudaMemcpyDeviceToHost, vs->stream));
  RT(vs, cudaStreamSynchronize(vs->stream));
  return TFW_OK;
}

tfw_status tfw_vspace_write(tfw_vspace* vs, uint32_t region, uint64_t off, const void* src, uint64_t nbytes) {
  if (!vs || region >= vs->n || !src || off > vs->R || nbytes > vs->R - off) return TFW_ERR_INVALID;
  const Region& r = vs->regions[region];
  cudaSetDevice(vs->cfg.home_device);
  if (r.tier == TFW_TIER_HOST) { std::memcpy(vs->host_pool + (uint64_t)r.host_slot * vs->R + off, src, nbytes); return TFW_OK; }
  if (r.tier == TFW_TIER_NONE) return TFW_ERR_NOT_SUPPORTED;
  RT(vs, cudaMemcpyAsync(reinterpret_cast<void*>(va_of(vs, region) + off), src, nbytes, cudaMemcpyHostToDevice, vs->stream));
  RT(vs, cudaStreamSynchronize(vs->stream));
  return TFW_OK;
}

}  // extern "C"
