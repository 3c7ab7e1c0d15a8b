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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
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

struct Transit;

struct Region {
  uint32_t tier = TFW_TIER_NONE;
  int32_t peer_slot = -1;
  Phys* phys = nullptr;
  int host_slot = -1;
  std::list<uint32_t>::iterator lru;  // valid while the region is accounted to HOME
  uint32_t pinned = 0;         // pin count: pinned regions are never chosen as victims
  bool mapped = false;         // the region's own VA is mapped (a PEER region evicted by the policy is not, see finish_move)
  Transit* transit = nullptr;  // non-null while a migration of this region is in flight
  uint64_t last_use = 0;       // access sequence number of the last touch
};

// One region on its way to another tier.  The copy is enqueued when the move begins; what is
// left for the host (re-pointing the VA of an evicted region, releasing the old backing,
// accounting) happens in finish_move once `done` has completed.
struct Transit {
  uint32_t region = 0, from = 0, to = 0;
  int32_t from_slot = -1, to_slot = -1;
  Phys* ophys = nullptr;
  Phys* nphys = nullptr;
  int ohost = -1, nhost = -1;
  cudaEvent_t done = nullptr;
  int ev_dev = 0;       // device `done` was recorded on
  bool va_done = false; // the region's VA already names the new backing (moves INTO the home GPU)
  bool lazy = false;    // policy eviction: the region's VA is left unmapped at its new tier until it is asked for again
};

constexpr uint32_t kWindowSlots = tfw::kInlineDescs;  // regions moved by one mover launch

// TFW_VS_FIXED_FRAMES: one of the home GPU's backings.  Frame f is mapped, for ever, at the VA of every region r with
// r % frames == f; `occupant` is the region whose bytes it holds (or is about to hold), `leaving` the eviction of the
// previous occupant while its copy is in flight (the copy of the next occupant is ordered behind it).
struct Frame {
  Phys* ph = nullptr;
  int32_t occupant = -1;
  Transit* leaving = nullptr;
};

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
  struct DevCtx { cudaStream_t stream = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr; bool used = false; std::vector<cudaEvent_t> ev_pool; };
  std::vector<DevCtx> dev;           // indexed by CUDA ordinal; [home] aliases `stream`
  // ---- asynchronous migrations ----
  std::list<Transit> transits;       // in issue order
  uint32_t evictions_in_flight = 0;
  cudaStream_t client = nullptr;     // the vGPU's execution stream (tfw_vspace_bind_stream); null: callers quiesce themselves
  struct Mark { uint64_t seq; cudaEvent_t ev; };
  std::deque<Mark> marks;            // events on the client stream: mark m covers every client kernel enqueued before access #m
  uint64_t seq = 0;                  // access counter
  uint32_t last_access = ~0u;        // for the sequential detector
  uint32_t ahead = 0;                // prefetch depth (cfg.prefetch_ahead)
  int peer_ctas = 0;                 // CTAs per SM of a peer-tier copy kernel (0 = one tile per CTA); TFW_VS_PEER_CTAS
  bool remap_late = false;           // TFW_VS_REMAP_LATE=1: a prefetched region's VA is re-pointed when its copy has completed, not when it is issued
  bool fixed = false;                // TFW_VS_FIXED_FRAMES: region VAs never change their backing, no VMM call after create
  std::vector<Frame> frames;
  tfw_vspace_stats st{};
  // TFW_VS_DEBUG=1: what the VMM calls cost, by kind (printed by tfw_vspace_destroy)
  struct VmmDbg { uint64_t n = 0, ns = 0, max_ns = 0, slow = 0; } dbg_unmap, dbg_map, dbg_access;
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
  const auto tv0 = std::chrono::steady_clock::now();
  struct Tick { tfw_vspace* v; std::chrono::steady_clock::time_point t; ~Tick() { v->st.vmm_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t).count(); } } tick{vs, tv0};
  vs->st.phys_created++;
  ph->device = device;
  ph->alias = vs->window + slot * vs->R;
  CUmemAllocationProp p = prop_for(device);
  CUresult r = g_drv.cuMemCreate(&ph->h, vs->R, &p, 0);
  if (r != CUDA_SUCCESS) { delete ph; vs->alias_free.push_back(slot); return vfail(vs, r == CUDA_ERROR_OUT_OF_MEMORY ? TFW_ERR_EXHAUSTED : TFW_ERR_FAILED, "cuMemCreate failed"); }
  r = g_drv.cuMemMap(ph->alias, vs->R, 0, ph->h, 0);
  if (r == CUDA_SUCCESS && set_access(vs, ph->alias, !(vs->cfg.flags & TFW_VS_HOME_DRIVEN)) != TFW_OK) { g_drv.cuMemUnmap(ph->alias, vs->R); r = CUDA_ERROR_UNKNOWN; }
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
  const auto tv0 = std::chrono::steady_clock::now();
  struct Tick { tfw_vspace* v; std::chrono::steady_clock::time_point t; ~Tick() { v->st.vmm_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t).count(); } } tick{vs, tv0};
  vs->st.phys_destroyed++;
  g_drv.cuMemUnmap(ph->alias, vs->R);
  g_drv.cuMemRelease(ph->h);
  vs->alias_free.push_back((uint64_t)(ph->alias - vs->window) / vs->R);
  delete ph;
}

// Backing no longer referenced by a region: keep it for reuse while the tier budget allows.
void release_phys(tfw_vspace* vs, Phys* ph, uint64_t used_bytes, uint64_t budget_bytes) {
  auto& pl = vs->pool[ph->device];
  // Within the tier's budget everything is kept.  Beyond it a few spares per GPU still are: a vGPU that runs at its
  // budget swaps one region in for every region it swaps out, and creating + mapping + granting access to a fresh
  // 1 GiB allocation on every miss costs more than moving the gigabyte (cuMemCreate / cuMemMap / cuMemSetAccess over
  // all GPUs of the space: ~2 ms at 8 GPUs against 1.4 ms of NVLink time).
  const size_t spares = (size_t)vs->ahead + 16;  // (moves finish in bursts: a handful of spares would still be destroyed and re-created)
  if (used_bytes + (pl.size() + 1) * vs->R <= budget_bytes || pl.size() < spares) pl.push_back(ph);
  else destroy_phys(vs, ph);
}

void dbg_add(tfw_vspace::VmmDbg& d, uint64_t ns) { d.n++; d.ns += ns; if (ns > d.max_ns) d.max_ns = ns; if (ns > 500000) d.slow++; }

tfw_status unmap_va(tfw_vspace* vs, uint32_t region) {
  const auto tv0 = std::chrono::steady_clock::now();
  const CUresult r = g_drv.cuMemUnmap(vs->base + (uint64_t)region * vs->R, vs->R);
  const uint64_t dn = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - tv0).count();
  vs->st.vmm_ns += dn;
  dbg_add(vs->dbg_unmap, dn);
  if (r != CUDA_SUCCESS) return vfail(vs, TFW_ERR_FAILED, "cuMemUnmap of a region failed");
  return TFW_OK;
}

// Point the region's own VA at `ph` (the alias mapping stays).
tfw_status point_region(tfw_vspace* vs, uint32_t region, Phys* ph) {
  const auto tv0 = std::chrono::steady_clock::now();
  struct Tick { tfw_vspace* v; std::chrono::steady_clock::time_point t; ~Tick() { v->st.vmm_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t).count(); } } tick{vs, tv0};
  const CUdeviceptr va = vs->base + (uint64_t)region * vs->R;
  const auto tm0 = std::chrono::steady_clock::now();
  DRV(vs, g_drv.cuMemMap(va, vs->R, 0, ph->h, 0));
  const auto tm1 = std::chrono::steady_clock::now();
  const tfw_status sa = set_access(vs, va);
  dbg_add(vs->dbg_map, (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(tm1 - tm0).count());
  dbg_add(vs->dbg_access, (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - tm1).count());
  return sa;
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

Frame& frame_of(tfw_vspace* vs, uint32_t region) { return vs->frames[region % vs->frames.size()]; }

// Where the home GPU finds the region's bytes: its own VA -- except in fixed-frames mode for a region that lives on a
// peer, whose VA keeps naming its home frame (somebody else's bytes by now): the backing's alias mapping then.
CUdeviceptr data_ptr(const tfw_vspace* vs, uint32_t region) {
  const Region& r = vs->regions[region];
  return vs->fixed && r.tier == TFW_TIER_PEER && r.phys ? r.phys->alias : va_of(vs, region);
}

cudaEvent_t get_event(tfw_vspace* vs, int device) {
  auto& pool = vs->dev[device].ev_pool;
  if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaSetDevice(device);
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return e;
}

// The event on the client stream that covers every client kernel which may have touched `r`
// (enqueued between access #last_use and the access after it); null = nothing to wait for.
cudaEvent_t use_mark(tfw_vspace* vs, const Region& r) {
  if (!vs->client || !r.last_use) return nullptr;
  for (const auto& m : vs->marks)
    if (m.seq > r.last_use) return m.ev;
  return nullptr;  // touched by the current access only: nothing of it is enqueued yet
}

// Record "everything the client has enqueued so far" (start of an access that may migrate).
tfw_status push_mark(tfw_vspace* vs) {
  if (!vs->client) return TFW_OK;
  if (!vs->marks.empty() && vs->marks.back().seq == vs->seq) return TFW_OK;
  cudaEvent_t ev;
  if (vs->marks.size() >= 1024) { ev = vs->marks.front().ev; vs->marks.pop_front(); }  // (older regions fall back to a younger mark: conservative)
  else ev = get_event(vs, vs->cfg.home_device);
  RT(vs, cudaSetDevice(vs->cfg.home_device));
  RT(vs, cudaEventRecord(ev, vs->client));
  vs->marks.push_back({vs->seq, ev});
  return TFW_OK;
}

// Host-side wait until no client kernel can still be using the region's current mapping.
tfw_status wait_last_use(tfw_vspace* vs, const Region& r) {
  if (cudaEvent_t m = use_mark(vs, r)) RT(vs, cudaEventSynchronize(m));
  return TFW_OK;
}

// Enqueue the copy of one region towards tier `to` and, for moves INTO the home GPU, re-point the
// region's VA right away (the client stream is made to wait for the copy before it uses the
// region; copies address backing through alias mappings, so they do not care).  Budgets are the
// caller's business.  The destination is accounted now, the source when the move finishes.
tfw_status begin_move(tfw_vspace* vs, uint32_t region, uint32_t to, int32_t slot, bool lazy = false) {
  Region& r = vs->regions[region];
  vs->transits.emplace_back();
  Transit& t = vs->transits.back();
  t.lazy = lazy;
  const bool to_frame = vs->fixed && to == TFW_TIER_HOME;  // the destination is the region's own frame (ensure_frame made it ours)
  auto undo = [&](tfw_status s) {
    if (t.nphys && !to_frame) vs->pool[t.nphys->device].push_back(t.nphys);
    if (to_frame && frame_of(vs, region).occupant == (int32_t)region) frame_of(vs, region).occupant = -1;
    if (t.nhost >= 0) vs->host_free.push_back(t.nhost);
    vs->transits.pop_back();
    return s;
  };
  t.region = region; t.from = r.tier; t.from_slot = r.peer_slot; t.to = to; t.to_slot = to == TFW_TIER_PEER ? slot : -1;
  t.ophys = r.phys; t.ohost = r.host_slot;
  const int home = vs->cfg.home_device;
  cudaEvent_t mark = use_mark(vs, r);
  cudaStream_t st = nullptr;
  if (to == TFW_TIER_HOST) {
    if (vs->host_free.empty()) return undo(vfail(vs, TFW_ERR_EXHAUSTED, "no free host slot"));
    t.nhost = vs->host_free.back();
    vs->host_free.pop_back();
    t.ev_dev = home;
    st = vs->stream;
    RT(vs, cudaSetDevice(home));
    if (mark) RT(vs, cudaStreamWaitEvent(st, mark, 0));
    RT(vs, cudaMemcpyAsync(vs->host_pool + (uint64_t)t.nhost * vs->R, reinterpret_cast<const void*>(r.phys->alias), vs->R, cudaMemcpyDeviceToHost, st));
  } else {
    cudaEvent_t frame_free = nullptr;  // fixed frames: the previous occupant's bytes must have left before ours arrive
    if (to_frame) {
      Frame& f = frame_of(vs, region);
      t.nphys = f.ph;
      if (f.leaving) frame_free = f.leaving->done;
    } else {
      tfw_status s = acquire_phys(vs, device_of(vs, to, slot), &t.nphys);
      if (s != TFW_OK) return undo(s);
    }
    if (t.from == TFW_TIER_HOST) {
      t.ev_dev = home;
      st = vs->stream2;  // host -> device on its own stream: evictions (device -> host) use the other PCIe direction at once
      RT(vs, cudaSetDevice(home));
      if (frame_free) RT(vs, cudaStreamWaitEvent(st, frame_free, 0));
      RT(vs, cudaMemcpyAsync(reinterpret_cast<void*>(t.nphys->alias), vs->host_pool + (uint64_t)r.host_slot * vs->R, vs->R, cudaMemcpyHostToDevice, st));
    } else {
      // receiver-driven: a copy INTO GPU d runs ON GPU d (SM-initiated NVLink reads beat writes on B200)
      t.ev_dev = (vs->cfg.flags & TFW_VS_HOME_DRIVEN) ? home
                 : (vs->cfg.flags & TFW_VS_SENDER_DRIVEN) ? r.phys->device : (vs->cfg.flags & TFW_VS_PUSH_EVICT) ? home : t.nphys->device;
      st = vs->dev[t.ev_dev].stream;
      if ((vs->cfg.flags & TFW_VS_HOME_DRIVEN) && to != TFW_TIER_HOME) st = vs->stream2;  // pushes out and pulls in overlap: one stream each
      RT(vs, cudaSetDevice(t.ev_dev));
      if (mark) RT(vs, cudaStreamWaitEvent(st, mark, 0));
      if (frame_free) RT(vs, cudaStreamWaitEvent(st, frame_free, 0));
      if (vs->cfg.flags & TFW_VS_COPY_ENGINE) {
        RT(vs, cudaMemcpyAsync(reinterpret_cast<void*>(t.nphys->alias), reinterpret_cast<const void*>(r.phys->alias), vs->R, cudaMemcpyDeviceToDevice, st));
      } else {
        tfw_move_desc d{};
        d.dst = (uint64_t)t.nphys->alias; d.src = (uint64_t)r.phys->alias; d.len = vs->R; d.tile0 = 0;
        RT(vs, tfw::launch_mover_inline(&d, 1, tfw::mover_tiles(d.dst, d.len), vs->sm_count, vs->peer_ctas, st));
        vs->st.mover_launches++;
      }
    }
  }
  t.done = get_event(vs, t.ev_dev);
  RT(vs, cudaSetDevice(t.ev_dev));
  RT(vs, cudaEventRecord(t.done, st));
  RT(vs, cudaSetDevice(home));
  if (to_frame) {
    t.va_done = true;  // the region's VA has named this frame all along
  } else if (to == TFW_TIER_HOME && !vs->remap_late) {  // re-point now: by the time the client may use the region its bytes have arrived (access() orders that)
    tfw_status s = wait_last_use(vs, r);  // in-place users of the old (peer) mapping
    if (s != TFW_OK) return s;
    if (r.mapped) { tfw_status u_ = unmap_va(vs, region); if (u_ != TFW_OK) return u_; r.mapped = false; }
    s = point_region(vs, region, t.nphys);
    if (s != TFW_OK) return s;
    r.mapped = true;
    t.va_done = true;
    vs->st.remaps++;
  } else if (to != TFW_TIER_HOME) {
    vs->evictions_in_flight++;
  }
  account(vs, region, to, t.to_slot, +1);
  if (vs->fixed && t.from == TFW_TIER_HOME && to != TFW_TIER_HOME) frame_of(vs, region).leaving = &t;
  r.transit = &t;
  return TFW_OK;
}

// The copy has completed: finish the book-keeping of a move.
tfw_status finish_move(tfw_vspace* vs, Transit* t) {
  Region& r = vs->regions[t->region];
  RT(vs, cudaSetDevice(vs->cfg.home_device));
  if (t->from == TFW_TIER_PEER && t->to == TFW_TIER_HOME) vs->st.prefetch_bytes_peer += vs->R;
  if (t->from == TFW_TIER_HOME && t->to == TFW_TIER_PEER) vs->st.evict_bytes_peer += vs->R;
  if (t->from == TFW_TIER_PEER && t->to == TFW_TIER_PEER) { vs->st.evict_bytes_peer += vs->R; vs->st.prefetch_bytes_peer += vs->R; }
  if (t->to == TFW_TIER_HOST) vs->st.evict_bytes_host += vs->R;
  if (t->from == TFW_TIER_HOST) vs->st.prefetch_bytes_host += vs->R;
  if (!t->va_done && vs->fixed) {  // nothing to re-point: the eviction copy itself was ordered behind the region's last users
    if (t->to != TFW_TIER_HOME) vs->evictions_in_flight--;
  } else if (!t->va_done) {  // the region leaves the home GPU: nobody may still be running on its old mapping
    tfw_status s = wait_last_use(vs, r);
    if (s != TFW_OK) return s;
    if (r.mapped) { tfw_status u_ = unmap_va(vs, t->region); if (u_ != TFW_OK) return u_; r.mapped = false; }
    // A region the POLICY evicted to a peer stays unmapped there: the policy brings a region home before it is used
    // (tfw_vspace_access), so pointing its VA at the peer backing would only be undone by the next prefetch -- and
    // granting the home GPU access to peer-located memory is the expensive VMM call (ms, against 0.1 ms for local memory).
    if (t->to != TFW_TIER_HOST && !t->lazy) {
      s = point_region(vs, t->region, t->nphys);
      if (s != TFW_OK) return s;
      r.mapped = true;
    }
    vs->st.remaps++;
    if (t->to != TFW_TIER_HOME) vs->evictions_in_flight--;
  }
  // source accounting (the destination was accounted when the move began)
  if (t->from == TFW_TIER_HOME) {
    vs->home_used -= vs->R; vs->st.regions_home--;
    if (t->to != TFW_TIER_HOME) vs->lru.erase(r.lru);
  } else if (t->from == TFW_TIER_PEER) { vs->peer_used[t->from_slot] -= vs->R; vs->st.regions_peer--; }
  else if (t->from == TFW_TIER_HOST) { vs->host_used -= vs->R; vs->st.regions_host--; vs->host_free.push_back(t->ohost); }
  if (t->ophys && vs->fixed && t->from == TFW_TIER_HOME) {  // the frame stays where it is; it is free unless somebody claimed it meanwhile
    Frame& f = frame_of(vs, t->region);
    if (f.leaving == t) f.leaving = nullptr;
    if (f.occupant == (int32_t)t->region) f.occupant = -1;
  } else if (t->ophys) {
    const uint64_t used = t->from == TFW_TIER_HOME ? vs->home_used : vs->peer_used[t->from_slot];
    const uint64_t budget = t->from == TFW_TIER_HOME ? vs->cfg.home_budget_bytes : vs->cfg.peer_budget_bytes;
    release_phys(vs, t->ophys, used, budget);
  }
  r.phys = t->nphys;
  r.host_slot = t->nhost;
  r.tier = t->to;
  r.peer_slot = t->to_slot;
  r.transit = nullptr;
  vs->dev[t->ev_dev].ev_pool.push_back(t->done);
  for (auto it = vs->transits.begin(); it != vs->transits.end(); ++it)
    if (&*it == t) { vs->transits.erase(it); break; }
  return TFW_OK;
}

tfw_status wait_transit(tfw_vspace* vs, Transit* t) {
  const auto t0 = std::chrono::steady_clock::now();
  RT(vs, cudaEventSynchronize(t->done));
  vs->st.stall_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  return finish_move(vs, t);
}

// Finish every move whose copy has completed (never blocks).
tfw_status retire_ready(tfw_vspace* vs) {
  for (auto it = vs->transits.begin(); it != vs->transits.end();) {
    Transit* t = &*it;
    ++it;
    const cudaError_t q = cudaEventQuery(t->done);
    if (q == cudaErrorNotReady) { cudaGetLastError(); continue; }
    if (q != cudaSuccess) return vfail(vs, TFW_ERR_FAILED, std::string("migration copy failed: ") + cudaGetErrorString(q));
    tfw_status s = finish_move(vs, t);
    if (s != TFW_OK) return s;
  }
  return TFW_OK;
}

tfw_status quiesce(tfw_vspace* vs) {
  while (!vs->transits.empty()) {
    tfw_status s = wait_transit(vs, &vs->transits.front());
    if (s != TFW_OK) return s;
  }
  return TFW_OK;
}

// A region that is on its way somewhere must have arrived before anything else happens to it; a peer-resident region
// the policy left unmapped gets its VA back if somebody wants to address it there.
tfw_status settle(tfw_vspace* vs, uint32_t region) {
  if (Transit* t = vs->regions[region].transit) { tfw_status s = wait_transit(vs, t); if (s != TFW_OK) return s; }
  Region& r = vs->regions[region];
  if (r.tier == TFW_TIER_PEER && !r.mapped && r.phys && !vs->fixed) {
    tfw_status s = point_region(vs, region, r.phys);
    if (s != TFW_OK) return s;
    r.mapped = true;
  }
  return TFW_OK;
}

// The least-recently-used HOME region that may leave (not pinned, not already moving, not `keep`).
int pick_victim(tfw_vspace* vs, uint32_t keep) {
  for (auto it = vs->lru.rbegin(); it != vs->lru.rend(); ++it) {
    const Region& r = vs->regions[*it];
    if (r.pinned || r.transit || *it == keep || r.tier != TFW_TIER_HOME) continue;
    return (int)*it;
  }
  return -1;
}

// Start evicting HOME region `v`: emptiest peer first, else host.
tfw_status evict_region(tfw_vspace* vs, int v);

// Start evicting one LRU region.  TFW_ERR_NOT_FOUND = no victim.
tfw_status evict_one(tfw_vspace* vs, uint32_t keep) {
  const int v = pick_victim(vs, keep);
  if (v < 0) return TFW_ERR_NOT_FOUND;
  return evict_region(vs, v);
}

tfw_status evict_region(tfw_vspace* vs, int v) {
  int best = -1;
  for (uint32_t p = 0; p < vs->cfg.n_peers; ++p)
    if (vs->peer_used[p] + vs->R <= vs->cfg.peer_budget_bytes && (best < 0 || vs->peer_used[p] < vs->peer_used[best])) best = (int)p;
  if (best < 0 && vs->host_free.empty()) return vfail(vs, TFW_ERR_EXHAUSTED, "no tier has room for an evicted region");
  tfw_status s = begin_move(vs, (uint32_t)v, best >= 0 ? TFW_TIER_PEER : TFW_TIER_HOST, best, !(vs->cfg.flags & TFW_VS_PEER_IN_PLACE));
  if (s == TFW_OK) vs->st.policy_evictions++;
  return s;
}

tfw_status wait_transit(tfw_vspace* vs, Transit* t);

// Fixed-frames mode: make `region`'s frame its own -- whoever lives there starts leaving now (the caller's copy into the
// frame is ordered behind that eviction by begin_move).  may_wait = false: never block the host (an occupant that is
// itself still arriving answers TFW_ERR_NOT_FOUND: try again later).
tfw_status ensure_frame(tfw_vspace* vs, uint32_t region, bool may_wait) {
  Frame& f = frame_of(vs, region);
  if (f.occupant == (int32_t)region) return TFW_OK;
  if (f.occupant >= 0) {
    const uint32_t v = (uint32_t)f.occupant;
    Region& o = vs->regions[v];
    if (o.pinned) return vfail(vs, TFW_ERR_EXHAUSTED, "the region's frame holds a pinned region (fixed-frames mode)");
    if (o.transit && o.transit->to == TFW_TIER_HOME) {  // the occupant is still arriving
      if (!may_wait) return TFW_ERR_NOT_FOUND;
      tfw_status s = wait_transit(vs, o.transit);
      if (s != TFW_OK) return s;
    }
    if (!o.transit && o.tier == TFW_TIER_HOME) {
      tfw_status s = evict_region(vs, (int)v);
      // every cold slot taken: the moves in flight are about to free theirs (a region coming home gives its slot back
      // when its copy has completed)
      while (s == TFW_ERR_EXHAUSTED && may_wait && !vs->transits.empty()) {
        s = wait_transit(vs, &vs->transits.front());
        if (s != TFW_OK) return s;
        s = evict_region(vs, (int)v);
      }
      if (s != TFW_OK) return s;
    }  // else: it is leaving already (f.leaving names that move)
  }
  f.occupant = (int32_t)region;
  return TFW_OK;
}

// How many evictions out of the home GPU have finished COPYING but not yet their book-keeping: their HOME backing is
// as good as free (finish_move will release it), so the next prefetch need not wait for the VA re-mapping.
uint32_t evictions_copied(tfw_vspace* vs) {
  uint32_t n = 0;
  for (auto& t : vs->transits)
    if (!t.va_done && t.from == TFW_TIER_HOME && cudaEventQuery(t.done) == cudaSuccess) ++n;
  cudaGetLastError();
  return n;
}

// Make room for one more HOME region, blocking on an eviction's COPY if it must (never on its re-mapping: the
// cuMemUnmap / cuMemMap / cuMemSetAccess of the region that left happen after the next prefetch has been enqueued,
// so the links stay busy while the host does VMM calls).
tfw_status ensure_home_room(tfw_vspace* vs, uint32_t keep, bool strict = false) {
  for (;;) {
    if (vs->home_used + vs->R <= vs->cfg.home_budget_bytes) return TFW_OK;
    if (vs->home_used + vs->R <= vs->cfg.home_budget_bytes + (uint64_t)evictions_copied(vs) * vs->R) {
      if (!strict) return TFW_OK;
      tfw_status s = retire_ready(vs);  // the caller needs the room in the books too (a first touch allocates against them)
      if (s != TFW_OK) return s;
      continue;
    }
    Transit* oldest = nullptr;
    for (auto& t : vs->transits)
      if (!t.va_done && t.from == TFW_TIER_HOME && cudaEventQuery(t.done) != cudaSuccess) { oldest = &t; break; }
    cudaGetLastError();
    if (!oldest) {
      tfw_status s = evict_one(vs, keep);
      if (s == TFW_ERR_NOT_FOUND) return vfail(vs, TFW_ERR_EXHAUSTED, "home budget exhausted by pinned regions");
      if (s != TFW_OK) return s;
      continue;
    }
    const auto t0 = std::chrono::steady_clock::now();
    RT(vs, cudaSetDevice(oldest->ev_dev));
    RT(vs, cudaEventSynchronize(oldest->done));
    RT(vs, cudaSetDevice(vs->cfg.home_device));
    vs->st.stall_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  }
}

// Keep `slack` regions worth of HOME budget free (or being freed) so that the next misses find room at once.
tfw_status top_up_slack(tfw_vspace* vs, uint32_t keep) {
  const uint64_t slack = (uint64_t)(vs->ahead + 1) * vs->R;
  if (slack >= vs->cfg.home_budget_bytes) return TFW_OK;
  while (vs->home_used + slack > vs->cfg.home_budget_bytes + (uint64_t)vs->evictions_in_flight * vs->R) {
    tfw_status s = evict_one(vs, keep);
    if (s == TFW_ERR_NOT_FOUND || s == TFW_ERR_EXHAUSTED) return TFW_OK;  // best effort
    if (s != TFW_OK) return s;
  }
  return TFW_OK;
}

// The client stream (or, without one, this thread) waits until the region's bytes have arrived.
tfw_status order_after(tfw_vspace* vs, Transit* t) {
  if (vs->client) { RT(vs, cudaSetDevice(vs->cfg.home_device)); RT(vs, cudaStreamWaitEvent(vs->client, t->done, 0)); return TFW_OK; }
  return wait_transit(vs, t);
}

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
  vs->ahead = std::min<uint32_t>(cfg->prefetch_ahead, 8);
  if (const char* e = getenv("TFW_VS_PEER_CTAS")) vs->peer_ctas = std::max(0, atoi(e));
  vs->remap_late = (cfg->flags & TFW_VS_REMAP_LATE) != 0;
  if (const char* e = getenv("TFW_VS_REMAP_LATE")) vs->remap_late = e[0] == '1';
  vs->fixed = (cfg->flags & TFW_VS_FIXED_FRAMES) != 0;
  if (vs->fixed) {
    // a region's VA names its frame for ever: it cannot also name the region's peer backing
    if ((cfg->flags & TFW_VS_PEER_IN_PLACE) || cfg->home_budget_bytes < cfg->region_bytes) { delete vs; return TFW_ERR_INVALID; }
    vs->remap_late = false;
  }
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
  vs->alias_slots = (uint64_t)vs->n + kWindowSlots + (vs->fixed ? cfg->home_budget_bytes / vs->R : 0);
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
  if (vs->fixed) {
    // every frame is created now and mapped at the VA of every region that will ever use it; one cuMemSetAccess over
    // the whole space; from here on no migration makes a VMM call
    const uint32_t nframes = (uint32_t)std::min<uint64_t>(cfg->home_budget_bytes / vs->R, vs->n);
    vs->frames.resize(nframes);
    for (Frame& f : vs->frames) {
      tfw_status s = acquire_phys(vs, cfg->home_device, &f.ph);
      if (s != TFW_OK) return bail(s);
    }
    uint32_t mapped = 0;
    for (; mapped < vs->n; ++mapped)
      if (g_drv.cuMemMap(va_of(vs, mapped), vs->R, 0, frame_of(vs, mapped).ph->h, 0) != CUDA_SUCCESS) break;
    for (uint32_t r = 0; r < mapped; ++r) vs->regions[r].mapped = true;
    if (mapped < vs->n) return bail(TFW_ERR_FAILED);
    CUmemAccessDesc a{};
    a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    a.location.id = cfg->home_device;
    a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (g_drv.cuMemSetAccess(vs->base, cfg->va_bytes, &a, 1) != CUDA_SUCCESS) return bail(TFW_ERR_FAILED);
    vs->ahead = std::min<uint32_t>(vs->ahead, nframes - 1);
  }
  *out = vs;
  return TFW_OK;
}

tfw_status tfw_vspace_destroy(tfw_vspace* vs) {
  if (!vs) return TFW_ERR_INVALID;
  cudaSetDevice(vs->cfg.home_device);
  quiesce(vs);
  if (getenv("TFW_VS_DEBUG")) {
    auto pr = [](const char* what, const tfw_vspace::VmmDbg& d) {
      fprintf(stderr, "[tfw_vspace] %-14s calls %8llu  avg %8.1f us  max %9.1f us  over 0.5 ms: %llu\n", what, (unsigned long long)d.n, d.n ? d.ns / 1e3 / d.n : 0.0,
              d.max_ns / 1e3, (unsigned long long)d.slow);
    };
    pr("cuMemUnmap", vs->dbg_unmap); pr("cuMemMap", vs->dbg_map); pr("cuMemSetAccess", vs->dbg_access);
  }
  if (vs->stream) cudaStreamSynchronize(vs->stream);
  for (auto& m : vs->marks) cudaEventDestroy(m.ev);
  for (auto& dc : vs->dev) for (cudaEvent_t e : dc.ev_pool) cudaEventDestroy(e);
  for (uint32_t i = 0; i < vs->n && i < vs->regions.size(); ++i) {
    Region& r = vs->regions[i];
    if (vs->fixed) {  // region VAs belong to the frames; a HOME region's backing is its frame
      if (r.mapped) g_drv.cuMemUnmap(va_of(vs, i), vs->R);
      if (r.tier == TFW_TIER_PEER && r.phys) destroy_phys(vs, r.phys);
      continue;
    }
    if ((r.tier == TFW_TIER_HOME || r.tier == TFW_TIER_PEER) && r.phys) {
      if (r.mapped) g_drv.cuMemUnmap(va_of(vs, i), vs->R);
      destroy_phys(vs, r.phys);
    }
  }
  for (Frame& f : vs->frames) if (f.ph) destroy_phys(vs, f.ph);
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
  // a region on its way INTO the home GPU already answers at its HOME address (tfw_vspace_access orders the client
  // behind the copy); one on its way out still lives where it was
  const uint32_t t = r.transit && r.transit->va_done ? (uint32_t)TFW_TIER_HOME : r.tier;
  if (tier) *tier = t;
  if (device) *device = t == TFW_TIER_HOME ? vs->cfg.home_device : t == TFW_TIER_PEER ? vs->cfg.peer_devices[r.peer_slot] : -1;
  return TFW_OK;
}

tfw_status tfw_vspace_populate(tfw_vspace* vs, uint32_t region, uint32_t tier, int32_t peer_slot) {
  if (!vs || region >= vs->n || tier == TFW_TIER_NONE || tier > TFW_TIER_HOST) return TFW_ERR_INVALID;
  Region& r = vs->regions[region];
  if (r.tier != TFW_TIER_NONE || r.transit) return vfail(vs, TFW_ERR_INVALID, "region already backed");
  const bool to_frame = vs->fixed && tier == TFW_TIER_HOME;  // (the frames are the home budget)
  if (!to_frame && !budget_ok(vs, tier, peer_slot)) return vfail(vs, TFW_ERR_EXHAUSTED, "tier budget exhausted");
  cudaSetDevice(vs->cfg.home_device);
  if (tier == TFW_TIER_HOST) {
    r.host_slot = vs->host_free.back();
    vs->host_free.pop_back();
    std::memset(vs->host_pool + (uint64_t)r.host_slot * vs->R, 0, vs->R);
  } else {
    tfw_status s = TFW_OK;
    if (to_frame) {
      s = ensure_frame(vs, region, true);
      if (s != TFW_OK) return s;
      Frame& f = frame_of(vs, region);
      r.phys = f.ph;
      if (f.leaving) RT(vs, cudaStreamWaitEvent(vs->stream, f.leaving->done, 0));  // scrub after the previous occupant has left
    } else {
      s = acquire_phys(vs, device_of(vs, tier, peer_slot), &r.phys);
      if (s != TFW_OK) return s;
      if (!vs->fixed) {
        s = point_region(vs, region, r.phys);
        if (s != TFW_OK) return s;
        r.mapped = true;
      }
    }
    tfw_move_desc d{};
    d.dst = (uint64_t)(vs->fixed ? r.phys->alias : va_of(vs, region));
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
  tfw_status rc = quiesce(vs);  // an explicit batch starts from a settled state (its timing means something then)
  if (rc != TFW_OK) return rc;
  ++vs->seq;                    // with a bound client stream: everything it has enqueued so far happens before these moves
  rc = push_mark(vs);
  if (rc != TFW_OK) return rc;
  tfw_migrate_result acc{};
  for (uint32_t off = 0; off < n; off += kWindowSlots) {
    const uint32_t m = std::min(kWindowSlots, n - off);
    struct Req { uint32_t region; uint32_t to; int32_t slot; bool noop; };
    std::vector<Req> mv(m);
    // ---- validate -------------------------------------------------------------------
    for (uint32_t k = 0; k < m; ++k) {
      Req& x = mv[k];
      x.region = regions[off + k];
      x.to = tiers[off + k];
      x.slot = peer_slots ? peer_slots[off + k] : -1;
      if (x.region >= vs->n || x.to == TFW_TIER_NONE || x.to > TFW_TIER_HOST) return vfail(vs, TFW_ERR_INVALID, "bad migrate request");
      for (uint32_t j = 0; j < k; ++j) if (mv[j].region == x.region) return vfail(vs, TFW_ERR_INVALID, "region listed twice in one batch");
      const Region& r = vs->regions[x.region];
      if (r.tier == TFW_TIER_NONE) return vfail(vs, TFW_ERR_INVALID, "region not populated");
      x.noop = r.tier == x.to && (x.to != TFW_TIER_PEER || r.peer_slot == x.slot);
    }
    // budgets are checked against the state after the moves of this window (regions leaving a tier make room in it)
    uint64_t home_plan = vs->home_used, host_plan = vs->host_free.size();
    std::vector<uint64_t> peer_plan = vs->peer_used;
    for (uint32_t k = 0; k < m; ++k) {
      if (mv[k].noop) continue;
      const Region& r = vs->regions[mv[k].region];
      if (r.tier == TFW_TIER_HOME) home_plan -= vs->R;
      else if (r.tier == TFW_TIER_PEER) peer_plan[r.peer_slot] -= vs->R;
    }
    uint64_t host_need = 0;
    for (uint32_t k = 0; k < m; ++k) {
      const Req& x = mv[k];
      if (x.noop) continue;
      bool ok = true;
      if (x.to == TFW_TIER_HOME) { ok = home_plan + vs->R <= vs->cfg.home_budget_bytes; home_plan += vs->R; }
      else if (x.to == TFW_TIER_PEER) { ok = x.slot >= 0 && (uint32_t)x.slot < vs->cfg.n_peers && peer_plan[x.slot] + vs->R <= vs->cfg.peer_budget_bytes; if (ok) peer_plan[x.slot] += vs->R; }
      else { ok = ++host_need <= host_plan; }  // host slots freed by this window only return when it has completed
      if (!ok) return vfail(vs, TFW_ERR_EXHAUSTED, x.to == TFW_TIER_HOST ? "no free host slot until this batch completes; split the batch" : "target tier budget exhausted");
    }
    // ---- copies: every GPU that receives data runs its own, all at once ------------------
    for (auto& dc : vs->dev) dc.used = false;
    for (const Req& x : mv) {
      if (x.noop) continue;
      const Region& r = vs->regions[x.region];
      const int d = (x.to == TFW_TIER_HOST || r.tier == TFW_TIER_HOST || (vs->cfg.flags & TFW_VS_HOME_DRIVEN)) ? vs->cfg.home_device
                    : (vs->cfg.flags & TFW_VS_SENDER_DRIVEN) ? r.phys->device
                    : (vs->cfg.flags & TFW_VS_PUSH_EVICT) ? vs->cfg.home_device : device_of(vs, x.to, x.slot);
      vs->dev[d].used = true;
    }
    for (size_t d = 0; d < vs->dev.size(); ++d) {
      if (!vs->dev[d].used) continue;
      RT(vs, cudaSetDevice((int)d));
      RT(vs, cudaEventRecord(vs->dev[d].e0, vs->dev[d].stream));
    }
    RT(vs, cudaSetDevice(vs->cfg.home_device));
    RT(vs, cudaStreamWaitEvent(vs->stream2, vs->e0, 0));  // host -> device DMAs of the window start with it
    const uint64_t launches0 = vs->st.mover_launches;
    // fixed frames: the moves out of the home GPU are begun first, so that a region coming home in the same window finds
    // its frame's occupant already leaving (begin_move orders the copy in behind that copy out)
    for (int pass = vs->fixed ? 0 : 1; pass < 2; ++pass) {
      for (const Req& x : mv) {
        if (x.noop) continue;
        if (vs->fixed) {
          const bool leaves_home = vs->regions[x.region].tier == TFW_TIER_HOME && x.to != TFW_TIER_HOME;
          if (leaves_home != (pass == 0)) continue;
          if (x.to == TFW_TIER_HOME) {
            Frame& f = frame_of(vs, x.region);
            if (f.occupant >= 0 && f.occupant != (int32_t)x.region) {
              const Region& o = vs->regions[(uint32_t)f.occupant];
              if (!(o.transit && o.transit->to != TFW_TIER_HOME)) {
                quiesce(vs);
                return vfail(vs, TFW_ERR_EXHAUSTED, "the region's frame is occupied (fixed-frames mode): move its occupant out in the same batch or before");
              }
            }
            f.occupant = (int32_t)x.region;
          }
        }
        rc = begin_move(vs, x.region, x.to, x.slot);
        if (rc != TFW_OK) { quiesce(vs); return rc; }
        acc.bytes += vs->R;
      }
    }
    acc.launches += (uint32_t)(vs->st.mover_launches - launches0);
    RT(vs, cudaSetDevice(vs->cfg.home_device));
    RT(vs, cudaEventRecord(vs->e2, vs->stream2));
    RT(vs, cudaStreamWaitEvent(vs->stream, vs->e2, 0));
    for (size_t d = 0; d < vs->dev.size(); ++d) {
      if (!vs->dev[d].used) continue;
      RT(vs, cudaSetDevice((int)d));
      RT(vs, cudaEventRecord(vs->dev[d].e1, vs->dev[d].stream));
    }
    // Finish each move as its copy completes: re-pointing region k (cuMemUnmap / cuMemMap / cuMemSetAccess, ~0.1 ms each
    // and growing with the number of GPUs mapped) overlaps the copies of the regions behind it instead of trailing the batch.
    rc = quiesce(vs);
    if (rc != TFW_OK) return rc;
    float batch_ms = 0;
    for (size_t d = 0; d < vs->dev.size(); ++d) {
      if (!vs->dev[d].used) continue;
      RT(vs, cudaSetDevice((int)d));
      RT(vs, cudaEventSynchronize(vs->dev[d].e1));
      float ms = 0;
      RT(vs, cudaEventElapsedTime(&ms, vs->dev[d].e0, vs->dev[d].e1));
      batch_ms = std::max(batch_ms, ms);  // the GPUs copy concurrently: the batch takes as long as the slowest
    }
    RT(vs, cudaSetDevice(vs->cfg.home_device));
    acc.copy_ms += batch_ms;
  }
  acc.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (res) *res = acc;
  return TFW_OK;
}

tfw_status tfw_vspace_bind_stream(tfw_vspace* vs, void* cuda_stream) {
  if (!vs) return TFW_ERR_INVALID;
  tfw_status s = quiesce(vs);
  if (s != TFW_OK) return s;
  vs->client = static_cast<cudaStream_t>(cuda_stream);
  for (auto& m : vs->marks) vs->dev[vs->cfg.home_device].ev_pool.push_back(m.ev);
  vs->marks.clear();
  return TFW_OK;
}

tfw_status tfw_vspace_quiesce(tfw_vspace* vs) {
  if (!vs) return TFW_ERR_INVALID;
  cudaSetDevice(vs->cfg.home_device);
  return quiesce(vs);
}

// Policy entry.  Migrations are asynchronous: a miss enqueues the wanted region's copy (and, on a
// sequential sweep, those of the next `prefetch_ahead` regions), re-points its VA, makes the client
// stream wait for the bytes, and starts evicting least-recently-used regions so that the NEXT miss
// finds room at once; evictions and prefetches run in opposite NVLink directions at the same time
// and nothing here waits for them unless the home budget is truly exhausted.
tfw_status tfw_vspace_access(tfw_vspace* vs, uint32_t region) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  cudaSetDevice(vs->cfg.home_device);
  Region& r = vs->regions[region];
  ++vs->seq;
  tfw_status s = TFW_OK;
  const bool sequential = region == vs->last_access + 1;
  vs->last_access = region;
  if (r.tier == TFW_TIER_HOME && !r.transit) {  // hit: nothing to move, nothing to record
    vs->lru.erase(r.lru);
    vs->lru.push_front(region);
    r.lru = vs->lru.begin();
    r.last_use = vs->seq;
    vs->st.policy_hits++;
    if (!sequential || !vs->ahead) return TFW_OK;
  }
  s = push_mark(vs);
  if (s != TFW_OK) return s;
  if (vs->fixed) {
    // Nothing else makes this thread wait in fixed-frames mode (a region's arrival is ordered behind its victim's departure
    // on the GPU), so a client that runs ahead of the links would queue moves -- and the backings they hold -- without
    // bound: keep the pipeline a few regions deep in each direction.
    const size_t depth = 2 * ((size_t)vs->ahead + 1) + 2;
    while (vs->transits.size() >= depth) {
      s = wait_transit(vs, &vs->transits.front());
      if (s != TFW_OK) return s;
    }
  }
  bool arrived = false;
  if (r.transit && !r.transit->va_done) {
    // on its way OUT: let it arrive, then treat it as the miss it is; on its way IN with the VA not re-pointed yet
    // (TFW_VS_REMAP_LATE): its copy must have completed before the VA can name it
    arrived = r.transit->to == TFW_TIER_HOME;
    s = wait_transit(vs, r.transit);
    if (s != TFW_OK) return s;
  }
  if (arrived) {
    vs->lru.erase(r.lru);
    vs->lru.push_front(region);
    r.lru = vs->lru.begin();
    vs->st.policy_hits_inflight++;
  } else if (r.transit) {  // on its way IN (prefetched ahead): the client stream waits for the bytes, the host does not
    s = order_after(vs, r.transit);
    if (s != TFW_OK) return s;
    vs->lru.erase(r.lru);
    vs->lru.push_front(region);
    r.lru = vs->lru.begin();
    vs->st.policy_hits_inflight++;
  } else if (r.tier == TFW_TIER_NONE) {  // first touch: fresh zero-filled HOME backing
    if (!vs->fixed) s = ensure_home_room(vs, region, true);  // (populate claims the frame itself)
    if (s != TFW_OK) return s;
    s = tfw_vspace_populate(vs, region, TFW_TIER_HOME, -1);
    if (s != TFW_OK) return s;
  } else if (r.tier != TFW_TIER_HOME) {  // miss
    s = vs->fixed ? ensure_frame(vs, region, true) : ensure_home_room(vs, region);
    if (s != TFW_OK) return s;
    s = begin_move(vs, region, TFW_TIER_HOME, -1);
    if (s != TFW_OK) return s;
    vs->st.policy_prefetches++;
    s = vs->remap_late ? wait_transit(vs, vs->regions[region].transit) : order_after(vs, vs->regions[region].transit);
    if (s != TFW_OK) return s;
  }
  vs->regions[region].last_use = vs->seq;
  // sequential sweep: start on the regions the client will want next, while it works on this one
  if (sequential && vs->ahead) {
    for (uint32_t j = 1; j <= vs->ahead && region + j < vs->n; ++j) {
      Region& nx = vs->regions[region + j];
      if (nx.transit || (nx.tier != TFW_TIER_PEER && nx.tier != TFW_TIER_HOST)) continue;
      if (vs->fixed) {  // its frame's occupant starts leaving now, its own copy follows that one on the GPU
        if (ensure_frame(vs, region + j, false) != TFW_OK) break;
      } else if (vs->home_used + vs->R > vs->cfg.home_budget_bytes + (uint64_t)evictions_copied(vs) * vs->R) break;  // no room yet: the evictions below make it
      if (begin_move(vs, region + j, TFW_TIER_HOME, -1) != TFW_OK) break;
      vs->st.policy_prefetches++;
      vs->st.policy_prefetch_ahead++;
    }
  }
  if (!vs->fixed) s = top_up_slack(vs, region);  // (fixed frames: a region's arrival names its own victim)
  if (s != TFW_OK) return s;
  // last, with every copy of this access enqueued: the book-keeping (VMM calls) of the moves that have completed
  return vs->transits.empty() ? TFW_OK : retire_ready(vs);
}

tfw_status tfw_vspace_sweep(tfw_vspace* vs, uint32_t first, uint32_t count, uint64_t* digests, double* seconds) {
  if (!vs || !count || !digests || first >= vs->n) return TFW_ERR_INVALID;
  cudaSetDevice(vs->cfg.home_device);
  cudaStream_t bound = vs->client;
  cudaStream_t st = bound;
  if (!st) {  // no client stream bound: sweep on one of our own, bound for the duration
    RT(vs, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    tfw_status b = tfw_vspace_bind_stream(vs, st);
    if (b != TFW_OK) { cudaStreamDestroy(st); return b; }
  }
  unsigned long long* d_sums = nullptr;
  std::vector<unsigned long long> sums(count);
  tfw_status rc = TFW_OK;
  if (cudaMalloc(reinterpret_cast<void**>(&d_sums), sizeof(unsigned long long) * count) != cudaSuccess) { cudaGetLastError(); rc = vfail(vs, TFW_ERR_EXHAUSTED, "no memory for the sweep's digests"); }
  if (rc == TFW_OK && (cudaMemsetAsync(d_sums, 0, sizeof(unsigned long long) * count, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)) rc = vfail(vs, TFW_ERR_FAILED, "sweep setup failed");
  const auto t0 = std::chrono::steady_clock::now();
  for (uint32_t i = 0; i < count && rc == TFW_OK; ++i) {
    const uint32_t region = (first + i) % vs->n;
    rc = tfw_vspace_access(vs, region);
    if (rc != TFW_OK) break;
    // the client's work on the region: read every byte of it (through the region's own VA)
    if (tfw::launch_digest(reinterpret_cast<void*>(va_of(vs, region)), vs->R, d_sums + i, vs->sm_count, st) != cudaSuccess) rc = vfail(vs, TFW_ERR_FAILED, "digest launch failed");
  }
  if (rc == TFW_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = vfail(vs, TFW_ERR_FAILED, "sweep failed on the device");
  if (rc == TFW_OK) rc = quiesce(vs);  // the evictions the sweep started are part of its cost
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc == TFW_OK && cudaMemcpy(sums.data(), d_sums, sizeof(unsigned long long) * count, cudaMemcpyDeviceToHost) != cudaSuccess) rc = vfail(vs, TFW_ERR_FAILED, "sweep read-back failed");
  for (uint32_t i = 0; i < count && rc == TFW_OK; ++i) digests[i] = tfw::digest_mix((uint64_t)sums[i] ^ (vs->R * tfw::kDigestK1));
  cudaGetLastError();
  if (d_sums) cudaFree(d_sums);
  if (!bound) { quiesce(vs); cudaStreamSynchronize(st); tfw_vspace_bind_stream(vs, nullptr); cudaStreamDestroy(st); }
  return rc;
}

tfw_status tfw_vspace_unpopulate(tfw_vspace* vs, uint32_t region) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  cudaSetDevice(vs->cfg.home_device);
  tfw_status st = settle(vs, region);
  if (st != TFW_OK) return st;
  Region& r = vs->regions[region];
  if (r.tier == TFW_TIER_NONE) return TFW_OK;
  st = wait_last_use(vs, r);
  if (st != TFW_OK) return st;
  const uint32_t from = r.tier;
  const int32_t from_slot = r.peer_slot;
  if (from == TFW_TIER_HOST) {
    vs->host_free.push_back(r.host_slot);
    r.host_slot = -1;
  } else if (!vs->fixed) {
    if (r.mapped) { tfw_status u_ = unmap_va(vs, region); if (u_ != TFW_OK) return u_; r.mapped = false; }
  }
  account(vs, region, from, from_slot, -1);
  if (r.phys && vs->fixed && from == TFW_TIER_HOME) {
    Frame& f = frame_of(vs, region);
    if (f.occupant == (int32_t)region) f.occupant = -1;
    r.phys = nullptr;
  } else if (r.phys) {
    const uint64_t used = from == TFW_TIER_HOME ? vs->home_used : vs->peer_used[from_slot];
    const uint64_t budget = from == TFW_TIER_HOME ? vs->cfg.home_budget_bytes : vs->cfg.peer_budget_bytes;
    release_phys(vs, r.phys, used, budget);
    r.phys = nullptr;
  }
  r.tier = TFW_TIER_NONE;
  r.peer_slot = -1;
  r.pinned = 0;
  return TFW_OK;
}

tfw_status tfw_vspace_pin(tfw_vspace* vs, uint32_t region, int pinned) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  Region& r = vs->regions[region];
  if (pinned) ++r.pinned;
  else if (r.pinned) --r.pinned;
  return TFW_OK;
}

tfw_status tfw_vspace_get_stats(tfw_vspace* vs, tfw_vspace_stats* out) {
  if (!vs || !out) return TFW_ERR_INVALID;
  *out = vs->st;
  return TFW_OK;
}

tfw_status tfw_vspace_fill_pattern(tfw_vspace* vs, uint32_t region, uint64_t seed) {
  if (!vs || region >= vs->n) return TFW_ERR_INVALID;
  { cudaSetDevice(vs->cfg.home_device); tfw_status st_ = settle(vs, region); if (st_ != TFW_OK) return st_; }
  const Region& r = vs->regions[region];
  if (r.tier != TFW_TIER_HOME && r.tier != TFW_TIER_PEER) return TFW_ERR_NOT_SUPPORTED;
  cudaSetDevice(vs->cfg.home_device);
  RT(vs, tfw::launch_pattern(reinterpret_cast<void*>(data_ptr(vs, region)), vs->R, seed, vs->sm_count, vs->stream));
  RT(vs, cudaStreamSynchronize(vs->stream));
  return TFW_OK;
}

tfw_status tfw_vspace_digest(tfw_vspace* vs, uint32_t region, uint64_t* digest) {
  if (!vs || region >= vs->n || !digest) return TFW_ERR_INVALID;
  { cudaSetDevice(vs->cfg.home_device); tfw_status st_ = settle(vs, region); if (st_ != TFW_OK) return st_; }
  const Region& r = vs->regions[region];
  if (r.tier != TFW_TIER_HOME && r.tier != TFW_TIER_PEER) return TFW_ERR_NOT_SUPPORTED;
  cudaSetDevice(vs->cfg.home_device);
  RT(vs, cudaMemsetAsync(vs->d_digest, 0, 8, vs->stream));
  RT(vs, tfw::launch_digest(reinterpret_cast<void*>(data_ptr(vs, region)), vs->R, vs->d_digest, vs->sm_count, vs->stream));
  unsigned long long sum = 0;
  RT(vs, cudaMemcpyAsync(&sum, vs->d_digest, 8, cudaMemcpyDeviceToHost, vs->stream));
  RT(vs, cudaStreamSynchronize(vs->stream));
  *digest = tfw::digest_mix((uint64_t)sum ^ (vs->R * tfw::kDigestK1));
  return TFW_OK;
}

tfw_status tfw_vspace_read(tfw_vspace* vs, uint32_t region, uint64_t off, void* dst, uint64_t nbytes) {
  if (!vs || region >= vs->n || !dst || off > vs->R || nbytes > vs->R - off) return TFW_ERR_INVALID;
  { cudaSetDevice(vs->cfg.home_device); tfw_status st_ = settle(vs, region); if (st_ != TFW_OK) return st_; }
  const Region& r = vs->regions[region];
  cudaSetDevice(vs->cfg.home_device);
  if (r.tier == TFW_TIER_HOST) { std::memcpy(dst, vs->host_pool + (uint64_t)r.host_slot * vs->R + off, nbytes); return TFW_OK; }
  if (r.tier == TFW_TIER_NONE) return TFW_ERR_NOT_SUPPORTED;
  RT(vs, cudaMemcpyAsync(dst, reinterpret_cast<void*>(data_ptr(vs, region) + off), nbytes, cudaMemcpyDeviceToHost, vs->stream));
  RT(vs, cudaStreamSynchronize(vs->stream));
  return TFW_OK;
}

tfw_status tfw_vspace_write(tfw_vspace* vs, uint32_t region, uint64_t off, const void* src, uint64_t nbytes) {
  if (!vs || region >= vs->n || !src || off > vs->R || nbytes > vs->R - off) return TFW_ERR_INVALID;
  { cudaSetDevice(vs->cfg.home_device); tfw_status st_ = settle(vs, region); if (st_ != TFW_OK) return st_; }
  const Region& r = vs->regions[region];
  cudaSetDevice(vs->cfg.home_device);
  if (r.tier == TFW_TIER_HOST) { std::memcpy(vs->host_pool + (uint64_t)r.host_slot * vs->R + off, src, nbytes); return TFW_OK; }
  if (r.tier == TFW_TIER_NONE) return TFW_ERR_NOT_SUPPORTED;
  RT(vs, cudaMemcpyAsync(reinterpret_cast<void*>(data_ptr(vs, region) + off), src, nbytes, cudaMemcpyHostToDevice, vs->stream));
  RT(vs, cudaStreamSynchronize(vs->stream));
  return TFW_OK;
}

}  // extern "C"
