// worker.cu -- host side of the B200 vGPU worker data path (libtfw_b200.so).
//
// Role in the reference architecture: this is the body of the process the
// operator starts as `./tensor-fusion-worker -p 8000`
// (reference: internal/utils/compose.go:1304-1325); the reference keeps it
// closed-source (README.md:131), so behaviour is defined by DESIGN.md and
// pinned by oracle/replay_oracle.c.
//
// Pipeline per batch (north_star (a)):
//   host thread : deserialize frames -> descriptors (+ hazard tracking)
//   copy stream : cudaMemcpyAsync(pinned ring / caller's pinned memory -> HBM slot)
//   exec stream : wait(dma) -> tfw_mover kernel (slot -> client buffers) -> event
// so deserialize(n+1) overlaps DMA(n) overlaps unpack(n-1).  Everything a
// client can observe executes in stream order on the exec stream.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "gate.h"
#include "kernels.h"
#include "tfw_stats_file.h"
#include "tfw_vram.h"
#include "tfw_worker.h"

namespace {

using tfw::kInlineDescs;

// Disjoint, merged half-open intervals of device addresses.
class IntervalSet {
 public:
  bool overlaps(uint64_t a, uint64_t b) const {
    if (a >= b || m_.empty()) return false;
    auto it = m_.upper_bound(a);  // first start > a
    if (it != m_.begin()) {
      auto p = std::prev(it);
      if (p->second > a) return true;
    }
    return it != m_.end() && it->first < b;
  }
  void add(uint64_t a, uint64_t b) {
    if (a >= b) return;
    auto it = m_.upper_bound(a);
    if (it != m_.begin()) {
      auto p = std::prev(it);
      if (p->second >= a) { a = p->first; b = std::max(b, p->second); it = m_.erase(p); }
    }
    while (it != m_.end() && it->first <= b) { b = std::max(b, it->second); it = m_.erase(it); }
    m_.emplace(a, b);
  }
  void clear() { m_.clear(); }

 private:
  std::map<uint64_t, uint64_t> m_;
};

struct Buffer {
  uint64_t ptr = 0, size = 0;
  bool live = false;
  bool tiered = false;          // lives in the tiered vGPU address space
  uint32_t region0 = 0, nregions = 0;
  uint8_t* parked = nullptr;    // host copy while the vGPU is frozen (plain buffers; ptr is 0 then)
};

struct Slot {
  uint8_t* host = nullptr;  // pinned staging for non-pinned input
  uint8_t* dev = nullptr;   // HBM staging slot
  tfw_move_desc* h_descs = nullptr;
  tfw_move_desc* d_descs = nullptr;
  cudaEvent_t dma_done = nullptr, exec_done = nullptr;
  bool busy = false;
};

struct Response {
  tfcs_frame_hdr hdr;
  uint8_t* host = nullptr;
  uint64_t len = 0;
  cudaEvent_t ev = nullptr;
  uint64_t sent = 0;  // bytes of the framed response (header + padded payload) already handed out
  bool owned = false; // `host` is a malloc'd blob of this response (RESP_FUNCTION), not arena memory
};

// Client memory the worker has mapped and page-locked too (TFCS_OP_HOST_REGISTER).
struct Arena {
  uint8_t* base = nullptr;
  uint64_t size = 0;
};

// A piece of the worker -> client ring whose bytes become valid when `ev` completes (response sink).
struct SinkPiece {
  uint64_t upto = 0;          // ring cursor after this piece
  cudaEvent_t ev = nullptr;   // null: valid as soon as everything before it is
};

// A loaded user module / resolved kernel (TFCS_OP_MODULE_LOAD, MODULE_GET_FUNCTION).
struct UserFunction {
  CUfunction fn = nullptr;
  uint32_t module = 0;
  uint32_t param_bytes = 0;
  std::vector<std::pair<uint32_t, uint32_t>> params;  // offset, size
};

enum StepKind : uint32_t { kStepMover, kStepLaunch, kStepD2H, kStepSync };
struct Step {
  StepKind kind;
  // mover
  uint64_t desc_off = 0;  // index into the trace's descriptor array
  uint32_t ndesc = 0, tiles = 0;
  bool bulk_ok = false;   // every copy of the batch is congruent modulo 16: eligible for the TMA mover
  // launch / d2h
  tfcs_frame_hdr hdr;
  uint64_t ptr = 0;
};

constexpr uint32_t kMaxDescsPerBatch = 1u << 16;
constexpr uint64_t kDefaultChunk = 32ull << 20;
constexpr uint32_t kDefaultSlots = 4;
constexpr uint64_t kSlotSlack = 64;  // alignment slack in front of each slot
// A resident trace needs no staging slots, so its batches are cut only by hazards,
// launches and this cap: big launches amortise the ramp-up/tail of a grid
// (12 us for a 32 MiB batch vs 6.7 TB/s sustained on GiB-sized ones, profiles/r01).
constexpr uint64_t kResidentBatchBytes = 4ull << 30;
constexpr uint64_t kMaxModuleBytes = 512ull << 20;  // one code image (cubin / PTX / fatbin)

// Driver entry points for user modules, resolved through the runtime so the library keeps
// loading on hosts without libcuda (it then answers TFW_ERR_NO_DEVICE like everything else).
struct ModDrv {
  CUresult (*cuModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*cuModuleUnload)(CUmodule) = nullptr;
  CUresult (*cuModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*cuFuncGetParamInfo)(CUfunction, size_t, size_t*, size_t*) = nullptr;
  CUresult (*cuFuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*cuGetErrorString)(CUresult, const char**) = nullptr;
  // green contexts: the vGPU's kernels are confined to a share of the SMs (hard compute isolation)
  CUresult (*cuDeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*cuDeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*cuDevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
  CUresult (*cuDevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*cuGreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*cuGreenCtxDestroy)(CUgreenCtx) = nullptr;
  CUresult (*cuGreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  bool green = false;
  bool ok = false, tried = false;
  bool load() {
    if (tried) return ok;
    tried = true;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    ok = get("cuModuleLoadData", (void**)&cuModuleLoadData) && get("cuModuleUnload", (void**)&cuModuleUnload) &&
         get("cuModuleGetFunction", (void**)&cuModuleGetFunction) && get("cuFuncGetParamInfo", (void**)&cuFuncGetParamInfo) &&
         get("cuFuncSetAttribute", (void**)&cuFuncSetAttribute) && get("cuLaunchKernel", (void**)&cuLaunchKernel) &&
         get("cuGetErrorString", (void**)&cuGetErrorString);
    green = get("cuDeviceGet", (void**)&cuDeviceGet) && get("cuDeviceGetDevResource", (void**)&cuDeviceGetDevResource) &&
            get("cuDevSmResourceSplitByCount", (void**)&cuDevSmResourceSplitByCount) &&
            get("cuDevResourceGenerateDesc", (void**)&cuDevResourceGenerateDesc) && get("cuGreenCtxCreate", (void**)&cuGreenCtxCreate) &&
            get("cuGreenCtxDestroy", (void**)&cuGreenCtxDestroy) && get("cuGreenCtxStreamCreate", (void**)&cuGreenCtxStreamCreate);
    cudaGetLastError();
    return ok;
  }
};
ModDrv g_mod;

}  // namespace

struct tfw_trace {
  std::vector<Step> steps;
  std::vector<tfw_move_desc> descs;  // host copy
  tfw_move_desc* d_descs = nullptr;
  uint8_t* d_stream = nullptr;  // resident copy of the wire bytes
  std::vector<uint64_t> allocs;  // device buffers owned by the trace
  std::vector<Buffer> bufs;      // handle table at the end of the trace
  uint64_t payload_bytes = 0, mover_launches = 0, algo_bytes = 0, staged_in_batch = 0;
  const uint8_t* host_base = nullptr;
  uint64_t dev_base = 0;  // device address corresponding to host_base
};

struct tfw_worker {
  int device = 0;
  int sm_count = 148;
  tfw_config cfg{};
  cudaStream_t copy_stream = nullptr, exec_stream = nullptr;
  CUgreenCtx green = nullptr;      // non-null: exec_stream belongs to a green context holding `sm_count` SMs
  int device_sms = 148;            // SMs of the whole GPU
  uint32_t sm_percent = 0;         // hard compute limit in force (0 = the whole GPU)
  std::vector<Slot> slots;
  uint32_t cur = 0;
  uint64_t chunk_bytes = kDefaultChunk;
  tfw::MoverKind mover = tfw::kMoverLdg;
  int ctas_per_sm = tfw::kMoverMinCtas;

  std::vector<Buffer> bufs;  // indexed by handle

  // ---- current batch ----
  std::vector<tfw_move_desc> descs;
  IntervalSet wr, rd;
  bool chunk_open = false, chunk_inplace = false;
  const uint8_t* chunk_host = nullptr;  // in-place: first host byte of the chunk
  uint64_t chunk_len = 0;               // bytes of the slot used (in-place: host span)
  uint64_t chunk_align = 0;             // slot offset of the first byte (in-place: host_ptr & 15)
  // ---- parser state (payload in progress) ----
  bool in_payload = false, pay_valid = false;
  tfcs_frame_hdr pay_hdr{};
  uint64_t pay_done = 0, pay_pad = 0, pay_dst = 0;
  bool input_pinned = false;

  // ---- responses ----
  std::deque<Response> resp;
  std::vector<cudaEvent_t> ev_pool;
  uint8_t* arena = nullptr;
  uint64_t arena_size = 0, arena_used = 0;

  struct Fence { cudaEvent_t copy = nullptr, exec = nullptr; };
  std::vector<Fence> fences;      // ring of 8, indexed by ticket % 8
  uint64_t fence_next = 1;
  unsigned long long* d_digest = nullptr;
  tfw_gate* gate = nullptr;
  tfw_trace* rec = nullptr;  // non-null while tfw_trace_load is recording
  tfw_stats_record* pub = nullptr;  // mmap of <dir of shm_path>/tfw_stats (metrics channel to the provider)
  // ---- tiered address space (optional) ----
  tfw_vspace* vs = nullptr;
  uint64_t vs_base = 0, vs_R = 0;
  std::vector<uint8_t> vs_used;  // region allocation bitmap
  std::vector<uint32_t> batch_pins;  // regions named by the open batch: pinned until it is enqueued
  bool frozen = false, frozen_auto = false;
  uint64_t frozen_unix_ms = 0, auto_freezes = 0, auto_resumes = 0;
  uint64_t parked_bytes = 0, last_moved = 0;
  uint64_t ctl_seen = 0;  // last ctl_request handled
  // ---- client memory shared with the worker (HOST_REGISTER) ----
  std::string arena_prefix;  // arena k is the file <prefix>.a<k>; empty = the transport has no shared memory
  Arena arenas[TFCS_MAX_ARENAS + 1];
  // ---- response sink: responses are produced straight into the worker -> client ring ----
  uint8_t* sink_ring = nullptr;
  uint64_t sink_size = 0;
  uint64_t* sink_head = nullptr;        // published cursor (written here, release)
  const uint64_t* sink_tail = nullptr;  // consumer cursor (written by the client)
  uint64_t sink_wr = 0;                 // reserved cursor: [*sink_head, sink_wr) is written or being DMA'd
  std::deque<SinkPiece> sink_pending;
  bool d2h_active = false, d2h_hdr_done = false;  // a D2H that is waiting for ring space (resumable)
  tfcs_frame_hdr d2h_hdr{};
  uint64_t d2h_ptr = 0, d2h_done = 0;
  // ---- frames with a small payload that is assembled on the host (modules, user launches) ----
  bool in_blob = false, blob_valid = false;
  tfcs_frame_hdr blob_hdr{};
  std::vector<uint8_t> blob;
  uint64_t blob_skip = 0;  // payload + padding bytes still to pass over
  std::map<uint32_t, CUmodule> modules;
  std::map<uint32_t, UserFunction> functions;
  tfw_stats st{};
  std::string err = "";
};

namespace {

#define CU_OK(w, call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      (w)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                           \
      return TFW_ERR_FAILED;                                                                    \
    }                                                                                           \
  } while (0)

tfw_status fail(tfw_worker* w, tfw_status s, const char* msg) {
  w->err = msg;
  return s;
}

tfw_status touch_range(tfw_worker* w, uint64_t ptr, uint64_t len, bool unpin = true);  // tiered address space, defined below
void unpin_range(tfw_worker* w, uint64_t ptr, uint64_t len);
tfw_status do_blob_frame(tfw_worker* w, const tfcs_frame_hdr& h, std::vector<uint8_t>& blob);

void drop_arena(tfw_worker* w, uint32_t id) {
  Arena& a = w->arenas[id];
  if (!a.base) return;
  cudaHostUnregister(a.base);
  cudaGetLastError();
  munmap(a.base, a.size);
  a = Arena{};
}

cudaEvent_t get_event(tfw_worker* w) {
  if (!w->ev_pool.empty()) {
    cudaEvent_t e = w->ev_pool.back();
    w->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return e;
}

// ---- response sink -----------------------------------------------------------------------
// With a sink (tfw_set_response_sink) responses are produced directly in the worker -> client
// ring: headers by the CPU, D2H payloads by the copy engine (the ring is page-locked), and the
// ring's head cursor is advanced over a piece once the event behind it has completed.
uint64_t sink_room(const tfw_worker* w) {
  return w->sink_size - (w->sink_wr - __atomic_load_n(w->sink_tail, __ATOMIC_ACQUIRE));
}

void sink_put(tfw_worker* w, const void* src, uint64_t n) {  // caller checked sink_room
  const uint8_t* s = static_cast<const uint8_t*>(src);
  const uint64_t pos = w->sink_wr % w->sink_size, first = std::min(n, w->sink_size - pos);
  if (src) { std::memcpy(w->sink_ring + pos, s, first); if (n > first) std::memcpy(w->sink_ring, s + first, n - first); }
  else { std::memset(w->sink_ring + pos, 0, first); if (n > first) std::memset(w->sink_ring, 0, n - first); }
  w->sink_wr += n;
}

// Advance the published cursor over every leading piece whose event has completed.
void sink_publish(tfw_worker* w) {
  uint64_t upto = 0;
  while (!w->sink_pending.empty()) {
    SinkPiece& p = w->sink_pending.front();
    if (p.ev) {
      if (cudaEventQuery(p.ev) != cudaSuccess) { cudaGetLastError(); break; }
      w->ev_pool.push_back(p.ev);
    }
    upto = p.upto;
    w->sink_pending.pop_front();
  }
  if (upto) __atomic_store_n(w->sink_head, upto, __ATOMIC_RELEASE);
}

// Move queued small responses into the ring while there is room, then publish.
void sink_pump(tfw_worker* w) {
  while (!w->resp.empty()) {
    Response& r = w->resp.front();
    const uint64_t total = TFCS_HDR_BYTES + tfcs_pad16(r.len);
    if (sink_room(w) < total) break;
    sink_put(w, &r.hdr, TFCS_HDR_BYTES);
    if (r.len) { sink_put(w, r.host, r.len); sink_put(w, nullptr, tfcs_pad16(r.len) - r.len); }
    w->sink_pending.push_back({w->sink_wr, r.ev});
    if (r.owned) std::free(r.host);
    w->resp.pop_front();
  }
  sink_publish(w);
}

// Every response that is small (header, or header + a host blob) goes through here.
void push_response(tfw_worker* w, const Response& r) {
  w->resp.push_back(r);
  if (w->sink_ring) sink_pump(w);
}

void push_error(tfw_worker* w, const tfcs_frame_hdr& h, tfw_status code) {
  Response r{};
  r.hdr = h;
  r.hdr.opcode = TFCS_OP_RESP_ERROR;
  r.hdr.arg0 = (uint32_t)code;
  r.hdr.arg1 = h.opcode;
  r.hdr.length = 0;
  push_response(w, r);
}

Buffer* find(tfw_worker* w, uint32_t h) {
  if (h >= w->bufs.size() || !w->bufs[h].live) return nullptr;
  return &w->bufs[h];
}

uint32_t assign_tiles(tfw_move_desc* d, uint32_t n) {
  uint64_t t = 0;
  for (uint32_t i = 0; i < n; ++i) {
    d[i].tile0 = (uint32_t)t;
    t += tfw::mover_tiles(d[i].dst, d[i].len);
  }
  return (uint32_t)t;
}

// Publish the counters for the provider (seqlock: readers retry while seq is odd).
void publish_stats(tfw_worker* w) {
  tfw_stats_record* r = w->pub;
  if (!r) return;
  __atomic_store_n(&r->seq, r->seq + 1, __ATOMIC_RELEASE);
  r->updated_unix_secs = (uint64_t)time(nullptr);
  r->frames = w->st.frames; r->payload_bytes = w->st.payload_bytes; r->h2d_dma_bytes = w->st.h2d_dma_bytes;
  r->d2h_bytes = w->st.d2h_bytes; r->d2d_bytes = w->st.d2d_bytes; r->fill_bytes = w->st.fill_bytes;
  r->mover_launches = w->st.mover_launches; r->client_launches = w->st.client_launches; r->gate_launches = w->st.gate_launches;
  r->vram_bytes = w->st.vram_bytes; r->vram_peak_bytes = w->st.vram_peak_bytes; r->live_buffers = w->st.live_buffers;
  if (w->gate) {
    tfw_gate_state g{};
    if (tfw_gate_get_state(w->gate, &g) == TFW_OK) { r->gate_admitted = g.admitted; r->gate_blocked = g.blocked_gates; r->gate_timeouts = g.timeouts; }
  }
  r->sm_limit_percent = w->sm_percent;
  r->sm_count = (uint64_t)w->sm_count;
  r->vram_limit_bytes = w->cfg.vram_limit_bytes;
  r->ctl_frozen = w->frozen ? 1 : 0;
  r->frozen_unix_ms = w->frozen ? w->frozen_unix_ms : 0;
  r->frozen_auto = w->frozen && w->frozen_auto ? 1 : 0;
  r->auto_freezes = w->auto_freezes;
  r->auto_resumes = w->auto_resumes;
  r->ctl_moved_bytes = w->last_moved;
  r->parked_bytes = w->parked_bytes;
  if (w->frozen && !w->vs) r->vram_bytes = 0;  // plain buffers are in host memory now
  __atomic_store_n(&r->seq, r->seq + 1, __ATOMIC_RELEASE);
}

// Issue the current batch: DMA of the open chunk + mover launch (or record it).
void release_batch_pins(tfw_worker* w) {
  for (uint32_t r : w->batch_pins) tfw_vspace_pin(w->vs, r, 0);
  w->batch_pins.clear();
}

tfw_status flush_batch(tfw_worker* w) {
  if (w->descs.empty()) {
    w->chunk_open = false;
    w->chunk_len = 0;
    if (!w->batch_pins.empty()) release_batch_pins(w);
    return TFW_OK;
  }
  const uint32_t n = (uint32_t)w->descs.size();
  if (w->rec) {  // recording for resident replay
    tfw_trace* t = w->rec;
    Step s{};
    s.kind = kStepMover;
    s.desc_off = t->descs.size();
    s.ndesc = n;
    s.tiles = assign_tiles(w->descs.data(), n);
    s.bulk_ok = tfw::mover_bulk_ok(w->descs.data(), n);
    t->descs.insert(t->descs.end(), w->descs.begin(), w->descs.end());
    t->steps.push_back(s);
    t->mover_launches++;
    t->staged_in_batch = 0;
  } else {
    Slot& s = w->slots[w->cur];
    const uint32_t tiles = assign_tiles(w->descs.data(), n);
    const bool inline_descs = n <= kInlineDescs;
    cudaStream_t up = w->chunk_open ? w->copy_stream : w->exec_stream;
    if (!inline_descs) {
      std::memcpy(s.h_descs, w->descs.data(), sizeof(tfw_move_desc) * n);
      CU_OK(w, cudaMemcpyAsync(s.d_descs, s.h_descs, sizeof(tfw_move_desc) * n, cudaMemcpyHostToDevice, up));
    }
    if (w->chunk_open) {
      const uint8_t* src = w->chunk_inplace ? w->chunk_host : s.host + w->chunk_align;
      CU_OK(w, cudaMemcpyAsync(s.dev + w->chunk_align, src, w->chunk_len, cudaMemcpyHostToDevice, w->copy_stream));
      CU_OK(w, cudaEventRecord(s.dma_done, w->copy_stream));
      CU_OK(w, cudaStreamWaitEvent(w->exec_stream, s.dma_done, 0));
      w->st.h2d_dma_bytes += w->chunk_len;
    }
    if (inline_descs)
      CU_OK(w, tfw::launch_mover_inline(w->descs.data(), n, tiles, w->sm_count, w->ctas_per_sm, w->exec_stream));
    else
      CU_OK(w, tfw::launch_mover(s.d_descs, n, tiles, w->sm_count, w->ctas_per_sm,
                                 w->mover == tfw::kMoverTma && tfw::mover_bulk_ok(w->descs.data(), n) ? tfw::kMoverTma : tfw::kMoverLdg, w->exec_stream));
    CU_OK(w, cudaEventRecord(s.exec_done, w->exec_stream));
    s.busy = true;
    w->st.mover_launches++;
    w->cur = (w->cur + 1) % (uint32_t)w->slots.size();
    Slot& nx = w->slots[w->cur];
    if (nx.busy) {  // back-pressure: the slot we are about to fill must have been drained
      CU_OK(w, cudaEventSynchronize(nx.exec_done));
      nx.busy = false;
    }
  }
  w->descs.clear();
  w->wr.clear();
  w->rd.clear();
  w->chunk_open = false;
  w->chunk_len = 0;
  if (!w->batch_pins.empty()) release_batch_pins(w);  // its kernels are enqueued now: the tiering engine sees them
  return TFW_OK;
}

// Append one mover descriptor, cutting the batch first if it would race with
// an earlier descriptor of the same batch (the kernel runs them concurrently).
tfw_status add_desc(tfw_worker* w, uint64_t dst, uint64_t src, uint64_t len, uint32_t fill, bool src_is_client) {
  if (len == 0) return TFW_OK;
  const bool hazard = w->wr.overlaps(dst, dst + len) || w->rd.overlaps(dst, dst + len) ||
                      (src_is_client && w->wr.overlaps(src, src + len));
  if (hazard || w->descs.size() >= kMaxDescsPerBatch) {
    if (hazard) w->st.batches_hazard++;
    // NB: the caller re-derives any slot address after a flush; see stage_piece.
    tfw_status s = flush_batch(w);
    if (s != TFW_OK) return s;
  }
  tfw_move_desc d{};
  d.dst = dst;
  d.src = src;
  d.len = len;
  d.fill = fill;
  w->descs.push_back(d);
  w->wr.add(dst, dst + len);
  if (src_is_client) w->rd.add(src, src + len);
  return TFW_OK;
}

// Stage `len` payload bytes found at host address `p` and schedule their
// scatter to device address `dst`.
tfw_status stage_piece(tfw_worker* w, const uint8_t* p, uint64_t len, uint64_t dst) {
  w->st.payload_bytes += len;
  if (w->rec) {
    tfw_trace* t = w->rec;
    while (len) {
      uint64_t room = kResidentBatchBytes > t->staged_in_batch ? kResidentBatchBytes - t->staged_in_batch : 0;
      if (room == 0) {
        tfw_status s = flush_batch(w);
        if (s != TFW_OK) return s;
        continue;
      }
      const uint64_t take = std::min(len, room);
      tfw_status s = add_desc(w, dst, t->dev_base + (uint64_t)(p - t->host_base), take, 0, false);
      if (s != TFW_OK) return s;
      t->staged_in_batch += take;
      t->payload_bytes += take;
      t->algo_bytes += 2 * take;
      p += take; dst += take; len -= take;
    }
    return TFW_OK;
  }
  while (len) {
    if (w->vs) {  // per staged piece (<= one chunk), so a payload larger than the HBM budget streams through
      tfw_status ts = touch_range(w, dst, std::min(len, w->chunk_bytes));
      if (ts != TFW_OK) return ts;
    }
    // hazard check up front: it may flush, which closes the chunk
    if (w->wr.overlaps(dst, dst + len) || w->rd.overlaps(dst, dst + len) || w->descs.size() >= kMaxDescsPerBatch) {
      w->st.batches_hazard++;
      tfw_status s = flush_batch(w);
      if (s != TFW_OK) return s;
    }
    Slot& s = w->slots[w->cur];
    if (w->chunk_open && w->chunk_inplace != w->input_pinned) {
      tfw_status st = flush_batch(w);
      if (st != TFW_OK) return st;
      continue;
    }
    uint64_t src_dev = 0, take = 0;
    if (w->input_pinned) {
      if (w->chunk_open) {
        const uint8_t* end = w->chunk_host + w->chunk_len;
        if (p < end || (uint64_t)(p - end) > 4096) {  // not (nearly) contiguous with the open chunk
          tfw_status st = flush_batch(w);
          if (st != TFW_OK) return st;
          continue;
        }
      } else {
        w->chunk_open = true;
        w->chunk_inplace = true;
        w->chunk_host = p;
        w->chunk_len = 0;
        w->chunk_align = (uint64_t)(reinterpret_cast<uintptr_t>(p) & 15u);
      }
      const uint64_t off = (uint64_t)(p - w->chunk_host);
      if (off >= w->chunk_bytes) {
        tfw_status st = flush_batch(w);
        if (st != TFW_OK) return st;
        continue;
      }
      take = std::min(len, w->chunk_bytes - off);
      src_dev = reinterpret_cast<uint64_t>(s.dev) + w->chunk_align + off;
      w->chunk_len = off + take;
    } else {
      if (!w->chunk_open) {
        w->chunk_open = true;
        w->chunk_inplace = false;
        w->chunk_len = 0;
        w->chunk_align = 0;
      }
      const uint64_t cursor = (w->chunk_len + 15u) & ~(uint64_t)15u;
      if (cursor >= w->chunk_bytes) {
        tfw_status st = flush_batch(w);
        if (st != TFW_OK) return st;
        continue;
      }
      take = std::min(len, w->chunk_bytes - cursor);
      std::memcpy(s.host + cursor, p, take);
      src_dev = reinterpret_cast<uint64_t>(s.dev) + cursor;
      w->chunk_len = cursor + take;
    }
    tfw_move_desc d{};
    d.dst = dst; d.src = src_dev; d.len = take;
    w->descs.push_back(d);
    w->wr.add(dst, dst + take);
    p += take; dst += take; len -= take;
  }
  return TFW_OK;
}

tfw_status ensure_arena(tfw_worker* w, uint64_t need) {
  if (w->arena_used + need <= w->arena_size) return TFW_OK;
  if (!w->resp.empty() && w->arena_used) return TFW_ERR_EXHAUSTED;  // caller must poll first
  const uint64_t want = std::max<uint64_t>(need, 64ull << 20);
  if (want > w->arena_size) {
    if (w->arena) cudaFreeHost(w->arena);
    w->arena = nullptr;
    w->arena_size = 0;
    CU_OK(w, cudaHostAlloc(reinterpret_cast<void**>(&w->arena), want, cudaHostAllocDefault));
    w->arena_size = want;
  }
  w->arena_used = 0;
  return TFW_OK;
}

// Tokens a launch costs: blocks x warps per block, the unit the LD_PRELOAD limiter charges too
// (cuda_hook.cc).  Computed HERE from the launch geometry: the figure a client puts on the wire is
// only a lower bound, so a tenant cannot talk its way past the limiter (the gate clamps the
// cost to the bucket capacity, so an over-sized launch costs one full bucket and never starves).
double launch_cost(uint64_t blocks, uint64_t threads_per_block, uint64_t client_says) {
  const uint64_t own = std::max<uint64_t>(1, blocks) * std::max<uint64_t>(1, (threads_per_block + 31) / 32);
  return (double)std::max(own, client_says);
}

tfw_status charge_launch(tfw_worker* w, double cost) {
  if (!w->gate || (w->cfg.flags & TFW_F_NO_LIMITER)) return TFW_OK;
  tfw_status s = tfw_gate_enqueue(w->gate, cost, w->exec_stream);
  if (s != TFW_OK) return fail(w, s, "gate enqueue failed");
  w->st.gate_launches++;
  return TFW_OK;
}

tfw_status issue_launch(tfw_worker* w, const tfcs_frame_hdr& h, uint64_t ptr) {
  uint32_t grid = h.arg1, block = h.arg2;
  tfw::clamp_client_launch(h.arg0, h.length, &grid, &block);
  tfw_status gs = charge_launch(w, launch_cost(grid, block, h.arg3));
  if (gs != TFW_OK) return gs;
  CU_OK(w, tfw::launch_client_kernel(h.arg0, h.arg1, h.arg2, reinterpret_cast<uint8_t*>(ptr), h.length, h.off1,
                                     w->exec_stream));
  w->st.client_launches++;
  return TFW_OK;
}

// D2H into the sink: header by the CPU, payload by the copy engine in pieces of 4 MiB, each
// published on its own event so that the client copies piece k out while piece k+1 is in flight.
// Resumable: when the ring is full it answers TFW_ERR_EXHAUSTED and continues at the next call.
constexpr uint64_t kSinkPiece = 4ull << 20;
tfw_status continue_d2h_sink(tfw_worker* w) {
  const tfcs_frame_hdr& h = w->d2h_hdr;
  if (!w->d2h_hdr_done) {
    if (!w->resp.empty()) { sink_pump(w); if (!w->resp.empty()) return TFW_ERR_EXHAUSTED; }  // keep the order
    if (sink_room(w) < TFCS_HDR_BYTES) { sink_publish(w); return TFW_ERR_EXHAUSTED; }
    tfcs_frame_hdr r = h;
    r.opcode = TFCS_OP_RESP_D2H;
    sink_put(w, &r, TFCS_HDR_BYTES);
    w->sink_pending.push_back({w->sink_wr, nullptr});
    w->d2h_hdr_done = true;
    w->st.d2h_bytes += h.length;
  }
  while (w->d2h_done < h.length) {
    const uint64_t pos = w->sink_wr % w->sink_size;
    const uint64_t left = h.length - w->d2h_done;
    uint64_t k = std::min<uint64_t>(std::min<uint64_t>(left, kSinkPiece), w->sink_size - pos);
    const uint64_t room = sink_room(w);
    if (room < std::min<uint64_t>(k, 1u << 20) + 16) { sink_publish(w); return TFW_ERR_EXHAUSTED; }
    if (k + 16 > room) k = (room - 16) & ~(uint64_t)15;
    CU_OK(w, cudaMemcpyAsync(w->sink_ring + pos, reinterpret_cast<const void*>(w->d2h_ptr + w->d2h_done), k, cudaMemcpyDeviceToHost, w->exec_stream));
    w->sink_wr += k;
    w->d2h_done += k;
    if (w->d2h_done == h.length) sink_put(w, nullptr, tfcs_pad16(h.length) - h.length);  // padding rides on the last piece
    cudaEvent_t ev = get_event(w);
    CU_OK(w, cudaEventRecord(ev, w->exec_stream));
    w->sink_pending.push_back({w->sink_wr, ev});
  }
  w->d2h_active = false;
  sink_publish(w);
  return TFW_OK;
}

tfw_status issue_d2h(tfw_worker* w, const tfcs_frame_hdr& h, uint64_t ptr) {
  if (w->sink_ring) {
    w->d2h_active = true;
    w->d2h_hdr_done = false;
    w->d2h_hdr = h;
    w->d2h_ptr = ptr;
    w->d2h_done = 0;
    return continue_d2h_sink(w);
  }
  const uint64_t padded = tfcs_pad16(h.length);
  tfw_status s = ensure_arena(w, padded);
  if (s != TFW_OK) return s;
  Response r{};
  r.hdr = h;
  r.hdr.opcode = TFCS_OP_RESP_D2H;
  r.host = w->arena + w->arena_used;
  r.len = h.length;
  w->arena_used += padded;
  if (h.length) CU_OK(w, cudaMemcpyAsync(r.host, reinterpret_cast<void*>(ptr), h.length, cudaMemcpyDeviceToHost, w->exec_stream));
  r.ev = get_event(w);
  CU_OK(w, cudaEventRecord(r.ev, w->exec_stream));
  w->resp.push_back(r);
  w->st.d2h_bytes += h.length;
  return TFW_OK;
}

// header-only response that becomes valid when the exec stream reaches this point
tfw_status issue_marker(tfw_worker* w, const tfcs_frame_hdr& h, uint16_t opcode) {
  Response r{};
  r.hdr = h;
  r.hdr.opcode = opcode;
  r.hdr.arg0 = 0;
  r.hdr.length = 0;
  r.ev = get_event(w);
  CU_OK(w, cudaEventRecord(r.ev, w->exec_stream));
  push_response(w, r);
  return TFW_OK;
}

tfw_status issue_sync(tfw_worker* w, const tfcs_frame_hdr& h) { return issue_marker(w, h, TFCS_OP_RESP_SYNC); }

// ---- tiered address space --------------------------------------------------------------
// Make the regions under [ptr, ptr+len) usable by the next kernels: HOME regions get an LRU
// bump, PEER regions are used in place over NVLink, HOST regions are prefetched (which may evict
// colder regions).  A migration re-maps memory, so the vGPU stream is drained first.
tfw_status touch_range(tfw_worker* w, uint64_t ptr, uint64_t len, bool unpin) {
  if (!w->vs || !len || ptr < w->vs_base || ptr >= w->vs_base + (uint64_t)w->vs_used.size() * w->vs_R) return TFW_OK;
  const uint32_t r0 = (uint32_t)((ptr - w->vs_base) / w->vs_R), r1 = (uint32_t)((ptr + len - 1 - w->vs_base) / w->vs_R);
  tfw_status rc = TFW_OK;
  for (uint32_t r = r0; r <= r1; ++r) tfw_vspace_pin(w->vs, r, 1);  // the op's own regions may not evict each other
  for (uint32_t r = r0; r <= r1 && rc == TFW_OK; ++r) {
    uint32_t tier = 0;
    tfw_vspace_residency(w->vs, r, &tier, nullptr);
    if (tier == TFW_TIER_PEER) continue;
    // The tiering engine orders its copies against the kernels ENQUEUED on the exec stream (it is bound to it), so
    // the stream is never drained for a migration.  What it cannot see is the open batch: its descriptors are not
    // enqueued yet.  Regions they name stay pinned until the batch is flushed (batch_pins), and a miss -- which may
    // evict -- enqueues the open batch first.
    if (tier != TFW_TIER_HOME && !w->descs.empty()) {
      rc = flush_batch(w);
      if (rc != TFW_OK) break;
    }
    rc = tfw_vspace_access(w->vs, r);
    if (rc == TFW_ERR_EXHAUSTED && !w->batch_pins.empty()) {  // the open batch held the home budget: let it go and try again
      rc = flush_batch(w);
      if (rc == TFW_OK) rc = tfw_vspace_access(w->vs, r);
    }
    if (rc != TFW_OK) w->err = std::string("tiering: ") + tfw_vspace_last_error(w->vs);
    else if (!w->rec) { tfw_vspace_pin(w->vs, r, 1); w->batch_pins.push_back(r); }
  }
  if (unpin || rc != TFW_OK) for (uint32_t r = r0; r <= r1; ++r) tfw_vspace_pin(w->vs, r, 0);
  return rc;
}

void unpin_range(tfw_worker* w, uint64_t ptr, uint64_t len) {
  if (!w->vs || !len) return;
  const uint32_t r0 = (uint32_t)((ptr - w->vs_base) / w->vs_R), r1 = (uint32_t)((ptr + len - 1 - w->vs_base) / w->vs_R);
  for (uint32_t r = r0; r <= r1; ++r) tfw_vspace_pin(w->vs, r, 0);
}

tfw_status tiered_malloc(tfw_worker* w, uint64_t size, Buffer* out) {
  const uint32_t need = (uint32_t)((size + w->vs_R - 1) / w->vs_R);
  const uint32_t n = (uint32_t)w->vs_used.size();
  uint32_t run = 0, start = 0;
  bool found = false;
  for (uint32_t i = 0; i < n; ++i) {
    run = w->vs_used[i] ? 0 : run + 1;
    if (run == need) { start = i + 1 - need; found = true; break; }
  }
  if (!found) return TFW_ERR_EXHAUSTED;
  for (uint32_t i = start; i < start + need; ++i) w->vs_used[i] = 1;
  out->ptr = w->vs_base + (uint64_t)start * w->vs_R;
  out->size = size;
  out->live = true;
  out->tiered = true;
  out->region0 = start;
  out->nregions = need;
  // Regions are populated (zero-filled) on first touch; a fresh buffer must read as zeros even
  // if it is never written, so touch it now -- in pieces, so buffers larger than the HBM budget work.
  tfw_status rc = TFW_OK;
  for (uint32_t i = 0; i < need && rc == TFW_OK; ++i) rc = touch_range(w, out->ptr + (uint64_t)i * w->vs_R, 1);
  if (rc != TFW_OK) {
    for (uint32_t i = start; i < start + need; ++i) { tfw_vspace_unpopulate(w->vs, i); w->vs_used[i] = 0; }
    *out = Buffer{};
  }
  return rc;
}

// Execute (or record) one non-payload frame.
tfw_status do_frame(tfw_worker* w, const tfcs_frame_hdr& h) {
  switch (h.opcode) {
    case TFCS_OP_NOP: return TFW_OK;
    case TFCS_OP_MALLOC: {
      if (h.h0 >= TFCS_MAX_HANDLES || h.length == 0 || h.length > TFCS_MAX_BUFFER_BYTES) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      if (h.h0 < w->bufs.size() && w->bufs[h.h0].live) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      if (w->cfg.vram_limit_bytes && w->st.vram_bytes + h.length > w->cfg.vram_limit_bytes) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
      if (w->vs) {
        if (w->rec) return fail(w, TFW_ERR_NOT_SUPPORTED, "resident traces are not supported on a tiered worker");
        Buffer nb;
        tfw_status ts = tiered_malloc(w, h.length, &nb);
        if (ts == TFW_ERR_EXHAUSTED) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
        if (ts != TFW_OK) return ts;
        if (h.h0 >= w->bufs.size()) w->bufs.resize(h.h0 + 1);
        w->bufs[h.h0] = nb;
        w->st.vram_bytes += h.length;
        w->st.vram_peak_bytes = std::max(w->st.vram_peak_bytes, w->st.vram_bytes);
        w->st.live_buffers++;
        return TFW_OK;
      }
      void* p = nullptr;
      cudaError_t e = w->rec ? cudaMalloc(&p, h.length) : cudaMallocAsync(&p, h.length, w->exec_stream);
      if (e != cudaSuccess) { cudaGetLastError(); push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
      if (w->rec) w->rec->allocs.push_back(reinterpret_cast<uint64_t>(p));
      if (h.h0 >= w->bufs.size()) w->bufs.resize(h.h0 + 1);
      w->bufs[h.h0] = Buffer{reinterpret_cast<uint64_t>(p), h.length, true};
      w->st.vram_bytes += h.length;
      w->st.vram_peak_bytes = std::max(w->st.vram_peak_bytes, w->st.vram_bytes);
      w->st.live_buffers++;
      if (!(w->cfg.flags & TFW_F_NO_ZERO_FILL)) {
        w->st.fill_bytes += h.length;
        if (w->rec) w->rec->algo_bytes += h.length;
        return add_desc(w, reinterpret_cast<uint64_t>(p), 0, h.length, 0, false);
      }
      return TFW_OK;
    }
    case TFCS_OP_FREE: {
      Buffer* b = find(w, h.h0);
      if (!b) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      tfw_status s = flush_batch(w);
      if (s != TFW_OK) return s;
      if (b->tiered) {
        CU_OK(w, cudaStreamSynchronize(w->exec_stream));  // no kernel may still use the regions we unmap
        for (uint32_t i = b->region0; i < b->region0 + b->nregions; ++i) { tfw_vspace_unpopulate(w->vs, i); w->vs_used[i] = 0; }
      } else if (!w->rec) CU_OK(w, cudaFreeAsync(reinterpret_cast<void*>(b->ptr), w->exec_stream));
      w->st.vram_bytes -= b->size;
      w->st.live_buffers--;
      *b = Buffer{};
      return TFW_OK;
    }
    case TFCS_OP_MEMCPY_D2D: {
      Buffer* d = find(w, h.h0);
      Buffer* s = find(w, h.h1);
      if (!d || !s) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      if (h.off0 > d->size || h.length > d->size - h.off0 || h.off1 > s->size || h.length > s->size - h.off1) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      const uint64_t da = d->ptr + h.off0, sa = s->ptr + h.off1;
      if (h.length && da < sa + h.length && sa < da + h.length) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }  // overlapping D2D is undefined in CUDA
      w->st.d2d_bytes += h.length;
      if (w->rec) w->rec->algo_bytes += 2 * h.length;
      if (w->vs && h.length) {  // both ranges must be resident together: keep the first pinned while touching the second
        tfw_status ts = touch_range(w, sa, h.length, false);
        if (ts == TFW_OK) {
          ts = touch_range(w, da, h.length, false);
          unpin_range(w, sa, h.length);
          if (ts == TFW_OK) unpin_range(w, da, h.length);
        }
        if (ts == TFW_ERR_EXHAUSTED) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
        if (ts != TFW_OK) return ts;
      }
      return add_desc(w, da, sa, h.length, 0, true);
    }
    case TFCS_OP_MEMSET: {
      Buffer* d = find(w, h.h0);
      if (!d) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      if (h.off0 > d->size || h.length > d->size - h.off0) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      w->st.fill_bytes += h.length;
      if (w->rec) w->rec->algo_bytes += h.length;
      if (w->vs) {
        tfw_status ts = touch_range(w, d->ptr + h.off0, h.length);
        if (ts == TFW_ERR_EXHAUSTED) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
        if (ts != TFW_OK) return ts;
      }
      return add_desc(w, d->ptr + h.off0, 0, h.length, (h.arg0 & 0xffu) * 0x01010101u, false);
    }
    case TFCS_OP_MEMCPY_D2H: {
      Buffer* b = find(w, h.h0);
      if (!b) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      if (h.off0 > b->size || h.length > b->size - h.off0) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      tfw_status s = flush_batch(w);
      if (s != TFW_OK) return s;
      if (w->vs) {
        s = touch_range(w, b->ptr + h.off0, h.length);
        if (s == TFW_ERR_EXHAUSTED) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
        if (s != TFW_OK) return s;
      }
      if (w->rec) { Step st{}; st.kind = kStepD2H; st.hdr = h; st.ptr = b->ptr + h.off0; w->rec->steps.push_back(st); return TFW_OK; }
      return issue_d2h(w, h, b->ptr + h.off0);
    }
    case TFCS_OP_LAUNCH: {
      uint64_t ptr = 0;
      if (h.arg0 > TFCS_KERNEL_XOR_IDX) { push_error(w, h, TFW_ERR_NOT_SUPPORTED); return TFW_OK; }
      if (h.length) {
        Buffer* b = find(w, h.h0);
        if (!b) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
        if (h.off0 > b->size || h.length > b->size - h.off0) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
        ptr = b->ptr + h.off0;
      }
      tfw_status s = flush_batch(w);
      if (s != TFW_OK) return s;
      if (w->vs && h.length) {
        s = touch_range(w, ptr, h.length);
        if (s == TFW_ERR_EXHAUSTED) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
        if (s != TFW_OK) return s;
      }
      if (w->rec) { Step st{}; st.kind = kStepLaunch; st.hdr = h; st.ptr = ptr; w->rec->steps.push_back(st); return TFW_OK; }
      return issue_launch(w, h, ptr);
    }
    case TFCS_OP_SYNC: {
      tfw_status s = flush_batch(w);
      if (s != TFW_OK) return s;
      if (w->rec) { Step st{}; st.kind = kStepSync; st.hdr = h; w->rec->steps.push_back(st); return TFW_OK; }
      return issue_sync(w, h);
    }
    case TFCS_OP_HOST_REGISTER: {
      if (w->rec || w->arena_prefix.empty()) { push_error(w, h, TFW_ERR_NOT_SUPPORTED); return TFW_OK; }  // no shared memory on this transport
      if (h.h0 == 0 || h.h0 > TFCS_MAX_ARENAS || h.length == 0 || w->arenas[h.h0].base) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      // the name is derived, never taken from the wire: <ring file>.a<id>
      const std::string path = w->arena_prefix + ".a" + std::to_string(h.h0);
      const int fd = ::open(path.c_str(), O_RDWR | O_NOFOLLOW);
      struct stat sb{};
      if (fd < 0 || fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || (uint64_t)sb.st_size < h.length) {
        if (fd >= 0) ::close(fd);
        push_error(w, h, TFW_ERR_NOT_FOUND);
        return TFW_OK;
      }
      void* m = mmap(nullptr, h.length, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
      ::close(fd);
      if (m == MAP_FAILED) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
      if (cudaHostRegister(m, h.length, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(m, h.length);
        push_error(w, h, TFW_ERR_EXHAUSTED);
        return TFW_OK;
      }
      w->arenas[h.h0].base = static_cast<uint8_t*>(m);
      w->arenas[h.h0].size = h.length;
      return TFW_OK;
    }
    case TFCS_OP_HOST_UNREGISTER: {
      if (h.h0 == 0 || h.h0 > TFCS_MAX_ARENAS || !w->arenas[h.h0].base) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      tfw_status s = flush_batch(w);
      if (s != TFW_OK) return s;
      CU_OK(w, cudaStreamSynchronize(w->copy_stream));
      CU_OK(w, cudaStreamSynchronize(w->exec_stream));  // no DMA may still touch the pages we are about to unmap
      drop_arena(w, h.h0);
      return TFW_OK;
    }
    case TFCS_OP_MEMCPY_H2D_REF:
    case TFCS_OP_MEMCPY_D2H_REF: {
      const bool up = h.opcode == TFCS_OP_MEMCPY_H2D_REF;
      if (w->rec) return fail(w, TFW_ERR_NOT_SUPPORTED, "by-reference copies cannot be part of a resident trace");
      Buffer* b = find(w, h.h0);
      if (!b || h.h1 == 0 || h.h1 > TFCS_MAX_ARENAS || !w->arenas[h.h1].base) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      const Arena& a = w->arenas[h.h1];
      if (h.off0 > b->size || h.length > b->size - h.off0 || h.off1 > a.size || h.length > a.size - h.off1) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      tfw_status s = flush_batch(w);  // everything issued before is enqueued: stream order does the rest
      if (s != TFW_OK) return s;
      // The copy engine moves the bytes between the client's own page-locked pages and the buffer: no
      // staging slot, no unpack kernel.  In pieces for a tiered buffer larger than the HBM budget.
      uint64_t done = 0;
      while (done < h.length) {
        uint64_t k = h.length - done;
        if (w->vs && b->tiered) {
          k = std::min(k, w->chunk_bytes);
          s = touch_range(w, b->ptr + h.off0 + done, k);
          if (s == TFW_ERR_EXHAUSTED) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
          if (s != TFW_OK) return s;
        }
        if (up) CU_OK(w, cudaMemcpyAsync(reinterpret_cast<void*>(b->ptr + h.off0 + done), a.base + h.off1 + done, k, cudaMemcpyHostToDevice, w->exec_stream));
        else CU_OK(w, cudaMemcpyAsync(a.base + h.off1 + done, reinterpret_cast<const void*>(b->ptr + h.off0 + done), k, cudaMemcpyDeviceToHost, w->exec_stream));
        done += k;
      }
      if (up) { w->st.payload_bytes += h.length; w->st.h2d_dma_bytes += h.length; w->st.h2d_ref_bytes += h.length; }
      else { w->st.d2h_bytes += h.length; w->st.d2h_ref_bytes += h.length; }
      if (!up && (h.flags & TFCS_F_ACK)) return issue_marker(w, h, TFCS_OP_RESP_ACK);
      return TFW_OK;
    }
    case TFCS_OP_MODULE_UNLOAD: {
      auto it = w->modules.find(h.h0);
      if (it == w->modules.end()) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      tfw_status s = flush_batch(w);
      if (s != TFW_OK) return s;
      CU_OK(w, cudaStreamSynchronize(w->exec_stream));  // its kernels may still be running
      g_mod.cuModuleUnload(it->second);
      for (auto f = w->functions.begin(); f != w->functions.end();) f = f->second.module == h.h0 ? w->functions.erase(f) : std::next(f);
      w->modules.erase(it);
      return TFW_OK;
    }
    default: push_error(w, h, TFW_ERR_NOT_SUPPORTED); return TFW_OK;
  }
}

// Frames whose payload was assembled on the host: MODULE_LOAD, MODULE_GET_FUNCTION, LAUNCH_USER.
tfw_status do_blob_frame(tfw_worker* w, const tfcs_frame_hdr& h, std::vector<uint8_t>& blob) {
  if (w->rec) return fail(w, TFW_ERR_NOT_SUPPORTED, "user modules cannot be part of a resident trace");
  if (!g_mod.load()) { push_error(w, h, TFW_ERR_NOT_SUPPORTED); return TFW_OK; }
  switch (h.opcode) {
    case TFCS_OP_MODULE_LOAD: {
      if (h.h0 == 0 || h.h0 > TFCS_MAX_MODULES || blob.empty() || w->modules.count(h.h0)) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      blob.push_back(0);  // PTX is a C string; harmless behind a cubin / fatbin
      CUmodule m = nullptr;
      const CUresult r = g_mod.cuModuleLoadData(&m, blob.data());
      if (r != CUDA_SUCCESS) {
        const char* msg = nullptr;
        g_mod.cuGetErrorString(r, &msg);
        w->err = std::string("cuModuleLoadData: ") + (msg ? msg : "?");
        push_error(w, h, r == CUDA_ERROR_OUT_OF_MEMORY ? TFW_ERR_EXHAUSTED : TFW_ERR_INVALID);
        return TFW_OK;
      }
      w->modules[h.h0] = m;
      return TFW_OK;
    }
    case TFCS_OP_MODULE_GET_FUNCTION: {
      auto it = w->modules.find(h.h0);
      if (it == w->modules.end() || blob.empty()) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      if (h.h1 >= TFCS_MAX_FUNCTIONS) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      const std::string name(blob.begin(), blob.end());
      UserFunction f;
      f.module = h.h0;
      if (g_mod.cuModuleGetFunction(&f.fn, it->second, name.c_str()) != CUDA_SUCCESS) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      for (size_t i = 0;; ++i) {
        size_t off = 0, size = 0;
        if (g_mod.cuFuncGetParamInfo(f.fn, i, &off, &size) != CUDA_SUCCESS) break;
        f.params.emplace_back((uint32_t)off, (uint32_t)size);
        f.param_bytes = std::max<uint32_t>(f.param_bytes, (uint32_t)(off + size));
      }
      if (f.param_bytes > TFCS_MAX_PARAM_BYTES) { push_error(w, h, TFW_ERR_NOT_SUPPORTED); return TFW_OK; }
      Response r{};
      r.hdr = h;
      r.hdr.opcode = TFCS_OP_RESP_FUNCTION;
      r.hdr.arg0 = (uint32_t)f.params.size();
      r.hdr.arg1 = f.param_bytes;
      r.len = r.hdr.length = 8 * f.params.size();
      if (r.len) {
        r.host = static_cast<uint8_t*>(std::malloc(r.len));
        if (!r.host) return fail(w, TFW_ERR_EXHAUSTED, "out of host memory");
        r.owned = true;
        for (size_t i = 0; i < f.params.size(); ++i) { std::memcpy(r.host + 8 * i, &f.params[i].first, 4); std::memcpy(r.host + 8 * i + 4, &f.params[i].second, 4); }
      }
      w->functions[h.h1] = f;
      push_response(w, r);
      return TFW_OK;
    }
    case TFCS_OP_LAUNCH_USER: {
      auto it = w->functions.find(h.h1);
      if (it == w->functions.end()) { push_error(w, h, TFW_ERR_NOT_FOUND); return TFW_OK; }
      const UserFunction& f = it->second;
      tfcs_launch_params lp;
      if (blob.size() < sizeof lp) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      std::memcpy(&lp, blob.data(), sizeof lp);
      const uint64_t blocks = (uint64_t)lp.grid[0] * lp.grid[1] * lp.grid[2], threads = (uint64_t)lp.block[0] * lp.block[1] * lp.block[2];
      if (!blocks || !threads || threads > 1024 || lp.param_bytes != blob.size() - sizeof lp || lp.param_bytes < f.param_bytes) { push_error(w, h, TFW_ERR_INVALID); return TFW_OK; }
      // client-side device pointers -> real addresses: every aligned 8-byte word that carries the stub's tag and
      // names a live buffer (kernels receive pointers inside by-value structs too, so all words are looked at)
      alignas(16) uint8_t params[TFCS_MAX_PARAM_BYTES];
      std::memcpy(params, blob.data() + sizeof lp, lp.param_bytes);
      std::vector<std::pair<uint64_t, uint64_t>> touched;
      for (uint32_t o = 0; o + 8 <= lp.param_bytes; o += 8) {
        uint64_t v;
        std::memcpy(&v, params + o, 8);
        if (!TFCS_PTR_IS_TAGGED(v)) continue;
        Buffer* b = find(w, TFCS_PTR_HANDLE(v));
        if (!b || TFCS_PTR_OFFSET(v) > b->size) continue;
        const uint64_t real = b->ptr + TFCS_PTR_OFFSET(v);
        std::memcpy(params + o, &real, 8);
        if (b->tiered) touched.emplace_back(b->ptr, b->size);
      }
      tfw_status s = flush_batch(w);
      if (s != TFW_OK) return s;
      for (auto& t : touched) {  // every buffer the kernel may dereference must be resident together
        s = touch_range(w, t.first, t.second, false);
        if (s != TFW_OK) break;
      }
      if (s != TFW_OK) {
        for (auto& t : touched) unpin_range(w, t.first, t.second);
        if (s == TFW_ERR_EXHAUSTED) { push_error(w, h, TFW_ERR_EXHAUSTED); return TFW_OK; }
        return s;
      }
      s = charge_launch(w, launch_cost(blocks, threads, h.arg3));
      if (s != TFW_OK) return s;
      void* args[512];
      if (f.params.size() > 512) { push_error(w, h, TFW_ERR_NOT_SUPPORTED); return TFW_OK; }
      for (size_t i = 0; i < f.params.size(); ++i) args[i] = params + f.params[i].first;
      if (lp.shared_bytes > (48u << 10)) g_mod.cuFuncSetAttribute(f.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)lp.shared_bytes);
      const CUresult r = g_mod.cuLaunchKernel(f.fn, lp.grid[0], lp.grid[1], lp.grid[2], lp.block[0], lp.block[1], lp.block[2], lp.shared_bytes,
                                              reinterpret_cast<CUstream>(w->exec_stream), args, nullptr);
      for (auto& t : touched) unpin_range(w, t.first, t.second);
      if (r != CUDA_SUCCESS) { push_error(w, h, TFW_ERR_FAILED); return TFW_OK; }
      w->st.client_launches++;
      w->st.user_launches++;
      return TFW_OK;
    }
    default: push_error(w, h, TFW_ERR_NOT_SUPPORTED); return TFW_OK;
  }
}

// The deserializer: walks frames in [p, p+n), resumable in the middle of a payload.
tfw_status parse(tfw_worker* w, const uint8_t* p, size_t n, size_t* consumed) {
  size_t pos = 0;
  tfw_status rc = TFW_OK;
  if (w->d2h_active) {  // a D2H that ran out of ring space: nothing after it may be enqueued before it
    rc = continue_d2h_sink(w);
    if (rc != TFW_OK) { if (consumed) *consumed = 0; return rc; }
    w->st.frames++;
  }
  while (pos < n) {
    if (w->in_blob) {  // payload assembled on the host (module image, kernel name, launch parameters)
      const uint64_t want = w->blob_hdr.length - std::min<uint64_t>(w->blob_hdr.length, w->blob.size());
      const uint64_t take = std::min<uint64_t>(w->blob_valid ? want : 0, n - pos);
      if (take) w->blob.insert(w->blob.end(), p + pos, p + pos + take);
      const uint64_t skip = std::min<uint64_t>(w->blob_skip, n - pos);
      pos += skip;
      w->blob_skip -= skip;
      if (w->blob_skip) break;  // need more bytes
      w->in_blob = false;
      if (w->blob_valid) {
        rc = do_blob_frame(w, w->blob_hdr, w->blob);
        std::vector<uint8_t>().swap(w->blob);
        if (rc != TFW_OK) break;
      }
      w->st.frames++;
      continue;
    }
    if (w->in_payload) {
      const uint64_t remaining = w->pay_hdr.length - w->pay_done;
      const uint64_t take = std::min<uint64_t>(remaining, n - pos);
      if (take) {
        if (w->pay_valid) {
          rc = stage_piece(w, p + pos, take, w->pay_dst + w->pay_done);
          if (rc != TFW_OK) break;
        }
        pos += take;
        w->pay_done += take;
      }
      if (w->pay_done == w->pay_hdr.length) {
        const uint64_t skip = std::min<uint64_t>(w->pay_pad, n - pos);
        pos += skip;
        w->pay_pad -= skip;
        if (w->pay_pad == 0) { w->in_payload = false; w->st.frames++; }
      }
      continue;
    }
    if (n - pos < TFCS_HDR_BYTES) break;
    tfcs_frame_hdr h;
    std::memcpy(&h, p + pos, sizeof(h));
    if (h.magic != TFCS_MAGIC || h.version != TFCS_VERSION) { rc = fail(w, TFW_ERR_PROTOCOL, "bad frame magic/version"); break; }
    if (h.opcode == TFCS_OP_MEMCPY_H2D) {
      pos += TFCS_HDR_BYTES;
      Buffer* b = find(w, h.h0);
      w->pay_valid = false;
      if (!b) push_error(w, h, TFW_ERR_NOT_FOUND);
      else if (h.off0 > b->size || h.length > b->size - h.off0) push_error(w, h, TFW_ERR_INVALID);
      else { w->pay_valid = true; w->pay_dst = b->ptr + h.off0; }
      w->pay_hdr = h;
      w->pay_done = 0;
      w->pay_pad = tfcs_pad16(h.length) - h.length;
      w->in_payload = true;
      if (h.length == 0 && w->pay_pad == 0) { w->in_payload = false; w->st.frames++; }
      continue;
    }
    if (tfcs_has_payload(h.opcode)) {  // MODULE_LOAD, MODULE_GET_FUNCTION, LAUNCH_USER (and frames a client must not send)
      pos += TFCS_HDR_BYTES;
      const uint64_t cap = h.opcode == TFCS_OP_MODULE_LOAD ? kMaxModuleBytes : h.opcode == TFCS_OP_LAUNCH_USER ? sizeof(tfcs_launch_params) + TFCS_MAX_PARAM_BYTES : 1024;
      w->blob_hdr = h;
      w->blob.clear();
      w->blob_valid = h.opcode >= TFCS_OP_MODULE_LOAD && h.opcode <= TFCS_OP_LAUNCH_USER && h.length <= cap;
      if (!w->blob_valid) push_error(w, h, h.opcode >= TFCS_OP_MODULE_LOAD && h.opcode <= TFCS_OP_LAUNCH_USER ? TFW_ERR_INVALID : TFW_ERR_NOT_SUPPORTED);
      else w->blob.reserve(h.length + 1);
      w->blob_skip = tfcs_pad16(h.length);
      w->in_blob = true;
      if (w->blob_skip == 0) {  // empty payload
        w->in_blob = false;
        if (w->blob_valid) { rc = do_blob_frame(w, h, w->blob); if (rc != TFW_OK) break; }
        w->st.frames++;
      }
      continue;
    }
    if (h.opcode == TFCS_OP_MEMCPY_D2H) {  // may need arena / ring space: check before consuming
      const size_t save = pos;
      pos += TFCS_HDR_BYTES;
      rc = do_frame(w, h);
      if (rc == TFW_ERR_EXHAUSTED) {
        if (!w->d2h_active) pos = save;  // (sink mode keeps the frame and resumes it: see the top of parse)
        break;
      }
      if (rc != TFW_OK) break;
      w->st.frames++;
      continue;
    }
    pos += TFCS_HDR_BYTES;
    rc = do_frame(w, h);
    if (rc != TFW_OK) break;
    w->st.frames++;
  }
  if (consumed) *consumed = pos;
  return rc;
}

// The vGPU's execution stream.  percent in 1..99: the stream belongs to a green context that owns
// ceil(percent % of the SMs) (rounded up to the part's partition granularity), so every kernel of the
// tenant -- client launches, user modules, the byte mover -- runs on that share of the GPU and nowhere
// else: TF_CUDA_SM_PERCENT_LIMIT / AccelSetComputeUnitHardLimit (internal/utils/compose.go:1287-1295,
// provider/accelerator.h:351-358) as hard isolation.  0 or >= 100: an ordinary stream on the whole GPU.
tfw_status make_exec_stream(tfw_worker* w, uint32_t percent, cudaStream_t* out_stream, CUgreenCtx* out_green, int* out_sms) {
  *out_green = nullptr;
  *out_sms = w->device_sms;
  if (percent == 0 || percent >= 100) {
    CU_OK(w, cudaStreamCreateWithFlags(out_stream, cudaStreamNonBlocking));
    return TFW_OK;
  }
  g_mod.load();
  if (!g_mod.green) return fail(w, TFW_ERR_NOT_SUPPORTED, "this driver has no green contexts: the SM limit cannot be enforced");
  CUdevice dev;
  CUdevResource all{}, part{}, rest{};
  unsigned groups = 1;
  const unsigned want = std::max(1u, (unsigned)((w->device_sms * (uint64_t)percent + 99) / 100));
  CUdevResourceDesc desc = nullptr;
  CUstream st = nullptr;
  CUgreenCtx g = nullptr;
  if (g_mod.cuDeviceGet(&dev, w->device) != CUDA_SUCCESS || g_mod.cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS ||
      g_mod.cuDevSmResourceSplitByCount(&part, &groups, &all, &rest, 0, want) != CUDA_SUCCESS || groups < 1 ||
      g_mod.cuDevResourceGenerateDesc(&desc, &part, 1) != CUDA_SUCCESS || g_mod.cuGreenCtxCreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS)
    return fail(w, TFW_ERR_FAILED, "cannot partition the SMs for the compute limit");
  if (g_mod.cuGreenCtxStreamCreate(&st, g, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
    g_mod.cuGreenCtxDestroy(g);
    return fail(w, TFW_ERR_FAILED, "cannot create the vGPU stream inside its SM partition");
  }
  *out_stream = reinterpret_cast<cudaStream_t>(st);
  *out_green = g;
  *out_sms = (int)part.sm.smCount;
  return TFW_OK;
}


// ---- parking plain buffers at PCIe speed ------------------------------------------------------
// A frozen plain vGPU keeps its bytes in ordinary (pageable) host memory: it may be as large as the
// GPU, and nobody can page-lock that much up front.  A cudaMemcpy into fresh pageable memory crawls
// (2.9 GB/s in round 1: one thread taking a page fault per 4 KiB behind the driver's own staging).
// Here the copy engine streams the buffer through a ring of page-locked bounce slots at PCIe speed
// while a pool of threads moves slot -> destination, so the first-touch page faults (on transparent
// huge pages where the kernel grants them) are taken by many cores at once; the way back is the mirror
// image.  The bounce ring lives only for the duration of one freeze / resume.
struct ParkPipe {
  uint64_t chunk = 4ull << 20;  // bytes per bounce slot (at most; a smaller staging slot shrinks it): 32 slots in the default 128 MiB of staging
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<uint8_t*> slot;           // page-locked bounce buffers
  std::vector<cudaEvent_t> dma;         // D2H: data has landed in the slot / H2D: the slot has been read
  std::vector<int> state;               // 0 free, 1 job queued or running
  std::vector<char> recorded;           // dma[s] has been recorded at least once (a later buffer may find a slot's DMA still in flight)
  size_t ring_pos = 0;                  // unpark: chunks are laid on the ring continuously, buffer after buffer
  struct Job { int slot; uint8_t* host; uint64_t n; bool to_host; };
  std::vector<Job> jobs;
  size_t next = 0;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::vector<std::thread> threads;
  bool stop = false, failed = false;
  double t_fill_wait = 0, t_ring_wait = 0, t_issue = 0;  // where unpark's issuing thread spent its time (TFW_LOG_PARK)

  // The bounce ring is carved out of the worker's own page-locked staging slots (idle while the vGPU is drained):
  // page-locking fresh memory costs more than moving a small vGPU.
  bool start(int dev, cudaStream_t st, const std::vector<std::pair<uint8_t*, uint64_t>>& pinned) {
    device = dev;
    stream = st;
    for (const auto& pb : pinned) {
      chunk = std::min<uint64_t>(chunk, pb.second & ~(uint64_t)4095);
    }
    if (chunk < (1u << 20)) return false;
    for (const auto& pb : pinned)
      for (uint64_t o = 0; o + chunk <= pb.second; o += chunk) slot.push_back(pb.first + o);
    if (slot.size() < 2) return false;
    const unsigned hw = std::thread::hardware_concurrency();
    unsigned want = std::min(16u, hw ? hw / 4 : 4u);  // first-touch page faults scale with cores up to ~16 (38 GB/s on the 128-thread box; 32 threads were slower)
    if (const char* e = getenv("TFW_PARK_THREADS")) { const int v = atoi(e); if (v > 0) want = (unsigned)v; }
    const unsigned nthreads = (unsigned)std::max<size_t>(2, std::min<size_t>(want, slot.size()));
    dma.assign(slot.size(), nullptr);
    state.assign(slot.size(), 0);
    recorded.assign(slot.size(), 0);
    for (auto& e : dma)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
    for (unsigned i = 0; i < nthreads; ++i) threads.emplace_back([this] { loop(); });
    return true;
  }
  void loop() {
    cudaSetDevice(device);
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv_job.wait(lk, [&] { return stop || next < jobs.size(); });
      if (next >= jobs.size()) { if (stop) return; continue; }
      const Job j = jobs[next++];
      lk.unlock();
      bool ok = true;
      if (j.to_host) {  // the DMA into the slot must have landed
        ok = cudaEventSynchronize(dma[j.slot]) == cudaSuccess;
        if (ok) std::memcpy(j.host, slot[j.slot], j.n);
      } else {
        std::memcpy(slot[j.slot], j.host, j.n);
      }
      lk.lock();
      if (!ok) failed = true;
      state[j.slot] = j.to_host ? 0 : 2;  // 2 = filled, waiting for its H2D DMA
      cv_done.notify_all();
    }
  }
  // device -> host
  bool park(uint64_t dev_ptr, uint8_t* host, uint64_t size) {
    uint64_t off = 0;
    size_t k = 0;
    while (off < size) {
      const int s = (int)(k++ % slot.size());
      const uint64_t n = std::min(chunk, size - off);
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return state[s] == 0; });
        state[s] = 1;
      }
      if (cudaMemcpyAsync(slot[s], reinterpret_cast<const void*>(dev_ptr + off), n, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
          cudaEventRecord(dma[s], stream) != cudaSuccess) return false;
      {
        std::lock_guard<std::mutex> lk(mu);
        jobs.push_back({s, host + off, n, true});
      }
      cv_job.notify_one();
      off += n;
    }
    return true;  // the ring keeps streaming into the next buffer; the caller drains once at the end
  }
  // host -> device: threads fill the slots (chunk k -> slot k % ring), the DMA follows in order
  bool unpark(uint64_t dev_ptr, const uint8_t* host, uint64_t size) {
    const size_t nchunks = (size_t)((size + chunk - 1) / chunk), ring = slot.size();
    size_t issued = 0, queued = 0;
    while (issued < nchunks) {
      while (queued < nchunks && queued - issued < ring) {
        const int s = (int)((ring_pos + queued) % ring);
        // the slot last carried the chunk one ring earlier (maybe of the previous buffer), whose DMA was issued: wait until it has been read
        if (recorded[s]) {
          const auto a = std::chrono::steady_clock::now();
          if (cudaEventSynchronize(dma[s]) != cudaSuccess) return false;
          t_ring_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
        }
        {
          std::lock_guard<std::mutex> lk(mu);
          state[s] = 1;
          jobs.push_back({s, const_cast<uint8_t*>(host) + queued * chunk, std::min<uint64_t>(chunk, size - queued * chunk), false});
        }
        cv_job.notify_one();
        ++queued;
      }
      const int s = (int)((ring_pos + issued) % ring);
      const auto a = std::chrono::steady_clock::now();
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return state[s] == 2 || failed; });
        if (failed) return false;
        state[s] = 0;
      }
      const auto b = std::chrono::steady_clock::now();
      const uint64_t n = std::min<uint64_t>(chunk, size - issued * chunk);
      if (cudaMemcpyAsync(reinterpret_cast<void*>(dev_ptr + issued * chunk), slot[s], n, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
          cudaEventRecord(dma[s], stream) != cudaSuccess) return false;
      recorded[s] = 1;
      t_fill_wait += std::chrono::duration<double>(b - a).count();
      t_issue += std::chrono::duration<double>(std::chrono::steady_clock::now() - b).count();
      ++issued;
    }
    ring_pos = (ring_pos + nchunks) % ring;
    return true;  // (the caller synchronises the stream once, after the last buffer)
  }
  bool drain() {
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { for (int st : state) if (st != 0) return false; return true; });
    return !failed;
  }
  ~ParkPipe() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv_job.notify_all();
    for (auto& t : threads) t.join();
    for (auto e : dma) if (e) cudaEventDestroy(e);
  }
};

std::vector<std::pair<uint8_t*, uint64_t>> staging_slots(tfw_worker* w);

// host memory for a parked buffer: anonymous pages, huge where the kernel grants them (fewer, cheaper first touches)
uint8_t* park_alloc(uint64_t n) {
  void* m = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (m == MAP_FAILED) return nullptr;
#ifdef MADV_HUGEPAGE
  madvise(m, n, MADV_HUGEPAGE);
#endif
  return static_cast<uint8_t*>(m);
}
void park_free(uint8_t* p, uint64_t n) {
  if (!p) return;
  // giving hundreds of thousands of touched pages back takes the kernel a while: not on the resume path
  if (n >= (64ull << 20)) std::thread([p, n] { munmap(p, n); }).detach();
  else munmap(p, n);
}

std::vector<std::pair<uint8_t*, uint64_t>> staging_slots(tfw_worker* w) {
  std::vector<std::pair<uint8_t*, uint64_t>> v;
  for (auto& sl : w->slots) if (sl.host) v.emplace_back(sl.host, w->chunk_bytes);
  return v;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

}  // namespace

// ===========================================================================
// C-ABI
// ===========================================================================
extern "C" {

uint32_t tfw_abi_version(void) { return 2; }

const char* tfw_last_error(const tfw_worker* w) { return w ? w->err.c_str() : "null worker"; }

tfw_status tfw_worker_create(const tfw_config* cfg, tfw_worker** out) {
  if (!cfg || !out || (cfg->struct_size != sizeof(tfw_config) && cfg->struct_size != TFW_CONFIG_SIZE_V1 && cfg->struct_size != TFW_CONFIG_SIZE_V2)) return TFW_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return TFW_ERR_NO_DEVICE; }
  if (cfg->device < 0 || cfg->device >= ndev) return TFW_ERR_INVALID;
  tfw_worker* w = new (std::nothrow) tfw_worker();
  if (!w) return TFW_ERR_EXHAUSTED;
  std::memcpy(&w->cfg, cfg, cfg->struct_size);  // a V1 caller leaves `tiering` NULL
  w->device = cfg->device;
  w->chunk_bytes = cfg->chunk_bytes ? (cfg->chunk_bytes + 4095) & ~4095ull : kDefaultChunk;
  const uint32_t nslots = cfg->num_slots ? std::max(2u, cfg->num_slots) : kDefaultSlots;
  auto bail = [&](tfw_status s) { tfw_worker_destroy(w); return s; };
#define CR(call) do { if ((call) != cudaSuccess) { cudaGetLastError(); return bail(TFW_ERR_FAILED); } } while (0)
  CR(cudaSetDevice(w->device));
  cudaDeviceProp prop{};
  CR(cudaGetDeviceProperties(&prop, w->device));
  if (prop.major < 10) return bail(TFW_ERR_NOT_SUPPORTED);  // sm_100a only, no fallback
  w->sm_count = w->device_sms = prop.multiProcessorCount;
  CR(tfw::preload_kernels());
  CR(tfw::preload_gate_kernels());
  w->mover = (cfg->flags & TFW_F_MOVER_TMA) ? tfw::kMoverTma : tfw::kMoverLdg;
  if (const char* mv = getenv("TFW_MOVER")) {  // pick the mover without touching the caller (profiling both through the same bench command)
    if (!strcmp(mv, "tma")) w->mover = tfw::kMoverTma;
    else if (!strcmp(mv, "ldg")) w->mover = tfw::kMoverLdg;
  }
  w->ctas_per_sm = cfg->mover_ctas_per_sm ? (int)cfg->mover_ctas_per_sm : 0;  // 0 = one tile per CTA (both movers); > 0: persistent grid
  CR(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
  {
    const uint32_t pct = cfg->struct_size >= sizeof(tfw_config) ? cfg->sm_percent_limit : 0;
    if (pct > 100) return bail(TFW_ERR_INVALID);
    tfw_status es = make_exec_stream(w, pct, &w->exec_stream, &w->green, &w->sm_count);
    if (es != TFW_OK) return bail(es);
    w->sm_percent = pct >= 100 ? 0 : pct;
  }
  cudaMemPool_t pool;
  CR(cudaDeviceGetDefaultMemPool(&pool, w->device));
  uint64_t thr = ~0ull;
  CR(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  w->slots.resize(nslots);
  for (auto& s : w->slots) {
    CR(cudaHostAlloc(reinterpret_cast<void**>(&s.host), w->chunk_bytes + kSlotSlack, cudaHostAllocDefault));
    CR(cudaMalloc(reinterpret_cast<void**>(&s.dev), w->chunk_bytes + kSlotSlack));
    CR(cudaHostAlloc(reinterpret_cast<void**>(&s.h_descs), sizeof(tfw_move_desc) * kMaxDescsPerBatch, cudaHostAllocDefault));
    CR(cudaMalloc(reinterpret_cast<void**>(&s.d_descs), sizeof(tfw_move_desc) * kMaxDescsPerBatch));
    CR(cudaEventCreateWithFlags(&s.dma_done, cudaEventDisableTiming));
    CR(cudaEventCreateWithFlags(&s.exec_done, cudaEventDisableTiming));
  }
  CR(cudaMalloc(reinterpret_cast<void**>(&w->d_digest), sizeof(unsigned long long)));
#undef CR
  if (cfg->shm_path && !(cfg->flags & TFW_F_NO_LIMITER)) {
    tfw_status gs = tfw_gate_create(w->device, cfg->shm_path, cfg->shm_device_index, &w->gate);
    if (gs != TFW_OK) return bail(gs);
    if (cfg->flags & TFW_F_GATE_FAIL_CLOSED) tfw_gate_set_policy(w->gate, 1, 0);
  }
  if (w->cfg.tiering) {
    tfw_vspace_config vc;
    std::memcpy(&vc, w->cfg.tiering, sizeof vc);
    if (vc.struct_size != sizeof vc) return bail(TFW_ERR_INVALID);
    vc.home_device = w->device;
    vc.flags |= TFW_VS_PEER_IN_PLACE;  // touch_range uses peer-resident regions where they are
    tfw_status ts = tfw_vspace_create(&vc, &w->vs);
    if (ts != TFW_OK) return bail(ts);
    uint32_t nreg = 0;
    tfw_vspace_info(w->vs, &w->vs_base, &w->vs_R, &nreg);
    tfw_vspace_bind_stream(w->vs, w->exec_stream);  // migrations are ordered against the vGPU's kernels on the GPU, not by draining
    w->vs_used.assign(nreg, 0);
  }
  {  // metrics channel: next to the quota file, or wherever TFW_STATS_PATH says
    std::string path;
    if (const char* e = getenv("TFW_STATS_PATH")) path = e;
    else if (cfg->shm_path) { path = cfg->shm_path; const size_t k = path.find_last_of('/'); path = (k == std::string::npos ? std::string(".") : path.substr(0, k)) + "/" + TFW_STATS_FILE_NAME; }
    if (!path.empty()) {
      int fd = ::open(path.c_str(), O_RDWR | O_CREAT, 0644);
      if (fd >= 0 && ftruncate(fd, sizeof(tfw_stats_record)) == 0) {
        void* m = mmap(nullptr, sizeof(tfw_stats_record), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m != MAP_FAILED) {
          w->pub = static_cast<tfw_stats_record*>(m);
          std::memset(w->pub, 0, sizeof(*w->pub));
          w->pub->magic = TFW_STATS_MAGIC;
          w->pub->version = TFW_STATS_VERSION;
          w->pub->pid = (uint64_t)getpid();
          {  // who this worker is to the hypervisor (FreezeWorker & co. name workers, not processes)
            std::string id;
            const char* ns = getenv("POD_NAMESPACE");
            const char* pod = getenv("POD_NAME");
            if (const char* e = getenv("TF_WORKER_ID")) id = e;
            else if (const char* u = getenv("POD_UID")) id = u;
            else if (ns && pod) id = std::string(ns) + "/" + pod;
            snprintf(w->pub->worker_id, sizeof(w->pub->worker_id), "%s", id.c_str());
          }
          const unsigned char* u = reinterpret_cast<const unsigned char*>(prop.uuid.bytes);
          snprintf(w->pub->device_uuid, sizeof(w->pub->device_uuid),
                   "GPU-%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", u[0], u[1], u[2], u[3], u[4], u[5], u[6],
                   u[7], u[8], u[9], u[10], u[11], u[12], u[13], u[14], u[15]);
          publish_stats(w);
        }
      }
      if (fd >= 0) ::close(fd);
    }
  }
  *out = w;
  return TFW_OK;
}

tfw_status tfw_worker_destroy(tfw_worker* w) {
  if (!w) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  if (w->exec_stream) cudaStreamSynchronize(w->exec_stream);
  if (w->copy_stream) cudaStreamSynchronize(w->copy_stream);
  if (w->pub) { publish_stats(w); munmap(w->pub, sizeof(tfw_stats_record)); }
  if (w->gate) tfw_gate_destroy(w->gate);
  for (auto& b : w->bufs) {
    if (b.live && !b.tiered && b.ptr) cudaFree(reinterpret_cast<void*>(b.ptr));
    if (b.parked) park_free(b.parked, b.size);
  }
  if (w->vs) tfw_vspace_destroy(w->vs);
  for (auto& s : w->slots) {
    if (s.host) cudaFreeHost(s.host);
    if (s.dev) cudaFree(s.dev);
    if (s.h_descs) cudaFreeHost(s.h_descs);
    if (s.d_descs) cudaFree(s.d_descs);
    if (s.dma_done) cudaEventDestroy(s.dma_done);
    if (s.exec_done) cudaEventDestroy(s.exec_done);
  }
  for (auto& f : w->fences) { if (f.copy) cudaEventDestroy(f.copy); if (f.exec) cudaEventDestroy(f.exec); }
  for (auto& r : w->resp) { if (r.ev) cudaEventDestroy(r.ev); if (r.owned) std::free(r.host); }
  for (auto& sp : w->sink_pending) if (sp.ev) cudaEventDestroy(sp.ev);
  for (auto e : w->ev_pool) cudaEventDestroy(e);
  for (auto& m : w->modules) g_mod.cuModuleUnload(m.second);
  for (uint32_t a = 1; a <= TFCS_MAX_ARENAS; ++a) drop_arena(w, a);
  if (w->arena) cudaFreeHost(w->arena);
  if (w->d_digest) cudaFree(w->d_digest);
  if (w->copy_stream) cudaStreamDestroy(w->copy_stream);
  if (w->exec_stream) cudaStreamDestroy(w->exec_stream);
  if (w->green) g_mod.cuGreenCtxDestroy(w->green);
  delete w;
  return TFW_OK;
}

tfw_status tfw_host_alloc(size_t bytes, void** out) {
  if (!out || !bytes) return TFW_ERR_INVALID;
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
  if (e == cudaSuccess) return TFW_OK;
  cudaGetLastError();
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? TFW_ERR_NO_DEVICE : TFW_ERR_EXHAUSTED;
}
tfw_status tfw_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? TFW_OK : TFW_ERR_FAILED; }
tfw_status tfw_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return TFW_ERR_INVALID;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e == cudaSuccess) return TFW_OK;
  cudaGetLastError();
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? TFW_ERR_NO_DEVICE : TFW_ERR_FAILED;
}
tfw_status tfw_host_unregister(void* p) { return cudaHostUnregister(p) == cudaSuccess ? TFW_OK : TFW_ERR_FAILED; }

tfw_status tfw_set_arena_prefix(tfw_worker* w, const char* prefix) {
  if (!w) return TFW_ERR_INVALID;
  w->arena_prefix = prefix ? prefix : "";
  return TFW_OK;
}

tfw_status tfw_set_response_sink(tfw_worker* w, const tfw_response_sink* sink) {
  if (!w) return TFW_ERR_INVALID;
  if (!w->resp.empty() || !w->sink_pending.empty() || w->d2h_active) return fail(w, TFW_ERR_INVALID, "responses are still in flight");
  if (!sink) { w->sink_ring = nullptr; return TFW_OK; }
  if (sink->struct_size != sizeof *sink || !sink->ring || !sink->head || !sink->tail || sink->ring_bytes < 4096 || (sink->ring_bytes & 63u) ||
      !is_pinned(sink->ring))
    return fail(w, TFW_ERR_INVALID, "response sink must be a page-locked ring (tfw_host_register) of a multiple of 64 bytes");
  w->sink_ring = static_cast<uint8_t*>(sink->ring);
  w->sink_size = sink->ring_bytes;
  w->sink_head = sink->head;
  w->sink_tail = sink->tail;
  w->sink_wr = __atomic_load_n(sink->head, __ATOMIC_ACQUIRE);
  return TFW_OK;
}

tfw_status tfw_submit(tfw_worker* w, const void* stream, size_t nbytes, size_t* consumed) {
  if (!w || (!stream && nbytes)) return TFW_ERR_INVALID;
  if (consumed) *consumed = 0;
  if (!nbytes && !w->d2h_active) return TFW_OK;
  if (w->frozen) return fail(w, TFW_ERR_NOT_SUPPORTED, "vGPU is frozen: call tfw_worker_resume first");
  cudaSetDevice(w->device);
  w->input_pinned = nbytes && is_pinned(stream);
  tfw_status rc = parse(w, static_cast<const uint8_t*>(stream), nbytes, consumed);
  tfw_status fl = flush_batch(w);  // kick: every submit ends with its work enqueued
  return rc != TFW_OK ? rc : fl;
}

tfw_status tfw_worker_freeze(tfw_worker* w, uint64_t* moved_bytes) {
  if (!w) return TFW_ERR_INVALID;
  if (moved_bytes) *moved_bytes = 0;
  if (w->rec) return fail(w, TFW_ERR_NOT_SUPPORTED, "freeze while a trace is being recorded");
  tfw_status s = tfw_flush(w);
  if (s != TFW_OK) return s;
  if (w->frozen) return TFW_OK;
  uint64_t moved = 0;
  if (w->vs) {
    std::vector<uint32_t> regs;
    for (uint32_t r = 0; r < w->vs_used.size(); ++r) {
      uint32_t tier = 0;
      if (w->vs_used[r] && tfw_vspace_residency(w->vs, r, &tier, nullptr) == TFW_OK && (tier == TFW_TIER_HOME || tier == TFW_TIER_PEER)) regs.push_back(r);
    }
    for (size_t off = 0; off < regs.size(); off += 16) {  // batches of 16 regions
      const uint32_t n = (uint32_t)std::min<size_t>(16, regs.size() - off);
      std::vector<uint8_t> tiers(n, (uint8_t)TFW_TIER_HOST);
      tfw_migrate_result res{};
      s = tfw_vspace_migrate(w->vs, regs.data() + off, tiers.data(), nullptr, n, &res);
      if (s != TFW_OK) { w->err = std::string("freeze: ") + tfw_vspace_last_error(w->vs); return s; }
      moved += res.bytes;
    }
  } else {
    // plain buffers: park each in host memory, give the HBM back.  Two passes so that a host
    // allocation failure leaves the vGPU untouched.
    for (Buffer& b : w->bufs) {
      if (!b.live || b.tiered) continue;
      b.parked = park_alloc(b.size);
      if (!b.parked) {
        for (Buffer& u : w->bufs) { if (u.live && !u.tiered) { park_free(u.parked, u.size); u.parked = nullptr; } }
        return fail(w, TFW_ERR_EXHAUSTED, "freeze: not enough host memory to park the vGPU");
      }
    }
    {
      const auto tp0 = std::chrono::steady_clock::now();
      ParkPipe pipe;
      bool ok = pipe.start(w->device, w->exec_stream, staging_slots(w));
      for (Buffer& b : w->bufs) {
        if (!ok) break;
        if (!b.live || b.tiered) continue;
        ok = pipe.park(b.ptr, b.parked, b.size);
      }
      ok = ok && pipe.drain();
      if (getenv("TFW_LOG_PARK")) fprintf(stderr, "[tfw] park: copy phase %.1f ms (%zu slots of %llu MiB, %zu threads)\n",
          std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count(), pipe.slot.size(), (unsigned long long)(pipe.chunk >> 20), pipe.threads.size());
      if (!ok || cudaStreamSynchronize(w->exec_stream) != cudaSuccess) {
        cudaGetLastError();
        for (Buffer& u : w->bufs) { if (u.live && !u.tiered) { park_free(u.parked, u.size); u.parked = nullptr; } }
        return fail(w, TFW_ERR_FAILED, "freeze: copying the vGPU out of HBM failed");
      }
    }
    for (Buffer& b : w->bufs) {
      if (!b.live || b.tiered) continue;
      CU_OK(w, cudaFreeAsync(reinterpret_cast<void*>(b.ptr), w->exec_stream));
      b.ptr = 0;
      moved += b.size;
      w->parked_bytes += b.size;
      w->st.d2h_bytes += b.size;
    }
    CU_OK(w, cudaStreamSynchronize(w->exec_stream));
    cudaMemPool_t pool = nullptr;  // hand the freed blocks back to the driver now, not at some later sync
    if (cudaDeviceGetDefaultMemPool(&pool, w->device) == cudaSuccess && pool) cudaMemPoolTrimTo(pool, 0);
  }
  w->frozen = true;
  {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    w->frozen_unix_ms = (uint64_t)ts.tv_sec * 1000 + (uint64_t)ts.tv_nsec / 1000000;
  }
  w->last_moved = moved;
  if (moved_bytes) *moved_bytes = moved;
  publish_stats(w);
  return TFW_OK;
}

tfw_status tfw_worker_resume(tfw_worker* w) {
  if (!w) return TFW_ERR_INVALID;
  if (!w->frozen) return TFW_OK;
  cudaSetDevice(w->device);
  if (!w->vs) {
    // allocate everything first: a partial resume would leave the vGPU half on the GPU
    const auto tr0 = std::chrono::steady_clock::now();
    std::vector<std::pair<Buffer*, void*>> fresh;
    for (Buffer& b : w->bufs) {
      if (!b.live || b.tiered || !b.parked) continue;
      void* p = nullptr;
      if (cudaMallocAsync(&p, b.size, w->exec_stream) != cudaSuccess) {
        cudaGetLastError();
        for (auto& f : fresh) cudaFreeAsync(f.second, w->exec_stream);
        cudaStreamSynchronize(w->exec_stream);
        return fail(w, TFW_ERR_EXHAUSTED, "resume: the vGPU's HBM is not available yet");
      }
      fresh.emplace_back(&b, p);
    }
    {
      const auto tr1 = std::chrono::steady_clock::now();
      ParkPipe pipe;
      bool ok = pipe.start(w->device, w->exec_stream, staging_slots(w));
      const auto tr2 = std::chrono::steady_clock::now();
      for (auto& f : fresh) {
        if (!ok) break;
        ok = pipe.unpark(reinterpret_cast<uint64_t>(f.second), f.first->parked, f.first->size);
        w->st.h2d_dma_bytes += f.first->size;
      }
      if (getenv("TFW_LOG_PARK"))
        fprintf(stderr, "[tfw] resume: allocation %.1f ms, ring %.1f ms, copy phase %.1f ms (waiting for fills %.1f, for the ring %.1f, issuing %.1f)\n",
                std::chrono::duration<double, std::milli>(tr1 - tr0).count(), std::chrono::duration<double, std::milli>(tr2 - tr1).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr2).count(), pipe.t_fill_wait * 1e3, pipe.t_ring_wait * 1e3, pipe.t_issue * 1e3);
      if (!ok || cudaStreamSynchronize(w->exec_stream) != cudaSuccess) {
        cudaGetLastError();
        for (auto& f : fresh) cudaFreeAsync(f.second, w->exec_stream);
        cudaStreamSynchronize(w->exec_stream);
        return fail(w, TFW_ERR_FAILED, "resume: copying the vGPU back into HBM failed");
      }
    }
    // Giving the parked pages back takes the kernel ~10 ms per 256 MiB with the address space locked: one thread for all
    // of them, started last (a thread per buffer made every later pthread_create wait for the previous munmap).
    std::vector<std::pair<uint8_t*, uint64_t>> gone;
    for (auto& f : fresh) {
      Buffer& b = *f.first;
      b.ptr = reinterpret_cast<uint64_t>(f.second);
      gone.emplace_back(b.parked, b.size);
      b.parked = nullptr;
    }
    std::thread([gone] { for (auto& g : gone) munmap(g.first, g.second); }).detach();
    w->parked_bytes = 0;
  }
  w->frozen = false;
  w->frozen_auto = false;
  publish_stats(w);
  return TFW_OK;
}

// The worker's own idle policy (auto_freeze.freeze_to_mem_ttl, api/http_types.go:82-100): the same freeze, marked
// as self-inflicted so that the next client byte may undo it (a freeze ordered by the provider stays until the
// provider resumes it).
tfw_status tfw_worker_auto_freeze(tfw_worker* w, uint64_t* moved_bytes) {
  if (!w) return TFW_ERR_INVALID;
  if (w->frozen) return TFW_OK;
  tfw_status s = tfw_worker_freeze(w, moved_bytes);
  if (s == TFW_OK) { w->frozen_auto = true; w->auto_freezes++; publish_stats(w); }
  return s;
}
tfw_status tfw_worker_auto_resume(tfw_worker* w) {
  if (!w) return TFW_ERR_INVALID;
  if (!w->frozen || !w->frozen_auto) return TFW_OK;  // not frozen, or frozen by the provider: not ours to undo
  tfw_status s = tfw_worker_resume(w);
  if (s == TFW_OK) { w->auto_resumes++; publish_stats(w); }
  return s;
}

// Change the hard compute limit of a running vGPU: drain, build the new SM partition and its stream, switch.
tfw_status tfw_worker_set_sm_limit(tfw_worker* w, uint32_t percent) {
  if (!w || percent > 100) return TFW_ERR_INVALID;
  if (percent == 100) percent = 0;
  if (percent == w->sm_percent) return TFW_OK;
  if (w->rec) return fail(w, TFW_ERR_NOT_SUPPORTED, "compute limit change while a trace is being recorded");
  tfw_status s = tfw_flush(w);
  if (s != TFW_OK) return s;
  cudaStream_t ns = nullptr;
  CUgreenCtx ng = nullptr;
  int nsms = w->device_sms;
  s = make_exec_stream(w, percent, &ns, &ng, &nsms);
  if (s != TFW_OK) return s;
  if (w->vs) tfw_vspace_bind_stream(w->vs, ns);
  cudaStreamDestroy(w->exec_stream);
  if (w->green) g_mod.cuGreenCtxDestroy(w->green);
  w->exec_stream = ns;
  w->green = ng;
  w->sm_count = nsms;
  w->sm_percent = percent;
  for (auto& f : w->fences) { if (f.copy) cudaEventDestroy(f.copy); if (f.exec) cudaEventDestroy(f.exec); }
  w->fences.clear();  // tickets handed out before the switch are complete (we drained); tfw_fence re-creates the ring
  publish_stats(w);
  return TFW_OK;
}

tfw_status tfw_worker_set_vram_limit(tfw_worker* w, uint64_t bytes) {
  if (!w) return TFW_ERR_INVALID;
  w->cfg.vram_limit_bytes = bytes;  // live buffers stay; further MALLOCs are checked against the new quota
  return TFW_OK;
}

tfw_status tfw_worker_poll_control(tfw_worker* w, int* frozen) {
  if (!w) return TFW_ERR_INVALID;
  tfw_status rc = TFW_OK;
  if (tfw_stats_record* r = w->pub) {
    const uint64_t req = __atomic_load_n(&r->ctl_request, __ATOMIC_ACQUIRE);
    if (req != w->ctl_seen) {
      w->ctl_seen = req;
      const uint32_t cmd = (uint32_t)(req & 0xff);
      if (cmd == TFW_CTL_FREEZE) { rc = tfw_worker_freeze(w, nullptr); w->frozen_auto = false; }
      else if (cmd == TFW_CTL_RESUME) rc = tfw_worker_resume(w);
      else if (cmd == TFW_CTL_SM_LIMIT) rc = tfw_worker_set_sm_limit(w, (uint32_t)r->ctl_arg);
      else if (cmd == TFW_CTL_MEM_LIMIT) rc = tfw_worker_set_vram_limit(w, r->ctl_arg);
      else rc = TFW_ERR_INVALID;
      r->ctl_status = (uint64_t)rc;
      publish_stats(w);
      __atomic_store_n(&r->ctl_ack, req, __ATOMIC_RELEASE);
    }
    // an idle or frozen worker submits nothing: keep the record fresh so the provider still sees it
    else if ((uint64_t)time(nullptr) > r->updated_unix_secs + 2) publish_stats(w);
  }
  if (frozen) *frozen = w->frozen ? (w->frozen_auto ? 2 : 1) : 0;
  return rc;
}

tfw_status tfw_fence(tfw_worker* w, uint64_t* ticket) {
  if (!w || !ticket) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  tfw_status s = flush_batch(w);
  if (s != TFW_OK) return s;
  if (w->fences.empty()) {
    w->fences.resize(8);
    for (auto& f : w->fences) {
      CU_OK(w, cudaEventCreateWithFlags(&f.copy, cudaEventDisableTiming));
      CU_OK(w, cudaEventCreateWithFlags(&f.exec, cudaEventDisableTiming));
    }
  }
  tfw_worker::Fence& f = w->fences[w->fence_next % w->fences.size()];
  CU_OK(w, cudaEventRecord(f.copy, w->copy_stream));
  CU_OK(w, cudaEventRecord(f.exec, w->exec_stream));
  *ticket = w->fence_next++;
  return TFW_OK;
}

tfw_status tfw_fence_wait(tfw_worker* w, uint64_t ticket) {
  if (!w || ticket == 0 || ticket >= w->fence_next) return TFW_ERR_INVALID;
  if (w->fences.empty()) return TFW_OK;  // the ring was reset by a drain (tfw_worker_set_sm_limit): everything before it is complete
  // A slot that was re-used holds a younger fence of the same two streams: its completion implies ours.
  tfw_worker::Fence& f = w->fences[ticket % w->fences.size()];
  CU_OK(w, cudaEventSynchronize(f.copy));
  CU_OK(w, cudaEventSynchronize(f.exec));
  return TFW_OK;
}

tfw_status tfw_fence_query(tfw_worker* w, uint64_t ticket, int* done) {
  if (!w || !done || ticket == 0 || ticket >= w->fence_next) return TFW_ERR_INVALID;
  *done = 1;
  if (w->fences.empty()) return TFW_OK;
  tfw_worker::Fence& f = w->fences[ticket % w->fences.size()];  // (a re-used slot holds a younger fence: see tfw_fence_wait)
  for (cudaEvent_t e : {f.copy, f.exec}) {
    const cudaError_t q = cudaEventQuery(e);
    if (q == cudaErrorNotReady) { *done = 0; return TFW_OK; }
    if (q != cudaSuccess) { w->err = std::string("cudaEventQuery: ") + cudaGetErrorString(q); return TFW_ERR_FAILED; }
  }
  return TFW_OK;
}

tfw_status tfw_flush(tfw_worker* w) {
  if (!w) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  tfw_status s = flush_batch(w);
  if (s != TFW_OK) return s;
  CU_OK(w, cudaStreamSynchronize(w->copy_stream));
  CU_OK(w, cudaStreamSynchronize(w->exec_stream));
  for (auto& sl : w->slots) sl.busy = false;
  publish_stats(w);
  return TFW_OK;
}

tfw_status tfw_poll_responses(tfw_worker* w, void* out, size_t cap, size_t* nbytes) {
  if (!w || !nbytes || (!out && cap)) return TFW_ERR_INVALID;
  *nbytes = 0;
  if (w->sink_ring) {  // responses are produced in the ring itself: move what is queued, publish what has completed
    sink_pump(w);
    return TFW_OK;
  }
  uint8_t* o = static_cast<uint8_t*>(out);
  while (!w->resp.empty()) {
    Response& r = w->resp.front();
    if (r.ev && cudaEventQuery(r.ev) != cudaSuccess) { cudaGetLastError(); break; }
    // The wire is a byte stream: a response larger than the caller's buffer leaves in pieces.
    const uint64_t padded = tfcs_pad16(r.len), total = TFCS_HDR_BYTES + padded;
    while (r.sent < total && *nbytes < cap) {
      const uint64_t room = cap - *nbytes;
      uint64_t k;
      if (r.sent < TFCS_HDR_BYTES) {
        k = std::min<uint64_t>(TFCS_HDR_BYTES - r.sent, room);
        std::memcpy(o + *nbytes, reinterpret_cast<const uint8_t*>(&r.hdr) + r.sent, k);
      } else {
        const uint64_t off = r.sent - TFCS_HDR_BYTES;
        if (off < r.len) {
          k = std::min<uint64_t>(r.len - off, room);
          std::memcpy(o + *nbytes, r.host + off, k);
        } else {
          k = std::min<uint64_t>(padded - off, room);
          std::memset(o + *nbytes, 0, k);
        }
      }
      r.sent += k;
      *nbytes += k;
    }
    if (r.sent < total) break;  // buffer full: the rest follows at the next call
    if (r.ev) w->ev_pool.push_back(r.ev);
    if (r.owned) std::free(r.host);
    w->resp.pop_front();
  }
  if (w->resp.empty()) w->arena_used = 0;
  return TFW_OK;
}

// ---- resident trace ---------------------------------------------------------
tfw_status tfw_trace_load(tfw_worker* w, const void* stream, size_t nbytes, tfw_trace** out) {
  if (!w || !stream || !nbytes || !out) return TFW_ERR_INVALID;
  *out = nullptr;
  cudaSetDevice(w->device);
  tfw_status s = tfw_flush(w);
  if (s != TFW_OK) return s;
  if (w->in_payload) return fail(w, TFW_ERR_PROTOCOL, "trace load in the middle of a payload");
  if (w->vs) return fail(w, TFW_ERR_NOT_SUPPORTED, "resident traces are not supported on a tiered worker");
  tfw_trace* t = new (std::nothrow) tfw_trace();
  if (!t) return TFW_ERR_EXHAUSTED;
  const uint64_t mis = reinterpret_cast<uintptr_t>(stream) & 15u;
  if (cudaMalloc(reinterpret_cast<void**>(&t->d_stream), nbytes + 32) != cudaSuccess) { cudaGetLastError(); delete t; return fail(w, TFW_ERR_EXHAUSTED, "no HBM for the resident trace"); }
  t->host_base = static_cast<const uint8_t*>(stream);
  t->dev_base = reinterpret_cast<uint64_t>(t->d_stream) + mis;
  if (cudaMemcpyAsync(reinterpret_cast<void*>(t->dev_base), stream, nbytes, cudaMemcpyHostToDevice, w->exec_stream) != cudaSuccess ||
      cudaStreamSynchronize(w->exec_stream) != cudaSuccess) {
    cudaGetLastError(); cudaFree(t->d_stream); delete t; return fail(w, TFW_ERR_FAILED, "resident trace upload failed");
  }
  const tfw_stats saved = w->st;
  const std::vector<Buffer> saved_bufs = w->bufs;
  w->rec = t;
  size_t consumed = 0;
  s = parse(w, t->host_base, nbytes, &consumed);
  if (s == TFW_OK) s = flush_batch(w);
  w->rec = nullptr;
  if (s == TFW_OK && (consumed != nbytes || w->in_payload)) { w->in_payload = false; s = fail(w, TFW_ERR_PROTOCOL, "trace ends inside a frame"); }
  // the live session's handle table is not disturbed by loading a trace
  t->bufs = w->bufs;
  w->bufs = saved_bufs;
  w->st = saved;
  w->descs.clear(); w->wr.clear(); w->rd.clear();
  if (s == TFW_OK && !t->descs.empty()) {
    if (cudaMalloc(reinterpret_cast<void**>(&t->d_descs), t->descs.size() * sizeof(tfw_move_desc)) != cudaSuccess ||
        cudaMemcpyAsync(t->d_descs, t->descs.data(), t->descs.size() * sizeof(tfw_move_desc), cudaMemcpyHostToDevice, w->exec_stream) != cudaSuccess ||
        cudaStreamSynchronize(w->exec_stream) != cudaSuccess) {
      cudaGetLastError();
      s = fail(w, TFW_ERR_EXHAUSTED, "no HBM for the trace descriptor tables");
    }
  }
  if (s != TFW_OK) { tfw_trace_free(w, t); return s; }
  *out = t;
  return TFW_OK;
}

tfw_status tfw_trace_replay(tfw_worker* w, tfw_trace* t) {
  if (!w || !t) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  for (const Step& s : t->steps) {
    switch (s.kind) {
      case kStepMover:
        CU_OK(w, tfw::launch_mover(t->d_descs + s.desc_off, s.ndesc, s.tiles, w->sm_count, w->ctas_per_sm,
                                   w->mover == tfw::kMoverTma && s.bulk_ok ? tfw::kMoverTma : tfw::kMoverLdg, w->exec_stream));
        w->st.mover_launches++;
        break;
      case kStepLaunch: { tfw_status r = issue_launch(w, s.hdr, s.ptr); if (r != TFW_OK) return r; break; }
      case kStepD2H: { tfw_status r = issue_d2h(w, s.hdr, s.ptr); if (r != TFW_OK) return r; break; }
      case kStepSync: { tfw_status r = issue_sync(w, s.hdr); if (r != TFW_OK) return r; break; }
    }
  }
  return TFW_OK;
}

tfw_status tfw_trace_info(const tfw_trace* t, uint64_t* payload_bytes, uint64_t* mover_launches, uint64_t* algorithmic_bytes) {
  if (!t) return TFW_ERR_INVALID;
  if (payload_bytes) *payload_bytes = t->payload_bytes;
  if (mover_launches) *mover_launches = t->mover_launches;
  if (algorithmic_bytes) *algorithmic_bytes = t->algo_bytes;
  return TFW_OK;
}

tfw_status tfw_trace_buffer_info(const tfw_trace* t, uint32_t handle, uint64_t* size, uint64_t* dev_ptr) {
  if (!t) return TFW_ERR_INVALID;
  if (handle >= t->bufs.size() || !t->bufs[handle].live) return TFW_ERR_NOT_FOUND;
  if (size) *size = t->bufs[handle].size;
  if (dev_ptr) *dev_ptr = t->bufs[handle].ptr;
  return TFW_OK;
}

tfw_status tfw_trace_free(tfw_worker* w, tfw_trace* t) {
  if (!w || !t) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  cudaStreamSynchronize(w->exec_stream);
  for (uint64_t p : t->allocs) cudaFree(reinterpret_cast<void*>(p));
  if (t->d_descs) cudaFree(t->d_descs);
  if (t->d_stream) cudaFree(t->d_stream);
  delete t;
  return TFW_OK;
}

// ---- introspection ------------------------------------------------------------
tfw_status tfw_buffer_info(tfw_worker* w, uint32_t handle, uint64_t* size, uint64_t* dev_ptr) {
  if (!w) return TFW_ERR_INVALID;
  Buffer* b = find(w, handle);
  if (!b) return TFW_ERR_NOT_FOUND;
  if (size) *size = b->size;
  if (dev_ptr) *dev_ptr = b->ptr;
  return TFW_OK;
}

tfw_status tfw_buffer_read(tfw_worker* w, uint32_t handle, uint64_t off, void* dst, uint64_t n) {
  if (!w || (!dst && n)) return TFW_ERR_INVALID;
  Buffer* b = find(w, handle);
  if (!b) return TFW_ERR_NOT_FOUND;
  if (off > b->size || n > b->size - off) return TFW_ERR_INVALID;
  if (b->parked) {  // frozen vGPU: the bytes are in host memory
    std::memcpy(dst, b->parked + off, n);
    return TFW_OK;
  }
  tfw_status s = tfw_flush(w);
  if (s != TFW_OK) return s;
  // a tiered buffer may be larger than the HBM budget: read it region by region
  uint64_t o = 0;
  while (o < n) {
    const uint64_t start = b->ptr + off + o;
    uint64_t piece = n - o;
    if (b->tiered) {
      piece = std::min(piece, w->vs_R - (start - w->vs_base) % w->vs_R);
      s = touch_range(w, start, piece);
      if (s != TFW_OK) return s;
    }
    // on the vGPU's own stream: a region that was just prefetched is ordered behind its copy there (and only there)
    CU_OK(w, cudaMemcpyAsync(static_cast<uint8_t*>(dst) + o, reinterpret_cast<void*>(start), piece, cudaMemcpyDeviceToHost, w->exec_stream));
    CU_OK(w, cudaStreamSynchronize(w->exec_stream));
    o += piece;
  }
  return TFW_OK;
}

tfw_status tfw_dev_digest(tfw_worker* w, uint64_t dev_ptr, uint64_t bytes, uint64_t* digest) {
  if (!w || !digest || !dev_ptr || (dev_ptr & 7u)) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  tfw_status s = flush_batch(w);
  if (s != TFW_OK) return s;
  CU_OK(w, cudaMemsetAsync(w->d_digest, 0, sizeof(unsigned long long), w->exec_stream));
  CU_OK(w, tfw::launch_digest(reinterpret_cast<void*>(dev_ptr), bytes, w->d_digest, w->sm_count, w->exec_stream));
  unsigned long long sum = 0;
  CU_OK(w, cudaMemcpyAsync(&sum, w->d_digest, sizeof(sum), cudaMemcpyDeviceToHost, w->exec_stream));
  CU_OK(w, cudaStreamSynchronize(w->exec_stream));
  w->st.other_launches++;
  *digest = tfw::digest_mix((uint64_t)sum ^ (bytes * tfw::kDigestK1));
  return TFW_OK;
}

tfw_status tfw_buffer_digest(tfw_worker* w, uint32_t handle, uint64_t* digest) {
  if (!w || !digest) return TFW_ERR_INVALID;
  Buffer* b = find(w, handle);
  if (!b) return TFW_ERR_NOT_FOUND;
  if (b->parked) return fail(w, TFW_ERR_NOT_SUPPORTED, "vGPU is frozen: the buffer is parked in host memory");
  if (b->tiered) {  // the digest kernel reads the whole buffer: all of it must be resident at once
    tfw_status s = touch_range(w, b->ptr, b->size);
    if (s != TFW_OK) return s;
  }
  return tfw_dev_digest(w, b->ptr, b->size, digest);
}

tfw_status tfw_get_stats(tfw_worker* w, tfw_stats* out) {
  if (!w || !out) return TFW_ERR_INVALID;
  *out = w->st;
  return TFW_OK;
}

tfw_status tfw_worker_gate_state(tfw_worker* w, tfw_gate_state* out) {
  if (!w || !out) return TFW_ERR_INVALID;
  if (!w->gate) return TFW_ERR_NOT_FOUND;  // no limiter configured for this vGPU
  return tfw_gate_get_state(w->gate, out);
}

void* tfw_exec_stream(tfw_worker* w) { return w ? static_cast<void*>(w->exec_stream) : nullptr; }

// ---- kernel-level entry points ------------------------------------------------
tfw_status tfw_move_batch(tfw_worker* w, tfw_move_desc* descs, uint32_t n, float* ms) {
  if (!w || !descs || !n) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  tfw_status s = flush_batch(w);
  if (s != TFW_OK) return s;
  const uint32_t tiles = assign_tiles(descs, n);
  tfw_move_desc* d = nullptr;
  CU_OK(w, cudaMallocAsync(reinterpret_cast<void**>(&d), sizeof(tfw_move_desc) * n, w->exec_stream));
  CU_OK(w, cudaMemcpyAsync(d, descs, sizeof(tfw_move_desc) * n, cudaMemcpyHostToDevice, w->exec_stream));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ms) {
    CU_OK(w, cudaEventCreate(&e0));
    CU_OK(w, cudaEventCreate(&e1));
    CU_OK(w, cudaStreamSynchronize(w->exec_stream));
    CU_OK(w, cudaEventRecord(e0, w->exec_stream));
  }
  CU_OK(w, tfw::launch_mover(d, n, tiles, w->sm_count, w->ctas_per_sm,
                             w->mover == tfw::kMoverTma && tfw::mover_bulk_ok(descs, n) ? tfw::kMoverTma : tfw::kMoverLdg, w->exec_stream));
  w->st.mover_launches++;
  if (ms) {
    CU_OK(w, cudaEventRecord(e1, w->exec_stream));
    CU_OK(w, cudaEventSynchronize(e1));
    CU_OK(w, cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  CU_OK(w, cudaFreeAsync(d, w->exec_stream));
  return TFW_OK;
}

tfw_status tfw_dev_alloc(tfw_worker* w, uint64_t bytes, uint64_t* dev_ptr) {
  if (!w || !dev_ptr || !bytes) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return fail(w, TFW_ERR_EXHAUSTED, "cudaMalloc failed"); }
  *dev_ptr = reinterpret_cast<uint64_t>(p);
  return TFW_OK;
}
tfw_status tfw_dev_free(tfw_worker* w, uint64_t dev_ptr) {
  if (!w || !dev_ptr) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  cudaStreamSynchronize(w->exec_stream);
  CU_OK(w, cudaFree(reinterpret_cast<void*>(dev_ptr)));
  return TFW_OK;
}
tfw_status tfw_dev_write(tfw_worker* w, uint64_t dev_ptr, const void* src, uint64_t n) {
  if (!w || !dev_ptr || (!src && n)) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  // stream-ordered on the exec stream: a plain cudaMemcpy from pageable memory may return
  // before the DMA lands, and the exec stream (non-blocking) would not wait for it
  if (n) {
    CU_OK(w, cudaMemcpyAsync(reinterpret_cast<void*>(dev_ptr), src, n, cudaMemcpyHostToDevice, w->exec_stream));
    CU_OK(w, cudaStreamSynchronize(w->exec_stream));
  }
  return TFW_OK;
}
tfw_status tfw_dev_read(tfw_worker* w, uint64_t dev_ptr, void* dst, uint64_t n) {
  if (!w || !dev_ptr || (!dst && n)) return TFW_ERR_INVALID;
  cudaSetDevice(w->device);
  CU_OK(w, cudaStreamSynchronize(w->exec_stream));
  if (n) CU_OK(w, cudaMemcpy(dst, reinterpret_cast<void*>(dev_ptr), n, cudaMemcpyDeviceToHost));
  return TFW_OK;
}

}  // extern "C"
