// limiter_api.cc -- the limiter half of the provider C-ABI (provider/limiter.h:71-106).
//
// Hypervisor-facing functions manage the per-pod quota files and run the ERL
// step; the reference's example provider stubs all of them with NOT_SUPPORTED
// (provider/example/accelerator.c:212-256) and the Go hypervisor currently
// does the same work in pure Go (worker/controller.go:501) -- these are the
// native equivalents.  Worker-facing functions are the CPU-side gate used by
// local soft mode (LD_PRELOAD hook); the remote worker of
// this repo gates launches on the GPU instead (gate.cu).
#include <strings.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "erl.h"
#include "provider_log.h"
#include "shm_quota.h"
#include "worker_ctl.h"
#include "tf_provider_abi.h"
#include "tfw_stats_file.h"

namespace {

struct WorkerEntry {
  std::unique_ptr<tfq::QuotaFile> file;
  tferl::State erl[TF_SHM_MAX_DEVICES];
};

std::mutex g_mu;
std::string g_base;                                   // LimiterInit(shmBasePath)
tferl::Config g_cfg;
std::map<std::string, WorkerEntry> g_workers;         // "ns/pod"

// worker-side state (one quota file per process: TF_SHM_PATH)
std::unique_ptr<tfq::QuotaFile> g_self;
bool g_self_tried = false;
int64_t g_local_bytes[TF_SHM_MAX_DEVICES] = {0};
std::map<std::string, WorkerFreezeState> g_frozen;

AccelResult map_status(tfq::Status s) {
  switch (s) {
    case tfq::kOk: return ACCEL_SUCCESS;
    case tfq::kInvalid: return ACCEL_ERROR_INVALID_PARAM;
    case tfq::kNotFound: return ACCEL_ERROR_NOT_FOUND;
    default: return ACCEL_ERROR_OPERATION_FAILED;
  }
}

// caller holds g_mu
AccelResult get_worker(const char* ns, const char* pod, WorkerEntry** out) {
  if (!ns || !pod || !*ns || !*pod) return ACCEL_ERROR_INVALID_PARAM;
  if (g_base.empty()) return ACCEL_ERROR_OPERATION_FAILED;  // LimiterInit not called
  const std::string key = std::string(ns) + "/" + pod;
  auto it = g_workers.find(key);
  if (it == g_workers.end()) {
    tfq::QuotaFile* q = nullptr;
    std::string err;
    tfq::Status s = tfq::QuotaFile::open(g_base, ns, pod, &q, &err);
    if (s != tfq::kOk) {
      tfprov::log("WARN", ("limiter: cannot open quota file of " + key + ": " + err).c_str());
      return map_status(s);
    }
    it = g_workers.emplace(key, WorkerEntry{}).first;
    it->second.file.reset(q);
  }
  *out = &it->second;
  return ACCEL_SUCCESS;
}

// UUIDs appear as "GPU-xxxx" (legacy.go:524-529) or stripped + upper-cased
// (worker/controller.go:562-570): compare case-insensitively without the prefix.
std::string canon_uuid(const char* u) {
  std::string s = u ? u : "";
  if (s.size() >= 4 && strncasecmp(s.c_str(), "GPU-", 4) == 0) s = s.substr(4);
  for (auto& c : s) c = (char)toupper((unsigned char)c);
  return s;
}

tfq::QuotaFile* self_file() {
  if (!g_self_tried) {
    g_self_tried = true;
    const char* p = getenv("TF_SHM_PATH");  // pkg/constants/env.go:133-138
    if (p && *p) {
      tfq::QuotaFile* q = nullptr;
      std::string err;
      if (tfq::QuotaFile::open_file(p, &q, &err) == tfq::kOk) g_self.reset(q);
      else tfprov::log("WARN", (std::string("limiter: TF_SHM_PATH unusable: ") + err).c_str());
    }
  }
  return g_self.get();
}

int device_index_of(tfq::QuotaFile* q, const char* uuid) {
  const std::string want = canon_uuid(uuid);
  for (uint32_t i = 0; i < TF_SHM_MAX_DEVICES; ++i) {
    if (!q->has_device(i)) continue;
    char buf[TF_SHM_UUID_LEN + 1] = {0};
    std::memcpy(buf, q->raw()->devices[i].uuid, TF_SHM_UUID_LEN);
    if (canon_uuid(buf) == want) return (int)i;
  }
  return -1;
}

uint64_t now_ms() {
  timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  return (uint64_t)ts.tv_sec * 1000 + (uint64_t)ts.tv_nsec / 1000000;
}

}  // namespace

namespace tfprov {
std::string limiter_base() {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_base;
}

bool self_limits(const char* uuid, uint64_t* mem_limit, uint64_t* mem_used, uint32_t* up_limit) {
  std::lock_guard<std::mutex> lk(g_mu);
  tfq::QuotaFile* q = self_file();
  if (!q) return false;
  const int idx = device_index_of(q, uuid);
  if (idx < 0) return false;
  uint64_t used = q->pod_memory_used((uint32_t)idx);
  if (g_local_bytes[idx] > 0 && (uint64_t)g_local_bytes[idx] > used) used = (uint64_t)g_local_bytes[idx];
  if (mem_limit) *mem_limit = q->raw()->devices[idx].mem_limit;
  if (mem_used) *mem_used = used;
  if (up_limit) *up_limit = q->raw()->devices[idx].up_limit;
  return true;
}

void* self_bucket(const char* uuid, int* device_index) {
  std::lock_guard<std::mutex> lk(g_mu);
  tfq::QuotaFile* q = self_file();
  if (!q) return nullptr;
  const int idx = device_index_of(q, uuid);
  if (idx < 0) return nullptr;
  *device_index = idx;
  return q;  // lives as long as the process (g_self is never reset)
}

double self_charge(void* bucket, int device_index, double cost) {
  return static_cast<tfq::QuotaFile*>(bucket)->fetch_sub((uint32_t)device_index, cost);  // soft_limiter_shm.go:715-731
}
// FetchSub never admits a cost above the bucket's capacity (the controller keeps capacity = clamp(rate * 0.5 s,
// 200, 200000), quota_controller.go:425-433), so one launch of 16K blocks x 8 warps would wait for ever: a launch is
// charged at most one full bucket.
double clamp_cost(void* bucket, int device_index, double cost) {
  const double cap = static_cast<tfq::QuotaFile*>(bucket)->capacity((uint32_t)device_index);
  return cap > 0.0 && cost > cap ? cap : cost;
}
}  // namespace tfprov

extern "C" {

// ---------------------------------------------------------------- hypervisor-facing
AccelResult LimiterInit(const char* shmBasePath) {
  if (!shmBasePath || !*shmBasePath) return ACCEL_ERROR_INVALID_PARAM;
  std::lock_guard<std::mutex> lk(g_mu);
  g_base = shmBasePath;
  while (g_base.size() > 1 && g_base.back() == '/') g_base.pop_back();
  g_cfg = tferl::Config::from_json(getenv("TF_HYPERVISOR_SCHEDULING_CONFIG"));  // quota_controller.go:143-178
  tfprov::log("INFO", ("limiter: quota files under " + g_base).c_str());
  return ACCEL_SUCCESS;
}

AccelResult LimiterShutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_workers.clear();
  g_base.clear();
  return ACCEL_SUCCESS;
}

AccelResult LimiterCreateWorker(const char* namespace_, const char* podName, const LimiterDeviceConfig* configs,
                                size_t configCount) {
  if (!namespace_ || !podName || (!configs && configCount) || configCount > TF_SHM_MAX_DEVICES) return ACCEL_ERROR_INVALID_PARAM;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_base.empty()) return ACCEL_ERROR_OPERATION_FAILED;
  std::vector<tfq::DeviceConfig> cfgs;
  for (size_t i = 0; i < configCount; ++i) {
    tfq::DeviceConfig c;
    c.device_idx = configs[i].deviceIdx;
    c.uuid.assign(configs[i].deviceUUID, strnlen(configs[i].deviceUUID, sizeof(configs[i].deviceUUID)));
    c.up_limit = configs[i].upLimit;
    c.mem_limit = configs[i].memLimit;
    c.total_cuda_cores = configs[i].totalCudaCores;
    cfgs.push_back(c);
  }
  const std::string key = std::string(namespace_) + "/" + podName;
  g_workers.erase(key);
  tfq::QuotaFile* q = nullptr;
  std::string err;
  tfq::Status s = tfq::QuotaFile::create(g_base, namespace_, podName, cfgs, &q, &err);
  if (s != tfq::kOk) {
    tfprov::log("ERROR", ("limiter: create " + key + " failed: " + err).c_str());
    return map_status(s);
  }
  g_workers[key].file.reset(q);
  return ACCEL_SUCCESS;
}

AccelResult LimiterRemoveWorker(const char* namespace_, const char* podName) {
  std::lock_guard<std::mutex> lk(g_mu);
  WorkerEntry* w = nullptr;
  AccelResult r = get_worker(namespace_, podName, &w);
  if (r != ACCEL_SUCCESS) return r;
  // the worker's metrics/control record lives in the same directory and would keep it alive (the walk up only
  // removes empty directories); it is this stack's own file, so it goes with the quota file
  unlink((g_base + "/" + namespace_ + "/" + podName + "/" + TFW_STATS_FILE_NAME).c_str());
  tfq::Status s = w->file->cleanup(g_base);
  g_workers.erase(std::string(namespace_) + "/" + podName);
  return map_status(s);
}

AccelResult LimiterRegisterPID(const char* namespace_, const char* podName, uint32_t hostPID) {
  if (hostPID == 0) return ACCEL_ERROR_INVALID_PARAM;
  std::lock_guard<std::mutex> lk(g_mu);
  WorkerEntry* w = nullptr;
  AccelResult r = get_worker(namespace_, podName, &w);
  if (r != ACCEL_SUCCESS) return r;
  w->file->add_pid(hostPID);  // idempotent (InsertIfAbsent)
  return ACCEL_SUCCESS;
}

AccelResult LimiterUpdateERL(const char* namespace_, const char* podName, uint32_t deviceIdx, uint32_t upLimit,
                             double utilizationPercent, uint64_t timestampMicros) {
  if (deviceIdx >= TF_SHM_MAX_DEVICES || upLimit > 100) return ACCEL_ERROR_INVALID_PARAM;
  std::lock_guard<std::mutex> lk(g_mu);
  WorkerEntry* w = nullptr;
  AccelResult r = get_worker(namespace_, podName, &w);
  if (r != ACCEL_SUCCESS) return r;
  if (!w->file->has_device(deviceIdx)) return ACCEL_ERROR_NOT_FOUND;
  tferl::tick(*w->file, deviceIdx, w->erl[deviceIdx], g_cfg, upLimit, utilizationPercent, (double)timestampMicros / 1e6);
  return ACCEL_SUCCESS;
}

AccelResult LimiterUpdateHeartbeat(const char* namespace_, const char* podName, uint64_t timestampSecs) {
  std::lock_guard<std::mutex> lk(g_mu);
  WorkerEntry* w = nullptr;
  AccelResult r = get_worker(namespace_, podName, &w);
  if (r != ACCEL_SUCCESS) return r;
  w->file->update_heartbeat(timestampSecs);
  return ACCEL_SUCCESS;
}

AccelResult LimiterSetPodMemoryUsed(const char* namespace_, const char* podName, uint32_t deviceIdx, uint64_t memoryUsed) {
  std::lock_guard<std::mutex> lk(g_mu);
  WorkerEntry* w = nullptr;
  AccelResult r = get_worker(namespace_, podName, &w);
  if (r != ACCEL_SUCCESS) return r;
  return w->file->set_pod_memory_used(deviceIdx, memoryUsed) ? ACCEL_SUCCESS : ACCEL_ERROR_NOT_FOUND;
}

// ---------------------------------------------------------------- worker-facing
AccelResult AddWorkerProcess(const char* deviceUUID, const char* processId) {
  if (!deviceUUID || !processId) return ACCEL_ERROR_INVALID_PARAM;
  std::lock_guard<std::mutex> lk(g_mu);
  tfq::QuotaFile* q = self_file();
  if (!q) return ACCEL_SUCCESS;  // limiter not configured for this process: nothing to track
  const long pid = strtol(processId, nullptr, 10);
  if (pid > 0) q->add_pid((uint64_t)pid);
  return ACCEL_SUCCESS;
}

AccelResult CheckAndRecordMemoryOps(const char* processId, const char* deviceUUID, int64_t bytesDiff, MemoryOpRecord* record) {
  (void)processId;
  if (!record) return ACCEL_ERROR_INVALID_PARAM;
  std::memset(record, 0, sizeof(*record));
  if (deviceUUID) snprintf(record->deviceUUID, sizeof(record->deviceUUID), "%s", deviceUUID);
  record->bytesDiff = bytesDiff;
  std::lock_guard<std::mutex> lk(g_mu);
  tfq::QuotaFile* q = self_file();
  if (!q) { record->availableBytes = ~0ull; return ACCEL_SUCCESS; }
  const int idx = device_index_of(q, deviceUUID);
  if (idx < 0) return ACCEL_ERROR_NOT_FOUND;
  const uint64_t limit = q->raw()->devices[idx].mem_limit;
  // the hypervisor refreshes pod_memory_used at 2 Hz (worker/controller.go:454-498); between
  // refreshes this process's own running total is the fresher lower bound
  uint64_t used = q->pod_memory_used((uint32_t)idx);
  if (g_local_bytes[idx] > 0 && (uint64_t)g_local_bytes[idx] > used) used = (uint64_t)g_local_bytes[idx];
  const uint64_t avail = limit > used ? limit - used : 0;
  if (bytesDiff > 0 && (uint64_t)bytesDiff > avail) {
    record->shouldBlock = true;
    record->availableBytes = avail;
    return ACCEL_SUCCESS;
  }
  g_local_bytes[idx] += bytesDiff;
  if (g_local_bytes[idx] < 0) g_local_bytes[idx] = 0;
  if (bytesDiff >= 0) {
    record->availableBytes = avail - (uint64_t)bytesDiff;
  } else {
    const uint64_t freed = (uint64_t)(-bytesDiff);
    const uint64_t now_used = used > freed ? used - freed : 0;
    record->availableBytes = limit > now_used ? limit - now_used : 0;
  }
  return ACCEL_SUCCESS;
}

AccelResult CheckAndRecordComputeOps(const char* processId, const char* deviceUUID, uint64_t computeTokens, ComputeOpRecord* record) {
  (void)processId;
  if (!record) return ACCEL_ERROR_INVALID_PARAM;
  std::memset(record, 0, sizeof(*record));
  if (deviceUUID) snprintf(record->deviceUUID, sizeof(record->deviceUUID), "%s", deviceUUID);
  record->computeTokens = computeTokens;
  std::lock_guard<std::mutex> lk(g_mu);
  tfq::QuotaFile* q = self_file();
  if (!q) { record->availableTokens = ~0ull; return ACCEL_SUCCESS; }
  const int idx = device_index_of(q, deviceUUID);
  if (idx < 0) return ACCEL_ERROR_NOT_FOUND;
  double cost = (double)computeTokens;
  const double cap = q->capacity((uint32_t)idx);
  if (cap > 0.0 && cost > cap) cost = cap;  // a launch larger than the bucket is charged one full bucket (see tfprov::clamp_cost)
  const double before = q->fetch_sub((uint32_t)idx, cost);  // soft_limiter_shm.go:715-731
  record->shouldBlock = before < cost;
  const double left = record->shouldBlock ? before : before - cost;
  record->availableTokens = left > 0 ? (uint64_t)left : 0;
  return ACCEL_SUCCESS;
}

// FreezeWorker / ResumeWorker / AutoFreeze / AutoResume (provider/limiter.h:77-81): the same control words
// AccelSnapshot / AccelResume use (worker_ctl.h).  A worker of this stack that publishes a record under
// <shm base>/<namespace>/<pod>/ really moves its vGPU out of HBM (tiered regions -> host tier, plain buffers ->
// pinned host memory) and acknowledges; `state` then carries what the worker reports.  For an id nobody
// publishes under (the reference's stub contract, provider/example/accelerator.c:206-256) the state is only
// remembered here, as the stub does.
static std::string control_base() {
  if (!g_base.empty()) return g_base;
  const char* b = getenv("TF_SHM_BASE_PATH");
  return b && *b ? b : "/run/tensor-fusion/shm";
}

static AccelResult set_frozen(const char* workerId, WorkerFreezeState* state, bool frozen) {
  if (!workerId || !state) return ACCEL_ERROR_INVALID_PARAM;
  std::string base;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    base = control_base();
  }
  const auto recs = tfctl::find_worker(base, workerId);  // file system walk + waiting for the worker: outside the lock
  if (!recs.empty()) {
    std::vector<std::string> files;
    for (const auto& r : recs) files.push_back(r.first);
    const int rc = tfctl::send_control(files, frozen ? TFW_CTL_FREEZE : TFW_CTL_RESUME);
    const auto after = tfctl::find_worker(base, workerId);
    std::memset(state, 0, sizeof *state);
    snprintf(state->workerId, sizeof(state->workerId), "%s", workerId);
    for (const auto& r : after) {
      if (r.second.ctl_frozen) { state->isFrozen = true; state->freezeTimeMs = r.second.frozen_unix_ms ? r.second.frozen_unix_ms : now_ms(); }
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_frozen[workerId] = *state;
    return rc == 0 ? ACCEL_SUCCESS : rc == 4 ? ACCEL_ERROR_RESOURCE_EXHAUSTED : ACCEL_ERROR_OPERATION_FAILED;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  WorkerFreezeState& s = g_frozen[workerId];
  snprintf(s.workerId, sizeof(s.workerId), "%s", workerId);
  if (frozen && !s.isFrozen) s.freezeTimeMs = now_ms();
  if (!frozen) s.freezeTimeMs = 0;
  s.isFrozen = frozen;
  *state = s;
  return ACCEL_SUCCESS;
}
uint32_t LimiterComputeUpLimit(int64_t computePercent, double tflopsLimit, double maxTflops) {
  return tferl::up_limit_percent(computePercent, tflopsLimit, maxTflops);
}

AccelResult FreezeWorker(const char* workerId, WorkerFreezeState* state) { return set_frozen(workerId, state, true); }
AccelResult ResumeWorker(const char* workerId, WorkerFreezeState* state) { return set_frozen(workerId, state, false); }

// The hook calls these when a resource of the worker runs dry / comes back ("compute" or "memory"); the policy
// driven by time (auto_freeze.freeze_to_mem_ttl of RemotePodInfo, api/http_types.go:82-100) lives in the worker
// executable itself, which knows when its client went quiet.
AccelResult AutoFreeze(const char* workerId, const char* deviceUUID, const char* resourceType) {
  if (!workerId || !deviceUUID || !resourceType) return ACCEL_ERROR_INVALID_PARAM;
  WorkerFreezeState s;
  return set_frozen(workerId, &s, true);
}
AccelResult AutoResume(const char* workerId, const char* deviceUUID, const char* resourceType) {
  if (!workerId || !deviceUUID || !resourceType) return ACCEL_ERROR_INVALID_PARAM;
  WorkerFreezeState s;
  return set_frozen(workerId, &s, false);
}

}  // extern "C"
