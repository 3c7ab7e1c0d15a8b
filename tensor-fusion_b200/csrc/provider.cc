// provider.cc -- the accelerator half of the provider C-ABI for NVIDIA B200
// (provider/accelerator.h:275-439), backed by NVML loaded at run time.
//
// What the reference ships for this boundary is a stub with four fake devices
// (provider/example/accelerator.c:272-389).  This implementation reports the
// real devices: the Go hypervisor dlopen()s it through purego and turns the
// answers into GPU CRs, scheduler topology tiers and the ERL controller's
// utilisation input (pkg/hypervisor/device/accelerator.go:408-806,
// quota_controller.go:388-395).  No CUDA context is created here: the
// hypervisor process must stay light, so everything comes from NVML.
#include <dlfcn.h>
#include <nvml.h>
#include <signal.h>
#include <strings.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>

#include "provider_log.h"
#include "shm_quota.h"
#include "tf_provider_abi.h"
#include "tfw_stats_file.h"
#include "worker_ctl.h"

namespace {

std::atomic<LogCallbackFunc> g_log{nullptr};

// ---- NVML, resolved lazily so the library loads on machines without a driver ----
struct Nvml {
  void* so = nullptr;
#define NV(name) decltype(&::name) name = nullptr
  NV(nvmlInit_v2); NV(nvmlShutdown); NV(nvmlErrorString);
  NV(nvmlDeviceGetCount_v2); NV(nvmlDeviceGetHandleByIndex_v2); NV(nvmlDeviceGetUUID); NV(nvmlDeviceGetName);
  NV(nvmlSystemGetDriverVersion); NV(nvmlDeviceGetVbiosVersion); NV(nvmlDeviceGetMinorNumber);
  NV(nvmlDeviceGetMemoryInfo); NV(nvmlDeviceGetNumGpuCores); NV(nvmlDeviceGetCudaComputeCapability);
  NV(nvmlDeviceGetMaxPcieLinkGeneration); NV(nvmlDeviceGetMaxPcieLinkWidth); NV(nvmlDeviceGetMaxClockInfo);
  NV(nvmlDeviceGetPowerManagementLimit); NV(nvmlDeviceGetPowerUsage); NV(nvmlDeviceGetTemperature);
  NV(nvmlDeviceGetUtilizationRates); NV(nvmlDeviceGetPcieThroughput); NV(nvmlDeviceGetComputeRunningProcesses_v3);
  NV(nvmlDeviceGetProcessUtilization); NV(nvmlDeviceGetTopologyCommonAncestor); NV(nvmlDeviceGetP2PStatus);
  NV(nvmlDeviceGetNumaNodeId); NV(nvmlDeviceGetEccMode); NV(nvmlDeviceGetPersistenceMode);
  NV(nvmlDeviceGetClockInfo); NV(nvmlDeviceGetSamples);
#undef NV
  bool load() {
    if (so) return true;
    for (const char* n : {"libnvidia-ml.so.1", "libnvidia-ml.so"}) {
      so = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (so) break;
    }
    if (!so) return false;
#define NV(name) name = reinterpret_cast<decltype(name)>(dlsym(so, #name))
    NV(nvmlInit_v2); NV(nvmlShutdown); NV(nvmlErrorString);
    NV(nvmlDeviceGetCount_v2); NV(nvmlDeviceGetHandleByIndex_v2); NV(nvmlDeviceGetUUID); NV(nvmlDeviceGetName);
    NV(nvmlSystemGetDriverVersion); NV(nvmlDeviceGetVbiosVersion); NV(nvmlDeviceGetMinorNumber);
    NV(nvmlDeviceGetMemoryInfo); NV(nvmlDeviceGetNumGpuCores); NV(nvmlDeviceGetCudaComputeCapability);
    NV(nvmlDeviceGetMaxPcieLinkGeneration); NV(nvmlDeviceGetMaxPcieLinkWidth); NV(nvmlDeviceGetMaxClockInfo);
    NV(nvmlDeviceGetPowerManagementLimit); NV(nvmlDeviceGetPowerUsage); NV(nvmlDeviceGetTemperature);
    NV(nvmlDeviceGetUtilizationRates); NV(nvmlDeviceGetPcieThroughput); NV(nvmlDeviceGetComputeRunningProcesses_v3);
    NV(nvmlDeviceGetProcessUtilization); NV(nvmlDeviceGetTopologyCommonAncestor); NV(nvmlDeviceGetP2PStatus);
    NV(nvmlDeviceGetNumaNodeId); NV(nvmlDeviceGetEccMode); NV(nvmlDeviceGetPersistenceMode);
    NV(nvmlDeviceGetClockInfo); NV(nvmlDeviceGetSamples);
#undef NV
    return nvmlInit_v2 && nvmlDeviceGetCount_v2 && nvmlDeviceGetHandleByIndex_v2 && nvmlDeviceGetUUID &&
           nvmlDeviceGetName && nvmlDeviceGetMemoryInfo;
  }
};

struct Dev {
  nvmlDevice_t h{};
  std::string uuid, model;
  unsigned sms = 0;
  uint64_t mem = 0;
};

// nvmlDeviceGetPcieThroughput integrates over 20 ms per call: two directions x 8 GPUs would add
// 320 ms to every AccelGetDeviceMetrics, which the hypervisor calls from its 500 ms ERL tick
// (quota_controller.go:388-395).  A sampler thread (started by AccelInit, never by a library
// constructor) refreshes the counters every 2 s; the ABI call only reads the cache.
struct PcieCache { std::atomic<uint64_t> rx{0}, tx{0}; };
PcieCache g_pcie[MAX_TOPOLOGY_DEVICES];
std::atomic<bool> g_pcie_stop{false};
// joined on AccelShutdown, and by this object's destructor if the host exits without calling it
// (a joinable std::thread destroyed at exit would call std::terminate)
struct SamplerThread {
  std::thread t;
  bool joinable() const { return t.joinable(); }
  void stop() {
    if (t.joinable()) {
      g_pcie_stop.store(true, std::memory_order_release);
      t.join();
    }
  }
  ~SamplerThread() { stop(); }
} g_pcie_thread;

std::mutex g_mu;
Nvml g_nv;
bool g_inited = false;
std::vector<Dev> g_devs;
std::map<std::string, int> g_alias;                    // TF_PROVIDER_DEVICE_ALIASES="name=index,..."
std::map<std::string, uint64_t> g_mem_hard;            // AccelSetMemHardLimit, keyed by canonical uuid
std::map<std::string, uint32_t> g_cu_hard;             // AccelSetComputeUnitHardLimit
std::map<std::string, std::string> g_partitions;       // partition uuid -> device uuid
unsigned g_partition_seq = 0;
unsigned long long g_last_util_ts[64] = {0};
unsigned long long g_last_gpu_sample_ts[64] = {0};  // AccelGetDeviceMetrics: newest GPU-utilisation sample already averaged

void put(char* dst, size_t cap, const std::string& s) { snprintf(dst, cap, "%s", s.c_str()); }

// Sum of the counters every live vGPU worker published for `uuid` under
// <base>/<namespace>/<pod>/tfw_stats (include/tfw_stats_file.h).
struct WorkerTotals { uint64_t workers = 0, payload = 0, h2d = 0, d2h = 0, movers = 0, launches = 0, throttled = 0, timeouts = 0, vram = 0, frozen = 0, parked = 0; };

std::string shm_base() {
  const char* b = getenv("TF_SHM_BASE_PATH");
  return tfprov::limiter_base().empty() ? std::string(b && *b ? b : "/run/tensor-fusion/shm") : tfprov::limiter_base();
}

// One walk over <base>/<namespace>/<pod>/ per metrics call (the hypervisor asks for all GPUs at once, every
// 500 ms): totals keyed by the upper-cased device UUID.
std::string upper(std::string s) {
  for (auto& c : s) c = (char)toupper((unsigned char)c);
  return s;
}
std::map<std::string, WorkerTotals> collect_worker_stats(const std::string& base) {
  std::map<std::string, WorkerTotals> all;
  tfctl::for_each_worker_record(base, [&](const std::string&, const std::string&, const tfw_stats_record& r) {
    WorkerTotals& t = all[upper(r.device_uuid)];
    t.workers++; t.payload += r.payload_bytes; t.h2d += r.h2d_dma_bytes; t.d2h += r.d2h_bytes; t.movers += r.mover_launches;
    t.launches += r.client_launches; t.throttled += r.gate_blocked; t.timeouts += r.gate_timeouts; t.vram += r.vram_bytes;
    t.frozen += r.ctl_frozen ? 1 : 0; t.parked += r.parked_bytes;
  });
  return all;
}

// dense bf16/fp16 TFLOPS by model (charts/tensor-fusion/templates/gpu-public-gpu-info.yaml:380-385
// lists B200 at 2250; the scheduler divides tflops limits by this number)
double model_tflops(const std::string& model, unsigned sms, unsigned max_sm_mhz) {
  struct { const char* key; double tf; } table[] = {{"B200", 2250}, {"B300", 2250}, {"GB200", 2500}, {"H200", 989},
                                                    {"H100", 989},  {"H20", 148},   {"A100", 312},   {"L40S", 362}};
  for (auto& e : table) if (model.find(e.key) != std::string::npos) return e.tf;
  // unknown part: tensor throughput ~ 1024 dense fp16 FMA/clk/SM on recent parts
  return (double)sms * 2048.0 * (double)(max_sm_mhz ? max_sm_mhz : 1500) * 1e6 / 1e12;
}

// returns index or -1; accepts native case, any case, with or without "GPU-", and aliases
int find_dev(const char* uuid) {
  if (!uuid || !*uuid) return -1;
  for (size_t i = 0; i < g_devs.size(); ++i)
    if (strcasecmp(g_devs[i].uuid.c_str(), uuid) == 0) return (int)i;
  for (size_t i = 0; i < g_devs.size(); ++i)
    if (g_devs[i].uuid.size() > 4 && strcasecmp(g_devs[i].uuid.c_str() + 4, uuid) == 0) return (int)i;
  auto a = g_alias.find(uuid);
  if (a != g_alias.end() && a->second >= 0 && (size_t)a->second < g_devs.size()) return a->second;
  auto p = g_partitions.find(uuid);
  if (p != g_partitions.end()) return find_dev(p->second.c_str());
  return -1;
}

void add_prop(DeviceProperties* p, const char* k, const std::string& v) {
  if (p->count >= MAX_DEVICE_PROPERTIES) return;
  put(p->properties[p->count].key, sizeof(p->properties[0].key), k);
  put(p->properties[p->count].value, sizeof(p->properties[0].value), v);
  p->count++;
}

bool refresh_devices() {  // caller holds g_mu, NVML initialised
  unsigned n = 0;
  if (g_nv.nvmlDeviceGetCount_v2(&n) != NVML_SUCCESS) return false;
  std::vector<Dev> devs;
  for (unsigned i = 0; i < n && i < MAX_TOPOLOGY_DEVICES; ++i) {
    Dev d;
    if (g_nv.nvmlDeviceGetHandleByIndex_v2(i, &d.h) != NVML_SUCCESS) continue;
    char buf[NVML_DEVICE_UUID_V2_BUFFER_SIZE] = {0};
    if (g_nv.nvmlDeviceGetUUID(d.h, buf, sizeof(buf)) != NVML_SUCCESS) continue;
    d.uuid = buf;
    char name[NVML_DEVICE_NAME_V2_BUFFER_SIZE] = {0};
    if (g_nv.nvmlDeviceGetName(d.h, name, sizeof(name)) == NVML_SUCCESS) d.model = name;
    nvmlMemory_t m{};
    if (g_nv.nvmlDeviceGetMemoryInfo(d.h, &m) == NVML_SUCCESS) d.mem = m.total;
    unsigned cores = 0;
    if (g_nv.nvmlDeviceGetNumGpuCores && g_nv.nvmlDeviceGetNumGpuCores(d.h, &cores) == NVML_SUCCESS && cores) d.sms = cores / 128;
    if (!d.sms) d.sms = d.model.find("B200") != std::string::npos ? 148 : 132;
    devs.push_back(d);
  }
  g_devs.swap(devs);
  return true;
}

}  // namespace

namespace tfprov {
void log(const char* level, const char* msg) {
  LogCallbackFunc f = g_log.load(std::memory_order_acquire);
  if (f) f(level, msg);
}
}  // namespace tfprov

extern "C" {

AccelResult AccelRegisterLogCallback(LogCallbackFunc callback) {
  g_log.store(callback, std::memory_order_release);
  return ACCEL_SUCCESS;
}
// accelerator_unix.go:102-106 falls back to this legacy name
TF_ABI_EXPORT AccelResult RegisterLogCallback(LogCallbackFunc callback) { return AccelRegisterLogCallback(callback); }

AccelResult AccelInit(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_inited) return ACCEL_SUCCESS;
  if (!g_nv.load()) {
    tfprov::log("ERROR", "libaccelerator_b200: libnvidia-ml.so.1 not found; no NVIDIA driver on this node");
    return ACCEL_ERROR_OPERATION_FAILED;
  }
  nvmlReturn_t r = g_nv.nvmlInit_v2();
  if (r != NVML_SUCCESS) {
    std::string m = std::string("libaccelerator_b200: nvmlInit failed: ") + (g_nv.nvmlErrorString ? g_nv.nvmlErrorString(r) : "?");
    tfprov::log("ERROR", m.c_str());
    return ACCEL_ERROR_OPERATION_FAILED;
  }
  g_alias.clear();
  if (const char* a = getenv("TF_PROVIDER_DEVICE_ALIASES")) {
    std::string s(a);
    size_t i = 0;
    while (i < s.size()) {
      size_t j = s.find(',', i);
      if (j == std::string::npos) j = s.size();
      const std::string kv = s.substr(i, j - i);
      const size_t e = kv.find('=');
      if (e != std::string::npos) g_alias[kv.substr(0, e)] = atoi(kv.c_str() + e + 1);
      i = j + 1;
    }
  }
  if (!refresh_devices()) { g_nv.nvmlShutdown(); return ACCEL_ERROR_OPERATION_FAILED; }
  g_inited = true;
  if (g_nv.nvmlDeviceGetPcieThroughput && !g_pcie_thread.joinable()) {
    g_pcie_stop.store(false);
    std::vector<nvmlDevice_t> handles;
    for (const Dev& d : g_devs) handles.push_back(d.h);
    g_pcie_thread.t = std::thread([handles] {
      while (!g_pcie_stop.load(std::memory_order_acquire)) {
        for (size_t i = 0; i < handles.size() && i < MAX_TOPOLOGY_DEVICES && !g_pcie_stop.load(); ++i) {
          unsigned v = 0;
          if (g_nv.nvmlDeviceGetPcieThroughput(handles[i], NVML_PCIE_UTIL_RX_BYTES, &v) == NVML_SUCCESS) g_pcie[i].rx.store((uint64_t)v * 1024);
          if (g_nv.nvmlDeviceGetPcieThroughput(handles[i], NVML_PCIE_UTIL_TX_BYTES, &v) == NVML_SUCCESS) g_pcie[i].tx.store((uint64_t)v * 1024);
        }
        for (int k = 0; k < 20 && !g_pcie_stop.load(std::memory_order_acquire); ++k) usleep(100000);
      }
    });
  }
  char msg[128];
  snprintf(msg, sizeof msg, "libaccelerator_b200: NVML up, %zu device(s)", g_devs.size());
  tfprov::log("INFO", msg);
  return ACCEL_SUCCESS;
}

AccelResult AccelShutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_pcie_thread.stop();
  if (g_inited) {
    g_nv.nvmlShutdown();
    g_inited = false;
    g_devs.clear();
  }
  return ACCEL_SUCCESS;
}

// The reference's own test calls the API after AccelShutdown and expects it to
// work (provider/test/test_accelerator.c:40-49 then :52 onwards); the header
// says use-before-init "will trigger a TF_PANIC" but no such macro exists
// (SURVEY.md App. E-2).  We re-initialise on demand instead of panicking.
static AccelResult ensure_init() {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited) return ACCEL_SUCCESS;
  }
  return AccelInit();
}

AccelResult AccelGetDeviceCount(size_t* deviceCount) {
  if (!deviceCount) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  std::lock_guard<std::mutex> lk(g_mu);
  refresh_devices();
  *deviceCount = g_devs.size();
  return ACCEL_SUCCESS;
}

AccelResult AccelGetAllDevices(ExtendedDeviceInfo* devices, size_t maxCount, size_t* deviceCount) {
  if (!devices || !deviceCount || maxCount == 0) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  std::lock_guard<std::mutex> lk(g_mu);
  char driver[NVML_SYSTEM_DRIVER_VERSION_BUFFER_SIZE] = {0};
  if (g_nv.nvmlSystemGetDriverVersion) g_nv.nvmlSystemGetDriverVersion(driver, sizeof(driver));
  const size_t n = g_devs.size() < maxCount ? g_devs.size() : maxCount;
  for (size_t i = 0; i < n; ++i) {
    const Dev& d = g_devs[i];
    ExtendedDeviceInfo* e = &devices[i];
    std::memset(e, 0, sizeof(*e));
    put(e->basic.uuid, sizeof(e->basic.uuid), d.uuid);
    put(e->basic.vendor, sizeof(e->basic.vendor), "NVIDIA");  // quota_controller.go:199-202 requires exactly this vendor
    put(e->basic.model, sizeof(e->basic.model), d.model);
    put(e->basic.driverVersion, sizeof(e->basic.driverVersion), driver);
    char vbios[NVML_DEVICE_VBIOS_VERSION_BUFFER_SIZE] = {0};
    if (g_nv.nvmlDeviceGetVbiosVersion) g_nv.nvmlDeviceGetVbiosVersion(d.h, vbios, sizeof(vbios));
    put(e->basic.firmwareVersion, sizeof(e->basic.firmwareVersion), vbios);
    unsigned minor = (unsigned)i;
    if (g_nv.nvmlDeviceGetMinorNumber) g_nv.nvmlDeviceGetMinorNumber(d.h, &minor);
    put(e->basic.deviceNode, sizeof(e->basic.deviceNode), "/dev/nvidia" + std::to_string(minor));
    e->basic.index = (int32_t)i;
    e->basic.numaNode = -1;
    unsigned numa = 0;
    if (g_nv.nvmlDeviceGetNumaNodeId && g_nv.nvmlDeviceGetNumaNodeId(d.h, &numa) == NVML_SUCCESS) e->basic.numaNode = (int32_t)numa;
    e->basic.totalMemoryBytes = d.mem;
    e->basic.totalComputeUnits = d.sms;
    unsigned max_sm = 0, max_mem = 0, gen = 0, width = 0, plimit = 0;
    if (g_nv.nvmlDeviceGetMaxClockInfo) { g_nv.nvmlDeviceGetMaxClockInfo(d.h, NVML_CLOCK_SM, &max_sm); g_nv.nvmlDeviceGetMaxClockInfo(d.h, NVML_CLOCK_MEM, &max_mem); }
    e->basic.maxTflops = model_tflops(d.model, d.sms, max_sm);
    if (g_nv.nvmlDeviceGetMaxPcieLinkGeneration) g_nv.nvmlDeviceGetMaxPcieLinkGeneration(d.h, &gen);
    if (g_nv.nvmlDeviceGetMaxPcieLinkWidth) g_nv.nvmlDeviceGetMaxPcieLinkWidth(d.h, &width);
    e->basic.pcieGen = gen;
    e->basic.pcieWidth = width;
    int cc_major = 0, cc_minor = 0;
    if (g_nv.nvmlDeviceGetCudaComputeCapability) g_nv.nvmlDeviceGetCudaComputeCapability(d.h, &cc_major, &cc_minor);
    if (g_nv.nvmlDeviceGetPowerManagementLimit) g_nv.nvmlDeviceGetPowerManagementLimit(d.h, &plimit);
    add_prop(&e->props, "clockSM", std::to_string(max_sm));
    add_prop(&e->props, "clockMem", std::to_string(max_mem));
    add_prop(&e->props, "powerLimit", std::to_string(plimit / 1000));
    // "computeCapability" drives coresPerSM in handlers/legacy.go:690-713 (10.x -> 128)
    add_prop(&e->props, "computeCapability", std::to_string(cc_major) + "." + std::to_string(cc_minor));
    // "totalComputeUnits" is read back as SM count in worker/controller.go:545-549
    add_prop(&e->props, "totalComputeUnits", std::to_string(d.sms));
    add_prop(&e->props, "chipType", "NVIDIA");
    add_prop(&e->props, "interconnect", "NVLink5/NVSwitch");
    nvmlEnableState_t ecc_cur = NVML_FEATURE_DISABLED, ecc_pend = NVML_FEATURE_DISABLED;
    if (g_nv.nvmlDeviceGetEccMode && g_nv.nvmlDeviceGetEccMode(d.h, &ecc_cur, &ecc_pend) == NVML_SUCCESS)
      add_prop(&e->props, "eccEnabled", ecc_cur == NVML_FEATURE_ENABLED ? "true" : "false");
    VirtualizationCapabilities& v = e->virtualizationCapabilities;
    v.supportsPartitioning = true;   // template partitions enforced by the worker's hard limits (see AccelAssignPartition)
    v.supportsSoftIsolation = true;  // quota file + device-resident token bucket
    v.supportsHardIsolation = true;  // TF_CUDA_MEMORY_LIMIT / TF_CUDA_SM_PERCENT_LIMIT
    v.supportsSnapshot = true;   // AccelSnapshot/AccelResume freeze and thaw this stack's vGPU workers
    v.supportsMetrics = true;
    v.supportsRemoting = true;       // the TFCS worker
    v.maxPartitions = 7;
    v.maxWorkersPerDevice = 16;      // one quota-file slot per worker device index (soft_limiter_shm.go:21)
  }
  *deviceCount = n;
  return ACCEL_SUCCESS;
}

AccelResult AccelGetAllDevicesTopology(ExtendedDeviceTopology* topology) {
  if (!topology) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  std::lock_guard<std::mutex> lk(g_mu);
  const size_t n = g_devs.size() < MAX_TOPOLOGY_DEVICES ? g_devs.size() : MAX_TOPOLOGY_DEVICES;
  topology->deviceCount = n;  // written at offset 300 032: the in-tree header, not the larger Go mirror (App. E-1)
  for (size_t i = 0; i < n; ++i) {
    DeviceTopologyInfo* t = &topology->devices[i];
    std::memset(t, 0, sizeof(*t));
    put(t->deviceUUID, sizeof(t->deviceUUID), g_devs[i].uuid);
    t->deviceIndex = (int32_t)i;
    t->numaNode = -1;
    unsigned numa = 0;
    if (g_nv.nvmlDeviceGetNumaNodeId && g_nv.nvmlDeviceGetNumaNodeId(g_devs[i].h, &numa) == NVML_SUCCESS) t->numaNode = (int32_t)numa;
    for (size_t j = 0; j < n; ++j) {
      if (i == j) continue;
      DeviceTopoNode* p = &t->peers[t->peerCount++];
      put(p->peerUUID, sizeof(p->peerUUID), g_devs[j].uuid);
      p->peerIndex = (int32_t)j;
      p->topoLevel = TOPO_LEVEL_UNKNOWN;
      nvmlGpuP2PStatus_t st = NVML_P2P_STATUS_UNKNOWN;
      if (g_nv.nvmlDeviceGetP2PStatus && g_nv.nvmlDeviceGetP2PStatus(g_devs[i].h, g_devs[j].h, NVML_P2P_CAPS_INDEX_NVLINK, &st) == NVML_SUCCESS &&
          st == NVML_P2P_STATUS_OK) {
        p->topoLevel = TOPO_LEVEL_INTERNAL;  // NVSwitch: every peer is tier 0 (accelerator.go:323-334)
        continue;
      }
      nvmlGpuTopologyLevel_t lvl;
      if (g_nv.nvmlDeviceGetTopologyCommonAncestor && g_nv.nvmlDeviceGetTopologyCommonAncestor(g_devs[i].h, g_devs[j].h, &lvl) == NVML_SUCCESS) {
        switch (lvl) {
          case NVML_TOPOLOGY_INTERNAL: p->topoLevel = TOPO_LEVEL_INTERNAL; break;
          case NVML_TOPOLOGY_SINGLE: p->topoLevel = TOPO_LEVEL_SINGLE_SWITCH; break;
          case NVML_TOPOLOGY_MULTIPLE: p->topoLevel = TOPO_LEVEL_MULTI_SWITCH; break;
          case NVML_TOPOLOGY_HOSTBRIDGE: p->topoLevel = TOPO_LEVEL_HOST_BRIDGE; break;
          case NVML_TOPOLOGY_NODE: p->topoLevel = TOPO_LEVEL_NUMA_NODE; break;
          default: p->topoLevel = TOPO_LEVEL_SYSTEM; break;
        }
      }
    }
  }
  return ACCEL_SUCCESS;
}

// "<N>g.<M>gb" (MIG-style template ids as the reference's tests use, e.g. "1g.10gb"):
// N sevenths of the SMs, M GiB of memory.
static bool parse_template(const char* id, unsigned* sevenths, uint64_t* bytes) {
  unsigned g = 0, gb = 0;
  if (sscanf(id, "%ug.%ugb", &g, &gb) != 2 || g == 0 || g > 7 || gb == 0) return false;
  *sevenths = g;
  *bytes = (uint64_t)gb << 30;
  return true;
}

AccelResult AccelAssignPartition(const char* templateId, const char* deviceUUID, PartitionResult* partitionResult) {
  if (!templateId || !deviceUUID || !partitionResult || !*templateId || !*deviceUUID) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  std::lock_guard<std::mutex> lk(g_mu);
  const int di = find_dev(deviceUUID);
  if (di < 0) return ACCEL_ERROR_NOT_FOUND;
  unsigned sevenths = 0;
  uint64_t bytes = 0;
  if (!parse_template(templateId, &sevenths, &bytes)) return ACCEL_ERROR_INVALID_PARAM;
  if (bytes > g_devs[di].mem) return ACCEL_ERROR_RESOURCE_EXHAUSTED;
  std::memset(partitionResult, 0, sizeof(*partitionResult));
  partitionResult->type = PARTITION_TYPE_ENVIRONMENT_VARIABLE;
  char puuid[64];
  snprintf(puuid, sizeof puuid, "%.40s-p%u", g_devs[di].uuid.c_str(), g_partition_seq++);
  put(partitionResult->deviceUUID, sizeof(partitionResult->deviceUUID), puuid);
  g_partitions[puuid] = g_devs[di].uuid;
  // The partition is enforced by the worker's hard limits (internal/utils/compose.go:1287-1295)
  snprintf(partitionResult->envVars[0], MAX_ENV_VALUE_LENGTH, "NVIDIA_VISIBLE_DEVICES=%s", g_devs[di].uuid.c_str());
  snprintf(partitionResult->envVars[1], MAX_ENV_VALUE_LENGTH, "TF_CUDA_SM_PERCENT_LIMIT=%u", (sevenths * 100 + 6) / 7);
  snprintf(partitionResult->envVars[2], MAX_ENV_VALUE_LENGTH, "TF_CUDA_MEMORY_LIMIT=%llu", (unsigned long long)(bytes >> 20));
  snprintf(partitionResult->envVars[3], MAX_ENV_VALUE_LENGTH, "TF_PARTITION_TEMPLATE=%s", templateId);
  return ACCEL_SUCCESS;
}

AccelResult AccelRemovePartition(const char* templateId, const char* deviceUUID) {
  if (!templateId || !deviceUUID) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_partitions.find(deviceUUID);
  if (it != g_partitions.end()) { g_partitions.erase(it); return ACCEL_SUCCESS; }
  return find_dev(deviceUUID) >= 0 ? ACCEL_SUCCESS : ACCEL_ERROR_NOT_FOUND;  // removing an absent partition is idempotent
}

// The hard limits are enforced where the tenant's work runs: every live vGPU worker on the device gets the new
// limit through the control words of its stats record (worker_ctl.h) -- the memory limit becomes its MALLOC quota,
// the compute limit moves its stream into an SM partition of that size (green context).  Workers started later
// get theirs from the operator's environment (TF_CUDA_MEMORY_LIMIT / TF_CUDA_SM_PERCENT_LIMIT, compose.go:1287-1295).
static AccelResult apply_to_workers(const std::string& uuid, uint32_t cmd, uint64_t arg) {
  std::vector<std::string> files;
  tfctl::for_each_worker_record(shm_base(), [&](const std::string&, const std::string& file, const tfw_stats_record& r) {
    if (strcasecmp(r.device_uuid, uuid.c_str()) == 0) files.push_back(file);
  });
  if (files.empty()) return ACCEL_SUCCESS;  // nobody to tell yet
  switch (tfctl::send_control(files, cmd, arg)) {
    case 0: return ACCEL_SUCCESS;
    case 3: return ACCEL_ERROR_NOT_SUPPORTED;
    default: return ACCEL_ERROR_OPERATION_FAILED;
  }
}

AccelResult AccelSetMemHardLimit(const char* deviceUUID, uint64_t memoryLimitBytes) {
  if (!deviceUUID || memoryLimitBytes == 0) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  std::unique_lock<std::mutex> lk(g_mu);
  const int di = find_dev(deviceUUID);
  if (di < 0) return ACCEL_ERROR_NOT_FOUND;
  g_mem_hard[g_devs[di].uuid] = memoryLimitBytes;
  const std::string uuid = g_devs[di].uuid;
  lk.unlock();
  return apply_to_workers(uuid, TFW_CTL_MEM_LIMIT, memoryLimitBytes);
}

AccelResult AccelSetComputeUnitHardLimit(const char* deviceUUID, uint32_t computeUnitLimit) {
  if (!deviceUUID || computeUnitLimit == 0 || computeUnitLimit > 100) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  std::unique_lock<std::mutex> lk(g_mu);
  const int di = find_dev(deviceUUID);
  if (di < 0) return ACCEL_ERROR_NOT_FOUND;
  g_cu_hard[g_devs[di].uuid] = computeUnitLimit;
  const std::string uuid = g_devs[di].uuid;
  lk.unlock();
  return apply_to_workers(uuid, TFW_CTL_SM_LIMIT, computeUnitLimit);
}


static AccelResult check_snapshot_ctx(SnapshotContext* c) {
  if (!c) return ACCEL_ERROR_INVALID_PARAM;
  if (c->processIds && c->processCount > 0) {
    if (c->processCount > MAX_PROCESSES) return ACCEL_ERROR_INVALID_PARAM;
    for (size_t i = 0; i < c->processCount; ++i)
      if (kill(c->processIds[i], 0) != 0) return ACCEL_ERROR_NOT_FOUND;
    return ACCEL_SUCCESS;
  }
  if (c->deviceUUID) return ACCEL_SUCCESS;
  return ACCEL_ERROR_INVALID_PARAM;
}
// Snapshot / resume.  The reference leaves the transport open ("send snapshot command
// to worker via shared memory", handlers/worker.go:94-129); here it is the control words of the worker's
// stats record (include/tfw_stats_file.h): the provider writes a request, the worker freezes its vGPU
// (every byte of the tenant leaves HBM: tiered regions -> host tier, plain buffers -> host memory) and
// acknowledges.  Process-level: a PID names a worker either directly (record.pid, same PID namespace) or
// through the pod's quota file, whose PID set holds the host PIDs the hypervisor registered (legacy.go:576).
static AccelResult control_workers(SnapshotContext* c, uint32_t cmd) {
  const bool by_pid = c->processIds && c->processCount > 0;
  std::string want_uuid;
  if (!by_pid) {
    std::lock_guard<std::mutex> lk(g_mu);
    const int di = find_dev(c->deviceUUID);
    want_uuid = di >= 0 ? g_devs[(size_t)di].uuid : std::string(c->deviceUUID);
  }
  std::vector<std::string> files;
  tfctl::for_each_worker_record(shm_base(), [&](const std::string& poddir, const std::string& file, const tfw_stats_record& r) {
    bool match = false;
    if (by_pid) {
      for (size_t i = 0; i < c->processCount && !match; ++i) match = r.pid == (uint64_t)c->processIds[i];
      if (!match) {
        tfq::QuotaFile* q = nullptr;
        std::string err;
        if (tfq::QuotaFile::open_file(poddir + "/shm", &q, &err) == tfq::kOk) {
          const std::vector<uint64_t> pids = q->pids();
          for (size_t i = 0; i < c->processCount && !match; ++i)
            for (uint64_t p : pids) if (p == (uint64_t)c->processIds[i]) { match = true; break; }
          delete q;
        }
      }
    } else {
      match = strcasecmp(r.device_uuid, want_uuid.c_str()) == 0;
    }
    if (match) files.push_back(file);
  });
  if (files.empty()) return by_pid ? ACCEL_ERROR_NOT_SUPPORTED : ACCEL_SUCCESS;  // not vGPU workers of this stack / idle device
  switch (tfctl::send_control(files, cmd)) {
    case 0: return ACCEL_SUCCESS;
    case 4: return ACCEL_ERROR_RESOURCE_EXHAUSTED;
    default: return ACCEL_ERROR_OPERATION_FAILED;
  }
}

AccelResult AccelSnapshot(SnapshotContext* context) {
  AccelResult r = check_snapshot_ctx(context);
  return r == ACCEL_SUCCESS ? control_workers(context, TFW_CTL_FREEZE) : r;
}
AccelResult AccelResume(SnapshotContext* context) {
  AccelResult r = check_snapshot_ctx(context);
  return r == ACCEL_SUCCESS ? control_workers(context, TFW_CTL_RESUME) : r;
}

AccelResult AccelGetProcessInformation(ProcessInformation* processInfos, size_t maxCount, size_t* processInfoCount) {
  if (!processInfos || !processInfoCount || maxCount == 0) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  std::lock_guard<std::mutex> lk(g_mu);
  size_t out = 0;
  for (size_t i = 0; i < g_devs.size() && out < maxCount; ++i) {
    const Dev& d = g_devs[i];
    unsigned cnt = 0;
    if (!g_nv.nvmlDeviceGetComputeRunningProcesses_v3) break;
    nvmlReturn_t rr = g_nv.nvmlDeviceGetComputeRunningProcesses_v3(d.h, &cnt, nullptr);
    if (rr != NVML_SUCCESS && rr != NVML_ERROR_INSUFFICIENT_SIZE) continue;
    if (cnt == 0) continue;
    std::vector<nvmlProcessInfo_t> procs(cnt + 8);
    cnt = (unsigned)procs.size();
    if (g_nv.nvmlDeviceGetComputeRunningProcesses_v3(d.h, &cnt, procs.data()) != NVML_SUCCESS) continue;
    // per-process SM utilisation samples since the previous query
    std::map<unsigned, unsigned> sm_util;
    std::map<unsigned, std::pair<unsigned long long, unsigned>> sm_sum;  // pid -> (sum, samples): the mean over the polling interval
    if (g_nv.nvmlDeviceGetProcessUtilization) {
      unsigned ns = 0;
      nvmlReturn_t ur = g_nv.nvmlDeviceGetProcessUtilization(d.h, nullptr, &ns, g_last_util_ts[i]);
      if (ur == NVML_ERROR_INSUFFICIENT_SIZE && ns) {
        std::vector<nvmlProcessUtilizationSample_t> s(ns);
        if (g_nv.nvmlDeviceGetProcessUtilization(d.h, s.data(), &ns, g_last_util_ts[i]) == NVML_SUCCESS) {
          for (unsigned k = 0; k < ns; ++k) {
            auto& acc = sm_sum[s[k].pid];
            acc.first += s[k].smUtil;
            acc.second++;
            if (s[k].timeStamp > g_last_util_ts[i]) g_last_util_ts[i] = s[k].timeStamp;
          }
          for (auto& kv : sm_sum) sm_util[kv.first] = (unsigned)((kv.second.first + kv.second.second / 2) / kv.second.second);
        }
      }
    }
    for (unsigned k = 0; k < cnt && out < maxCount; ++k) {
      ProcessInformation* p = &processInfos[out++];
      std::memset(p, 0, sizeof(*p));
      snprintf(p->processId, sizeof(p->processId), "%u", procs[k].pid);
      put(p->deviceUUID, sizeof(p->deviceUUID), d.uuid);
      const unsigned u = sm_util.count(procs[k].pid) ? sm_util[procs[k].pid] : 0;
      p->computeUtilizationPercent = u > 100 ? 100.0 : (double)u;
      p->totalSMs = d.sms;
      p->activeSMs = (uint64_t)((double)d.sms * p->computeUtilizationPercent / 100.0);
      p->memoryUsedBytes = procs[k].usedGpuMemory == (unsigned long long)NVML_VALUE_NOT_AVAILABLE ? 0 : procs[k].usedGpuMemory;
      p->memoryReservedBytes = 0;
      p->memoryUtilizationPercent = d.mem ? (double)p->memoryUsedBytes / (double)d.mem * 100.0 : 0.0;
    }
  }
  *processInfoCount = out;
  return ACCEL_SUCCESS;
}

AccelResult AccelGetDeviceMetrics(const char** deviceUUIDs, size_t deviceCount, DeviceMetrics* metrics) {
  if (!deviceUUIDs || deviceCount == 0 || !metrics) return ACCEL_ERROR_INVALID_PARAM;
  AccelResult r = ensure_init();
  if (r != ACCEL_SUCCESS) return r;
  const std::map<std::string, WorkerTotals> workers = collect_worker_stats(shm_base());  // file system walk: outside the lock
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t i = 0; i < deviceCount; ++i) {
    DeviceMetrics* m = &metrics[i];
    std::memset(m, 0, sizeof(*m));
    if (deviceUUIDs[i]) snprintf(m->deviceUUID, sizeof(m->deviceUUID), "%s", deviceUUIDs[i]);
    const int di = find_dev(deviceUUIDs[i]);
    if (di < 0) continue;  // unknown device: zeroed row, same shape as the reference's fallback
    const Dev& d = g_devs[di];
    unsigned mw = 0, temp = 0, rx = 0, tx = 0, smclk = 0;
    if (g_nv.nvmlDeviceGetPowerUsage && g_nv.nvmlDeviceGetPowerUsage(d.h, &mw) == NVML_SUCCESS) m->powerUsageWatts = mw / 1000.0;
    if (g_nv.nvmlDeviceGetTemperature && g_nv.nvmlDeviceGetTemperature(d.h, NVML_TEMPERATURE_GPU, &temp) == NVML_SUCCESS) m->temperatureCelsius = temp;
    (void)rx; (void)tx;
    m->pcieRxBytes = g_pcie[di].rx.load(std::memory_order_relaxed);  // bytes/s, refreshed by the sampler thread
    m->pcieTxBytes = g_pcie[di].tx.load(std::memory_order_relaxed);
    nvmlUtilization_t u{};
    if (g_nv.nvmlDeviceGetUtilizationRates && g_nv.nvmlDeviceGetUtilizationRates(d.h, &u) == NVML_SUCCESS) m->utilizationPercent = u.gpu;
    // nvmlDeviceGetUtilizationRates is ONE sample of the driver's utilisation counter (the latest ~1/6 s window): with
    // tenants that run in bursts a 2 Hz reader sees 0 % or 99 % at random, and the ERL controller chases the noise.  The
    // driver keeps a buffer of those samples: report the mean of the ones taken since the previous call, i.e. the
    // utilisation over the caller's own polling interval.  (TF_UTIL_SINGLE_SAMPLE=1: the single latest sample.)
    unsigned util_samples = 0;
    if (g_nv.nvmlDeviceGetSamples && di < 64 && !getenv("TF_UTIL_SINGLE_SAMPLE")) {
      nvmlValueType_t vt;
      unsigned cnt = 0;
      if (g_nv.nvmlDeviceGetSamples(d.h, NVML_GPU_UTILIZATION_SAMPLES, g_last_gpu_sample_ts[di], &vt, &cnt, nullptr) == NVML_SUCCESS && cnt) {
        std::vector<nvmlSample_t> sm(cnt);
        if (g_nv.nvmlDeviceGetSamples(d.h, NVML_GPU_UTILIZATION_SAMPLES, g_last_gpu_sample_ts[di], &vt, &cnt, sm.data()) == NVML_SUCCESS && cnt) {
          double sum = 0;
          const bool first = g_last_gpu_sample_ts[di] == 0;  // the whole buffer (many seconds): only its timestamps are of use
          for (unsigned q = 0; q < cnt; ++q) {
            const double v = vt == NVML_VALUE_TYPE_DOUBLE ? sm[q].sampleValue.dVal : vt == NVML_VALUE_TYPE_UNSIGNED_LONG ? (double)sm[q].sampleValue.ulVal
                           : vt == NVML_VALUE_TYPE_UNSIGNED_LONG_LONG ? (double)sm[q].sampleValue.ullVal : (double)sm[q].sampleValue.uiVal;
            sum += v;
            if (sm[q].timeStamp > g_last_gpu_sample_ts[di]) g_last_gpu_sample_ts[di] = sm[q].timeStamp;
          }
          if (!first) { m->utilizationPercent = std::min(100.0, sum / cnt); util_samples = cnt; }
        }
      }
    }
    nvmlMemory_t mem{};
    if (g_nv.nvmlDeviceGetMemoryInfo(d.h, &mem) == NVML_SUCCESS) m->memoryUsedBytes = mem.used;
    size_t k = 0;
    auto extra = [&](const char* key, double v) {
      if (k >= MAX_EXTRA_METRICS) return;
      snprintf(m->extraMetrics[k].key, sizeof(m->extraMetrics[k].key), "%s", key);
      m->extraMetrics[k].value = v;
      ++k;
    };
    extra("memoryBandwidthUtilPercent", (double)u.memory);
    extra("utilizationSamplesAveraged", (double)util_samples);
    if (g_nv.nvmlDeviceGetClockInfo && g_nv.nvmlDeviceGetClockInfo(d.h, NVML_CLOCK_SM, &smclk) == NVML_SUCCESS) extra("clockSMMHz", smclk);
    extra("memoryTotalBytes", (double)d.mem);
    {
      const auto wi = workers.find(upper(d.uuid));
      const WorkerTotals t = wi == workers.end() ? WorkerTotals{} : wi->second;
      extra("tfwWorkers", (double)t.workers);
      extra("tfwStagedPayloadBytesTotal", (double)t.payload);
      extra("tfwH2DDmaBytesTotal", (double)t.h2d);
      extra("tfwD2HBytesTotal", (double)t.d2h);
      extra("tfwMoverLaunchesTotal", (double)t.movers);
      extra("tfwClientLaunchesTotal", (double)t.launches);
      extra("computeThrottledCnt", (double)t.throttled);   // internal/metrics/types.go:181
      extra("tfwGateTimeoutsTotal", (double)t.timeouts);
      extra("tfwVramBytes", (double)t.vram);
      extra("tfwFrozenWorkers", (double)t.frozen);
      extra("tfwParkedBytes", (double)t.parked);
    }
    m->extraMetricsCount = k;
  }
  return ACCEL_SUCCESS;
}

AccelResult AccelGetVendorMountLibs(MountPath* mounts, size_t maxCount, size_t* mountCount) {
  if (!mounts || maxCount == 0 || !mountCount) return ACCEL_ERROR_INVALID_PARAM;
  *mountCount = 0;  // driver libraries reach the pod through the NVIDIA container runtime
  return ACCEL_SUCCESS;
}

}  // extern "C"
