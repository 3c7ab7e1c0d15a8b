/* mock_nvml.c -- a stand-in libnvidia-ml.so.1 for CPU tests of the provider (libaccelerator_b200.so dlopen()s NVML
 * at AccelInit).  Prototypes come from the real nvml.h, so a signature drift fails to compile.  It describes
 * MOCK_NVML_DEVICES (default 2) B200s on one NVSwitch; utilisation, memory, power and the process list can be
 * steered with MOCK_NVML_UTIL / MOCK_NVML_MEM_USED / MOCK_NVML_PIDS ("pid:bytes:smUtil,..." on device 0).
 * Test infrastructure, never shipped. */
#include <nvml.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))
#define MAXDEV 16

static int ndev(void) {
  const char* e = getenv("MOCK_NVML_DEVICES");
  int n = e ? atoi(e) : 2;
  return n < 0 ? 0 : n > MAXDEV ? MAXDEV : n;
}
static long env_long(const char* k, long dflt) { const char* e = getenv(k); return e && *e ? atol(e) : dflt; }
static int idx_of(nvmlDevice_t d) { return (int)((char*)d - (char*)0x1000) - 1; }   /* handles are 0x1001, 0x1002, ... */
static int bad(nvmlDevice_t d) { const int i = idx_of(d); return i < 0 || i >= ndev(); }

EXPORT nvmlReturn_t nvmlInit_v2(void) { return getenv("MOCK_NVML_FAIL_INIT") ? NVML_ERROR_DRIVER_NOT_LOADED : NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlShutdown(void) { return NVML_SUCCESS; }
EXPORT const char* nvmlErrorString(nvmlReturn_t r) { return r == NVML_SUCCESS ? "Success" : "mock NVML error"; }
EXPORT nvmlReturn_t nvmlDeviceGetCount_v2(unsigned* n) { *n = (unsigned)ndev(); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetHandleByIndex_v2(unsigned i, nvmlDevice_t* d) {
  if ((int)i >= ndev()) return NVML_ERROR_INVALID_ARGUMENT;
  *d = (nvmlDevice_t)((char*)0x1000 + i + 1);
  return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetUUID(nvmlDevice_t d, char* s, unsigned n) {
  if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT;
  snprintf(s, n, "GPU-%08x-aaaa-bbbb-cccc-0123456789ab", 0xb2000000u + (unsigned)idx_of(d));
  return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetName(nvmlDevice_t d, char* s, unsigned n) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; snprintf(s, n, "NVIDIA B200"); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlSystemGetDriverVersion(char* s, unsigned n) { snprintf(s, n, "580.159.00"); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetVbiosVersion(nvmlDevice_t d, char* s, unsigned n) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; snprintf(s, n, "97.00.88.00.0F"); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetMinorNumber(nvmlDevice_t d, unsigned* m) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *m = (unsigned)idx_of(d); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo(nvmlDevice_t d, nvmlMemory_t* m) {
  if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT;
  m->total = 183359ull << 20;  /* what a B200 reports: 183 359 MiB */
  m->used = (unsigned long long)env_long("MOCK_NVML_MEM_USED", 1024) << 20;
  m->free = m->total - m->used;
  return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetNumGpuCores(nvmlDevice_t d, unsigned* c) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *c = 148 * 128; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetCudaComputeCapability(nvmlDevice_t d, int* major, int* minor) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *major = 10; *minor = 0; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetMaxPcieLinkGeneration(nvmlDevice_t d, unsigned* g) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *g = 5; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetMaxPcieLinkWidth(nvmlDevice_t d, unsigned* w) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *w = 16; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetMaxClockInfo(nvmlDevice_t d, nvmlClockType_t t, unsigned* mhz) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *mhz = t == NVML_CLOCK_MEM ? 3996 : 1965; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetClockInfo(nvmlDevice_t d, nvmlClockType_t t, unsigned* mhz) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *mhz = t == NVML_CLOCK_MEM ? 3996 : 1500; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetPowerManagementLimit(nvmlDevice_t d, unsigned* mw) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *mw = 1000000; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetPowerUsage(nvmlDevice_t d, unsigned* mw) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *mw = (unsigned)env_long("MOCK_NVML_POWER_MW", 180000); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetTemperature(nvmlDevice_t d, nvmlTemperatureSensors_t s, unsigned* t) { (void)s; if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *t = 41; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetUtilizationRates(nvmlDevice_t d, nvmlUtilization_t* u) {
  if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT;
  u->gpu = (unsigned)env_long("MOCK_NVML_UTIL", 37) + (unsigned)idx_of(d);
  u->memory = 11;
  return NVML_SUCCESS;
}
/* MOCK_NVML_UTIL_SAMPLES="0,99,0,21": the driver's utilisation sample buffer; every call hands out the whole list with
 * time stamps newer than anything seen before (unset: NOT_SUPPORTED, the provider falls back to nvmlDeviceGetUtilizationRates) */
EXPORT nvmlReturn_t nvmlDeviceGetSamples(nvmlDevice_t d, nvmlSamplingType_t type, unsigned long long last, nvmlValueType_t* vt, unsigned* count, nvmlSample_t* samples) {
  if (bad(d) || !count || !vt) return NVML_ERROR_INVALID_ARGUMENT;
  const char* e = getenv("MOCK_NVML_UTIL_SAMPLES");
  if (!e || type != NVML_GPU_UTILIZATION_SAMPLES) return NVML_ERROR_NOT_SUPPORTED;
  unsigned vals[64], n = 0;
  while (*e && n < 64) { vals[n++] = (unsigned)strtoul(e, (char**)&e, 10); if (*e == ',') ++e; }
  *vt = NVML_VALUE_TYPE_UNSIGNED_INT;
  if (samples) {
    if (*count < n) return NVML_ERROR_INSUFFICIENT_SIZE;
    for (unsigned i = 0; i < n; ++i) { samples[i].timeStamp = last + 1000ull * (i + 1); samples[i].sampleValue.uiVal = vals[i]; }
  }
  *count = n;
  return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetPcieThroughput(nvmlDevice_t d, nvmlPcieUtilCounter_t c, unsigned* kbps) {
  if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT;
  *kbps = c == NVML_PCIE_UTIL_TX_BYTES ? 2000000 : 3000000;
  return NVML_SUCCESS;
}
/* "pid:bytes:smUtil,pid:bytes:smUtil" on device 0 */
static unsigned parse_pids(unsigned* pids, unsigned long long* bytes, unsigned* util, unsigned cap) {
  const char* e = getenv("MOCK_NVML_PIDS");
  unsigned n = 0;
  while (e && *e && n < cap) {
    unsigned p = 0, u = 0;
    unsigned long long b = 0;
    if (sscanf(e, "%u:%llu:%u", &p, &b, &u) < 1) break;
    pids[n] = p; bytes[n] = b; util[n] = u; ++n;
    e = strchr(e, ',');
    if (e) ++e;
  }
  return n;
}
EXPORT nvmlReturn_t nvmlDeviceGetComputeRunningProcesses_v3(nvmlDevice_t d, unsigned* count, nvmlProcessInfo_t* infos) {
  if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT;
  unsigned pids[32], util[32];
  unsigned long long bytes[32];
  const unsigned n = idx_of(d) == 0 ? parse_pids(pids, bytes, util, 32) : 0;
  if (*count < n || (!infos && n)) { *count = n; return NVML_ERROR_INSUFFICIENT_SIZE; }
  for (unsigned i = 0; i < n; ++i) { memset(&infos[i], 0, sizeof infos[i]); infos[i].pid = pids[i]; infos[i].usedGpuMemory = bytes[i]; }
  *count = n;
  return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetProcessUtilization(nvmlDevice_t d, nvmlProcessUtilizationSample_t* s, unsigned* count, unsigned long long last) {
  (void)last;
  if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT;
  unsigned pids[32], util[32];
  unsigned long long bytes[32];
  const unsigned n = idx_of(d) == 0 ? parse_pids(pids, bytes, util, 32) : 0;
  if (!n) { *count = 0; return NVML_ERROR_NOT_FOUND; }
  if (!s || *count < n) { *count = n; return NVML_ERROR_INSUFFICIENT_SIZE; }
  for (unsigned i = 0; i < n; ++i) { memset(&s[i], 0, sizeof s[i]); s[i].pid = pids[i]; s[i].smUtil = util[i]; s[i].timeStamp = 1000 + i; }
  *count = n;
  return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetTopologyCommonAncestor(nvmlDevice_t a, nvmlDevice_t b, nvmlGpuTopologyLevel_t* l) {
  if (bad(a) || bad(b)) return NVML_ERROR_INVALID_ARGUMENT;
  *l = NVML_TOPOLOGY_NODE;
  return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetP2PStatus(nvmlDevice_t a, nvmlDevice_t b, nvmlGpuP2PCapsIndex_t i, nvmlGpuP2PStatus_t* st) {
  if (bad(a) || bad(b)) return NVML_ERROR_INVALID_ARGUMENT;
  /* MOCK_NVML_NO_NVLINK: a PCIe-only box, the provider falls back to the common-ancestor level */
  *st = (i == NVML_P2P_CAPS_INDEX_NVLINK && !getenv("MOCK_NVML_NO_NVLINK")) ? NVML_P2P_STATUS_OK : NVML_P2P_STATUS_NOT_SUPPORTED;
  return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetNumaNodeId(nvmlDevice_t d, unsigned* node) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *node = (unsigned)idx_of(d) / 4; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetEccMode(nvmlDevice_t d, nvmlEnableState_t* cur, nvmlEnableState_t* pend) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *cur = *pend = NVML_FEATURE_ENABLED; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetPersistenceMode(nvmlDevice_t d, nvmlEnableState_t* m) { if (bad(d)) return NVML_ERROR_INVALID_ARGUMENT; *m = NVML_FEATURE_ENABLED; return NVML_SUCCESS; }
