/* hypervisor_harness.c -- replays, from compiled code, the exact call sequence the Go hypervisor
 * issues against a provider library through purego:
 *   boot      pkg/hypervisor/device/accelerator_unix.go:42-124  dlopen(RTLD_NOW|RTLD_GLOBAL), 14 mandatory
 *             dlsym()s, AccelRegisterLogCallback (fallback name RegisterLogCallback), AccelInit
 *   discover  pkg/hypervisor/device/accelerator.go:408-470,471-520  count -> all devices (<=64) -> topology
 *   2 Hz loop quota_controller.go:388-395 (AccelGetDeviceMetrics) and worker/controller.go:587-632
 *             (AccelGetProcessInformation, 1024 entries)
 *   shutdown  accelerator.go:291-310  AccelShutdown, RegisterLogCallback(NULL)
 * Usage: hypervisor_harness <provider.so> [ticks]     exit 0 = every call answered as the Go side requires. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../include/tf_provider_abi.h"

static int n_logs, n_fatal;
static void on_log(const char* level, const char* msg) {
  ++n_logs;
  if (!strcmp(level, "FATAL")) ++n_fatal; /* klog.Fatal would terminate the hypervisor */
  fprintf(stderr, "  [provider %s] %s\n", level, msg);
}

#define MUST(sym) do { *(void**)(&sym) = dlsym(h, #sym); if (!sym) { fprintf(stderr, "missing mandatory symbol %s\n", #sym); return 2; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <provider.so> [ticks]\n", argv[0]); return 64; }
  const int ticks = argc > 2 ? atoi(argv[2]) : 3;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  AccelResult (*AccelInit)(void), (*AccelShutdown)(void), (*AccelGetDeviceCount)(size_t*);
  AccelResult (*AccelGetAllDevices)(ExtendedDeviceInfo*, size_t, size_t*), (*AccelGetAllDevicesTopology)(ExtendedDeviceTopology*);
  AccelResult (*AccelAssignPartition)(const char*, const char*, PartitionResult*), (*AccelRemovePartition)(const char*, const char*);
  AccelResult (*AccelSetMemHardLimit)(const char*, uint64_t), (*AccelSetComputeUnitHardLimit)(const char*, uint32_t);
  AccelResult (*AccelSnapshot)(SnapshotContext*), (*AccelResume)(SnapshotContext*);
  AccelResult (*AccelGetProcessInformation)(ProcessInformation*, size_t, size_t*);
  AccelResult (*AccelGetDeviceMetrics)(const char**, size_t, DeviceMetrics*), (*AccelGetVendorMountLibs)(MountPath*, size_t, size_t*);
  AccelResult (*RegisterLog)(LogCallbackFunc);
  MUST(AccelInit); MUST(AccelShutdown); MUST(AccelGetDeviceCount); MUST(AccelGetAllDevices); MUST(AccelGetAllDevicesTopology);
  MUST(AccelAssignPartition); MUST(AccelRemovePartition); MUST(AccelSetMemHardLimit); MUST(AccelSetComputeUnitHardLimit);
  MUST(AccelSnapshot); MUST(AccelResume); MUST(AccelGetProcessInformation); MUST(AccelGetDeviceMetrics); MUST(AccelGetVendorMountLibs);
  *(void**)(&RegisterLog) = dlsym(h, "AccelRegisterLogCallback");
  if (!RegisterLog) *(void**)(&RegisterLog) = dlsym(h, "RegisterLogCallback");
  if (RegisterLog) RegisterLog(on_log);
  (void)AccelAssignPartition; (void)AccelRemovePartition; (void)AccelSetMemHardLimit; (void)AccelSetComputeUnitHardLimit; (void)AccelSnapshot; (void)AccelResume;

  AccelResult r = AccelInit();
  if (r != ACCEL_SUCCESS) { fprintf(stderr, "AccelInit -> %d (hypervisor would refuse to start)\n", r); return 3; }
  size_t count = 0;
  if (AccelGetDeviceCount(&count) != ACCEL_SUCCESS || count == 0) { fprintf(stderr, "no devices\n"); return 4; }
  const size_t cap = count < 64 ? count : 64;
  ExtendedDeviceInfo* devs = calloc(cap, sizeof *devs);
  size_t got = 0;
  if (AccelGetAllDevices(devs, cap, &got) != ACCEL_SUCCESS || got == 0) return 5;
  ExtendedDeviceTopology* topo = calloc(1, sizeof *topo);
  r = AccelGetAllDevicesTopology(topo);
  if (r != ACCEL_SUCCESS && r != ACCEL_ERROR_NOT_SUPPORTED && r != ACCEL_ERROR_NOT_FOUND) return 6; /* accelerator.go:55-69 */
  printf("devices=%zu first={uuid=%s vendor=%s model=%s sms=%llu tflops=%.0f mem=%.1fGiB} topo_devices=%zu\n", got, devs[0].basic.uuid,
         devs[0].basic.vendor, devs[0].basic.model, (unsigned long long)devs[0].basic.totalComputeUnits, devs[0].basic.maxTflops,
         devs[0].basic.totalMemoryBytes / 1073741824.0, topo->deviceCount);
  MountPath* mounts = calloc(64, sizeof *mounts);
  size_t nm = 0;
  r = AccelGetVendorMountLibs(mounts, 64, &nm);
  if (r != ACCEL_SUCCESS && r != ACCEL_ERROR_NOT_SUPPORTED && r != ACCEL_ERROR_NOT_FOUND) return 7; /* allocation.go:79-90 */
  const char* uuids[64];
  for (size_t i = 0; i < got; ++i) uuids[i] = devs[i].basic.uuid;
  DeviceMetrics* dm = calloc(got, sizeof *dm);
  ProcessInformation* pi = calloc(1024, sizeof *pi);
  for (int t = 0; t < ticks; ++t) {
    if (AccelGetDeviceMetrics(uuids, got, dm) != ACCEL_SUCCESS) return 8;
    size_t np = 0;
    if (AccelGetProcessInformation(pi, 1024, &np) != ACCEL_SUCCESS) return 9;
    printf("tick %d: util=%u%% mem=%.1fGiB power=%.0fW extra=%zu procs=%zu\n", t, dm[0].utilizationPercent, dm[0].memoryUsedBytes / 1073741824.0,
           dm[0].powerUsageWatts, dm[0].extraMetricsCount, np);
    usleep(500000); /* erlUpdateInterval */
  }
  if (AccelShutdown() != ACCEL_SUCCESS) return 10;
  if (RegisterLog) RegisterLog(NULL);
  printf("ok logs=%d fatal=%d\n", n_logs, n_fatal);
  return n_fatal ? 11 : 0; /* the library is never dlclose()d by the hypervisor */
}
