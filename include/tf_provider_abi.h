/*
 * tf_provider_abi.h -- the tensor-fusion provider C-ABI as implemented by
 * libaccelerator_b200.so (drop-in boundary #1).
 *
 * This header RESTATES the binary contract of the reference's
 *   provider/accelerator.h  (types :47-261, functions :275-439)
 *   provider/limiter.h      (types :36-65,  functions :71-106)
 * so that the Go hypervisor can dlopen this library through purego exactly as
 * it does the vendor libraries (pkg/hypervisor/device/accelerator_unix.go:42-124):
 * same exported names, same argument order, same POD layouts.  Every size and
 * offset the Go mirror structs rely on (pkg/hypervisor/device/accelerator.go:71-234,
 * SURVEY.md App. A) is pinned by a static assertion at the bottom.
 *
 * Where header and Go mirror disagree (App. E-1: DeviceTopoNode is 72 bytes in
 * the header, 80 in Go) the in-tree HEADER is the contract.
 *
 * Ownership: the caller allocates every output buffer; strings are fixed-size,
 * NUL-terminated char arrays; the only pointer the library keeps is the log
 * callback.  All entry points are thread-safe.
 */
#ifndef TF_PROVIDER_ABI_H
#define TF_PROVIDER_ABI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TF_ABI_EXPORT __attribute__((visibility("default")))

/* ---- result codes (accelerator.h:47-55) -------------------------------- */
typedef enum {
  ACCEL_SUCCESS = 0,
  ACCEL_ERROR_INVALID_PARAM = 1,
  ACCEL_ERROR_NOT_FOUND = 2,
  ACCEL_ERROR_NOT_SUPPORTED = 3,
  ACCEL_ERROR_RESOURCE_EXHAUSTED = 4,
  ACCEL_ERROR_OPERATION_FAILED = 5,
  ACCEL_ERROR_INTERNAL = 6
} AccelResult;

/* Levels the Go side understands: "DEBUG" "INFO" "WARN" "ERROR" "FATAL"
 * (accelerator_unix.go:138-151).  "FATAL" makes klog terminate the
 * hypervisor, so this library never emits it. */
typedef void (*LogCallbackFunc)(const char* level, const char* message);

/* ---- capacities --------------------------------------------------------- */
#define MAX_DEVICE_PROPERTIES 64
#define MAX_TOPOLOGY_DEVICES 64
#define MAX_PARTITION_ENVS 16
#define MAX_ENV_KEY_LENGTH 64
#define MAX_ENV_VALUE_LENGTH 256
#define MAX_PROCESSES 1024
#define MAX_EXTRA_METRICS 64
#define MAX_DEVICE_UUIDS 64
#define UUID_STRING_LENGTH 64
#define MAX_MOUNT_PATH 512
#define MAX_PARTITION_DEVICE_NODES 16
#define MAX_PARTITION_DEVICE_NODE_LENGTH (MAX_MOUNT_PATH * 2 + 2)

/* ---- device description (accelerator.h:69-122) -------------------------- */
typedef struct {
  bool supportsPartitioning;
  bool supportsSoftIsolation;
  bool supportsHardIsolation;
  bool supportsSnapshot;
  bool supportsMetrics;
  bool supportsRemoting;
  uint32_t maxPartitions;
  uint32_t maxWorkersPerDevice;
} VirtualizationCapabilities; /* 16 B */

typedef struct {
  char uuid[64];
  char vendor[32];
  char model[128];
  char driverVersion[80];
  char firmwareVersion[64];
  char deviceNode[64];
  int32_t index;
  int32_t numaNode; /* -1 when unknown */
  uint64_t totalMemoryBytes;
  uint64_t totalComputeUnits; /* SMs */
  double maxTflops;
  uint32_t pcieGen;
  uint32_t pcieWidth;
} DeviceBasicInfo; /* 472 B */

typedef struct {
  char key[64];
  char value[256];
} DevicePropertyKV; /* 320 B */

typedef struct {
  DevicePropertyKV properties[MAX_DEVICE_PROPERTIES];
  size_t count;
} DeviceProperties; /* 20 488 B */

typedef struct {
  DeviceBasicInfo basic;
  DeviceProperties props;
  VirtualizationCapabilities virtualizationCapabilities;
} ExtendedDeviceInfo; /* 20 976 B */

typedef struct { /* declared by the reference, used by no function (accelerator.h:124-134) */
  char templateId[64];
  char name[128];
  uint64_t memoryBytes;
  uint64_t computeUnits;
  double tflops;
  uint32_t sliceCount;
  bool isDefault;
  char description[256];
} PartitionTemplate; /* 480 B */

/* ---- topology (accelerator.h:136-172) ----------------------------------- */
typedef enum {
  TOPO_LEVEL_INTERNAL = 0, /* NVLink / same board: scheduler tier 0 (accelerator.go:323-334) */
  TOPO_LEVEL_SINGLE_SWITCH = 1,
  TOPO_LEVEL_MULTI_SWITCH = 2,
  TOPO_LEVEL_HOST_BRIDGE = 3,
  TOPO_LEVEL_NUMA_NODE = 4,
  TOPO_LEVEL_SYSTEM = 5,
  TOPO_LEVEL_SELF = 6,
  TOPO_LEVEL_UNKNOWN = 7
} TopoLevelType;

typedef struct {
  char peerUUID[64];
  int32_t peerIndex;
  TopoLevelType topoLevel;
} DeviceTopoNode; /* 72 B (header wins over the 80-byte Go mirror) */

typedef struct {
  char deviceUUID[64];
  int32_t deviceIndex;
  int32_t numaNode;
  DeviceTopoNode peers[MAX_TOPOLOGY_DEVICES];
  size_t peerCount;
} DeviceTopologyInfo; /* 4 688 B */

typedef struct {
  DeviceTopologyInfo devices[MAX_TOPOLOGY_DEVICES];
  size_t deviceCount;
} ExtendedDeviceTopology; /* 300 040 B -- never write past this */

/* ---- snapshot / partition (accelerator.h:176-261) ------------------------ */
typedef struct {
  pid_t* processIds;      /* process-level snapshot, else NULL */
  size_t processCount;
  const char* deviceUUID; /* device-level snapshot, else NULL */
} SnapshotContext; /* 24 B */

typedef enum {
  PARTITION_TYPE_ENVIRONMENT_VARIABLE = 0,
  PARTITION_TYPE_DEVICE_NODE = 1
} PartitionResultType;

typedef struct {
  PartitionResultType type;
  char deviceUUID[64];
  char envVars[MAX_PARTITION_ENVS][MAX_ENV_VALUE_LENGTH];                         /* "KEY=VALUE" */
  char deviceNodes[MAX_PARTITION_DEVICE_NODES][MAX_PARTITION_DEVICE_NODE_LENGTH]; /* "host=guest" */
} PartitionResult; /* 20 580 B */

/* ---- metrics (accelerator.h:195-242) -------------------------------------- */
typedef struct {
  char key[64];
  double value;
} ExtraMetric; /* 72 B */

typedef struct {
  char processId[32];
  char deviceUUID[64];
  double computeUtilizationPercent;
  uint64_t activeSMs;
  uint64_t totalSMs;
  uint64_t memoryUsedBytes;
  uint64_t memoryReservedBytes;
  double memoryUtilizationPercent;
} ProcessInformation; /* 144 B */

typedef struct {
  char deviceUUID[64];
  double powerUsageWatts;
  double temperatureCelsius;
  uint64_t pcieRxBytes;
  uint64_t pcieTxBytes;
  uint32_t utilizationPercent; /* whole-device; the ERL controller's input (quota_controller.go:388-395) */
  uint64_t memoryUsedBytes;
  ExtraMetric extraMetrics[MAX_EXTRA_METRICS];
  size_t extraMetricsCount;
} DeviceMetrics; /* 4 728 B */

typedef struct {
  char hostPath[MAX_MOUNT_PATH];
  char guestPath[MAX_MOUNT_PATH];
} MountPath; /* 1 024 B */

/* ---- limiter records (limiter.h:36-65) ------------------------------------- */
typedef struct {
  char deviceUUID[64];
  int64_t bytesDiff;
  bool shouldBlock;
  uint64_t availableBytes;
} MemoryOpRecord; /* 88 B */

typedef struct {
  char deviceUUID[64];
  uint64_t computeTokens;
  bool shouldBlock;
  uint64_t availableTokens;
} ComputeOpRecord; /* 88 B */

typedef struct {
  char workerId[64];
  bool isFrozen;
  uint64_t freezeTimeMs;
} WorkerFreezeState; /* 80 B */

typedef struct {
  uint32_t deviceIdx;
  char deviceUUID[64];
  uint32_t upLimit; /* percent 0..100 */
  uint64_t memLimit;
  uint32_t totalCudaCores;
} LimiterDeviceConfig; /* 88 B */

/* ============================ accelerator.h:275-439 ========================== */
TF_ABI_EXPORT AccelResult AccelInit(void);
TF_ABI_EXPORT AccelResult AccelShutdown(void);
TF_ABI_EXPORT AccelResult AccelGetDeviceCount(size_t* deviceCount);
TF_ABI_EXPORT AccelResult AccelGetAllDevices(ExtendedDeviceInfo* devices, size_t maxCount, size_t* deviceCount);
TF_ABI_EXPORT AccelResult AccelGetAllDevicesTopology(ExtendedDeviceTopology* topology);
TF_ABI_EXPORT AccelResult AccelAssignPartition(const char* templateId, const char* deviceUUID,
                                               PartitionResult* partitionResult);
TF_ABI_EXPORT AccelResult AccelRemovePartition(const char* templateId, const char* deviceUUID);
TF_ABI_EXPORT AccelResult AccelSetMemHardLimit(const char* deviceUUID, uint64_t memoryLimitBytes);
TF_ABI_EXPORT AccelResult AccelSetComputeUnitHardLimit(const char* deviceUUID, uint32_t computeUnitLimit);
TF_ABI_EXPORT AccelResult AccelSnapshot(SnapshotContext* context);
TF_ABI_EXPORT AccelResult AccelResume(SnapshotContext* context);
TF_ABI_EXPORT AccelResult AccelGetProcessInformation(ProcessInformation* processInfos, size_t maxCount,
                                                     size_t* processInfoCount);
TF_ABI_EXPORT AccelResult AccelGetDeviceMetrics(const char** deviceUUIDs, size_t deviceCount, DeviceMetrics* metrics);
TF_ABI_EXPORT AccelResult AccelGetVendorMountLibs(MountPath* mounts, size_t maxCount, size_t* mountCount);
TF_ABI_EXPORT AccelResult AccelRegisterLogCallback(LogCallbackFunc callback);

/* ============================== limiter.h:71-106 ============================== */
/* worker-facing (called from the CUDA hook / this repo's worker) */
TF_ABI_EXPORT AccelResult CheckAndRecordMemoryOps(const char* processId, const char* deviceUUID, int64_t bytesDiff,
                                                  MemoryOpRecord* record);
TF_ABI_EXPORT AccelResult CheckAndRecordComputeOps(const char* processId, const char* deviceUUID,
                                                   uint64_t computeTokens, ComputeOpRecord* record);
TF_ABI_EXPORT AccelResult FreezeWorker(const char* workerId, WorkerFreezeState* state);
TF_ABI_EXPORT AccelResult ResumeWorker(const char* workerId, WorkerFreezeState* state);
TF_ABI_EXPORT AccelResult AutoFreeze(const char* workerId, const char* deviceUUID, const char* resourceType);
TF_ABI_EXPORT AccelResult AutoResume(const char* workerId, const char* deviceUUID, const char* resourceType);
TF_ABI_EXPORT AccelResult AddWorkerProcess(const char* deviceUUID, const char* processId);
/* hypervisor-facing (quota file management + ERL) */
TF_ABI_EXPORT AccelResult LimiterInit(const char* shmBasePath);
TF_ABI_EXPORT AccelResult LimiterShutdown(void);
TF_ABI_EXPORT AccelResult LimiterCreateWorker(const char* namespace_, const char* podName,
                                              const LimiterDeviceConfig* configs, size_t configCount);
TF_ABI_EXPORT AccelResult LimiterRemoveWorker(const char* namespace_, const char* podName);
TF_ABI_EXPORT AccelResult LimiterRegisterPID(const char* namespace_, const char* podName, uint32_t hostPID);
TF_ABI_EXPORT AccelResult LimiterUpdateERL(const char* namespace_, const char* podName, uint32_t deviceIdx,
                                           uint32_t upLimit, double utilizationPercent, uint64_t timestampMicros);
TF_ABI_EXPORT AccelResult LimiterUpdateHeartbeat(const char* namespace_, const char* podName, uint64_t timestampSecs);
TF_ABI_EXPORT AccelResult LimiterSetPodMemoryUsed(const char* namespace_, const char* podName, uint32_t deviceIdx,
                                                  uint64_t memoryUsed);
/* Extension (not in provider/limiter.h): computeUpLimit of the hypervisor (pkg/hypervisor/worker/controller.go:307-325
 * == computeLimitPercent, server/handlers/legacy.go:643-661) for hosts that are not the Go hypervisor -- the worker uses
 * it to turn RemotePodInfo.tflops_limit into its SM partition under hard isolation, tools/limiter_c3.py to play the
 * hypervisor.  computePercent > 0 wins; else ceil(tflopsLimit / maxTflops * 100) clamped to [1, 100]; else 100. */
TF_ABI_EXPORT uint32_t LimiterComputeUpLimit(int64_t computePercent, double tflopsLimit, double maxTflops);

/* ---- layout pins (gcc x86-64) ------------------------------ */
#if defined(__cplusplus)
#define TF_ABI_ASSERT(c, m) static_assert(c, m)
#else
#define TF_ABI_ASSERT(c, m) _Static_assert(c, m)
#endif
TF_ABI_ASSERT(sizeof(AccelResult) == 4, "AccelResult");
TF_ABI_ASSERT(sizeof(VirtualizationCapabilities) == 16 && offsetof(VirtualizationCapabilities, maxPartitions) == 8, "caps");
TF_ABI_ASSERT(sizeof(DeviceBasicInfo) == 472 && offsetof(DeviceBasicInfo, index) == 432 &&
              offsetof(DeviceBasicInfo, totalMemoryBytes) == 440 && offsetof(DeviceBasicInfo, maxTflops) == 456 &&
              offsetof(DeviceBasicInfo, pcieWidth) == 468, "DeviceBasicInfo");
TF_ABI_ASSERT(sizeof(DevicePropertyKV) == 320, "DevicePropertyKV");
TF_ABI_ASSERT(sizeof(DeviceProperties) == 20488 && offsetof(DeviceProperties, count) == 20480, "DeviceProperties");
TF_ABI_ASSERT(sizeof(ExtendedDeviceInfo) == 20976 && offsetof(ExtendedDeviceInfo, props) == 472 &&
              offsetof(ExtendedDeviceInfo, virtualizationCapabilities) == 20960, "ExtendedDeviceInfo");
TF_ABI_ASSERT(sizeof(PartitionTemplate) == 480, "PartitionTemplate");
TF_ABI_ASSERT(sizeof(DeviceTopoNode) == 72 && offsetof(DeviceTopoNode, topoLevel) == 68, "DeviceTopoNode");
TF_ABI_ASSERT(sizeof(DeviceTopologyInfo) == 4688 && offsetof(DeviceTopologyInfo, peers) == 72 &&
              offsetof(DeviceTopologyInfo, peerCount) == 4680, "DeviceTopologyInfo");
TF_ABI_ASSERT(sizeof(ExtendedDeviceTopology) == 300040 && offsetof(ExtendedDeviceTopology, deviceCount) == 300032, "topology");
TF_ABI_ASSERT(sizeof(SnapshotContext) == 24, "SnapshotContext");
TF_ABI_ASSERT(sizeof(ExtraMetric) == 72 && offsetof(ExtraMetric, value) == 64, "ExtraMetric");
TF_ABI_ASSERT(sizeof(ProcessInformation) == 144 && offsetof(ProcessInformation, memoryUsedBytes) == 120, "ProcessInformation");
TF_ABI_ASSERT(sizeof(DeviceMetrics) == 4728 && offsetof(DeviceMetrics, utilizationPercent) == 96 &&
              offsetof(DeviceMetrics, extraMetrics) == 112 && offsetof(DeviceMetrics, extraMetricsCount) == 4720, "DeviceMetrics");
TF_ABI_ASSERT(sizeof(MountPath) == 1024, "MountPath");
TF_ABI_ASSERT(sizeof(PartitionResult) == 20580 && offsetof(PartitionResult, envVars) == 68 &&
              offsetof(PartitionResult, deviceNodes) == 4164, "PartitionResult");
TF_ABI_ASSERT(sizeof(MemoryOpRecord) == 88 && offsetof(MemoryOpRecord, shouldBlock) == 72 &&
              offsetof(MemoryOpRecord, availableBytes) == 80, "MemoryOpRecord");
TF_ABI_ASSERT(sizeof(ComputeOpRecord) == 88, "ComputeOpRecord");
TF_ABI_ASSERT(sizeof(WorkerFreezeState) == 80 && offsetof(WorkerFreezeState, freezeTimeMs) == 72, "WorkerFreezeState");
TF_ABI_ASSERT(sizeof(LimiterDeviceConfig) == 88 && offsetof(LimiterDeviceConfig, upLimit) == 68 &&
              offsetof(LimiterDeviceConfig, memLimit) == 72 && offsetof(LimiterDeviceConfig, totalCudaCores) == 80, "LimiterDeviceConfig");

#ifdef __cplusplus
}
#endif
#endif /* TF_PROVIDER_ABI_H */
