"""ctypes mirror of the provider C-ABI (include/tf_provider_abi.h), i.e. what the Go
hypervisor binds through purego (pkg/hypervisor/device/accelerator_unix.go:42-124).
Struct layouts are checked against the sizes the Go mirror relies on."""
import ctypes as C
import os

from . import PACKAGE_DIR

LIB_PATH = os.path.join(PACKAGE_DIR, "lib", "libaccelerator_b200.so")

SUCCESS, INVALID_PARAM, NOT_FOUND, NOT_SUPPORTED, RESOURCE_EXHAUSTED, OPERATION_FAILED, INTERNAL = range(7)
MAX_DEVICE_PROPERTIES = MAX_TOPOLOGY_DEVICES = MAX_EXTRA_METRICS = 64

LogCallback = C.CFUNCTYPE(None, C.c_char_p, C.c_char_p)


class VirtualizationCapabilities(C.Structure):
    _fields_ = [(n, C.c_bool) for n in ("supportsPartitioning", "supportsSoftIsolation", "supportsHardIsolation",
                                        "supportsSnapshot", "supportsMetrics", "supportsRemoting")] + \
               [("maxPartitions", C.c_uint32), ("maxWorkersPerDevice", C.c_uint32)]


class DeviceBasicInfo(C.Structure):
    _fields_ = [("uuid", C.c_char * 64), ("vendor", C.c_char * 32), ("model", C.c_char * 128),
                ("driverVersion", C.c_char * 80), ("firmwareVersion", C.c_char * 64), ("deviceNode", C.c_char * 64),
                ("index", C.c_int32), ("numaNode", C.c_int32), ("totalMemoryBytes", C.c_uint64),
                ("totalComputeUnits", C.c_uint64), ("maxTflops", C.c_double), ("pcieGen", C.c_uint32),
                ("pcieWidth", C.c_uint32)]


class DevicePropertyKV(C.Structure):
    _fields_ = [("key", C.c_char * 64), ("value", C.c_char * 256)]


class DeviceProperties(C.Structure):
    _fields_ = [("properties", DevicePropertyKV * 64), ("count", C.c_size_t)]


class ExtendedDeviceInfo(C.Structure):
    _fields_ = [("basic", DeviceBasicInfo), ("props", DeviceProperties),
                ("virtualizationCapabilities", VirtualizationCapabilities)]


class DeviceTopoNode(C.Structure):
    _fields_ = [("peerUUID", C.c_char * 64), ("peerIndex", C.c_int32), ("topoLevel", C.c_int32)]


class DeviceTopologyInfo(C.Structure):
    _fields_ = [("deviceUUID", C.c_char * 64), ("deviceIndex", C.c_int32), ("numaNode", C.c_int32),
                ("peers", DeviceTopoNode * 64), ("peerCount", C.c_size_t)]


class ExtendedDeviceTopology(C.Structure):
    _fields_ = [("devices", DeviceTopologyInfo * 64), ("deviceCount", C.c_size_t)]


class TfwStatsRecord(C.Structure):
    """include/tfw_stats_file.h: the record a vGPU worker publishes next to its quota file; the last
    words are the provider -> worker control channel used by AccelSnapshot / AccelResume."""
    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint32), ("seq", C.c_uint64), ("pid", C.c_uint64),
                ("updated_unix_secs", C.c_uint64), ("device_uuid", C.c_char * 64)] + \
               [(k, C.c_uint64) for k in ("frames", "payload_bytes", "h2d_dma_bytes", "d2h_bytes", "d2d_bytes", "fill_bytes",
                                          "mover_launches", "client_launches", "gate_launches", "vram_bytes", "vram_peak_bytes",
                                          "live_buffers", "gate_admitted", "gate_blocked", "gate_timeouts", "ctl_request", "ctl_ack",
                                          "ctl_status", "ctl_frozen", "ctl_moved_bytes", "parked_bytes")] + \
               [("ctl_arg", C.c_uint64), ("reserved", C.c_uint64), ("worker_id", C.c_char * 64)] + \
               [(k, C.c_uint64) for k in ("frozen_unix_ms", "frozen_auto", "auto_freezes", "auto_resumes", "sm_limit_percent", "sm_count",
                                          "vram_limit_bytes")]


TFW_STATS_MAGIC, TFW_STATS_VERSION, TFW_CTL_FREEZE, TFW_CTL_RESUME = 0x53574654, 2, 1, 2


class SnapshotContext(C.Structure):
    _fields_ = [("processIds", C.POINTER(C.c_int32)), ("processCount", C.c_size_t), ("deviceUUID", C.c_char_p)]


class PartitionResult(C.Structure):
    _fields_ = [("type", C.c_int32), ("deviceUUID", C.c_char * 64), ("envVars", (C.c_char * 256) * 16),
                ("deviceNodes", (C.c_char * 1026) * 16)]


class ExtraMetric(C.Structure):
    _fields_ = [("key", C.c_char * 64), ("value", C.c_double)]


class ProcessInformation(C.Structure):
    _fields_ = [("processId", C.c_char * 32), ("deviceUUID", C.c_char * 64), ("computeUtilizationPercent", C.c_double),
                ("activeSMs", C.c_uint64), ("totalSMs", C.c_uint64), ("memoryUsedBytes", C.c_uint64),
                ("memoryReservedBytes", C.c_uint64), ("memoryUtilizationPercent", C.c_double)]


class DeviceMetrics(C.Structure):
    _fields_ = [("deviceUUID", C.c_char * 64), ("powerUsageWatts", C.c_double), ("temperatureCelsius", C.c_double),
                ("pcieRxBytes", C.c_uint64), ("pcieTxBytes", C.c_uint64), ("utilizationPercent", C.c_uint32),
                ("memoryUsedBytes", C.c_uint64), ("extraMetrics", ExtraMetric * 64), ("extraMetricsCount", C.c_size_t)]


class MountPath(C.Structure):
    _fields_ = [("hostPath", C.c_char * 512), ("guestPath", C.c_char * 512)]


class MemoryOpRecord(C.Structure):
    _fields_ = [("deviceUUID", C.c_char * 64), ("bytesDiff", C.c_int64), ("shouldBlock", C.c_bool),
                ("availableBytes", C.c_uint64)]


class ComputeOpRecord(C.Structure):
    _fields_ = [("deviceUUID", C.c_char * 64), ("computeTokens", C.c_uint64), ("shouldBlock", C.c_bool),
                ("availableTokens", C.c_uint64)]


class WorkerFreezeState(C.Structure):
    _fields_ = [("workerId", C.c_char * 64), ("isFrozen", C.c_bool), ("freezeTimeMs", C.c_uint64)]


class LimiterDeviceConfig(C.Structure):
    _fields_ = [("deviceIdx", C.c_uint32), ("deviceUUID", C.c_char * 64), ("upLimit", C.c_uint32),
                ("memLimit", C.c_uint64), ("totalCudaCores", C.c_uint32)]


# sizes the Go mirror structs assume
EXPECTED_SIZES = {
    VirtualizationCapabilities: 16, DeviceBasicInfo: 472, DevicePropertyKV: 320, DeviceProperties: 20488,
    ExtendedDeviceInfo: 20976, DeviceTopoNode: 72, DeviceTopologyInfo: 4688, ExtendedDeviceTopology: 300040,
    SnapshotContext: 24, PartitionResult: 20580, ExtraMetric: 72, ProcessInformation: 144, DeviceMetrics: 4728,
    MountPath: 1024, MemoryOpRecord: 88, ComputeOpRecord: 88, WorkerFreezeState: 80, LimiterDeviceConfig: 88,
}

_P, _S = C.c_void_p, C.c_char_p
SIGS = {
    "AccelInit": [], "AccelShutdown": [], "AccelGetDeviceCount": [C.POINTER(C.c_size_t)],
    "AccelGetAllDevices": [C.POINTER(ExtendedDeviceInfo), C.c_size_t, C.POINTER(C.c_size_t)],
    "AccelGetAllDevicesTopology": [C.POINTER(ExtendedDeviceTopology)],
    "AccelAssignPartition": [_S, _S, C.POINTER(PartitionResult)], "AccelRemovePartition": [_S, _S],
    "AccelSetMemHardLimit": [_S, C.c_uint64], "AccelSetComputeUnitHardLimit": [_S, C.c_uint32],
    "AccelSnapshot": [C.POINTER(SnapshotContext)], "AccelResume": [C.POINTER(SnapshotContext)],
    "AccelGetProcessInformation": [C.POINTER(ProcessInformation), C.c_size_t, C.POINTER(C.c_size_t)],
    "AccelGetDeviceMetrics": [C.POINTER(_S), C.c_size_t, C.POINTER(DeviceMetrics)],
    "AccelGetVendorMountLibs": [C.POINTER(MountPath), C.c_size_t, C.POINTER(C.c_size_t)],
    "AccelRegisterLogCallback": [LogCallback],
    "CheckAndRecordMemoryOps": [_S, _S, C.c_int64, C.POINTER(MemoryOpRecord)],
    "CheckAndRecordComputeOps": [_S, _S, C.c_uint64, C.POINTER(ComputeOpRecord)],
    "FreezeWorker": [_S, C.POINTER(WorkerFreezeState)], "ResumeWorker": [_S, C.POINTER(WorkerFreezeState)],
    "AutoFreeze": [_S, _S, _S], "AutoResume": [_S, _S, _S], "AddWorkerProcess": [_S, _S],
    "LimiterInit": [_S], "LimiterShutdown": [],
    "LimiterCreateWorker": [_S, _S, C.POINTER(LimiterDeviceConfig), C.c_size_t], "LimiterRemoveWorker": [_S, _S],
    "LimiterRegisterPID": [_S, _S, C.c_uint32],
    "LimiterUpdateERL": [_S, _S, C.c_uint32, C.c_uint32, C.c_double, C.c_uint64],
    "LimiterUpdateHeartbeat": [_S, _S, C.c_uint64], "LimiterSetPodMemoryUsed": [_S, _S, C.c_uint32, C.c_uint64],
}
# the 14 symbols the hypervisor refuses to start without (accelerator_unix.go:57-98)
MANDATORY = [n for n in SIGS if n.startswith("Accel") and n != "AccelRegisterLogCallback"]


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `make`")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)  # purego: RTLD_NOW|RTLD_GLOBAL
    for name, args in SIGS.items():
        f = getattr(lib, name)
        f.restype = C.c_int32
        f.argtypes = args
    return lib


def all_devices(lib, cap=64):
    """AcceleratorInterface.GetAllDevices (device/accelerator.go:408-470): count, then fill min(count, 64)."""
    n = C.c_size_t()
    rc = lib.AccelGetDeviceCount(C.byref(n))
    if rc != SUCCESS:
        return rc, []
    if n.value == 0:
        return SUCCESS, []
    k = min(n.value, cap)
    buf = (ExtendedDeviceInfo * k)()
    got = C.c_size_t()
    rc = lib.AccelGetAllDevices(buf, k, C.byref(got))
    out = []
    for i in range(got.value if rc == SUCCESS else 0):
        d = buf[i]
        props = {d.props.properties[j].key.decode(): d.props.properties[j].value.decode() for j in range(d.props.count)}
        out.append({"uuid": d.basic.uuid.decode(), "vendor": d.basic.vendor.decode(), "model": d.basic.model.decode(),
                    "driver": d.basic.driverVersion.decode(), "node": d.basic.deviceNode.decode(), "index": d.basic.index,
                    "numa": d.basic.numaNode, "memory": d.basic.totalMemoryBytes, "sms": d.basic.totalComputeUnits,
                    "tflops": d.basic.maxTflops, "pcie": (d.basic.pcieGen, d.basic.pcieWidth), "props": props,
                    "caps": {n: getattr(d.virtualizationCapabilities, n) for n, _ in VirtualizationCapabilities._fields_}})
    return rc, out
