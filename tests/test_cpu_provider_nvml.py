"""The provider's device paths on the CPU: libaccelerator_b200.so dlopen()s NVML, so a stand-in
libnvidia-ml.so.1 (tools/mock_nvml.c, compiled against the real nvml.h) in front of LD_LIBRARY_PATH lets
AccelInit, discovery, topology, metrics, process information and the snapshot-by-device lookup run here.
Everything runs in subprocesses: the library resolves NVML once per process."""
import json
import os
import subprocess
import sys
import time

import pytest

import conftest

ROOT = conftest.ROOT
MOCK = os.path.join(ROOT, "build", "mock")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(os.path.join(MOCK, "libnvidia-ml.so.1")):
        subprocess.run(["make", "-s", "build/mock/libnvidia-ml.so.1"], cwd=ROOT, check=True)


def run_py(code, **env_extra):
    env = dict(os.environ, LD_LIBRARY_PATH=MOCK)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r)\n" % ROOT + code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout


def test_the_gpu_provider_tests_pass_against_the_mock_driver():
    """tests/test_gpu_provider.py minus the two tests that need a CUDA worker: discovery fields, topology, partitions,
    hard limits, the compiled hypervisor harness and the reference's own ABI suite -- here, without a GPU."""
    env = dict(os.environ, LD_LIBRARY_PATH=MOCK, TFW_RUN_GPU_MARKED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_provider.py"), "-q", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "not metrics_and_processes and not worker_counters"], env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:]
    assert int(tail.split(" passed")[0].split()[-1]) >= 5, tail


COMMON = r'''
import ctypes as C, json, os
from tensor_fusion_b200 import provider as P
lib = P.load()
assert lib.AccelInit() == P.SUCCESS
rc, devs = P.all_devices(lib)
assert rc == P.SUCCESS
n = len(devs)
'''


def test_discovery_topology_and_metrics_values():
    out = json.loads(run_py(COMMON + r'''
topo = P.ExtendedDeviceTopology()
assert lib.AccelGetAllDevicesTopology(C.byref(topo)) == P.SUCCESS
levels = sorted({topo.devices[i].peers[j].topoLevel for i in range(topo.deviceCount) for j in range(topo.devices[i].peerCount)})
uu = (C.c_char_p * n)(*[d["uuid"].encode() for d in devs])
dm = (P.DeviceMetrics * n)()
assert lib.AccelGetDeviceMetrics(uu, n, dm) == P.SUCCESS
pi = (P.ProcessInformation * 8)()
cnt = C.c_size_t()
rc = lib.AccelGetProcessInformation(pi, 8, C.byref(cnt))
print(json.dumps({"n": n, "devs": devs, "levels": levels, "peer_counts": [topo.devices[i].peerCount for i in range(topo.deviceCount)],
  "util": [dm[i].utilizationPercent for i in range(n)], "mem": [dm[i].memoryUsedBytes for i in range(n)], "power": dm[0].powerUsageWatts,
  "procs": [(pi[i].processId.decode(), pi[i].deviceUUID.decode(), pi[i].memoryUsedBytes, pi[i].computeUtilizationPercent) for i in range(cnt.value)], "prc": rc}))
lib.AccelShutdown()
''', MOCK_NVML_DEVICES="4", MOCK_NVML_UTIL="40", MOCK_NVML_MEM_USED="2048", MOCK_NVML_PIDS="4242:1073741824:17,4343:2147483648:5"))
    assert out["n"] == 4 and [d["uuid"] for d in out["devs"]] == ["GPU-b200000%d-aaaa-bbbb-cccc-0123456789ab" % i for i in range(4)]
    d0 = out["devs"][0]
    assert d0["vendor"] == "NVIDIA" and d0["model"] == "NVIDIA B200" and d0["sms"] == 148 and d0["tflops"] == 2250 and d0["pcie"] == [5, 16]
    assert out["levels"] == [0] and out["peer_counts"] == [3, 3, 3, 3]          # NVSwitch: every peer TOPO_LEVEL_INTERNAL
    assert out["util"] == [40, 41, 42, 43] and out["mem"] == [2048 << 20] * 4 and out["power"] == 180
    assert out["prc"] == 0 and out["procs"] == [["4242", out["devs"][0]["uuid"], 1 << 30, 17.0], ["4343", out["devs"][0]["uuid"], 2 << 30, 5.0]]


def test_device_utilisation_is_the_mean_over_the_polling_interval_not_one_sample():
    """nvmlDeviceGetUtilizationRates is the driver's latest ~1/6 s sample: bursty tenants read as 0 or 99 at random (C3 on
    the B200 box) and the ERL controller chases that.  AccelGetDeviceMetrics averages the driver's sample buffer since the
    previous call; the first call only learns the time stamps."""
    out = json.loads(run_py(COMMON + r'''
uu = (C.c_char_p * 1)(devs[0]["uuid"].encode())
dm = (P.DeviceMetrics * 1)()
got = []
for _ in range(3):
    assert lib.AccelGetDeviceMetrics(uu, 1, dm) == P.SUCCESS
    ex = {dm[0].extraMetrics[k].key.decode(): dm[0].extraMetrics[k].value for k in range(dm[0].extraMetricsCount)}
    got.append([dm[0].utilizationPercent, ex["utilizationSamplesAveraged"]])
print(json.dumps(got))
''', MOCK_NVML_UTIL="99", MOCK_NVML_UTIL_SAMPLES="0,99,0,21"))
    assert out == [[99.0, 0.0], [30.0, 4.0], [30.0, 4.0]]
    single = json.loads(run_py(COMMON + r'''
uu = (C.c_char_p * 1)(devs[0]["uuid"].encode())
dm = (P.DeviceMetrics * 1)()
for _ in range(2):
    assert lib.AccelGetDeviceMetrics(uu, 1, dm) == P.SUCCESS
print(json.dumps(dm[0].utilizationPercent))
''', MOCK_NVML_UTIL="99", MOCK_NVML_UTIL_SAMPLES="0,99,0,21", TF_UTIL_SINGLE_SAMPLE="1"))
    assert single == 99.0


def test_pcie_only_box_reports_the_common_ancestor_level():
    out = json.loads(run_py(COMMON + r'''
topo = P.ExtendedDeviceTopology()
assert lib.AccelGetAllDevicesTopology(C.byref(topo)) == P.SUCCESS
print(json.dumps(sorted({topo.devices[i].peers[j].topoLevel for i in range(topo.deviceCount) for j in range(topo.devices[i].peerCount)})))
''', MOCK_NVML_NO_NVLINK="1"))
    assert out == [4]     # NVML_TOPOLOGY_NODE -> TOPO_LEVEL_NUMA_NODE: the scheduler must not treat these peers as tier 0


def test_init_fails_loudly_when_the_driver_does_not_come_up():
    out = run_py(r'''
from tensor_fusion_b200 import provider as P
import ctypes as C
lib = P.load()
rc = lib.AccelInit()
n = C.c_size_t(99)
print(rc, lib.AccelGetDeviceCount(C.byref(n)), n.value)
''', MOCK_NVML_FAIL_INIT="1")
    rc, rc2, n = out.split()
    assert int(rc) == 5 and int(rc2) != 0 and int(n) in (0, 99)   # OPERATION_FAILED, and no fake devices afterwards


def test_extra_metrics_sum_worker_records_per_gpu(tmp_path):
    """Several pods on two GPUs publish tfw_stats records; AccelGetDeviceMetrics folds the live ones into each
    GPU's extraMetrics (a stale record and a record of another GPU do not leak in)."""
    out = json.loads(run_py(COMMON + r'''
import mmap, time
base = os.environ["TF_SHM_BASE_PATH"]
def record(ns, pod, uuid, payload, throttled, frozen, parked, age=0):
    d = os.path.join(base, ns, pod); os.makedirs(d)
    r = P.TfwStatsRecord()
    r.magic, r.version, r.pid, r.updated_unix_secs = P.TFW_STATS_MAGIC, P.TFW_STATS_VERSION, 1, int(time.time()) - age
    r.device_uuid = uuid.encode()
    r.payload_bytes, r.gate_blocked, r.ctl_frozen, r.parked_bytes, r.vram_bytes, r.mover_launches = payload, throttled, frozen, parked, 1000, 3
    open(os.path.join(d, "tfw_stats"), "wb").write(bytes(r))
u0, u1 = devs[0]["uuid"], devs[1]["uuid"]
record("a", "p1", u0, 100, 1, 0, 0)
record("a", "p2", u0, 200, 2, 1, 4096)
record("b", "p3", u1.upper(), 1000, 7, 0, 0)          # UUID case does not matter
record("b", "p4", u0, 5000, 50, 1, 1, age=3600)       # a dead worker's record
open(os.path.join(base, "b", "p4", "tfw_stats"), "ab").close()
os.makedirs(os.path.join(base, "c", "empty-pod"))
uu = (C.c_char_p * 2)(u0.encode(), u1.encode())
dm = (P.DeviceMetrics * 2)()
assert lib.AccelGetDeviceMetrics(uu, 2, dm) == P.SUCCESS
print(json.dumps([{dm[k].extraMetrics[i].key.decode(): dm[k].extraMetrics[i].value for i in range(dm[k].extraMetricsCount)} for k in range(2)]))
''', TF_SHM_BASE_PATH=str(tmp_path)))
    g0, g1 = out
    assert g0["tfwWorkers"] == 2 and g0["tfwStagedPayloadBytesTotal"] == 300 and g0["computeThrottledCnt"] == 3
    assert g0["tfwFrozenWorkers"] == 1 and g0["tfwParkedBytes"] == 4096 and g0["tfwVramBytes"] == 2000 and g0["tfwMoverLaunchesTotal"] == 6
    assert g1["tfwWorkers"] == 1 and g1["tfwStagedPayloadBytesTotal"] == 1000 and g1["computeThrottledCnt"] == 7 and g1["tfwFrozenWorkers"] == 0


def test_device_level_snapshot_resolves_aliases_and_uuid_spellings(tmp_path):
    """AccelSnapshot(deviceUUID): the UUID may come without the "GPU-" prefix, in another case, or as an alias
    (TF_PROVIDER_DEVICE_ALIASES); the request lands in every live worker record of that GPU only."""
    out = json.loads(run_py(COMMON + r'''
import mmap, threading, time
base = os.environ["TF_SHM_BASE_PATH"]
recs = {}
for pod, dev in (("w0", 0), ("w1", 1)):
    d = os.path.join(base, "ns", pod); os.makedirs(d)
    f = open(os.path.join(d, "tfw_stats"), "w+b"); f.truncate(C.sizeof(P.TfwStatsRecord))
    mm = mmap.mmap(f.fileno(), C.sizeof(P.TfwStatsRecord))
    r = P.TfwStatsRecord.from_buffer(mm)
    r.magic, r.version, r.pid, r.updated_unix_secs = P.TFW_STATS_MAGIC, P.TFW_STATS_VERSION, 4000000000 + dev, int(time.time())
    r.device_uuid = devs[dev]["uuid"].encode()
    recs[pod] = (r, mm, f)
stop = False
def ack():
    while not stop:
        for r, _, _ in recs.values():
            if r.ctl_ack != r.ctl_request: r.ctl_status = 0; r.ctl_ack = r.ctl_request
        time.sleep(0.001)
th = threading.Thread(target=ack); th.start()
seen = []
for spelling in (devs[1]["uuid"][4:].upper(), "second-gpu", devs[1]["uuid"]):
    ctx = P.SnapshotContext(processIds=None, processCount=0, deviceUUID=spelling.encode())
    rc = lib.AccelSnapshot(C.byref(ctx))
    seen.append((rc, recs["w0"][0].ctl_request, recs["w1"][0].ctl_request >> 8, recs["w1"][0].ctl_request & 0xff))
stop = True; th.join()
print(json.dumps(seen))
''', TF_SHM_BASE_PATH=str(tmp_path), TF_PROVIDER_DEVICE_ALIASES="second-gpu=1"))
    assert out == [[0, 0, 1, 1], [0, 0, 2, 1], [0, 0, 3, 1]]      # three freezes reached w1, none reached w0
