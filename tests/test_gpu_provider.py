"""Provider C-ABI on a real B200 node: what the Go hypervisor would see through purego."""
import ctypes as C
import os
import subprocess

import pytest

import conftest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prov():
    from tensor_fusion_b200 import provider as P
    lib = P.load()
    logs = []
    cb = P.LogCallback(lambda lvl, msg: logs.append((lvl, msg)))
    assert lib.AccelRegisterLogCallback(cb) == P.SUCCESS       # before AccelInit, like accelerator_unix.go:102-117
    assert lib.AccelInit() == P.SUCCESS
    yield P, lib, logs
    lib.AccelShutdown()
    lib.AccelRegisterLogCallback(P.LogCallback())
    del cb


def test_devices_as_the_hypervisor_sees_them(prov):
    P, lib, logs = prov
    rc, devs = P.all_devices(lib)
    assert rc == P.SUCCESS and len(devs) >= 1
    d = devs[0]
    assert d["uuid"].startswith("GPU-") and d["vendor"] == "NVIDIA"       # quota_controller.go:199-202
    assert "B200" in d["model"] and d["sms"] == 148 and d["tflops"] == 2250.0
    assert 170 << 30 < d["memory"] < 200 << 30
    assert d["props"]["computeCapability"] == "10.0"                       # coresPerSM lookup, legacy.go:690-713
    assert d["props"]["totalComputeUnits"] == "148"                        # worker/controller.go:545-549
    assert d["node"].startswith("/dev/nvidia") and d["index"] == 0
    assert d["caps"]["supportsSoftIsolation"] and d["caps"]["supportsRemoting"] and d["caps"]["maxWorkersPerDevice"] == 16
    assert all(l != b"FATAL" for l, _ in logs)


def test_topology_all_peers_tier0(prov):
    P, lib, _ = prov
    topo = P.ExtendedDeviceTopology()
    assert lib.AccelGetAllDevicesTopology(C.byref(topo)) == P.SUCCESS
    n = topo.deviceCount
    assert n >= 1
    for i in range(n):
        t = topo.devices[i]
        assert t.peerCount == n - 1
        for j in range(t.peerCount):
            assert t.peers[j].topoLevel == 0           # TOPO_LEVEL_INTERNAL over NVSwitch


def test_metrics_and_processes(prov):
    import numpy as np
    P, lib, _ = prov
    from tensor_fusion_b200.worker import Worker
    _, devs = P.all_devices(lib)
    with Worker() as w:                                 # a CUDA context with ~0.6 GiB so this pid shows up
        p = w.dev_alloc(512 << 20)
        w.move_batch([(p, 0, 512 << 20, 7)])
        uu = (C.c_char_p * 1)(devs[0]["uuid"].lower().encode())      # Go passes native case but be lenient
        dm = (P.DeviceMetrics * 1)()
        assert lib.AccelGetDeviceMetrics(uu, 1, dm) == P.SUCCESS
        m = dm[0]
        assert 10 < m.powerUsageWatts < 1200 and 10 < m.temperatureCelsius < 100
        assert m.utilizationPercent <= 100 and m.memoryUsedBytes > 512 << 20
        keys = [m.extraMetrics[i].key.decode() for i in range(m.extraMetricsCount)]
        assert "clockSMMHz" in keys
        pi = (P.ProcessInformation * 1024)()
        n = C.c_size_t()
        assert lib.AccelGetProcessInformation(pi, 1024, C.byref(n)) == P.SUCCESS
        mine = [pi[i] for i in range(n.value) if pi[i].processId.decode() == str(os.getpid())]
        assert mine and mine[0].memoryUsedBytes > 512 << 20 and mine[0].totalSMs == 148
        assert mine[0].deviceUUID.decode() == devs[0]["uuid"]
        w.dev_free(p)
    # unknown device: zeroed row, still SUCCESS (reference falls back the same way)
    uu = (C.c_char_p * 1)(b"GPU-does-not-exist")
    assert lib.AccelGetDeviceMetrics(uu, 1, dm) == P.SUCCESS and dm[0].powerUsageWatts == 0


def test_partition_and_hard_limits(prov):
    P, lib, _ = prov
    _, devs = P.all_devices(lib)
    uuid = devs[0]["uuid"].encode()
    pr = P.PartitionResult()
    assert lib.AccelAssignPartition(b"1g.10gb", uuid, C.byref(pr)) == P.SUCCESS
    env = [bytes(pr.envVars[i]).split(b"\0")[0].decode() for i in range(4)]
    assert env[0] == "NVIDIA_VISIBLE_DEVICES=" + devs[0]["uuid"]
    assert env[1] == "TF_CUDA_SM_PERCENT_LIMIT=15" and env[2] == "TF_CUDA_MEMORY_LIMIT=10240"
    part = pr.deviceUUID
    assert part.startswith(uuid[:40]) and pr.type == 0
    assert lib.AccelRemovePartition(b"1g.10gb", part) == P.SUCCESS
    assert lib.AccelAssignPartition(b"bogus", uuid, C.byref(pr)) == P.INVALID_PARAM
    assert lib.AccelAssignPartition(b"7g.900gb", uuid, C.byref(pr)) == P.RESOURCE_EXHAUSTED
    assert lib.AccelAssignPartition(b"1g.10gb", b"GPU-nope", C.byref(pr)) == P.NOT_FOUND
    assert lib.AccelSetMemHardLimit(uuid, 4 << 30) == P.SUCCESS
    assert lib.AccelSetComputeUnitHardLimit(uuid, 50) == P.SUCCESS
    assert lib.AccelSetMemHardLimit(b"GPU-nope", 1) == P.NOT_FOUND
    ctx = P.SnapshotContext()
    ctx.deviceUUID = uuid
    assert lib.AccelSnapshot(C.byref(ctx)) == P.SUCCESS        # device level, no vGPU worker on it: nothing to freeze
    me = (C.c_int32 * 1)(os.getpid())
    ctx = P.SnapshotContext(processIds=C.cast(me, C.POINTER(C.c_int32)), processCount=1, deviceUUID=None)
    assert lib.AccelSnapshot(C.byref(ctx)) == P.NOT_SUPPORTED  # a live process that is not one of this stack's workers


def test_reference_abi_suite_passes_against_our_library():
    """The reference's own 49-assertion known-answer test (provider/test/test_accelerator.c),
    compiled from the reference tree against libaccelerator_b200.so (oracle/Makefile)."""
    exe = os.path.join(conftest.ROOT, "oracle", "_ref", "test_accelerator_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/test_accelerator_b200 not built: `make -C oracle` needs /root/reference (build container only)")
    # the suite addresses a device called "stub-device-0": alias it to GPU 0
    env = dict(os.environ, TF_PROVIDER_DEVICE_ALIASES="stub-device-0=0")
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
    assert "Failed:       0" in r.stdout and r.returncode == 0, r.stdout[-3000:]
    total = int(r.stdout.split("Total tests:")[1].split()[0])
    assert total >= 49          # 49 + 3 more when a GPU process is visible


def test_reference_provider_baseline_still_passes():
    """Sanity: the unmodified reference provider + its suite, built from its own sources."""
    exe = os.path.join(conftest.ROOT, "oracle", "_ref", "test_accelerator_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, cwd=os.path.dirname(exe))
    assert r.returncode == 0 and "Failed:       0" in r.stdout


def test_worker_counters_reach_extra_metrics(prov, tmp_path):
    """what the worker staged / throttled shows up in AccelGetDeviceMetrics.extraMetrics
    (and from there in the hypervisor's tf_gpu_usage line) with no Go change."""
    P, lib, _ = prov
    from tensor_fusion_b200 import trace
    from tensor_fusion_b200.worker import Worker
    _, devs = P.all_devices(lib)
    uuid = devs[0]["uuid"].encode()
    base = str(tmp_path / "shm")
    assert lib.LimiterInit(base.encode()) == P.SUCCESS
    cfg = (P.LimiterDeviceConfig * 1)()
    cfg[0].deviceIdx, cfg[0].deviceUUID, cfg[0].upLimit, cfg[0].memLimit = 0, uuid, 100, 1 << 40
    assert lib.LimiterCreateWorker(b"ns", b"metrics-pod", cfg, 1) == P.SUCCESS
    raw = trace.gen_c1(ncalls=300, error_permille=0)

    def extras():
        uu = (C.c_char_p * 1)(uuid)
        dm = (P.DeviceMetrics * 1)()
        assert lib.AccelGetDeviceMetrics(uu, 1, dm) == P.SUCCESS
        return {dm[0].extraMetrics[i].key.decode(): dm[0].extraMetrics[i].value for i in range(dm[0].extraMetricsCount)}

    before = extras()
    assert before["tfwWorkers"] == 0
    with Worker(shm_path=os.path.join(base, "ns", "metrics-pod", "shm")) as w:
        w.run(raw)
        st = w.stats()
        got = extras()
        assert got["tfwWorkers"] == 1
        assert got["tfwStagedPayloadBytesTotal"] == st["payload_bytes"] > 0
        assert got["tfwMoverLaunchesTotal"] == st["mover_launches"] and got["tfwClientLaunchesTotal"] == st["client_launches"]
        assert got["tfwVramBytes"] == st["vram_bytes"] and "computeThrottledCnt" in got
    assert os.path.exists(os.path.join(base, "ns", "metrics-pod", "tfw_stats"))
    lib.LimiterShutdown()


def test_compiled_hypervisor_harness_runs_the_purego_sequence():
    """tools/hypervisor_harness.c: dlopen(RTLD_NOW|RTLD_GLOBAL), the 14 mandatory symbols, log callback,
    AccelInit, discovery, the 2 Hz metric loops, shutdown -- against our library and the reference stub."""
    exe = os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "hypervisor_harness")
    ours = os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "libaccelerator_b200.so")
    r = subprocess.run([exe, ours, "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "vendor=NVIDIA" in r.stdout and "B200" in r.stdout and "sms=148" in r.stdout and "fatal=0" in r.stdout
    ref = os.path.join(conftest.ROOT, "oracle", "_ref", "libaccelerator_example.so")
    if not os.path.exists(ref):
        return
    r = subprocess.run([exe, ref, "1"], capture_output=True, text=True, timeout=120, cwd=os.path.dirname(ref))
    assert r.returncode == 0 and "vendor=STUB" in r.stdout and "devices=4" in r.stdout      # hypervisor_suite_test.go:213-218
