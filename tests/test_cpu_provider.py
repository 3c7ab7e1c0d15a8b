"""Provider / limiter C-ABI on the CPU: layout, exports, argument validation, and the
product's quota-file + ERL implementation diffed against the oracle restatement."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import conftest
import oracle
from oracle import lib as O

ROOT = conftest.ROOT


@pytest.fixture(scope="module")
def prov():
    from tensor_fusion_b200 import provider as P
    return P, P.load()


def test_struct_sizes_match_go_mirror(prov):
    P, _ = prov
    for t, n in P.EXPECTED_SIZES.items():
        assert C.sizeof(t) == n, t
    assert P.DeviceBasicInfo.maxTflops.offset == 456 and P.DeviceMetrics.utilizationPercent.offset == 96
    assert P.ExtendedDeviceTopology.deviceCount.offset == 300032      # header layout, not the Go mirror's 332800 (App. E-1)
    assert P.PartitionResult.deviceNodes.offset == 4164 and P.LimiterDeviceConfig.memLimit.offset == 72


def test_exports_every_symbol_of_both_reference_headers(prov):
    P, lib = prov
    txt = open(os.path.join(ROOT, "include", "tf_provider_abi.h")).read()
    declared = set(re.findall(r"TF_ABI_EXPORT AccelResult (\w+)\(", txt))
    assert len(declared) == 30 and declared == set(P.SIGS)
    for name in declared:
        assert hasattr(lib, name)
    assert len(P.MANDATORY) == 14                       # accelerator_unix.go:57-98
    assert hasattr(lib, "RegisterLogCallback")          # legacy fallback name, accelerator_unix.go:102-106


def test_argument_validation_matches_reference_tests(prov):
    """provider/test/test_accelerator.c: every NULL / 0 / >100 case answers INVALID_PARAM."""
    P, lib = prov
    n = C.c_size_t()
    assert lib.AccelGetDeviceCount(None) == P.INVALID_PARAM
    dev = (P.ExtendedDeviceInfo * 1)()
    assert lib.AccelGetAllDevices(None, 256, C.byref(n)) == P.INVALID_PARAM
    assert lib.AccelGetAllDevices(dev, 0, C.byref(n)) == P.INVALID_PARAM
    assert lib.AccelGetAllDevicesTopology(None) == P.INVALID_PARAM
    pr = P.PartitionResult()
    assert lib.AccelAssignPartition(None, b"d", C.byref(pr)) == P.INVALID_PARAM
    assert lib.AccelAssignPartition(b"1g.10gb", None, C.byref(pr)) == P.INVALID_PARAM
    assert lib.AccelAssignPartition(b"1g.10gb", b"d", None) == P.INVALID_PARAM
    assert lib.AccelAssignPartition(b"", b"d", C.byref(pr)) == P.INVALID_PARAM
    assert lib.AccelRemovePartition(None, b"d") == P.INVALID_PARAM and lib.AccelRemovePartition(b"t", None) == P.INVALID_PARAM
    assert lib.AccelSetMemHardLimit(None, 1 << 30) == P.INVALID_PARAM and lib.AccelSetMemHardLimit(b"d", 0) == P.INVALID_PARAM
    assert lib.AccelSetComputeUnitHardLimit(b"d", 150) == P.INVALID_PARAM
    assert lib.AccelSetComputeUnitHardLimit(None, 50) == P.INVALID_PARAM and lib.AccelSetComputeUnitHardLimit(b"d", 0) == P.INVALID_PARAM
    pi = (P.ProcessInformation * 1)()
    assert lib.AccelGetProcessInformation(None, 256, C.byref(n)) == P.INVALID_PARAM
    assert lib.AccelGetProcessInformation(pi, 0, C.byref(n)) == P.INVALID_PARAM
    dm = (P.DeviceMetrics * 1)()
    uu = (C.c_char_p * 1)(b"d")
    assert lib.AccelGetDeviceMetrics(None, 1, dm) == P.INVALID_PARAM and lib.AccelGetDeviceMetrics(uu, 0, dm) == P.INVALID_PARAM
    mp = (P.MountPath * 1)()
    assert lib.AccelGetVendorMountLibs(None, 64, C.byref(n)) == P.INVALID_PARAM
    assert lib.AccelGetVendorMountLibs(mp, 0, C.byref(n)) == P.INVALID_PARAM
    assert lib.AccelSnapshot(None) == P.INVALID_PARAM and lib.AccelResume(None) == P.INVALID_PARAM
    ctx = P.SnapshotContext()
    assert lib.AccelSnapshot(C.byref(ctx)) == P.INVALID_PARAM          # neither pids nor device
    assert lib.AccelRegisterLogCallback(P.LogCallback()) == P.SUCCESS  # NULL callback unregisters
    assert lib.CheckAndRecordMemoryOps(b"p", b"d", 0, None) == P.INVALID_PARAM
    assert lib.CheckAndRecordComputeOps(b"p", b"d", 1, None) == P.INVALID_PARAM
    assert lib.FreezeWorker(b"w", None) == P.INVALID_PARAM and lib.ResumeWorker(None, None) == P.INVALID_PARAM
    assert lib.LimiterInit(None) == P.INVALID_PARAM


@pytest.mark.skipif(conftest.HAS_GPU, reason="checks the no-driver behaviour")
def test_accel_init_fails_loudly_without_a_driver(prov):
    P, lib = prov
    seen = []
    cb = P.LogCallback(lambda lvl, msg: seen.append((lvl, msg)))
    lib.AccelRegisterLogCallback(cb)
    assert lib.AccelInit() == P.OPERATION_FAILED
    n = C.c_size_t()
    assert lib.AccelGetDeviceCount(C.byref(n)) == P.OPERATION_FAILED    # no fake devices, ever
    assert seen and seen[0][0] == b"ERROR" and all(l != b"FATAL" for l, _ in seen)
    lib.AccelRegisterLogCallback(P.LogCallback())


def _mk_cfg(P, rows):
    arr = (P.LimiterDeviceConfig * len(rows))()
    for i, (idx, uuid, up, mem, cores) in enumerate(rows):
        arr[i].deviceIdx, arr[i].deviceUUID, arr[i].upLimit, arr[i].memLimit, arr[i].totalCudaCores = idx, uuid, up, mem, cores
    return arr


def test_quota_file_bytes_equal_oracle_image(prov, tmp_path):
    P, lib = prov
    base = str(tmp_path / "shm")
    assert lib.LimiterInit(base.encode()) == P.SUCCESS
    rows = [(0, b"GPU-3f1c", 25, 40 << 30, 0), (3, b"x" * 64, 100, 1 << 40, 18944), (15, b"", 1, 1, 1)]
    assert lib.LimiterCreateWorker(b"ns-a", b"pod-1", _mk_cfg(P, rows), 3) == P.SUCCESS
    path = os.path.join(base, "ns-a", "pod-1", "shm")
    got = np.fromfile(path, dtype=np.uint8)
    assert got.nbytes == 35504
    now = int(got[0x890:0x898].view(np.uint64)[0])
    want = np.zeros(35504, dtype=np.uint8)
    ocfg = (oracle.DevCfg * 3)()
    for i, (idx, uuid, up, mem, cores) in enumerate(rows):
        ocfg[i].device_idx, ocfg[i].uuid, ocfg[i].up_limit, ocfg[i].mem_limit, ocfg[i].total_cuda_cores = idx, uuid, up, mem, cores
    assert O.tfo_shm_init_image(C.c_void_p(want.ctypes.data), ocfg, 3, now, os.getpid()) == 0
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]
    # the oracle (Go semantics) opens the product's file, and vice versa
    h = C.c_void_p()
    assert O.tfo_shm_open(base.encode(), b"ns-a", b"pod-1", C.byref(h)) == 0
    O.tfo_shm_close(h)
    assert O.tfo_shm_create(base.encode(), b"ns-b", b"pod-2", ocfg, 3, C.byref(h)) == 0
    assert lib.LimiterRegisterPID(b"ns-b", b"pod-2", 4242) == P.SUCCESS
    out = (C.c_uint64 * 8)()
    assert O.tfo_shm_pid_values(O.tfo_shm_data(h), out, 8) == 1 and out[0] == 4242
    assert lib.LimiterSetPodMemoryUsed(b"ns-b", b"pod-2", 3, 777) == P.SUCCESS
    assert O.tfo_shm_pod_memory_used(O.tfo_shm_data(h), 3) == 777
    assert lib.LimiterSetPodMemoryUsed(b"ns-b", b"pod-2", 4, 1) == P.NOT_FOUND
    assert lib.LimiterUpdateHeartbeat(b"ns-b", b"pod-2", 123456) == P.SUCCESS
    assert O.tfo_shm_is_healthy(O.tfo_shm_data(h), 30, 123460) == 1
    O.tfo_shm_close(h)
    # path safety + legacy layout (soft_limiter_shm_test.go:197-270)
    assert lib.LimiterCreateWorker(b"../escape", b"pod", _mk_cfg(P, rows[:1]), 1) == P.INVALID_PARAM
    assert lib.LimiterRegisterPID(b"namespace", b"pod/name", 1) == P.INVALID_PARAM
    d = tmp_path / "shm" / "legacy" / "p"
    d.mkdir(parents=True)
    (d / "shm").write_bytes(b"\0" * 35496)
    assert lib.LimiterRegisterPID(b"legacy", b"p", 1) == P.OPERATION_FAILED
    assert lib.LimiterUpdateHeartbeat(b"nope", b"p", 1) == P.NOT_FOUND
    # remove prunes the empty directories but not the base
    assert lib.LimiterRemoveWorker(b"ns-a", b"pod-1") == P.SUCCESS
    assert not os.path.exists(os.path.join(base, "ns-a")) and os.path.isdir(base)
    assert lib.LimiterShutdown() == P.SUCCESS


def test_erl_steps_bit_identical_to_oracle(prov, tmp_path):
    """2 000 controller ticks with a wandering utilisation signal: rate, capacity,
    tokens and timestamp words of the product's file equal the oracle's, bit for bit."""
    P, lib = prov
    base = str(tmp_path / "erl")
    assert lib.LimiterInit(base.encode()) == P.SUCCESS
    assert lib.LimiterCreateWorker(b"n", b"p", _mk_cfg(P, [(2, b"GPU-z", 25, 1 << 30, 0)]), 1) == P.SUCCESS
    path = os.path.join(base, "n", "p", "shm")
    prod = np.memmap(path, dtype=np.uint8, mode="r")
    img = np.array(prod)                       # oracle works on a private copy of the same initial image
    f = C.c_void_p(img.ctypes.data)
    cfg = oracle.ErlCfg(); O.tfo_erl_default_cfg(C.byref(cfg))
    st = oracle.ErlState(); O.tfo_erl_new_state(C.byref(st))
    rng = np.random.default_rng(3)
    t0 = int(O.tfo_shm_get(f, 2, 3)) * 1_000_000
    util, up = 0.0, 25
    e = 8 + 2 * 136
    for k in range(2000):
        r = rng.random()
        util = float(np.clip(util + rng.normal(0, 12), 0, 100)) if r > 0.05 else float(rng.choice([0.0, 100.0, 24.9]))
        if k % 400 == 399:
            up = int(rng.choice([1, 25, 50, 100]))
        ts = t0 + (k + 1) * 500_000 + int(rng.integers(-20_000, 20_000))
        if k % 97 == 0:                        # the limiter consumes tokens between ticks
            cost = float(rng.random() * 50)
            O.tfo_shm_fetch_sub(f, 2, cost)
            w = np.memmap(path, dtype=np.uint8, mode="r+")
            cur = w[e + 112: e + 120].view(np.float64)
            if cur[0] >= cost:
                cur[0] = max(0.0, cur[0] - cost)
            w.flush(); del w
        assert lib.LimiterUpdateERL(b"n", b"p", 2, up, util, ts) == P.SUCCESS
        O.tfo_erl_tick(f, 2, C.byref(st), C.byref(cfg), up, util, ts / 1e6)
        assert np.array_equal(np.array(prod[e + 96: e + 128]), img[e + 96: e + 128]), f"tick {k}"
    assert lib.LimiterUpdateERL(b"n", b"p", 5, 25, 1.0, 1) == P.NOT_FOUND
    assert lib.LimiterUpdateERL(b"n", b"p", 2, 101, 1.0, 1) == P.INVALID_PARAM
    lib.LimiterShutdown()


def test_worker_facing_gate_matches_fetch_sub(tmp_path):
    """CheckAndRecordComputeOps / MemoryOps against a quota file named by TF_SHM_PATH
    (separate process: the path is read once)."""
    base = str(tmp_path)
    h = C.c_void_p()
    cfg = (oracle.DevCfg * 1)()
    cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 1, b"3F1C-AA", 50, 1000
    assert O.tfo_shm_create(base.encode(), b"ns", b"pod", cfg, 1, C.byref(h)) == 0
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from tensor_fusion_b200 import provider as P
lib = P.load()
rec = P.ComputeOpRecord()
out = []
for cost in (30, 30, 30, 30, 10, 1):
    rc = lib.CheckAndRecordComputeOps(b"1", b"GPU-3f1c-aa", cost, C.byref(rec))
    out.append((rc, rec.shouldBlock, rec.availableTokens))
m = P.MemoryOpRecord()
for diff in (600, 600, -600, 600):
    rc = lib.CheckAndRecordMemoryOps(b"1", b"3f1c-aa", diff, C.byref(m))
    out.append((rc, m.shouldBlock, m.availableBytes))
out.append(lib.CheckAndRecordComputeOps(b"1", b"GPU-other", 1, C.byref(rec)))
f = P.WorkerFreezeState()
out.append((lib.FreezeWorker(b"w1", C.byref(f)), f.isFrozen, f.freezeTimeMs > 0))
out.append((lib.ResumeWorker(b"w1", C.byref(f)), f.isFrozen, f.freezeTimeMs))
print(out)
''' % ROOT
    env = dict(os.environ, TF_SHM_PATH=os.path.join(base, "ns", "pod", "shm"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = eval(r.stdout)
    # tokens start at 100: 30,30,30 admitted (100->70->40->10), 4th denied (10 < 30), then 10 admitted, then 1 denied
    assert out[:6] == [(0, False, 70), (0, False, 40), (0, False, 10), (0, True, 10), (0, False, 0), (0, True, 0)]
    assert out[6:10] == [(0, False, 400), (0, True, 400), (0, False, 1000), (0, False, 400)]
    assert out[10] == 2
    assert out[11] == (0, True, True) and out[12] == (0, False, 0)
    assert O.tfo_shm_get(O.tfo_shm_data(h), 1, 2) == 0.0
    O.tfo_shm_close(h)


def test_snapshot_and_resume_reach_the_worker_through_the_stats_record(prov, tmp_path, monkeypatch):
    """AccelSnapshot / AccelResume (provider/accelerator.h:364-390): the command travels through the control
    words of the worker's stats record (include/tfw_stats_file.h); here a thread plays the worker's
    tfw_worker_poll_control."""
    import mmap
    import threading
    import time
    P, lib = prov
    base = tmp_path / "shm"
    pod = base / "ns" / "pod-a"
    pod.mkdir(parents=True)
    monkeypatch.setenv("TF_SHM_BASE_PATH", str(base))
    assert lib.LimiterShutdown() == P.SUCCESS  # the base of an earlier LimiterInit would win over the env
    f = open(pod / "tfw_stats", "w+b")
    f.truncate(C.sizeof(P.TfwStatsRecord))
    mm = mmap.mmap(f.fileno(), C.sizeof(P.TfwStatsRecord))
    rec = P.TfwStatsRecord.from_buffer(mm)
    rec.magic, rec.version, rec.pid, rec.updated_unix_secs = P.TFW_STATS_MAGIC, P.TFW_STATS_VERSION, 4_000_000_000, int(time.time())
    rec.device_uuid = b"GPU-0a0b0c0d-0000-1111-2222-333344445555"
    seen, stop, status = [], threading.Event(), [0]

    def worker():
        last = 0
        while not stop.is_set():
            req = rec.ctl_request
            if req != last:
                last = req
                seen.append(req)
                rec.ctl_status = status[0]
                rec.ctl_frozen = 1 if (req & 0xff) == P.TFW_CTL_FREEZE and status[0] == 0 else 0
                rec.ctl_ack = req
            time.sleep(0.001)

    th = threading.Thread(target=worker)
    th.start()
    try:
        me = (C.c_int * 1)(os.getpid())
        ctx = P.SnapshotContext(processIds=C.cast(me, C.POINTER(C.c_int)), processCount=1, deviceUUID=None)
        # a live PID that is no vGPU worker of this stack
        assert lib.AccelSnapshot(C.byref(ctx)) == P.NOT_SUPPORTED and seen == []
        # same PID namespace: the record names the worker
        rec.pid = os.getpid()
        assert lib.AccelSnapshot(C.byref(ctx)) == P.SUCCESS and seen == [(1 << 8) | P.TFW_CTL_FREEZE] and rec.ctl_frozen == 1
        assert lib.AccelResume(C.byref(ctx)) == P.SUCCESS and seen[-1] == (2 << 8) | P.TFW_CTL_RESUME and rec.ctl_frozen == 0
        # containers: the record holds the container PID, the pod's quota file the host PID (legacy.go:576)
        rec.pid = 4_000_000_000
        h = C.c_void_p()
        cfg = (oracle.DevCfg * 1)()
        cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 0, b"GPU-0a0b0c0d-0000-1111-2222-333344445555", 50, 1 << 30
        assert O.tfo_shm_create(str(base).encode(), b"ns", b"pod-a", cfg, 1, C.byref(h)) == 0
        assert lib.AccelSnapshot(C.byref(ctx)) == P.NOT_SUPPORTED
        assert O.tfo_shm_pid_insert(O.tfo_shm_data(h), os.getpid()) == 1
        assert lib.AccelSnapshot(C.byref(ctx)) == P.SUCCESS and seen[-1] == (3 << 8) | P.TFW_CTL_FREEZE
        # device level: every worker on that GPU
        dctx = P.SnapshotContext(processIds=None, processCount=0, deviceUUID=b"gpu-0A0B0C0D-0000-1111-2222-333344445555")
        assert lib.AccelResume(C.byref(dctx)) == P.SUCCESS and seen[-1] == (4 << 8) | P.TFW_CTL_RESUME
        other = P.SnapshotContext(processIds=None, processCount=0, deviceUUID=b"GPU-ffffffff-0000-0000-0000-000000000000")
        assert lib.AccelSnapshot(C.byref(other)) == P.SUCCESS and len(seen) == 4   # an idle GPU has nothing to freeze
        # the worker's verdict comes back: no HBM to resume into -> RESOURCE_EXHAUSTED
        status[0] = 4
        assert lib.AccelResume(C.byref(dctx)) == P.RESOURCE_EXHAUSTED
        status[0] = 0
        # FreezeWorker / ResumeWorker / AutoFreeze / AutoResume (provider/limiter.h:77-81) name the worker, and travel
        # the same control words: by "<namespace>/<pod>", by pod name, or by the id the worker published
        fs = P.WorkerFreezeState()
        n0 = len(seen)
        assert lib.FreezeWorker(b"ns/pod-a", C.byref(fs)) == P.SUCCESS and seen[-1] & 0xff == P.TFW_CTL_FREEZE and len(seen) == n0 + 1
        assert fs.isFrozen and fs.freezeTimeMs > 0 and fs.workerId == b"ns/pod-a"
        assert lib.ResumeWorker(b"pod-a", C.byref(fs)) == P.SUCCESS and seen[-1] & 0xff == P.TFW_CTL_RESUME and not fs.isFrozen
        rec.worker_id = b"8d0c7a1e-uid"
        assert lib.AutoFreeze(b"8d0c7a1e-uid", b"GPU-0a0b0c0d-0000-1111-2222-333344445555", b"compute") == P.SUCCESS and seen[-1] & 0xff == P.TFW_CTL_FREEZE
        status[0] = 4
        assert lib.ResumeWorker(b"8d0c7a1e-uid", C.byref(fs)) == P.RESOURCE_EXHAUSTED and len(seen) == n0 + 4
        status[0] = 0
        assert lib.AutoResume(b"8d0c7a1e-uid", b"GPU-0a0b0c0d-0000-1111-2222-333344445555", b"compute") == P.SUCCESS and rec.ctl_frozen == 0
        # a worker nobody publishes a record for: remembered only, as the reference stub does (accelerator.c:206-256)
        assert lib.FreezeWorker(b"somebody-else", C.byref(fs)) == P.SUCCESS and fs.isFrozen and len(seen) == n0 + 5
        # a worker that does not answer
        stop.set()
        th.join()
        monkeypatch.setenv("TF_SNAPSHOT_TIMEOUT_MS", "150")
        t0 = time.time()
        assert lib.AccelSnapshot(C.byref(dctx)) == P.OPERATION_FAILED and 0.1 < time.time() - t0 < 5
        # a dead worker's record (stale) is not a target
        rec.updated_unix_secs = int(time.time()) - 3600
        assert lib.AccelSnapshot(C.byref(dctx)) == P.SUCCESS
        O.tfo_shm_close(h)
    finally:
        stop.set()
        th.join()
        del rec
        mm.close()
        f.close()


def test_token_bucket_conserves_tokens_across_processes(tmp_path):
    """FetchSubERLTokens / FetchAddERLTokens are CAS loops on a float64 stored in a u64 of a file shared by
    the hypervisor and every process of the pod (soft_limiter_shm.go:715-748).  Four processes charge the
    bucket through the product's CheckAndRecordComputeOps while this one refills it through the oracle:
    tokens granted + tokens left == initial + tokens added, exactly (all amounts are integers < 2^53)."""
    import threading
    base = str(tmp_path)
    h = C.c_void_p()
    cfg = (oracle.DevCfg * 1)()
    cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 3, b"GPU-stress", 50, 1 << 30
    assert O.tfo_shm_create(base.encode(), b"ns", b"pod", cfg, 1, C.byref(h)) == 0
    f = O.tfo_shm_data(h)
    O.tfo_shm_set(f, 3, 1, 1e15)       # capacity far away: no refill is clamped
    O.tfo_shm_set(f, 3, 2, 5000.0)
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from tensor_fusion_b200 import provider as P
lib = P.load()
rec = P.ComputeOpRecord()
granted = denied = 0
cost = int(sys.argv[1])
for i in range(60000):
    assert lib.CheckAndRecordComputeOps(b"1", b"GPU-stress", cost, C.byref(rec)) == 0
    if rec.shouldBlock: denied += 1
    else: granted += cost
print(granted, denied)
''' % ROOT
    env = dict(os.environ, TF_SHM_PATH=os.path.join(base, "ns", "pod", "shm"))
    procs = [subprocess.Popen([sys.executable, "-c", code, str(cost)], env=env, stdout=subprocess.PIPE, text=True) for cost in (1, 3, 7, 50)]
    added, stop = [0], threading.Event()

    def refill():
        while not stop.is_set():
            O.tfo_shm_fetch_add(f, 3, 977.0)
            added[0] += 977
            time.sleep(0.0005)

    import time
    th = threading.Thread(target=refill)
    th.start()
    outs = [p.communicate(timeout=300)[0].split() for p in procs]
    stop.set()
    th.join()
    assert all(p.returncode == 0 for p in procs)
    granted = sum(int(o[0]) for o in outs)
    denied = sum(int(o[1]) for o in outs)
    left = O.tfo_shm_get(f, 3, 2)
    assert granted + left == 5000 + added[0], (granted, left, added[0])
    assert granted > 0 and denied > 0          # both outcomes were exercised
    O.tfo_shm_close(h)


def test_pid_set_survives_concurrent_registration(tmp_path):
    """The PID set of the quota file is guarded by a spin lock whose owner is a PID (soft_limiter_shm.go:791-840).
    Four processes register 300 PIDs each through the product (AddWorkerProcess) while the oracle's
    implementation inserts 300 more from here: the set ends up with exactly the 1 500 distinct values."""
    base = str(tmp_path)
    h = C.c_void_p()
    cfg = (oracle.DevCfg * 1)()
    cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 0, b"GPU-pids", 50, 1 << 30
    assert O.tfo_shm_create(base.encode(), b"ns", b"pod", cfg, 1, C.byref(h)) == 0
    f = O.tfo_shm_data(h)
    code = r'''
import sys
sys.path.insert(0, %r)
from tensor_fusion_b200 import provider as P
lib = P.load()
k = int(sys.argv[1])
for i in range(300):
    assert lib.AddWorkerProcess(b"GPU-pids", str(100000 * k + i).encode()) == 0
    assert lib.AddWorkerProcess(b"GPU-pids", str(100000 * k + i).encode()) == 0   # InsertIfAbsent: idempotent
''' % ROOT
    env = dict(os.environ, TF_SHM_PATH=os.path.join(base, "ns", "pod", "shm"))
    procs = [subprocess.Popen([sys.executable, "-c", code, str(k)], env=env) for k in (1, 2, 3, 4)]
    for i in range(300):
        O.tfo_shm_pid_insert(f, 900000 + i)
    assert all(p.wait(timeout=300) == 0 for p in procs)
    out = (C.c_uint64 * 4096)()
    n = O.tfo_shm_pid_values(f, out, 4096)
    got = sorted(out[i] for i in range(n))
    want = sorted([100000 * k + i for k in (1, 2, 3, 4) for i in range(300)] + [900000 + i for i in range(300)])
    assert got == want
    O.tfo_shm_close(h)


def test_remove_worker_cleans_up_empty_parents_like_the_reference(prov, tmp_path):
    """Mirrors TestSharedMemoryHandleCleanup and the three TestCleanupEmptyParentDirectories* cases
    (soft_limiter_shm_test.go:413-522, 683-702) through LimiterRemoveWorker: the quota file goes, empty
    pod / namespace directories go, the base (stop-at path) stays, a non-empty directory stops the walk."""
    P, lib = prov
    base = tmp_path / "shm"
    base.mkdir()
    assert lib.LimiterInit(str(base).encode()) == P.SUCCESS
    rows = [(0, b"GPU-c", 50, 1 << 30, 1024)]
    # 1. empty parents: pod and namespace directories disappear, the base remains
    assert lib.LimiterCreateWorker(b"test-namespace", b"test-pod", _mk_cfg(P, rows), 1) == P.SUCCESS
    shm = base / "test-namespace" / "test-pod" / "shm"
    assert shm.exists()
    assert lib.LimiterRemoveWorker(b"test-namespace", b"test-pod") == P.SUCCESS
    assert not shm.exists() and not (base / "test-namespace").exists() and base.exists()
    # 1b. the worker's own tfw_stats record (include/tfw_stats_file.h) does not keep the directory alive
    assert lib.LimiterCreateWorker(b"test-namespace", b"test-pod", _mk_cfg(P, rows), 1) == P.SUCCESS
    (base / "test-namespace" / "test-pod" / "tfw_stats").write_bytes(b"\0" * 280)
    assert lib.LimiterRemoveWorker(b"test-namespace", b"test-pod") == P.SUCCESS
    assert not (base / "test-namespace").exists() and base.exists()
    # 2. a second pod in the namespace keeps the namespace directory
    assert lib.LimiterCreateWorker(b"ns2", b"pod-a", _mk_cfg(P, rows), 1) == P.SUCCESS
    assert lib.LimiterCreateWorker(b"ns2", b"pod-b", _mk_cfg(P, rows), 1) == P.SUCCESS
    assert lib.LimiterRemoveWorker(b"ns2", b"pod-a") == P.SUCCESS
    assert not (base / "ns2" / "pod-a").exists() and (base / "ns2" / "pod-b" / "shm").exists()
    # 3. another file in the pod directory (this repo's own tfw_stats record, for one) stops the walk at once
    (base / "ns2" / "pod-b" / "other_file").write_bytes(b"other data")
    assert lib.LimiterRemoveWorker(b"ns2", b"pod-b") == P.SUCCESS
    assert not (base / "ns2" / "pod-b" / "shm").exists() and (base / "ns2" / "pod-b" / "other_file").exists()
    # removing what is not there: NOT_FOUND, nothing else touched
    assert lib.LimiterRemoveWorker(b"ns2", b"pod-a") == P.NOT_FOUND
    assert base.exists()
    lib.LimiterShutdown()


def test_device_entries_are_addressed_by_index_not_position(prov, tmp_path):
    """TestDeviceIterationMethods (soft_limiter_shm_test.go:315-360): configs for device 0 and device 2 activate
    exactly those two of the 16 entries; the ERL step of an inactive index answers NOT_FOUND."""
    P, lib = prov
    base = str(tmp_path)
    assert lib.LimiterInit(base.encode()) == P.SUCCESS
    rows = [(0, b"device-0", 80, 1 << 30, 1024), (2, b"device-2", 70, 2 << 30, 2048)]
    assert lib.LimiterCreateWorker(b"ns", b"iter", _mk_cfg(P, rows), 2) == P.SUCCESS
    h = C.c_void_p()
    assert O.tfo_shm_open(base.encode(), b"ns", b"iter", C.byref(h)) == 0
    f = O.tfo_shm_data(h)
    assert [i for i in range(16) if O.tfo_shm_has_device(f, i)] == [0, 2]
    assert lib.LimiterUpdateERL(b"ns", b"iter", 0, 80, 10.0, 1_000_000) == P.SUCCESS
    assert lib.LimiterUpdateERL(b"ns", b"iter", 2, 70, 10.0, 1_000_000) == P.SUCCESS
    assert lib.LimiterUpdateERL(b"ns", b"iter", 1, 70, 10.0, 1_000_000) == P.NOT_FOUND
    assert lib.LimiterUpdateERL(b"ns", b"iter", 16, 70, 10.0, 1_000_000) == P.INVALID_PARAM
    assert lib.LimiterSetPodMemoryUsed(b"ns", b"iter", 2, 12345) == P.SUCCESS and O.tfo_shm_pod_memory_used(f, 2) == 12345
    assert lib.LimiterSetPodMemoryUsed(b"ns", b"iter", 1, 1) == P.NOT_FOUND
    O.tfo_shm_close(h)
    lib.LimiterShutdown()


def test_compute_up_limit_of_the_product_equals_the_oracle(prov):
    """LimiterComputeUpLimit (extension of include/tf_provider_abi.h) restates computeUpLimit
    (worker/controller.go:307-325): the reference's own cases, then a sweep against the oracle's restatement."""
    import numpy as np
    P, lib = prov
    f = lib.LimiterComputeUpLimit
    f.restype, f.argtypes = C.c_uint32, [C.c_int64, C.c_double, C.c_double]
    assert f(25, 0, 2250) == 25 and f(0, 562.5, 2250) == 25 and f(0, 563, 2250) == 26 and f(0, 1, 2250) == 1
    assert f(0, 9999, 2250) == 100 and f(0, 0, 2250) == 100 and f(0, 100, 0) == 100 and f(40, 1000, 2250) == 40
    rng = np.random.default_rng(11)
    for _ in range(2000):
        cp = int(rng.integers(0, 3)) * int(rng.integers(0, 101))
        tf = float(rng.choice([0.0, rng.uniform(0, 3000), rng.uniform(0, 30)]))
        mx = float(rng.choice([0.0, 2250.0, 989.0, rng.uniform(1, 3000)]))
        assert f(cp, tf, mx) == O.tfo_compute_up_limit(cp, tf, mx), (cp, tf, mx)
