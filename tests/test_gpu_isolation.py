"""Isolation that is enforced, not advertised (round-2 verdict item 4): the hard compute limit as an SM partition
(green context), hard limits reaching live workers from the provider ABI, the limiter's cost clamp and its
fail-open / fail-closed policy, the worker's idle auto-freeze, and plain-buffer freezes at PCIe speed."""
import ctypes as C
import os
import struct
import subprocess
import time

import numpy as np
import pytest

import conftest

pytestmark = pytest.mark.gpu
MOCK = os.path.join(conftest.ROOT, "build", "mock")


def _cubin():
    path = os.path.join(MOCK, "user_kernels.cubin")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "build/mock/user_kernels.cubin"], cwd=conftest.ROOT, check=True)
    return open(path, "rb").read()


def _smids(w, blocks=2048, spin_ns=300_000):
    """Which SMs a grid of `blocks` CTAs runs on inside worker `w` (tools/user_kernels.cu: where_am_i)."""
    from tensor_fusion_b200 import wire
    b = wire.Builder().malloc(1, 4 * blocks).module_load(1, _cubin()).get_function(1, 1, b"where_am_i")
    b.launch_user(1, (blocks,), (64,), struct.pack("<QQ", wire.tagged_ptr(1), spin_ns)).sync()
    n, resp = w.run(bytes(b))
    frames = list(wire.parse_frames(resp))
    assert [h["opcode"] for h, _ in frames] == [wire.OP_RESP_FUNCTION, wire.OP_RESP_SYNC], frames
    assert frames[0][0]["arg0"] == 2 and frames[0][0]["arg1"] == 16            # (uint32_t*, uint64_t)
    ids = np.frombuffer(w.read(1).tobytes(), dtype=np.uint32)
    w.run(bytes(wire.Builder().module_unload(1).free(1).sync()))
    return set(int(x) for x in ids)


def test_sm_percent_limit_confines_the_tenant_to_its_share_of_the_sms():
    """TF_CUDA_SM_PERCENT_LIMIT / tfw_config.sm_percent_limit (compose.go:1287-1295): every kernel of the vGPU runs
    inside a green context that owns ceil(25 % of 148) = 37 SMs (rounded up to the partition granularity), whatever
    the grid asks for; the unlimited vGPU spreads over the whole GPU; the limit can be changed while it runs."""
    from tensor_fusion_b200.worker import Worker
    from tensor_fusion_b200 import _native as N
    with Worker() as w:
        free = _smids(w)
    assert len(free) > 120, len(free)
    with Worker(sm_percent=25) as w:
        part = _smids(w)
        assert 8 <= len(part) <= 48, sorted(part)                      # 37, or the next multiple the hardware partitions in
        # data path and built-in kernels run inside the partition too, bit-exact
        import oracle
        from tensor_fusion_b200 import trace
        raw = trace.gen_c1(ncalls=300, seed=99)
        rep = oracle.Replay(raw)
        n, resp = w.run(raw)
        assert resp == rep.responses()
        for h in rep.live_handles():
            assert np.array_equal(w.read(h), rep.buffer(h))
        # tighten, then lift, the limit on the running vGPU
        N.check(N.lib.tfw_worker_set_sm_limit(w.h, 10), "set_sm_limit", w.h)
        tight = _smids(w)
        assert 1 <= len(tight) <= 24 and len(tight) < len(part), (len(tight), len(part))
    with Worker(sm_percent=50) as w:
        half = _smids(w)
        assert len(part) < len(half) <= 80, (len(part), len(half))
        N.check(N.lib.tfw_worker_set_sm_limit(w.h, 0), "set_sm_limit", w.h)
        assert len(_smids(w)) > 120


def test_hard_limits_of_the_provider_abi_reach_the_running_worker(tmp_path, monkeypatch):
    """AccelSetComputeUnitHardLimit / AccelSetMemHardLimit (provider/accelerator.h:343-358) travel through the control
    words of the worker's stats record: the SM partition and the MALLOC quota of the live vGPU change."""
    import threading
    from tensor_fusion_b200 import provider as P
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    from tensor_fusion_b200 import _native as N
    base = tmp_path / "shm"
    (base / "ns" / "pod").mkdir(parents=True)
    monkeypatch.setenv("TF_SHM_BASE_PATH", str(base))
    monkeypatch.setenv("TFW_STATS_PATH", str(base / "ns" / "pod" / "tfw_stats"))
    lib = P.load()
    lib.LimiterShutdown()                                              # a base path left by an earlier LimiterInit would win over the env
    assert lib.AccelInit() == P.SUCCESS
    rc, devs = P.all_devices(lib)
    uuid = devs[0]["uuid"].encode()
    with Worker() as w:
        stop = threading.Event()
        results = {}

        def hypervisor():
            results["sm"] = lib.AccelSetComputeUnitHardLimit(uuid, 25)
            results["mem"] = lib.AccelSetMemHardLimit(uuid, 64 << 20)
            stop.set()

        th = threading.Thread(target=hypervisor)
        th.start()
        while not stop.is_set():                                       # the worker's owning thread polls its control words
            N.lib.tfw_worker_poll_control(w.h, None)
            time.sleep(0.001)
        th.join()
        assert results == {"sm": P.SUCCESS, "mem": P.SUCCESS}
        assert 8 <= len(_smids(w)) <= 48
        _, resp = w.run(bytes(wire.Builder().malloc(7, 32 << 20).malloc(8, 48 << 20).sync()))
        codes = [(h["call_id"], h["arg0"]) for h, _ in wire.parse_frames(resp) if h["opcode"] == wire.OP_RESP_ERROR]
        assert codes == [(1, 4)]                                       # the second MALLOC breaks the 64 MiB quota
    lib.AccelShutdown()


def test_a_launch_larger_than_the_bucket_is_charged_one_bucket_not_the_fail_open_timer():
    """blocks x warps of one large grid exceeds any capacity the reference controller sets
    (200..200000); FetchSub never admits cost > capacity.  The gate clamps the cost to the capacity."""
    import torch
    from tensor_fusion_b200.gate import Gate
    g = Gate()
    try:
        g.set_capacity(200.0)
        g.set_tokens(200.0)
        st = torch.cuda.Stream()
        t0 = time.time()
        g.enqueue(16384 * 8, st.cuda_stream)                         # 131 072 tokens asked of a bucket that holds 200
        st.synchronize()
        assert time.time() - t0 < 2.0                                 # (the fail-open timer is 5 s)
        s = g.state()
        assert s["tokens"] == 0.0 and s["admitted"] == 1 and s["timeouts"] == 0
    finally:
        g.close()


def test_fail_open_and_fail_closed_are_a_policy():
    """A gate whose tokens never come: open (default) -> released by the watchdog and counted once; closed -> it waits
    for the refill, however long (TFW_GATE_FAIL_POLICY / TFW_F_GATE_FAIL_CLOSED)."""
    import torch
    from tensor_fusion_b200 import _native as N
    from tensor_fusion_b200.gate import Gate
    g = Gate()
    try:
        g.set_capacity(100.0)
        g.set_tokens(0.0)
        N.check(N.lib.tfw_gate_set_policy(g.h, 0, 300.0), "set_policy")
        st = torch.cuda.Stream()
        t0 = time.time()
        g.enqueue(50.0, st.cuda_stream)
        g.enqueue(10.0, st.cuda_stream)                               # a cheaper gate BEHIND the stuck one
        st.synchronize()
        dt = time.time() - t0
        s = g.state()
        assert 0.25 < dt < 4.0 and s["timeouts"] == 2 and s["admitted"] == 2, (dt, s)   # each released once, each counted once
        # closed: nothing releases it but a refill
        N.check(N.lib.tfw_gate_set_policy(g.h, 1, 200.0), "set_policy")
        g.set_tokens(0.0)
        g.enqueue(40.0, st.cuda_stream)
        time.sleep(1.0)
        assert not st.query()                                          # still waiting after 5x the bound
        g.refill(40.0)
        st.synchronize()
        assert g.state()["timeouts"] == 2
    finally:
        g.close()


def test_idle_vgpu_freezes_itself_and_the_next_byte_brings_it_back(tmp_path):
    """auto_freeze.freeze_to_mem_ttl (RemotePodInfo, api/http_types.go:82-100; here TF_AUTO_FREEZE_TTL_MS): a session
    that has been silent for the TTL gives its HBM back; the client's next call finds every byte where it was."""
    exe = os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "tensor-fusion-worker")
    lib = C.CDLL(os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so"))
    lib.tfc_connect.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.tfc_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.tfc_memcpy_h2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.tfc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_sync.argtypes = [C.c_void_p]
    lib.tfc_close.argtypes = [C.c_void_p]
    stats = tmp_path / "tfw_stats"
    env = dict(os.environ, TFW_ONESHOT="1", TFW_BIND="127.0.0.1", TF_ENABLE_LOG="1", TF_AUTO_FREEZE_TTL_MS="400", TFW_STATS_PATH=str(stats),
               POD_NAMESPACE="ns", POD_NAME="pod-x", TFW_NO_SHM_UPGRADE="1")     # the socket loop's idle policy is what is under test
    p = subprocess.Popen([exe, "-p", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, text=True)
    try:
        port = int(p.stdout.readline().split()[-1])
        c = C.c_void_p()
        assert lib.tfc_connect(f"native+127.0.0.1+{port}+x".encode(), C.byref(c)) == 0
        n = 96 << 20
        a = C.c_uint32()
        src = np.random.default_rng(3).integers(0, 256, n, dtype=np.uint8)
        assert lib.tfc_malloc(c, n, C.byref(a)) == 0 and lib.tfc_memcpy_h2d(c, a, 0, src.ctypes.data, n) == 0 and lib.tfc_sync(c) == 0
        from tensor_fusion_b200 import provider as P
        deadline = time.time() + 10
        rec = None
        while time.time() < deadline:                                  # the client goes quiet; the worker notices
            time.sleep(0.1)
            rec = P.TfwStatsRecord.from_buffer_copy(stats.read_bytes())
            if rec.ctl_frozen:
                break
        assert rec.ctl_frozen == 1 and rec.frozen_auto == 1 and rec.parked_bytes == n and rec.vram_bytes == 0, (rec.ctl_frozen, rec.parked_bytes)
        assert rec.worker_id == b"ns/pod-x" and rec.frozen_unix_ms > 0
        got = np.empty(n, dtype=np.uint8)
        assert lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n) == 0 and np.array_equal(got, src)   # thawed by the request itself
        rec = P.TfwStatsRecord.from_buffer_copy(stats.read_bytes())
        assert rec.ctl_frozen == 0 and rec.auto_freezes == 1 and rec.auto_resumes == 1
        lib.tfc_close(c)
        _, err = p.communicate(timeout=60)
        assert "frozen to memory" in err
    finally:
        if p.poll() is None:
            p.kill()


def test_plain_buffers_are_parked_at_pcie_speed_not_page_fault_speed():
    """Round 1 parked plain buffers with a cudaMemcpy into fresh pageable memory: 2.9 GB/s out, 4.2 GB/s back.  The
    bounce ring + copy pool (worker.cu: ParkPipe) takes the first-touch page faults on many cores at once."""
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    n, k = 256 << 20, 8
    rng = np.random.default_rng(12)
    data = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(2)]
    os.environ["TFW_LOG_PARK"] = "1"
    with Worker() as w:
        b = wire.Builder()
        for h in range(1, k + 1):
            b.malloc(h, n)
        w.run(bytes(b.sync()))
        for h in range(1, k + 1):
            w.run(bytes(wire.Builder().h2d(h, 0, data[h % 2].tobytes()).sync()))
        t0 = time.time()
        moved = w.freeze()
        t_out = time.time() - t0
        assert moved == k * n
        t0 = time.time()
        w.resume()
        t_in = time.time() - t0
        for h in (1, 2, k):
            assert np.array_equal(w.read(h), data[h % 2])
        out_gbps, in_gbps = moved / t_out / 1e9, moved / t_in / 1e9
        print(f"park {out_gbps:.1f} GB/s, unpark {in_gbps:.1f} GB/s")
        assert out_gbps > 12.0 and in_gbps > 8.0, (out_gbps, in_gbps)   # round 1: 2.9 / 4.2 GB/s; the figures measured on the B200 box are in DESIGN 7c
