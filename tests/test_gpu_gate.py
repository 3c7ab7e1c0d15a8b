"""Device-resident token bucket vs the CPU restatement of FetchSub/FetchAddERLTokens."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.float64(x).view(np.uint64)


def test_reference_golden_vectors_on_device():
    """soft_limiter_shm_test.go:139-151 and :624-638, executed by the gate kernels."""
    from tensor_fusion_b200.gate import Gate
    g = Gate()
    g.set_tokens(1.5)
    before, ok = g.try_acquire(2.0)
    assert (before, ok) == (1.5, False) and g.state()["tokens"] == 1.5
    g.set_tokens(5.0)
    before, ok = g.try_acquire(2.0)
    assert (before, ok) == (5.0, True) and g.state()["tokens"] == 3.0
    g.set_capacity(100.0)
    g.set_tokens(50.0)
    assert g.refill(30.0) == 50.0 and g.state()["tokens"] == 80.0
    assert g.refill(50.0) == 80.0 and g.state()["tokens"] == 100.0
    g.close()


def test_sequence_bit_exact_vs_oracle():
    """A recorded 20k-op sequence of refills / gate requests with awkward float
    costs: every pre-value and the final token word must be bit-identical."""
    import oracle
    from tensor_fusion_b200.gate import Gate
    rng = np.random.default_rng(11)
    img = np.zeros(oracle.lib.tfo_shm_file_bytes(), dtype=np.uint8)
    cfg = (oracle.DevCfg * 1)()
    cfg[0].device_idx, cfg[0].uuid = 0, b"GPU-x"
    oracle.lib.tfo_shm_init_image(C.c_void_p(img.ctypes.data), cfg, 1, 1000, 1)
    f = C.c_void_p(img.ctypes.data)
    ops, want = [], []
    for _ in range(20_000):
        r = rng.random()
        if r < 0.6:
            amt = float(rng.choice([0.1, 1.0, 1e-9, 3.3333333333333335, 17.25, 1e6])) * float(rng.random() + 0.01)
            ops.append((0, amt)); want.append(oracle.lib.tfo_shm_fetch_sub(f, 0, amt))
        elif r < 0.95:
            amt = float(rng.random() * 40.0)
            ops.append((1, amt)); want.append(oracle.lib.tfo_shm_fetch_add(f, 0, amt))
        else:
            cap = float(rng.choice([50.0, 100.0, 1234.5]))
            ops.append((2, cap)); want.append(oracle.lib.tfo_shm_get(f, 0, 1)); oracle.lib.tfo_shm_set(f, 0, 1, cap)
    g = Gate()
    got = g.run_sequence(ops)
    assert np.array_equal(_bits(np.array(got)), _bits(np.array(want)))
    assert _bits(g.state()["tokens"]) == _bits(oracle.lib.tfo_shm_get(f, 0, 2))
    g.close()


def test_contended_cas_conserves_tokens():
    """592 CTAs hammer the same bucket: admitted*cost + left == initial, exactly."""
    from tensor_fusion_b200.gate import Gate
    g = Gate()
    g.set_capacity(1e9)
    g.set_tokens(100_000.0)
    admitted = g.contend(592, 500, 1.0)
    st = g.state()
    assert admitted == 100_000 and st["tokens"] == 0.0
    g.set_tokens(50_000.0)
    admitted = g.contend(148, 100, 3.0)
    st = g.state()
    assert admitted * 3.0 + st["tokens"] == 50_000.0 and admitted == 148 * 100
    g.close()


def test_blocking_gate_orders_the_stream():
    """A launch behind an empty bucket starts only after the refill arrives."""
    import time
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.gate import Gate
    from tensor_fusion_b200.worker import Worker
    with Worker() as w:
        g = Gate()
        g.set_tokens(0.0)
        from tensor_fusion_b200._native import lib
        s = lib.tfw_exec_stream(w.h)
        b = wire.Builder()
        b.malloc(1, 4096).memset(1, 0, 4096, 1)
        w.run(bytes(b))
        g.enqueue(10.0, s)                                  # blocks the exec stream
        w.submit(bytes(wire.Builder().launch(wire.K_ADD_U8, h=1, n=4096, scalar=1)))
        time.sleep(0.2)
        assert g.state()["admitted"] == 0                   # still waiting
        g.refill(25.0)
        assert np.all(w.read(1) == 2)                       # flush returns => kernel ran after the gate
        st = g.state()
        assert st["admitted"] == 1 and st["blocked_gates"] == 1 and st["tokens"] == 15.0 and st["timeouts"] == 0
        g.close()


def test_quota_file_bridge_enforces_the_hypervisor_rate(tmp_path):
    """End to end (b): the hypervisor side (played here through the oracle's Go-semantics file
    operations) refills the quota file at `rate`; the worker's LAUNCH frames carry a token cost;
    the device-resident gate admits them at that rate and every token is accounted for."""
    import threading
    import time
    import oracle
    from oracle import lib as O
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    base = str(tmp_path)
    h = C.c_void_p()
    cfg = (oracle.DevCfg * 1)()
    cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 0, b"GPU-test", 25, 1 << 40
    assert O.tfo_shm_create(base.encode(), b"ns", b"pod", cfg, 1, C.byref(h)) == 0
    f = O.tfo_shm_data(h)
    rate, cost, launches = 2000.0, 100, 30
    O.tfo_shm_set(f, 0, 0, rate)            # refill rate
    O.tfo_shm_set(f, 0, 1, 400.0)           # capacity
    O.tfo_shm_set(f, 0, 2, 200.0)           # current tokens
    stop = threading.Event()
    added = [200.0]

    def hypervisor():                        # rebalanceTokenBucket's refill (quota_controller.go:355-360) + heartbeat
        last = time.time()
        while not stop.is_set():
            time.sleep(0.05)
            now = time.time()
            before = O.tfo_shm_fetch_add(f, 0, rate * (now - last))
            after = min(400.0, before + rate * (now - last))
            added[0] += after - before
            last = now
            img = np.ctypeslib.as_array((C.c_uint8 * 35504).from_address(f))
            img[0x890:0x898].view(np.uint64)[0] = int(now)

    th = threading.Thread(target=hypervisor)
    th.start()
    try:
        with Worker(shm_path=os.path.join(base, "ns", "pod", "shm"), shm_device_index=0) as w:
            b = wire.Builder()
            b.malloc(1, 4096)
            for _ in range(launches):
                b.launch(wire.K_ADD_U8, h=1, n=4096, scalar=1, cost=cost)
            t0 = time.time()
            _, resp = w.run(bytes(b.sync()))
            dt = time.time() - t0
            assert np.all(w.read(1) == launches)                 # every launch ran, in order
            st = w.stats()
            assert st["gate_launches"] == launches
        stop.set()
        th.join()
        need = launches * cost - 200.0                            # tokens that had to be refilled
        # rate-limited (not instantaneous), and not stuck on the 5 s fail-open either; generous bounds: the
        # file may have refilled to its capacity (400) while the worker's CUDA context was being created
        assert (launches * cost - 400.0) / rate * 0.6 < dt < need / rate * 3.0 + 2.0, (dt, added[0])
        # conservation: what the hypervisor put in == what the launches consumed + what is left in the file
        left = O.tfo_shm_get(f, 0, 2)
        spill = added[0] - launches * cost - left
        if left >= 400.0 - 1e-9:
            # the bucket was full when the worker handed its unspent prepaid window back: FetchAddERLTokens clamps at
            # the capacity (soft_limiter_shm.go:734-748), so up to one window may spill -- never the other way round
            assert -1e-3 <= spill <= 400.0, (added[0], left)
        else:
            assert abs(spill) < 1e-6 * added[0] + 1e-3, (added[0], left)
    finally:
        stop.set()
        th.join()
        O.tfo_shm_close(h)


def test_a_multi_device_worker_charges_the_bucket_of_its_own_device_index(tmp_path):
    """A worker that spans several GPUs gets one DeviceEntry -- one token bucket -- per device index in the same quota
    file (soft_limiter_shm.go:21,201-205; AllocateWorkerDevices, worker/allocation.go:46-139).  The vGPU session bound
    to index 1 takes its tokens from entry 1 and leaves entry 0 alone."""
    import oracle
    from oracle import lib as O
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    base = str(tmp_path)
    h = C.c_void_p()
    cfg = (oracle.DevCfg * 2)()
    cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 0, b"GPU-aaa", 50, 1 << 40
    cfg[1].device_idx, cfg[1].uuid, cfg[1].up_limit, cfg[1].mem_limit = 1, b"GPU-bbb", 50, 1 << 40
    assert O.tfo_shm_create(base.encode(), b"ns", b"pod2", cfg, 2, C.byref(h)) == 0
    f = O.tfo_shm_data(h)
    for idx in (0, 1):
        O.tfo_shm_set(f, idx, 0, 1000.0)        # rate
        O.tfo_shm_set(f, idx, 1, 5000.0)        # capacity
        O.tfo_shm_set(f, idx, 2, 5000.0)        # tokens
    img = np.ctypeslib.as_array((C.c_uint8 * 35504).from_address(f))
    img[0x890:0x898].view(np.uint64)[0] = int(__import__("time").time())      # fresh heartbeat
    with Worker(shm_path=os.path.join(base, "ns", "pod2", "shm"), shm_device_index=1) as w:
        b = wire.Builder().malloc(1, 4096)
        for _ in range(20):
            b.launch(wire.K_ADD_U8, grid=4, block=64, h=1, n=4096, scalar=1)   # no cost on the wire: the worker computes 4 x 2 = 8
        w.run(bytes(b.sync()))
        assert np.all(w.read(1) == 20)
        st = w.gate_state()
        assert st["admitted"] == 20 and st["timeouts"] == 0
    assert O.tfo_shm_get(f, 0, 2) == 5000.0                               # device 0's bucket was never touched
    assert O.tfo_shm_get(f, 1, 2) == 5000.0 - 20 * 8                      # 20 launches x (4 blocks x 2 warps), unspent prepaid tokens handed back
    O.tfo_shm_close(h)
