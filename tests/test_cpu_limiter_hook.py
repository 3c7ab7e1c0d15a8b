"""libcuda_limiter.so (LD_PRELOAD limiter of local soft mode) on the CPU:
a counting stand-in for libcuda.so.1 (tools/mock_cuda.c) and an application that reaches it the way
libcudart does (tools/hook_probe.c).  The quota file is written by the oracle, the hook charges it
through the product's CheckAndRecord* implementation; the arithmetic expected here is the Go
FetchSub/FetchAdd semantics of soft_limiter_shm.go:715-748."""
import ctypes as C
import json
import os
import subprocess
import threading
import time

import pytest

import conftest
import oracle
from oracle import lib as O

ROOT = conftest.ROOT
MOCK = os.path.join(ROOT, "build", "mock")
HOOK = os.path.join(ROOT, "tensor-fusion_b200", "lib", "libcuda_limiter.so")
MOCK_UUID = b"GPU-10111213-1415-1617-1819-1a1b1c1d1e1f"  # tools/mock_cuda.c: cuDeviceGetUuid_v2
RATE, CAPACITY, TOKENS = 0, 1, 2


@pytest.fixture(scope="module", autouse=True)
def built():
    if not (os.path.exists(HOOK) and os.path.exists(os.path.join(MOCK, "libcuda.so.1")) and os.path.exists(os.path.join(MOCK, "hook_probe"))):
        subprocess.run(["make", "-s", HOOK.replace(ROOT + "/", ""), "build/mock/libcuda.so.1", "build/mock/hook_probe"], cwd=ROOT, check=True)


def quota(tmp_path, tokens=1000.0, capacity=1000.0, mem_limit=1 << 30, uuid=MOCK_UUID):
    h = C.c_void_p()
    cfg = (oracle.DevCfg * 1)()
    cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 0, uuid, 50, mem_limit
    assert O.tfo_shm_create(str(tmp_path).encode(), b"ns", b"pod", cfg, 1, C.byref(h)) == 0
    d = O.tfo_shm_data(h)
    O.tfo_shm_set(d, 0, CAPACITY, capacity)
    O.tfo_shm_set(d, 0, TOKENS, tokens)
    return h, d, os.path.join(str(tmp_path), "ns", "pod", "shm")


def probe(args, shm=None, preload=True, **env_extra):
    env = dict(os.environ, LD_LIBRARY_PATH=MOCK)
    env.pop("TF_SHM_PATH", None)
    env.pop("HYPERVISOR_IP", None)
    if preload:
        env["LD_PRELOAD"] = HOOK
    if shm:
        env["TF_SHM_PATH"] = shm
        env["TF_ISOLATION_MODE"] = "soft"
    env.update(env_extra)
    r = subprocess.run([os.path.join(MOCK, "hook_probe")] + [str(a) for a in args], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stderr)
    return json.loads(r.stdout), r.stderr


def test_hook_exports_the_interposed_entry_points_and_the_worker_facing_abi():
    lib = C.CDLL(HOOK)
    for name in ("dlsym", "cuGetProcAddress", "cuGetProcAddress_v2", "cuLaunchKernel", "cuLaunchKernel_ptsz", "cuLaunchKernelEx",
                 "cuLaunchCooperativeKernel", "cuMemAlloc_v2", "cuMemFree_v2", "cuMemAllocAsync", "cuMemFreeAsync", "cuMemCreate",
                 "cuMemRelease", "cuMemGetInfo_v2", "cuDeviceTotalMem_v2", "CheckAndRecordComputeOps", "CheckAndRecordMemoryOps",
                 "AddWorkerProcess", "FreezeWorker", "ResumeWorker", "AutoFreeze", "AutoResume"):
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", HOOK], capture_output=True, text=True).stdout
    assert "_ZN" not in out and "_ZS" not in out  # no C++ (libstdc++ is static and hidden): safe to preload anywhere
    assert "libstdc++" not in subprocess.run(["ldd", HOOK], capture_output=True, text=True).stdout


@pytest.mark.parametrize("mode", ["procaddr", "procaddr_ptsz", "dlsym"])
def test_forwarding_without_a_quota_file(mode):
    base, _ = probe([mode, 25, 4, 128, 4096], preload=False)
    out, _ = probe([mode, 25, 4, 128, 4096])
    assert out["hooked"] == 1 and out["hook_active"] == 0 and out["libc_ok"] == 1
    for k in ("launch_rc", "allocs", "free", "total", "driver_launches", "driver_launches_ptsz", "driver_allocs", "driver_frees"):
        assert out[k] == base[k], k
    assert out["driver_launches"] + out["driver_launches_ptsz"] == 25


@pytest.mark.parametrize("mode", ["procaddr", "procaddr_ptsz", "dlsym"])
def test_launches_are_charged_blocks_times_warps(tmp_path, mode):
    h, d, shm = quota(tmp_path, tokens=1000.0)
    out, err = probe([mode, 10, 6, 96], shm=shm, TF_LIMITER_LOG="1")  # 6 blocks x 3 warps = 18 tokens per launch
    assert out["hook_active"] == 1 and out["hook_launches"] == 10 and out["hook_blocked"] == 0, err
    assert out["hook_tokens"] == 180
    assert O.tfo_shm_get(d, 0, TOKENS) == 1000.0 - 180.0
    assert out["driver_launches" + ("_ptsz" if mode == "procaddr_ptsz" else "")] == 10  # each reached the right driver entry
    O.tfo_shm_close(h)


def test_short_bucket_blocks_until_the_refill(tmp_path):
    """40 tokens, 10 launches of 16: two pass, the third waits for a refill written like the
    hypervisor's rebalance does (FetchAddERLTokens); nothing is lost or minted."""
    h, d, shm = quota(tmp_path, tokens=40.0, capacity=200.0)
    added = []

    def refill():
        for _ in range(4):
            time.sleep(0.15)
            O.tfo_shm_fetch_add(d, 0, 40.0)
            added.append(40.0)

    t = threading.Thread(target=refill)
    t0 = time.time()
    t.start()
    out, _ = probe(["procaddr", 10, 4, 128], shm=shm)  # 4 x 4 = 16 tokens per launch, 160 in total
    t.join()
    assert out["launch_rc"] == 0 and out["driver_launches"] == 10 and out["hook_timeouts"] == 0
    assert out["hook_blocked"] >= 2
    assert out["launch_ms"] >= 400.0  # 160 tokens need the 3rd refill (40 + 3 x 40), which lands at ~450 ms
    assert time.time() - t0 < 20
    assert O.tfo_shm_get(d, 0, TOKENS) == 40.0 + sum(added) - 160.0
    O.tfo_shm_close(h)


def test_a_launch_larger_than_the_bucket_costs_one_full_bucket(tmp_path):
    """FetchSub never admits a cost above the capacity, and the reference controller keeps the capacity between 200 and
    200 000 tokens (quota_controller.go:425-433) while one large grid is worth more (16K blocks x 8 warps = 131 072):
    such a launch is charged one full bucket instead of waiting for the fail-open timer."""
    h, d, shm = quota(tmp_path, tokens=10.0, capacity=10.0)
    out, _ = probe(["dlsym", 1, 64, 1024], shm=shm, TF_LIMITER_MAX_WAIT_MS="3000")  # 2048 tokens asked, 10 is all there can ever be
    assert out["launch_rc"] == 0 and out["driver_launches"] == 1
    assert out["hook_timeouts"] == 0 and out["hook_blocked"] == 0 and out["launch_ms"] < 1000.0
    assert O.tfo_shm_get(d, 0, TOKENS) == 0.0  # the whole bucket, not nothing
    O.tfo_shm_close(h)


def test_large_grids_under_the_reference_controllers_bounds(tmp_path):
    """The real tferl::tick bounds: capacity starts at 200 (capacity_min) with rate 10/s.  Four launches of 16 384 blocks x
    8 warps each wait for a full bucket (refilled here as the hypervisor would), none falls through the fail-open timer."""
    import threading
    import time
    h, d, shm = quota(tmp_path, tokens=200.0, capacity=200.0)
    stop = threading.Event()

    def refill():
        while not stop.is_set():
            time.sleep(0.02)
            O.tfo_shm_fetch_add(d, 0, 40.0)          # 2000 tokens/s: a bucket every 100 ms

    t = threading.Thread(target=refill)
    t.start()
    try:
        out, _ = probe(["procaddr", 4, 16384, 256], shm=shm, TF_LIMITER_MAX_WAIT_MS="5000")
    finally:
        stop.set()
        t.join()
    assert out["launch_rc"] == 0 and out["driver_launches"] == 4 and out["hook_timeouts"] == 0
    assert out["hook_blocked"] >= 2 and 150.0 <= out["launch_ms"] < 4000.0   # buckets 2..4 had to be waited for
    O.tfo_shm_close(h)


def test_memory_limit_is_enforced_before_the_driver(tmp_path):
    h, d, shm = quota(tmp_path, mem_limit=1000)
    out, _ = probe(["procaddr", 0, 1, 32, 600, 600, 300, 200], shm=shm)
    assert out["allocs"] == [0, 2, 0, 2]  # CUDA_ERROR_OUT_OF_MEMORY for what does not fit in 1000 bytes
    assert out["driver_allocs"] == 2 and out["driver_frees"] == 2 and out["hook_denied"] == 2
    assert out["total"] == 1000 and out["free"] == 100 and out["free_after"] == 1000  # the pod's view, not the GPU's
    O.tfo_shm_close(h)


def test_hypervisor_reported_usage_counts_against_the_limit(tmp_path):
    h, d, shm = quota(tmp_path, mem_limit=1000)
    O.tfo_shm_set_pod_memory_used(d, 0, 700)  # another process of the pod, refreshed by the hypervisor at 2 Hz
    out, _ = probe(["procaddr", 0, 1, 32, 400, 300], shm=shm)
    assert out["allocs"] == [2, 0] and out["free"] == 300 and out["total"] == 1000
    O.tfo_shm_close(h)


def test_disabled_by_env_and_foreign_device(tmp_path):
    h, d, shm = quota(tmp_path, tokens=100.0)
    out, _ = probe(["procaddr", 5, 4, 128], shm=shm, DISABLE_GPU_LIMITER="1")
    assert out["hook_active"] == 0 and O.tfo_shm_get(d, 0, TOKENS) == 100.0
    out, _ = probe(["procaddr", 5, 4, 128], shm=shm, TF_ISOLATION_MODE="hard")  # not the mode this library implements
    assert out["hook_active"] == 0 and O.tfo_shm_get(d, 0, TOKENS) == 100.0
    O.tfo_shm_close(h)
    # the GPU the process runs on is not in the pod's quota file: forwarded, never charged
    sub = tmp_path / "other"
    sub.mkdir()
    h, d, shm = quota(sub, tokens=100.0, uuid=b"GPU-deadbeef-0000-0000-0000-000000000000")
    out, _ = probe(["procaddr", 5, 4, 128, 4096], shm=shm)
    assert out["hook_active"] == 1 and out["hook_launches"] == 0 and out["driver_launches"] == 5 and out["allocs"] == [0]
    assert O.tfo_shm_get(d, 0, TOKENS) == 100.0
    O.tfo_shm_close(h)


def test_registers_with_the_hypervisor(tmp_path):
    """POST /api/v1/process?container_pid= lets the hypervisor add the host PID to the quota file
    (handlers/legacy.go:319-384, :576)."""
    import http.server

    seen = []

    class H(http.server.BaseHTTPRequestHandler):
        def _ok(self, body):
            seen.append((self.command, self.path, self.headers.get("Authorization")))
            self.send_response(200)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def do_GET(self):
            self._ok(b'{"gpu_uuids":[],"vram_limit":1048576}')

        def do_POST(self):
            self._ok(b'{"success":true}')

        def log_message(self, *a):
            pass

    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    tok = tmp_path / "token"
    tok.write_text("jwt-abc\n")
    h, d, shm = quota(tmp_path, tokens=100.0)
    try:
        probe(["procaddr", 3, 1, 32], shm=shm, HYPERVISOR_IP="127.0.0.1", HYPERVISOR_PORT=str(srv.server_address[1]), CONTAINER_NAME="trainer",
              TFW_SA_TOKEN_FILE=str(tok), TF_LIMITER_LOG="1")
        deadline = time.time() + 5
        while len(seen) < 2 and time.time() < deadline:
            time.sleep(0.05)
    finally:
        srv.shutdown()
    paths = [s[1] for s in seen]
    # the handshake thread is detached: a short-lived process may exit before the POST -- the GET must be there
    assert any(p.startswith("/api/v1/pod?container_name=trainer") for p in paths), seen
    assert all(s[2] == "Bearer jwt-abc" for s in seen)
    O.tfo_shm_close(h)


@pytest.mark.parametrize("mode", ["procaddr", "procaddr_ptsz", "dlsym"])
def test_graph_replays_are_charged_unless_opted_out(tmp_path, mode):
    """A replayed CUDA graph costs the sum of its kernel nodes (child graphs included), computed once at
    instantiation -- 16 + 4 + 1 + 64 = 85 tokens for the stand-in driver's graph.  TF_LIMITER_CHARGE_GRAPHS=0 opts
    out: the graph entry points are not even substituted then."""
    h, d, shm = quota(tmp_path, tokens=100000.0)
    out, err = probe([mode, 50, 1, 32, "graph"], shm=shm, TF_LIMITER_LOG="1")   # on by default
    key = "driver_graph_launches" + ("_ptsz" if mode == "procaddr_ptsz" else "")
    assert out["graph_rc"] == 0 and out[key] == 50 and out["driver_graph_destroys"] == 1, err
    assert out["hook_launches"] == 50 and out["hook_tokens"] == 50 * 85 and out["hook_blocked"] == 0
    assert O.tfo_shm_get(d, 0, TOKENS) == 100000.0 - 50 * 85
    out, _ = probe([mode, 50, 1, 32, "graph"], shm=shm, TF_LIMITER_CHARGE_GRAPHS="0")   # opted out: forwarded untouched
    assert out[key] == 50
    assert out["hook_launches"] == 0 and out["hook_tokens"] == 0
    assert O.tfo_shm_get(d, 0, TOKENS) == 100000.0 - 50 * 85
    O.tfo_shm_close(h)


def test_graph_replays_wait_for_tokens_like_launches(tmp_path):
    h, d, shm = quota(tmp_path, tokens=100.0, capacity=1000.0)             # one replay (85) fits, the second does not
    out, _ = probe(["procaddr", 2, 1, 32, "graph"], shm=shm, TF_LIMITER_CHARGE_GRAPHS="1", TF_LIMITER_MAX_WAIT_MS="150")
    assert out["graph_rc"] == 0 and out["driver_graph_launches"] == 2 and out["hook_blocked"] == 1 and out["graph_ms"] >= 100.0
    assert O.tfo_shm_get(d, 0, TOKENS) == 15.0                             # the replay that timed out took nothing
    O.tfo_shm_close(h)
