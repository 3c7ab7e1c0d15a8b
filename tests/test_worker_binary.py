"""`tensor-fusion-worker -p <port>` over loopback TCP (BASELINE config 0/1 harness)."""
import os
import socket
import subprocess
import threading

import pytest

import conftest

EXE = os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "tensor-fusion-worker")


def _start(extra_env=None):
    env = dict(os.environ, TFW_ONESHOT="1", TFW_BIND="127.0.0.1", TF_ENABLE_LOG="1")
    env.update(extra_env or {})
    p = subprocess.Popen([EXE, "-p", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, text=True)
    line = p.stdout.readline()
    assert "listening on port" in line, line
    return p, int(line.split()[-1])


def test_cli_contract(tmp_path):
    """compose.go:1311-1324: `-p 8000` (TCP) and `-n shmem -m tf_shm -M 1024`."""
    env = dict(os.environ, TFW_SHM_DIR=str(tmp_path))
    r = subprocess.run([EXE, "-n", "shmem", "-m", "tf_shm", "-M", "1"], capture_output=True, text=True, timeout=20, env=env)
    assert r.returncode == 2 and "shmem" in r.stderr
    r = subprocess.run([EXE, "-n", "shmem", "-m", "../x", "-M", "64"], capture_output=True, text=True, timeout=20, env=env)
    assert r.returncode == 2
    r = subprocess.run([EXE, "--help"], capture_output=True, text=True, timeout=20)
    assert r.returncode == 0 and "-p <port>" in r.stdout and "-n shmem" in r.stdout


@pytest.mark.skipif(conftest.HAS_GPU, reason="checks the no-GPU behaviour")
def test_shmem_transport_refused_without_a_gpu(tmp_path):
    env = dict(os.environ, TFW_SHM_DIR=str(tmp_path))
    r = subprocess.run([EXE, "-n", "shmem", "-m", "tf_shm", "-M", "8"], capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == 4 and "no CUDA device" in r.stderr       # the rings cannot be page-locked: no CPU fallback
    import ctypes as C
    from tensor_fusion_b200 import shm_ring as R
    raw = (tmp_path / "tf_shm").read_bytes()
    h = R.TfsrHeader.from_buffer_copy(raw[:C.sizeof(R.TfsrHeader)])
    assert len(raw) == 8 << 20 and h.magic == 0 and h.worker_ready == 0   # no client may attach to it


@pytest.mark.skipif(conftest.HAS_GPU, reason="checks the no-GPU behaviour")
def test_session_refused_without_a_gpu():
    p, port = _start()
    s = socket.create_connection(("127.0.0.1", port), timeout=20)
    s.settimeout(20)
    assert s.recv(16) == b""                 # closed at once: no CPU fallback
    s.close()
    _, err = p.communicate(timeout=30)
    assert "no CUDA device" in err


@pytest.mark.gpu
def test_loopback_replay_matches_oracle():
    import oracle
    from tensor_fusion_b200 import trace
    raw = trace.gen_c1().tobytes()
    rep = oracle.Replay(raw)
    p, port = _start({"TF_CUDA_MEMORY_LIMIT": "65536"})
    s = socket.create_connection(("127.0.0.1", port), timeout=60)
    got = bytearray()

    def reader():
        while True:
            b = s.recv(1 << 20)
            if not b:
                break
            got.extend(b)

    t = threading.Thread(target=reader)
    t.start()
    s.sendall(raw)
    s.shutdown(socket.SHUT_WR)
    t.join(timeout=120)
    s.close()
    _, err = p.communicate(timeout=60)
    assert bytes(got) == rep.responses(), err[-2000:]
    assert "session closed" in err and f"{rep.stat(0)} frames" in err


def test_bootstrap_handshake_with_the_hypervisor(tmp_path):
    """GET /api/v1/pod then POST /api/v1/process with the service-account bearer token."""
    import http.server
    import json
    seen = []

    class H(http.server.BaseHTTPRequestHandler):
        def _reply(self, data):
            body = json.dumps({"success": True, "data": data, "message": ""}).encode()
            self.send_response(200)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def do_GET(self):
            seen.append(("GET", self.path, self.headers.get("Authorization")))
            self._reply({"pod_name": "p", "namespace": "ns", "gpu_uuids": ["GPU-1234"], "vram_limit": 123456789, "qos_level": "Low", "compute_shard": False})

        def do_POST(self):
            seen.append(("POST", self.path, self.headers.get("Authorization")))
            self._reply({"host_pid": 22, "container_pid": 1, "container_name": "tensorfusion-worker", "pod_name": "p", "namespace": "ns"})

        def log_message(self, *a):
            pass

    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    tok = tmp_path / "token"
    tok.write_text("header.payload.sig\n")
    p, port = _start({"HYPERVISOR_IP": "127.0.0.1", "HYPERVISOR_PORT": str(srv.server_address[1]),
                      "CONTAINER_NAME": "tensorfusion-worker", "TFW_SA_TOKEN_FILE": str(tok)})
    s = socket.create_connection(("127.0.0.1", port), timeout=20)
    s.close()
    _, err = p.communicate(timeout=60)
    srv.shutdown()
    assert [x[0] for x in seen] == ["GET", "POST"], (seen, err)
    assert seen[0][1] == "/api/v1/pod?container_name=tensorfusion-worker" and seen[0][2] == "Bearer header.payload.sig"
    assert seen[1][1].startswith("/api/v1/process?container_name=tensorfusion-worker&container_pid=") and seen[1][1].endswith(str(p.pid))
    assert "/api/v1/pod ->" in err and "/api/v1/process ->" in err


@pytest.mark.parametrize("compact,pod,want", [
    (True, {"vram_limit": 123456789, "tflops_limit": 562.5, "isolation": "hard", "auto_freeze": {"freeze_to_mem_ttl": "1m30s", "enable": True}},
     "vram_limit=123456789 tflops_limit=562.5 isolation=hard auto_freeze_ttl_ms=90000 sm_percent=25"),
    (False, {"vram_limit": 1 << 34, "tflops_limit": 100, "isolation": "hard", "auto_freeze": {"enable": True, "freeze_to_disk_ttl": "1h", "freeze_to_mem_ttl": "250ms"}},
     "vram_limit=17179869184 tflops_limit=100 isolation=hard auto_freeze_ttl_ms=250 sm_percent=5"),
    (False, {"vram_limit": 5, "tflops_limit": 900.0, "isolation": "soft", "auto_freeze": {"freeze_to_mem_ttl": "5m", "enable": False}},
     "vram_limit=5 tflops_limit=900 isolation=soft auto_freeze_ttl_ms=0 sm_percent=0"),
    (True, {"qos_level": "Low"}, "vram_limit=0 tflops_limit=0 isolation=soft auto_freeze_ttl_ms=0 sm_percent=0"),
])
def test_pod_info_fields_the_worker_acts_on(compact, pod, want):
    """RemotePodInfo (pkg/hypervisor/api/http_types.go:82-100) as Go's encoder writes it (compact) and as any other JSON
    writer may (spaces): the VRAM quota, the TFLOPS limit that becomes an SM partition under hard isolation
    (computeUpLimit, controller.go:307-325: ceil(limit / 2250 * 100)), and the auto-freeze TTL (a Go duration)."""
    import http.server
    import json

    class H(http.server.BaseHTTPRequestHandler):
        def _reply(self, data):
            body = json.dumps({"success": True, "data": data, "message": ""}, separators=(",", ":") if compact else None).encode()
            self.send_response(200)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def do_GET(self):
            self._reply(dict({"pod_name": "p", "namespace": "ns", "gpu_uuids": ["GPU-1234"], "compute_shard": False}, **pod))

        def do_POST(self):
            self._reply({"host_pid": 22, "container_pid": 1})

        def log_message(self, *a):
            pass

    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    env = {"HYPERVISOR_IP": "127.0.0.1", "HYPERVISOR_PORT": str(srv.server_address[1]), "TFW_SA_TOKEN_FILE": "/nonexistent"}
    p, port = _start(env)
    socket.create_connection(("127.0.0.1", port), timeout=20).close()
    _, err = p.communicate(timeout=60)
    srv.shutdown()
    assert "pod info: " + want in err, err[-1500:]


@pytest.mark.gpu
def test_client_library_end_to_end_over_tcp(monkeypatch):
    monkeypatch.setenv("TFC_NO_SHM_UPGRADE", "1")     # this test is about the socket path
    """Remote vGPU path: libtfc_client.so (host only) -> TCP -> tensor-fusion-worker -> GPU, using the
    reference's connection URL format native+<ip>+<port>+<name>-<rv>
    (internal/controller/tensorfusionconnection_controller.go:136-138)."""
    import ctypes as C
    import numpy as np
    lib = C.CDLL(os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so"))
    lib.tfc_connect.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.tfc_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.tfc_memcpy_h2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.tfc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_memcpy_d2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_memset.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_uint64]
    lib.tfc_launch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    lib.tfc_free.argtypes = [C.c_void_p, C.c_uint32]
    lib.tfc_sync.argtypes = [C.c_void_p]
    lib.tfc_close.argtypes = [C.c_void_p]
    lib.tfc_last_error_code.argtypes = [C.c_void_p]
    p, port = _start()
    c = C.c_void_p()
    assert lib.tfc_connect(f"native+127.0.0.1+{port}+tf-worker-abc-12345".encode(), C.byref(c)) == 0
    rng = np.random.default_rng(12)
    n = 3_000_001
    a, b = C.c_uint32(), C.c_uint32()
    assert lib.tfc_malloc(c, n, C.byref(a)) == 0 and lib.tfc_malloc(c, n, C.byref(b)) == 0
    src = rng.integers(0, 256, n, dtype=np.uint8)
    assert lib.tfc_memcpy_h2d(c, a, 0, src.ctypes.data, n) == 0                       # large: sent in place
    small = rng.integers(0, 256, 1000, dtype=np.uint8)
    assert lib.tfc_memcpy_h2d(c, b, 7, small.ctypes.data, 1000) == 0                  # small: coalesced
    assert lib.tfc_launch(c, 2, 32, 128, a, 5, n - 5, 9, 0) == 0                      # add_u8 +9
    assert lib.tfc_memcpy_d2d(c, b, 2000, a, 1, 1_000_000) == 0
    assert lib.tfc_memset(c, b, 1_500_000, 0xEE, 333) == 0
    want_a = src.copy(); want_a[5:] += np.uint8(9)
    want_b = np.zeros(n, dtype=np.uint8); want_b[7:1007] = small; want_b[2000:1_002_000] = want_a[1:1_000_001]; want_b[1_500_000:1_500_333] = 0xEE
    got = np.empty(n, dtype=np.uint8)
    assert lib.tfc_memcpy_d2h(c, got.ctypes.data, b, 0, n) == 0 and np.array_equal(got, want_b)
    assert lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n) == 0 and np.array_equal(got, want_a)
    assert lib.tfc_sync(c) == 0
    assert lib.tfc_memcpy_d2h(c, got.ctypes.data, 999, 0, 16) == 2                     # NOT_FOUND comes back as the call's result
    assert lib.tfc_memset(c, a, n, 1, 10) == 0 and lib.tfc_sync(c) == 1                # INVALID surfaces at the next sync
    assert lib.tfc_free(c, a) == 0 and lib.tfc_free(c, b) == 0 and lib.tfc_sync(c) == 0
    lib.tfc_close(c)
    _, err = p.communicate(timeout=60)
    assert "session closed" in err


@pytest.mark.gpu
def test_accel_snapshot_parks_the_vgpu_in_host_memory_and_resume_restores_it(tmp_path, monkeypatch):
    """Hypervisor -> AccelSnapshot(pid of the worker) -> control words of the worker's stats record ->
    the worker parks every buffer in host memory and frees its HBM; the client is back-pressured, not
    failed; AccelResume brings the bytes back bit for bit (accelerator.h:364-390, handlers/worker.go:94-129)."""
    import ctypes as C
    import mmap
    import time
    import numpy as np
    import oracle
    from oracle import lib as O
    from tensor_fusion_b200 import provider as P
    base = tmp_path / "shm"
    base.mkdir()
    h = C.c_void_p()
    cfg = (oracle.DevCfg * 1)()
    cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 0, b"GPU-any", 100, 1 << 40
    assert O.tfo_shm_create(str(base).encode(), b"ns", b"snap-pod", cfg, 1, C.byref(h)) == 0
    pod = base / "ns" / "snap-pod"
    monkeypatch.setenv("TF_SHM_BASE_PATH", str(base))
    prov = P.load()
    prov.LimiterShutdown()

    lib = C.CDLL(os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so"))
    lib.tfc_connect.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.tfc_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.tfc_memcpy_h2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.tfc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_launch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    lib.tfc_sync.argtypes = [C.c_void_p]
    lib.tfc_close.argtypes = [C.c_void_p]
    p, port = _start({"TF_SHM_PATH": str(pod / "shm"), "DISABLE_GPU_LIMITER": ""})
    c = C.c_void_p()
    assert lib.tfc_connect(f"native+127.0.0.1+{port}+tf-worker-snap-1".encode(), C.byref(c)) == 0
    n = 96 << 20
    rng = np.random.default_rng(5)
    a, b = C.c_uint32(), C.c_uint32()
    assert lib.tfc_malloc(c, n, C.byref(a)) == 0 and lib.tfc_malloc(c, n, C.byref(b)) == 0
    src_a = rng.integers(0, 256, n, dtype=np.uint8)
    src_b = rng.integers(0, 256, n, dtype=np.uint8)
    assert lib.tfc_memcpy_h2d(c, a, 0, src_a.ctypes.data, n) == 0 and lib.tfc_memcpy_h2d(c, b, 0, src_b.ctypes.data, n) == 0
    assert lib.tfc_launch(c, 2, 148, 256, a, 0, n, 3, 0) == 0      # add_u8 +3 on all of a
    assert lib.tfc_sync(c) == 0

    f = open(pod / "tfw_stats", "r+b")
    mm = mmap.mmap(f.fileno(), C.sizeof(P.TfwStatsRecord))
    rec = P.TfwStatsRecord.from_buffer(mm)
    assert rec.magic == P.TFW_STATS_MAGIC and rec.pid == p.pid and rec.ctl_frozen == 0

    pids = (C.c_int32 * 1)(p.pid)
    ctx = P.SnapshotContext(processIds=C.cast(pids, C.POINTER(C.c_int32)), processCount=1, deviceUUID=None)
    t0 = time.time()
    assert prov.AccelSnapshot(C.byref(ctx)) == P.SUCCESS
    snap_s = time.time() - t0
    assert rec.ctl_frozen == 1 and rec.parked_bytes == 2 * n and rec.ctl_moved_bytes == 2 * n and rec.vram_bytes == 0
    assert prov.AccelSnapshot(C.byref(ctx)) == P.SUCCESS and rec.parked_bytes == 2 * n   # idempotent

    # the client keeps working against a frozen vGPU: its call simply waits for the resume
    got = np.empty(n, dtype=np.uint8)
    res = []
    th = threading.Thread(target=lambda: res.append(lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n)), daemon=True)
    th.start()
    th.join(timeout=1.0)
    assert th.is_alive()
    t0 = time.time()
    assert prov.AccelResume(C.byref(ctx)) == P.SUCCESS
    resume_s = time.time() - t0
    th.join(timeout=60)
    if res != [0]:
        p.kill()
    assert res == [0] and np.array_equal(got, src_a + np.uint8(3))
    assert lib.tfc_memcpy_d2h(c, got.ctypes.data, b, 0, n) == 0 and np.array_equal(got, src_b)
    assert rec.ctl_frozen == 0 and rec.parked_bytes == 0 and rec.vram_bytes == 2 * n
    print(f"snapshot {2 * n / snap_s / 1e9:.2f} GB/s ({snap_s * 1e3:.0f} ms), resume {2 * n / resume_s / 1e9:.2f} GB/s ({resume_s * 1e3:.0f} ms)")
    del rec
    mm.close()
    f.close()
    lib.tfc_close(c)
    _, err = p.communicate(timeout=60)
    assert "session closed" in err
    O.tfo_shm_close(h)


@pytest.mark.gpu
def test_shared_memory_transport_end_to_end(request, monkeypatch):
    """`-n shmem -m <name> -M <MiB>` + "shmem+<name>+<MiB>+1" (compose.go:1311-1317, pod_webhook.go:584): the client
    writes TFCS frames into page-locked rings, the copy engine reads the payloads in place.  Rings of 6 MiB / 2 MiB,
    so a 20 MB copy laps them several times and headers land on every phase of the wrap; two sessions in a row."""
    import ctypes as C
    import shutil
    import tempfile
    import numpy as np
    # page-locking needs anonymous / tmpfs pages: the rings live in /dev/shm like in the pod (constants.go:291)
    tmp_path = tempfile.mkdtemp(dir="/dev/shm", prefix="tfw-test-")
    request.addfinalizer(lambda: shutil.rmtree(tmp_path, ignore_errors=True))
    monkeypatch.setenv("TFC_SHM_DIR", str(tmp_path))
    lib = C.CDLL(os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so"))
    lib.tfc_connect.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.tfc_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.tfc_memcpy_h2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.tfc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_memcpy_d2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_memset.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_uint64]
    lib.tfc_launch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    lib.tfc_free.argtypes = [C.c_void_p, C.c_uint32]
    lib.tfc_sync.argtypes = [C.c_void_p]
    lib.tfc_close.argtypes = [C.c_void_p]
    env = dict(os.environ, TFW_ONESHOT="2", TFW_SHM_DIR=str(tmp_path), TF_ENABLE_LOG="1")
    p = subprocess.Popen([EXE, "-n", "shmem", "-m", "tfring", "-M", "8"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, text=True)
    try:
        line = p.stdout.readline()
        assert "serving shmem" in line, line + p.stderr.read()
        rng = np.random.default_rng(44)
        for session in range(2):
            c = C.c_void_p()
            assert lib.tfc_connect(b"shmem+tfring+8+1", C.byref(c)) == 0
            n = 20_000_003
            a, b = C.c_uint32(), C.c_uint32()
            assert lib.tfc_malloc(c, n, C.byref(a)) == 0 and lib.tfc_malloc(c, n, C.byref(b)) == 0
            src = rng.integers(0, 256, n, dtype=np.uint8)
            assert lib.tfc_memcpy_h2d(c, a, 0, src.ctypes.data, n) == 0
            want_b = np.zeros(n, dtype=np.uint8)
            for i in range(400):                      # small frames of every length class between the big ones
                piece = rng.integers(0, 256, 1 + (i * 7) % 300, dtype=np.uint8)
                off = (i * 4099) % (n - 400)
                assert lib.tfc_memcpy_h2d(c, b, off, piece.ctypes.data, len(piece)) == 0
                want_b[off:off + len(piece)] = piece
                if i % 100 == 50:
                    assert lib.tfc_memcpy_h2d(c, a, i, src[i:].ctypes.data, 3_000_000) == 0   # rewrite with the same bytes
            assert lib.tfc_launch(c, 2, 64, 256, a, 5, n - 5, 9, 0) == 0                          # add_u8 +9
            want_a = src.copy()
            want_a[5:] += np.uint8(9)
            assert lib.tfc_memcpy_d2d(c, b, 2_000_000, a, 1, 1_000_000) == 0
            want_b[2_000_000:3_000_000] = want_a[1:1_000_001]
            assert lib.tfc_memset(c, b, 1_500_000, 0xEE, 333) == 0
            want_b[1_500_000:1_500_333] = 0xEE
            got = np.empty(n, dtype=np.uint8)
            assert lib.tfc_memcpy_d2h(c, got.ctypes.data, b, 0, n) == 0 and np.array_equal(got, want_b)   # 10x the downstream ring
            assert lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n) == 0 and np.array_equal(got, want_a)
            assert lib.tfc_memcpy_d2h(c, got.ctypes.data, 999, 0, 16) == 2
            assert lib.tfc_free(c, a) == 0 and lib.tfc_free(c, b) == 0 and lib.tfc_sync(c) == 0
            lib.tfc_close(c)
        _, err = p.communicate(timeout=60)
        assert p.returncode == 0 and err.count("session closed") == 2, err[-2000:]
    finally:
        if p.poll() is None:
            p.kill()


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["tcp", "shmem"])
def test_driver_api_application_through_the_client_stub(transport, request):
    """tools/cuda_remote_probe.c is linked against "libcuda.so.1"; with the stub directory in front that is
    libcuda_remote.so, and cuMemAlloc / cuMemcpy / cuLaunchKernel / cuMemset run on the B200 behind the worker."""
    import json
    import shutil
    import tempfile
    stub = os.path.join(conftest.ROOT, "build", "stub")
    probe = os.path.join(conftest.ROOT, "build", "mock", "cuda_remote_probe")
    if not (os.path.exists(probe) and os.path.exists(os.path.join(stub, "libcuda.so.1"))):
        subprocess.run(["make", "-s", "build/stub/libcuda.so.1", "build/mock/cuda_remote_probe"], cwd=conftest.ROOT, check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=stub, TF_ENABLE_LOG="1", TF_CUDA_MEMORY_LIMIT="4096", TFC_NO_SHM_UPGRADE="1")
    if transport == "tcp":
        p, port = _start({"TF_CUDA_MEMORY_LIMIT": "4096"})
        env["TENSOR_FUSION_OPERATOR_CONNECTION_INFO"] = f"native+127.0.0.1+{port}+probe-1"
    else:
        d = tempfile.mkdtemp(dir="/dev/shm", prefix="tfw-stub-")
        request.addfinalizer(lambda: shutil.rmtree(d, ignore_errors=True))
        wenv = dict(os.environ, TFW_ONESHOT="1", TFW_SHM_DIR=d, TF_ENABLE_LOG="1", TF_CUDA_MEMORY_LIMIT="4096")
        p = subprocess.Popen([EXE, "-n", "shmem", "-m", "tf_shm", "-M", "64"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=wenv, text=True)
        assert "serving shmem" in p.stdout.readline()
        env.update(TENSOR_FUSION_OPERATOR_CONNECTION_INFO="shmem+tf_shm+64+1", TFC_SHM_DIR=d)
    try:
        r = subprocess.run([probe, "40000003"], env=env, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-2000:]
        out = json.loads(r.stdout)
        assert out["ok_a"] == 1 and out["ok_b"] == 1
        assert out["oom"] in (1, 2) and out["not_found"] == 500 and out["bad_ptr"] == 1 and out["double_free"] == 1
        assert out["total"] == 4096 << 20
        _, err = p.communicate(timeout=60)
        assert "session closed" in err
    finally:
        if p.poll() is None:
            p.kill()


@pytest.mark.gpu
@pytest.mark.parametrize("transport,kind", [("shmem", "cubin"), ("shmem", "ptx"), ("tcp", "fatbin")])
def test_application_kernels_through_the_stub_match_native_cuda(transport, kind, request):
    """tools/cuda_user_probe.c loads tools/user_kernels.cu as a code image and launches it: once on the real driver,
    once through libcuda_remote.so -> tensor-fusion-worker (MODULE_LOAD / LAUNCH_USER, tagged pointers translated on
    the worker, host buffers in page-locked arenas over shmem).  Integer kernels: the two runs must agree bit for bit."""
    import json
    import shutil
    import tempfile
    stub = os.path.join(conftest.ROOT, "build", "stub")
    mock = os.path.join(conftest.ROOT, "build", "mock")
    image = os.path.join(mock, f"user_kernels.{kind}")
    n = "3000017"
    native = subprocess.run([os.path.join(mock, "cuda_user_probe_native"), image, n], capture_output=True, text=True, timeout=120)
    assert native.returncode == 0, native.stderr[-2000:]
    want = json.loads(native.stdout)
    assert want["ok_saxpy"] == 1 and want["ok_vec_add_struct"] == 1
    env = dict(os.environ, LD_LIBRARY_PATH=stub, TF_ENABLE_LOG="1", TFC_NO_SHM_UPGRADE="1")
    if transport == "tcp":
        p, port = _start()
        env["TENSOR_FUSION_OPERATOR_CONNECTION_INFO"] = f"native+127.0.0.1+{port}+probe-1"
    else:
        d = tempfile.mkdtemp(dir="/dev/shm", prefix="tfw-user-")
        request.addfinalizer(lambda: shutil.rmtree(d, ignore_errors=True))
        wenv = dict(os.environ, TFW_ONESHOT="1", TFW_SHM_DIR=d, TF_ENABLE_LOG="1")
        p = subprocess.Popen([EXE, "-n", "shmem", "-m", "tf_shm", "-M", "16"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=wenv, text=True)
        assert "serving shmem" in p.stdout.readline()
        env.update(TENSOR_FUSION_OPERATOR_CONNECTION_INFO="shmem+tf_shm+16+1", TFC_SHM_DIR=d)
    try:
        r = subprocess.run([os.path.join(mock, "cuda_user_probe"), image, n], env=env, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-2000:]
        got = json.loads(r.stdout)
        # digests of both outputs, flags, the NOT_FOUND code: everything but the code for a garbage image (the real driver
        # may try it as PTX first and say INVALID_PTX where the stub says INVALID_IMAGE)
        assert got["bad_image"] != 0 and want["bad_image"] != 0
        assert {k: v for k, v in got.items() if k != "bad_image"} == {k: v for k, v in want.items() if k != "bad_image"}
        _, err = p.communicate(timeout=60)
        assert "session closed" in err
    finally:
        if p.poll() is None:
            p.kill()


@pytest.mark.gpu
def test_page_locked_arenas_and_ring_dma_through_the_real_worker(request, monkeypatch):
    """Client memory from tfc_host_alloc is mapped and page-locked by the worker too: H2D / D2H from / to it are
    by-reference DMAs (tfw_stats h2d_ref_bytes / d2h_ref_bytes through the session log); pageable destinations get
    their bytes through the worker -> client ring, written there by the copy engine (a 50 MB read through a 4 MiB ring)."""
    import ctypes as C
    import shutil
    import tempfile
    import numpy as np
    d = tempfile.mkdtemp(dir="/dev/shm", prefix="tfw-arena-")
    request.addfinalizer(lambda: shutil.rmtree(d, ignore_errors=True))
    monkeypatch.setenv("TFC_SHM_DIR", d)
    lib = C.CDLL(os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so"))
    lib.tfc_connect.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.tfc_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.tfc_memcpy_h2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.tfc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_memcpy_d2h_async.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_launch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    lib.tfc_host_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.tfc_host_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.tfc_sync.argtypes = [C.c_void_p]
    lib.tfc_close.argtypes = [C.c_void_p]
    env = dict(os.environ, TFW_ONESHOT="1", TFW_SHM_DIR=d, TF_ENABLE_LOG="1")
    p = subprocess.Popen([EXE, "-n", "shmem", "-m", "ring", "-M", "16"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, text=True)
    try:
        assert "serving shmem" in p.stdout.readline()
        c = C.c_void_p()
        assert lib.tfc_connect(b"shmem+ring+16+1", C.byref(c)) == 0
        n = 50_000_017
        hp, hq = C.c_void_p(), C.c_void_p()
        assert lib.tfc_host_alloc(c, n, C.byref(hp)) == 0 and lib.tfc_host_alloc(c, n, C.byref(hq)) == 0   # 64 MiB arena + a second one
        src = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), (n,))
        back = np.ctypeslib.as_array(C.cast(hq, C.POINTER(C.c_uint8)), (n,))
        src[:] = np.random.default_rng(77).integers(0, 256, n, dtype=np.uint8)
        a = C.c_uint32()
        assert lib.tfc_malloc(c, n, C.byref(a)) == 0
        assert lib.tfc_memcpy_h2d(c, a, 0, hp, n) == 0                                   # by reference
        assert lib.tfc_memcpy_h2d(c, a, 7, C.c_void_p(hp.value + 4099), 1_000_001) == 0   # odd source and destination
        want = src.copy()
        want[7:7 + 1_000_001] = src[4099:4099 + 1_000_001]
        assert lib.tfc_launch(c, 3, 128, 256, a, 0, n, 5, 0) == 0                        # xor_idx: ordered after the DMAs
        want ^= ((np.arange(n, dtype=np.uint64) * np.uint64(5)) >> np.uint64(3)).astype(np.uint8)
        assert lib.tfc_memcpy_d2h_async(c, hq, a, 0, n) == 0 and lib.tfc_sync(c) == 0    # by reference, asynchronous
        assert np.array_equal(back, want)
        page = np.empty(n, dtype=np.uint8)
        assert lib.tfc_memcpy_d2h(c, page.ctypes.data, a, 0, n) == 0                     # through the 4 MiB ring, DMA'd in pieces
        assert np.array_equal(page, want)
        back[:] = 0
        assert lib.tfc_memcpy_d2h(c, hq, a, 3, n - 3) == 0 and np.array_equal(back[:n - 3], want[3:])   # synchronous by-reference form
        assert lib.tfc_host_free(c, hp) == 0 and lib.tfc_host_free(c, hq) == 0
        assert not [f for f in os.listdir(d) if ".a" in f]
        lib.tfc_close(c)
        _, err = p.communicate(timeout=60)
        assert p.returncode == 0 and "session closed" in err, err[-2000:]
    finally:
        if p.poll() is None:
            p.kill()


@pytest.mark.gpu
def test_loopback_tcp_connection_upgrades_itself_to_shared_memory_rings(request, monkeypatch):
    """`native+127.0.0.1+<port>+...`: client and worker are on one node, so the first frame offers a ring file the client
    created (TFCS_OP_UPGRADE_SHM); the worker maps and page-locks it and the session -- page-locked arenas included --
    runs over the rings, with the socket kept as its lifeline.  With TFW_NO_SHM_UPGRADE the same URL stays on TCP."""
    import ctypes as C
    import shutil
    import tempfile
    import numpy as np
    d = tempfile.mkdtemp(dir="/dev/shm", prefix="tfw-upg-")
    request.addfinalizer(lambda: shutil.rmtree(d, ignore_errors=True))
    monkeypatch.setenv("TFC_SHM_DIR", d)
    monkeypatch.setenv("TFC_UPGRADE_MIB", "32")
    lib = C.CDLL(os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so"))
    lib.tfc_connect.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.tfc_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.tfc_memcpy_h2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.tfc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_host_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.tfc_host_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.tfc_sync.argtypes = [C.c_void_p]
    lib.tfc_close.argtypes = [C.c_void_p]
    for upgrade in (True, False):
        p, port = _start({"TFW_SHM_DIR": d} if upgrade else {"TFW_SHM_DIR": d, "TFW_NO_SHM_UPGRADE": "1"})
        try:
            c = C.c_void_p()
            assert lib.tfc_connect(f"native+127.0.0.1+{port}+up-1".encode(), C.byref(c)) == 0
            n = 40_000_003
            a = C.c_uint32()
            src = np.random.default_rng(9).integers(0, 256, n, dtype=np.uint8)
            assert lib.tfc_malloc(c, n, C.byref(a)) == 0 and lib.tfc_memcpy_h2d(c, a, 0, src.ctypes.data, n) == 0
            hp = C.c_void_p()
            rc = lib.tfc_host_alloc(c, 1 << 20, C.byref(hp))
            assert rc == (0 if upgrade else 3)                         # page-locked sharing exists on the rings only
            if upgrade:
                assert [f for f in os.listdir(d) if f.startswith("tfw-up-")]
                assert lib.tfc_host_free(c, hp) == 0
            got = np.empty(n, dtype=np.uint8)
            assert lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n) == 0 and np.array_equal(got, src)
            assert lib.tfc_sync(c) == 0
            lib.tfc_close(c)
            _, err = p.communicate(timeout=60)
            assert ("upgraded to shared-memory rings" in err) == upgrade and "session closed" in err, err[-1500:]
            assert not [f for f in os.listdir(d) if f.startswith("tfw-up-")]     # the ring file is gone with the session
        finally:
            if p.poll() is None:
                p.kill()


def test_sigterm_stops_the_listener_gracefully():
    """Pod deletion sends SIGTERM: the worker stops accepting and exits 0 (sessions drain and close first)."""
    import signal
    import time
    p, port = _start({"TFW_ONESHOT": "-1"})          # no connection budget: would serve for ever
    time.sleep(0.3)
    assert p.poll() is None
    socket.create_connection(("127.0.0.1", port), timeout=2).close()   # it is accepting
    t0 = time.time()
    p.send_signal(signal.SIGTERM)
    assert p.wait(timeout=10) == 0 and time.time() - t0 < 5
    with pytest.raises(OSError):
        socket.create_connection(("127.0.0.1", port), timeout=2)


@pytest.mark.gpu
def test_client_side_driver_api_surface_through_the_worker(request):
    """tools/cuda_api_probe.c on the B200: 16/32-bit pattern memsets (a seed block + doubling D2D copies inside the buffer),
    unified-addressing cuMemcpy, an asynchronous D2H into page-locked memory covered by an event, pointer attributes."""
    import json
    import shutil
    import tempfile
    stub = os.path.join(conftest.ROOT, "build", "stub")
    probe = os.path.join(conftest.ROOT, "build", "mock", "cuda_api_probe")
    if not (os.path.exists(probe) and os.path.exists(os.path.join(stub, "libcuda.so.1"))):
        subprocess.run(["make", "-s", "build/stub/libcuda.so.1", "build/mock/cuda_api_probe"], cwd=conftest.ROOT, check=True)
    d = tempfile.mkdtemp(dir="/dev/shm", prefix="tfw-api-")
    request.addfinalizer(lambda: shutil.rmtree(d, ignore_errors=True))
    wenv = dict(os.environ, TFW_ONESHOT="1", TFW_SHM_DIR=d, TF_ENABLE_LOG="1")
    p = subprocess.Popen([EXE, "-n", "shmem", "-m", "tf_shm", "-M", "64"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=wenv, text=True)
    try:
        assert "serving shmem" in p.stdout.readline()
        env = dict(os.environ, LD_LIBRARY_PATH=stub, TENSOR_FUSION_OPERATOR_CONNECTION_INFO="shmem+tf_shm+64+1", TFC_SHM_DIR=d)
        r = subprocess.run([probe, "10000019"], env=env, capture_output=True, text=True, timeout=120)      # 40 MB per buffer
        assert r.returncode == 0, (r.stderr + r.stdout)[-2000:]
        out = json.loads(r.stdout)
        assert out["ok_d32"] == 1 and out["ok_d32_bytes"] == 1 and out["ok_d16"] == 1 and out["ok_memcpy"] == 1
        assert out["misaligned"] == 1 and out["past_end"] == 1 and out["elapsed"] == 801 and out["stale"] == 400
        assert out["mtype"] == 2 and out["range"] == 10000019 * 4 and out["called"] == 1
        _, err = p.communicate(timeout=60)
        assert "session closed" in err
    finally:
        if p.poll() is None:
            p.kill()
