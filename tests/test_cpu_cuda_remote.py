"""libcuda_remote.so -- the client stub of remote vGPU mode -- on the CPU: a driver-API
application (tools/cuda_remote_probe.c, linked against "libcuda.so.1") runs against a stand-in worker
(tests/test_cpu_client_shm.FakeWorker) over the shared-memory rings; the connection is named either directly
(TENSOR_FUSION_OPERATOR_CONNECTION_INFO, pod_webhook.go:580-586) or by the operator's
GET /api/connection endpoint (internal/server/router/connection.go:46-100)."""
import http.server
import json
import os
import subprocess
import threading

import pytest

import conftest
from test_cpu_client_shm import FakeWorker
from tensor_fusion_b200 import wire

ROOT = conftest.ROOT
STUB = os.path.join(ROOT, "build", "stub")
PROBE = os.path.join(ROOT, "build", "mock", "cuda_remote_probe")
LIB = os.path.join(ROOT, "tensor-fusion_b200", "lib", "libcuda_remote.so")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not (os.path.exists(LIB) and os.path.exists(PROBE) and os.path.exists(os.path.join(STUB, "libcuda.so.1"))):
        subprocess.run(["make", "-s", "tensor-fusion_b200/lib/libcuda_remote.so", "build/stub/libcuda.so.1", "build/mock/cuda_remote_probe"], cwd=ROOT, check=True)


def run_probe(env_extra, n=None):
    env = dict(os.environ, LD_LIBRARY_PATH=STUB, TF_ENABLE_LOG="1")
    for k in ("TENSOR_FUSION_OPERATOR_CONNECTION_INFO", "TENSOR_FUSION_OPERATOR_GET_CONNECTION_URL", "TF_CUDA_MEMORY_LIMIT"):
        env.pop(k, None)
    env.update(env_extra)
    return subprocess.run([PROBE] + ([str(n)] if n else []), env=env, capture_output=True, text=True, timeout=120)


def test_stub_exports_a_driver_api_and_nothing_else():
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    names = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for need in ("cuInit", "cuDeviceGet", "cuCtxCreate_v2", "cuMemAlloc_v2", "cuMemFree_v2", "cuMemcpyHtoD_v2", "cuMemcpyDtoH_v2", "cuMemcpyDtoD_v2",
                 "cuMemsetD8_v2", "cuLaunchKernel", "cuModuleGetFunction", "cuCtxSynchronize", "cuStreamSynchronize", "cuGetProcAddress_v2", "cuMemGetInfo_v2"):
        assert need in names, need
    assert all(n.startswith(("cu", "tfc_")) for n in names), sorted(n for n in names if not n.startswith(("cu", "tfc_")))
    assert "libstdc++" not in subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout


def test_no_connection_means_no_device():
    r = run_probe({})
    assert r.returncode == 10 and "CUDA_ERROR_NO_DEVICE" in r.stderr
    r = run_probe({"TENSOR_FUSION_OPERATOR_CONNECTION_INFO": "shmem+nobody-home+1+1", "TFC_CONNECT_TIMEOUT_MS": "200", "TFC_SHM_DIR": "/tmp"})
    assert r.returncode == 10 and "CUDA_ERROR_NO_DEVICE" in r.stderr and "cannot reach the worker" in r.stderr


def test_driver_api_application_over_the_rings(tmp_path):
    w = FakeWorker(str(tmp_path / "tf_shm"), 4 << 20, vram_quota=1 << 30)
    w.start()
    r = run_probe({"TENSOR_FUSION_OPERATOR_CONNECTION_INFO": "shmem+tf_shm+4+1", "TFC_SHM_DIR": str(tmp_path), "TF_CUDA_MEMORY_LIMIT": "1024"})
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["ok_a"] == 1 and out["ok_b"] == 1                       # copies, both kernels, D2D, memset: bit-exact
    assert out["count"] == 1 and "B200" in out["name"] and out["sms"] == 148 and out["cc_major"] == 10
    assert out["total"] == 1 << 30 and out["free_after_alloc"] == (1 << 30) - 2 * 5000003
    assert out["oom"] == 2                                             # CUDA_ERROR_OUT_OF_MEMORY from the worker's quota
    assert out["not_found"] == 500                                     # only the worker's built-in kernels exist
    assert out["bad_ptr"] == 1 and out["past_end"] == 1 and out["double_free"] == 1
    w.join(timeout=10)
    assert not w.is_alive() and w.h.worker_closed == 1                 # the stub's destructor closed the session
    # launches carry blocks x warps as their cost in tokens, the unit the limiter charges
    assert w.launches == [(wire.K_ADD_U8, 64, 256, 64 * 8), (wire.K_XOR_IDX, 16, 128, 16 * 4)]


def test_connection_url_from_the_operator(tmp_path):
    seen = []

    class H(http.server.BaseHTTPRequestHandler):
        def do_GET(self):
            seen.append((self.path, self.headers.get("Authorization")))
            body = b"shmem+via-operator+4+1"
            if not self.path.startswith("/api/connection?name=c1&namespace=ns1"):
                self.send_response(404)
                body = b'{"error":"connection not found"}'
            else:
                self.send_response(200)
            self.send_header("Content-Type", "text/plain; charset=utf-8")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *a):
            pass

    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    tok = tmp_path / "token"
    tok.write_text("sa-jwt\n")
    w = FakeWorker(str(tmp_path / "via-operator"), 4 << 20, vram_quota=1 << 30)
    w.start()
    base = f"http://127.0.0.1:{srv.server_address[1]}/api/connection"
    try:
        r = run_probe({"TENSOR_FUSION_OPERATOR_GET_CONNECTION_URL": base + "?name=c1&namespace=ns1", "TFC_SHM_DIR": str(tmp_path),
                       "TFW_SA_TOKEN_FILE": str(tok)}, n=1000000)
        assert r.returncode == 0, r.stderr
        assert seen == [("/api/connection?name=c1&namespace=ns1", "Bearer sa-jwt")]
        r = run_probe({"TENSOR_FUSION_OPERATOR_GET_CONNECTION_URL": base + "?name=other&namespace=ns1", "TFC_SHM_DIR": str(tmp_path)})
        assert r.returncode == 10 and "CUDA_ERROR_NO_DEVICE" in r.stderr
    finally:
        srv.shutdown()
    w.join(timeout=10)
