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


USER_PROBE = os.path.join(ROOT, "build", "mock", "cuda_user_probe")


@pytest.mark.parametrize("kind", ["cubin", "ptx", "fatbin"])
def test_application_with_its_own_kernels_over_the_rings(tmp_path, kind):
    """tools/cuda_user_probe.c ships tools/user_kernels.cu as a code image: cuModuleLoadData finds the image's length
    from its own header (the API carries none), page-locked host memory comes from an arena the worker maps, and the
    launches' parameter blocks (pointers at top level and inside a by-value struct) reach the stand-in worker, which
    executes the two integer kernels in numpy."""
    image = os.path.join(ROOT, "build", "mock", f"user_kernels.{kind}")
    if not (os.path.exists(USER_PROBE) and os.path.exists(image)):
        subprocess.run(["make", "-s", "build/mock/cuda_user_probe", f"build/mock/user_kernels.{kind}"], cwd=ROOT, check=True)
    w = FakeWorker(str(tmp_path / "tf_shm"), 4 << 20)
    w.start()
    env = dict(os.environ, LD_LIBRARY_PATH=STUB, TENSOR_FUSION_OPERATOR_CONNECTION_INFO="shmem+tf_shm+4+1", TFC_SHM_DIR=str(tmp_path))
    n = 200003
    r = subprocess.run([USER_PROBE, image, str(n)], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["ok_saxpy"] == 1 and out["ok_vec_add_struct"] == 1 and out["n"] == n
    assert out["not_found"] == 500 and out["bad_image"] == 200          # CUDA_ERROR_NOT_FOUND / CUDA_ERROR_INVALID_IMAGE
    w.join(timeout=10)
    raw = open(image, "rb").read()
    # the whole image and nothing but the image crossed the wire (PTX: the text up to its NUL)
    assert not w.modules                                                # unloaded at the end
    assert w.loaded_images == [raw.rstrip(b"\0") if kind == "ptx" else raw]
    names = [u[0] for u in w.user_launches]
    assert names == [b"saxpy_u32", b"vec_add_struct"]
    assert w.user_launches[0][1:3] == ((592, 1, 1), (256, 1, 1)) and w.user_launches[0][5] == 592 * 8
    assert w.user_launches[1][4] == 32                                  # one by-value struct of 32 bytes
    # host buffers were page-locked arenas: 3 x n x 4 bytes up and n x 4 down travelled by reference;
    # only the image, two names, two parameter blocks and one pageable D2H of n x 4 bytes were payload
    assert w.payload_bytes_in < len(raw) + 4096


def test_client_side_driver_api_surface_over_the_rings(tmp_path):
    """What the stub builds on the client out of wire operations the worker already has (tools/cuda_api_probe.c): pattern
    memsets (seed block + doubling D2D copies), unified-addressing cuMemcpy, events on the one ordered stream, pointer
    attributes, host functions -- against the numpy-executing stand-in worker."""
    probe = os.path.join(ROOT, "build", "mock", "cuda_api_probe")
    if not os.path.exists(probe):
        subprocess.run(["make", "-s", "build/mock/cuda_api_probe"], cwd=ROOT, check=True)
    w = FakeWorker(str(tmp_path / "tf_shm"), 8 << 20, vram_quota=1 << 30)
    w.start()
    env = dict(os.environ, LD_LIBRARY_PATH=STUB, TENSOR_FUSION_OPERATOR_CONNECTION_INFO="shmem+tf_shm+8+1", TFC_SHM_DIR=str(tmp_path))
    r = subprocess.run([probe], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr + r.stdout
    out = json.loads(r.stdout)
    assert out["ok_d32"] == 1 and out["ok_d32_bytes"] == 1 and out["ok_d16"] == 1 and out["ok_memcpy"] == 1
    assert out["misaligned"] == 1 and out["past_end"] == 1                 # CUDA_ERROR_INVALID_VALUE, as the driver answers
    assert out["query_done"] == 0 and out["elapsed"] == 801 and out["stale"] == 400
    assert out["mtype"] == 2 and out["base_ok"] == 1 and out["range"] == 300007 * 4 and out["host_ptr"] == 1
    assert out["cc"] == [10, 0] and out["called"] == 1
    w.join(timeout=10)
    assert not w.is_alive()
