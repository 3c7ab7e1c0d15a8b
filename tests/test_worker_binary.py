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


def test_cli_contract():
    """compose.go:1311-1324: `-p 8000` (TCP) and `-n shmem -m tf_shm -M 1024`."""
    r = subprocess.run([EXE, "-n", "shmem", "-m", "tf_shm", "-M", "1024"], capture_output=True, text=True, timeout=20)
    assert r.returncode == 3 and "shmem" in r.stderr
    r = subprocess.run([EXE, "--help"], capture_output=True, text=True, timeout=20)
    assert r.returncode == 0 and "-p <port>" in r.stdout


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
