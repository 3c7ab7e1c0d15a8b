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
