"""libcuda_limiter.so preloaded into a real CUDA application (a PyTorch process) on a B200:
every kernel launch is charged to the pod's quota file, the launch loop is paced by the refill,
the memory limit is the pod's, results are untouched."""
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import pytest

import conftest
import oracle
from oracle import lib as O

pytestmark = pytest.mark.gpu
ROOT = conftest.ROOT
HOOK = os.path.join(ROOT, "tensor-fusion_b200", "lib", "libcuda_limiter.so")
RATE, CAPACITY, TOKENS = 0, 1, 2

APP = r'''
import ctypes, json, sys, time
import torch
n_launch = int(sys.argv[1])
x = torch.zeros(1 << 20, device="cuda")
torch.cuda.synchronize()
class S(ctypes.Structure):
    _fields_ = [(k, ctypes.c_uint64) for k in ("launches", "blocked", "timeouts", "wait_ns", "tokens", "denied", "active")]
def stats():
    s = S()
    try:
        ctypes.CDLL(None).tf_hook_get_stats(ctypes.byref(s))
    except AttributeError:
        return None
    return {k: getattr(s, k) for k, _ in S._fields_}
before = stats()
t0 = time.time()
for _ in range(n_launch):
    x.add_(1.0)
torch.cuda.synchronize()
loop_s = time.time() - t0
after = stats()
free, total = torch.cuda.mem_get_info()
oom = False
try:
    big = torch.empty(int(sys.argv[2]), dtype=torch.uint8, device="cuda")
except torch.OutOfMemoryError:
    oom = True
ok = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
print(json.dumps({"sum": float(x.sum().item()), "loop_s": loop_s, "before": before, "after": after, "free": free, "total": total,
                  "oom": oom, "uuid": str(torch.cuda.get_device_properties(0).uuid)}))
'''


def run_app(n_launch, big_bytes, shm=None, preload=True, **extra):
    env = dict(os.environ)
    env.pop("TF_SHM_PATH", None)
    env.pop("HYPERVISOR_IP", None)
    if preload:
        env["LD_PRELOAD"] = HOOK
    if shm:
        env.update(TF_SHM_PATH=shm, TF_ISOLATION_MODE="soft", TF_LIMITER_LOG="1")
    env.update(extra)
    r = subprocess.run([sys.executable, "-c", APP, str(n_launch), str(big_bytes)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1]), r.stderr


def gpu_uuid():
    out, _ = run_app(1, 1 << 20, preload=False)
    u = out["uuid"]
    return u if u.startswith("GPU-") else "GPU-" + u


def quota(tmp_path, uuid, tokens, capacity, mem_limit):
    h = C.c_void_p()
    cfg = (oracle.DevCfg * 1)()
    cfg[0].device_idx, cfg[0].uuid, cfg[0].up_limit, cfg[0].mem_limit = 0, uuid.encode(), 50, mem_limit
    assert O.tfo_shm_create(str(tmp_path).encode(), b"ns", b"pod", cfg, 1, C.byref(h)) == 0
    d = O.tfo_shm_data(h)
    O.tfo_shm_set(d, 0, CAPACITY, capacity)
    O.tfo_shm_set(d, 0, TOKENS, tokens)
    return h, d, os.path.join(str(tmp_path), "ns", "pod", "shm")


def test_pytorch_under_the_preloaded_limiter(tmp_path):
    uuid = gpu_uuid()
    N = 400
    # 1. plenty of tokens: nothing blocks, every launch of the loop is charged, the result is right
    h, d, shm = quota(tmp_path, uuid, tokens=1e12, capacity=1e12, mem_limit=8 << 30)
    out, err = run_app(N, 16 << 30, shm=shm)
    assert out["sum"] == float(N) * (1 << 20), err[-1500:]
    a, b = out["after"], out["before"]
    assert a is not None and a["active"] == 1, err[-1500:]
    launches = a["launches"] - b["launches"]
    tokens = a["tokens"] - b["tokens"]
    assert launches == N, (launches, err[-1500:])          # one cuLaunchKernel per add_
    assert a["blocked"] == 0 and a["timeouts"] == 0
    per_launch = tokens // N
    assert tokens == per_launch * N and per_launch >= (1 << 20) // 32 // 16  # >= one warp per 16 warps' worth of elements
    lost = 1e12 - O.tfo_shm_get(d, 0, TOKENS)
    assert a["tokens"] <= lost <= a["tokens"] + 64 * per_launch   # the file lost what the hook charged (+ the script's tail: sum())
    # the pod's memory view: 8 GiB, the 16 GiB tensor does not fit, a 64 MiB one does
    assert out["total"] == 8 << 30 and out["free"] <= 8 << 30 and out["oom"] is True
    O.tfo_shm_close(h)

    # 2. a bucket that covers the start-up and half of the loop; the rest arrives 0.5 s after the bucket ran dry:
    #    the launching thread waits for it, every launch still runs, nothing is minted or lost
    sub = tmp_path / "paced"
    sub.mkdir()
    need = float(per_launch * N)
    h, d, shm = quota(sub, uuid, tokens=1e9, capacity=1e12, mem_limit=8 << 30)
    run_app(N, 16 << 30, shm=shm)
    whole_run = 1e9 - O.tfo_shm_get(d, 0, TOKENS)   # context creation and the tail of the script launch kernels too
    assert whole_run >= need
    O.tfo_shm_set(d, 0, TOKENS, whole_run - need / 2)
    refilled_at = []
    SLACK = 1e6  # a start-up that launches a few more kernels than the calibration run must not block at the end

    def late():
        deadline = time.time() + 240
        while O.tfo_shm_get(d, 0, TOKENS) >= per_launch and time.time() < deadline:
            time.sleep(0.005)
        time.sleep(0.5)
        O.tfo_shm_fetch_add(d, 0, need / 2 + SLACK)
        refilled_at.append(time.time())

    th = threading.Thread(target=late)
    th.start()
    out, err = run_app(N, 16 << 30, shm=shm, TF_LIMITER_MAX_WAIT_MS="60000")
    th.join()
    assert out["sum"] == float(N) * (1 << 20)
    assert out["after"]["blocked"] >= 1 and out["after"]["timeouts"] == 0, err[-1500:]
    assert out["after"]["wait_ns"] >= 0.4e9 and out["loop_s"] >= 0.4
    assert 0.0 <= O.tfo_shm_get(d, 0, TOKENS) <= 2 * SLACK   # granted ~= charged: nothing minted
    O.tfo_shm_close(h)


def test_worker_binary_is_not_double_charged(tmp_path):
    """Soft isolation mounts the limiter into the worker container too (compose.go:1441-1447); the
    worker gates on the GPU, so the preloaded library must only forward there."""
    exe = os.path.join(ROOT, "tensor-fusion_b200", "lib", "tensor-fusion-worker")
    env = dict(os.environ, LD_PRELOAD=HOOK, TF_LIMITER_LOG="1", TF_SHM_PATH=str(tmp_path / "none"))
    r = subprocess.run([exe, "-h"], env=env, capture_output=True, text=True, timeout=30)
    assert "inside tensor-fusion-worker" in r.stderr and "forwarding only" in r.stderr
