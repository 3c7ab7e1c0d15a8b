"""Client -> worker transports side by side, through libtfc_client.so (the call a client shim makes):
TCP loopback (`-p`) vs the page-locked shared-memory rings (`-n shmem`).  Bulk leg: 64 copies of 64 MiB
into 16 device buffers + sync (GB/s of payload).  Small-call leg: 4 KiB copies + one sync (us per call).
D2H leg: 16 reads of 64 MiB."""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tensor-fusion_b200", "lib", "tensor-fusion-worker")
lib = C.CDLL(os.path.join(ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so"))
lib.tfc_connect.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
lib.tfc_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
lib.tfc_memcpy_h2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
lib.tfc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
lib.tfc_sync.argtypes = [C.c_void_p]
lib.tfc_close.argtypes = [C.c_void_p]


def run(url, proc):
    c = C.c_void_p()
    assert lib.tfc_connect(url.encode(), C.byref(c)) == 0, url
    each, nbuf, ncopies = 64 << 20, 16, 64
    hs = []
    for _ in range(nbuf):
        h = C.c_uint32()
        assert lib.tfc_malloc(c, each, C.byref(h)) == 0
        hs.append(h)
    assert lib.tfc_sync(c) == 0                     # context creation, allocation and zero-fill are not part of the numbers
    src = np.random.default_rng(1).integers(0, 256, each, dtype=np.uint8)
    out = {}
    t0 = time.perf_counter()
    for i in range(ncopies):
        assert lib.tfc_memcpy_h2d(c, hs[i % nbuf], 0, src.ctypes.data, each) == 0
    assert lib.tfc_sync(c) == 0
    dt = time.perf_counter() - t0
    out["h2d_bulk_GBps"] = round(ncopies * each / dt / 1e9, 2)
    dst = np.empty(each, dtype=np.uint8)
    t0 = time.perf_counter()
    for i in range(16):
        assert lib.tfc_memcpy_d2h(c, dst.ctypes.data, hs[i % nbuf], 0, each) == 0
    out["d2h_bulk_GBps"] = round(16 * each / (time.perf_counter() - t0) / 1e9, 2)
    assert np.array_equal(dst, src)
    calls = 20000
    t0 = time.perf_counter()
    for i in range(calls):
        assert lib.tfc_memcpy_h2d(c, hs[i % nbuf], (i * 4096) % (each - 4096), src.ctypes.data, 4096) == 0
    assert lib.tfc_sync(c) == 0
    out["h2d_4KiB_us_per_call"] = round((time.perf_counter() - t0) / calls * 1e6, 3)
    t0 = time.perf_counter()
    for i in range(200):
        assert lib.tfc_sync(c) == 0
    out["sync_round_trip_us"] = round((time.perf_counter() - t0) / 200 * 1e6, 1)
    lib.tfc_close(c)
    proc.wait(timeout=60)
    return out


res = {}
p = subprocess.Popen([EXE, "-p", "0"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                     env=dict(os.environ, TFW_ONESHOT="1", TFW_BIND="127.0.0.1"))
port = int(p.stdout.readline().split()[-1])
res["tcp_loopback"] = run(f"native+127.0.0.1+{port}+bench-1", p)
with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
    os.environ["TFC_SHM_DIR"] = d
    p = subprocess.Popen([EXE, "-n", "shmem", "-m", "tf_shm", "-M", "1024"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                         env=dict(os.environ, TFW_ONESHOT="1", TFW_SHM_DIR=d))
    assert "serving shmem" in p.stdout.readline()
    res["shmem_1024MiB"] = run("shmem+tf_shm+1024+1", p)
print(json.dumps(res))
