"""The C2 stream through the REAL process boundary: a client process speaking TFCS through libtfc_client.so to
`tensor-fusion-worker` (the executable the operator starts, compose.go:1304-1325) over
  (a) the page-locked shared-memory rings  (-n shmem, "shmem+<name>+<MiB>+1", pod_webhook.go:584) and
  (b) TCP loopback                         (-p <port>, "native+<ip>+<port>+...", tensorfusionconnection_controller.go:136-138)
next to the identical calls issued straight to the CUDA runtime by this process (tfw_native_copy /
tfw_native_replay): north_star's "<= 4 % added wall-clock over native CUDA".

Bulk legs: 256 copies of 64 MiB (16 GiB) cycling over 16 host buffers and 16 device buffers,
  host -> device and device -> host, with the host side (i) pageable and (ii) page-locked -- natively
  cudaHostAlloc, through the worker an arena from tfc_host_alloc, which the worker maps and page-locks too, so
  the copy engine moves the bytes between the client's own pages and HBM with no CPU copy at all.
Latency legs: 4 KiB H2D calls (C loop, tools/transport_lab.c) and the synchronise round trip.

Used by bench.py (overhead_vs_native.through_worker_*); runs stand-alone too: prints one JSON object."""
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tensor-fusion_b200", "lib", "tensor-fusion-worker")
LAB = os.path.join(ROOT, "build", "mock", "transport_lab")
MIB = 1 << 20


def client_lib():
    lib = C.CDLL(os.path.join(ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so"))
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
    return lib


def start_worker(transport, shm_dir, ring_mib, device):
    env = dict(os.environ, TFW_ONESHOT="-1", TFW_BIND="127.0.0.1", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(device)),
               TFW_SHM_DIR=shm_dir)      # (both ends default to /dev/shm; the bench keeps its files in a directory of its own)
    if transport == "tcp":
        p = subprocess.Popen([EXE, "-p", "0"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env)
        return p, f"native+127.0.0.1+{int(p.stdout.readline().split()[-1])}+bench-1"
    env["TFW_SHM_DIR"] = shm_dir
    p = subprocess.Popen([EXE, "-n", "shmem", "-m", "tf_shm", "-M", str(ring_mib)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env)
    line = p.stdout.readline()
    assert "serving shmem" in line, line
    return p, f"shmem+tf_shm+{ring_mib}+1"


def through_worker(lib, url, each, nsrc, nbuf, ncopies, passes, arena, check=True):
    """One client session: bulk H2D and D2H legs; returns seconds per pass for each and whether the bytes survived."""
    c = C.c_void_p()
    assert lib.tfc_connect(url.encode(), C.byref(c)) == 0, url
    hs = []
    for _ in range(nbuf):
        h = C.c_uint32()
        assert lib.tfc_malloc(c, each, C.byref(h)) == 0
        hs.append(h)
    assert lib.tfc_sync(c) == 0                       # context creation, allocation and scrubbing are not part of the numbers
    if arena:
        base = C.c_void_p()
        rc = lib.tfc_host_alloc(c, nsrc * each, C.byref(base))
        if rc != 0:
            lib.tfc_close(c)
            return {"unavailable": f"tfc_host_alloc -> {rc} (no shared memory on this transport, or /dev/shm too small)"}
        host = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), (nsrc * each,))
    else:
        host = np.empty(nsrc * each, dtype=np.uint8)
    pat = np.random.default_rng(1).integers(0, 256, each, dtype=np.uint8)
    for i in range(nsrc):
        host[i * each:(i + 1) * each] = pat ^ np.uint8(i)        # also the first touch of every page
    addr = host.ctypes.data
    out = {}
    up, down = [], []
    for p in range(passes + 1):                                      # pass 0 = warm-up
        t0 = time.perf_counter()
        for i in range(ncopies):
            assert lib.tfc_memcpy_h2d(c, hs[i % nbuf], 0, addr + (i % nsrc) * each, each) == 0
        assert lib.tfc_launch(c, 0, 1, 32, 0, 0, 0, 0, 0) == 0     # the stream's noop launch
        assert lib.tfc_sync(c) == 0
        if p:
            up.append(time.perf_counter() - t0)
    ok = True
    if check:                                                        # device buffer k now holds source (last copy into it)
        probe = np.empty(each, dtype=np.uint8)
        for k in (0, nbuf - 1):
            last = max(i for i in range(ncopies) if i % nbuf == k)
            assert lib.tfc_memcpy_d2h(c, probe.ctypes.data, hs[k], 0, each) == 0
            ok &= bool(np.array_equal(probe, pat ^ np.uint8(last % nsrc)))
    host[:] = 0
    for p in range(passes + 1):
        t0 = time.perf_counter()
        for i in range(ncopies):
            dst = addr + (i % nsrc) * each
            if arena:
                assert lib.tfc_memcpy_d2h_async(c, dst, hs[i % nbuf], 0, each) == 0
            else:
                assert lib.tfc_memcpy_d2h(c, dst, hs[i % nbuf], 0, each) == 0
        assert lib.tfc_sync(c) == 0
        if p:
            down.append(time.perf_counter() - t0)
    if check:
        for k in (0, nsrc - 1):
            last = max(i for i in range(ncopies) if i % nsrc == k)      # host buffer k got device buffer (last % nbuf)
            src_dev = last % nbuf
            last_up = max(i for i in range(ncopies) if i % nbuf == src_dev)
            ok &= bool(np.array_equal(host[k * each:(k + 1) * each], pat ^ np.uint8(last_up % nsrc)))
    if arena:
        del host
        lib.tfc_host_free(c, base)
    lib.tfc_close(c)
    out["h2d_s"], out["d2h_s"], out["bytes_ok"] = min(up), min(down), ok
    return out


def native(N, device, direction, pinned, each, nsrc, nbuf, ncopies, passes):
    s = C.c_double()
    rc = N.lib.tfw_native_copy(device, direction, 1 if pinned else 0, each, nsrc, nbuf, ncopies, passes, C.byref(s))
    assert rc == 0, rc
    return s.value


def lab(url, calls, nbytes, shm_dir):
    env = dict(os.environ, TFC_SHM_DIR=shm_dir)
    r = subprocess.run([LAB, url, str(calls), str(nbytes)], capture_output=True, text=True, timeout=300, env=env)
    return json.loads(r.stdout) if r.returncode == 0 and r.stdout.strip() else {"error": r.stderr[-300:]}


def run(device=0, each=64 * MIB, nsrc=16, nbuf=16, ncopies=256, passes=2, ring_mib=1024, latency_calls=200000, tcp_copies=64):
    sys.path.insert(0, ROOT)
    from tensor_fusion_b200 import _native as N
    lib = client_lib()
    total = ncopies * each
    res = {"workload": f"{ncopies} x {each // MIB} MiB copies ({total / 2**30:.0f} GiB) over {nsrc} host and {nbuf} device buffers + noop launch + sync",
           "client": "libtfc_client.so in this process", "worker": "tensor-fusion-worker, separate process", "copy_threads": os.environ.get("TFC_COPY_THREADS", "default")}
    gb = lambda s, n=total: round(n / s / 1e9, 2)
    pct = lambda w, n: round((w / n - 1) * 100, 2)
    nat = {}
    for direction, dname in ((0, "h2d"), (1, "d2h")):
        for pinned in (False, True):
            nat[(dname, pinned)] = native(N, device, direction, pinned, each, nsrc, nbuf, ncopies, passes)
    res["native"] = {f"{d}_{'pinned' if p else 'pageable'}_GBps": gb(s) for (d, p), s in nat.items()}
    nat_small = native(N, device, 0, False, 4096, 1, 1, latency_calls, 1)      # the same 4 KiB calls issued to the CUDA runtime (pageable source)
    res["native"]["h2d_4KiB_us_per_call"] = round(nat_small / latency_calls * 1e6, 3)
    shm_dir = tempfile.mkdtemp(dir="/dev/shm", prefix="tfw-bench-")
    os.environ["TFC_SHM_DIR"] = shm_dir
    try:
        w, url = start_worker("shmem", shm_dir, ring_mib, device)
        try:
            legs = {}
            for arena in (False, True):
                r = through_worker(lib, url, each, nsrc, nbuf, ncopies, passes, arena)
                kind = "pinned" if arena else "pageable"
                if "unavailable" in r:
                    legs[kind] = r
                    continue
                legs[kind] = {"h2d_GBps": gb(r["h2d_s"]), "d2h_GBps": gb(r["d2h_s"]), "bytes_ok": r["bytes_ok"],
                              "h2d_added_percent": pct(r["h2d_s"], nat[("h2d", arena)]), "d2h_added_percent": pct(r["d2h_s"], nat[("d2h", arena)]),
                              "host_memory": "arena from tfc_host_alloc: page-locked by the worker, DMA on the client's pages (no CPU copy)" if arena
                                             else "pageable numpy: copied into / out of the page-locked rings by the client's copy threads"}
            legs["latency"] = lab(url, latency_calls, 4096, shm_dir)
            if "h2d_us_per_call" in legs["latency"]:
                legs["latency"]["h2d_4KiB_added_percent_vs_native"] = pct(legs["latency"]["h2d_us_per_call"], res["native"]["h2d_4KiB_us_per_call"])
            res["through_worker_shm"] = legs
            # the headline of the boundary: bulk H2D with the application's memory page-locked, as the native 55 GB/s figure needs too
            if "h2d_added_percent" in legs.get("pinned", {}):
                res["through_worker_shm"]["added_percent"] = legs["pinned"]["h2d_added_percent"]
        finally:
            w.terminate()
            w.wait(timeout=30)
        # the same URL a remote client dials, on the worker's own node: the connection upgrades itself to shared-memory
        # rings (TFCS_OP_UPGRADE_SHM), so "over loopback" costs what the rings cost
        w, url = start_worker("tcp", shm_dir, ring_mib, device)
        try:
            os.environ["TFC_UPGRADE_MIB"] = str(ring_mib)
            legs = {}
            for arena in (False, True):
                r = through_worker(lib, url, each, nsrc, nbuf, ncopies, passes, arena)
                kind = "pinned" if arena else "pageable"
                legs[kind] = r if "unavailable" in r else {
                    "h2d_GBps": gb(r["h2d_s"]), "d2h_GBps": gb(r["d2h_s"]), "bytes_ok": r["bytes_ok"],
                    "h2d_added_percent": pct(r["h2d_s"], nat[("h2d", arena)]), "d2h_added_percent": pct(r["d2h_s"], nat[("d2h", arena)])}
            legs["url"] = "native+127.0.0.1+<port>+... (TCP connect, session moved onto shared-memory rings by TFCS_OP_UPGRADE_SHM)"
            if "h2d_added_percent" in legs.get("pinned", {}):
                legs["added_percent"] = legs["pinned"]["h2d_added_percent"]
            res["through_worker_loopback_upgraded"] = legs
        finally:
            w.terminate()
            w.wait(timeout=30)
        os.environ["TFC_NO_SHM_UPGRADE"] = "1"      # and the socket path itself, bytes through the TCP stack
        w, url = start_worker("tcp", shm_dir, ring_mib, device)
        try:
            r = through_worker(lib, url, each, nsrc, nbuf, tcp_copies, 1, False)
            n = tcp_copies * each
            nat_h = native(N, device, 0, False, each, nsrc, nbuf, tcp_copies, 1)
            nat_d = native(N, device, 1, False, each, nsrc, nbuf, tcp_copies, 1)
            res["through_worker_tcp_loopback"] = {"h2d_GBps": gb(r["h2d_s"], n), "d2h_GBps": gb(r["d2h_s"], n), "bytes_ok": r["bytes_ok"], "copies": tcp_copies,
                                                  "h2d_added_percent": pct(r["h2d_s"], nat_h), "d2h_added_percent": pct(r["d2h_s"], nat_d),
                                                  "bound": "two kernel socket copies per byte on one core each; the cross-node transport, where the NIC is the bound",
                                                  "latency": lab(url, min(latency_calls, 50000), 4096, shm_dir)}
        finally:
            w.terminate()
            w.wait(timeout=30)
    finally:
        shutil.rmtree(shm_dir, ignore_errors=True)
    return res


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--copies", type=int, default=256)
    ap.add_argument("--each-mib", type=int, default=64)
    ap.add_argument("--passes", type=int, default=2)
    a = ap.parse_args()
    print(json.dumps(run(device=a.device, each=a.each_mib * MIB, ncopies=a.copies, passes=a.passes)))
