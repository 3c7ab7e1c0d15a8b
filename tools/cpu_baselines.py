"""CPU baselines of BASELINE.md section 3 (context, not targets), printed as JSON lines.
B1 reference mock driver handling the control calls of the C1 trace (oracle/_ref/libdriver_mock.so)
B2 oracle replay payload legs (memcpy GB/s), 1 thread and all cores
B3 FetchSubERLTokens restatement: ops/s, 1 and 4 contending threads
B4 provider ABI round trips the 2 Hz loops issue: reference stub vs this repo's NVML provider
This file is measurement tooling: it may load oracle/ (same rule as bench.py's cpu_baseline leg)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from tensor_fusion_b200 import trace, wire  # noqa: E402

ncpu = os.cpu_count() or 1


def b1():
    so = os.path.join(ROOT, "oracle", "_ref", "libdriver_mock.so")
    if not os.path.exists(so):
        return {"baseline": "B1", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
    cwd = os.getcwd()
    os.chdir(os.path.join(ROOT, "oracle", "_ref"))      # the mock keeps its state file in the cwd
    try:
        m = C.CDLL(so)
        m.hipInit(0)
        frames = [h for h, _ in wire.parse_frames(trace.gen_c1())]
        ptrs, n = {}, 0
        t0 = time.perf_counter()
        for rep in range(20):
            for h in frames:
                if h["opcode"] == wire.OP_MALLOC:
                    p = C.c_void_p()
                    if m.hipMalloc(C.byref(p), C.c_size_t(h["length"])) == 0:
                        ptrs[h["h0"]] = p
                    n += 1
                elif h["opcode"] == wire.OP_FREE and h["h0"] in ptrs:
                    m.hipFree(ptrs.pop(h["h0"]))
                    n += 1
                elif h["opcode"] == wire.OP_LAUNCH:
                    m.hipLaunchKernel(None, h["arg1"], 1, 1, h["arg2"], 1, 1, 0, None, None, None)   # 100/s cap applies
                    n += 1
            for p in ptrs.values():
                m.hipFree(p)
            ptrs.clear()
        dt = time.perf_counter() - t0
        return {"baseline": "B1", "what": "reference mock driver hipMalloc/hipFree/hipLaunchKernel on the C1 trace's control calls (via ctypes)",
                "calls": n, "calls_per_s": round(n / dt), "us_per_call": round(dt / n * 1e6, 3), "threads": 1}
    finally:
        os.chdir(cwd)


def b2():
    out = []
    raw = trace.gen_bulk(8, 16, 64 << 20, nthreads=min(32, ncpu))
    for th in (1, ncpu):
        oracle.lib.tfo_set_threads(th)
        oracle.Replay(raw).close()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < 6:
            oracle.Replay(raw).close()
            n += 1
        dt = time.perf_counter() - t0
        out.append({"baseline": "B2", "what": "oracle replay of 16 x 64 MiB H2D (memcpy into calloc'd buffers)", "threads": th,
                    "payload_GBps": round(n * (1 << 30) / dt / 1e9, 2), "read_write_GBps": round(2 * n * (1 << 30) / dt / 1e9, 2)})
    oracle.lib.tfo_set_threads(1)
    return out


def b3():
    f = oracle.lib.tfo_gate_bench
    f.restype, f.argtypes = C.c_double, [C.c_int, C.c_uint64, C.c_double, C.POINTER(C.c_double)]
    out = []
    for th in (1, 4):
        deny = C.c_double()
        ops = f(th, 10_000_000 // th, 1.0, C.byref(deny))
        out.append({"baseline": "B3", "what": "FetchSubERLTokens restatement, cost 1.0, refill thread at 2 Hz", "threads": th,
                    "ops_per_s": round(ops), "deny_ratio": round(deny.value, 4)})
    return out


def b4():
    from tensor_fusion_b200 import provider as P
    out = []
    for name, path in (("reference stub provider", os.path.join(ROOT, "oracle", "_ref", "libaccelerator_example.so")), ("libaccelerator_b200 (NVML)", P.LIB_PATH)):
        if not os.path.exists(path):
            continue
        cwd = os.getcwd()
        os.chdir(os.path.dirname(path))
        try:
            lib = P.load(path)
            if lib.AccelInit() != P.SUCCESS:
                out.append({"baseline": "B4", "provider": name, "unavailable": "AccelInit failed (no driver here)"})
                continue
            rc, devs = P.all_devices(lib)
            k = max(1, min(4, len(devs)))
            uu = (C.c_char_p * k)(*[d["uuid"].encode() for d in devs[:k]])
            dm = (P.DeviceMetrics * k)()
            pi = (P.ProcessInformation * 1024)()
            n = C.c_size_t()
            t0 = time.perf_counter()
            for _ in range(200):
                lib.AccelGetDeviceMetrics(uu, k, dm)
            t_m = (time.perf_counter() - t0) / 200
            t0 = time.perf_counter()
            for _ in range(200):
                lib.AccelGetProcessInformation(pi, 1024, C.byref(n))
            t_p = (time.perf_counter() - t0) / 200
            out.append({"baseline": "B4", "provider": name, "devices": k, "AccelGetDeviceMetrics_us": round(t_m * 1e6, 1),
                        "AccelGetProcessInformation_us": round(t_p * 1e6, 1)})
        finally:
            os.chdir(cwd)
    return out


if __name__ == "__main__":
    print(json.dumps({"host_cores": ncpu}))
    for fn in (b1, b2, b3, b4):
        r = fn()
        for row in (r if isinstance(r, list) else [r]):
            print(json.dumps(row), flush=True)
