"""BASELINE config 3: 4 vGPU workers @25 % on one B200 under the ERL limiter.

The parent plays the hypervisor: every 500 ms it reads utilisation through the provider ABI
(AccelGetDeviceMetrics / AccelGetProcessInformation) and runs the reference's controller step
through LimiterUpdateERL for every worker -- exactly the loop of
pkg/hypervisor/worker/computing/quota_controller.go:378-458.  Each child is one vGPU worker
(libtfw_b200.so + quota file) issuing a saturating stream of 200 us spin kernels whose
launches are gated by the device-resident token bucket.

feedback=device : the reference's semantics (whole-device NVML utilisation vs each worker's target)
feedback=process: per-process SM utilisation as the feedback signal (what a per-vGPU share needs)
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker_main(idx, shm_path, seconds, kernel_us, cost, q, start_evt, limiter):
    import numpy as np
    from tensor_fusion_b200 import _native as N
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    w = Worker(shm_path=shm_path if limiter else None, flags=0 if limiter else N.TFW_F_NO_LIMITER)
    batch = wire.Builder()
    nb = 16
    for _ in range(nb):
        batch.launch(wire.K_SPIN, grid=1, block=32, scalar=kernel_us * 1000, cost=cost)
    raw = np.frombuffer(bytes(batch), dtype=np.uint8)
    w.submit(raw); w.flush()                      # warm-up
    q.put({"ready": idx})
    start_evt.wait()
    t0 = time.time()
    done, lat, late, slow = 0, [], [], []
    while time.time() - t0 < seconds:
        t1 = time.time()
        w.submit(raw)
        w.flush()
        t2 = time.time()
        lat.append((t2 - t1) / nb)
        if t1 - t0 >= seconds / 2:
            late.append(lat[-1])                   # the controller has settled: steady state
        if limiter and t2 - t1 > 0.1 and len(slow) < 12:   # a stall: what did the bucket look like?
            g = w.gate_state()
            slow.append({"at_s": round(t1 - t0, 2), "batch_ms": round((t2 - t1) * 1e3, 1), "rate": round(g["refill_rate"], 1),
                         "bucket": round(g["tokens"], 1), "capacity": round(g["capacity"], 1)})
        done += nb
    dt = time.time() - t0
    lat.sort()
    late.sort()
    st = w.gate_state() if limiter else {}
    q.put({"worker": idx, "pid": os.getpid(), "launches": done, "seconds": round(dt, 2), "gate_timeouts": st.get("timeouts", 0),
           "throttled_gates": st.get("blocked_gates", 0),
           "busy_share_percent": round(done * kernel_us * 1e-6 / dt * 100, 2),
           "per_launch_ms_p50": round(lat[len(lat) // 2] * 1e3, 4), "per_launch_ms_p99": round(lat[int(len(lat) * 0.99)] * 1e3, 4),
           "per_launch_ms_mean": round(sum(lat) / len(lat) * 1e3, 4), "batches": len(lat),
           "steady_per_launch_ms_p99": round(late[int(len(late) * 0.99)] * 1e3, 4) if late else None,
           "steady_per_launch_ms_max": round(late[-1] * 1e3, 4) if late else None,
           "steady_per_launch_ms_p50": round(late[len(late) // 2] * 1e3, 4) if late else None,
           # a batch that waited more than 100 ms sat out (most of) a controller tick
           "steady_stalled_batches_percent": round(100.0 * sum(1 for x in late if x * nb > 0.1) / len(late), 2) if late else None,
           "stalls": slow})
    w.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--limit", type=int, default=25)
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--kernel-us", type=int, default=200)
    ap.add_argument("--cost", type=int, default=1)
    ap.add_argument("--feedback", choices=["device", "process"], default="device")
    ap.add_argument("--no-limiter", action="store_true")
    a = ap.parse_args()
    mp.set_start_method("spawn")
    from tensor_fusion_b200 import provider as P
    lib = P.load()
    assert lib.AccelInit() == P.SUCCESS
    rc, devs = P.all_devices(lib)
    uuid = devs[0]["uuid"].encode()
    base = tempfile.mkdtemp(prefix="tf_c3_")
    assert lib.LimiterInit(base.encode()) == P.SUCCESS
    cfg = (P.LimiterDeviceConfig * 1)()
    cfg[0].deviceIdx, cfg[0].deviceUUID, cfg[0].upLimit, cfg[0].memLimit = 0, uuid, a.limit, 40 << 30
    q, evt, procs = mp.Queue(), mp.Event(), []
    for i in range(a.workers):
        assert lib.LimiterCreateWorker(b"c3", f"w{i}".encode(), cfg, 1) == P.SUCCESS
        p = mp.Process(target=worker_main, args=(i, os.path.join(base, "c3", f"w{i}", "shm"), a.seconds, a.kernel_us, a.cost, q, evt, not a.no_limiter))
        p.start()
        procs.append(p)
    ready = 0
    while ready < a.workers:                        # children create their CUDA contexts and warm up
        if "ready" in q.get(timeout=120):
            ready += 1
    evt.set()
    uu = (C.c_char_p * 1)(uuid)
    dm = (P.DeviceMetrics * 1)()
    pi = (P.ProcessInformation * 1024)()
    n = C.c_size_t()
    utils, ticks, t_begin, t_end = [], [], time.time(), time.time() + a.seconds
    pids = {p.pid: i for i, p in enumerate(procs)}
    while time.time() < t_end:
        time.sleep(0.5)
        lib.AccelGetDeviceMetrics(uu, 1, dm)
        dev_util = float(dm[0].utilizationPercent)
        per = {}
        if a.feedback == "process":
            lib.AccelGetProcessInformation(pi, 1024, C.byref(n))
            for k in range(n.value):
                pid = int(pi[k].processId.decode())
                if pid in pids:
                    per[pids[pid]] = pi[k].computeUtilizationPercent
        utils.append(dev_util)
        nsamp = next((dm[0].extraMetrics[k].value for k in range(dm[0].extraMetricsCount) if dm[0].extraMetrics[k].key == b"utilizationSamplesAveraged"), -1)
        ticks.append([round(time.time() - t_begin, 2), round(dev_util, 1), int(nsamp)] + [round(per.get(i, 0.0), 1) for i in range(a.workers) if a.feedback == "process"])
        now_us = int(time.time() * 1e6)
        for i in range(a.workers):
            u = per.get(i, 0.0) if a.feedback == "process" else dev_util
            lib.LimiterUpdateERL(b"c3", f"w{i}".encode(), 0, a.limit, u, now_us)
            lib.LimiterUpdateHeartbeat(b"c3", f"w{i}".encode(), int(time.time()))
    res = sorted((q.get(timeout=60) for _ in procs), key=lambda r: r["worker"])
    for p in procs:
        p.join(timeout=30)
    tail = utils[len(utils) // 2:]
    shares = [r["busy_share_percent"] for r in res]
    mean = sum(shares) / len(shares)
    out = {"share_percent_each": shares, "share_mean_percent": round(mean, 2), "target_percent": a.limit,
           "share_error_vs_equal_percent": round(max(abs(x - mean) for x in shares) / mean * 100, 2) if mean else None,
           "share_error_vs_target_points": round(max(abs(x - a.limit) for x in shares), 2),
           "per_launch_ms_p50_max": max(r["per_launch_ms_p50"] for r in res), "per_launch_ms_p99_max": max(r["per_launch_ms_p99"] for r in res),
           "per_launch_ms_mean_max": max(r["per_launch_ms_mean"] for r in res),
           "steady_per_launch_ms_p99_max": max((r["steady_per_launch_ms_p99"] or 0.0) for r in res),
           "steady_per_launch_ms_p50_max": max((r["steady_per_launch_ms_p50"] or 0.0) for r in res),
           "steady_stalled_batches_percent_max": max((r["steady_stalled_batches_percent"] or 0.0) for r in res),
           "ticks_t_util_nsamples": ticks,
           "gate_timeouts": sum(r.get("gate_timeouts", 0) for r in res),
           "config": f"{a.workers} vGPU @ {a.limit} %, {a.kernel_us} us spin kernels, cost {a.cost} token/launch, feedback={a.feedback}, limiter={'off' if a.no_limiter else 'on'}",
           "device_util_percent_mean_2nd_half": round(sum(tail) / max(1, len(tail)), 1), "workers": res,
           "total_busy_share_percent": round(sum(r["busy_share_percent"] for r in res), 2)}
    print(json.dumps(out), flush=True)
    lib.LimiterShutdown()


if __name__ == "__main__":
    main()
