"""BASELINE config 4: one vGPU asking for more VRAM than the GPU has (default 256 GiB on a 180 GB
B200): 1 GiB regions, ~168 GiB resident budget, the rest in pinned host DRAM.
Access pattern: 3 sequential sweeps, then Zipf(1.1) over regions, seed 42.
Every region's digest is checked each time it is touched (it is HOME-resident then)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from tensor_fusion_b200 import vram as V  # noqa: E402

GIB = 1 << 30


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--va-gib", type=int, default=256)
    ap.add_argument("--home-gib", type=int, default=168)
    ap.add_argument("--host-gib", type=int, default=96)
    ap.add_argument("--zipf", type=int, default=400)
    ap.add_argument("--peers", type=int, default=0, help="number of peer GPUs (1..N) holding cold regions (config C5)")
    ap.add_argument("--peer-gib", type=int, default=0, help="HBM budget per peer GPU")
    ap.add_argument("--sweeps", type=int, default=3)
    a = ap.parse_args()
    n = a.va_gib
    t0 = time.time()
    with V.VSpace(home=0, va_bytes=n * GIB, region_bytes=GIB, home_budget=a.home_gib * GIB, host_budget=a.host_gib * GIB,
                  peers=list(range(1, a.peers + 1)), peer_budget=a.peer_gib * GIB) as vs:
        setup_s = time.time() - t0
        want = {}
        t0 = time.time()
        for r in range(n):               # first touch: populate (evicting LRU regions to host once HBM is full)
            vs.access(r)
            vs.fill_pattern(r, 4200 + r)
            want[r] = vs.digest(r)
        populate_s = time.time() - t0
        s0 = vs.stats()
        sweeps = []
        for sweep in range(a.sweeps):
            t0 = time.time()
            for r in range(n):
                vs.access(r)
                assert vs.digest(r) == want[r], f"sweep {sweep} region {r} corrupted"
            sweeps.append(time.time() - t0)
        s1 = vs.stats()
        rng = np.random.default_rng(42)
        seq = [int(min(n - 1, z - 1)) for z in rng.zipf(1.1, a.zipf)]
        t0 = time.time()
        for r in seq:
            vs.access(r)
            assert vs.digest(r) == want[r]
        zipf_s = time.time() - t0
        s2 = vs.stats()
    moved = (s1["evict_bytes_host"] - s0["evict_bytes_host"]) + (s1["prefetch_bytes_host"] - s0["prefetch_bytes_host"])
    moved_peer = (s1["evict_bytes_peer"] - s0["evict_bytes_peer"]) + (s1["prefetch_bytes_peer"] - s0["prefetch_bytes_peer"])
    out = {"config": f"1 vGPU, {n} GiB VA, home GPU0 {a.home_gib} GiB HBM budget, {a.peers} peer GPUs x {a.peer_gib} GiB, host tier {a.host_gib} GiB, 1 GiB regions",
           "sweep_bytes_over_nvlink_both_directions": moved_peer,
           "sweep_swap_GBps_nvlink_both_directions": round(moved_peer / max(1e-9, sum(sweeps)) / 1e9, 1),
           "setup_s": round(setup_s, 1), "populate_s": round(populate_s, 1), "sweep_s": [round(x, 2) for x in sweeps],
           "sweep_bytes_over_pcie_both_directions": moved,
           "sweep_swap_GBps_both_directions": round(moved / sum(sweeps) / 1e9, 1),
           "zipf": {"accesses": len(seq), "seconds": round(zipf_s, 2), "hits": s2["policy_hits"] - s1["policy_hits"],
                    "prefetches": s2["policy_prefetches"] - s1["policy_prefetches"]},
           "all_digests_verified": True, "regions_home": s2["regions_home"], "regions_peer": s2["regions_peer"], "regions_host": s2["regions_host"]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
