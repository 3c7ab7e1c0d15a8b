"""VRAM-tier swap bandwidth on one box, one process: home GPU 0, peers 1..N-1 (and host DRAM).
Prints JSON lines: evict / prefetch GB/s (CUDA-event copy time, and wall-clock incl. re-mapping)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (device count only)
from tensor_fusion_b200 import vram as V  # noqa: E402

GIB = 1 << 30
ndev = torch.cuda.device_count()
R = int(os.environ.get("REGION_MIB", "1024")) << 20
K = int(os.environ.get("REGIONS", "16"))


def run(peers, flags, label, tier):
    with V.VSpace(home=0, va_bytes=K * R, region_bytes=R, home_budget=K * R, peer_budget=K * R,
                  host_budget=(K * R if tier == V.HOST else 0), peers=peers, flags=flags) as vs:
        for r in range(K):
            vs.populate(r, V.HOME)
            vs.fill_pattern(r, 77 + r)
        want = [vs.digest(r) for r in (0, K - 1)]
        slots = [r % max(1, len(peers)) for r in range(K)]
        out = {"leg": label, "regions": K, "region_mib": R >> 20, "peers": peers,
               "mover": "copy_engine" if flags & 1 else "tma_kernel" if flags & 2 else "kernel"}
        for rep in range(3):
            ev = vs.migrate(list(range(K)), [tier] * K, slots)
            if tier == V.PEER:
                assert vs.digest(K - 1) == want[1]      # read back through NVLink
            pf = vs.migrate(list(range(K)), [V.HOME] * K)
            assert [vs.digest(0), vs.digest(K - 1)] == want
        out.update({"evict_GBps_copy": round(ev["bytes"] / ev["copy_ms"] / 1e6, 1), "evict_GBps_wall": round(ev["bytes"] / ev["total_ms"] / 1e6, 1),
                    "prefetch_GBps_copy": round(pf["bytes"] / pf["copy_ms"] / 1e6, 1), "prefetch_GBps_wall": round(pf["bytes"] / pf["total_ms"] / 1e6, 1),
                    "launches_per_batch": ev["launches"]})
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    run([], 0, "host_tier_pcie", V.HOST)
    for n in (2, 4, 8):
        if ndev >= n:
            run(list(range(1, n)), 0, f"peer_tier_{n}gpu", V.PEER)
            run(list(range(1, n)), 1, f"peer_tier_{n}gpu", V.PEER)
            run(list(range(1, n)), 2, f"peer_tier_{n}gpu", V.PEER)
