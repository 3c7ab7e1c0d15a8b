"""Two peer-tier migrations for ncu (run under `gpurun --gpus 2`): K x 1 GiB regions homed on GPU 0 are evicted to
GPU 1 (one mover launch ON GPU 1: it pulls the bytes over NVLink) and prefetched back (one launch ON GPU 0).
    ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum -k regex:tfw_mover --csv python tools/peer_ncu_probe.py
The evict launch should show nvlrx ~ K GiB on GPU 1, the prefetch launch nvlrx ~ K GiB on GPU 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensor_fusion_b200 import vram as V  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
R = 1 << 30
flags = (V.COPY_ENGINE if "--ce" in sys.argv else 0) | (V.SENDER_DRIVEN if "--sender" in sys.argv else 0)
with V.VSpace(home=0, va_bytes=K * R, region_bytes=R, home_budget=K * R, peer_budget=K * R, peers=[1], flags=flags) as vs:
    for r in range(K):
        vs.populate(r, V.HOME)
        vs.fill_pattern(r, 31 + r)
    want = [vs.digest(r) for r in range(K)]
    out = []
    for rep in range(3):
        ev = vs.migrate(list(range(K)), [V.PEER] * K, [0] * K)
        pf = vs.migrate(list(range(K)), [V.HOME] * K)
        out.append({"evict_GBps": round(K * R / ev["copy_ms"] / 1e6, 1), "prefetch_GBps": round(K * R / pf["copy_ms"] / 1e6, 1)})
    assert [vs.digest(r) for r in range(K)] == want
print(json.dumps({"regions_gib": K, "copy_engine": bool(flags & V.COPY_ENGINE), "sender_driven": bool(flags & V.SENDER_DRIVEN), "reps": out}))
