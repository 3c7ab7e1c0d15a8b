"""Mover kernel sweep on one GPU: GB/s (algorithmic 2N / CUDA-event time) for both
mover variants, several CTAs/SM, aligned / misaligned / fill.  Prints JSON lines."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor_fusion_b200 import _native as N
from tensor_fusion_b200.worker import Worker

PEAK = 6591.9
total = 4 << 30  # 4 GiB per launch: far larger than the 126 MB L2
each = 64 << 20


def run(mode, ctas, kind):
    flags = N.TFW_F_MOVER_TMA if mode == "tma" else N.TFW_F_MOVER_LDG
    with Worker(flags=flags, ctas_per_sm=ctas) as w:
        s = w.dev_alloc(total + 256)
        d = w.dev_alloc(total + 256)
        n = total // each
        if kind == "aligned":
            descs = [(d + i * each, s + i * each, each, 0) for i in range(n)]
            algo = 2 * total
        elif kind == "misaligned":
            descs = [(d + i * each + 3, s + i * each + 9, each - 16, 0) for i in range(n)]
            algo = 2 * (each - 16) * n
        elif kind == "dstmis":
            descs = [(d + i * each + 5, s + i * each + 5, each - 16, 0) for i in range(n)]
            algo = 2 * (each - 16) * n
        else:
            descs = [(d + i * each, 0, each, 0x11) for i in range(n)]
            algo = total
        for _ in range(3):
            w.move_batch(descs, timed=True)
        ts = [w.move_batch(descs, timed=True) for _ in range(5)]
        best, med = min(ts), sorted(ts)[len(ts) // 2]
        w.dev_free(s)
        w.dev_free(d)
    gbs = algo / (med * 1e-3) / 1e9
    print(json.dumps({"mover": mode, "ctas_per_sm": ctas, "kind": kind, "ms_med": round(med, 4), "ms_best": round(best, 4),
                      "GBps_algorithmic": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 4)}), flush=True)


if __name__ == "__main__":
    for kind in ("aligned", "misaligned", "dstmis", "fill"):
        for mode, ctas_list in (("ldg", (0, 3)), ("tma", (0,))):   # 0 = one tile per CTA (one-shot); > 0 = persistent grid (vector kernel only)
            for c in ctas_list:
                try:
                    run(mode, c, kind)
                except Exception as e:  # keep sweeping
                    print(json.dumps({"mover": mode, "ctas_per_sm": c, "kind": kind, "error": str(e)}), flush=True)
