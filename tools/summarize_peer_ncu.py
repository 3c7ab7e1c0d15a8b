#!/usr/bin/env python
"""gpurun_out/<tag>_peer_ncu.csv (ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,... of tools/peer_ncu_probe.py)
-> profiles/<tag>_peer_ncu.md: one row per captured mover launch with its NVLink bytes, duration and GB/s."""
import collections
import csv
import json
import os
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", f"{tag}_peer_ncu.csv")
lines = [l for l in open(src) if not l.startswith("==")]
launches = collections.OrderedDict()
for row in csv.DictReader(lines):
    key = (row["ID"], row["Kernel Name"].split("(")[0], row.get("Device", row.get("Context", "")))
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = row["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "second": 1}.get(unit, 1)
    launches.setdefault(key, {})[row["Metric Name"]] = v * scale
probe = {}
try:
    probe = json.loads(open(os.path.join(ROOT, "gpurun_out", f"{tag}_peer_ncu_probe.json")).read().strip().splitlines()[-1])
except Exception:
    pass
out = os.path.join(ROOT, "profiles", f"{tag}_peer_ncu.md")
with open(out, "w") as f:
    f.write(f"# {tag}: NVLink counters of the peer-tier mover (ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,..., --clock-control none)\n\n"
            "`tools/peer_ncu_probe.py 2` under `CUDA_VISIBLE_DEVICES=0,1`: 2 x 1 GiB regions homed on GPU 0 are evicted to GPU 1 (each region one mover\n"
            "launch ON GPU 1, which pulls the bytes over NVLink) and prefetched back (launches ON GPU 0).  `nvlrx` = bytes the launching GPU received\n"
            "over NVLink during the launch (payload + protocol overhead), `nvltx` = bytes it sent (read requests).  Durations under ncu are\n"
            "serialised single launches; the un-profiled rates of the same probe are below.\n\n"
            "| # | kernel | device | duration ms | nvlrx GB | nvltx GB | nvlrx GB/s | dram read GB | dram write GB |\n|---|---|---|---:|---:|---:|---:|---:|---:|\n")
    for (i, name, dev), m in launches.items():
        d = m.get("gpu__time_duration.sum", 0.0)
        rx, tx = m.get("nvlrx__bytes.sum", 0.0), m.get("nvltx__bytes.sum", 0.0)
        f.write(f"| {i} | `{name}` | {dev} | {d * 1e3:.3f} | {rx / 1e9:.3f} | {tx / 1e9:.3f} | {(rx / d / 1e9) if d else 0:.1f} | "
                f"{m.get('dram__bytes_read.sum', 0) / 1e9:.3f} | {m.get('dram__bytes_write.sum', 0) / 1e9:.3f} |\n")
    if probe:
        f.write(f"\nSame probe without the profiler (CUDA events, per direction): `{json.dumps(probe)}`\n")
print(open(out).read())
