#!/usr/bin/env python
"""Summarise gpurun_out/<tag>_launches.csv and <tag>_mover.ncu-rep into profiles/ (tracked)."""
import collections
import csv
import io
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------------
src = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        k = row["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v_us; a[2] = min(a[2], v_us); a[3] = max(a[3], v_us)
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(os.path.join(P, f"{tag}_launches_summary.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n"
                "Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes.\n\n"
                "| kernel | launches | total us | avg us | min us | max us | share |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1]/a[0]:.2f} | {a[2]:.2f} | {a[3]:.2f} | {a[1]/tot*100:.1f} % |\n")
    shutil.copy(src, os.path.join(P, f"{tag}_launches.csv"))
    print(open(os.path.join(P, f"{tag}_launches_summary.md")).read())

# ---- full capture of the mover ---------------------------------------------------------------------
rep = os.path.join(G, f"{tag}_mover.ncu-rep")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    traffic_rows = []
    with open(os.path.join(P, f"{tag}_mover_ncu.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full of the byte mover (one row per captured launch)\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            f.write(f"## `{name}`  grid {r[hdr.index('launch__grid_size')] if 'launch__grid_size' in hdr else '?'}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| {w} | {r[i]} | {units[i]} |\n")
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")); ru = units[hdr.index("dram__bytes_read.sum")]
                wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", "")); wu = units[hdr.index("dram__bytes_write.sum")]
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                t = rd * scale[ru] + wr * scale[wu]
                f.write(f"| **traffic = dram read + write** | {t:.0f} | byte |\n")
                grid = int(r[hdr.index("launch__grid_size")].replace(",", ""))
                traffic_rows.append({"kernel": name, "grid": grid, "dram_bytes": int(t), "algorithmic_bytes_if_copy": grid * 32768 * 2,
                                     "duration_ms": float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))})
            except Exception:
                pass
            f.write("\n")
    print(open(os.path.join(P, f"{tag}_mover_ncu.md")).read()[:1500])
    if traffic_rows:
        import json
        best = max(traffic_rows, key=lambda x: x["grid"])
        best["source"] = f"profiles/{tag}_mover_ncu.md (ncu --set full --clock-control none, one launch)"
        json.dump(best, open(os.path.join(P, f"{tag}_mover_traffic.json"), "w"), indent=1)
for extra in (f"{tag}_bench.json", "sweep.jsonl", "copy_lab.jsonl"):
    s = os.path.join(G, extra)
    if os.path.exists(s):
        dst = extra if extra.startswith(tag) else f"{tag}_" + extra.replace("sweep", "mover_sweep")
        shutil.copy(s, os.path.join(P, dst))
