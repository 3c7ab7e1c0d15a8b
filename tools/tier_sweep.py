"""The policy path of VRAM expansion as a client sees it: ONE vGPU whose address space is larger
than its GPU -- 1 TiB over 8 GPUs' HBM, or 256 GiB on one 180 GB GPU with the cold part in pinned host DRAM -- swept
sequentially through tfw_vspace_access (tfw_vspace_sweep: access + a kernel reading the whole region on the client
stream).  Every access of a cold region is one 1 GiB prefetch INTO the home GPU plus one 1 GiB eviction OUT of it,
asynchronous and overlapped (prefetch-ahead 2); every region's digest is checked on every lap, a sample of them
against the CPU oracle.  Prints one JSON object; bench.py embeds it as swap.c4_policy_sweep / swap.c5_policy_sweep."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
R = 1 << 30


def pattern_digests(seeds, nbytes):
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(16, len(seeds))) as ex:      # ctypes releases the GIL: regions in parallel
        return list(ex.map(lambda sd: oracle.digest(oracle.pattern(sd, nbytes)), seeds))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--va-gib", type=int, default=0)
    ap.add_argument("--home-device", type=int, default=0)
    ap.add_argument("--ahead", type=int, default=2)
    ap.add_argument("--laps", type=int, default=4)
    ap.add_argument("--engine", choices=["copy-engine", "kernel"], default="copy-engine",
                    help="who moves a region: the copy engines (default: they leave the SMs to the tenant) or the byte-mover kernel")
    ap.add_argument("--receiver-driven", action="store_true", help="copies run on the GPU that receives them (pulls) instead of the one that holds the source (pushes)")
    ap.add_argument("--early-remap", action="store_true", help="re-point a prefetched region's VA when its copy is issued instead of when it has completed")
    ap.add_argument("--kernel-ctas", type=int, default=2, help="kernel engine: CTAs per SM of a copy (0 = one tile per CTA, the whole GPU)")
    ap.add_argument("--home-driven", action="store_true", help="every copy driven by the home GPU (pull prefetch, push evict); backings mapped for the home GPU only")
    ap.add_argument("--fixed-frames", action="store_true", help="TFW_VS_FIXED_FRAMES: home backings mapped once at every VA that will use them, direct-mapped replacement, no VMM call per migration")
    ap.add_argument("--push-evict", action="store_true", help="evictions pushed by the home GPU, prefetches pulled by it (every copy kernel on the home GPU)")
    a = ap.parse_args()
    a.copy_engine, a.sender_driven = a.engine == "copy-engine", not a.receiver_driven
    if not a.copy_engine:
        os.environ.setdefault("TFW_VS_PEER_CTAS", str(a.kernel_ctas))
    from tensor_fusion_b200 import multi
    from tensor_fusion_b200 import vram as V
    plan = multi.vgpu_plan(a.gpus, a.va_gib)
    npeers, home_gib, peer_gib, host_gib, va = plan["n_peers"], plan["home"], plan["peer_each"], plan["host"], plan["va"]
    nreg = va
    tier = "peer" if npeers else "host"
    peers = [d for d in range(a.gpus) if d != a.home_device] if npeers else []
    with V.VSpace(home=a.home_device, va_bytes=nreg * R, region_bytes=R, home_budget=home_gib * R, peer_budget=peer_gib * R, host_budget=host_gib * R,
                  peers=peers, prefetch_ahead=a.ahead,
                  flags=(V.COPY_ENGINE if a.copy_engine else 0) | (V.SENDER_DRIVEN if a.sender_driven else 0) | (V.PUSH_EVICT if a.push_evict else 0) | (V.HOME_DRIVEN if a.home_driven else 0)
                  | (V.FIXED_FRAMES if a.fixed_frames else 0 if a.early_remap else V.REMAP_LATE)) as vs:
        t0 = time.perf_counter()
        want = []
        for r in range(nreg):
            vs.access(r)                      # first touch; colder regions are evicted as we go
            vs.fill_pattern(r, 77000 + r)
            want.append(vs.digest(r))         # known answer while the region has never moved (kernels pinned to the oracle by tests/)
        vs.quiesce()
        populate_s = time.perf_counter() - t0
        sample = sorted({0, 1, nreg // 2, nreg - 1})
        assert [want[r] for r in sample] == pattern_digests([77000 + r for r in sample], R), "pattern/digest kernels disagree with the CPU oracle"
        st0 = vs.stats()
        laps, lap_detail = [], []
        for _ in range(a.laps):
            sa = vs.stats()
            got, secs = vs.sweep(0, nreg)
            sb = vs.stats()
            bad = [r for r in range(nreg) if got[r] != want[r]]
            assert not bad, f"sweep: {len(bad)} regions changed their bytes, first {bad[:4]}"
            laps.append(secs)
            lap_detail.append({"s": round(secs, 3), "vmm_ms": round((sb["vmm_ns"] - sa["vmm_ns"]) / 1e6, 1), "stall_ms": round((sb["stall_ns"] - sa["stall_ns"]) / 1e6, 1),
                               "created": sb["phys_created"] - sa["phys_created"], "hits_inflight": sb["policy_hits_inflight"] - sa["policy_hits_inflight"]})
        st1 = vs.stats()
    secs = min(laps)
    med = sorted(laps)[len(laps) // 2]
    pf = (st1[f"prefetch_bytes_{tier}"] - st0[f"prefetch_bytes_{tier}"]) / len(laps)
    ev = (st1[f"evict_bytes_{tier}"] - st0[f"evict_bytes_{tier}"]) / len(laps)
    out = {"what": f"1 vGPU of {va} GiB on {a.gpus} GPU(s) ({home_gib} GiB home budget, " +
                   (f"{npeers} peers x {peer_gib} GiB over NVLink" if npeers else f"{host_gib} GiB pinned host DRAM over PCIe") +
                   f"), sequential sweep of all {nreg} x 1 GiB regions through tfw_vspace_access + a kernel reading each region; best of {len(laps)} laps",
           "va_gib": va, "regions": nreg, "prefetch_ahead": a.ahead, "engine": a.engine + ("" if a.copy_engine else f" ({os.environ.get('TFW_VS_PEER_CTAS')} CTAs/SM)"),
           "copies_driven_by": "home GPU" if a.home_driven else "sender (push)" if a.sender_driven else "receiver (pull)", "va_repointed": "never (fixed frames, direct-mapped)" if a.fixed_frames else "at issue" if a.early_remap else "at completion", "sweep_seconds": round(secs, 3),
           "lap_seconds": [round(x, 3) for x in laps], "laps": lap_detail, "populate_seconds": round(populate_s, 2),
           "prefetch_GBps_into_home_gpu": round(pf / secs / 1e9, 1), "evict_GBps_out_of_home_gpu": round(ev / secs / 1e9, 1),
           "both_directions_GBps": round((pf + ev) / secs / 1e9, 1), "regions_brought_home_per_lap": round(pf / R, 1),
           "median_lap_seconds": round(med, 3), "median_lap_GBps_per_direction": round(pf / med / 1e9, 1)}
    if npeers:
        out["prefetch_frac_of_nvlink_nominal_900"] = round(pf / secs / 1e9 / 900.0, 3)
        out["evict_frac_of_nvlink_nominal_900"] = round(ev / secs / 1e9 / 900.0, 3)
        out["median_lap_frac_of_nvlink_nominal_900"] = round(pf / med / 1e9 / 900.0, 3)
    out.update({"hits_inflight": st1["policy_hits_inflight"] - st0["policy_hits_inflight"],
                "host_stall_ms_per_lap": round((st1["stall_ns"] - st0["stall_ns"]) / 1e6 / len(laps), 1),
                "vmm_ms_per_lap": round((st1["vmm_ns"] - st0["vmm_ns"]) / 1e6 / len(laps), 1),
                "backings_created_in_sweeps": st1["phys_created"] - st0["phys_created"], "backings_destroyed_in_sweeps": st1["phys_destroyed"] - st0["phys_destroyed"],
                "verified": f"every region's digest after each lap; {len(sample)} regions cross-checked against the CPU oracle"})
    if a.fixed_frames:
        out["replacement"] = (f"direct-mapped: region r lives in frame r % {home_gib}; on a cyclic sweep LRU misses on every region, "
                              f"direct-mapped only on the regions that share a frame ({round(pf / R)} of {nreg} per lap)")
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
