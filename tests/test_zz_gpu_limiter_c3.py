"""BASELINE config 3 in miniature, under the driver's `-m gpu` run: 4 vGPU worker processes @ upLimit 25 on one
B200, the parent playing the hypervisor's 2 Hz ERL loop (quota_controller.go:378-458) through the provider ABI.
(Named to sort last: these are the suite's only assertions on wall-clock behaviour of four competing processes, and a
`pytest -x` run should have checked every bit-exactness test before it gets here.)"""
import json
import os
import subprocess
import sys

import pytest

import conftest

pytestmark = pytest.mark.gpu
TOOL = os.path.join(conftest.ROOT, "tools", "limiter_c3.py")


def run(*extra, timeout=300, seconds="16"):
    r = subprocess.run([sys.executable, TOOL, "--seconds", seconds, *extra], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def _short(out):
    return {k: v for k, v in out.items() if k not in ("workers", "ticks_t_util_nsamples")}


def test_four_vgpus_at_25_percent_share_the_gpu_equally_and_smoothly():
    """The reference's semantics: every worker's target is compared with WHOLE-device utilisation (quota_controller.go:388-436),
    so four tenants at upLimit 25 settle at ~25 % of the GPU together."""
    out = run("--workers", "4", "--limit", "25", "--feedback", "device")
    shares = out["share_percent_each"]
    assert len(shares) == 4 and all(s > 1.0 for s in shares), _short(out)
    assert out["share_error_vs_equal_percent"] < 10.0, _short(out)       # the four tenants get the same share (measured: 0.4 %)
    # the provider reports utilisation averaged over the polling interval, so the loop settles ON the target (measured 23-26 %;
    # with the driver's single latest sample the controller saw 0 / 99 at random and overshot to 34 % of busy time)
    assert 15.0 < out["device_util_percent_mean_2nd_half"] < 35.0, _short(out)
    assert sum(shares) < 40.0, _short(out)
    assert out["gate_timeouts"] == 0, _short(out)                        # nobody fell through the fail-open timer
    # Tokens are handed over at the controller's rate in 50 ms bursts, not as one lump per 500 ms tick.  At a 5.8 % share a
    # 200 us launch comes round every 3.5 ms on average; once the controller has settled (second half of the run) no launch
    # waits for the rest of a tick (round 1, and round 2 before the pacing fix: p99 24-28 ms = a 450 ms stall per tick)
    # (measured: p50 3.0 ms, mean 3.5 ms; an occasional overshoot of the 2 Hz loop still drains a bucket for the rest of that
    # tick -- the controller's own throttle, quota_controller.go:349-376 -- so the tail is asserted as a fraction of batches)
    assert out["per_launch_ms_mean_max"] < 6.0, _short(out)
    assert out["steady_per_launch_ms_p50_max"] < 5.0, _short(out)
    assert out["steady_stalled_batches_percent_max"] < 5.0, _short(out)


def test_per_process_feedback_gives_each_vgpu_its_own_share_with_low_launch_latency():
    """feedback=process: each worker's own SM utilisation is its feedback signal -- every tenant is regulated towards ITS 25 %."""
    out = run("--workers", "4", "--limit", "25", "--feedback", "process")
    shares = out["share_percent_each"]
    assert all(s > 8.0 for s in shares) and out["share_error_vs_equal_percent"] < 30.0, _short(out)   # measured 15.7-16.1 % each (NVML's per-process samples are coarse)
    assert out["gate_timeouts"] == 0, _short(out)
    assert out["steady_per_launch_ms_p99_max"] < 5.0, _short(out)        # measured 1.3 ms (p50 0.9 ms)
    assert out["steady_stalled_batches_percent_max"] < 2.0, _short(out)


def test_without_the_limiter_the_four_tenants_take_the_whole_gpu():
    out = run("--workers", "4", "--limit", "25", "--no-limiter", seconds="6")
    assert sum(out["share_percent_each"]) > 80.0, out
