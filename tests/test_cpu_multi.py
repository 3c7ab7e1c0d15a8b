"""N > 1 host logic on CPU: world_size 2 over gloo (rendezvous on 127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conftest  # noqa: F401  (puts the repo root on sys.path)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensor_fusion_b200 import multi
    local_seconds = 1.0 + rank            # rank 1 is the slow one
    rate, slowest = multi.whole_job_rate(10.0, local_seconds)
    mx = multi.max_over_ranks([rank * 2.0, 5.0 - rank])
    peers = multi.peers_of(rank, world)
    slots = multi.stripe_slots(8, len(peers), rank)
    dist.barrier()
    out.put((rank, rate, slowest, mx, peers, slots))
    dist.destroy_process_group()


def test_two_rank_aggregation_and_striping():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rate, slowest, mx, peers, slots in res:
        assert slowest == 2.0 and rate == 2 * 10.0 / 2.0          # whole job / slowest rank
        assert mx == [2.0, 5.0]
        assert peers == [1 - rank] and slots == [0] * 8


def test_striping_is_balanced_for_every_world_size():
    from tensor_fusion_b200 import multi
    for world in (2, 4, 8):
        for n in (7, 8, 56):
            got = multi.incoming_regions(world, n)
            assert sum(got) == world * n
            if n % (world - 1) == 0:
                assert len(set(got)) == 1, (world, n, got)        # perfectly even
            else:
                assert max(got) - min(got) <= world               # never worse than one region per rank
    assert multi.stripe_slots(4, 0, 3) == [-1] * 4                 # single GPU: host tier only
    assert multi.max_over_ranks([1.5]) == [1.5]                     # not distributed: identity


def test_vgpu_plan_spans_the_box():
    """bench.py's C4 / C5 sizing: 256 GiB on one GPU with a host tier, 1 TiB over 8 GPUs' HBM, scaled with fewer."""
    from tensor_fusion_b200 import multi
    one = multi.vgpu_plan(1)
    assert one == {"va": 256, "home": 160, "peer_each": 0, "host": 104, "n_peers": 0}
    eight = multi.vgpu_plan(8)
    assert eight["va"] == 1024 and eight["home"] == 150 and eight["peer_each"] == 128 and eight["n_peers"] == 7 and eight["host"] == 0
    for n in (2, 4):
        p = multi.vgpu_plan(n)
        assert p["va"] == p["home"] - 8 + (n - 1) * p["peer_each"] and p["va"] > p["home"]      # larger than one GPU, fits the tiers
        assert p["home"] + 12 <= 180 and p["peer_each"] + 30 <= 180
    assert multi.vgpu_plan(8, va_gib=64)["va"] == 64
    assert multi.vgpu_plan(2, va_gib=5000)["va"] == 150 - 8 + 150
