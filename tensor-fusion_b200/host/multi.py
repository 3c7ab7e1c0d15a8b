"""Multi-GPU plumbing shared by bench.py and the tests (one process per GPU, torch.distributed).

The payload-stage and limiter paths do not shard ("replicas only"): every rank is
an independent vGPU worker.  The VRAM-expansion path shards by region: rank r is the home of its
own vGPU and stripes cold regions over all other GPUs with one-sided P2P -- no collective is on the
data path; torch.distributed only carries the barrier and the max-over-ranks of the timings.
"""


def peers_of(rank, world):
    """CUDA ordinals that can hold rank's cold regions: every other GPU of the box."""
    return [d for d in range(world) if d != rank]


def stripe_slots(n_regions, n_peers, rank):
    """Peer slot (index into peers_of) for each region: round-robin, rotated by rank so that when
    every rank evicts at once each GPU receives the same number of regions."""
    if n_peers <= 0:
        return [-1] * n_regions
    return [(r + rank) % n_peers for r in range(n_regions)]


def incoming_regions(world, n_regions):
    """How many regions land on each GPU when all ranks evict n_regions with stripe_slots."""
    got = [0] * world
    for rank in range(world):
        peers = peers_of(rank, world)
        for s in stripe_slots(n_regions, len(peers), rank):
            got[peers[s]] += 1
    return got


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def whole_job_rate(units_per_rank, seconds_local, device=None):
    """Aggregate throughput of `world` replicas: all units / the slowest rank's time."""
    import torch.distributed as dist
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    (slowest,) = max_over_ranks([seconds_local], device)
    return world * units_per_rank / slowest, slowest


def vgpu_plan(n_gpus, va_gib=0, hbm_gib=180):
    """Sizing of the one-vGPU-spans-the-box legs of bench.py, in GiB.

    One vGPU homed on GPU 0.  With peers (C5): 150 GiB resident at home, the cold part striped over the other GPUs'
    HBM -- 1 TiB on 8 GPUs (7 x 128 GiB), less with fewer peers (the other ranks of the bench keep a few GiB of their
    own on those GPUs).  Alone (C4): 256 GiB on one 180 GB GPU, 160 GiB resident, the rest in pinned host DRAM.
    Returns dict(va, home, peer_each, host, n_peers); va is capped by what the tiers can hold."""
    n_peers = max(0, n_gpus - 1)
    home = min(150 if n_peers else 160, hbm_gib - 12)
    peer_each = 0 if not n_peers else min(128 if n_gpus >= 8 else 150, hbm_gib - 30)
    host = 0 if n_peers else 104
    want = va_gib or (min(1024, home + n_peers * peer_each) if n_peers else 256)
    room = home - 8 + n_peers * peer_each + host            # prefetch slack and regions in transit stay free
    return {"va": max(1, min(want, room)), "home": home, "peer_each": peer_each, "host": host, "n_peers": n_peers}
