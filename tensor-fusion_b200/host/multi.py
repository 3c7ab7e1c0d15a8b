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
