"""Multi-GPU plumbing: sample-range sharding of a progressive render and the single reduce
of the radiance/G-buffer sums.  torch.distributed is plumbing only — every
rank renders with its own context; the only exchange is one `reduce` of 32 B/pixel."""
from __future__ import annotations


def iteration_range(rank: int, world: int, spp_total: int, first_iteration: int = 0):
    """Strong-scaling split of iterations [first, first+spp_total) into `world` contiguous
    ranges that differ in length by at most one; the seeds are exactly the ones a single GPU
    would use (hash(hash(pixel) ^ iteration), ray_gen.cu:18), so the union is the same image."""
    base, rem = divmod(spp_total, world)
    start = first_iteration + rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def weak_range(rank: int, spp_per_rank: int, first_iteration: int = 0):
    """Weak-scaling split: every rank renders spp_per_rank iterations of the same frame."""
    return first_iteration + rank * spp_per_rank, spp_per_rank


def reduce_sums(sums, dst: int = 0):
    """In-place sum of the per-rank accumulation buffers onto rank `dst` (NCCL over NVLink on
    GPUs, gloo in the CPU tests).  `sums` holds, per pixel, colour.rgb + sample count and
    normal.xyz + depth — all plain sums, hence associative."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sums, dst=dst)
    return sums


def means_from_sums(sums, pixels: int):
    """(colour mean [P,3], normal mean [P,3], depth mean [P]) from a reduced sums buffer."""
    s = sums.view(2, pixels, 4)
    n = s[0, :, 3:4]
    return s[0, :, :3] / n, s[1, :, :3] / n, s[1, :, 3] / n[:, 0]
