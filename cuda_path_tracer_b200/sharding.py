"""Multi-GPU plumbing.  torch.distributed is plumbing only — every rank renders
with its own context.
  * sample-range sharding of a progressive render: the only exchange is one `reduce` of the
    radiance/G-buffer sums, 32 B/pixel;
  * row-band sharding of ONE frame (1 spp + denoise): each rank renders its band, the bands'
    edge rows are exchanged once (halo for the a-trous denoiser), each rank denoises its band,
    and the finished RGBA8 rows are gathered."""
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


# ------------------------------------------------------------------ row bands of one frame
def band_rows(rank: int, world: int, height: int):
    """Rows [begin, end) of rank's band: whole 8x4 primary-ray tiles, sizes differ by <= 4 rows."""
    tiles = (height + 3) // 4
    base, rem = divmod(tiles, world)
    t0 = rank * base + min(rank, rem)
    t1 = t0 + base + (1 if rank < rem else 0)
    return min(height, 4 * t0), min(height, 4 * t1)


def _overlap(a0, a1, b0, b1):
    lo, hi = max(a0, b0), min(a1, b1)
    return (lo, hi) if lo < hi else None


def halo_plan(rank: int, world: int, height: int, halo: int):
    """(sends, recvs): lists of (peer, row_begin, row_end).  Rank q needs rows
    [begin_q - halo, begin_q) and [end_q, end_q + halo) (clipped to the frame) from whoever owns
    them; with thin bands a halo can span several neighbours."""
    bands = [band_rows(r, world, height) for r in range(world)]

    def needs(q):
        b0, b1 = bands[q]
        return [(max(0, b0 - halo), b0), (b1, min(height, b1 + halo))]

    sends, recvs = [], []
    for q in range(world):
        if q == rank:
            continue
        for n0, n1 in needs(q):                      # what q needs from me
            o = _overlap(n0, n1, *bands[rank])
            if o:
                sends.append((q, o[0], o[1]))
        for n0, n1 in needs(rank):                   # what I need from q
            o = _overlap(n0, n1, *bands[q])
            if o:
                recvs.append((q, o[0], o[1]))
    return sends, recvs


def exchange_halo(sums, width: int, height: int, halo: int):
    """Halo exchange for the band-sharded denoiser.  `sums` is the frame-sized accumulation buffer
    of this rank (flat float32: colour plane then G-buffer plane, 4 floats per pixel) in which
    only the rank's own band is valid; on return the `halo` rows beyond either end of the band
    hold their owners' values.  Point-to-point (NCCL send/recv over NVLink; gloo in the CPU
    tests): 2 planes x halo x width x 16 B per neighbour and direction."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sums
    rank, world = dist.get_rank(), dist.get_world_size()
    planes = sums.view(2, height, width * 4)
    sends, recvs = halo_plan(rank, world, height, halo)
    ops = []
    for peer, r0, r1 in sends:
        for pl in range(2):
            ops.append(dist.P2POp(dist.isend, planes[pl, r0:r1], peer))
    for peer, r0, r1 in recvs:
        for pl in range(2):
            ops.append(dist.P2POp(dist.irecv, planes[pl, r0:r1], peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return sums


def gather_rows(image, height: int, dst: int = 0):
    """Collect every rank's band of a frame-sized [H, ...] tensor on rank `dst` (in place)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return image
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == dst:
        for q in range(world):
            if q != dst:
                b0, b1 = band_rows(q, world, height)
                if b0 < b1:
                    ops.append(dist.P2POp(dist.irecv, image[b0:b1], q))
    else:
        b0, b1 = band_rows(rank, world, height)
        if b0 < b1:
            ops.append(dist.P2POp(dist.isend, image[b0:b1], dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return image
