"""N > 1 host logic on CPU: world_size-2 gloo run of the sample-range sharding + sums reduce.
The per-rank renderer is the CPU oracle here (tests only); on GPUs bench.py runs the same
helpers over NCCL with the CUDA path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, spp_total, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cuda_path_tracer_b200 as pt
    from cuda_path_tracer_b200 import sharding
    from tests.oracle_lib import load_oracle
    sd = pt.three_balls(40, 24)
    w, h = sd.resolution
    first, n = sharding.iteration_range(rank, world, spp_total)
    osc = load_oracle().scene(sd)
    sums = torch.zeros(2, w * h, 4, dtype=torch.float32)
    for it in range(first, first + n):
        # one iteration (seed = absolute iteration index) into fresh zero buffers; the oracle
        # folds it as the running mean (0*(it) + x)/(it+1) (final_gather), so x = (it+1) * mean
        color = np.zeros((h, w, 3), np.float32)
        normal = np.zeros((h, w, 3), np.float32)
        depth = np.zeros((h, w), np.float32)
        osc.render(sd.camera, w, h, 1, 6, first_iteration=it, state=(color, normal, depth))
        scale = float(it + 1)
        sums[0, :, :3] += torch.from_numpy(color.reshape(-1, 3)) * scale
        sums[1, :, :3] += torch.from_numpy(normal.reshape(-1, 3)) * scale
        sums[1, :, 3] += torch.from_numpy(depth.reshape(-1)) * scale
        sums[0, :, 3] += 1.0
    sharding.reduce_sums(sums, dst=0)
    if rank == 0:
        color, normal, depth = sharding.means_from_sums(sums, w * h)
        np.savez(out_path, color=color.numpy(), count=sums[0, :, 3].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_iteration_ranges_partition_exactly():
    from cuda_path_tracer_b200.sharding import iteration_range, weak_range
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 7, 64, 1024):
            seen = []
            for r in range(world):
                s, n = iteration_range(r, world, spp, first_iteration=5)
                seen += list(range(s, s + n))
            assert seen == list(range(5, 5 + spp))
    assert weak_range(3, 16, 2) == (50, 16)


def test_two_rank_gloo_sharding_equals_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    import cuda_path_tracer_b200 as pt
    from tests.oracle_lib import load_oracle
    spp = 6
    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, spp, out), nprocs=2, join=True)
    got = np.load(out)
    sd = pt.three_balls(40, 24)
    w, h = sd.resolution
    ref = load_oracle().scene(sd).render(sd.camera, w, h, spp, 6)[0].reshape(-1, 3)
    assert np.all(got["count"] == spp)
    assert np.abs(got["color"] - ref).max() < 1e-5


# ------------------------------------------------------------------ row bands of one frame
def test_band_rows_partition_the_frame_in_whole_tiles():
    from cuda_path_tracer_b200.sharding import band_rows, halo_plan
    for height in (4, 30, 54, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                b0, b1 = band_rows(r, world, height)
                assert b0 % 4 == 0 and (b1 % 4 == 0 or b1 == height)
                rows += list(range(b0, b1))
            assert rows == list(range(height))
    # every needed halo row is received exactly once, from its owner, and each send has a receiver
    for height, world, halo in ((1080, 8, 62), (64, 8, 62), (54, 2, 14), (1080, 2, 62)):
        plans = [halo_plan(r, world, height, halo) for r in range(world)]
        for r in range(world):
            b0, b1 = band_rows(r, world, height)
            need = set(range(max(0, b0 - halo), b0)) | set(range(b1, min(height, b1 + halo)))
            got = []
            for peer, r0, r1 in plans[r][1]:
                p0, p1 = band_rows(peer, world, height)
                assert p0 <= r0 and r1 <= p1
                got += list(range(r0, r1))
                assert (r, r0, r1) in plans[peer][0]
            assert sorted(got) == sorted(need)


def _band_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuda_path_tracer_b200 import sharding
    W, H, halo = 12, 54, 14
    truth = torch.arange(2 * H * W * 4, dtype=torch.float32)
    sums = torch.full_like(truth, -1.0)
    b0, b1 = sharding.band_rows(rank, world, H)
    sums.view(2, H, W * 4)[:, b0:b1] = truth.view(2, H, W * 4)[:, b0:b1]       # own band only
    sharding.exchange_halo(sums, W, H, halo)
    lo, hi = max(0, b0 - halo), min(H, b1 + halo)
    ok = bool(torch.equal(sums.view(2, H, W * 4)[:, lo:hi], truth.view(2, H, W * 4)[:, lo:hi]))
    untouched = sums.view(2, H, W * 4)
    ok = ok and bool((untouched[:, :lo] == -1).all()) and bool((untouched[:, hi:] == -1).all())
    img = torch.zeros(H, W, 4, dtype=torch.uint8)
    img[b0:b1] = rank + 1
    sharding.gather_rows(img, H, dst=0)
    if rank == 0:
        owner = torch.zeros(H, dtype=torch.uint8)
        for q in range(world):
            q0, q1 = sharding.band_rows(q, world, H)
            owner[q0:q1] = q + 1
        ok = ok and bool(torch.equal(img[:, 0, 0], owner))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(out_path, flag.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_halo_exchange_and_row_gather(tmp_path, world):
    out = str(tmp_path / "ok.npy")
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_band_worker, args=(world, port, out), nprocs=world, join=True)
    assert int(np.load(out)[0]) == 1
