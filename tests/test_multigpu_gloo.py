"""N>1 host logic on CPU: world_size-2 (and 3) gloo runs of the spp-shard + single-reduce plan that
bench.py and the NCCL path use.  Each rank "renders" its shard with a deterministic per-(pixel,
sample) function standing in for K1, so the reduced buffer has a closed form."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uecraytracing_b200.distributed import item_range, item_shard, reduce_accum, sample_shard


def test_sample_shards_tile_the_range():
    for spp in (1, 2, 7, 100, 500, 1024, 4096):
        for world in (1, 2, 3, 4, 8):
            shards = [sample_shard(spp, r, world) for r in range(world)]
            ids = [i for lo, n in shards for i in range(lo, lo + n)]
            assert ids == list(range(spp))
            assert max(n for _, n in shards) - min(n for _, n in shards) <= 1
    assert sample_shard(500, 3, 8) == (187, 63)
    with pytest.raises(ValueError):
        sample_shard(10, 2, 2)


def test_item_shards_tile_the_frame():
    """Work-item shards (uecrt_params.shard_index / shard_count): equal to within one item, contiguous, complete."""
    for total in (0, 1, 7, 810000 * 500, 25920 * 32 * 500, 2 ** 40 + 3):
        for world in (1, 2, 3, 4, 8):
            r = [item_range(total, i, world) for i in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert item_shard(0, 1) == (0, 0) and item_shard(3, 8) == (3, 8)
    with pytest.raises(ValueError):
        item_shard(2, 2)


def _fake_path_radiance(pixel, sample):
    # exactly representable, depends on the GLOBAL sample id like the Philox counter does
    return ((pixel * 31 + sample * 7) % 64) / 64.0


def _worker(rank, world, port, spp, npix, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, n = sample_shard(spp, rank, world)
    acc = torch.zeros(npix, 4)
    for p in range(npix):
        for s in range(lo, lo + n):
            r = _fake_path_radiance(p, s)
            acc[p] += torch.tensor([r, 2 * r, 3 * r, 1.0])
    reduce_accum(acc, dst=0)
    dist.barrier()
    if rank == 0:
        q.put(acc.numpy())
    dist.destroy_process_group()


def _worker_items(rank, world, port, spp, npix, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = item_range(npix * spp, *item_shard(rank, world))
    acc = torch.zeros(npix, 4)
    for w in range(lo, hi):                       # item w = sample * npix + pixel (sample-major, kernels.cu slot_assign)
        s, p = divmod(w, npix)
        r = _fake_path_radiance(p, s)
        acc[p] += torch.tensor([r, 2 * r, 3 * r, 1.0])
    reduce_accum(acc, dst=0)
    dist.barrier()
    if rank == 0:
        q.put(acc.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_item_sharded_reduce_gloo(world):
    """The plan bench.py and uecrt_render(num_gpus=N) run: every rank takes an equal contiguous range of the frame's
    (pixel, sample) items, one reduce sums the buffers; every pixel ends up with every sample exactly once."""
    spp, npix = 13, 24
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_items, args=(r, world, port, spp, npix, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros((npix, 4), np.float32)
    for p in range(npix):
        for s in range(spp):
            r = _fake_path_radiance(p, s)
            want[p] += [r, 2 * r, 3 * r, 1.0]
    assert np.array_equal(got, want)
    assert np.all(got[:, 3] == spp)


@pytest.mark.parametrize("world", [2, 3])
def test_spp_sharded_reduce_gloo(world):
    spp, npix = 13, 24
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, spp, npix, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros((npix, 4), np.float32)
    for p in range(npix):
        for s in range(spp):
            r = _fake_path_radiance(p, s)
            want[p] += [r, 2 * r, 3 * r, 1.0]
    assert np.array_equal(got, want)          # the sums are exact in FP32
    assert np.all(got[:, 3] == spp)           # every pixel saw every sample id exactly once
