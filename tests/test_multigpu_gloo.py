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

from uecraytracing_b200.distributed import reduce_accum, sample_shard


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
