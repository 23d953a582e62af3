"""One-process-per-GPU plumbing for the spp-sharded render.

Every (pixel, sample) path is independent under the counter-based RNG, so rank g of G renders the
GLOBAL sample ids [g*S/G, (g+1)*S/G) of every pixel into its own float4 accumulation buffer and the
buffers are summed by ONE reduce to rank 0 (NCCL over NVLink on GPUs; gloo in the CPU tests).  The
Philox counters use the global sample id, so the image is independent of G up to FP32 summation
order.  There is no other exchange step on this path.
"""


def sample_shard(spp, rank, world):
    """(sample_begin, sample_count) of `rank`; shards tile [0, spp) exactly, sizes differ by <= 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = spp * rank // world
    hi = spp * (rank + 1) // world
    return lo, hi - lo


def item_shard(rank, world):
    """(shard_index, shard_count) for uecrt_params: rank's equal range of the frame's (pixel, sample) work items.
    Unlike sample_shard every rank gets the same number of paths (to within one), whatever spp is."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank, world) if world > 1 else (0, 0)


def item_range(total_items, index, count):
    """[lo, hi) of shard `index` of `count` over `total_items` work items -- the rule uecrt_api.cu applies."""
    q, r = divmod(total_items, count)
    return q * index + r * index // count, q * (index + 1) + r * (index + 1) // count


def reduce_accum(accum, dst=0):
    """Sum the per-rank accumulation tensors onto `dst` (in place there). No-op for world size 1."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
