"""Sample-split data parallelism of the progressive renderer.

A pixel-sample depends only on (pixel, frame, read-only scene) and `frame` is the RNG seed (globals.glsl:116-120), so the
frames of a job are dealt round-robin to the ranks: rank g of G renders frames first+g, first+g+G, ...  Every rank holds
the whole scene and its own W*H*3 accumulator; the only exchange is one sum of the accumulators at the end (NCCL on the
GPU path via lfcuda_reduce / torch.distributed; gloo in the CPU tests).  No per-step collective exists on the data path.
"""


def rank_frames(first_frame, nframes, rank, world):
    """(first, count, stride) of the frames rank `rank` renders out of [first_frame, first_frame + nframes)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    count = (nframes - rank + world - 1) // world if nframes > rank else 0
    return first_frame + rank, count, world


def all_frames(first_frame, nframes, world):
    """The per-rank frame lists (for checks): disjoint, covering, each an arithmetic progression."""
    out = []
    for g in range(world):
        f0, n, st = rank_frames(first_frame, nframes, g, world)
        out.append([f0 + k * st for k in range(n)])
    return out


def reduce_accumulators(accum, dist=None):
    """Sum the per-rank accumulation buffers in place.  `accum` is a torch tensor (CPU with gloo, CUDA with nccl)."""
    if dist is None:
        import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM)
    return accum
