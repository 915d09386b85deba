"""Sample partitioning across the GPUs of one box: one process per GPU, the scene replicated, every rank renders a
contiguous range of iterations (the reference seeds its RNG with the iteration number, Renderer.cpp:435, so the union over ranks
is exactly the sample set of a one-GPU run), and the per-rank films are combined with ONE reduce (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  There is no other data-path collective."""
from __future__ import annotations

import numpy as np


def iteration_range(rank: int, world: int, iters: int) -> tuple[int, int]:
    """[begin, end) of rank `rank` when `iters` iterations are split over `world` ranks (remainder to the first ranks)."""
    if not (0 <= rank < world) or iters < 0:
        raise ValueError("iteration_range: need 0 <= rank < world and iters >= 0")
    base, rem = divmod(iters, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def film_tensor(renderer):
    """Zero-copy torch view of the renderer's device film (W*H*3 floats, the un-normalised sum) for the NCCL reduce."""
    import torch
    ptr, n = renderer.film_device_ptr()

    class _Film:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Film(), device=f"cuda:{torch.cuda.current_device()}")


def reduce_film(film, dst: int = 0):
    """Sum the per-rank films onto rank `dst` with one collective.  `film` is a torch tensor (device film view, or a host tensor
    under gloo) and is reduced in place on `dst`."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def render_partitioned(renderer, iters: int, rank: int, world: int):
    """renderLoop over this rank's share of `iters` iterations, then the reduce; returns the film tensor (complete on rank 0).
    The collective is issued under the renderer's own stream, so it is ordered after the render without a host synchronisation."""
    import torch
    b, e = iteration_range(rank, world, iters)
    renderer.frame_begin()
    if e > b:
        renderer.render(b, e)
    t = film_tensor(renderer)
    if world > 1:
        with torch.cuda.stream(torch.cuda.ExternalStream(renderer.stream_ptr())):
            reduce_film(t, 0)
    # after the reduce rank 0's film holds all `iters` samples: renderImage divides by this count (Renderer.cpp:42)
    renderer._iters_done = iters if rank == 0 else e - b
    return t
