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


def nccl_unique_id() -> bytes:
    """The 128-byte NCCL id made by the library (ptap_nccl_unique_id); rank 0 calls this and hands the bytes to every rank."""
    import ctypes as C
    from . import _native as N
    buf = (C.c_char * 128)()
    rc = N.lib().ptap_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise N.PtapError(f"ptap_nccl_unique_id failed with {rc}: libnccl.so.2 could not be loaded (set PTAP_NCCL_LIB)")
    return bytes(buf)


def nccl_join(renderer, rank: int, world: int):
    """Creates the library's own NCCL communicator on every rank (ptap_nccl_init), the id travelling over torch.distributed - the only
    thing torch does for the data plane.  Returns a description string, or None when the library could not load NCCL (the caller then
    falls back to torch.distributed.reduce on the film view; plumbing, not compute)."""
    import torch
    import torch.distributed as dist
    from . import _native as N
    ok = torch.ones(1, dtype=torch.int32, device=f"cuda:{torch.cuda.current_device()}")
    idt = torch.zeros(128, dtype=torch.uint8, device=ok.device)
    if rank == 0:
        try:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        except N.PtapError:
            ok.zero_()
    dist.broadcast(ok, 0)
    if int(ok.item()) == 0:
        return None
    dist.broadcast(idt, 0)
    try:
        renderer.nccl_init(bytes(idt.cpu().numpy().tobytes()), world, rank)
    except N.PtapError:
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        renderer.nccl_finalize()
        return None
    return "ptap_reduce (ncclReduce through the C ABI)"


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
