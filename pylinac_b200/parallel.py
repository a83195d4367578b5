"""Multi-GPU plumbing of the batched entry points.

Frames are independent, so a batch shards by frame index with no data-path collective; the only exchange is the
final gather of the fixed-size per-frame result rows.  One process per GPU (torchrun: RANK / LOCAL_RANK / WORLD_SIZE).
``torch.distributed`` is plumbing only (rendezvous, barrier, the gloo fallback of the gather used by the CPU tests);
on GPUs the gather is ONE ``ncclAllGather`` issued by libepid (``epid_gather_results``).
"""
from __future__ import annotations

import os

import numpy as np


def world_info():
    """(world_size, rank, local_rank) from the torchrun environment (1, 0, 0 when launched plainly)."""
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device: int, *, sysfs: str = "/sys") -> dict:
    """Pin this process to the CPU cores of the NUMA node CUDA device `device` hangs off, so that the page-locked frame /
    result buffers it allocates afterwards are node-local (first touch) and the H2D DMA does not cross the socket
    interconnect.  With one process per GPU and ~54 GB/s of H2D per GPU this is what keeps 8 ranks from sharing one
    socket's memory controllers.  Returns {"numa_node", "cpus"}; a no-op (numa_node None) when sysfs has no answer."""
    import ctypes as C

    from . import _native as nat

    info = {"numa_node": None, "cpus": None}
    try:
        buf = C.create_string_buffer(32)
        nat.check(nat.lib().epid_device_pci_bus_id(int(device), buf, 32))
        bus = buf.value.decode().lower()
        with open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")) as f:
            cpus = _parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return info


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of frame indices owned by `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_total: int, world: int) -> list[int]:
    return [shard_range(n_total, world, r)[1] - shard_range(n_total, world, r)[0] for r in range(world)]


def gather_rows(rows: np.ndarray, n_total: int, *, dist=None, ctx=None) -> np.ndarray:
    """All ranks contribute their shard's result rows (a numpy structured array); every rank receives all n_total rows in
    frame order.  Shards are padded to the largest shard so that the exchange is one fixed-size all-gather.

    ctx  : a pylinac_b200._native.Context whose NCCL communicator is initialised -> ncclAllGather (the GPU path)
    dist : an initialised torch.distributed module (gloo) -> all_gather of byte tensors (CPU tests / no NCCL)
    """
    world, rank, _ = world_info() if dist is None else (dist.get_world_size(), dist.get_rank(), 0)
    if world == 1:
        return rows
    if ctx is None and dist is None:
        raise ValueError(f"world size is {world}: pass ctx= (NCCL communicator initialised with init_comm) or dist= (torch.distributed)")
    sizes = shard_sizes(n_total, world)
    if len(rows) != sizes[rank]:
        raise ValueError(f"rank {rank} holds {len(rows)} rows, its shard has {sizes[rank]}")
    width = max(sizes)
    padded = np.zeros(width, rows.dtype)
    padded[: len(rows)] = rows
    if ctx is not None:
        from . import _native as nat

        nranks, crank = ctx.comm_info()
        if (nranks, crank) != (world, rank):
            raise RuntimeError(f"the context's communicator has {nranks} rank(s) (this one is {crank}) but the job has {world} "
                               f"(this one is {rank}): call parallel.init_comm(ctx, dist) on every rank first")
        allbuf = np.zeros(world * width, rows.dtype)
        nat.check(nat.lib().epid_gather_results(ctx.handle, padded.ctypes.data, padded.nbytes, allbuf.ctypes.data))
    else:
        import torch

        mine = torch.from_numpy(padded.view(np.uint8).copy())
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        allbuf = np.concatenate([p.numpy() for p in parts]).view(rows.dtype)
    out = np.empty(n_total, rows.dtype)
    pos = 0
    for r in range(world):
        out[pos: pos + sizes[r]] = allbuf[r * width: r * width + sizes[r]]
        pos += sizes[r]
    return out


def init_comm(ctx, dist) -> None:
    """Create the NCCL communicator of `ctx` across the ranks of an initialised torch.distributed group: rank 0 makes the
    ncclUniqueId, the 128 bytes travel through `dist.broadcast` (plumbing), every rank calls epid_comm_init."""
    import torch

    from . import _native as nat

    world, rank = dist.get_world_size(), dist.get_rank()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = np.zeros(128, np.uint8)
        nat.check(nat.lib().epid_comm_unique_id(buf.ctypes.data))
        uid = torch.from_numpy(buf)
    dist.broadcast(uid, src=0)
    nat.check(nat.lib().epid_comm_init(ctx.handle, world, rank, uid.numpy().ctypes.data))


def analyze_sharded(analyze_rows, frames, *, dist=None, ctx=None) -> np.ndarray:
    """Run a batched analysis on this rank's contiguous shard of ``frames`` and return ALL ranks' result rows in frame order.

    analyze_rows : callable(frames_shard) -> numpy structured array, one row per frame (e.g.
                   ``lambda f: winston_lutz.analyze_batch(f, dpmm).rows``); the frames of a batch are independent, so no
                   data-path collective is involved -- the only exchange is the final gather of the fixed-size rows.
    frames       : anything sliceable by frame index whose length is the total frame count (an ndarray, a memmap, a lazy loader);
                   only ``frames[start:end]`` of this rank's shard is touched.
    """
    world, rank, _ = world_info() if dist is None else (dist.get_world_size(), dist.get_rank(), 0)
    n_total = len(frames)
    start, end = shard_range(n_total, world, rank)
    rows = analyze_rows(frames[start:end])
    if len(rows) != end - start:
        raise ValueError(f"analyze_rows returned {len(rows)} rows for a shard of {end - start} frames")
    return gather_rows(rows, n_total, dist=dist, ctx=ctx)
