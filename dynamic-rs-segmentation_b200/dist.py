"""One-process-per-GPU plumbing (torch.distributed is used for rendezvous and the NCCL communicator only).

The reference is single-process; the two shardings below are what its hot path offers:

* training shards by batch: every rank runs the same host policy (same seeds -> same patch size, same batch
  selection) on its slice of the global batch; ``libdrs`` calls back ONE sum-allreduce per step over the flat
  buffer [gradients ++ loss numerator ++ confusion counts] (plus the per-layer BN sums when ``sync_bn``),
  and every rank applies the identical momentum update.
* full-scene inference shards by output row stripe: a rank evaluates every patch that intersects its stripe, so
  each pixel's contributions are all local and are added in the reference's visiting order (bit-exact, no halo
  exchange); the uint8 label stripes are then gathered to rank 0.
"""
import numpy as np


class _DevBuf:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def stripe_bounds(H, world, rank=None):
    """Row stripes [r*H/world, (r+1)*H/world)."""
    cuts = [(r * H) // world for r in range(world + 1)]
    if rank is None:
        return cuts
    return cuts[rank], cuts[rank + 1]


def stripe_rows_needed(H, crop, row_begin, row_end):
    """Scene rows a rank must hold to evaluate every patch that intersects its output stripe [row_begin, row_end):
    patch origins lie on the lattice {i*stride} U {H-crop} (isprs:1243, 366-375)."""
    stride = crop // 2
    origins = sorted(set(list(range(0, max(H - crop, 0) + 1, stride)) + [H - crop]))
    hit = [o for o in origins if o < row_end and o + crop > row_begin]
    return (min(hit), max(hit) + crop) if hit else (row_begin, row_end)


def attach_allreduce(session, group=None, sync_bn=False):
    """Route the library's exchange step through torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    cache = {}

    streams = {}

    def allreduce(ptr, count, stream):
        key = (ptr, count)
        t = cache.get(key)
        if t is None:
            t = torch.as_tensor(_DevBuf(ptr, count), device="cuda")
            cache[key] = t
        # NCCL orders itself after torch's CURRENT stream: make that the stream libdrs enqueues on (0 = legacy default)
        if stream:
            ext = streams.get(stream)
            if ext is None:
                ext = streams[stream] = torch.cuda.ExternalStream(stream)
            with torch.cuda.stream(ext):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            with torch.cuda.stream(torch.cuda.default_stream()):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    session.set_allreduce(allreduce if world > 1 else None, world, sync_bn)
    return world


def comm_unique_id():
    import ctypes as C
    from . import lib as L
    buf = (C.c_uint8 * 128)()
    L.check(L.load().drs_comm_unique_id(buf))
    return bytes(buf)


def attach_nccl(session, sync_bn=False, group=None):
    """The library's own NCCL communicator (csrc/drs_comm.cuh): rank 0 creates the id, torch.distributed only carries the 128
    bytes to the other ranks.  Every exchange of the step is then an ncclAllReduce enqueued by the library itself."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world <= 1:
        return world
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0, group=group)
    session.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world, sync_bn)
    return world


def _dist_device(group=None):
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def sync_host_rng(group=None):
    """Every data-parallel rank must run the SAME host policy (patch size, batch selection), i.e. the same two random
    streams.  The reference never seeds and cache files created by rank 0 only make the ranks consume different
    amounts of the streams: if the ranks' ``random`` / ``np.random`` states differ, rank 0 draws one seed from its own stream
    and every rank re-seeds both with it; if they already agree (a harness seeded every rank alike) nothing is touched.
    Returns True when a re-seed happened."""
    import random
    import zlib
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) <= 1:
        return False
    dev = _dist_device(group)
    st = np.random.get_state()
    h = zlib.crc32(st[1].tobytes() + repr(st[2:]).encode() + repr(random.getstate()).encode())
    v = torch.tensor([h, -h], dtype=torch.int64, device=dev)
    dist.all_reduce(v, op=dist.ReduceOp.MAX, group=group)
    if int(v[0]) == -int(v[1]):
        return False
    t = torch.tensor([np.random.randint(0, 2 ** 31 - 1) if dist.get_rank(group) == 0 else 0], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=0, group=group)
    np.random.seed(int(t.item()))
    random.seed(int(t.item()))
    return True


def rank0_first(fn, group=None):
    """Cache files in the working directory (isprs:2087-2115, 1634-1639): rank 0 runs ``fn`` (which may create and save), the
    other ranks wait at a barrier and run it afterwards (finding the files and loading them)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) <= 1:
        return fn()
    if dist.get_rank(group) == 0:
        r = fn()
        dist.barrier(group=group)
        return r
    dist.barrier(group=group)
    return fn()


def check_same_plan(crop, digest, group=None):
    """DRS_DP_DEBUG: raise unless patch size and batch digest are identical on every rank."""
    import torch
    import torch.distributed as dist
    v = torch.tensor([crop, digest, -crop, -digest], dtype=torch.int64, device=_dist_device(group))
    dist.all_reduce(v, op=dist.ReduceOp.MAX, group=group)
    if int(v[0]) != -int(v[2]) or int(v[1]) != -int(v[3]):
        raise RuntimeError("data-parallel ranks disagree on the step plan (patch size / batch): host RNG streams diverged")


def rank_slice(batch, rank, world):
    """This rank's share of a global batch (contiguous, equal sizes)."""
    per = len(batch) // world
    return batch[rank * per:(rank + 1) * per]


def gather_label_stripes(stripe, H, W, rank, world, device=None, group=None, all_ranks=False):
    """Collect every rank's uint8 label stripe on rank 0 -> [H, W] (None elsewhere unless all_ranks)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return stripe
    cuts = stripe_bounds(H, world)
    rows = max(b - a for a, b in zip(cuts[:-1], cuts[1:]))
    backend = dist.get_backend(group)
    dev = device if device is not None else ("cuda" if backend == "nccl" else "cpu")
    mine = torch.zeros((rows, W), dtype=torch.uint8, device=dev)
    mine[:stripe.shape[0]].copy_(torch.from_numpy(np.ascontiguousarray(stripe)))
    if backend == "nccl":
        out = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(out, mine, group=group)
        if rank != 0 and not all_ranks:
            return None
    elif all_ranks:
        out = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(out, mine, group=group)
    else:
        out = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, out, dst=0, group=group)
        if rank != 0:
            return None
    full = np.empty((H, W), dtype=np.uint8)
    for r in range(world):
        full[cuts[r]:cuts[r + 1]] = out[r][:cuts[r + 1] - cuts[r]].cpu().numpy()
    return full
