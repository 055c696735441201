"""Multi-GPU layout of a stream batch: contiguous blocks of streams per rank, no data-path collective.

Streams never interact and their state stays resident on the GPU that owns them, so the only traffic
between ranks is I/O when a deployment has a single ingest point: rank `src` scatters int16[N][640] PCM rows and gathers
uint8[N][cap] payloads + int16[N][2] length fields (decoder: the mirror image).  Both are grouped point-to-point
transfers (ncclSend/ncclRecv under the "nccl" backend, i.e. NVLink 5 / NVSwitch peer copies; "gloo" on CPU for tests).
Ranks that ingest their own streams (the layout bench.py times) never call into this module."""
import torch
import torch.distributed as dist


def _bytes(t):
    """Raw byte view of a contiguous tensor: NCCL has no int16 type, and these transfers are plain byte moves anyway."""
    return t.view(torch.uint8)


def shard_bounds(n_streams, world):
    """[(lo, hi)) per rank: rank r owns streams r*N//W .. (r+1)*N//W (contiguous so every transfer is one message)."""
    return [(r * n_streams // world, (r + 1) * n_streams // world) for r in range(world)]


def local_count(n_streams, rank, world):
    lo, hi = shard_bounds(n_streams, world)[rank]
    return hi - lo


def scatter_streams(full, n_streams, row_shape, dtype, device, src=0, group=None):
    """Rank `src` passes `full` = tensor [N, *row_shape]; every rank (src included) gets its own rows [n_local, *row_shape]."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = shard_bounds(n_streams, world)
    lo, hi = bounds[rank]
    out = torch.empty((hi - lo,) + tuple(row_shape), dtype=dtype, device=device)
    ops = []
    if rank == src:
        assert full.shape[0] == n_streams and full.dtype == dtype
        for r, (a, b) in enumerate(bounds):
            if r == src:
                out.copy_(full[a:b])
            elif b > a:
                ops.append(dist.P2POp(dist.isend, _bytes(full[a:b].contiguous()), r, group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.irecv, _bytes(out), src, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def gather_streams(local, n_streams, dst=0, group=None):
    """Inverse of scatter_streams: rank `dst` returns [N, *row_shape], the others None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = shard_bounds(n_streams, world)
    lo, hi = bounds[rank]
    assert local.shape[0] == hi - lo
    ops = []
    full = None
    if rank == dst:
        full = torch.empty((n_streams,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        full[lo:hi].copy_(local)
        for r, (a, b) in enumerate(bounds):
            if r != dst and b > a:
                ops.append(dist.P2POp(dist.irecv, _bytes(full[a:b]), r, group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.isend, _bytes(local.contiguous()), dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full
