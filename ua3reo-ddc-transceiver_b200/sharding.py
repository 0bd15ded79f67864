"""Multi-GPU plumbing: channels shard by rank, the ADC block is broadcast, results stay sharded.

The path has no cross-channel term (one NCO, one filter chain, one TRX block per channel), so the
only exchange is the shared ADC block: rank `src` ingests it and broadcasts it to every rank before any compute.
torch.distributed is the transport (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def channel_slab(n_total, rank, world):
    """Contiguous slab [lo, hi) of channels owned by `rank`; slabs differ by at most one channel."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(channel, n_total, world):
    for r in range(world):
        lo, hi = channel_slab(n_total, r, world)
        if lo <= channel < hi:
            return r
    raise ValueError("channel out of range")


def torch_uint8():
    import torch
    return torch.uint8


class AdcBroadcaster:
    """Double-buffered broadcast of ADC blocks from the ingest rank (torch tensors on the compute device)."""

    def __init__(self, block_samples, device, src=0, dist=None):
        import torch
        self.dist = dist
        self.src = src
        self.bufs = [torch.empty(block_samples, dtype=torch.int16, device=device) for _ in range(2)]
        self.i = 0

    def next_block(self, local_block=None):
        """Rank `src` passes its block (tensor, same device); every rank gets the broadcast block back."""
        buf = self.bufs[self.i & 1]
        self.i += 1
        if self.dist is None or self.dist.get_world_size() == 1:
            if local_block is not None:
                buf.copy_(local_block, non_blocking=True)
            return buf
        if self.dist.get_rank() == self.src:
            buf.copy_(local_block, non_blocking=True)
        self.dist.broadcast(buf.view(torch_uint8()), src=self.src)   # NCCL/gloo have no int16: move the bytes
        return buf


def gather_rows(local_rows, n_total, dist):
    """All ranks contribute their slab's rows ([n_local, ...] tensors); rank 0 gets [n_total, ...] back (others None)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [channel_slab(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((pad,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    padded[:local_rows.shape[0]] = local_rows
    outs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, outs, dst=0)
    if rank != 0:
        return None
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
