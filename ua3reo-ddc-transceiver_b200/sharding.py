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
    """Pipelined broadcast of ADC blocks from the ingest rank into a small ring of device buffers.

    The library consumes a device push in place on its own stream (include/ua3reo_b200.h, DEVICE PUSH CONTRACT), so the
    broadcaster owns both orderings: the consumer stream waits for the broadcast that filled a buffer (`acquire`), and
    the broadcast that refills a buffer waits for the consumer work that read it (`release`).  Broadcasts run on a side
    stream, so that the transfer of block i+1 overlaps the kernels of block i:

        bc.prefetch(block0)
        for i in range(n):
            if i + 1 < n: bc.prefetch(block[i + 1])     # rank `src` passes its block, the others None
            buf = bc.acquire()                          # consumer stream now waits for the broadcast of block i
            rx.push(buf, assume_ordered=True)           # enqueued on the consumer stream
            bc.release(buf)                             # buffer may be refilled once that push has finished

    On CPU tensors (gloo, the unit tests) everything is synchronous and the events are skipped."""

    def __init__(self, block_samples, device, src=0, dist=None, consumer_stream=None, n_buffers=2, group=None):
        import torch
        self.torch = torch
        self.dist = dist
        self.group = group
        self.src = src
        self.cuda = torch.device(device).type == "cuda"
        self.bufs = [torch.empty(block_samples, dtype=torch.int16, device=device) for _ in range(n_buffers)]
        self.n_filled = 0          # blocks broadcast so far
        self.n_taken = 0           # blocks handed to the consumer so far
        if self.cuda:
            assert consumer_stream is not None, "AdcBroadcaster on a GPU needs the stream that consumes the blocks (rx.stream())"
            self.consumer = consumer_stream if isinstance(consumer_stream, torch.cuda.Stream) else \
                torch.cuda.ExternalStream(int(consumer_stream), device=device)
            self.side = torch.cuda.Stream(device=device)
            self.ev_ready = [torch.cuda.Event() for _ in range(n_buffers)]
            self.ev_free = [torch.cuda.Event() for _ in range(n_buffers)]
            self.released = [False] * n_buffers

    def _world(self):
        return 1 if self.dist is None else self.dist.get_world_size(self.group)

    def prefetch(self, local_block=None):
        """Enqueues the broadcast of the next block.  Rank `src` passes the block (same device, or pinned host memory)."""
        assert self.n_filled - self.n_taken < len(self.bufs), "AdcBroadcaster: every buffer holds a block that was not acquired yet"
        b = self.n_filled % len(self.bufs)
        buf = self.bufs[b]
        is_src = self._world() == 1 or self.dist.get_rank(self.group) == self.src
        if not self.cuda:
            if is_src:
                buf.copy_(local_block)
            if self._world() > 1:
                self.dist.broadcast(buf.view(torch_uint8()), src=self.src, group=self.group)
        else:
            torch = self.torch
            with torch.cuda.stream(self.side):
                if self.released[b]:
                    self.side.wait_event(self.ev_free[b])          # the push that read this buffer has finished
                    self.released[b] = False
                if is_src:
                    buf.copy_(local_block, non_blocking=True)
                if self._world() > 1:
                    self.dist.broadcast(buf.view(torch_uint8()), src=self.src, group=self.group)   # no int16 in NCCL: move bytes
                self.ev_ready[b].record(self.side)
        self.n_filled += 1

    def acquire(self):
        """The oldest prefetched block; the consumer stream is made to wait for its broadcast."""
        assert self.n_taken < self.n_filled, "AdcBroadcaster.acquire without a prefetch"
        b = self.n_taken % len(self.bufs)
        self.n_taken += 1
        if self.cuda:
            self.consumer.wait_event(self.ev_ready[b])
        return self.bufs[b]

    def release(self, buf):
        """Call after the consumer's work on `buf` has been enqueued on the consumer stream."""
        if self.cuda:
            b = next(i for i, x in enumerate(self.bufs) if x is buf)
            self.ev_free[b].record(self.consumer)
            self.released[b] = True

    def next_block(self, local_block=None):
        """prefetch + acquire in one call (no overlap).  On a GPU the caller still calls release(buf) after the push."""
        self.prefetch(local_block)
        return self.acquire()


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
