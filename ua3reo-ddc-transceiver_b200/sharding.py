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


def _ipc_open(lib, kind, create_args, dist, group):
    """Collective set-up of one ua3reo_fanout / ua3reo_gather end per rank: create, all_gather of the 64-byte handles,
    connect, and a second all_gather of the outcome so that EVERY rank either returns a handle or raises UA3Error."""
    import ctypes
    from . import UA3Error
    fn = lambda name: getattr(lib, "ua3reo_%s_%s" % (kind, name))
    world = 1 if dist is None else dist.get_world_size(group)
    h = ctypes.c_void_p()
    rc = fn("create")(*create_args, ctypes.byref(h))
    msg = "" if rc == 0 else lib.ua3reo_last_error().decode()
    handle = h if rc == 0 else None
    if world == 1:
        if rc != 0:
            raise UA3Error("%s: %s" % (kind, msg))
        return handle
    mine = (ctypes.c_uint8 * 64)()
    if rc == 0:
        rc = fn("handle")(handle, mine)
        if rc != 0:
            msg = lib.ua3reo_last_error().decode()
    payload = [None] * world
    dist.all_gather_object(payload, bytes(mine) if rc == 0 else None, group=group)     # tiny, once
    if rc == 0 and all(p is not None for p in payload):
        allh = (ctypes.c_uint8 * (64 * world)).from_buffer_copy(b"".join(payload))
        rc = fn("connect")(handle, allh)
        if rc != 0:
            msg = lib.ua3reo_last_error().decode()
    elif rc == 0:
        rc, msg = -5, ""                           # a peer could not create its end: its own message names the cause
    oks = [None] * world
    dist.all_gather_object(oks, (rc == 0, msg), group=group)
    bad = sorted(((r, m) for r, (ok, m) in enumerate(oks) if not ok), key=lambda rm: (rm[1] == "", rm[0]))
    if bad:                                        # every rank sees the same verdict: unmap, meet, free
        if handle is not None:
            fn("disconnect")(handle)
        dist.barrier(group=group)
        if handle is not None:
            fn("destroy")(handle)
        raise UA3Error("%s: rank %d: %s" % ((kind,) + bad[0]))
    return handle


def _ipc_close(lib, kind, handle, dist, group):
    """COLLECTIVE tear-down: peers are unmapped first, the arenas are freed only after a barrier."""
    if handle is None:
        return
    multi = dist is not None and dist.get_world_size(group) > 1
    if multi:
        dist.barrier(group=group)                  # every rank's streams have been synchronised by its caller
    getattr(lib, "ua3reo_%s_disconnect" % kind)(handle)
    if multi:
        dist.barrier(group=group)
    getattr(lib, "ua3reo_%s_destroy" % kind)(handle)


class AdcFanout:
    """The same pipeline as AdcBroadcaster without a collective kernel: ua3reo_fanout_* (csrc/fanout.cu) - the ingest rank's
    copy engines write every block into each rank's slot over NVLink (CUDA IPC mappings) and the consumer STREAMS wait on
    a flag word, so no SM has to be kept free for the transfer and the ranks' host threads never meet.  Same calls:

        fo.prefetch(block)            # rank `src` passes its block (device tensor, pinned host tensor or numpy view), the others None
        buf = fo.acquire()            # DeviceBlock; the consumer stream now waits for that block's flag
        rx.push(buf)
        fo.release(buf)               # the slot's credit returns behind the push

    Construction is collective (one all_gather of the 64-byte handles); it raises UA3Error on EVERY rank when any rank
    could not map its peers, so that the caller can fall back to AdcBroadcaster on all of them."""

    def __init__(self, lib, block_samples, device_index, consumer_stream, src=0, dist=None, n_buffers=3, group=None):
        import ctypes
        from . import UA3Error, DeviceBlock
        self._ct, self._DeviceBlock, self._err = ctypes, DeviceBlock, UA3Error
        self.lib, self.dist, self.group, self.src = lib, dist, group, src
        self.block = int(block_samples)
        self.world = 1 if dist is None else dist.get_world_size(group)
        self.rank = 0 if dist is None else dist.get_rank(group)
        self.consumer = ctypes.c_void_p(int(consumer_stream))
        self.n_buffers = int(n_buffers)
        self.n_filled = self.n_taken = 0
        self._keep = []
        self._h = _ipc_open(lib, "fanout", (int(device_index), self.rank, self.world, int(src), self.block, self.n_buffers), dist, group)

    def _chk(self, rc):
        if rc != 0:
            raise self._err("ua3reo fan-out error %d: %s" % (rc, self.lib.ua3reo_last_error().decode()))

    def prefetch(self, local_block=None):
        assert self.n_filled - self.n_taken < self.n_buffers, "AdcFanout: every slot holds a block that was not acquired yet"
        if self.rank == self.src:
            ptr = local_block.ctypes.data if isinstance(local_block, np.ndarray) else local_block.data_ptr()
            n = local_block.size if isinstance(local_block, np.ndarray) else local_block.numel()
            self._keep = self._keep[-2 * self.n_buffers:] + [local_block]        # the copy reads it later
            self._chk(self.lib.ua3reo_fanout_send(self._h, ptr, n))
        self.n_filled += 1

    def acquire(self):
        assert self.n_taken < self.n_filled, "AdcFanout.acquire without a prefetch"
        p = self._ct.c_void_p()
        self._chk(self.lib.ua3reo_fanout_acquire(self._h, self.consumer, self._ct.byref(p)))
        self.n_taken += 1
        return self._DeviceBlock(p.value, self.block)

    def release(self, buf=None):
        self._chk(self.lib.ua3reo_fanout_release(self._h, self.consumer))

    def next_block(self, local_block=None):
        self.prefetch(local_block)
        return self.acquire()

    def info(self):
        d, s, a = self._ct.c_int(), self._ct.c_uint64(), self._ct.c_uint64()
        self._chk(self.lib.ua3reo_fanout_info(self._h, self._ct.byref(d), self._ct.byref(s), self._ct.byref(a)))
        return {"direct_remote_store": bool(d.value), "sent": s.value, "acquired": a.value}

    def close(self):
        """COLLECTIVE: every rank calls it once its consumer stream is idle."""
        _ipc_close(self.lib, "fanout", self._h, self.dist, self.group)
        self._h = None


class SlabGather:
    """Every rank's slab of results into one buffer on the root rank without a collective kernel (ua3reo_gather_*):

        g.send(slab_ptr, stream)                 # every rank, enqueued on the stream that produced the slab
        ptr, stride = g.acquire(stream)          # root: `stream` waits for all ranks; rank r's slab is at ptr + r * stride
        ...
        g.release(stream)                        # root: credits return behind the work on `stream`

    Construction and close() are collective, as for AdcFanout."""

    def __init__(self, lib, slab_bytes, device_index, root=0, dist=None, n_buffers=4, group=None):
        import ctypes
        from . import UA3Error
        self._ct, self._err = ctypes, UA3Error
        self.lib, self.dist, self.group, self.root = lib, dist, group, root
        self.world = 1 if dist is None else dist.get_world_size(group)
        self.rank = 0 if dist is None else dist.get_rank(group)
        self.slab_bytes = int(slab_bytes)
        self._h = _ipc_open(lib, "gather", (int(device_index), self.rank, self.world, int(root), self.slab_bytes, int(n_buffers)), dist, group)

    def _chk(self, rc):
        if rc != 0:
            raise self._err("ua3reo gather error %d: %s" % (rc, self.lib.ua3reo_last_error().decode()))

    def send(self, slab_ptr, stream):
        self._chk(self.lib.ua3reo_gather_send(self._h, self._ct.c_void_p(int(slab_ptr)), self._ct.c_void_p(int(stream))))

    def acquire(self, stream):
        p, stride = self._ct.c_void_p(), self._ct.c_size_t()
        self._chk(self.lib.ua3reo_gather_acquire(self._h, self._ct.c_void_p(int(stream)), self._ct.byref(p), self._ct.byref(stride)))
        return p.value, stride.value

    def release(self, stream):
        self._chk(self.lib.ua3reo_gather_release(self._h, self._ct.c_void_p(int(stream))))

    def close(self):
        _ipc_close(self.lib, "gather", self._h, self.dist, self.group)
        self._h = None


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
