"""ua3reo_fanout_* (csrc/fanout.cu; sharding.AdcFanout): the ADC block fanned out to one process per GPU with copy engines
over CUDA IPC mappings and stream-side flag waits.  Two PROCESSES share cuda:0 here (CUDA IPC works between processes on
one device as well), gloo carries the handles; more blocks than slots, so that the credit path is exercised; every rank's
frames are compared with the golden model bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
N_CH, BLOCK, N_BLOCKS = 6, 1 << 16, 9


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, pinned):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ua3reo_loader
    pkg = ua3reo_loader.load()
    torch.cuda.set_device(0)
    lo, hi = pkg.sharding.channel_slab(N_CH, rank, world)
    fcw = pkg.random_fcw(N_CH, seed=5)
    rx = pkg.Receiver(hi - lo, BLOCK, device=0)
    rx.set_fcw(fcw[lo:hi])
    fo = pkg.sharding.AdcFanout(rx.lib, BLOCK, 0, rx.stream(), src=0, dist=dist, n_buffers=3)
    adc = pkg.synth_adc(N_BLOCKS * BLOCK, seed=11).reshape(N_BLOCKS, BLOCK)
    if rank == 0:
        src = torch.from_numpy(adc).pin_memory() if pinned else torch.from_numpy(adc).cuda()
    frames = []
    fo.prefetch(src[0] if rank == 0 else None)
    for i in range(N_BLOCKS):
        if i + 1 < N_BLOCKS:
            fo.prefetch(src[i + 1] if rank == 0 else None)
        buf = fo.acquire()
        assert rx.push(buf) == BLOCK // 1024
        fo.release(buf)
        if rank == 1 and i == 3:
            import time
            time.sleep(0.3)              # a slow consumer: the ingest rank must wait for its credit, not overrun the slot
        frames.append(rx.read_frames())
    info = fo.info()
    assert info["acquired"] == N_BLOCKS and (rank != 0 or info["sent"] == N_BLOCKS)
    torch.cuda.synchronize()
    fo.close()
    rx.close()
    np.save(os.path.join(out_dir, "frames_%d.npy" % rank), np.concatenate(frames, axis=1))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("pinned", [False, True])
def test_two_process_fanout_frames_are_bit_exact(pkg, oracle, tmp_path, pinned):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), pinned), nprocs=2, join=True)
    fcw = pkg.random_fcw(N_CH, seed=5)
    adc = pkg.synth_adc(N_BLOCKS * BLOCK, seed=11)
    want = oracle.golden_frames(adc, fcw)
    got = np.concatenate([np.load(str(tmp_path / ("frames_%d.npy" % r))) for r in range(2)], axis=0)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_single_rank_fanout_is_a_plain_ring(pkg, oracle):
    import torch
    rx = pkg.Receiver(3, BLOCK, device=0)
    fcw = pkg.random_fcw(3, seed=6)
    rx.set_fcw(fcw)
    fo = pkg.sharding.AdcFanout(rx.lib, BLOCK, 0, rx.stream())
    adc = pkg.synth_adc(5 * BLOCK, seed=12).reshape(5, BLOCK)
    dev = torch.from_numpy(adc).cuda()
    frames = []
    for i in range(5):
        buf = fo.next_block(dev[i])
        rx.push(buf)
        fo.release(buf)
        frames.append(rx.read_frames())
    torch.cuda.synchronize()
    fo.close()
    assert np.array_equal(np.concatenate(frames, axis=1), oracle.golden_frames(adc.reshape(-1), fcw))
    with pytest.raises(pkg.UA3Error):
        pkg.sharding.AdcFanout(rx.lib, BLOCK, 0, rx.stream(), n_buffers=1)
    rx.close()


def _gather_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ua3reo_loader
    pkg = ua3reo_loader.load()
    torch.cuda.set_device(0)
    lib = pkg.load_library()
    n_words, n_steps = 5000, 11                                  # more steps than slots: credits come back from the root
    g = pkg.sharding.SlabGather(lib, n_words * 4, 0, root=0, dist=dist, n_buffers=3)
    st = torch.cuda.Stream()
    got = []
    srcs = [torch.empty(n_words, dtype=torch.int32, device="cuda") for _ in range(2)]

    class Raw:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    for s in range(n_steps):
        with torch.cuda.stream(st):
            srcs[s & 1].copy_(torch.arange(n_words, dtype=torch.int32, device="cuda") * (rank + 1) + 1000 * s)
        g.send(srcs[s & 1].data_ptr(), st.cuda_stream)
        if rank == 0:
            ptr, stride = g.acquire(st.cuda_stream)
            with torch.cuda.stream(st):
                allr = torch.as_tensor(Raw(ptr, world * stride), device="cuda").view(world, stride)[:, :n_words * 4]
                got.append(allr.contiguous().view(torch.int32).view(world, n_words).cpu().numpy().copy())
            g.release(st.cuda_stream)
            if s == 4:
                import time
                time.sleep(0.3)                                  # a slow root: the senders must wait for their credits
    torch.cuda.synchronize()
    g.close()
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), np.stack(got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_process_gather(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_gather_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = np.load(str(tmp_path / "gathered.npy"))
    assert got.shape == (11, 2, 5000)
    base = np.arange(5000, dtype=np.int64)
    for s in range(11):
        for r in range(2):
            assert np.array_equal(got[s, r], base * (r + 1) + 1000 * s), (s, r)
