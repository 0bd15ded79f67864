"""N>1 host logic on CPU: two gloo ranks shard channels, broadcast the ADC block, gather frames.
The per-rank compute is the golden model (no GPU here); what is under test is the slab arithmetic and the
collective plumbing bench.py uses with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, n_samples, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ua3reo_loader
    from oracle import pyoracle
    pkg = ua3reo_loader.load()
    fcw = pkg.random_fcw(n_total, seed=42)
    lo, hi = pkg.sharding.channel_slab(n_total, rank, world)
    bc = pkg.sharding.AdcBroadcaster(n_samples, "cpu", src=0, dist=dist)
    blocks = []
    for b in range(2):
        local = torch.from_numpy(pkg.synth_adc(n_samples, seed=100 + b)) if rank == 0 else None
        blocks.append(bc.next_block(local).clone())
    banks = [pyoracle.GoldenDDC(w) for w in fcw[lo:hi]]
    rows = []
    for blk in blocks:
        rows.append(np.stack([g.push(blk.numpy()) for g in banks]) if banks else np.zeros((0, n_samples // 1024, 8), np.uint8))
    local_rows = torch.from_numpy(np.concatenate(rows, axis=1))
    full = pkg.sharding.gather_rows(local_rows, n_total, dist)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_slabs_cover_and_balance(pkg):
    for n, w in [(1024, 8), (10, 3), (7, 8), (8192, 8), (1, 2)]:
        slabs = [pkg.sharding.channel_slab(n, r, w) for r in range(w)]
        assert slabs[0][0] == 0 and slabs[-1][1] == n
        assert all(slabs[i][1] == slabs[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in slabs]
        assert max(sizes) - min(sizes) <= 1
        for c in (0, n - 1, n // 2):
            r = pkg.sharding.owner_of(c, n, w)
            assert slabs[r][0] <= c < slabs[r][1]


@pytest.mark.timeout(120)
def test_two_rank_broadcast_shard_gather(pkg, oracle, tmp_path):
    n_total, n_samples = 5, 4096        # odd channel count: uneven slabs
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_total, n_samples, out), nprocs=2, join=True)
    got = np.load(out)
    fcw = pkg.random_fcw(n_total, seed=42)
    adc = np.concatenate([pkg.synth_adc(n_samples, seed=100 + b) for b in range(2)])
    assert np.array_equal(got, oracle.golden_frames(adc, fcw))
