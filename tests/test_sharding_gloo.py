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


class _FakeLib:
    """Stands in for libua3reo_b200.so in the rendezvous logic of sharding._ipc_open: create() fails on `bad_rank`,
    connect() records the handles it was given."""

    def __init__(self, rank, bad_rank):
        self.rank, self.bad_rank, self.calls = rank, bad_rank, []

    def ua3reo_last_error(self):
        return b"no peer mapping (fake)"

    def ua3reo_fanout_create(self, *args):
        self.calls.append("create")
        return -2 if self.rank == self.bad_rank else 0

    def ua3reo_fanout_handle(self, h, buf):
        for i in range(64):
            buf[i] = (self.rank * 7 + i) & 0xFF
        return 0

    def ua3reo_fanout_connect(self, h, allh):
        self.calls.append(("connect", bytes(allh)))
        return 0

    def ua3reo_fanout_disconnect(self, h):
        self.calls.append("disconnect")
        return 0

    def ua3reo_fanout_destroy(self, h):
        self.calls.append("destroy")
        return 0


def _rendezvous_worker(rank, world, port, bad_rank, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ua3reo_loader
    pkg = ua3reo_loader.load()
    lib = _FakeLib(rank, bad_rank)
    try:
        pkg.sharding._ipc_open(lib, "fanout", (0, rank, world, 0, 1024, 3), dist, None)
        verdict = "ok"
    except pkg.UA3Error as e:
        verdict = "error: %s" % e
    with open(os.path.join(out_dir, "r%d.txt" % rank), "w") as f:
        f.write(verdict + "\n" + repr([c if isinstance(c, str) else (c[0], len(c[1])) for c in lib.calls]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("bad_rank", [-1, 1])
def test_ipc_rendezvous_is_collective(tmp_path, bad_rank):
    """sharding._ipc_open (what AdcFanout / SlabGather are built with): all ranks connect with all handles in rank order, or -
    when any rank cannot create its end - ALL of them raise (so that bench.py falls back to the NCCL transport on every rank
    instead of hanging), after the ranks that did create have unmapped and destroyed theirs."""
    mp.spawn(_rendezvous_worker, args=(2, _free_port(), bad_rank, str(tmp_path)), nprocs=2, join=True)
    res = [open(str(tmp_path / ("r%d.txt" % r))).read().split("\n") for r in range(2)]
    if bad_rank < 0:
        assert [r[0] for r in res] == ["ok", "ok"]
        assert all("('connect', 128)" in r[1] and "destroy" not in r[1] for r in res)
    else:
        assert all(r[0].startswith("error: fanout: rank 1: no peer mapping") for r in res), res        # the root cause, on both ranks
        assert "disconnect" in res[0][1] and "destroy" in res[0][1]          # the healthy rank cleaned up
        assert "connect" not in res[1][1].replace("disconnect", "")
