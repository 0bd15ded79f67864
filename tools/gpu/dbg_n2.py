"""debug: the N>1 parity leg of bench.py on its own (torchrun --nproc-per-node 2 tools/gpu/dbg_n2.py)"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ua3reo_loader
from oracle import pyoracle
pkg = ua3reo_loader.load()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_ch, block = int(os.environ.get('NCH', '1024')), 1 << 20
rx = pkg.Receiver(n_ch, block, device=local)
fcw = pkg.random_fcw(n_ch * world, 20261018)[rank * n_ch:(rank + 1) * n_ch]
rx.set_fcw(fcw)
ext = torch.cuda.ExternalStream(rx.stream(), device=local)
host_np = pkg.synth_adc(2 * block, 20261018).reshape(2, block)
dev = torch.from_numpy(host_np).cuda() if rank == 0 else None
bc = pkg.sharding.AdcBroadcaster(block, torch.device("cuda", local), src=0, dist=dist, consumer_stream=ext)
# a few pipelined steps first, as the bench does
bc.prefetch(dev[0] if rank == 0 else None)
for i in range(6):
    if i + 1 < 6:
        bc.prefetch(dev[(i + 1) % 2] if rank == 0 else None)
    b_ = bc.acquire(); rx.push(b_, assume_ordered=True); bc.release(b_)
rx.sync()
for trial in range(3):
    rx.reset()
    bc.prefetch(dev[trial % 2] if rank == 0 else None)
    buf = bc.acquire()
    rx.push(buf, assume_ordered=True)
    bc.release(buf)
    got = rx.read_frames()
    torch.cuda.synchronize()
    same_input = np.array_equal(buf.cpu().numpy(), host_np[trial % 2])
    ref = pyoracle.GoldenDDC(int(fcw[0])).push(host_np[trial % 2])
    bad = np.argwhere(got[0] != ref)
    refl = pyoracle.GoldenDDC(int(fcw[n_ch - 1])).push(host_np[trial % 2])
    print('rank', rank, 'last channel equal:', np.array_equal(got[n_ch - 1], refl), flush=True)
    print("rank", rank, "trial", trial, "broadcast buffer == block:", same_input, "frames equal:", bad.size == 0, "first bad", bad[:3].tolist(), flush=True)
del bc
rx.close()
dist.destroy_process_group()
