#!/usr/bin/env python3
"""DEVELOPMENT AID: runs the CUDA sources' host-emulation build (tools/emu) against the golden model."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ua3reo_loader
from oracle import pyoracle

pkg = ua3reo_loader.load()
EMU = os.path.join(ROOT, "tools", "emu", "_build", "libua3reo_emu.so")

def run(n_ch, pushes, max_block=1 << 14, seed=1, clocking=None):
    rng = np.random.default_rng(seed)
    fcw = rng.integers(1, 1 << 21, n_ch).astype(np.uint32)
    total = sum(pushes)
    adc = pyoracle.synth_adc(total, seed=seed)
    adc[:5] = -2048   # exercise the (-2048)*(-2048) mixer wrap
    rx = pkg.Receiver(n_ch, max_block, _lib_path=EMU)
    rx.set_fcw(fcw)
    if clocking is not None:
        rx.set_clocking(*clocking)
    got = []
    off = 0
    for n in pushes:
        nf = rx.push(adc[off:off + n]); off += n
        got.append(rx.read_frames())
    got = np.concatenate(got, axis=1)
    ref = pyoracle.golden_frames(adc, fcw, clocking)[:, :got.shape[1]]
    ok = np.array_equal(got, ref)
    print("n_ch=%d pushes=%s clocking=%s frames=%d match=%s" % (n_ch, pushes, clocking, got.shape[1], ok))
    if not ok:
        bad = np.argwhere(got != ref)
        print(" first mismatches:", bad[:8].tolist())
        print(" got", got[bad[0][0], bad[0][1]], "ref", ref[bad[0][0], bad[0][1]])
    return ok

if __name__ == "__main__":
    t = time.time()
    ok = run(3, [4096, 2048, 8192])
    ok &= run(33, [16384], seed=2)
    ok &= run(2, [1000, 24, 3000, 5192, 1024 * 3 + 5], seed=3)
    for cl in ((0, 3, 129), (0, 3, 130), (0, 2, 129), (1, 2, 129), (1, 3, 130), (1, 3, 129), (0, 0, 130)):
        ok &= run(2, [1 << 16, 1 << 16, 3000, (1 << 16) - 3000, 1 << 15], max_block=1 << 16, seed=4, clocking=cl)
    print("ALL OK" if ok else "FAILED", "%.1fs" % (time.time() - t))
    sys.exit(0 if ok else 1)
