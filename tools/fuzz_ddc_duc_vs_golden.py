#!/usr/bin/env python3
"""One-off fuzz (needs a GPU): random tuning words / amplitudes / push patterns, GPU DDC and DUC against the golden C model."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402
from oracle import pyoracle  # noqa: E402

pkg = ua3reo_loader.load()
pyoracle.build()
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(seed)
bad = 0
for trial in range(4):
    n_ch = int(rng.choice([5, 64, 300, 1024]))
    n = int(rng.integers(40, 200)) * 1024 + int(rng.integers(0, 1024))
    amp = float(rng.choice([2047, 600, 30]))
    adc = np.clip(np.rint(amp * np.sin(2 * np.pi * rng.uniform(0.01, 0.49) * np.arange(n)) + rng.normal(0, amp / 8, n)), -2048, 2047).astype(np.int16)
    fcw = rng.integers(0, 1 << 22, n_ch, dtype=np.uint32)
    fcw[:3] = [0, (1 << 22) - 1, 1 << 21][:min(3, n_ch)]
    rx = pkg.Receiver(n_ch, 1 << 18)
    rx.set_fcw(fcw)
    cuts = np.sort(rng.integers(0, n, 3)).tolist()
    got, off = [], 0
    for c in cuts + [n]:
        rx.push(adc[off:c]); off = c
        got.append(rx.read_frames())
    got = np.concatenate(got, 1)
    pick = rng.choice(n_ch, min(n_ch, 12), replace=False)
    ref = pyoracle.golden_frames(adc, [int(fcw[c]) for c in pick])
    # the register-transfer model emits frame k part-way through its last 1024 samples, the block-based kernels when all
    # of them have arrived: with a ragged total the model may be one frame ahead
    assert 0 <= ref.shape[1] - got.shape[1] <= 1 and got.shape[1] == n // 1024
    ok = np.array_equal(got[pick], ref[:, :got.shape[1]])
    # DUC on the same context
    rx.duc_enable(8)
    iq = rng.integers(-32768, 32768, (n_ch, 8, 2)).astype(np.int16)
    rx.duc_push(iq); d1 = rx.duc_read_dac()
    rx.duc_push(iq[:, :5]); d2 = rx.duc_read_dac()
    okd = True
    for c in pick[:6]:
        g = pyoracle.GoldenDUC(int(fcw[c]))
        r1, _ = g.push(iq[c, :, 0], iq[c, :, 1]); r2, _ = g.push(iq[c, :5, 0], iq[c, :5, 1])
        okd &= np.array_equal(d1[c], r1) and np.array_equal(d2[c, :5 * 1024], r2)
    rx.close()
    print("trial %d: n_ch=%d n=%d amp=%g cuts=%s  ddc %s  duc %s" % (trial, n_ch, n, amp, cuts, "ok" if ok else "MISMATCH", "ok" if okd else "MISMATCH"))
    bad += (not ok) + (not okd)
print("seed %d: %d mismatches" % (seed, bad))
