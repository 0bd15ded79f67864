#!/usr/bin/env python3
"""processTxAudio() throughput on the GPU (supporting measurement for DESIGN.md 4.5): channel x 48 kHz samples / s for the
SSB, AM and FM modulators, host buffers in and I/Q words out (the ua3reo_tx_process / ua3reo_tx_read_iq calls)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402

pkg = ua3reo_loader.load()
out = []
n_blocks = 8
for n_ch in (1024, 4096, 16384):
    for name, mode, width in (("USB", 1, 2700), ("AM", 10, 6000), ("NFM", 8, 8000)):
        rx = pkg.Receiver(n_ch, 1 << 14)
        rx.tx_enable(n_blocks)
        rx.tx_set(rx.tx_defaults(mode=mode, filter_width=width))
        mic = np.random.default_rng(0).integers(-8000, 8000, (n_ch, n_blocks * 192, 2)).astype(np.int16)
        rx.tx_process(mic)
        t = time.time()
        for _ in range(3):
            rx.tx_process(mic)
        dt = (time.time() - t) / 3
        out.append({"channels": n_ch, "mode": name, "ms_per_%d_blocks" % n_blocks: dt * 1e3,
                    "channel_samples_per_s": n_ch * n_blocks * 192 / dt, "real_time_channels": n_ch * n_blocks * 192 / dt / 48000.0})
        rx.close()
print(json.dumps(out))
