#!/usr/bin/env python3
"""TX DUC throughput (supporting measurement for DESIGN.md 4.5): channel x DAC-samples / s for several channel counts."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ua3reo_loader
pkg = ua3reo_loader.load()
out = []
for n_ch in (1024, 4096, 16384):
    rx = pkg.Receiver(n_ch, 1 << 14); rx.set_fcw(pkg.random_fcw(n_ch)); rx.duc_enable(16)
    iq = np.random.default_rng(0).integers(-20000, 20000, (n_ch, 16, 2)).astype(np.int16)
    rx.duc_push(iq); rx.sync()
    t = time.time()
    for _ in range(3):
        rx.duc_push(iq)
    rx.sync(); dt = (time.time() - t) / 3
    out.append({"channels": n_ch, "ms_per_16_tx_samples": dt * 1e3, "channel_dac_samples_per_s": n_ch * 16 * 1024 / dt,
                "real_time_channels": n_ch * 16 * 1024 / dt / 49152000})
    rx.close()
print(json.dumps(out))
