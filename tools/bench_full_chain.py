#!/usr/bin/env python3
"""BASELINE config 5 shape: full per-channel RX chain (DDC + bandpass/notch + mixed SSB/AM/FM demod + AGC + LMS DNR +
panorama FFT) for N channels, per-kernel device times.  Not the contract bench (bench.py measures the DDC metric);
this is the supporting measurement for DESIGN.md section 4.3/4.4."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402


def main():
    import torch
    pkg = ua3reo_loader.load()
    n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    block = 1 << 20
    rx = pkg.Receiver(n_ch, block)
    rx.set_fcw(pkg.random_fcw(n_ch))
    rx.rx_enable(True)
    modes = [(0, 2700), (1, 2700), (4, 500), (10, 6000), (8, 15000)]      # LSB, USB, CW_U, AM, NFM round-robin
    sets = []
    for c in range(n_ch):
        m, w = modes[c % 5]
        half = (c // 5) % 2
        sets.append(rx.rx_defaults(mode=m, filter_width=w, dnr=half, notch=half))
    rx.rx_set(sets)
    adc = torch.from_numpy(pkg.synth_adc(2 * block).reshape(2, block)).cuda()
    for i in range(3):
        rx.push(adc[i & 1])
    rx.sync()
    rx.profile_begin(steps)
    t0 = time.time()
    for i in range(steps):
        rx.push(adc[i & 1])
    rx.sync()
    dt = time.time() - t0
    kms, nb = rx.profile_end()
    per = {k: v / nb for k, v in kms.items()}
    total = sum(per.values())
    print(json.dumps({"workload": "config 5: %d channels, full RX chain, modes LSB/USB/CW_U/AM/NFM round-robin, DNR+notch on half" % n_ch,
                      "block_samples": block, "steps": steps, "ms_per_step_wall": 1e3 * dt / steps, "kernel_ms_per_step": per,
                      "kernel_ms_total": total, "channel_samples_per_s": n_ch * block / (total * 1e-3),
                      "real_time_channels": n_ch * block / (total * 1e-3) / 49152000.0}))
    rx.close()


if __name__ == "__main__":
    main()
