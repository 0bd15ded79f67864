#!/usr/bin/env python3
"""Runs the STM32-stage kernels a few times on a mode-mixed bank (for ncu captures): python tools/gpu/rx_kernels_once.py [n_ch] [pushes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402

pkg = ua3reo_loader.load()
n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pushes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mix = [(0, 2700), (1, 2700), (4, 500), (10, 6000), (8, 15000)]
rx = pkg.Receiver(n_ch, 1 << 20, _lib_path=os.environ.get("UA3REO_LIB") or None)
rx.set_fcw(pkg.random_fcw(n_ch))
rx.rx_enable(True)
rx.rx_set([rx.rx_defaults(mode=mix[c % 5][0], filter_width=mix[c % 5][1], dnr=(c // 5) % 2, notch=(c // 5) % 2) for c in range(n_ch)])
adc = pkg.synth_adc(1 << 20)
for _ in range(pushes):
    rx.push(adc)
    rx.sync()
print("audio", rx.read_audio().shape, "spectra", rx.read_spectra().shape)
rx.close()
