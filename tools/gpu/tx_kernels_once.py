#!/usr/bin/env python3
"""Runs processTxAudio + the DUC a few times (for ncu captures): python tools/gpu/tx_kernels_once.py [n_ch]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402

pkg = ua3reo_loader.load()
n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mix = [(0, 2700), (1, 2700), (4, 500), (10, 6000), (8, 15000)]
rx = pkg.Receiver(n_ch, 1 << 14)
rx.set_fcw(pkg.random_fcw(n_ch))
rx.tx_enable(1)
rx.duc_enable(192)
rx.tx_set([rx.tx_defaults(mode=mix[c % 5][0], filter_width=mix[c % 5][1]) for c in range(n_ch)])
mic = np.random.default_rng(0).integers(-8000, 8000, (n_ch, 192, 2)).astype(np.int16)
for _ in range(3):
    rx.tx_process(mic)
    rx.tx_feed_duc()
    rx.sync()
print("dac", rx.duc_read_dac().shape)
rx.close()
