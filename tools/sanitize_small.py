#!/usr/bin/env python3
"""DEVELOPMENT AID: a small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck):
both front-kernel variants, ragged pushes, the STM32 stage with pipelined reads, the DUC and TX audio."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402

pkg = ua3reo_loader.load()
rng = np.random.default_rng(3)
n_ch = 300
rx = pkg.Receiver(n_ch, 1 << 16)
rx.set_fcw(rng.integers(0, 1 << 22, n_ch, dtype=np.uint32))
rx.rx_enable(True)
modes = [(0, 2700), (1, 2700), (4, 500), (10, 6000), (8, 15000)]
rx.rx_set([rx.rx_defaults(mode=modes[c % 5][0], filter_width=modes[c % 5][1], dnr=c & 1, notch=(c >> 1) & 1, fft_zoom=[1, 2][c % 2]) for c in range(n_ch)])
adc = rng.integers(-2048, 2048, 5 << 16, dtype=np.int16)
off = 0
for n in (1 << 16, 40000, 1000, 1 << 16, 30000):
    rx.push(adc[off:off + n]); off += n
    f = rx.read_frames(); a = rx.read_audio(); s = rx.read_spectra()
rx.move_waterfall(3000)
rx.push(adc[:1 << 16])
h = rx.read_waterfall_history()
print("rx ok", f.shape, a.shape, s.shape, h.shape)
rx.duc_enable(4)
rx.duc_push(rng.integers(-20000, 20000, (n_ch, 4, 2), dtype=np.int16))
dac = rx.duc_read_dac()
rx.tx_enable(2)
rx.tx_set(rx.tx_defaults(mode=1))
w, fl = rx.tx_process(rng.integers(-8000, 8000, (n_ch, 384, 2), dtype=np.int16))
print("tx ok", w.shape)
rx.close()
# small-table front kernel
os.environ["UA3REO_FRONT_VARIANT"] = "1"
rx = pkg.Receiver(40, 1 << 15)
rx.set_fcw(rng.integers(0, 1 << 22, 40, dtype=np.uint32))
rx.push(adc[:1 << 15]); rx.read_frames()
rx.close()
print("done")
