#!/usr/bin/env python3
"""One-off fuzz (needs a GPU and oracle/_ref/fw_tx): random settings and microphone levels per channel, GPU processTxAudio
against live runs of the reference firmware."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402
from oracle import pyoracle  # noqa: E402

pkg = ua3reo_loader.load()
if os.environ.get("UA3REO_DEV_EMU") == "1":
    pkg.LIB_PATH = os.path.join(ROOT, "tools", "emu", "_build", "libua3reo_emu.so")
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_ch = int(sys.argv[2]) if len(sys.argv) > 2 else 48
rng = np.random.default_rng(seed)
widths = [300, 500, 1400, 1800, 2100, 2700, 3000, 3400, 4000, 5000, 6000, 7000, 8000, 9000, 9500, 10000, 15000, 0]
n = 192 * 10
t = np.arange(n)
cases, mics = [], []
for _ in range(n_ch):
    cases.append(dict(mode=int(rng.choice([0, 1, 2, 3, 4, 5, 6, 8, 9, 10])), filter_width=int(rng.choice(widths)),
                      ssb_hpf_pass=int(rng.choice([100, 200, 300, 400, 500])), rf_power=int(rng.integers(0, 101)),
                      mute=int(rng.random() < 0.1), tune=int(rng.random() < 0.1), key_down=int(rng.integers(2))))
    a = float(rng.choice([30, 800, 8000, 32000]))
    m = np.zeros((n, 2))
    m[:, 0] = a * np.sin(2 * np.pi * rng.uniform(100, 4000) * t / 48000) * (1 + 0.5 * np.sin(2 * np.pi * rng.uniform(1, 8) * t / 48000)) + rng.normal(0, a / 10, n)
    m[:, 1] = a * 0.6 * np.sin(2 * np.pi * rng.uniform(100, 4000) * t / 48000) + rng.normal(0, a / 10, n)
    mics.append(np.clip(np.rint(m), -32768, 32767).astype(np.int16))
mic = np.stack(mics)
rx = pkg.Receiver(n_ch, 1024)
rx.tx_enable(10)
rx.tx_set([rx.tx_defaults(**c) for c in cases])
w, f = rx.tx_process(mic)
fails, exact = [], 0
for i, c in enumerate(cases):
    rw, rf = pyoracle.run_fw_tx(mic[i], rx.tx_defaults(**c).as_dict())
    peak = max(np.abs(rf).max(), 1e-30)
    err = np.abs(f[i].astype(np.float64) - rf).max() / peak if np.any(rf) else float(np.any(f[i]))
    dw = np.abs(w[i].astype(np.int32) - rw.astype(np.int32)).max()
    exact += int(np.array_equal(w[i], rw))
    if err > 1e-5 or dw > 1:
        fails.append((i, c, err, int(dw)))
rx.close()
print("seed %d: %d channels, %d with identical wire words, failures: %d" % (seed, n_ch, exact, len(fails)))
for x in fails[:10]:
    print("  FAIL", x)
