#!/usr/bin/env python3
"""One-off fuzz (needs a GPU and oracle/_ref/fw_rx): random settings per channel, GPU full chain against live runs of the
reference firmware on each channel's own frames.  Prints the worst relative error per quantity and any failing settings."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ua3reo_loader  # noqa: E402
from oracle import pyoracle  # noqa: E402
from test_rx_gpu import _run_frames_through_gpu, stats  # noqa: E402

pkg = ua3reo_loader.load()
if os.environ.get("UA3REO_DEV_EMU") == "1":        # development aid: host emulation build of the CUDA sources
    pkg.LIB_PATH = os.path.join(ROOT, "tools", "emu", "_build", "libua3reo_emu.so")
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_ch = int(sys.argv[2]) if len(sys.argv) > 2 else 96
rng = np.random.default_rng(seed)
widths = [300, 500, 1400, 1600, 1800, 2100, 2300, 2500, 2700, 2900, 3000, 3200, 3400, 3600, 3800, 4000,
          4200, 4400, 4600, 4800, 5000, 5500, 6000, 6500, 7000, 7500, 8000, 8500, 9000, 9500, 10000, 15000, 0]
cases = []
for _ in range(n_ch):
    cases.append(dict(mode=int(rng.choice([0, 1, 2, 3, 4, 5, 6, 8, 9, 10])), filter_width=int(rng.choice(widths)),
                      ssb_hpf_pass=int(rng.choice([100, 200, 300, 400, 500])), dnr=int(rng.integers(2)), notch=int(rng.integers(2)),
                      notch_fc=int(rng.integers(200, 3500)), agc=int(rng.integers(2)), agc_speed=int(rng.integers(1, 11)),
                      rf_gain=int(rng.integers(5, 250)), volume=int(rng.integers(1, 101)), mute=int(rng.random() < 0.1),
                      iq_swap=int(rng.integers(2)), fm_sql_threshold=int(rng.integers(0, 10)), fft_zoom=int(rng.choice([1, 2, 4, 8, 16])),
                      fft_averaging=int(rng.integers(1, 9)), cw_decoder=int(rng.integers(2))))
n = 1024 * (192 * 8 + 1)
frames, audio, spec, sm = _run_frames_through_gpu(pkg, pyoracle, cases, n, [n // 3 + 11, n - (n // 3 + 11)], seed)
extra = _run_frames_through_gpu.extra
rx0 = pkg.Receiver(1, 1024)
worst = {"audio": 0.0, "spectra": 0.0, "cw": 0.0}
fails = []
exact = 0
for i, c in enumerate(cases):
    ref = pyoracle.run_fw_rx(frames[i], rx0.rx_defaults(**c).as_dict())
    nb = min(audio.shape[1], ref["audio"].shape[0]); nf = min(spec.shape[1], ref["spectra"].shape[0])
    ea, sa = stats(audio[i, :nb], ref["audio"][:nb]) if np.any(ref["audio"][:nb]) else (0.0 if not np.any(audio[i, :nb]) else 1.0, np.inf)
    es, ss = stats(spec[i, :nf], ref["spectra"][:nf])
    lsb = 1.0 / max(np.abs(ref["audio"][:nb]).max(), 1.0)
    ecw = np.abs(extra["cw"][i, :nb] - ref["cw"][:nb]).max() / max(np.abs(ref["cw"][:nb]).max(), 1.0)
    exact += int(np.array_equal(audio[i, :nb], ref["audio"][:nb]))
    worst["audio"] = max(worst["audio"], ea - lsb); worst["spectra"] = max(worst["spectra"], es); worst["cw"] = max(worst["cw"], ecw)
    if ea > 1e-5 + lsb or es > 1e-5 or ecw > 1e-5:
        fails.append((i, c, ea, es, ecw))
rx0.close()
print("seed %d: %d channels, %d bit-exact audio, worst rel err beyond one LSB: %s, failures: %d" % (seed, n_ch, exact, worst, len(fails)))
for f in fails[:10]:
    print("  FAIL", f)
