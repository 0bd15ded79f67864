#!/usr/bin/env python3
"""One-off fuzz (needs a GPU, oracle/_ref/fw_rx and fw_rx_b200): the same firmware driver with the reference's DSP units
(CPU) and with the firmware-name shim over the library (GPU), over random mid-stream events - settings written without
re-initialisation, ReinitAudioFilters(), InitNotchFilter(), FFT_Init(), retunes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle  # noqa: E402
from test_fw_shim_gpu import BASE, _frames  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_runs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rng = np.random.default_rng(seed)
env = dict(os.environ)
if os.environ.get("UA3REO_DEV_EMU") == "1":
    d = "/tmp/ua3_emulib"
    os.makedirs(d, exist_ok=True)
    link = os.path.join(d, "libua3reo_b200.so")
    if not os.path.exists(link):
        os.symlink(os.path.join(ROOT, "tools", "emu", "_build", "libua3reo_emu.so"), link)
    env["LD_LIBRARY_PATH"] = d
widths = [300, 500, 1800, 2700, 3000, 3400, 6000, 8000, 10000, 15000, 0]
bad = 0
for run in range(n_runs):
    s = dict(BASE)
    # (initial width never 0: switching the LPF on later without ReinitAudioFilters() runs the firmware's lattice on an
    #  instance that was never initialised - a null dereference in the reference, not a parity case)
    s.update(mode=int(rng.choice([0, 1, 3, 4, 8, 10])), filter_width=int(rng.choice(widths[:-1])), dnr=int(rng.integers(2)),
             notch=int(rng.integers(2)), fft_zoom=int(rng.choice([1, 2, 4])))
    events = []
    for at in sorted(rng.integers(200, 3600, 10).tolist()):
        k = rng.choice(["volume", "agc", "mode", "notch", "rf_gain", "dnr", "mute", "fft_averaging", "filter_width", "notch_fc",
                        "reinit", "notch_init", "fft_zoom+init", "freq", "agc_speed", "agc_speed+init", "fm_sql_threshold", "smeter_reset"])
        if k == "volume": events.append((at, k, int(rng.integers(1, 101))))
        elif k in ("agc", "notch", "dnr", "mute"): events.append((at, k, int(rng.integers(2))))
        elif k == "mode": events.append((at, k, int(rng.choice([0, 1, 3, 4, 8, 10]))))
        elif k == "rf_gain": events.append((at, k, int(rng.integers(5, 250))))
        elif k == "fft_averaging": events.append((at, k, int(rng.integers(1, 9))))
        elif k == "agc_speed": events.append((at, k, int(rng.integers(1, 11))))          # no effect until InitAGC()
        elif k == "agc_speed+init": events += [(at, "agc_speed", int(rng.integers(1, 11))), (at, "agc_init", 0)]
        elif k == "fm_sql_threshold": events.append((at, k, int(rng.integers(0, 10))))
        elif k == "filter_width": events.append((at, k, int(rng.choice(widths))))
        elif k == "notch_fc": events.append((at, k, int(rng.integers(200, 3500))))
        elif k == "reinit": events.append((at, "reinit", 0))
        elif k == "smeter_reset": events.append((at, "smeter_reset", 0))
        elif k == "notch_init": events.append((at, "notch_init", 0))
        elif k == "fft_zoom+init": events += [(at, "fft_zoom", int(rng.choice([1, 2, 4, 8]))), (at, "fft_init", 0)]
        # retunes stay below 256 waterfall pixels ((diff / 187) * zoom): beyond that FFT_moveWaterfall reads past FFTOutput_mean
        elif k == "freq": events.append((at, "freq", int(rng.integers(0, 5000))))
    fr = _frames(192 * 20, int(rng.integers(1 << 30)))
    try:
        ref = pyoracle.run_fw_rx(fr, s, events=events)
    except Exception as e:                       # the reference itself died (undefined behaviour in the firmware): not a parity case
        print("run %d: REFERENCE CRASHED (%s) settings=%s events=%s" % (run, type(e).__name__, {k: s[k] for k in ("mode", "filter_width", "fft_zoom")}, events))
        continue
    got = pyoracle.run_fw_rx(fr, s, events=events, binary=pyoracle.FW_RX_B200, env=env)
    res = {}
    for k in ("audio", "usb", "smeter", "cw", "spectra", "waterfall", "wtf_history"):
        a, b = ref[k].astype(np.float64), got[k].astype(np.float64)
        res[k] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1.0)) if a.shape == b.shape else 9.9
    ok = res["audio"] <= 1e-5 and res["spectra"] <= 1e-5 and res["usb"] <= 1e-4 and res["cw"] <= 1e-5 and \
        np.mean(ref["wtf_history"] != got["wtf_history"]) <= 2e-3
    bad += (not ok)
    print("run %d: %s  %s  events=%s" % (run, "ok" if ok else "MISMATCH", {k: "%.1e" % v for k, v in res.items()}, events if not ok else len(events)))
print("seed %d: %d mismatching runs of %d" % (seed, bad, n_runs))
