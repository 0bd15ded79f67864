#!/usr/bin/env python3
"""Generates tests/golden/tx_cases.npz: codec input samples + settings and the outputs of the REFERENCE FIRMWARE's
own processTxAudio() (host-built from /root/reference by oracle/ref_harness, binary oracle/_ref/fw_tx).
Run:  make -C oracle ref && python tools/gen_golden_tx.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

DEFAULTS = dict(mode=1, mute=0, tune=0, key_down=0, rf_power=20, filter_width=2700, ssb_hpf_pass=300)
CASES = [
    dict(name="usb", settings=dict(mode=1)),
    dict(name="lsb_3k0_hpf100", settings=dict(mode=0, filter_width=3000, ssb_hpf_pass=100)),
    dict(name="digi_u", settings=dict(mode=6)),
    dict(name="digi_l_full_power", settings=dict(mode=5, rf_power=100)),
    dict(name="am_6k", settings=dict(mode=10, filter_width=6000)),
    dict(name="nfm_9k", settings=dict(mode=8, filter_width=9000)),
    dict(name="wfm_15k", settings=dict(mode=9, filter_width=15000, rf_power=100)),
    dict(name="cw_u_key_down", settings=dict(mode=4, key_down=1, filter_width=500)),
    dict(name="cw_l_key_up", settings=dict(mode=3, key_down=0, filter_width=500)),
    dict(name="iq", settings=dict(mode=2)),
    dict(name="usb_tune", settings=dict(mode=1, tune=1)),
    dict(name="usb_mute", settings=dict(mode=1, mute=1)),
    dict(name="lsb_lpf_off", settings=dict(mode=0, filter_width=0)),
]


def make_mic(seed, n):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = 7000 * np.sin(2 * np.pi * 1000 * t / 48000) + 3000 * np.sin(2 * np.pi * 1700 * t / 48000) + rng.normal(0, 50, n)
    x[n // 2:n // 2 + 192] = 0.0          # one silent block: exercises the ALC noise gate
    mic = np.zeros((n, 2), np.int16)
    mic[:, 0] = np.rint(x)
    mic[:, 1] = mic[:, 0] // 3
    return mic


def main():
    assert pyoracle.have_fw_tx(), "build oracle/_ref/fw_tx first (make -C oracle ref)"
    mic = make_mic(20261018, 192 * 8)
    out = {"mic": mic}
    for c in CASES:
        s = {**DEFAULTS, **c["settings"]}
        w, f = pyoracle.run_fw_tx(mic, s)
        out[c["name"] + "/words"] = w
        out[c["name"] + "/float"] = f
        print("%-22s rms I %.1f Q %.1f" % (c["name"], f[192 * 3:, 0].std(), f[192 * 3:, 1].std()))
    out["meta"] = np.frombuffer(json.dumps({"cases": [dict(name=c["name"], settings={**DEFAULTS, **c["settings"]}) for c in CASES],
                                            "generator": "tools/gen_golden_tx.py",
                                            "source": "oracle/_ref/fw_tx (reference firmware C, host-built)"}).encode(), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "tx_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
