#!/usr/bin/env python3
"""Generates tests/golden/rx_cases.npz: input I/Q frames + settings and the outputs of the REFERENCE
FIRMWARE's own processRxAudio()/FFT_doFFT() (host-built from /root/reference by oracle/ref_harness,
binary oracle/_ref/fw_rx).  The fixtures travel to the GPU box, where the reference tree does not exist.

Run:  make -C oracle ref && python tools/gen_golden_rx.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

CASES = [
    dict(name="usb_default", settings=dict(mode=1)),
    dict(name="lsb_dnr", settings=dict(mode=0, dnr=1)),
    dict(name="lsb_notch_hpf100", settings=dict(mode=0, notch=1, notch_fc=1900, ssb_hpf_pass=100, filter_width=3000)),
    dict(name="cw_u_500", settings=dict(mode=4, filter_width=500)),
    dict(name="cw_l_300_dnr", settings=dict(mode=3, filter_width=300, dnr=1)),
    dict(name="digi_u_agc_forced_off", settings=dict(mode=6, filter_width=3000)),
    dict(name="am_6k_notch", settings=dict(mode=10, filter_width=6000, notch=1, notch_fc=1000)),
    dict(name="nfm_15k", settings=dict(mode=8, filter_width=15000)),
    dict(name="wfm_15k_sql0", settings=dict(mode=9, filter_width=15000, fm_sql_threshold=0)),
    dict(name="iq_passthrough", settings=dict(mode=2)),
    dict(name="usb_iqswap_agcoff_vol100", settings=dict(mode=1, iq_swap=1, agc=0, volume=100, rf_gain=20)),
    dict(name="usb_mute_fftavg1", settings=dict(mode=1, mute=1, fft_averaging=1)),
    dict(name="lsb_lpf_off_agcfast", settings=dict(mode=0, filter_width=0, agc_speed=10)),
    dict(name="usb_fft_off", settings=dict(mode=1, fft_enabled=0)),
    dict(name="cw_u_decoder", settings=dict(mode=4, filter_width=500, cw_decoder=1)),
    dict(name="usb_strong_rf_gain", settings=dict(mode=1, rf_gain=250, fft_averaging=2)),
    dict(name="usb_zoom2", settings=dict(mode=1, fft_zoom=2)),
    dict(name="lsb_zoom8_notch", settings=dict(mode=0, fft_zoom=8, notch=1)),
    dict(name="am_zoom16", settings=dict(mode=10, filter_width=6000, fft_zoom=16)),
    # retune seen by the first FFT_printFFT(): FFT_moveWaterfall() rotates the averages in place (fft.c:347-351,458-504)
    dict(name="usb_retune_up", settings=dict(mode=1), retune_hz=3000),
    dict(name="lsb_zoom2_retune_down", settings=dict(mode=0, fft_zoom=2), retune_hz=-4000),
]
DEFAULTS = dict(mode=0, agc=1, agc_speed=3, dnr=0, notch=0, mute=0, volume=20, rf_gain=50, fm_sql_threshold=1, fft_enabled=1,
                fft_averaging=4, fft_zoom=1, iq_swap=0, cw_decoder=0, filter_width=2700, ssb_hpf_pass=300, notch_fc=1000)


def make_frames(seed, n_frames):
    """I/Q frames from the golden DDC over a synthetic ADC stream: two SSB-like tones around the NCO
    frequency (+1.0 kHz, -1.9 kHz), an AM carrier with 400 Hz modulation at +6 kHz offset and noise."""
    fs = 49152000.0
    n = 1024 * n_frames
    t = np.arange(n, dtype=np.float64)
    f0 = 605867 * fs / 2 ** 22
    rng = np.random.default_rng(seed)
    x = 600 * np.cos(2 * np.pi * (f0 + 1000.0) / fs * t) + 250 * np.cos(2 * np.pi * (f0 - 1900.0) / fs * t)
    x += 200 * (1 + 0.5 * np.cos(2 * np.pi * 400.0 / fs * t)) * np.cos(2 * np.pi * (f0 + 6000.0) / fs * t)
    x += rng.normal(0, 8.0, n)
    adc = np.clip(np.rint(x), -2048, 2047).astype(np.int16)
    return pyoracle.GoldenDDC(605867).push(adc)


def main():
    assert pyoracle.have_fw_rx(), "build oracle/_ref/fw_rx first (make -C oracle ref)"
    n_frames = 192 * 7 + 1            # 7 audio blocks, 2 FFT frames
    frames = make_frames(20261018, n_frames)
    out = {"frames": frames, "meta": np.frombuffer(json.dumps(
        {"cases": [dict(name=c["name"], settings={**DEFAULTS, **c["settings"]}, retune_hz=c.get("retune_hz", 0)) for c in CASES],
         "n_frames": n_frames, "generator": "tools/gen_golden_rx.py", "source": "oracle/_ref/fw_rx (reference firmware C, host-built)"}
    ).encode(), dtype=np.uint8)}
    for c in CASES:
        s = {**DEFAULTS, **c["settings"]}
        # the VFO frequency is a uint32: a downward retune is the wrapped difference, as in Freq - currentFFTFreq
        ev = ((100, "freq", c["retune_hz"] & 0xFFFFFFFF),) if c.get("retune_hz") else ()
        r = pyoracle.run_fw_rx(frames, s, events=ev)
        for k in ("audio", "smeter", "cw", "usb", "spectra", "waterfall", "fft_max", "wtf_history"):
            out[c["name"] + "/" + k] = r[k]
        print("%-28s audio %s spectra %s  rms L %.1f" % (c["name"], r["audio"].shape, r["spectra"].shape,
                                                         r["audio"][:, 0::2].astype(float).std()))
    path = os.path.join(ROOT, "tests", "golden", "rx_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
