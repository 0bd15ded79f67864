#!/usr/bin/env python3
"""BASELINE configs[0]: the STM32 RX audio chain (processRxAudio + FFT_doFFT) on 48 kSPS synthetic I/Q, USB demodulation +
panorama FFT.  Times (a) the reference firmware's own C, host-built (oracle/_ref/fw_rx, one channel, one core) and
(b) the CUDA stage fed with the same frames for many channels (ua3reo_rx_push_frames).  Supporting measurement."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ua3reo_loader  # noqa: E402
from oracle import pyoracle  # noqa: E402


def synth_iq_frames(seconds, seed=20261018):
    """48 kSPS I/Q: USB two-tone (1.0 kHz + 1.9 kHz above the carrier) + noise, as int16 frames (SPEC = VOICE)."""
    n = int(48000 * seconds)
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 48000.0
    z = 6000 * np.exp(2j * np.pi * 1000 * t) + 4000 * np.exp(2j * np.pi * 1900 * t) + rng.normal(0, 30, n) + 1j * rng.normal(0, 30, n)
    i, q = np.rint(z.real).astype(np.int16), np.rint(z.imag).astype(np.int16)
    f = np.zeros((n, 8), np.uint8)
    for w, v in ((0, q), (1, i), (2, q), (3, i)):
        f[:, 2 * w] = (v.view(np.uint16) >> 8).astype(np.uint8)
        f[:, 2 * w + 1] = (v.view(np.uint16) & 0xFF).astype(np.uint8)
    return f


def main():
    pkg = ua3reo_loader.load()
    n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    frames = synth_iq_frames(seconds)
    out = {"workload": "configs[0]: STM32 RX chain, 48 kSPS synthetic I/Q, USB 2.7 kHz + panorama FFT, %.0f s of signal" % seconds}
    if pyoracle.have_fw_rx():
        t0 = time.time()
        ref = pyoracle.run_fw_rx(frames, dict(mode=1))
        dt = time.time() - t0
        out["cpu_reference"] = {"kind": "reference", "cores": 1, "seconds": dt, "frames_per_s": frames.shape[0] / dt,
                                "x_real_time": frames.shape[0] / dt / 48000.0}
    block = 1024 * 48          # 1.024 s of frames per push
    rx = pkg.Receiver(n_ch, 1024 * block)
    rx.rx_enable(True)
    rx.rx_set(rx.rx_defaults(mode=1))
    tile = np.repeat(frames[None, :block], n_ch, 0)
    rx.rx_push_frames(tile); rx.read_audio()
    t0 = time.time()
    nb = 0
    for p in range(0, frames.shape[0] - block + 1, block):
        rx.rx_push_frames(tile)          # same frames for every channel and push: the arithmetic does not care
        a = rx.read_audio()
        nb += block
    dt = time.time() - t0
    out["gpu"] = {"channels": n_ch, "seconds": dt, "channel_frames_per_s": n_ch * nb / dt, "x_real_time_channels": n_ch * nb / dt / 48000.0,
                  "note": "host-facing call: H2D of the frames + kernels + D2H of the audio, per push"}
    if pyoracle.have_fw_rx():
        g = rx.read_audio()
    rx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
