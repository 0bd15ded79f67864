#!/usr/bin/env python3
"""DEVELOPMENT AID: host-emulated CUDA RX audio/FFT kernels vs the host-built reference firmware (oracle/_ref/fw_rx)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ua3reo_loader
from oracle import pyoracle
pkg = ua3reo_loader.load()
EMU = os.path.join(ROOT, "tools", "emu", "_build", "libua3reo_emu.so")

def stats(got, ref):
    got = got.astype(np.float64); ref = ref.astype(np.float64)
    peak = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max() / peak
    den = ((got - ref) ** 2).sum()
    snr = 10 * np.log10((ref ** 2).sum() / den) if den > 0 else float("inf")
    return err, snr

def run(cases, n_frames=192 * 8 + 1, seed=1, pushes=None):
    fs = 49152000.0
    n = 1024 * n_frames
    t = np.arange(n)
    f0 = 605867 * fs / 2 ** 22
    rng = np.random.default_rng(seed)
    adc = np.rint(700 * np.cos(2 * np.pi * (f0 + 1000.0) / fs * t) + 300 * np.cos(2 * np.pi * (f0 - 1900.0) / fs * t)
                  + rng.normal(0, 8, n)).astype(np.int16)
    n_ch = len(cases)
    rx = pkg.Receiver(n_ch, 1 << 21, _lib_path=EMU)
    rx.set_fcw([605867] * n_ch)
    rx.rx_enable(True)
    rx.rx_set([rx.rx_defaults(**c) for c in cases])
    pushes = pushes or [n]
    audio, spec, frames, off = [], [], [], 0
    for p in pushes:
        rx.push(adc[off:off + p]); off += p
        frames.append(rx.read_frames()); audio.append(rx.read_audio()); spec.append(rx.read_spectra())
    frames = np.concatenate(frames, 1); audio = np.concatenate(audio, 1); spec = np.concatenate(spec, 1)
    ok = True
    for c, case in enumerate(cases):
        s = rx.rx_defaults(**case).as_dict()
        ref = pyoracle.run_fw_rx(frames[c], s)
        nb, nf = min(audio.shape[1], ref["audio"].shape[0]), min(spec.shape[1], ref["spectra"].shape[0])
        ea, sa = stats(audio[c, :nb], ref["audio"][:nb])
        ef, sf = stats(spec[c, :nf], ref["spectra"][:nf])
        exact = np.array_equal(audio[c, :nb], ref["audio"][:nb])
        good = ea <= 1e-5 and sa >= 120 and ef <= 1e-5 and sf >= 120
        ok &= good
        print("%-60s blocks=%d fft=%d audio: exact=%s err=%.2e snr=%.1f | fft: err=%.2e snr=%.1f  %s" % (
            case, nb, nf, exact, ea, sa, ef, sf, "OK" if good else "FAIL"))
    return ok

if __name__ == "__main__":
    t = time.time()
    ok = run([dict(mode=1), dict(mode=0, dnr=1), dict(mode=10, filter_width=6000, notch=1), dict(mode=8, filter_width=15000),
              dict(mode=2), dict(mode=3, filter_width=500), dict(mode=9, filter_width=15000), dict(mode=5, agc=0, iq_swap=1)])
    ok &= run([dict(mode=1, dnr=1, notch=1, notch_fc=1500), dict(mode=0)], n_frames=1024 + 300, pushes=[1024 * 700, 1024 * 11, 1024 * 613])
    print("ALL OK" if ok else "FAILED", "%.1fs" % (time.time() - t))
    sys.exit(0 if ok else 1)
