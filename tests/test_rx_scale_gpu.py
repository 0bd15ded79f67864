"""BASELINE-size and long-run parity of the STM32 stage against the reference firmware's own C (oracle/_ref/fw_rx).

  * configs[4] at full size: 4096 channels x 2^20 ADC samples through DDC + processRxAudio + FFT_doFFT; 64 sampled
    channels (every mode / DNR / notch class, first and last warps) are compared with the firmware run on that channel's
    own frames, and the frames of those channels with the golden DDC model.
  * configs[0] long run: 60 s of 48 kSPS I/Q (15 000 audio blocks) through ua3reo_rx_push_frames - USB + DNR + AGC and NFM -
    whole-run SNR and worst error; AGC and NLMS are feedback loops, a drift would show here.
  * which outputs are bit-identical: SSB/CW/AM/IQ audio, spectra, waterfall rows and the USB int16 packing are IEEE
    +,-,*,/,sqrt only and must be EXACT; NFM/WFM audio goes through atan2f, where CUDA's libdevice and glibc differ by an
    ulp - the one irreducible libm call, inside the 1e-5 tolerance.
"""
import numpy as np
import pytest

from test_rx_gpu import check, stats

pytestmark = pytest.mark.gpu
MIX = [(0, 2700), (1, 2700), (4, 500), (10, 6000), (8, 15000)]     # bench.py's configs[4] mode mix


def _settings(c):
    return dict(mode=MIX[c % 5][0], filter_width=MIX[c % 5][1], dnr=(c // 5) % 2, notch=(c // 5) % 2)


def test_config5_full_size_sampled_channels_against_firmware(pkg, oracle):
    if not oracle.have_fw_rx():
        pytest.skip("oracle/_ref/fw_rx did not travel with this snapshot")
    n_ch, n = 4096, 1 << 20
    rx = pkg.Receiver(n_ch, n)
    fcw = pkg.random_fcw(n_ch, 20261018)
    rx.set_fcw(fcw)
    rx.rx_enable(True)
    rx.rx_set([rx.rx_defaults(**_settings(c)) for c in range(n_ch)])
    adc = pkg.synth_adc(2 * n, 20261018)
    frames, audio, spec, wf = [], [], [], []
    for b in range(2):
        rx.push(adc[b * n:(b + 1) * n])
        frames.append(rx.read_frames()); audio.append(rx.read_audio()); spec.append(rx.read_spectra()); wf.append(rx.read_waterfall())
    rx.close()
    frames, audio, spec, wf = (np.concatenate(x, 1) for x in (frames, audio, spec, wf))
    assert frames.shape == (n_ch, 2048, 8) and audio.shape[1] == 10 and spec.shape[1] == 4
    rng = np.random.default_rng(5)
    picks = sorted(set(list(range(10)) + list(range(n_ch - 10, n_ch)) + list(rng.integers(0, n_ch, 50))))[:64]
    assert len(picks) >= 60
    rx0 = pkg.Receiver(1, 1024)
    exact_audio = 0
    for c in picks:
        assert np.array_equal(frames[c], oracle.GoldenDDC(int(fcw[c])).push(adc)), "DDC frames of channel %d" % c
        ref = oracle.run_fw_rx(frames[c], rx0.rx_defaults(**_settings(c)).as_dict())
        check(audio[c], ref["audio"][:10], "config 5 channel %d audio" % c)
        assert np.array_equal(spec[c], ref["spectra"][:4]), "config 5 channel %d spectrum not bit-identical" % c
        assert np.array_equal(wf[c], ref["waterfall"][:4]), "config 5 channel %d waterfall rows" % c
        same = np.array_equal(audio[c], ref["audio"][:10])
        exact_audio += int(same)
        if MIX[c % 5][0] != 8:
            assert same, "config 5 channel %d (mode %d): audio not bit-identical" % (c, MIX[c % 5][0])
    rx0.close()
    assert exact_audio >= len(picks) * 4 // 5 - 1


@pytest.mark.parametrize("name,settings", [("usb_dnr_agc", dict(mode=1, dnr=1, agc=1)), ("nfm", dict(mode=8, filter_width=15000))])
def test_sixty_seconds_against_firmware(pkg, oracle, name, settings):
    """>= 60 s of 48 kSPS I/Q; SNR over the whole run."""
    if not oracle.have_fw_rx():
        pytest.skip("oracle/_ref/fw_rx did not travel with this snapshot")
    n = 192 * 15000                                   # 60 s
    rng = np.random.default_rng(77)
    t = np.arange(n) / 48000.0
    fade = 0.55 + 0.45 * np.sin(2 * np.pi * 0.37 * t)             # slow fading keeps the AGC moving for the whole minute
    z = fade * (6000 * np.exp(2j * np.pi * 1000 * t) + 4000 * np.exp(2j * np.pi * (1900 * t + 30 * np.sin(2 * np.pi * 3 * t))))
    z = z + rng.normal(0, 60, n) + 1j * rng.normal(0, 60, n)
    i, q = np.rint(z.real).astype(np.int16), np.rint(z.imag).astype(np.int16)
    frames = np.zeros((n, 8), np.uint8)
    for w, v in ((0, q), (1, i), (2, q), (3, i)):
        frames[:, 2 * w] = (v.view(np.uint16) >> 8).astype(np.uint8)
        frames[:, 2 * w + 1] = (v.view(np.uint16) & 0xFF).astype(np.uint8)
    rx = pkg.Receiver(1, 1 << 20)
    rx.rx_enable(True)
    s = rx.rx_defaults(**settings)
    rx.rx_set(s)
    audio, spec, usb = [], [], []
    step = 192 * 5                                   # a 2^20-sample block holds 1024 frames
    for a in range(0, n, step):
        rx.rx_push_frames(frames[None, a:a + step])
        audio.append(rx.read_audio()); spec.append(rx.read_spectra()); usb.append(rx.read_audio_usb())
    rx.close()
    audio, spec, usb = np.concatenate(audio, 1)[0], np.concatenate(spec, 1)[0], np.concatenate(usb, 1)[0]
    ref = oracle.run_fw_rx(frames, s.as_dict())
    nb = min(audio.shape[0], ref["audio"].shape[0])          # the firmware harness leaves the last block unfinished
    assert nb >= 14999
    audio, ref["audio"] = audio[:nb], ref["audio"][:nb]
    err, snr = stats(audio, ref["audio"])
    print("%s: 60 s, max|d|/peak %.3g, whole-run SNR %s dB, %d of %d blocks bit-identical"
          % (name, err, "inf" if np.isinf(snr) else "%.1f" % snr, int((audio == ref["audio"]).all(axis=1).sum()), nb))
    check(audio, ref["audio"], name + " 60 s audio")
    nf = min(spec.shape[0], ref["spectra"].shape[0])
    assert nf >= 5600
    assert np.array_equal(spec[:nf], ref["spectra"][:nf]), name + ": spectra not bit-identical over 60 s"
    if settings["mode"] != 8:
        assert np.array_equal(audio, ref["audio"]), name + ": audio not bit-identical over 60 s"
        assert np.array_equal(usb[:nb], ref["usb"][:nb]), name + ": USB int16 packing not bit-identical"
