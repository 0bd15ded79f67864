"""processRxAudio()/FFT_doFFT() on the GPU against the reference firmware's own C code:
(a) the committed golden fixtures (tests/golden/rx_cases.npz, produced by oracle/_ref/fw_rx from the
    reference sources; generator tools/gen_golden_rx.py), every mode the firmware demodulates;
(b) when the host-built firmware binary travelled with the snapshot, live runs on fresh random input.
Tolerance (BASELINE.json north_star): max |delta| / peak <= 1e-5 and SNR >= 120 dB per channel."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
TOL_REL, TOL_SNR = 1e-5, 120.0


def stats(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    peak = max(np.abs(ref).max(), 1e-30)
    den = ((got - ref) ** 2).sum()
    snr = 10 * np.log10(max((ref ** 2).sum(), 1e-300) / den) if den > 0 else np.inf
    return np.abs(got - ref).max() / peak, snr


def check(got, ref, what):
    if not np.any(ref):
        assert not np.any(got), what + ": reference is all zero"
        return
    err, snr = stats(got, ref)
    # int32 audio: one LSB of truncation noise is allowed on top of the relative tolerance
    lsb = 1.0 / max(np.abs(np.asarray(ref, np.float64)).max(), 1.0) if np.issubdtype(np.asarray(ref).dtype, np.integer) else 0.0
    assert err <= TOL_REL + lsb, "%s: max|d|/peak = %.3e" % (what, err)
    assert snr >= TOL_SNR or err <= lsb, "%s: SNR = %.1f dB" % (what, snr)


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "rx_cases.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def _run_frames_through_gpu(pkg, oracle, cases, n_adc, pushes, seed, retune_hz=None):
    """Frames come from the GPU DDC itself (all channels share tuning word and ADC stream)."""
    fs = 49152000.0
    t = np.arange(n_adc, dtype=np.float64)
    f0 = 605867 * fs / 2 ** 22
    rng = np.random.default_rng(seed)
    x = 600 * np.cos(2 * np.pi * (f0 + 1000.0) / fs * t) + 250 * np.cos(2 * np.pi * (f0 - 1900.0) / fs * t)
    x += 200 * (1 + 0.5 * np.cos(2 * np.pi * 400.0 / fs * t)) * np.cos(2 * np.pi * (f0 + 6000.0) / fs * t)
    x += rng.normal(0, 8.0, n_adc)
    adc = np.clip(np.rint(x), -2048, 2047).astype(np.int16)
    rx = pkg.Receiver(len(cases), 1 << 21)
    rx.set_fcw([605867] * len(cases))
    rx.rx_enable(True)
    rx.rx_set([rx.rx_defaults(**c) for c in cases])
    if retune_hz is not None and any(retune_hz):
        rx.move_waterfall(np.asarray(retune_hz, np.int32))     # seen by the first FFT frame's display pass
    frames, audio, spec, off = [], [], [], 0
    extra = {}
    for p in pushes:
        rx.push(adc[off:off + p]); off += p
        frames.append(rx.read_frames()); audio.append(rx.read_audio()); spec.append(rx.read_spectra())
        extra.setdefault("waterfall", []).append(rx.read_waterfall())
        extra.setdefault("cw", []).append(rx.read_cw())
        extra.setdefault("usb", []).append(rx.read_audio_usb())
    sm = rx.read_smeter()
    hist = rx.read_waterfall_history()
    rx.close()
    _run_frames_through_gpu.extra = {k: np.concatenate(v, 1) for k, v in extra.items()}
    _run_frames_through_gpu.extra["wtf_history"] = hist
    return np.concatenate(frames, 1), np.concatenate(audio, 1), np.concatenate(spec, 1), sm


def test_against_reference_firmware_fixtures(pkg, oracle, golden):
    z, meta = golden
    cases = meta["cases"]
    n_frames = meta["n_frames"]
    # same ADC recipe as tools/gen_golden_rx.py -> the GPU DDC must reproduce the fixture's frames bit-exactly
    keys = ("mode", "agc", "agc_speed", "dnr", "notch", "mute", "volume", "rf_gain", "fm_sql_threshold", "fft_enabled",
            "fft_averaging", "fft_zoom", "iq_swap", "cw_decoder", "filter_width", "ssb_hpf_pass", "notch_fc")
    frames, audio, spec, sm = _run_frames_through_gpu(
        pkg, oracle, [{k: c["settings"][k] for k in keys} for c in cases], 1024 * n_frames, [1024 * n_frames], 20261018,
        retune_hz=[c.get("retune_hz", 0) for c in cases])
    assert np.array_equal(frames[0], z["frames"]), "GPU DDC frames differ from the fixture's golden frames"
    exact = 0
    for i, c in enumerate(cases):
        ra, rs = z[c["name"] + "/audio"], z[c["name"] + "/spectra"]
        assert audio.shape[1] == ra.shape[0] == 7
        check(audio[i], ra, c["name"] + " audio")
        exact += int(np.array_equal(audio[i], ra))
        if c["settings"]["fft_enabled"]:
            assert spec.shape[1] == rs.shape[0] == 2
            check(spec[i], rs, c["name"] + " spectrum")
        rsm = z[c["name"] + "/smeter"][-1]
        assert np.allclose(sm[i], rsm, rtol=1e-5, atol=1e-3), c["name"] + " s-meter"
        # next rows of SURVEY.md 8(f): waterfall colour rows (display half of fft.c) and the CW Goertzel front end
        if c["settings"]["fft_enabled"]:
            wf, rwf = _run_frames_through_gpu.extra["waterfall"][i], z[c["name"] + "/waterfall"]
            assert np.array_equal(wf, rwf), c["name"] + " waterfall row"          # index work: bit-exact
            # the 50-row history ring (wtf_buffer, fft.c:29,353-358) incl. the sideways move of a retune (:458-504)
            hist, rhist = _run_frames_through_gpu.extra["wtf_history"][i], z[c["name"] + "/wtf_history"]
            assert hist.shape == rhist.shape == (50, 256)
            assert np.array_equal(hist, rhist), c["name"] + " waterfall history"
            assert np.array_equal(hist[0], wf[-1]) and not hist[2:].any()
        usb, rusb = _run_frames_through_gpu.extra["usb"][i], z[c["name"] + "/usb"]
        if np.array_equal(audio[i], ra):                               # exact int32 audio in -> exact int16 packing out
            assert np.array_equal(usb, rusb), c["name"] + " USB audio packing"
        else:
            assert np.abs(usb.astype(np.int32) - rusb.astype(np.int32)).max() <= 1, c["name"] + " USB audio packing"
        cwm, rcw = _run_frames_through_gpu.extra["cw"][i], z[c["name"] + "/cw"]
        assert np.allclose(cwm, rcw, rtol=1e-5, atol=1e-4), c["name"] + " CW Goertzel magnitude"
    # SSB/CW/DIGI/IQ/AM use only IEEE +,-,*,/,sqrt: those channels must be bit-identical; NFM/WFM go through atan2f, the one
    # libm call where libdevice and glibc differ by an ulp
    for i, c in enumerate(cases):
        if c["settings"]["mode"] not in (8, 9):
            assert np.array_equal(audio[i], z[c["name"] + "/audio"]), c["name"] + ": audio not bit-identical"
        if c["settings"]["fft_enabled"]:
            assert np.array_equal(spec[i], z[c["name"] + "/spectra"]), c["name"] + ": spectrum not bit-identical"


def test_state_carry_across_pushes(pkg, oracle, golden):
    """Splitting the ADC stream into ragged pushes changes nothing (ring buffer + per-channel state)."""
    z, meta = golden
    cases = [dict(mode=1, dnr=1, notch=1, notch_fc=1500), dict(mode=0), dict(mode=8, filter_width=15000), dict(mode=10, filter_width=6000)]
    n = 1024 * (192 * 9 + 5)
    _, a1, s1, _ = _run_frames_through_gpu(pkg, oracle, cases, n, [n], 7)
    _, a2, s2, _ = _run_frames_through_gpu(pkg, oracle, cases, n, [1024 * 700 + 3, 1024 * 11, 1021, n - (1024 * 711 + 3 + 1021)], 7)
    assert a1.shape[1] == 9 and s1.shape[1] == 3
    assert np.array_equal(a1, a2) and np.array_equal(s1, s2)


def test_live_against_host_built_firmware(pkg, oracle):
    if not oracle.have_fw_rx():
        pytest.skip("oracle/_ref/fw_rx did not travel with this snapshot")
    cases = [dict(mode=m, filter_width=w, dnr=d, notch=nt, notch_fc=fc, agc_speed=sp, rf_gain=g, iq_swap=sw)
             for m, w, d, nt, fc, sp, g, sw in [
                 (0, 2700, 0, 0, 1000, 3, 50, 0), (1, 3400, 1, 1, 700, 5, 30, 0), (3, 500, 1, 0, 1000, 3, 50, 1),
                 (4, 300, 0, 1, 600, 1, 80, 0), (5, 2900, 0, 0, 1000, 3, 50, 0), (6, 1800, 1, 0, 1000, 7, 10, 1),
                 (10, 10000, 0, 1, 2000, 3, 50, 0), (8, 9000, 0, 0, 1000, 3, 50, 0), (9, 15000, 1, 0, 1000, 3, 50, 0),
                 (2, 2700, 0, 0, 1000, 3, 50, 1), (1, 0, 0, 0, 1000, 3, 50, 0), (0, 5000, 0, 0, 1000, 9, 100, 0)]]
    for i, zoom in enumerate([1, 2, 4, 8, 16, 1, 2, 4, 8, 16, 1, 2]):
        cases[i]["fft_zoom"] = zoom
    n = 1024 * (192 * 12 + 1)
    frames, audio, spec, sm = _run_frames_through_gpu(pkg, oracle, cases, n, [n // 2 + 17, n - (n // 2 + 17)], 99)
    rx0 = pkg.Receiver(1, 1024)
    for i, c in enumerate(cases):
        s = rx0.rx_defaults(**c).as_dict()
        ref = oracle.run_fw_rx(frames[i], s)
        nb = min(audio.shape[1], ref["audio"].shape[0])
        assert nb == 12
        check(audio[i, :nb], ref["audio"][:nb], "live %s audio" % c)
        nf = min(spec.shape[1], ref["spectra"].shape[0])
        assert nf >= 4
        check(spec[i, :nf], ref["spectra"][:nf], "live %s spectrum" % c)
    rx0.close()


def test_every_filter_table_against_host_built_firmware(pkg, oracle):
    """All 32 lattice LPF tables (audio_filters.c:59-122, selected by Filter_Width, :141-303) and the five SSB HPF corners
    (:305-333), each on its own channel, against live runs of the reference firmware on that channel's frames."""
    if not oracle.have_fw_rx():
        pytest.skip("oracle/_ref/fw_rx did not travel with this snapshot")
    widths = [300, 500, 1400, 1600, 1800, 2100, 2300, 2500, 2700, 2900, 3000, 3200, 3400, 3600, 3800, 4000,
              4200, 4400, 4600, 4800, 5000, 5500, 6000, 6500, 7000, 7500, 8000, 8500, 9000, 9500, 10000, 15000]
    cases = []
    for i, w in enumerate(widths):
        mode = 3 if w <= 500 else (i % 2 if w <= 3400 else (10 if i % 2 else 8))      # CW_L, LSB/USB, AM/NFM
        cases.append(dict(mode=mode, filter_width=w))
    for hp in (100, 200, 300, 400, 500):
        cases.append(dict(mode=1, filter_width=3000, ssb_hpf_pass=hp))
    n = 1024 * (192 * 6 + 1)
    frames, audio, spec, sm = _run_frames_through_gpu(pkg, oracle, cases, n, [n], 4242)
    rx0 = pkg.Receiver(1, 1024)
    exact = 0
    for i, c in enumerate(cases):
        ref = oracle.run_fw_rx(frames[i], rx0.rx_defaults(**c).as_dict())
        assert ref["audio"].shape[0] == audio.shape[1] == 6
        check(audio[i], ref["audio"], "filter sweep %s audio" % c)
        exact += int(np.array_equal(audio[i], ref["audio"]))
    rx0.close()
    assert exact >= len(cases) - 8, "only %d of %d filter cases bit-exact" % (exact, len(cases))     # FM (atan2f) may differ by an ulp


def test_pipelined_stage_and_async_reads(pkg, oracle):
    """The STM32 stage runs on its own stream one push behind the DDC; results read with the *_async calls while the next
    push is already running must equal the synchronous reads of an identical receiver."""
    import torch
    cases = [dict(mode=1, dnr=1), dict(mode=0, notch=1, notch_fc=1300), dict(mode=10, filter_width=6000), dict(mode=8, filter_width=15000),
             dict(mode=4, filter_width=500, cw_decoder=1), dict(mode=1, fft_zoom=4)] * 20          # 120 channels: two rx_audio CTAs
    block = 1024 * 192 * 2
    sizes = [block, block, 1024 * 700 + 5, 1019, block - 1024 * 300, 1024 * 64, block, 1024 * 513 + 1000]      # ragged pushes too
    adc = oracle.synth_adc(sum(sizes), seed=21)
    fcw = [605867 + 997 * i for i in range(len(cases))]
    offs = np.concatenate([[0], np.cumsum(sizes)])

    def make():
        rx = pkg.Receiver(len(cases), 1 << 20)
        rx.set_fcw(fcw)
        rx.rx_enable(True)
        rx.rx_set([rx.rx_defaults(**c) for c in cases])
        return rx

    ref = make()
    want_a, want_s = [], []
    for b in range(len(sizes)):
        ref.push(adc[offs[b]:offs[b + 1]])
        want_a.append(ref.read_audio()); want_s.append(ref.read_spectra())
    ref.close()

    rx = make()
    dev = torch.from_numpy(adc).cuda()
    got_a = [torch.empty((len(cases) * 6 * 384,), dtype=torch.int32).pin_memory() for _ in sizes]
    got_s = [torch.empty((len(cases) * 3 * 256,), dtype=torch.float32).pin_memory() for _ in sizes]
    counts = []
    for b in range(len(sizes)):                  # no host synchronisation inside the loop
        rx.push(dev[offs[b]:offs[b + 1]])
        na = rx.read_audio_async(got_a[b]); ns = rx.read_spectra_async(got_s[b])
        counts.append((na, ns))
    rx.sync()
    rx.close()
    assert sum(c[0] for c in counts) == sum(sizes) // 1024 // 192 and sum(c[1] for c in counts) == sum(sizes) // 1024 // 512
    for b in range(len(sizes)):
        na, ns = counts[b]
        assert (na, ns) == (want_a[b].shape[1], want_s[b].shape[1])
        if na:
            assert np.array_equal(got_a[b].numpy()[:len(cases) * na * 384].reshape(len(cases), na, 384), want_a[b]), "push %d audio" % b
        if ns:
            assert np.array_equal(got_s[b].numpy()[:len(cases) * ns * 256].reshape(len(cases), ns, 256), want_s[b]), "push %d spectra" % b


def test_many_small_pushes_wrap_the_ring(pkg, oracle):
    """80 pushes of 2^16 samples into a receiver whose frame ring holds 2048 frames: the ring wraps more than twice while the
    STM32 stage trails one push behind and every result leaves through the pipelined reads.  Frames must equal the golden
    DDC and the audio must equal a receiver that got the same stream in three big pushes."""
    import torch
    n_push, block = 80, 1 << 16
    cases = [dict(mode=1, dnr=1), dict(mode=0, notch=1), dict(mode=10, filter_width=6000), dict(mode=8, filter_width=15000)] * 10
    adc = oracle.synth_adc(n_push * block, seed=77)
    fcw = [605867 + 1009 * i for i in range(len(cases))]

    small = pkg.Receiver(len(cases), block)
    small.set_fcw(fcw); small.rx_enable(True); small.rx_set([small.rx_defaults(**c) for c in cases])
    dev = torch.from_numpy(adc.reshape(n_push, block)).cuda()
    fr = [torch.empty((len(cases), block // 1024, 8), dtype=torch.uint8).pin_memory() for _ in range(n_push)]
    au = [torch.empty((len(cases) * 2 * 384,), dtype=torch.int32).pin_memory() for _ in range(n_push)]
    na = []
    for b in range(n_push):
        small.push(dev[b])
        small.read_frames_async(fr[b])
        na.append(small.read_audio_async(au[b]))
    small.sync()
    small.close()
    frames = np.concatenate([f.numpy() for f in fr], 1)
    audio = np.concatenate([a.numpy()[:len(cases) * n * 384].reshape(len(cases), n, 384) for a, n in zip(au, na) if n], 1)

    for ch in (0, 17, 39):
        assert np.array_equal(frames[ch], oracle.golden_frames(adc, [fcw[ch]])[0]), "frames of channel %d" % ch
    big = pkg.Receiver(len(cases), 1 << 21)
    big.set_fcw(fcw); big.rx_enable(True); big.rx_set([big.rx_defaults(**c) for c in cases])
    ref = []
    cuts = [0, 2097152, 4194304, adc.size]
    for a, b in zip(cuts[:-1], cuts[1:]):
        big.push(adc[a:b]); ref.append(big.read_audio())
    big.close()
    ref = np.concatenate(ref, 1)
    assert audio.shape == ref.shape and np.array_equal(audio, ref)


def test_cw_decoder_end_to_end(pkg, oracle):
    """Keyed carrier -> STM32 stage in CW_U with the decoder on (device Goertzel front end per 192-sample block)
    -> host state machine (ua3reo_cw_decoder_step): the Morse text comes out, as the firmware shows it in its text bar."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_golden_cw as g
    audio = g.keyed_audio("CQ CQ DE UA3REO UA3REO K", 20, amp=1.0, noise=0.0)
    n = audio.size
    # recover the on/off envelope (the generator's tone is sin(2 pi 350 t) * env) and key a +350 Hz complex exponential with it
    t = np.arange(n) / 48000.0
    key = np.zeros(n)
    blk = 48                                                   # 1 ms resolution is plenty for 60 ms dits
    a = np.abs(audio).reshape(-1, blk).max(1) > 0.1
    key = np.repeat(a, blk).astype(np.float64)
    rng = np.random.default_rng(5)
    z = 6000.0 * key * np.exp(2j * np.pi * 350.0 * t) + rng.normal(0, 20, n) + 1j * rng.normal(0, 20, n)
    words = np.stack([np.rint(z.imag), np.rint(z.real), np.rint(z.imag), np.rint(z.real)], 1).astype(np.int16)   # SPEC_Q, SPEC_I, VOICE_Q, VOICE_I
    frames = words.astype(">i2").view(np.uint8).reshape(n, 8)
    rx = pkg.Receiver(2, 1 << 22)
    rx.rx_enable(True)
    rx.rx_set([rx.rx_defaults(mode=4, filter_width=500, cw_decoder=1, agc=1), rx.rx_defaults(mode=1)])
    dec = pkg.CwDecoder(pkg.load_library())
    text = ""
    step = 192 * 20
    for off in range(0, n - step + 1, step):
        rx.rx_push_frames(np.repeat(frames[None, off:off + step], 2, 0))
        cw = rx.read_cw()
        assert not cw[1].any()                                  # the decoder only runs in CW modes with TRX.CWDecoder set
        for m in cw[0]:
            text += dec.step(float(m))
    rx.close()
    assert "UA3REO UA3REO" in text, text          # (the first letters train the adaptive thresholds; AGC pumps the noise up in the tail)
    assert 12 <= dec.wpm <= 26, dec.wpm


def test_rejects_settings_without_firmware_table(pkg):
    rx = pkg.Receiver(2, 1024)
    rx.rx_enable(True)
    for bad in (dict(filter_width=2750), dict(mode=12), dict(agc_speed=0), dict(fft_zoom=3), dict(fft_averaging=0)):
        with pytest.raises(pkg.UA3Error):
            rx.rx_set(rx.rx_defaults(**bad))
    rx.rx_set(rx.rx_defaults(mode=1, ssb_hpf_pass=60))     # no branch in ReinitAudioFilters: keeps the previous HPF
    rx.close()


def test_full_chain_many_channels_consistency(pkg, oracle):
    """BASELINE config 5 shape at reduced length: 512 channels, modes round-robin; channels that share
    tuning word and settings must agree bit for bit, whatever lane pair / warp they land in."""
    n_ch = 512
    modes = [0, 1, 4, 10, 8]
    widths = {0: 2700, 1: 2700, 4: 500, 10: 6000, 8: 15000}
    rx = pkg.Receiver(n_ch, 1 << 19)
    fcw = np.where(np.arange(n_ch) % 2 == 0, 605867, 620407).astype(np.uint32)
    rx.set_fcw(fcw)
    rx.rx_enable(True)
    sets = []
    for c in range(n_ch):
        m = modes[(c // 2) % 5]
        sets.append(rx.rx_defaults(mode=m, filter_width=widths[m], dnr=(c // 10) % 2, notch=(c // 10) % 2))
    rx.rx_set(sets)
    adc = oracle.synth_adc(1 << 19, seed=5)
    rx.push(adc)
    a, s = rx.read_audio(), rx.read_spectra()
    rx.close()
    assert a.shape == (n_ch, 2, 384) and s.shape == (n_ch, 1, 256)
    for c in range(20, n_ch):
        assert np.array_equal(a[c], a[c - 20]) and np.array_equal(s[c], s[c - 20])   # period lcm(10,20)=20 in settings, 2 in fcw


def test_stm32_stage_alone_on_external_frames(pkg, golden):
    """BASELINE config 1 shape: the STM32 stage fed with I/Q frames directly (no DDC), ragged pushes; equals the
    reference firmware fixtures exactly like the DDC-fed run."""
    z, meta = golden
    frames = z["frames"]
    cases = [c for c in meta["cases"] if c["name"] in ("usb_default", "lsb_dnr", "am_6k_notch", "nfm_15k", "usb_zoom2")]
    keys = ("mode", "agc", "agc_speed", "dnr", "notch", "mute", "volume", "rf_gain", "fm_sql_threshold", "fft_enabled",
            "fft_averaging", "fft_zoom", "iq_swap", "cw_decoder", "filter_width", "ssb_hpf_pass", "notch_fc")
    rx = pkg.Receiver(len(cases), 1 << 20)
    rx.rx_enable(True)
    rx.rx_set([rx.rx_defaults(**{k: c["settings"][k] for k in keys}) for c in cases])
    audio, spec = [], []
    for a, b in [(0, 100), (100, 700), (700, 1023), (1023, frames.shape[0])]:
        rx.rx_push_frames(np.repeat(frames[None, a:b], len(cases), 0))
        audio.append(rx.read_audio()); spec.append(rx.read_spectra())
    rx.close()
    audio, spec = np.concatenate(audio, 1), np.concatenate(spec, 1)
    for i, c in enumerate(cases):
        check(audio[i], z[c["name"] + "/audio"], c["name"] + " audio (external frames)")
        check(spec[i], z[c["name"] + "/spectra"], c["name"] + " spectrum (external frames)")
