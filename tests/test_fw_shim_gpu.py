"""The firmware-name shim (ua3reo-ddc-transceiver_b200/host/ua3reo_fw_shim.c) as a drop-in: the SAME firmware driver
(oracle/ref_harness/fw_rx.c / fw_tx.c around the reference's own fpga.c bus driver and functions.c) is built twice -
once with the reference's DSP translation units (oracle/_ref/fw_rx, all CPU) and once with those units replaced by
the shim over libua3reo_b200.so (oracle/_ref/fw_rx_b200, numbers from the GPU) - and both run over the same frames.
Every processRxAudio()/processTxAudio()/FFT_doFFT() result must agree: bit-exact int32 audio / int16 USB audio /
RGB565 rows, floats within the north_star tolerance (1e-5 relative, 120 dB SNR).
Also covers the settings semantics the shim relies on: values the firmware reads on every call change without
clearing state (ua3reo_rx_set_live), ReinitAudioFilters()/InitNotchFilter()/FFT_Init() do what the firmware's do."""
import os

import numpy as np
import pytest

from conftest import ROOT
from test_rx_gpu import check

pytestmark = pytest.mark.gpu

BASE = dict(mode=0, filter_width=2700, ssb_hpf_pass=300, rf_gain=50, agc=1, agc_speed=3, dnr=0, notch=0, notch_fc=1000,
            volume=20, mute=0, fm_sql_threshold=1, fft_enabled=1, fft_zoom=1, fft_averaging=4, iq_swap=0, cw_decoder=0)


def _need(oracle):
    for b in (oracle.FW_RX, oracle.FW_RX_B200, oracle.FW_TX, oracle.FW_TX_B200):
        if not os.path.exists(b):
            pytest.skip("host-built firmware harness %s did not travel with the snapshot" % os.path.basename(b))


def _shim_env(pkg, tmp_path_factory):
    """Normally the shim binary finds lib/libua3reo_b200.so through its RUNPATH; the emulation aid swaps the library."""
    env = dict(os.environ)
    if os.path.basename(pkg.LIB_PATH) != "libua3reo_b200.so":
        d = tmp_path_factory.mktemp("emulib")
        os.symlink(pkg.LIB_PATH, os.path.join(d, "libua3reo_b200.so"))
        env["LD_LIBRARY_PATH"] = str(d)
    return env


def _frames(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    w = np.zeros((n, 4))
    for k in range(4):
        w[:, k] = 2500 * np.sin(2 * np.pi * (0.011 + 0.004 * k) * t + k) + 900 * np.sin(2 * np.pi * 0.19 * t) + rng.normal(0, 250, n)
    w *= (1.0 + 0.6 * np.sin(2 * np.pi * t / 4096.0))[:, None]
    return np.clip(np.rint(w), -32768, 32767).astype(np.int16).astype(">i2").view(np.uint8).reshape(n, 8)


def _compare(ref, got, what):
    for k in ("audio", "usb", "waterfall", "wtf_history"):
        assert ref[k].shape == got[k].shape, "%s: %s shape" % (what, k)
        if k in ("waterfall", "wtf_history"):
            assert np.array_equal(ref[k], got[k]), "%s: waterfall rows differ" % what          # index work: bit-exact
        else:
            check(got[k], ref[k], "%s: %s" % (what, k))
    for k in ("smeter", "cw", "spectra"):
        assert ref[k].shape == got[k].shape, "%s: %s shape" % (what, k)
        check(got[k], ref[k], "%s: %s" % (what, k))


RX_CASES = [
    ("lsb", dict(mode=0), ()),
    ("usb_notch_dnr", dict(mode=1, notch=1, dnr=1, notch_fc=1400), ()),
    ("am_noagc", dict(mode=10, agc=0, filter_width=6000), ()),
    ("nfm", dict(mode=8, filter_width=15000), ()),
    ("cw_decoder", dict(mode=3, cw_decoder=1, filter_width=500), ()),
    ("iq_swap_zoom4", dict(mode=2, iq_swap=1, fft_zoom=4, filter_width=0), ()),
    # values the firmware reads on every call, changed mid-stream with no re-initialisation
    ("live_changes", dict(mode=0), ((700, "volume", 55), (900, "agc", 0), (1300, "mode", 1), (1500, "notch", 1),
                                    (1700, "rf_gain", 30), (2100, "dnr", 1), (2500, "mute", 1), (2700, "mute", 0),
                                    (2900, "fft_averaging", 2))),
    # TRX_setMode(): new mode and filter, then ReinitAudioFilters(); the notch corner moves and InitNotchFilter() follows later
    ("reinit", dict(mode=0, notch=1), ((1000, "mode", 10), (1000, "filter_width", 6000), (1000, "reinit", 0),
                                       (1600, "notch_fc", 2200), (2000, "notch_init", 0),
                                       (2400, "fft_zoom", 2), (2400, "fft_init", 0))),
    # TRX.Agc_speed only counts from the next InitAGC() (agc.c:14-19); FFT_Init() clears the ZoomFFT filter states even when the
    # zoom is unchanged (fft.c:189-207)
    ("agc_init_and_fft_init", dict(mode=0, fft_zoom=2), ((600, "agc_speed", 9), (1200, "agc_init", 0), (1500, "fft_init", 0),
                                                         (2200, "agc_speed", 1), (2600, "fft_zoom", 4), (2600, "fft_init", 0))),
    # the housekeeping tick zeroes the S-meter extremes after showing them (stm32f4xx_it.c:398-409)
    ("smeter_reset", dict(mode=1), ((800, "smeter_reset", 0), (2000, "smeter_reset", 0), (2001, "rf_gain", 20))),
    # retunes as FFT_printFFT() sees them: rows and averages move sideways (FFT_moveWaterfall, fft.c:458-504)
    ("retune_up_down", dict(mode=0), ((1100, "freq", 3000), (2300, "freq", 1000), (3300, "freq", 1100))),
    ("retune_zoom2_far", dict(mode=0, fft_zoom=2), ((1500, "freq", 20000), (2800, "freq", 2000))),
]


@pytest.mark.parametrize("name,over,events", RX_CASES, ids=[c[0] for c in RX_CASES])
def test_rx_firmware_driver_cpu_vs_gpu_shim(pkg, oracle, tmp_path_factory, name, over, events):
    _need(oracle)
    s = dict(BASE); s.update(over)
    fr = _frames(192 * 20, 77)
    ref = oracle.run_fw_rx(fr, s, events=events)
    got = oracle.run_fw_rx(fr, s, events=events, binary=oracle.FW_RX_B200, env=_shim_env(pkg, tmp_path_factory))
    assert ref["audio"].shape[0] == 19 and ref["spectra"].shape[0] >= 6
    _compare(ref, got, name)


TX_BASE = dict(mode=1, filter_width=2700, ssb_hpf_pass=300, rf_power=20, mute=0, tune=0, key_down=0)
TX_CASES = [("usb", dict(mode=1)), ("lsb_1k8", dict(mode=0, filter_width=1800)), ("am", dict(mode=10, filter_width=6000)),
            ("nfm", dict(mode=8, filter_width=8000)), ("cw_key", dict(mode=3, key_down=1)), ("tune", dict(mode=1, tune=1))]


@pytest.mark.parametrize("name,over", TX_CASES, ids=[c[0] for c in TX_CASES])
def test_tx_firmware_driver_cpu_vs_gpu_shim(pkg, oracle, tmp_path_factory, name, over):
    _need(oracle)
    s = dict(TX_BASE); s.update(over)
    rng = np.random.default_rng(5)
    t = np.arange(192 * 12)
    mic = np.stack([6000 * np.sin(2 * np.pi * 0.02 * t) + rng.normal(0, 500, t.size),
                    3000 * np.sin(2 * np.pi * 0.031 * t) + rng.normal(0, 500, t.size)], 1).astype(np.int16)
    ref_w, ref_f = oracle.run_fw_tx(mic, s)
    got_w, got_f = oracle.run_fw_tx(mic, s, binary=oracle.FW_TX_B200, env=_shim_env(pkg, tmp_path_factory))
    assert np.array_equal(ref_w, got_w), name + ": I/Q words on the wire differ"
    check(got_f, ref_f, name + ": FPGA_Audio_SendBuffer")


@pytest.mark.parametrize("iq_swap,freq", [(0, 7100000), (1, 30000000)])
def test_fpga_bus_driver_cpu_vs_gpu_shim(pkg, oracle, tmp_path, tmp_path_factory, iq_swap, freq):
    """The FPGA half of the binding (host/ua3reo_fpga_shim.c): FPGA_Init / FPGA_fpgadata_stuffclock / FPGA_fpgadata_iqclock
    over ua3reo_ddc_push.  The same driver (oracle/ref_harness/fw_fpga.c) runs (a) with the reference's own fpga.c, fed the
    golden DDC's frames byte by byte over the bus stub, and (b) with the shim, fed the ADC samples: rings, FFT input
    buffers, ring index, FFT_buff_index and NeedFFTInputBuffer must agree after EVERY tick and at the end."""
    ref_bin, shim_bin = os.path.join(ROOT, "oracle", "_ref", "fw_fpga"), os.path.join(ROOT, "oracle", "_ref", "fw_fpga_b200")
    for b in (ref_bin, shim_bin):
        if not os.path.exists(b):
            pytest.skip("host-built firmware harness %s did not travel with the snapshot" % os.path.basename(b))
    import subprocess
    n_ticks = int(os.environ.get("UA3_TEST_TICKS", "2000"))
    fcw, _ = pkg.phrase_from_frequency(freq)             # what FPGA_fpgadata_sendparam() puts on the bus (fpga.c:173-220)
    t = np.arange(1024 * n_ticks, dtype=np.float64)
    tone = 600 * np.cos(2 * np.pi * (fcw / 2.0 ** 22 + 1500.0 / 49152000.0) * t)          # 1.5 kHz above the tuned frequency
    adc = np.clip(oracle.synth_adc(1024 * n_ticks, seed=3 + iq_swap).astype(np.float64) * 0.5 + np.rint(tone), -2048, 2047).astype(np.int16)
    frames = oracle.GoldenDDC(int(fcw)).push(adc)
    assert frames.shape == (n_ticks, 8)
    fa, ff = tmp_path / "adc.bin", tmp_path / "frames.bin"
    adc.tofile(fa); frames.tofile(ff)
    outs = []
    for binary, inp, env in ((ref_bin, ff, None), (shim_bin, fa, _shim_env(pkg, tmp_path_factory))):
        out = tmp_path / (os.path.basename(binary) + ".out")
        subprocess.check_call([binary, str(iq_swap), str(freq), str(inp), str(out)], env=env)
        outs.append(np.fromfile(out, dtype=np.uint8))
    assert outs[0].size == outs[1].size == n_ticks * 12 + 4 * 384 * 4 + 2 * 512 * 4 + 8
    per_tick = [o[:n_ticks * 12].view(np.uint32).reshape(n_ticks, 3) for o in outs]
    assert np.array_equal(per_tick[0], per_tick[1]), "ring index / FFT_buff_index / NeedFFTInputBuffer per tick"
    assert per_tick[0][:, 2].min() == 0 and per_tick[0][:, 1].max() == 511          # the FFT buffer filled and was re-armed
    tail = [o[n_ticks * 12:] for o in outs]
    assert np.array_equal(tail[0], tail[1]), "rings / FFTInput_I/Q / FPGA_samples / underrun flag"
    rings = tail[1][:4 * 384 * 4].view(np.float32)
    assert np.abs(rings).max() > 100                                                   # real signal went through


@pytest.mark.parametrize("agc,speed,dnr,mode", [(1, 3, 1, 1), (1, 1, 0, 0), (0, 3, 1, 3), (1, 5, 1, 5)])
def test_stage_functions_cpu_vs_gpu_shim(pkg, tmp_path, tmp_path_factory, agc, speed, dnr, mode):
    """dc_filter() / DoAGC() / processNoiseReduction() (audio_filters.h:51, agc.h:9, noise_reduction.h:16) called on their own:
    the reference's translation units against host/ua3reo_fw_shim.c over ua3reo_rx_stage, same driver (fw_stage.c), 40 blocks
    with the gain ramp, the clip branch and the NLMS weights in motion.  IEEE +,-,*,/ only: bit-exact."""
    import subprocess
    ref_bin, shim_bin = os.path.join(ROOT, "oracle", "_ref", "fw_stage"), os.path.join(ROOT, "oracle", "_ref", "fw_stage_b200")
    for b in (ref_bin, shim_bin):
        if not os.path.exists(b):
            pytest.skip("host-built firmware harness %s did not travel with the snapshot" % os.path.basename(b))
    rng = np.random.default_rng(agc * 8 + dnr * 4 + mode)
    n = 192 * 40
    t = np.arange(n)
    x = (2000 * np.sin(2 * np.pi * 0.02 * t) * (1 + 0.9 * np.sin(2 * np.pi * t / 2000.0)) + rng.normal(0, 150, n) + 300).astype(np.float32)
    x[192 * 20:192 * 21] *= 40.0                                   # a burst: the clip branch of DoAGC (agc.c:41-45)
    fin = tmp_path / "in.f32"
    x.tofile(fin)
    outs = []
    for binary, env in ((ref_bin, None), (shim_bin, _shim_env(pkg, tmp_path_factory))):
        out = tmp_path / (os.path.basename(binary) + ".out")
        subprocess.check_call([binary, str(agc), str(speed), str(dnr), str(mode), str(fin), str(out)], env=env)
        outs.append(np.fromfile(out, dtype=np.float32).reshape(-1, 2, 192))
    assert outs[0].shape == outs[1].shape == (40, 2, 192)
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32)), "stage outputs differ (max |d| %g)" % np.abs(outs[0] - outs[1]).max()
    assert np.abs(outs[0][:, 0]).max() > 1 and not np.array_equal(outs[0][:, 0], outs[0][:, 1])
