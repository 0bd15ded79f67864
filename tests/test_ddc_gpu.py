"""Parity of the CUDA DDC (through the C ABI) with the golden model: bit-exact frames."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fcws(n, seed):
    return np.random.default_rng(seed).integers(1, 1 << 21, n).astype(np.uint32)


def _run(pkg, oracle, n_ch, pushes, max_block, seed, fcw=None):
    fcw = _fcws(n_ch, seed) if fcw is None else np.asarray(fcw, np.uint32)
    adc = oracle.synth_adc(int(sum(pushes)), seed=seed)
    adc[:7] = -2048                               # reach the (-2048)*(-2048) mixer wrap when sin12 = -2048
    rx = pkg.Receiver(n_ch, max_block)
    rx.set_fcw(fcw)
    got, off = [], 0
    for n in pushes:
        rx.push(adc[off:off + n]); off += n
        got.append(rx.read_frames())
    rx.close()
    got = np.concatenate(got, axis=1)
    ref = oracle.golden_frames(adc, fcw)[:, :got.shape[1]]
    return got, ref


def test_single_channel_full_ddc(pkg, oracle):
    """BASELINE config 2: one channel, default tuning word, bit-exact 48 kSPS I/Q frames."""
    got, ref = _run(pkg, oracle, 1, [1 << 19], 1 << 19, seed=11, fcw=[605867])
    assert got.shape == (1, 512, 8)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n_ch", [1, 3, 32, 33, 100])
def test_channel_counts_and_padding(pkg, oracle, n_ch):
    got, ref = _run(pkg, oracle, n_ch, [1 << 16], 1 << 16, seed=n_ch)
    assert np.array_equal(got, ref)


def test_state_carry_across_blocks(pkg, oracle):
    """Blocks of different sizes, including ragged ones that leave a remainder, equal one long push."""
    pushes = [1024 * 300, 1024 * 7 + 13, 1011, 1024 * 64, 2048, 1024 * 130 + 1, 1023]
    got, ref = _run(pkg, oracle, 5, pushes, 1 << 19, seed=5)
    assert got.shape[1] == sum(pushes) // 1024
    assert np.array_equal(got, ref)


def test_empty_and_tiny_pushes(pkg, oracle):
    adc = oracle.synth_adc(4096, seed=9)
    rx = pkg.Receiver(2, 4096)
    rx.set_fcw([1, (1 << 22) - 1])
    assert rx.push(adc[:0]) == 0
    assert rx.read_frames().shape == (2, 0, 8)
    assert rx.push(adc[:1]) == 0
    assert rx.push(adc[1:1024]) == 1
    f0 = rx.read_frames()
    assert rx.push(adc[1024:4096]) == 3
    f1 = rx.read_frames()
    ref = oracle.golden_frames(adc, [1, (1 << 22) - 1])
    assert np.array_equal(np.concatenate([f0, f1], axis=1), ref)
    with pytest.raises(pkg.UA3Error):
        rx.push(np.zeros(4096 + 1024, np.int16))
    rx.close()


def test_extreme_tuning_words_and_full_scale(pkg, oracle):
    """FCW 0 (DC), 1, 2^21, 2^22-1 with a full-scale square wave: exercises every wrap in the chain."""
    n = 1 << 17
    adc = np.where((np.arange(n) // 3) % 2 == 0, 2047, -2048).astype(np.int16)
    fcw = np.array([0, 1, 1 << 21, (1 << 22) - 1, 605867, 620407], np.uint32)
    rx = pkg.Receiver(len(fcw), n)
    rx.set_fcw(fcw)
    rx.push(adc)
    got = rx.read_frames()
    rx.close()
    assert np.array_equal(got, oracle.golden_frames(adc, fcw))


def test_reset_and_retune(pkg, oracle):
    adc = oracle.synth_adc(1 << 16, seed=21)
    rx = pkg.Receiver(4, 1 << 16)
    f1, f2 = _fcws(4, 1), _fcws(4, 2)
    rx.set_fcw(f1); rx.push(adc); a = rx.read_frames()
    rx.reset(); rx.set_fcw(f2); rx.push(adc); b = rx.read_frames()
    rx.reset(); rx.set_fcw(f1); rx.push(adc); c = rx.read_frames()
    rx.close()
    assert np.array_equal(a, c)
    assert np.array_equal(b, oracle.golden_frames(adc, f2))


def test_device_resident_input_and_frames_view(pkg, oracle):
    import torch
    adc = oracle.synth_adc(1 << 16, seed=31)
    fcw = _fcws(64, 31)
    rx = pkg.Receiver(64, 1 << 16)
    rx.set_fcw(fcw)
    d = torch.from_numpy(adc).cuda()
    assert rx.push(d) == 64
    got = rx.read_frames()
    base, first, nf, ring, stride = rx.frames_device()
    assert nf == 64 and first == 0 and ring >= 64 + 1024 and ring & (ring - 1) == 0 and stride == ring * 8 and base
    rx.close()
    assert np.array_equal(got, oracle.golden_frames(adc, fcw))


def test_full_size_block_properties(pkg, oracle):
    """BASELINE config 3 shape (1024 channels x 2^20 samples): too big for the scalar oracle on every
    channel, so check (a) 16 sampled channels bit-exactly, (b) channels sharing a tuning word agree,
    (c) a checksum of all frames is independent of how the stream is split into blocks."""
    n = 1 << 20
    adc = oracle.synth_adc(n, seed=20261018)
    fcw = _fcws(1024, 20261018)
    fcw[512:] = fcw[:512]                                   # duplicates
    rx = pkg.Receiver(1024, n)
    rx.set_fcw(fcw)
    rx.push(adc)
    whole = rx.read_frames()
    assert whole.shape == (1024, 1024, 8)
    assert np.array_equal(whole[:512], whole[512:])
    pick = np.random.default_rng(1).choice(512, 16, replace=False)
    assert np.array_equal(whole[pick], oracle.golden_frames(adc, fcw[pick]))
    rx.reset()
    parts = []
    for a, b in [(0, 1 << 18), (1 << 18, (1 << 18) + 5000), ((1 << 18) + 5000, n)]:
        if rx.push(adc[a:b]):
            parts.append(rx.read_frames())
    rx.close()
    assert np.array_equal(np.concatenate(parts, axis=1), whole)


@pytest.mark.parametrize("n_ch", [1024, 8192])
def test_baseline_size_every_channel_against_golden(pkg, oracle, n_ch):
    """BASELINE configs[2] / [3] at their full size - 1024 and 8192 channels x 2^20 ADC samples - with EVERY channel's frames
    compared bit for bit against the golden model (all host cores run it: about 2 s and 16 s of CPU work on 16 cores)."""
    n = 1 << 20
    adc = oracle.synth_adc(n, seed=20261018)
    fcw = pkg.random_fcw(n_ch, seed=20261018 + n_ch)
    fcw[:4] = [1, (1 << 22) - 1, 1 << 21, 620407]                     # the extremes ride along
    rx = pkg.Receiver(n_ch, n)
    rx.set_fcw(fcw)
    assert rx.push(adc) == n // 1024
    got = rx.read_frames()
    rx.close()
    _, ref = oracle.GoldenBank(fcw).push(adc, os.cpu_count() or 8, want_frames=True)
    bad = np.flatnonzero((got != ref).reshape(n_ch, -1).any(axis=1))
    assert bad.size == 0, "channels with differing frames: %s" % bad[:10]
    assert got.reshape(n_ch, -1).any(axis=1).all()                    # no channel was left silent


def test_front_kernel_variants_agree(pkg, oracle, monkeypatch):
    """The 8 KB-table kernel and the 208 KB big-table kernel (TMA-staged, one CTA per SM) are both exact:
    same frames for 300 channels (not a multiple of the 256-channel CTA tile) over ragged pushes."""
    n_ch, pushes = 300, [1024 * 37, 1024 * 5 + 7, 1017, 1024 * 64]
    adc = oracle.synth_adc(sum(pushes), seed=77)
    adc[100:110] = -2048
    fcw = _fcws(n_ch, 77)
    outs = {}
    for variant in ("1", "2"):
        monkeypatch.setenv("UA3REO_FRONT_VARIANT", variant)
        rx = pkg.Receiver(n_ch, 1 << 16)
        rx.set_fcw(fcw)
        got, off = [], 0
        for n in pushes:
            rx.push(adc[off:off + n]); off += n
            got.append(rx.read_frames())
        rx.close()
        outs[variant] = np.concatenate(got, axis=1)
    assert np.array_equal(outs["1"], outs["2"])
    pick = [0, 1, 31, 32, 255, 256, 299]
    assert np.array_equal(outs["2"][pick], oracle.golden_frames(adc, fcw[pick])[:, :outs["2"].shape[1]])


def test_tensor_core_front_kernel_is_exact(pkg, oracle, monkeypatch):
    """ddc_front_tc_kernel (tcgen05 int8 MMAs for the integrators, FP32-pipe mixer) against the CUDA-core big-table
    kernel and the golden model: 300 channels (a partly filled 128-channel warpgroup), ragged pushes, tuning words
    that park the NCO on sin = -1 / cos = -1, and ADC samples of -2048 - the one product, (-2048) * (-2048) = 2^22,
    that wraps the 23-bit mixer register - both in flagged chunks and next to unflagged ones."""
    n_ch, pushes = 300, [1024 * 37, 1024 * 5 + 7, 1017, 1024 * 64]
    adc = oracle.synth_adc(sum(pushes), seed=79)
    adc[100:110] = -2048
    adc[5000] = -2048
    adc[1024 * 40:1024 * 41] = -2048
    adc[1024 * 50:1024 * 50 + 512] = 2047
    fcw = _fcws(n_ch, 79)
    fcw[0] = 0                      # phase stays 0: cos = +max, sin = 0
    fcw[1] = 1 << 20                # 0, 1/4, 1/2, 3/4 of a turn: sin and cos visit -1
    fcw[2] = 1 << 21                # 0, 1/2 turn
    fcw[3] = 3 << 20
    fcw[299] = (1 << 22) - 1
    outs = {}
    for variant in ("2", "3"):
        monkeypatch.setenv("UA3REO_FRONT_VARIANT", variant)
        rx = pkg.Receiver(n_ch, 1 << 16)
        rx.set_fcw(fcw)
        got, off = [], 0
        for n in pushes:
            rx.push(adc[off:off + n]); off += n
            got.append(rx.read_frames())
        rx.close()
        outs[variant] = np.concatenate(got, axis=1)
    assert np.array_equal(outs["2"], outs["3"])
    pick = [0, 1, 2, 3, 31, 127, 128, 255, 256, 299]
    assert np.array_equal(outs["3"][pick], oracle.golden_frames(adc, fcw[pick])[:, :outs["3"].shape[1]])


def test_tensor_core_front_kernel_small_and_odd_banks(pkg, oracle, monkeypatch):
    """Forced onto banks it is not the default for: 1, 33 and 129 channels (idle lanes in the MMA's 128 rows, a second
    warpgroup tile with one live channel), a 512-sample push (one chunk) and a long one."""
    monkeypatch.setenv("UA3REO_FRONT_VARIANT", "3")
    for n_ch in (1, 33, 129):
        n = 1024 * 48
        adc = oracle.synth_adc(n, seed=n_ch)
        adc[::97] = -2048
        fcw = _fcws(n_ch, n_ch)
        rx = pkg.Receiver(n_ch, 1 << 15)
        rx.set_fcw(fcw)
        got = []
        for a, b in [(0, 1024), (1024, 1024 * 30), (1024 * 30, n)]:
            rx.push(adc[a:b])
            got.append(rx.read_frames())
        rx.close()
        assert np.array_equal(np.concatenate(got, axis=1), oracle.golden_frames(adc, fcw)), n_ch


def test_async_frame_reads_overlap_pushes(pkg, oracle):
    """ua3reo_ddc_read_frames_async: results of push k are copied while push k+1 runs; pinned host buffers."""
    import torch
    n_ch, block, n_blocks = 64, 1 << 15, 7
    adc = oracle.synth_adc(block * n_blocks, seed=123)
    fcw = _fcws(n_ch, 123)
    rx = pkg.Receiver(n_ch, block)
    rx.set_fcw(fcw)
    host_adc = torch.from_numpy(adc.reshape(n_blocks, block)).pin_memory()
    outs = [torch.empty((n_ch, block // 1024, 8), dtype=torch.uint8).pin_memory() for _ in range(n_blocks)]
    for b in range(n_blocks):
        rx.push(host_adc[b])
        rx.read_frames_async(outs[b])
    rx.sync()
    got = np.concatenate([o.numpy() for o in outs], axis=1)
    rx.close()
    assert np.array_equal(got, oracle.golden_frames(adc, fcw))


def test_adc_min_max_tracking(pkg, oracle):
    """ADC_MIN / ADC_MAX of stm32_interface.v:384-397: +2000 / -2000 after a reset, running extremes afterwards."""
    rx = pkg.Receiver(4, 1 << 15)
    assert rx.adc_stats() == (2000, -2000, 0)            # first call arms the tracker with the FPGA's reset values
    adc = oracle.synth_adc(1 << 15, seed=3)
    adc[100] = 2047
    adc[200] = -2048
    rx.push(adc[:20000])
    rx.push(adc[20000:])
    n_proc = (adc.size // 1024) * 1024                   # only whole frames have been through the chain
    seen = adc[:n_proc]
    mn, mx, rail = rx.adc_stats(reset=True)
    assert (mn, mx) == (int(seen.min()), int(seen.max()))
    assert rail == int(((seen <= -2048) | (seen >= 2047)).sum())
    assert rx.adc_stats() == (2000, -2000, 0)
    small = np.full(2048, 5, np.int16)
    rx.push(small)
    assert rx.adc_stats() == (5, 5, 0)                   # min(2000, 5), max(-2000, 5)
    rx.close()


def test_get_params_packet(pkg, oracle):
    """Command 2 of the wire protocol (stm32_interface.v:172-205) and its decode by FPGA_fpgadata_getparam() (fpga.c:222-284)."""
    rx = pkg.Receiver(2, 1 << 14)
    rx.adc_stats()                                        # arm
    adc = np.clip(oracle.synth_adc(1 << 14, seed=9) // 4, -400, 700).astype(np.int16)
    rx.push(adc)
    pkt, mn, mx = rx.get_params()
    lo, hi = int(adc.min()), int(adc.max())
    assert lo < 0 < hi
    assert pkt[0] == 0                                    # no sample at a rail, no DAC overflow, no keys
    assert pkt[1] == (((lo & 0xFFF) >> 8) << 4 | ((hi & 0xFFF) >> 8)) and pkt[2] == (lo & 0xFF) and pkt[3] == (hi & 0xFF) and pkt[4] == (hi & 0xE0)
    assert (mn, mx) == (lo, hi)
    assert rx.adc_stats() == (2000, -2000, 0)             # the read reset the extremes (ADC_MINMAX_RESET, k == 203)
    rx.push(np.full(1 << 12, -7, np.int16))
    pkt, mn, mx = rx.get_params(dac_otr=True)
    assert pkt[0] == 2 and mn == -7
    assert mx == ((-7) & 0xFFF)                           # firmware quirk: the maximum is not sign-extended (fpga.c:270)
    rail = np.zeros(1 << 12, np.int16); rail[5] = 2047
    rx.push(rail)
    pkt, mn, mx = rx.get_params()
    assert pkt[0] == 1 and (mn, mx) == (0, 2047)          # ADC_OTR
    rx.close()


def test_get_params_equals_the_executed_bus(pkg):
    """ua3reo_get_params against tests/golden/bus_cases.npz: the five bytes the firmware's own FPGA_fpgadata_getparam() read
    from the reference's stm32_interface.v (executed from its source, tools/gen_golden_bus.py) for the same ADC samples, and
    the two amplitudes it decoded."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bus_cases.npz"))
    rx = pkg.Receiver(2, 1 << 12)
    for adc, (otr, dac), packet, decoded in zip(g["gp_adc"], g["gp_flags"], g["gp_packet"], g["gp_decoded"]):
        rx.get_params()                                   # the read before: ADC_MINMAX_RESET
        rx.push(adc)
        pkt, mn, mx = rx.get_params(dac_otr=bool(dac))
        assert np.array_equal(pkt, packet), (pkt, packet)
        assert (mn, mx) == (int(np.int16(decoded[0])), int(np.int16(decoded[1])))
        assert (pkt[0] & 1) == otr
    rx.close()


def test_c_host_driver_runs(pkg):
    """The C host program over the C ABI (host/ua3reo_rx_host.c, built by build.sh): full chain for a small bank."""
    import subprocess
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "ua3reo_rx_host")
    if os.path.basename(pkg.LIB_PATH) != "libua3reo_b200.so" or not os.path.exists(exe):
        pytest.skip("C host driver not built beside the library")
    out = subprocess.run([exe, "40", "3", str(1 << 17), "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "channels" in out.stdout and "checksum" in out.stdout.lower()
