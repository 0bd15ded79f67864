"""Transmit DUC on the GPU against the golden model: bit-exact 14-bit DAC words."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_duc_bit_exact_with_state_carry(pkg, oracle):
    rng = np.random.default_rng(11)
    n_ch = 37
    fcw = rng.integers(1, 1 << 21, n_ch).astype(np.uint32)
    fcw[:3] = [605867, 1, (1 << 22) - 1]
    rx = pkg.Receiver(n_ch, 1 << 14)
    rx.set_fcw(fcw)
    rx.duc_enable(16)
    gold = [oracle.GoldenDUC(w) for w in fcw]
    for push, n in enumerate([7, 16, 1, 0, 5]):
        iq = rng.integers(-32768, 32768, (n_ch, n, 2)).astype(np.int16)
        if push == 1:
            iq[:, :, 0] = 32767
            iq[::2, :, 1] = -32768
        rx.duc_push(iq)
        got = rx.duc_read_dac()
        assert got.shape == (n_ch, n * 1024)
        for c in range(n_ch):
            ref, _ = gold[c].push(iq[c, :, 0], iq[c, :, 1])
            assert np.array_equal(got[c], ref), "push %d channel %d" % (push, c)
    assert not rx.duc_read_otr().any()
    with pytest.raises(pkg.UA3Error):
        rx.duc_push(np.zeros((n_ch, 17, 2), np.int16))
    rx.close()


def test_duc_then_ddc_loopback(pkg, oracle):
    """TX DUC -> DAC words -> (12-bit ADC) -> RX DDC on the same tuning word recovers the baseband tone:
    an end-to-end property that needs no oracle (TRX_MODE_LOOPBACK of the firmware does this over the air)."""
    n = 64
    t = np.arange(n)
    i = np.rint(14000 * np.cos(2 * np.pi * 1500 * t / 48000)).astype(np.int16)
    q = np.rint(14000 * np.sin(2 * np.pi * 1500 * t / 48000)).astype(np.int16)
    rx = pkg.Receiver(1, 1 << 16)
    rx.set_fcw([605867])
    rx.duc_enable(n)
    rx.duc_push(np.stack([i, q], axis=1)[None])
    dac = rx.duc_read_dac()[0]
    adc = ((dac.astype(np.int32) - 8191) >> 2).astype(np.int16)      # 14-bit DAC word -> 12-bit ADC range
    rx.push(adc)
    iq = pkg.frames_to_iq(rx.read_frames()[0])
    rx.close()
    z = iq["spec_i"][24:].astype(np.float64) + 1j * iq["spec_q"][24:].astype(np.float64)
    sp = np.abs(np.fft.fft(z * np.hanning(z.size)))
    k = int(np.argmax(sp))
    f = (k if k < z.size / 2 else k - z.size) * 48000.0 / z.size
    assert abs(abs(f) - 1500.0) < 1300.0 and np.abs(z).mean() > 300


def test_duc_wire_order_entry(pkg, oracle):
    """ua3reo_duc_push_wire takes the command-3 bytes (Q hi, Q lo, I hi, I lo: fpga.c:403-436, stm32_interface.v:206-227)."""
    rng = np.random.default_rng(12)
    n_ch, n = 5, 9
    fcw = rng.integers(1, 1 << 21, n_ch).astype(np.uint32)
    iq = rng.integers(-32768, 32768, (n_ch, n, 2)).astype(np.int16)
    iq[0, 0] = [-32768, 32767]
    wire = np.zeros((n_ch, n, 4), np.uint8)
    u = iq.view(np.uint16)
    wire[..., 0], wire[..., 1] = u[..., 1] >> 8, u[..., 1] & 0xFF          # Q first
    wire[..., 2], wire[..., 3] = u[..., 0] >> 8, u[..., 0] & 0xFF
    out = []
    for use_wire in (False, True):
        rx = pkg.Receiver(n_ch, 1 << 14)
        rx.set_fcw(fcw)
        rx.duc_enable(16)
        rx.duc_push_wire(wire) if use_wire else rx.duc_push(iq)
        out.append(rx.duc_read_dac())
        rx.close()
    assert np.array_equal(out[0], out[1])
    ref, _ = oracle.GoldenDUC(fcw[0]).push(iq[0, :, 0], iq[0, :, 1])
    assert np.array_equal(out[1][0], ref)


def test_duc_wire_entry_equals_the_executed_bus(pkg, oracle):
    """tests/golden/bus_cases.npz: the command-3 bytes as the firmware's own FPGA_fpgadata_sendiq() put them on the bus and the
    TX_I / TX_Q that the reference's stm32_interface.v (executed from its source text) latched from them.  The product's
    wire-order entry fed those bytes must produce the DAC stream of the golden DUC fed the latched words."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bus_cases.npz"))
    wire, latched = g["tx_wire"][:32], g["tx_latched"][:32]
    assert np.array_equal(latched, g["tx_iq"][:32])
    fcw = np.array([620407, 1234567], np.uint32)
    rx = pkg.Receiver(2, 1 << 14)
    rx.set_fcw(fcw)
    rx.duc_enable(32)
    rx.duc_push_wire(np.broadcast_to(wire, (2,) + wire.shape).copy())
    dac = rx.duc_read_dac()
    rx.close()
    for c in range(2):
        ref, _ = oracle.GoldenDUC(int(fcw[c])).push(latched[:, 0], latched[:, 1])
        assert np.array_equal(dac[c], ref)
