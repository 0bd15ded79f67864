"""Known-answer identities that pin the golden DDC model (the reference ships no vectors)."""
import ctypes

import numpy as np
import pytest


class CIC(ctypes.Structure):
    _fields_ = [("cnt", ctypes.c_uint32), ("inreg", ctypes.c_int32), ("s", ctypes.c_uint64 * 5),
                ("d", ctypes.c_uint64 * 5), ("outreg", ctypes.c_int16)]


class COMP(ctypes.Structure):
    _fields_ = [("p0", ctypes.c_int16 * 33), ("p1", ctypes.c_int16 * 33), ("n_in", ctypes.c_uint32)]


def _comp_h():
    import os, re
    from conftest import ROOT
    txt = open(os.path.join(ROOT, "oracle", "tables", "ddc_tables.h")).read()
    def arr(name):
        m = re.search(name + r"\[\d+\] = \{([^}]*)\}", txt)
        return np.array([int(t) for t in m.group(1).replace("\n", "").split(",") if t.strip()])
    return arr("UA3_RXCOMP_H"), arr("UA3_RXHILB_C")


@pytest.mark.parametrize("x", [123456, -98765, 4194303, -4194304, 255, -1])
def test_cic_dc_gain(oracle, x):
    """Constant input x -> output 2*(x>>8) once the 5 combs have filled (gain 512^5 = 2^45, slice [59:44])."""
    L = oracle.lib()
    c = CIC()
    L.ua3g_rx_cic_reset(ctypes.byref(c))
    outs = []
    for _ in range(512 * 9):
        if L.ua3g_rx_cic_clock(ctypes.byref(c), x):
            outs.append(c.outreg)
    assert outs[0] == 0 and len(outs) == 9
    want = ((2 * (x >> 8) + 32768) % 65536) - 32768
    assert outs[-1] == want and outs[-2] == want


def test_cic_rate_and_wrap(oracle):
    """512 clocks -> exactly one output; full-scale alternating input never disturbs the count."""
    L = oracle.lib()
    c = CIC()
    L.ua3g_rx_cic_reset(ctypes.byref(c))
    n = sum(L.ua3g_rx_cic_clock(ctypes.byref(c), (-1) ** t * 4194303) for t in range(512 * 20))
    assert n == 20


def test_comp_impulse_response(oracle):
    """Impulse of 2^15-1 through the compensator: outputs follow h (convergent-rounded), polyphase by input parity."""
    L = oracle.lib()
    h, _ = _comp_h()
    for parity in (0, 1):
        c = COMP()
        L.ua3g_rx_ciccomp_reset(ctypes.byref(c))
        y = ctypes.c_int16()
        outs = []
        for m in range(80 + parity):
            u = 16384 if m == parity else 0
            if L.ua3g_rx_ciccomp_push(ctypes.byref(c), u, ctypes.byref(y)):
                outs.append(y.value)
        # y[k] = h[2k+1-parity] * 16384 / 32768 with convergent rounding
        taps = h[(1 - parity)::2]
        want = []
        for t in taps:
            v = int(t) * 16384
            q, r = divmod(v, 32768)
            if r > 16384 or (r == 16384 and (q & 1)):
                q += 1
            want.append(q)
        assert outs[:len(want)] == want


def test_comp_dc_gain(oracle):
    L = oracle.lib()
    c = COMP()
    L.ua3g_rx_ciccomp_reset(ctypes.byref(c))
    y = ctypes.c_int16()
    last = None
    for m in range(200):
        if L.ua3g_rx_ciccomp_push(ctypes.byref(c), 16368, ctypes.byref(y)):
            last = y.value
    assert last == round(16368 * 32700 / 32768)   # 16334, SURVEY.md 8c item 3


def test_hilbert_impulse_and_delay(oracle):
    L = oracle.lib()
    _, hc = _comp_h()
    L.ua3g_rx_hilb_push.restype = ctypes.c_int16
    L.ua3g_delay_push.restype = ctypes.c_int16
    hs = ctypes.create_string_buffer(512)
    ds = ctypes.create_string_buffer(260)
    L.ua3g_rx_hilb_reset(hs)
    L.ua3g_delay_reset(ds)
    outs = [L.ua3g_rx_hilb_push(hs, 16384 if n == 0 else 0) for n in range(256)]
    # product rounding: (c*16384 + bit1) >> 1 = c*8192 exactly; output (acc + 0x1FFF + bit14) >> 14 -> c/2 convergent
    want = []
    for cval in hc:
        acc = int(cval) * 8192
        want.append(((acc & 0x3FFFFFFF) + 0x1FFF + ((acc >> 14) & 1)) >> 14)
    want = [((w + 32768) % 65536) - 32768 if w < 32768 else w - 65536 for w in want]
    want = [((int(cval) * 8192 + 0x1FFF + ((int(cval) * 8192 >> 14) & 1)) >> 14) for cval in hc]
    assert outs == want
    d = [L.ua3g_delay_push(ds, n + 1) for n in range(300)]
    assert d[:130] == [0] * 130 and d[130:] == list(range(1, 171))


def test_frame_byte_order(oracle):
    L = oracle.lib()
    f = (ctypes.c_uint8 * 8)()
    L.ua3g_frame_pack(f, ctypes.c_int16(0x1234), ctypes.c_int16(-2), ctypes.c_int16(0x7FFF), ctypes.c_int16(-32768))
    assert list(f) == [0x12, 0x34, 0xFF, 0xFE, 0x7F, 0xFF, 0x80, 0x00]


def test_phrase_from_frequency(oracle):
    L = oracle.lib()
    s = ctypes.c_int()
    assert L.ua3g_phrase_from_frequency(7100000, ctypes.byref(s)) == 605867 and s.value == 0
    assert L.ua3g_phrase_from_frequency(7270394, ctypes.byref(s)) in (620406, 620407)
    # second Nyquist zone is inverted, third is not
    w = L.ua3g_phrase_from_frequency(30000000, ctypes.byref(s))
    assert s.value == 1 and w == round((49152000 - 30000000) / 49152000 * 4194304)
    w = L.ua3g_phrase_from_frequency(50000000, ctypes.byref(s))
    assert s.value == 0 and w == round((50000000 - 49152000) / 49152000 * 4194304)


def test_ddc_rates_and_state_carry(oracle):
    """1024 ADC samples -> one frame; splitting the stream anywhere gives the identical frames."""
    adc = oracle.synth_adc(1024 * 40, seed=3)
    one = oracle.GoldenDDC(605867).push(adc)
    assert one.shape == (40, 8)
    g = oracle.GoldenDDC(605867)
    parts = [g.push(adc[a:b]) for a, b in [(0, 1), (1, 1000), (1000, 5000), (5000, 1024 * 40)]]
    assert np.array_equal(np.concatenate(parts), one)


def test_ddc_tone_lands_in_passband(oracle):
    """A -6 dBFS tone 1 kHz above the NCO frequency comes out as a 1 kHz complex tone at ~half scale."""
    fs, f0 = 49152000.0, 7100000.0
    n = 1024 * 700
    t = np.arange(n)
    adc = np.rint(1023 * np.cos(2 * np.pi * (f0 + 1000.0) / fs * t)).astype(np.int16)
    fr = oracle.GoldenDDC(605867).push(adc)
    w = (fr[:, 0::2].astype(np.uint16) << 8 | fr[:, 1::2]).astype(np.int16).astype(np.float64)
    z = w[300:, 1] + 1j * w[300:, 0]       # SPEC_I + j SPEC_Q
    spec = np.abs(np.fft.fft(z * np.hanning(len(z))))
    k = int(np.argmax(spec))
    f_est = (k if k < len(z) / 2 else k - len(z)) * 48000.0 / len(z)
    assert abs(abs(f_est) - 1000.0 - (605867 * fs / 2 ** 22 - f0) * 0) < 150.0
    assert 3000 < np.abs(z).mean() < 9000
