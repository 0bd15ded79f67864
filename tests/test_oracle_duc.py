"""Known answers that pin the golden transmit DUC (oracle/duc_golden.c)."""
import ctypes

import numpy as np


class TXCIC(ctypes.Structure):
    _fields_ = [("cnt", ctypes.c_uint32), ("wreg", ctypes.c_int16), ("d", ctypes.c_int64 * 5), ("up", ctypes.c_int64),
                ("i", ctypes.c_int64 * 5), ("out14", ctypes.c_int16)]


def test_tx_cic_dc_gain_and_rate(oracle):
    """A constant s16 input w settles to w/4 at the 14-bit output (gain 512^4 * 2^43 >> 46 with the pruning shifts)."""
    L = oracle.lib()
    L.ua3g_tx_cic_clock.restype = ctypes.c_int16
    for w in (16000, -16000, 32767, -32768, 4):
        c = TXCIC()
        L.ua3g_tx_cic_reset(ctypes.byref(c))
        outs = [L.ua3g_tx_cic_clock(ctypes.byref(c), ctypes.c_int16(w)) for _ in range(512 * 10)]
        assert outs[-1] == w >> 2 and outs[-600] == w >> 2
        assert outs[0] == 0


def test_tx_ciccomp_impulse_is_the_coefficient_table(oracle):
    import os, re
    from conftest import ROOT
    txt = open(os.path.join(ROOT, "oracle", "tables", "ddc_tables.h")).read()
    def arr(name):
        m = re.search(name + r"\[\d+\] = \{([^}]*)\}", txt)
        return [int(t) for t in m.group(1).replace("\n", "").split(",") if t.strip()]
    c1, c2 = arr("UA3_TXCOMP_C1"), arr("UA3_TXCOMP_C2")
    L = oracle.lib()
    st = ctypes.create_string_buffer(64)
    L.ua3g_tx_ciccomp_reset(st)
    z = (ctypes.c_int16 * 2)()
    got = []
    for n in range(24):
        L.ua3g_tx_ciccomp_push(st, ctypes.c_int16(16384 if n == 0 else 0), z)
        got += [z[0], z[1]]
    want = []
    for i in range(24):
        want += [c1[i], c2[i]]          # 16384 * c >> 14 = c exactly
    assert got == want


def test_dac_word_and_overflow(oracle):
    L = oracle.lib()
    L.ua3g_dac_word.restype = ctypes.c_uint16
    ov = ctypes.c_int()
    assert L.ua3g_dac_word(0, 0, 8190, 8190, ctypes.byref(ov)) == 8191 and ov.value == 0       # mid scale
    assert L.ua3g_dac_word(8191, 0, 8190, 0, ctypes.byref(ov)) == ((8191 * 8190) >> 14) + 8191
    assert L.ua3g_dac_word(-8192, 0, 8190, 0, ctypes.byref(ov)) == ((-8192 * 8190) >> 14) + 8191
    L.ua3g_dac_word(-8192, -8192, -8192, -8192, ctypes.byref(ov))
    assert ov.value == 1                                                                       # 2^27 does not fit s28


def test_duc_tone_spectrum(oracle):
    """A complex baseband tone at +1 kHz comes out as a real RF tone next to the NCO frequency."""
    n = 96
    t = np.arange(n)
    i = np.rint(12000 * np.cos(2 * np.pi * 1000 * t / 48000)).astype(np.int16)
    q = np.rint(12000 * np.sin(2 * np.pi * 1000 * t / 48000)).astype(np.int16)
    dac, otr = oracle.GoldenDUC(605867).push(i, q)
    assert otr.sum() == 0 and dac.max() < 16384
    x = dac[20 * 1024:].astype(np.float64) - 8191
    sp = np.abs(np.fft.rfft(x * np.hanning(x.size)))
    f = np.argmax(sp) * 49152000.0 / x.size
    f0 = 605867 * 49152000.0 / 2 ** 22
    assert abs(abs(f - f0) - 1000.0) < 700.0          # bin width ~630 Hz at this length
    assert 1500 < np.abs(x).max() < 4000


def test_cos_path_is_sin_path_quarter_turn_ahead(oracle):
    """The DUC kernel's Q lane evaluates the sine path at phase + 2^20: must equal the cosine path for 14 bits."""
    L = oracle.lib()
    rng = np.random.default_rng(5)
    for p in list(rng.integers(0, 1 << 22, 3000)) + [0, 1, (1 << 20) - 1, 1 << 20, (1 << 22) - 1, 3 << 20]:
        s, c, s2, c2 = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        L.ua3g_nco(int(p), ctypes.byref(s), ctypes.byref(c))
        L.ua3g_nco(int((p + (1 << 20)) & 0x3FFFFF), ctypes.byref(s2), ctypes.byref(c2))
        assert c.value == s2.value
