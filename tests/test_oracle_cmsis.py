"""The CMSIS-DSP restatement (oracle/cmsis_min.c) against independent references, and the committed golden
fixtures against a fresh run of the host-built reference firmware when it is available."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def L(oracle):
    lib = oracle.lib()
    lib.arm_cos_f32.restype = ctypes.c_float
    lib.arm_cos_f32.argtypes = [ctypes.c_float]
    lib.arm_sin_f32.restype = ctypes.c_float
    lib.arm_sin_f32.argtypes = [ctypes.c_float]
    return lib


class CFFT(ctypes.Structure):
    _fields_ = [("fftLen", ctypes.c_uint16), ("tw", ctypes.c_void_p), ("br", ctypes.c_void_p), ("brl", ctypes.c_uint16)]


def test_cfft_512_matches_numpy(L):
    inst = CFFT.in_dll(L, "arm_cfft_sR_f32_len512")
    rng = np.random.default_rng(0)
    for _ in range(4):
        x = ((rng.normal(size=512) + 1j * rng.normal(size=512)) * 3000).astype(np.complex64)
        buf = np.empty(1024, np.float32)
        buf[0::2], buf[1::2] = x.real, x.imag
        L.arm_cfft_f32(ctypes.byref(inst), buf.ctypes.data_as(ctypes.c_void_p), 0, 1)
        y = buf[0::2].astype(np.float64) + 1j * buf[1::2].astype(np.float64)
        ref = np.fft.fft(x.astype(np.complex128))
        snr = 10 * np.log10((np.abs(ref) ** 2).sum() / (np.abs(y - ref) ** 2).sum())
        assert snr > 130.0


def test_table_cosine_error_profile(L):
    """arm_cos_f32 is a 512-entry table with linear interpolation: ~1.9e-5 worst-case error,
    exact at table nodes, and periodic."""
    xs = np.linspace(-20, 20, 40001, dtype=np.float32)
    err = max(abs(L.arm_cos_f32(float(v)) - np.cos(float(v))) for v in xs[::7])
    assert 1.0e-5 < err < 2.5e-5
    assert L.arm_cos_f32(0.0) == 1.0 and abs(L.arm_sin_f32(0.0)) == 0.0
    assert abs(L.arm_cos_f32(3.0) - L.arm_cos_f32(3.0 + 2 * np.float32(np.pi))) < 2e-5


def test_lattice_matches_direct_form(L):
    """arm_iir_lattice_f32 with the firmware's 2.7 kHz elliptic LPF: the lattice's impulse response must be a
    stable low-pass (DC gain ~1 within the 1 dB ripple, stop band 80 dB down)."""
    import re
    txt = open(os.path.join(ROOT, "oracle", "tables", "audio_tables.h")).read()

    def arr(name):
        m = re.search(name + r"\[\d+\] = \{[^\n]*\n([^}]*)\}", txt)
        return np.array([int(t, 16) for t in re.findall(r"0x([0-9a-f]{8})u", m.group(1))], dtype=np.uint32).view(np.float32)
    widths = [int(t) for t in re.search(r"UA3_LPF_WIDTH\[\d+\] = \{([^}]*)\}", txt).group(1).replace("\n", "").split(",") if t.strip()]
    idx = widths.index(2700)
    pk = arr("UA3_LPF_PK")[idx * 11:idx * 11 + 11].copy()
    pv = arr("UA3_LPF_PV")[idx * 12:idx * 12 + 12].copy()

    class LAT(ctypes.Structure):
        _fields_ = [("numStages", ctypes.c_uint16), ("pState", ctypes.c_void_p), ("pk", ctypes.c_void_p), ("pv", ctypes.c_void_p)]
    S = LAT()
    state = np.zeros(11 + 64, np.float32)
    L.arm_iir_lattice_init_f32(ctypes.byref(S), 11, pk.ctypes.data_as(ctypes.c_void_p), pv.ctypes.data_as(ctypes.c_void_p),
                               state.ctypes.data_as(ctypes.c_void_p), 64)
    h = []
    for b in range(64):
        x = np.zeros(64, np.float32)
        if b == 0:
            x[0] = 1.0
        L.arm_iir_lattice_f32(ctypes.byref(S), x.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), 64)
        h.append(x.copy())
    h = np.concatenate(h).astype(np.float64)
    H = np.abs(np.fft.rfft(h, 1 << 16))
    f = np.arange(H.size) * 48000.0 / (1 << 16)
    assert 0.85 < H[0] < 1.05
    assert H[(f > 200) & (f < 2600)].min() > 0.85
    assert H[f > 4000].max() < 3e-4


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "tests", "golden", "rx_cases.npz")), reason="fixtures missing")
def test_golden_fixture_reproducible(oracle):
    """The committed fixtures are exactly what the host-built reference firmware produces today."""
    if not oracle.have_fw_rx():
        pytest.skip("oracle/_ref/fw_rx not built (needs the reference tree)")
    z = np.load(os.path.join(ROOT, "tests", "golden", "rx_cases.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    for c in meta["cases"][:6]:
        r = oracle.run_fw_rx(z["frames"], c["settings"])
        assert np.array_equal(r["audio"], z[c["name"] + "/audio"])
        assert np.array_equal(r["spectra"], z[c["name"] + "/spectra"])
