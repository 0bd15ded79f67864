"""ctypes binding of the CPU checker (oracle/_build/libua3_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libua3_oracle.so")
DDC_STATE_BYTES = 4096   # >= sizeof(ua3g_ddc)


def build(force=False):
    if force or not os.path.exists(LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(LIB)
            for f in os.listdir(_HERE) if f.endswith((".c", ".h"))):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        L.ua3g_ddc_push.restype = ctypes.c_size_t
        L.ua3g_ddc_push.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.ua3g_ddc_init.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.ua3g_ddc_init_clocking.restype = ctypes.c_int
        L.ua3g_ddc_init_clocking.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ua3g_phrase_from_frequency.restype = ctypes.c_uint32
        L.ua3g_phrase_from_frequency.argtypes = [ctypes.c_uint32, ctypes.POINTER(ctypes.c_int)]
        L.ua3g_nco.argtypes = [ctypes.c_uint32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
        L.ua3g_rx_mix.restype = ctypes.c_int32
        L.ua3g_rx_mix.argtypes = [ctypes.c_int32, ctypes.c_int32]
        L.ua3g_bank_create.restype = ctypes.c_void_p
        L.ua3g_bank_create.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.ua3g_bank_destroy.argtypes = [ctypes.c_void_p]
        L.ua3g_bank_push.restype = ctypes.c_double
        L.ua3g_bank_push.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                     ctypes.c_int]
        L.ua3g_duc_init.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.ua3g_duc_push.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                    ctypes.c_void_p]
        _lib = L
    return _lib


class GoldenDDC:
    """One channel of the register-transfer golden DDC (ddc_golden.c); state persists across push()."""

    def __init__(self, fcw, clocking=None):
        """clocking: None = the model's default class (B, 3, 129), else (align_b, d_i, d_q)"""
        self._st = ctypes.create_string_buffer(DDC_STATE_BYTES)
        if clocking is None:
            lib().ua3g_ddc_init(self._st, int(fcw) & 0x3FFFFF)
        elif lib().ua3g_ddc_init_clocking(self._st, int(fcw) & 0x3FFFFF, *[int(v) for v in clocking]) != 0:
            raise ValueError("bad clocking class %r" % (clocking,))

    def push(self, adc, want_cic=False):
        adc = np.ascontiguousarray(adc, dtype=np.int16)
        cap = adc.size // 1024 + 2
        frames = np.zeros((cap, 8), np.uint8)
        ncic = ctypes.c_size_t(0)
        if want_cic:
            ccap = adc.size // 512 + 2
            ci = np.zeros(ccap, np.int16); cq = np.zeros(ccap, np.int16)
            nf = lib().ua3g_ddc_push(self._st, adc.ctypes.data, adc.size, frames.ctypes.data, cap,
                                     ci.ctypes.data, cq.ctypes.data, ccap, ctypes.byref(ncic))
            return frames[:nf].copy(), ci[:ncic.value].copy(), cq[:ncic.value].copy()
        nf = lib().ua3g_ddc_push(self._st, adc.ctypes.data, adc.size, frames.ctypes.data, cap, None, None, 0, None)
        return frames[:nf].copy()


def golden_frames(adc, fcws, clocking=None):
    """uint8 [n_ch, n_frames, 8] for a list of tuning words over one ADC stream."""
    out = []
    for w in fcws:
        out.append(GoldenDDC(w, clocking).push(adc))
    return np.stack(out)


def golden_nco(n, fcw):
    """(sin14, cos14) of the golden NCO (ua3g_nco: start phase 0, the documented output convention) for n samples"""
    L = lib()
    phase = (np.arange(n, dtype=np.int64) * int(fcw)) & 0x3FFFFF
    uniq, inv = np.unique(phase, return_inverse=True)
    s14 = np.zeros(uniq.size, np.int64)
    c14 = np.zeros(uniq.size, np.int64)
    s = ctypes.c_int32(0)
    c = ctypes.c_int32(0)
    for i, ph in enumerate(uniq):
        L.ua3g_nco(int(ph), ctypes.byref(s), ctypes.byref(c))
        s14[i], c14[i] = s.value, c.value
    return s14[inv], c14[inv]


def executed_mixer(adc, fcw):
    """golden_mixer with the restated arithmetic replaced by the reference's own nco_shift.v, mixer.v and rx_mixer_shift.v,
    executed (oracle/vlog_ref.py); only the NCO in front of them stays the golden model's."""
    from . import vlog_ref
    s14, c14 = golden_nco(len(adc), fcw)
    return vlog_ref.rx_mixer_path(adc, s14), vlog_ref.rx_mixer_path(adc, c14)


def golden_mixer(adc, fcw):
    """(x_i, x_q): the s23 words RX_CIC_I/Q.filter_in see for every ADC sample (golden NCO + mixer conventions)"""
    L = lib()
    adc = np.asarray(adc, np.int64)
    n = adc.size
    phase = (np.arange(n, dtype=np.int64) * int(fcw)) & 0x3FFFFF
    uniq, inv = np.unique(phase, return_inverse=True)
    s14 = np.zeros(uniq.size, np.int64)
    c14 = np.zeros(uniq.size, np.int64)
    s = ctypes.c_int32(0)
    c = ctypes.c_int32(0)
    for i, ph in enumerate(uniq):
        L.ua3g_nco(int(ph), ctypes.byref(s), ctypes.byref(c))
        s14[i], c14[i] = s.value, c.value

    def mix(nco14):
        p = adc * (nco14[inv] >> 2)
        p &= 0x7FFFFF
        return np.where(p >> 22, p - (1 << 23), p)
    return mix(s14), mix(c14)


def synth_adc(n, seed=20261018, tones=8, noise_lsb=8.0, level_dbfs=-6.0):
    """Synthetic 12-bit ADC stream of SURVEY.md 8(d): K tones + white Gaussian noise, clipped to 12 bit."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    f = rng.uniform(0.02, 0.48, tones)
    ph = rng.uniform(0, 2 * np.pi, tones)
    amp = 2047.0 * 10 ** (level_dbfs / 20.0) / tones
    x = np.zeros(n)
    for k in range(tones):
        x += amp * np.sin(2 * np.pi * f[k] * t + ph[k])
    x += rng.normal(0.0, noise_lsb, n)
    return np.clip(np.rint(x), -2048, 2047).astype(np.int16)


class GoldenBank:
    """n_ch golden DDC channels driven by n_threads host threads (bench baseline only)."""

    def __init__(self, fcws):
        self.fcw = np.ascontiguousarray(fcws, dtype=np.uint32)
        self.n_ch = self.fcw.size
        self._b = lib().ua3g_bank_create(self.n_ch, self.fcw.ctypes.data)

    def push(self, adc, n_threads, want_frames=False):
        adc = np.ascontiguousarray(adc, dtype=np.int16)
        frames = np.zeros((self.n_ch, adc.size // 1024, 8), np.uint8) if want_frames else None
        dt = lib().ua3g_bank_push(self._b, self.n_ch, adc.ctypes.data, adc.size,
                                  frames.ctypes.data if want_frames else None, int(n_threads))
        return (dt, frames) if want_frames else dt

    def __del__(self):
        try:
            lib().ua3g_bank_destroy(self._b)
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# The reference firmware's own receive-audio / FFT code, host-built (oracle/_ref/fw_rx)
# ---------------------------------------------------------------------------------------------
FW_RX = os.path.join(_HERE, "_ref", "fw_rx")
RX_PARAM_KEYS = {"mode": "mode", "filter_width": "filter_width", "ssb_hpf_pass": "hpf_pass", "rf_gain": "rf_gain",
                 "agc": "agc", "agc_speed": "agc_speed", "dnr": "dnr", "notch": "notch", "notch_fc": "notch_fc",
                 "volume": "volume", "mute": "mute", "fm_sql_threshold": "fm_sql", "fft_enabled": "fft_enabled",
                 "fft_zoom": "fft_zoom", "fft_averaging": "fft_averaging", "iq_swap": "iq_swap", "cw_decoder": "cw_decoder", "freq": "freq"}


def have_fw_rx():
    return os.path.exists(FW_RX)


FW_RX_B200 = os.path.join(_HERE, "_ref", "fw_rx_b200")     # same driver, firmware DSP units replaced by the GPU shim
FW_TX_B200 = os.path.join(_HERE, "_ref", "fw_tx_b200")


def run_fw_rx(frames, settings, workdir=None, events=(), binary=None, env=None):
    """Runs the host-built reference firmware over uint8 [n_frames, 8] frames for one channel.
    settings: dict with the ua3reo_rx_settings field names.  events: (frame_index, key, value) applied mid-stream
    without re-initialisation; keys are settings names or the pseudo keys reinit / notch_init / fft_init.
    binary: another build of the same driver (FW_RX_B200).  Returns dict(audio int32 [nb, 384],
    smeter float32 [nb, 2], cw float32 [nb], spectra float32 [nf, 256], waterfall uint16 [nf, 256], fft_max float32 [nf])."""
    import tempfile
    frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, 8)
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        pp, fp, ap, sp, wp = (os.path.join(d, n) for n in ("p.txt", "frames.bin", "audio.bin", "fft.bin", "wtf.bin"))
        with open(pp, "w") as f:
            for k, v in settings.items():
                if k in RX_PARAM_KEYS:
                    f.write("%s %d\n" % (RX_PARAM_KEYS[k], int(v)))
            for at, k, v in events:
                f.write("at %d %s %d\n" % (int(at), RX_PARAM_KEYS.get(k, k), int(v)))
        frames.tofile(fp)
        subprocess.check_call([binary or FW_RX, pp, fp, ap, sp, wp], env=env)
        a = np.fromfile(ap, dtype=np.int32).reshape(-1, 387 + 192)
        raw = np.fromfile(sp, dtype=np.uint8)
        rec = 256 * 4 + 256 * 2 + 4
        raw = raw[:rec * (raw.size // rec)].reshape(-1, rec)
        return {
            "audio": a[:, :384].copy(),
            "smeter": a[:, 384:386].copy().view(np.float32),
            "cw": a[:, 386].copy().view(np.float32),
            "usb": np.ascontiguousarray(a[:, 387:]).view(np.int16).reshape(-1, 384),
            "spectra": np.ascontiguousarray(raw[:, :1024]).view(np.float32).reshape(-1, 256),
            "waterfall": np.ascontiguousarray(raw[:, 1024:1536]).view(np.uint16).reshape(-1, 256),
            "fft_max": np.ascontiguousarray(raw[:, 1536:1540]).view(np.float32).reshape(-1),
            "wtf_history": np.fromfile(wp, dtype=np.uint16).reshape(50, 256),
        }


class GoldenDUC:
    """One channel of the register-transfer golden transmit DUC (duc_golden.c)."""

    def __init__(self, fcw):
        self._st = ctypes.create_string_buffer(DDC_STATE_BYTES)
        lib().ua3g_duc_init(self._st, int(fcw) & 0x3FFFFF)

    def push(self, tx_i, tx_q):
        tx_i = np.ascontiguousarray(tx_i, dtype=np.int16)
        tx_q = np.ascontiguousarray(tx_q, dtype=np.int16)
        n = tx_i.size
        dac = np.zeros(n * 1024, np.uint16)
        otr = np.zeros(n * 1024, np.uint8)
        lib().ua3g_duc_push(self._st, tx_i.ctypes.data, tx_q.ctypes.data, n, dac.ctypes.data, otr.ctypes.data)
        return dac, otr


FW_TX = os.path.join(_HERE, "_ref", "fw_tx")
TX_PARAM_KEYS = {"mode": "mode", "filter_width": "filter_width", "ssb_hpf_pass": "hpf_pass", "rf_power": "rf_power",
                 "mute": "mute", "tune": "tune", "key_down": "key", "volume": "volume"}


def have_fw_tx():
    return os.path.exists(FW_TX)


def run_fw_tx(mic, settings, workdir=None, binary=None, env=None, want_codec=False):
    """Runs the host-built reference firmware's processTxAudio() over int16 [n, 2] codec samples (n multiple of 192).
    Returns (iq_words int16 [n, 2], iq_float float32 [n, 2]) and, with want_codec, the int32 [n, 2] the loopback branch
    hands to the codec."""
    import tempfile
    mic = np.ascontiguousarray(mic, dtype=np.int16).reshape(-1, 2)
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        pp, mp_, op = (os.path.join(d, n) for n in ("p.txt", "mic.bin", "out.bin"))
        with open(pp, "w") as f:
            for k, v in settings.items():
                if k in TX_PARAM_KEYS:
                    f.write("%s %d\n" % (TX_PARAM_KEYS[k], int(v)))
        mic.tofile(mp_)
        subprocess.check_call([binary or FW_TX, pp, mp_, op], env=env)
        raw = np.fromfile(op, dtype=np.uint8).reshape(-1, 192 * 2 * 4 + 192 * 2 * 2 + 192 * 2 * 4)
        f = np.ascontiguousarray(raw[:, :1536]).view(np.float32).reshape(-1, 2)
        w = np.ascontiguousarray(raw[:, 1536:2304]).view(np.int16).reshape(-1, 2)
        if want_codec:
            return w, f, np.ascontiguousarray(raw[:, 2304:]).view(np.int32).reshape(-1, 2)
        return w, f
