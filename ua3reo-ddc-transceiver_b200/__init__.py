"""ua3reo-ddc-transceiver_b200 - host-side Python binding of libua3reo_b200.so.

This is plumbing (ctypes over the C ABI in include/ua3reo_b200.h) used by tests/ and bench.py; the
product is the shared library.  There is NO CPU fallback: if the CUDA library has not been built
(`__graft_entry__.build()` / `ua3reo-ddc-transceiver_b200/build.sh`) importing this module raises.

The directory name carries a hyphen (the repository layout asks for it), so import it through
`ua3reo_loader.load()` at the repo root, which registers it as `ua3reo_ddc_transceiver_b200`.
"""
import ctypes
import os

import numpy as np

from . import sharding  # noqa: F401
from . import synth  # noqa: F401  (synthetic ADC stream / tuning words of SURVEY.md 8d)
from .synth import random_fcw, synth_adc  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libua3reo_b200.so")

ADC_PER_FRAME = 1024
FRAME_BYTES = 8
AUDIO_BLOCK = 192
FFT_SIZE = 512
FFT_BINS = 256
ADC_CLOCK_HZ = 49152000
DEFAULT_FCW = 620407  # stm32_interface.v:56


class UA3Error(RuntimeError):
    pass


class DeviceBlock:
    """n int16 ADC samples at device address ptr, already ordered for the context's stream (Receiver.push takes it as is)."""
    __slots__ = ("ptr", "n")

    def __init__(self, ptr, n):
        self.ptr, self.n = int(ptr), int(n)


# trx_manager.h:11-24
MODE_LSB, MODE_USB, MODE_IQ, MODE_CW_L, MODE_CW_U, MODE_DIGI_L, MODE_DIGI_U, MODE_NO_TX, MODE_NFM, MODE_WFM, MODE_AM, \
    MODE_LOOPBACK = range(12)


class TxSettings(ctypes.Structure):
    """struct ua3reo_tx_settings (include/ua3reo_b200.h)."""
    _fields_ = [("mode", ctypes.c_uint8), ("mute", ctypes.c_uint8), ("tune", ctypes.c_uint8), ("key_down", ctypes.c_uint8),
                ("rf_power", ctypes.c_uint8), ("volume", ctypes.c_uint8), ("reserved", ctypes.c_uint8 * 2),
                ("filter_width", ctypes.c_uint16), ("ssb_hpf_pass", ctypes.c_uint16)]
    FIELDS = ("mode", "mute", "tune", "key_down", "rf_power", "volume", "filter_width", "ssb_hpf_pass")

    def as_dict(self):
        return {k: int(getattr(self, k)) for k in self.FIELDS}


class RxSettings(ctypes.Structure):
    """struct ua3reo_rx_settings (include/ua3reo_b200.h): the TRX fields processRxAudio()/FFT_doFFT() read."""
    _fields_ = [("mode", ctypes.c_uint8), ("agc", ctypes.c_uint8), ("agc_speed", ctypes.c_uint8), ("dnr", ctypes.c_uint8),
                ("notch", ctypes.c_uint8), ("mute", ctypes.c_uint8), ("volume", ctypes.c_uint8), ("rf_gain", ctypes.c_uint8),
                ("fm_sql_threshold", ctypes.c_uint8), ("fft_enabled", ctypes.c_uint8), ("fft_averaging", ctypes.c_uint8),
                ("fft_zoom", ctypes.c_uint8), ("iq_swap", ctypes.c_uint8), ("cw_decoder", ctypes.c_uint8),
                ("reserved", ctypes.c_uint8 * 2),
                ("filter_width", ctypes.c_uint16), ("ssb_hpf_pass", ctypes.c_uint16), ("notch_fc", ctypes.c_uint16),
                ("reserved2", ctypes.c_uint16)]

    FIELDS = ("mode", "agc", "agc_speed", "dnr", "notch", "mute", "volume", "rf_gain", "fm_sql_threshold", "fft_enabled",
              "fft_averaging", "fft_zoom", "iq_swap", "cw_decoder", "filter_width", "ssb_hpf_pass", "notch_fc")

    def as_dict(self):
        return {k: int(getattr(self, k)) for k in self.FIELDS}


def _bind(lib):
    c = ctypes
    vp, u32, sz = c.c_void_p, c.c_uint32, c.c_size_t
    sigs = {
        "ua3reo_version": (c.c_char_p, []),
        "ua3reo_last_error": (c.c_char_p, []),
        "ua3reo_create": (c.c_int, [c.c_int, u32, u32, c.POINTER(vp)]),
        "ua3reo_destroy": (c.c_int, [vp]),
        "ua3reo_reset": (c.c_int, [vp]),
        "ua3reo_ddc_set_clocking": (c.c_int, [vp, c.c_int, c.c_int, c.c_int]),
        "ua3reo_ddc_get_clocking": (c.c_int, [vp, c.POINTER(c.c_int), c.POINTER(c.c_int), c.POINTER(c.c_int)]),
        "ua3reo_reserve_sms": (c.c_int, [vp, c.c_int]),
        "ua3reo_n_channels": (u32, [vp]),
        "ua3reo_max_block_samples": (u32, [vp]),
        "ua3reo_set_fcw": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_get_fcw": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_phrase_from_frequency": (u32, [u32, c.POINTER(c.c_int)]),
        "ua3reo_set_frequency": (c.c_int, [vp, u32, u32]),
        "ua3reo_ddc_push": (c.c_int, [vp, vp, sz, c.POINTER(sz)]),
        "ua3reo_ddc_push_device": (c.c_int, [vp, vp, sz, c.POINTER(sz)]),
        "ua3reo_ddc_read_frames": (c.c_int, [vp, vp, sz]),
        "ua3reo_ddc_read_frames_async": (c.c_int, [vp, vp, sz]),
        "ua3reo_ddc_frames_device": (c.c_int, [vp, c.POINTER(vp), c.POINTER(sz), c.POINTER(sz), c.POINTER(sz), c.POINTER(sz)]),
        "ua3reo_rx_defaults": (None, [c.POINTER(RxSettings)]),
        "ua3reo_rx_enable": (c.c_int, [vp, c.c_int]),
        "ua3reo_rx_set": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_rx_set_live": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_rx_set_notch": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_rx_set_agc_speed": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_rx_fft_init": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_rx_push_frames": (c.c_int, [vp, vp, sz]),
        "ua3reo_rx_counts": (c.c_int, [vp, c.POINTER(sz), c.POINTER(sz)]),
        "ua3reo_rx_read_audio": (c.c_int, [vp, vp, sz]),
        "ua3reo_rx_read_spectra": (c.c_int, [vp, vp, sz]),
        "ua3reo_rx_read_audio_usb": (c.c_int, [vp, vp, sz]),
        "ua3reo_rx_read_smeter": (c.c_int, [vp, vp, c.c_int]),
        "ua3reo_rx_stage": (c.c_int, [vp, u32, c.c_int, vp, vp, sz, c.c_int]),
        "ua3reo_rx_read_waterfall": (c.c_int, [vp, vp, sz]),
        "ua3reo_copy_stream": (c.c_int, [vp, c.POINTER(vp)]),
        "ua3reo_rx_read_audio_async": (c.c_int, [vp, vp, sz]),
        "ua3reo_rx_read_spectra_async": (c.c_int, [vp, vp, sz]),
        "ua3reo_rx_read_waterfall_history": (c.c_int, [vp, vp]),
        "ua3reo_rx_move_waterfall": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_rx_read_cw": (c.c_int, [vp, vp, sz]),
        "ua3reo_adc_stats": (c.c_int, [vp, c.POINTER(c.c_int16), c.POINTER(c.c_int16), c.POINTER(u32), c.c_int]),
        "ua3reo_smeter_dbm": (c.c_int16, [c.c_float, c.c_float, c.c_uint8]),
        "ua3reo_get_params": (c.c_int, [vp, vp, c.POINTER(c.c_int16), c.POINTER(c.c_int16), c.c_int]),
        "ua3reo_cw_decoder_init": (None, [vp]),
        "ua3reo_cw_decoder_step": (c.c_int, [vp, c.c_float, u32, vp, c.c_int]),
        "ua3reo_autogain_init": (None, [vp]),
        "ua3reo_autogain_step": (None, [vp, c.c_int16]),
        "ua3reo_duc_enable": (c.c_int, [vp, u32]),
        "ua3reo_duc_push": (c.c_int, [vp, vp, sz]),
        "ua3reo_duc_push_wire": (c.c_int, [vp, vp, sz]),
        "ua3reo_duc_read_dac": (c.c_int, [vp, vp, sz]),
        "ua3reo_duc_dac_device": (c.c_int, [vp, c.POINTER(vp), c.POINTER(sz), c.POINTER(sz)]),
        "ua3reo_duc_read_otr": (c.c_int, [vp, vp]),
        "ua3reo_tx_defaults": (None, [c.POINTER(TxSettings)]),
        "ua3reo_tx_enable": (c.c_int, [vp, u32]),
        "ua3reo_tx_set": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_tx_set_live": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_tx_process": (c.c_int, [vp, vp, sz]),
        "ua3reo_tx_read_iq": (c.c_int, [vp, vp, vp, sz]),
        "ua3reo_tx_read_loopback": (c.c_int, [vp, vp, sz]),
        "ua3reo_tx_feed_duc": (c.c_int, [vp]),
        "ua3reo_bank_create": (c.c_int, [c.c_int, c.POINTER(c.c_int), u32, u32, c.POINTER(vp)]),
        "ua3reo_bank_destroy": (c.c_int, [vp]),
        "ua3reo_bank_n_devices": (c.c_int, [vp]),
        "ua3reo_bank_context": (c.c_int, [vp, c.c_int, c.POINTER(vp), c.POINTER(u32), c.POINTER(u32)]),
        "ua3reo_bank_set_fcw": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_bank_rx_enable": (c.c_int, [vp, c.c_int]),
        "ua3reo_bank_rx_set": (c.c_int, [vp, u32, u32, vp]),
        "ua3reo_bank_push": (c.c_int, [vp, vp, sz, c.POINTER(sz)]),
        "ua3reo_bank_read_frames": (c.c_int, [vp, vp, sz]),
        "ua3reo_bank_rx_counts": (c.c_int, [vp, c.POINTER(sz), c.POINTER(sz)]),
        "ua3reo_bank_rx_read_audio": (c.c_int, [vp, vp, sz]),
        "ua3reo_bank_rx_read_spectra": (c.c_int, [vp, vp, sz]),
        "ua3reo_bank_sync": (c.c_int, [vp]),
        "ua3reo_fanout_create": (c.c_int, [c.c_int, c.c_int, c.c_int, c.c_int, sz, c.c_int, c.POINTER(vp)]),
        "ua3reo_fanout_destroy": (c.c_int, [vp]),
        "ua3reo_fanout_disconnect": (c.c_int, [vp]),
        "ua3reo_fanout_handle": (c.c_int, [vp, vp]),
        "ua3reo_fanout_connect": (c.c_int, [vp, vp]),
        "ua3reo_fanout_send": (c.c_int, [vp, vp, sz]),
        "ua3reo_fanout_acquire": (c.c_int, [vp, vp, c.POINTER(vp)]),
        "ua3reo_fanout_release": (c.c_int, [vp, vp]),
        "ua3reo_fanout_sync": (c.c_int, [vp]),
        "ua3reo_fanout_info": (c.c_int, [vp, c.POINTER(c.c_int), c.POINTER(c.c_uint64), c.POINTER(c.c_uint64)]),
        "ua3reo_gather_create": (c.c_int, [c.c_int, c.c_int, c.c_int, c.c_int, sz, c.c_int, c.POINTER(vp)]),
        "ua3reo_gather_disconnect": (c.c_int, [vp]),
        "ua3reo_gather_destroy": (c.c_int, [vp]),
        "ua3reo_gather_handle": (c.c_int, [vp, vp]),
        "ua3reo_gather_connect": (c.c_int, [vp, vp]),
        "ua3reo_gather_send": (c.c_int, [vp, vp, vp]),
        "ua3reo_gather_acquire": (c.c_int, [vp, vp, c.POINTER(vp), c.POINTER(sz)]),
        "ua3reo_gather_release": (c.c_int, [vp, vp]),
        "ua3reo_sync": (c.c_int, [vp]),
        "ua3reo_stream": (c.c_int, [vp, c.POINTER(vp)]),
        "ua3reo_launch_count": (c.c_uint64, [vp]),
        "ua3reo_profile_begin": (c.c_int, [vp, u32]),
        "ua3reo_profile_begin_kernel": (c.c_int, [vp, u32, u32]),
        "ua3reo_profile_end": (c.c_int, [vp, c.POINTER(c.c_double), u32, c.POINTER(u32)]),
        "ua3reo_measure_int32_peak": (c.c_int, [c.c_int, c.POINTER(c.c_double)]),
        "ua3reo_measure_lds_peak": (c.c_int, [c.c_int, c.POINTER(c.c_double)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load_library(path=None):
    """Loads the CUDA library.  Raises UA3Error when it is missing: no fallback exists."""
    path = path or LIB_PATH   # module global, read at call time
    if not os.path.exists(path):
        raise UA3Error(
            "%s not found: build it with ua3reo-ddc-transceiver_b200/build.sh (nvcc, sm_100a). "
            "There is no CPU fallback." % path)
    return _bind(ctypes.CDLL(path))


def declared_symbols(header_path):
    """Names of the functions include/*.h declares (used by the symbol-export test)."""
    import re
    txt = open(header_path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ua3reo_\w+|processRxAudio|processTxAudio|initAudioProcessor|FFT_Init|FFT_doFFT|"
                                 r"FPGA_fpgadata_iqclock|getPhraseFromFrequency)\s*\(", txt)))


def phrase_from_frequency(freq_hz, lib=None):
    lib = lib or load_library()
    swap = ctypes.c_int(0)
    w = lib.ua3reo_phrase_from_frequency(int(freq_hz), ctypes.byref(swap))
    return int(w), bool(swap.value)


class CwDecoder(ctypes.Structure):
    """Host half of the CW decoder (ua3reo_cw_decoder): feed one Goertzel magnitude per 192-sample block (Receiver.read_cw)."""
    _fields_ = ([("level_hi", ctypes.c_float), ("level_lo", ctypes.c_float)] +
                [(n, ctypes.c_uint8) for n in ("raw", "raw_prev", "key", "key_prev", "flushed", "reserved")] +
                [("wpm", ctypes.c_uint16)] +
                [(n, ctypes.c_int64) for n in ("t_raw_edge", "t_key_down", "t_key_up", "mark_ms", "space_ms", "dit_ms")] +
                [("code", ctypes.c_char * 24)])

    def __init__(self, lib=None):
        super().__init__()
        self._lib = lib or load_library()
        self._lib.ua3reo_cw_decoder_init(ctypes.byref(self))
        self._buf = ctypes.create_string_buffer(16)
        self.tick_ms = 0

    def step(self, magnitude, tick_ms=None):
        """One audio block: returns the characters decoded by this block (a space marks a word gap)."""
        self.tick_ms = self.tick_ms + 4 if tick_ms is None else int(tick_ms)
        n = self._lib.ua3reo_cw_decoder_step(ctypes.byref(self), float(magnitude), self.tick_ms, self._buf, 16)
        return self._buf.raw[:min(n, 16)].decode("ascii")


class AutoGain(ctypes.Structure):
    """TRX_DoAutoGain() state (ua3reo_autogain): stage, wait counter and the Preamp / ATT / LPF / BPF decisions."""
    _fields_ = [(n, ctypes.c_uint8) for n in ("stage", "wait", "preamp", "att", "lpf", "bpf")]

    def __init__(self, lib=None):
        super().__init__()
        self._lib = lib or load_library()
        self._lib.ua3reo_autogain_init(ctypes.byref(self))

    def step(self, adc_max_amplitude):
        self._lib.ua3reo_autogain_step(ctypes.byref(self), int(adc_max_amplitude))
        return (self.stage, self.wait, self.preamp, self.att, self.lpf, self.bpf)


class Receiver:
    """Batched receiver bank: n_channels DDCs over one shared ADC stream (one context = one GPU)."""

    def __init__(self, n_channels, max_block_samples=1 << 20, device=0, _lib_path=None):
        self.lib = load_library(_lib_path)
        h = ctypes.c_void_p()
        self._h = None
        self._chk(self.lib.ua3reo_create(int(device), int(n_channels), int(max_block_samples), ctypes.byref(h)))
        self._h = h
        self.n_channels = int(n_channels)
        self.max_block_samples = int(self.lib.ua3reo_max_block_samples(h))
        self.device = int(device)
        self.last_frames = 0

    def _chk(self, rc):
        if rc != 0:
            raise UA3Error("ua3reo error %d: %s" % (rc, self.lib.ua3reo_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.lib.ua3reo_destroy(self._h)      # synchronises the context's streams first
            self._h = None
            self._inflight = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._chk(self.lib.ua3reo_reset(self._h))

    def set_clocking(self, align_b=1, d_i=3, d_q=129):
        """Clocking class of the I/Q frame (include/ua3reo_b200.h: ua3reo_ddc_set_clocking); before the first push."""
        self._chk(self.lib.ua3reo_ddc_set_clocking(self._h, int(align_b), int(d_i), int(d_q)))

    def get_clocking(self):
        a, i, q = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._chk(self.lib.ua3reo_ddc_get_clocking(self._h, ctypes.byref(a), ctypes.byref(i), ctypes.byref(q)))
        return a.value, i.value, q.value

    def reserve_sms(self, n_sms):
        """Keep n_sms SMs out of the front kernel's grid for the caller's concurrent kernels (an NCCL broadcast)."""
        self._chk(self.lib.ua3reo_reserve_sms(self._h, int(n_sms)))

    def set_fcw(self, fcw, first=0):
        a = np.ascontiguousarray(fcw, dtype=np.uint32)
        self._chk(self.lib.ua3reo_set_fcw(self._h, int(first), a.size, a.ctypes.data))

    def get_fcw(self):
        a = np.zeros(self.n_channels, np.uint32)
        self._chk(self.lib.ua3reo_get_fcw(self._h, 0, a.size, a.ctypes.data))
        return a

    def set_frequency(self, channel, freq_hz):
        self._chk(self.lib.ua3reo_set_frequency(self._h, int(channel), int(freq_hz)))

    def push(self, adc, assume_ordered=False):
        """adc: 1-D int16 numpy array (host) or a CUDA torch tensor of dtype int16 on this device.
        assume_ordered: the caller has already ordered the library stream after the tensor's producer and keeps the
        tensor alive and unchanged until the push completes (sharding.AdcBroadcaster does)."""
        n = ctypes.c_size_t(0)
        if isinstance(adc, DeviceBlock):       # a raw device block whose ordering the producer took care of (sharding.AdcFanout)
            self._chk(self.lib.ua3reo_ddc_push_device(self._h, adc.ptr, adc.n, ctypes.byref(n)))
        elif isinstance(adc, np.ndarray):
            a = np.ascontiguousarray(adc, dtype=np.int16)
            self._keep = a
            self._chk(self.lib.ua3reo_ddc_push(self._h, a.ctypes.data, a.size, ctypes.byref(n)))
        else:  # torch tensor
            if adc.is_cuda:
                assert adc.dtype.itemsize == 2 and adc.is_contiguous()
                # The library consumes whole-block device pushes IN PLACE on its own non-blocking stream: order that
                # stream after the tensor's producer (torch's current stream) and keep the tensor alive until an event
                # recorded behind the push has completed.  (Tensor.record_stream is not used: the caching allocator
                # would touch the library's stream when the tensor dies, possibly after close() destroyed it.)
                lib_stream = None
                if not assume_ordered:
                    import torch
                    lib_stream = torch.cuda.ExternalStream(self.stream(), device=adc.device)
                    lib_stream.wait_stream(torch.cuda.current_stream(adc.device))
                self._chk(self.lib.ua3reo_ddc_push_device(self._h, adc.data_ptr(), adc.numel(), ctypes.byref(n)))
                if lib_stream is not None:
                    ev = torch.cuda.Event()
                    ev.record(lib_stream)
                    self._inflight = [(e, t) for e, t in getattr(self, "_inflight", []) if not e.query()] + [(ev, adc)]
            else:
                assert adc.dtype.itemsize == 2 and adc.is_contiguous()
                self._keep = adc
                self._chk(self.lib.ua3reo_ddc_push(self._h, adc.data_ptr(), adc.numel(), ctypes.byref(n)))
        self.last_frames = int(n.value)
        return self.last_frames

    def read_frames(self, out=None):
        """Returns uint8 [n_channels, frames_of_last_push, 8] in stm32_interface byte order."""
        nf = self.last_frames
        if out is None:
            out = np.empty((self.n_channels, nf, FRAME_BYTES), np.uint8)
        ptr = out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()
        self._chk(self.lib.ua3reo_ddc_read_frames(self._h, ptr, nf))
        return out

    def read_frames_async(self, out):
        """Enqueue the copy of the last push's frames into `out` (pinned torch tensor or numpy array) and return;
        `out` is valid after sync().  Lets the next push overlap the device-to-host copy."""
        ptr = out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()
        self._chk(self.lib.ua3reo_ddc_read_frames_async(self._h, ptr, self.last_frames))
        return out

    def frames_device(self):
        """(base pointer, first ring index, frames of last push, ring size in frames, bytes between channels)"""
        base, first, nf, ring, stride = (ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t(),
                                         ctypes.c_size_t())
        self._chk(self.lib.ua3reo_ddc_frames_device(self._h, ctypes.byref(base), ctypes.byref(first), ctypes.byref(nf),
                                                    ctypes.byref(ring), ctypes.byref(stride)))
        return base.value, int(first.value), int(nf.value), int(ring.value), int(stride.value)

    # ---- STM32 stage: processRxAudio / FFT_doFFT ----
    def rx_defaults(self, **overrides):
        s = RxSettings()
        self.lib.ua3reo_rx_defaults(ctypes.byref(s))
        for k, v in overrides.items():
            setattr(s, k, v)
        return s

    def rx_enable(self, on=True):
        self._chk(self.lib.ua3reo_rx_enable(self._h, 1 if on else 0))

    def rx_set(self, settings, first=0, live=False):
        """settings: one RxSettings (applied to every channel from `first`) or a list, one per channel.
        live=True: only the fields the firmware reads on every call (no filter reselection, no state cleared)."""
        if isinstance(settings, RxSettings):
            settings = [settings] * (self.n_channels - first)
        arr = (RxSettings * len(settings))(*settings)
        fn = self.lib.ua3reo_rx_set_live if live else self.lib.ua3reo_rx_set
        self._chk(fn(self._h, int(first), len(settings), ctypes.cast(arr, ctypes.c_void_p)))

    def rx_set_notch(self, notch_fc, first=0):
        """InitNotchFilter() alone: notch_fc is one corner in Hz or one per channel from `first`."""
        a = np.atleast_1d(np.asarray(notch_fc, dtype=np.uint16))
        if a.size == 1:
            a = np.repeat(a, self.n_channels - first)
        a = np.ascontiguousarray(a)
        self._chk(self.lib.ua3reo_rx_set_notch(self._h, int(first), a.size, a.ctypes.data))

    def rx_set_agc_speed(self, speed, first=0):
        """InitAGC() alone: one speed (1..) or one per channel from `first`."""
        a = np.atleast_1d(np.asarray(speed, dtype=np.uint8))
        if a.size == 1:
            a = np.repeat(a, self.n_channels - first)
        a = np.ascontiguousarray(a)
        self._chk(self.lib.ua3reo_rx_set_agc_speed(self._h, int(first), a.size, a.ctypes.data))

    def rx_fft_init(self, zoom, first=0):
        """FFT_Init() alone: one zoom factor or one per channel from `first`."""
        a = np.atleast_1d(np.asarray(zoom, dtype=np.uint8))
        if a.size == 1:
            a = np.repeat(a, self.n_channels - first)
        a = np.ascontiguousarray(a)
        self._chk(self.lib.ua3reo_rx_fft_init(self._h, int(first), a.size, a.ctypes.data))

    def rx_push_frames(self, frames):
        """frames: uint8 [n_channels, n, 8] I/Q frames (stm32_interface byte order) fed straight to the STM32 stage."""
        a = np.ascontiguousarray(frames, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[0] == self.n_channels and a.shape[2] == FRAME_BYTES
        self._keep_frames = a
        self._chk(self.lib.ua3reo_rx_push_frames(self._h, a.ctypes.data, a.shape[1]))
        self.last_frames = a.shape[1]

    def rx_counts(self):
        a, f = ctypes.c_size_t(), ctypes.c_size_t()
        self._chk(self.lib.ua3reo_rx_counts(self._h, ctypes.byref(a), ctypes.byref(f)))
        return int(a.value), int(f.value)

    def read_audio(self, out=None):
        """int32 [n_channels, blocks_of_last_push, 384] (L/R interleaved, Processor_AudioBuffer layout)."""
        nb, _ = self.rx_counts()
        if out is None:
            out = np.empty((self.n_channels, nb, 2 * AUDIO_BLOCK), np.int32)
        ptr = out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()
        self._chk(self.lib.ua3reo_rx_read_audio(self._h, ptr, nb))
        return out

    def read_audio_usb(self):
        """int16 [n_channels, blocks_of_last_push, 384]: the USB-audio packing of the same blocks."""
        nb, _ = self.rx_counts()
        out = np.empty((self.n_channels, nb, 2 * AUDIO_BLOCK), np.int16)
        self._chk(self.lib.ua3reo_rx_read_audio_usb(self._h, out.ctypes.data, nb))
        return out

    def read_spectra(self, out=None):
        """float32 [n_channels, fft_frames_of_last_push, 256] (FFTOutput_mean after each FFT_doFFT)."""
        _, nf = self.rx_counts()
        if out is None:
            out = np.empty((self.n_channels, nf, FFT_BINS), np.float32)
        ptr = out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()
        self._chk(self.lib.ua3reo_rx_read_spectra(self._h, ptr, nf))
        return out

    def read_waterfall(self):
        """uint16 [n_channels, fft_frames_of_last_push, 256]: RGB565 waterfall rows (fft-shifted)."""
        _, nf = self.rx_counts()
        out = np.empty((self.n_channels, nf, FFT_BINS), np.uint16)
        self._chk(self.lib.ua3reo_rx_read_waterfall(self._h, out.ctypes.data, nf))
        return out

    def read_audio_async(self, out):
        """Enqueue the copy of the last push's audio into `out` (pinned torch/numpy int32 [n_ch, blocks, 384]); valid after sync()."""
        nb, _ = self.rx_counts()
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        self._chk(self.lib.ua3reo_rx_read_audio_async(self._h, ptr, nb))
        return nb

    def read_spectra_async(self, out):
        """Enqueue the copy of the last push's spectra into `out` (pinned float32 [n_ch, frames, 256]); valid after sync()."""
        _, nf = self.rx_counts()
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        self._chk(self.lib.ua3reo_rx_read_spectra_async(self._h, ptr, nf))
        return nf

    def read_waterfall_history(self):
        """uint16 [n_channels, 50, 256]: the firmware's wtf_buffer per channel, row 0 the newest."""
        out = np.empty((self.n_channels, 50, FFT_BINS), np.uint16)
        self._chk(self.lib.ua3reo_rx_read_waterfall_history(self._h, out.ctypes.data))
        return out

    def move_waterfall(self, freq_diff_hz, first=0):
        """A retune as FFT_printFFT() sees it: Hz difference per channel (or one for all), applied by the next FFT frame."""
        a = np.atleast_1d(np.asarray(freq_diff_hz, dtype=np.int32))
        if a.size == 1:
            a = np.repeat(a, self.n_channels - first)
        a = np.ascontiguousarray(a)
        self._chk(self.lib.ua3reo_rx_move_waterfall(self._h, int(first), a.size, a.ctypes.data))

    def read_cw(self):
        """float32 [n_channels, blocks_of_last_push]: Goertzel magnitude of the CW decoder front end."""
        nb, _ = self.rx_counts()
        out = np.empty((self.n_channels, nb), np.float32)
        self._chk(self.lib.ua3reo_rx_read_cw(self._h, out.ctypes.data, nb))
        return out

    def adc_stats(self, reset=False):
        mn, mx, nr = ctypes.c_int16(), ctypes.c_int16(), ctypes.c_uint32()
        self._chk(self.lib.ua3reo_adc_stats(self._h, ctypes.byref(mn), ctypes.byref(mx), ctypes.byref(nr), 1 if reset else 0))
        return int(mn.value), int(mx.value), int(nr.value)

    def get_params(self, dac_otr=False):
        """Command 2 of the wire protocol: (packet uint8[5], TRX_ADC_MINAMPLITUDE, TRX_ADC_MAXAMPLITUDE); resets the extremes."""
        pkt = np.zeros(5, np.uint8)
        mn, mx = ctypes.c_int16(), ctypes.c_int16()
        self._chk(self.lib.ua3reo_get_params(self._h, pkt.ctypes.data, ctypes.byref(mn), ctypes.byref(mx), 1 if dac_otr else 0))
        return pkt, int(mn.value), int(mx.value)

    def rx_stage(self, channel, stage, buf, arg=0):
        """dc_filter (stage 0, arg = stateNum), DoAGC (1) or processNoiseReduction (2, 64 samples) on a float32 buffer with the
        channel's state; returns the result (include/ua3reo_b200.h: ua3reo_rx_stage)."""
        a = np.array(buf, dtype=np.float32, copy=True)
        out = a.copy()
        self._chk(self.lib.ua3reo_rx_stage(self._h, int(channel), int(stage), a.ctypes.data, out.ctypes.data, a.size, int(arg)))
        return out if stage == 2 else a

    def read_smeter(self, reset=False):
        out = np.empty((self.n_channels, 2), np.float32)
        self._chk(self.lib.ua3reo_rx_read_smeter(self._h, out.ctypes.data, 1 if reset else 0))
        return out

    # ---- transmit DUC ----
    def duc_enable(self, max_tx_samples=64):
        self._chk(self.lib.ua3reo_duc_enable(self._h, int(max_tx_samples)))

    def duc_push(self, iq):
        """iq: int16 [n_channels, n, 2] (I, Q) 48 kHz TX samples; returns n."""
        a = np.ascontiguousarray(iq, dtype=np.int16)
        assert a.ndim == 3 and a.shape[0] == self.n_channels and a.shape[2] == 2
        self._keep_tx = a
        self._chk(self.lib.ua3reo_duc_push(self._h, a.ctypes.data, a.shape[1]))
        self._last_tx = a.shape[1]
        return a.shape[1]

    def duc_push_wire(self, wire):
        """wire: uint8 [n_channels, n, 4] = Q hi, Q lo, I hi, I lo per sample (command 3 of the bus); returns n."""
        a = np.ascontiguousarray(wire, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[0] == self.n_channels and a.shape[2] == 4
        self._chk(self.lib.ua3reo_duc_push_wire(self._h, a.ctypes.data, a.shape[1]))
        self._last_tx = a.shape[1]
        return a.shape[1]

    def duc_read_dac(self):
        out = np.empty((self.n_channels, self._last_tx * 1024), np.uint16)
        self._chk(self.lib.ua3reo_duc_read_dac(self._h, out.ctypes.data, self._last_tx))
        return out

    def duc_read_otr(self):
        out = np.empty(self.n_channels, np.uint32)
        self._chk(self.lib.ua3reo_duc_read_otr(self._h, out.ctypes.data))
        return out

    # ---- transmit audio (processTxAudio) ----
    def tx_defaults(self, **overrides):
        s = TxSettings()
        self.lib.ua3reo_tx_defaults(ctypes.byref(s))
        for k, v in overrides.items():
            setattr(s, k, v)
        return s

    def tx_enable(self, max_blocks=8):
        self._chk(self.lib.ua3reo_tx_enable(self._h, int(max_blocks)))

    def tx_set(self, settings, first=0, live=False):
        if isinstance(settings, TxSettings):
            settings = [settings] * (self.n_channels - first)
        arr = (TxSettings * len(settings))(*settings)
        fn = self.lib.ua3reo_tx_set_live if live else self.lib.ua3reo_tx_set
        self._chk(fn(self._h, int(first), len(settings), ctypes.cast(arr, ctypes.c_void_p)))

    def tx_process(self, mic):
        """mic: int16 [n_channels, n_blocks*192, 2] (left, right).  Returns (iq_words int16, iq_float float32), same shape."""
        a = np.ascontiguousarray(mic, dtype=np.int16)
        assert a.ndim == 3 and a.shape[0] == self.n_channels and a.shape[1] % AUDIO_BLOCK == 0 and a.shape[2] == 2
        nb = a.shape[1] // AUDIO_BLOCK
        self._chk(self.lib.ua3reo_tx_process(self._h, a.ctypes.data, nb))
        self._tx_blocks = nb
        w = np.empty(a.shape, np.int16)
        f = np.empty(a.shape, np.float32)
        self._chk(self.lib.ua3reo_tx_read_iq(self._h, w.ctypes.data, f.ctypes.data, nb))
        return w, f

    def tx_read_loopback(self, n_blocks):
        """int32 [n_channels, n_blocks*192, 2]: what TRX_MODE_LOOPBACK hands to the codec (zeros for other modes)."""
        out = np.empty((self.n_channels, int(n_blocks) * AUDIO_BLOCK, 2), np.int32)
        self._chk(self.lib.ua3reo_tx_read_loopback(self._h, out.ctypes.data, int(n_blocks)))
        return out

    def tx_feed_duc(self):
        self._chk(self.lib.ua3reo_tx_feed_duc(self._h))
        self._last_tx = getattr(self, "_tx_blocks", 0) * AUDIO_BLOCK

    def sync(self):
        self._chk(self.lib.ua3reo_sync(self._h))
        self._inflight = []

    def stream(self):
        s = ctypes.c_void_p()
        self._chk(self.lib.ua3reo_stream(self._h, ctypes.byref(s)))
        return s.value or 0

    def copy_stream(self):
        """cudaStream_t of the pipelined result reads, as an integer (torch.cuda.ExternalStream)."""
        s = ctypes.c_void_p()
        self._chk(self.lib.ua3reo_copy_stream(self._h, ctypes.byref(s)))
        return s.value or 0

    def profile_begin(self, max_blocks, kernel=None):
        """kernel: None = events around every kernel; "front" etc. = only the two events around that kernel."""
        if kernel is None:
            self._chk(self.lib.ua3reo_profile_begin(self._h, int(max_blocks)))
        else:
            names = ["adc_expand", "front", "ciccomp", "hilb", "rotate", "rx_audio", "rx_fft"]
            self._chk(self.lib.ua3reo_profile_begin_kernel(self._h, int(max_blocks), names.index(kernel)))

    def profile_end(self):
        ms = (ctypes.c_double * 7)()
        nb = ctypes.c_uint32(0)
        self._chk(self.lib.ua3reo_profile_end(self._h, ms, 7, ctypes.byref(nb)))
        names = ["adc_expand", "front", "ciccomp", "hilb", "rotate", "rx_audio", "rx_fft"]
        return {k: float(v) for k, v in zip(names, ms)}, int(nb.value)

    def launch_count(self):
        return int(self.lib.ua3reo_launch_count(self._h))


class Bank:
    """Channels sharded over several devices from ONE host process (ua3reo_bank_*): the ADC block is fanned out by copy
    engines (cudaMemcpyPeerAsync), results are read slab by slab into [n_channels, ...] arrays."""

    def __init__(self, devices, n_channels, max_block_samples=1 << 20, _lib_path=None):
        self.lib = load_library(_lib_path)
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = ctypes.c_void_p()
        self._h = None
        self._chk(self.lib.ua3reo_bank_create(len(devices), devs, int(n_channels), int(max_block_samples), ctypes.byref(h)))
        self._h = h
        self.n_channels = int(n_channels)
        self.last_frames = 0

    def _chk(self, rc):
        if rc != 0:
            raise UA3Error("ua3reo error %d: %s" % (rc, self.lib.ua3reo_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.lib.ua3reo_bank_destroy(self._h)
            self._h = None

    def slabs(self):
        out = []
        for i in range(self.lib.ua3reo_bank_n_devices(self._h)):
            f, n = ctypes.c_uint32(), ctypes.c_uint32()
            self._chk(self.lib.ua3reo_bank_context(self._h, i, None, ctypes.byref(f), ctypes.byref(n)))
            out.append((int(f.value), int(n.value)))
        return out

    def set_fcw(self, fcw, first=0):
        a = np.ascontiguousarray(fcw, dtype=np.uint32)
        self._chk(self.lib.ua3reo_bank_set_fcw(self._h, int(first), a.size, a.ctypes.data))

    def rx_enable(self, on=True):
        self._chk(self.lib.ua3reo_bank_rx_enable(self._h, 1 if on else 0))

    def rx_set(self, settings, first=0):
        arr = (RxSettings * len(settings))(*settings)
        self._chk(self.lib.ua3reo_bank_rx_set(self._h, int(first), len(settings), ctypes.cast(arr, ctypes.c_void_p)))

    def push(self, adc):
        a = np.ascontiguousarray(adc, dtype=np.int16)
        self._keep = a
        n = ctypes.c_size_t(0)
        self._chk(self.lib.ua3reo_bank_push(self._h, a.ctypes.data, a.size, ctypes.byref(n)))
        self.last_frames = int(n.value)
        return self.last_frames

    def read_frames(self):
        out = np.empty((self.n_channels, self.last_frames, FRAME_BYTES), np.uint8)
        self._chk(self.lib.ua3reo_bank_read_frames(self._h, out.ctypes.data, self.last_frames))
        return out

    def rx_counts(self):
        a, f = ctypes.c_size_t(), ctypes.c_size_t()
        self._chk(self.lib.ua3reo_bank_rx_counts(self._h, ctypes.byref(a), ctypes.byref(f)))
        return int(a.value), int(f.value)

    def read_audio(self):
        nb, _ = self.rx_counts()
        out = np.empty((self.n_channels, nb, 2 * AUDIO_BLOCK), np.int32)
        self._chk(self.lib.ua3reo_bank_rx_read_audio(self._h, out.ctypes.data, nb))
        return out

    def read_spectra(self):
        _, nf = self.rx_counts()
        out = np.empty((self.n_channels, nf, FFT_BINS), np.float32)
        self._chk(self.lib.ua3reo_bank_rx_read_spectra(self._h, out.ctypes.data, nf))
        return out

    def sync(self):
        self._chk(self.lib.ua3reo_bank_sync(self._h))


def frames_to_iq(frames):
    """uint8 [..., 8] frames -> dict of int16 arrays, as FPGA_fpgadata_getiq rebuilds them (fpga.c:286-385)."""
    f = np.asarray(frames, dtype=np.uint8)
    w = (f[..., 0::2].astype(np.uint16) << 8) | f[..., 1::2].astype(np.uint16)
    w = w.astype(np.int16)
    return {"spec_q": w[..., 0], "spec_i": w[..., 1], "voice_q": w[..., 2], "voice_i": w[..., 3]}


def measure_lds_peak(device=0, lib=None):
    lib = lib or load_library()
    v = ctypes.c_double(0.0)
    rc = lib.ua3reo_measure_lds_peak(int(device), ctypes.byref(v))
    if rc != 0:
        raise UA3Error("ua3reo error %d: %s" % (rc, lib.ua3reo_last_error().decode()))
    return float(v.value)


def measure_int32_peak(device=0, lib=None):
    lib = lib or load_library()
    v = ctypes.c_double(0.0)
    rc = lib.ua3reo_measure_int32_peak(int(device), ctypes.byref(v))
    if rc != 0:
        raise UA3Error("ua3reo error %d: %s" % (rc, lib.ua3reo_last_error().decode()))
    return float(v.value)
