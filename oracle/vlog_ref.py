"""ctypes binding of oracle/_ref/libua3_vlog.so - the reference's hand-written Verilog units (mixer.v, tx_mixer.v,
tx_summator.v, rx_mixer_shift.v, nco_shift.v, data_delay.v, DAC_corrector.v, stm32_interface.v), translated to C by
tools/verilog_eval.py and compiled by oracle/hdl/Makefile.  TEST INFRASTRUCTURE ONLY.

    m = VModule("mixer")
    m["dataa"], m["datab"], m["clken"] = -2048, -2048, 1     # values are taken modulo the port width
    m.settle()                                               # continuous assignments after an input change
    m.clock("clock")                                         # one rising edge
    m["result"], m.signed("result")                          # raw bits / as a two's complement number
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libua3_vlog.so")
REF_FPGA = "/root/reference/FPGA"
_lib = None


def available():
    return os.path.exists(LIB) or os.path.isdir(REF_FPGA)


def lib():
    global _lib
    if _lib is None:
        if os.path.isdir(REF_FPGA):
            subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "hdl")])
        if not os.path.exists(LIB):
            raise RuntimeError("oracle/_ref/libua3_vlog.so is missing and /root/reference is not present to build it")
        L = ctypes.CDLL(LIB)
        vp = ctypes.c_void_p
        L.vl_find.restype, L.vl_find.argtypes = vp, [ctypes.c_char_p]
        L.vl_size.restype, L.vl_size.argtypes = ctypes.c_size_t, [vp]
        L.vl_init.restype, L.vl_init.argtypes = None, [vp, vp]
        L.vl_settle.restype, L.vl_settle.argtypes = None, [vp, vp]
        L.vl_n_fields.restype, L.vl_n_fields.argtypes = ctypes.c_int, [vp]
        L.vl_field_name.restype, L.vl_field_name.argtypes = ctypes.c_char_p, [vp, ctypes.c_int]
        L.vl_field_offset.restype, L.vl_field_offset.argtypes = ctypes.c_size_t, [vp, ctypes.c_int]
        for f in ("vl_field_width", "vl_field_length", "vl_field_signed"):
            getattr(L, f).restype, getattr(L, f).argtypes = ctypes.c_int, [vp, ctypes.c_int]
        L.vl_clock_edge.restype, L.vl_clock_edge.argtypes = ctypes.c_int, [vp, vp, ctypes.c_char_p]
        L.vl_run.restype = ctypes.c_int
        L.vl_run.argtypes = [vp, vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, vp, vp, ctypes.c_int, vp, vp]
        _lib = L
    return _lib


class VModule:
    def __init__(self, name):
        L = lib()
        self._L = L
        self._m = L.vl_find(name.encode())
        if not self._m:
            raise KeyError("no translated Verilog module %r" % name)
        self._buf = ctypes.create_string_buffer(L.vl_size(self._m))
        self._fields = {}
        for i in range(L.vl_n_fields(self._m)):
            self._fields[L.vl_field_name(self._m, i).decode()] = (
                L.vl_field_offset(self._m, i), L.vl_field_width(self._m, i), L.vl_field_length(self._m, i),
                bool(L.vl_field_signed(self._m, i)))
        self.reset()

    def reset(self):
        self._L.vl_init(self._m, self._buf)

    def _word(self, name, index):
        off, width, length, _ = self._fields[name]
        if length:
            assert index is not None and 0 <= index < length
            off += 8 * index
        return ctypes.c_uint64.from_buffer(self._buf, off), width

    def __getitem__(self, key):
        name, index = key if isinstance(key, tuple) else (key, None)
        w, _ = self._word(name, index)
        return int(w.value)

    def __setitem__(self, key, value):
        name, index = key if isinstance(key, tuple) else (key, None)
        w, width = self._word(name, index)
        w.value = int(value) & ((1 << width) - 1)

    def signed(self, name, index=None):
        w, width = self._word(name, index)
        v = int(w.value)
        return v - (1 << width) if v >> (width - 1) else v

    def width(self, name):
        return self._fields[name][1]

    def settle(self):
        self._L.vl_settle(self._m, self._buf)

    def clock(self, name):
        if self._L.vl_clock_edge(self._m, self._buf, name.encode()) != 0:
            raise KeyError("module has no process on posedge %s" % name)

    def run(self, inputs, outputs, clock=None):
        """Vector form: inputs {name: int array of n}, outputs [names]; per step the inputs are applied, the continuous
        assignments settle, `clock` rises once (if given).  Returns {name: int64 array}, two's complement for signed signals."""
        names = list(inputs)
        n = len(np.asarray(inputs[names[0]]))
        order = list(self._fields)
        a = np.ascontiguousarray(np.stack([np.asarray(inputs[k], dtype=np.int64).astype(np.uint64) for k in names]))
        out = np.zeros((len(outputs), n), np.uint64)
        ii = np.array([order.index(k) for k in names], np.int32)
        oi = np.array([order.index(k) for k in outputs], np.int32)
        rc = self._L.vl_run(self._m, self._buf, clock.encode() if clock else None, n, len(names), ii.ctypes.data, a.ctypes.data,
                            len(outputs), oi.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise KeyError("vl_run failed (%d): unknown clock or a memory used as a scalar" % rc)
        res = {}
        for k, row in zip(outputs, out):
            _, width, _, sgn = self._fields[k]
            v = row.astype(np.int64)
            res[k] = np.where(v >> (width - 1) != 0, v - (1 << width), v) if sgn and width < 64 else v
        return res


def rx_mixer_path(adc12, nco14):
    """nco_shift.v -> mixer.v (lpm_mult, one pipeline register) -> rx_mixer_shift.v for whole streams, executed: the s23 word
    RX_CIC.filter_in sees for every (ADC sample, NCO output) pair.  The one-clock latency of the multiplier is taken out."""
    n12 = VModule("nco_shift").run({"in": np.asarray(nco14, np.int64) & 0x3FFF}, ["out"])["out"]
    res = VModule("mixer").run({"dataa": np.asarray(adc12, np.int64) & 0xFFF, "datab": n12, "clken": np.ones(len(n12), np.int64)},
                               ["result"], clock="clock")["result"]
    out = VModule("rx_mixer_shift").run({"in": res & 0xFFFFFF}, ["out"])["out"]
    return np.where(out >> 22 != 0, out - (1 << 23), out)             # out is declared unsigned [22:0]: read it as the s23 the CIC takes
