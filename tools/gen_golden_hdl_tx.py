#!/usr/bin/env python3
"""tests/golden/hdl_tx_cases.npz: DAC words of the transmit chain with the reference's own HDL executed on every stage but the
NCO - tx_ciccomp.vhd on PLL c3 and tx_cic.vhd on clk_sys across their clock domains (oracle/hdl_ref.tx_chain), then
tx_mixer.v x 2, tx_summator.v and DAC_corrector.v (oracle/vlog_ref, vector runner), the NCO being the golden model's
(ua3g_nco: its Altera submodules are encrypted).  The HDL stream is the golden stream delayed by a pure latency
(tests/test_hdl_pin.py::test_golden_duc_composition_equals_hdl_chain); the vectors are stored with that latency taken out,
NCO phase 0 at the first DAC word of the first TX sample (the golden model's and the product's convention).

  <name>_iq   int16 [n, 2]        TX_I, TX_Q words (what stm32_interface.v latches from command 3)
  <name>_meta int64 [4]           fcw, t_tx, tau, latency found (clk_sys ticks)
  <name>_dac  uint16 [m]          DAC words (14 bits), m = (n - 2) * 1024
  <name>_otr  uint8 [m]           tx_summator overflow (DAC_OTR)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hdl_ref, pyoracle, vlog_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "hdl_tx_cases.npz")
N_WORDS = 26
# (name, kind, seed, fcw, t_tx, tau)
CASES = (("random_a", "random", 1, 620407, 0, 0), ("random_b", "random", 2, 2345678, 517, 300), ("square", "square", 0, 1234567, 100, 900),
         ("impulse", "impulse", 0, 605867, 990, 64), ("fullscale", "fullscale", 0, (1 << 22) - 1, 333, 512), ("slow", "random", 3, 1, 7, 700))


def make_iq(kind, seed, n=N_WORDS):
    rng = np.random.default_rng(500 + seed)
    if kind == "random":
        return rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
    if kind == "square":
        t = np.arange(n)
        return np.stack([np.where(t & 1, -32768, 32767), np.where((t >> 1) & 1, 32767, -32768)], axis=1).astype(np.int16)
    if kind == "impulse":
        a = np.zeros((n, 2), np.int16)
        a[3] = [32767, -32768]
        return a
    if kind == "fullscale":
        return np.tile(np.array([[-32768, -32768]], np.int16), (n, 1))
    raise ValueError(kind)


def golden_cic(words):
    """the golden model's interpolator stream for one rail (ua3g_tx_ciccomp_push -> ua3g_tx_cic_clock, duc_golden.c)"""
    dac_unused, _ = None, None
    import ctypes
    L = pyoracle.lib()
    L.ua3g_tx_cic_clock.restype = ctypes.c_int16

    class TxCic(ctypes.Structure):
        _fields_ = [("cnt", ctypes.c_uint32), ("wreg", ctypes.c_int16), ("d", ctypes.c_int64 * 5), ("up", ctypes.c_int64),
                    ("i", ctypes.c_int64 * 5), ("out14", ctypes.c_int16)]
    dp, z, c = (ctypes.c_int16 * 24)(), (ctypes.c_int16 * 2)(), TxCic()
    L.ua3g_tx_cic_reset(ctypes.byref(c))
    out = []
    for v in words:
        L.ua3g_tx_ciccomp_push(ctypes.byref(dp), ctypes.c_int16(int(v)), z)
        for h in range(2):
            zz = ctypes.c_int16(z[h])
            out += [L.ua3g_tx_cic_clock(ctypes.byref(c), zz) for _ in range(512)]
    return np.array(out, np.int64)


def output_stage(i14, q14, sin14, cos14):
    """tx_mixer.v (I x sin, Q x cos: UA3REO.bdf), tx_summator.v, DAC_corrector.v, executed stage by stage over whole streams"""
    ones = np.ones(len(i14), np.int64)
    mi = vlog_ref.VModule("tx_mixer").run({"dataa": i14 & 0x3FFF, "datab": sin14 & 0x3FFF, "clken": ones}, ["result"], clock="clock")["result"]
    mq = vlog_ref.VModule("tx_mixer").run({"dataa": q14 & 0x3FFF, "datab": cos14 & 0x3FFF, "clken": ones}, ["result"], clock="clock")["result"]
    sm = vlog_ref.VModule("tx_summator").run({"dataa": mi & 0xFFFFFFF, "datab": mq & 0xFFFFFFF, "clken": ones}, ["result", "overflow"], clock="clock")
    dac = vlog_ref.VModule("DAC_corrector").run({"DATA_IN": sm["result"] & 0xFFFFFFF}, ["DATA_OUT"], clock="clk_in")["DATA_OUT"]
    return dac.astype(np.uint16), sm["overflow"].astype(np.uint8)


def generate():
    out = {}
    for name, kind, seed, fcw, t_tx, tau in CASES:
        iq = make_iq(kind, seed)
        streams, lag_found = [], None
        for rail in (0, 1):
            h = hdl_ref.tx_chain(iq[:, rail], t_tx=t_tx, tau=tau)
            g = golden_cic(iq[:, rail])
            lags = [lag for lag in (1024, 2048) if np.array_equal(h[lag:], g[:g.size - lag])]
            assert len(lags) == 1, (name, rail)
            assert lag_found in (None, lags[0])
            lag_found = lags[0]
            streams.append(h[lag_found:][:(N_WORDS - 2) * 1024])
        m = streams[0].size
        s14, c14 = pyoracle.golden_nco(m, fcw)
        dac, otr = output_stage(streams[0], streams[1], s14, c14)
        out[name + "_iq"], out[name + "_meta"] = iq, np.array([fcw, t_tx, tau, lag_found], np.int64)
        out[name + "_dac"], out[name + "_otr"] = dac, otr
        print("%-10s fcw=%7d t_tx=%4d tau=%3d: latency %d ticks, %d DAC words, %d overflows" % (name, fcw, t_tx, tau, lag_found, m, int(otr.sum())))
    return out


if __name__ == "__main__":
    data = generate()
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
