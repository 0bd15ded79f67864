#!/usr/bin/env python3
"""Generates tests/golden/hdl_cases.npz: frames produced by the reference's OWN VHDL (tools/vhdl_eval.py ->
oracle/_ref/libua3_hdl.so, clock domains composed by oracle/hdl_ref.py) for a set of ADC streams, tuning words and
clocking instants.  Needs /root/reference; the vectors travel to the GPU box, where the CUDA DDC is compared with
them directly (tests/test_hdl_pin.py).  The NCO and the mixer multiply are the golden model's (the NCO IP is
encrypted; mixer.v is an lpm_mult wizard file, i.e. a signed multiply).

Every case stores: the recipe of the ADC stream (kind, seed), fcw, t_rx, tau, the clocking class the frames show
(align_b, d_i, d_q), the number of leading HDL frames before golden frame 0 (lag), and the HDL frames from there on."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import hdl_ref, pyoracle  # noqa: E402
from hdl_cases import CASES, N_ADC, make_adc  # noqa: E402


def classify(hdl_frames, adc, fcw):
    """(align_b, d_i, d_q, lag): the golden clocking class whose frames equal the HDL's, hdl[lag + n] == golden[n]"""
    n_cmp = 1500
    for align_b in (1, 0):
        for d_i in (3, 2):
            for d_q in (129, 130):
                g = pyoracle.GoldenDDC(fcw, (align_b, d_i, d_q)).push(adc)
                for lag in range(0, 4):
                    if np.array_equal(hdl_frames[lag:lag + n_cmp], g[:n_cmp]):
                        return align_b, d_i, d_q, lag
    return None


def main():
    out = {}
    names = []
    for name, kind, seed, fcw, t_rx, tau in CASES:
        adc = make_adc(kind, seed)
        x_i, x_q = pyoracle.executed_mixer(adc, fcw)          # nco_shift.v, mixer.v, rx_mixer_shift.v executed; the NCO is the golden model's
        g_i, g_q = pyoracle.golden_mixer(adc, fcw)
        assert np.array_equal(x_i, g_i) and np.array_equal(x_q, g_q), name
        ch = hdl_ref.rx_chain(x_i, x_q, t_rx=t_rx)
        hf = hdl_ref.frames_at(ch, tau)
        cls = classify(hf, adc, fcw)
        assert cls is not None, "no clocking class reproduces the HDL frames of case " + name
        align_b, d_i, d_q, lag = cls
        g = pyoracle.GoldenDDC(fcw, (align_b, d_i, d_q)).push(adc)
        n = min(len(g), len(hf) - lag)
        assert np.array_equal(hf[lag:lag + n], g[:n]), name
        print("%-10s fcw=%7d t_rx=%4d tau=%3d -> class (%s, %d, %d), lag %d, %d frames identical to the golden model"
              % (name, fcw, t_rx, tau, "AB"[align_b], d_i, d_q, lag, n))
        names.append(name)
        out[name + "_frames"] = hf[lag:lag + n]
        out[name + "_meta"] = np.array([fcw, t_rx, tau, align_b, d_i, d_q, lag, seed], np.int64)
        out[name + "_cic_i"] = ch["cic_i"][1::512][:2 * n].astype(np.int16)      # output_register loads (96 kHz)
        out[name + "_cic_q"] = ch["cic_q"][1::512][:2 * n].astype(np.int16)
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "hdl_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
