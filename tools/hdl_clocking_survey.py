#!/usr/bin/env python3
"""Which frame does the board assemble?  Sweeps the one free clocking parameter of the receive chain - T_rx, the
instant the MCU raises RX relative to the PLL clocks - through the reference's own HDL (oracle/hdl_ref.py: C
translation of rx_cic / rx_ciccomp / rx_hilb, data_delay restated) and, for every T_rx and bus read instant tau
(clk_sys ticks after the 48 kHz edge that interrupts the MCU, stm32_interface.v:228-271), expresses the four words
the MCU would read in terms of the golden model's streams:
    SPEC_x  = Y_x[k]            Y = golden compensator output, polyphase alignment A (golden convention) or B
    VOICE_I = H[k - dI]         H = golden Hilbert output of Y_I (zero-latency convention)
    VOICE_Q = Y_Q[k - dQ]
TEST INFRASTRUCTURE (documentation of the conventions in DESIGN.md 2); needs /root/reference."""
import ctypes
import os
import sys
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hdl_ref, pyoracle  # noqa: E402


class _Comp(ctypes.Structure):
    _fields_ = [("p0", ctypes.c_int16 * 33), ("p1", ctypes.c_int16 * 33), ("n_in", ctypes.c_uint32)]


def golden_comp(u):
    L = pyoracle.lib()
    c = _Comp()
    y = ctypes.c_int16(0)
    out = []
    for v in u:
        if L.ua3g_rx_ciccomp_push(ctypes.byref(c), ctypes.c_int16(int(v)), ctypes.byref(y)):
            out.append(y.value)
    return np.array(out, np.int64)


def golden_hilb(y):
    L = pyoracle.lib()
    L.ua3g_rx_hilb_push.restype = ctypes.c_int16
    h = (ctypes.c_int16 * 256)()
    return np.array([L.ua3g_rx_hilb_push(ctypes.byref(h), ctypes.c_int16(int(v))) for v in y], np.int64)


def find_lag(seq, ref, lo=-4, hi=140, n=120, start=150):
    """k such that seq[j] == ref[j - k] for j in [start, start+n)"""
    for k in range(lo, hi):
        if start - k < 0:
            continue
        if np.array_equal(seq[start:start + n], ref[start - k:start - k + n]):
            return k
    return None


def survey(t_rx_values, taus, n_frames=330, seed=5):
    rng = np.random.default_rng(seed)
    n = n_frames * 1024
    # band-limited-ish random mixer outputs (full 23-bit range) - any input will do, the relation is structural
    x_i = rng.integers(-(1 << 22), 1 << 22, n)
    x_q = rng.integers(-(1 << 22), 1 << 22, n)
    cic_i = hdl_ref._sx(hdl_ref.run("rx_cic", x_i & 0x7FFFFF), 16)
    u_i = cic_i[1::512]            # output_register after edges 1, 513, ... = golden CIC outputs (checked elsewhere)
    cic_q = hdl_ref._sx(hdl_ref.run("rx_cic", x_q & 0x7FFFFF), 16)
    u_q = cic_q[1::512]
    gold = {}
    for name, pre in (("A", 0), ("B", 1)):
        yi = golden_comp(np.concatenate([np.zeros(pre, np.int64), u_i]))
        yq = golden_comp(np.concatenate([np.zeros(pre, np.int64), u_q]))
        gold[name] = (yi, yq, golden_hilb(yi))
    rows = []
    for t_rx in t_rx_values:
        ch = hdl_ref.rx_chain(x_i, x_q, t_rx=t_rx)
        c2 = ch["c2_abs"]
        for tau in taus:
            t_read = c2 + tau                       # SPEC latched at k=400, VOICE at k=404: same instant here
            spec_i = ch["comp_seen"]("i", t_read + 1)   # +1: value after the last edge at or before t_read
            spec_q = ch["comp_seen"]("q", t_read + 1)
            voice_i = ch["hilb_seen"](t_read + 1)
            voice_q = ch["voice_q_after_c2"]
            hit = None
            for name, (yi, yq, hh) in gold.items():
                k = find_lag(spec_i, yi)
                if k is None or find_lag(spec_q, yq) != k:
                    continue
                d_i = find_lag(voice_i, hh)
                d_q = find_lag(voice_q, yq)
                if d_i is None or d_q is None:
                    continue
                hit = (name, k, d_i - k, d_q - k)
            rows.append((t_rx, tau, hit))
    return rows


def main():
    step = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    taus = (0, 64, 160, 320, 640)
    rows = survey(range(0, 1024, step), taus)
    cnt = Counter()
    for t_rx, tau, hit in rows:
        cnt[(tau, hit[0], hit[2], hit[3]) if hit else (tau, None)] += 1
    print("tau  alignment  dI  dQ   share of T_rx values")
    for key in sorted(cnt, key=str):
        print(key, cnt[key] * len(taus) / len(rows))
    return rows


if __name__ == "__main__":
    main()
