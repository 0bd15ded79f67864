"""ctypes binding of oracle/_ref/libua3_hdl.so - the reference's own VHDL filters, translated to C by
tools/vhdl_eval.py and compiled by oracle/hdl/Makefile - plus the clock-domain composition of the receive chain.
TEST INFRASTRUCTURE ONLY.

Clocking model (what the reference fixes, read from the sources):
  * every RX module has reset = RX_N and clk_enable = RX (UA3REO.bdf via tools/bdf_netlist.py), so all counters
    leave reset at the same instant T_rx;
  * rx_cic runs on clk_sys (49.152 MHz), rx_ciccomp on MAIN_PLL c0 = clk_sys/32, rx_hilb on c1 = clk_sys/4,
    data_delay on c2 = clk_sys/1024, all with phase shift "0" (MAIN_PLL.v:105-116).
  [convention] ideal PLL: a derived clock rises exactly on a clk_sys edge, and a register clocked by an edge that
  coincides with an upstream register's edge samples the upstream value from BEFORE that edge (RTL semantics).
  T_rx (when the MCU raises RX, asynchronous to everything) is the one free parameter: `t_rx` below is the number
  of clk_sys edges between the last common rising edge of all PLL clocks and the first clk_sys edge after release.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libua3_hdl.so")
REF_FPGA = "/root/reference/FPGA"
_lib = None


def available():
    return os.path.exists(LIB) or os.path.isdir(REF_FPGA)


def build():
    if os.path.isdir(REF_FPGA):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "hdl")])
    if not os.path.exists(LIB):
        raise RuntimeError("oracle/_ref/libua3_hdl.so is missing and /root/reference is not present to build it")
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        for m in ("rx_cic", "rx_ciccomp", "rx_hilb", "tx_cic", "tx_ciccomp"):
            f = getattr(L, "hdl_%s_run" % m)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def run(module, x, want_ce=False):
    """Reset, release, then one rising edge per element of x (filter_in at that edge).
    Returns filter_out after every edge (and ce_out after every edge)."""
    x = np.ascontiguousarray(x, dtype=np.int64)
    out = np.zeros(x.size, np.int64)
    ce = np.zeros(x.size, np.int64) if want_ce else None
    rc = getattr(lib(), "hdl_%s_run" % module)(None, x.ctypes.data, x.size, out.ctypes.data,
                                                ce.ctypes.data if want_ce else None)
    assert rc == 0
    return (out, ce) if want_ce else out


def _sx(v, bits):
    v = np.asarray(v, np.int64) & ((1 << bits) - 1)
    return np.where(v >> (bits - 1), v - (1 << bits), v)


def seen_before(trace, t):
    """value of a per-edge trace (trace[e] = value after the producer's edge e, producer edges at times 0,1,2,..)
    as sampled by a consumer edge at producer-time t: the value after the last producer edge strictly before t;
    0 (the reset value) before the first edge."""
    t = np.asarray(t, np.int64)
    idx = t - 1
    return np.where(idx >= 0, trace[np.clip(idx, 0, trace.size - 1)], 0)


def rx_chain(x_i, x_q, t_rx=0, delay_length=130):
    """The receive chain downstream of rx_mixer_shift, all four clock domains, for one channel.
    x_i/x_q[e]: filter_in of RX_CIC_I/Q at clk_sys edge e (e = 0 is the first edge after reset release).
    Returns a dict of per-48-kHz-edge register values (what stm32_interface sees right AFTER each c2 edge) and the
    raw per-edge traces.  Time axis: clk_sys edge e happens at absolute time t_rx + e; the PLL clocks rise at
    absolute times that are multiples of 32 / 4 / 1024."""
    n = len(x_i)
    cic = {}
    for rail, x in (("i", x_i), ("q", x_q)):
        cic[rail] = _sx(run("rx_cic", np.asarray(x, np.int64) & 0x7FFFFF), 16)       # after sys edge e
    # c0 edges (clk_sys/32) strictly after release: absolute times 32*m >= t_rx  -> sys-edge-time = 32*m - t_rx
    first = -(-t_rx // 32)
    t_c0 = np.arange(first, (t_rx + n) // 32) * 32 - t_rx
    comp = {}
    for rail in ("i", "q"):
        comp[rail] = _sx(run("rx_ciccomp", seen_before(cic[rail], t_c0) & 0xFFFF), 16)   # after c0 edge m
    # c1 edges (clk_sys/4): sample RX_CICCOMP_I.filter_out
    t_c1_abs = np.arange(-(-t_rx // 4), (t_rx + n) // 4) * 4
    c0_abs = np.arange(first, (t_rx + n) // 32) * 32

    def comp_seen(rail, t_abs):
        # value after the last c0 edge strictly before absolute time t_abs
        k = np.searchsorted(c0_abs, t_abs, side="left") - 1
        return np.where(k >= 0, comp[rail][np.clip(k, 0, comp[rail].size - 1)], 0)

    hilb = _sx(run("rx_hilb", comp_seen("i", t_c1_abs) & 0xFFFF), 16)                # after c1 edge
    # c2 edges (clk_sys/1024): data_delay.v:16-30 - fifo_buf[0] <= data_in, fifo_buf[k] <= fifo_buf[k-1],
    # data_out = fifo_buf[delay_length-1]; no reset, power-up contents taken as 0
    t_c2_abs = np.arange(-(-t_rx // 1024), (t_rx + n) // 1024) * 1024
    q_in = comp_seen("q", t_c2_abs)
    voice_q = np.zeros_like(q_in)
    voice_q[delay_length - 1:] = q_in[:q_in.size - (delay_length - 1)]              # after c2 edge j: in[j-(N-1)]
    if delay_length == 130:
        # ... and the same from data_delay.v itself (translated by tools/verilog_eval.py with the schematic's parameters)
        from . import vlog_ref
        if vlog_ref.available():
            d = vlog_ref.VModule("data_delay_q")
            executed = np.zeros_like(q_in)
            for j, v in enumerate(q_in):
                d["data_in"] = int(v)
                d.clock("clk_in")
                executed[j] = d.signed("data_out")
            assert np.array_equal(executed, voice_q), "data_delay.v disagrees with its restatement"
            voice_q = executed

    def hilb_seen(t_abs):
        k = np.searchsorted(t_c1_abs, t_abs, side="left") - 1
        return np.where(k >= 0, hilb[np.clip(k, 0, hilb.size - 1)], 0)

    return {"cic_i": cic["i"], "cic_q": cic["q"], "comp_i": comp["i"], "comp_q": comp["q"], "hilb": hilb,
            "c0_abs": c0_abs, "c1_abs": t_c1_abs, "c2_abs": t_c2_abs, "voice_q_after_c2": voice_q,
            "comp_seen": comp_seen, "hilb_seen": hilb_seen}


def frames_at(ch, tau, n_frames=None):
    """The 8-byte frames the MCU reads `tau` clk_sys ticks after every 48 kHz (c2) edge (stm32_interface.v:228-271
    latches SPEC at k=400 and VOICE at k=404; both are taken at the same instant here).  uint8 [n, 8]."""
    t_read = ch["c2_abs"] + tau + 1                 # +1: include an edge that coincides with the read instant
    spec_i = ch["comp_seen"]("i", t_read)
    spec_q = ch["comp_seen"]("q", t_read)
    voice_i = ch["hilb_seen"](t_read)
    voice_q = ch["voice_q_after_c2"]
    words = np.stack([spec_q, spec_i, voice_q, voice_i], axis=1) & 0xFFFF
    out = np.zeros((words.shape[0], 8), np.uint8)
    out[:, 0::2] = words >> 8
    out[:, 1::2] = words & 0xFF
    return out if n_frames is None else out[:n_frames]


def tx_chain(words, t_tx=0, tau=0):
    """The transmit interpolators across their two clock domains, for one rail: TX_I / TX_Q as stm32_interface holds it (a new
    word every 1024 clk_sys ticks, written `tau` ticks after the 48 kHz grid) -> tx_ciccomp on MAIN_PLL c3 = clk_sys x 23 / 512
    (46 clocks per word, MAIN_PLL.v:117-119) -> tx_cic on clk_sys; both leave reset `t_tx` ticks after an instant at which
    all clocks rise together (TX_N, UA3REO.bdf via tools/bdf_netlist.py).  Same [convention] as rx_chain: ideal PLL, a
    register clocked at the same instant as its producer sees the producer's previous value.  Returns tx_cic.filter_out
    (14 bits) after every clk_sys edge."""
    words = np.asarray(words, np.int64)
    n = words.size * 1024
    j = np.arange(int(np.ceil(t_tx * 23 / 512.0)), int((t_tx + n) * 23 / 512.0))
    t_c3 = j * 512.0 / 23.0 - t_tx                                      # c3 edges, in clk_sys ticks after the release
    k = np.floor((t_c3 + t_tx - tau) / 1024.0).astype(np.int64)
    w = np.where((k >= 0) & (k < words.size), words[np.clip(k, 0, words.size - 1)], 0)
    comp = _sx(run("tx_ciccomp", w & 0xFFFF), 16)                       # after every c3 edge
    e = np.arange(n, dtype=np.float64)
    idx = np.searchsorted(t_c3, e, side="left") - 1                     # the last c3 edge strictly before the clk_sys edge
    xin = np.where(idx >= 0, comp[np.clip(idx, 0, comp.size - 1)], 0)
    return _sx(run("tx_cic", xin & 0xFFFF), 14)
