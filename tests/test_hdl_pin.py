"""Pins the integer golden model - and through it the CUDA DDC / DUC - to the reference's OWN HDL.

The reference's FPGA filters are executed from their VHDL source text by tools/vhdl_eval.py (a Python interpreter and
a C translation compiled into oracle/_ref/libua3_hdl.so by oracle/hdl/Makefile); the schematic top level is read by
tools/bdf_netlist.py.  Three layers:
  * needs /root/reference (this container): netlist facts, interpreter == C translation;
  * needs oracle/_ref/libua3_hdl.so (built here, travels to the GPU box): every golden filter function equals the HDL
    module edge for edge / sample for sample, and the golden chain equals the four-clock-domain HDL chain for all six
    clocking classes;
  * needs nothing: golden model == committed HDL vectors tests/golden/hdl_cases.npz (tools/gen_golden_hdl.py);
    on the GPU (-m gpu): CUDA frames == the same HDL vectors, through the C ABI.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, have_reference
from hdl_cases import CASES, make_adc

sys.path.insert(0, os.path.join(ROOT, "tools"))

GOLDEN = os.path.join(ROOT, "tests", "golden", "hdl_cases.npz")
needs_ref = pytest.mark.skipif(not have_reference(), reason="needs the reference tree")


def _hdl():
    from oracle import hdl_ref
    if not hdl_ref.available():
        pytest.skip("oracle/_ref/libua3_hdl.so not built and no reference tree")
    hdl_ref.lib()
    return hdl_ref


# ------------------------------------------------------------------------------------------------------------------
# the schematic: what the golden model's wiring claims rest on (UA3REO.bdf, extracted mechanically)
# ------------------------------------------------------------------------------------------------------------------
@needs_ref
def test_bdf_netlist_matches_the_golden_model_wiring():
    from bdf_netlist import Netlist
    nl = Netlist("/root/reference/FPGA/UA3REO.bdf")
    d = nl.describe
    # I <- sin, Q <- cos, both mixers see the raw ADC word, NCO outputs are truncated by nco_shift first
    assert d("MIXER_I", "datab") == "NCO_SHIFT_SIN.out" and d("MIXER_Q", "datab") == "NCO_SHIFT_COS.out"
    assert d("MIXER_I", "dataa") == d("MIXER_Q", "dataa") == "ADC_INPUT[11..0]"
    assert d("NCO_SHIFT_SIN", "in").startswith("NCO.fsin_o") and d("NCO_SHIFT_COS", "in").startswith("NCO.fcos_o")
    assert d("NCO", "phi_inc_i") == "STM32_INTERFACE.freq_out"
    for rail in "IQ":
        assert d("MIXER_SHIFT_" + rail, "in") == "MIXER_%s.result" % rail
        assert d("RX_CIC_" + rail, "filter_in") == "MIXER_SHIFT_%s.out" % rail
        assert d("RX_CICCOMP_" + rail, "filter_in") == "RX_CIC_%s.filter_out" % rail
        assert d("STM32_INTERFACE", "SPEC_" + rail) == "RX_CICCOMP_%s.filter_out" % rail
    assert d("RX_VOICE_HILBERT_I", "filter_in") == "RX_CICCOMP_I.filter_out"
    assert d("RX_VOICE_DELAY_Q", "data_in") == "RX_CICCOMP_Q.filter_out"
    assert d("STM32_INTERFACE", "VOICE_I") == "RX_VOICE_HILBERT_I.filter_out"
    assert d("STM32_INTERFACE", "VOICE_Q") == "RX_VOICE_DELAY_Q.data_out"
    assert nl.instances["RX_VOICE_DELAY_Q"]["params"] == {"bus_length": "16", "delay_length": "130"}
    # one reset / enable for the whole receive chain; clocks: clk_sys, c0 = /32, c1 = /4, c2 = /1024 (MAIN_PLL.v:105-116)
    for inst in ("RX_CIC_I", "RX_CIC_Q", "RX_CICCOMP_I", "RX_CICCOMP_Q", "RX_VOICE_HILBERT_I"):
        assert d(inst, "reset") == "RX_NOT.OUT (RX_N)" and d(inst, "clk_enable") == "STM32_INTERFACE.rx (RX)"
    assert d("RX_NOT", "IN") == "STM32_INTERFACE.rx (RX)"
    assert d("RX_CIC_I", "clk") == "clk_sys" and d("RX_CICCOMP_I", "clk") == "MAIN_PLL.c0"
    assert d("RX_VOICE_HILBERT_I", "clk").startswith("MAIN_PLL.c1") and d("RX_VOICE_DELAY_Q", "clk_in").startswith("MAIN_PLL.c2")
    pll = open("/root/reference/FPGA/MAIN_PLL.v").read()
    for k, div in ((0, 32), (1, 4), (2, 1024)):
        assert "clk%d_divide_by = %d," % (k, div) in pll and "clk%d_multiply_by = 1," % k in pll
        assert 'clk%d_phase_shift = "0"' % k in pll
    # transmit mirror
    assert d("TX_MIXER_I", "datab").startswith("NCO.fsin_o") and d("TX_MIXER_Q", "datab").startswith("NCO.fcos_o")
    for rail in "IQ":
        assert d("TX_CICCOMP_" + rail, "filter_in") == "STM32_INTERFACE.TX_" + rail
        assert d("TX_CIC_" + rail, "filter_in") == "TX_CICCOMP_%s.filter_out" % rail
        assert d("TX_MIXER_" + rail, "dataa") == "TX_CIC_%s.filter_out" % rail
    assert d("TX_SUMMATOR", "dataa") == "TX_MIXER_I.result" and d("TX_SUMMATOR", "datab") == "TX_MIXER_Q.result"
    assert d("DAC_CORRECTOR", "DATA_IN") == "TX_SUMMATOR.result"


# ------------------------------------------------------------------------------------------------------------------
# the evaluator's two back ends against each other, on the reference's five VHDL files
# ------------------------------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("module,in_bits,hold,n", [("rx_cic", 23, 1, 1600), ("rx_ciccomp", 16, 16, 700),
                                                     ("rx_hilb", 16, 256, 1300), ("tx_cic", 16, 512, 1600),
                                                     ("tx_ciccomp", 16, 46, 700)])
def test_interpreter_equals_c_translation(module, in_bits, hold, n):
    import vhdl_eval
    hdl = _hdl()
    rng = np.random.default_rng(hash(module) & 0xFFFF)
    vals = rng.integers(0, 1 << in_bits, n // hold + 2)
    vals[:2] = [(1 << (in_bits - 1)), (1 << (in_bits - 1)) - 1]          # most negative / most positive word first
    x = np.repeat(vals, hold)[:n]
    out_c, ce_c = hdl.run(module, x, want_ce=True)
    design = vhdl_eval.Design("/root/reference/FPGA/%s.vhd" % module)
    inst = design.instance()
    vhdl_eval.reset_instance(inst)
    has_ce = "ce_out" in design.sym
    for e in range(n):
        inst.set("filter_in", int(x[e]))
        inst.clock()
        got = inst.get("filter_out")
        want = int(out_c[e]) & ((1 << design.sym["filter_out"].w) - 1)
        assert got == want, "%s: edge %d interpreter %d, C translation %d" % (module, e, got, want)
        if has_ce:
            assert inst.get("ce_out") == ce_c[e]


@needs_ref
def test_type_checker_sees_every_assignment_width():
    """VHDL demands equal widths on both sides of an assignment; the evaluator derives the right-hand widths from its
    numeric_std rules, so parsing the reference's files IS a test of those rules.  Also: what it parsed is the lot."""
    import vhdl_eval
    sizes = {}
    for m in vhdl_eval.MODULES:
        d = vhdl_eval.Design("/root/reference/FPGA/%s.vhd" % m)
        sizes[m] = (len(d.signals), len(d.conc), len(d.procs))
    assert sizes == {"rx_cic": (67, 55, 14), "rx_ciccomp": (91, 84, 9), "rx_hilb": (16, 12, 5),
                     "tx_cic": (74, 63, 13), "tx_ciccomp": (63, 60, 5)}


def test_numeric_std_rules_on_a_synthetic_design(tmp_path):
    """The rules that matter here, each isolated: signed resize keeps the SIGN bit when truncating, '&' and '+' associate
    left to right at equal precedence, unary minus wraps, slices keep the kind, registers see pre-edge values."""
    import vhdl_eval
    src = """
LIBRARY IEEE; USE IEEE.std_logic_1164.all; USE IEEE.numeric_std.ALL;
ENTITY t IS PORT( clk : IN std_logic; clk_enable : IN std_logic; reset : IN std_logic;
  filter_in : IN std_logic_vector(7 DOWNTO 0); filter_out : OUT std_logic_vector(7 DOWNTO 0) ); END t;
ARCHITECTURE rtl OF t IS
  SIGNAL a : signed(7 DOWNTO 0); SIGNAL r4 : signed(3 DOWNTO 0); SIGNAL cat : signed(8 DOWNTO 0);
  SIGNAL neg : signed(7 DOWNTO 0); SIGNAL sl : signed(3 DOWNTO 0); SIGNAL q1 : signed(7 DOWNTO 0);
  SIGNAL q2 : signed(7 DOWNTO 0); SIGNAL prod : signed(15 DOWNTO 0); SIGNAL rnd : signed(8 DOWNTO 0);
BEGIN
  a <= signed(filter_in);
  r4 <= resize(a, 4);
  cat <= a(7) & a(7 DOWNTO 0) + ( "0" & (a(1)));
  neg <= -a;
  sl <= a(7 DOWNTO 4);
  prod <= a * a;
  rnd <= shift_right(cat, 1);
  p : PROCESS (clk, reset) BEGIN
    IF reset = '1' THEN q1 <= (OTHERS => '0'); q2 <= (OTHERS => '0');
    ELSIF clk'event AND clk = '1' THEN
      IF clk_enable = '1' THEN q1 <= a; q2 <= q1; END IF;
    END IF;
  END PROCESS p;
  filter_out <= std_logic_vector(q2);
END rtl;
"""
    path = tmp_path / "t.vhd"
    path.write_text(src)
    d = vhdl_eval.Design(str(path))
    inst = d.instance()
    vhdl_eval.reset_instance(inst)
    seq = [0x80, 0x7F, 0xB5, 0x03, 0xFE]
    outs = []
    for v in seq:
        inst.set("filter_in", v)
        inst.clock()
        s8 = v - 256 if v & 0x80 else v
        assert inst.get_signed("r4") == (-(8 if s8 < 0 else 0) + (v & 7))              # sign bit + 3 low bits
        assert inst.get_signed("cat") == s8 + ((v >> 1) & 1)                           # (a(7) & a) + ("0" & a(1))
        assert inst.get_signed("neg") == (-128 if s8 == -128 else -s8)
        assert inst.get_signed("sl") == (s8 >> 4)
        assert inst.get_signed("prod") == s8 * s8
        assert inst.get_signed("rnd") == (s8 + ((v >> 1) & 1)) >> 1
        outs.append(inst.get("filter_out"))
    assert outs == [0, 0x80, 0x7F, 0xB5, 0x03]                                         # two registers: q2 lags a by one edge after its own
    assert "RESIZE_S" in d.emit_c()


# ------------------------------------------------------------------------------------------------------------------
# golden filter functions == HDL modules
# ------------------------------------------------------------------------------------------------------------------
class _Cic(ctypes.Structure):
    _fields_ = [("cnt", ctypes.c_uint32), ("inreg", ctypes.c_int32), ("s", ctypes.c_uint64 * 5),
                ("d", ctypes.c_uint64 * 5), ("outreg", ctypes.c_int16)]


class _Comp(ctypes.Structure):
    _fields_ = [("p0", ctypes.c_int16 * 33), ("p1", ctypes.c_int16 * 33), ("n_in", ctypes.c_uint32)]


class _TxCic(ctypes.Structure):
    _fields_ = [("cnt", ctypes.c_uint32), ("wreg", ctypes.c_int16), ("d", ctypes.c_int64 * 5), ("up", ctypes.c_int64),
                ("i", ctypes.c_int64 * 5), ("out14", ctypes.c_int16)]


def _golden_comp(L, u, n_in0=0):
    c = _Comp()
    c.n_in = n_in0
    y = ctypes.c_int16(0)
    out = []
    for v in u:
        if L.ua3g_rx_ciccomp_push(ctypes.byref(c), ctypes.c_int16(int(v)), ctypes.byref(y)):
            out.append(y.value)
    return np.array(out, np.int64)


def _golden_hilb(L, y):
    L.ua3g_rx_hilb_push.restype = ctypes.c_int16
    h = (ctypes.c_int16 * 256)()
    return np.array([L.ua3g_rx_hilb_push(ctypes.byref(h), ctypes.c_int16(int(v))) for v in y], np.int64)


def _extremes(rng, n, bits):
    x = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), n)
    x[: n // 8] = -(1 << (bits - 1))                      # a run of the most negative word, then the most positive
    x[n // 8: n // 4] = (1 << (bits - 1)) - 1
    return x


def test_golden_rx_cic_equals_hdl_edge_for_edge(oracle):
    hdl = _hdl()
    L = oracle.lib()
    x = _extremes(np.random.default_rng(21), 40 * 512, 23)
    out = hdl._sx(hdl.run("rx_cic", x & 0x7FFFFF), 16)
    c = _Cic()
    L.ua3g_rx_cic_reset(ctypes.byref(c))
    g = np.empty(x.size, np.int64)
    for e in range(x.size):
        L.ua3g_rx_cic_clock(ctypes.byref(c), ctypes.c_int32(int(x[e])))
        g[e] = c.outreg
    assert np.array_equal(g, out)
    assert np.abs(out).max() > 30000                     # the extremes reach the output range


def test_golden_rx_ciccomp_equals_hdl_for_every_input_phase(oracle):
    """96 kHz words held for 16 compensator clocks each; the first word arrives `off` clocks after reset release.
    off = 1..16: alignment A (golden n_in = 0); off = 0, 17..32: alignment B (n_in = 1).  Output latency: one sample
    (none when the first word is already there at release)."""
    hdl = _hdl()
    L = oracle.lib()
    K = 420
    u = _extremes(np.random.default_rng(22), K, 16)
    gold = {0: _golden_comp(L, u, 0), 1: _golden_comp(L, u, 1)}
    for off in range(0, 33):
        e = np.arange(16 * K)
        k = (e - off) // 16
        xin = np.where(k >= 0, u[np.clip(k, 0, K - 1)], 0)
        out, ce = hdl.run("rx_ciccomp", xin & 0xFFFF, want_ce=True)
        ys = hdl._sx(out, 16)[np.nonzero(ce)[0]]
        align_b = 0 if 1 <= off <= 16 else 1
        g = gold[align_b]
        lag = 0 if off == 0 else 1
        n = min(len(g), len(ys) - lag) - 2
        assert n > 190 and np.array_equal(ys[lag:lag + n], g[:n]), "off=%d" % off


def test_golden_rx_hilb_equals_hdl(oracle):
    hdl = _hdl()
    L = oracle.lib()
    K = 600
    y = _extremes(np.random.default_rng(23), K, 16)
    g = _golden_hilb(L, y)
    for off, lag in ((0, 1), (1, 2), (128, 2), (256, 2)):
        e = np.arange(256 * K)
        k = (e - off) // 256
        xin = np.where(k >= 0, y[np.clip(k, 0, K - 1)], 0)
        out = hdl._sx(hdl.run("rx_hilb", xin & 0xFFFF), 16)[256::256]     # after every output-register edge
        n = K - 4
        assert np.array_equal(out[lag:lag + n - lag], g[:n - lag]), "off=%d" % off


def test_golden_tx_filters_equal_hdl(oracle):
    hdl = _hdl()
    L = oracle.lib()
    rng = np.random.default_rng(24)
    # tx_cic: edge for edge
    M = 48
    w = _extremes(rng, M, 16)
    xin = np.repeat(w, 512)
    out = hdl._sx(hdl.run("tx_cic", xin & 0xFFFF), 14)
    c = _TxCic()
    L.ua3g_tx_cic_reset(ctypes.byref(c))
    L.ua3g_tx_cic_clock.restype = ctypes.c_int16
    g = np.array([L.ua3g_tx_cic_clock(ctypes.byref(c), ctypes.c_int16(int(v))) for v in xin], np.int64)
    assert np.array_equal(g, out)
    # tx_ciccomp: 46 clocks per 48 kHz word, two outputs per word
    K = 300
    x = _extremes(rng, K, 16)
    dp = (ctypes.c_int16 * 24)()
    z = (ctypes.c_int16 * 2)()
    gz = []
    for v in x:
        L.ua3g_tx_ciccomp_push(ctypes.byref(dp), ctypes.c_int16(int(v)), z)
        gz += [z[0], z[1]]
    gz = np.array(gz, np.int64)
    for off in (0, 1, 23, 46):
        e = np.arange(46 * K)
        k = (e - off) // 46
        xi = np.where(k >= 0, x[np.clip(k, 0, K - 1)], 0)
        o = hdl._sx(hdl.run("tx_ciccomp", xi & 0xFFFF), 16)
        first = 26 if off <= 1 else 72                     # the delay line loads at edges 1, 47, ..; first output 25 edges later
        seq = o[first::23][: 2 * K - 6]                   # phase_23_1: one output every 23 clocks
        assert np.array_equal(seq, gz[: seq.size]), "off=%d" % off


# ------------------------------------------------------------------------------------------------------------------
# golden chain == HDL chain (four clock domains) for all six clocking classes
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("t_rx,tau,cls", [(517, 64, (1, 3, 129)), (990, 640, (1, 3, 130)), (45, 64, (1, 2, 129)),
                                          (127, 64, (0, 3, 129)), (0, 64, (0, 3, 130)), (63, 64, (0, 2, 129))])
def test_golden_chain_equals_hdl_chain(oracle, t_rx, tau, cls):
    hdl = _hdl()
    n = 400 * 1024
    rng = np.random.default_rng(t_rx)
    adc = rng.integers(-2048, 2048, n).astype(np.int16)
    fcw = int(rng.integers(1, 1 << 22))
    x_i, x_q = oracle.golden_mixer(adc, fcw)
    from oracle import vlog_ref
    if vlog_ref.available():              # the mixer path from the reference's own Verilog: the HDL chain then runs executed
        e_i, e_q = oracle.executed_mixer(adc, fcw)                      # sources from the mixer inputs to the frame
        assert np.array_equal(e_i, x_i) and np.array_equal(e_q, x_q)
        x_i, x_q = e_i, e_q
    hf = hdl.frames_at(hdl.rx_chain(x_i, x_q, t_rx=t_rx), tau)
    g = oracle.GoldenDDC(fcw, cls).push(adc)
    lags = [lag for lag in range(4) if np.array_equal(hf[lag:lag + 390], g[:390])]
    assert len(lags) == 1, "class %r does not reproduce the HDL frames at t_rx=%d tau=%d" % (cls, t_rx, tau)
    # and no other class does
    for other in ((1, 3, 129), (1, 3, 130), (1, 2, 129), (0, 3, 129), (0, 3, 130), (0, 2, 129), (0, 0, 130)):
        if other != cls:
            go = oracle.GoldenDDC(fcw, other).push(adc)
            assert not any(np.array_equal(hf[lag:lag + 390], go[:390]) for lag in range(4))


@pytest.mark.parametrize("t_tx,tau", [(0, 0), (7, 300), (100, 900), (511, 300), (700, 0), (1023, 900), (333, 512), (990, 64)])
def test_golden_duc_composition_equals_hdl_chain(oracle, t_tx, tau):
    """The golden DUC feeds each compensator output to the interpolator exactly once, in order (duc_golden.c: ua3g_duc_push).
    The HDL does that across two clock domains - tx_ciccomp on PLL c3 (46 clocks per 48 kHz word), tx_cic on clk_sys - that leave
    reset at an arbitrary instant, with the MCU writing its words at an arbitrary offset: the composed VHDL stream must equal
    the golden stream up to a pure latency of one or two 48 kHz periods.  (A sweep of 420 alignments, 37 x 73 tick grid, found
    that for every one except a write that coincides with the compensator's load edge.)"""
    hdl = _hdl()
    L = oracle.lib()
    L.ua3g_tx_cic_clock.restype = ctypes.c_int16
    words = np.random.default_rng(t_tx + tau).integers(-20000, 20001, 28)
    words[:3] = [32767, -32768, 1]
    dp = (ctypes.c_int16 * 24)()
    z = (ctypes.c_int16 * 2)()
    c = _TxCic()
    L.ua3g_tx_cic_reset(ctypes.byref(c))
    g = []
    for v in words:
        L.ua3g_tx_ciccomp_push(ctypes.byref(dp), ctypes.c_int16(int(v)), z)
        for h in range(2):
            zz = ctypes.c_int16(z[h])
            g += [L.ua3g_tx_cic_clock(ctypes.byref(c), zz) for _ in range(512)]
    g = np.array(g, np.int64)
    got = hdl.tx_chain(words, t_tx=t_tx, tau=tau)
    assert np.abs(g).max() > 4000                                        # the interpolator is driven well into its range
    lags = [lag for lag in (1024, 2048) if np.array_equal(got[lag:], g[:g.size - lag])]
    assert len(lags) == 1, "no pure latency reproduces the golden stream at t_tx=%d tau=%d" % (t_tx, tau)


# ------------------------------------------------------------------------------------------------------------------
# committed HDL vectors
# ------------------------------------------------------------------------------------------------------------------
def _cases():
    z = np.load(GOLDEN)
    for name, kind, seed, fcw, t_rx, tau in CASES:
        meta = z[name + "_meta"]
        assert (int(meta[0]), int(meta[1]), int(meta[2]), int(meta[7])) == (fcw, t_rx, tau, seed), "regenerate hdl_cases.npz"
        yield name, make_adc(kind, seed), fcw, tuple(int(v) for v in meta[3:6]), z[name + "_frames"], z


def test_hdl_vectors_cover_every_clocking_class_and_the_edge_cases():
    classes = {cls for _, _, _, cls, _, _ in _cases()}
    assert classes == {(1, 3, 129), (1, 3, 130), (1, 2, 129), (0, 3, 129), (0, 3, 130), (0, 2, 129)}
    z = np.load(GOLDEN)
    assert all(z[n + "_frames"].shape[0] >= 2045 and z[n + "_cic_i"].shape[0] >= 4090 for n in z["names"])
    # the (-2048) x (-2048) product wraps to -2^22 in rx_mixer_shift.v:9: the I rail of the "wrap" case sees it
    assert np.abs(z["wrap_cic_i"].astype(np.int64)).max() > 8000


def test_golden_model_equals_hdl_vectors(oracle):
    for name, adc, fcw, cls, frames, z in _cases():
        g, ci, cq = oracle.GoldenDDC(fcw, cls).push(adc, want_cic=True)
        assert np.array_equal(g[: frames.shape[0]], frames), name
        n = z[name + "_cic_i"].shape[0]
        assert np.array_equal(ci[:n], z[name + "_cic_i"]) and np.array_equal(cq[:n], z[name + "_cic_q"]), name


@pytest.mark.gpu
def test_cuda_ddc_equals_hdl_vectors(pkg):
    """CUDA frames against the frames the reference's VHDL produced, through the C ABI, one receiver per class."""
    by_class = {}
    for name, adc, fcw, cls, frames, _ in _cases():
        by_class.setdefault(cls, []).append((name, adc, fcw, frames))
    for cls, cases in by_class.items():
        for name, adc, fcw, frames in cases:
            rx = pkg.Receiver(3, 1 << 19)
            rx.set_clocking(*cls)
            assert rx.get_clocking() == cls
            rx.set_fcw([fcw, (fcw + 1) & 0x3FFFFF, fcw])
            got = []
            for off in range(0, adc.size, 1 << 19):
                rx.push(adc[off:off + (1 << 19)])
                got.append(rx.read_frames())
            rx.close()
            got = np.concatenate(got, axis=1)
            assert np.array_equal(got[0, : frames.shape[0]], frames), name
            assert np.array_equal(got[2], got[0])


# ------------------------------------------------------------------------------------------------------------------
# committed TRANSMIT vectors: every stage but the NCO executed from the reference's HDL (tools/gen_golden_hdl_tx.py)
# ------------------------------------------------------------------------------------------------------------------
TX_GOLDEN = os.path.join(ROOT, "tests", "golden", "hdl_tx_cases.npz")


def _tx_cases():
    z = np.load(TX_GOLDEN)
    for name in sorted(k[:-3] for k in z.files if k.endswith("_iq")):
        yield name, z[name + "_iq"], int(z[name + "_meta"][0]), z[name + "_dac"], z[name + "_otr"]


def test_golden_duc_equals_hdl_tx_vectors(oracle):
    n_cases = 0
    for name, iq, fcw, dac, otr in _tx_cases():
        g_dac, g_otr = oracle.GoldenDUC(fcw).push(iq[:, 0], iq[:, 1])
        assert np.array_equal(g_dac[:dac.size], dac), name
        assert np.array_equal(g_otr[:otr.size], otr), name
        assert dac.size == (iq.shape[0] - 2) * 1024 and dac.max() < (1 << 14)
        n_cases += 1
    assert n_cases == 6


@needs_ref
def test_committed_hdl_tx_vectors_are_current():
    import gen_golden_hdl_tx
    if not _hdl():
        pytest.skip("no HDL library")
    fresh = gen_golden_hdl_tx.generate()
    stored = np.load(TX_GOLDEN)
    assert sorted(fresh) == sorted(stored.files)
    for k in fresh:
        assert np.array_equal(fresh[k], stored[k]), k


@pytest.mark.gpu
def test_cuda_duc_equals_hdl_tx_vectors(pkg):
    """The CUDA DUC against DAC words produced by the reference's own HDL (tx_ciccomp.vhd, tx_cic.vhd across their clock domains,
    tx_mixer.v, tx_summator.v, DAC_corrector.v executed; NCO = the golden convention), through the C ABI, bit for bit."""
    cases = list(_tx_cases())
    n = cases[0][1].shape[0]
    rx = pkg.Receiver(len(cases), 1 << 14)
    rx.set_fcw(np.array([c[2] for c in cases], np.uint32))
    rx.duc_enable(n)
    rx.duc_push(np.stack([c[1] for c in cases]))
    got = rx.duc_read_dac()
    rx.close()
    for ch, (name, iq, fcw, dac, otr) in enumerate(cases):
        assert np.array_equal(got[ch, :dac.size], dac), name
