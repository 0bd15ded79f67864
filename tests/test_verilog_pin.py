"""Pins the restated glue of the golden model (and of the product's host code) to the reference's hand-written VERILOG,
executed from its own source text: tools/verilog_eval.py translates mixer.v, tx_mixer.v, tx_summator.v, rx_mixer_shift.v,
nco_shift.v, data_delay.v, DAC_corrector.v and stm32_interface.v to C (oracle/hdl/Makefile -> oracle/_ref/libua3_vlog.so);
for the MCU <-> FPGA byte bus the reference sits on BOTH ends: the firmware's unmodified bus driver fpga.c is linked against
the translated stm32_interface.v (oracle/_ref/fw_bus_hdl, glue oracle/ref_harness/bus_hdl.c).

  * needs oracle/_ref/libua3_vlog.so / fw_bus_hdl (built in this container from /root/reference, travel with the snapshot);
  * the evaluator's own expression rules are pinned by a synthetic module with known answers (needs nothing but gcc)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, have_reference

sys.path.insert(0, os.path.join(ROOT, "tools"))
FW_BUS_HDL = os.path.join(ROOT, "oracle", "_ref", "fw_bus_hdl")
FW_FPGA = os.path.join(ROOT, "oracle", "_ref", "fw_fpga")


def _vl():
    from oracle import vlog_ref
    if not vlog_ref.available():
        pytest.skip("oracle/_ref/libua3_vlog.so not built and no reference tree")
    vlog_ref.lib()
    return vlog_ref


def _sx(v, w):
    v = np.asarray(v, dtype=np.int64) & ((1 << w) - 1)
    return np.where(v >> (w - 1) != 0, v - (1 << w), v)


# ------------------------------------------------------------------------------------------------------------------
# the evaluator itself: IEEE 1364 sizing / signedness corner rules on a synthetic module with known answers
# ------------------------------------------------------------------------------------------------------------------
SYNTH = r"""
module corner(clk, a, b, u, y_sum, y_mixed, y_cmp_s, y_cmp_u, y_cat, y_part, y_shift, y_tern, y_neg, y_nba0, y_nba1, y_mul);
input clk;
input signed [7:0] a;
input signed [7:0] b;
input [7:0] u;
output reg signed [15:0] y_sum = 0;
output reg signed [15:0] y_mixed = 0;
output reg y_cmp_s = 0;
output reg y_cmp_u = 0;
output reg [15:0] y_cat = 0;
output reg [3:0] y_part = 0;
output reg signed [15:0] y_shift = 0;
output reg signed [15:0] y_tern = 0;
output reg signed [11:0] y_neg = 0;
output reg [7:0] y_nba0 = 1;
output reg [7:0] y_nba1 = 2;
output reg signed [19:0] y_mul = 0;
integer n = 0;
always @ (posedge clk)
begin
    y_sum = a + b;              // all signed: operands sign-extended to 16 bits before the add
    y_mixed = a + u;            // one unsigned operand: both zero-extended
    y_cmp_s = (a < b);          // signed compare
    y_cmp_u = (a < u);          // unsigned compare
    y_cat = {a, b[3:0], 4'd9};  // concatenation is unsigned, operands self-determined
    y_part = y_cat[11:8];
    y_shift = a >>> 2;          // arithmetic because the expression is signed
    y_tern = (u == 'd0) ? a : -a;
    y_neg = -2000;
    y_mul = a * b;              // 20-bit context: full signed product
    y_nba0 <= y_nba1;           // swap through non-blocking assignments
    y_nba1 <= y_nba0;
    n = n + 1;
end
endmodule
"""


def test_evaluator_expression_rules(tmp_path):
    import verilog_eval
    src = tmp_path / "corner.v"
    src.write_text(SYNTH)
    c = tmp_path / "corner.c"
    c.write_text(verilog_eval.translate([(str(src), {}, None)]))
    so = tmp_path / "corner.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", str(so), str(c)])
    L = ctypes.CDLL(str(so))
    vp = ctypes.c_void_p
    L.vl_find.restype, L.vl_find.argtypes = vp, [ctypes.c_char_p]
    L.vl_size.restype, L.vl_size.argtypes = ctypes.c_size_t, [vp]
    L.vl_init.argtypes = [vp, vp]
    L.vl_n_fields.restype, L.vl_n_fields.argtypes = ctypes.c_int, [vp]
    L.vl_field_name.restype, L.vl_field_name.argtypes = ctypes.c_char_p, [vp, ctypes.c_int]
    L.vl_field_offset.restype, L.vl_field_offset.argtypes = ctypes.c_size_t, [vp, ctypes.c_int]
    L.vl_clock_edge.restype, L.vl_clock_edge.argtypes = ctypes.c_int, [vp, vp, ctypes.c_char_p]
    m = L.vl_find(b"corner")
    buf = ctypes.create_string_buffer(L.vl_size(m))
    L.vl_init(m, buf)
    off = {L.vl_field_name(m, i).decode(): L.vl_field_offset(m, i) for i in range(L.vl_n_fields(m))}

    def word(name):
        return ctypes.c_uint64.from_buffer(buf, off[name])

    def s16(v):
        return v - 65536 if v & 0x8000 else v

    for a, b, u in [(-3, 5, 0), (-128, -128, 200), (127, -1, 255), (0, 0, 1), (-77, 90, 0)]:
        word("a").value, word("b").value, word("u").value = a & 0xFF, b & 0xFF, u
        swap_before = (word("y_nba0").value, word("y_nba1").value)
        assert L.vl_clock_edge(m, buf, b"clk") == 0
        assert s16(word("y_sum").value) == a + b
        assert word("y_mixed").value == ((a & 0xFF) + u) & 0xFFFF
        assert word("y_cmp_s").value == int(a < b)
        assert word("y_cmp_u").value == int((a & 0xFF) < u)
        assert word("y_cat").value == ((a & 0xFF) << 8) | ((b & 0xF) << 4) | 9
        assert word("y_part").value == (a & 0xF)
        assert s16(word("y_shift").value) == a >> 2
        assert s16(word("y_tern").value) == (a if u == 0 else -a)
        assert word("y_neg").value == (-2000) & 0xFFF
        v = word("y_mul").value
        assert (v - (1 << 20) if v >> 19 else v) == a * b
        assert (word("y_nba0").value, word("y_nba1").value) == (swap_before[1], swap_before[0])
    assert word("n").value == 5


# ------------------------------------------------------------------------------------------------------------------
# receive glue: nco_shift.v, mixer.v, rx_mixer_shift.v == ua3g_rx_mix (oracle/ddc_golden.c)
# ------------------------------------------------------------------------------------------------------------------
def test_rx_mixer_chain(oracle):
    vl = _vl()
    sh, mx, ms = vl.VModule("nco_shift"), vl.VModule("mixer"), vl.VModule("rx_mixer_shift")
    # nco_shift: all 2^14 inputs - the golden model's `nco14 >> 2`
    for v in range(0, 1 << 14, 7):
        sh["in"] = v
        sh.settle()
        assert sh["out"] == v >> 2
    assert sh.width("in") == 14 and sh.width("out") == 12 and mx.width("result") == 24 and ms.width("out") == 23
    L = oracle.lib()
    # mixer (lpm_mult signed 12 x 12 -> 24, one pipeline register) then the shift: corners, the (-2048)^2 wrap, a random sweep
    rng = np.random.default_rng(1)
    pairs = [(-2048, -2048), (-2048, 2047), (2047, 2047), (0, -2048), (-1, -1), (1, -2048), (-2048, 1)]
    pairs += [tuple(int(x) for x in rng.integers(-2048, 2048, 2)) for _ in range(20000)]
    mx["clken"] = 1
    for adc, nco12 in pairs:
        mx["dataa"], mx["datab"] = adc, nco12
        mx.clock("clock")
        assert mx.signed("result") == adc * nco12                       # 24 bits hold every product, 2^22 included
        ms["in"] = mx["result"]
        ms.settle()
        got = ms.signed("out")
        assert got == int(_sx(adc * nco12, 23))
        for low in (0, 3):                                              # the two bits nco_shift drops do not matter
            assert got == L.ua3g_rx_mix(adc, (nco12 << 2) | low)
    mx["dataa"], mx["datab"] = -2048, -2048
    mx.clock("clock")
    ms["in"] = mx["result"]
    ms.settle()
    assert mx.signed("result") == 1 << 22 and ms.signed("out") == -(1 << 22)    # the one product that wraps in 23 bits
    # clken low holds the pipeline register
    mx["clken"] = 0
    mx["dataa"], mx["datab"] = 5, 7
    mx.clock("clock")
    assert mx.signed("result") == 1 << 22


def test_rx_mixer_every_product(oracle):
    """All 2^24 (ADC sample, 12-bit NCO word) pairs through mixer.v and rx_mixer_shift.v (vector runner of the translated
    library): the golden model's product-and-wrap for every one of them; every 97th pair also through ua3g_rx_mix itself."""
    vl = _vl()
    mx, ms = vl.VModule("mixer"), vl.VModule("rx_mixer_shift")
    L = oracle.lib()
    b = np.arange(-2048, 2048, dtype=np.int64)
    ones = np.ones(b.size, np.int64)
    n_wrap = 0
    for a in range(-2048, 2048):
        res = mx.run({"dataa": np.full(b.size, a & 0xFFF, np.int64), "datab": b & 0xFFF, "clken": ones}, ["result"], clock="clock")["result"]
        out = ms.run({"in": res & 0xFFFFFF}, ["out"])["out"]
        out = np.where(out >> 22 != 0, out - (1 << 23), out)
        want = a * b
        want = ((want + (1 << 22)) & ((1 << 23) - 1)) - (1 << 22)
        assert np.array_equal(out, want), a
        n_wrap += int((a * b != want).sum())
        for j in range((a * 31) % 97, b.size, 97):
            assert int(out[j]) == L.ua3g_rx_mix(a, int(b[j]) << 2)
    assert n_wrap == 1                                                    # (-2048)^2 is the only product that leaves 23 bits


def test_q_delay_is_130_registers(oracle):
    vl = _vl()
    d = vl.VModule("data_delay_q")                                       # parameters from UA3REO.bdf: bus_length 16, delay_length 130
    x = np.random.default_rng(2).integers(-32768, 32768, 700)
    got = []
    for v in x:
        d["data_in"] = int(v)
        d.clock("clk_in")
        got.append(d.signed("data_out"))
    got = np.array(got)
    # after the k-th edge data_out = the sample of edge k - 129 (130 registers: the value sampled at an edge is visible after it)
    assert np.array_equal(got[129:], x[:-129]) and not got[:129].any()
    # the golden model's delay line (ua3g_delay_push returns the word BEFORE the push: in[n - 130])
    L = oracle.lib()
    st = ctypes.create_string_buffer(4096)
    L.ua3g_delay_reset(st)
    L.ua3g_delay_push.restype = ctypes.c_int16
    L.ua3g_delay_push.argtypes = [ctypes.c_void_p, ctypes.c_int16]
    gold = np.array([L.ua3g_delay_push(st, int(v)) for v in x])
    assert np.array_equal(gold[1:], got[:-1])
    dflt = vl.VModule("data_delay")                                      # the file's own defaults: 16 x 32
    for k in range(40):
        dflt["data_in"] = k + 1
        dflt.clock("clk_in")
    assert dflt["data_out"] == 40 - 31


# ------------------------------------------------------------------------------------------------------------------
# transmit output stage: tx_mixer.v x2, tx_summator.v, DAC_corrector.v == ua3g_dac_word (oracle/duc_golden.c)
# ------------------------------------------------------------------------------------------------------------------
def test_tx_output_stage(oracle):
    vl = _vl()
    mi, mq, sm, dc = vl.VModule("tx_mixer"), vl.VModule("tx_mixer"), vl.VModule("tx_summator"), vl.VModule("DAC_corrector")
    L = oracle.lib()
    L.ua3g_dac_word.restype = ctypes.c_uint16
    L.ua3g_dac_word.argtypes = [ctypes.c_int32] * 4 + [ctypes.POINTER(ctypes.c_int)]
    rng = np.random.default_rng(3)
    cases = [(-8192, -8192, -8192, -8192), (8191, 8191, 8191, 8191), (-8192, 8191, 8191, -8192), (0, 0, 0, 0),
             (-8192, -8192, 8191, 8191), (8191, -8192, 8191, -8192), (-1, 1, 1, -1)]
    cases += [tuple(int(v) for v in rng.integers(-8192, 8192, 4)) for _ in range(20000)]
    cases += [tuple(int(v) for v in rng.choice([-8192, -8191, 8190, 8191], 4)) for _ in range(2000)]    # drive the summator over
    n_over = 0
    for m in (mi, mq, sm):
        m["clken"] = 1
    for i14, q14, s14, c14 in cases:
        mi["dataa"], mi["datab"] = i14, s14
        mq["dataa"], mq["datab"] = q14, c14
        mi.clock("clock"); mq.clock("clock")
        assert mi.signed("result") == i14 * s14 and mq.signed("result") == q14 * c14
        sm["dataa"], sm["datab"] = mi["result"], mq["result"]
        sm.clock("clock")
        dc["DATA_IN"] = sm["result"]
        dc.clock("clk_in")
        ov = ctypes.c_int(0)
        want = L.ua3g_dac_word(i14, q14, s14, c14, ctypes.byref(ov))
        assert dc["DATA_OUT"] == want
        assert sm["overflow"] == ov.value
        n_over += ov.value
    assert n_over >= 5                # (-8192)^2 + (-8192)^2 = 2^27 is the only sum that leaves 28 bits: the wrap / overflow branch ran


# ------------------------------------------------------------------------------------------------------------------
# the byte bus: fpga.c <-> stm32_interface.v, the reference on both ends
# ------------------------------------------------------------------------------------------------------------------
def _need(*bins):
    for b in bins:
        if not os.path.exists(b):
            pytest.skip("%s not built (needs the reference tree) / did not travel" % os.path.basename(b))


@pytest.mark.parametrize("iq_swap", [0, 1])
def test_rx_frames_over_the_executed_bus(oracle, tmp_path, iq_swap):
    """RX IQ (command 4): the four 16-bit filter outputs presented to stm32_interface.v, clocked out by the firmware's own
    FPGA_fpgadata_getiq(), must leave the firmware's rings / FFT buffers exactly as the golden model's 8-byte frames do when
    they are served byte by byte (oracle/_ref/fw_fpga) - i.e. ua3g_frame_pack IS the wire format of k = 400..407."""
    _need(FW_BUS_HDL, FW_FPGA)
    n = 1500
    rng = np.random.default_rng(10 + iq_swap)
    w = rng.integers(-32768, 32768, (n, 4)).astype(np.int16)            # SPEC_I, SPEC_Q, VOICE_I, VOICE_Q
    w[:8] = [[-32768, 32767, -1, 0], [0x1234, 0x5678, -0x1234, -0x5678], [255, 256, -255, -256], [1, 2, 3, 4],
             [-2, -3, -4, -5], [0x7F80, -0x7F80, 0x00FF, -0x0100], [0, 0, 0, 0], [32767, -32768, 32767, -32768]]
    L = oracle.lib()
    frames = np.zeros((n, 8), np.uint8)
    for i in range(n):
        L.ua3g_frame_pack(frames[i].ctypes.data_as(ctypes.c_void_p), int(w[i, 1]), int(w[i, 0]), int(w[i, 3]), int(w[i, 2]))
    fw, ff = tmp_path / "words.bin", tmp_path / "frames.bin"
    w.tofile(fw); frames.tofile(ff)
    o1, o2 = tmp_path / "hdl.out", tmp_path / "stub.out"
    subprocess.check_call([FW_BUS_HDL, "rx", str(iq_swap), "7100000", str(fw), str(o1)], stderr=subprocess.DEVNULL)
    subprocess.check_call([FW_FPGA, str(iq_swap), "7100000", str(ff), str(o2)])
    a, b = np.fromfile(o1, np.uint8), np.fromfile(o2, np.uint8)
    assert a.size == b.size == n * 12 + 4 * 384 * 4 + 2 * 512 * 4 + 8
    assert np.array_equal(a, b)


def run_bus_tx(iq_int16, tmp):
    """FPGA_fpgadata_sendiq() against the executed stm32_interface.v: (TX_I/TX_Q the module latched [n, 2], wire bytes [n, 4])"""
    fi, fo = os.path.join(tmp, "iq.f32"), os.path.join(tmp, "tx.out")
    np.asarray(iq_int16, dtype=np.float32).tofile(fi)           # integer-valued floats: what processTxAudio leaves in the send buffers
    subprocess.check_call([FW_BUS_HDL, "tx", "7100000", fi, fo], stderr=subprocess.DEVNULL)
    rec = np.fromfile(fo, dtype=np.dtype([("tx", "<i2", 2), ("wire", "u1", 4)]))
    return rec["tx"].copy(), rec["wire"].copy()


def run_bus_params(freq, mode, preamp, ptt, adc, adc_otr, dac_otr, tmp):
    fa = os.path.join(tmp, "adc.bin")
    np.asarray(adc, dtype=np.int16).tofile(fa)
    out = subprocess.check_output([FW_BUS_HDL, "params", str(freq), str(mode), str(preamp), str(ptt), fa, str(adc_otr), str(dac_otr)]).decode()
    return {k: int(v) for k, v in (line.split("=") for line in out.split())}


def test_tx_words_over_the_executed_bus(tmp_path):
    """TX IQ (command 3): FPGA_fpgadata_sendiq() -> stm32_interface.v k = 300..303 -> TX_I / TX_Q.  The bytes that cross the
    bus are Q hi, Q lo, I hi, I lo - the order include/ua3reo_b200.h documents for ua3reo_duc_push_wire - and the module
    reassembles exactly the int16 pair the firmware sent."""
    _need(FW_BUS_HDL)
    n = 600
    rng = np.random.default_rng(20)
    iq = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
    iq[:4] = [[-32768, 32767], [1, -1], [255, 256], [-256, -255]]
    tx, wire = run_bus_tx(iq, str(tmp_path))
    assert tx.shape == (n, 2) and np.array_equal(tx, iq)
    u = iq.astype(np.uint16)
    assert np.array_equal(wire, np.stack([u[:, 1] >> 8, u[:, 1] & 255, u[:, 0] >> 8, u[:, 0] & 255], axis=1))


@pytest.mark.parametrize("freq,mode,preamp,ptt,lo,hi", [(7100000, 1, 1, 0, -1500, 1199), (14200000, 0, 0, 1, -2048, 2047),
                                                        (28500000, 4, 1, 0, -900, -3), (3700000, 3, 0, 0, 5, 1800)])
def test_params_over_the_executed_bus(pkg, oracle, tmp_path, freq, mode, preamp, ptt, lo, hi):
    """SEND PARAMS (command 1) and GET PARAMS (command 2) through the executed state machine: the tuning word the FPGA latches
    is getPhraseFromFrequency() of what the firmware tunes == the product's ua3reo_phrase_from_frequency and the golden
    model's; ADC min / max / OTR come back through FPGA_fpgadata_getparam() as the firmware decodes them - the minimum sign
    extended, the maximum NOT (fpga.c:270), the quirk ua3reo_get_params keeps."""
    _need(FW_BUS_HDL)
    rng = np.random.default_rng(freq % 1000)
    adc = rng.integers(lo, hi + 1, 5000).astype(np.int16)
    adc[100], adc[200] = lo, hi
    kv = run_bus_params(freq, mode, preamp, ptt, adc, 1, 0, str(tmp_path))
    fcw, _ = pkg.phrase_from_frequency(freq)                           # host function of the product (no device needed)
    assert kv["freq_out"] == fcw == oracle.lib().ua3g_phrase_from_frequency(freq, None)
    assert kv["tx"] == ptt and kv["rx"] == 1 - ptt
    assert kv["preamp_enable"] == (preamp if not ptt else 0)
    assert kv["hdl_adc_min"] == min(lo, 2000) and kv["hdl_adc_max"] == max(hi, -2000)
    assert kv["packet_bytes"] == 5
    mn, mx = kv["hdl_adc_min"] & 0xFFF, kv["hdl_adc_max"] & 0xFFF
    # byte 4 = encoder / key: only DATA_BUS_OUT[4:0] is assigned at k == 204, bits 7:5 are those of byte 3
    assert [kv["packet%d" % i] for i in range(5)] == [1, (mn >> 8) << 4 | (mx >> 8), mn & 255, mx & 255, mx & 0xE0]
    assert kv["TRX_ADC_OTR"] == 1 and kv["TRX_DAC_OTR"] == 0
    assert kv["TRX_ADC_MINAMPLITUDE"] == kv["hdl_adc_min"]
    assert kv["TRX_ADC_MAXAMPLITUDE"] == mx                            # 12 raw bits: a negative maximum reads as 4096 + max


# ------------------------------------------------------------------------------------------------------------------
# committed vectors of the executed bus (tools/gen_golden_bus.py) for the GPU box, where neither the reference tree nor
# necessarily the built harness exists: regenerated here and compared, so that they cannot go stale
# ------------------------------------------------------------------------------------------------------------------
def test_committed_bus_vectors_are_current(tmp_path):
    _need(FW_BUS_HDL)
    import gen_golden_bus
    fresh = gen_golden_bus.generate(str(tmp_path))
    stored = np.load(os.path.join(ROOT, "tests", "golden", "bus_cases.npz"))
    assert sorted(fresh) == sorted(stored.files)
    for k in fresh:
        assert np.array_equal(fresh[k], stored[k]), k


# ------------------------------------------------------------------------------------------------------------------
# two evaluation paths of the evaluator: the compiled C translation and a Python interpreter over the same typed tree
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not have_reference(), reason="needs the reference tree")
@pytest.mark.parametrize("fname,cname,overrides,clocks,steps", [
    ("DAC_corrector.v", "DAC_corrector", {}, ["clk_in"], 400),
    ("data_delay.v", "data_delay", {}, ["clk_in"], 200),
    ("data_delay.v", "data_delay_q", {"bus_length": 16, "delay_length": 130}, ["clk_in"], 300),
    ("mixer.v", "mixer", {}, ["clock"], 400),
    ("tx_summator.v", "tx_summator", {}, ["clock"], 400),
    ("nco_shift.v", "nco_shift", {}, [], 200),
    ("stm32_interface.v", "stm32_interface", {}, ["clk_in", "adcclk_in"], 3000)])
def test_interpreter_equals_c_translation_verilog(fname, cname, overrides, clocks, steps):
    import verilog_eval as V
    vl = _vl()
    mod = V.parse_file(os.path.join(V.REF, fname))
    it = V.Interp(mod, overrides)
    cm = vl.VModule(cname)
    inputs = [s for s in it.e.sigs.values() if s.kind in ("input", "inout") and s.name not in clocks]
    rng = np.random.default_rng(len(fname) + steps)
    loop_vars = set()                                 # `for` counters are locals of the C translation (nothing else reads them)
    for _, body in mod["always"]:
        loop_vars |= it.e._loopvars_in(body)
    names = [s.name for s in it.e.sigs.values() if s.name not in loop_vars]
    for step in range(steps):
        for s in inputs:
            if s.name == "DATA_BUS":                                  # mostly commands and small data so that the state machine moves
                v = int(rng.integers(0, 8)) if rng.random() < 0.5 else int(rng.integers(0, 256))
                it.ext_in["DATA_BUS"] = v
                cm["DATA_BUS__ext"] = v
                continue
            v = int(rng.integers(0, 1 << s.width)) if s.width < 63 else int(rng.integers(0, 1 << 62))
            if s.name == "DATA_SYNC":
                v = int(rng.random() < 0.2)
            it.v[s.name] = v
            cm[s.name] = v
        it.settle()
        cm.settle()
        if clocks:
            ck = clocks[int(rng.integers(0, len(clocks)))]
            it.clock(ck)
            cm.clock(ck)
        for nme in names:
            s = it.e.sigs[nme]
            if s.kind == "input" and nme in clocks:
                continue
            if s.length:
                assert [cm[(nme, i)] for i in range(s.length)] == list(it.v[nme]), "%s step %d: %s" % (cname, step, nme)
            else:
                assert cm[nme] == it.v[nme], "%s step %d: %s (C %d, interpreter %d)" % (cname, step, nme, cm[nme], it.v[nme])
