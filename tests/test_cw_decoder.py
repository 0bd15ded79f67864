"""The host half of the CW decoder, ua3reo_cw_decoder_step(): compared block by block with what the reference firmware's
own CWDecoder_Process() (cw_decoder.c:69-170, host-built unmodified into oracle/_ref/fw_cw by oracle/ref_harness, tracing
wrapper cw_trace_wrap.c) leaves in its element buffer, CW_Decoder_WPM and the text bar, for the Goertzel magnitudes the
reference computed (tests/golden/cw_cases.npz, generator tools/gen_golden_cw.py).  No device involved."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _trace(pkg, magnitudes):
    d = pkg.CwDecoder(pkg.load_library())
    text, wpm, code = [], [], []
    for m in magnitudes:
        text.append(d.step(float(m)))
        wpm.append(int(d.wpm)); code.append(d.code.decode("ascii"))
    return text, np.array(wpm, np.uint16), code


def test_cw_decoder_against_reference_firmware_fixture(pkg):
    z = np.load(os.path.join(ROOT, "tests", "golden", "cw_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) >= 5
    for n in names:
        text, wpm, code = _trace(pkg, z[n + "/magnitude"])
        want_text, want_code = [str(x) for x in z[n + "/text"]], [str(x) for x in z[n + "/code"]]
        assert text == want_text, "%s: decoded %r, firmware %r" % (n, "".join(text), "".join(want_text))
        assert np.array_equal(wpm, z[n + "/wpm"]), n + ": CW_Decoder_WPM"
        assert code == want_code, n + ": element buffer"
    assert "UA3REO UA3REO K" in "".join(str(x) for x in z["cq_20wpm/text"])


def test_cw_decoder_live_against_host_built_firmware(pkg):
    fw = os.path.join(ROOT, "oracle", "_ref", "fw_cw")
    if not os.path.exists(fw):
        pytest.skip("oracle/_ref/fw_cw not built here")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_golden_cw as g
    rng = np.random.default_rng(20261018)     # fixed seed: the round-end run must be reproducible
    for _ in range(3):
        words = " ".join("".join(rng.choice(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"), rng.integers(2, 6))) for _ in range(4))
        audio = g.keyed_audio(words, int(rng.integers(10, 36)), amp=float(rng.uniform(500, 6000)), noise=float(rng.uniform(5, 200)),
                              seed=int(rng.integers(1 << 30)))
        mag, wpm, code, text = g.run_reference(audio)
        got_text, got_wpm, got_code = _trace(pkg, mag)
        assert got_text == text and np.array_equal(got_wpm, wpm) and got_code == code, words
