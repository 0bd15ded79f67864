"""TRX_DoAutoGain() (trx_manager.c:268-356) as the pure host function ua3reo_autogain_step(): compared step by step
with what the reference firmware's own function leaves behind (tests/golden/autogain_cases.npz, made by
tools/gen_golden_autogain.py from the host-built oracle/_ref/fw_autogain) and, when that binary is present, with live
runs on fresh random sequences.  No device is involved: the library must load and run this without a GPU."""
import os
import subprocess

import numpy as np

from conftest import ROOT


def _run(pkg, seq):
    ag = pkg.AutoGain(pkg.load_library())
    return np.array([ag.step(int(v)) for v in seq], np.uint8)


def test_autogain_against_reference_firmware_fixture(pkg):
    z = np.load(os.path.join(ROOT, "tests", "golden", "autogain_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) >= 7
    for n in names:
        got, want = _run(pkg, z[n + "/in"]), z[n + "/out"]
        assert np.array_equal(got, want), "%s: first difference at step %d" % (n, int(np.argmax((got != want).any(1))))


def test_autogain_live_against_host_built_firmware(pkg):
    fw = os.path.join(ROOT, "oracle", "_ref", "fw_autogain")
    if not os.path.exists(fw):
        import pytest
        pytest.skip("oracle/_ref/fw_autogain not built here")
    rng = np.random.default_rng(20261018)     # fixed seed: the round-end run must be reproducible
    for _ in range(5):
        seq = np.clip(np.cumsum(rng.integers(-200, 205, 600)) + rng.integers(0, 1500), -100, 2047).astype(np.int16)
        out = subprocess.run([fw], input=" ".join(map(str, seq)) + "\n", capture_output=True, text=True, check=True).stdout
        want = np.array([[int(x) for x in l.split()] for l in out.strip().splitlines()], np.uint8)
        assert np.array_equal(_run(pkg, seq), want)
